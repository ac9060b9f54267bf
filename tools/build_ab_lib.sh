#!/bin/bash
# Builds the library of another commit into tools/_ab/liblsnf_old.so for tools/run_ab.sh (same ABI assumed):
#   bash tools/build_ab_lib.sh <commit>
set -e
C=${1:-HEAD~1}
T=$(mktemp -d)
git archive "$C" latent-space-normalizing-flow_b200/csrc include | tar -x -C "$T"
mkdir -p tools/_ab
for f in plan gen_aux flow adam wgrad tapgemm_simt tapgemm_tc; do
  nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC -I"$T/include" \
       -I"$T/latent-space-normalizing-flow_b200/csrc" -c "$T/latent-space-normalizing-flow_b200/csrc/$f.cu" -o "$T/$f.o" 2>/dev/null &
done
wait
nvcc -shared -o tools/_ab/liblsnf_old.so "$T"/*.o -gencode arch=compute_100a,code=sm_100a
ls -la tools/_ab/liblsnf_old.so
