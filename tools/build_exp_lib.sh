#!/bin/bash
# builds experiment variants of the library next to the product one: tools/_ab/liblsnf_fwd2pass{1,2}.so
# (two-pass forward: one cross term of the hi|lo split dropped -- see DESIGN.md section 4.1)
set -e
cd "$(dirname "$0")/.."
CS=latent-space-normalizing-flow_b200/csrc
for v in 1 2; do
  objs=""
  for f in plan gen_aux flow adam wgrad tapgemm_simt tapgemm_tc; do
    o=/tmp/exp_${v}_$f.o
    nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC -Iinclude -I$CS -DLSNF_EXP_FWD2PASS=$v -c $CS/$f.cu -o $o &
    objs="$objs $o"
  done
  wait
  nvcc -shared -o tools/_ab/liblsnf_fwd2pass$v.so $objs -gencode arch=compute_100a,code=sm_100a
done
ls -la tools/_ab/
