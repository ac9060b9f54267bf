#!/bin/bash
mkdir -p gpurun_out
for r in 1 2; do
for v in old new V_NOBITS V_NOWAIT V_OLDRING ALL; do
  if [ $v = new ]; then unset LSNF_LIB; else export LSNF_LIB=$PWD/tools/_ab/liblsnf_$v.so; fi
  echo -n "$v: "; EXPS="0" timeout 300 python tools/exp_epi.py 1 2 2>gpurun_out/ab2_$v.err | tr -d '\n'; echo
done; done
