#!/bin/bash
mkdir -p gpurun_out
for v in old new; do
  if [ $v = old ]; then export LSNF_LIB=$PWD/tools/_ab/liblsnf_old.so; else unset LSNF_LIB; fi
  LSNF_NO_GRAPH=1 LSNF_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/trace_$v.json 2> gpurun_out/trace_$v.err; echo "== $v"; grep "lsnf trace" gpurun_out/trace_$v.err | tail -16
done
