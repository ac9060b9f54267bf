#!/bin/bash
# A/B of two library builds on the same box: tools/_ab/liblsnf_old.so (tools/build_ab_lib.sh <commit>) against the in-tree build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_ab.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_gpu_ab.log
for r in 1 2; do
for v in old new; do
  if [ $v = new ]; then unset LSNF_LIB; else export LSNF_LIB=$PWD/tools/_ab/liblsnf_$v.so; fi
  echo -n "$v: "; EXPS="0" timeout 300 python tools/exp_epi.py 2>gpurun_out/ab_$v.err | tr -d '\n'; echo
done; done
for i in 1 2; do
for v in old new; do
  if [ $v = old ]; then export LSNF_LIB=$PWD/tools/_ab/liblsnf_old.so; else unset LSNF_LIB; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_ab_$v$i.json > gpurun_out/bench_ab_$v$i.json 2> gpurun_out/bench_ab_$v$i.err || echo "bench $v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$v$i.json')); s=json.load(open('gpurun_out/stages_ab_$v$i.json'))
print('$v$i', round(d['value']), 'ls/s', round(d['roofline']['iteration_us'],1), 'us/iter', d['clocks']['sm_mhz'], d['clocks']['power_w_max'], [round(x['us'],1) for x in s['stages']], round(s['flow_prior_kernel_us'],1))"
done; done
