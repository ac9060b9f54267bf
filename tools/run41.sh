#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu41.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_gpu41.log
EXPS="0 1 64 0 1" timeout 300 python tools/exp_epi.py 1 2 3 5 6 > gpurun_out/exp41.json 2> gpurun_out/exp41.err; echo "exp rc=$?"; cat gpurun_out/exp41.json | tr -d '\n'; echo
for e in 0 1 0 1; do
LSNF_EXP=$e timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_cifar41_$e.json > gpurun_out/bench_cifar41_$e.json 2> gpurun_out/bench_cifar41_$e.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar41_$e.json')); print('exp$e', round(d['value']), 'ls/s', round(d['roofline']['iteration_us'],1), 'us/iter', d['clocks'])
s=json.load(open('gpurun_out/stages_cifar41_$e.json')); print([round(x['us'],1) for x in s['stages']], s['flow_prior_kernel_us'])"
done
LSNF_NO_GRAPH=1 LSNF_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/trace_new.json 2> gpurun_out/trace_new.err; grep "lsnf trace" gpurun_out/trace_new.err | tail -12
