#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu42.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_gpu42.log
for wl in cifar10 svhn celeba_crop celeba_hq256; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_${wl}42.json > gpurun_out/bench_${wl}42.json 2> gpurun_out/bench_${wl}42.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}42.json')); print('$wl', round(d['value']), 'ls/s', round(d['roofline']['iteration_us'],1), 'us/iter', d['clocks'])
s=json.load(open('gpurun_out/stages_${wl}42.json')); print([round(x['us'],1) for x in s['stages']], s['flow_prior_kernel_us'])"
done
LSNF_NO_GRAPH=1 LSNF_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/trace_new.json 2> gpurun_out/trace_new.err; grep "lsnf trace" gpurun_out/trace_new.err | tail -12
