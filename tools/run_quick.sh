#!/bin/bash
# quick GPU sanity: parity tests, smoke, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_quick.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_gpu_quick.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_quick.json > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(round(d['value']), 'ls/s', round(d['roofline']['iteration_us'],1), 'us/iter e2e', round(d['e2e']['value']), d['clocks'])
s=json.load(open('gpurun_out/stages_quick.json')); print([round(x['us'],1) for x in s['stages']], s['flow_prior_kernel_us'])"
