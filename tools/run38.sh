#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu38.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_gpu38.log
EXPS="0 4 32" timeout 300 python tools/exp_epi.py 0 3 4 5 6 7 > gpurun_out/exp38.json 2> gpurun_out/exp38.err; echo "exp rc=$?"; cat gpurun_out/exp38.json | tr -d '\n'; echo
for wl in cifar10 svhn celeba_crop celeba_hq256; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_${wl}38.json > gpurun_out/bench_${wl}38.json 2> gpurun_out/bench_${wl}38.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${wl}38.json')); print('$wl', round(d['value']), 'ls/s', round(d['roofline']['iteration_us'],1), 'us/iter', d['clocks'])
s=json.load(open('gpurun_out/stages_${wl}38.json')); print([round(x['us'],1) for x in s['stages']], s['flow_prior_kernel_us'])"
done
