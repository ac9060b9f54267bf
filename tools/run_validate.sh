#!/bin/bash
# full validation: GPU tests, smoke, benches of every BASELINE config on 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_final.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/t_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for wl in cifar10 svhn celeba_crop celeba_hq256; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --stage-table gpurun_out/stages_$wl.json > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', round(d['value']), 'ls/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value']), 'frac', round(d['config']['frac_of_tensor_roofline'],3), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1), d['clocks'])"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
