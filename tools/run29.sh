#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py all svhn32 cifar32 celeba64 svhn64_b130 > gpurun_out/diag29.log 2>&1; echo "diag rc=$?"; grep " tc \| simt " gpurun_out/diag29.log | cut -c1-330
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu29.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_gpu29.log
for wl in cifar10 svhn celeba_crop celeba_hq256; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_$wl.json > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', round(d['value']), 'ls/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value']), 'frac', round(d['config']['frac_of_tensor_roofline'],3)); print([(r['stage'], round(r['us'],1)) for r in json.load(open('gpurun_out/stages_$wl.json'))['stages']])"
done
