#!/bin/bash
mkdir -p gpurun_out
export LSNF_BWD_PASSES=1
timeout 300 python tools/gpu_diag.py all svhn32 cifar32 celeba64 > gpurun_out/diag32.log 2>&1; echo "diag rc=$?"; grep " tc \| simt " gpurun_out/diag32.log | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_langevin.py -q -m gpu -k "long_chain or fixture or one_step or sharding or training" -s 2>&1 | grep -v "^$" | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_cifar32.json > gpurun_out/bench_cifar32.json 2> gpurun_out/bench_cifar32.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar32.json')); print('cifar10 1-pass bwd', round(d['value']), 'ls/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value']), 'frac', round(d['config']['frac_of_tensor_roofline'],3)); print([(r['stage'], round(r['us'],1)) for r in json.load(open('gpurun_out/stages_cifar32.json'))['stages']])"
