#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py all svhn32 cifar32 celeba64 > gpurun_out/diag5.log 2>&1; echo "diag rc=$?"; tail -12 gpurun_out/diag5.log
timeout 600 python tools/kink_diag.py > gpurun_out/kink5.log 2>&1; echo "kink rc=$?"; grep -o '"err_vs_reference": \[[^]]*\]' gpurun_out/kink5.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu5.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/t_gpu5.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_cifar5.json > gpurun_out/bench_cifar5.json 2> gpurun_out/bench_cifar5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cifar5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])
for r in json.load(open('gpurun_out/stages_cifar5.json')): print(r['stage'], r['kind'], r['layer'], round(r['us'],1), r['block_n'])"
timeout 600 python bench.py --workload svhn --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_svhn5.json > gpurun_out/bench_svhn5.json 2> gpurun_out/bench_svhn5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_svhn5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])
for r in json.load(open('gpurun_out/stages_svhn5.json')): print(r['stage'], r['kind'], r['layer'], round(r['us'],1), r['block_n'])"
