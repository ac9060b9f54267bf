#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py all svhn32 cifar32 celeba64 svhn64_b130 > gpurun_out/diag7.log 2>&1; echo "diag rc=$?"; grep -E "simt_vs_tc| tc " gpurun_out/diag7.log | cut -c1-600
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu7.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_gpu7.log
LSNF_NO_GRAPH=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_cifar7.json > gpurun_out/bench_cifar7.json 2> gpurun_out/bench_cifar7.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cifar7.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar7.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])
for r in json.load(open('gpurun_out/stages_cifar7.json'))['stages']: print(r['stage'], r['kind'], r['layer'], round(r['us'],1), r['block_n'])"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cifar7g.json 2> gpurun_out/bench_cifar7g.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar7g.json')); print('graph', d['value'], d['ms_per_step'])"
LSNF_NO_GRAPH=1 timeout 600 python bench.py --workload svhn --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_svhn7.json > gpurun_out/bench_svhn7.json 2> gpurun_out/bench_svhn7.err
python -c "
import json; d=json.load(open('gpurun_out/bench_svhn7.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])
for r in json.load(open('gpurun_out/stages_svhn7.json'))['stages']: print(r['stage'], r['kind'], r['layer'], round(r['us'],1), r['block_n'])"
