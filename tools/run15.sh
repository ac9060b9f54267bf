#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu15.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/t_gpu15.log
LSNF_NO_GRAPH=1 LSNF_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/trace_cifar.json 2> gpurun_out/trace_cifar.err; grep "lsnf trace" gpurun_out/trace_cifar.err | tail -14
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cifar15.json 2> gpurun_out/bench_cifar15.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cifar15.json')); print('graph', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['flow_prior_kernel_us'])"
timeout 600 python bench.py --workload svhn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_svhn15.json 2> gpurun_out/bench_svhn15.err; python -c "
import json; d=json.load(open('gpurun_out/bench_svhn15.json')); print('graph', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['flow_prior_kernel_us'])"
