#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu14.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/t_gpu14.log
LSNF_NO_GRAPH=1 LSNF_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/trace_cifar.json 2> gpurun_out/trace_cifar.err; grep "lsnf trace" gpurun_out/trace_cifar.err | tail -14
timeout 600 python bench.py --steps 5 --warmup 3 --stage-table gpurun_out/stages_cifar14.json > gpurun_out/bench_cifar14.json 2> gpurun_out/bench_cifar14.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cifar14.json')); print('graph', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['flow_prior_kernel_us'], d['cpu_baseline'], d['clocks'])"
timeout 600 python bench.py --workload svhn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_svhn14.json 2> gpurun_out/bench_svhn14.err; python -c "
import json; d=json.load(open('gpurun_out/bench_svhn14.json')); print('graph', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['flow_prior_kernel_us'])"
timeout 600 python bench.py --workload celeba_crop --steps 3 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_celeba14.json > gpurun_out/bench_celeba14.json 2> gpurun_out/bench_celeba14.err; python -c "
import json; d=json.load(open('gpurun_out/bench_celeba14.json')); print('celeba', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['frac_of_tensor_roofline'])"; tail -3 gpurun_out/bench_celeba14.err
timeout 900 python bench.py --workload celeba_hq256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hq14.json 2> gpurun_out/bench_hq14.err; python -c "
import json; d=json.load(open('gpurun_out/bench_hq14.json')); print('hq256', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['frac_of_tensor_roofline'])"; tail -3 gpurun_out/bench_hq14.err
