#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_langevin.py -q -m gpu > gpurun_out/t_gpu6.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_gpu6.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cifar6.json 2> gpurun_out/bench_cifar6.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cifar6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar6.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])"
LSNF_NO_GRAPH=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cifar6ng.json 2> gpurun_out/bench_cifar6ng.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar6ng.json')); print('nograph', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --workload svhn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_svhn6.json 2> gpurun_out/bench_svhn6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_svhn6.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])"
python tools/prof_stage.py 1 4 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc_kernel -s 4 -c 2 -o gpurun_out/prof_r1_stage1_4 python tools/prof_stage.py 1 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
