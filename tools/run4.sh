#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kink_diag.py > gpurun_out/kink.log 2>&1; echo "kink rc=$?"; tail -12 gpurun_out/kink.log
timeout 900 python -m pytest tests -q -m gpu -k "update or philox or fixture" > gpurun_out/t_gpu4.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_gpu4.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_cifar4.json > gpurun_out/bench_cifar4.json 2> gpurun_out/bench_cifar4.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])"
timeout 600 python bench.py --workload svhn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_svhn4.json 2> gpurun_out/bench_svhn4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_svhn4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['iteration_us'], d['roofline']['all_gemm_stages_us'])"
