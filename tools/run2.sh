#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/summary2.txt
tail -25 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --stage-table gpurun_out/stages_cifar.json > gpurun_out/bench_cifar.json 2> gpurun_out/bench_cifar.err; echo "bench rc=$?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench_cifar.json; tail -5 gpurun_out/bench_cifar.err; cat gpurun_out/stages_cifar.json
timeout 600 python bench.py --workload svhn --steps 5 --warmup 3 --no-cpu-baseline --stage-table gpurun_out/stages_svhn.json > gpurun_out/bench_svhn.json 2> gpurun_out/bench_svhn.err; cat gpurun_out/bench_svhn.json; cat gpurun_out/stages_svhn.json
