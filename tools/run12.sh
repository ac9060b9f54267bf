#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"; tail -5 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 --stage-table gpurun_out/stages_cifar12.json > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "1gpu rc=$?"; cat gpurun_out/bench_1gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; echo "ref rc=$?"; cat gpurun_out/bench_ref2.json
