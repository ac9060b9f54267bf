#!/bin/bash
# the driver's scaling command at N GPUs (N = $1): python -m torch.distributed.run ... bench.py --gpus N --steps 20 --warmup 5
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_cifar10_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_reference_${N}gpu.json 2> gpurun_out/bench_reference_${N}gpu.err; echo "reference N=$N rc=$?"; cut -c1-160 gpurun_out/bench_reference_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_cifar10_${N}gpu.json')); print(d['n_gpus'], round(d['value']), 'ls/s | per GPU', round(d['value']/d['n_gpus']), '| 1p', round(d['value_bwd1pass']['value']), '| e2e', round(d['e2e']['value']), '| timed', round(d['details']['timed_region_s'],2), '| clk', d['clocks'])"
