#!/bin/bash
# bench.py --mode train on N GPUs of one box (N = $1)
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --mode train --steps 20 --warmup 5 > gpurun_out/bench_train_cifar10_${N}gpu.json 2> gpurun_out/bench_train_${N}gpu.err; echo "train N=$N rc=$?"; tail -2 gpurun_out/bench_train_${N}gpu.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_train_cifar10_${N}gpu.json')); det=d['details']; print(d['n_gpus'], round(d['value']), 'ls/s | ms/iteration', round(det['ms_per_iteration'],2), '| exposed', det['allreduce_exposed_ms_per_iteration'], '| alone', det['allreduce_alone'] and (round(det['allreduce_alone']['ms'],3), round(det['allreduce_alone']['busbw_gbs'])))"
