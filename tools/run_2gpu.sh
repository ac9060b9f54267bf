#!/bin/bash
# 2-GPU checks: data-parallel training iteration vs single process, and the N=2 bench line
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_ddp_check.py > gpurun_out/ddp2.log 2>&1; echo "ddp rc=$?"; grep "DDP check" gpurun_out/ddp2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-400 gpurun_out/bench_2gpu.json
