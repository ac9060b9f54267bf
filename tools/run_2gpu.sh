#!/bin/bash
# 2-GPU checks: data-parallel training iteration vs single process, two devices in one process, N=2 bench lines
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_ddp_check.py > gpurun_out/ddp2.log 2>&1; echo "ddp rc=$?"; grep "DDP check" gpurun_out/ddp2.log | cut -c1-400; tail -3 gpurun_out/ddp2.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_sampling_and_long_chains.py -q -s -k "two_devices or two_plans" > gpurun_out/t_two_devices.log 2>&1; echo "two devices rc=$?"; tail -2 gpurun_out/t_two_devices.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode train --steps 20 --warmup 5 > gpurun_out/bench_train_cifar10_2gpu.json 2> gpurun_out/bench_train_2gpu.err; echo "train2 rc=$?"; tail -2 gpurun_out/bench_train_2gpu.err; cut -c1-300 gpurun_out/bench_train_cifar10_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cifar10_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/bench_2gpu.err; cut -c1-200 gpurun_out/bench_cifar10_2gpu.json
python - <<'PY'
import json
for f in ('gpurun_out/bench_train_cifar10_2gpu.json',):
    try:
        d = json.load(open(f)); print(f, round(d['value']), d['ms_per_step'], json.dumps(d['details'])[:900])
    except Exception as e: print(f, e)
PY
