#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/train_ddp_check.py 2>&1 | grep "DDP\|Error\|error" | head
timeout 600 python -m pytest tests/test_gpu_langevin.py -q -m gpu -k training 2>&1 | tail -3
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    fi
    python -c "
import json; d=json.load(open('gpurun_out/scale_$n.json')); print('N=$n', round(d['value']), 'ls/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']), d['clocks'])"
  fi
done
