#!/bin/bash
# N=2 bench lines (Langevin loop and whole training iteration) of BASELINE configs 1, 3, 5 -- `gpurun --gpus 2`
mkdir -p gpurun_out
P=29560
for wl in svhn celeba_crop celeba_hq256; do
  P=$((P+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-eager-ref > gpurun_out/bench_${wl}_2gpu.json 2> gpurun_out/bench_${wl}_2gpu.err; echo "bench $wl rc=$?"; tail -1 gpurun_out/bench_${wl}_2gpu.err | cut -c1-200
  P=$((P+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --workload $wl --mode train --steps 20 --warmup 5 > gpurun_out/bench_train_${wl}_2gpu.json 2> gpurun_out/bench_train_${wl}_2gpu.err; echo "train $wl rc=$?"; tail -1 gpurun_out/bench_train_${wl}_2gpu.err | cut -c1-200
  python -c "
import json
d=json.load(open('gpurun_out/bench_${wl}_2gpu.json')); print('$wl langevin N=2', round(d['value']), 'ls/s', d['clocks']['sm_mhz'], d['clocks']['reasons'])
d=json.load(open('gpurun_out/bench_train_${wl}_2gpu.json')); det=d['details']; print('$wl train N=2', round(d['value']), 'ls/s | ms/iteration', round(det['ms_per_iteration'],2), '| exposed', det['allreduce_exposed_ms_per_iteration'], '| alone', det['allreduce_alone'] and (round(det['allreduce_alone']['ms'],3), round(det['allreduce_alone']['busbw_gbs'])))"
done
