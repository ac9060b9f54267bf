#!/bin/bash
# bench.py --mode train (one whole training iteration per step) for every BASELINE training workload on 1 GPU
mkdir -p gpurun_out
for wl in svhn cifar10 celeba_crop celeba_hq256; do
  timeout 300 python bench.py --workload $wl --mode train --steps 20 --warmup 5 > gpurun_out/bench_train_${wl}_1gpu.json 2> gpurun_out/bench_train_${wl}_1gpu.err; echo "train $wl rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_train_${wl}_1gpu.json')); det=d['details']; print('$wl', round(d['value']), 'ls/s | ms/iteration', round(det['ms_per_iteration'],2), {k: round(v,2) for k,v in det.items() if k.endswith('_ms') and isinstance(v,(int,float))}, d['clocks']['sm_mhz'])"
done
