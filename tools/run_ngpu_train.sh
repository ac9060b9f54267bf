#!/bin/bash
# N-GPU runs of one box (N = $1): training-iteration bench with the collectives in the timed region, config 4 over all
# 50 000 latents sharded across the ranks
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29521 --mode train --steps 20 --warmup 5 > gpurun_out/bench_train_cifar10_${N}gpu.json 2> gpurun_out/bench_train_${N}gpu.err; echo "train N=$N rc=$?"; tail -2 gpurun_out/bench_train_${N}gpu.err | cut -c1-300
run 29522 --workload svhn_test --full-50k --warmup 3 --no-secondary --no-cpu-baseline --no-eager-ref > gpurun_out/bench_svhn_test_50k_${N}gpu.json 2> gpurun_out/bench_svhn_test_${N}gpu.err; echo "svhn_test N=$N rc=$?"; tail -2 gpurun_out/bench_svhn_test_${N}gpu.err | cut -c1-300
python - <<PY
import json
for f in ('gpurun_out/bench_train_cifar10_${N}gpu.json', 'gpurun_out/bench_svhn_test_50k_${N}gpu.json'):
    try:
        d = json.load(open(f)); print(f, round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'steps', d['steps'], json.dumps(d['details'])[:700], d.get('prior_sampling'))
    except Exception as e: print(f, e)
PY
