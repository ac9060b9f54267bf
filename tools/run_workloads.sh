#!/bin/bash
# parameter-update tests, the training-iteration bench, and the bench lines of every
# BASELINE workload (configs 1-5) plus the saturating-batch CIFAR-10 lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_param_updates.py tests/test_gpu_langevin.py -q -s -x -k "update or iteration or gradients or adam" > gpurun_out/t_gpu_updates.log 2>&1; echo "updates rc=$?"; grep -v "^$" gpurun_out/t_gpu_updates.log | tail -12 | cut -c1-300
timeout 900 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/bench_train_cifar10_1gpu.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"; tail -3 gpurun_out/bench_train.err; cut -c1-1500 gpurun_out/bench_train_cifar10_1gpu.json
for wl in svhn celeba_crop celeba_hq256; do
  timeout 900 python bench.py --workload $wl --steps 20 --warmup 5 --stage-table gpurun_out/stages_$wl.json > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err
done
timeout 900 python bench.py --workload svhn_test --steps 8 --warmup 3 --no-secondary > gpurun_out/bench_svhn_test.json 2> gpurun_out/bench_svhn_test.err; echo "bench svhn_test rc=$?"; tail -2 gpurun_out/bench_svhn_test.err
for b in 1024 4096; do
  timeout 900 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-eager-ref > gpurun_out/bench_cifar10_b$b.json 2> gpurun_out/bench_cifar10_b$b.err; echo "bench cifar b=$b rc=$?"; tail -2 gpurun_out/bench_cifar10_b$b.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    if d.get('impl') == 'reference' or 'details' not in d: continue
    det = d['details']
    print(f.split('/')[-1], round(d['value']), d['unit'], '| frac', det.get('frac_of_tensor_roofline') and round(det['frac_of_tensor_roofline'], 3),
          '| 1p', d.get('value_bwd1pass') and round(d['value_bwd1pass']['value']), '| e2e', round(d['e2e']['value']), '| timed s', round(det.get('timed_region_s', 0), 2),
          '| eager', d.get('reference_cuda_eager') and d['reference_cuda_eager'].get('value') and round(d['reference_cuda_eager']['value']),
          '| cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'], 1), '| prior', d.get('prior_sampling') and round(d['prior_sampling']['samples_per_sec']),
          '| clk', d['clocks'] and (d['clocks']['sm_mhz'], d['clocks']['reasons']))
PY
