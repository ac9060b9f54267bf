#!/bin/bash
# after a change that touches csrc/: the whole GPU suite, smoke, the dominant-kernel capture that ties
# roofline.traffic to the source hash, and the bench lines of configs 2 (the driver's command), 1, 3, 5
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_final.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_gpu_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python tools/prof_stage.py 1 2 5 6 > gpurun_out/plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:tapgemm_tc2_kernel -s 4 -c 4 -f -o gpurun_out/prof_r2_dominant python tools/prof_stage.py 1 2 5 6 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_r2_dominant.ncu-rep gpurun_out/r2_ncu_full_tapgemm_pair.json --dominant cifar10 100 1,2,5,6; rm -f gpurun_out/prof_r2_dominant.ncu-rep
cp gpurun_out/ncu_dominant_kernel.json profiles/ncu_dominant_kernel.json   # so that the bench line below carries the traffic
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --stage-table gpurun_out/stages_cifar10.json > gpurun_out/bench_cifar10.json 2> gpurun_out/bench_cifar10.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_cifar10.err
for wl in svhn celeba_hq256 celeba_crop; do
  timeout 900 python bench.py --workload $wl --steps 20 --warmup 5 --stage-table gpurun_out/stages_$wl.json > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err
done
python - <<'PY'
import json
for wl in ('cifar10', 'svhn', 'celeba_crop', 'celeba_hq256'):
    try:
        d = json.load(open(f'gpurun_out/bench_{wl}.json')); det = d['details']
        print(wl, round(d['value']), 'ls/s | frac', round(det['frac_of_tensor_roofline'], 3), '| 1p', d.get('value_bwd1pass') and round(d['value_bwd1pass']['value']), '| e2e', round(d['e2e']['value']), '| traffic', d['roofline']['traffic'], '| eager', d.get('reference_cuda_eager') and round(d['reference_cuda_eager']['value']), '| cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'], 1), '| clk', d['clocks'] and (d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(wl, 'unreadable', e)
PY
