#!/bin/bash
# round 2, first GPU pass: the whole GPU suite (new real-shape parity tests included), smoke, the default bench line
# (primary 3-pass + secondary 1-pass), the reference arm, and the ncu evidence of the same build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_r2.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/t_gpu_r2.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --stage-table gpurun_out/stages_cifar10.json > gpurun_out/bench_cifar10.json 2> gpurun_out/bench_cifar10.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cifar10.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_cifar10.json'))
    print('cifar10', round(d['value']), 'ls/s frac', round(d['details']['frac_of_tensor_roofline'], 3), '| 1-pass', d['value_bwd1pass'] and round(d['value_bwd1pass']['value']), '| e2e', round(d['e2e']['value']), '| R', d['details']['langevin_calls_per_step'], 'timed', round(d['details']['timed_region_s'], 2), 's | eager', d['reference_cuda_eager'], '| cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 1), d['clocks'])
    s = json.load(open('gpurun_out/stages_cifar10.json')); print([round(x['us'], 1) for x in s['stages']], s['flow_prior_kernel_us'])
except Exception as e:
    print('bench parse failed', e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
# ncu: launch list of an eager loop, then full captures (only after the same commands exited 0 without ncu)
LSNF_NO_GRAPH=1 timeout 600 python bench.py --steps 1 --warmup 3 --calls-per-step 1 --no-cpu-baseline --no-secondary --no-eager-ref > gpurun_out/plain_launch.log 2>&1 &&
LSNF_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 130 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --calls-per-step 1 --no-cpu-baseline --no-secondary --no-eager-ref > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch rc=$?"; wc -l gpurun_out/launches_r2.csv
timeout 300 python tools/prof_stage.py 1 2 5 6 > gpurun_out/plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc2_kernel -s 4 -c 4 -f -o gpurun_out/prof_r2_dominant python tools/prof_stage.py 1 2 5 6 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
python tools/ncu_summarize.py gpurun_out/prof_r2_dominant.ncu-rep gpurun_out/r2_ncu_full_tapgemm_pair.json --dominant cifar10 100 1,2,5,6
FLOW=1 timeout 300 python tools/prof_stage.py 0 > gpurun_out/plain_prof_flow.log 2>&1 &&
FLOW=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:flow_forward_kernel|flow_inverse_kernel" -c 2 -f -o gpurun_out/prof_r2_flow python tools/prof_stage.py 0 > gpurun_out/ncu_flow.log 2>&1
echo "ncu flow rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_r2_flow.ncu-rep gpurun_out/r2_ncu_full_flow.json
