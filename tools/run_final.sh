#!/bin/bash
# final GPU pass on the final sources: whole GPU suite, smoke, the driver's bench command and the reference
# arm, config 4 in full on one GPU (50 000 latents; one 8 000-iteration call), then the ncu evidence of the same build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_final.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_gpu_final.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --stage-table gpurun_out/stages_cifar10.json > gpurun_out/bench_cifar10.json 2> gpurun_out/bench_cifar10.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_cifar10.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 900 python bench.py --workload svhn_test --full-50k --warmup 3 --no-secondary --no-cpu-baseline --no-eager-ref > gpurun_out/bench_svhn_test_50k_1gpu.json 2> gpurun_out/bench_svhn_test_50k.err; echo "svhn_test 50k rc=$?"; tail -2 gpurun_out/bench_svhn_test_50k.err
timeout 900 python bench.py --workload svhn_test --test-steps 8000 --steps 1 --calls-per-step 1 --warmup 3 --no-secondary --no-cpu-baseline --no-eager-ref > gpurun_out/bench_svhn_test_t8000_1gpu.json 2> gpurun_out/bench_svhn_test_t8000.err; echo "svhn_test T=8000 rc=$?"; tail -2 gpurun_out/bench_svhn_test_t8000.err
python - <<'PY'
import json, glob
for f in ['gpurun_out/bench_cifar10.json', 'gpurun_out/bench_svhn_test_50k_1gpu.json', 'gpurun_out/bench_svhn_test_t8000_1gpu.json']:
    try:
        d = json.load(open(f)); det = d['details']
        print(f.split('/')[-1], round(d['value']), 'ls/s | ms/step', round(d['ms_per_step'], 2), '| steps', d['steps'], '| frac', round(det['frac_of_tensor_roofline'], 3), '| 1p', d.get('value_bwd1pass') and round(d['value_bwd1pass']['value']), '| e2e', round(d['e2e']['value']), '| eager', d.get('reference_cuda_eager') and d['reference_cuda_eager'].get('value'), '| cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'], 1), '| prior', d.get('prior_sampling') and round(d['prior_sampling']['samples_per_sec']), '| clk', d['clocks'] and (d['clocks']['sm_mhz'], d['clocks']['reasons']))
    except Exception as e:
        print(f, 'unreadable', e)
PY
# ---- ncu (each capture only after the same command exited 0 without ncu) ----
LSNF_NO_GRAPH=1 timeout 600 python bench.py --steps 1 --warmup 3 --calls-per-step 1 --no-cpu-baseline --no-secondary --no-eager-ref > gpurun_out/plain_launch.log 2>&1 &&
LSNF_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 130 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --calls-per-step 1 --no-cpu-baseline --no-secondary --no-eager-ref > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch rc=$?"; wc -l gpurun_out/launches_r2.csv
timeout 300 python tools/prof_stage.py 1 2 5 6 > gpurun_out/plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:tapgemm_tc2_kernel -s 4 -c 4 -f -o gpurun_out/prof_r2_dominant python tools/prof_stage.py 1 2 5 6 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -1 gpurun_out/ncu_full.log
python tools/ncu_summarize.py gpurun_out/prof_r2_dominant.ncu-rep gpurun_out/r2_ncu_full_tapgemm_pair.json --dominant cifar10 100 1,2,5,6; rm -f gpurun_out/prof_r2_dominant.ncu-rep
timeout 300 python tools/prof_stage.py 0 3 4 7 > gpurun_out/plain_prof_short.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:tapgemm_tc_kernel -s 4 -c 4 -f -o gpurun_out/prof_r2_short python tools/prof_stage.py 0 3 4 7 > gpurun_out/ncu_short.log 2>&1
echo "ncu short rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_r2_short.ncu-rep gpurun_out/r2_ncu_full_short_stages.json; rm -f gpurun_out/prof_r2_short.ncu-rep
FLOW=1 timeout 300 python tools/prof_stage.py 0 > gpurun_out/plain_prof_flow.log 2>&1 &&
FLOW=1 timeout 900 ncu --set full --clock-control none -k "regex:flow_forward_kernel|flow_inverse_kernel" -s 2 -c 2 -f -o gpurun_out/prof_r2_flow python tools/prof_stage.py 0 > gpurun_out/ncu_flow.log 2>&1
echo "ncu flow rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_r2_flow.ncu-rep gpurun_out/r2_ncu_full_flow.json; rm -f gpurun_out/prof_r2_flow.ncu-rep
timeout 300 python tools/prof_train.py > gpurun_out/plain_prof_train.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -k "regex:tapgemm_tc_kernel|transpose_hl|wgrad_finalize|bias_rowsum|flow_param_grad|flow_logdet_inverse|flow_forward_kernel|adam_kernel|mse_sum" -s 30 -c 40 -f -o gpurun_out/prof_r2_train python tools/prof_train.py > gpurun_out/ncu_train.log 2>&1
echo "ncu train rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_r2_train.ncu-rep gpurun_out/r2_ncu_full_param_updates.json; rm -f gpurun_out/prof_r2_train.ncu-rep
du -sh gpurun_out
