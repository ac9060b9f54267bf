#!/bin/bash
# A/B of programmatic dependent launch across the kernel boundaries of the Langevin loop (LSNF_PDL modes 0-4, see
# plan.cu: langevin_loop) on one box: Langevin parity tests under LSNF_PDL=1, then per-stage tables and loop
# throughput of every BASELINE workload per mode -> gpurun_out/pdl_*.json
mkdir -p gpurun_out
LSNF_PDL=1 timeout 600 python -m pytest tests/test_gpu_langevin.py tests/test_gpu_parity_real_shapes.py tests/test_gpu_sampling_and_long_chains.py \
  -q -m gpu -x -k "not multi_seed and not two_devices and not two_plans and not small_sigma" > gpurun_out/t_gpu_pdl.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_gpu_pdl.log | cut -c1-200
run() {  # mode workload
  LSNF_PDL=$1 timeout 300 python bench.py --workload $2 --steps 5 --warmup 3 --calls-per-step 4 --no-cpu-baseline --no-secondary --no-eager-ref \
    --stage-table gpurun_out/pdl_stages_$1_$2.json > gpurun_out/pdl_bench_$1_$2.json 2> gpurun_out/pdl_bench_$1_$2.err || echo "bench $1 $2 rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/pdl_bench_$1_$2.json')); s=json.load(open('gpurun_out/pdl_stages_$1_$2.json'))
print('pdl=$1 $2', round(d['value']), 'ls/s', round(1e3*d['ms_per_step']/d['details']['langevin_calls_per_step']/d['config']['g_l_steps'],1), 'us/iter in the loop | stage sum', round(sum(x['us'] for x in s['stages']),1), d['clocks']['sm_mhz'])"
}
for wl in svhn celeba_hq256 celeba_crop cifar10; do
  for m in 0 1 2 3 4; do run $m $wl; done
done
