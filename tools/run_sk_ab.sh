#!/bin/bash
# A/B of the stream-K finisher on one box: tools/_ab/liblsnf_old.so (previous commit: one helper's partial in flight),
# the in-tree build (two helpers in flight), tools/_ab/liblsnf_b3.so (three), and the cost constant of the
# stream-K decision (LSNF_SK_COST).  Per-stage tables of every BASELINE workload land in gpurun_out/sk_*.json.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_real_shapes.py tests/test_gpu_langevin.py -q -m gpu -x \
  -k "true_widths or one_step or hq256 or fixture or sharding" > gpurun_out/t_gpu_sk.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_gpu_sk.log | cut -c1-200
run() {  # tag workload
  timeout 300 python bench.py --workload $2 --steps 5 --warmup 3 --calls-per-step 2 --no-cpu-baseline --no-secondary --no-eager-ref \
    --stage-table gpurun_out/sk_stages_$1_$2.json > gpurun_out/sk_bench_$1_$2.json 2> gpurun_out/sk_bench_$1_$2.err || echo "bench $1 $2 rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/sk_bench_$1_$2.json')); s=json.load(open('gpurun_out/sk_stages_$1_$2.json'))
print('$1 $2', round(d['value']), 'ls/s', round(s['iteration_us'],1), 'us/iter', d['clocks']['sm_mhz'], [round(x['us'],1) for x in s['stages']])"
}
for wl in svhn celeba_hq256 celeba_crop cifar10; do
  export LSNF_LIB=$PWD/tools/_ab/liblsnf_old.so; run old $wl
  unset LSNF_LIB; run new $wl
  if [ $wl = svhn ] || [ $wl = celeba_hq256 ]; then export LSNF_LIB=$PWD/tools/_ab/liblsnf_b3.so; run b3 $wl; fi
  unset LSNF_LIB
done
for wl in celeba_crop cifar10; do LSNF_SK_COST=15 run new15 $wl; done
