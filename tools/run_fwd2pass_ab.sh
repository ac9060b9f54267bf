#!/bin/bash
# A/B of a two-pass forward: z_T parity at the headline shape (five seeds, B=100, T=40) and speed,
# for the product library and the two experiment builds of tools/build_exp_lib.sh
mkdir -p gpurun_out
for v in product fwd2pass1 fwd2pass2; do
  if [ $v = product ]; then unset LSNF_LIB; else export LSNF_LIB=$PWD/tools/_ab/liblsnf_$v.so; fi
  timeout 900 python -m pytest tests/test_gpu_parity_real_shapes.py -q -s -k multi_seed > gpurun_out/ab_$v.log 2>&1
  echo "== $v"; grep "^{'seed'" gpurun_out/ab_$v.log | cut -c1-200
  cp gpurun_out/parity_cifar10_b100_t40.json gpurun_out/ab_parity_$v.json
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-ref --no-secondary > gpurun_out/ab_bench_$v.json 2> gpurun_out/ab_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_bench_$v.json')); print('$v', round(d['value']), 'latent-steps/s, frac', round(d['details']['frac_of_tensor_roofline'],3), 'iteration us', round(d['roofline']['iteration_us'],1))"
done
