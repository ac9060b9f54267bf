#!/bin/bash
# round 2, second GPU pass: the parameter-update kernels (flow gradients, weight gradients, fused Adam), then the
# whole suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_param_updates.py -q -s -x > gpurun_out/t_gpu_updates.log 2>&1; echo "updates rc=$?"; grep -v "^$" gpurun_out/t_gpu_updates.log | tail -40 | cut -c1-400
timeout 2400 python -m pytest tests -q -m gpu -s --deselect tests/test_gpu_param_updates.py > gpurun_out/t_gpu_r2b.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/t_gpu_r2b.log | cut -c1-300
