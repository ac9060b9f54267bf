#!/bin/bash
# first GPU round: safe kernels first, then the tensor-core path under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_flow_update.py -q -x > gpurun_out/t_flow.log 2>&1; echo "flow tests rc=$?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/t_flow.log
timeout 600 python tools/gpu_diag.py simt svhn32 cifar32 > gpurun_out/diag_simt.log 2>&1; echo "diag simt rc=$?" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/diag_simt.log
timeout 600 python tools/gpu_diag.py all svhn32 cifar32 celeba64 svhn64_b130 > gpurun_out/diag_all.log 2>&1; echo "diag all rc=$?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/diag_all.log
