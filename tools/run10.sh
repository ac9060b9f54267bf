#!/bin/bash
mkdir -p gpurun_out
LSNF_NO_GRAPH=1 LSNF_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/trace_cifar.json 2> gpurun_out/trace_cifar.err; grep "lsnf trace" gpurun_out/trace_cifar.err | tail -16
LSNF_NO_GRAPH=1 LSNF_TRACE=1 timeout 600 python bench.py --workload svhn --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/trace_svhn.json 2> gpurun_out/trace_svhn.err; grep "lsnf trace" gpurun_out/trace_svhn.err | tail -16
