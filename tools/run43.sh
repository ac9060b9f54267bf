#!/bin/bash
mkdir -p gpurun_out
LSNF_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:last_fused -s 3 -c 1 -o gpurun_out/prof_r1_fused -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_fused.log
