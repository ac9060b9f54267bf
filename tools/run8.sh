#!/bin/bash
mkdir -p gpurun_out
python tools/prof_stage.py 1 5 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc2_kernel -s 5 -c 2 -o gpurun_out/prof_r1_pair python tools/prof_stage.py 1 5 > gpurun_out/ncu_full8.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full8.log; ls -la gpurun_out/*.ncu-rep
