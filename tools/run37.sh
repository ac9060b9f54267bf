#!/bin/bash
mkdir -p gpurun_out
python tools/prof_stage.py 4 5 > gpurun_out/plain_prof37.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc2_kernel -s 5 -c 2 -o gpurun_out/prof_r1_d3 -f python tools/prof_stage.py 4 5 > gpurun_out/ncu_full37.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full37.log
