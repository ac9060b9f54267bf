#!/bin/bash
# ncu evidence for profiles/: launch list of one Langevin call (eager, so every kernel is a launch) and a full
# capture of the wide tap-GEMM launches (stages 1 2 = 3-pass forward, 5 6 = single-pass data gradient)
mkdir -p gpurun_out
LSNF_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 &&
LSNF_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 130 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch rc=$?"; wc -l gpurun_out/launches_r1.csv
python tools/prof_stage.py 1 2 5 6 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc2_kernel -s 4 -c 4 -f -o gpurun_out/prof_r1_dominant python tools/prof_stage.py 1 2 5 6 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
