#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_langevin.py -q -m gpu -s -k "full_chain or long_chain" 2>&1 | grep "rel-l2\|divergence\|passed\|failed"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
LSNF_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain13.log 2>&1 &&
LSNF_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 130 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch13.log 2>&1
echo "ncu launch rc=$?"; wc -l gpurun_out/launches_r1.csv
python tools/prof_stage.py 1 2 5 6 > gpurun_out/plain_prof13.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc2_kernel -s 5 -c 4 -o gpurun_out/prof_r1_dominant python tools/prof_stage.py 1 2 5 6 > gpurun_out/ncu_full13.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full13.log
