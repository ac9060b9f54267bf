#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/summary3.txt
tail -40 gpurun_out/t_gpu.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 120 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_r1a.csv
