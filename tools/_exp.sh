mkdir -p gpurun_out
python tools/time_update.py cifar10 2>&1 | grep "ensure_flow\|flow_gradients"
for wl in svhn cifar10; do
timeout 900 python bench.py --mode train --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_train_${wl}_1gpu.json 2> gpurun_out/bench_train_$wl.err; echo "train $wl rc=$?"; tail -2 gpurun_out/bench_train_$wl.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_train_${wl}_1gpu.json')); det=d['details']; print('$wl', round(d['value']), 'ls/s | ms/iteration', round(det['ms_per_iteration'],2), '| langevin', round(det['langevin_call_ms'],2), '| updates', round(det['updates_ms'],2))"
done
bash tools/run_refresh.sh
for wl in celeba_crop celeba_hq256; do
timeout 900 python bench.py --mode train --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_train_${wl}_1gpu.json 2> gpurun_out/bench_train_$wl.err; echo "train $wl rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_train_${wl}_1gpu.json')); det=d['details']; print('$wl', round(d['value']), 'ls/s | ms/iteration', round(det['ms_per_iteration'],2), '| langevin', round(det['langevin_call_ms'],2), '| updates', round(det['updates_ms'],2))"
done
