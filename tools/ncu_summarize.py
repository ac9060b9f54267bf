"""Summarise an .ncu-rep on the GPU box right after the capture:  python tools/ncu_summarize.py REP OUT.json
[--dominant WORKLOAD BATCH stage,stage,...]

Reads `ncu -i REP --page raw --csv`, keeps the metrics the roofline discussion uses, one entry per captured launch.
With --dominant, also writes profiles/ncu_dominant_kernel.json: DRAM bytes per launch of the listed stages (in capture
order) together with the hash of the CUDA sources, so bench.py can refuse a capture of other kernels."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def to_bytes(v, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * mul


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    launches = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        e = {}
        for n, u, v in zip(names, units, r):
            if n in KEEP:
                try:
                    e[n] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    e[n] = {"value": v, "unit": u}
        launches.append(e)
    json.dump(launches, open(out, "w"), indent=1)
    print(f"{len(launches)} launches -> {out}")
    if "--dominant" in sys.argv:
        i = sys.argv.index("--dominant")
        workload, batch, stages = sys.argv[i + 1], int(sys.argv[i + 2]), sys.argv[i + 3].split(",")
        import bench
        per = {}
        for st, e in zip(stages, launches):   # the first pass over the listed stages
            rd, wr = e["dram__bytes_read.sum"], e["dram__bytes_write.sum"]
            per[st] = to_bytes(rd["value"], rd["unit"]) + to_bytes(wr["value"], wr["unit"])
        json.dump({"workload": workload, "batch_per_gpu": batch, "kernel": "tap-GEMM stages " + ",".join(stages),
                   "dram_bytes_per_launch_by_stage": per, "kernel_source_sha": bench.kernel_source_sha(),
                   "source": f"{os.path.basename(out)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                             "captured in the same gpurun call as the kernels' plain run)"},
                  open(os.path.join(ROOT, "gpurun_out", "ncu_dominant_kernel.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
