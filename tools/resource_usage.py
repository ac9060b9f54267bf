"""Per-kernel resource usage of the shipped library:  python tools/resource_usage.py [out.json]

`cuobjdump --dump-resource-usage` on latent-space-normalizing-flow_b200/_lib/liblsnf_b200.so: registers per thread,
stack bytes (> 0 = spills or a local array), static shared memory, constant bank 0 (kernel parameters; the tap-GEMM
kernels carry their tensor maps there as __grid_constant__).  Static evidence only -- no GPU needed."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "latent-space-normalizing-flow_b200", "_lib", "liblsnf_b200.so")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_resource_usage.json")
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("lsnf::", "")
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+) .*?CONSTANT\[0\]:(\d+)", line)
        if m and name:
            res[name] = {"registers": int(m.group(1)), "stack_bytes": int(m.group(2)), "static_shared_bytes": int(m.group(3)),
                         "local_bytes": int(m.group(4)), "constant0_bytes": int(m.group(5))}
            name = None
    src = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import bench; "
                          "print(bench.kernel_source_sha())" % ROOT], capture_output=True, text=True).stdout.strip()
    doc = {"what": "cuobjdump --dump-resource-usage of liblsnf_b200.so (sm_100a), per kernel", "kernel_source_sha": src,
           "kernels_with_stack": sorted(k for k, v in res.items() if v["stack_bytes"]), "kernels": res}
    json.dump(doc, open(out, "w"), indent=1)
    for k, v in sorted(res.items(), key=lambda kv: -kv[1]["registers"])[:12]:
        print(f"{v['registers']:4d} regs {v['stack_bytes']:4d} B stack  {k}")
    print("kernels:", len(res), " with stack:", doc["kernels_with_stack"])


if __name__ == "__main__":
    main()
