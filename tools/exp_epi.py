"""Times single generator stages, one synchronised launch at a time:  python tools/exp_epi.py [stage ...]
EXPS lists values of the LSNF_EXP environment variable to time each stage under (a hook for temporary experiment
switches in the kernels; the library ignores it otherwise).  With LSNF_LIB pointing at another build of the library
this gives A/B timings of two builds on one box (tools/run_ab3.sh)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import lsnf_b200
from lsnf_b200.plan import default_bwd_passes

w = dict(bench.WORKLOADS[os.environ.get("WORKLOAD", "cifar10")])
dev = torch.device("cuda:0")
args, netG, netF, gsd, fsd = bench.build_models(w, dev)
plan = lsnf_b200.langevin_plan(netG, netF, w["B"], dev, default_bwd_passes(noisy_chain=True))
plan.ensure_generator(netG)
plan.ensure_flow(netF)
x_np, z0_np, _ = lsnf_b200.synth.inputs(w["B"], w["nz"], 3, w["img"], 1, seed=1)
z = torch.from_numpy(z0_np).to(dev).reshape(w["B"], w["nz"]).contiguous()
x = torch.from_numpy(x_np).to(dev)
plan.generator_forward(z)
plan.generator_dgrad(x, w["sigma"])
torch.cuda.synchronize()


def time_stage(i, reps=15):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run_stage(i)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


variants = [int(a) for a in os.environ.get("EXPS", "0").split()]
stages = [int(a) for a in sys.argv[1:]] or list(range(len(plan.stages())))
res = {}
for i in stages:
    for v in variants:
        os.environ["LSNF_EXP"] = str(v)
        res[f"stage{i}/exp{v}"] = round(time_stage(i), 1)
    os.environ["LSNF_EXP"] = "0"
print(json.dumps(res, indent=1))
