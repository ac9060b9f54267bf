"""Runs the generator stages of the CIFAR-10 configuration a few times (for ncu): python tools/prof_stage.py [stage ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import lsnf_b200

w = dict(bench.ALL_WORKLOADS[os.environ.get("WORKLOAD", "cifar10")])
if os.environ.get("BATCH"):
    w["B"] = int(os.environ["BATCH"])
dev = torch.device("cuda:0")
args, netG, netF, gsd, fsd = bench.build_models(w, dev)
from lsnf_b200.plan import default_bwd_passes
plan = lsnf_b200.langevin_plan(netG, netF, w["B"], dev, default_bwd_passes())
plan.ensure_generator(netG)
plan.ensure_flow(netF)
x_np, z0_np, _ = lsnf_b200.synth.inputs(w["B"], w["nz"], 3, w["img"], 1, seed=1)
z = torch.from_numpy(z0_np).to(dev).reshape(w["B"], w["nz"]).contiguous()
x = torch.from_numpy(x_np).to(dev)
plan.generator_forward(z)
plan.generator_dgrad(x, w["sigma"])
torch.cuda.synchronize()
stages = [int(a) for a in sys.argv[1:]] or list(range(len(plan.stages())))
for rep in range(3):
    for i in stages:
        plan.run_stage(i)
if os.environ.get("FLOW"):   # the flow-prior kernel alone (forward + analytic backward), and the inverse
    plan.ensure_flow(netF, need_inverse=True)
    for rep in range(3):
        plan.flow_forward(z, want_grad=True)
        plan.flow_inverse(z)
torch.cuda.synchronize()
print("ok")
