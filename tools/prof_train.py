"""One generator update + one flow update at a BASELINE configuration (for ncu): python tools/prof_train.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import lsnf_b200

w = dict(bench.ALL_WORKLOADS[os.environ.get("WORKLOAD", "cifar10")])
dev = torch.device("cuda:0")
args, netG, netF, gsd, fsd = bench.build_models(w, dev)
netG.train(); netF.train()
optG, optF = lsnf_b200.make_optimizers(netG, netF, args)
x_np, z0_np, _ = lsnf_b200.synth.inputs(w["B"], w["nz"], 3, w["img"], 1, seed=1)
z = torch.from_numpy(z0_np).to(dev)
x = torch.from_numpy(x_np).to(dev)
for rep in range(2):   # the first repetition warms up (module load, plan creation); ncu skips it with -s
    lsnf_b200.generator_update(netG, optG, z, x, args)
    lsnf_b200.flow_update(netF, optF, z, args)
    torch.cuda.synchronize()
print("ok")
