"""Host vs device time of the pieces of one parameter update: python tools/time_update.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import lsnf_b200
from lsnf_b200 import train as T

wl = sys.argv[1] if len(sys.argv) > 1 else "svhn"
w = dict(bench.ALL_WORKLOADS[wl])
dev = torch.device("cuda:0")
args, netG, netF, gsd, fsd = bench.build_models(w, dev)
netG.train(); netF.train()
optG, optF = lsnf_b200.make_optimizers(netG, netF, args)
x_np, z0_np, _ = lsnf_b200.synth.inputs(w["B"], w["nz"], 3, w["img"], 1, seed=1)
z = torch.from_numpy(z0_np).to(dev); x = torch.from_numpy(x_np).to(dev)
B = w["B"]


def timed(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); host = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    print(f"{name:34s} host {host:7.3f} ms   device {e0.elapsed_time(e1) / reps:7.3f} ms")


plan = netG._plan(B, dev, train=True)
timed("generator_gradients", lambda: T.generator_gradients(netG, z, x, B, plan=plan))
flat, pairs, loss = T.generator_gradients(netG, z, x, B, plan=plan)
params, grads = [p for p, _ in pairs], [v for _, v in pairs]
timed("optG.fused_step", lambda: optG.fused_step(params, grads))
timed("ensure_generator (re-pack)", lambda: (lsnf_b200.optim.bump_versions(params), plan.ensure_generator(netG)))
timed("flow_gradients (incl. ensure_flow)", lambda: (lsnf_b200.optim.bump_versions(list(netF.parameters())), T.flow_gradients(netF, z, B)))
fflat, fpairs, floss = T.flow_gradients(netF, z, B)
fparams, fgrads = [p for p, _ in fpairs], [v for _, v in fpairs]
timed("optF.fused_step", lambda: optF.fused_step(fparams, fgrads))
fplan = netF._plan(B, dev)
timed("ensure_flow alone (LU + pack)", lambda: (lsnf_b200.optim.bump_versions(list(netF.parameters())), fplan.ensure_flow(netF, need_inverse=True)))
timed("generator_update + flow_update", lambda: (T.generator_update(netG, optG, z, x, args, plan=plan), T.flow_update(netF, optF, z, args)))
