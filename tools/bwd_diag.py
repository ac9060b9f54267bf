"""Locate mismatches between the tensor-core path and its CUDA-core twin in the data-gradient stages."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import lsnf_b200
from lsnf_b200 import _cabi, synth
from helpers import to_torch
from gpu_diag import hl_view

def main():
    ds, nz, ngf, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    dev = torch.device("cuda:0")
    args = lsnf_b200.make_args(dataset=ds, nz=nz, ngf=ngf)
    gsd = synth.generator_state(ds, nz, ngf)
    img = synth.image_size(ds)
    x_np, z0_np, _ = synth.inputs(B, nz, 3, img, 1, seed=21)
    plans = {}
    for impl in ("simt", "tc"):
        netG = lsnf_b200._netG(args).to(dev).eval(); netG.load_state_dict(to_torch(gsd))
        netG.gemm_impl = _cabi.GEMM_SIMT if impl == "simt" else _cabi.GEMM_TCGEN05
        plan = netG._plan(B, dev); plan.ensure_generator(netG)
        z = torch.from_numpy(z0_np).to(dev).reshape(B, nz).contiguous(); x = torch.from_numpy(x_np).to(dev)
        plan.generator_forward(z); g = plan.generator_dgrad(x, 0.3); torch.cuda.synchronize()
        plans[impl] = (plan, g.cpu())
    ps, pt = plans["simt"][0], plans["tc"][0]
    infos = ps.stages()
    L = len(infos) // 2
    for i, info in enumerate(infos):
        if info.epilogue not in (0, 2):
            continue
        hh = info.grid_h * info.out_mul
        if info.kind == 0 and info.layer == 0:
            hh = int(round((info.n_valid // info.out_channels) ** 0.5))
        C = info.out_channels
        if info.out_phase_split:
            shape = (4, B, hh // 2, hh // 2, 2 * C)
        else:
            shape = (1, B, hh, hh, 2 * C)
        a = hl_view(ps, info.out_offset, shape, fp16=info.epilogue == 0); b = hl_view(pt, info.out_offset, shape, fp16=info.epilogue == 0)
        d = (a - b).abs()
        scale = a.abs().max().item()
        bad = d > 1e-3 * scale
        nb = int(bad.sum())
        msg = f"stage {i} kind {info.kind} layer {info.layer} split {info.out_phase_split} shape {tuple(a.shape)} max|a| {scale:.3g} maxdiff {d.max().item():.3g} bad {nb} / {a.numel()}"
        if nb:
            idx = bad.nonzero()
            msg += f"\n   planes {sorted(set(idx[:,0].tolist()))[:8]} batch {sorted(set(idx[:,1].tolist()))[:12]}.. rows {sorted(set(idx[:,2].tolist()))[:16]} cols {sorted(set(idx[:,3].tolist()))[:16]} ch {sorted(set(idx[:,4].tolist()))[:16]}..({len(set(idx[:,4].tolist()))} distinct)"
            msg += f"\n   first bad: idx {idx[0].tolist()} simt {a[tuple(idx[0])].item():.6g} tc {b[tuple(idx[0])].item():.6g}"
        print(msg, flush=True)
    g0, g1 = plans["simt"][1], plans["tc"][1]
    e = (g0 - g1).norm(dim=1) / g0.norm(dim=1)
    print("grad_g per-sample rel-l2 simt vs tc: median %.3g max %.3g" % (e.median().item(), e.max().item()))

main()
