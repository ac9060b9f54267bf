"""GPU diagnostic: runs every kernel family on small cases, compares with the CPU oracle and -- stage by stage --
the tcgen05 tap-GEMM with its SIMT twin.  Never stops at the first mismatch; writes gpurun_out/diag.json.

    python tools/gpu_diag.py [simt|tc|all] [case ...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import lsnf_b200
from lsnf_b200 import _cabi, synth
from helpers import rel_err, rel_l2, to_torch
from oracle import refpath

CASES = {
    "svhn32": dict(dataset="svhn", nz=100, ngf=32, B=6),
    "cifar32": dict(dataset="cifar10", nz=128, ngf=32, B=5),
    "celeba64": dict(dataset="celeba_crop", nz=100, ngf=64, B=3),
    "svhn64_b130": dict(dataset="svhn", nz=100, ngf=64, B=130),
}


def hl_view(plan, offset, shape_elems, fp16=False):
    """float32 view of a 16-bit hi|lo buffer in the workspace: shape [..., 2C] -> value[..., C].
    Forward activations are fp16 pairs, data-gradient tensors bf16 pairs."""
    n = int(np.prod(shape_elems))
    start = plan._ws_ptr - plan._ws.data_ptr() + offset
    raw = plan._ws[start:start + 2 * n].view(torch.float16 if fp16 else torch.bfloat16).reshape(shape_elems).float()
    c = shape_elems[-1] // 2
    return raw[..., :c] + raw[..., c:]


def run_case(name, c, impls, out):
    dev = torch.device("cuda:0")
    args = lsnf_b200.make_args(dataset=c["dataset"], nz=c["nz"], ngf=c["ngf"])
    gsd = synth.generator_state(c["dataset"], c["nz"], c["ngf"])
    fsd = synth.flow_state(c["nz"], 64)
    img = synth.image_size(c["dataset"])
    x_np, z0_np, eps_np = synth.inputs(c["B"], c["nz"], 3, img, 3, seed=5)
    layers = refpath.generator_layers(c["dataset"], c["nz"], c["ngf"])
    xh_ref, gg_ref = refpath.recon_grad(torch.from_numpy(z0_np), torch.from_numpy(x_np), to_torch(gsd), layers, 0.3)
    zT_ref, gn_ref, fn_ref = refpath.langevin(torch.from_numpy(z0_np), torch.from_numpy(x_np), to_torch(gsd),
                                              to_torch(fsd), layers, depth=5, steps=3, step_size=0.1, sigma=0.3,
                                              eps=torch.from_numpy(eps_np))
    plans = {}
    for impl in impls:
        rec = {}
        try:
            netG = lsnf_b200._netG(args).to(dev)
            netF = lsnf_b200._netF(args, nz=c["nz"]).to(dev)
            netG.load_state_dict(to_torch(gsd))
            netF.load_state_dict(to_torch(fsd))
            netG.gemm_impl = _cabi.GEMM_SIMT if impl == "simt" else _cabi.GEMM_TCGEN05
            plan = lsnf_b200.langevin_plan(netG, netF, c["B"], dev)
            plan.ensure_generator(netG)
            plan.ensure_flow(netF)
            z = torch.from_numpy(z0_np).to(dev).reshape(c["B"], c["nz"]).contiguous()
            x = torch.from_numpy(x_np).to(dev)
            t0 = time.time()
            xh = plan.generator_forward(z)
            torch.cuda.synchronize()
            rec["x_hat_rel"] = rel_err(xh.cpu().numpy(), xh_ref.numpy())
            gg = plan.generator_dgrad(x, 0.3)
            torch.cuda.synchronize()
            rec["grad_g_rel"] = rel_err(gg.cpu().numpy(), gg_ref.numpy().reshape(c["B"], c["nz"]))
            rec["fwd_bwd_s"] = time.time() - t0
            zT, gn, fn = lsnf_b200.sample_langevin_post_z_with_flow(
                torch.from_numpy(z0_np).to(dev), x, netG, netF, lsnf_b200.make_args(g_l_steps=3),
                eps=torch.from_numpy(eps_np).to(dev))
            torch.cuda.synchronize()
            rec["z_T_rel_l2"] = rel_l2(zT.cpu().numpy(), zT_ref.numpy())
            rec["z_T_rel_max"] = rel_err(zT.cpu().numpy(), zT_ref.numpy())
            rec["gnorm_g"] = [float(gn), float(gn_ref)]
            rec["gnorm_f"] = [float(fn), float(fn_ref)]
            # re-run forward+dgrad on z0 so that the workspaces of the two impls hold comparable buffers
            plan.generator_forward(z)
            plan.generator_dgrad(x, 0.3)
            torch.cuda.synchronize()
            plans[impl] = plan
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)
        out[f"{name}/{impl}"] = rec
        print(name, impl, json.dumps(rec), flush=True)
    if len(plans) == 2:
        B = c["B"]
        cmp = {}
        ps, pt = plans["simt"], plans["tc"]
        for i, info in enumerate(ps.stages()):
            if info.epilogue in (0, 2):
                hh = info.grid_h * info.out_mul
                if info.kind == 0 and info.layer == 0:
                    hh = int(round((info.n_valid // info.out_channels) ** 0.5))
                shape = (B * hh * hh, 2 * info.out_channels)
                a = hl_view(ps, info.out_offset, shape, fp16=info.epilogue == 0)
                b = hl_view(pt, info.out_offset, shape, fp16=info.epilogue == 0)
            else:
                n = info.k_splits * B * info.grid_h * info.grid_w * info.n_pad
                st = ps._ws_ptr - ps._ws.data_ptr() + info.out_offset
                a = ps._ws[st:st + 4 * n].view(torch.float32)
                st = pt._ws_ptr - pt._ws.data_ptr() + info.out_offset
                b = pt._ws[st:st + 4 * n].view(torch.float32)
            d = (a - b).abs().max().item()
            cmp[f"stage{i}_k{info.kind}_l{info.layer}"] = [d, a.abs().max().item()]
        out[f"{name}/simt_vs_tc"] = cmp
        print(name, "simt_vs_tc", json.dumps(cmp), flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    impls = ["simt", "tc"] if which == "all" else [which]
    names = sys.argv[2:] or list(CASES)
    out = {}
    for n in names:
        run_case(n, CASES[n], impls, out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"diag_{which}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
