"""Are the per-sample gradient mismatches of the tensor-core path LeakyReLU-kink flips?  For each fixture: count the
hidden units whose sign differs between the kernel's saved activations and an fp64 forward pass, and recompute the
fp64 gradient with the KERNEL's masks: if that matches the kernel's gradient, the mismatch is entirely the kinks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import torch
import torch.nn.functional as F

import lsnf_b200
from lsnf_b200 import _cabi, synth
from helpers import load_golden, per_sample_rel_l2, to_torch
from gpu_diag import hl_view
from oracle import refpath


def main():
    out = {}
    dev = torch.device("cuda:0")
    for name in sys.argv[1:] or ["svhn_small", "cifar_small", "celeba_small"]:
        g = load_golden(name)
        c = g["config"]
        B, nz = c["B"], c["nz"]
        layers = refpath.generator_layers(c["dataset"], nz, c["ngf"])
        gsd = synth.generator_state(c["dataset"], nz, c["ngf"])
        gp64 = to_torch(gsd, torch.float64)
        # fp64 forward with pre-activations
        h = torch.from_numpy(g["z0"]).double()
        pre = []
        for i, (ci, co, k, s, p) in enumerate(layers):
            h = F.conv_transpose2d(h, gp64[f"gen.{3*i}.weight"], gp64[f"gen.{3*i}.bias"], stride=s, padding=p)
            if i < len(layers) - 1:
                pre.append(h)
                h = F.leaky_relu(h, 0.2)
        xhat = torch.tanh(h)
        x = torch.from_numpy(g["x"]).double()
        for impl in ("tcgen05", "simt"):
            args = lsnf_b200.make_args(dataset=c["dataset"], nz=nz, ngf=c["ngf"])
            netG = lsnf_b200._netG(args).to(dev).eval()
            netG.load_state_dict(to_torch(gsd))
            netG.gemm_impl = _cabi.GEMM_SIMT if impl == "simt" else _cabi.GEMM_TCGEN05
            plan = netG._plan(B, dev)
            plan.ensure_generator(netG)
            z = torch.from_numpy(g["z0"]).to(dev).reshape(B, nz).contiguous()
            plan.generator_forward(z)
            gg = plan.generator_dgrad(torch.from_numpy(g["x"]).to(dev), c["sigma"]).cpu().double()
            infos = plan.stages()
            masks, flips = [], []
            for l in range(len(layers) - 1):
                info = infos[l]
                hh = pre[l].shape[-1]
                a = hl_view(plan, info.out_offset, (B, hh, hh, 2 * info.out_channels), fp16=True).cpu().double().permute(0, 3, 1, 2)
                mism = (a > 0) != (pre[l] > 0)
                flips.append(dict(layer=l, n=int(mism.sum()), per_sample=mism.flatten(1).sum(1).tolist(),
                                  abs_pre=[float(v) for v in pre[l][mism].abs()[:6]],
                                  max_abs_err=float((a - F.leaky_relu(pre[l], 0.2)).abs().max())))
                masks.append(torch.where(a > 0, 1.0, 0.2))
            gr = (xhat - x) / c["sigma"] ** 2 * (1 - xhat ** 2)
            for l in range(len(layers) - 1, 0, -1):
                ci, co, k, s, p = layers[l]
                gr = F.conv2d(gr, gp64[f"gen.{3*l}.weight"], stride=s, padding=p) * masks[l - 1]
            gz = F.conv2d(gr, gp64["gen.0.weight"], stride=1, padding=0).reshape(B, nz)
            e_ref = per_sample_rel_l2(gg.numpy(), g["grad_g"].reshape(B, nz))
            e_masked = per_sample_rel_l2(gg.numpy(), gz.numpy())
            out[f"{name}/{impl}"] = dict(flips=flips, err_vs_reference=e_ref.tolist(), err_vs_fp64_with_kernel_masks=e_masked.tolist())
            print(name, impl, json.dumps(out[f"{name}/{impl}"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kink_diag.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
