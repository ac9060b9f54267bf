"""The operand-split emulation behind profiles/r2_precision_emulation_cpu.json (tools/precision_emulation.py) does what
its docstring says: hi|lo pairs carry 22 / 16 significant bits, the e4m3 roundings stay within their format's error
bound, the emulated ConvTranspose2d has the autograd data gradient, and on a small chain the schemes order as on the
B200 (three passes ~ fp32, a dropped cross term is an order of magnitude worse)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import precision_emulation as pe  # noqa: E402
from lsnf_b200 import synth  # noqa: E402
from oracle import refpath  # noqa: E402


def test_hi_lo_pairs_carry_the_stated_bits():
    g = torch.Generator().manual_seed(0)
    t = torch.randn(4096, generator=g)
    hi, lo = pe.split(t, torch.float16)
    assert float(((hi + lo) - t).abs().max() / t.abs().max()) < 2.0 ** -21
    hi, lo = pe.split(t, torch.bfloat16)
    assert float((((hi + lo) - t).abs() / t.abs()).max()) < 2.0 ** -15
    assert torch.equal(hi, t.to(torch.bfloat16).float())


def test_e4m3_roundings_stay_within_the_format_error():
    g = torch.Generator().manual_seed(1)
    t = torch.randn(8, 64, 5, 5, generator=g) * torch.logspace(-3, 0, 64).view(1, 64, 1, 1)
    q = pe.q8_block(t, axis=1)
    assert q.shape == t.shape
    blocks = t.movedim(1, -1).reshape(-1, 2, 32)
    err = (q - t).movedim(1, -1).reshape(-1, 2, 32).abs().amax(-1)
    assert bool((err <= blocks.abs().amax(-1) * 2.0 ** -4).all())          # half an ulp of the block's largest binade
    big = t.abs() >= t.abs().amax() * 2.0 ** -6                               # normal range of the per-tensor scale
    qt = pe.q8_tensor(t)
    assert float((((qt - t).abs() / t.abs())[big]).max()) <= 2.0 ** -4
    # ragged contraction axis (100 channels: nz = 100) is padded, not dropped
    r = torch.randn(3, 100, generator=g)
    assert pe.q8_block(r, axis=1).shape == r.shape


def test_emulated_conv_transpose_matches_torch_forward_and_data_gradient():
    g = torch.Generator().manual_seed(2)
    a = torch.randn(2, 8, 4, 4, generator=g, requires_grad=True)
    w = torch.randn(8, 6, 4, 4, generator=g) * 0.1
    b = torch.randn(6, generator=g)
    fp32 = pe.FWD["fp32"]
    y = pe.EmuConvT.apply(a, w, b, 2, 1, fp32, pe.BWD["fp32"])
    ref = F.conv_transpose2d(a, w, b, 2, 1)
    assert torch.allclose(y, ref, atol=1e-6)
    go = torch.randn_like(ref)
    (ga,) = torch.autograd.grad(y, a, go)
    (gr,) = torch.autograd.grad(ref, a, go)
    assert torch.allclose(ga, gr, atol=1e-5)
    # the 3-pass splits reproduce both to a few 1e-7 relative, a single fp16 pass only to ~1e-3
    y3 = pe.EmuConvT.apply(a, w, b, 2, 1, pe.FWD["f16x3"], pe.BWD["bf16x3"])
    y1 = pe.EmuConvT.apply(a, w, b, 2, 1, pe.FWD["f16x1"], pe.BWD["f16x1"])
    e3 = float((y3 - ref).detach().norm() / ref.detach().norm())
    e1 = float((y1 - ref).detach().norm() / ref.detach().norm())
    assert e3 < 2e-6 < 1e-4 < e1 < 2e-3, (e3, e1)
    (g3,) = torch.autograd.grad(y3, a, go)
    assert float((g3 - gr).norm() / gr.norm()) < 2e-5                         # bf16 hi|lo: 16 significant bits


def test_schemes_order_on_a_small_chain_as_on_the_b200():
    ds, nz, ngf, B, T = "svhn", 100, 16, 4, 6
    layers = synth.generator_layers(ds, nz, ngf)
    to_t = lambda sd: {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
    gp, fp = to_t(synth.generator_state(ds, nz, ngf)), to_t(synth.flow_state(nz, 64))
    x, z0, eps = (torch.from_numpy(v) for v in synth.inputs(B, nz, 3, 32, T))
    kw = dict(depth=5, steps=T, step_size=0.1, sigma=0.3, leak=0.2)
    ref, _, _ = refpath.langevin(z0, x, gp, fp, layers, eps=eps, **kw)
    err = {}
    for f, b in (("fp32", "fp32"), ("f16x3", "bf16x3"), ("f16+f8b", "f16+f8b"), ("f16x1", "f16x1")):
        zt = pe.langevin_emulated(z0, x, gp, fp, layers, eps=eps, fwd=pe.FWD[f], bwd=pe.BWD[b], **kw)
        err[f] = pe.rel_l2(zt, ref)
    assert err["fp32"] < 1e-6                       # the emulated graph IS the oracle's when nothing is rounded
    assert err["f16x3"] < 1e-4
    assert err["f16+f8b"] < 1e-4
    assert err["f16x1"] > 3 * max(err["f16x3"], 1e-6)
    assert pe.FWD["f16x3"].passes + pe.BWD["bf16x3"].passes == 6.0
    assert pe.FWD["f16+f8b"].passes + pe.BWD["f16+f8b"].passes == 4.0
