"""GPU parity of the training-mode parameter updates against torch.autograd +
torch.optim.Adam run on the CPU oracle: flow parameter gradients (train.py:403-411, incl. d log|det W| / dW = W^-T,
model.py:182), generator weight / bias gradients (train.py:390-394) and the fused Adam step (train.py:294-295)."""
import copy

import numpy as np
import pytest
import torch

import lsnf_b200
from lsnf_b200 import synth
from lsnf_b200.train import flow_params_in_order
from oracle import refpath
from helpers import REL_TOL, build_nets, rel_l2, to_torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _flow_net(nz, w, coupling, perm, seed):
    args = lsnf_b200.make_args(nz=nz, f_width=w, f_flow_coupling=coupling, f_flow_permutation=perm)
    netF = lsnf_b200._netF(args, nz=nz).to(DEV)
    sd = synth.flow_state(nz, w, 5, coupling, perm, seed=seed)
    netF.load_state_dict(to_torch(sd))
    return args, netF, sd


def _oracle_flow_loss(fp, z, coupling, perm, global_batch):
    ll, _, _ = refpath.log_prior(fp, z, 5, coupling, perm)        # train.py:406-409
    return -ll.sum() / global_batch                                 # train.py:410 (mean over the global batch)


def _param_keys(sd):
    return [k for k, v in sd.items() if v.dtype.kind == "f" and not k.endswith(".bias") and not k.endswith("fc_1.b")
            and not k.endswith("fc_2.b")]


@pytest.mark.parametrize("nz,w,coupling,perm,B", [
    (128, 64, 1, 2, 100),     # CIFAR-10 configuration
    (100, 64, 1, 2, 57),      # SVHN / CelebA, ragged batch
    (100, 128, 1, 2, 8),      # CelebA-HQ256 (f_width 128)
    (100, 64, 0, 2, 33),      # additive coupling
    (100, 64, 1, 1, 33),      # shuffle permutation (no W)
])
def test_flow_parameter_gradients_match_autograd(nz, w, coupling, perm, B):
    args, netF, sd = _flow_net(nz, w, coupling, perm, seed=3)
    z = torch.randn(B, nz, generator=torch.Generator().manual_seed(B))
    flat, pairs, loss = lsnf_b200.flow_gradients(netF, z.to(DEV), global_batch=B)
    keys = _param_keys(sd)
    fp = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
    leaves = {k: fp[k].clone().requires_grad_(True) for k in keys}
    want_loss = _oracle_flow_loss({**fp, **leaves}, z, coupling, perm, B)
    want_loss.backward()
    assert abs(loss.item() - want_loss.item()) < REL_TOL * abs(want_loss.item())
    got = {id(p): g for p, g in pairs}
    named = dict(netF.named_parameters())
    worst = 0.0
    for k in keys:
        g = got[id(named[k])].cpu()
        e = rel_l2(g, leaves[k].grad)
        worst = max(worst, e)
        assert e < REL_TOL, (k, e)          # 1e-4 relative (fp32), per parameter tensor
    print(f"flow parameter gradients nz={nz} w={w} coupling={coupling} perm={perm} B={B}: worst rel-l2 {worst:.2e}")
    # a two-rank split of the same batch sums to the full-batch gradient (what the all-reduce does)
    h = B // 2
    fa, _, la = lsnf_b200.flow_gradients(netF, z[:h].to(DEV), global_batch=B)
    fa = fa.clone()
    fb, _, lb = lsnf_b200.flow_gradients(netF, z[h:].to(DEV), global_batch=B)
    full, _, _ = lsnf_b200.flow_gradients(netF, z.to(DEV), global_batch=B)
    assert rel_l2((fa + fb).cpu(), full.cpu()) < 1e-5 and abs((la + lb).item() - want_loss.item()) < REL_TOL * abs(want_loss.item())


@pytest.mark.parametrize("clamp", [False, True])
def test_flow_update_matches_torch_adam_over_several_iterations(clamp):
    nz, w, B = 128, 64, 100
    args, netF, sd = _flow_net(nz, w, 1, 2, seed=5)
    args.update(f_lr=0.0004, f_is_grad_clamp=clamp, f_max_norm=2.0, f_decay=1e-4 if clamp else 0.0)
    _, optF = lsnf_b200.make_optimizers(lsnf_b200._netG(lsnf_b200.make_args(dataset="svhn", nz=nz, ngf=64)), netF, args)
    keys = _param_keys(sd)
    fp = {k: torch.from_numpy(np.ascontiguousarray(v)).clone() for k, v in sd.items()}
    leaves = [fp[k].requires_grad_(True) for k in keys]
    ref_opt = torch.optim.Adam(leaves, lr=0.0004, betas=(0.5, 0.999), weight_decay=1e-4 if clamp else 0.0)   # train.py:295
    start = {k: fp[k].detach().clone() for k in keys}
    for it in range(4):
        z = torch.randn(B, nz, generator=torch.Generator().manual_seed(40 + it))
        loss = lsnf_b200.flow_update(netF, optF, z.to(DEV), args)
        ref_opt.zero_grad()
        want = _oracle_flow_loss(fp, z, 1, 2, B)
        want.backward()
        if clamp:
            torch.nn.utils.clip_grad_norm_(leaves, 2.0)                # train.py:411-412
        ref_opt.step()
        assert abs(loss.item() - want.item()) < 1e-3 * abs(want.item())
    named = dict(netF.named_parameters())
    moved = 0.0
    for k in keys:
        a, b = named[k].detach().cpu(), fp[k].detach()
        moved = max(moved, float((b - start[k]).abs().max()))
        assert_params_close(a, b, 0.0004, 4, k)        # parameters moved by ~4 * lr = 1.6e-3
    assert moved > 5e-4
    # the optimizer state is torch.optim.Adam's: it loads into a plain Adam
    plain = torch.optim.Adam(netF.parameters(), lr=0.0004, betas=(0.5, 0.999))
    plain.load_state_dict(copy.deepcopy(optF.state_dict()))
    st = plain.state[named["revnet2d_s.0.revnet2d_step_s.0.invertible_1x1_conv.w"]]
    assert float(st["step"]) == 4.0 and st["exp_avg"].abs().sum() > 0
    # kernels re-packed the updated parameters: log p(z) through the kernel path matches the oracle on the new weights
    z = torch.randn(B, nz, generator=torch.Generator().manual_seed(99))
    netF.eval()
    _, _, logp, _ = netF.log_prior(z.to(DEV))
    ll, _, _ = refpath.log_prior({k: v.detach() for k, v in fp.items()}, z, 5)
    assert rel_l2(logp.cpu(), ll) < REL_TOL


def test_fused_adam_matches_torch_adam_elementwise():
    # the multi-tensor kernel alone: odd sizes, weight decay, several steps, a tap-major gradient layout
    gen = torch.Generator().manual_seed(0)
    shapes = [(7,), (33, 5), (128, 64, 4, 4), (3, 1000)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=gen).to(DEV)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ps]
    opt = lsnf_b200.FusedAdam(ps, lr=1e-3, betas=(0.5, 0.999), weight_decay=0.01)
    ref = torch.optim.Adam(qs, lr=1e-3, betas=(0.5, 0.999), weight_decay=0.01)
    for it in range(5):
        gs = [torch.randn(s, generator=gen) * (10.0 ** (it - 2)) for s in shapes]
        for q, g in zip(qs, gs):
            q.grad = g.clone()
        ref.step()
        dev_g, kk, inner = [], [], []
        for g, s in zip(gs, shapes):
            if len(s) == 4:   # [C_in, C_out, k, k] parameter, gradient delivered tap-major [k*k][C_in][C_out]
                dev_g.append(g.permute(2, 3, 0, 1).contiguous().to(DEV)); kk.append(s[2] * s[3]); inner.append(s[1])
            else:
                dev_g.append(g.to(DEV)); kk.append(1); inner.append(1)
        opt.fused_step(ps, dev_g, grad_kk=kk, grad_inner=inner)
    for p, q in zip(ps, qs):
        assert float((p.detach().cpu() - q.detach()).abs().max()) < 1e-6
    sd = opt.state_dict()
    assert float(sd["state"][0]["step"]) == 5.0


def assert_params_close(got, want, lr, steps, what):
    """Adam's update is lr * m / (sqrt(v) + eps): where a gradient element is zero to within its rounding error its
    SIGN decides a full +-lr step, so two fp32-accurate gradient computations legitimately differ by up to 2 * lr per
    step on a handful of elements.  Everything else must agree to 2e-5; nothing may differ by more than that bound."""
    d = (got - want).abs().flatten()
    frac = float((d > 2e-5).float().mean())
    print(f"{what}: max |diff| {float(d.max()):.2e}, fraction above 2e-5: {frac:.2e}")
    assert frac < 2e-3, (what, frac)
    assert float(d.max()) <= 2.2 * lr * steps, (what, float(d.max()))


@pytest.mark.parametrize("leak", [1.0, 0.2])
@pytest.mark.parametrize("c,B", [
    (dict(dataset="svhn", nz=100, ngf=64), 100),            # BASELINE config 1
    (dict(dataset="cifar10", nz=128, ngf=128), 100),        # BASELINE config 2 (headline)
    (dict(dataset="celeba_crop", nz=100, ngf=64), 9),       # five layers, ragged K padding
    (dict(dataset="svhn", nz=100, ngf=32), 37),             # C_in = 64 < one M tile in the last layer
    (dict(dataset="celeba_hq256", nz=100, ngf=64), 2),      # seven layers, 128-wide last grid
])
def test_generator_parameter_gradients_match_autograd(c, B, leak):
    # leak = 1.0 makes the generator kink-free (LeakyReLU(1.0) is the identity): every weight and bias gradient must
    # then agree with autograd to 1e-4 relative -- that is the arithmetic of the transposes, the weight-gradient
    # tap-GEMMs, the split-K finalize and the bias row sums.  With the reference's leak = 0.2 the gradient of a hidden
    # unit whose pre-activation is zero to within rounding error jumps by 1/leak (helpers.assert_grad_close); one such
    # unit moves a small-batch weight gradient by ~1e-3 relative, and the reference's own fp32 and fp64 gradients
    # differ by up to 7e-4 at the CIFAR-10 shape.  There the last layer (nothing kinked
    # downstream of it) is held to 1e-4 and the hidden layers to a kink's worth.
    c = dict(c, leak=leak)
    args, netG, netF = build_nets(c, DEV, seed=4)
    img = synth.image_size(c["dataset"])
    x_np, z_np, _ = synth.inputs(B, c["nz"], 3, img, 1, seed=9)
    z, x = torch.from_numpy(z_np), torch.from_numpy(x_np)
    flat, pairs, loss = lsnf_b200.generator_gradients(netG, z.to(DEV), x.to(DEV), B)
    gp = to_torch(synth.generator_state(c["dataset"], c["nz"], c["ngf"], 3, seed=4))
    leaves = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    layers = refpath.generator_layers(c["dataset"], c["nz"], c["ngf"])
    x_hat = refpath.generator_forward(leaves, z, layers, leak)                         # train.py:392
    want = torch.nn.functional.mse_loss(x_hat, x, reduction="sum") / B                 # train.py:393
    want.backward()
    assert abs(loss.item() - want.item()) < REL_TOL * want.item()
    named = dict(netG.named_parameters())
    got = {id(p): g for p, g in pairs}
    errs = {k: rel_l2(got[id(named[k])].cpu(), leaves[k].grad) for k in leaves}
    print(f"generator parameter gradients {c['dataset']} ngf={c['ngf']} B={B} leak={leak}: " +
          ", ".join(f"{k} {e:.1e}" for k, e in errs.items()))
    last = f"gen.{3 * (len(layers) - 1)}."
    for k, e in errs.items():
        assert e < (REL_TOL if (leak == 1.0 or k.startswith(last)) else 1e-2), (k, e)


def assert_updates_agree(got, want, start, lr, steps, what, strict):
    """strict (kink-free generator): element-wise agreement as assert_params_close.  With LeakyReLU kinks the gradient
    itself is only defined up to the sign decisions of near-zero pre-activations (see the gradient test above), and
    Adam turns every element's gradient into a step of size ~lr whatever its magnitude, so element-wise agreement is
    not a meaningful demand; the two updates must then point the same way (cosine > 0.98; measured 0.996 ... 1.000) and
    respect the step bound."""
    if strict:
        return assert_params_close(got, want, lr, steps, what)
    du, dr = (got - start).flatten().double(), (want - start).flatten().double()
    cos = float((du @ dr) / (du.norm() * dr.norm() + 1e-30))
    d = (got - want).abs()
    print(f"{what}: update cosine {cos:.4f}, max |diff| {float(d.max()):.2e}, fraction above 2e-5: {float((d > 2e-5).float().mean()):.2e}")
    assert cos > 0.98 and float(d.max()) <= 2.2 * lr * steps, (what, cos)


@pytest.mark.parametrize("leak", [1.0, 0.2])
def test_generator_update_matches_torch_adam_over_several_iterations(leak):
    c = dict(dataset="svhn", nz=100, ngf=64, leak=leak)
    B, lr = 100, 0.0004
    args, netG, netF = build_nets(c, DEV, seed=6)
    netG.train()
    optG, _ = lsnf_b200.make_optimizers(netG, netF, args)
    gp = to_torch(synth.generator_state("svhn", 100, 64, 3, seed=6))
    start = {k: v.clone() for k, v in gp.items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    ref_opt = torch.optim.Adam(list(leaves.values()), lr=lr, betas=(0.5, 0.999))      # train.py:294
    layers = refpath.generator_layers("svhn", 100, 64)
    for it in range(3):
        x_np, z_np, _ = synth.inputs(B, 100, 3, 32, 1, seed=50 + it)
        z, x = torch.from_numpy(z_np), torch.from_numpy(x_np)
        loss = lsnf_b200.generator_update(netG, optG, z.to(DEV), x.to(DEV), args)
        ref_opt.zero_grad()
        want = torch.nn.functional.mse_loss(refpath.generator_forward(leaves, z, layers, leak), x, reduction="sum") / B
        want.backward()
        ref_opt.step()
        assert abs(loss.item() - want.item()) < 1e-3 * want.item()
    named = dict(netG.named_parameters())
    for k in leaves:
        assert_updates_agree(named[k].detach().cpu(), leaves[k].detach(), start[k], lr, 3, k, strict=(leak == 1.0))
    # the kernels see the updated weights (re-packed on the version bump)
    netG.eval()
    z = torch.randn(5, 100, 1, 1, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        xh = netG(z.to(DEV))
    want = refpath.generator_forward({k: v.detach().cpu() for k, v in named.items()}, z, layers, leak)
    assert float((xh.cpu() - want).abs().max()) < 1e-4


@pytest.mark.parametrize("leak", [1.0, 0.2])
def test_training_iteration_matches_the_reference_iteration_on_the_oracle(leak):
    # one whole iteration of train.py:376-415 -- Langevin, G step, F step -- kernels vs oracle + autograd + torch Adam
    c = dict(dataset="svhn", nz=100, ngf=64, T=20, sigma=0.3, leak=leak)
    B, lr = 100, 0.0004
    args, netG, netF = build_nets(c, DEV, seed=2)
    optG, optF = lsnf_b200.make_optimizers(netG, netF, args)
    x_np, z0_np, _ = synth.inputs(B, 100, 3, 32, 1, seed=77)
    from oracle import philox
    seed = 0xABC
    eps = np.stack([philox.langevin_noise(seed, 0, B, 100, t) for t in range(20)]).reshape(20, B, 100, 1, 1).astype(np.float32)
    lg, lf, gn, fn, zk = lsnf_b200.training_iteration(torch.from_numpy(x_np).to(DEV), netG, netF, optG, optF, args,
                                                      seed=seed, z0=torch.from_numpy(z0_np).to(DEV))
    g0 = to_torch(synth.generator_state("svhn", 100, 64, 3, seed=2))
    gp = {k: v.clone().requires_grad_(True) for k, v in g0.items()}
    fsd = synth.flow_state(100, 64, 5, 1, 2, seed=2)
    fkeys = _param_keys(fsd)
    fp = {k: torch.from_numpy(np.ascontiguousarray(v)).clone() for k, v in fsd.items()}
    fleaves = [fp[k].requires_grad_(True) for k in fkeys]
    layers = refpath.generator_layers("svhn", 100, 64)
    z_ref, _, _ = refpath.langevin(torch.from_numpy(z0_np), torch.from_numpy(x_np), {k: v.detach() for k, v in gp.items()},
                                   {k: v.detach() for k, v in fp.items()}, layers, depth=5, steps=20, step_size=0.1,
                                   sigma=0.3, eps=torch.from_numpy(eps), leak=leak)
    assert rel_l2(zk.cpu(), z_ref) < REL_TOL
    og = torch.optim.Adam(list(gp.values()), lr=lr, betas=(0.5, 0.999))
    of = torch.optim.Adam(fleaves, lr=lr, betas=(0.5, 0.999))
    x = torch.from_numpy(x_np)
    loss_g = torch.nn.functional.mse_loss(refpath.generator_forward(gp, z_ref, layers, leak), x, reduction="sum") / B
    loss_g.backward(); og.step()
    loss_f = _oracle_flow_loss(fp, z_ref.reshape(B, 100), 1, 2, B)
    loss_f.backward(); of.step()
    assert abs(lg.item() - loss_g.item()) < 1e-3 * loss_g.item() and abs(lf.item() - loss_f.item()) < 1e-3 * abs(loss_f.item())
    ng, nf = dict(netG.named_parameters()), dict(netF.named_parameters())
    for k in gp:
        assert_updates_agree(ng[k].detach().cpu(), gp[k].detach(), g0[k], lr, 1, k, strict=(leak == 1.0))
    for k in fkeys:
        assert_params_close(nf[k].detach().cpu(), fp[k].detach(), lr, 1, k)


def test_generator_gradients_with_the_single_pass_data_gradient(monkeypatch):
    # opt-in reduced-precision mode (LSNF_BWD_PASSES=1): the gradient tensors hold fp16 hi halves only; the transposes
    # and weight-gradient GEMMs must read them as such.  Kink-free generator; tolerance = the 11-bit significand.
    monkeypatch.setenv("LSNF_BWD_PASSES", "1")
    lsnf_b200.clear_plans()
    try:
        c = dict(dataset="svhn", nz=100, ngf=64, leak=1.0)
        B = 64
        args, netG, netF = build_nets(c, DEV, seed=4)
        x_np, z_np, _ = synth.inputs(B, 100, 3, 32, 1, seed=9)
        z, x = torch.from_numpy(z_np), torch.from_numpy(x_np)
        flat, pairs, loss = lsnf_b200.generator_gradients(netG, z.to(DEV), x.to(DEV), B)
        leaves = {k: v.clone().requires_grad_(True) for k, v in to_torch(synth.generator_state("svhn", 100, 64, 3, seed=4)).items()}
        layers = refpath.generator_layers("svhn", 100, 64)
        want = torch.nn.functional.mse_loss(refpath.generator_forward(leaves, z, layers, 1.0), x, reduction="sum") / B
        want.backward()
        named = dict(netG.named_parameters())
        got = {id(p): g for p, g in pairs}
        errs = {k: rel_l2(got[id(named[k])].cpu(), leaves[k].grad) for k in leaves}
        print("single-pass data gradient, weight gradients:", ", ".join(f"{k} {e:.1e}" for k, e in errs.items()))
        assert max(errs.values()) < 2e-3 and errs["gen.9.weight"] < REL_TOL
    finally:
        lsnf_b200.clear_plans()
