"""GPU parity: fused flow-prior kernels and the fused Langevin update, through the C ABI, against the reference
fixtures (tests/golden, written from the reference's own modules) and the CPU oracle."""
import numpy as np
import pytest
import torch

import lsnf_b200
from lsnf_b200 import synth
from oracle import philox, refpath
from helpers import REL_TOL, load_golden, rel_err, to_torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def flow_from_case(c):
    args = lsnf_b200.make_args(nz=c["nz"], f_width=c["f_width"], f_flow_coupling=c["coupling"])
    netF = lsnf_b200._netF(args, nz=c["nz"]).to(DEV).eval()
    netF.load_state_dict(to_torch(synth.flow_state(c["nz"], c["f_width"], 5, c["coupling"], 2, seed=1)))
    return netF


@pytest.mark.parametrize("name", ["svhn_small", "cifar_small", "svhn_additive", "hq_flow_w128"])
def test_flow_forward_logdet_logp_grad_and_inverse_match_reference(name):
    g = load_golden(name)
    c = g["config"]
    netF = flow_from_case(c)
    z = torch.from_numpy(g["z0"]).reshape(c["B"], c["nz"]).to(DEV)
    z1, logdet, logp, grad = netF.log_prior(z, want_grad=True)
    assert rel_err(z1.cpu(), g["flow_z1"]) < REL_TOL           # tolerance: 1e-4 relative, fp32 (north_star)
    assert rel_err(logdet.cpu(), g["flow_logdet"]) < REL_TOL
    assert rel_err(logp.cpu(), g["flow_ll"]) < REL_TOL
    assert rel_err(grad.cpu(), g["flow_grad"]) < REL_TOL
    e = torch.from_numpy(g["eps"][0]).reshape(c["B"], c["nz"]).to(DEV)
    e_copy = e.clone()
    zi, negobj = netF.inverse(e)
    assert torch.equal(e, e_copy), "the inverse must not modify its input (the reference does, model.py:436)"
    assert rel_err(zi.cpu(), g["flow_inv_z"]) < REL_TOL
    assert rel_err(negobj.cpu(), g["flow_inv_negobj"]) < REL_TOL
    # module-level signature of the reference (model.py:473-498)
    with torch.no_grad():
        zo, obj, extra = netF(z, objective=torch.zeros(c["B"], device=DEV))
        assert extra == [] and rel_err(obj.cpu(), g["flow_logdet"]) < REL_TOL and rel_err(zo.cpu(), g["flow_z1"]) < REL_TOL
        zr = netF(e, objective=torch.zeros(c["B"], device=DEV), reverse=True)
        assert rel_err(zr.cpu(), g["flow_inv_z"]) < REL_TOL


@pytest.mark.parametrize("batch", [1, 57, 100, 333, 1000, 3000])
def test_flow_roundtrip_and_batch_tiling(batch):
    # size-independent properties at BASELINE sizes: F^-1(F(z)) == z, -objective == logdet, every samples-per-CTA
    # variant (1, 2, 4, 8) agrees with the single-sample kernel
    c = dict(nz=128, f_width=64, coupling=1)
    netF = flow_from_case(c)
    z = torch.randn(batch, 128, device=DEV, generator=torch.Generator(DEV).manual_seed(batch))
    z1, logdet, logp, grad = netF.log_prior(z, want_grad=True)
    zr, negobj = netF.inverse(z1)
    assert rel_err(zr.cpu(), z.cpu()) < 1e-4
    assert rel_err(negobj.cpu(), logdet.cpu()) < 1e-4
    one = netF.log_prior(z[:1].contiguous(), want_grad=True)
    assert rel_err(one[0].cpu(), z1[:1].cpu()) < 1e-5 and rel_err(one[3].cpu(), grad[:1].cpu()) < 1e-5
    ll, _, _, gref = refpath.prior_grad(z[:8].cpu(), {k: v.cpu() for k, v in netF.state_dict().items()}, 5)
    assert rel_err(logp[:8].cpu(), ll) < REL_TOL and rel_err(grad[:8].cpu(), gref) < REL_TOL


def test_flow_shuffle_permutation_is_bit_exact():
    # f_flow_permutation=1: pure data movement through int32 indices -> bit-exact against index_select
    args = lsnf_b200.make_args(nz=100, f_flow_permutation=1)
    sd = synth.flow_state(100, 64, 5, 1, 1, seed=2)
    netF = lsnf_b200._netF(args, nz=100).to(DEV).eval()
    netF.load_state_dict(to_torch(sd))
    z = torch.randn(9, 100, device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    z1, logdet, logp, grad = netF.log_prior(z, want_grad=True)
    ll, z1r, ldr, gr = refpath.prior_grad(z.cpu(), to_torch(sd), 5, 1, 1)
    assert rel_err(z1.cpu(), z1r) < REL_TOL and rel_err(logdet.cpu(), ldr) < REL_TOL and rel_err(grad.cpu(), gr) < REL_TOL
    # a flow whose coupling MLP is switched off reduces to actnorm + gather: compare bit for bit
    for k in list(sd):
        if "fc_zeros" in k:
            sd[k] = np.zeros_like(sd[k])
    netF.load_state_dict(to_torch(sd))
    fp = to_torch(sd)
    z1, _, _, _ = netF.log_prior(z)
    zc = z.cpu()
    sig2 = torch.sigmoid(torch.tensor(2.0))
    for i in range(5):
        pre = f"revnet2d_s.0.revnet2d_step_s.{i}."
        zc = (zc + fp[pre + "actnorm.b"]) * torch.exp(fp[pre + "actnorm.logs"] * 3.0)
        zc = zc.index_select(1, fp[pre + "shuffle_features.indices"].long())
        zc = torch.cat([zc[:, :50], zc[:, 50:] * sig2], 1)
    # identical up to the library expf/sigmoid ulps; the gather itself moves values untouched:
    perm_only = z.cpu().index_select(1, fp["revnet2d_s.0.revnet2d_step_s.0.shuffle_features.indices"].long())
    assert rel_err(z1.cpu(), zc) < 1e-5
    assert torch.equal(perm_only.sort(dim=1).values, z.cpu().sort(dim=1).values)


def test_langevin_update_injected_noise_and_norms():
    B, nz = 37, 100
    plan = lsnf_b200.get_plan(arch="none", batch=B, nz=nz, ngf=0, nc=3, f_depth=1, f_width=4, f_permutation=2,
                              f_coupling=1, leak=0.2, device=DEV)
    gen = torch.Generator(DEV).manual_seed(3)
    z = torch.randn(B, nz, device=DEV, generator=gen)
    gg = torch.randn(B, nz, device=DEV, generator=gen) * 7
    gf = torch.randn(B, nz, device=DEV, generator=gen)
    eps = torch.randn(B, nz, device=DEV, generator=gen)
    want = z - 0.5 * 0.1 * 0.1 * (gg + gf) + 0.1 * eps           # train.py:324-326
    zz = z.clone()
    norms = plan.langevin_update(zz, gg, gf, 0.1, eps=eps)
    assert rel_err(zz.cpu(), want.cpu()) < 1e-6
    assert abs(norms[0].item() - gg.norm(dim=1).mean().item()) < 1e-4 * gg.norm(dim=1).mean().item()
    assert abs(norms[1].item() - gf.norm(dim=1).mean().item()) < 1e-4 * gf.norm(dim=1).mean().item()
    zz = z.clone()
    plan.langevin_update(zz, gg, gf, 0.1, eps=None, with_noise=False, want_norms=False)
    assert rel_err(zz.cpu(), (z - 0.005 * (gg + gf)).cpu()) < 1e-6   # test-mode variant, train.py:623


def test_langevin_update_philox_matches_oracle_and_is_shard_invariant():
    B, nz = 64, 128
    plan = lsnf_b200.get_plan(arch="none", batch=B, nz=nz, ngf=0, nc=3, f_depth=1, f_width=4, f_permutation=2,
                              f_coupling=1, leak=0.2, device=DEV)
    zero = torch.zeros(B, nz, device=DEV)
    z = torch.zeros(B, nz, device=DEV)
    plan.langevin_update(z, zero, zero, 1.0, eps=None, with_noise=True, seed=0x1234567812345678, sample_offset=1000,
                         step=7, want_norms=False)
    want = philox.langevin_noise(0x1234567812345678, 1000, B, nz, 7)
    assert np.abs(z.cpu().numpy() - want).max() < 2e-6
    # the same global samples drawn by a differently sharded call are bit-identical
    half = lsnf_b200.get_plan(arch="none", batch=B // 2, nz=nz, ngf=0, nc=3, f_depth=1, f_width=4, f_permutation=2,
                              f_coupling=1, leak=0.2, device=DEV)
    z2 = torch.zeros(B // 2, nz, device=DEV)
    half.langevin_update(z2, zero[: B // 2], zero[: B // 2], 1.0, eps=None, with_noise=True,
                         seed=0x1234567812345678, sample_offset=1000 + B // 2, step=7, want_norms=False)
    assert torch.equal(z2, z[B // 2:])
    big = torch.zeros(4096, nz, device=DEV)
    p2 = lsnf_b200.get_plan(arch="none", batch=4096, nz=nz, ngf=0, nc=3, f_depth=1, f_width=4, f_permutation=2,
                            f_coupling=1, leak=0.2, device=DEV)
    p2.langevin_update(big, torch.zeros_like(big), torch.zeros_like(big), 1.0, eps=None, with_noise=True, seed=9,
                       want_norms=False)
    assert abs(big.mean().item()) < 0.01 and abs(big.std().item() - 1.0) < 0.01
