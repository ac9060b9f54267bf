"""Pin the oracle against the reference's own modules and write tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference, which is absent on the GPU box):

    python oracle/make_golden.py            # checks + (re)writes fixtures

For every case it
  1. builds the reference ``_netG`` / ``_netF`` (imported unmodified from /root/reference/model.py),
     loads the deterministic synthetic parameters of ``lsnf_b200.synth`` into them,
  2. runs the reference closure -- train.py:307-335 restated around the *reference modules*, noise injected --
     plus single evaluations of x_hat, the two gradients, log p(z), log-det and the flow inverse,
  3. asserts that ``oracle/refpath.py`` (which never touches the reference modules) reproduces them,
  4. stores the reference outputs, fp64 truth for z_T, and checksums of the synthetic parameters.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("LSNF_REFERENCE", "/root/reference")

import lsnf_b200.synth as synth  # noqa: E402
from oracle import refpath  # noqa: E402

CASES = {
    # name: dataset, nz, ngf, f_width, B, T, sigma, coupling
    "svhn_small": dict(dataset="svhn", nz=100, ngf=32, f_width=64, B=6, T=4, sigma=0.3, coupling=1),
    "cifar_small": dict(dataset="cifar10", nz=128, ngf=32, f_width=64, B=5, T=3, sigma=0.3, coupling=1),
    "celeba_small": dict(dataset="celeba_crop", nz=100, ngf=64, f_width=64, B=3, T=2, sigma=0.3, coupling=1),
    "svhn_additive": dict(dataset="svhn", nz=100, ngf=32, f_width=64, B=4, T=2, sigma=0.3, coupling=0),
    "hq_flow_w128": dict(dataset=None, nz=100, ngf=0, f_width=128, B=7, T=0, sigma=1.0, coupling=1),
}


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def ref_args(c):
    return AttrDict(dataset=c["dataset"], nz=c["nz"], ngf=c["ngf"], nc=3, g_activation="lrelu",
                    g_activation_leak=0.2, g_batchnorm=False, f_n_levels=1, f_depth=5, f_flow_permutation=2,
                    f_width=c["f_width"], f_flow_coupling=c["coupling"])


def to_torch(sd, dtype=torch.float32):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in sd.items()}


def reference_langevin(netG, netF, z, x, steps, step_size, sigma, eps):
    """train.py:307-335 around the reference modules; the one change is eps[t] for randn_like."""
    mse = torch.nn.MSELoss(reduction="sum")
    z = z.clone().detach()
    z.requires_grad = True
    bsz = z.shape[0]
    for i in range(steps):
        x_hat = netG(z)
        g_log_lkhd = 1.0 / (2.0 * sigma * sigma) * mse(x_hat, x)
        z_grad_g = torch.autograd.grad(g_log_lkhd, z)[0]
        z1, logdet, _ = netF(torch.squeeze(z), objective=torch.zeros(int(z.shape[0]), dtype=z.dtype), init=False)
        prior_ll = -0.5 * (z1 ** 2)
        prior_ll = prior_ll.flatten(1).sum(-1) + np.log(2 * np.pi)
        ll = prior_ll + logdet
        f_log_lkhd = -ll.sum()
        z_grad_f = torch.autograd.grad(f_log_lkhd, z)[0]
        z.data = z.data - 0.5 * step_size * step_size * (z_grad_g + z_grad_f)
        if eps is not None:
            z.data += step_size * eps[i]
        gn = z_grad_g.view(bsz, -1).norm(dim=1).mean()
        fn = z_grad_f.view(bsz, -1).norm(dim=1).mean()
    return z.detach(), gn, fn


def reference_parameter_updates(netG, netF, optG, optF, z_k, x):
    """train.py:390-415 around the reference modules and torch.optim.Adam."""
    mse = torch.nn.MSELoss(reduction="sum")
    optG.zero_grad()
    x_hat = netG(z_k.detach())
    loss_g = mse(x_hat, x) / x.shape[0]
    loss_g.backward()
    optG.step()
    optF.zero_grad()
    z1, logdet, _ = netF(torch.squeeze(z_k), objective=torch.zeros(int(z_k.shape[0])), init=False)
    prior_ll = -0.5 * (z1 ** 2)
    prior_ll = prior_ll.flatten(1).sum(-1) + np.log(2 * np.pi)
    ll = prior_ll + logdet
    loss_f = -ll.mean()
    loss_f.backward()
    optF.step()
    return loss_g.detach(), loss_f.detach()


TRAIN_CASE = dict(dataset="svhn", nz=100, ngf=32, f_width=64, B=6, T=4, sigma=0.3, coupling=1, lr=0.0004, iters=2)
TRAIN_KEEP = ["gen.9.weight", "gen.9.bias", "gen.0.bias", "revnet2d_s.0.revnet2d_step_s.0.invertible_1x1_conv.w",
              "revnet2d_s.0.revnet2d_step_s.0.actnorm.logs", "revnet2d_s.0.revnet2d_step_s.4.f.fc_zeros.b",
              "revnet2d_s.0.revnet2d_step_s.2.f.fc_1.w"]


def training_iteration_case(ref_model, out_dir):
    """Two whole training iterations (train.py:384-415: Langevin, generator step, flow step; Adam as train.py:294-295)
    on the reference modules against ``refpath.langevin`` + ``refpath.parameter_updates`` on plain leaf tensors."""
    c = TRAIN_CASE
    args = ref_args(c)
    gsd = synth.generator_state(c["dataset"], c["nz"], c["ngf"], 3, seed=1)
    fsd = synth.flow_state(c["nz"], c["f_width"], 5, c["coupling"], 2, seed=1)
    netG, netF = ref_model._netG(args), ref_model._netF(args, nz=c["nz"])
    netG.load_state_dict(to_torch(gsd))
    netF.load_state_dict(to_torch(fsd))
    adam = lambda ps: torch.optim.Adam(ps, lr=c["lr"], weight_decay=0, betas=(0.5, 0.999))
    optG, optF = adam(netG.parameters()), adam(netF.parameters())
    gp = {k: v.clone().requires_grad_(True) for k, v in to_torch(gsd).items()}
    fp = to_torch(fsd)
    fkeys = refpath.trainable_flow_keys(fp)
    for k in fkeys:
        fp[k] = fp[k].clone().requires_grad_(True)
    o_optG, o_optF = adam(list(gp.values())), adam([fp[k] for k in fkeys])
    layers = refpath.generator_layers(c["dataset"], c["nz"], c["ngf"], 3)
    losses = []
    for it in range(c["iters"]):
        x_np, z0_np, eps_np = synth.inputs(c["B"], c["nz"], 3, 32, c["T"], seed=100 + it)
        x, z0, eps = torch.from_numpy(x_np), torch.from_numpy(z0_np), torch.from_numpy(eps_np)
        z_k, _, _ = reference_langevin(netG, netF, z0, x, c["T"], 0.1, c["sigma"], eps)
        lg, lf = reference_parameter_updates(netG, netF, optG, optF, z_k, x)
        o_zk, _, _ = refpath.langevin(z0, x, {k: v.detach() for k, v in gp.items()}, {k: v.detach() for k, v in fp.items()},
                                      layers, depth=5, steps=c["T"], step_size=0.1, sigma=c["sigma"], eps=eps)
        o_lg, o_lf = refpath.parameter_updates(gp, fp, o_zk, x, layers, o_optG, o_optF, depth=5)
        close(o_zk, z_k, 1e-6, "z_k")
        close(o_lg, lg, 1e-6, "loss_g")
        close(o_lf, lf, 1e-6, "loss_f")
        losses.append([lg.item(), lf.item()])
    rg, rf = netG.state_dict(), netF.state_dict()
    worst = 0.0
    for k in gp:
        worst = max(worst, close(gp[k].detach(), rg[k], 1e-6, k))
    for k in fp:
        if fp[k].is_floating_point():
            # ``actnorm.bias`` is the same Parameter as ``actnorm.b`` in the module (model.py:231): a plain dict holds
            # two tensors, only ``b`` is read and updated
            src = k[:-len("bias")] + "b" if k.endswith("actnorm.bias") else k
            worst = max(worst, close(fp[src].detach(), rf[k], 1e-6, k))
    moved = max(float((rg[k] - to_torch(gsd)[k]).abs().max()) for k in gp)
    print("train_update", "losses", losses, "parameters after", c["iters"], "iterations: max rel err", worst,
          "largest generator update", moved)
    out = dict(config=np.array(repr(c)), losses=np.array(losses, np.float64),
               g_checksum_after=np.array(synth.checksum({k: v.numpy() for k, v in rg.items()})),
               f_checksum_after=np.array(synth.checksum({k: v.numpy() for k, v in rf.items() if v.is_floating_point()})))
    for k in TRAIN_KEEP:
        out["after:" + k] = (rg[k] if k in rg else rf[k]).numpy()
    np.savez_compressed(os.path.join(out_dir, "train_update_svhn_small.npz"), **out)
    print("wrote train_update_svhn_small")


def close(a, b, tol, what):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol}"
    return err


def main():
    sys.path.insert(0, REF)
    import model as ref_model  # the unmodified reference

    torch.manual_seed(0)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        args = ref_args(c)
        fsd = synth.flow_state(c["nz"], c["f_width"], 5, c["coupling"], 2, seed=1)
        netF = ref_model._netF(args, nz=c["nz"])
        assert sorted(netF.state_dict().keys()) == sorted(fsd.keys()), "flow state_dict keys differ"
        netF.load_state_dict(to_torch(fsd))
        fp = to_torch(fsd)
        out = dict(case=np.array(name), f_checksum=np.array(synth.checksum(fsd)))
        B, nz, T = c["B"], c["nz"], c["T"]
        img = synth.image_size(c["dataset"]) if c["dataset"] else 32
        x_np, z0_np, eps_np = synth.inputs(B, nz, 3, img, max(T, 1), seed=1)
        z0 = torch.from_numpy(z0_np)
        z2d = z0.reshape(B, nz)

        # flow forward / log p / gradient / inverse: reference modules
        zr = z2d.clone().requires_grad_(True)
        z1, logdet, _ = netF(zr, objective=torch.zeros(B), init=False)
        ll = (-0.5 * z1 ** 2).flatten(1).sum(-1) + np.log(2 * np.pi) + logdet
        gf = torch.autograd.grad(-ll.sum(), zr)[0]
        with torch.no_grad():
            e = torch.from_numpy(eps_np[0].reshape(B, nz))
            zinv, negobj = netF(e.clone(), objective=torch.zeros(B), reverse=True, return_obj=True)
        o_ll, o_z1, o_ld, o_gf = refpath.prior_grad(z2d, fp, 5, c["coupling"])
        o_zinv, o_obj = refpath.flow_reverse(fp, e.clone(), torch.zeros(B), 5, c["coupling"])
        print(name, "flow fwd z1", close(o_z1, z1.detach(), 1e-6, "z1"), "logdet", close(o_ld, logdet.detach(), 1e-6, "ld"),
              "ll", close(o_ll, ll.detach(), 1e-6, "ll"), "grad_f", close(o_gf, gf, 1e-5, "gf"),
              "inv", close(o_zinv, zinv, 1e-5, "zinv"), "negobj", close(-o_obj, negobj, 1e-5, "obj"))
        if c["coupling"] == 1:
            ga = refpath.prior_grad_analytic(z2d.numpy(), fsd, 5, 1)
            zr64 = z2d.double().clone()
            _, _, _, gf64 = refpath.prior_grad(zr64, to_torch(fsd, torch.float64), 5, 1)
            print(name, "analytic grad vs autograd fp64", close(ga, gf64, 1e-10, "analytic"))
        out.update(z0=z0_np, eps=eps_np, flow_z1=z1.detach().numpy(), flow_logdet=logdet.detach().numpy(),
                   flow_ll=ll.detach().numpy(), flow_grad=gf.numpy(), flow_inv_z=zinv.numpy(),
                   flow_inv_negobj=negobj.numpy())

        if c["dataset"]:
            gsd = synth.generator_state(c["dataset"], nz, c["ngf"], 3, seed=1)
            netG = ref_model._netG(args)
            assert sorted(netG.state_dict().keys()) == sorted(gsd.keys()), "generator state_dict keys differ"
            netG.load_state_dict(to_torch(gsd))
            gp = to_torch(gsd)
            layers = refpath.generator_layers(c["dataset"], nz, c["ngf"], 3)
            x = torch.from_numpy(x_np)
            eps = torch.from_numpy(eps_np)
            zr = z0.clone().requires_grad_(True)
            x_hat = netG(zr)
            loss = 1.0 / (2.0 * c["sigma"] ** 2) * torch.nn.functional.mse_loss(x_hat, x, reduction="sum")
            gg = torch.autograd.grad(loss, zr)[0]
            o_xhat, o_gg = refpath.recon_grad(z0, x, gp, layers, c["sigma"])
            zT, gn, fn = reference_langevin(netG, netF, z0, x, T, 0.1, c["sigma"], eps)
            o_zT, o_gn, o_fn = refpath.langevin(z0, x, gp, fp, layers, depth=5, steps=T, step_size=0.1,
                                                sigma=c["sigma"], eps=eps, coupling=c["coupling"])
            zT_nn, _, _ = reference_langevin(netG, netF, z0, x, T, 0.1, c["sigma"], None)
            # fp64 truth for the error budget
            z64, _, _ = refpath.langevin(z0.double(), x.double(), to_torch(gsd, torch.float64),
                                         to_torch(fsd, torch.float64), layers, depth=5, steps=T, step_size=0.1,
                                         sigma=c["sigma"], eps=eps.double(), coupling=c["coupling"])
            print(name, "x_hat", close(o_xhat, x_hat.detach(), 1e-6, "xhat"), "grad_g", close(o_gg, gg, 1e-5, "gg"),
                  "z_T", close(o_zT, zT, 1e-6, "zT"), "|gg|", close(o_gn, gn, 1e-5, "gn"), "|gf|", close(o_fn, fn, 1e-5, "fn"),
                  "ref fp32 vs fp64 z_T", np.abs(zT.numpy() - z64.numpy()).max() / np.abs(z64.numpy()).max())
            out.update(g_checksum=np.array(synth.checksum(gsd)), x=x_np, x_hat=x_hat.detach().numpy(),
                       grad_g=gg.numpy(), z_T=zT.numpy(), z_T_nonoise=zT_nn.numpy(), z_T_fp64=z64.numpy(),
                       gnorm_g=np.array(gn.item()), gnorm_f=np.array(fn.item()))
        out["config"] = np.array(repr(c))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
        print("wrote", name)

    # state_dict key inventory: 85 flow keys at depth 5, generator keys per arch
    keys = {"flow_nz100_w64": sorted(ref_model._netF(ref_args(CASES["svhn_small"]), nz=100).state_dict().keys())}
    for ds in ("svhn", "cifar10", "celeba_crop", "celeba_hq256"):
        a = ref_args(dict(dataset=ds, nz=100, ngf=8, f_width=64, coupling=1))
        m = ref_model._netG(a)
        keys["gen_" + ds] = [f"{k}:{tuple(v.shape)}" for k, v in m.state_dict().items()]
    np.savez_compressed(os.path.join(out_dir, "state_dict_keys.npz"), **{k: np.array(v) for k, v in keys.items()})
    print("wrote state_dict_keys")
    training_iteration_case(ref_model, out_dir)


if __name__ == "__main__":
    main()
