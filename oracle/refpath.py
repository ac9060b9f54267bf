"""Restatement of the reference hot path on CPU (torch, fp32 or fp64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function names the
reference lines it follows (paths relative to ``/root/reference``).  Parameters
are passed as plain ``{state_dict key: tensor}`` mappings using the reference's
own key names, so the same function consumes a reference checkpoint, the
reference modules' ``state_dict()`` or the product modules' ``state_dict()``.

The only semantic change w.r.t. the reference closure is that the Langevin
noise is an argument (``eps[t]``) instead of ``torch.randn_like``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

LOG_2PI = float(np.log(2 * np.pi))  # the literal additive constant of train.py:318


# --------------------------------------------------------------------------
# generator (model.py:48-157)
# --------------------------------------------------------------------------
def generator_layers(dataset: str, nz: int, ngf: int, nc: int = 3) -> List[Tuple[int, int, int, int, int]]:
    """(C_in, C_out, kernel, stride, pad) per ConvTranspose2d of ``_netG``.

    svhn model.py:56-71, cifar10 :77-92, celeba_crop :98-117, celeba_hq256 :123-151.
    """
    if dataset == "svhn":
        return [(nz, ngf * 8, 4, 1, 0), (ngf * 8, ngf * 4, 4, 2, 1), (ngf * 4, ngf * 2, 4, 2, 1),
                (ngf * 2, nc, 4, 2, 1)]
    if dataset == "cifar10":
        return [(nz, ngf * 8, 8, 1, 0), (ngf * 8, ngf * 4, 4, 2, 1), (ngf * 4, ngf * 2, 4, 2, 1),
                (ngf * 2, nc, 3, 1, 1)]
    if dataset == "celeba_crop":
        return [(nz, ngf * 8, 4, 1, 0), (ngf * 8, ngf * 4, 4, 2, 1), (ngf * 4, ngf * 2, 4, 2, 1),
                (ngf * 2, ngf, 4, 2, 1), (ngf, nc, 4, 2, 1)]
    if dataset == "celeba_hq256":
        return [(nz, ngf * 16, 4, 1, 0), (ngf * 16, ngf * 8, 4, 2, 1), (ngf * 8, ngf * 4, 4, 2, 1),
                (ngf * 4, ngf * 2, 4, 2, 1), (ngf * 2, ngf, 4, 2, 1), (ngf, ngf, 4, 2, 1),
                (ngf, nc, 4, 2, 1)]
    raise ValueError(dataset)  # model.py:154


def generator_forward(gp: Params, z: torch.Tensor, layers: Sequence[Tuple[int, int, int, int, int]],
                      leak: float = 0.2) -> torch.Tensor:
    """``_netG.forward`` (model.py:156-157): [ConvT, Identity, LeakyReLU]*(L-1) + [ConvT, Tanh].

    Parameter keys are those of the ``nn.Sequential``: ``gen.{3*i}.weight|bias``.
    """
    h = z
    last = len(layers) - 1
    for i, (_ci, _co, _k, s, p) in enumerate(layers):
        h = F.conv_transpose2d(h, gp[f"gen.{3 * i}.weight"], gp[f"gen.{3 * i}.bias"], stride=s, padding=p)
        h = torch.tanh(h) if i == last else F.leaky_relu(h, leak)
    return h


# --------------------------------------------------------------------------
# flow prior (model.py:171-498)
# --------------------------------------------------------------------------
def _step_prefix(i: int) -> str:
    return f"revnet2d_s.0.revnet2d_step_s.{i}."


def _actnorm_fwd(x, b, logs):
    """actnorm forward without logdet (model.py:243-244, :264-268): (x + b) * exp(3*logs)."""
    return (x + b) * torch.exp(logs * 3.0)


def coupling_mlp(fp: Params, pre: str, h: torch.Tensor) -> torch.Tensor:
    """``f.forward`` (model.py:306-310) = relu(fc_1) -> relu(fc_2) -> fc_zeros.

    fc.forward model.py:321-332 (matmul, then actnorm); fc_zeros.forward :344-350.
    """
    h = F.relu(_actnorm_fwd(h @ fp[pre + "f.fc_1.w"], fp[pre + "f.fc_1.actnorm.b"], fp[pre + "f.fc_1.actnorm.logs"]))
    h = F.relu(_actnorm_fwd(h @ fp[pre + "f.fc_2.w"], fp[pre + "f.fc_2.actnorm.b"], fp[pre + "f.fc_2.actnorm.logs"]))
    h = (h @ fp[pre + "f.fc_zeros.w"] + fp[pre + "f.fc_zeros.b"]) * torch.exp(fp[pre + "f.fc_zeros.logs"] * 3.0)
    return h


def log_abs_det(w: torch.Tensor) -> torch.Tensor:
    """model.py:182 -- determinant in fp64, cast back to the working dtype."""
    return torch.log(torch.abs(torch.det(w.double()))).to(w.dtype)


def flow_forward(fp: Params, z: torch.Tensor, logdet: torch.Tensor, depth: int, coupling: int = 1,
                 permutation: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    """``_netF.forward(reverse=False)`` (model.py:474-483) -> revnet2d (:357-360) ->
    revnet2d_step.forward (:391-422).  Returns (z_out, logdet)."""
    n = z.shape[-1]
    for i in range(depth):
        pre = _step_prefix(i)
        logs = fp[pre + "actnorm.logs"]
        z = _actnorm_fwd(z, fp[pre + "actnorm.b"], logs)                       # model.py:392
        logdet = logdet + torch.sum(logs * 3.0)                                # model.py:273-276
        if permutation == 2:                                                   # model.py:399-400
            w = fp[pre + "invertible_1x1_conv.w"]
            z = z @ w                                                          # model.py:187
            logdet = logdet + log_abs_det(w)                                   # model.py:189
        elif permutation == 1:
            # shuffle_features (model.py:214-225) is broken upstream;
            # intended semantics: a fixed int32 channel permutation, h[:, idx].
            z = z.index_select(1, fp[pre + "shuffle_features.indices"].long())
        else:
            raise Exception()                                                  # model.py:379
        z1, z2 = z[:, : n // 2], z[:, n // 2:]
        if coupling == 0:                                                      # model.py:407-408
            z2 = z2 + coupling_mlp(fp, pre, z1)
        elif coupling == 1:                                                    # model.py:409-418
            h = coupling_mlp(fp, pre, z1)
            shift = h[:, 0::2]
            scale = torch.sigmoid(h[:, 1::2] + 2.0)
            z2 = (z2 + shift) * scale
            logdet = logdet + torch.sum(torch.log(scale), dim=1)
        else:
            raise Exception()                                                  # model.py:420
        z = torch.cat([z1, z2], 1)
    return z, logdet


def flow_reverse(fp: Params, z: torch.Tensor, logdet: torch.Tensor, depth: int, coupling: int = 1,
                 permutation: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    """``_netF.forward(reverse=True)`` (model.py:484-498) -> revnet2d (:361-363) ->
    revnet2d_step.forward reverse branch (:424-456).  Returns (z, objective); the reference
    returns ``-objective`` when ``return_obj`` (model.py:498).  Does NOT mutate its input
    (the reference does, model.py:436-437)."""
    n = z.shape[-1]
    for i in reversed(range(depth)):
        pre = _step_prefix(i)
        z1, z2 = z[:, : n // 2], z[:, n // 2:]
        if coupling == 0:                                                      # model.py:429-430
            z2 = z2 - coupling_mlp(fp, pre, z1)
        elif coupling == 1:                                                    # model.py:431-438
            h = coupling_mlp(fp, pre, z1)
            shift = h[:, 0::2]
            scale = torch.sigmoid(h[:, 1::2] + 2.0)
            z2 = z2 / scale - shift
            logdet = logdet - torch.sum(torch.log(scale), dim=1)
        else:
            raise Exception()
        z = torch.cat([z1, z2], 1)
        if permutation == 2:                                                   # model.py:451-452, :192-198
            w = fp[pre + "invertible_1x1_conv.w"]
            z = z @ torch.inverse(w)
            logdet = logdet - log_abs_det(w)
        elif permutation == 1:
            z = z.index_select(1, fp[pre + "shuffle_features.indices_inverse"].long())
        else:
            raise Exception()
        logs = fp[pre + "actnorm.logs"]                                        # model.py:456 -> :288-291
        z = z * torch.exp(-logs * 3.0) - fp[pre + "actnorm.b"]
        logdet = logdet - torch.sum(logs * 3.0)
    return z, logdet


def log_prior(fp: Params, z: torch.Tensor, depth: int, coupling: int = 1, permutation: int = 2):
    """train.py:316-319: ll_b = sum_j(-0.5 z1^2) + log(2 pi) + logdet_b.  Returns (ll, z1, logdet)."""
    z1, logdet = flow_forward(fp, z, torch.zeros(z.shape[0], dtype=z.dtype, device=z.device), depth, coupling,
                              permutation)   # train.py:316: torch.zeros(B).to(device)
    ll = (-0.5 * z1 ** 2).flatten(1).sum(-1) + LOG_2PI + logdet
    return ll, z1, logdet


# --------------------------------------------------------------------------
# Langevin closure (train.py:307-335; test variant :602-634)
# --------------------------------------------------------------------------
def langevin(z0: torch.Tensor, x: torch.Tensor, gp: Params, fp: Params, layers, *, depth: int, steps: int,
             step_size: float, sigma: float, eps: Optional[torch.Tensor], leak: float = 0.2, coupling: int = 1,
             permutation: int = 2, trace: Optional[list] = None):
    """``sample_langevin_post_z_with_flow`` with injected noise.

    ``z0`` [B,nz,1,1], ``x`` [B,nc,H,W], ``eps`` [steps,B,nz,1,1] or None (the test-mode variant,
    train.py:623-625, has no noise; its step count ``g_l_steps*20`` is the caller's business).
    Returns (z [B,nz,1,1], mean_b |grad_g|_2, mean_b |grad_f|_2) of the LAST step (train.py:335).
    The diagnostics use the real batch size (the reference's ``view(args.batch_size, -1)``,
    train.py:328-329, mis-shapes on a ragged last batch).
    """
    z = z0.clone().detach()
    z.requires_grad_(True)
    bsz = z.shape[0]
    gn = fn = None
    for t in range(steps):
        x_hat = generator_forward(gp, z, layers, leak)                                   # train.py:312
        g_log_lkhd = 1.0 / (2.0 * sigma * sigma) * F.mse_loss(x_hat, x, reduction="sum")  # :313
        z_grad_g = torch.autograd.grad(g_log_lkhd, z)[0]                                 # :314
        ll, _z1, _ld = log_prior(fp, z.reshape(bsz, -1), depth, coupling, permutation)   # :316-319
        f_log_lkhd = -ll.sum()                                                           # :320
        z_grad_f = torch.autograd.grad(f_log_lkhd, z)[0]                                 # :323
        z.data = z.data - 0.5 * step_size * step_size * (z_grad_g + z_grad_f)            # :324
        if eps is not None:
            z.data += step_size * eps[t]                                                 # :325-326
        gn = z_grad_g.view(bsz, -1).norm(dim=1).mean()                                   # :328
        fn = z_grad_f.view(bsz, -1).norm(dim=1).mean()                                   # :329
        if trace is not None:
            trace.append(z.detach().clone())
    return z.detach(), gn, fn


def trainable_flow_keys(fp: Params) -> List[str]:
    """Keys of the flow ``state_dict`` that ``optF`` owns and that receive a gradient: every float tensor except the
    ``actnorm.bias`` aliases of ``actnorm.b`` (model.py:231 registers the same Parameter twice) and the ``fc.b`` the
    forward pass never reads (model.py:329-330; its .grad stays None, Adam skips it)."""
    return [k for k in fp if fp[k].is_floating_point() and not k.endswith("actnorm.bias")
            and not (k.endswith(".b") and (".fc_1." in k or ".fc_2." in k) and "actnorm" not in k)]


def parameter_updates(gp: Params, fp: Params, z_k: torch.Tensor, x: torch.Tensor, layers, optG, optF, *, depth: int,
                      leak: float = 0.2, coupling: int = 1, permutation: int = 2,
                      g_max_norm: Optional[float] = None, f_max_norm: Optional[float] = None):
    """The two parameter updates that follow the Langevin call in one training iteration (train.py:390-415).

    ``gp`` / ``fp`` map state_dict keys to LEAF tensors with ``requires_grad`` that ``optG`` / ``optF``
    (``torch.optim.Adam``, train.py:294-295) were built on.  Returns (loss_g, loss_f) as 0-d tensors.
    """
    bsz = x.shape[0]
    optG.zero_grad()                                                                     # train.py:390
    x_hat = generator_forward(gp, z_k.detach(), layers, leak)                            # :392
    loss_g = F.mse_loss(x_hat, x, reduction="sum") / bsz                                 # :393
    loss_g.backward()                                                                    # :394
    if g_max_norm is not None:                                                           # :396-397 (intended: args.g_max_norm)
        torch.nn.utils.clip_grad_norm_([v for v in gp.values() if v.grad is not None], g_max_norm)
    optG.step()                                                                          # :398
    optF.zero_grad()                                                                     # :403
    ll, _z1, _ld = log_prior(fp, torch.squeeze(z_k.detach()).reshape(bsz, -1), depth, coupling, permutation)  # :405-409
    loss_f = -ll.mean()                                                                  # :410
    loss_f.backward()                                                                    # :411
    if f_max_norm is not None:                                                           # :412-413
        torch.nn.utils.clip_grad_norm_([v for v in fp.values() if v.grad is not None], f_max_norm)
    optF.step()                                                                          # :415
    return loss_g.detach(), loss_f.detach()


def recon_grad(z: torch.Tensor, x: torch.Tensor, gp: Params, layers, sigma: float, leak: float = 0.2):
    """One evaluation of train.py:312-314: (x_hat, d/dz [1/(2 sigma^2) * sum (G(z)-x)^2])."""
    z = z.clone().detach().requires_grad_(True)
    x_hat = generator_forward(gp, z, layers, leak)
    loss = 1.0 / (2.0 * sigma * sigma) * F.mse_loss(x_hat, x, reduction="sum")
    return x_hat.detach(), torch.autograd.grad(loss, z)[0]


def prior_grad(z: torch.Tensor, fp: Params, depth: int, coupling: int = 1, permutation: int = 2):
    """One evaluation of train.py:316-323: (ll, z1, logdet, d/dz [-sum_b ll_b]) for z [B,nz]."""
    z = z.clone().detach().requires_grad_(True)
    ll, z1, logdet = log_prior(fp, z, depth, coupling, permutation)
    g = torch.autograd.grad(-ll.sum(), z)[0]
    return ll.detach(), z1.detach(), logdet.detach(), g


# --------------------------------------------------------------------------
# analytic backward of -sum_b ll_b through the coupling layers
# numpy fp64; used to pin the formulas the CUDA kernel implements against autograd.
# --------------------------------------------------------------------------
def prior_grad_analytic(z: np.ndarray, fp: Dict[str, np.ndarray], depth: int, coupling: int = 1) -> np.ndarray:
    def P(k):
        return np.asarray(fp[k], dtype=np.float64)

    n = z.shape[1]
    x = np.asarray(z, dtype=np.float64)
    saved = []
    for i in range(depth):
        pre = _step_prefix(i)
        x = (x + P(pre + "actnorm.b")) * np.exp(3.0 * P(pre + "actnorm.logs"))
        x = x @ P(pre + "invertible_1x1_conv.w")
        x1, x2 = x[:, : n // 2], x[:, n // 2:]
        e1 = np.exp(3.0 * P(pre + "f.fc_1.actnorm.logs"))
        e2 = np.exp(3.0 * P(pre + "f.fc_2.actnorm.logs"))
        e3 = np.exp(3.0 * P(pre + "f.fc_zeros.logs"))
        a1 = np.maximum((x1 @ P(pre + "f.fc_1.w") + P(pre + "f.fc_1.actnorm.b")) * e1, 0.0)
        a2 = np.maximum((a1 @ P(pre + "f.fc_2.w") + P(pre + "f.fc_2.actnorm.b")) * e2, 0.0)
        h = (a2 @ P(pre + "f.fc_zeros.w") + P(pre + "f.fc_zeros.b")) * e3
        if coupling == 1:
            shift = h[:, 0::2]
            scale = 1.0 / (1.0 + np.exp(-(h[:, 1::2] + 2.0)))
            y2 = (x2 + shift) * scale
        else:
            shift, scale = h, None
            y2 = x2 + shift
        saved.append((a1, a2, x2 + shift if coupling == 1 else None, scale, e1, e2, e3))
        x = np.concatenate([x1, y2], axis=1)
    g = x.copy()          # d/dz_out of sum(0.5 z_out^2)
    g_ld = -1.0           # d/dlogdet of -sum ll
    for i in reversed(range(depth)):
        pre = _step_prefix(i)
        a1, a2, x2s, scale, e1, e2, e3 = saved[i]
        g1, g2 = g[:, : n // 2], g[:, n // 2:]
        if coupling == 1:
            g_x2 = g2 * scale
            g_shift = g2 * scale
            g_scale = g2 * x2s + g_ld / scale
            g_h = np.empty((g.shape[0], n))
            g_h[:, 0::2] = g_shift
            g_h[:, 1::2] = g_scale * scale * (1.0 - scale)
        else:
            g_x2 = g2
            g_h = g2
        g_a2 = ((g_h * e3) @ P(pre + "f.fc_zeros.w").T) * (a2 > 0)
        g_a1 = ((g_a2 * e2) @ P(pre + "f.fc_2.w").T) * (a1 > 0)
        g_x1 = g1 + (g_a1 * e1) @ P(pre + "f.fc_1.w").T
        g = np.concatenate([g_x1, g_x2], axis=1) @ P(pre + "invertible_1x1_conv.w").T
        g = g * np.exp(3.0 * P(pre + "actnorm.logs"))
    return g
