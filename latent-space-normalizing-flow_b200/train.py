"""One training iteration of the reference (train.py:376-415) with the Langevin posterior inference on the CUDA
path and data parallelism over the latent batch.

Only the Langevin call is in the hot-path scope; the generator and flow parameter updates stay
in torch autograd exactly as upstream (section 8f lists their kernels as next).  What this module adds is the
glue a multi-GPU run needs: each rank infers the latents of its shard (no collective), computes its local losses,
and the parameter gradients are summed across ranks with one flat all-reduce per network (NCCL over NVLink).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

from .dist import allreduce_grads
from .langevin import sample_langevin_post_z_with_flow


def make_optimizers(netG, netF, args):
    """train.py:294-295."""
    g = lambda k, d: args.get(k, d) if isinstance(args, dict) else getattr(args, k, d)
    optG = torch.optim.Adam(netG.parameters(), lr=g("g_lr", 0.0004), weight_decay=g("g_decay", 0),
                            betas=(g("g_beta1", 0.5), g("g_beta2", 0.999)))
    optF = torch.optim.Adam(netF.parameters(), lr=g("f_lr", 0.0004), weight_decay=g("f_decay", 0),
                            betas=(g("f_beta1", 0.5), g("f_beta2", 0.999)))
    return optG, optF


_iter_counter = [0]


def training_iteration(x, netG, netF, optG, optF, args, *, global_batch=None, sample_offset=None, seed=None,
                       z0=None, group=None, data_parallel=None):
    """x: this rank's shard [B_local, nc, H, W] on the GPU.  Returns (loss_g, loss_f, |grad_g|, |grad_f|, z_k).

    Mirrors train.py:378-415: z_0 ~ N(0, I); z_k = Langevin(z_0, x); generator step on mse_sum(G(z_k), x) / B;
    flow step on -mean log p(z_k) (with the f_is_grad_clamp / f_max_norm clipping of train.py:411-412).  With several
    ranks the losses are normalised by the GLOBAL batch, so the summed gradients equal the single-process ones.

    Data parallelism: every rank must draw DIFFERENT chains.  ``sample_offset`` (index of this shard's first sample in
    the global batch) defaults to the contiguous sharding of ``dist.shard_range``; the Philox noise is keyed by
    (seed, global sample index, step) and z_0 is the rank's slice of ONE global draw from a generator seeded with
    ``seed``, so an N-rank iteration reproduces the single-process one.  ``seed`` defaults to
    (args.seed << 32) ^ iteration counter -- identical on every rank, as it must be."""
    b_local = x.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if data_parallel is False:
        world = 1   # e.g. a single-process reference run inside an initialised process group
    g = lambda k, d: args.get(k, d) if isinstance(args, dict) else getattr(args, k, d)
    if seed is None:
        seed = (int(g("seed", 1)) << 32) ^ _iter_counter[0]
    _iter_counter[0] += 1
    if world > 1:
        rank = dist.get_rank(group)
        if global_batch is None:
            sizes = torch.tensor([b_local], device=x.device)
            dist.all_reduce(sizes, group=group)
            global_batch = int(sizes.item())
        if sample_offset is None:
            from .dist import shard_range
            a, b = shard_range(int(global_batch), rank, world)
            if b - a != b_local:
                raise ValueError(f"rank {rank} holds {b_local} samples but the contiguous sharding of a global batch of "
                                 f"{global_batch} gives it {b - a}; pass sample_offset explicitly")
            sample_offset = a
    b_global = b_local if global_batch is None else int(global_batch)
    sample_offset = 0 if sample_offset is None else int(sample_offset)
    netG.train()
    netF.train()
    if z0 is None:                                                                       # train.py:384
        gen = torch.Generator(x.device).manual_seed(int(seed) & (2 ** 63 - 1))
        z0 = torch.randn(b_global, netG.nz, 1, 1, device=x.device, generator=gen)[sample_offset:sample_offset + b_local]
    z_k, gn, fn = sample_langevin_post_z_with_flow(z0, x, netG, netF, args, seed=seed,
                                                   sample_offset=sample_offset)          # train.py:387
    # generator update (train.py:390-398); autograd branch of _netG.forward
    optG.zero_grad()
    x_hat = netG(z_k.detach())
    loss_g = torch.nn.functional.mse_loss(x_hat, x, reduction="sum") / b_global
    loss_g.backward()
    if world > 1:
        allreduce_grads(netG.parameters(), group=group)
    if g("g_is_grad_clamp", False):   # train.py:396-397 (upstream names an undefined `opt` there; intended semantics)
        torch.nn.utils.clip_grad_norm_(netG.parameters(), g("g_max_norm", 100))
    optG.step()
    # flow update (train.py:403-415)
    optF.zero_grad()
    z1, logdet, _ = netF(torch.squeeze(z_k).reshape(b_local, -1), objective=torch.zeros(b_local, device=x.device))
    ll = (-0.5 * z1 ** 2).flatten(1).sum(-1) + np.log(2 * np.pi) + logdet
    loss_f = -ll.sum() / b_global
    loss_f.backward()
    if world > 1:
        allreduce_grads(netF.parameters(), group=group)
    if g("f_is_grad_clamp", False):   # train.py:411-412
        torch.nn.utils.clip_grad_norm_(netF.parameters(), g("f_max_norm", 100))
    optF.step()
    return loss_g.detach(), loss_f.detach(), gn, fn, z_k
