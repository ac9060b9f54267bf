"""The reference's README command lines on synthetic data, through the B200 path.

    python examples/run_synthetic.py --dataset cifar10 --g_l_steps 40 --img_size 32 --nz 128 --ngf 128 \\
        --g_lr 0.00038 --f_lr 0.00038 --n_epochs 2 --iters_per_epoch 10 --ckpt_dir /tmp/lsnf_ckpt
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/run_synthetic.py --dataset cifar10 ...   # data parallel
    python examples/run_synthetic.py --dataset svhn --test_mode --g_l_steps 20 --nz 100 --ngf 64 \\
        --path_check_point /tmp/lsnf_ckpt/ckpt_000001.pth --n_fid_samples 1000 --testing_reconstruct

Every flag of the reference's ``train.py`` (train.py:37-99) is accepted with its default; ``--iters_per_epoch``,
``--ckpt_dir`` and ``--n_test_batches`` are the only additions (the reference takes these from its datasets and its
output directory).  What runs:

* training (train.py:362-505): per iteration ``lsnf_b200.training_iteration`` = Langevin posterior inference +
  generator update + flow update on the CUDA kernels, data-parallel when launched under torchrun; per epoch the two
  ``ExponentialLR`` schedules (train.py:297-298, :484-485) and a checkpoint with the reference's keys
  ``epoch / netG / optG / netF / optF`` (train.py:493-503), which the reference's own ``train.py`` can resume from;
* ``--test_mode`` (train.py:520-662): load ``--path_check_point``, draw ``--n_fid_samples`` prior samples
  eps -> F^-1 -> G -> [0,1] (train.py:565-586), and with ``--testing_reconstruct`` report the reconstruction error of
  ``g_l_steps * 20`` noise-free Langevin iterations per batch (train.py:606, :641-662).

Status: written in the last session of round 2 after the GPU budget was spent -- the host side (flags, network
construction, checkpoints, resume, the epoch / test-mode loops with the device entry points stubbed, the no-GPU error)
is covered by tests/test_example_cli.py on the CPU; the GPU legs have NOT been executed on a B200.  They only call entry points that the GPU suite exercises with the same arguments
(``training_iteration``: tests/test_gpu_langevin.py, tools/train_ddp_check.py; ``sample_x`` / ``reconstruction_error``:
tests/test_gpu_langevin.py, tests/test_gpu_sampling_and_long_chains.py).

Datasets, FID and image dumps are outside the path this repository rebuilds: images are
x ~ U(-1, 1), as in bench.py.  There is no CPU fallback: without a CUDA device the script stops with an error.
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import lsnf_b200  # noqa: E402
from lsnf_b200 import cli, synth  # noqa: E402

EXTRA = {"iters_per_epoch": (int, 20), "ckpt_dir": (str, None), "n_test_batches": (int, 4)}


def parse(argv=None):
    p = cli.build_parser()
    for name, (kind, default) in EXTRA.items():
        p.add_argument("--" + name, type=kind, default=default)
    a = lsnf_b200.AttrDict(vars(p.parse_args(argv)))
    if a.img_size != synth.image_size(a.dataset):
        raise SystemExit(f"--img_size {a.img_size} does not match the {a.dataset} generator "
                         f"({synth.image_size(a.dataset)}x{synth.image_size(a.dataset)}, model.py:52-151)")
    return a


def set_seed(seed: int) -> None:
    """train.py:723-730."""
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def build(args, device):
    """train.py:268-272 / :540-551: the two networks, Xavier-initialised generator."""
    netG = lsnf_b200._netG(args)
    netF = lsnf_b200._netF(args, nz=args.nz)
    netG.apply(lsnf_b200.weights_init_xavier)
    netF.apply(lsnf_b200.weights_init_xavier)
    return netG.to(device), netF.to(device)


def checkpoint_dict(epoch, netG, netF, optG, optF):
    """train.py:495-501."""
    return {"epoch": epoch, "netF": netF.state_dict(), "optF": optF.state_dict(), "netG": netG.state_dict(),
            "optG": optG.state_dict()}


def load_checkpoint(path, netG, netF, optG=None, optF=None, map_location=None) -> int:
    """train.py:342-349 (resume) / :546-548 (test).  Returns the epoch to continue from."""
    ckp = torch.load(path, map_location=map_location)
    netG.load_state_dict(ckp["netG"])
    netF.load_state_dict(ckp["netF"])
    if optG is not None:
        optG.load_state_dict(ckp["optG"])
    if optF is not None:
        optF.load_state_dict(ckp["optF"])
    lsnf_b200.invalidate_plans()   # parameters were written through .data: re-pack at the next call
    return int(ckp["epoch"]) + 1


def synthetic_batch(args, batch, device, generator):
    """x ~ U(-1, 1) [B, nc, H, W] (images are normalised to [-1, 1], train.py:138)."""
    return torch.rand(batch, args.nc, args.img_size, args.img_size, device=device, generator=generator) * 2.0 - 1.0


def cuda_device(local_rank: int) -> torch.device:
    """This rank's GPU, made current.  No GPU -> error: the lsnf_b200 path has no CPU fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("run_synthetic.py needs a CUDA device: the lsnf_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    return torch.device("cuda", local_rank)


def device_sync(device) -> None:
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def train(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = cuda_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    set_seed(args.seed)                      # every rank builds the same parameters
    netG, netF = build(args, device)
    optG, optF = lsnf_b200.make_optimizers(netG, netF, args)                      # train.py:294-295
    schedG = torch.optim.lr_scheduler.ExponentialLR(optG, args.g_gamma)           # train.py:297-298
    schedF = torch.optim.lr_scheduler.ExponentialLR(optF, args.f_gamma)
    epoch_start = 0
    if args.path_check_point:
        epoch_start = load_checkpoint(args.path_check_point, netG, netF, optG, optF, map_location=device)
    if args.batch_size % world:
        raise SystemExit("--batch_size must be divisible by the number of ranks")
    b_local = args.batch_size // world
    data_rng = torch.Generator(device).manual_seed(args.seed * 1000 + rank)       # each rank its own shard of the data
    it = 0
    for epoch in range(epoch_start, args.n_epochs):
        t0 = time.perf_counter()
        for i in range(args.iters_per_epoch):
            x = synthetic_batch(args, b_local, device, data_rng)
            loss_g, loss_f, gn, fn, _z = lsnf_b200.training_iteration(
                x, netG, netF, optG, optF, args, global_batch=args.batch_size, sample_offset=rank * b_local,
                seed=(args.seed << 32) ^ it)
            it += 1
            if i % args.n_printout == 0 and rank == 0:   # train.py:421-461 (its log line; the .item() calls synchronise)
                print("{:5d}/{:5d} {:5d}/{:5d} loss_g={:8.3f}, loss_f={:8.3f}, z_g_grad_norm={:8.3f}, "
                      "z_f_grad_norm={:8.3f}, lr_g={:8.6f}, lr_f={:8.6f}".format(
                          epoch, args.n_epochs, i, args.iters_per_epoch, loss_g.item(), loss_f.item(), gn.item(),
                          fn.item(), optG.param_groups[0]["lr"], optF.param_groups[0]["lr"]), flush=True)
        schedG.step()                                                              # train.py:484-485
        schedF.step()
        device_sync(device)
        if rank == 0:
            dt = time.perf_counter() - t0
            print(f"epoch {epoch}: {args.iters_per_epoch} iterations in {dt:.2f} s = "
                  f"{args.batch_size * args.g_l_steps * args.iters_per_epoch / dt:,.0f} latent-steps/s "
                  f"({world} GPU(s), batch {args.batch_size})", flush=True)
            if args.ckpt_dir and (epoch == args.n_epochs - 1 or epoch % args.n_ckpt == 0):   # train.py:493-503
                os.makedirs(args.ckpt_dir, exist_ok=True)
                path = os.path.join(args.ckpt_dir, "ckpt_{:>06d}.pth".format(epoch))
                torch.save(checkpoint_dict(epoch, netG, netF, optG, optF), path)
                print("wrote", path, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test(args):
    device = cuda_device(int(os.environ.get("LOCAL_RANK", "0")))
    set_seed(args.seed)
    netG, netF = build(args, device)
    if args.path_check_point:
        load_checkpoint(args.path_check_point, netG, netF, map_location=device)
    netG.eval()
    netF.eval()
    # train.py:565-586: n_fid_samples prior samples in batches of batch_size, already mapped to [0, 1]
    n_batches = max(1, args.n_fid_samples // args.batch_size)
    rng = torch.Generator(device).manual_seed(args.seed)
    device_sync(device)
    t0 = time.perf_counter()
    stats = torch.zeros(2, device=device)
    for _ in range(n_batches):
        xs = lsnf_b200.sample_x(netG, netF, args.batch_size, device, generator=rng)
        stats += torch.stack([xs.mean(), xs.var()])
    device_sync(device)
    dt = time.perf_counter() - t0
    m, v = (stats / n_batches).tolist()
    print(f"{n_batches * args.batch_size} prior samples in {dt:.3f} s = {n_batches * args.batch_size / dt:,.0f} samples/s; "
          f"pixel mean {m:.4f}, variance {v:.4f} (FID is outside this repository's scope)", flush=True)
    if args.testing_reconstruct:                                                   # train.py:641-662
        batches = (synthetic_batch(args, args.batch_size, device, rng) for _ in range(args.n_test_batches))
        t0 = time.perf_counter()
        err = lsnf_b200.reconstruction_error(batches, netG, netF, args, generator=rng)
        dt = time.perf_counter() - t0
        print(f"reconstruction error={err} ({args.n_test_batches} batches of {args.batch_size}, "
              f"{args.g_l_steps * 20} noise-free Langevin iterations each, {dt:.2f} s)", flush=True)


def main(argv=None):
    args = parse(argv)
    (test if args.test_mode else train)(args)


if __name__ == "__main__":
    main()
