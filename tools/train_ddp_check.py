"""torchrun --nproc-per-node 2 tools/train_ddp_check.py: one data-parallel training iteration (Langevin, weight-gradient
and flow-gradient kernels per shard, NCCL all-reduce of the two flat gradient buffers, fused Adam) must reproduce the
single-process iteration on the full batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import lsnf_b200
from lsnf_b200 import synth
from lsnf_b200.dist import shard_range
from helpers import to_torch

def build(dev):
    args = lsnf_b200.make_args(dataset="svhn", nz=100, ngf=32, g_l_steps=5)
    netG = lsnf_b200._netG(args).to(dev); netF = lsnf_b200._netF(args, nz=100).to(dev)
    netG.load_state_dict(to_torch(synth.generator_state("svhn", 100, 32))); netF.load_state_dict(to_torch(synth.flow_state(100)))
    return args, netG, netF

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B = 64
    x_np, z0_np, _ = synth.inputs(B, 100, 3, 32, 1, seed=5)
    x, z0 = torch.from_numpy(x_np).to(dev), torch.from_numpy(z0_np).to(dev)
    args, netG, netF = build(dev)
    optG, optF = lsnf_b200.make_optimizers(netG, netF, args)
    a, b = shard_range(B, rank, world)
    lg, lf, gn, fn, zk = lsnf_b200.training_iteration(x[a:b], netG, netF, optG, optF, args, global_batch=B,
                                                      sample_offset=a, seed=3, z0=z0[a:b])
    # loss_g / loss_f come back already summed over ranks (the flat gradient buffers and the two loss scalars are
    # all-reduced inside generator_update / flow_update)
    if rank == 0:
        args1, g1, f1 = build(dev)
        o1, o2 = lsnf_b200.make_optimizers(g1, f1, args1)
        lg1, lf1, _, _, zk1 = lsnf_b200.training_iteration(x, g1, f1, o1, o2, args1, seed=3, z0=z0, data_parallel=False)
        dz = (zk1[a:b] - zk).norm() / zk1[a:b].norm()
        # Adam's first step is lr * g / (|g| + eps): where g ~ 0 a last-bit difference of the summation order (two
        # half-batch partial sums vs one) flips the update, so parameters are compared against the size of one update
        lr = max(getattr(args, "g_lr", 0.0004), getattr(args, "f_lr", 0.0004))
        ps = list(netG.parameters()) + list(netF.parameters())
        qs = list(g1.parameters()) + list(f1.parameters())
        d = torch.cat([(p - q).abs().flatten() for p, q in zip(ps, qs)])
        frac = (d > 2e-5).float().mean().item()
        print(f"DDP check world={world}: loss_g {lg.item():.6f} vs {lg1.item():.6f}; loss_f {lf.item():.6f} vs {lf1.item():.6f}; "
              f"z_k shard rel diff {dz.item():.2e}; max |param diff| {d.max().item():.2e} (lr {lr:.1e}); "
              f"fraction of parameters differing by more than 2e-5: {frac:.2e}")
        assert abs(lg.item() - lg1.item()) < 1e-4 * abs(lg1.item()) and abs(lf.item() - lf1.item()) < 1e-4 * abs(lf1.item())
        assert dz.item() < 1e-6 and d.max().item() <= 2.5 * lr and frac < 2e-3
        print("DDP check ok")
    dist.barrier(); dist.destroy_process_group()

main()
