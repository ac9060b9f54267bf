"""torchrun --nproc-per-node 2 tools/train_ddp_check.py: one data-parallel training iteration (Langevin on the CUDA path
per shard, NCCL all-reduce of the parameter gradients) must reproduce the single-process iteration on the full batch."""
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
    # the parameter updates run through torch autograd (cuDNN); TF32 convolutions, PyTorch's default, make the
    # weight gradients depend on the algorithm cuDNN picks for a batch size, which is not what this check is about
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
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
    tot = torch.stack([lg, lf]); dist.all_reduce(tot)
    if rank == 0:
        args1, g1, f1 = build(dev)
        o1, o2 = lsnf_b200.make_optimizers(g1, f1, args1)
        lg1, lf1, _, _, zk1 = lsnf_b200.training_iteration(x, g1, f1, o1, o2, args1, seed=3, z0=z0, data_parallel=False)
        dz = (zk1[a:b] - zk).norm() / zk1[a:b].norm()
        # Adam's first step is lr * g / (|g| + eps): where g ~ 0 a last-bit difference of the summation order flips the
        # update, so the parameters are compared through the all-reduced GRADIENTS (the optimizers keep them after
        # step()), and the parameters themselves only against the size of one update
        def rel(p, q):
            return ((p - q).norm() / (q.norm() + 1e-12)).item()
        dg = max(rel(p.grad, q.grad) for p, q in zip(netG.parameters(), g1.parameters()) if q.grad is not None)
        df = max(rel(p.grad, q.grad) for p, q in zip(netF.parameters(), f1.parameters()) if q.grad is not None)
        lr = max(getattr(args, "g_lr", 0.0004), getattr(args, "f_lr", 0.0004))
        dp = max((p - q).abs().max().item() for p, q in zip(list(netG.parameters()) + list(netF.parameters()),
                                                             list(g1.parameters()) + list(f1.parameters())))
        print(f"DDP check world={world}: loss_g {tot[0].item():.6f} vs {lg1.item():.6f}; loss_f {tot[1].item():.6f} vs {lf1.item():.6f}; "
              f"z_k shard rel diff {dz.item():.2e}; max grad rel diff G {dg:.2e} F {df:.2e}; max |param diff| {dp:.2e} (lr {lr:.1e})")
        assert abs(tot[0].item() - lg1.item()) < 1e-3 * abs(lg1.item()) and dg < 1e-3 and df < 1e-3 and dp <= 2.5 * lr
        print("DDP check ok")
    dist.barrier(); dist.destroy_process_group()

main()
