"""N>1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding and the parameter-gradient
all-reduce that training mode adds around the (collective-free) Langevin loop."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lsnf_b200
from lsnf_b200 import dist as ldist
from lsnf_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_every_batch():
    for n in (1, 7, 100, 101, 50000, 73257):
        for world in (1, 2, 4, 8):
            spans = [ldist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert ldist.shard_range(50000, 3, 8) == (18750, 25000)   # config 4: 6 250 latents per GPU


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    args = lsnf_b200.make_args(dataset="svhn", nz=100, ngf=4)     # tiny widths: eager autograd branch only
    netG = lsnf_b200._netG(args).train()
    netF = lsnf_b200._netF(args, nz=100).train()
    netG.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state("svhn", 100, 4).items()})
    netF.load_state_dict({k: torch.from_numpy(v) for k, v in synth.flow_state(100).items()})
    B = 10
    x_np, z_np, _ = synth.inputs(B, 100, 3, 32, 1, seed=3)
    x, z = torch.from_numpy(x_np), torch.from_numpy(z_np)
    a, b = ldist.shard_range(B, rank, world)
    # generator update of train.py:390-394 on the local shard, then the all-reduce
    loss_g = torch.nn.functional.mse_loss(netG(z[a:b]), x[a:b], reduction="sum") / (b - a)
    loss_g.backward()
    nbytes = ldist.allreduce_grads(netG.parameters(), scale=(b - a) / B)
    # flow update of train.py:403-411
    z1, logdet, _ = netF(z[a:b].reshape(b - a, 100), objective=torch.zeros(b - a))
    ll = (-0.5 * z1 ** 2).flatten(1).sum(-1) + np.log(2 * np.pi) + logdet
    (-ll.mean()).backward()
    ldist.allreduce_grads(netF.parameters(), scale=(b - a) / B)
    if rank == 0:
        torch.save({"g": [p.grad.clone() for p in netG.parameters()],
                    "f": [p.grad.clone() for p in netF.parameters() if p.grad is not None], "nbytes": nbytes}, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_matches_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    tmp = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
    got = torch.load(tmp)
    args = lsnf_b200.make_args(dataset="svhn", nz=100, ngf=4)
    netG = lsnf_b200._netG(args).train()
    netF = lsnf_b200._netF(args, nz=100).train()
    netG.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state("svhn", 100, 4).items()})
    netF.load_state_dict({k: torch.from_numpy(v) for k, v in synth.flow_state(100).items()})
    x_np, z_np, _ = synth.inputs(10, 100, 3, 32, 1, seed=3)
    x, z = torch.from_numpy(x_np), torch.from_numpy(z_np)
    (torch.nn.functional.mse_loss(netG(z), x, reduction="sum") / 10).backward()
    z1, logdet, _ = netF(z.reshape(10, 100), objective=torch.zeros(10))
    (-((-0.5 * z1 ** 2).flatten(1).sum(-1) + np.log(2 * np.pi) + logdet).mean()).backward()
    for p, g in zip(netG.parameters(), got["g"]):
        assert torch.allclose(p.grad, g, rtol=1e-4, atol=1e-6)
    for p, g in zip([p for p in netF.parameters() if p.grad is not None], got["f"]):
        assert torch.allclose(p.grad, g, rtol=1e-4, atol=1e-6)
    assert got["nbytes"] == sum(p.numel() for p in netG.parameters()) * 4


def test_reference_arm_non_zero_ranks_exit_silently():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_eager_training_branch_matches_oracle_on_cpu():
    # the autograd branch used for the (out-of-scope) parameter updates computes the same function as the oracle
    from oracle import refpath
    args = lsnf_b200.make_args(dataset="svhn", nz=100, ngf=4)
    netF = lsnf_b200._netF(args, nz=100).train()
    sd = synth.flow_state(100)
    netF.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    z = torch.randn(5, 100, generator=torch.Generator().manual_seed(0))
    a, b, _ = netF(z, objective=torch.zeros(5))
    ra, rb = refpath.flow_forward({k: torch.from_numpy(v) for k, v in sd.items()}, z, torch.zeros(5), 5)
    assert torch.allclose(a, ra, atol=1e-6) and torch.allclose(b, rb, atol=1e-5)
    zi = netF(a.detach(), objective=torch.zeros(5), reverse=True)
    assert torch.allclose(zi, z, atol=1e-4)
    netF.eval()
    with pytest.raises(RuntimeError):
        netF(z, objective=torch.zeros(5))          # inference on CPU tensors: no fallback
