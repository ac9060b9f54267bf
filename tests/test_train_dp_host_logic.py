"""Data-parallel bookkeeping of ``training_iteration`` (train.py of the package) on the CPU, world_size 2 over gloo, with
the three device entry points replaced by recorders: every rank must infer DIFFERENT chains -- its shard's
offset into the global batch, the global batch size, one seed shared by all ranks, and z_0 = the rank's slice of ONE
global draw, so that an N-rank iteration draws what the single-process one draws."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lsnf_b200
from lsnf_b200 import train as ltrain


class _Pending:
    def __init__(self, loss):
        self.loss = loss

    def finish(self):
        return self.loss


def _install_recorders(monkeypatch_setattr, log):
    def fake_langevin(z, x, netG, netF, args, verbose=False, **kw):
        log["langevin"] = dict(z0=z.clone(), seed=kw.get("seed"), sample_offset=kw.get("sample_offset"), train=kw.get("train"))
        return z.clone(), torch.tensor(1.0), torch.tensor(2.0)

    def fake_gen_begin(netG, optG, z_k, x, args, *, global_batch=None, group=None, world=1, plan=None):
        log["gen"] = dict(global_batch=global_batch, world=world)
        return _Pending(torch.tensor(3.0))

    def fake_flow_update(netF, optF, z_k, args, *, global_batch=None, group=None, world=1, plan=None):
        log["flow"] = dict(global_batch=global_batch, world=world)
        return torch.tensor(4.0)

    monkeypatch_setattr(ltrain, "sample_langevin_post_z_with_flow", fake_langevin)
    monkeypatch_setattr(ltrain, "generator_update_begin", fake_gen_begin)
    monkeypatch_setattr(ltrain, "flow_update", fake_flow_update)
    import lsnf_b200.langevin as llang
    monkeypatch_setattr(llang, "langevin_plan", lambda *a, **k: None)


def _nets():
    args = lsnf_b200.make_args(dataset="svhn", nz=100, ngf=4, seed=5)
    return args, lsnf_b200._netG(args), lsnf_b200._netF(args, nz=100)


def _worker(rank, world, port, tmp, sizes):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log = {}
    _install_recorders(lambda obj, name, val: setattr(obj, name, val), log)
    args, netG, netF = _nets()
    start = sum(sizes[:rank])
    x = torch.zeros(sizes[rank], 3, 32, 32)
    out = {}
    # 1. defaults: global batch by all-reduce, contiguous sharding, shared default seed
    lg, lf, gn, fn, zk = lsnf_b200.training_iteration(x, netG, netF, None, None, args)
    out["default"] = dict(log["langevin"], gen=log["gen"], flow=log["flow"], start=start,
                          ret=(float(lg), float(lf), float(gn), float(fn)))
    # 2. explicit seed: z_0 must be the rank's slice of the single-process draw
    lsnf_b200.training_iteration(x, netG, netF, None, None, args, seed=1234, global_batch=sum(sizes))
    out["seeded"] = dict(log["langevin"])
    # 3. a shard that is not the contiguous one must be refused unless its offset is given
    if rank == 0:
        bad = torch.zeros(sizes[0] + 1, 3, 32, 32)
        try:
            lsnf_b200.training_iteration(bad, netG, netF, None, None, args, global_batch=sum(sizes))
            out["refused"] = False
        except ValueError:
            out["refused"] = True
        lsnf_b200.training_iteration(bad, netG, netF, None, None, args, global_batch=sum(sizes), sample_offset=3)
        out["explicit_offset"] = log["langevin"]["sample_offset"]
    # 4. data_parallel=False inside an initialised group: a single-process iteration
    lsnf_b200.training_iteration(x, netG, netF, None, None, args, data_parallel=False)
    out["single"] = dict(sample_offset=log["langevin"]["sample_offset"], world=log["gen"]["world"],
                         global_batch=log["gen"]["global_batch"])
    torch.save(out, f"{tmp}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sizes", [(6, 6), (7, 6)])
def test_training_iteration_shards_chains_across_ranks(tmp_path, sizes):
    port = 31500 + (os.getpid() + sum(sizes)) % 2000
    tmp = str(tmp_path / "dp")
    mp.spawn(_worker, args=(2, port, tmp, sizes), nprocs=2, join=True)
    r = [torch.load(f"{tmp}.{k}", weights_only=False) for k in range(2)]
    B = sum(sizes)
    for k in range(2):
        d = r[k]["default"]
        assert d["sample_offset"] == d["start"] and d["train"] is True
        assert d["gen"] == {"global_batch": B, "world": 2} and d["flow"] == {"global_batch": B, "world": 2}
        assert d["ret"] == (3.0, 4.0, 1.0, 2.0)
        assert d["z0"].shape == (sizes[k], 100, 1, 1)
    assert r[0]["default"]["seed"] == r[1]["default"]["seed"]                  # one seed for all ranks ...
    assert r[0]["default"]["seed"] >> 32 == 5                                  # ... derived from args.seed
    assert not torch.equal(r[0]["default"]["z0"][:6], r[1]["default"]["z0"][:6])   # ... but different chains
    # the two shards together are the single-process draw of the same seed
    gen = torch.Generator().manual_seed(1234)
    z_single = torch.randn(B, 100, 1, 1, generator=gen)
    assert torch.equal(torch.cat([r[0]["seeded"]["z0"], r[1]["seeded"]["z0"]]), z_single)
    assert r[0]["seeded"]["seed"] == r[1]["seeded"]["seed"] == 1234
    assert r[0]["refused"] is True and r[0]["explicit_offset"] == 3
    for k in range(2):
        assert r[k]["single"] == {"sample_offset": 0, "world": 1, "global_batch": sizes[k]}


def test_single_process_defaults(monkeypatch):
    log = {}
    _install_recorders(monkeypatch.setattr, log)
    args, netG, netF = _nets()
    x = torch.zeros(9, 3, 32, 32)
    lsnf_b200.training_iteration(x, netG, netF, None, None, args, seed=77)
    assert log["langevin"]["sample_offset"] == 0 and log["gen"] == {"global_batch": 9, "world": 1}
    assert torch.equal(log["langevin"]["z0"], torch.randn(9, 100, 1, 1, generator=torch.Generator().manual_seed(77)))
    s0 = log["langevin"]["seed"]
    lsnf_b200.training_iteration(x, netG, netF, None, None, args)
    s1 = log["langevin"]["seed"]
    lsnf_b200.training_iteration(x, netG, netF, None, None, args)
    assert s0 == 77 and log["langevin"]["seed"] == s1 ^ (s1 & 0xFFFFFFFF) ^ ((s1 & 0xFFFFFFFF) + 1)   # counter advances
