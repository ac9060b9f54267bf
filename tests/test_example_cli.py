"""examples/run_synthetic.py: the reference's README command lines parse unchanged, its checkpoints carry the
reference's keys and round-trip through the host mirror (modules + FusedAdam), and without a CUDA device the script
stops with an error instead of falling back to the CPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

import lsnf_b200  # noqa: E402
import run_synthetic as ex  # noqa: E402

# /root/reference/README.md:30-70, with `python train.py` dropped
README_COMMANDS = [
    "--dataset svhn --g_l_steps 20 --img_size 32 --nz 100 --ngf 64 --g_lr 0.0004 --f_lr 0.0004",
    "--dataset cifar10 --g_l_steps 40 --img_size 32 --nz 128 --ngf 128 --g_lr 0.00038 --f_lr 0.00038",
    "--dataset celeba_crop --g_l_steps 20 --img_size 64 --nz 100 --ngf 128 --g_lr 0.0003 --f_lr 0.0003",
    "--dataset svhn --test_mode --g_l_steps 400 --img_size 32 --nz 100 --ngf 64 --g_lr 0.0004 --f_lr 0.0004 "
    "--path_check_point ./ckpt/ckpt_000115.pth --n_fid_samples 50000",
    "--dataset cifar10 --test_mode --g_l_steps 800 --img_size 32 --nz 128 --ngf 128 --g_lr 0.00038 --f_lr 0.00038 "
    "--path_check_point ./ckpt/ckpt_000093.pth --n_fid_samples 50000",
    "--dataset celeba_hq256 --g_l_steps 20 --img_size 256 --nz 100 --ngf 128 --g_llhd_sigma 1.0 --f_width 128 "
    "--g_lr 0.0003 --f_lr 0.0003",
]


@pytest.mark.parametrize("line", README_COMMANDS)
def test_readme_command_lines_parse_and_build_the_networks(line):
    a = ex.parse(line.split())
    assert a.f_depth == 5 and a.g_l_step_size == 0.1 and a.batch_size == 100          # reference defaults
    assert a.test_mode == ("--test_mode" in line)
    if a.ngf <= 64:   # constructing the wide generators on the CPU is slow and adds nothing
        netG, netF = ex.build(a, "cpu")
        assert netG.nz == netF.nz == a.nz
        assert list(netG.state_dict())[0] == "gen.0.weight"


def test_mismatched_image_size_is_rejected():
    with pytest.raises(SystemExit):
        ex.parse("--dataset cifar10 --img_size 64".split())


def test_checkpoint_has_the_reference_keys_and_round_trips(tmp_path):
    a = ex.parse("--dataset svhn --nz 100 --ngf 8".split())
    netG, netF = ex.build(a, "cpu")
    optG, optF = lsnf_b200.make_optimizers(netG, netF, a)
    # give the optimizers state, as after one update (plain torch step on the CPU: FusedAdam IS a torch.optim.Adam)
    for p in list(netG.parameters()) + list(netF.parameters()):
        p.grad = torch.full_like(p, 1e-3)
    optG.step()
    optF.step()
    d = ex.checkpoint_dict(7, netG, netF, optG, optF)
    assert sorted(d) == ["epoch", "netF", "netG", "optF", "optG"]                      # train.py:495-501
    path = str(tmp_path / "ckpt_000007.pth")
    torch.save(d, path)
    netG2, netF2 = ex.build(a, "cpu")
    optG2, optF2 = lsnf_b200.make_optimizers(netG2, netF2, a)
    assert ex.load_checkpoint(path, netG2, netF2, optG2, optF2) == 8                   # resumes at epoch + 1
    for k, v in netG.state_dict().items():
        assert torch.equal(v, netG2.state_dict()[k])
    for k, v in netF.state_dict().items():
        assert torch.equal(v, netF2.state_dict()[k])
    s1, s2 = optG.state_dict()["state"], optG2.state_dict()["state"]
    assert s1.keys() == s2.keys() and all(torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"]) for i in s1)
    # a plain torch.optim.Adam (the reference's optimizer) loads the same state
    ref_opt = torch.optim.Adam(netG2.parameters(), lr=a.g_lr, betas=(a.g_beta1, a.g_beta2))
    ref_opt.load_state_dict(d["optG"])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_without_a_gpu_the_script_stops_instead_of_falling_back():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.main("--dataset svhn --nz 100 --ngf 8 --n_epochs 1 --iters_per_epoch 1".split())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.main("--dataset svhn --nz 100 --ngf 8 --test_mode".split())


def test_train_and_test_loops_run_end_to_end_with_the_device_calls_stubbed(tmp_path, monkeypatch, capsys):
    """The script's own logic -- epochs, the reference's log line, ExponentialLR per epoch, checkpoint files, resume,
    test mode -- with lsnf_b200's three device entry points replaced by CPU recorders (the entry points themselves are
    covered by the GPU suite)."""
    calls = {"train": [], "sample": 0, "recon": 0}

    def fake_iteration(x, netG, netF, optG, optF, args, **kw):
        calls["train"].append(dict(batch=x.shape[0], **{k: kw[k] for k in ("global_batch", "sample_offset", "seed")}))
        for opt in (optG, optF):          # what fused_step does, as far as the schedulers and checkpoints can tell
            for grp in opt.param_groups:
                for p in grp["params"]:
                    p.grad = torch.zeros_like(p)
            opt.step()
        return torch.tensor(10.0), torch.tensor(5.0), torch.tensor(1.0), torch.tensor(2.0), None

    def fake_sample_x(netG, netF, n, device, generator=None, **kw):
        calls["sample"] += 1
        return torch.rand(n, 3, 32, 32, generator=generator)

    def fake_recon(batches, netG, netF, args, generator=None):
        calls["recon"] += sum(1 for _ in batches)
        return 0.125

    monkeypatch.setattr(ex, "cuda_device", lambda local: torch.device("cpu"))
    monkeypatch.setattr(lsnf_b200, "training_iteration", fake_iteration)
    monkeypatch.setattr(lsnf_b200, "sample_x", fake_sample_x)
    monkeypatch.setattr(lsnf_b200, "reconstruction_error", fake_recon)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    ckpt = tmp_path / "ckpt"
    base = f"--dataset svhn --nz 100 --ngf 8 --batch_size 12 --iters_per_epoch 3 --n_printout 2 --ckpt_dir {ckpt}"
    ex.main((base + " --n_epochs 2 --g_gamma 0.5").split())
    out = capsys.readouterr().out
    assert len(calls["train"]) == 6 and calls["train"][0] == dict(batch=12, global_batch=12, sample_offset=0, seed=1 << 32)
    assert len({c["seed"] for c in calls["train"]}) == 6                      # a fresh noise seed every iteration
    assert "loss_g=  10.000, loss_f=   5.000, z_g_grad_norm=   1.000, z_f_grad_norm=   2.000" in out
    assert "lr_g=0.000400" in out and "lr_g=0.000200" in out                   # ExponentialLR stepped once per epoch
    assert sorted(os.listdir(ckpt)) == ["ckpt_000000.pth", "ckpt_000001.pth"]
    ck = torch.load(ckpt / "ckpt_000001.pth")
    assert ck["epoch"] == 1 and ck["optG"]["param_groups"][0]["lr"] == pytest.approx(0.0001)
    # resume (train.py:342-349): continues at epoch 2 with the optimizer state of the checkpoint
    calls["train"].clear()
    ex.main((base + f" --n_epochs 3 --path_check_point {ckpt / 'ckpt_000001.pth'}").split())
    assert len(calls["train"]) == 3 and "ckpt_000002.pth" in os.listdir(ckpt)
    # test mode (train.py:520-662): n_fid_samples / batch_size sampling calls, then the reconstruction report
    ex.main((base + f" --test_mode --testing_reconstruct --n_fid_samples 48 --n_test_batches 2 "
                    f"--path_check_point {ckpt / 'ckpt_000002.pth'}").split())
    out = capsys.readouterr().out
    assert calls["sample"] == 4 and calls["recon"] == 2
    assert "48 prior samples" in out and "reconstruction error=0.125" in out
