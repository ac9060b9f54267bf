"""The command-line surface of the reference (train.py:37-99) is kept flag for flag: names, defaults and argparse
semantics are compared with a fixture written from the reference's own parser (oracle/make_cli_fixture.py)."""
import json
import os

import pytest

import lsnf_b200
from lsnf_b200 import cli

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cli_flags.json")))["flags"]


def test_every_reference_flag_exists_with_the_reference_default():
    got = lsnf_b200.parse_args([])
    assert sorted(got) == sorted(FIX)
    for k, v in FIX.items():
        assert got[k] == v["default"] and type(got[k]).__name__ == v["type"], k
    assert lsnf_b200.make_args() == got


def test_a_reference_command_line_parses_unchanged():
    # README command of the reference's CIFAR-10 run (BASELINE.json configs[1]) plus the test-mode switch
    a = lsnf_b200.parse_args("--dataset cifar10 --nz 128 --ngf 128 --g_l_steps 40 --f_width 64 --g_llhd_sigma 0.3 "
                             "--test_mode --path_check_point ckpt.pth".split())
    assert (a.dataset, a.nz, a.ngf, a.g_l_steps, a.f_width, a.g_llhd_sigma) == ("cifar10", 128, 128, 40, 64, 0.3)
    assert a.test_mode is True and a.path_check_point == "ckpt.pth" and a.g_l_with_noise is True
    assert a["nz"] == a.nz                     # AttrDict access pattern of train.py:743-746
    # argparse's type=bool quirk is kept: any non-empty string is True (train.py:56)
    assert lsnf_b200.parse_args(["--g_l_with_noise", "False"]).g_l_with_noise is True
    with pytest.raises(SystemExit):
        lsnf_b200.parse_args(["--dataset", "mnist"])
    with pytest.raises(SystemExit):
        lsnf_b200.parse_args(["--no_such_flag", "1"])


def test_make_args_takes_the_flag_names_and_rejects_unknown_ones():
    a = lsnf_b200.make_args(dataset="celeba_hq256", f_width=128, g_llhd_sigma=1.0)
    assert (a.dataset, a.f_width, a.g_llhd_sigma, a.nz) == ("celeba_hq256", 128, 1.0, cli.FLAGS["nz"][1])
    with pytest.raises(TypeError):
        lsnf_b200.make_args(latent_dim=3)
