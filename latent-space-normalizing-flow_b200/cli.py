"""The reference's command-line flags (train.py:37-99) as a table, so that a command line written for the
reference's ``train.py`` parses unchanged: ``args = lsnf_b200.parse_args()`` replaces ``args = parse_args()``.

The hot path reads ``--test_mode --dataset --nz --nc --ngf --g_llhd_sigma --g_activation --g_activation_leak
--g_l_steps --g_l_step_size --g_l_with_noise --g_batchnorm --f_n_levels --f_depth --f_flow_permutation --f_width
--f_flow_coupling`` (and ``--g_lr --f_lr --g_decay --f_decay --*_beta*`` for the two optimizers); the remaining
flags (schedules, logging, FID, checkpoints) are accepted and carried so existing launch scripts keep working.
Names, defaults and argparse semantics are pinned against the reference's own parser by
``tests/golden/cli_flags.json`` (``oracle/make_cli_fixture.py``).  As upstream, ``type=bool`` flags turn any
non-empty string into ``True``.
"""
from __future__ import annotations

import argparse
from typing import Dict, Optional, Sequence, Tuple

# name -> (kind, default); kind is a type, or "flag" for store_true switches
FLAGS: Dict[str, Tuple[object, object]] = {
    "test_mode": ("flag", False), "seed": (int, 1), "gpu_deterministic": (bool, False),
    "dataset": (str, "svhn"), "img_size": (int, 32), "batch_size": (int, 100),
    "nz": (int, 100), "nc": (int, 3), "ngf": (int, 64),
    "g_llhd_sigma": (float, 0.3), "g_activation": (str, "lrelu"), "g_activation_leak": (float, 0.2),
    "g_l_steps": (int, 20), "g_l_step_size": (float, 0.1), "g_l_with_noise": (bool, True),
    "g_batchnorm": (bool, False),
    "f_n_levels": (int, 1), "f_depth": (int, 5), "f_flow_permutation": (int, 2), "f_width": (int, 64),
    "f_flow_coupling": (int, 1),
    "g_lr": (float, 0.0004), "f_lr": (float, 0.0004),
    "g_is_grad_clamp": (bool, False), "f_is_grad_clamp": (bool, False),
    "g_max_norm": (float, 100), "f_max_norm": (float, 100),
    "g_decay": (float, 0), "f_decay": (float, 0), "g_gamma": (float, 0.998), "f_gamma": (float, 0.998),
    "g_beta1": (float, 0.5), "g_beta2": (float, 0.999), "f_beta1": (float, 0.5), "f_beta2": (float, 0.999),
    "n_epochs": (int, 201), "n_printout": (int, 20), "n_plot": (int, 1), "n_ckpt": (int, 1), "n_metrics": (int, 10),
    "n_stats": (int, 1), "n_fid_samples": (int, 50000),
    "path_check_point": (str, None), "testing_reconstruct": ("flag", False),
}
DATASETS = ("svhn", "cifar10", "celeba_crop", "celeba_hq256")


def defaults() -> Dict[str, object]:
    return {k: d for k, (_, d) in FLAGS.items()}


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="flags of the reference's train.py (train.py:37-99)")
    for name, (kind, default) in FLAGS.items():
        if kind == "flag":
            p.add_argument("--" + name, action="store_true", default=default)
        elif name == "dataset":
            p.add_argument("--dataset", type=str, default=default, choices=list(DATASETS))
        else:
            p.add_argument("--" + name, type=kind, default=default)
    return p


def parse_args(argv: Optional[Sequence[str]] = None):
    """``train.py:37-99``; returns an ``AttrDict`` (the reference wraps the namespace the same way, train.py:743)."""
    from .langevin import AttrDict
    return AttrDict(vars(build_parser().parse_args(argv)))
