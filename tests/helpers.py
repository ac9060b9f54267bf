"""Shared helpers for the test-suite (fixture loading, synthetic cases, tolerances)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-4  # north_star: z_T, log p(z), log-det within 1e-4 relative (fp32)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {k: d[k] for k in d.files}
    out["config"] = ast.literal_eval(str(out["config"]))
    return out


def to_torch(sd, dtype=torch.float32, device="cpu"):
    out = {}
    for k, v in sd.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        out[k] = (t.to(dtype) if t.is_floating_point() else t).to(device)
    return out


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
