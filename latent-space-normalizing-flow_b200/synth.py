"""Deterministic synthetic parameters and inputs for the Langevin path.

Distributions follow the reference's initialisation -- Xavier-normal ConvTranspose2d weights
(train.py:271 -> model.py:39-42), PyTorch's default ConvTranspose2d bias, actnorm b/logs ~ N(0, 0.05^2)
(model.py:230-233), QR-orthogonal 1x1-conv matrix (model.py:176), fc.w ~ N(0, 0.05^2) (model.py:318) --
except that the zero-initialised ``fc_zeros.{w,b,logs}`` (model.py:340-342) are perturbed with
N(0, 0.05^2): with zeros the coupling MLP contributes no gradient and the flow kernels would see a
degenerate problem.

Values come from numpy's Philox bit generator, so they are identical on every machine; fixtures under
``tests/golden`` therefore store only outputs plus checksums of these parameters.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

GEN_CONV_IDX_STRIDE = 3  # nn.Sequential index of the i-th ConvTranspose2d is 3*i (model.py:56-71)


def generator_layers(dataset: str, nz: int, ngf: int, nc: int = 3) -> List[Tuple[int, int, int, int, int]]:
    """(C_in, C_out, kernel, stride, pad) of every ConvTranspose2d in ``_netG`` (model.py:52-151)."""
    if dataset == "svhn":
        return [(nz, ngf * 8, 4, 1, 0), (ngf * 8, ngf * 4, 4, 2, 1), (ngf * 4, ngf * 2, 4, 2, 1),
                (ngf * 2, nc, 4, 2, 1)]
    if dataset == "cifar10":
        return [(nz, ngf * 8, 8, 1, 0), (ngf * 8, ngf * 4, 4, 2, 1), (ngf * 4, ngf * 2, 4, 2, 1),
                (ngf * 2, nc, 3, 1, 1)]
    if dataset == "celeba_crop":
        return [(nz, ngf * 8, 4, 1, 0), (ngf * 8, ngf * 4, 4, 2, 1), (ngf * 4, ngf * 2, 4, 2, 1),
                (ngf * 2, ngf, 4, 2, 1), (ngf, nc, 4, 2, 1)]
    if dataset == "celeba_hq256":
        return [(nz, ngf * 16, 4, 1, 0), (ngf * 16, ngf * 8, 4, 2, 1), (ngf * 8, ngf * 4, 4, 2, 1),
                (ngf * 4, ngf * 2, 4, 2, 1), (ngf * 2, ngf, 4, 2, 1), (ngf, ngf, 4, 2, 1),
                (ngf, nc, 4, 2, 1)]
    raise ValueError(dataset)


def image_size(dataset: str) -> int:
    return {"svhn": 32, "cifar10": 32, "celeba_crop": 64, "celeba_hq256": 256}[dataset]


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def generator_state(dataset: str, nz: int, ngf: int, nc: int = 3, seed: int = 1) -> Dict[str, np.ndarray]:
    rng = _rng(seed, 1)
    sd = {}
    for i, (ci, co, k, _s, _p) in enumerate(generator_layers(dataset, nz, ngf, nc)):
        std = np.sqrt(2.0 / ((ci + co) * k * k))          # xavier_normal_ on a [ci, co, k, k] tensor
        bound = 1.0 / np.sqrt(co * k * k)                 # torch's default bias init for ConvTranspose2d
        sd[f"gen.{GEN_CONV_IDX_STRIDE * i}.weight"] = (rng.standard_normal((ci, co, k, k)) * std).astype(np.float32)
        sd[f"gen.{GEN_CONV_IDX_STRIDE * i}.bias"] = rng.uniform(-bound, bound, size=(co,)).astype(np.float32)
    return sd


def flow_state(nz: int, f_width: int = 64, f_depth: int = 5, coupling: int = 1, permutation: int = 2,
               seed: int = 1, perturb: float = 0.05) -> Dict[str, np.ndarray]:
    """All 17 keys per step of the reference ``_netF.state_dict()``."""
    rng = _rng(seed, 2)
    sd = {}
    n_out = nz if coupling == 1 else nz // 2

    def actnorm(pre, n):
        b = (rng.standard_normal((1, n)) * 0.05).astype(np.float32)
        sd[pre + "b"] = b
        sd[pre + "bias"] = b                                # registered alias of ``b`` (model.py:231)
        sd[pre + "logs"] = (rng.standard_normal((1, n)) * 0.05).astype(np.float32)

    for i in range(f_depth):
        pre = f"revnet2d_s.0.revnet2d_step_s.{i}."
        actnorm(pre + "actnorm.", nz)
        if permutation == 2:
            q = np.linalg.qr(rng.standard_normal((nz, nz)))[0].astype(np.float32)
            sd[pre + "invertible_1x1_conv.w"] = q
        elif permutation == 1:
            idx = rng.permutation(nz).astype(np.int32)
            inv = np.empty_like(idx)
            inv[idx] = np.arange(nz, dtype=np.int32)
            sd[pre + "shuffle_features.indices"] = idx
            sd[pre + "shuffle_features.indices_inverse"] = inv
        else:
            raise Exception()
        for name, (a, b) in (("fc_1", (nz // 2, f_width)), ("fc_2", (f_width, f_width))):
            sd[pre + f"f.{name}.w"] = (rng.standard_normal((a, b)) * 0.05).astype(np.float32)
            sd[pre + f"f.{name}.b"] = np.zeros((1, b), np.float32)   # never read (model.py:329-330)
            actnorm(pre + f"f.{name}.actnorm.", b)
        sd[pre + "f.fc_zeros.w"] = (rng.standard_normal((f_width, n_out)) * perturb).astype(np.float32)
        sd[pre + "f.fc_zeros.b"] = (rng.standard_normal((1, n_out)) * perturb).astype(np.float32)
        sd[pre + "f.fc_zeros.logs"] = (rng.standard_normal((1, n_out)) * perturb).astype(np.float32)
    return sd


def inputs(batch: int, nz: int, nc: int, img: int, steps: int, seed: int = 1, with_noise: bool = True):
    """x ~ U(-1,1) [B,nc,H,W]; z0 ~ N(0,I) [B,nz,1,1]; eps ~ N(0,I) [T,B,nz,1,1] (or None)."""
    rng = _rng(seed, 3)
    x = rng.uniform(-1.0, 1.0, size=(batch, nc, img, img)).astype(np.float32)
    z0 = rng.standard_normal((batch, nz, 1, 1)).astype(np.float32)
    eps = rng.standard_normal((steps, batch, nz, 1, 1)).astype(np.float32) if with_noise else None
    return x, z0, eps


def checksum(sd: Dict[str, np.ndarray]) -> Tuple[float, float]:
    """(sum, sum|.|) over all float entries in key order, fp64 -- stored in fixtures to detect drift."""
    s = a = 0.0
    for k in sorted(sd):
        v = np.asarray(sd[k], dtype=np.float64)
        s += float(v.sum())
        a += float(np.abs(v).sum())
    return s, a
