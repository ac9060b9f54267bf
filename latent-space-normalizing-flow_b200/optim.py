"""``FusedAdam``: ``torch.optim.Adam`` (train.py:294-295) whose update can be applied by the library's fused
multi-tensor kernel (``lsnf_adam_step``) straight from the flat gradient buffers the gradient kernels fill.

It IS a ``torch.optim.Adam``: same constructor, ``param_groups``, per-parameter state (``step``, ``exp_avg``,
``exp_avg_sq``) and ``state_dict()``, so checkpoints written by the reference (``ckpt['optG']``, ``ckpt['optF']``,
train.py:497-503) load into it and ``ExponentialLR`` (train.py:297-298) drives it unchanged; ``step()`` on autograd
gradients still works.  ``fused_step`` is the path ``train.training_iteration`` takes."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _cabi


def bump_versions(params: Sequence[torch.Tensor]) -> None:
    """The kernels write parameters through raw pointers; tell autograd / the plans' change tracking about it."""
    torch._C._increment_version(list(params))


class FusedAdam(torch.optim.Adam):
    def _ensure_state(self, p: torch.Tensor):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def fused_step(self, params: Sequence[torch.nn.Parameter], grads: Sequence[torch.Tensor], *,
                   grad_kk: Optional[Sequence[int]] = None, grad_inner: Optional[Sequence[int]] = None,
                   grad_scale: Optional[torch.Tensor] = None) -> None:
        """One Adam step of ``params`` from ``grads`` (fp32 CUDA tensors / views, one per parameter, same element
        count).  ``grad_kk[i] > 1`` declares gradient i tap-major ``[kk][C_in][C_out]`` for a ``[C_in, C_out, k, k]``
        weight (``grad_inner[i]`` = C_out).  ``grad_scale``: 0-d CUDA tensor multiplied into every gradient (norm
        clipping).  Hyper-parameters come from the parameter group each tensor belongs to."""
        lib = _cabi.load()
        key = tuple(id(p) for p in params)
        cached = getattr(self, "_fused_groups", {}).get(key)
        if cached is None:   # first use of this parameter set: group lookup and tensor validation, done once
            group_of = {}
            for gi, g in enumerate(self.param_groups):
                if g.get("amsgrad") or g.get("maximize"):
                    raise NotImplementedError("FusedAdam.fused_step: amsgrad / maximize are not used by the reference")
                for p in g["params"]:
                    group_of[id(p)] = gi
            by_group = {}
            for i, p in enumerate(params):
                if id(p) not in group_of:
                    raise ValueError("parameter does not belong to this optimizer")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam.fused_step needs contiguous float32 CUDA parameters and gradients")
                by_group.setdefault(group_of[id(p)], []).append(i)
            cached = list(by_group.items())
            if not hasattr(self, "_fused_groups"):
                self._fused_groups = {}
            self._fused_groups[key] = cached
        for gi, idxs in cached:
            g = self.param_groups[gi]
            states = [self._ensure_state(params[i]) for i in idxs]
            step_tensors = [st["step"] for st in states]
            torch._foreach_add_(step_tensors, 1)
            step = int(step_tensors[0].item())
            if int(step_tensors[-1].item()) != step:
                raise RuntimeError("FusedAdam.fused_step: parameters of one group are at different step counts")
            n = len(idxs)
            for i in idxs:
                gr = grads[i]
                if not (gr.is_cuda and gr.dtype == torch.float32 and gr.is_contiguous() and gr.numel() == params[i].numel()):
                    raise RuntimeError("FusedAdam.fused_step needs contiguous float32 CUDA parameters and gradients")
            ptr = C.c_void_p * n
            dev = params[idxs[0]].device
            with torch.cuda.device(dev):
                _cabi.check(lib.lsnf_adam_step(
                    n, ptr(*[params[i].data_ptr() for i in idxs]), ptr(*[grads[i].data_ptr() for i in idxs]),
                    ptr(*[st["exp_avg"].data_ptr() for st in states]), ptr(*[st["exp_avg_sq"].data_ptr() for st in states]),
                    (C.c_int64 * n)(*[params[i].numel() for i in idxs]),
                    (C.c_int32 * n)(*[int(grad_kk[i]) if grad_kk is not None else 1 for i in idxs]),
                    (C.c_int32 * n)(*[int(grad_inner[i]) if grad_inner is not None else 1 for i in idxs]),
                    float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                    float(g["weight_decay"]), step,
                    C.c_void_p(grad_scale.data_ptr()) if grad_scale is not None else None,
                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "lsnf_adam_step")
        self._opt_called = True   # what torch's LR schedulers look at to tell "scheduler stepped before the optimizer"
        bump_versions(params)
