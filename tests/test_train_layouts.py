"""Host logic of the parameter-update path (no GPU): the flat gradient layouts the library reports line up with the
modules' parameters, training plans only add workspace, FusedAdam keeps torch.optim.Adam's state format."""
import ctypes as C

import pytest
import torch

import lsnf_b200
from lsnf_b200 import _cabi
from lsnf_b200.train import flow_params_in_order

import bench


def _plan(arch, batch, nz, ngf, f_width=64, coupling=1, perm=2, train=1):
    lib = _cabi.load()
    cfg = _cabi.Config(arch=_cabi.ARCH[arch], batch=batch, nz=nz, ngf=ngf, nc=3, f_depth=5, f_width=f_width,
                       f_permutation=perm, f_coupling=coupling, leak=0.2, gemm_impl=0, bwd_passes=3, train=train)
    h = C.c_void_p()
    _cabi.check(lib.lsnf_plan_create(C.byref(cfg), C.byref(h)), "lsnf_plan_create")
    return lib, h


@pytest.mark.parametrize("name", sorted(bench.WORKLOADS))
def test_generator_gradient_layout_matches_the_parameters(name):
    w = bench.WORKLOADS[name]
    lib, h = _plan(w["dataset"], w["B"], w["nz"], w["ngf"], w["f_width"])
    netG = lsnf_b200._netG(lsnf_b200.make_args(dataset=w["dataset"], nz=w["nz"], ngf=w["ngf"]))
    convs = [m for m in netG.gen if isinstance(m, torch.nn.ConvTranspose2d)]
    n = 2 * len(convs)
    off, size = (C.c_int64 * n)(), (C.c_int64 * n)()
    _cabi.check(lib.lsnf_generator_grad_layout(h, off, size), "layout")
    params = [p for m in convs for p in (m.weight, m.bias)]
    end = 0
    for i, p in enumerate(params):
        assert size[i] == p.numel() and off[i] >= end and off[i] % 4 == 0     # natural layouts, 16-byte aligned
        end = off[i] + size[i]
    assert lib.lsnf_generator_grad_floats(h) >= end
    assert sum(size) == sum(p.numel() for p in netG.parameters())
    # a layer's weight and bias gradients are adjacent: one all-reduce bucket per layer
    for l in range(len(convs)):
        assert off[2 * l + 1] - (off[2 * l] + size[2 * l]) < 4
    lib.lsnf_plan_destroy(h)
    # an inference plan has no training buffers and refuses the call
    lib, h0 = _plan(w["dataset"], w["B"], w["nz"], w["ngf"], w["f_width"], train=0)
    lib1, h1 = _plan(w["dataset"], w["B"], w["nz"], w["ngf"], w["f_width"], train=1)
    assert lib.lsnf_workspace_bytes(h0) < lib.lsnf_workspace_bytes(h1)
    assert lib.lsnf_generator_grad_layout(h0, off, size) == -3 and b"train" in lib.lsnf_last_error()
    assert lib.lsnf_generator_param_grads(h0, None, None, 1, None, None, -1, None) == -3     # not bound
    lib.lsnf_plan_destroy(h0)
    lib.lsnf_plan_destroy(h1)


@pytest.mark.parametrize("nz,w,coupling,perm", [(128, 64, 1, 2), (100, 64, 1, 2), (100, 128, 1, 2), (100, 64, 0, 2),
                                                (100, 64, 1, 1)])
def test_flow_gradient_layout_matches_the_parameters(nz, w, coupling, perm):
    lib, h = _plan("none", 7, nz, 0, w, coupling, perm)
    args = lsnf_b200.make_args(nz=nz, f_width=w, f_flow_coupling=coupling, f_flow_permutation=perm)
    netF = lsnf_b200._netF(args, nz=nz)
    n = 5 * _cabi.FLOW_PTRS_PER_STEP
    off, size = (C.c_int64 * n)(), (C.c_int64 * n)()
    _cabi.check(lib.lsnf_flow_grad_layout(h, off, size), "layout")
    params = flow_params_in_order(netF)
    assert len(params) == n
    end, covered = 0, set()
    for i, p in enumerate(params):
        assert off[i] >= end
        end = off[i] + size[i]
        if p is None:
            assert perm == 1 and i % 12 == 2          # the 1x1-conv slot of a shuffle step
            continue
        assert size[i] == p.numel()
        covered.add(id(p))
    assert lib.lsnf_flow_grad_floats(h) >= end
    # every parameter that receives a gradient upstream is covered: all but the never-read fc.b (model.py:329-330)
    # and the integer shuffle indices
    for k, p in netF.named_parameters():
        used = p.requires_grad and not k.endswith("fc_1.b") and not k.endswith("fc_2.b")
        assert (id(p) in covered) == used, k
    lib.lsnf_plan_destroy(h)


def test_fused_adam_is_a_torch_adam():
    p = [torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(3, 2))]
    opt = lsnf_b200.FusedAdam(p, lr=4e-4, betas=(0.5, 0.999), weight_decay=0.0)
    assert isinstance(opt, torch.optim.Adam)
    for q in p:
        q.grad = torch.ones_like(q)
    opt.step()                                              # the autograd-gradient path still works (CPU tensors)
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 2))], lr=4e-4,
                           betas=(0.5, 0.999))
    ref.load_state_dict(opt.state_dict())                   # reference checkpoints: ckpt['optG'] (train.py:499-502)
    assert float(ref.state[ref.param_groups[0]["params"][0]]["step"]) == 1.0
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, 0.998)      # train.py:297-298
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - 4e-4 * 0.998) < 1e-12
    with pytest.raises(RuntimeError):
        opt.fused_step(p, [torch.ones_like(q) for q in p])  # the fused path needs CUDA tensors: no CPU fallback
