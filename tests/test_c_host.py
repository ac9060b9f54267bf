"""include/lsnf.h is valid C99 and the library is usable from a plain-C host (no Python, no torch types): the program
tests/c_host/host_check.c dlopen()s the library, resolves the entry points by name, plans BASELINE config 2 and walks
its stage tables.  Its view of the struct layouts and of the plan must equal the ctypes binding's (``_cabi.py``), and a
compute call on an unbound plan must fail with a status and a message -- there is no CPU fallback behind the ABI."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

from lsnf_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "latent-space-normalizing-flow_b200", "_lib", "liblsnf_b200.so")


@pytest.fixture(scope="module")
def host_output(tmp_path_factory):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    _cabi.load()   # raises if the library has not been built
    exe = str(tmp_path_factory.mktemp("c_host") / "host_check")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_host", "host_check.c"), "-o", exe, "-ldl"])
    out = subprocess.run([exe, os.environ.get("LSNF_LIB", LIB)], check=True, capture_output=True, text=True).stdout
    kv, stages = {}, []
    for line in out.splitlines():
        t = line.split()
        if t[0] == "stage":
            stages.append({t[i]: [int(v) for v in t[i + 1:i + 4]] if t[i] == "grid" else int(t[i + 1])
                           for i in range(0, len(t)) if t[i].isalpha() or "_" in t[i]})
        else:
            kv[t[0]] = int(t[1])
    return kv, stages


def test_header_is_c99_and_struct_layouts_match_the_ctypes_binding(host_output):
    kv, _ = host_output
    assert kv["abi_version"] == kv["header_abi_version"] == 1
    assert kv["sizeof_config"] == C.sizeof(_cabi.Config)
    assert kv["sizeof_tap"] == C.sizeof(_cabi.Tap)
    assert kv["sizeof_stage_info"] == C.sizeof(_cabi.StageInfo)
    assert kv["sizeof_launch_info"] == C.sizeof(_cabi.LaunchInfo)
    assert kv["offsetof_config_leak"] == _cabi.Config.leak.offset
    assert kv["offsetof_config_train"] == _cabi.Config.train.offset
    assert kv["offsetof_stage_taps"] == _cabi.StageInfo.taps.offset
    assert kv["offsetof_stage_passes"] == _cabi.StageInfo.passes.offset
    assert kv["offsetof_stage_a_offset"] == _cabi.StageInfo.a_offset.offset
    assert kv["offsetof_stage_flops"] == _cabi.StageInfo.flops.offset
    assert kv["offsetof_launch_smem_bytes"] == _cabi.LaunchInfo.smem_bytes.offset


def test_c_host_plans_the_cifar10_configuration_like_the_python_binding(host_output):
    kv, stages = host_output
    assert kv["odd_nz_rc"] == -1 and kv["odd_nz_has_message"] == 1          # LSNF_ERR_INVALID (model.py:383)
    assert kv["create_rc"] == 0
    lib = _cabi.load()
    cfg = _cabi.Config(arch=_cabi.ARCH["cifar10"], batch=100, nz=128, ngf=128, nc=3, f_depth=5, f_width=64,
                       f_permutation=2, f_coupling=1, leak=0.2, gemm_impl=_cabi.GEMM_TCGEN05, bwd_passes=0, train=0)
    h = C.c_void_p()
    _cabi.check(lib.lsnf_plan_create(C.byref(cfg), C.byref(h)), "lsnf_plan_create")
    try:
        assert kv["workspace_bytes"] == lib.lsnf_workspace_bytes(h)
        assert kv["num_stages"] == lib.lsnf_plan_num_stages(h) == len(stages) == 8
        assert kv["launches_per_40_steps"] == lib.lsnf_langevin_launch_count(h, 40)
        for i, s in enumerate(stages):
            info, li = _cabi.StageInfo(), _cabi.LaunchInfo()
            _cabi.check(lib.lsnf_plan_stage_info(h, i, C.byref(info)), "lsnf_plan_stage_info")
            _cabi.check(lib.lsnf_plan_stage_launch_info(h, i, 148, C.byref(li)), "lsnf_plan_stage_launch_info")
            assert (s["kind"], s["layer"], s["block_n"], s["passes"], s["k_splits"], s["flops"]) == \
                   (info.kind, info.layer, info.block_n, info.passes, info.k_splits, info.flops)
            assert (s["kernel"], s["grid"], s["smem"], s["tmem"]) == \
                   (li.kernel, [li.grid_x, li.grid_y, li.grid_z], li.smem_bytes, li.tmem_columns)
            assert s["smem"] <= 227 * 1024 and s["tmem"] <= 512
    finally:
        lib.lsnf_plan_destroy(h)


def test_compute_on_an_unbound_plan_fails_loudly(host_output):
    kv, _ = host_output
    assert kv["unbound_run_rc"] in (-2, -3)      # LSNF_ERR_CUDA / LSNF_ERR_STATE: never a silent host-side "success"
    assert kv["unbound_run_has_message"] == 1
