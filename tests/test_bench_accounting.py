"""bench.py's ALGORITHMIC work per latent-step -- the numerator of `roofline.achieved` and of
`frac_of_tensor_roofline` -- is the figure SURVEY.md section 8(d) states: 2 x the generator's exact forward FLOPs (taps
that land inside the output; forward + data gradient, ONE pass each) + the flow term; extra split-precision passes and
zero-filled border taps never count.  Also: the library's nominal per-stage FLOPs are torch FlopCounter's."""
import ctypes as C

import pytest

import bench
from lsnf_b200 import _cabi, synth

# SURVEY.md section 8(a) layer table, "exact MFLOP" column, and section 8(d) "ALGORITHMIC work per latent-step"
EXACT_MFLOP = {
    "svhn": [1.64, 51.38, 58.98, 2.95],
    "cifar10": [16.78, 943.72, 1007.68, 13.57],
    "celeba_crop": [3.28, 205.52, 235.93, 251.92, 12.19],
    "celeba_hq256": [6.55, 822.08, 943.72, 1007.68, 1040.45, 2114.06, 199.76],
}
NOMINAL_MFLOP = {"svhn": 139.0, "cifar10": 2178.4, "celeba_crop": 821.2, "celeba_hq256": 6650.3}
ALG_GFLOP_PER_LATENT_STEP = {"svhn": 0.2304, "cifar10": 3.9641, "celeba_crop": 1.4181, "celeba_hq256": 12.2695}
FLOW_MFLOP = {"svhn": 0.474, "cifar10": 0.655, "celeba_crop": 0.474, "celeba_hq256": 0.912}


@pytest.mark.parametrize("name", sorted(bench.WORKLOADS))
def test_algorithmic_flops_are_the_survey_figures(name):
    w = bench.WORKLOADS[name]
    layers = synth.generator_layers(w["dataset"], w["nz"], w["ngf"])
    exact = bench.exact_layer_flops(layers)
    assert [round(e / 1e6, 2) for e in exact] == EXACT_MFLOP[name]
    flow = bench.flow_flops(w["nz"], w["f_width"])
    assert flow / 1e6 == pytest.approx(FLOW_MFLOP[name], abs=6e-4)
    alg = 2.0 * sum(exact) + flow
    assert alg / 1e9 == pytest.approx(ALG_GFLOP_PER_LATENT_STEP[name], abs=1e-4)
    # the roofline table of BASELINE.md section 3 (latent-steps/s at 100 % of the sustained bf16 peak) follows from it
    assert 1376.8e12 / alg == pytest.approx(bench.ROOFLINE_LATENT_STEPS[name], rel=5e-3)


@pytest.mark.parametrize("name", sorted(bench.WORKLOADS))
def test_library_stage_flops_are_nominal_and_never_below_the_algorithmic_ones(name):
    w = bench.WORKLOADS[name]
    lib = _cabi.load()
    cfg = _cabi.Config(arch=_cabi.ARCH[w["dataset"]], batch=w["B"], nz=w["nz"], ngf=w["ngf"], nc=3, f_depth=5,
                       f_width=w["f_width"], f_permutation=2, f_coupling=1, leak=0.2, gemm_impl=0, bwd_passes=3, train=0)
    h = C.c_void_p()
    _cabi.check(lib.lsnf_plan_create(C.byref(cfg), C.byref(h)), "lsnf_plan_create")
    try:
        layers = synth.generator_layers(w["dataset"], w["nz"], w["ngf"])
        exact = bench.exact_layer_flops(layers)
        n = lib.lsnf_plan_num_stages(h)
        assert n == 2 * len(layers)
        fwd_nominal = 0.0
        for i in range(n):
            info = _cabi.StageInfo()
            _cabi.check(lib.lsnf_plan_stage_info(h, i, C.byref(info)), "lsnf_plan_stage_info")
            assert info.passes == 3
            # what the tensor pipe is asked to do (padding of K / N tiles and zero-filled taps included) is never
            # less than the algorithmic work the roofline claim counts
            assert info.flops >= exact[info.layer] * w["B"] * (1 - 1e-9)
            if info.kind == 0:
                fwd_nominal += info.flops / w["B"]
        # hidden layers are counted exactly as torch.utils.flop_counter does (2 H_in^2 C_in C_out k^2); the first and
        # last layers carry K / N padding on top, so the sum is bounded below by the FlopCounter figure
        assert fwd_nominal / 1e6 >= NOMINAL_MFLOP[name] * (1 - 1e-3)
        assert fwd_nominal / 1e6 <= NOMINAL_MFLOP[name] * 1.35
    finally:
        lib.lsnf_plan_destroy(h)


def test_reference_arm_times_one_whole_training_iteration_of_config_1():
    """`bench.py --impl reference --mode train --workload svhn`: BASELINE.json's config 1 ("one training iteration on
    CPU: flow prior + short-run Langevin + generator update") on the host cores, same JSON contract as the CUDA arm's
    `--mode train` line."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--mode", "train",
                        "--workload", "svhn", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_iteration_latent_steps_per_sec"
    assert line["mode"] == "train" and line["gpu_launches"] == 0 and line["higher_is_better"] is True
    assert line["config"] == bench.workload_config("svhn", bench.WORKLOADS["svhn"])
    d = line["details"]
    assert d["ms_per_iteration"] == pytest.approx(d["langevin_call_ms"] + d["updates_ms"], rel=1e-9)
    assert line["value"] == pytest.approx(100 * 20 / (d["ms_per_iteration"] * 1e-3), rel=1e-9)
    assert d["langevin_call_ms"] > d["updates_ms"] > 0           # 1 467 ms against 110 ms on 8 vCPU
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "latent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
