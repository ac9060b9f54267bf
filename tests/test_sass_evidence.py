"""The library that was just built really contains the Blackwell paths DESIGN.md claims, kernel by kernel (static:
`cuobjdump -sass` of the in-tree .so, no GPU): tcgen05.mma (UTCHMMA, `.2CTA` in the CTA-pair kernel), TMEM loads (LDTM),
TMA tensor loads / stores (UTMALDG / UTMASTG), bulk copies with mbarriers in the flow kernels (UBLKCP, SYNCS),
programmatic dependent launch (ACQBULK / PREEXIT) -- and no legacy HMMA anywhere.  Guards against a build that silently
lost an instruction path, and keeps profiles/r2_sass_histogram.json honest."""
import collections
import json
import os
import re
import shutil
import subprocess

import pytest

from lsnf_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not available")
    _cabi.load()
    txt = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    mangled = re.findall(r"Function : (\S+)", txt)
    names = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(mangled, names))
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = re.sub(r"\(.*", "", demangle[m.group(1)]).replace("void ", "").replace("lsnf::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur[op.split(".")[0]] += 1
            if op.startswith("UTCHMMA.2CTA"):
                cur["UTCHMMA.2CTA"] += 1
    return kernels


def test_tap_gemm_kernels_are_tcgen05_tmem_tma(sass):
    single = [k for k in sass if k.startswith("tapgemm_tc_kernel<")]
    pair = [k for k in sass if k.startswith("tapgemm_tc2_kernel<")]
    assert sorted(single) == [f"tapgemm_tc_kernel<{n}>" for n in (128, 16, 256, 32, 64)] and len(pair) == 2
    for k in single + pair:
        c = sass[k]
        assert c["UTCHMMA"] >= 24, k          # tcgen05.mma: 3 passes x 4 K steps x accumulate / overwrite variants
        assert c["LDTM"] >= 1, k              # tcgen05.ld: the epilogue reads the accumulator from TMEM
        assert c["UTMALDG"] >= 6, k           # cp.async.bulk.tensor loads of both operands
        assert c["SYNCS"] >= 1, k             # mbarrier traffic of the operand ring
        assert c["ACQBULK"] >= 1 and c["PREEXIT"] >= 1, k   # griddepcontrol.wait / launch_dependents
        assert c["HMMA"] == 0, k              # no mma.sync / wmma anywhere
    for k in pair:
        assert sass[k]["UTCHMMA.2CTA"] >= 24, k               # cta_group::2 MMAs issued by the leader CTA
        assert sass[k]["UTMASTG"] >= 1, k                     # tensor-store epilogue
    assert all(sass[k]["UTCHMMA.2CTA"] == 0 for k in single)
    for n in (128, 256):
        assert sass[f"tapgemm_tc_kernel<{n}>"]["UTMASTG"] >= 1


def test_flow_kernels_stream_their_matrices_with_bulk_copies(sass):
    flows = [k for k in sass if k.startswith(("flow_forward_kernel<", "flow_inverse_kernel<"))]
    assert len(flows) == 12
    for k in flows:
        assert sass[k]["UBLKCP"] >= 1 and sass[k]["SYNCS"] >= 1, k
        assert sass[k]["UTCHMMA"] == 0 and sass[k]["HMMA"] == 0, k   # fp32 CUDA-core mat-vecs: 0.02 % of the FLOPs


def test_no_kernel_of_the_library_uses_legacy_tensor_core_instructions(sass):
    assert len(sass) >= 38
    assert sum(c["HMMA"] + c["IMMA"] + c["QGMMA"] + c["HGMMA"] for c in sass.values()) == 0


def test_committed_histogram_agrees_with_this_build(sass):
    doc = json.load(open(os.path.join(ROOT, "profiles", "r2_sass_histogram.json")))
    for name, rec in doc["kernels"].items():
        for op in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP"):
            assert sass[name][op] == rec.get(op, 0), (name, op)
