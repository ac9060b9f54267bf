"""The C-ABI library loads without a GPU and exports every symbol include/lsnf.h declares."""
import os
import re

from lsnf_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "lsnf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsnf_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _cabi.load()
    names = declared_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in lsnf.h but not exported"
        assert n in _cabi.EXPORTS, f"{n} has no ctypes prototype in _cabi.py"
    assert sorted(_cabi.EXPORTS) == names
    assert lib.lsnf_abi_version() == 1


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_cabi.Config) == 4 * 16
    assert C.sizeof(_cabi.Tap) == 16
    # kind..n_phases (12) + n_taps (4) + taps (4*16*4) + out_mul (1) + off_y/x (8) + 11 ints, then 4 int64
    ints = 12 + 4 + 4 * 16 * 4 + 1 + 8 + 12
    assert C.sizeof(_cabi.StageInfo) == (ints * 4 + 7) // 8 * 8 + 32
