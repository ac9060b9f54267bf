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


def parse_prototypes():
    """{name: (return type, [parameter types])} of every function include/lsnf.h declares (comments stripped,
    parameter names dropped, `const` ignored)."""
    src = open(os.path.join(ROOT, "include", "lsnf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    for m in re.finditer(r"^[ \t]*((?:const[ \t]+)?[A-Za-z_]\w*[ \t\*]*)\b(lsnf_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", src, flags=re.M):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        types = []
        if params and params != "void":
            for p in params.split(","):
                p = re.sub(r"\bconst\b", "", p).strip()
                t = re.sub(r"\s*\b[A-Za-z_]\w*$", "", p) if not p.endswith("*") else p      # drop the parameter name
                types.append(re.sub(r"\s+", "", t))
        protos[name] = (re.sub(r"\s+", "", re.sub(r"\bconst\b", "", ret)), types)
    return protos


def test_ctypes_prototypes_agree_with_the_header_argument_by_argument():
    import ctypes as C
    scalar = {"int": C.c_int, "int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "uint64_t": C.c_uint64,
              "size_t": C.c_size_t, "float": C.c_float, "lsnf_stream": C.c_void_p}
    protos = parse_prototypes()
    assert sorted(protos) == sorted(_cabi.EXPORTS)
    for name, (ret, params) in protos.items():
        res, args = _cabi.EXPORTS[name]
        assert len(args) == len(params), f"{name}: header has {len(params)} parameters, _cabi.py {len(args)}"
        if ret == "void":
            assert res is None, name
        elif ret == "char*":
            assert res is C.c_char_p, name
        else:
            assert res is scalar[ret], f"{name}: return type {ret}"
        for i, (t, a) in enumerate(zip(params, args)):
            if t.endswith("*"):
                # any pointer: an opaque c_void_p or a typed ctypes pointer -- never a by-value scalar
                assert a is C.c_void_p or issubclass(a, C._Pointer), f"{name} arg {i}: {t} bound as {a}"
            else:
                assert a is scalar[t], f"{name} arg {i}: header says {t}, _cabi.py binds {a}"
