"""CPU-side checks of the C-ABI boundary: libecgmm.so loads, exports every symbol that
include/ecgmm.h declares, and the ctypes prototypes in ecgmm.lib match the header parameter by
parameter.  No compute entry point is called (there is no GPU here)."""
import ctypes
import os
import re

import pytest

import ecgmm
from ecgmm import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ecgmm.h")


def parse_header():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(unsigned long long|long long|int|void|const char\*)\s+(ecgmm_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, params)
    return protos


def ctype_of(param: str):
    if "*" in param:
        return "ptr"
    t = param.rsplit(" ", 1)[0].strip()
    return {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
            "unsigned long long": ctypes.c_ulonglong, "double": ctypes.c_double}[t]


def test_header_parses_all_entry_points():
    protos = parse_header()
    assert len(protos) >= 40
    assert set(protos) == set(lib.SIGNATURES), set(protos) ^ set(lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    so = lib.load()
    for name in parse_header():
        assert hasattr(so, name), f"{name} declared in ecgmm.h but not exported by libecgmm.so"
    assert so.ecgmm_version() >= 101
    assert so.ecgmm_adam_chunk_bytes() == 40


@pytest.mark.parametrize("name", sorted(lib.SIGNATURES))
def test_ctypes_prototype_matches_header(name):
    ret, params = parse_header()[name]
    argtypes = lib.SIGNATURES[name]
    assert len(argtypes) == len(params), f"{name}: header has {len(params)} parameters, ctypes {len(argtypes)}"
    for i, (p, a) in enumerate(zip(params, argtypes)):
        want = ctype_of(p)
        if want == "ptr":
            assert a is ctypes.c_void_p or issubclass(a, ctypes._Pointer), f"{name} arg {i} ({p}) should be a pointer"
        else:
            assert a is want, f"{name} arg {i} ({p}): ctypes has {a}"


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(ImportError):
        lib.load()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.EcgmmError):
        lib.require_device()
    from ecgmm import ops

    with pytest.raises(lib.EcgmmError):
        ops.nchw_to_nhwc_bf16(torch.zeros(1, 8, 2, 2))


def test_smoke_entry_resolves_its_imports_before_asking_for_a_device():
    """__graft_entry__.smoke() must get as far as the device check on a CPU-only box (a `tests` package in
    site-packages once shadowed its `from tests.parity_util import ...`)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import __graft_entry__ as g

    with pytest.raises(lib.EcgmmError):
        g.smoke()


def test_product_package_never_touches_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under ecg-multimodal-model_b200/ may import or read oracle/ or
    /root/reference, and nothing the GPU box runs may read /root/reference."""
    import glob

    pkg = glob.glob(os.path.join(ROOT, "ecg-multimodal-model_b200", "*.py"))
    assert pkg
    for f in pkg:
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
        assert "/root/reference" not in src.replace("/root/reference/multimodal", "REFDOC") or f.endswith("model.py"), f
    for f in [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + \
            glob.glob(os.path.join(ROOT, "tests", "test_*gpu.py")):
        code = "\n".join(l for l in open(f).read().splitlines() if not l.strip().startswith("#"))
        assert 'open("/root/reference' not in code and "sys.path.insert(0, \"/root/reference\")" not in code, f
