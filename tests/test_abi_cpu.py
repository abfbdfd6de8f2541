"""CPU suite: the C-ABI library loads and exports every symbol include/ragfin.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from ragfin_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        ge.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from ragfin_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ragfin.h")).read()
    declared = set(re.findall(r"\b(ragfin_[a-z_]+)\s*\(", hdr))
    declared.discard("ragfin_t")
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ragfin_abi_version() == int(re.search(r"RAGFIN_ABI_VERSION (\d+)", hdr).group(1))


def test_argument_validation_without_device(lib):
    h = ctypes.c_void_p()
    assert lib.ragfin_create(ctypes.byref(h), 0, 0, 10, 0) == -1       # EINVAL: dim
    assert b"dim" in lib.ragfin_last_error()
    assert lib.ragfin_create(ctypes.byref(h), 8, 7, 10, 0) == -1       # EINVAL: dtype
    assert lib.ragfin_create(ctypes.byref(h), 8, 0, 0, 0) == -1        # EINVAL: capacity
    assert lib.ragfin_search(None, None, 1, 1, None, None, None) == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import ragfin_b200
    with pytest.raises(ragfin_b200.RagfinError) as e:
        ragfin_b200.Index(384, "f32", 16)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ragfin_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "ragfin_oracle" not in src.replace("oracle/ragfin_oracle", ""), fn


def test_view_argument_validation_without_device(lib):
    out = ctypes.c_void_p()
    assert lib.ragfin_create_view(None, ctypes.byref(out)) == -1       # EINVAL: no parent
    assert out.value is None


def test_ctypes_prototypes_match_the_header(lib):
    """Every entry point's ctypes argtypes agree with its prototype in include/ragfin.h, parameter by parameter: pointer,
    32-bit or 64-bit integer.  (A Python int passed without argtypes is truncated to 32 bits - a device pointer would not
    survive it - and a miscounted parameter shifts everything behind it.)"""
    from ragfin_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ragfin.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = dict(re.findall(r"\b(ragfin_[a-z_]+)\s*\(([^()]*)\)\s*;", hdr))
    assert set(protos) == set(_lib.SYMBOLS)

    def kind_of_c(param):
        p = " ".join(param.split())
        if p in ("void", ""):
            return None
        if "*" in p:
            return "ptr"
        base = p.replace("const ", "").split(" ")[0]
        return {"int32_t": "i32", "int": "i32", "uint32_t": "i32", "int64_t": "i64", "uint64_t": "i64", "size_t": "i64"}[base]

    def kind_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
            return "ptr"
        return {4: "i32", 8: "i64"}[ctypes.sizeof(t)]

    for name, params in protos.items():
        want = [k for k in (kind_of_c(p) for p in params.split(",")) if k is not None]
        fn = getattr(lib, name)
        assert fn.argtypes is not None, f"{name}: no argtypes"
        got = [kind_of_ctypes(t) for t in fn.argtypes]
        assert got == want, f"{name}: header {want}, ctypes {got}"
