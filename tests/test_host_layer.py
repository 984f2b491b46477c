"""CPU tests of the host layer: the C-ABI library loads and exports every declared symbol, and the
product path refuses to run without a CUDA device (no fallback)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as graft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(zkl):
    graft.build()
    header = open(os.path.join(ROOT, "include", "zkl_b200.h")).read()
    declared = set(re.findall(r"\b(zkl_[a-z0-9_]+)\s*\(", header))
    declared -= {"zkl_batch_proof"}
    lib = ctypes.CDLL(zkl.capi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    assert declared == set(zkl.capi.EXPORTED_SYMBOLS)


def test_no_cpu_fallback(zkl):
    if _has_cuda():
        pytest.skip("CUDA device present")
    with pytest.raises(zkl.capi.ZklError) as e:
        zkl.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The product package must never reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "zk-lisp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f


def test_public_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/zkl_b200.h must compile as C99 (what cgo / bindgen / ctypes-style binders read)
    and as C++, with nothing but plain pointers and sizes in the signatures."""
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "zkl_b200.h"\nint main(void) { zkl_child_challenges c; (void)c; return (int)sizeof(zkl_fe) - 16; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
    header = open(os.path.join(inc, "zkl_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # comments may mention callers; declarations may not
    assert "torch" not in code.lower() and "std::" not in code and "at::" not in code
