"""The C-ABI library loads on a GPU-less host and exports every symbol include/eon_kzg.h declares;
the product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "eon_kzg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eon_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_header():
    from plonky3_eon_b200 import lib
    handle = lib.load()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(handle, s), f"libeon_kzg.so does not export {s}"
    # and the ctypes prototype table covers the whole header
    assert set(syms) == set(lib.EXPORTED_SYMBOLS)
    assert b"sm_100a" in handle.eon_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from plonky3_eon_b200 import Context, EonError
    with pytest.raises(EonError):
        Context(0)
    raw = ctypes.CDLL(os.path.join(ROOT, "plonky3_eon_b200", "libeon_kzg.so"))
    out = ctypes.c_void_p()
    assert raw.eon_ctx_create(0, None, ctypes.byref(out)) < 0 and not out.value


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "plonky3_eon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import|from)\s+oracle|oracle[./]c|liboracle", src), f"{f} uses the oracle"
