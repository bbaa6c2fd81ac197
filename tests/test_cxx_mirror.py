"""The C++ host mirror (include/eon_kzg.hpp: GpuDft = TwoAdicSubgroupDft<Fr>, GpuKzgPcs = KzgPcs as Pcs<Fr, _>,
RowMajorMatrix, TwoAdicMultiplicativeCoset, G1::multi_exp) built with g++ against libeon_kzg.so.

CPU: the header and the C++ twin of the reference's tests compile and link against the C ABI, and the binary
refuses to run without a GPU (no CPU fallback).  GPU: the twin (tests/host/cxx_mirror_test.cpp — multi_exp KATs,
NaiveDft::basic, DFT round trips, pcs_roundtrip, degree / height guards, commit_quotient) passes on the device."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "cxx_mirror_test.cpp")
BIN = os.path.join(HERE, "host", "cxx_mirror_test")
PKG = os.path.join(ROOT, "plonky3_eon_b200")


def build():
    deps = [SRC, os.path.join(ROOT, "include", "eon_kzg.hpp"), os.path.join(ROOT, "include", "eon_kzg.h"),
            os.path.join(PKG, "libeon_kzg.so")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas", "-Werror", "-o", BIN, SRC,
                               "-L" + PKG, "-leon_kzg", "-Wl,-rpath," + PKG])
    return BIN


def test_cxx_mirror_builds_and_has_no_cpu_fallback():
    import torch
    exe = build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 2 and "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_cxx_mirror_on_gpu():
    out = subprocess.run([build()], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
