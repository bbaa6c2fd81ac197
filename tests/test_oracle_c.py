"""Pin the C restatement (oracle/c/oracle.c) against the big-int oracle and the committed goldens."""
import os

import numpy as np
import pytest

from oracle import cport, dft, fr, g1, kzg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kzg_small.npz")
P = fr.P


def test_field_ops():
    rng = np.random.default_rng(3)
    for which, mod in ((0, fr.P), (1, g1.Q)):
        vals = [int.from_bytes(rng.bytes(40), "little") % mod for _ in range(200)] + [0, 1, mod - 1]
        a = np.array([[(v >> (64 * k)) & (2**64 - 1) for k in range(4)] for v in vals], dtype=np.uint64)
        b = a[::-1].copy()
        Rinv = pow(1 << 256, -1, mod)
        ia = [sum(int(r[k]) << (64 * k) for k in range(4)) for r in a]
        ib = ia[::-1]
        for op, fn in ((0, lambda x, y: x * y * Rinv % mod), (4, lambda x, y: x * y * Rinv % mod),
                       (1, lambda x, y: (x + y) % mod),
                       (2, lambda x, y: (x - y) % mod)):
            r = cport.field_op(which, op, a, b)
            for i in range(len(ia)):
                assert sum(int(r[i, k]) << (64 * k) for k in range(4)) == fn(ia[i], ib[i])
        r = cport.field_op(which, 3, a[:20], a[:20])
        for i in range(20):
            got = sum(int(r[i, k]) << (64 * k) for k in range(4))
            assert got * ia[i] * Rinv % mod == (1 << 256) % mod


def test_golden_dft_family():
    g = np.load(GOLD)
    x = g["dft_in"]
    assert np.array_equal(cport.dft_batch(x), g["dft_out"])
    assert np.array_equal(cport.coset_dft_batch(x, fr.GENERATOR), g["coset_dft_out"])
    assert np.array_equal(cport.idft_batch(x), g["idft_out"])
    assert np.array_equal(cport.coset_idft_batch(x, fr.GENERATOR), g["coset_idft_out"])
    assert np.array_equal(cport.coset_lde_batch(x, 1, fr.GENERATOR), g["coset_lde1_out"])
    assert np.array_equal(cport.coset_lde_batch(x, 2, 1), g["lde2_out"])


@pytest.mark.parametrize("log_h,w", [(0, 2), (1, 3), (9, 3), (12, 2), (13, 16)])
def test_dft_vs_python(log_h, w):
    rng = np.random.default_rng(log_h)
    h = 1 << log_h
    xw = fr.random_wire(rng, h * w).reshape(h, w, 4)
    x = dft.mat_from_wire(xw)
    assert dft.mat_from_wire(cport.coset_dft_batch(xw, 7)) == dft.coset_dft_batch(x, 7, fast=True)
    assert dft.mat_from_wire(cport.coset_idft_batch(xw, 7)) == dft.coset_idft_batch(x, 7, fast=True)
    if log_h <= 9:
        assert dft.mat_from_wire(cport.coset_lde_batch(xw, 1, 5)) == dft.coset_lde_batch(x, 1, 5, fast=True)


def test_golden_kzg():
    g = np.load(GOLD)
    commits, coeffs = cport.kzg_commit(g["kzg_evals"], 1, g["srs"])
    assert np.array_equal(coeffs, g["kzg_coeffs"])
    assert np.array_equal(commits, g["kzg_commit"])
    commits2, _ = cport.kzg_commit(g["kzg_evals"], g["kzg_shift"][0], g["srs"])
    assert np.array_equal(commits2, g["kzg_commit_shifted"])
    for p in range(2):
        for c in range(2):
            q, v = cport.quotient_and_eval(coeffs, c, g["kzg_points"][p])
            assert np.array_equal(v, g["kzg_opened"][p, c])
            wit = cport.msm(g["srs"][:7], q)
            assert np.array_equal(wit[0], g["kzg_witness"][p, c])
    assert np.array_equal(cport.srs_generate(12345, 16), g["srs"])


def test_msm_kats_and_dlog():
    G = g1.G
    assert g1.from_wire(cport.msm(np.zeros((0, 8), np.uint64), np.zeros((0, 4), np.uint64)))[0] is None
    assert g1.from_wire(cport.msm(g1.to_wire([G, G]), fr.to_wire([2, 3])))[0] == g1.mul(G, 5)
    assert g1.from_wire(cport.msm(g1.to_wire([g1.mul(G, 7), g1.mul(G, 11)]), fr.to_wire([3, 5])))[0] == g1.mul(G, 76)
    assert g1.from_wire(cport.msm(g1.to_wire([G, g1.neg(G)]), fr.to_wire([9, 9])))[0] is None
    rng = np.random.default_rng(8)
    n, ncols = 3000, 2
    alpha = 424242
    srs = cport.srs_generate(alpha, n)
    assert g1.from_wire(srs[:3]) == kzg.init_srs_unsafe(2, alpha)
    sc = fr.random_wire(rng, n * ncols).reshape(n, ncols, 4)
    out = cport.msm(srs, sc, ncols=ncols)
    dl = kzg.srs_dlogs(n - 1, alpha)
    for c in range(ncols):
        assert g1.from_wire(out[c:c + 1])[0] == g1.msm_via_dlog(dl, fr.from_wire(sc[:, c, :]))
