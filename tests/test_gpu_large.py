"""Parity at the sizes BASELINE.json names (not only at sizes the Python oracle finishes in seconds):

* the standalone MSM of configs[2] at 2^24 points x 1 column (window bits 20, 13 table levels, no slice schedule —
  a code path the 2^20 x 16 commit never takes), against the discrete-log shortcut: the synthetic SRS is
  g1_powers[i] = alpha^i G (kzg/src/params.rs:123-139), so the MSM must equal p(alpha) G with p(alpha) from the C
  restatement's Horner loop (oracle/c, util.rs:63-68);
* the transform sizes of configs[4]: inverse NTT at 2^24 rows and the blow-up-4 coset LDE to 2^26 rows (4 HBM
  passes per transform), against Horner evaluations of the same polynomial at spot rows and an inverse round trip.
"""
import os
import sys

import numpy as np
import pytest

from oracle import cport, fr, g1

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALPHA = 12345


def uniform_fr(seed, n):
    """n uniform Fr as Montgomery limbs (vectorised numpy: fr.random_wire loops in Python)."""
    sys.path.insert(0, ROOT)
    import bench
    rng_backup = bench.synth_column
    return rng_backup(seed - 1000, n)


@pytest.fixture(scope="module")
def ctx():
    from plonky3_eon_b200 import Context
    c = Context(0)
    yield c
    c.close()


def test_msm_2p24_single_column_matches_dlog_shortcut(ctx):
    from plonky3_eon_b200 import GpuKzgPcs
    n = 1 << 24
    GpuKzgPcs.new(n - 1, ALPHA, ctx=ctx)                       # SRS + window tables on the device
    assert int(ctx.lib.eon_srs_window_bits(ctx.h)) == 20
    sc = uniform_fr(4242, n).reshape(n, 1, 4)
    sc[5:9] = 0                                                 # a few zero scalars
    out = np.zeros((1, 8), dtype=np.uint64)
    ctx.call("eon_msm_srs", sc, n, 1, 1, out)
    assert int(ctx.lib.eon_msm_window_bits_used(ctx.h)) == 20 and int(ctx.lib.eon_msm_rounds_used(ctx.h)) >= 1
    _, val = cport.quotient_and_eval(sc, 0, ALPHA)               # p(alpha), Horner in C
    assert g1.from_wire(out)[0] == g1.mul(g1.G, fr.from_wire(val.reshape(1, 4))[0])
    # the same points through an index-range shard with its own tables (what one of 8 GPUs computes) plus the rest
    first, cnt = n // 8 * 3, n // 8
    ctx.call("eon_srs_set_range_tables", first, cnt, 0)
    import ctypes as C
    d = ctx.dev_alloc(cnt * 32)
    ctx.h2d(d, np.ascontiguousarray(sc[first:first + cnt]))
    part = np.zeros((1, 8), dtype=np.uint64)
    ctx.call("eon_msm_srs_range_dev", C.c_void_p(d), first, cnt, 1, 1, part)
    assert int(ctx.lib.eon_msm_window_bits_used(ctx.h)) == 17      # sized for the 2^21-point shard
    ctx.dev_free(d)
    _, v = cport.quotient_and_eval(np.ascontiguousarray(sc[first:first + cnt]), 0, ALPHA)
    want = g1.mul(g1.G, fr.from_wire(v.reshape(1, 4))[0] * pow(ALPHA, first, fr.P) % fr.P)
    assert g1.from_wire(part)[0] == want
    ctx.call("eon_srs_generate_unsafe", fr.to_wire([ALPHA])[0], 1 << 10)   # give the 14 GiB of tables back


def test_lde_2p24_to_2p26_rows_matches_horner(ctx):
    from plonky3_eon_b200 import GpuDft, field
    log_h, w, added = 24, 2, 2
    h = 1 << log_h
    m = np.stack([uniform_fr(7000 + c, h) for c in range(w)], axis=1)      # [h, w, 4]
    dft = GpuDft(ctx)
    coeffs = dft.idft_batch(m)
    lde = dft.coset_lde_batch(m, added, fr.GENERATOR)
    assert lde.shape == (h << added, w, 4)
    assert np.array_equal(dft.dft_batch(coeffs), m)                         # inverse round trip, bit for bit
    w_h = field.two_adic_generator(log_h)
    w_l = field.two_adic_generator(log_h + added)
    n_l = h << added
    for col in range(w):
        col_coeffs = np.ascontiguousarray(coeffs[:, col:col + 1, :])
        # the coefficients really are the polynomial through the evaluations: p(omega^i) == m[i]
        for i in (1, h // 2 + 12345):
            _, v = cport.quotient_and_eval(col_coeffs, 0, pow(w_h, i, fr.P))
            assert np.array_equal(v, m[i, col])
        # LDE row j == p(5 * omega_(4h)^j): the Horner value of get_evaluations_on_domain (kzg/src/pcs.rs:278-286)
        for j in (0, 3, n_l // 2 + 777, n_l - 1):
            x = fr.GENERATOR * pow(w_l, j, fr.P) % fr.P
            _, v = cport.quotient_and_eval(col_coeffs, 0, x)
            assert np.array_equal(v, lde[j, col]), (col, j)
