"""GPU parity: TwoAdicSubgroupDft<Fr> through the C ABI vs the oracle.

Mirrors the reference's DFT conformance template (field-testing/src/dft_testing.rs:9-305):
every method vs NaiveDft on h = 1..16, w = 3, shift = GENERATOR; round-trips and cross-checks at
larger sizes; plus the KATs of dft/src/naive.rs:49-85 and dft/src/util.rs:49-138.
Bit-exact: outputs are compared limb for limb (integer work, tolerance 0).
"""
import os

import numpy as np
import pytest

from oracle import dft as odft
from oracle import fr, kzg as okzg

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kzg_small.npz")
P = fr.P


@pytest.fixture(scope="module")
def gdft():
    from plonky3_eon_b200 import GpuDft
    return GpuDft()


def rand_mat(seed, h, w):
    rng = np.random.default_rng(seed)
    return fr.random_wire(rng, h * w).reshape(h, w, 4)


def test_naive_basic_kat(gdft):
    # dft/src/naive.rs:49-85
    mat = odft.mat_to_wire([[5, 2, 0], [4, 3, 0]])
    out = gdft.dft_batch(mat)
    assert odft.mat_from_wire(out) == [[9, 5, 0], [1, P - 1, 0]]


def test_util_kats(gdft):
    # divide_by_height [2,4,6,8] -> [1,2,3,4] is idft's scale; check through idft(dft(x)) == x and
    # coset_shift_cols shift 2 on 3x2 through coset_dft == dft(shifted)  (dft/src/util.rs:49-138)
    x = odft.mat_to_wire([[1, 2], [3, 4], [5, 6], [7, 8]])
    assert np.array_equal(gdft.idft_batch(gdft.dft_batch(x)), x)
    shifted = odft.mat_to_wire(odft.coset_shift_cols([[1, 2], [3, 4], [5, 6], [7, 8]], 2))
    assert np.array_equal(gdft.coset_dft_batch(x, 2), gdft.dft_batch(shifted))


@pytest.mark.parametrize("log_h", range(0, 5))
def test_all_methods_vs_naive(gdft, log_h):
    # field-testing/src/dft_testing.rs:9-112: h = 1..16, width 3, shift = GENERATOR, seed 1
    h, w = 1 << log_h, 3
    xw = rand_mat(1, h, w)
    x = odft.mat_from_wire(xw)
    s = fr.GENERATOR
    assert odft.mat_from_wire(gdft.dft_batch(xw)) == odft.dft_batch(x)
    assert odft.mat_from_wire(gdft.coset_dft_batch(xw, s)) == odft.coset_dft_batch(x, s)
    assert odft.mat_from_wire(gdft.idft_batch(xw)) == odft.idft_batch(x)
    assert odft.mat_from_wire(gdft.coset_idft_batch(xw, s)) == odft.coset_idft_batch(x, s)
    assert odft.mat_from_wire(gdft.lde_batch(xw, 1)) == odft.lde_batch(x, 1)
    assert odft.mat_from_wire(gdft.coset_lde_batch(xw, 1, s)) == odft.coset_lde_batch(x, 1, s)
    assert odft.mat_from_wire(gdft.coset_lde_batch(xw, 2, s)) == odft.coset_lde_batch(x, 2, s)
    # single-vector forms
    col = np.ascontiguousarray(xw[:, 0, :])
    assert odft.mat_from_wire(gdft.dft(col).reshape(h, 1, 4)) == odft.dft_batch([[r[0]] for r in x])


def test_golden_fixture(gdft):
    g = np.load(GOLD)
    x = g["dft_in"]
    s = fr.GENERATOR
    assert np.array_equal(gdft.dft_batch(x), g["dft_out"])
    assert np.array_equal(gdft.coset_dft_batch(x, s), g["coset_dft_out"])
    assert np.array_equal(gdft.idft_batch(x), g["idft_out"])
    assert np.array_equal(gdft.coset_idft_batch(x, s), g["coset_idft_out"])
    assert np.array_equal(gdft.coset_lde_batch(x, 1, s), g["coset_lde1_out"])
    assert np.array_equal(gdft.lde_batch(x, 2), g["lde2_out"])


@pytest.mark.parametrize("log_h,w", [(5, 1), (6, 2), (7, 16), (8, 5), (9, 16), (10, 3), (11, 1), (12, 4), (12, 33)])
def test_mid_sizes_vs_fast_oracle(gdft, log_h, w):
    # multi-pass kernels (1, 2 and 3 HBM passes, ragged column tiles) vs the O(n log n) oracle
    h = 1 << log_h
    xw = rand_mat(100 + log_h * 64 + w, h, w)
    x = odft.mat_from_wire(xw)
    s = 7
    assert odft.mat_from_wire(gdft.coset_dft_batch(xw, s)) == odft.coset_dft_batch(x, s, fast=True)
    assert odft.mat_from_wire(gdft.coset_idft_batch(xw, s)) == odft.coset_idft_batch(x, s, fast=True)
    if log_h <= 10:
        assert odft.mat_from_wire(gdft.coset_lde_batch(xw, 1, fr.GENERATOR)) == \
            odft.coset_lde_batch(x, 1, fr.GENERATOR, fast=True)


def _check_rows_by_horner(colvals, out_w, shift, log_n, rows):
    """colvals: {col: coefficient list}; checks out_w[j, col] == poly_col(shift * omega^j)."""
    g = fr.two_adic_generator(log_n)
    for j in rows:
        pt = shift * pow(g, j, P) % P
        for c, coeffs in colvals.items():
            assert fr.from_wire(out_w[j, c])[0] == okzg.eval_poly(coeffs, pt), (j, c)


@pytest.mark.parametrize("log_h,w", [(14, 16), (15, 16), (16, 8), (17, 2), (18, 1)])
def test_large_roundtrip_and_spot_checks(gdft, log_h, w):
    # dft_testing.rs:260-305 (round trips at 2^14..2^17) + Horner spot checks of dft / lde rows
    h = 1 << log_h
    xw = rand_mat(7 + log_h, h, w)
    ev = gdft.dft_batch(xw)
    assert np.array_equal(gdft.idft_batch(ev), xw)
    s = fr.GENERATOR
    cev = gdft.coset_dft_batch(xw, s)
    assert np.array_equal(gdft.coset_idft_batch(cev, s), xw)
    rows = [0, 1, h // 2, h - 1, 12345 % h]
    colvals = {c: fr.from_wire(xw[:, c, :]) for c in sorted({0, w - 1})}
    _check_rows_by_horner(colvals, ev, 1, log_h, rows)
    _check_rows_by_horner(colvals, cev, s, log_h, rows)
    # coset LDE of the evaluations: rows are the same polynomials on s * <omega_2h>
    lde = gdft.coset_lde_batch(ev, 1, s)
    assert lde.shape == (2 * h, w, 4)
    _check_rows_by_horner(colvals, lde, s, log_h + 1, [0, 1, h - 1, h + 3, 2 * h - 1])
    # even rows of the blow-up on shift 1 are the original evaluations (lde consistency)
    lde1 = gdft.lde_batch(ev, 1)
    assert np.array_equal(lde1[0::2], ev)


def test_linearity_full_size_config2_shape(gdft):
    # config-2 shape (2^20 x 16) is too big for the oracle: size-independent properties instead.
    log_h, w = 20, 16
    h = 1 << log_h
    rng = np.random.default_rng(99)
    a = fr.random_wire(rng, h * w).reshape(h, w, 4)
    s = fr.GENERATOR
    lde = gdft.coset_lde_batch(a, 1, s)
    # (1) inverse: coset_idft over the size-2h coset gives the zero-padded coefficients
    co = gdft.coset_idft_batch(lde, s)
    assert not co[h:].any()
    assert np.array_equal(co[:h], gdft.idft_batch(a))
    # (2) spot-check rows against Horner on one column
    coeff_col = fr.from_wire(co[:h, 5, :])
    g = fr.two_adic_generator(log_h + 1)
    for j in (0, 1, h + 7, 2 * h - 1):
        pt = s * pow(g, j, P) % P
        assert fr.from_wire(lde[j, 5])[0] == okzg.eval_poly(coeff_col, pt)


def test_errors(gdft):
    with pytest.raises(ValueError):  # non power of two height: log2_strict panics in the reference
        gdft.dft_batch(np.zeros((3, 1, 4), dtype=np.uint64))
    from plonky3_eon_b200 import EonError
    with pytest.raises(EonError):    # zero shift is not a coset
        gdft.coset_dft_batch(np.zeros((4, 1, 4), dtype=np.uint64), 0)
    # zero width is a no-op
    assert gdft.dft_batch(np.zeros((4, 0, 4), dtype=np.uint64)).shape == (4, 0, 4)
