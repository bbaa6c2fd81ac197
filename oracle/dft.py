"""TwoAdicSubgroupDft semantics — big-int restatement (TEST INFRASTRUCTURE ONLY).

Matrices are lists of rows of canonical ints (row-major, like RowMajorMatrix,
matrix/src/dense.rs:24-37).  Every function follows the *trait default* of
dft/src/traits.rs (which is what defines the behaviour every Dft impl must
reproduce) on top of NaiveDft (dft/src/naive.rs:15-31).  `fast=True` swaps the
O(n^2) NaiveDft for an O(n log n) radix-2 evaluation that is checked against the
naive one in tests/test_oracle.py.
"""
from . import fr

P = fr.P


def log2_strict(n):
    """util/src/lib.rs:39 — panics (here: ValueError) on non powers of two."""
    if n <= 0 or n & (n - 1):
        raise ValueError(f"Not a power of two: {n}")
    return n.bit_length() - 1


def naive_dft_batch(mat):
    """dft/src/naive.rs:15-31: res[r][c] = sum_s g^(r*s) * mat[s][c]."""
    h = len(mat)
    if h == 0:
        return []
    w = len(mat[0])
    g = fr.two_adic_generator(log2_strict(h))
    res = [[0] * w for _ in range(h)]
    point = 1
    for r in range(h):
        pp = 1
        for s in range(h):
            row = mat[s]
            for c in range(w):
                res[r][c] = (res[r][c] + pp * row[c]) % P
            pp = pp * point % P
        point = point * g % P
    return res


def _ntt_col(col, omega):
    n = len(col)
    if n == 1:
        return col[:]
    e = _ntt_col(col[0::2], omega * omega % P)
    o = _ntt_col(col[1::2], omega * omega % P)
    out = [0] * n
    t = 1
    half = n // 2
    for k in range(half):
        x = t * o[k] % P
        out[k] = (e[k] + x) % P
        out[k + half] = (e[k] - x) % P
        t = t * omega % P
    return out


def fast_dft_batch(mat):
    h = len(mat)
    if h == 0:
        return []
    w = len(mat[0])
    g = fr.two_adic_generator(log2_strict(h))
    cols = [_ntt_col([mat[r][c] for r in range(h)], g) for c in range(w)]
    return [[cols[c][r] for c in range(w)] for r in range(h)]


def dft_batch(mat, fast=False):
    return fast_dft_batch(mat) if fast else naive_dft_batch(mat)


def coset_shift_cols(mat, shift):
    """dft/src/util.rs:28-36: row i *= shift^i."""
    out = []
    wgt = 1
    for row in mat:
        out.append([v * wgt % P for v in row])
        wgt = wgt * shift % P
    return out


def divide_by_height(mat):
    """dft/src/util.rs:15-25."""
    h = len(mat)
    log2_strict(h)
    hinv = fr.inv(h % P)
    return [[v * hinv % P for v in row] for row in mat]


def coset_dft_batch(mat, shift, fast=False):
    """traits.rs:83-91."""
    return dft_batch(coset_shift_cols(mat, shift), fast)


def idft_batch(mat, fast=False):
    """traits.rs:111-122: dft, divide by height, swap rows i <-> h-i."""
    d = divide_by_height(dft_batch(mat, fast))
    h = len(d)
    for row in range(1, h // 2):
        d[row], d[h - row] = d[h - row], d[row]
    return d


def coset_idft_batch(mat, shift, fast=False):
    """traits.rs:144-153."""
    return coset_shift_cols(idft_batch(mat, fast), fr.inv(shift))


def coset_lde_batch(mat, added_bits, shift, fast=False):
    """traits.rs:226-249: idft, zero-pad to h<<added_bits rows, coset_dft(shift)."""
    coeffs = idft_batch(mat, fast)
    h = len(coeffs)
    w = len(coeffs[0]) if h else 0
    coeffs = coeffs + [[0] * w for _ in range((h << added_bits) - h)]
    return coset_dft_batch(coeffs, shift, fast)


def lde_batch(mat, added_bits, fast=False):
    """traits.rs:187-192."""
    return coset_lde_batch(mat, added_bits, 1, fast)


# --- wire helpers -------------------------------------------------------------
def mat_to_wire(mat):
    """list-of-rows of canonical ints -> np.uint64 [h, w, 4] Montgomery limbs."""
    import numpy as np
    h = len(mat)
    w = len(mat[0]) if h else 0
    flat = [v for row in mat for v in row]
    return fr.to_wire(flat).reshape(h, w, 4) if h * w else np.zeros((h, w, 4), dtype=np.uint64)


def mat_from_wire(arr):
    import numpy as np
    a = np.asarray(arr, dtype=np.uint64)
    h, w = a.shape[0], a.shape[1]
    flat = fr.from_wire(a.reshape(-1, 4)) if h * w else []
    return [flat[r * w:(r + 1) * w] for r in range(h)]
