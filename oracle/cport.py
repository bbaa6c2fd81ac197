"""ctypes view of oracle/c/liboracle.so — the C restatement of the reference hot path
(TEST / BASELINE INFRASTRUCTURE ONLY; see oracle/c/oracle.c for the reference file:line map).

Arrays are numpy uint64 wire arrays: Fr [.., 4] Montgomery limbs, G1 [.., 8] affine Montgomery.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_LIB = os.path.join(_DIR, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _DIR])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_DIR, "oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.oc_num_threads.restype = C.c_int
        _lib.oc_init()
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def num_threads():
    return int(lib().oc_num_threads())


def field_op(which, op, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    r = np.zeros_like(a)
    lib().oc_field_op(C.c_int(which), C.c_int(op), _p(a), _p(b), _p(r), C.c_size_t(a.shape[0]))
    return r


def _mat(m):
    a = np.array(m, dtype=np.uint64, order="C", copy=True)
    assert a.ndim == 3 and a.shape[2] == 4
    return a


def _log2(h):
    assert h > 0 and h & (h - 1) == 0, f"Not a power of two: {h}"
    return h.bit_length() - 1


def _fr(x):
    from . import fr
    if isinstance(x, (int, np.integer)):
        return fr.to_wire([int(x)])[0].copy()
    return np.ascontiguousarray(x, dtype=np.uint64).reshape(4)


def dft_batch(m):
    a = _mat(m)
    lib().oc_dft_batch(_p(a), C.c_uint(_log2(a.shape[0])), C.c_size_t(a.shape[1]))
    return a


def coset_dft_batch(m, shift):
    a = _mat(m)
    s = _fr(shift)
    lib().oc_coset_dft_batch(_p(a), C.c_uint(_log2(a.shape[0])), C.c_size_t(a.shape[1]), _p(s))
    return a


def idft_batch(m):
    a = _mat(m)
    lib().oc_idft_batch(_p(a), C.c_uint(_log2(a.shape[0])), C.c_size_t(a.shape[1]))
    return a


def coset_idft_batch(m, shift):
    a = _mat(m)
    s = _fr(shift)
    lib().oc_coset_idft_batch(_p(a), C.c_uint(_log2(a.shape[0])), C.c_size_t(a.shape[1]), _p(s))
    return a


def coset_lde_batch(m, added_bits, shift):
    a = np.ascontiguousarray(m, dtype=np.uint64)
    h, w = a.shape[0], a.shape[1]
    out = np.empty((h << added_bits, w, 4), dtype=np.uint64)
    s = _fr(shift)
    lib().oc_coset_lde_batch(_p(a), _p(out), C.c_uint(_log2(h)), C.c_size_t(w), C.c_uint(added_bits), _p(s))
    return out


def msm(points, scalars, ncols=1, ld=None):
    """out[c] = sum_i scalars[i, c] * points[i]; scalars [n, ld, 4] (or [n, 4] for one column)."""
    p = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8)
    s = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = p.shape[0]
    if s.ndim == 2:
        s = s.reshape(n, 1, 4)
    ld = s.shape[1] if ld is None else ld
    out = np.zeros((ncols, 8), dtype=np.uint64)
    lib().oc_msm(_p(p), _p(s), C.c_size_t(n), C.c_size_t(ncols), C.c_size_t(ld), _p(out))
    return out


def srs_generate(alpha, n):
    out = np.zeros((n, 8), dtype=np.uint64)
    a = _fr(alpha)
    lib().oc_srs_generate(_p(a), C.c_size_t(n), _p(out))
    return out


def quotient_and_eval(coeffs, col, z):
    a = np.ascontiguousarray(coeffs, dtype=np.uint64)
    h, w = a.shape[0], a.shape[1]
    q = np.zeros((max(h - 1, 0), 4), dtype=np.uint64)
    v = np.zeros(4, dtype=np.uint64)
    zz = _fr(z)
    lib().oc_quotient_and_eval(_p(a), C.c_size_t(h), C.c_size_t(w), C.c_size_t(col), _p(zz), _p(q), _p(v))
    return q, v


def kzg_commit(evals, shift, srs, ncols_msm=None):
    a = np.ascontiguousarray(evals, dtype=np.uint64)
    h, w = a.shape[0], a.shape[1]
    ncols_msm = w if ncols_msm is None else ncols_msm
    coeffs = np.empty_like(a)
    commits = np.zeros((ncols_msm, 8), dtype=np.uint64)
    s = _fr(shift)
    p = np.ascontiguousarray(srs, dtype=np.uint64).reshape(-1, 8)
    assert p.shape[0] >= h, "DegreeTooLarge"
    lib().oc_kzg_commit(_p(a), C.c_uint(_log2(h)), C.c_size_t(w), _p(s), _p(p), C.c_size_t(ncols_msm),
                        _p(coeffs), _p(commits))
    return commits, coeffs
