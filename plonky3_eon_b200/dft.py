"""GpuDft — host mirror of `p3_dft::TwoAdicSubgroupDft<Fr>` (dft/src/traits.rs:27-507).

Same method names, argument meaning and panics as the trait; matrices are numpy uint64 arrays
of shape [height, width, 4] (RowMajorMatrix<Fr> bytes, matrix/src/dense.rs:24-37), shifts are
canonical Python ints or 4-limb wire arrays.  Every method is one C-ABI call
(include/eon_kzg.h); the `*_algebra_*` forms with D = 1 (Challenge = Fr) are aliases.
"""
import numpy as np

from . import field
from .lib import default_context


def _as_matrix(mat):
    a = np.ascontiguousarray(mat, dtype=np.uint64)
    if a.ndim == 2 and a.shape[1] == 4:      # a single column given as [h, 4]
        a = a.reshape(a.shape[0], 1, 4)
    assert a.ndim == 3 and a.shape[2] == 4, "matrix must have shape [height, width, 4]"
    return a


def _shift_wire(shift):
    if isinstance(shift, (int, np.integer)):
        return field.to_wire(int(shift))
    return np.ascontiguousarray(shift, dtype=np.uint64).reshape(4)


class GpuDft:
    """Clone + Default like the trait requires: all instances share the per-device context
    (twiddle caches live there, cf. Radix2Dit's cache dft/src/radix_2_dit.rs:33-58)."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or default_context(device)

    # -- required method -------------------------------------------------------------------
    def dft_batch(self, mat):
        """traits.rs:61 — evaluate each column (coefficients) on <omega_h>."""
        a = _as_matrix(mat)
        h, w = a.shape[0], a.shape[1]
        log_h = field.log2_strict(h)
        out = np.empty_like(a)
        self.ctx.call("eon_dft_batch", a, out, log_h, w)
        return out

    def coset_dft_batch(self, mat, shift):
        """traits.rs:83-91."""
        a = _as_matrix(mat)
        log_h = field.log2_strict(a.shape[0])
        out = np.empty_like(a)
        self.ctx.call("eon_coset_dft_batch", a, out, log_h, a.shape[1], _shift_wire(shift))
        return out

    def idft_batch(self, mat):
        """traits.rs:111-122."""
        a = _as_matrix(mat)
        log_h = field.log2_strict(a.shape[0])
        out = np.empty_like(a)
        self.ctx.call("eon_idft_batch", a, out, log_h, a.shape[1])
        return out

    def coset_idft_batch(self, mat, shift):
        """traits.rs:144-153."""
        a = _as_matrix(mat)
        log_h = field.log2_strict(a.shape[0])
        out = np.empty_like(a)
        self.ctx.call("eon_coset_idft_batch", a, out, log_h, a.shape[1], _shift_wire(shift))
        return out

    def lde_batch(self, mat, added_bits):
        """traits.rs:187-192."""
        return self.coset_lde_batch(mat, added_bits, 1)

    def coset_lde_batch(self, mat, added_bits, shift):
        """traits.rs:226-249."""
        a = _as_matrix(mat)
        h, w = a.shape[0], a.shape[1]
        log_h = field.log2_strict(h)
        out = np.empty((h << added_bits, w, 4), dtype=np.uint64)
        self.ctx.call("eon_coset_lde_batch", a, out, log_h, w, int(added_bits), _shift_wire(shift))
        return out

    # -- single-vector conveniences (traits.rs:41-45,70-75,99-101,131-134,172-176,206-210) ----
    def dft(self, vec):
        return self.dft_batch(_as_matrix(vec))[:, 0, :]

    def coset_dft(self, vec, shift):
        return self.coset_dft_batch(_as_matrix(vec), shift)[:, 0, :]

    def idft(self, vec):
        return self.idft_batch(_as_matrix(vec))[:, 0, :]

    def coset_idft(self, vec, shift):
        return self.coset_idft_batch(_as_matrix(vec), shift)[:, 0, :]

    def lde(self, vec, added_bits):
        return self.lde_batch(_as_matrix(vec), added_bits)[:, 0, :]

    def coset_lde(self, vec, added_bits, shift):
        return self.coset_lde_batch(_as_matrix(vec), added_bits, shift)[:, 0, :]

    # -- algebra forms (traits.rs:269-507): Challenge = Fr => DIMENSION = 1 => identical -------
    dft_algebra_batch = dft_batch
    coset_dft_algebra_batch = coset_dft_batch
    idft_algebra_batch = idft_batch
    coset_idft_algebra_batch = coset_idft_batch
    lde_algebra_batch = lde_batch
    coset_lde_algebra_batch = coset_lde_batch
