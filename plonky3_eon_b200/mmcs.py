"""Host-side mirror of `KzgMmcs` (kzg/src/mmcs.rs:58-295, `impl Mmcs<Fr>`), prover half.

commit      every column of every input matrix is a polynomial in COEFFICIENT form (no iDFT) and is
            committed with commit_column (mmcs.rs:155-190) -> eon_kzg_commit_coeffs, one batched MSM
            per matrix; the matrix stays on the device behind the prover-data handle.
open_batch  row `index` (scaled to each matrix's height, mmcs.rs:192-237): value = f_c(local_index),
            witness = commit((f_c - value)/(X - local_index)) -> eon_kzg_open on the handle.
verify_batch (pairings, mmcs.rs:243-295) stays on the CPU side, like KzgPcs::verify.
"""
import ctypes as C
import weakref
from dataclasses import dataclass, field as dc_field

import numpy as np

from . import field
from .lib import default_context
from .pcs import _as_matrix, _shift_wire, release_handle


@dataclass
class KzgMmcsProverData:
    """mmcs.rs:88-94 — the committed matrices (host copy, as `get_matrices` hands them back) plus
    the device handles of the same coefficients."""
    matrices: list
    handles: list
    _ctx: object = dc_field(default=None, repr=False)

    def __post_init__(self):
        # dropped like the reference's ProverData: the device buffers go back when this object is collected
        self._fins = [weakref.finalize(self, release_handle, self._ctx, h) for h in self.handles if h and self._ctx]

    def free(self):
        for f in self._fins:
            f.detach()
        self._fins = []
        for h in self.handles:
            if h:
                self._ctx.call("eon_handle_free", C.c_uint64(h))
        self.handles = [0] * len(self.handles)


def _log2_ceil(n):
    """`n.next_power_of_two().trailing_zeros()` (mmcs.rs:203,209); next_power_of_two(0) == 1."""
    lg = 0
    while (1 << lg) < n:
        lg += 1
    return lg


class GpuKzgMmcs:
    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or default_context(device)

    @classmethod
    def from_srs(cls, g1_powers_wire, ctx=None, device=0):
        """mmcs.rs:128-130; points are affine wire points, normalised once."""
        self = cls(ctx, device)
        a = np.ascontiguousarray(g1_powers_wire, dtype=np.uint64).reshape(-1, 8)
        self.ctx.call("eon_srs_load_affine", a, a.shape[0])
        return self

    @classmethod
    def new(cls, max_degree, alpha, ctx=None, device=0):
        """mmcs.rs:150-153 -> init_srs_unsafe(max_degree, alpha), generated on the device."""
        self = cls(ctx, device)
        self.ctx.call("eon_srs_generate_unsafe", _shift_wire(alpha), int(max_degree) + 1)
        return self

    def commit(self, inputs):
        """mmcs.rs:174-190.  inputs: list of [h, w, 4] wire matrices (any h).  Returns
        (commitment: list of uint64 [w, 8] per matrix, KzgMmcsProverData)."""
        commitments, mats, handles = [], [], []
        for m in inputs:
            a = _as_matrix(m)
            h, w = a.shape[0], a.shape[1]
            cols = np.zeros((w, 8), dtype=np.uint64)
            handle = C.c_uint64(0)
            self.ctx.call("eon_kzg_commit_coeffs", a, h, w, cols, C.byref(handle))
            commitments.append(cols)
            mats.append(a)
            handles.append(int(handle.value))
        return commitments, KzgMmcsProverData(mats, handles, self.ctx)

    @staticmethod
    def local_index(index, height, log2_max_height):
        """mmcs.rs:209-214."""
        lg = _log2_ceil(height)
        li = index >> (log2_max_height - lg) if log2_max_height >= lg else index
        return li % height

    def open_batch(self, index, prover_data):
        """mmcs.rs:192-237.  Returns (opened_values[matrix] = uint64 [w, 4],
        witnesses[matrix] = uint64 [w, 8])."""
        max_height = max((m.shape[0] for m in prover_data.matrices), default=0)
        lmax = _log2_ceil(max_height)
        # every matrix at its own local index through ONE eon_kzg_open_batch (heights differ: the shorter
        # quotients are zero-filled below their length and share the single batched MSM)
        mats = prover_data.matrices
        if not mats:
            return [], []
        pts = np.zeros((len(mats), 4), dtype=np.uint64)
        for i, m in enumerate(mats):
            pts[i] = field.to_wire(self.local_index(index, m.shape[0], lmax))   # Fr::new(local_index as u64)
        total = sum(m.shape[1] for m in mats)
        vals = np.zeros((max(total, 1), 4), dtype=np.uint64)
        wits = np.zeros((max(total, 1), 8), dtype=np.uint64)
        self.ctx.call("eon_kzg_open_batch", len(mats), np.array(prover_data.handles, dtype=np.uint64),
                      np.ones(len(mats), dtype=np.uint64), pts, vals, wits)
        opened, witnesses, k = [], [], 0
        for m in mats:
            w = m.shape[1]
            opened.append(vals[k:k + w])
            witnesses.append(wits[k:k + w])
            k += w
        return opened, witnesses

    def get_matrices(self, prover_data):
        """mmcs.rs:239-241."""
        return list(prover_data.matrices)

    def verify_batch(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("verification (pairings) stays on the CPU side: kzg/src/mmcs.rs:243-295")
