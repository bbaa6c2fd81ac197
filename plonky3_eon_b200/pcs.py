"""GpuKzgPcs — host mirror of `KzgPcs` as a `p3_commit::Pcs<Fr, Challenger>`
(kzg/src/pcs.rs:143-335, trait commit/src/pcs.rs:21-187) over the C ABI.

Same method names, argument nesting, result nesting and failure behaviour as the reference:
  commit(iter of (domain, evals))            -> (KzgCommitment, ProverData)      pcs.rs:223-265
  commit_quotient(domain, evals, num_chunks) -> same (trait default)             commit/src/pcs.rs:82-102
  get_evaluations_on_domain(pd, idx, domain) -> RowMajorMatrix<Fr>               pcs.rs:267-287
  open(rounds, challenger)                   -> (OpenedValues, KzgProof)         pcs.rs:289-335
`verify` (pairings) stays on the CPU/Rust side and is out of scope.

Data: Fr matrices are numpy uint64 [h, w, 4]; G1 points are uint64 [8] affine Montgomery wire
points (identity = zeros); opening points are canonical ints or 4-limb wire arrays.
Prover-side failures raise (the reference panics): AssertionError for a height/domain mismatch
(pcs.rs:233-237), DegreeTooLarge for a short SRS (pcs.rs:238-240).
"""
import ctypes as C
import weakref
from dataclasses import dataclass, field as dc_field

import numpy as np

from . import field
from .dft import _as_matrix, _shift_wire
from .lib import default_context, pinned_empty


@dataclass(frozen=True)
class TwoAdicMultiplicativeCoset:
    """field/src/coset.rs:55-90: the coset shift * <omega_{2^log_size}> (shift: canonical int)."""
    shift: int
    log_size: int

    def __post_init__(self):
        if self.shift % field.P == 0 or self.log_size > field.TWO_ADICITY:
            raise ValueError("invalid coset")  # TwoAdicMultiplicativeCoset::new returns None

    def size(self):
        return 1 << self.log_size

    def subgroup_generator(self):
        return field.two_adic_generator(self.log_size)

    def first_point(self):
        return self.shift

    def next_point(self, x):
        """commit/src/domain.rs:144-146."""
        return x * self.subgroup_generator() % field.P

    def create_disjoint_domain(self, min_size):
        """commit/src/domain.rs:155-168: shift * GENERATOR, size 2^ceil(log2(min_size))."""
        log = max(min_size - 1, 0).bit_length()
        return TwoAdicMultiplicativeCoset(self.shift * field.GENERATOR % field.P, log)

    def split_domains(self, num_chunks):
        """commit/src/domain.rs:174-186."""
        lc = field.log2_strict(num_chunks)
        g = self.subgroup_generator()
        return [TwoAdicMultiplicativeCoset(self.shift * pow(g, i, field.P) % field.P, self.log_size - lc)
                for i in range(num_chunks)]

    def split_evals(self, num_chunks, evals):
        """commit/src/domain.rs:188-221: row r -> chunk r mod num_chunks."""
        a = _as_matrix(evals)
        assert a.shape[0] == self.size()
        return [np.ascontiguousarray(a[i::num_chunks]) for i in range(num_chunks)]


def release_handle(ctx, handle):
    """Finalizer of the prover-data classes: give a device coefficient buffer back, unless its context is gone."""
    try:
        if handle and getattr(ctx, "h", None) is not None and ctx.h.value:
            ctx.call("eon_handle_free", C.c_uint64(handle))
    except Exception:
        pass


@dataclass
class MatrixProverData:
    """kzg/src/pcs.rs:52-61 — evals stay on the host (the prover reads them back for the
    same-domain fast path); coefficients stay on the device behind `handle`."""
    domain: TwoAdicMultiplicativeCoset
    evals: np.ndarray
    handle: int
    _ctx: object = dc_field(default=None, repr=False)
    # (log_size, shift, evaluations) produced ahead of time by commit() under an LDE hint
    lde: tuple = dc_field(default=None, repr=False)

    def __post_init__(self):
        # the reference's ProverData is dropped with its scope; here the coefficients live in HBM behind `handle`,
        # so the handle goes back to the context when this object is collected (free() is the eager path)
        self._fin = weakref.finalize(self, release_handle, self._ctx, self.handle) if self.handle and self._ctx else None

    def coeffs(self):
        h, w = self.evals.shape[0], self.evals.shape[1]
        out = np.empty((h, w, 4), dtype=np.uint64)
        self._ctx.call("eon_kzg_read_coeffs", C.c_uint64(self.handle), out)
        return out

    def free(self):
        if self.handle:
            if self._fin is not None:
                self._fin.detach()
            self._ctx.call("eon_handle_free", C.c_uint64(self.handle))
            self.handle = 0


def observe_commitment(commitment, ctx=None, enc=0):
    """What `CanObserve<KzgCommitment> for DuplexChallenger` feeds the sponge (kzg/src/pcs.rs:409-438): every
    column commitment of every matrix -> its 32 compressed bytes (G1::to_bytes) -> four little-endian u64 ->
    `Fr::from_u64`.  Returns the canonical integers in observation order (a Fr challenger absorbs them
    as-is; the permutation itself is out of scope).  `commitment` = list of uint64 [w, 8]
    arrays as returned by GpuKzgPcs.commit."""
    ctx = ctx or default_context()
    out = []
    for cols in commitment:
        b = ctx.g1_to_bytes(cols, enc)
        out.extend(int(v) for v in b.view("<u8").reshape(-1))
    return out


class GpuKzgPcs:
    ZK = False  # pcs.rs:216

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or default_context(device)
        self.lde_hint = None

    def with_lde_hint(self, added_bits, shift=field.GENERATOR):
        """The prover evaluates every committed trace on the quotient coset right after committing it
        (eon-uni-stark/src/prover.rs:186-187 then :307-322); its size (trace height << added_bits) and
        shift (Fr::GENERATOR, commit/src/domain.rs:167) are known before the commit.  With the hint,
        commit() produces that matrix in the same call (eon_kzg_commit_lde: the download hides under the
        MSM) and get_evaluations_on_domain() hands it out."""
        self.lde_hint = (int(added_bits), int(shift) % field.P)
        return self

    # -- constructors (pcs.rs:170-203) -------------------------------------------------------
    @classmethod
    def from_srs(cls, g1_powers_wire, ctx=None, device=0):
        """g1_powers_wire: uint64 [n, 8] affine points (normalised once, not per MSM as in
        bn254/src/curve.rs:170)."""
        self = cls(ctx, device)
        a = np.ascontiguousarray(g1_powers_wire, dtype=np.uint64).reshape(-1, 8)
        self.ctx.call("eon_srs_load_affine", a, a.shape[0])
        return self

    @classmethod
    def from_srs_bytes(cls, g1_powers_bytes, ctx=None, device=0, enc=0):
        """A deserialised StructuredReferenceString (serde, kzg/src/params.rs:56): g1_powers as uint8 [n, 32]
        compressed points (Serialize for G1, bn254/src/curve.rs:84-88).  Decompression (one Fq square root per
        point) runs on the device; raises InvalidG1Point like serde's "Invalid G1 point"."""
        self = cls(ctx, device)
        self.ctx.srs_load_compressed(g1_powers_bytes, enc)
        return self

    def g1_powers_bytes(self, first=0, n=None, enc=0):
        """Serialize for G1 over g1_powers[first : first + n]."""
        return self.ctx.g1_to_bytes(self.g1_powers(first, n), enc)

    @classmethod
    def new(cls, max_degree, alpha, ctx=None, device=0):
        """KzgPcs::new -> init_srs_unsafe(max_degree, alpha) (params.rs:123-139), on the device."""
        self = cls(ctx, device)
        self.ctx.call("eon_srs_generate_unsafe", _shift_wire(alpha), int(max_degree) + 1)
        return self

    @property
    def max_degree(self):
        return self.ctx.srs_size() - 1

    def g1_powers(self, first=0, n=None):
        n = self.ctx.srs_size() - first if n is None else n
        out = np.empty((n, 8), dtype=np.uint64)
        self.ctx.call("eon_srs_read", first, n, out)
        return out

    # -- Pcs ---------------------------------------------------------------------------------
    def natural_domain_for_degree(self, degree):
        """pcs.rs:218-221."""
        npow = 1
        while npow < degree:
            npow <<= 1
        return TwoAdicMultiplicativeCoset(1, field.log2_strict(npow))

    def commit(self, evaluations, _use_hint=True):
        """pcs.rs:223-265.  Returns (commitment, prover_data): commitment[m] is a uint64 [w, 8]
        array (MatrixCommitment.columns), prover_data[m] a MatrixProverData."""
        commitments, prover = [], []
        for domain, evals in evaluations:
            a = _as_matrix(evals)
            h, w = a.shape[0], a.shape[1]
            assert h == domain.size(), "evaluation height must match domain size"
            cols = np.zeros((w, 8), dtype=np.uint64)
            handle = C.c_uint64(0)
            lde = None
            if _use_hint and self.lde_hint is not None and w > 0:
                added, lshift = self.lde_hint
                out = pinned_empty((h << added, w, 4))
                self.ctx.call("eon_kzg_commit_lde", a, domain.log_size, w, field.to_wire(domain.shift), cols,
                              C.byref(handle), domain.log_size + added, field.to_wire(lshift), out)
                lde = (domain.log_size + added, lshift, out)
            else:
                self.ctx.call("eon_kzg_commit", a, domain.log_size, w, field.to_wire(domain.shift),
                              cols, C.byref(handle))
            commitments.append(cols)
            prover.append(MatrixProverData(domain, a, int(handle.value), self.ctx, lde))
        return commitments, prover

    def commit_quotient(self, quotient_domain, quotient_evaluations, num_chunks):
        """commit/src/pcs.rs:82-102 (trait default)."""
        doms = quotient_domain.split_domains(num_chunks)
        a = _as_matrix(quotient_evaluations)
        assert a.shape[0] == quotient_domain.size(), "evaluation height must match domain size"
        w = a.shape[1]
        # one call: the chunks are pitched views of the uploaded matrix (no split_evals copy on the host,
        # domain.rs:188-221) and one batched MSM commits all of their columns (eon_kzg_commit_quotient)
        cols = np.zeros((num_chunks, w, 8), dtype=np.uint64)
        handles = np.zeros(num_chunks, dtype=np.uint64)
        self.ctx.call("eon_kzg_commit_quotient", a, quotient_domain.log_size, w, field.log2_strict(num_chunks),
                      field.to_wire(quotient_domain.shift), cols, handles)
        # MatrixProverData.evals of chunk i = rows i, i + num_chunks, ... (a strided view, not a copy)
        prover = [MatrixProverData(doms[i], a[i::num_chunks], int(handles[i]), self.ctx, None)
                  for i in range(num_chunks)]
        return [cols[i] for i in range(num_chunks)], prover

    def get_evaluations_on_domain(self, prover_data, idx, domain):
        """pcs.rs:267-287; the quadratic Horner loop of the reference is replaced by
        zero-pad + coset NTT from the device-resident coefficients (bit-identical)."""
        m = prover_data[idx]
        if m.domain.shift == domain.shift and m.domain.size() == domain.size():
            return m.evals.copy()
        if m.lde is not None and m.lde[0] == domain.log_size and m.lde[1] == domain.shift % field.P:
            return m.lde[2]
        w = m.evals.shape[1]
        out = pinned_empty((domain.size(), w, 4))
        self.ctx.call("eon_kzg_evals_on_coset", C.c_uint64(m.handle), domain.log_size,
                      field.to_wire(domain.shift), out)
        return out

    def open(self, commitment_data_with_opening_points, challenger=None):
        """pcs.rs:289-335.  Input: list of (prover_data, points_per_matrix).  Returns
        (opened_values[round][matrix][point] = uint64 [w, 4],
         proof[round][matrix][point]        = uint64 [w, 8] witnesses)."""
        # every (round, matrix, point, column) of the call goes through ONE eon_kzg_open_batch: all quotients
        # side by side, one batched MSM for all witnesses (the reference commits them one column at a time)
        mats, counts, pts = [], [], []
        for prover_data, points_per_matrix in commitment_data_with_opening_points:
            assert len(prover_data) == len(points_per_matrix)
            for m, points in zip(prover_data, points_per_matrix):
                mats.append(m)
                counts.append(len(points))
                pts.extend(_shift_wire(z) for z in points)
        widths = [m.evals.shape[1] for m in mats]
        total = sum(c * w for c, w in zip(counts, widths))
        vals = np.zeros((max(total, 1), 4), dtype=np.uint64)
        wits = np.zeros((max(total, 1), 8), dtype=np.uint64)
        if mats:
            handles = np.array([m.handle for m in mats], dtype=np.uint64)
            npts = np.array(counts, dtype=np.uint64)
            parr = np.ascontiguousarray(np.array(pts, dtype=np.uint64).reshape(-1, 4)) if pts else np.zeros((1, 4), np.uint64)
            self.ctx.call("eon_kzg_open_batch", len(mats), handles, npts, parr, vals, wits)
        opened_values, rounds, k, i = [], [], 0, 0
        for prover_data, _ in commitment_data_with_opening_points:
            mv, mp = [], []
            for _m in prover_data:
                w, c = widths[i], counts[i]
                mv.append([vals[k + p * w:k + (p + 1) * w] for p in range(c)])
                mp.append([wits[k + p * w:k + (p + 1) * w] for p in range(c)])
                k += c * w
                i += 1
            opened_values.append(mv)
            rounds.append(mp)
        return opened_values, rounds

    def open_matrix(self, m, points):
        """One matrix at `points` through eon_kzg_open (what open() did before the batched call; kept for the
        parity tests of the two entry points against each other)."""
        w = m.evals.shape[1]
        npts = len(points)
        pts = np.zeros((max(npts, 1), 4), dtype=np.uint64)
        for i, z in enumerate(points):
            pts[i] = _shift_wire(z)
        vals = np.zeros((npts, w, 4), dtype=np.uint64)
        wits = np.zeros((npts, w, 8), dtype=np.uint64)
        self.ctx.call("eon_kzg_open", C.c_uint64(m.handle), pts, npts, vals, wits)
        return [vals[i] for i in range(npts)], [wits[i] for i in range(npts)]

    def verify(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("verification (pairings) stays on the CPU side: kzg/src/pcs.rs:337-401")

    # -- bn254 G1::multi_exp (bn254/src/curve.rs:158-180) --------------------------------------
    def multi_exp(self, points_wire, scalars_wire):
        p = np.ascontiguousarray(points_wire, dtype=np.uint64).reshape(-1, 8)
        s = np.ascontiguousarray(scalars_wire, dtype=np.uint64).reshape(-1, 4)
        assert p.shape[0] == s.shape[0], "points and scalars must have the same length"
        out = np.zeros(8, dtype=np.uint64)
        self.ctx.call("eon_msm_points", p, s, p.shape[0], out)
        return out

    def commit_column(self, coeffs_wire):
        """kzg/src/util.rs:37-40 on the resident SRS."""
        s = np.ascontiguousarray(coeffs_wire, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(8, dtype=np.uint64)
        self.ctx.call("eon_msm_srs", s, s.shape[0], 1, 1, out)
        return out
