"""GPU parity of the two batched Pcs entry points against the oracle and against the per-matrix calls:

* eon_kzg_commit_quotient = Pcs::commit_quotient (trait default, commit/src/pcs.rs:82-102; split_evals /
  split_domains, commit/src/domain.rs:174-221): chunks as pitched views, one MSM over all chunk columns;
* eon_kzg_open_batch = KzgPcs::open (kzg/src/pcs.rs:289-335) over every (round, matrix, point) of a call.

Bit-exact: affine G1 wire points and canonical Fr limbs are compared directly.
"""

import numpy as np
import pytest

from oracle import dft as odft
from oracle import fr, g1, kzg as okzg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from plonky3_eon_b200 import Context
    c = Context(0)
    yield c
    c.close()


def pcs_new(ctx, max_degree, alpha):
    from plonky3_eon_b200 import GpuKzgPcs
    return GpuKzgPcs.new(max_degree, alpha, ctx=ctx)


@pytest.mark.parametrize("log_size,width,num_chunks", [(4, 1, 2), (5, 3, 4), (3, 2, 1), (3, 1, 8), (7, 2, 2)])
def test_commit_quotient_vs_oracle(ctx, log_size, width, num_chunks):
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    alpha = 12345
    h = (1 << log_size) // num_chunks
    pcs = pcs_new(ctx, max(h - 1, 1), alpha)
    srs = okzg.init_srs_unsafe(max(h - 1, 1), alpha)
    rng = np.random.default_rng(1000 + log_size * 16 + width)
    qdom = TwoAdicMultiplicativeCoset(fr.GENERATOR, log_size)
    qw = fr.random_wire(rng, (1 << log_size) * width).reshape(1 << log_size, width, 4)
    q = odft.mat_from_wire(qw)
    commit, pdata = pcs.commit_quotient(qdom, qw, num_chunks)
    doms = okzg.split_domains((fr.GENERATOR, log_size), num_chunks)
    subs = okzg.split_evals(num_chunks, q)
    ocommit, opdata = okzg.commit(srs, list(zip(doms, subs)))
    assert len(commit) == num_chunks and len(pdata) == num_chunks
    for m in range(num_chunks):
        assert (pdata[m].domain.shift, pdata[m].domain.log_size) == doms[m]
        assert g1.from_wire(commit[m]) == ocommit[m]
        assert odft.mat_from_wire(pdata[m].coeffs()) == opdata[m]["coeffs"]
        assert odft.mat_from_wire(np.ascontiguousarray(pdata[m].evals)) == subs[m]
    # the chunks open like any committed matrix (prover.rs:416-442 opens every chunk at zeta)
    zeta = int.from_bytes(rng.bytes(31), "little")
    opened, proof = pcs.open([(pdata, [[zeta]] * num_chunks)])
    oopened, owits = okzg.open_(srs, [(opdata, [[zeta]] * num_chunks)])
    for m in range(num_chunks):
        assert fr.from_wire(opened[0][m][0]) == oopened[0][m][0]
        assert g1.from_wire(proof[0][m][0]) == owits[0][m][0]
    for m in pdata:
        m.free()


def test_commit_quotient_matches_per_chunk_commits_large(ctx):
    """2^15 x 2 quotient in 2 and 4 chunks: fused call == commit() of the host-split chunks (multi-pass NTTs,
    sorted MSM path), commitments and retained coefficients."""
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    log_size, width = 15, 2
    pcs = pcs_new(ctx, (1 << log_size) - 1, 777)
    rng = np.random.default_rng(5)
    qw = fr.random_wire(rng, (1 << log_size) * width).reshape(1 << log_size, width, 4)
    qdom = TwoAdicMultiplicativeCoset(fr.GENERATOR, log_size)
    for num_chunks in (2, 4):
        commit, pdata = pcs.commit_quotient(qdom, qw, num_chunks)
        subs = qdom.split_evals(num_chunks, qw)
        doms = qdom.split_domains(num_chunks)
        ref_commit, ref_pdata = pcs.commit(list(zip(doms, subs)), _use_hint=False)
        for m in range(num_chunks):
            assert np.array_equal(commit[m], ref_commit[m])
            assert np.array_equal(pdata[m].coeffs(), ref_pdata[m].coeffs())
        for m in pdata + ref_pdata:
            m.free()


def test_commit_quotient_errors(ctx):
    from plonky3_eon_b200 import DegreeTooLarge, EonError, TwoAdicMultiplicativeCoset
    pcs = pcs_new(ctx, 3, 5)                      # SRS of 4 points
    rng = np.random.default_rng(9)
    qw = fr.random_wire(rng, 16).reshape(16, 1, 4)
    qdom = TwoAdicMultiplicativeCoset(fr.GENERATOR, 4)
    with pytest.raises(DegreeTooLarge):            # chunks of 8 rows need 8 SRS points (pcs.rs:238-240)
        pcs.commit_quotient(qdom, qw, 2)
    with pytest.raises(Exception):                 # log2_strict_usize(3) panics (domain.rs:175)
        pcs.commit_quotient(qdom, qw, 3)
    handles = np.zeros(32, dtype=np.uint64)
    cols = np.zeros((32, 8), dtype=np.uint64)
    with pytest.raises(EonError):                  # more chunks than rows
        ctx.call("eon_kzg_commit_quotient", qw, 4, 1, 5, fr.to_wire([fr.GENERATOR])[0].copy(), cols, handles)
    commit, pdata = pcs.commit_quotient(qdom, qw, 4)   # 4 rows per chunk fits
    assert len(commit) == 4
    for m in pdata:
        m.free()


def test_open_batch_matches_per_matrix_open_and_oracle(ctx):
    """Two rounds, matrices of different heights and widths, different numbers of points: the batched call,
    the per-matrix eon_kzg_open and the oracle agree."""
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    alpha = 4242
    pcs = pcs_new(ctx, 63, alpha)
    srs = okzg.init_srs_unsafe(63, alpha)
    rng = np.random.default_rng(21)
    shapes = [(6, 3), (3, 2), (0, 1), (6, 1)]      # (log_h, width)
    wire = [fr.random_wire(rng, (1 << lh) * w).reshape(1 << lh, w, 4) for lh, w in shapes]
    doms = [TwoAdicMultiplicativeCoset(1, lh) for lh, _ in shapes]
    c0, pd0 = pcs.commit(list(zip(doms[:2], wire[:2])))
    c1, pd1 = pcs.commit(list(zip(doms[2:], wire[2:])))
    _, opd0 = okzg.commit(srs, [((1, lh), odft.mat_from_wire(a)) for (lh, _), a in zip(shapes[:2], wire[:2])])
    _, opd1 = okzg.commit(srs, [((1, lh), odft.mat_from_wire(a)) for (lh, _), a in zip(shapes[2:], wire[2:])])
    z = [int.from_bytes(rng.bytes(31), "little") for _ in range(4)]
    pts0 = [[z[0], z[1], z[2]], [z[0]]]
    pts1 = [[z[3], z[0]], []]
    opened, proof = pcs.open([(pd0, pts0), (pd1, pts1)])
    oopened, owits = okzg.open_(srs, [(opd0, pts0), (opd1, pts1)])
    for r, (pd, pts) in enumerate([(pd0, pts0), (pd1, pts1)]):
        for m in range(len(pd)):
            assert len(opened[r][m]) == len(pts[m]) and len(proof[r][m]) == len(pts[m])
            v1, w1 = pcs.open_matrix(pd[m], pts[m])
            for p in range(len(pts[m])):
                assert np.array_equal(opened[r][m][p], v1[p])
                assert np.array_equal(proof[r][m][p], w1[p])
                assert fr.from_wire(opened[r][m][p]) == oopened[r][m][p]
                assert g1.from_wire(proof[r][m][p]) == owits[r][m][p]
    for m in pd0 + pd1:
        m.free()


def test_open_batch_mixed_heights_large(ctx):
    """2^14 x 3 and 2^12 x 2 in one call (zero-filled tails, sorted MSM path) == per-matrix opens."""
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    pcs = pcs_new(ctx, (1 << 14) - 1, 31337)
    rng = np.random.default_rng(8)
    shapes = [(14, 3), (12, 2)]
    wire = [fr.random_wire(rng, (1 << lh) * w).reshape(1 << lh, w, 4) for lh, w in shapes]
    doms = [TwoAdicMultiplicativeCoset(1, lh) for lh, _ in shapes]
    _, pd = pcs.commit(list(zip(doms, wire)))
    z = [int.from_bytes(rng.bytes(31), "little") for _ in range(2)]
    pts = [[z[0], z[1]], [z[1]]]
    opened, proof = pcs.open([(pd, pts)])
    for m in range(2):
        v1, w1 = pcs.open_matrix(pd[m], pts[m])
        for p in range(len(pts[m])):
            assert np.array_equal(opened[0][m][p], v1[p])
            assert np.array_equal(proof[0][m][p], w1[p])
    # p(z) against a direct Horner evaluation of one column (size-independent check)
    coeffs = fr.from_wire(pd[1].coeffs()[:, 1, :])
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z[1] + c) % fr.P
    assert fr.from_wire(opened[0][1][0][1:2])[0] == acc
    for m in pd:
        m.free()


def test_open_batch_errors(ctx):
    from plonky3_eon_b200 import EonError
    vals = np.zeros((4, 4), dtype=np.uint64)
    wits = np.zeros((4, 8), dtype=np.uint64)
    pts = np.zeros((1, 4), dtype=np.uint64)
    with pytest.raises(EonError):                  # unknown handle
        ctx.call("eon_kzg_open_batch", 1, np.array([987654321], dtype=np.uint64), np.array([1], dtype=np.uint64),
                 pts, vals, wits)
    ctx.call("eon_kzg_open_batch", 0, np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64), pts, vals, wits)
    pcs = pcs_new(ctx, 7, 3)
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    rng = np.random.default_rng(2)
    ev = fr.random_wire(rng, 8).reshape(8, 1, 4)
    _, pd = pcs.commit([(TwoAdicMultiplicativeCoset(1, 3), ev)])
    bad = np.full((1, 4), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)   # not a canonical Fr
    with pytest.raises(EonError):
        ctx.call("eon_kzg_open_batch", 1, np.array([pd[0].handle], dtype=np.uint64), np.array([1], dtype=np.uint64),
                 bad, vals, wits)
    pd[0].free()
