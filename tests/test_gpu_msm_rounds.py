"""GPU parity of the batched-affine pairwise rounds (csrc/msm_tree.cu): every MSM parity case of
tests/test_gpu_kzg.py again with the round count forced to 1, 2, 3 and 5 (the automatic policy only
switches them on for large inputs), so that the degenerate pairs — identity operands, P + P, P + (-P),
empty and oversized buckets, segment tails — go through k_tree_fwd / k_tree_bwd.  Results must be the
same affine points as with the XYZZ-only accumulation (rounds = 0) and as the oracle's."""
import numpy as np
import pytest

import test_gpu_kzg as T
from oracle import fr, g1, kzg as okzg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from plonky3_eon_b200 import Context
    c = Context(0)
    yield c
    c.close()


# (rounds, slice schedule): the slice schedule (round 0 walked by 64 MiB table slice, msm_tree.cu) only
# switches itself on for tables beyond the L2, so it is forced here to put every degenerate case through it
@pytest.fixture(params=[(1, 0), (2, 0), (3, 0), (5, 0), (1, 1), (3, 1)], ids=lambda p: f"r{p[0]}{'s' if p[1] else ''}")
def rctx(ctx, request):
    ctx.call("eon_msm_set_rounds", request.param[0])
    ctx.call("eon_msm_set_slice_schedule", request.param[1])
    yield ctx
    ctx.call("eon_msm_set_rounds", -1)
    ctx.call("eon_msm_set_slice_schedule", -1)


def test_kats_and_edge_cases(rctx):
    T.test_multi_exp_kats(rctx)
    T.test_multi_exp_edge_cases(rctx)


@pytest.mark.parametrize("n", [1, 3, 64, 257])
def test_random_small(rctx, n):
    T.test_multi_exp_random_small(rctx, n)


@pytest.mark.parametrize("kind", ["zeros", "ones", "equal", "one_bit", "small64", "fib", "few_buckets"])
def test_skewed(rctx, kind):
    T.test_msm_skewed_scalars(rctx, kind)


@pytest.mark.parametrize("log_n,ncols", [(10, 3), (16, 2)])
def test_dlog_shortcut(rctx, log_n, ncols):
    T.test_msm_srs_dlog_shortcut(rctx, log_n, ncols)


@pytest.mark.parametrize("log_n,bits", [(15, 0), (15, 13), (16, 17)])
def test_tables_and_ranges(rctx, log_n, bits):
    T.test_window_tables_and_index_ranges(rctx, log_n, bits)


def test_duplicate_bases_double_inside_rounds(rctx):
    """all bases equal: every pair of every round is a doubling (P + P), plus sign-flipped entries
    that cancel (P + (-P)) when digits are negative."""
    from plonky3_eon_b200 import GpuKzgPcs
    pcs = GpuKzgPcs(rctx)
    n = 300
    A = g1.mul(g1.G, 424242)
    rng = np.random.default_rng(3)
    sc = [int(v) for v in rng.integers(1, 1 << 62, size=n)]
    sc[::7] = [fr.P - 5] * len(sc[::7])            # negative top digits
    want = g1.mul(A, sum(sc) % fr.P)
    got = T.pt(pcs.multi_exp(g1.to_wire([A] * n), fr.to_wire(sc)))
    assert got == want


def test_commit_matches_rounds_off(ctx):
    """the same commit with rounds forced to 3 and to 0 gives byte-identical commitments"""
    h, w, alpha = 1 << 12, 4, 12345
    pcs = T.pcs_new(ctx, h - 1, alpha)
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    ev = fr.random_wire(np.random.default_rng(8), h * w).reshape(h, w, 4)
    dom = TwoAdicMultiplicativeCoset(1, 12)
    outs = []
    for r, sl in ((0, 0), (3, 0), (3, 1)):
        ctx.call("eon_msm_set_rounds", r)
        ctx.call("eon_msm_set_slice_schedule", sl)
        c, pd = pcs.commit([(dom, ev)])
        outs.append(c[0].copy())
        pd[0].free()
    ctx.call("eon_msm_set_rounds", -1)
    ctx.call("eon_msm_set_slice_schedule", -1)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_slice_schedule_multi_slice_table(ctx):
    """2^18 points x 6 columns with a 16-bit window: the table has 16 levels = 4 slices of 2^20 points, so
    the pair records really are binned; commitments equal the slot-order schedule's and the dlog shortcut."""
    n, ncols, alpha = 1 << 18, 6, 12345
    pcs = T.pcs_new(ctx, n - 1, alpha)
    ctx.call("eon_srs_set_window_tables", 16)
    rng = np.random.default_rng(12)
    sc = fr.random_wire(rng, n * ncols).reshape(n, ncols, 4)
    outs = []
    for sl in (0, 1):
        ctx.call("eon_msm_set_slice_schedule", sl)
        out = np.zeros((ncols, 8), dtype=np.uint64)
        ctx.call("eon_msm_srs", sc, n, ncols, ncols, out)
        outs.append(out)
    ctx.call("eon_msm_set_slice_schedule", -1)
    assert np.array_equal(outs[0], outs[1])
    assert ctx.lib.eon_msm_rounds_used(ctx.h) >= 1
    col0 = fr.from_wire(np.ascontiguousarray(sc[:, 0, :]))
    dl, a = [], 1
    for _ in range(n):
        dl.append(a)
        a = a * alpha % fr.P
    assert T.pt(outs[1][0]) == g1.msm_via_dlog(dl, col0)


@pytest.mark.parametrize("log_n,ncols,bits", [(6, 2, 0), (10, 3, 0), (12, 5, 0), (15, 4, 13), (16, 2, 17), (17, 3, 17)])
def test_two_stream_split_matches_one_batch(ctx, log_n, ncols, bits):
    """An MSM over few columns runs as two half-batches on two streams (csrc/msm.cu msm_run, workspace bank 1):
    forced on for every shape here (odd column counts, table and plain paths, rounds on) it must give the same
    points as the single batch, and the first column must equal the dlog shortcut."""
    n, alpha = 1 << log_n, 12345
    pcs = T.pcs_new(ctx, n - 1, alpha)
    ctx.call("eon_srs_set_window_tables", bits)
    rng = np.random.default_rng(100 + log_n + ncols)
    sc = fr.random_wire(rng, n * ncols).reshape(n, ncols, 4)
    sc[: n // 3, ncols - 1] = 0                      # a sparse last column: very different bucket loads per half
    outs = []
    for split, rounds in ((0, -1), (1, -1), (1, 2), (1, 0)):
        ctx.call("eon_msm_set_split", split)
        ctx.call("eon_msm_set_rounds", rounds)
        out = np.zeros((ncols, 8), dtype=np.uint64)
        ctx.call("eon_msm_srs", sc, n, ncols, ncols, out)
        outs.append(out)
    ctx.call("eon_msm_set_split", -1)
    ctx.call("eon_msm_set_rounds", -1)
    ctx.call("eon_srs_set_window_tables", 0)
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)
    col0 = fr.from_wire(np.ascontiguousarray(sc[:, 0, :]))
    dl, a = [], 1
    for _ in range(n):
        dl.append(a)
        a = a * alpha % fr.P
    assert T.pt(outs[1][0]) == g1.msm_via_dlog(dl, col0)


@pytest.mark.parametrize("fused", [1, 0])
def test_fused_sort_records_with_skewed_columns(ctx, fused):
    """The sort of a sliced MSM emits the pair records of round 0 itself (csrc/msm_sort.cu: k_sort_count2 /
    k_sort_place2 / k_pair_tail) once the table spans several slices: 2^16 points with a 14-bit window = 19 levels
    = 3 slices of 2^19 points.  Columns: all scalars equal (19 buckets of 2^16 entries each: windows beyond the
    shared-memory image, placed in global memory), three distinct values, sparse random with a zero block, dense
    random.  Every column must equal the discrete-log shortcut; the unfused flow (EON_SORT_FUSED=0 in a child
    process is not needed: the per-context switch below) gives the same bytes."""
    n, alpha = 1 << 16, 424243
    pcs = T.pcs_new(ctx, n - 1, alpha)
    ctx.call("eon_srs_set_window_tables", 14)
    rng = np.random.default_rng(77)
    cols = []
    cols.append([int.from_bytes(rng.bytes(31), "little")] * n)
    cols.append([int(v) * 0x10001 for v in rng.integers(1, 4, size=n)])
    sparse = fr.from_wire(fr.random_wire(rng, n))
    sparse[: n // 2] = [0] * (n // 2)
    cols.append(sparse)
    cols.append(fr.from_wire(fr.random_wire(rng, n)))
    sc = np.stack([fr.to_wire(c) for c in cols], axis=1)
    ctx.call("eon_msm_set_rounds", 3)
    ctx.call("eon_msm_set_slice_schedule", 1)
    ctx.call("eon_msm_set_split", 0)
    ctx.call("eon_msm_set_sort_mode", 1 if fused else 2)
    out = np.zeros((len(cols), 8), dtype=np.uint64)
    try:
        ctx.call("eon_msm_srs", np.ascontiguousarray(sc), n, len(cols), len(cols), out)
    finally:
        ctx.call("eon_msm_set_rounds", -1)
        ctx.call("eon_msm_set_slice_schedule", -1)
        ctx.call("eon_msm_set_split", -1)
        ctx.call("eon_msm_set_sort_mode", -1)
        ctx.call("eon_srs_set_window_tables", 0)
    dl = okzg.srs_dlogs(n - 1, alpha)
    for c, vals in enumerate(cols):
        assert T.pt(out[c]) == g1.msm_via_dlog(dl, vals), c
