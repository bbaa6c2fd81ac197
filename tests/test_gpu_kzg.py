"""GPU parity: G1::multi_exp, SRS generation and the KzgPcs commit / evaluate / open path through the
C ABI vs the oracle.  Bit-exact (affine G1 wire points and canonical Fr limbs compared directly).

Mirrors bn254/src/curve.rs:597-628 (multi_exp KATs), kzg/src/tests.rs:19-171 (KZG vectors) and
the shapes of eon-uni-stark/tests/fib_air.rs:112-136 (heights 1 and 8, alpha = 12345).
"""
import os

import numpy as np
import pytest

from oracle import dft as odft
from oracle import fr, g1, kzg as okzg

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kzg_small.npz")
P = fr.P


@pytest.fixture(scope="module")
def ctx():
    from plonky3_eon_b200 import Context
    c = Context(0)
    yield c
    c.close()


def pcs_new(ctx, max_degree, alpha):
    from plonky3_eon_b200 import GpuKzgPcs
    return GpuKzgPcs.new(max_degree, alpha, ctx=ctx)


def pt(wire):
    return g1.from_wire(np.asarray(wire).reshape(1, 8))[0]


def kG(k):
    return g1.mul(g1.G, k)


# ---- bn254/src/curve.rs:597-628 ----------------------------------------------------------------
def test_multi_exp_kats(ctx):
    from plonky3_eon_b200 import GpuKzgPcs
    pcs = GpuKzgPcs(ctx)
    assert pt(pcs.multi_exp(np.zeros((0, 8), np.uint64), np.zeros((0, 4), np.uint64))) is None
    assert pt(pcs.multi_exp(g1.to_wire([g1.G]), fr.to_wire([5]))) == kG(5)
    assert pt(pcs.multi_exp(g1.to_wire([g1.G, g1.G]), fr.to_wire([2, 3]))) == kG(5)
    assert pt(pcs.multi_exp(g1.to_wire([kG(7), kG(11)]), fr.to_wire([3, 5]))) == kG(76)
    with pytest.raises(AssertionError):
        pcs.multi_exp(g1.to_wire([g1.G]), fr.to_wire([1, 2]))


def test_multi_exp_edge_cases(ctx):
    from plonky3_eon_b200 import GpuKzgPcs
    pcs = GpuKzgPcs(ctx)
    A, B = kG(1234567), kG(987654321)
    cases = [
        ([A, g1.neg(A)], [9, 9], None),                          # P + (-P)
        ([A, A, A], [1, 1, 1], g1.mul(A, 3)),                    # doubling inside one bucket
        ([A, B], [0, 0], None),                                  # zero scalars
        ([None, A, None], [5, 7, 9], g1.mul(A, 7)),              # identity points
        ([A, B], [P - 1, 1], g1.add(g1.neg(A), B)),              # -1 scalar (top signed digits)
        ([A], [(1 << 253) + 12345], g1.mul(A, (1 << 253) + 12345)),
        ([A, B], [0xFFFF, 0x8000], g1.add(g1.mul(A, 0xFFFF), g1.mul(B, 0x8000))),  # digit carry edges
    ]
    for pts, sc, want in cases:
        got = pt(pcs.multi_exp(g1.to_wire(pts), fr.to_wire(sc)))
        assert got == want, (sc,)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 100, 257])
def test_multi_exp_random_small(ctx, n):
    from plonky3_eon_b200 import GpuKzgPcs
    pcs = GpuKzgPcs(ctx)
    rng = np.random.default_rng(n)
    dl = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(n)]
    pts = [kG(d) for d in dl]
    scw = fr.random_wire(rng, n)
    sc = fr.from_wire(scw)
    assert pt(pcs.multi_exp(g1.to_wire(pts), scw)) == g1.msm_via_dlog(dl, sc)


# ---- SRS generation (kzg/src/params.rs:123-139) -------------------------------------------------
def test_srs_generate_matches_oracle(ctx):
    pcs = pcs_new(ctx, 40, 12345)
    assert pcs.max_degree == 40
    got = g1.from_wire(pcs.g1_powers())
    assert got == okzg.init_srs_unsafe(40, 12345)
    pcs0 = pcs_new(ctx, 3, 0)  # alpha = 0: [G, O, O, O]
    assert g1.from_wire(pcs0.g1_powers()) == [g1.G, None, None, None]


# ---- MSM over the SRS at scale, checked with the discrete-log shortcut -----------------------------
@pytest.mark.parametrize("log_n,ncols", [(10, 3), (13, 2), (16, 2), (18, 1)])
def test_msm_srs_dlog_shortcut(ctx, log_n, ncols):
    n = 1 << log_n
    alpha = 987654321
    pcs = pcs_new(ctx, n - 1, alpha)
    rng = np.random.default_rng(log_n)
    sc = fr.random_wire(rng, n * ncols).reshape(n, ncols, 4)
    out = np.zeros((ncols, 8), dtype=np.uint64)
    ctx.call("eon_msm_srs", sc, n, ncols, ncols, out)
    dl = okzg.srs_dlogs(n - 1, alpha)
    for c in range(ncols):
        col = fr.from_wire(sc[:, c, :])
        assert pt(out[c]) == g1.msm_via_dlog(dl, col), c


@pytest.mark.parametrize("kind", ["zeros", "ones", "equal", "one_bit", "small64", "fib", "few_buckets"])
def test_msm_skewed_scalars(ctx, kind):
    # SURVEY §7 hard part 4: structured traces concentrate in few buckets (oversized-bucket path)
    n = 1 << 14
    alpha = 31337
    pcs = pcs_new(ctx, n - 1, alpha)
    rng = np.random.default_rng(5)
    if kind == "zeros":
        vals = [0] * n
    elif kind == "ones":
        vals = [1] * n
    elif kind == "equal":
        vals = [int.from_bytes(rng.bytes(31), "little")] * n
    elif kind == "one_bit":
        vals = [1 << int(b) for b in rng.integers(0, 253, size=n)]
    elif kind == "small64":
        vals = [int(v) for v in rng.integers(0, 1 << 63, size=n)]
    elif kind == "fib":
        vals, a, b = [], 0, 1
        for _ in range(n):
            vals.append(a)
            a, b = b, (a + b) % P
    else:
        vals = [int(v) * 0x10001 for v in rng.integers(1, 4, size=n)]
    out = pcs.commit_column(fr.to_wire(vals))
    assert pt(out) == g1.msm_via_dlog(okzg.srs_dlogs(n - 1, alpha), vals)


@pytest.mark.parametrize("log_n,bits", [(15, 0), (15, 8), (15, 13), (16, 17), (18, 20)])
def test_window_tables_and_index_ranges(ctx, log_n, bits):
    """MSMs over the resident SRS with explicit window tables (all windows share one bucket set) and
    over index sub-ranges (the multi-GPU shard form) agree with the discrete-log shortcut."""
    import ctypes as C
    n = 1 << log_n
    alpha = 55555
    pcs = pcs_new(ctx, n - 1, alpha)
    ctx.call("eon_srs_set_window_tables", bits)
    assert int(ctx.lib.eon_srs_window_bits(ctx.h)) == bits
    rng = np.random.default_rng(bits + log_n)
    sc = fr.random_wire(rng, n * 2).reshape(n, 2, 4)
    dl = okzg.srs_dlogs(n - 1, alpha)
    out = np.zeros((2, 8), dtype=np.uint64)
    ctx.call("eon_msm_srs", sc, n, 2, 2, out)
    cols = [fr.from_wire(sc[:, c, :]) for c in range(2)]
    for c in range(2):
        assert pt(out[c]) == g1.msm_via_dlog(dl, cols[c]), c
    # index range [first, first + m): scalars row r pairs with srs[first + r]
    first, m = n // 3 + 1, n // 2 + 5
    d = ctx.dev_alloc(m * 2 * 32)
    ctx.h2d(d, np.ascontiguousarray(sc[:m]))
    ctx.call("eon_msm_srs_range_dev", C.c_void_p(d), first, m, 2, 2, out)
    ctx.dev_free(d)
    for c in range(2):
        assert pt(out[c]) == g1.msm_via_dlog(dl[first:first + m], cols[c][:m]), c


# ---- kzg/src/tests.rs ---------------------------------------------------------------------------
def test_kzg_batch_verification_vectors(ctx):
    # tests.rs:73-137: alpha = 42
    pcs = pcs_new(ctx, 16, 42)
    assert pt(pcs.commit_column(fr.to_wire([1, 2, 3]))) == kG(5377)
    assert pt(pcs.commit_column(fr.to_wire([5, 7, 11]))) == kG(19703)
    assert pt(pcs.commit_column(fr.to_wire([8, 3]))) == kG(134)
    assert pt(pcs.commit_column(fr.to_wire([40, 11]))) == kG(502)
    # tests.rs:149-171: alpha = 999
    pcs = pcs_new(ctx, 8, 999)
    assert pt(pcs.commit_column(fr.to_wire([1, 2]))) == kG(1999)
    assert pt(pcs.commit_column(fr.to_wire([2]))) == kG(2)
    assert pt(pcs.commit_column(np.zeros((0, 4), np.uint64))) is None


def test_pcs_roundtrip_vector(ctx):
    # tests.rs:19-48: alpha = 7, evals x + 1 on the size-8 subgroup, opened at 2
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    pcs = pcs_new(ctx, 8, 7)
    dom = TwoAdicMultiplicativeCoset(1, 3)
    evals = odft.mat_to_wire([[(x + 1) % P] for x in okzg.coset_points(1, 3)])
    commit, pdata = pcs.commit([(dom, evals)])
    assert odft.mat_from_wire(pdata[0].coeffs()) == [[1], [1]] + [[0]] * 6
    assert pt(commit[0][0]) == kG(8)
    opened, proof = pcs.open([(pdata, [[2]])])
    assert fr.from_wire(opened[0][0][0]) == [3]
    assert pt(proof[0][0][0][0]) == g1.G
    # degree guard: pcs.rs:238-240 / params.rs:164-173
    from plonky3_eon_b200 import DegreeTooLarge
    big = odft.mat_to_wire([[1]] * 16)
    with pytest.raises(DegreeTooLarge):
        pcs.commit([(TwoAdicMultiplicativeCoset(1, 4), big)])
    with pytest.raises(AssertionError):  # pcs.rs:233-237
        pcs.commit([(TwoAdicMultiplicativeCoset(1, 2), evals)])


def test_golden_fixture(ctx):
    from plonky3_eon_b200 import GpuKzgPcs, TwoAdicMultiplicativeCoset
    g = np.load(GOLD)
    pcs = GpuKzgPcs.from_srs(g["srs"], ctx=ctx)
    dom = TwoAdicMultiplicativeCoset(1, 3)
    commit, pdata = pcs.commit([(dom, g["kzg_evals"])])
    assert np.array_equal(commit[0], g["kzg_commit"])
    assert np.array_equal(pdata[0].coeffs(), g["kzg_coeffs"])
    opened, proof = pcs.open([(pdata, [[g["kzg_points"][0], g["kzg_points"][1]]])])
    assert np.array_equal(np.stack(opened[0][0]), g["kzg_opened"])
    assert np.array_equal(np.stack(proof[0][0]), g["kzg_witness"])
    qdom = dom.create_disjoint_domain(16)
    assert (qdom.shift, qdom.log_size) == (fr.GENERATOR, 4)
    assert np.array_equal(pcs.get_evaluations_on_domain(pdata, 0, qdom), g["kzg_evals_on_quotient_domain"])
    assert np.array_equal(pcs.get_evaluations_on_domain(pdata, 0, dom), g["kzg_evals"])
    sh = fr.from_wire(g["kzg_shift"])[0]
    commit2, _ = pcs.commit([(TwoAdicMultiplicativeCoset(sh, 3), g["kzg_evals"])])
    assert np.array_equal(commit2[0], g["kzg_commit_shifted"])


@pytest.mark.parametrize("log_h,w,log_small", [(3, 2, 0), (3, 2, 2), (6, 3, 4), (10, 16, 7), (11, 5, 8)])
def test_evaluations_on_a_domain_smaller_than_the_polynomial(ctx, log_h, w, log_small):
    """get_evaluations_on_domain is a Horner loop over ANY coset (kzg/src/pcs.rs:278-286), also one with fewer
    points than the committed polynomial has coefficients; here: fold mod X^n - s^n, then a size-n coset NTT."""
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    h = 1 << log_h
    pcs = pcs_new(ctx, h - 1, 12345)
    rng = np.random.default_rng(1000 + log_h * 31 + log_small)
    evw = fr.random_wire(rng, h * w).reshape(h, w, 4)
    dom = TwoAdicMultiplicativeCoset(1, log_h)
    _, pdata = pcs.commit([(dom, evw)])
    coeffs = odft.mat_from_wire(pdata[0].coeffs())
    for shift in (1, fr.GENERATOR, 0x1234567890ABCDEF % P):
        small = TwoAdicMultiplicativeCoset(shift, log_small)
        got = pcs.get_evaluations_on_domain(pdata, 0, small)
        assert got.shape == (1 << log_small, w, 4)
        want = okzg.get_evaluations_on_domain({"coeffs": coeffs, "domain": None, "evals": None}, (shift, log_small))
        assert odft.mat_from_wire(got) == want
    pdata[0].free()


@pytest.mark.parametrize("log_h,w", [(0, 2), (1, 1), (3, 3), (6, 4)])
def test_commit_open_vs_oracle(ctx, log_h, w):
    # heights 1 and 8 are the reference's end-to-end shapes (eon-uni-stark/tests/fib_air.rs:127-131)
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    h = 1 << log_h
    alpha = 12345
    pcs = pcs_new(ctx, max(h - 1, 1), alpha)
    srs = okzg.init_srs_unsafe(max(h - 1, 1), alpha)
    rng = np.random.default_rng(log_h * 10 + w)
    evw = fr.random_wire(rng, h * w).reshape(h, w, 4)
    ev = odft.mat_from_wire(evw)
    dom = TwoAdicMultiplicativeCoset(1, log_h)
    commit, pdata = pcs.commit([(dom, evw)])
    ocommit, opdata = okzg.commit(srs, [((1, log_h), ev)])
    assert g1.from_wire(commit[0]) == ocommit[0]
    assert odft.mat_from_wire(pdata[0].coeffs()) == opdata[0]["coeffs"]
    zeta = int.from_bytes(rng.bytes(31), "little")
    pts = [zeta, dom.next_point(zeta)]
    opened, proof = pcs.open([(pdata, [pts])])
    oopened, owits = okzg.open_(srs, [(opdata, [pts])])
    for i in range(2):
        assert fr.from_wire(opened[0][0][i]) == oopened[0][0][i]
        assert g1.from_wire(proof[0][0][i]) == owits[0][0][i]
    qdom = dom.create_disjoint_domain(2 * h)
    got = pcs.get_evaluations_on_domain(pdata, 0, qdom)
    assert odft.mat_from_wire(got) == okzg.get_evaluations_on_domain(opdata[0], (qdom.shift, qdom.log_size))
    pdata[0].free()


def test_commit_quotient_chunks(ctx):
    # commit/src/pcs.rs:82-102 + commit/src/domain.rs:174-221
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    alpha = 12345
    pcs = pcs_new(ctx, 15, alpha)
    srs = okzg.init_srs_unsafe(15, alpha)
    rng = np.random.default_rng(77)
    qdom = TwoAdicMultiplicativeCoset(fr.GENERATOR, 4)
    qw = fr.random_wire(rng, 16).reshape(16, 1, 4)
    q = odft.mat_from_wire(qw)
    commit, pdata = pcs.commit_quotient(qdom, qw, 2)
    doms = okzg.split_domains((fr.GENERATOR, 4), 2)
    subs = okzg.split_evals(2, q)
    ocommit, _ = okzg.commit(srs, list(zip(doms, subs)))
    assert len(commit) == 2
    for m in range(2):
        assert (pdata[m].domain.shift, pdata[m].domain.log_size) == doms[m]
        assert g1.from_wire(commit[m]) == ocommit[m]


def test_quotient_scan_large(ctx):
    # quotient_and_eval (kzg/src/util.rs:100-111) at a size that exercises all three scan levels
    h, w = (1 << 17) + 0, 3
    rng = np.random.default_rng(4)
    cw = fr.random_wire(rng, h * w).reshape(h, w, 4)
    z = int.from_bytes(rng.bytes(31), "little")
    d_c = ctx.dev_alloc(cw.nbytes)
    d_q = ctx.dev_alloc(cw.nbytes)
    ctx.h2d(d_c, cw)
    vals = np.zeros((w, 4), dtype=np.uint64)
    zw = fr.to_wire([z])[0].copy()
    ctx.call("eon_quotient_and_eval_dev", d_c, h, w, zw, d_q, vals)
    qw = np.zeros_like(cw)
    ctx.d2h(qw, d_q)
    ctx.dev_free(d_c)
    ctx.dev_free(d_q)
    for c in range(w):
        coeffs = fr.from_wire(cw[:, c, :])
        q, v = okzg.quotient_and_eval(coeffs, z)
        assert fr.from_wire(vals[c])[0] == v
        assert fr.from_wire(qw[:h - 1, c, :]) == q
        assert not qw[h - 1, c].any()


def test_config2_shape_commit_and_open_properties(ctx):
    """2^20 x 16 (BASELINE config 2): commitments checked through the discrete-log shortcut on
    every column; one opening checked as witness == (f(alpha) - f(z)) / (alpha - z) * G."""
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    log_h, w = 20, 16
    h = 1 << log_h
    alpha = 12345
    pcs = pcs_new(ctx, h - 1, alpha)
    rng = np.random.default_rng(2)
    evw = fr.random_wire(rng, h * w).reshape(h, w, 4)
    dom = TwoAdicMultiplicativeCoset(1, log_h)
    commit, pdata = pcs.commit([(dom, evw)])
    coeffs = pdata[0].coeffs()
    # commitment[c] = f_c(alpha) * G
    f_alpha = []
    for c in range(w):
        col = fr.from_wire(coeffs[:, c, :])
        fa = okzg.eval_poly(col, alpha)
        f_alpha.append(fa)
        assert pt(commit[0][c]) == kG(fa), c
    # coefficients really are the interpolation of the evaluations: spot-check f_c(omega^j) == evals[j][c]
    g = fr.two_adic_generator(log_h)
    col0 = fr.from_wire(coeffs[:, 0, :])
    for j in (0, 1, h - 1, 777777):
        assert okzg.eval_poly(col0, pow(g, j, P)) == fr.from_wire(evw[j, 0])[0]
    # open all columns at one point; verify algebraically
    z = int.from_bytes(rng.bytes(31), "little") % P
    opened, proof = pcs.open([(pdata, [[z]])])
    inv = pow((alpha - z) % P, -1, P)
    for c in (0, 7, 15):
        col = col0 if c == 0 else fr.from_wire(coeffs[:, c, :])
        fz = okzg.eval_poly(col, z)
        assert fr.from_wire(opened[0][0][0][c])[0] == fz
        assert pt(proof[0][0][0][c]) == kG((f_alpha[c] - fz) * inv % P)
    pdata[0].free()


# ---- KzgMmcs (kzg/src/mmcs.rs:155-237; kzg/src/tests.rs:50-70) -----------------------------------
def test_mmcs_roundtrip_kat(ctx):
    from plonky3_eon_b200 import GpuKzgMmcs
    mmcs = GpuKzgMmcs.new(4, 5, ctx=ctx)
    m = fr.to_wire([1, 2, 3, 4]).reshape(2, 2, 4)
    com, pd = mmcs.commit([m])
    assert pt(com[0][0]) == kG(16) and pt(com[0][1]) == kG(22)
    opened, wits = mmcs.open_batch(0, pd)
    assert fr.from_wire(opened[0]) == [1, 2]
    assert pt(wits[0][0]) == kG(3) and pt(wits[0][1]) == kG(4)
    assert np.array_equal(mmcs.get_matrices(pd)[0], m)
    pd.free()


def test_mmcs_mixed_heights_vs_oracle(ctx):
    """matrices of heights 8, 6 (not a power of two), 4, 1 and widths 3, 2, 1, 2, every index."""
    from plonky3_eon_b200 import GpuKzgMmcs
    alpha = 999
    mmcs = GpuKzgMmcs.new(16, alpha, ctx=ctx)
    srs = okzg.init_srs_unsafe(16, alpha)
    rng = np.random.default_rng(11)
    shapes = [(8, 3), (6, 2), (4, 1), (1, 2)]
    wire = [fr.random_wire(rng, h * w).reshape(h, w, 4) for h, w in shapes]
    mats = [odft.mat_from_wire(a) for a in wire]
    com, pd = mmcs.commit(wire)
    ocom = okzg.mmcs_commit(srs, mats)
    for a, b in zip(com, ocom):
        assert g1.from_wire(a) == b
    for index in range(8):
        opened, wits = mmcs.open_batch(index, pd)
        oopened, owits = okzg.mmcs_open_batch(srs, index, mats)
        for i in range(len(shapes)):
            assert fr.from_wire(opened[i]) == oopened[i], (index, i)
            assert g1.from_wire(wits[i]) == owits[i], (index, i)
    pd.free()


def test_mmcs_degree_too_large(ctx):
    from plonky3_eon_b200 import DegreeTooLarge, GpuKzgMmcs
    mmcs = GpuKzgMmcs.new(3, 7, ctx=ctx)
    with pytest.raises(DegreeTooLarge):
        mmcs.commit([fr.to_wire(list(range(5))).reshape(5, 1, 4)])


def test_pipelined_host_entry_points_match_unpipelined(ctx):
    """eon_kzg_commit / eon_kzg_evals_on_coset move column groups over PCIe under the compute of the
    previous group for matrices >= 32 MiB.  Same bytes as the single-copy path (EON_NO_PIPELINE), from
    pinned and from pageable host memory, and spot checks against the oracle's Horner evaluation."""
    import torch
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    log_h, w, alpha = 17, 12, 4242
    h = 1 << log_h
    pcs = pcs_new(ctx, h - 1, alpha)
    evw = fr.random_wire(np.random.default_rng(21), h * w).reshape(h, w, 4)
    pinned = torch.from_numpy(evw.view(np.int64)).pin_memory().numpy().view(np.uint64)
    dom = TwoAdicMultiplicativeCoset(1, log_h)
    qdom = dom.create_disjoint_domain(2 * h)
    res = {}
    for name, src in (("plain", evw), ("piped", evw), ("piped_pinned", pinned)):
        if name == "plain":
            os.environ["EON_NO_PIPELINE"] = "1"
        else:
            os.environ.pop("EON_NO_PIPELINE", None)
        commit, pd = pcs.commit([(dom, src)])
        lde = pcs.get_evaluations_on_domain(pd, 0, qdom)
        res[name] = (commit[0].copy(), pd[0].coeffs(), lde)
        pd[0].free()
    os.environ.pop("EON_NO_PIPELINE", None)
    for name in ("piped", "piped_pinned"):
        for a, b in zip(res["plain"], res[name]):
            assert np.array_equal(a, b), name
    coeffs, lde = res["piped"][1], res["piped"][2]
    g2 = fr.two_adic_generator(log_h + 1)
    for c in (0, 5, 11):
        col = fr.from_wire(coeffs[:, c, :])
        for j in (0, 3, 2 * h - 1):
            x = fr.GENERATOR * pow(g2, j, P) % P
            assert fr.from_wire(lde[j, c])[0] == okzg.eval_poly(col, x), (c, j)


@pytest.mark.parametrize("log_h,w", [(3, 2), (10, 5), (17, 12)])
def test_commit_with_lde_hint_matches_two_calls(ctx, log_h, w):
    """GpuKzgPcs.with_lde_hint: commit() also produces the quotient-coset evaluations
    (eon_kzg_commit_lde).  Same commitments, coefficients and LDE bytes as commit() followed by
    get_evaluations_on_domain(); a non-matching domain still goes through eon_kzg_evals_on_coset."""
    from plonky3_eon_b200 import GpuKzgPcs, TwoAdicMultiplicativeCoset
    h, alpha = 1 << log_h, 777
    pcs = pcs_new(ctx, h - 1, alpha)
    evw = fr.random_wire(np.random.default_rng(log_h), h * w).reshape(h, w, 4)
    dom = TwoAdicMultiplicativeCoset(1, log_h)
    qdom = dom.create_disjoint_domain(2 * h)
    c0, pd0 = pcs.commit([(dom, evw)])
    lde0 = pcs.get_evaluations_on_domain(pd0, 0, qdom)
    hinted = GpuKzgPcs(ctx).with_lde_hint(1)
    c1, pd1 = hinted.commit([(dom, evw)])
    assert pd1[0].lde is not None
    lde1 = hinted.get_evaluations_on_domain(pd1, 0, qdom)
    assert lde1 is pd1[0].lde[2]
    assert np.array_equal(c0[0], c1[0])
    assert np.array_equal(pd0[0].coeffs(), pd1[0].coeffs())
    assert np.array_equal(lde0, lde1)
    other = dom.create_disjoint_domain(4 * h)
    assert np.array_equal(hinted.get_evaluations_on_domain(pd1, 0, other),
                          pcs.get_evaluations_on_domain(pd0, 0, other))
    pd0[0].free()
    pd1[0].free()


@pytest.mark.parametrize("log_h,w", [(0, 1), (4, 3), (15, 6)])
def test_commit_lde_dev_matches_two_calls(ctx, log_h, w):
    """eon_kzg_commit_lde_dev (device buffers; the LDE transform runs on a second stream beside the MSM)
    returns the same commitments, coefficients and LDE bytes as eon_kzg_commit_dev followed by
    eon_kzg_evals_on_coset_dev."""
    import ctypes as C
    from plonky3_eon_b200 import field
    h, alpha = 1 << log_h, 31
    pcs_new(ctx, max(h - 1, 1), alpha)
    evw = fr.random_wire(np.random.default_rng(log_h + 40), h * w).reshape(h, w, 4)
    d_ev = ctx.dev_alloc(evw.nbytes)
    ctx.h2d(d_ev, evw)
    lde_bytes = 2 * evw.nbytes
    d_l1, d_l2 = ctx.dev_alloc(lde_bytes), ctx.dev_alloc(lde_bytes)
    one, five = field.to_wire(1), field.to_wire(5)
    c1, c2 = np.zeros((w, 8), np.uint64), np.zeros((w, 8), np.uint64)
    h1, h2 = C.c_uint64(0), C.c_uint64(0)
    ctx.call("eon_kzg_commit_dev", C.c_void_p(d_ev), log_h, w, one, c1, C.byref(h1))
    ctx.call("eon_kzg_evals_on_coset_dev", h1, log_h + 1, five, C.c_void_p(d_l1))
    ctx.call("eon_kzg_commit_lde_dev", C.c_void_p(d_ev), log_h, w, one, c2, C.byref(h2), log_h + 1, five, C.c_void_p(d_l2))
    l1 = np.empty((2 * h, w, 4), np.uint64)
    l2 = np.empty((2 * h, w, 4), np.uint64)
    ctx.d2h(l1, d_l1)
    ctx.d2h(l2, d_l2)
    k1 = np.empty((h, w, 4), np.uint64)
    k2 = np.empty((h, w, 4), np.uint64)
    ctx.call("eon_kzg_read_coeffs", h1, k1)
    ctx.call("eon_kzg_read_coeffs", h2, k2)
    assert np.array_equal(c1, c2) and np.array_equal(l1, l2) and np.array_equal(k1, k2)
    assert odft.mat_from_wire(l2) == odft.coset_lde_batch(odft.mat_from_wire(evw), 1, fr.GENERATOR) if log_h <= 4 else True
    for p in (d_ev, d_l1, d_l2):
        ctx.dev_free(p)
    ctx.call("eon_handle_free", h1)
    ctx.call("eon_handle_free", h2)


def test_context_shared_by_threads(ctx):
    """callers may run several proofs from different threads against one Pcs.  Four host threads
    commit + open different traces through ONE context at the same time (ctypes drops the GIL inside the
    call; the context serialises them); every result equals the single-threaded one."""
    import threading
    from plonky3_eon_b200 import TwoAdicMultiplicativeCoset
    h, w, alpha = 1 << 10, 3, 12345
    pcs = pcs_new(ctx, h - 1, alpha)
    dom = TwoAdicMultiplicativeCoset(1, 10)
    rng = np.random.default_rng(99)
    traces = [fr.random_wire(rng, h * w).reshape(h, w, 4) for _ in range(4)]
    zeta = 0x1234567

    def job(ev):
        c, pd = pcs.commit([(dom, ev)])
        o, p = pcs.open([(pd, [[zeta]])])
        lde = pcs.get_evaluations_on_domain(pd, 0, dom.create_disjoint_domain(2 * h))
        pd[0].free()
        return c[0].copy(), o[0][0][0].copy(), p[0][0][0].copy(), lde.copy()

    want = [job(ev) for ev in traces]
    got = [None] * 4
    errs = []

    def run(i):
        try:
            for _ in range(3):
                got[i] = job(traces[i])
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=run, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for a, b in zip(want, got):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
