"""Multi-device context (eon_mctx_*, include/eon_kzg.h) vs the single-device entry points: byte-identical results.

The reference prover is one process handing whole matrices to the Pcs (eon-uni-stark/src/prover.rs:186-187,
371-372,424-442); the multi-device context splits their columns (kzg/src/pcs.rs:244-249: every column is
independent) or, for a lone MSM, the point index range.  On a box with one GPU the same ordinal is listed several
times (several shard contexts on one device), which exercises all of the sharding / scatter logic; with more GPUs
the real devices are used as well.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import fr, g1

pytestmark = pytest.mark.gpu
P = fr.P
ALPHA = 12345


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.fixture(scope="module")
def single():
    from plonky3_eon_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=range(3))
def multi(request):
    from plonky3_eon_b200 import MultiContext
    lists = device_lists()
    if request.param >= len(lists):
        pytest.skip("needs >= 2 GPUs")
    m = MultiContext(lists[request.param])
    yield m
    m.close()


def rand_matrix(seed, h, w):
    return fr.random_wire(np.random.default_rng(seed), h * w).reshape(h, w, 4)


@pytest.mark.parametrize("log_h,w", [(3, 1), (6, 5), (10, 16), (12, 7)])
def test_dft_family_matches_single_device(single, multi, log_h, w):
    from plonky3_eon_b200 import GpuDft
    a, b = GpuDft(single), GpuDft(multi)
    m = rand_matrix(log_h * 100 + w, 1 << log_h, w)
    shift = 0x1234567 % P
    assert np.array_equal(a.dft_batch(m), b.dft_batch(m))
    assert np.array_equal(a.coset_dft_batch(m, shift), b.coset_dft_batch(m, shift))
    assert np.array_equal(a.idft_batch(m), b.idft_batch(m))
    assert np.array_equal(a.coset_idft_batch(m, shift), b.coset_idft_batch(m, shift))
    assert np.array_equal(a.coset_lde_batch(m, 2, fr.GENERATOR), b.coset_lde_batch(m, 2, fr.GENERATOR))


@pytest.mark.parametrize("log_h,w,hint", [(0, 2, False), (4, 1, True), (8, 16, True), (11, 5, False), (14, 16, True)])
def test_pcs_sequence_matches_single_device(single, multi, log_h, w, hint):
    """commit -> get_evaluations_on_domain -> commit_quotient -> open, as eon_uni_stark::prove calls them."""
    from plonky3_eon_b200 import GpuKzgPcs, TwoAdicMultiplicativeCoset
    h = 1 << log_h
    results = []
    for ctx in (single, multi):
        pcs = GpuKzgPcs.new(max(2 * h - 1, 1), ALPHA, ctx=ctx)
        if hint:
            pcs = pcs.with_lde_hint(1)
        dom = TwoAdicMultiplicativeCoset(1, log_h)
        evals = rand_matrix(7 + log_h, h, w)
        commit, pd = pcs.commit([(dom, evals)])
        qdom = dom.create_disjoint_domain(2 * h)
        lde = np.array(pcs.get_evaluations_on_domain(pd, 0, qdom))
        other = np.array(pcs.get_evaluations_on_domain(pd, 0, TwoAdicMultiplicativeCoset(7, log_h + 1)))
        quot = rand_matrix(99 + log_h, 2 * h, 1)
        qcommit, qpd = pcs.commit_quotient(qdom, quot, 2)
        zeta = 0x1234567890ABCDEF1234567890ABCDEF
        opened, proof = pcs.open([(pd, [[zeta, dom.next_point(zeta)]]), (qpd, [[zeta], [zeta]])])
        coeffs = pd[0].coeffs()
        results.append((commit, lde, other, qcommit, opened, proof, coeffs, [q.coeffs() for q in qpd]))
        for m in pd + qpd:
            m.free()

    def flat(x):
        if isinstance(x, np.ndarray):
            return [x]
        out = []
        for y in x:
            out.extend(flat(y))
        return out
    a, b = (flat(list(r)) for r in results)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)


def test_commit_kat_through_the_multi_device_context(multi):
    """kzg/src/tests.rs:73-137: alpha = 42, coefficients [1,2,3] and [5,7,11] -> 5377 G and 19703 G."""
    from plonky3_eon_b200 import GpuKzgMmcs
    mm = GpuKzgMmcs.new(2, 42, ctx=multi)
    mat = fr.to_wire([1, 5, 2, 7, 3, 11]).reshape(3, 2, 4)
    commit, pd = mm.commit([mat])
    got = g1.from_wire(commit[0])
    assert got[0] == g1.mul(g1.G, 5377) and got[1] == g1.mul(g1.G, 19703)
    pd.free()


@pytest.mark.parametrize("log_n,ncols", [(10, 1), (16, 1), (17, 2), (16, 9)])
def test_msm_srs_index_and_column_sharding(single, multi, log_n, ncols):
    """G1::multi_exp over the resident SRS: fewer columns than devices -> index-range shards with per-shard window
    tables and the partial sums added on the first device; otherwise column shards.  Checked against the
    single-device result and, through the discrete logs of the synthetic SRS (P_i = alpha^i G), against
    (sum_i s_i alpha^i) G."""
    n = 1 << log_n
    sc = rand_matrix(5 + log_n + ncols, n, ncols)
    out_s = np.zeros((ncols, 8), dtype=np.uint64)
    out_m = np.zeros((ncols, 8), dtype=np.uint64)
    for ctx, out in ((single, out_s), (multi, out_m)):
        ctx.call("eon_srs_generate_unsafe", fr.to_wire([ALPHA])[0], n)
        ctx.call("eon_msm_srs", sc, n, ncols, ncols, out)
    assert np.array_equal(out_s, out_m)
    vals = np.array(fr.from_wire(sc.reshape(-1, 4)), dtype=object).reshape(n, ncols)
    apow, acc = 1, [0] * ncols
    for i in range(n):
        for c in range(ncols):
            acc[c] = (acc[c] + int(vals[i, c]) * apow) % P
        apow = apow * ALPHA % P
    for c in range(ncols):
        assert g1.from_wire(out_m[c:c + 1])[0] == g1.mul(g1.G, acc[c])


def test_msm_points_and_errors(single, multi):
    from plonky3_eon_b200 import DegreeTooLarge, EonError, GpuKzgPcs, TwoAdicMultiplicativeCoset
    rng = np.random.default_rng(3)
    n = 1 << 13
    ks = [int.from_bytes(rng.bytes(16), "little") for _ in range(64)]
    pts = g1.to_wire([g1.mul(g1.G, k) for k in ks] * (n // 64))
    sc = rand_matrix(11, n, 1).reshape(n, 4)
    a = GpuKzgPcs(single).multi_exp(pts, sc)
    b = GpuKzgPcs(multi).multi_exp(pts, sc)
    assert np.array_equal(a, b)
    # empty MSM -> identity (bn254/src/curve.rs:165-167)
    assert not GpuKzgPcs(multi).multi_exp(np.zeros((0, 8), np.uint64), np.zeros((0, 4), np.uint64)).any()
    pcs = GpuKzgPcs.new(7, ALPHA, ctx=multi)
    with pytest.raises(DegreeTooLarge):       # kzg/src/pcs.rs:238-240
        pcs.commit([(TwoAdicMultiplicativeCoset(1, 4), rand_matrix(1, 16, 3))])
    with pytest.raises(EonError):
        multi.call("eon_handle_free", C.c_uint64(123456))


def test_prover_data_is_released_with_its_python_object(single):
    """ProverData is dropped with its scope in the reference; here the finalizer returns the HBM buffer."""
    import gc

    import torch
    from plonky3_eon_b200 import GpuKzgPcs, TwoAdicMultiplicativeCoset
    log_h, w = 16, 16
    pcs = GpuKzgPcs.new((1 << log_h) - 1, ALPHA, ctx=single)
    dom = TwoAdicMultiplicativeCoset(1, log_h)
    evals = rand_matrix(1, 1 << log_h, w)
    used = []
    for i in range(12):
        _, pd = pcs.commit([(dom, evals)])
        del pd
        gc.collect()
        free, total = torch.cuda.mem_get_info(0)
        used.append(total - free)
    # 32 MiB of coefficients per commit: a leak would add 32 MiB per cycle after the pool of 4 is in use
    assert used[-1] - used[5] < (16 << 20), used


def test_ragged_shapes_and_mmcs_through_the_multi_device_context(single, multi):
    """Fewer columns than devices, zero-width matrices, non-power-of-two KzgMmcs heights (kzg/src/mmcs.rs:155-237:
    coefficient-form matrices of mixed heights opened at one index): same bytes as the single-device calls."""
    from plonky3_eon_b200 import GpuKzgMmcs, GpuKzgPcs, TwoAdicMultiplicativeCoset
    outs = []
    for ctx in (single, multi):
        pcs = GpuKzgPcs.new(63, ALPHA, ctx=ctx)
        dom = TwoAdicMultiplicativeCoset(1, 5)
        res = []
        for w in (1, 0):                                   # one column (fewer than devices), then none
            evals = rand_matrix(300 + w, 32, w) if w else np.zeros((32, 0, 4), dtype=np.uint64)
            commit, pd = pcs.commit([(dom, evals)])
            res.append(commit[0])
            if w:
                res.append(pd[0].coeffs())
                res.append(np.array(pcs.get_evaluations_on_domain(pd, 0, TwoAdicMultiplicativeCoset(3, 3))))  # smaller
                opened, proof = pcs.open([(pd, [[5, 6, 7]])])
                res.extend(opened[0][0] + proof[0][0])
            pd[0].free()
        mm = GpuKzgMmcs(ctx)
        mats = [rand_matrix(400, 48, 5), rand_matrix(401, 7, 2), rand_matrix(402, 64, 1)]   # heights 48, 7, 64
        commit, pdm = mm.commit(mats)
        res.extend(commit)
        opened, wits = mm.open_batch(37, pdm)
        res.extend(opened + wits)
        pdm.free()
        outs.append(res)
    assert len(outs[0]) == len(outs[1])
    for a, b in zip(*outs):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_multi_device_context_rejects_bad_calls(multi):
    from plonky3_eon_b200 import EonError
    out = np.zeros((4, 4), dtype=np.uint64)
    with pytest.raises(EonError):
        multi.call("eon_kzg_evals_on_coset", C.c_uint64(987654321), 4, fr.to_wire([3])[0], out)
    with pytest.raises(EonError):                           # shift 0 is not a coset (field/src/coset.rs:77-86)
        multi.call("eon_coset_dft_batch", rand_matrix(1, 4, 2), np.zeros((4, 2, 4), dtype=np.uint64), 2, 2,
                   np.zeros(4, dtype=np.uint64))
    with pytest.raises(EonError):                           # ld < ncols
        multi.call("eon_msm_srs", rand_matrix(2, 4, 2), 4, 2, 1, out)
