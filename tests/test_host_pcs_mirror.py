"""CPU test of the host mirror's marshalling (plonky3_eon_b200/pcs.py, mmcs.py): argument order, the nesting of
rounds / matrices / points, the slicing of the flat ABI outputs, chunk views of commit_quotient.

The ABI calls are served by an oracle-backed stand-in that follows the contracts written in include/eon_kzg.h
(this checks the Python side of the boundary, not the kernels: those are the -m gpu tests)."""

import numpy as np

from oracle import dft as odft
from oracle import fr, g1, kzg as okzg


def _as_int(a):
    return int(a.value) if hasattr(a, "value") else int(a)


class OracleAbi:
    """Stand-in for lib.Context: serves the eon_kzg_* calls the Pcs / Mmcs mirrors make, from the oracle."""

    def __init__(self, max_degree, alpha):
        self.srs = okzg.init_srs_unsafe(max_degree, alpha)
        self.mats = {}          # handle -> coefficient matrix (canonical ints, rows x width)
        self.next = 1
        self.calls = []

    def srs_size(self):
        return len(self.srs)

    def _new(self, coeffs):
        h = self.next
        self.next += 1
        self.mats[h] = coeffs
        return h

    def call(self, name, *a):
        self.calls.append(name)
        getattr(self, "_" + name)(*a)

    def _eon_kzg_commit(self, evals, log_h, width, shift, cols, handle_ref):
        shift = fr.from_wire(np.asarray(shift).reshape(1, 4))[0]
        ev = odft.mat_from_wire(evals)
        assert len(ev) == 1 << log_h and evals.shape[1] == width
        c, pd = okzg.commit(self.srs, [((shift, log_h), ev)])
        if width:
            cols[:] = g1.to_wire(c[0])
        handle_ref._obj.value = self._new(pd[0]["coeffs"])

    def _eon_kzg_commit_quotient(self, evals, log_size, width, log_chunks, shift, cols, handles):
        shift = fr.from_wire(np.asarray(shift).reshape(1, 4))[0]
        ev = odft.mat_from_wire(evals)
        n = 1 << log_chunks
        doms = okzg.split_domains((shift, log_size), n)
        subs = okzg.split_evals(n, ev)
        c, pd = okzg.commit(self.srs, list(zip(doms, subs)))
        for i in range(n):
            if width:
                cols[i] = g1.to_wire(c[i])
            handles[i] = self._new(pd[i]["coeffs"])

    def _eon_kzg_commit_coeffs(self, coeffs, rows, width, cols, handle_ref):
        m = odft.mat_from_wire(coeffs) if rows else []
        c = okzg.mmcs_commit(self.srs, [m])
        if width and rows:
            cols[:] = g1.to_wire(c[0])
        handle_ref._obj.value = self._new(m)

    def _eon_kzg_open_batch(self, nmat, handles, npoints, points, vals, wits):
        k, pt = 0, 0
        for m in range(nmat):
            coeffs = self.mats[int(handles[m])]
            h = len(coeffs)
            w = len(coeffs[0]) if h else 0
            for _ in range(int(npoints[m])):
                z = fr.from_wire(points[pt:pt + 1])[0]
                pt += 1
                for c in range(w):
                    q, v = okzg.quotient_and_eval([coeffs[r][c] for r in range(h)], z)
                    vals[k] = fr.to_wire([v])[0]
                    wits[k] = g1.to_wire([okzg.commit_column(self.srs, q)])[0]
                    k += 1

    def _eon_handle_free(self, h):
        del self.mats[_as_int(h)]


def make_pcs(abi):
    from plonky3_eon_b200.pcs import GpuKzgPcs
    pcs = GpuKzgPcs.__new__(GpuKzgPcs)
    pcs.ctx = abi
    pcs.lde_hint = None
    return pcs


def test_commit_quotient_and_open_nesting():
    from plonky3_eon_b200.pcs import TwoAdicMultiplicativeCoset
    alpha = 12345
    abi = OracleAbi(15, alpha)
    pcs = make_pcs(abi)
    rng = np.random.default_rng(3)
    dom = TwoAdicMultiplicativeCoset(1, 3)
    t0 = fr.random_wire(rng, 8 * 2).reshape(8, 2, 4)
    t1 = fr.random_wire(rng, 8 * 1).reshape(8, 1, 4)
    (c_t, pd_t) = pcs.commit([(dom, t0), (dom, t1)])
    qdom = dom.create_disjoint_domain(16)
    qw = fr.random_wire(rng, 16 * 2).reshape(16, 2, 4)
    c_q, pd_q = pcs.commit_quotient(qdom, qw, 2)
    assert abi.calls.count("eon_kzg_commit_quotient") == 1 and abi.calls.count("eon_kzg_commit") == 2

    # oracle, straight from the reference's definitions
    srs = abi.srs
    oc_t, opd_t = okzg.commit(srs, [((1, 3), odft.mat_from_wire(t0)), ((1, 3), odft.mat_from_wire(t1))])
    doms = okzg.split_domains((qdom.shift, 4), 2)
    subs = okzg.split_evals(2, odft.mat_from_wire(qw))
    oc_q, opd_q = okzg.commit(srs, list(zip(doms, subs)))
    for m in range(2):
        assert g1.from_wire(c_t[m]) == oc_t[m]
        assert g1.from_wire(c_q[m]) == oc_q[m]
        assert (pd_q[m].domain.shift, pd_q[m].domain.log_size) == doms[m]
        assert odft.mat_from_wire(np.ascontiguousarray(pd_q[m].evals)) == subs[m]
        # same-domain fast path of get_evaluations_on_domain (pcs.rs:275-277) hands the chunk's own rows out
        got = pcs.get_evaluations_on_domain(pd_q, m, pd_q[m].domain)
        assert odft.mat_from_wire(got) == subs[m]

    zeta = 0x123456789ABCDEF
    rounds = [(pd_t, [[zeta, dom.next_point(zeta)], [zeta]]), (pd_q, [[zeta], [zeta]])]
    opened, proof = pcs.open(rounds)
    assert abi.calls.count("eon_kzg_open_batch") == 1
    oopened, owits = okzg.open_(srs, [(opd_t, rounds[0][1]), (opd_q, rounds[1][1])])
    assert len(opened) == 2 and len(proof) == 2
    for r in range(2):
        assert len(opened[r]) == 2
        for m in range(2):
            assert len(opened[r][m]) == len(rounds[r][1][m])
            for p in range(len(rounds[r][1][m])):
                assert fr.from_wire(opened[r][m][p]) == oopened[r][m][p]
                assert g1.from_wire(proof[r][m][p]) == owits[r][m][p]
    for m in pd_t + pd_q:
        m.free()
    assert not abi.mats


def test_open_of_nothing():
    abi = OracleAbi(3, 7)
    pcs = make_pcs(abi)
    assert pcs.open([]) == ([], [])
    assert "eon_kzg_open_batch" not in abi.calls


def test_mmcs_open_batch_marshalling():
    from plonky3_eon_b200.mmcs import GpuKzgMmcs
    alpha = 999
    abi = OracleAbi(16, alpha)
    mmcs = GpuKzgMmcs.__new__(GpuKzgMmcs)
    mmcs.ctx = abi
    rng = np.random.default_rng(11)
    shapes = [(8, 3), (6, 2), (4, 1), (1, 2)]
    wire = [fr.random_wire(rng, h * w).reshape(h, w, 4) for h, w in shapes]
    mats = [odft.mat_from_wire(a) for a in wire]
    com, pd = mmcs.commit(wire)
    for a, b in zip(com, okzg.mmcs_commit(abi.srs, mats)):
        assert g1.from_wire(a) == b
    for index in (0, 3, 7):
        opened, wits = mmcs.open_batch(index, pd)
        oopened, owits = okzg.mmcs_open_batch(abi.srs, index, mats)
        for i in range(len(shapes)):
            assert fr.from_wire(opened[i]) == oopened[i], (index, i)
            assert g1.from_wire(wits[i]) == owits[i], (index, i)
    assert abi.calls.count("eon_kzg_open_batch") == 3
