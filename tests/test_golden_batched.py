"""The batched Pcs calls against the committed fixture tests/golden/kzg_batched.npz (made by
tests/golden/make_golden.py from the oracle): Pcs::commit_quotient of a 16 x 2 quotient in 2 chunks and one
KzgPcs::open over a trace round and the chunk round, compared limb for limb.

The same checker runs three ways: on the GPU through the C ABI (-m gpu), on the CPU with the host mirror served
by the oracle-backed ABI stand-in (marshalling only), and the C port of the reference path against the fixture."""
import os

import numpy as np
import pytest

from oracle import fr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kzg_batched.npz")


def check_against_fixture(pcs):
    from plonky3_eon_b200.pcs import TwoAdicMultiplicativeCoset
    g = np.load(GOLD)
    dom = TwoAdicMultiplicativeCoset(1, 3)
    qdom = dom.create_disjoint_domain(16)
    c_t, pd_t = pcs.commit([(dom, g["trace0"]), (dom, g["trace1"])])
    assert np.array_equal(c_t[0], g["trace_commit0"]) and np.array_equal(c_t[1], g["trace_commit1"])
    c_q, pd_q = pcs.commit_quotient(qdom, g["quotient"], 2)
    for i in range(2):
        assert np.array_equal(c_q[i], g["chunk_commits"][i])
        assert fr.from_wire(g["chunk_shifts"][i:i + 1])[0] == pd_q[i].domain.shift and pd_q[i].domain.log_size == 3
    zeta, zeta_next = (int(v) for v in fr.from_wire(g["points"]))
    assert zeta_next == dom.next_point(zeta)
    rounds = [(pd_t, [[zeta, zeta_next], [zeta]]), (pd_q, [[zeta], [zeta]])]
    opened, proof = pcs.open(rounds)
    flat_v = [row for r in opened for m in r for p in m for row in p]
    flat_w = [row for r in proof for m in r for p in m for row in p]
    assert np.array_equal(np.stack(flat_v), g["opened_flat"])
    assert np.array_equal(np.stack(flat_w), g["witness_flat"])
    return pd_t + pd_q


def test_fixture_through_the_host_mirror_cpu():
    from test_host_pcs_mirror import OracleAbi, make_pcs
    g = np.load(GOLD)
    pd = check_against_fixture(make_pcs(OracleAbi(15, int(g["alpha"][0]))))
    for m in pd:
        m.free()


def test_fixture_vs_c_port():
    """oracle/c (the cpu_baseline / --impl reference arm) reproduces the chunk commitments of the fixture."""
    from oracle import cport
    g = np.load(GOLD)
    srs = cport.srs_generate(int(g["alpha"][0]), 16)
    q = g["quotient"]
    for i in range(2):
        shift = fr.from_wire(g["chunk_shifts"][i:i + 1])[0]
        commits, coeffs = cport.kzg_commit(np.ascontiguousarray(q[i::2]), shift, srs)
        assert np.array_equal(np.asarray(commits).reshape(2, 8), g["chunk_commits"][i])
        assert np.array_equal(np.asarray(coeffs).reshape(8, 2, 4), g["chunk_coeffs"][i])


@pytest.mark.gpu
def test_fixture_on_gpu():
    from plonky3_eon_b200 import Context, GpuKzgPcs
    g = np.load(GOLD)
    ctx = Context(0)
    try:
        pcs = GpuKzgPcs.new(15, int(g["alpha"][0]), ctx=ctx)
        pd = check_against_fixture(pcs)
        for i in range(2):
            assert np.array_equal(pd[2 + i].coeffs(), g["chunk_coeffs"][i])
        for m in pd:
            m.free()
    finally:
        ctx.close()
