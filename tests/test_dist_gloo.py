"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in plonky3_eon_b200/dist.py.

The compute calls are served by an oracle-backed stand-in for the C ABI (this is a test of the
sharding / gather / combine logic, not of the kernels — those are covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def __init__(self, srs_wire):
        self.srs = srs_wire

    def msm_srs_range(self, scalars, first, n, ncols):
        from oracle import cport
        return cport.msm(self.srs[first:first + n], np.ascontiguousarray(scalars).reshape(n, ncols, 4), ncols=ncols)

    def g1_sum(self, points):
        from oracle import g1
        acc = None
        for p in g1.from_wire(points):
            acc = g1.add(acc, p)
        return g1.to_wire([acc])[0]


class OraclePcs:
    """commit() of GpuKzgPcs served by the C port."""

    def __init__(self, srs_wire):
        self.srs = srs_wire

    def commit(self, evaluations):
        from oracle import cport
        from plonky3_eon_b200.pcs import MatrixProverData
        cs, pd = [], []
        for dom, ev in evaluations:
            c, _ = cport.kzg_commit(ev, dom.shift, self.srs)
            cs.append(c)
            pd.append(MatrixProverData(dom, ev, 0, None))
        return cs, pd


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import cport, fr
    from plonky3_eon_b200 import dist as edist
    from plonky3_eon_b200.pcs import TwoAdicMultiplicativeCoset
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, ncols, alpha = 1000, 3, 777
        srs = cport.srs_generate(alpha, 1024)
        rng = np.random.default_rng(5)                      # same data on every rank
        sc = fr.random_wire(rng, n * ncols).reshape(n, ncols, 4)
        # index-range sharded MSM
        first, cnt = edist.index_shard(n, world, rank)
        got = edist.sharded_msm(OracleBackend(srs), sc[first:first + cnt], first, cnt, ncols)
        want = cport.msm(srs[:n], sc, ncols=ncols)
        ok_msm = bool(np.array_equal(got, want))
        # column-sharded commit of a 64 x 5 matrix (ragged: 3 + 2 columns)
        h, w = 64, 5
        ev = fr.random_wire(rng, h * w).reshape(h, w, 4)
        c0, c1 = edist.column_shard(w, world, rank)
        dom = TwoAdicMultiplicativeCoset(1, 6)
        commit, _ = edist.sharded_commit(OraclePcs(srs), dom, np.ascontiguousarray(ev[:, c0:c1]), w)
        want_c, _ = cport.kzg_commit(ev, 1, srs)
        ok_commit = bool(np.array_equal(commit, want_c))
        q.put((rank, ok_msm, ok_commit))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_helpers():
    from plonky3_eon_b200.dist import column_shard, index_shard
    for width in (0, 1, 5, 16, 17, 64):
        for world in (1, 2, 3, 8):
            spans = [column_shard(width, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == width
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert index_shard(10, 4, 3) == (8, 2)


@pytest.mark.timeout(300)
def test_world2_sharded_msm_and_commit():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_msm, ok_commit in res:
        assert ok_msm, f"rank {rank}: sharded MSM differs from the unsharded result"
        assert ok_commit, f"rank {rank}: column-sharded commitment differs"
