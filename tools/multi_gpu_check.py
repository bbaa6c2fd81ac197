#!/usr/bin/env python
"""Multi-GPU parity check of the two sharding axes on real GPUs, one process per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py [--log-n 18]

  1. point-index sharded MSM (NCCL all_gather of the per-rank partial sums, then eon_g1_sum) against
     the discrete-log shortcut  sum_i c_i * (alpha^i G) = (sum_i c_i alpha^i) G  computed by the oracle;
  2. column-sharded KzgPcs::commit (all_gather of the commitments) against the same commit done whole
     on every rank.
Rank 0 prints one JSON line; any mismatch exits non-zero.  tests/test_gpu_multi.py runs this under
torchrun when the box has >= 2 GPUs.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=18)
    ap.add_argument("--cols", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import plonky3_eon_b200 as eon
    from oracle import fr, g1
    from plonky3_eon_b200 import dist as edist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = eon.Context(local)
    alpha = 12345
    n, ncols = 1 << args.log_n, args.cols
    pcs = eon.GpuKzgPcs.new(n - 1, alpha, ctx=ctx)          # SRS replicated on every GPU
    backend = edist.GpuBackend(ctx)

    # ---- 1. index-range sharded MSM -----------------------------------------------------------
    rng = np.random.default_rng(99)                         # same scalars on every rank
    scw = fr.random_wire(rng, n * ncols).reshape(n, ncols, 4)
    first, cnt = edist.index_shard(n, world, rank)
    got = edist.sharded_msm(backend, scw[first:first + cnt], first, cnt, ncols, device=dev)
    ok_msm = True
    if rank == 0:
        sc = fr.from_wire(scw.reshape(-1, 4))
        for c in range(ncols):
            acc, ap_ = 0, 1
            for i in range(n):
                acc = (acc + sc[i * ncols + c] * ap_) % fr.P
                ap_ = ap_ * alpha % fr.P
            want = g1.mul(g1.G, acc)
            ok_msm &= g1.from_wire(got[c].reshape(1, 8))[0] == want

    # ---- 2. column-sharded commit ----------------------------------------------------------------
    h, width = 1 << 12, 8
    evw = fr.random_wire(np.random.default_rng(7), h * width).reshape(h, width, 4)
    dom = eon.TwoAdicMultiplicativeCoset(1, 12)
    c0, c1 = edist.column_shard(width, world, rank)
    local_ev = np.ascontiguousarray(evw[:, c0:c1])
    sharded, _ = edist.sharded_commit(pcs, dom, local_ev, width, device=dev)
    whole, _ = pcs.commit([(dom, evw)])
    ok_commit = bool(np.array_equal(sharded, whole[0]))

    flag = torch.tensor([int(ok_msm and ok_commit)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "msm_points": n, "msm_cols": ncols, "sharded_msm_ok": bool(ok_msm),
                          "sharded_commit_ok": ok_commit, "all_ranks_ok": bool(flag.item())}))
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
