"""Multi-GPU sharding of the KZG hot path for ONE-PROCESS-PER-GPU callers (torch.distributed for plumbing).

A single process that owns several GPUs needs none of this: `MultiContext` (lib.py; `eon_mctx_*` in
include/eon_kzg.h, csrc/multi.cu) takes whole host matrices and shards them inside the library.  The helpers here
are the same two axes for a launcher that already runs one rank per GPU (torchrun), each a thin caller of the C ABI:
`eon_kzg_commit[_lde]_ld` on this rank's columns of the shared host matrix, and for a lone MSM
`eon_srs_set_range_tables` + `eon_msm_srs_range_partial_dev` + ncclAllGather + `eon_g1_sum_cols_dev` (partial sums
never leave the devices).

The reference has no distributed layer; the path shards along two independent axes
(SURVEY §8e), neither of which needs a data-path collective:

  * columns      every column's iDFT / LDE / MSM is independent (kzg/src/pcs.rs:244-249):
                 rank r owns columns column_shard(width, world, r); the SRS is replicated.
                 The only exchange is an all_gather of the per-column commitments (64 B each).
  * point index  one big MSM: rank r holds SRS[index_shard(n, world, r)] and the matching scalar
                 rows, produces ONE partial affine point per column; partial sums are
                 all-gathered (world x ncols x 64 B) and added on every rank
                 (EC addition is not an NCCL reduction op).

`backend` is anything with the two methods of `GpuBackend` below; tests/test_dist_gloo.py runs the
same logic at world_size 2 on CPU (gloo) with an oracle-backed stand-in, since only the
sharding/gather logic is being exercised there.
"""
import numpy as np


def column_shard(width, world, rank):
    """Contiguous, balanced column range [c0, c1) of rank `rank` (first `width % world` ranks get one more)."""
    base, extra = divmod(width, world)
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


def index_shard(n, world, rank):
    """Contiguous, balanced point-index range (first, count) of rank `rank`."""
    c0, c1 = column_shard(n, world, rank)
    return c0, c1 - c0


class GpuBackend:
    """The C-ABI calls the sharded algorithms need (see include/eon_kzg.h)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def msm_srs_range(self, scalars, first, n, ncols):
        """scalars: uint64 [n, ncols, 4] host rows [first, first+n) -> uint64 [ncols, 8]."""
        import ctypes as C
        s = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.zeros((ncols, 8), dtype=np.uint64)
        d = self.ctx.dev_alloc(max(s.nbytes, 32))
        try:
            if s.nbytes:
                self.ctx.h2d(d, s)
            self.ctx.call("eon_msm_srs_range_dev", C.c_void_p(d), first, n, ncols, ncols, out)
        finally:
            self.ctx.dev_free(d)
        return out

    def g1_sum(self, points):
        p = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8)
        out = np.zeros(8, dtype=np.uint64)
        self.ctx.call("eon_g1_sum", p, p.shape[0], out)
        return out

    # -- device path of the index-range MSM: partial sums stay in HBM between the MSM, the all_gather and the add --
    def msm_partial_device(self, scalars, first, n, ncols, device):
        """Partial sums of this rank's shard as an int64 tensor [ncols * 8] on `device` (window tables sized for
        the shard are built on first use)."""
        import ctypes as C

        import torch
        if n >= (1 << 14) and getattr(self, "_range", None) != (first, n):
            self.ctx.call("eon_srs_set_range_tables", first, n, 0)
            self._range = (first, n)
        d_sc = torch.from_numpy(np.ascontiguousarray(scalars, dtype=np.uint64).view(np.int64)).to(device)
        part = torch.zeros(ncols * 8, dtype=torch.int64, device=device)
        torch.cuda.current_stream(device).synchronize()      # the context may run on another stream
        self.ctx.call("eon_msm_srs_range_partial_dev", C.c_void_p(d_sc.data_ptr()), first, n, ncols, ncols,
                      C.c_void_p(part.data_ptr()))
        self.ctx.sync()
        return part

    def sum_cols_device(self, parts, nparts, ncols):
        """parts: int64 tensor [nparts * ncols * 8] on the device -> uint64 [ncols, 8] on the host."""
        import ctypes as C
        out = np.zeros((ncols, 8), dtype=np.uint64)
        self.ctx.call("eon_g1_sum_cols_dev", C.c_void_p(parts.data_ptr()), nparts, ncols, out)
        return out


def _all_gather_u64(arr, group=None, device=None):
    """all_gather of a uint64 numpy array (same shape on every rank) -> list of arrays, rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [o.cpu().numpy().view(np.uint64).reshape(arr.shape) for o in outs]


def sharded_msm(backend, scalars_local, first, n_local, ncols, group=None, device=None):
    """Index-range sharded MSM.  Every rank passes its own scalar rows [first, first + n_local) and
    gets the full result [ncols, 8] (identical on all ranks)."""
    if hasattr(backend, "msm_partial_device") and device is not None and getattr(device, "type", "") == "cuda":
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group)
        part = backend.msm_partial_device(scalars_local, first, n_local, ncols, device)
        allp = torch.empty(world * ncols * 8, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allp, part, group=group)    # NCCL over NVLink: world x ncols x 64 bytes
        torch.cuda.current_stream(device).synchronize()
        return backend.sum_cols_device(allp, world, ncols)
    partial = backend.msm_srs_range(scalars_local, first, n_local, ncols)
    parts = _all_gather_u64(partial, group, device)            # world x [ncols, 8]
    out = np.zeros((ncols, 8), dtype=np.uint64)
    for c in range(ncols):
        out[c] = backend.g1_sum(np.stack([p[c] for p in parts]))
    return out


def gather_column_commitments(local_commit, width, group=None, device=None):
    """Column-sharded commit: every rank committed its column_shard(); returns the [width, 8]
    commitment of the whole matrix in column order on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    widths = [column_shard(width, world, r) for r in range(world)]
    maxw = max(c1 - c0 for c0, c1 in widths)
    pad = np.zeros((maxw, 8), dtype=np.uint64)
    pad[:local_commit.shape[0]] = local_commit
    parts = _all_gather_u64(pad, group, device)
    out = np.zeros((width, 8), dtype=np.uint64)
    for r, (c0, c1) in enumerate(widths):
        out[c0:c1] = parts[r][:c1 - c0]
    return out


def sharded_commit(pcs, domain, evals_local, width, group=None, device=None):
    """KzgPcs::commit of one h x width matrix whose columns are sharded: `evals_local` holds this
    rank's columns column_shard(width, world, rank) as [h, local_w, 4].  Returns
    (commitment [width, 8] on every rank, this rank's MatrixProverData)."""
    commit, pdata = pcs.commit([(domain, evals_local)])
    return gather_column_commitments(commit[0], width, group, device), pdata[0]


class RowBlockCommitLde:
    """Host-buffer `commit` + hinted LDE of ONE row-major host matrix over N ranks that own FEW columns each.

    With 2 columns per rank (16 trace columns over 8 GPUs) the strided copies of the `_ld` entry points move
    64-byte row pieces, and the LDE download — the long pole of the step, and the transfer the host serves slowest
    when every GPU copies at once — reaches two thirds of the contiguous rate at best.  Here every rank moves
    CONTIGUOUS row blocks over PCIe and the column <-> row exchange happens over NVLink (two `all_to_all`s, the path's
    one real exchange step, SURVEY §8e "row-block H2D + NVLink all-to-all"):

        host rows [r h/N, (r+1) h/N) x W  --H2D-->  pack by destination  --all_to_all-->  all rows of my W/N columns
        coset LDE of my columns (queued first: its way back is the long pole)
        LDE [2h x W/N]  --all_to_all (row chunks)-->  my row block of every column shard  --unpack-->  --D2H--> host rows
        KzgPcs::commit of my columns (coset iDFT + MSM) meanwhile, on the main stream

    Results are the same bytes as `eon_kzg_commit_lde` on the whole matrix: every column goes through the same
    `eon_coset_lde_batch_dev` / `eon_kzg_commit_dev` calls, only the transport differs.
    """

    def __init__(self, ctx, log_rows, cols_total, added_bits, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.ctx, self.group, self.device = torch, dist, ctx, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        rows = 1 << log_rows
        if rows % self.world or cols_total % self.world:
            raise ValueError("rows and columns must divide evenly over the ranks")
        self.log_rows, self.ab, self.W = log_rows, added_bits, cols_total
        self.w, self.rb, self.lrb = cols_total // self.world, rows // self.world, (rows << added_bits) // self.world
        i64 = torch.int64
        self.d_blk = torch.empty((self.rb, cols_total, 4), dtype=i64, device=device)
        self.d_send = torch.empty((self.world, self.rb, self.w, 4), dtype=i64, device=device)
        self.d_evals = torch.empty((rows, self.w, 4), dtype=i64, device=device)
        self.d_lde = torch.empty((rows << added_bits, self.w, 4), dtype=i64, device=device)
        self.d_lrecv = torch.empty((self.world, self.lrb, self.w, 4), dtype=i64, device=device)
        self.d_lout = torch.empty((self.lrb, cols_total, 4), dtype=i64, device=device)
        self.side = torch.cuda.Stream(device)
        self.ev = torch.cuda.Event()

    def step(self, host_evals, host_lde, shift_wire, lde_shift_wire, commits_out):
        """host_evals / host_lde: pinned int64 tensors [rows, W, 4] / [rows << ab, W, 4] (the whole matrices);
        commits_out: uint64 numpy [W/N, 8].  Returns the prover-data handle of this rank's columns."""
        import ctypes as C
        torch, dist = self.torch, self.dist
        main = torch.cuda.current_stream(self.device)
        r0 = self.rank * self.rb
        self.d_blk.copy_(host_evals[r0:r0 + self.rb], non_blocking=True)                       # contiguous H2D
        self.d_send.copy_(self.d_blk.view(self.rb, self.world, self.w, 4).permute(1, 0, 2, 3))  # pack by destination
        dist.all_to_all_single(self.d_evals.view(-1), self.d_send.view(-1), group=self.group)   # rows of MY columns
        self.ctx.call("eon_coset_lde_batch_dev", C.c_void_p(self.d_evals.data_ptr()), C.c_void_p(self.d_lde.data_ptr()),
                      self.log_rows, self.w, self.ab, lde_shift_wire)                            # queued, not synchronised
        self.ev.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev)
            dist.all_to_all_single(self.d_lrecv.view(-1), self.d_lde.view(-1), group=self.group)  # row chunks out
            self.d_lout.copy_(self.d_lrecv.permute(1, 0, 2, 3).reshape(self.lrb, self.W, 4))     # [row][shard][col]
            l0 = self.rank * self.lrb
            host_lde[l0:l0 + self.lrb].copy_(self.d_lout, non_blocking=True)                     # contiguous D2H
        h = C.c_uint64(0)
        self.ctx.call("eon_kzg_commit_dev", C.c_void_p(self.d_evals.data_ptr()), self.log_rows, self.w, shift_wire,
                      commits_out, C.byref(h))                                                   # coset iDFT + MSM
        main.wait_stream(self.side)
        return h

    def shard_checksums(self):
        """Sum (mod 2^64) of every destination rank's row chunk of this rank's LDE columns: what the receiving rank
        must find in its host block (parity of the transport)."""
        return self.d_lde.view(self.world, -1).sum(dim=1)
