#!/usr/bin/env python
"""Probe: does splitting the 16-column commit into two 8-column halves on two streams (two contexts, two host
threads) overlap the memory-bound MSM phases of one half with the integer-bound phases of the other?
Prints ms per 16 columns for: one context x 16 columns, one context x 8 columns twice (serial), and two
contexts x 8 columns concurrently."""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import plonky3_eon_b200 as eon  # noqa: E402
from plonky3_eon_b200 import field  # noqa: E402

LOG, COLS, ALPHA = 20, 16, 12345
rows = 1 << LOG
torch.cuda.set_device(0)
rng = np.random.default_rng(1)


def synth(cols):
    a = rng.integers(0, 1 << 62, size=(rows, cols, 4), dtype=np.uint64)
    a[..., 3] &= (1 << 60) - 1
    return torch.from_numpy(a.view(np.int64)).cuda()


streams = [torch.cuda.Stream(), torch.cuda.Stream()]
ctxs = [eon.Context(0, stream=s.cuda_stream) for s in streams]
for c in ctxs:
    eon.GpuKzgPcs.new(rows - 1, ALPHA, ctx=c)
one = field.to_wire(1)
full = synth(COLS)
halves = [synth(COLS // 2), synth(COLS // 2)]


def commit(ctx, t, cols):
    out = np.zeros((cols, 8), dtype=np.uint64)
    h = C.c_uint64(0)
    ctx.call("eon_kzg_commit_dev", C.c_void_p(t.data_ptr()), LOG, cols, one, out, C.byref(h))
    ctx.call("eon_handle_free", h)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def serial16():
    commit(ctxs[0], full, COLS)


def serial8x2():
    commit(ctxs[0], halves[0], COLS // 2)
    commit(ctxs[0], halves[1], COLS // 2)


def concurrent8x2():
    th = [threading.Thread(target=commit, args=(ctxs[i], halves[i], COLS // 2)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()


res = {"one_ctx_16_cols_ms": timed(serial16), "one_ctx_8_cols_twice_ms": timed(serial8x2),
       "two_ctx_8_cols_concurrent_ms": timed(concurrent8x2)}
print(json.dumps(res))
