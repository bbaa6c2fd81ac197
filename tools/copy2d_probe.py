"""Probe: strided pinned<->device copy rates (cudaMemcpy2DAsync) for column-group pipelining."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import plonky3_eon_b200 as eon

ctx = eon.Context(0)
rows, pitch = 1 << 21, 512
host = torch.empty(rows * pitch, dtype=torch.uint8).pin_memory()
for width in (64, 128, 256, 512):
    for to_dev in (1, 0):
        ms = C.c_float()
        ctx.call("eon_bench_copy2d", C.c_void_p(host.data_ptr()), rows, width, pitch, to_dev, C.byref(ms))
        print(f"rows=2^21 width={width}B pitch={pitch}B {'H2D' if to_dev else 'D2H'}: {ms.value:.2f} ms "
              f"{rows * width / ms.value / 1e6:.1f} GB/s")
