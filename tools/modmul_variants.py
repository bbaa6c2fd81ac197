#!/usr/bin/env python
"""Montgomery-product throughput of each formulation in csrc/fp.cuh on this GPU (1e9 products/s):
word-serial CIOS vs split (Karatsuba + separate reduction) vs dedicated square, for Fr and Fq,
beside the measured IMAD peak.  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import plonky3_eon_b200 as eon  # noqa: E402

ctx = eon.Context(0)
out = {"imad_tops": ctx.imad_peak_tops(0), "imad_hi_tops": ctx.imad_peak_tops(1), "imad_wide_tops": ctx.imad_peak_tops(2)}
names = {0: "library", 1: "word_serial", 2: "split", 3: "square_plus_add", 4: "shoup_fixed_operand_lazy",
         5: "word_serial_lazy"}
for field, fname in ((0, "fr"), (1, "fq")):
    for v, vname in names.items():
        out[f"{fname}_{vname}_gmul_s"] = ctx.modmul_gmuls(field, v)
print(json.dumps(out))
