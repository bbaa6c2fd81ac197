#!/usr/bin/env python
"""profiles/traffic.json from ncu launch lists: DRAM bytes per step of the kernel families bench.py reports rooflines for.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/X_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-open --msm-log-n 0
    python tools/ncu_traffic.py commit gpurun_out/X_launches.csv [msm gpurun_out/Y_launches.csv]

One step = the launches between two consecutive k_msm_combine launches (every MSM batch ends with exactly one); the
last complete step of the list is taken (warm caches, workspace allocated).  bench.py reads the file and puts the
numbers into `roofline.traffic` / `roofline_ntt.traffic` — a value measured by ncu for THIS round's kernels, not a
constant in bench.py.
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    d = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        rec = dict(zip(hdr, r))
        m = re.search(r"\b(k_[a-z0-9_]+)", rec["Kernel Name"])   # "void eon::k_x<..>(..)" -> k_x; other kernels as they are
        e = d.setdefault(int(rec["ID"]), {"name": m.group(1) if m else rec["Kernel Name"][:40]})
        v = float(rec["Metric Value"].replace(",", ""))
        unit = rec.get("Metric Unit", "")
        if rec["Metric Name"].startswith("dram__bytes"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
        elif rec["Metric Name"].startswith("gpu__time"):
            v *= {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(unit, 1.0)
        e[rec["Metric Name"]] = v
    return list(d.values())


def last_step(ls):
    idx = [i for i, l in enumerate(ls) if l["name"] == "k_msm_combine"]
    if len(idx) < 2:
        raise SystemExit("need at least two MSM batches in the launch list")
    return ls[idx[-2] + 1: idx[-1] + 1]


def family(step, pred):
    sel = [l for l in step if pred(l["name"])]
    by = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in sel)
    ns = sum(l.get("gpu__time_duration.sum", 0) for l in sel)
    return by, ns, len(sel)


def main():
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    args = sys.argv[1:]
    for kind, path in zip(args[0::2], args[1::2]):
        step = last_step(launches(path))
        total_ns = sum(l.get("gpu__time_duration.sum", 0) for l in step)
        src = f"ncu dram__bytes_read.sum + dram__bytes_write.sum, last warm step of {os.path.basename(path)} (profiles/)"
        if kind == "commit":
            by, ns, n = family(step, lambda k: k.startswith("k_tree_bwd"))
            out["commit_tree_bwd"] = {"bytes": by, "launches": n, "ncu_ms": ns / 1e6, "share_of_step": ns / total_ns, "source": src}
            by, ns, n = family(step, lambda k: k.startswith("k_ntt_pass"))
            out["commit_ntt_passes"] = {"bytes": by, "launches": n, "ncu_ms": ns / 1e6, "share_of_step": ns / total_ns, "source": src}
        elif kind == "msm":
            by, ns, n = family(step, lambda k: k.startswith("k_tree_bwd"))
            out["msm_2p24_tree_bwd"] = {"bytes": by, "launches": n, "ncu_ms": ns / 1e6, "share_of_step": ns / total_ns, "source": src}
        elif kind == "open":
            by, ns, n = family(step, lambda k: k.startswith("k_tree_bwd"))
            out["open_tree_bwd"] = {"bytes": by, "launches": n, "ncu_ms": ns / 1e6, "share_of_step": ns / total_ns, "source": src}
        print(kind, {k: v for k, v in out.items()})
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
