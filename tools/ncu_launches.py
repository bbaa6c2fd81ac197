"""Condense an `ncu --csv --metrics ...` launch list into one line per launch (or per kernel with --agg).

    python tools/ncu_launches.py gpurun_out/launches.csv [--agg] [--skip N] [--take M]
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    agg = "--agg" in sys.argv
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    take = int(sys.argv[sys.argv.index("--take") + 1]) if "--take" in sys.argv else 1 << 30
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    d = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        rec = dict(zip(hdr, r))
        d.setdefault((int(rec["ID"]), rec["Kernel Name"].split("(")[0][-40:]), {})[rec["Metric Name"]] = \
            float(rec["Metric Value"].replace(",", ""))
    items = [(k, m) for k, m in d.items() if skip <= k[0] < skip + take]
    if agg:
        a = OrderedDict()
        for (i, k), m in items:
            e = a.setdefault(k, {"n": 0})
            e["n"] += 1
            for mk, mv in m.items():
                e[mk] = e.get(mk, 0.0) + mv
        tot = sum(e["gpu__time_duration.sum"] for e in a.values())
        for k, e in a.items():
            t = e["gpu__time_duration.sum"]
            print(f"{k:42s} n={e['n']:3d} {t / 1e6:9.3f} ms {100 * t / tot:5.1f}%  "
                  f"rd {e.get('dram__bytes_read.sum', 0) / 1e9:7.2f} GB wr {e.get('dram__bytes_write.sum', 0) / 1e9:7.2f} GB")
        print(f"total {tot / 1e6:.3f} ms")
    else:
        for (i, k), m in items:
            print(i, k, " ".join(f"{mk.split('.')[0][-24:]}={mv:.4g}" for mk, mv in m.items()))


if __name__ == "__main__":
    main()
