#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")[:80]
        v = float(x["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(x["Metric Unit"], 1.0)
        e = agg.setdefault(name, [0, 0.0, x["Grid Size"], x["Block Size"]])
        e[0] += 1
        e[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"launches: {sum(v[0] for v in agg.values())}, total device time {tot / 1e3:.2f} ms (cold-cache, serialised "
          f"under ncu: compare shares, not absolutes)\n")
    print("| kernel | launches | total us | share | grid | block |")
    print("|---|---:|---:|---:|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.0005:
            continue
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[2]} | {v[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
