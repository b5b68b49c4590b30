#!/usr/bin/env python
"""Markdown table of an `ncu -i X.ncu-rep --page raw --csv` export (one row per profiled launch):
    python scripts/ncu_table.py gpurun_out/r2t_ncu_conv_raw.csv "title" > profiles/r2t_ncu_conv.md"""
import csv
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
COLS = [("us", "gpu__time_duration.sum", TIME, 1.0, "{:.1f}"),
        ("DRAM rd MB", "dram__bytes_read.sum", UNIT, 1e-6, "{:.1f}"),
        ("DRAM wr MB", "dram__bytes_write.sum", UNIT, 1e-6, "{:.1f}"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None, 1.0, "{:.0f}"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", None, 1.0, "{:.0f}"),
        ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", None, 1.0, "{:.0f}"),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", None, 1.0, "{:.0f}"),
        ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", None, 1.0, "{:.0f}"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", None, 1.0, "{:.0f}"),
        ("regs", "launch__registers_per_thread", None, 1.0, "{:.0f}"),
        ("Minst", "smsp__inst_executed.sum", None, 1e-6, "{:.2f}")]


def main(path, title):
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if not l.startswith("=="))]
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print("`ncu --set full --clock-control none` (cold L2, serialised, each kernel replayed ~40 times): compare shares and "
          "counters, not absolute times.\n")
    print("| # | kernel | grid | block | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|---|---|" + "---:|" * len(COLS))
    for n, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("tmf::", "")
        cells = []
        for _, key, table, mul, fmt in COLS:
            if key not in col or r[col[key]] in ("", "n/a"):
                cells.append("-")
                continue
            v = float(r[col[key]].replace(",", ""))
            if table:
                v *= table.get(units[col[key]], 1.0)
            cells.append(fmt.format(v * mul))
        print(f"| {n} | `{name}` | {r[col['Grid Size']]} | {r[col['Block Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
