#!/usr/bin/env python
"""DRAM traffic of the tensor-bound conv launches of ONE train step, from `ncu --set full` raw-page CSV exports
(ncu -i X.ncu-rep --page raw --csv):   python scripts/ncu_traffic.py OUT.json TAG conv_raw.csv wgrad_raw.csv

Writes {"dram_bytes_per_step", "launches": [{kernel, us, read_mb, write_mb, tensor_pipe_pct}], "source"} -- bench.py copies
`dram_bytes_per_step` into roofline.traffic (the capture is per launch: cold L2, serialised)."""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def rows_of(path):
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if not l.startswith("=="))]
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, table):
        v = float(r[col[name]].replace(",", ""))
        return v * table.get(units[col[name]], 1.0)

    out = []
    for r in rows[2:]:
        out.append({
            "kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("tmf::", ""),
            "us": round(val(r, "gpu__time_duration.sum", TIME), 2),
            "read_mb": round(val(r, "dram__bytes_read.sum", UNIT) / 1e6, 2),
            "write_mb": round(val(r, "dram__bytes_write.sum", UNIT) / 1e6, 2),
            "tensor_pipe_pct": round(float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]), 1),
        })
    return out


def main(out, tag, *paths):
    launches = [x for p in paths for x in rows_of(p)]
    total = sum(x["read_mb"] + x["write_mb"] for x in launches) * 1e6
    with open(out, "w") as f:
        json.dump({"dram_bytes_per_step": total, "n_launches": len(launches),
                   "source": f"ncu --set full --clock-control none, {tag}: dram__bytes_read.sum + dram__bytes_write.sum "
                             f"summed over the {len(launches)} conv2.0..conv4.3 fwd/dgrad/wgrad launches of one step",
                   "launches": launches}, f, indent=1)
    print(f"{len(launches)} launches, {total / 1e9:.3f} GB per step")


if __name__ == "__main__":
    main(*sys.argv[1:])
