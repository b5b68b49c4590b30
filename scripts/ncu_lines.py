#!/usr/bin/env python
"""Warp-stall samples per CUDA source line in an .ncu-rep (needs -lineinfo and --import-source on).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep <function-substring> [top] [--sass]
"""
import csv
import io
import subprocess
import sys


def main(path, func, top, sass):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, fname, hdr, ix = None, None, None, None
    recs = {}
    seen_first = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
            continue
        if r[0] == "Function Name":
            fname = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
            continue
        if hdr is None or func not in (fname or "") or len(r) < len(hdr):
            continue
        # keep only the first launch of each function (reports hold several launches of the same kernel)
        is_line = r[0] != ""
        if not is_line and not sass:
            continue
        try:
            n = int(r[ix["# Samples"]])
        except ValueError:
            continue
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        key = (fname, fpath.split("/")[-1], r[0], r[1] if is_line else r[3])
        e = recs.setdefault(key, [0, st, r[ix["Instructions Executed"]]])
        e[0] += n
    tot = sum(v[0] for v in recs.values()) or 1
    print(f"total samples {tot}")
    for k, v in sorted(recs.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]:7d} {100 * v[0] / tot:5.1f}% {k[1]:16s}:{k[2]:>4s} {k[3].strip()[:100]:100s} inst={v[2]} {v[1]}")


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0], args[1], int(args[2]) if len(args) > 2 else 40, "--sass" in sys.argv)
