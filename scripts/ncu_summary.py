#!/usr/bin/env python
"""Key per-kernel metrics of an .ncu-rep (read on the CPU box):  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared", "sm__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path, extra):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = KEYS + extra
    for r in rows[2:]:
        print(f"== {r[idx['Kernel Name']][:90]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for h in hdr:
            if any(h.startswith(k) for k in keys):
                print(f"   {h:80s} {r[idx[h]]:>18s} {units[idx[h]]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
