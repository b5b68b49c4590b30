#!/usr/bin/env python
"""Key per-kernel metrics from an exported `ncu --page raw --csv` table:  python scripts/ncu_csv.py gpurun_out/ncu_conv_raw.csv [extra-metric-prefix ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_tensor", "launch__registers_per_thread",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed_op_shfl", "sm__pipe_tensor_subpipe"]


def main(path, extra):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[start], rows[start + 1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = KEYS + extra
    for r in rows[start + 2:]:
        if len(r) < len(hdr):
            continue
        print(f"== {r[idx['Kernel Name']][:100]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for h in hdr:
            if any(h.startswith(k) for k in keys):
                print(f"   {h:84s} {r[idx[h]]:>18s} {units[idx[h]]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
