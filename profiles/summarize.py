#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics the
roofline argument needs.  Usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_requests_srcunit_tex_op_red.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {r[i]:>18s} {units[i]}")
        extra = [h for h in hdr if "tensor" in h and h not in KEYS and "pct" in h]
        for h in extra[:12]:
            i = hdr.index(h)
            print(f"{h:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
