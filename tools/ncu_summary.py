#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [substring filters...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
    "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_tensor.sum",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.split(".")[0].isupper() or h.startswith("FBSP") else h
            if any(h.endswith(k) or short == k for k in KEYS) or any(x in h for x in extra) or "issue_stalled" in h and h.endswith("per_warp_active.pct"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if "issue_stalled" in h and v < 1.0:
                    continue
                print(f"   {h:110s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
