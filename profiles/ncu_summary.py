#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md and bench.py cite.
Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/<name>.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "inst_executed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_branch.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("kernel:", r[name_i][:120])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>16s} {units[i]}")
        stalls = [(float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
        print("  stall cycles per issued instruction:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
