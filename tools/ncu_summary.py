#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key SOL / pipe / stall metrics per captured launch.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x_summary.txt]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H = rows[0]
    U = rows[1]                      # units row
    data = rows[2:]
    kn = H.index("Kernel Name")
    for r in data:
        print("==", r[kn][:110])
        for w in WANT:
            if w in H:
                print(f"   {w:70s} {r[H.index(w)]} {U[H.index(w)]}")
        st = [(h, r[i]) for i, h in enumerate(H) if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")]
        def f(v):
            try:
                return float(v.replace(",", ""))
            except ValueError:
                return 0.0
        for h, v in sorted(st, key=lambda x: -f(x[1]))[:7]:
            print(f"   STALL {h:64s} {v}")
        pipes = [(h, r[i]) for i, h in enumerate(H) if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum")]
        for h, v in sorted(pipes, key=lambda x: -f(x[1]))[:8]:
            print(f"   PIPE  {h:64s} {v}")


if __name__ == "__main__":
    main(sys.argv[1])
