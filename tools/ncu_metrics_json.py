#!/usr/bin/env python
"""Per-launch figures bench.py quotes next to its live timings, extracted from .ncu-rep captures (read on the CPU box):
  python tools/ncu_metrics_json.py profiles/r02_ncu_metrics.json chi2_mixed_kernel=gpurun_out/a.ncu-rep channel_stream=gpurun_out/b.ncu-rep
`channel_stream` sums every launch in its capture (the zero-fill and the tile kernel of one cha_simulate_dev call)."""
import csv
import io
import json
import subprocess
import sys

M = {"inst_executed": "smsp__inst_executed.sum", "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
     "time_us_under_ncu": "gpu__time_duration.sum",
     "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
     "fp64_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
     "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "registers": "launch__registers_per_thread"}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[H.index("Kernel Name")][:80]}
        for k, m in M.items():
            if m in H:
                v = float(r[H.index(m)].replace(",", "")); u = U[H.index(m)]
                d[k] = v * SCALE.get(u, 1.0)
        res.append(d)
    return res


def main():
    dst, out = sys.argv[1], {}
    try:
        out = json.load(open(dst))
    except Exception:
        pass
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        launches = read(path)
        if name == "channel_stream":
            d = {"launches": [l["kernel"] for l in launches]}
            for k in ("dram_read_bytes", "dram_write_bytes", "time_us_under_ncu", "inst_executed"):
                d[k] = sum(l.get(k, 0.0) for l in launches)
        else:
            d = launches[0]
        d["capture"] = path.split("/")[-1]
        out[name] = d
    out["_note"] = ("per launch, from `ncu --set full --clock-control none` captures of `python bench.py --steps 2 --warmup 3` "
                    "(default workload: benzonitrile_k1, 8192 walkers; channel-stream: 256 walkers); summaries in profiles/")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
