#!/usr/bin/env python
"""bench.py -- walker log-prob evals/sec (BASELINE.json metric) on the benzonitrile config.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference ...                           the reference algorithm on the host cores
  (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one vectorised log_prob(theta[nwalkers, ndim]) over the whole walker batch of one GPU.
Workload: benzonitrile, all 3718 catalog lines in 7-30 GHz, synthetic GOTHAM-like
spectrum of 2^20 channels, 8192 walkers per GPU, inference.py 5-dim layout (free source size).
`value`  : inputs resident in HBM, CUDA-event timed on the engine's stream, max over ranks.
`e2e`    : the same call through the host-buffer C-ABI (cha_log_prob): H2D of theta and D2H of log-probs inside.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "walker log-prob evals/sec"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="benzonitrile_k1")
    ap.add_argument("--walkers", type=int, default=8192, help="walkers per GPU")
    ap.add_argument("--n-chan", type=int, default=1 << 20)
    ap.add_argument("--precision", default="mixed", choices=["mixed", "fp64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batches", type=int, default=4, help="distinct theta buffers cycled through the steps")
    ap.add_argument("--cpu-sample", type=int, default=0, help="walkers in the CPU baseline sample (0: 2 x cores)")
    return ap.parse_args()


def workload_config(args, prob, extra=None):
    cfg = {"workload": f"{args.workload}: {'+'.join(c.name for c in prob.cats)} LTE log-prob, "
                       f"{prob.freq.size} channels, K={prob.spec.K} components, ndim={prob.spec.ndim}",
           "molecules": [c.name for c in prob.cats], "n_channels": int(prob.freq.size),
           "walkers_per_gpu": int(args.walkers), "ndim": int(prob.spec.ndim), "components": int(prob.spec.K),
           "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"walkers sharded x{args.gpus}, no collective"}
    if args.walkers <= 4096:
        cfg["launch"] = "CUDA-graph replay (batches <= 4096 walkers); untimed warm-up shows each of the 4 input buffers twice"
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs ~0.1 s before its first sample: the timed region must not start before it polls."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.005)

    def mark(self):
        return time.perf_counter()

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if t_begin is not None:
            # samples taken inside the timed region (+ one polling period either side, 20 ms)
            inside = [r for r in rows if t_begin - 0.02 <= r[0] <= t_end + 0.02]
            rows = inside or rows[-1:]
        for _, r in rows:
            p = [s.strip() for s in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_exps(prob, theta, eng, zcut=6.0):
    """Work per evaluation, averaged over the batch (DESIGN.md "algorithmic work"):
    reference_mask : SURVEY.md 8(d)  N_exp = K*sum_i W_i + (K+1)*C_act + 2*L + S with W_i = #channels inside the
                     reference's mask |dv - mc| < 10 dV (inference.py:52); sum_i W_i exact for every walker (device count)
    relevant       : the same formula restricted to terms that are not numerically zero: channels closer than
                     zcut sigma to a component centre (terms beyond are < exp(-zcut^2/2) = 1.5e-8 of the line peak, below fp32 resolution);
                     this is what the mixed kernel evaluates.  128-walker sample."""
    from cha1_mcmc_b200.constants import ckm
    K = prob.spec.K
    pairs = eng.count_window_pairs(theta).astype(np.float64)
    lines = []
    for m, c in enumerate(prob.cats):
        i0, i1 = c.trim_bounds(prob.spec.ll, prob.spec.ul)
        f = c.frequency[i0:i1]
        lines.append(f if prob.line_idx[m] is None else f[prob.line_idx[m]])
    f = np.sort(np.concatenate(lines))
    x = np.sort(prob.freq)
    mc, al = prob.spec.mask_centre, prob.spec.aligned_velocity

    def union_len(lo, hi):
        o = np.argsort(lo, kind="stable"); lo = lo[o]; hi = hi[o]
        end = np.maximum.accumulate(hi)
        start = np.maximum(lo, np.r_[lo[0], end[:-1]])
        return float(np.sum(np.maximum(hi - start, 0)))

    cact, cact_rel, pairs_rel = [], [], []
    for t in theta[:: max(1, len(theta) // 128)][:128]:
        dv = t[prob.spec.idx_dv]
        lo = np.searchsorted(x, f * (1 - (mc + 10 * dv) / ckm), "right")
        hi = np.searchsorted(x, f * (1 - (mc - 10 * dv) / ckm), "left")
        cact.append(union_len(lo, hi))
        los, his, npair = [], [], 0.0
        for c in range(K):
            d = t[prob.spec.idx_vlsr[c]] - al - mc
            a = max(-10 * dv, d - zcut * dv / 2.355); b = min(10 * dv, d + zcut * dv / 2.355)
            l2 = np.searchsorted(x, f * (1 - (mc + b) / ckm), "right")
            h2 = np.searchsorted(x, f * (1 - (mc + a) / ckm), "left")
            npair += float(np.sum(np.maximum(h2 - l2, 0)))
            los.append(l2); his.append(h2)
        pairs_rel.append(npair)
        cact_rel.append(union_len(np.concatenate(los), np.concatenate(his)))
    S = sum(c.state_g.size for c in prob.cats)
    L = f.size
    ref = K * pairs.mean() + (K + 1) * float(np.mean(cact)) + 2 * L + S
    rel = float(np.mean(pairs_rel)) + (K + 1) * float(np.mean(cact_rel)) + 2 * L + S
    return {"n_exp_per_eval": rel, "pairs_per_eval": float(np.mean(pairs_rel)), "c_act_per_eval": float(np.mean(cact_rel)),
            "zcut_sigma": zcut, "lines": int(L), "states": int(S),
            "reference_mask": {"n_exp_per_eval": float(ref), "pairs_per_eval": float(pairs.mean()),
                               "c_act_per_eval": float(np.mean(cact))}}


def measured_peaks():
    out = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "ex2_per_s": 148 * 16 * 1.965e9,
           "ex2_src": "nominal 16 MUFU/clk/SM x 148 SM x 1.965 GHz"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            out["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"]); out["hbm_src"] = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    p = os.path.join(ROOT, "profiles", "r01_ubench_pipes.json")
    if os.path.exists(p):
        try:
            out["ex2_per_s"] = float(json.load(open(p))["ex2_per_s"])
            out["ex2_src"] = "measured on this pool (tools/ubench_pipes.cu -> profiles/r01_ubench_pipes.json)"
        except Exception:
            pass
    return out


def ncu_traffic(kernel, args):
    """dram bytes (read + write) per launch from the committed ncu capture -- only valid for the workload it was
    taken on (the default one); null otherwise."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if args.workload != "benzonitrile_k1" or args.walkers != 8192 or args.n_chan != (1 << 20) or not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))[kernel]
        return d["dram_read_bytes"] + d["dram_write_bytes"]
    except Exception:
        return None


def cpu_port_rate(prob, theta, n_sample, threads):
    """The reference ALGORITHM (O(K*L*C) full-grid masks, MolSim over the whole catalog per call) as the C port
    oracle/lte_oracle.c, walkers farmed to `threads` host threads like emcee's pool.map."""
    from oracle import lte_oracle as O
    from oracle.c_oracle import COracle
    spec = to_oracle_spec(prob.spec)
    cats = [O.parse_catalog(c.catalog_file, name_for_q=os.path.basename(c.catalog_file).replace(".gz", "")) for c in prob.cats]
    lidx = []
    for m, c in enumerate(prob.cats):
        i0, i1 = c.trim_bounds(prob.spec.ll, prob.spec.ul)
        lidx.append(np.arange(i1 - i0) if prob.line_idx[m] is None else np.asarray(prob.line_idx[m]))
    co = COracle(spec, cats, (prob.freq, prob.y, prob.yerr, lidx), prior=(prob.prior_stds, prob.prior_means))
    th = theta[:n_sample]
    t0 = time.perf_counter()
    out = co.lnprob(th, nthreads=threads)
    dt = time.perf_counter() - t0
    return len(th) / dt, dt, out


def to_oracle_spec(s):
    from oracle import lte_oracle as O
    o = O.ModelSpec(ndim=s.ndim, K=s.K, idx_ss=list(s.idx_ss), idx_ncol=[list(r) for r in s.idx_ncol], idx_tex=s.idx_tex,
                    idx_vlsr=list(s.idx_vlsr), idx_dv=s.idx_dv, fixed_ss=s.fixed_ss, dish_size=s.dish_size,
                    aligned_velocity=s.aligned_velocity, mask_centre=s.mask_centre, planck_eps=s.planck_eps,
                    ll=s.ll, ul=s.ul, lo=s.lo, hi=s.hi, vlsr_min_sep=s.vlsr_min_sep, vlsr_max_sep=s.vlsr_max_sep)
    return o


def run_reference(args, rank, world, real_stdout):
    """--impl reference: rank 0 alone times the CPU implementation; other ranks exit 0."""
    if rank != 0:
        return
    from cha1_mcmc_b200.synthetic import default_cat_folder
    prob, theta = reference_problem(args)
    cores = os.cpu_count() or 1
    # bounded sample: one probe evaluation per host thread sizes the per-step sample so that the whole
    # --steps/--warmup run stays near two minutes (at least one walker per step)
    _, t_probe, _ = cpu_port_rate(prob, theta, cores, cores)
    if args.cpu_sample:
        n_sample = args.cpu_sample
    else:
        budget = 120.0 / max(1, args.steps + args.warmup)
        n_sample = cores * max(1, int(budget / t_probe)) if budget >= t_probe else max(1, int(cores * budget / t_probe))
        n_sample = min(n_sample, len(theta))
    for _ in range(args.warmup):
        cpu_port_rate(prob, theta, n_sample, cores)
    t_tot = 0.0
    for s in range(args.steps):
        _, dt, _ = cpu_port_rate(prob, np.roll(theta, -s * n_sample, axis=0), n_sample, cores)
        t_tot += dt
    value = args.steps * n_sample / t_tot
    sample = f"{n_sample} walkers per step x {args.steps} steps of the same workload (full 2^20-channel grid, all lines)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, prob, {"l2": "n/a (CPU)", "parallelism": f"{cores} host threads over walkers"}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(real_stdout, line)


def reference_problem(args):
    """The reference arm needs the same synthetic spectrum; its noiseless truth comes from the C oracle itself
    when no GPU is present (windowed NumPy oracle), so the arm never touches the CUDA path."""
    from cha1_mcmc_b200.synthetic import default_cat_folder, SyntheticProblem, window_grid, _trimmed_freqs
    from cha1_mcmc_b200 import synthetic as SY
    from cha1_mcmc_b200 import MolCat, ModelSpec, find_catalog
    from oracle import lte_oracle as O
    name = args.workload
    if name != "benzonitrile_k1":
        raise SystemExit("--impl reference supports the headline workload only")
    cat = MolCat("benzonitrile", find_catalog(default_cat_folder(), "benzonitrile"))
    bounds = {'source_size': [0.0, 200.0], 'Ncol': [1e8, 1e14], 'Tex': [2.7, 15.0], 'vlsr': [5.0, 6.6], 'dV': [0.05, 0.3]}
    spec = ModelSpec.inference(None, bounds, 100, 5.8, 7000, 30000)
    theta = np.array([40.0, 2.15e11, 6.7, 5.8, 0.117]); stds = np.array([4.0, 0.3e11, 0.1, 0.002, 0.002])
    freq = window_grid(_trimmed_freqs(cat, spec.ll, spec.ul), args.n_chan, SY.GOTHAM_DNU, 0.0)
    ocat = O.parse_catalog(cat.catalog_file, name_for_q="benzonitrile.cat")
    i0, i1 = cat.trim_bounds(spec.ll, spec.ul)
    truth = O.simulate(to_oracle_spec(spec), [ocat], [np.arange(i1 - i0)], freq, theta, windowed=True)
    rng = np.random.default_rng(0)
    y = truth + rng.normal(0.0, 0.005, freq.size)
    yerr = np.sqrt(0.005 ** 2 + (0.1 * y) ** 2)
    prob = SyntheticProblem(name, spec, [cat], [None], freq, y, yerr, theta, theta.copy(), stds)
    return prob, prob.walkers(max(64, 4 * (os.cpu_count() or 1)), seed=1)


def _claim_stdout():
    """Route everything libraries write to fd 1 (the NCCL version banner, nvcc notes) to stderr and keep the real stdout
    for the ONE JSON line of the contract."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_fd, line):
    sys.stdout.flush()
    os.write(real_fd, (json.dumps(line) + "\n").encode())


def main():
    args = parse_args()
    real_stdout = _claim_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cha1_mcmc_b200.build import build_library
    if rank == 0:
        build_library()
    if world > 1:
        dist.barrier()
    from cha1_mcmc_b200.synthetic import make_problem, default_cat_folder

    prob = make_problem(args.workload, default_cat_folder(), n_chan=args.n_chan, device=local, seed=0)
    eng = prob.engine(device=local, precision=args.precision)
    nw, nd = args.walkers, prob.spec.ndim
    n_batches = max(1, args.batches)
    thetas = [prob.walkers(nw, seed=1 + 1000 * rank + b) for b in range(n_batches)]
    stream = torch.cuda.ExternalStream(eng._lib.cha_stream(eng._h), device=torch.device("cuda", local))
    d_thetas = [torch.from_numpy(t).to(f"cuda:{local}") for t in thetas]
    d_out = torch.empty(nw, dtype=torch.float64, device=f"cuda:{local}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    torch.cuda.synchronize()

    def step_dev(i):
        eng.log_prob_device(d_thetas[i % n_batches], out=d_out, with_prior=True, sync=False)

    # ---- value: device-resident inputs ---------------------------------------------------------------
    # batches of <= 4096 walkers are replayed as CUDA graphs from the second sighting of a (pointer, size) pair on:
    # every one of the n_batches buffers is shown twice before the timed region so that no capture falls inside it
    n_warm = max(args.warmup, 2 * n_batches) if nw <= 4096 else args.warmup
    for i in range(n_warm):
        step_dev(i)
        if nw <= 4096:
            eng.sync()          # the timed steps sync after every call, and the pending-call slot is part of the graph key
    eng.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local); clocks.start(); clocks.wait_first()
    t_clk0 = clocks.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.stat("launches")
    fused_ns = []
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xff)                       # L2 flush, outside the timed events
            ev[i][0].record(stream)
        step_dev(i)
        with torch.cuda.stream(stream):
            ev[i][1].record(stream)
        eng.sync()
        fused_ns.append(eng.stat("fused_ns"))
    torch.cuda.synchronize()
    launches = eng.stat("launches") - launches0
    t_dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t_wall = time.perf_counter() - t_wall0
    # ---- e2e: host buffers through the C-ABI, copies inside ---------------------------------------------
    for i in range(min(args.warmup, 3)):
        eng.log_prob(thetas[i % n_batches])
    t_e2e = 0.0
    lp_host = None
    for i in range(args.steps):
        flush.fill_(i & 0xff); torch.cuda.synchronize()
        t0 = time.perf_counter()
        lp_host = eng.log_prob(thetas[i % n_batches])
        t_e2e += time.perf_counter() - t0
    clk = clocks.stop(t_clk0, clocks.mark())
    # ---- reduce over ranks: max time ----------------------------------------------------------------------
    times = torch.tensor([t_dev_ms, t_e2e * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev_ms, t_e2e_ms = [float(v) for v in times.cpu()]
    total_evals = args.steps * nw * world
    value = total_evals / (t_dev_ms * 1e-3)
    e2e_value = total_evals / (t_e2e_ms * 1e-3)

    if rank == 0:
        peaks = measured_peaks()
        alg = algorithmic_exps(prob, thetas[0], eng)
        fused_s = float(np.mean(fused_ns)) * 1e-9
        st = eng.stats()
        ach = alg["n_exp_per_eval"] * nw / fused_s
        # algorithmic bytes of the fused kernel: 24*C_act + 24*L + 8*(ndim+1) per eval, no reuse
        alg_bytes = (24 * alg["c_act_per_eval"] + 24 * alg["lines"] + 8 * (nd + 1)) * nw
        roofline = {"kernel": "chi2_mixed_kernel" if args.precision == "mixed" else "chi2_fp64_kernel",
                    "bound": "sfu", "achieved": ach / 1e9, "peak": peaks["ex2_per_s"] / 1e9, "unit": "Gexp/s",
                    "frac": ach / peaks["ex2_per_s"],
                    "traffic": ncu_traffic("chi2_mixed_kernel", args) if args.precision == "mixed" else None,
                    "traffic_unit": "bytes/launch (dram read+write, ncu)", "peak_src": peaks["ex2_src"],
                    "avg_launch_ms": fused_s * 1e3, "share_of_step": fused_s / (t_dev_ms * 1e-3 / args.steps),
                    "algorithmic": alg, "algorithmic_bytes_per_launch": alg_bytes,
                    "hbm_equiv_gbs": alg_bytes / fused_s / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_src": peaks["hbm_src"],
                    "reference_mask_equiv_gexp_s": alg["reference_mask"]["n_exp_per_eval"] * nw / fused_s / 1e9,
                    "pair_list": {"pairs": st["pairs"], "active_channels": st["active_channels"], "tiles": st["tiles"],
                                  "dv_list": st["dv_list"]}}
        # ---- channel-stream kernel (model spectra written to HBM): HBM roofline ---------------------------
        stream_roof = None
        if world == 1:
            n_sim = max(1, min(256, (1 << 31) // (8 * prob.freq.size)))          # <= 2 GiB of spectra per launch
            d_sim = torch.empty((n_sim, prob.freq.size), dtype=torch.float64, device=f"cuda:{local}")
            th_sim = d_thetas[0][:n_sim].contiguous()
            for _ in range(2):
                eng.simulate_device(th_sim, out=d_sim, sync=True)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for a_, b_ in evs:
                with torch.cuda.stream(stream):
                    a_.record(stream)
                eng.simulate_device(th_sim, out=d_sim, sync=False)
                with torch.cuda.stream(stream):
                    b_.record(stream)
                eng.sync()
            t_sim = min(a_.elapsed_time(b_) for a_, b_ in evs) * 1e-3
            sim_bytes = n_sim * prob.freq.size * 8
            stream_roof = {"kernel": "cha_simulate_dev: walker_prep + zero-fill (memset) + simulate_tiles_kernel", "bound": "hbm",
                           "achieved": sim_bytes / t_sim / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": sim_bytes / t_sim / 1e9 / peaks["hbm_gbs"], "peak_src": peaks["hbm_src"],
                           "algorithmic_bytes_per_launch": sim_bytes, "walkers": n_sim, "ms": t_sim * 1e3,
                           "traffic": None}
            del d_sim
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_sample = args.cpu_sample or 2 * cores
            rate, dt, cpu_lp = cpu_port_rate(prob, thetas[0], n_sample, cores)
            gpu_lp = eng.log_prob(thetas[0][:n_sample])
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_sample} of the {nw} walkers, full grid and line list, {dt:.1f} s on {cores} threads",
                   "max_abs_dlogp_vs_gpu": float(np.max(np.abs(cpu_lp - gpu_lp)))}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32-mufu+f64-acc" if args.precision == "mixed" else "f64", "data": "synthetic",
                "config": workload_config(args, prob), "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nw * nd * 8, "d2h_bytes_per_step": nw * 8,
                        "ms_per_step": t_e2e_ms / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_stream": stream_roof, "cpu_baseline": cpu,
                "wall_s_timed_region": t_wall, "finite_logp_frac": float(np.mean(np.isfinite(lp_host)))}
        _emit(real_stdout, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
