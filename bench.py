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
Further blocks of the same JSON line (what an MCMC run sees):
`sampler`        : the resident stretch-move sampler in steady state (>= 60 steps in), walkers sharded over the ranks,
                   the per-half-step all-gather of positions (NCCL, enqueued by the engine) INSIDE the timed region
`posterior_batch`: log_prob of a theta batch taken from that chain (the wide-list regime of a running sampler)
`sustained`      : >= 2 s of back-to-back steps with clocks/power sampled inside
`fp64`           : the same workload through the all-fp64 kernels (the reference's arithmetic)
Other workloads: --workload {hc5n_dsn, hc7n_hfs_k4, benzonitrile_k4, joint_k4, survey}; --mode sampler makes the
resident sampler the headline `value` (config 4: --workload joint_k4 --mode sampler --scaling strong --walkers 65536).
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
    ap.add_argument("--mode", default="logprob", choices=["logprob", "sampler"],
                    help="what the headline `value` times: vectorised log_prob calls, or resident-sampler steps")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --walkers per GPU; strong: --walkers is the global ensemble, divided over the ranks")
    ap.add_argument("--sampler-steps", type=int, default=40)
    ap.add_argument("--sampler-burn", type=int, default=60, help="untimed steps before the sampler block is timed")
    ap.add_argument("--sustained-s", type=float, default=2.0)
    ap.add_argument("--no-extras", action="store_true", help="skip the sampler/posterior/sustained/fp64/stream blocks")
    ap.add_argument("--walkers-total", type=int, default=262144, help="--workload survey: walkers over all fits")
    return ap.parse_args()


def workload_config(args, prob):
    """Names the workload only -- identical in both arms (how a run was executed goes into `run`)."""
    return {"workload": f"{args.workload}: {'+'.join(c.name for c in prob.cats)} LTE log-prob, "
                        f"{prob.freq.size} channels, K={prob.spec.K} components, ndim={prob.spec.ndim}",
            "molecules": [c.name for c in prob.cats], "n_channels": int(prob.freq.size),
            "walkers": int(args.walkers), "walkers_are": "per GPU" if args.scaling == "weak" else "global ensemble",
            "ndim": int(prob.spec.ndim), "components": int(prob.spec.K)}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        self.index = index; self.rows = []; self.proc = None; self.period_ms = int(period_ms)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if t_begin is not None:
            # samples taken inside the timed region (+ one polling period either side, 20 ms)
            pad = self.period_ms * 1e-3
            inside = [r for r in rows if t_begin - pad <= r[0] <= t_end + pad]
            rows = inside or rows[-1:]
        for _, r in rows:
            p = [s.strip() for s in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            try:
                pw.append(float(p[2]))
            except ValueError:
                pass
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w_median": float(np.median(pw)) if pw else None, "power_w_max": max(pw) if pw else None}


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


def ncu_metrics(kernel, args):
    """Per-launch figures of `kernel` from the committed `ncu --set full` capture of this command (profiles/):
    dram bytes, warp instructions executed, pipe utilisations.  Only valid for the workload the capture was taken on
    (the default one); {} otherwise."""
    if args.workload != "benzonitrile_k1" or args.walkers != 8192 or args.n_chan != (1 << 20) or args.precision != "mixed":
        return {}
    for name in ("r02_ncu_metrics.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                d = dict(json.load(open(p))[kernel])
                d["source"] = "profiles/" + name
                if "dram_read_bytes" in d:
                    d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
                return d
            except Exception:
                continue
    return {}


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
            "config": workload_config(args, prob),
            "run": {"l2": "n/a (CPU)", "parallelism": f"{cores} host threads over walkers", "mode": "logprob"},
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


class DeviceTimer:
    """CUDA-event pairs on the ENGINE's stream (torch.cuda.Event only sees the stream it is recorded on)."""

    def __init__(self, torch, stream):
        self.torch, self.stream, self.pairs = torch, stream, []

    def begin(self):
        e = self.torch.cuda.Event(enable_timing=True)
        with self.torch.cuda.stream(self.stream):
            e.record(self.stream)
        self._b = e

    def end(self):
        e = self.torch.cuda.Event(enable_timing=True)
        with self.torch.cuda.stream(self.stream):
            e.record(self.stream)
        self.pairs.append((self._b, e))

    def total_ms(self):
        return sum(a.elapsed_time(b) for a, b in self.pairs)

    def each_ms(self):
        return [a.elapsed_time(b) for a, b in self.pairs]


def max_over_ranks(torch, dist, world, local, values):
    t = torch.tensor(list(values), dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def sampler_block(torch, dist, args, prob, eng, stream, rank, world, local, nw_local, steps, warmup):
    """Resident stretch-move sampler in steady state: `--sampler-burn` (>= 60) untimed steps, `warmup` more, then
    `steps` steps timed with CUDA events on the engine stream.  Walkers sharded over the ranks; the all-gather of
    positions (one per half-step) is enqueued by the engine and lies inside the timed region.  One stretch-move step
    evaluates every walker once."""
    from cha1_mcmc_b200.sampler import DeviceEnsembleSampler, shard_range
    nwg = nw_local * world
    p0 = prob.walkers(nwg, seed=11)
    w0, w1 = shard_range(nwg, world, rank)
    smp = DeviceEnsembleSampler(eng, nwg, p0[w0:w1], w0=w0, seed=5, dist=dist if world > 1 else None)
    burn = max(args.sampler_burn, 60)
    smp.step(burn + warmup)
    eng.sync(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    st0 = eng.stats()
    tm = DeviceTimer(torch, stream)
    t_host = time.perf_counter()
    tm.begin()
    smp.step(steps)
    t_queue = time.perf_counter() - t_host
    # the synchronisation point validates the queued half-steps and runs again whatever the lists had not covered: it
    # belongs inside the timed region (the end event is recorded once the stream is idle)
    eng.sync()
    tm.end()
    torch.cuda.synchronize()
    ms, = max_over_ranks(torch, dist, world, local, [tm.total_ms()])
    st1 = eng.stats()
    coords, lp, nacc = smp.state()
    nd = prob.spec.ndim
    blk = {"value": nwg * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "burn_in_steps": burn + warmup, "walkers_global": nwg, "walkers_per_gpu": nw_local,
           "evals_per_step": nwg,
           "collective": ("ncclAllGather of positions, one per half-step, enqueued by the engine on its own stream"
                          if world > 1 else "none (one rank: the resident positions are the ensemble)"),
           "collectives_in_timed_region": st1["collectives"] - st0["collectives"],
           "collective_bytes_per_step": (st1["collective_bytes"] - st0["collective_bytes"]) / steps,
           "timing": "CUDA events on the engine stream around the queued steps, max over ranks",
           "host_queue_ms_per_step": 1e3 * t_queue / steps,
           "launches_per_step": (st1["launches"] - st0["launches"]) / steps,
           "list_rebuilds_in_timed_region": st1["rebuilds"] - st0["rebuilds"],
           "narrow_list_builds_in_timed_region": st1["tight_builds"] - st0["tight_builds"],
           "list_build_ms_in_timed_region": (st1["build_us"] - st0["build_us"]) * 1e-3,
           "uncovered_events_in_timed_region": st1["uncovered_events"] - st0["uncovered_events"],
           "sync_points_in_timed_region": st1["syncs"] - st0["syncs"],
           "half_steps_rerun_in_timed_region": st1["reruns"] - st0["reruns"],
           "acceptance_rank0": nacc / (nw_local * (burn + warmup + steps)),
           "half_steps_replayed_as_graphs": st1["graph_launches"] - st0["graph_launches"],
           "fused_ms_last_half_step": st1["fused_ns"] * 1e-6 if st1["graph_launches"] == st0["graph_launches"] else None,
           "lists": {k: st1[k] for k in ("pairs", "active_channels", "tiles", "records", "dv_list", "hv_list",
                                         "tight_pairs", "tight_tiles", "tight_hv", "tight_builds")},
           "all_finite_rank0": bool(np.all(np.isfinite(lp)))}
    return blk, coords


def run_survey(torch, dist, args, rank, world, local, real_stdout):
    """--workload survey: BASELINE config 5 (every shipped catalog x {DSN-like, GOTHAM-like} fit, walkers split evenly
    over the fits, whole fits sharded over the ranks by cost; strong scaling, no collective)."""
    from cha1_mcmc_b200.synthetic import default_cat_folder
    from cha1_mcmc_b200 import survey as SV
    folder = default_cat_folder()
    mols = SV.list_molecules(folder)
    t_setup = time.perf_counter()
    probs = [SV.survey_problem(m, k, folder, device=local, seed=7) for m in mols for k in SV.KINDS]
    costs = [SV.fit_cost(p) for p in probs]
    per_fit = args.walkers_total // len(probs)
    # whole fits are too coarse a unit at 8 ranks (one fit alone is a seventh of the survey): the expensive ones are cut
    # into walker blocks that land on different ranks
    mine = SV.shard_fit_walkers(costs, world, per_fit)[rank]
    sv = SV.MoleculeSurvey([probs[i] for i, _, _ in mine], per_fit, device=local, precision=args.precision,
                           ranges=[(a, b) for _, a, b in mine], seeds=[1 + i for i, _, _ in mine])
    t_setup = time.perf_counter() - t_setup
    for _ in range(max(args.warmup, 3)):
        sv.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local); clocks.start(); clocks.wait_first()
    t_clk0 = clocks.mark()
    l0 = sum(f.eng.stat("launches") for f in sv.fits)
    # every fit runs on its own (non-blocking) stream, so no single stream's events bracket a pass: host clock with a
    # device synchronisation on both sides
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sv.step(sync=False)
        sv.sync()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clk = clocks.stop(t_clk0, clocks.mark())
    launches = sum(f.eng.stat("launches") for f in sv.fits) - l0
    list_build_s = 1e-6 * sum(f.eng.stat("build_us") for f in sv.fits)     # host time in the list builder, all fits of this rank
    ms, = max_over_ranks(torch, dist, world, local, [wall_ms])
    finite = all(bool(torch.isfinite(f.out).all()) for f in sv.fits)
    n_mine, = [sv.n_evals]
    tot = torch.tensor([float(n_mine)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tot)
    if rank == 0:
        n_eval = per_fit * len(probs)
        assert int(tot.item()) == n_eval, (int(tot.item()), n_eval)      # every walker of every fit is on exactly one rank
        line = {"metric": METRIC, "value": n_eval * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32-mufu+f64-acc" if args.precision == "mixed" else "f64",
                "data": "synthetic",
                "config": {"workload": f"survey (config 5): {len(mols)} catalogs x (DSN-like 30.5 kHz 18-25 GHz K=1, GOTHAM-like "
                                       f"1.4 kHz 7-30 GHz K=4) = {len(probs)} fits, {per_fit} walkers per fit",
                           "walkers": n_eval, "walkers_are": "global, split evenly over the fits",
                           "channels_total": int(sum(p.freq.size for p in probs)),
                           "lines_total": int(sum(p.line_idx[0].size for p in probs))},
                "run": {"parallelism": f"fits sharded over {world} GPU(s) by (line, channel) pair count, the expensive ones cut "
                                       f"into walker blocks; no collective",
                        "timing": "host clock around the passes (every fit runs on its own stream), max over ranks",
                        "fit_pieces_on_rank0": len(mine), "setup_s": round(t_setup, 2),
                        "setup_note": "synthetic problems (catalog parsing, fp64 truth spectra, noise) + handles; of which "
                                      "host list building: list_build_s", "list_build_s": round(list_build_s, 4)},
                "clocks": clk, "gpu_launches": int(launches), "all_finite": finite}
        _emit(real_stdout, line)
    sv.close()


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
    if args.workload == "survey":
        run_survey(torch, dist, args, rank, world, local, real_stdout)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    from cha1_mcmc_b200.synthetic import make_problem, default_cat_folder

    prob = make_problem(args.workload, default_cat_folder(), n_chan=args.n_chan, device=local, seed=0)
    eng = prob.engine(device=local, precision=args.precision)
    nw = args.walkers if args.scaling == "weak" else max(2, args.walkers // world)
    nd = prob.spec.ndim
    dev = f"cuda:{local}"
    stream = torch.cuda.ExternalStream(eng._lib.cha_stream(eng._h), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_batches = max(1, args.batches)
    thetas = [prob.walkers(nw, seed=1 + 1000 * rank + b) for b in range(n_batches)]
    d_thetas = [torch.from_numpy(t).to(dev) for t in thetas]
    d_out = torch.empty(nw, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    line_extra, run_info = {}, {}

    def step_dev(i, bufs=None, eng=eng):
        eng.log_prob_device((bufs or d_thetas)[i % len(bufs or d_thetas)], out=d_out, with_prior=True, sync=False, wait_torch=False)

    def timed_logprob(bufs, steps, warm, eng=eng, stream=stream):
        """`warm` untimed steps, then `steps` steps, each bracketed by CUDA events on the engine stream, with a 256 MiB
        write between them (L2 flush, outside the events).  Returns (ms total, fused-kernel ns per step, launches)."""
        n_warm = max(warm, 2 * len(bufs)) if nw <= 4096 else warm      # small batches: graph capture outside the timing
        for i in range(n_warm):
            step_dev(i, bufs, eng)
            eng.sync()          # the timed steps sync after every call (the pending-call slot is part of the graph key), and
                                # the lists settle on this batch's extent before the timing starts
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tm = DeviceTimer(torch, stream)
        l0 = eng.stat("launches")
        fused = []
        for i in range(steps):
            with torch.cuda.stream(stream):
                flush.fill_(i & 0xff)                   # L2 flush, outside the timed events
            tm.begin()
            step_dev(i, bufs, eng)
            tm.end()
            eng.sync()
            fused.append(eng.stat("fused_ns"))
        torch.cuda.synchronize()
        timed_logprob.last_each_ms = tm.each_ms()
        return tm.total_ms(), fused, eng.stat("launches") - l0

    clocks = ClockSampler(local); clocks.start(); clocks.wait_first()
    t_clk0 = clocks.mark()
    t_wall0 = time.perf_counter()
    smp_blk = post_coords = None
    if args.mode == "logprob":
        t_dev_ms, fused_ns, launches = timed_logprob(d_thetas, args.steps, args.warmup)
        t_wall = time.perf_counter() - t_wall0
        # ---- e2e: host buffers through the C-ABI, copies inside --------------------------------------------
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
        t_dev_ms, t_e2e_ms = max_over_ranks(torch, dist, world, local, [t_dev_ms, t_e2e * 1e3])
        total_evals = args.steps * nw * world
        value = total_evals / (t_dev_ms * 1e-3)
        e2e = {"value": total_evals / (t_e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nw * nd * 8,
               "d2h_bytes_per_step": nw * 8, "ms_per_step": t_e2e_ms / args.steps}
        finite_frac = float(np.mean(np.isfinite(lp_host)))
        run_info = {"mode": "logprob: one vectorised log_prob(theta[walkers, ndim]) per step and GPU",
                    "l2": "flushed between timed steps (256 MiB write)",
                    "parallelism": f"walkers sharded x{world}, no collective on this path (see `sampler` for the path with one)"}
        if nw <= 4096:
            run_info["launch"] = "CUDA-graph replay (batches <= 4096 walkers); untimed warm-up shows each input buffer twice"
    else:
        # ---- headline = resident sampler steps (the path with the collective) -------------------------------
        smp_blk, post_coords = sampler_block(torch, dist, args, prob, eng, stream, rank, world, local, nw, args.steps, args.warmup)
        t_wall = time.perf_counter() - t_wall0
        clk = clocks.stop(t_clk0, clocks.mark())
        value = smp_blk["value"]; t_dev_ms = smp_blk["ms_per_step"] * args.steps
        launches = int(round(smp_blk["launches_per_step"] * args.steps)); fused_ns = [eng.stat("fused_ns")]
        # e2e of a resident sampler: the chain comes back to the host every step (the reference saves it every step,
        # inference.py:462): steps with store_every=1 and the D2H of the stored rows inside the timed region
        from cha1_mcmc_b200.sampler import DeviceEnsembleSampler
        smp_tmp_steps = max(1, min(args.steps, 64))
        # (one untimed pass first: the chain store in HBM and the host arrays are allocated on first use)
        first = eng.sampler_chain_len()
        eng.sampler_run(10 ** 6, smp_tmp_steps, 1)
        eng.sampler_chain_read(first)
        eng.sampler_chain_clear()
        eng.sync()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        first = eng.sampler_chain_len()
        eng.sampler_run(10 ** 6 + smp_tmp_steps, smp_tmp_steps, 1)
        c_host, _ = eng.sampler_chain_read(first)
        t_e2e = time.perf_counter() - t0
        eng.sampler_chain_clear()
        t_e2e_ms, = max_over_ranks(torch, dist, world, local, [t_e2e * 1e3])
        e2e = {"value": smp_tmp_steps * nw * world / (t_e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": nw * (nd + 1) * 8, "ms_per_step": t_e2e_ms / smp_tmp_steps,
               "note": "resident sampler: positions never travel to the device after init; every step's chain row "
                       "(positions + log-probs) is copied back to the host inside the timed region"}
        finite_frac = 1.0 if smp_blk["all_finite_rank0"] else 0.0
        run_info = {"mode": "sampler: one stretch-move step of the resident ensemble per step (every walker evaluated once)",
                    "l2": "not flushed: the lists (a few MB) are meant to stay L2-resident between half-steps",
                    "parallelism": f"walkers sharded x{world}; {smp_blk['collective']}",
                    "burn_in_steps": smp_blk["burn_in_steps"]}

    # ---- further blocks: what an MCMC run sees --------------------------------------------------------------
    if not args.no_extras and args.sustained_s > 0 and args.mode == "logprob":
        # (before the sampler block: the handle's lists are still the ones the log-prob calls settled on)
        ms_est = max(t_dev_ms / max(args.steps, 1), 0.02)
        n_sus = int(min(200000, max(16, args.sustained_s * 1e3 / ms_est * 1.05)))
        for i in range(4):
            step_dev(i)
        eng.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # (polled every 100 ms here: one nvidia-smi query per rank every 20 ms over seconds is itself a load on the driver)
        c2 = ClockSampler(local, period_ms=100); c2.start(); c2.wait_first()
        tb = c2.mark()
        tm = DeviceTimer(torch, stream)
        tm.begin()
        for i in range(n_sus):
            step_dev(i)                       # queued back to back; the engine validates every 256 calls
        tm.end()
        eng.sync(); torch.cuda.synchronize()
        clk2 = c2.stop(tb, c2.mark())
        ms_sus, = max_over_ranks(torch, dist, world, local, [tm.total_ms()])
        line_extra["sustained"] = {"value": n_sus * nw * world / (ms_sus * 1e-3), "unit": UNIT, "seconds": ms_sus * 1e-3,
                                   "steps": n_sus, "ms_per_step": ms_sus / n_sus, "clocks": clk2,
                                   "l2": "not flushed (back-to-back steps over the cycled theta buffers)"}
    if not args.no_extras and args.mode == "logprob":
        ss, sw = max(4, args.sampler_steps), 3
        smp_blk, post_coords = sampler_block(torch, dist, args, prob, eng, stream, rank, world, local, nw, ss, sw)
    if not args.no_extras and post_coords is not None:
        # log_prob of a batch drawn from the running chain: the lists must cover the spread of a live ensemble
        # (a fresh handle: what an emcee-style host sampler calling log_prob on its live ensemble sees -- the resident
        # sampler's handle keeps lists sized for ITS proposals, which is not the state such a caller would be in)
        d_post = [torch.from_numpy(np.ascontiguousarray(post_coords)).to(dev)]
        eng_p = prob.engine(device=local, precision=args.precision)
        stream_p = torch.cuda.ExternalStream(eng_p._lib.cha_stream(eng_p._h), device=torch.device("cuda", local))
        rb0 = eng_p.stat("rebuilds")
        t_ms, f_ns, _ = timed_logprob(d_post, max(5, args.steps), 6, eng=eng_p, stream=stream_p)
        each = sorted(timed_logprob.last_each_ms)
        t_ms, = max_over_ranks(torch, dist, world, local, [t_ms])
        st = eng_p.stats()
        line_extra["posterior_batch"] = {
            "value": max(5, args.steps) * nw * world / (t_ms * 1e-3), "unit": UNIT, "ms_per_step": t_ms / max(5, args.steps),
            "ms_per_step_median_rank0": each[len(each) // 2], "ms_per_step_max_rank0": each[-1],
            "list_rebuilds_incl_warmup": st["rebuilds"] - rb0, "reach_ordered_batches": st["sorted_batches"],
            "theta": f"positions of the resident chain after {smp_blk['burn_in_steps'] + smp_blk['steps']} steps",
            "fused_ms": float(np.mean(f_ns)) * 1e-6,
            "lists": {k: st[k] for k in ("pairs", "active_channels", "tiles", "records", "dv_list", "hv_list")}}
        eng_p.close()
    if not args.no_extras and args.precision == "mixed" and world == 1:
        # the same workload through the all-fp64 kernels (reference operation order, full 10 dV masks)
        eng64 = prob.engine(device=local, precision="fp64")
        d64 = torch.empty(nw, dtype=torch.float64, device=dev)
        s64 = torch.cuda.ExternalStream(eng64._lib.cha_stream(eng64._h), device=torch.device("cuda", local))
        eng64.log_prob_device(d_thetas[0], out=d64, sync=True)
        tm = DeviceTimer(torch, s64)
        n64 = 2
        for i in range(n64):
            tm.begin()
            eng64.log_prob_device(d_thetas[i % n_batches], out=d64, sync=False, wait_torch=False)
            tm.end()
            eng64.sync()
        lp64 = d64.cpu().numpy()
        eng.log_prob_device(d_thetas[(n64 - 1) % n_batches], out=d_out, sync=True)
        lpmx = d_out.cpu().numpy()
        m = np.isfinite(lp64) & np.isfinite(lpmx)
        line_extra["fp64"] = {"value": n64 * nw / (tm.total_ms() * 1e-3), "unit": UNIT, "ms_per_step": tm.total_ms() / n64,
                              "kernel": "chi2_fp64_kernel + line_tau_kernel<double>", "steps": n64,
                              "fused_ms": eng64.stat("fused_ns") * 1e-6,
                              "max_abs_dlogp_mixed_vs_fp64": float(np.max(np.abs(lp64[m] - lpmx[m]))) if m.any() else None,
                              "walkers_compared": int(m.sum())}
        eng64.close()

    if rank == 0:
        peaks = measured_peaks()
        alg = algorithmic_exps(prob, thetas[0], eng)
        fused_s = float(np.mean(fused_ns)) * 1e-9
        st = eng.stats()
        kname = "chi2_mixed_kernel" if args.precision == "mixed" else "chi2_fp64_kernel"
        ncu = ncu_metrics(kname, args) if args.mode == "logprob" else {}
        n_eval_launch = nw if args.mode == "logprob" else nw // 2
        exp_rate = alg["n_exp_per_eval"] * n_eval_launch / fused_s
        gauss_rate = alg["pairs_per_eval"] * n_eval_launch / fused_s
        issue_peak = 148 * 4 * (clk.get("sm_mhz") or 1965.0) * 1e6          # warp instructions / s: 4 schedulers per SM
        inst = ncu.get("inst_executed")
        # algorithmic bytes of the fused kernel: 24*C_act + 24*L + 8*(ndim+1) per eval, no reuse
        alg_bytes = (24 * alg["c_act_per_eval"] + 24 * alg["lines"] + 8 * (nd + 1)) * n_eval_launch
        roofline = {"kernel": kname, "bound": "issue",
                    "achieved": (inst / fused_s / 1e9) if inst else None, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s",
                    "frac": (inst / fused_s / issue_peak) if inst else None,
                    "peak_src": "148 SMs x 4 warp schedulers x 1 instruction/clk at the SM clock sampled during the run",
                    "inst_per_eval": (inst * 32.0 / n_eval_launch) if inst else None,
                    "inst_src": ncu.get("source"),
                    "ncu_pipes_pct": {k: ncu[k] for k in ("xu_pct", "fma_pct", "fp64_pct", "alu_pct", "lsu_pct", "issue_active_pct") if k in ncu},
                    "exp": {"unit": "Gexp/s", "peak": peaks["ex2_per_s"] / 1e9, "peak_src": peaks["ex2_src"],
                            "algorithmic": exp_rate / 1e9, "algorithmic_frac": exp_rate / peaks["ex2_per_s"],
                            "mufu_needed": gauss_rate / 1e9, "mufu_needed_frac": gauss_rate / peaks["ex2_per_s"],
                            "w_i_definition": "relevant: channels within 6 sigma of a component centre (what the mixed kernel "
                                              "evaluates); the Gaussian terms are the only ones that reach MUFU.EX2 -- "
                                              "1-exp(-tau) is a polynomial and Planck is one exponential per tile. "
                                              "SURVEY 8(d)'s literal 10 dV mask count is under algorithmic.reference_mask",
                            "reference_mask_equiv": alg["reference_mask"]["n_exp_per_eval"] * n_eval_launch / fused_s / 1e9},
                    "traffic": ncu.get("dram_bytes"), "traffic_unit": "bytes/launch (dram read+write, ncu)",
                    "avg_launch_ms": fused_s * 1e3, "share_of_step": fused_s / (t_dev_ms * 1e-3 / args.steps) if args.mode == "logprob" else None,
                    "evals_per_launch": n_eval_launch,
                    "algorithmic": alg, "algorithmic_bytes_per_launch": alg_bytes,
                    "hbm_equiv_gbs": alg_bytes / fused_s / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_src": peaks["hbm_src"],
                    "pair_list": {"pairs": st["pairs"], "active_channels": st["active_channels"], "tiles": st["tiles"],
                                  "dv_list": st["dv_list"]}}
        # ---- channel-stream kernel (model spectra written to HBM): HBM roofline ---------------------------
        stream_roof = None
        if world == 1 and not args.no_extras:
            n_sim = max(1, min(256, (1 << 31) // (8 * prob.freq.size)))          # <= 2 GiB of spectra per launch
            d_sim = torch.empty((n_sim, prob.freq.size), dtype=torch.float64, device=dev)
            th_sim = d_thetas[0][:n_sim].contiguous()
            for _ in range(2):
                eng.simulate_device(th_sim, out=d_sim, sync=True)
            best = None
            for _ in range(5):
                tm = DeviceTimer(torch, stream)
                tm.begin()
                eng.simulate_device(th_sim, out=d_sim, sync=False)
                tm.end()
                eng.sync()
                best = tm.total_ms() if best is None else min(best, tm.total_ms())
            t_sim = best * 1e-3
            sim_bytes = n_sim * prob.freq.size * 8
            k_ns = eng.stat("fused_ns")            # CUDA events around the span kernel (or zero-fill + tiles) of the last call
            sncu = ncu_metrics("channel_stream", args)
            stream_roof = {"kernel": "cha_simulate_dev (channel-stream path, whole sequence)", "bound": "hbm",
                           "achieved": sim_bytes / t_sim / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": sim_bytes / t_sim / 1e9 / peaks["hbm_gbs"], "peak_src": peaks["hbm_src"],
                           "algorithmic_bytes_per_launch": sim_bytes, "walkers": n_sim, "ms": t_sim * 1e3,
                           "sequence": "need reduction (16-byte read-back) -> walker_prep -> sim_line_tau -> sim_gcoef -> simulate_span_kernel",
                           "stream_kernel_ms": k_ns * 1e-6 if k_ns > 0 else None,
                           "stream_kernel_frac": (sim_bytes / (k_ns * 1e-9) / 1e9 / peaks["hbm_gbs"]) if k_ns > 0 else None,
                           "traffic": sncu.get("dram_bytes"), "traffic_src": sncu.get("source")}
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
                "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32-mufu+f64-acc" if args.precision == "mixed" else "f64", "data": "synthetic",
                "config": workload_config(args, prob), "run": run_info, "clocks": clk, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "roofline_stream": stream_roof, "cpu_baseline": cpu,
                "sampler": smp_blk, "wall_s_timed_region": t_wall, "finite_logp_frac": finite_frac}
        line.update(line_extra)
        _emit(real_stdout, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
