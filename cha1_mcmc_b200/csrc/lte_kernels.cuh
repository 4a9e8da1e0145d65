// lte_kernels.cuh -- hand-written sm_100a kernels of the walker log-probability path.
//
// Data layout in HBM:
//   lines   : K_i, El_i, nu_i, mol_i                 one entry per SELECTED line, frequency-sorted
//   channels: only ACTIVE channels (touched by >=1 line window of the current lists) are streamed;
//             the walker-independent part of the chi-square is folded into a constant
//   mixed path (production): GroupBlk (8 consecutive active channels: dx, sigma-scaled data y/sigma and -1/sigma in
//             fp32, record counts),
//             LineRec (one per (line, group): velocity offset of the group's first channel, ckm/nu, line id),
//             TileG (<= 32 groups / 512 records / 24 lines, contiguous: one CTA per (tile, 128 walkers))
//   fp64 path (exactness reference): CSR over (active channel, molecule) with u_p = (nu_i - x_j)/nu_i*ckm
//   walkers : lanes of a warp = 32 consecutive walkers; channel / line / record data are warp-uniform broadcast
//             loads; per-walker line strengths live in shared-memory columns (mixed) or tau0[line][walker] (fp64)
//   partial : chi-square partial sums [tile][walker]; reduced in a fixed order (deterministic)
#pragma once
#include "lte_common.cuh"

namespace lte {

constexpr int kWalkersPerBlock = 128;   // thread = walker in the fused kernels

// ------------------------------------------------------------------------------------------
// (1) catalog precompute, once per molecule: classes.py:90-98 folded into the line factor
//     K_i = (ccm/(nu*1e6))^2 * (aij*gup) / (8*pi*nu*1e6/ckm)      [gup, glow cancel: classes.py:349-354]
// ------------------------------------------------------------------------------------------
__global__ void catalog_terms_kernel(int n, const double* __restrict__ nu, const double* __restrict__ logint,
                                     const double* __restrict__ elower, double q_ct,
                                     double* __restrict__ Kfac) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double f = nu[i], el = elower[i];
  double eu = el + f / kMHzPerCm;                                               // classes.py:90
  double sijmu = 1.0 / (exp(-(el / kBoltzLit) / kCT) - exp(-(eu / kBoltzLit) / kCT))
                 * (pow(10.0, logint[i]) / f) * (1.0 / kSijConst) * q_ct;        // classes.py:95
  double aij_gup = kAijConst * f * f * f * sijmu;                               // classes.py:98 (x gup)
  double lam = kCcm / (f * 1e6);
  Kfac[i] = lam * lam * aij_gup / (8.0 * M_PI * (f * 1e6 / kCkm));              // classes.py:351-353
}

// ------------------------------------------------------------------------------------------
// (2a) partition function, state-sum branch (functions.py:263-323):
//      Qpart[chunk][w] = sum_{s in chunk} g_s * exp(-E_s/(kcm*T_w))
//      lanes = walkers, warps stride over the chunk's states, fixed-order block reduction.
// ------------------------------------------------------------------------------------------
constexpr int kQChunk = 2048;
constexpr double kZcutPrep = 6.0;   // == kZcut (declared next to the mixed kernel)
__global__ void __launch_bounds__(256)
q_state_sum_kernel(const double* __restrict__ theta, int nw, int ndim, int idx_tex,
                   const double* __restrict__ g, const double* __restrict__ E, int n_states,
                   double* __restrict__ qpart, int nwp) {
  __shared__ double red[8][32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int w = blockIdx.x * 32 + lane;
  int s0 = blockIdx.y * kQChunk, s1 = min(n_states, s0 + kQChunk);
  double acc = 0.0;
  if (w < nw) {
    double T = theta[(size_t)w * ndim + idx_tex];
    double nb = -1.0 / (kKcm * T);
    for (int s = s0 + warp; s < s1; s += 8) acc += g[s] * exp(E[s] * nb);      // functions.py:323
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    if (w < nwp) qpart[(size_t)blockIdx.y * nwp + w] = t;
  }
}

// ------------------------------------------------------------------------------------------
// (2b) per-walker preparation: bounds + prior (inference.py:169-236; TMC1:224-268), Q(Tex)
//      and 1/(Q*dV) per molecule.  One thread per walker.
// ------------------------------------------------------------------------------------------
struct PriorDev {               // device arrays of length ndim
  const double* lo; const double* hi; const double* mu; const double* sg; const int* gauss;
  double vmin_sep, vmax_sep;
};

__global__ void walker_prep_kernel(const double* __restrict__ theta, int nw, int nwp, ModelDev md, PriorDev pr,
                                   int with_prior, const QDesc* __restrict__ qd,
                                   const double* __restrict__ qpart, int n_qchunks_max,
                                   int* __restrict__ ok, double* __restrict__ lp,
                                   double* __restrict__ qinv /*[M][nwp]*/,
                                   float* __restrict__ wpf /*[2+K+M*K][nwp]*/, double* __restrict__ wpd /*[1+K][nwp]*/,
                                   unsigned long long* __restrict__ need /*[2] or nullptr*/, double need_lo, double need_hi,
                                   const int* __restrict__ dest /*nullptr, or [nw]: slot of input row r in the evaluation order*/) {
  const int w_in = blockIdx.x * blockDim.x + threadIdx.x;       // row of theta (and of the state-sum partials)
  double n_dv = 0.0, n_dc = 0.0;     // this walker's share of the batch maxima the pair list must cover (see `need` below)
  do {
  if (w_in >= nwp) break;
  // every per-walker table below is written at the row's slot in the evaluation order (reach-ordered batches); the
  // padding rows keep their own slots
  const int w = (dest && w_in < nw) ? dest[w_in] : w_in;
  if (w_in >= nw) { ok[w] = 0; lp[w] = -INFINITY; for (int m = 0; m < 2 * md.M; ++m) qinv[(size_t)m * nwp + w] = 0.0; break; }
  const double* th = theta + (size_t)w_in * md.ndim;
  bool good = true;
  double lprior = 0.0;
  for (int p = 0; p < md.ndim; ++p) if (!isfinite(th[p])) good = false;
  if (with_prior) {
    for (int p = 0; p < md.ndim; ++p) {
      double v = th[p];
      if (!(pr.lo[p] < v && v < pr.hi[p])) good = false;                         // strict: inference.py:175-178
    }
    if (!isnan(pr.vmin_sep))
      for (int c = 0; c + 1 < md.K; ++c)
        if (!(th[md.idx_vlsr[c]] < th[md.idx_vlsr[c + 1]] - pr.vmin_sep)) good = false;   // TMC1:229
    if (!isnan(pr.vmax_sep))
      for (int c = 0; c + 1 < md.K; ++c)
        if (!(th[md.idx_vlsr[c + 1]] < th[md.idx_vlsr[c]] + pr.vmax_sep)) good = false;   // TMC1:230
    if (good) {
      for (int p = 0; p < md.ndim; ++p) {
        if (!pr.gauss[p]) continue;                                              // flat in Ncol: inference.py:208
        double s = pr.sg[p], d = th[p] - pr.mu[p];
        lprior += log(1.0 / (sqrt(2.0 * M_PI) * s)) - 0.5 * (d * d / (s * s));   // inference.py:209-211
      }
    }
  }
  double T = th[md.idx_tex], dV = th[md.idx_dv];
  for (int m = 0; m < md.M; ++m) {
    double Q;
    if (qd[m].kind == 3) {
      Q = 0.0;
      int nch = (qd[m].n_states + kQChunk - 1) / kQChunk;
      for (int c = 0; c < nch; ++c) Q += qpart[((size_t)m * n_qchunks_max + c) * nwp + w_in];
    } else {
      Q = q_analytic(qd[m], T);
    }
    qinv[(size_t)m * nwp + w] = 1.0 / (Q * dV);                                  // classes.py:349,353
    qinv[(size_t)(md.M + m) * nwp + w] = -log2(Q * dV);                          // the same, for strengths formed in log2 space
  }
  // per-walker constants of chi2_mixed_kernel (see there): a, 10 dV, centre offsets, column densities (fp32);
  // Planck exponent per MHz and source_size^2 (fp64); bit 1 of ok = "the 10 dV mask is a no-op within kZcut sigma"
  bool maskfree = dV > 0.0;
  // bit 2: the model is non-negative (tau >= 0, Tex above Tbg: emission) -> packed fast path of chi2_mixed_kernel
  bool signsafe = dV > 0.0 && T > kTbg + 1e-4;
  for (int m = 0; m < md.M; ++m) if (!(qinv[(size_t)m * nwp + w] > 0.0)) signsafe = false;
  for (int i = 0; i < md.M * md.K; ++i) if (!(th[md.idx_ncol[i]] >= 0.0)) signsafe = false;
  if (wpf) {
    const double a64 = 0.84932180028801907 * kFwhm / dV;      // sqrt(log2(e)/2) / sigma_v
    wpf[w] = (float)a64;
    wpf[(size_t)nwp + w] = (float)(dV * 10);
    wpd[w] = (kH * 1e6) / (kK * T);
    for (int c = 0; c < md.K; ++c) {
      const double dc = th[md.idx_vlsr[c]] - md.al - md.mc;   // Gaussian centre relative to the mask centre
      wpf[(size_t)(2 + c) * nwp + w] = (float)(dc * a64);
      if (!(fabs(dc) <= dV * (10.0 - kZcutPrep / kFwhm))) maskfree = false;
      const double ss = md.idx_ss[c] < 0 ? md.fixed_ss : th[md.idx_ss[c]];
      wpd[(size_t)(1 + c) * nwp + w] = ss * ss;
      for (int m = 0; m < md.M; ++m)
        wpf[(size_t)(2 + md.K + m * md.K + c) * nwp + w] = (float)th[md.idx_ncol[m * md.K + c]];
    }
  }
  ok[w] = good ? (1 | (maskfree ? 2 : 0) | (signsafe ? 4 : 0)) : 0;
  lp[w] = good ? lprior : -INFINITY;
  if (need && isfinite(dV) && dV > 0.0 && dV > need_lo && dV < need_hi) {      // same rows as dv_max_kernel (lte_sampler.cuh)
    n_dv = dV;
    for (int c = 0; c < md.K; ++c) {
      const double x = fabs(th[md.idx_vlsr[c]] - md.al - md.mc);
      if (isfinite(x) && x > n_dc) n_dc = x;
    }
  }
  } while (0);
  // need[0] = max dV, need[1] = max_c |vlsr_c - al - mc| of the batch: what an optimistic launch (chalte.cu:
  // log_prob_dev_opt) checks against the pair list afterwards.  Non-negative doubles order like their bit patterns.
  if (need) {
    unsigned long long b0 = (unsigned long long)__double_as_longlong(n_dv), b1 = (unsigned long long)__double_as_longlong(n_dc);
    for (int o = 16; o; o >>= 1) {
      unsigned long long o0 = __shfl_xor_sync(0xffffffffu, b0, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o);
      b0 = o0 > b0 ? o0 : b0; b1 = o1 > b1 ? o1 : b1;
    }
    if ((threadIdx.x & 31) == 0) { if (b0) atomicMax(need, b0); if (b1) atomicMax(need + 1, b1); }
  }
}

// ------------------------------------------------------------------------------------------
// (2c) line optical depths per unit column density (classes.py:349-354):
//      tau0[i][w] = K_i * exp(-El_i/(0.695*T_w)) * (1 - exp(-h*nu_i*1e6/(k*T_w))) / (Q_m(T_w)*dV_w)
//      thread = walker (coalesced store of a 128-walker row), block.y strides over lines.
// ------------------------------------------------------------------------------------------
template <typename TauT>
__global__ void __launch_bounds__(kWalkersPerBlock)
line_tau_kernel(const double* __restrict__ theta, int nwp, int ndim, int idx_tex,
                const int* __restrict__ ok, const double* __restrict__ qinv,
                int n_lines, const double* __restrict__ Kfac, const double* __restrict__ El,
                const double* __restrict__ nu, const int* __restrict__ mol,
                TauT* __restrict__ tau0, int lines_per_block) {
  int w = blockIdx.x * kWalkersPerBlock + threadIdx.x;
  int i0 = blockIdx.y * lines_per_block, i1 = min(n_lines, i0 + lines_per_block);
  bool live = (ok[w] & 1) != 0;
  double T = live ? theta[(size_t)w * ndim + idx_tex] : 1.0;
  double a = -1.0 / (kBoltzLit * T);
  double b = -(kH * 1e6) / (kK * T);
  for (int i = i0; i < i1; ++i) {
    double v = 0.0;
    if (live) {
      double boltz = exp(El[i] * a);                                            // classes.py:349
      double stim = 1.0 - exp(nu[i] * b);                                       // classes.py:351
      v = Kfac[i] * boltz * stim * qinv[(size_t)mol[i] * nwp + w];
    }
    tau0[(size_t)i * nwp + w] = (TauT)v;
  }
}

// Line strength per unit column density for the mixed path (classes.py:349-354), rounded to fp32 at the end:
//   tau0 = K_i * exp(-El_i/(0.695 T)) * (1 - exp(-h nu_i/(k T))) / (Q(T) dV)
// The two exponentials are evaluated to ~5e-10 relative so that the only error left is the final fp32 rounding
// (an error here is coherent over all channels of a line):
//   exp(-El/(0.695 T)) = 2^n * e^g : t = El*a2, n = rint(t) (magic-number rounding), |g| <= 0.3466,
//                        degree-8 Taylor in fp64 (4.5e-10), 2^n applied through the exponent bits
//   1 - exp(-x), x = h nu/(k T) < 0.6 : x * (alternating series to degree 9) (< 2.8e-9 relative), else exp()
// ~30 fp64 instructions instead of two full exp() calls (~110).
//   a2 = -log2(e)/(0.695 T),  b = h*1e6/(k T),  qinv = 1/(Q dV)
// fp64 literals cost two UMOVs each wherever they are used; coefficients read from the constant bank are free
// DFMA operands
__constant__ double kInvFact[11] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                                    1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0};
__constant__ double kPow2Consts[2] = {6755399441055744.0 /* 1.5 * 2^52 */, 0.6931471805599453 /* ln 2 */};

__device__ __forceinline__ double boltzmann_pow2(double t) {       // 2^t, |t| < 1000, ~4.5e-10 relative
  const double magic = kPow2Consts[0];                                          // (t + magic) - magic == rint(t)
  const double tm = t + magic;
  const int n = __double2loint(tm);
  const double g = (t - (tm - magic)) * kPow2Consts[1];                         // |g| <= 0.3466
  double e = kInvFact[8];
  e = fma(e, g, kInvFact[7]); e = fma(e, g, kInvFact[6]); e = fma(e, g, kInvFact[5]); e = fma(e, g, kInvFact[4]);
  e = fma(e, g, kInvFact[3]); e = fma(e, g, kInvFact[2]); e = fma(e, g, kInvFact[1]); e = fma(e, g, kInvFact[0]);
  return e * __longlong_as_double((long long)(n + 1023) << 52);                 // exact 2^n, n in (-1000, 1000)
}

// 2^t for |t| < 2^31 with the exponent clamped in the INTEGER domain (fp64 min/max cost ~4 instructions each):
// t < -1023 gives +0, t > 1023 saturates at 2^1023 * e^g
__device__ __forceinline__ double pow2_clamped(double t) {
  const double magic = kPow2Consts[0];
  const double tm = t + magic;
  const int n = __double2loint(tm);
  const double g = (t - (tm - magic)) * kPow2Consts[1];
  // degree 7: |g| <= 0.3466, next term g^8/8! < 5.2e-9 relative -- a tenth of the fp32 rounding the strength ends in
  double e = kInvFact[7];
  e = fma(e, g, kInvFact[6]); e = fma(e, g, kInvFact[5]); e = fma(e, g, kInvFact[4]);
  e = fma(e, g, kInvFact[3]); e = fma(e, g, kInvFact[2]); e = fma(e, g, kInvFact[1]); e = fma(e, g, kInvFact[0]);
  const int ex = max(min(n + 1023, 2046), 0);
  return e * __hiloint2double(ex << 20, 0);
}

// same with the stimulated-emission factor 1 - exp(-h nu/(k T)) supplied by the caller
__device__ __forceinline__ float line_strength_stim(double Kfac, double El, double a2, double stim, double qinv) {
  const double t = El * a2;                                                     // classes.py:349
  if (t > -1000.0 && t < 1000.0) return (float)(Kfac * boltzmann_pow2(t) * stim * qinv);
  if (t >= 1000.0) return (float)(Kfac * exp(t * 0.6931471805599453) * stim * qinv);
  return 0.0f;
}

__device__ __forceinline__ float line_strength(double Kfac, double El, double nu, double a2, double b, double qinv) {
  const double t = El * a2;                                                     // classes.py:349
  const double x = nu * b;                                                      // classes.py:351
  double stim;
  if (x < 0.6 && x > -0.6) {
    double p = -1.0 / 3628800.0;                                                // 1/10!
    p = fma(p, x, 1.0 / 362880.0); p = fma(p, x, -1.0 / 40320.0); p = fma(p, x, 1.0 / 5040.0);
    p = fma(p, x, -1.0 / 720.0);   p = fma(p, x, 1.0 / 120.0);    p = fma(p, x, -1.0 / 24.0);
    p = fma(p, x, 1.0 / 6.0);      p = fma(p, x, -0.5);           p = fma(p, x, 1.0);
    stim = p * x;
  } else {
    stim = 1.0 - exp(-x);
  }
  if (t > -1000.0 && t < 1000.0) return (float)(Kfac * boltzmann_pow2(t) * stim * qinv);
  if (t >= 1000.0) return (float)(Kfac * exp(t * 0.6931471805599453) * stim * qinv);
  return 0.0f;
}

// selected-line tables resident in HBM (one entry per selected line, frequency-sorted across molecules)
struct LinesDev {
  const double* Kfac; const double* El; const double* nu; const int* mol;
  const double* qinv;           // [2M][nwp]: 1/(Q_m(Tex_w) dV_w) from walker_prep_kernel, then log2 of the same
  const double* lK2;            // log2(Kfac) per selected line
};

// table variant of line_strength() (channel-stream kernel of the mixed path): tau0[line][walker] in fp32
__global__ void __launch_bounds__(kWalkersPerBlock)
line_tau_fast_kernel(const double* __restrict__ theta, int nwp, int ndim, int idx_tex,
                     const int* __restrict__ ok, const double* __restrict__ qinv,
                     int n_lines, const double* __restrict__ Kfac, const double* __restrict__ El,
                     const double* __restrict__ nu, const int* __restrict__ mol,
                     float* __restrict__ tau0, int lines_per_block) {
  const int w = blockIdx.x * kWalkersPerBlock + threadIdx.x;
  const int i0 = blockIdx.y * lines_per_block, i1 = min(n_lines, i0 + lines_per_block);
  const bool live = (ok[w] & 1) != 0;
  const double T = live ? theta[(size_t)w * ndim + idx_tex] : 1.0;
  const double a2 = -1.4426950408889634 / (kBoltzLit * T);    // log2(e) * (-1/(0.695 T))
  const double b = (kH * 1e6) / (kK * T);
  for (int i = i0; i < i1; ++i)
    tau0[(size_t)i * nwp + w] = live ? line_strength(Kfac[i], El[i], nu[i], a2, b, qinv[(size_t)mol[i] * nwp + w]) : 0.0f;
}

// ------------------------------------------------------------------------------------------
// per-walker registers of the fused kernels
// ------------------------------------------------------------------------------------------
template <int K>
struct WalkerF64 {
  double dV, inv_sig, Tex, mask_hw;      // sigma_v = dV/2.355 ; mask half width 10*dV
  double vl[K];                          // vlsr_c
  double ss2[K];                         // source_size_c^2
};

// ------------------------------------------------------------------------------------------
// (3a) fused profile + chi-square, all fp64, reference operation order (inference.py:50-60,
//      157-160; TMC1:158-179).  grid = (tiles, walker blocks); thread = walker; channel /
//      line / pair data are warp-uniform loads.  Model spectrum never written to HBM.
// ------------------------------------------------------------------------------------------
struct TileDev { int c0, c1; double xc, hs; };   // active-channel range [c0,c1), centre, half span (MHz)

struct SpecDev {
  const TileDev* tiles;
  const int* pair_off;        // [(n_act * M) + 1]
  const int* pair_line;       // [P] selected-line id
  const double* pair_u64;     // [P] (nu-x)/nu*ckm            (fp64 path: reference value)
  const float* pair_u32;      // [P] (nu-x)/nu*ckm - mc       (mixed path)
  const double* x;            // [n_act] channel frequency (MHz)
  const double* y;            // [n_act]
  const double* w;            // [n_act] 1/yerr^2
  const double* jbg;          // [n_act] J(x, 2.7)
  const double* beam2;        // [n_act] beam_size(x)^2
  const float* tn;            // [n_act] (x - xc)/hs of its tile
};

template <int K>
__global__ void __launch_bounds__(kWalkersPerBlock)
chi2_fp64_kernel(const double* __restrict__ theta, int nwp, ModelDev md, const int* __restrict__ ok,
                 SpecDev sp, const double* __restrict__ tau0, double* __restrict__ partial) {
  const int w = blockIdx.y * kWalkersPerBlock + threadIdx.x;
  const TileDev tile = sp.tiles[blockIdx.x];
  double chi = 0.0;
  if (ok[w]) {
    const double* th = theta + (size_t)w * md.ndim;
    const double dV = th[md.idx_dv], Tex = th[md.idx_tex];
    const double sig = dV / kFwhm;                                              // inference.py:53
    const double hw = dV * 10;                                                  // inference.py:52
    double vl[K], ss2[K], ncol[kMaxM][K];
#pragma unroll
    for (int c = 0; c < K; ++c) {
      vl[c] = th[md.idx_vlsr[c]];
      double ss = md.idx_ss[c] < 0 ? md.fixed_ss : th[md.idx_ss[c]];
      ss2[c] = ss * ss;
#pragma unroll
      for (int m = 0; m < kMaxM; ++m) ncol[m][c] = m < md.M ? th[md.idx_ncol[m * md.K + c]] : 0.0;
    }
    for (int j = tile.c0; j < tile.c1; ++j) {
      double acc[K];
#pragma unroll
      for (int c = 0; c < K; ++c) acc[c] = 0.0;
#pragma unroll
      for (int m = 0; m < kMaxM; ++m) {
        if (m >= md.M) break;
        double S[K];
#pragma unroll
        for (int c = 0; c < K; ++c) S[c] = 0.0;
        const int p0 = sp.pair_off[j * md.M + m], p1 = sp.pair_off[j * md.M + m + 1];
        for (int p = p0; p < p1; ++p) {
          const double u = sp.pair_u64[p];
          const double vg = u + md.al;                                          // inference.py:51
          if (fabs(vg - md.al - md.mc) < hw) {                                  // inference.py:52 / TMC1:160
            const double t0 = tau0[(size_t)sp.pair_line[p] * nwp + w];
#pragma unroll
            for (int c = 0; c < K; ++c) {
              double z = (vg - vl[c]) / sig;
              S[c] += t0 * exp(-0.5 * (z * z));                                 // inference.py:53
            }
          }
        }
#pragma unroll
        for (int c = 0; c < K; ++c) acc[c] += ncol[m][c] * S[c];
      }
      const double x = sp.x[j];
      const double dJ = planck_j(x, Tex, md.eps) - sp.jbg[j];                   // inference.py:56-57
      const double b2 = sp.beam2[j];
      double model = 0.0;
#pragma unroll
      for (int c = 0; c < K; ++c)
        model += dJ * (1.0 - exp(-acc[c])) * (ss2[c] / (b2 + ss2[c]));          // inference.py:60, 39
      const double r = sp.y[j] - model;
      chi += r * r * sp.w[j];                                                   // inference.py:160
    }
  }
  partial[(size_t)blockIdx.x * nwp + w] = chi;
}

// ------------------------------------------------------------------------------------------
// (3b) fused profile + chi-square, mixed precision -- the production kernel.
//   * u_p (fp32) is the velocity offset from the mask centre, formed in fp64 on the host:
//     the cancellation nu_i - x_j never happens in fp32
//   * Gaussian:  exp(-z^2/2) = 2^-(a*(u - d_c))^2,  a = sqrt(log2(e)/2)/sigma_v : FFMA, FMUL, MUFU.EX2, FFMA
//   * (J(x,Tex)-J(x,Tbg)) * dilution_c(x) is a smooth function of x: per (walker, component,
//     tile) it is sampled in fp64 at 4 Chebyshev nodes and evaluated per channel as a cubic
//     in fp32 (3 FFMA); the tile builder bounds span/x so the interpolation error is <1e-9
//   * 1 - exp(-T) is evaluated without cancellation (series below 0.5, MUFU above)
//   * residual and chi-square accumulate in fp64
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 1 - exp(-t), relative accuracy ~2e-7 for every t >= 0 (and small negative t)
__device__ __forceinline__ float one_minus_exp_neg(float t) {
  if (fabsf(t) < 0.5f) {
    // t*(1 - t/2 + t^2/6 - ... ) Horner, degree 9 in t: |err| < 0.5^10/10! = 2.7e-10
    float p = -1.0f / 3628800.0f;
    p = fmaf(p, t, 1.0f / 362880.0f);
    p = fmaf(p, t, -1.0f / 40320.0f);
    p = fmaf(p, t, 1.0f / 5040.0f);
    p = fmaf(p, t, -1.0f / 720.0f);
    p = fmaf(p, t, 1.0f / 120.0f);
    p = fmaf(p, t, -1.0f / 24.0f);
    p = fmaf(p, t, 1.0f / 6.0f);
    p = fmaf(p, t, -0.5f);
    p = fmaf(p, t, 1.0f);
    return p * t;
  }
  return 1.0f - ex2_approx(-1.4426950408889634f * t);
}

// inverse Vandermonde of the 4 Chebyshev nodes t_n = cos((2n+1)pi/8): monomial coefficients
// g_k = sum_n kChebInv[k][n] * G(t_n)
__constant__ double kChebNodes[4] = {0.9238795325112867, 0.38268343236508984,
                                     -0.3826834323650897, -0.9238795325112867};
__constant__ double kChebInv[4][4] = {
    {-0.10355339059327377, 0.60355339059327362, 0.60355339059327384, -0.10355339059327384},
    {-0.11208538229199132, 1.5771610149494748, -1.5771610149494748, 0.11208538229199132},
    {0.70710678118654757, -0.70710678118654757, -0.70710678118654757, 0.70710678118654757},
    {0.76536686473017956, -1.847759065022573, 1.847759065022573, -0.76536686473017956}};

// ---- group/tile layout of the mixed kernel ------------------------------------------------------
// A GROUP is up to 8 consecutive active channels (<= 1 km/s wide).  For a line covering a group the
// velocity offsets of its channels are an arithmetic progression, u_j = u0 - dx_j * (ckm/nu): one
// 16-byte record per (line, group) instead of one per (line, channel), and 8*K independent
// FFMA/FMUL/MUFU.EX2/FFMA chains per record.  A TILE is up to 32 groups; its group blocks and records
// are contiguous in HBM and are staged into shared memory by two TMA bulk copies (cp.async.bulk,
// mbarrier complete_tx) issued by one thread while all threads set up their walker.
constexpr double kHk = kH * 1e6 / kK;   // Kelvin per MHz
constexpr int kGroupCh = 8;
constexpr int kTileMaxGroups = 32;
constexpr int kTileMaxRecs = 512;
constexpr int kTileMaxLines = 24;     // tau0 columns of the tile's lines staged per block: 24 x 128 walkers x 4 B = 12 KB

struct __align__(16) GroupBlk {
  float dx[kGroupCh];            // x_j - x_first (MHz); padding channels repeat the last offset
  // chi-square in RESIDUAL form on sigma-scaled data: r_j = y_j/sigma_j - m_j/sigma_j, chi_j = r_j^2
  // (inference.py:157-160).  y_j/sigma_j is formed in fp64 on the host and split into two floats (hi + lo: 48
  // bits); the default build uses the hi part alone -- the data rounded to fp32 after the sigma scaling, 6e-8
  // relative -- and CHA_YS_SPLIT=1 adds the lo part to every residual (see residual2()).
  float ysh[kGroupCh];           // hi part of y_j/sigma_j                  (0 for padding channels)
  float ysl[kGroupCh];           // lo part of y_j/sigma_j                  (read only with CHA_YS_SPLIT=1)
  float ns[kGroupCh];            // -1/sigma_j                              (0 for padding channels)
  double y2w;                    // sum_j y_j^2/sigma_j^2 of the group: its chi-square when the model is exactly 0
  int rec_off;                   // first record of the group relative to the tile's rec_begin
  unsigned short nrec[kMaxM];    // records per molecule
  float tn0;                     // (x_first - xc)/hs of the tile
  int opos[kGroupCh];            // caller's channel index of channel j, -1 for padding (channel-stream kernel)
  int pad[2];
};
static_assert(sizeof(GroupBlk) == 192, "GroupBlk must be 192 bytes (16-byte multiple for cp.async.bulk)");

struct __align__(16) LineRec { float u0, slope; int line; int lloc; };   // line: selected-line id; lloc: (line - tile.line0) * kWalkersPerBlock
static_assert(sizeof(LineRec) == 16, "LineRec must be 16 bytes");

struct __align__(16) TileG {
  int g0, ng, rec_begin, rec_count;
  int line0, nline, pad0, pad1;  // the tile's records reference selected lines [line0, line0 + nline)
  double xc, hs;
  double jbg[4];                 // J(x_n, 2.7 K) at the 4 Chebyshev nodes (walker independent)
  double beam2[4];               // beam_size(x_n)^2
  double jbg_hi, jbg_lo;         // the same at the tile's end points xc + hs, xc - hs (narrow tiles)
  double beam2_hi, beam2_lo;
  double line_span;              // max |nu_i - xc| over the tile's lines (MHz)
  double inv_hs;                 // 1 / hs
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// 1/d to ~4e-15 relative: MUFU.RCP seed + one Newton step (the interpolant is rounded to fp32 anyway)
__device__ __forceinline__ double fast_rcp(double d) {
  double r = (double)__frcp_rn((float)d);
  double e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// Gaussian terms further than kZcut sigma from every line centre are < exp(-kZcut^2/2) = 1.5e-8 of the line
// peak -- below the fp32 resolution (6e-8) the mixed path forms the model in: the pair list is truncated there
// (host: ensure_pairs) and the reference's 10*dV mask (inference.py:52) is applied explicitly only for walkers
// whose mask edge lies inside that range.
constexpr double kZcut = 6.0;
static_assert(kZcut == kZcutPrep, "walker_prep_kernel and the list builder must agree on the truncation");
constexpr float kVcut = (float)(kZcut * 0.84932180028801907);   // kZcut sigma as an argument v of exp2(-v^2)

// ---- packed fp32 pairs: Blackwell FFMA2/FMUL2 (PTX fma.rn.f32x2 / mul.rn.f32x2, sm_100+) -----------------
// one instruction issue for two channels; the kernel is issue-bound, not FMA-pipe bound
typedef unsigned long long f32x2;
template <int K> constexpr bool kGcInSmem = K > 1 && K <= 4;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
// acc += a * b with the accumulator tied to one register pair (ptxas otherwise forms the sum in a temporary pair and
// moves it back when the update sits under a branch: two MOVs per packed FMA in the multi-molecule path)
__device__ __forceinline__ void fma2_acc(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// Residual of a channel pair and its contribution to the group's chi-square, all packed fp32:
//   r = (ysh - m/sigma) + ysl  (FFMA2 + FADD2),  acc += r^2 (FFMA2)
// A model value outside the fp32 range makes r, the group sum and the fp64 partial non-finite; finalize_kernel
// maps that to -inf (inference.py:162-164) -- no magnitude sentinel.
// CHA_YS_SPLIT=1 adds the low part of the sigma-scaled data (hi + lo split, 48 bits) to every residual: one packed add
// and half an LDS.128 more per channel pair.  Without it the data enter rounded to fp32 AFTER the sigma scaling (6e-8
// relative, five orders below the noise), which moves lnlike by 2 sum_j (r_j/sigma_j)(y_j/sigma_j) 3e-8 with random
// signs -- measured next to the other fp32 terms in profiles/r02_parity_errors.json.
#ifndef CHA_YS_SPLIT
#define CHA_YS_SPLIT 0
#endif
#ifndef CHA_STRENGTH_MUFU
#define CHA_STRENGTH_MUFU 0
#endif
#ifndef CHA_K4_BLOCKS
#define CHA_K4_BLOCKS 5          // resident CTAs per SM the K = 3, 4 instantiations are compiled for (96 registers; measured against 4)
#endif
// Group chi-square (a sum of squares: >= +0, or non-finite) to fp64 without F2F.F64.F32, which shares the XU pipe with
// MUFU.EX2 (8 cycles per warp instruction): two integer instructions, hi = (bits >> 3) + 0x38000000, lo = bits << 29.
// Exact for normal values; +0 and denormals map to < 2^-126 (an additive error of no consequence).  Inf / NaN would
// come out finite, so the largest bit pattern seen travels beside the sum (one integer max per group) and the
// caller turns the partial into +inf when it reached 0x7f800000 (finalize_kernel: -inf, inference.py:162-164).
__device__ __forceinline__ double chi_group_to_double(float s, unsigned& smax) {
  const unsigned b = __float_as_uint(s);
  smax = max(smax, b);
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}

__device__ __forceinline__ f32x2 residual2(f32x2 model2, f32x2 ns2, f32x2 ysh2, f32x2 ysl2) {
#if CHA_YS_SPLIT
  return add2(fma2(model2, ns2, ysh2), ysl2);
#else
  (void)ysl2;
  return fma2(model2, ns2, ysh2);
#endif
}

// General path: reference mask applied explicitly, any sign of the model, records possibly in global
// memory.  Used only by blocks where some live walker needs it (see chi2_mixed_kernel).
template <int K>
__device__ __noinline__ double chi2_mixed_groups(const GroupBlk* __restrict__ s_grp, int ng,
                                                 const LineRec* __restrict__ rbase, int M, int nwp, int w,
                                                 const LinesDev ln, double a2, double cT, float a, const float (&sc)[K],
                                                 float hw, const float (&ncol)[kMaxM][K], const float (&gc)[K][4],
                                                 float inv_hs) {
  double chi = 0.0;
  for (int g = 0; g < ng; ++g) {
    const GroupBlk& gb = s_grp[g];
    float dx[kGroupCh];
    {
      const float4 d0 = *reinterpret_cast<const float4*>(&gb.dx[0]);
      const float4 d1 = *reinterpret_cast<const float4*>(&gb.dx[4]);
      dx[0] = d0.x; dx[1] = d0.y; dx[2] = d0.z; dx[3] = d0.w; dx[4] = d1.x; dx[5] = d1.y; dx[6] = d1.z; dx[7] = d1.w;
    }
    float T[K][kGroupCh];                     // optical depth per component and channel
#pragma unroll
    for (int c = 0; c < K; ++c)
#pragma unroll
      for (int j = 0; j < kGroupCh; ++j) T[c][j] = 0.0f;
    int r = gb.rec_off;
#pragma unroll
    for (int m = 0; m < kMaxM; ++m) {
      if (m >= M) break;
      const int n = gb.nrec[m];
      for (int q = 0; q < n; ++q, ++r) {
        const LineRec rc = rbase[r];
        const float t0 = line_strength(ln.Kfac[rc.line], ln.El[rc.line], ln.nu[rc.line], a2, cT,
                                       ln.qinv[(size_t)ln.mol[rc.line] * nwp + w]);
        const float B = rc.slope * a;
        float A[K], tn[K];
#pragma unroll
        for (int c = 0; c < K; ++c) { A[c] = fmaf(rc.u0, a, -sc[c]); tn[c] = t0 * ncol[m][c]; }   // classes.py:349 (x Ncol)
#pragma unroll
        for (int j = 0; j < kGroupCh; ++j) {
          const bool in = fabsf(fmaf(-dx[j], rc.slope, rc.u0)) < hw;                   // inference.py:52
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const float v = fmaf(-dx[j], B, A[c]);                                     // inference.py:51,53
            const float e = ex2_approx(-v * v);
            T[c][j] = fmaf(in ? tn[c] : 0.0f, e, T[c][j]);
          }
        }
      }
    }
    // G_c at the group's first channel and its slope per MHz: within a group (<= 1 km/s, dx/x <= 3.3e-6) G is
    // linear to 1e-11
    float G0[K], Gp[K];
#pragma unroll
    for (int c = 0; c < K; ++c) {
      G0[c] = fmaf(fmaf(fmaf(gc[c][3], gb.tn0, gc[c][2]), gb.tn0, gc[c][1]), gb.tn0, gc[c][0]);
      Gp[c] = fmaf(fmaf(3.0f * gc[c][3], gb.tn0, 2.0f * gc[c][2]), gb.tn0, gc[c][1]) * inv_hs;
    }
#pragma unroll
    for (int j = 0; j < kGroupCh; ++j) {
      float model = 0.0f;
#pragma unroll
      for (int c = 0; c < K; ++c)
        model = fmaf(fmaf(dx[j], Gp[c], G0[c]), one_minus_exp_neg(T[c][j]), model);    // inference.py:60
      // inference.py:160 on sigma-scaled data: ((y - m)/sigma)^2, residual in fp64 (any sign of the model)
      const double r = fma((double)model, (double)gb.ns[j], CHA_YS_SPLIT ? (double)gb.ysh[j] + (double)gb.ysl[j] : (double)gb.ysh[j]);
      chi = fma(r, r, chi);
    }
  }
  return chi;
}

// Fast path (the MCMC regime): every live walker of the block is mask-free within kZcut sigma, has non-negative
// column densities and Tex above Tbg (emission), so model = sum_c G_c (1 - e^-tau_c) >= 0 and the integer
// fp32 -> fp64 conversion applies.  Channels are processed as packed pairs (FFMA2/FMUL2); per channel:
//   pair loop   1.5 issues + 1 MUFU.EX2 per (line, channel, component)
//   epilogue    thin (all tau < 1/32): 3 packed issues/component; residual on sigma-scaled data, packed fp32:
//               r = y/sigma - m/sigma (FFMA2 per channel pair), sum r^2 over the group's 8 channels in fp32, one fp64
//               add per group (chi_group_to_double: integer conversion, no F2F on the XU pipe)
template <int K>
__device__ __forceinline__ double chi2_mixed_groups_fast(const GroupBlk* __restrict__ s_grp, int ng,
                                                         const LineRec* __restrict__ s_rec, int M,
                                                         const float* __restrict__ tau_col, float a,
                                                         const float (&sc)[K], const float (&ncol)[kMaxM][K],
                                                         const float* __restrict__ ncol_col,
                                                         const float (&gc)[K][4], const float* __restrict__ gc_col,
                                                         float inv_hs) {
  // K > 1: the column densities are read from ncol_col, this walker's column of the block's shared table,
  // ncol_col[(m * K + c) * kWalkersPerBlock] (M * K values in registers next to K x 4 packed optical depths left ptxas
  // at the 128-register cap moving every packed pair it formed: one MOV per MUFU; a thread reads only its own column,
  // so no barrier is needed).  K == 1: four registers, selected by the record's molecule.
  double chi = 0.0;
  unsigned smax = 0u;                       // largest fp32 bit pattern among the group sums (non-finite detector)
  const f32x2 c24 = pk2(-1.0f / 24.0f, -1.0f / 24.0f), c6 = pk2(1.0f / 6.0f, 1.0f / 6.0f);
  const f32x2 ch = pk2(-0.5f, -0.5f), c1 = pk2(1.0f, 1.0f);
  for (int g = 0; g < ng; ++g) {
    const GroupBlk& gb = s_grp[g];
    f32x2 dx2[4];
    {
      const ulonglong2 d0 = *reinterpret_cast<const ulonglong2*>(&gb.dx[0]);
      const ulonglong2 d1 = *reinterpret_cast<const ulonglong2*>(&gb.dx[4]);
      dx2[0] = d0.x; dx2[1] = d0.y; dx2[2] = d1.x; dx2[3] = d1.y;
    }
    f32x2 T2[K][4];                           // optical depth per component and channel pair
#pragma unroll
    for (int c = 0; c < K; ++c)
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) T2[c][jp] = 0ull;
    // K > 1: a component further than kVcut from every channel of the group is skipped for that record, and a
    // component no record reached is skipped in the epilogue (see chi2_mixed_groups_fast1)
    const float dxm = 0.5f * gb.dx[kGroupCh - 1];
    unsigned live = K == 1 ? 1u : 0u;
    // the group's records are ONE contiguous stream, molecule after molecule (nrec[m] = 0 for m >= M): a flat loop
    // with the record's molecule found from the running index, so that the optical depths stay in one set of
    // registers (a loop per molecule made ptxas move every packed sum back after forming it: 1 MOV per MUFU)
    const LineRec* __restrict__ rp = s_rec + gb.rec_off;
    const int cut1 = gb.nrec[0], cut2 = cut1 + gb.nrec[1], cut3 = cut2 + gb.nrec[2], nrecs = cut3 + gb.nrec[3];
    (void)M;
#pragma unroll 1
    for (int q = 0; q < nrecs; ++q, ++rp) {
      const LineRec rc = *rp;
      const float t0 = tau_col[rc.lloc];
      const float nB = -rc.slope * a;
      const f32x2 nB2 = pk2(nB, nB);
      const float reach = fmaf(dxm, fabsf(nB), kVcut);
      const int m = (q >= cut1) + (q >= cut2) + (q >= cut3);
      const float* __restrict__ ncm = K > 1 ? ncol_col + m * (K * kWalkersPerBlock) : nullptr;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const float A = fmaf(rc.u0, a, -sc[c]);
        if (K > 1 && !(fabsf(fmaf(dxm, nB, A)) < reach)) continue;
        live |= 1u << c;
        float nc;
        if constexpr (K > 1) nc = ncm[c * kWalkersPerBlock];
        else nc = m == 0 ? ncol[0][c] : (m == 1 ? ncol[1][c] : (m == 2 ? ncol[2][c] : ncol[3][c]));
        const float tn = t0 * nc;                                                      // classes.py:349 (x Ncol)
        const f32x2 A2 = pk2(A, A), tn2 = pk2(tn, tn);
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          const f32x2 v2 = fma2(dx2[jp], nB2, A2);                                     // inference.py:51,53
          float s0, s1;
          upk2(mul2(v2, v2), s0, s1);
          const f32x2 e2 = pk2(ex2_approx(-s0), ex2_approx(-s1));
          fma2_acc(T2[c][jp], tn2, e2);
        }
      }
    }
    f32x2 model2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int c = 0; c < K; ++c) {
      if (!(live & (1u << c))) continue;
      float g0c, g1c, g2c, g3c;
      if constexpr (kGcInSmem<K>) {
        g0c = gc_col[(4 * c + 0) * kWalkersPerBlock]; g1c = gc_col[(4 * c + 1) * kWalkersPerBlock];
        g2c = gc_col[(4 * c + 2) * kWalkersPerBlock]; g3c = gc_col[(4 * c + 3) * kWalkersPerBlock];
      } else { g0c = gc[c][0]; g1c = gc[c][1]; g2c = gc[c][2]; g3c = gc[c][3]; }
      const float G0 = fmaf(fmaf(fmaf(g3c, gb.tn0, g2c), gb.tn0, g1c), gb.tn0, g0c);
      const float Gp = fmaf(fmaf(3.0f * g3c, gb.tn0, 2.0f * g2c), gb.tn0, g1c) * inv_hs;
      const f32x2 G02 = pk2(G0, G0), Gp2 = pk2(Gp, Gp);
      float tmax = 0.0f;
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        float t0, t1;
        upk2(T2[c][jp], t0, t1);
        tmax = fmaxf(tmax, fmaxf(t0, t1));
      }
      if (tmax < 0.03125f) {
        // optically thin (the usual case): 1 - exp(-tau) = tau (1 - tau/2 + tau^2/6 - tau^3/24), next term < 8e-9
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          const f32x2 tau2 = T2[c][jp];
          f32x2 p2 = fma2(tau2, c24, c6);
          p2 = fma2(p2, tau2, ch);
          p2 = fma2(p2, tau2, c1);
          model2[jp] = fma2(fma2(dx2[jp], Gp2, G02), mul2(p2, tau2), model2[jp]);      // inference.py:60
        }
      } else {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          float t0, t1;
          upk2(T2[c][jp], t0, t1);
          const f32x2 e2 = pk2(one_minus_exp_neg(t0), one_minus_exp_neg(t1));
          model2[jp] = fma2(fma2(dx2[jp], Gp2, G02), e2, model2[jp]);
        }
      }
    }
    if (live == 0u) { chi += gb.y2w; continue; }     // no component reached the group: model == 0 exactly
    f32x2 acc2 = 0ull;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      const f32x2 r2 = residual2(model2[jp], *reinterpret_cast<const f32x2*>(&gb.ns[2 * jp]),
                                 *reinterpret_cast<const f32x2*>(&gb.ysh[2 * jp]),
                                 CHA_YS_SPLIT ? *reinterpret_cast<const f32x2*>(&gb.ysl[2 * jp]) : 0ull);
      acc2 = jp == 0 ? mul2(r2, r2) : fma2(r2, r2, acc2);                              // inference.py:160
    }
    float s0, s1;
    upk2(acc2, s0, s1);
    chi += chi_group_to_double(s0 + s1, smax);
  }
  return smax >= 0x7f800000u ? (double)INFINITY : chi;
}

// Single-molecule fast path (M == 1, the common fit): the tile's records are ONE contiguous stream that the
// groups consume in order, so the record pointer just advances.  The line strengths tau0[line][w] of the tile's
// lines are staged per block as shared-memory columns (thread = walker reads only its own column: conflict-free,
// no barrier), so the hot loop has no global load; the next record's descriptor and strength are fetched one
// record ahead.  s_rec holds rec_count + 1 records (the host appends a dummy after the last tile) so the
// look-ahead never leaves the staged range.
// The polynomial for 1 - exp(-tau) is chosen per group from the largest optical depth in it.
template <int K, bool NARROW, bool SKIP>
__device__ __forceinline__ double chi2_mixed_groups_fast1(const GroupBlk* __restrict__ gb, int ng,
                                                          const LineRec* __restrict__ rp,
                                                          const float* __restrict__ tau_col, float a,
                                                          const float (&sc)[K], const float (&ncol)[K],
                                                          const float (&gc)[K][4], const float* __restrict__ gc_col,
                                                          float inv_hs, float vcut1) {
  double chi = 0.0;
  unsigned smax = 0u;                       // largest fp32 bit pattern among the group sums (non-finite detector)
  const f32x2 c24 = pk2(-1.0f / 24.0f, -1.0f / 24.0f), c6 = pk2(1.0f / 6.0f, 1.0f / 6.0f);
  const f32x2 ch = pk2(-0.5f, -0.5f), c1 = pk2(1.0f, 1.0f);
  // (rcA, tA): the next record to process and this walker's line strength for it, fetched ahead of use.
  // lloc is the line's row in the staged columns, pre-multiplied by the row pitch (floats).
  LineRec rcA = *rp;
  float tA = tau_col[rcA.lloc];
  for (int g = 0; g < ng; ++g, ++gb) {
    f32x2 dx2[4];
    {
      const ulonglong2 d0 = *reinterpret_cast<const ulonglong2*>(&gb->dx[0]);
      const ulonglong2 d1 = *reinterpret_cast<const ulonglong2*>(&gb->dx[4]);
      dx2[0] = d0.x; dx2[1] = d0.y; dx2[2] = d1.x; dx2[3] = d1.y;
    }
    f32x2 T2[K][4];                           // optical depth per component and channel pair
#define CHA_RECORD(RC, T0)   CHA_RECORD_(RC, T0, false)
    // K > 1: a component whose centre is further than kVcut (= kZcut sigma in the units of the exponent) from every
    // channel of the group contributes < 2^-26 of its peak -- the same bound the pair list is truncated at -- and is
    // skipped for this record: one FFMA and a compare per component against 4 x (3 packed + 2 MUFU) instructions.
    // The test is on this walker's own centres; walkers of a warp sit close together, so it rarely diverges.
    // K == 1, SKIP: the same test per record, for warps holding walkers much narrower than the batch the list was
    // built for (an ensemble mid-run: the list covers the widest proposal, the bulk needs half of it).  vcut1 is
    // kVcut for such a walker and +inf for the others, so WHICH terms a walker sums depends on its own parameters
    // and the list only -- never on the walkers it shares a warp with.
    const float dxm = 0.5f * gb->dx[kGroupCh - 1];     // half extent of the group (padding repeats the last offset)
    unsigned live = 0u;                                // components some record of this group reached
#define CHA_RECORD_(RC, T0, FIRST)                                                                                \
    {                                                                                                        \
      const float nB = -(RC).slope * a;                                                                      \
      const f32x2 nB2 = pk2(nB, nB);                                                                         \
      if constexpr (K == 1) {                                                                                \
        const float A = fmaf((RC).u0, a, -sc[0]);                                                            \
        if (!SKIP || fabsf(fmaf(dxm, nB, A)) < fmaf(dxm, fabsf(nB), vcut1)) {                                \
          if (SKIP) live = 1u;                                                                               \
          const float tn = (T0) * ncol[0];                                   /* classes.py:349 (x Ncol) */   \
          const f32x2 A2 = pk2(A, A), tn2 = pk2(tn, tn);                                                     \
          _Pragma("unroll") for (int jp = 0; jp < 4; ++jp) {                                                 \
            const f32x2 v2 = fma2(dx2[jp], nB2, A2);                         /* inference.py:51,53 */        \
            float s0, s1;                                                                                    \
            upk2(mul2(v2, v2), s0, s1);                                                                      \
            const f32x2 e2 = pk2(ex2_approx(-s0), ex2_approx(-s1));                                          \
            T2[0][jp] = (FIRST) ? mul2(tn2, e2) : fma2(tn2, e2, T2[0][jp]);                                  \
          }                                                                                                  \
        } else if (FIRST) {                                                                                  \
          _Pragma("unroll") for (int jp = 0; jp < 4; ++jp) T2[0][jp] = 0ull;                                 \
        }                                                                                                    \
      } else {                                                                                               \
        const float reach = fmaf(dxm, fabsf(nB), kVcut);                                                     \
        _Pragma("unroll") for (int c = 0; c < K; ++c) {                                                      \
          const float A = fmaf((RC).u0, a, -sc[c]);                                                          \
          if (fabsf(fmaf(dxm, nB, A)) < reach) {                                                             \
            live |= 1u << c;                                                                                 \
            const float tn = (T0) * ncol[c];                                                                 \
            const f32x2 A2 = pk2(A, A), tn2 = pk2(tn, tn);                                                   \
            _Pragma("unroll") for (int jp = 0; jp < 4; ++jp) {                                               \
              const f32x2 v2 = fma2(dx2[jp], nB2, A2);                                                       \
              float s0, s1;                                                                                  \
              upk2(mul2(v2, v2), s0, s1);                                                                    \
              const f32x2 e2 = pk2(ex2_approx(-s0), ex2_approx(-s1));                                        \
              T2[c][jp] = (FIRST) ? mul2(tn2, e2) : fma2(tn2, e2, T2[c][jp]);                                \
            }                                                                                                \
          } else if (FIRST) {                                                                                \
            _Pragma("unroll") for (int jp = 0; jp < 4; ++jp) T2[c][jp] = 0ull;                               \
          }                                                                                                  \
        }                                                                                                    \
      }                                                                                                      \
    }
    // every group of a single-molecule list has >= 1 record (host): the first one initialises the optical depths
    int q = gb->nrec[0] - 1;
    CHA_RECORD_(rcA, tA, true)
    rcA = *++rp;
    tA = tau_col[rcA.lloc];
#pragma unroll 1
    for (; q >= 2; q -= 2) {                  // two records per trip, ping-pong registers: no rotation moves
      const LineRec rcB = rp[1];
      const float tB = tau_col[rcB.lloc];
      CHA_RECORD(rcA, tA)
      rcA = rp[2];
      tA = tau_col[rcA.lloc];
      rp += 2;
      CHA_RECORD(rcB, tB)
    }
    if (q > 0) {
      CHA_RECORD(rcA, tA)
      rcA = *++rp;
      tA = tau_col[rcA.lloc];
    }
#undef CHA_RECORD
#undef CHA_RECORD_
    // no record reached this walker: the model is exactly 0 on the group's channels
    if ((SKIP || K > 1) && live == 0u) { chi += gb->y2w; continue; }
    const float tn0 = gb->tn0;
    f32x2 acc2;
    // (1 - exp(-tau))/tau: 1 - tau/2 below 4e-4 (next term tau^2/6 < 2.7e-8), degree 3 below 1/32 (next term
    // tau^4/120 < 8e-9), MUFU.EX2 above
#define CHA_EPILOGUE_PAIR(MODEL2)                                                                      \
      {                                                                                                \
        const f32x2 r2 = residual2(MODEL2, *reinterpret_cast<const f32x2*>(&gb->ns[2 * jp]),           \
                                   *reinterpret_cast<const f32x2*>(&gb->ysh[2 * jp]),                  \
                                   CHA_YS_SPLIT ? *reinterpret_cast<const f32x2*>(&gb->ysl[2 * jp]) : 0ull);  \
        acc2 = jp == 0 ? mul2(r2, r2) : fma2(r2, r2, acc2);              /* inference.py:160 */        \
      }
    if constexpr (K == 1) {
      f32x2 G02[K], Gp2[K];
#pragma unroll
      for (int c = 0; c < K; ++c) {
        // narrow tiles: G is linear over the whole tile (gc[2] = gc[3] = 0)
        const float G0 = NARROW ? fmaf(gc[c][1], tn0, gc[c][0])
                                : fmaf(fmaf(fmaf(gc[c][3], tn0, gc[c][2]), tn0, gc[c][1]), tn0, gc[c][0]);
        const float Gp = NARROW ? gc[c][1] * inv_hs
                                : fmaf(fmaf(3.0f * gc[c][3], tn0, 2.0f * gc[c][2]), tn0, gc[c][1]) * inv_hs;
        G02[c] = pk2(G0, G0); Gp2[c] = pk2(Gp, Gp);
      }
      float tmax = 0.0f;
#pragma unroll
      for (int c = 0; c < K; ++c)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          float t0, t1;
          upk2(T2[c][jp], t0, t1);
          tmax = fmaxf(tmax, fmaxf(t0, t1));
        }
      if (tmax < 4e-4f) {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          f32x2 model2 = 0ull;
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const f32x2 tau2 = T2[c][jp];
            const f32x2 pt2 = mul2(fma2(tau2, ch, c1), tau2);
            const f32x2 g2 = fma2(dx2[jp], Gp2[c], G02[c]);
            model2 = c == 0 ? mul2(g2, pt2) : fma2(g2, pt2, model2);                     // inference.py:60
          }
          CHA_EPILOGUE_PAIR(model2)
        }
      } else if (tmax < 0.03125f) {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          f32x2 model2 = 0ull;
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const f32x2 tau2 = T2[c][jp];
            f32x2 p2 = fma2(tau2, c24, c6);
            p2 = fma2(p2, tau2, ch);
            p2 = fma2(p2, tau2, c1);
            const f32x2 pt2 = mul2(p2, tau2);
            const f32x2 g2 = fma2(dx2[jp], Gp2[c], G02[c]);
            model2 = c == 0 ? mul2(g2, pt2) : fma2(g2, pt2, model2);
          }
          CHA_EPILOGUE_PAIR(model2)
        }
      } else {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          f32x2 model2 = 0ull;
#pragma unroll
          for (int c = 0; c < K; ++c) {
            float t0, t1;
            upk2(T2[c][jp], t0, t1);
            const f32x2 e2 = pk2(one_minus_exp_neg(t0), one_minus_exp_neg(t1));
            const f32x2 g2 = fma2(dx2[jp], Gp2[c], G02[c]);
            model2 = c == 0 ? mul2(g2, e2) : fma2(g2, e2, model2);
          }
          CHA_EPILOGUE_PAIR(model2)
        }
      }
    } else {
      // a component no record of this group reached has zero optical depth on all 8 channels: skip its G and its
      // 1 - exp(-tau); the polynomial degree is chosen per component
      f32x2 model2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int c = 0; c < K; ++c) {
        if (!(live & (1u << c))) continue;
        // 1 < K <= 4: the interpolant's coefficients come from this walker's shared-memory column (16 registers at K = 4)
        float g0c, g1c, g2c = 0.f, g3c = 0.f;
        if constexpr (kGcInSmem<K>) {
          g0c = gc_col[(4 * c + 0) * kWalkersPerBlock]; g1c = gc_col[(4 * c + 1) * kWalkersPerBlock];
          if (!NARROW) { g2c = gc_col[(4 * c + 2) * kWalkersPerBlock]; g3c = gc_col[(4 * c + 3) * kWalkersPerBlock]; }
        } else { g0c = gc[c][0]; g1c = gc[c][1]; g2c = gc[c][2]; g3c = gc[c][3]; }
        const float G0 = NARROW ? fmaf(g1c, tn0, g0c) : fmaf(fmaf(fmaf(g3c, tn0, g2c), tn0, g1c), tn0, g0c);
        const float Gp = NARROW ? g1c * inv_hs : fmaf(fmaf(3.0f * g3c, tn0, 2.0f * g2c), tn0, g1c) * inv_hs;
        const f32x2 G02 = pk2(G0, G0), Gp2 = pk2(Gp, Gp);
        float tmax = 0.0f;
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          float t0, t1;
          upk2(T2[c][jp], t0, t1);
          tmax = fmaxf(tmax, fmaxf(t0, t1));
        }
        if (tmax < 4e-4f) {
#pragma unroll
          for (int jp = 0; jp < 4; ++jp) {
            const f32x2 tau2 = T2[c][jp];
            const f32x2 pt2 = mul2(fma2(tau2, ch, c1), tau2);
            model2[jp] = fma2(fma2(dx2[jp], Gp2, G02), pt2, model2[jp]);                // inference.py:60
          }
        } else if (tmax < 0.03125f) {
#pragma unroll
          for (int jp = 0; jp < 4; ++jp) {
            const f32x2 tau2 = T2[c][jp];
            f32x2 p2 = fma2(tau2, c24, c6);
            p2 = fma2(p2, tau2, ch);
            p2 = fma2(p2, tau2, c1);
            model2[jp] = fma2(fma2(dx2[jp], Gp2, G02), mul2(p2, tau2), model2[jp]);
          }
        } else {
#pragma unroll
          for (int jp = 0; jp < 4; ++jp) {
            float t0, t1;
            upk2(T2[c][jp], t0, t1);
            const f32x2 e2 = pk2(one_minus_exp_neg(t0), one_minus_exp_neg(t1));
            model2[jp] = fma2(fma2(dx2[jp], Gp2, G02), e2, model2[jp]);
          }
        }
      }
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) CHA_EPILOGUE_PAIR(model2[jp])
    }
#undef CHA_EPILOGUE_PAIR
    // the group's chi-square: 8 non-negative fp32 terms, one fp64 add per group
    float s0, s1;
    upk2(acc2, s0, s1);
    chi += chi_group_to_double(s0 + s1, smax);
  }
  return smax >= 0x7f800000u ? (double)INFINITY : chi;
}

// Which rows of a batch a launch serves when two list sets are resident (see chi2_mixed_kernel, finalize_kernel)
struct RowSplit {
  const int* split;     // device int: first row of the wide side (a multiple of kWalkersPerBlock); nullptr = one list
  int row_offset;       // row of this launch's first walker in the whole batch (chunked batches)
  int side;             // finalize_kernel only (0)
  const unsigned long long* void_flag;   // sampler: set once a queued half-step was not covered by the lists; every
                                         // later queued evaluation is void (the host re-runs it) and leaves at once
};

// One resident group / record / tile list set (device pointers)
struct ListsDev {
  const TileG* tiles; const GroupBlk* groups; const LineRec* recs;
  int n_tiles;
  float hv;             // half-width (km/s) of its line windows
};

// per-(walker, tile) state of the fused kernels
template <int K>
struct WalkerTile {
  float a, hw, sc[K], ncol[kMaxM][K], gc[K][4];
  double a2, cT;
  bool live, fast_ok;
};

// Prologue per (tile, walker), all fp64: (i) e0 = exp(h nu_c/(k Tex)) once, (ii) the line strengths of the tile's lines
// straight into this walker's shared-memory column tau_col[k * col_stride] (no HBM table; a thread reads only its
// own column, so no barrier separates producer and consumer), (iii) the interpolant of G_c over the tile.
// ok bits: 1 live, 2 the 10 dV mask is a no-op within kZcut sigma, 4 column densities >= 0 and Tex off Tbg
template <int K>
__device__ __forceinline__ void walker_tile_setup(WalkerTile<K>& W, int w, int nwp, const ModelDev& md,
                                                  const int* __restrict__ ok, const float* __restrict__ wpf,
                                                  const double* __restrict__ wpd, const TileG& tile, bool staged,
                                                  const LinesDev& ln, float* __restrict__ tau_col, int col_stride) {
  const int flags = ok[w];
  W.live = (flags & 1) != 0;
  W.fast_ok = (flags & 6) == 6;
  W.a = 0.f; W.hw = 0.f;
  // Planck exponent per MHz h*1e6/(k*Tex) and the Boltzmann exponent scale -log2(e)/(0.695*Tex)
  const double cT = W.live ? wpd[w] : 1.0;
  const double a2 = cT * (-1.4426950408889634 * kK / (kBoltzLit * kH * 1e6));
  W.cT = cT; W.a2 = a2;
  if (!W.live) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      W.sc[c] = 0.f;
#pragma unroll
      for (int m = 0; m < kMaxM; ++m) W.ncol[m][c] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) W.gc[c][k] = 0.f;
    }
    if (staged) for (int k = 0; k < tile.nline; ++k) tau_col[k * col_stride] = 0.0f;
    return;
  }
  W.a = wpf[w];
  W.hw = wpf[(size_t)nwp + w];
  double ss2[K];
#pragma unroll
  for (int c = 0; c < K; ++c) {
    W.sc[c] = wpf[(size_t)(2 + c) * nwp + w];
    ss2[c] = wpd[(size_t)(1 + c) * nwp + w];
#pragma unroll
    for (int m = 0; m < kMaxM; ++m) W.ncol[m][c] = m < md.M ? wpf[(size_t)(2 + K + m * K + c) * nwp + w] : 0.f;
  }
  // one fp64 exp per (walker, tile): e0 serves the Planck function of the G interpolant AND the stimulated-emission
  // factor of the tile's lines
  const double e0 = exp(cT * tile.xc);
  const double inv_e0 = fast_rcp(e0);
  if (staged) {
    // 1 - exp(-x_i) from e0: exp(-x_i) = exp(-x_c) exp(-z), z = (h/kTex)(nu_i - x_c); tiles span at most 0.4 % in
    // frequency, so |z| < 0.01 and the cubic Taylor polynomial of exp(-z) is exact to z^4/24 < 5e-10 (relative to
    // 1 - exp(-x) >= x_c/2: < 1e-8 for x_c > 0.1; the full series is used when that does not hold).
    // line constants are warp-uniform broadcast loads
    const bool near_lines = cT * tile.line_span < 0.01 && cT * tile.xc > 0.02;
    const double* __restrict__ Kp = ln.Kfac + tile.line0;
    const double* __restrict__ Ep = ln.El + tile.line0;
    const double* __restrict__ Np = ln.nu + tile.line0;
    if (md.M == 1 && near_lines && W.fast_ok) {
      // the common case, branch-free (fast_ok: Tex > 2.7 K, so |t| = El/(0.695 Tex) log2(e) stays far below 2^31)
#if CHA_STRENGTH_MUFU
      // strength in log2 space: log2 K_i + El_i a2 - log2(Q dV) formed in fp64, split into integer and fraction, the
      // fraction through MUFU.EX2 (|g| <= 1/2; 2^-22 relative, next to the fp32 rounding the strength ends in anyway),
      // the integer part through the exponent bits: 6 fp64 instructions per line instead of 20
      const double lq = ln.qinv[(size_t)nwp + w];
      const double* __restrict__ Lp = ln.lK2 + tile.line0;
#pragma unroll 2
      for (int k = 0; k < tile.nline; ++k) {
        const double tt = fma(Ep[k], a2, Lp[k]) + lq;                                 // classes.py:349-354
        const double tm = tt + kPow2Consts[0];
        const int n = __double2loint(tm);
        const float g = (float)(tt - (tm - kPow2Consts[0]));
        const float sc2 = __int_as_float(max(min(n + 127, 254), 0) << 23);            // 2^n (0 below the fp32 range)
        const double u = cT * (tile.xc - Np[k]);                                      // -z
        const double ez = fma(u, fma(0.5 * u, fma(u, kInvFact[3] * 2.0, 1.0), 1.0), 1.0);   // 1 + u + u^2/2 + u^3/6
        const float stim = (float)fma(-inv_e0, ez, 1.0);                              // classes.py:351
        tau_col[k * col_stride] = ex2_approx(g) * sc2 * stim;
      }
#else
      const double q = ln.qinv[w];
#pragma unroll 2
      for (int k = 0; k < tile.nline; ++k) {
        const double t = Ep[k] * a2;                                                  // classes.py:349
        const double u = cT * (tile.xc - Np[k]);                                      // -z
        const double ez = fma(u, fma(0.5 * u, fma(u, kInvFact[3] * 2.0, 1.0), 1.0), 1.0);   // 1 + u + u^2/2 + u^3/6
        const double stim = fma(-inv_e0, ez, 1.0);                                    // classes.py:351
        tau_col[k * col_stride] = (float)(Kp[k] * pow2_clamped(t) * stim * q);
      }
#endif
    } else {
      double qi[kMaxM];
#pragma unroll
      for (int m = 0; m < kMaxM; ++m) qi[m] = m < md.M ? ln.qinv[(size_t)m * nwp + w] : 0.0;
      for (int k = 0; k < tile.nline; ++k) {
        const int m = ln.mol[tile.line0 + k];
        const double q = m == 0 ? qi[0] : (m == 1 ? qi[1] : (m == 2 ? qi[2] : qi[3]));
        float st;
        if (near_lines) {
          const double u = cT * (tile.xc - Np[k]);
          const double ez = fma(u, fma(0.5 * u, fma(u, kInvFact[3] * 2.0, 1.0), 1.0), 1.0);
          st = line_strength_stim(Kp[k], Ep[k], a2, fma(-inv_e0, ez, 1.0), q);
        } else {
          st = line_strength(Kp[k], Ep[k], Np[k], a2, cT, q);
        }
        tau_col[k * col_stride] = st;
      }
    }
  }
  // interpolant of G_c(x) = (J(x,Tex) - J(x,Tbg)) * ss_c^2/(beam(x)^2 + ss_c^2) over the tile in tn = (x-xc)/hs:
  // Taylor factors of e0 at the nodes, MUFU.RCP+Newton reciprocals.  Narrow tiles (hs/xc < 5e-5, every tile of a
  // GOTHAM-like window grid): linear through the two end points (curvature term < 3e-9); else cubic through
  // 4 Chebyshev nodes.
  const double dmax = cT * tile.hs;
  const bool narrow = tile.hs <= 5e-5 * tile.xc;             // block-uniform
  // G_c at xc + dxn given the walker-independent J(x,Tbg) and beam^2 there
#define CHA_G_NODE(DXN, JBG, BEAM2, OUT)                                                                        \
  {                                                                                                             \
    const double dxn_ = (DXN);                                                                                  \
    const double xn_ = tile.xc + dxn_;                                                                          \
    double en_;                                                                                                 \
    if (fabs(dmax) < 0.01) {                                 /* |z| < 0.01: degree-6 Taylor, error < 1e-18 */   \
      const double z = cT * dxn_;                                                                               \
      en_ = e0 * fma(z, fma(z, fma(z, fma(z, fma(z, fma(z, kInvFact[6], kInvFact[5]), kInvFact[4]), kInvFact[3]), kInvFact[2]), kInvFact[1]), kInvFact[0]); \
    } else {                                                                                                    \
      en_ = exp(cT * xn_);                                                                                      \
    }                                                                                                           \
    const double dJ_ = (kHk * xn_) * fast_rcp(en_ - 1.0 + md.eps) - (JBG);        /* inference.py:56-57 */      \
    _Pragma("unroll") for (int c = 0; c < K; ++c) OUT[c] = dJ_ * ss2[c] * fast_rcp((BEAM2) + ss2[c]);  /* inference.py:39 */ \
  }
  if (narrow) {
    double Gh[K], Gl[K];
    CHA_G_NODE(tile.hs, tile.jbg_hi, tile.beam2_hi, Gh)
    CHA_G_NODE(-tile.hs, tile.jbg_lo, tile.beam2_lo, Gl)
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double g0 = 0.5 * (Gh[c] + Gl[c]), g1 = 0.5 * (Gh[c] - Gl[c]);
      W.gc[c][0] = (float)g0; W.gc[c][1] = (float)g1; W.gc[c][2] = 0.f; W.gc[c][3] = 0.f;
    }
  } else {
    double G0n[K], G1n[K], G2n[K], G3n[K];
    CHA_G_NODE(tile.hs * kChebNodes[0], tile.jbg[0], tile.beam2[0], G0n)
    CHA_G_NODE(tile.hs * kChebNodes[1], tile.jbg[1], tile.beam2[1], G1n)
    CHA_G_NODE(tile.hs * kChebNodes[2], tile.jbg[2], tile.beam2[2], G2n)
    CHA_G_NODE(tile.hs * kChebNodes[3], tile.jbg[3], tile.beam2[3], G3n)
#pragma unroll
    for (int c = 0; c < K; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double gsum = fma(kChebInv[k][3], G3n[c], fma(kChebInv[k][2], G2n[c], fma(kChebInv[k][1], G1n[c], kChebInv[k][0] * G0n[c])));
        W.gc[c][k] = (float)gsum;
      }
  }
#undef CHA_G_NODE
}

// general (scalar, masked) evaluation of one walker of a block that could not take a fast path
template <int K>
__device__ __forceinline__ double walker_tile_general(WalkerTile<K>& W, int w, int nwp, const ModelDev& md,
                                                      const GroupBlk* s_grp, const TileG& tile, const LineRec* rbase,
                                                      const LinesDev& ln, float inv_hs) {
  return chi2_mixed_groups<K>(s_grp, tile.ng, rbase, md.M, nwp, w, ln, W.a2, W.cT, W.a, W.sc, W.hw, W.ncol, W.gc, inv_hs);
}

template <int K>
__global__ void __launch_bounds__(kWalkersPerBlock, K == 1 ? 8 : (K == 2 ? 5 : (K <= 4 ? CHA_K4_BLOCKS : 2)))
chi2_mixed_kernel(int nwp, ModelDev md, const int* __restrict__ ok, const float* __restrict__ wpf,
                  const double* __restrict__ wpd, const ListsDev L0, const ListsDev L1, const LinesDev ln,
                  double* __restrict__ partial, RowSplit rs) {
  // Two resident list sets (sampler: bulk / outliers): the batch is ordered so that rows below *rs.split belong to the
  // narrow set L0 and the rows from it on to the wide set L1.  ONE launch serves both: blockIdx.x < L0.n_tiles is a
  // tile of L0, the rest are tiles of L1, and a (tile, walker block) pair whose rows belong to the other side leaves at
  // once -- so the few blocks of outliers run beside the bulk instead of after it.  The boundary is a multiple of the
  // block size (reach_sort_kernel pads it).  rs.split == nullptr: one list set, L0.
  if (rs.void_flag && *rs.void_flag != 0ull) return;
  int tile_idx = (int)blockIdx.x;
  const bool wide = rs.split && tile_idx >= L0.n_tiles;
  if (rs.split) {
    const int row0 = rs.row_offset + (int)blockIdx.y * kWalkersPerBlock;
    const int sp = *rs.split;
    if (wide ? row0 < sp : row0 >= sp) return;
    if (wide) tile_idx -= L0.n_tiles;
  }
  const TileG* __restrict__ tiles = wide ? L1.tiles : L0.tiles;
  const GroupBlk* __restrict__ groups = wide ? L1.groups : L0.groups;
  const LineRec* __restrict__ recs = wide ? L1.recs : L0.recs;
  const float hv_list = wide ? L1.hv : L0.hv;
  __shared__ __align__(128) GroupBlk s_grp[kTileMaxGroups];
  __shared__ __align__(128) LineRec s_rec[kTileMaxRecs + 1];
  __shared__ float s_tau[kTileMaxLines][kWalkersPerBlock];
  __shared__ __align__(8) unsigned long long s_bar;
  const int w = blockIdx.y * kWalkersPerBlock + threadIdx.x;
  const TileG tile = tiles[tile_idx];
  const bool staged = tile.rec_count <= kTileMaxRecs && tile.nline <= kTileMaxLines;
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gbytes = (unsigned)tile.ng * (unsigned)sizeof(GroupBlk);
    // + 1: the record after the tile's last one (next tile's first, or the host's trailing dummy) for the look-ahead
    const unsigned rbytes = (staged ? (unsigned)tile.rec_count + 1u : 0u) * (unsigned)sizeof(LineRec);
    mbar_expect_tx(&s_bar, gbytes + rbytes);
    bulk_g2s(s_grp, groups + tile.g0, gbytes, &s_bar);
    if (rbytes) bulk_g2s(s_rec, recs + tile.rec_begin, rbytes, &s_bar);
  }
  // ---- per-walker setup (overlaps the bulk copies) ----
  WalkerTile<K> W;
  walker_tile_setup<K>(W, w, nwp, md, ok, wpf, wpd, tile, staged, ln, &s_tau[0][threadIdx.x], kWalkersPerBlock);
  const float inv_hs = (float)(1.0 / tile.hs);
  // The code path is a function of the WALKER and the tile alone -- never of the walkers it shares a block with: a
  // fast_ok walker always runs the packed fast variant on a staged tile, any other live walker the general one
  // (a warp holding both kinds runs the two variants one after the other: rare, such rows sit at the mask edge,
  // have a negative column density or Tex <= Tbg).  Log-probs therefore do not depend on the position in the batch,
  // which the reach-sorted sampler batches and the sharding-independence of the chains rely on.
  const bool take_fast = staged && W.fast_ok;
  // K == 1: warps holding a walker whose own 6-sigma reach is well inside the list's windows take the variant that
  // tests every record against the walker's reach (either variant is exact for every walker; the vote only keeps a
  // warp on one code path)
  const bool skip_self = K == 1 && W.live && take_fast && (fabsf(W.sc[0]) + kVcut) < 0.8f * hv_list * W.a;
  const bool skip_warp = __any_sync(0xffffffffu, skip_self);
  const float vcut1 = skip_self ? kVcut : INFINITY;
  mbar_wait(&s_bar, 0);
  double chi = 0.0;
  // 1 < K <= 4: the G interpolants (4 K coefficients) live in this walker's column of a shared table, not in registers
  // (K >= 5 runs two CTAs per SM with registers to spare, and the static shared-memory limit is 48 KB)
  __shared__ float s_gc[kGcInSmem<K> ? 4 * K : 1][kGcInSmem<K> ? kWalkersPerBlock : 1];
  const float* gc_col = nullptr;
  if constexpr (kGcInSmem<K>) {
    if (W.live && take_fast) {
#pragma unroll
      for (int c = 0; c < K; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) s_gc[4 * c + k][threadIdx.x] = W.gc[c][k];
    }
    gc_col = &s_gc[0][threadIdx.x];
  }
  if (W.live) {
    if (take_fast && md.M == 1) {
      const bool narrow = tile.hs <= 5e-5 * tile.xc;       // same test as walker_tile_setup: linear G interpolant
      if (K == 1 && skip_warp)
        chi = narrow ? chi2_mixed_groups_fast1<K, true, true>(s_grp, tile.ng, s_rec, &s_tau[0][threadIdx.x], W.a, W.sc, W.ncol[0], W.gc, gc_col, inv_hs, vcut1)
                     : chi2_mixed_groups_fast1<K, false, true>(s_grp, tile.ng, s_rec, &s_tau[0][threadIdx.x], W.a, W.sc, W.ncol[0], W.gc, gc_col, inv_hs, vcut1);
      else
        chi = narrow ? chi2_mixed_groups_fast1<K, true, false>(s_grp, tile.ng, s_rec, &s_tau[0][threadIdx.x], W.a, W.sc, W.ncol[0], W.gc, gc_col, inv_hs, vcut1)
                     : chi2_mixed_groups_fast1<K, false, false>(s_grp, tile.ng, s_rec, &s_tau[0][threadIdx.x], W.a, W.sc, W.ncol[0], W.gc, gc_col, inv_hs, vcut1);
    } else if (take_fast) {
      if constexpr (K > 1) {
        // multi-molecule, multi-component fit: the walker's M x K column densities go to its column of a shared table
        __shared__ float s_ncol[kMaxM * K][kWalkersPerBlock];
#pragma unroll
        for (int m = 0; m < kMaxM; ++m)
#pragma unroll
          for (int c = 0; c < K; ++c) s_ncol[m * K + c][threadIdx.x] = W.ncol[m][c];
        chi = chi2_mixed_groups_fast<K>(s_grp, tile.ng, s_rec, md.M, &s_tau[0][threadIdx.x], W.a, W.sc, W.ncol, &s_ncol[0][threadIdx.x], W.gc, gc_col, inv_hs);
      } else {
        chi = chi2_mixed_groups_fast<K>(s_grp, tile.ng, s_rec, md.M, &s_tau[0][threadIdx.x], W.a, W.sc, W.ncol, nullptr, W.gc, gc_col, inv_hs);
      }
    } else {
      chi = walker_tile_general<K>(W, w, nwp, md, s_grp, tile, staged ? s_rec : recs + tile.rec_begin, ln, inv_hs);
    }
  }
  partial[(size_t)tile_idx * nwp + w] = chi;
}

// ------------------------------------------------------------------------------------------
// (3c) finalize: fixed-order sum over tiles, constants, prior, non-finite guard
//      (inference.py:160-166, 239-246)
// ------------------------------------------------------------------------------------------
constexpr int kFinSlices = 32;    // 32 walkers x 32 tile slices per block: the sum over ~300-800 tiles is latency bound, so many
                                  // short independent chains; the order of the additions is fixed (slice-wise, then slices 0..31)
__global__ void __launch_bounds__(32 * kFinSlices)
finalize_kernel(int nw, int nwp, int n_tiles, const double* __restrict__ partial,
                double chi_const, const int* __restrict__ ok, const double* __restrict__ lp,
                int with_prior, double* __restrict__ out,
                unsigned long long* __restrict__ need_dev, unsigned long long* __restrict__ need_host,
                RowSplit rs, int n_tiles_wide, double chi_const_wide,
                const int* __restrict__ inv /*nullptr, or [nw]: input row evaluated at slot w (reach-ordered batches)*/) {
  // The batch maxima walker_prep_kernel reduced into need_dev[0..1] are complete by now (earlier kernel of the same
  // stream): publish them to the host's pinned mirror (mapped memory) and leave the slot zeroed for its next use --
  // no memset / copy operation in the launch sequence.
  if (need_dev && blockIdx.x == 0 && threadIdx.x == 0) {
    need_host[0] = need_dev[0]; need_host[1] = need_dev[1];
    need_dev[0] = 0ull; need_dev[1] = 0ull;
  }
  // block = 32 walkers x kFinSlices tile slices; slice s sums tiles s, s + kFinSlices, ... ; slices combined in fixed order
  __shared__ double red[kFinSlices][32];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int w = blockIdx.x * 32 + lane;
  // two list sets: rows from *rs.split on were evaluated against the wide one (its tile count, its constant)
  if (rs.split && rs.row_offset + w >= *rs.split) { n_tiles = n_tiles_wide; chi_const = chi_const_wide; }
  double acc = 0.0;
  if (w < nw && ok[w])
    for (int t = sl; t < n_tiles; t += kFinSlices) acc += partial[(size_t)t * nwp + w];
  red[sl][lane] = acc;
  __syncthreads();
  if (sl != 0 || w >= nw) return;
  double res = -INFINITY;
  if (ok[w]) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < kFinSlices; ++k) tot += red[k][lane];
    tot += chi_const;
    double ll = -0.5 * tot;                                                     // inference.py:166
    if (isfinite(ll)) res = with_prior ? lp[w] + ll : ll;                       // inference.py:162-164, 246
  }
  out[inv ? inv[w] : w] = res;
}

__global__ void prior_only_kernel(int nw, const double* __restrict__ lp, double* __restrict__ out) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nw) out[w] = lp[w];
}

// ------------------------------------------------------------------------------------------
// (4) channel-stream kernel: the model spectrum make_model returns (inference.py:44-61),
//     lanes = channels so the fp64 stores are coalesced 256 B rows; HBM-write bound.
//     act_of[j] = active index of sorted channel j or -1; out_pos[j] = caller's channel index.
// ------------------------------------------------------------------------------------------
template <int K, bool MIXED>
__global__ void __launch_bounds__(256)
simulate_kernel(const double* __restrict__ theta, int nw, int nwp, ModelDev md, const int* __restrict__ ok,
                SpecDev sp, int n_chan, const int* __restrict__ act_of, const int* __restrict__ out_pos,
                const double* __restrict__ x_all,
                const void* __restrict__ tau0_v, double* __restrict__ out) {
  const int w = blockIdx.y;
  const int js = blockIdx.x * blockDim.x + threadIdx.x;
  if (js >= n_chan) return;
  double model = 0.0;
  const int j = act_of[js];
  if (ok[w] && j >= 0) {
    const double* th = theta + (size_t)w * md.ndim;
    const double dV = th[md.idx_dv], Tex = th[md.idx_tex];
    const double sig = dV / kFwhm, hw = dV * 10;
    const double x = x_all[js];
    const double dJ = planck_j(x, Tex, md.eps) - planck_j(x, kTbg, md.eps);
    const double bs = beam_size(x, md.dish);
    double acc[K];
#pragma unroll
    for (int c = 0; c < K; ++c) acc[c] = 0.0;
    for (int m = 0; m < md.M; ++m) {
      const int p0 = sp.pair_off[j * md.M + m], p1 = sp.pair_off[j * md.M + m + 1];
      if (MIXED) {
        const float* tau0 = (const float*)tau0_v;
        const double a64 = 0.84932180028801907 * kFwhm / dV;
        const float a = (float)a64, hwf = (float)hw;
        float S[K], sc[K];
#pragma unroll
        for (int c = 0; c < K; ++c) { S[c] = 0.f; sc[c] = (float)((th[md.idx_vlsr[c]] - md.al - md.mc) * a64); }
        for (int p = p0; p < p1; ++p) {
          const float u = sp.pair_u32[p];
          float t0 = tau0[(size_t)sp.pair_line[p] * nwp + w];
          t0 = (fabsf(u) < hwf) ? t0 : 0.0f;
#pragma unroll
          for (int c = 0; c < K; ++c) { float v = fmaf(u, a, -sc[c]); S[c] = fmaf(t0, ex2_approx(-v * v), S[c]); }
        }
#pragma unroll
        for (int c = 0; c < K; ++c) acc[c] += (double)((float)th[md.idx_ncol[m * md.K + c]] * S[c]);
      } else {
        const double* tau0 = (const double*)tau0_v;
        double S[K];
#pragma unroll
        for (int c = 0; c < K; ++c) S[c] = 0.0;
        for (int p = p0; p < p1; ++p) {
          const double u = sp.pair_u64[p];
          const double vg = u + md.al;
          if (fabs(vg - md.al - md.mc) < hw) {
            const double t0 = tau0[(size_t)sp.pair_line[p] * nwp + w];
#pragma unroll
            for (int c = 0; c < K; ++c) { double z = (vg - th[md.idx_vlsr[c]]) / sig; S[c] += t0 * exp(-0.5 * (z * z)); }
          }
        }
#pragma unroll
        for (int c = 0; c < K; ++c) acc[c] += th[md.idx_ncol[m * md.K + c]] * S[c];
      }
    }
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double ss = md.idx_ss[c] < 0 ? md.fixed_ss : th[md.idx_ss[c]];
      const double e = MIXED ? (double)one_minus_exp_neg((float)acc[c]) : 1.0 - exp(-acc[c]);
      model += dJ * e * (ss * ss / (bs * bs + ss * ss));
    }
  }
  out[(size_t)w * n_chan + out_pos[js]] = model;
}

// ------------------------------------------------------------------------------------------
// (4b) channel-stream kernel of the mixed path, tiled: the spectra make_model returns for every walker.
//      The output is zero-filled first (inactive channels: model == 0 exactly, a pure HBM write stream);
//      this kernel then fills the active channels.  block = one tile (<= 32 groups = 256 active channels) x
//      32 walkers.  Phase A: per-(walker, line) strengths and per-walker interpolants into shared memory.
//      Phase B: thread = channel (consecutive lanes = consecutive channels: coalesced fp64 stores), loop over
//      the block's walkers; the reference's 10 dV mask (inference.py:52) is applied per pair.
// ------------------------------------------------------------------------------------------
constexpr int kSimWalkers = 32;

template <int K>
__global__ void __launch_bounds__(256)
simulate_tiles_kernel(int nw, int nwp, ModelDev md, const int* __restrict__ ok, const float* __restrict__ wpf,
                      const double* __restrict__ wpd, const TileG* __restrict__ tiles,
                      const GroupBlk* __restrict__ groups, const LineRec* __restrict__ recs, const LinesDev ln,
                      size_t n_chan, double* __restrict__ out) {
  constexpr int kPar = 2 + K + kMaxM * K + 4 * K;            // a, 10 dV, sc[K], ncol[M][K], gc[K][4]
  __shared__ __align__(16) GroupBlk s_grp[kTileMaxGroups];
  __shared__ __align__(16) LineRec s_rec[kTileMaxRecs];
  __shared__ float s_tau[kTileMaxLines][kSimWalkers];
  __shared__ float s_par[kSimWalkers][kPar];
  __shared__ int s_live[kSimWalkers];
  const TileG tile = tiles[blockIdx.x];
  const int w0 = blockIdx.y * kSimWalkers;
  // stage the tile (plain cooperative copies: 16-byte words)
  {
    const uint4* src = reinterpret_cast<const uint4*>(groups + tile.g0);
    uint4* dst = reinterpret_cast<uint4*>(s_grp);
    for (int i = threadIdx.x; i < tile.ng * (int)(sizeof(GroupBlk) / 16); i += blockDim.x) dst[i] = src[i];
    const uint4* rs = reinterpret_cast<const uint4*>(recs + tile.rec_begin);
    uint4* rd = reinterpret_cast<uint4*>(s_rec);
    for (int i = threadIdx.x; i < tile.rec_count; i += blockDim.x) rd[i] = rs[i];
  }
  // phase A1: line strengths, thread = (walker, line)
  for (int idx = threadIdx.x; idx < kSimWalkers * tile.nline; idx += blockDim.x) {
    const int wl = idx % kSimWalkers, k = idx / kSimWalkers, w = w0 + wl;
    float v = 0.0f;
    if (w < nw && (ok[w] & 1)) {
      const double cT = wpd[w];
      const double a2 = cT * (-1.4426950408889634 * kK / (kBoltzLit * kH * 1e6));
      const int i = tile.line0 + k;
      v = line_strength(ln.Kfac[i], ln.El[i], ln.nu[i], a2, cT, ln.qinv[(size_t)ln.mol[i] * nwp + w]);
    }
    s_tau[k][wl] = v;
  }
  // phase A2: per-walker constants and the cubic interpolant of G_c over the tile, thread = walker
  if (threadIdx.x < kSimWalkers) {
    const int wl = threadIdx.x, w = w0 + wl;
    const bool live = w < nw && (ok[w] & 1);
    s_live[wl] = live ? 1 : 0;
    if (live) {
      float* par = s_par[wl];
      par[0] = wpf[w];
      par[1] = wpf[(size_t)nwp + w];
      const double cT = wpd[w];
      double ss2[K];
#pragma unroll
      for (int c = 0; c < K; ++c) {
        par[2 + c] = wpf[(size_t)(2 + c) * nwp + w];
        ss2[c] = wpd[(size_t)(1 + c) * nwp + w];
        for (int m = 0; m < kMaxM; ++m) par[2 + K + m * K + c] = m < md.M ? wpf[(size_t)(2 + K + m * K + c) * nwp + w] : 0.f;
      }
      double Gn[K][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const double xn = tile.xc + tile.hs * kChebNodes[n];
        const double dJ = (kHk * xn) / (exp(cT * xn) - 1.0 + md.eps) - tile.jbg[n];   // inference.py:56-57
#pragma unroll
        for (int c = 0; c < K; ++c) Gn[c][n] = dJ * ss2[c] / (tile.beam2[n] + ss2[c]);  // inference.py:39
      }
#pragma unroll
      for (int c = 0; c < K; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double gsum = 0.0;
#pragma unroll
          for (int n = 0; n < 4; ++n) gsum = fma(kChebInv[k][n], Gn[c][n], gsum);
          par[2 + K + kMaxM * K + 4 * c + k] = (float)gsum;
        }
    }
  }
  __syncthreads();
  // phase B: thread = channel of the tile.  The block's walkers are taken in sub-batches of kSimSub: a record (shared
  // by every walker) is fetched and its velocity offset formed once per sub-batch, the optical depths of the
  // sub-batch's walkers live in registers, per-walker parameters are warp-uniform shared-memory loads
  constexpr int kSimSub = K == 1 ? 8 : (K == 2 ? 4 : 2);
  const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
  if (g >= tile.ng) return;
  const GroupBlk& gb = s_grp[g];
  const int opos = gb.opos[j];
  if (opos < 0) return;
  const float dx = gb.dx[j];
  const float tn = gb.tn0 + dx * (float)(1.0 / tile.hs);
  int nrec_m[kMaxM];
#pragma unroll
  for (int m = 0; m < kMaxM; ++m) nrec_m[m] = m < md.M ? gb.nrec[m] : 0;
  double* po = out + (size_t)w0 * n_chan + opos;
  for (int wb = 0; wb < kSimWalkers && w0 + wb < nw; wb += kSimSub) {
    float T[kSimSub][K];
#pragma unroll
    for (int i = 0; i < kSimSub; ++i)
#pragma unroll
      for (int c = 0; c < K; ++c) T[i][c] = 0.0f;
    int r = gb.rec_off;
#pragma unroll
    for (int m = 0; m < kMaxM; ++m) {
      for (int q = 0; q < nrec_m[m]; ++q, ++r) {
        const LineRec rc = s_rec[r];
        const float u = fmaf(-dx, rc.slope, rc.u0);                                      // inference.py:51
        const float au = fabsf(u);
        const float* trow = &s_tau[rc.lloc / kWalkersPerBlock][wb];
#pragma unroll
        for (int i = 0; i < kSimSub; ++i) {
          const float* par = s_par[wb + i];
          const float t0 = au < par[1] ? trow[i] : 0.0f;                                 // inference.py:52
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const float v = fmaf(u, par[0], -par[2 + c]);
            T[i][c] = fmaf(t0 * par[2 + K + m * K + c], ex2_approx(-v * v), T[i][c]);    // inference.py:53
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kSimSub; ++i) {
      const int wl = wb + i;
      if (w0 + wl < nw) {
        float model = 0.0f;
        if (s_live[wl]) {
          const float* par = s_par[wl];
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const float* gcf = par + 2 + K + kMaxM * K + 4 * c;
            const float G = fmaf(fmaf(fmaf(gcf[3], tn, gcf[2]), tn, gcf[1]), tn, gcf[0]);
            model = fmaf(G, one_minus_exp_neg(T[i][c]), model);                          // inference.py:60
          }
        }
        po[(size_t)wl * n_chan] = (double)model;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// (4b) channel-stream path, one pass over the output, every byte written exactly once (spectra in
//      ascending channel order, the usual case).
//      sim_line_tau_kernel / sim_gcoef_kernel: the walker-dependent tables (line strengths tau0[line][walker],
//      cubic interpolant of G_c per (tile, walker)) -- the same arithmetic as phase A of simulate_tiles_kernel,
//      done once per launch instead of once per CTA, so the streaming kernel below carries no fp64 latency chain.
//      simulate_span_kernel: a CTA owns a SPAN of kSpanCh consecutive channels x 32 walkers.  It keeps a dense
//      buffer [kSpanRows walkers][kSpanCh channels] of fp64 in shared memory, zeroed once; for each sub-block of
//      kSpanRows walkers the active channels of the span (the tiles that intersect it: same group / record lists
//      and the same arithmetic as phase B of simulate_tiles_kernel) are overwritten in place -- the active
//      positions of a span do not depend on the walker, so the zeros between them are never rewritten -- and
//      every row leaves as ONE cp.async.bulk shared -> global store (TMA, SASS UBLKCP) of kSpanCh * 8 bytes.
//      No zero-fill pass, no memset in the launch sequence; HBM sees each byte of the output once.
// ------------------------------------------------------------------------------------------
#ifndef CHA_SPAN_CH
#define CHA_SPAN_CH 512
#endif
#ifndef CHA_SPAN_ROWS
#define CHA_SPAN_ROWS 8
#endif
#ifndef CHA_SPAN_BUFS
#define CHA_SPAN_BUFS 1
#endif
constexpr int kSpanCh = CHA_SPAN_CH;       // channels per span: 4 KB rows
constexpr int kSpanRows = CHA_SPAN_ROWS;   // walkers per sub-block (one bulk store each)
constexpr int kSpanBufs = CHA_SPAN_BUFS;   // row buffers: a sub-block is computed while the previous one drains
constexpr int kSpanWalkers = 32;           // walkers per CTA
constexpr int kSpanDynSmem = kSpanBufs * kSpanRows * kSpanCh * 8;

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// thread = (walker, line): line strengths exactly as phase A1 of simulate_tiles_kernel forms them
__global__ void __launch_bounds__(kWalkersPerBlock)
sim_line_tau_kernel(int nwp, const int* __restrict__ ok, const double* __restrict__ wpd, const LinesDev ln, int n_lines,
                    float* __restrict__ tau0, int lines_per_block) {
  const int w = blockIdx.x * kWalkersPerBlock + threadIdx.x;
  const int i0 = blockIdx.y * lines_per_block, i1 = min(n_lines, i0 + lines_per_block);
  const bool live = (ok[w] & 1) != 0;
  const double cT = live ? wpd[w] : 1.0;
  const double a2 = cT * (-1.4426950408889634 * kK / (kBoltzLit * kH * 1e6));
  for (int i = i0; i < i1; ++i)
    tau0[(size_t)i * nwp + w] = live ? line_strength(ln.Kfac[i], ln.El[i], ln.nu[i], a2, cT, ln.qinv[(size_t)ln.mol[i] * nwp + w]) : 0.0f;
}

// thread = (walker, tile): cubic interpolant of G_c over the tile, exactly as phase A2 of simulate_tiles_kernel
template <int K>
__global__ void __launch_bounds__(kWalkersPerBlock)
sim_gcoef_kernel(int nwp, ModelDev md, const int* __restrict__ ok, const double* __restrict__ wpd,
                 const TileG* __restrict__ tiles, float* __restrict__ gco) {
  const int w = blockIdx.x * kWalkersPerBlock + threadIdx.x;
  const int t = blockIdx.y;
  float* o = gco + ((size_t)t * nwp + w) * (4 * K);
  if (!(ok[w] & 1)) {
#pragma unroll
    for (int k = 0; k < 4 * K; ++k) o[k] = 0.f;
    return;
  }
  const TileG* tp = tiles + t;
  const double cT = wpd[w], xc = tp->xc, hs = tp->hs;
  double dJ[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const double xn = xc + hs * kChebNodes[n];
    dJ[n] = (kHk * xn) / (exp(cT * xn) - 1.0 + md.eps) - tp->jbg[n];                   // inference.py:56-57
  }
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const double ss2 = wpd[(size_t)(1 + c) * nwp + w];
    double Gn[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) Gn[n] = dJ[n] * ss2 / (tp->beam2[n] + ss2);             // inference.py:39
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double gsum = 0.0;
#pragma unroll
      for (int n = 0; n < 4; ++n) gsum = fma(kChebInv[k][n], Gn[n], gsum);
      o[4 * c + k] = (float)gsum;
    }
  }
}

// The part of one tile that lies in one span (host: build_pairs): groups [g_lo, g_lo + g_n) of the group array, their
// records [r_lo, r_lo + r_n) (contiguous), the selected lines [l_lo, l_lo + l_n) those records reference.
struct __align__(16) SpanSeg {
  int tile, g_lo, g_n, r_lo;
  int r_n, l_lo, l_n, rec_shift;   // rec_shift: tile.rec_begin - r_lo (group.rec_off is relative to the tile)
  int line_shift;                  // tile.line0 - l_lo (record.lloc is relative to the tile)
  float inv_hs;                    // 1 / tile.hs
  int pad[2];
};
static_assert(sizeof(SpanSeg) == 48, "SpanSeg must be 48 bytes");
// a segment is cut (host) so that it fits the CTA's small staging area; a span whose part of a tile is larger holds
// several segments (restaged per sub-block: slower, and rare on the sparse grids this kernel is used for)
#ifndef CHA_SEG_GROUPS
#define CHA_SEG_GROUPS 16
#endif
#ifndef CHA_SEG_RECS
#define CHA_SEG_RECS 128
#endif
constexpr int kSegGroups = CHA_SEG_GROUPS;
constexpr int kSegRecs = CHA_SEG_RECS;

template <int K>
__global__ void __launch_bounds__(256)
simulate_span_kernel(int nw, int nwp, ModelDev md, const int* __restrict__ ok, const float* __restrict__ wpf,
                     const GroupBlk* __restrict__ groups, const LineRec* __restrict__ recs,
                     const float* __restrict__ tau0, const float* __restrict__ gco,
                     const int* __restrict__ span_off, const SpanSeg* __restrict__ segs, size_t n_chan,
                     double* __restrict__ out) {
  constexpr int kPar = 2 + K + kMaxM * K + 4 * K;            // a, 10 dV, sc[K], ncol[M][K], gc[K][4]
  extern __shared__ __align__(128) unsigned char span_dyn[];
  double (*Ball)[kSpanRows][kSpanCh] = reinterpret_cast<double (*)[kSpanRows][kSpanCh]>(span_dyn);   // [kSpanBufs][kSpanRows][kSpanCh]
  __shared__ __align__(16) GroupBlk s_grp[kSegGroups];
  __shared__ __align__(16) LineRec s_rec[kSegRecs];
  __shared__ float s_tau[kTileMaxLines][kSpanWalkers];
  __shared__ float s_par[kSpanWalkers][kPar];
  __shared__ unsigned short s_act[kSegGroups * kGroupCh], s_pos[kSegGroups * kGroupCh];
  __shared__ int s_nact;
  static_assert(kSegGroups * kGroupCh <= 65536 && kSpanCh <= 65536, "codes and positions fit 16 bits");
  const int tid = threadIdx.x;
  const int c0 = (int)blockIdx.x * kSpanCh;
  const int nch = min(kSpanCh, (int)(n_chan - (size_t)c0));
  const int sg_lo = span_off[blockIdx.x], sg_hi = span_off[blockIdx.x + 1];      // segments of this span
  const int wbase = (int)blockIdx.y * kSpanWalkers;
  const int nrow = min(kSpanWalkers, nw - wbase);
  if (sg_lo >= sg_hi) {
    // no active channel in the span: one zero row serves every walker's store
    for (int i = tid; i < kSpanCh / 2; i += 256) reinterpret_cast<uint4*>(span_dyn)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < nrow) {
      bulk_s2g(out + (size_t)(wbase + tid) * n_chan + c0, span_dyn, (unsigned)nch * 8u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  SpanSeg sg = segs[sg_lo];                                   // (its latency overlaps the zero-fill)
  for (int i = tid; i < kSpanDynSmem / 16; i += 256) reinterpret_cast<uint4*>(span_dyn)[i] = make_uint4(0u, 0u, 0u, 0u);
  const int n_it = (nrow + kSpanRows - 1) / kSpanRows;
  int staged = -1;
  for (int it = 0; it < n_it; ++it) {
    const int r0 = it * kSpanRows;                           // first row (walker of the CTA) of this sub-block
    const int rows = min(kSpanRows, nrow - r0);
    double (*B)[kSpanCh] = Ball[it % kSpanBufs];
    // the rows stored from this buffer kSpanBufs sub-blocks ago must have been read by the copy engine
    if (it >= kSpanBufs && tid < kSpanRows) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kSpanBufs - 1) : "memory");
    for (int si = sg_lo; si < sg_hi; ++si) {
      __syncthreads();                                       // previous phase B is done with the staged tables
      if (si != staged) {
        // a span inside one tile (the usual case) stages its segment and the tables of all 32 walkers once
        if (si != sg_lo || staged >= 0) sg = segs[si];
        const uint4* src = reinterpret_cast<const uint4*>(groups + sg.g_lo);
        uint4* dst = reinterpret_cast<uint4*>(s_grp);
        for (int i = tid; i < sg.g_n * (int)(sizeof(GroupBlk) / 16); i += 256) dst[i] = src[i];
        const uint4* rs = reinterpret_cast<const uint4*>(recs + sg.r_lo);
        uint4* rd = reinterpret_cast<uint4*>(s_rec);
        for (int i = tid; i < sg.r_n; i += 256) rd[i] = rs[i];
        for (int idx = tid; idx < kSpanWalkers * sg.l_n; idx += 256) {
          const int wl = idx % kSpanWalkers, k = idx / kSpanWalkers;
          s_tau[k][wl] = wl < nrow ? tau0[(size_t)(sg.l_lo + k) * nwp + wbase + wl] : 0.0f;
        }
        for (int idx = tid; idx < kSpanWalkers * kPar; idx += 256) {
          const int wl = idx / kPar, q = idx % kPar, w = wbase + wl;
          float v = 0.0f;
          if (wl < nrow && (ok[w] & 1)) {
            if (q < 2 + K) v = wpf[(size_t)q * nwp + w];                                   // a, 10 dV, sc[c]
            else if (q < 2 + K + kMaxM * K) {
              const int m = (q - 2 - K) / K, c = (q - 2 - K) % K;
              v = m < md.M ? wpf[(size_t)(2 + K + m * K + c) * nwp + w] : 0.f;             // ncol[m][c]
            } else v = gco[((size_t)sg.tile * nwp + w) * (4 * K) + (q - 2 - K - kMaxM * K)];   // gc[c][k]
          }
          s_par[wl][q] = v;
        }
        staged = si;
        if (tid == 0) s_nact = 0;
        __syncthreads();
        // the segment's channels that lie in this span (walker independent): compacted once per staged segment
        for (int code = tid; code < sg.g_n * kGroupCh; code += 256) {
          const int opos = s_grp[code >> 3].opos[code & 7];
          if (opos >= c0 && opos < c0 + nch) {
            const int a = atomicAdd(&s_nact, 1);
            s_act[a] = (unsigned short)code; s_pos[a] = (unsigned short)(opos - c0);
          }
        }
        __syncthreads();
      }
      // phase B: thread = (active channel of the span, walker of the sub-block) -- a span holds a few dozen active
      // channels, so one thread per channel would leave most of the CTA idle behind a long serial chain.  Same
      // arithmetic per (channel, walker) as phase B of simulate_tiles_kernel.
      const int n_act = s_nact;
      for (int item = tid; item < n_act * kSpanRows; item += 256) {
        const int a = item / kSpanRows, wl = item % kSpanRows;
        if (wl >= rows) continue;
        const int code = s_act[a];
        const int g = code >> 3, j = code & 7;
        const GroupBlk& gb = s_grp[g];
        const float dx = gb.dx[j];
        const float tn = gb.tn0 + dx * sg.inv_hs;
        const float* par = s_par[r0 + wl];
        float T[K];
#pragma unroll
        for (int c = 0; c < K; ++c) T[c] = 0.0f;
        int r = gb.rec_off + sg.rec_shift;
#pragma unroll
        for (int m = 0; m < kMaxM; ++m) {
          const int nrec = m < md.M ? gb.nrec[m] : 0;
          for (int q = 0; q < nrec; ++q, ++r) {
            const LineRec rc = s_rec[r];
            const float u = fmaf(-dx, rc.slope, rc.u0);                                    // inference.py:51
            const float t0 = fabsf(u) < par[1] ? s_tau[rc.lloc / kWalkersPerBlock + sg.line_shift][r0 + wl] : 0.0f;   // inference.py:52
#pragma unroll
            for (int c = 0; c < K; ++c) {
              const float v = fmaf(u, par[0], -par[2 + c]);
              T[c] = fmaf(t0 * par[2 + K + m * K + c], ex2_approx(-v * v), T[c]);          // inference.py:53
            }
          }
        }
        float model = 0.0f;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const float* gcf = par + 2 + K + kMaxM * K + 4 * c;
          const float G = fmaf(fmaf(fmaf(gcf[3], tn, gcf[2]), tn, gcf[1]), tn, gcf[0]);
          model = fmaf(G, one_minus_exp_neg(T[c]), model);                                 // inference.py:60
        }
        B[wl][s_pos[a]] = (double)model;
      }
    }
    // generic-proxy writes of the rows (zero-fill included) become visible to the async proxy, then one store per row
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < kSpanRows) {
      if (tid < rows) bulk_s2g(out + (size_t)(wbase + r0 + tid) * n_chan + c0, &B[tid][0], (unsigned)nch * 8u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid < kSpanRows) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// (5) bookkeeping: exact number of (line, channel) pairs the reference's mask admits for a walker
//     (inference.py:52).  Integer atomics -> deterministic.  thread = (walker, line).
// ------------------------------------------------------------------------------------------
__global__ void count_window_pairs_kernel(const double* __restrict__ theta, int nw, int ndim, int idx_dv, double mc,
                                          int n_lines, const double* __restrict__ nu,
                                          int n_chan, const double* __restrict__ x_sorted,
                                          unsigned long long* __restrict__ out) {
  int w = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lines || w >= nw) return;
  double dV = theta[(size_t)w * ndim + idx_dv];
  if (!(dV > 0.0)) return;
  double f = nu[i];
  // |(f-x)/f*ckm - mc| < 10 dV  <=>  f*(1-(mc+10dV)/ckm) < x < f*(1-(mc-10dV)/ckm)
  double xlo = f * (1.0 - (mc + 10.0 * dV) / kCkm), xhi = f * (1.0 - (mc - 10.0 * dV) / kCkm);
  int lo = 0, hi = n_chan;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (x_sorted[mid] <= xlo) lo = mid + 1; else hi = mid; }
  int first = lo; hi = n_chan;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (x_sorted[mid] < xhi) lo = mid + 1; else hi = mid; }
  if (lo > first) atomicAdd(&out[w], (unsigned long long)(lo - first));
}


// ------------------------------------------------------------------------------------------
// (6) stand-alone restatements of two reference functions (setup / plotting, not the hot loop)
// ------------------------------------------------------------------------------------------
// MolSim.run_sim, gauss=False, one component: classes.py:347-377 over every catalog line
__global__ void stick_spectrum_kernel(int n, const double* __restrict__ nu, const double* __restrict__ logint,
                                      const double* __restrict__ elower, double q_ct, double Q, double ncol, double tex,
                                      double dv, double ss, double dish, double* __restrict__ out_tau,
                                      double* __restrict__ out_int) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f = nu[i], el = elower[i];
  const double eu = el + f / kMHzPerCm;
  const double sijmu = 1.0 / (exp(-(el / kBoltzLit) / kCT) - exp(-(eu / kBoltzLit) / kCT)) * (pow(10.0, logint[i]) / f) *
                       (1.0 / kSijConst) * q_ct;
  const double aij_gup = kAijConst * f * f * f * sijmu;
  const double nl = ncol * exp(-el / (kBoltzLit * tex)) / Q;                                     // classes.py:349 (/glow)
  const double lam = kCcm / (f * 1e6);
  const double num = lam * lam * aij_gup * nl * (1.0 - exp(-(kH * f * 1e6) / (kK * tex)));       // classes.py:351
  const double den = 8.0 * M_PI * (dv * f * 1e6 / kCkm);                                         // classes.py:353
  const double tau = num / den;
  const double jt = planck_j(f, tex, 0.0), jbg = planck_j(f, kTbg, 0.0);                         // classes.py:372-373
  const double b = beam_size(f, dish);
  out_tau[i] = tau;
  out_int[i] = (jt - jbg) * (1.0 - exp(-tau)) * (ss * ss / (b * b + ss * ss));                   // classes.py:375-377
}

// make_model_numba: inference.py:44-61 (one component, caller-supplied lines); thread = channel
__global__ void make_model_kernel(int L, const double* __restrict__ freqs, const double* __restrict__ taus, int C,
                                  const double* __restrict__ x, double vlsr, double dv, double tex, double ss, double al,
                                  double dish, double mc, double eps, double* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C) return;
  const double xj = x[j], sig = dv / kFwhm, hw = dv * 10;
  double acc = 0.0;
  for (int i = 0; i < L; ++i) {
    const double f = freqs[i];
    const double vg = (f - xj) / f * kCkm + al;                                                  // inference.py:51
    if (fabs(vg - al - mc) < hw) {                                                               // inference.py:52
      const double z = (vg - vlsr) / sig;
      acc += taus[i] * exp(-0.5 * (z * z));                                                      // inference.py:53
    }
  }
  const double dJ = planck_j(xj, tex, eps) - planck_j(xj, kTbg, eps);                            // inference.py:56-57
  const double b = beam_size(xj, dish);
  out[j] = dJ * (1.0 - exp(-acc)) * (ss * ss / (b * b + ss * ss));                               // inference.py:60
}

}  // namespace lte
