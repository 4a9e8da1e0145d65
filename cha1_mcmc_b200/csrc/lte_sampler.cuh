// lte_sampler.cuh -- on-device affine-invariant stretch move.
//
// Replaces emcee.EnsembleSampler.run_mcmc + StretchMove (call sites inference.py:456-473;
// emcee 3.1.6 itself is third-party and absent, algorithm restated from the published package:
// z = ((a-1)u+1)^2/a, q = c - (c - s) z, accept if (ndim-1) ln z + lp(q) - lp(s) > ln u').
// Differences by design (documented in DESIGN.md): the red/blue split is the parity of the global
// walker id (emcee: randomize_split) and the RNG is Philox4x32-10 keyed by (seed, step, walker id)
// instead of MT19937, so a chain is bit-reproducible for ANY sharding of walkers over GPUs.
#pragma once
#include "lte_common.cuh"

namespace lte {

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ inline double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

// number of ids in [0, g) whose parity equals `split`
__host__ __device__ inline int colour_count(int g, int split) { return split ? (g >> 1) : ((g + 1) >> 1); }

// out[0] = max dV, out[1] = max_c |vlsr_c - al - mc| over the rows that can reach the fused kernel
__global__ void dv_max_kernel(const double* __restrict__ theta, int nw, ModelDev md, double lo, double hi,
                              unsigned long long* __restrict__ out) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0, dc = 0.0;
  if (w < nw) {
    const double* th = theta + (size_t)w * md.ndim;
    double v = th[md.idx_dv];
    if (isfinite(v) && v > 0.0 && v > lo && v < hi) {
      d = v;
      for (int c = 0; c < md.K; ++c) {
        double x = fabs(th[md.idx_vlsr[c]] - md.al - md.mc);
        if (isfinite(x) && x > dc) dc = x;
      }
    }
  }
  // non-negative doubles order like their bit patterns
  unsigned long long b0 = (unsigned long long)__double_as_longlong(d), b1 = (unsigned long long)__double_as_longlong(dc);
  for (int o = 16; o; o >>= 1) {
    unsigned long long o0 = __shfl_xor_sync(0xffffffffu, b0, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o);
    b0 = o0 > b0 ? o0 : b0; b1 = o1 > b1 ? o1 : b1;
  }
  if ((threadIdx.x & 31) == 0) { if (b0) atomicMax(out, b0); if (b1) atomicMax(out + 1, b1); }
}

// The stretch-move proposal of global walker `gid` (colour `split`): q = c - (c - s) z with z = ((a-1)u+1)^2/a and the
// partner c drawn from the other colour; RNG keyed by (seed, step, gid).  Returns z.
__device__ __forceinline__ double stretch_proposal(const double* __restrict__ all_coords, int nw_global, int ndim, int gid,
                                                   int split, uint64_t seed, unsigned long long step, double a,
                                                   double* __restrict__ q) {
  uint32_t r[4];
  philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)gid, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double zr = (a - 1.0) * u01(r[0]) + 1.0;
  const double z = zr * zr / a;
  const int nc = nw_global >> 1;
  int j = (int)(u01(r[1]) * nc);
  if (j >= nc) j = nc - 1;
  const int partner = 2 * j + (1 - split);
  const double* s = all_coords + (size_t)gid * ndim;
  const double* c = all_coords + (size_t)partner * ndim;
  for (int p = 0; p < ndim; ++p) q[p] = c[p] - (c[p] - s[p]) * z;
  return z;
}

// Per-launch values of a half-step that is replayed as a CUDA graph (chalte.cu: sampler_half_step_impl): the graph's
// kernel arguments are frozen, so the step index and the need slot are read from this 16-byte device record, which
// the host refreshes in stream order before every replay.  dyn == nullptr: the by-value arguments are used.
struct SamplerDyn { unsigned long long step; int slot; int pad; };
// Reach class of a proposal: r = (|vlsr - al - mc|_max + kZcut sigma) / (half-width of the primary lists), in 16
// geometric steps of 2^(1/5): class 0: r <= 1/8, class c: 2^((c-1)/5)/8 < r <= 2^(c/5)/8, class 15: r > 0.87
constexpr int kReachClasses = 16;
__host__ __device__ inline double reach_class_upper(int c) { return c >= kReachClasses - 1 ? 1.0 : 0.125 * exp2(0.2 * c); }
__device__ __forceinline__ int reach_class(float r) {
  if (!(r > 0.125f)) return 0;
  const int c = (int)ceilf(log2f(8.0f * r) * 5.0f);
  return c > kReachClasses - 1 ? kReachClasses - 1 : (c < 0 ? 0 : c);
}

// What this half-step's proposals need from the pair list: max dV and max_c |vlsr_c - al - mc| over the proposals of
// ALL walkers of colour `split` of the GLOBAL ensemble (each rank recomputes every proposal: a few thousand threads),
// ignoring proposals outside the prior box (never evaluated).  Identical on every rank, so the list extent -- and
// with it every log-probability -- does not depend on how the walkers are sharded.
__global__ void proposal_need_kernel(const double* __restrict__ all_coords, int nw_global, ModelDev md, int split,
                                     uint64_t seed, unsigned long long step, double a,
                                     const double* __restrict__ lo, const double* __restrict__ hi,
                                     unsigned long long* __restrict__ out, const SamplerDyn* __restrict__ dyn,
                                     int w0, int nl, float inv_hv_ref, int* __restrict__ cls,
                                     int* __restrict__ hist /*[kReachClasses] or nullptr: classes of ALL proposals*/) {
  __shared__ int s_hist[kReachClasses];
  if (hist) { if (threadIdx.x < kReachClasses) s_hist[threadIdx.x] = 0; __syncthreads(); }
  if (dyn) { step = dyn->step; out += 2 * dyn->slot; }
  const int k = blockIdx.x * blockDim.x + threadIdx.x;          // k-th walker of this colour
  double d = 0.0, dc = 0.0;
  const int gid = 2 * k + split;
  if (gid < nw_global) {
    double q[kMaxNdim];
    stretch_proposal(all_coords, nw_global, md.ndim, gid, split, seed, step, a, q);
    bool inside = true;
    for (int p = 0; p < md.ndim; ++p) if (!(lo[p] < q[p] && q[p] < hi[p])) inside = false;
    const double v = q[md.idx_dv];
    if (inside && isfinite(v) && v > 0.0) {
      d = v;
      for (int c = 0; c < md.K; ++c) {
        const double x = fabs(q[md.idx_vlsr[c]] - md.al - md.mc);
        if (isfinite(x) && x > dc) dc = x;
      }
    }
    // reach class: how far from the mask centre the proposal's own kZcut-sigma range extends, relative to the
    // half-width of the (wide) list.  reach_sort_kernel orders the evaluation batch of the LOCAL proposals by it, so
    // that the walkers of a warp skip the same records and the bulk can be served by a narrower list set; the
    // histogram over ALL proposals of the ensemble (identical on every rank) is what the host sizes that set from.
    const int c = d > 0.0 ? reach_class((float)(dc + kZcut * d / kFwhm) * inv_hv_ref) : 0;
    if (cls && gid >= w0 && gid < w0 + nl) cls[colour_count(gid, split) - colour_count(w0, split)] = c;
    if (hist && d > 0.0) atomicAdd(&s_hist[c], 1);
  }
  if (hist) {
    __syncthreads();
    if (threadIdx.x < kReachClasses && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
  }
  unsigned long long b0 = (unsigned long long)__double_as_longlong(d), b1 = (unsigned long long)__double_as_longlong(dc);
  for (int o = 16; o; o >>= 1) {
    unsigned long long o0 = __shfl_xor_sync(0xffffffffu, b0, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o);
    b0 = o0 > b0 ? o0 : b0; b1 = o1 > b1 ? o1 : b1;
  }
  if ((threadIdx.x & 31) == 0) { if (b0) atomicMax(out, b0); if (b1) atomicMax(out + 1, b1); }
}

// Reach class of every row of a caller's batch (plain log-prob calls, see chalte.cu: eval_device), and how often the
// 32 consecutive rows of a warp differ by two classes or more: stat[0] += such warps, stat[1] += warps.  A warp of the
// fused kernel processes a record when ANY of its walkers reaches it, so a batch drawn from a spread-out ensemble in
// arbitrary order pays for its widest walker in every warp; evaluated in order of reach class it does not.
__global__ void row_reach_class_kernel(const double* __restrict__ theta, int nw, ModelDev md, float inv_hv_ref,
                                       int* __restrict__ cls, unsigned long long* __restrict__ stat) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  int c = -1;
  if (w < nw) {
    const double* th = theta + (size_t)w * md.ndim;
    const double v = th[md.idx_dv];
    c = 0;
    if (isfinite(v) && v > 0.0) {
      double dc = 0.0;
      for (int k = 0; k < md.K; ++k) {
        const double x = fabs(th[md.idx_vlsr[k]] - md.al - md.mc);
        if (isfinite(x) && x > dc) dc = x;
      }
      c = reach_class((float)(dc + kZcut * v / kFwhm) * inv_hv_ref);
    }
    cls[w] = c;
  }
  int lo = c < 0 ? kReachClasses : c, hi = c;
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && hi >= 0) {
    atomicAdd(stat + 1, 1ull);
    if (hi - lo >= 2) atomicAdd(stat, 1ull);
  }
}

// Stable counting sort of the n local proposals by reach class (one block): dest[k] = row of proposal k in the
// evaluation batch.  Log-probabilities do not depend on a walker's position in the batch, so the chain is unchanged.
// Two list sets (c_tight >= 0): proposals of class <= c_tight come first and are served by the narrow set; the others
// start at the next multiple of the walker-block size, so that no block mixes the two sides -- WHICH set a walker is
// evaluated against is a function of its own class, never of its neighbours or of the sharding.  The rows in between
// and after the last proposal (up to n_rows) are marked dead: NaN position (walker_prep_kernel drops it), idx -1.
// The class histogram of the whole ensemble (proposal_need_kernel) is published to the host mirror and re-zeroed.
constexpr int kSortThreads = 512;
__global__ void __launch_bounds__(kSortThreads)
reach_sort_kernel(int n, const int* __restrict__ cls, int* __restrict__ dest, int c_tight, int* __restrict__ split_out,
                  int n_rows, int ndim, double* __restrict__ prop, int* __restrict__ idx,
                  int* __restrict__ hist, int* __restrict__ hist_host,
                  int* __restrict__ inv /*nullptr or [n]: inv[dest[k]] = k*/,
                  unsigned long long* __restrict__ stat /*nullptr or [2]: published to stat_host and re-zeroed*/,
                  unsigned long long* __restrict__ stat_host) {
  __shared__ int s_cnt[kReachClasses][kSortThreads];
  __shared__ int s_tot[kReachClasses];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (hist && t < kReachClasses) { if (hist_host) hist_host[t] = hist[t]; hist[t] = 0; }
  if (stat && t < 2) { stat_host[t] = stat[t]; stat[t] = 0ull; }
  const int S = (n + kSortThreads - 1) / kSortThreads;
  const int k0 = min(t * S, n), k1 = min(k0 + S, n);
  for (int q = 0; q < kReachClasses; ++q) s_cnt[q][t] = 0;
  for (int k = k0; k < k1; ++k) s_cnt[cls[k]][t]++;
  __syncthreads();
  // exclusive scan of s_cnt[q][*] over the threads, one warp per class (kSortThreads / 32 == kReachClasses): every lane
  // owns 16 consecutive entries, a warp scan joins the lanes -- no block barrier per class
  static_assert(kSortThreads / 32 == kReachClasses && kSortThreads % 32 == 0, "one warp per reach class");
  {
    constexpr int kPer = kSortThreads / 32;
    int* row = s_cnt[wid];
    int loc[kPer];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) { loc[i] = sum; sum += row[lane * kPer + i]; }
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    const int excl = inc - sum;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kPer; ++i) row[lane * kPer + i] = excl + loc[i];
    if (lane == 31) s_tot[wid] = inc;
  }
  __syncthreads();
  int base[kReachClasses];
  int acc = 0, n_tight = 0, split_row = 0;
  for (int q = 0; q < kReachClasses; ++q) {
    if (c_tight >= 0 && q == c_tight + 1) { n_tight = acc; acc = (acc + 127) / 128 * 128; split_row = acc; }
    base[q] = acc; acc += s_tot[q];
  }
  if (c_tight >= kReachClasses - 1) { n_tight = acc; split_row = (acc + 127) / 128 * 128; }
  if (c_tight < 0) { n_tight = 0; split_row = 0; }
  const int n_end = acc;                                 // one past the last live row
  for (int k = k0; k < k1; ++k) {
    const int q = cls[k];
    int b = 0;
    for (int j = 0; j < kReachClasses; ++j) if (j == q) b = base[j];
    const int d = b + s_cnt[q][t]++;
    dest[k] = d;
    if (inv) inv[d] = k;
  }
  if (split_out) {
    if (t == 0) *split_out = split_row;
    // dead rows: the gap before the wide side and the tail of the launch
    for (int r = n_tight + t; r < min(split_row, n_rows); r += kSortThreads) { prop[(size_t)r * ndim] = nan(""); idx[r] = -1; }
    for (int r = max(n_end, split_row) + t; r < n_rows; r += kSortThreads) { prop[(size_t)r * ndim] = nan(""); idx[r] = -1; }
  }
}

// proposals for the local walkers of colour `split`, compacted in id order (or in the order reach_sort_kernel chose)
__global__ void stretch_propose_kernel(const double* __restrict__ all_coords, int nw_global, int w0, int nl, int ndim,
                                       int split, uint64_t seed, unsigned long long step, double a,
                                       double* __restrict__ prop, double* __restrict__ factor, int* __restrict__ idx,
                                       const SamplerDyn* __restrict__ dyn, const int* __restrict__ dest) {
  if (dyn) step = dyn->step;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nl) return;
  int gid = w0 + t;
  if ((gid & 1) != split) return;
  int k = colour_count(gid, split) - colour_count(w0, split);
  if (dest) k = dest[k];
  const double z = stretch_proposal(all_coords, nw_global, ndim, gid, split, seed, step, a, prop + (size_t)k * ndim);
  factor[k] = (ndim - 1.0) * log(z);
  idx[k] = t;
}

// What the pair list covered when the proposals' log-probs were launched optimistically, and where dv_max_kernel put
// the proposals' maxima.  The comparison is the host's (chalte.cu: hv_needed / drain) in the same arithmetic.
struct ListCover {
  const unsigned long long* need;   // [2]: max dV, max |vlsr_c - al - mc| (bit patterns); nullptr = list was checked up front
  double dv_cover, hv_cover, zc, fwhm;
  int mixed;
  unsigned long long* poison;       // sticky: set when a half-step was skipped; later half-steps skip too (until the
                                    // host re-runs them in order), so an invalid step never feeds a valid one
};

__device__ __forceinline__ bool list_covered(const ListCover& c) {
  if (!c.need) return true;
  const double dv = __longlong_as_double((long long)c.need[0]), dabs = __longlong_as_double((long long)c.need[1]);
  double hv = 10.0 * dv;
  if (c.mixed) hv = fmin(hv, dabs + c.zc * dv / c.fwhm);
  return dv <= c.dv_cover && hv <= c.hv_cover;
}

__global__ void stretch_accept_kernel(int n_move, int ndim, int w0, const int* __restrict__ idx,
                                      const double* __restrict__ prop, const double* __restrict__ new_lp,
                                      const double* __restrict__ factor, uint64_t seed, unsigned long long step,
                                      double* __restrict__ coords, double* __restrict__ logp,
                                      unsigned long long* __restrict__ n_acc, ListCover cov,
                                      const SamplerDyn* __restrict__ dyn, unsigned long long* __restrict__ need_host) {
  if (dyn) { step = dyn->step; cov.need += 2 * dyn->slot; need_host += 2 * dyn->slot; }
  // graph-replayed half-steps: the proposals' maxima (complete: proposal_need_kernel ran earlier in the stream) are
  // published to the host's pinned mirror here instead of by a copy node; drain() zeroes the slots
  if (dyn && blockIdx.x == 0 && threadIdx.x == 0) { need_host[0] = cov.need[0]; need_host[1] = cov.need[1]; }
  if (cov.need) {                           // log-probs not valid (or an earlier half-step was skipped): leave the state
    const bool skip = (cov.poison && *cov.poison != 0ull) || !list_covered(cov);
    if (skip) {
      if (cov.poison && blockIdx.x == 0 && threadIdx.x == 0) *cov.poison = 1ull;
      return;
    }
  }
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  bool acc = false;
  if (k < n_move && idx[k] >= 0) {                      // idx < 0: a dead row of the padded batch (reach_sort_kernel)
    int li = idx[k];
    int gid = w0 + li;
    uint32_t r[4];
    philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)gid, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    double diff = factor[k] + new_lp[k] - logp[li];
    acc = diff > log(u01(r[2]));
    if (acc) {
      for (int p = 0; p < ndim; ++p) coords[(size_t)li * ndim + p] = prop[(size_t)k * ndim + p];
      logp[li] = new_lp[k];
    }
  }
  unsigned m = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_acc, (unsigned long long)__popc(m));
}

// Chain store (the reference's chain.npy rows, inference.py:462/471, kept in HBM): after a full step the local
// positions and log-probs are appended to slot `slot` of the resident chain [slot][walker][ndim].  A step the lists did
// not cover (sticky flag set) stores nothing: the host re-runs it, store included, at the next synchronisation point.
__global__ void chain_store_kernel(int n_local, int ndim, const double* __restrict__ coords, const double* __restrict__ logp,
                                   double* __restrict__ chain_c, double* __restrict__ chain_l,
                                   const unsigned long long* __restrict__ poison) {
  if (poison && *poison != 0ull) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_local * ndim) chain_c[i] = coords[i];
  if (i < n_local) chain_l[i] = logp[i];
}

}  // namespace lte
