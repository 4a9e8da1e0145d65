// chalte.cu -- host side of libchalte.so: engine state, HBM residency, pair-list/tile builder,
// kernel launch sequence and the extern "C" boundary declared in include/chalte.h.
// There is deliberately NO CPU implementation of the path in this file: every evaluation
// launches the sm_100a kernels of lte_kernels.cuh; without a device cha_create fails.
#include "../../include/chalte.h"
#include "lte_kernels.cuh"
#include "lte_sampler.cuh"

#include <dlfcn.h>
#include <nccl.h>            // types and prototypes only: the library is dlopen'ed by cha_comm_* (no link-time dependency)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace lte;

namespace {

thread_local std::string g_create_error;
std::atomic<uint64_t> g_buf_epoch{0};      // bumped by every device reallocation: captured graphs hold raw pointers

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    g_buf_epoch.fetch_add(1, std::memory_order_relaxed);
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    // a buffer that is outgrown once tends to be outgrown again (lists follow the ensemble as it spreads): grow
    // geometrically so that reallocations -- cudaFree synchronises the device -- stay rare
    size_t want = std::max(bytes + bytes / 4 + 256, cap + cap / 2);
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return (T*)p; }
};

// A launch sequence of a small batch (walker_prep -> fused kernel -> finalize, plus its copies) replayed as one CUDA
// graph: at a few thousand walkers on a small grid the sequence is launch-latency bound (config 1: 128 walkers).
// A graph is valid for one (pointers, batch size, mode, need slot) and one configuration epoch.
struct GraphKey {
  int kind = -1;                 // 0 device-pointer log-prob, 1 host-buffer evaluation
  const void* a = nullptr; const void* b = nullptr;
  int64_t nw = 0; int mode = 0, slot = 0;
  uint64_t epoch = 0, buf_epoch = 0;
  bool operator==(const GraphKey& o) const {
    return kind == o.kind && a == o.a && b == o.b && nw == o.nw && mode == o.mode && slot == o.slot && epoch == o.epoch &&
           buf_epoch == o.buf_epoch;
  }
};
struct GraphEntry { GraphKey key; cudaGraphExec_t exec = nullptr; int launches = 0; uint64_t last_use = 0; };

// NCCL entry points, resolved on first use.  A process that already holds libnccl.so.2 (torch's bundled copy) keeps
// using that one; otherwise the system library is loaded.  A single-GPU user never needs NCCL on the machine.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    void* l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!l) l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!l) l = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!l) { err = std::string("NCCL not found: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(l, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(l, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(l, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(l, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(l, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString) {
      err = "libnccl.so.2 lacks one of ncclGetUniqueId/CommInitRank/CommDestroy/AllGather/GetErrorString";
      return false;
    }
    lib = l;
    return true;
  }
};
NcclApi g_nccl;

// Second, narrower group / record / tile list set of the resident sampler: the bulk of a half-step's proposals reach
// much less far from the mask centre than the widest one, which the primary lists must cover.
struct TightLists {
  bool valid = false;
  int cls = -1;                    // proposals of reach class <= cls are served by this set
  double hv = 0.0, chi_const = 0.0;
  int64_t n_tiles = 0, n_groups = 0, n_recs = 0, n_pairs = 0, n_unstaged = 0;
  DevBuf d_tiles, d_groups, d_recs;
};

struct HostMol {
  bool set = false;
  std::vector<double> nu, logint, elower;
  int q_kind = 0, n_qp = 0;
  double qp[8] = {0};
  std::vector<double> sg, sE;
  double ll = 0, ul = 0;
  bool all_lines = true;
  std::vector<int64_t> line_idx;
  DevBuf d_sg, d_sE;
};

}  // namespace

// Host image of one group / record / tile list set (the mixed kernel's layout) for the line windows of half-width hv
struct HostLists {
  std::vector<int> wa, wb;            // window of line i in channel index space: [wa, wb)
  std::vector<int> act_ch;            // active channels = union of the windows
  std::vector<double> ax, ay, aw, ais;
  std::vector<GroupBlk> gblk;
  std::vector<LineRec> recs;
  std::vector<TileG> tiles;
  std::vector<int> grp_c0, grp_c1;    // first / last sorted channel index of each group (span table of the channel stream)
  std::vector<int> grp_rec0, grp_rec1;  // its records [rec0, rec1) in the record array
  int64_t P = 0, n_unstaged = 0;
  double y2w_active = 0.0;
};

struct cha_engine {
  int dev = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t n_launch = 0, n_rebuild = 0, n_graph_launch = 0;
  uint64_t epoch = 0;                   // bumped whenever anything a captured graph baked in changes
  std::vector<GraphEntry> graphs; std::vector<GraphKey> seen_keys; uint64_t graph_clock = 0; bool capturing = false;
  float last_fused_ms = 0.f; bool ev_valid = false;   // ev0/ev1 bracket the last fused launch once one was made
  int prec = CHA_PREC_MIXED;

  // configuration (host copies)
  HostMol mol[kMaxM];
  bool model_set = false, prior_set = false, spec_set = false;
  ModelDev md{};
  std::vector<double> pr_lo, pr_hi, pr_mu, pr_sg;
  std::vector<int> pr_gauss;
  double vmin_sep = NAN, vmax_sep = NAN;
  std::vector<double> sx, sy, syerr;    // spectrum as given

  // derived host state
  bool lines_dirty = true, spec_dirty = true, pairs_dirty = true;
  std::vector<double> l_nu, l_logint, l_el; std::vector<int> l_mol;   // selected lines, frequency-sorted
  std::vector<double> xs, ys, ws, iss; std::vector<int> perm;         // channels sorted by frequency (iss = 1/yerr)
  double sum_neg_log_w = 0.0; std::vector<double> y2w_prefix;         // walker-independent chi-square pieces
  double build_ms_total = 0.0;                                        // host time spent (re)building the lists
  int64_t calls_since_rebuild = 0; double grow_margin = 1.02;         // rebuild policy state (ensure_pairs)
  int64_t slack_calls = 0;                                            // optimistic calls served by a too-wide list
  double dv_list = 0.0;     // largest dV the current pair list serves
  double hv_list = 0.0;     // half-width (km/s about the mask centre) of the line windows in the list
  int64_t n_act = 0, n_pairs = 0, n_tiles = 0;      // n_tiles: per-pair tiling (fp64 kernel)
  int64_t n_tiles_g = 0, n_groups = 0, n_recs = 0;   // group tiling (mixed kernel)
  int64_t n_tiles_unstaged = 0;                      // tiles too dense for shared-memory staging (general paths)
  double chi_const_fp64 = 0.0, chi_const_mixed = 0.0;

  // device residency
  DevBuf d_lnu, d_llogint, d_lel, d_lK, d_lK2, d_lmol, d_qdesc, d_prior, d_prior_i;
  DevBuf d_tiles, d_poff, d_pline, d_pu64, d_pu32, d_x, d_y, d_w, d_jbg, d_beam2, d_tn;
  DevBuf d_xall, d_actof, d_outpos;
  DevBuf d_tiles_g, d_groups, d_recs;
  DevBuf d_span_tiles, d_span_segs; int64_t n_spans = 0;   // per span of kSpanCh channels: offsets into its segments (SpanSeg)
  bool perm_identity = false;                    // the spectrum was given in ascending channel order
  int span_stream = 1;                           // CHALTE_SPAN_STREAM: 0 never (zero-fill + simulate_tiles_kernel), 1 on sparse
                                                 // grids (default), 2 wherever the span table exists (tests, A/B)
  bool span_sparse = false;
  // reach-ordered evaluation of plain log-prob batches (eval_device): -1 adaptive, 0 off, 1 always (CHALTE_SORT_ROWS)
  int sort_rows = -1; bool sort_on = false; int64_t sort_calls = 0, n_sorted = 0;
  DevBuf d_rcls, d_rdest, d_rinv, d_sortstat; unsigned long long* h_sortstat = nullptr;
  // workspace
  DevBuf d_theta, d_out, d_ok, d_lp, d_qinv, d_qpart, d_tau, d_gco, d_partial, d_scratch, d_sim, d_wpf, d_wpd;
  double* h_pin = nullptr; size_t h_pin_cap = 0;
  int n_qchunks_max = 1;

  // optimistic device-pointer calls: evaluated against the current pair list without waiting for the batch's
  // dV / |vlsr| maxima; the maxima land in pinned memory and are checked at the next synchronisation point,
  // where anything the list did not cover is re-run after a rebuild (drain())
  struct Pend {
    int kind;                      // 0 log_prob_dev, 1 sampler half-step
    const double* d_theta; int64_t nw; double* d_out; int with_prior;
    int64_t step; int split; const double* d_all;
    int64_t store_slot;            // chain slot this half-step appends to (-1: none)
    double dv_cover, hv_cover;     // what the list covered at launch
  };
  std::vector<Pend> pend;
  DevBuf d_need;                   // (kMaxPend + 1) x 2 u64: max dV, max |vlsr_c - al - mc| per pending call (the last
                                   // slot belongs to the synchronous path), then the sticky skip flag
  unsigned long long* h_need = nullptr;
  DevBuf d_dyn; SamplerDyn* h_dyn = nullptr;     // per-launch record of a graph-replayed half-step + its pinned ring
  bool in_redo = false;

  // sampler state (lte_sampler.cuh)
  int64_t s_nw_global = 0, s_w0 = 0, s_nw_local = 0, s_accepted = 0;
  bool s_logp_valid = false;     // local log-probs computed with the ensemble-sized list (first half-step)
  uint64_t s_seed = 0; double s_a = 2.0;
  DevBuf s_coords, s_logp, s_prop, s_newlp, s_factor, s_acc, s_idx, s_cls, s_dest;
  // walkers sharded over ranks: communicator, the gathered ensemble, counters
  ncclComm_t comm = nullptr; int comm_rank = 0, comm_world = 1;
  DevBuf s_all;                  // [nw_global][ndim], refreshed by the all-gather at the head of every half-step
  int64_t n_coll = 0, coll_bytes = 0, n_redo = 0;
  // chain resident in HBM: [slot][nw_local][ndim] and [slot][nw_local]
  DevBuf s_chain_c, s_chain_l; int64_t s_chain_cap = 0, s_chain_n = 0;
  // bulk / outlier split of the sampler's evaluation batches (lte_sampler.cuh: reach classes)
  TightLists tight; int64_t n_rebuild_tight = 0;
  HostLists host_wide, host_tight;   // host images, kept between rebuilds (tens of MB at 1e6 channels: no re-faulting)
  bool two_lists = true;           // CHALTE_TWO_LISTS=0 turns the split off (A/B measurements)
  bool debug = false;              // CHALTE_DEBUG=1: list builds and re-runs are reported on stderr
  bool sampler_graphs = true;      // CHALTE_SAMPLER_GRAPHS=0: half-steps as plain launches (A/B measurements)
  double drain_ms_total = 0.0; int64_t n_drain = 0, n_events = 0;
  double dv_hi = 0.0, dabs_hi = 0.0;   // slowly decaying maxima of what the queued calls needed (sizing of the primary lists)
  bool box_lists = false;              // the primary lists cover the whole prior box (no proposal can outgrow them)
  DevBuf d_hist, d_split; int* h_hist = nullptr;   // class histogram of ALL proposals of the last half-step (+ pinned mirror)
  int tight_want = -1, tight_want_streak = 0;
};

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      h->err = std::string(#call) + ": " + cudaGetErrorString(_e);                        \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

#define FAIL(msg)       \
  do {                  \
    h->err = (msg);     \
    return 1;           \
  } while (0)

static int upload(cha_handle h, DevBuf& b, const void* src, size_t bytes) {
  CK(b.ensure(bytes ? bytes : 8));
  if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

static int ensure_pin(cha_handle h, size_t bytes) {
  if (bytes <= h->h_pin_cap) return 0;
  if (h->h_pin) cudaFreeHost(h->h_pin);
  h->h_pin = nullptr; h->h_pin_cap = 0;
  CK(cudaMallocHost((void**)&h->h_pin, bytes + bytes / 4));
  h->h_pin_cap = bytes + bytes / 4;
  h->epoch++;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// lines: trim (functions.py:507-540) + selection (inference.py:142-144), merge molecules by frequency,
// device precompute of K_i (classes.py:90-98)
// ---------------------------------------------------------------------------------------------
static int prepare_lines(cha_handle h) {
  const int M = h->md.M;
  struct L { double nu, logint, el; int mol; };
  std::vector<L> all;
  std::vector<double> q_ct(M, 0.0);
  int nqc_max = 1;
  std::vector<QDesc> qd(M);
  for (int m = 0; m < M; ++m) {
    HostMol& hm = h->mol[m];
    if (!hm.set) FAIL("molecule " + std::to_string(m) + " not set");
    const int64_t N = (int64_t)hm.nu.size();
    // trim_array: first index with nu > ll ... first index with nu > ul
    int64_t i0 = N, i1 = N;
    for (int64_t i = 0; i < N; ++i) if (hm.nu[i] > hm.ll) { i0 = i; break; }
    if (i0 == N) { if (N && hm.nu[N - 1] < hm.ll) { i0 = 0; i1 = 0; } else i0 = 0; }   // functions.py:522-526
    if (!(i0 == 0 && i1 == 0)) { i1 = N; for (int64_t i = 0; i < N; ++i) if (hm.nu[i] > hm.ul) { i1 = i; break; } }
    if (i1 < i0) i1 = i0;
    const int64_t ntrim = i1 - i0;
    auto push = [&](int64_t k) { all.push_back({hm.nu[k], hm.logint[k], hm.elower[k], m}); };
    if (hm.all_lines) { for (int64_t k = i0; k < i1; ++k) push(k); }
    else {
      for (int64_t v : hm.line_idx) {
        int64_t k = v < 0 ? v + ntrim : v;                      // numpy negative indexing
        if (k < 0 || k >= ntrim) FAIL("line index out of range for the trimmed catalog (inference.py:142)");
        push(i0 + k);
      }
    }
    qd[m].kind = hm.q_kind; qd[m].n_params = hm.n_qp;
    for (int k = 0; k < 8; ++k) qd[m].p[k] = hm.qp[k];
    qd[m].n_states = (int)hm.sg.size();
    if (hm.q_kind == CHA_Q_SUM) {
      if (upload(h, hm.d_sg, hm.sg.data(), hm.sg.size() * 8)) return 1;
      if (upload(h, hm.d_sE, hm.sE.data(), hm.sE.size() * 8)) return 1;
      nqc_max = std::max(nqc_max, (qd[m].n_states + kQChunk - 1) / kQChunk);
    }
    qd[m].g = hm.d_sg.as<double>(); qd[m].E = hm.d_sE.as<double>();
  }
  h->n_qchunks_max = nqc_max;
  std::stable_sort(all.begin(), all.end(), [](const L& a, const L& b) { return a.nu < b.nu; });
  const size_t Ls = all.size();
  h->l_nu.resize(Ls); h->l_logint.resize(Ls); h->l_el.resize(Ls); h->l_mol.resize(Ls);
  for (size_t i = 0; i < Ls; ++i) { h->l_nu[i] = all[i].nu; h->l_logint[i] = all[i].logint; h->l_el[i] = all[i].el; h->l_mol[i] = all[i].mol; }
  if (upload(h, h->d_lnu, h->l_nu.data(), Ls * 8) || upload(h, h->d_llogint, h->l_logint.data(), Ls * 8) ||
      upload(h, h->d_lel, h->l_el.data(), Ls * 8) || upload(h, h->d_lmol, h->l_mol.data(), Ls * 4) ||
      upload(h, h->d_qdesc, qd.data(), sizeof(QDesc) * M))
    return 1;
  CK(h->d_lK.ensure(Ls * 8 + 8));
  // Q(CT=300) per molecule on the device (classes.py:94), then K_i per line
  CK(h->d_scratch.ensure(64 * 8 + (size_t)nqc_max * 128 * 8));
  for (int m = 0; m < M; ++m) {
    if (qd[m].kind == CHA_Q_SUM) {
      double T300 = kCT;
      double* d_t = h->d_scratch.as<double>();
      double* d_qp = d_t + 64;
      CK(cudaMemcpyAsync(d_t, &T300, 8, cudaMemcpyHostToDevice, h->stream));
      int nch = (qd[m].n_states + kQChunk - 1) / kQChunk;
      q_state_sum_kernel<<<dim3(1, nch), 256, 0, h->stream>>>(d_t, 1, 1, 0, qd[m].g, qd[m].E, qd[m].n_states, d_qp, 128);
      h->n_launch++;
      std::vector<double> part((size_t)nch * 128);
      CK(cudaMemcpyAsync(part.data(), d_qp, part.size() * 8, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      double q = 0.0;
      for (int c = 0; c < nch; ++c) q += part[(size_t)c * 128];
      q_ct[m] = q;
    } else {
      q_ct[m] = q_analytic(qd[m], kCT);
    }
  }
  // lines are interleaved by molecule after the sort: launch per molecule over the whole list with a
  // per-line q_ct would need a gather; instead run once per molecule on a compacted copy
  for (int m = 0; m < M; ++m) {
    std::vector<double> nu, li, el; std::vector<size_t> pos;
    for (size_t i = 0; i < Ls; ++i) if (h->l_mol[i] == m) { nu.push_back(h->l_nu[i]); li.push_back(h->l_logint[i]); el.push_back(h->l_el[i]); pos.push_back(i); }
    const size_t n = nu.size();
    if (!n) continue;
    CK(h->d_scratch.ensure(4 * n * 8));
    double* d = h->d_scratch.as<double>();
    CK(cudaMemcpyAsync(d, nu.data(), n * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + n, li.data(), n * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + 2 * n, el.data(), n * 8, cudaMemcpyHostToDevice, h->stream));
    catalog_terms_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>((int)n, d, d + n, d + 2 * n, q_ct[m], d + 3 * n);
    h->n_launch++;
    std::vector<double> Kf(n);
    CK(cudaMemcpyAsync(Kf.data(), d + 3 * n, n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (M == 1) {
      CK(cudaMemcpyAsync(h->d_lK.p, d + 3 * n, n * 8, cudaMemcpyDeviceToDevice, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    } else {
      // scatter back to merged order (tiny, once per molecule)
      std::vector<double> tmp(Ls);
      CK(cudaMemcpy(tmp.data(), h->d_lK.p, Ls * 8, cudaMemcpyDeviceToHost));
      for (size_t k = 0; k < n; ++k) tmp[pos[k]] = Kf[k];
      CK(cudaMemcpy(h->d_lK.p, tmp.data(), Ls * 8, cudaMemcpyHostToDevice));
    }
  }
  CK(cudaGetLastError());
  {
    // log2 of the line factors, for strengths formed in log2 space (once per catalog)
    std::vector<double> kf(Ls + 1, 0.0);
    if (Ls) CK(cudaMemcpy(kf.data(), h->d_lK.p, Ls * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < Ls; ++i) kf[i] = kf[i] > 0.0 ? std::log2(kf[i]) : -1.0e4;
    if (upload(h, h->d_lK2, kf.data(), (Ls + 1) * 8)) return 1;
    CK(cudaStreamSynchronize(h->stream));
  }
  h->lines_dirty = false;
  h->pairs_dirty = true;
  return 0;
}

static int prepare_spectrum(cha_handle h) {
  const size_t C = h->sx.size();
  h->perm.resize(C);
  for (size_t j = 0; j < C; ++j) h->perm[j] = (int)j;
  std::stable_sort(h->perm.begin(), h->perm.end(), [&](int a, int b) { return h->sx[a] < h->sx[b]; });
  h->xs.resize(C); h->ys.resize(C); h->ws.resize(C); h->iss.resize(C);
  for (size_t j = 0; j < C; ++j) {
    int o = h->perm[j];
    h->xs[j] = h->sx[o]; h->ys[j] = h->sy[o];
    h->ws[j] = 1.0 / (h->syerr[o] * h->syerr[o]);                                 // inference.py:157
    h->iss[j] = 1.0 / h->syerr[o];
  }
  h->perm_identity = true;
  for (size_t j = 0; j < C; ++j) if (h->perm[j] != (int)j) { h->perm_identity = false; break; }
  if (upload(h, h->d_xall, h->xs.data(), C * 8) || upload(h, h->d_outpos, h->perm.data(), C * 4)) return 1;
  // walker-independent pieces of the chi-square, once per spectrum: sum_j -ln(w_j) (inference.py:160) and the prefix
  // sums of y^2 w, from which any rebuild gets the chi-square of its inactive channels (model == 0 there) in O(runs)
  h->sum_neg_log_w = 0.0;
  h->y2w_prefix.assign(C + 1, 0.0);
  for (size_t j = 0; j < C; ++j) {
    h->sum_neg_log_w -= std::log(h->ws[j]);
    h->y2w_prefix[j + 1] = h->y2w_prefix[j] + h->ys[j] * h->ys[j] * h->ws[j];
  }
  h->spec_dirty = false;
  h->pairs_dirty = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// pair list (line x channel windows at dV = dv), active channels, tiles, walker-independent constants
// ---------------------------------------------------------------------------------------------
static constexpr int kTileMaxChan = 512;
static constexpr int kTileMaxPairs = 8192;
static constexpr double kTileMaxRelHalfSpan = 0.004;   // cubic interpolation error of G(x) < 2e-11

static int make_group_lists(cha_handle h, double hv, HostLists& L) {
  const int M = h->md.M;
  const double mc = h->md.mc;
  const size_t C = h->xs.size(), Ls = h->l_nu.size();
  const double* x = h->xs.data();
  // window of line i in channel index space; nu sorted -> brackets monotone
  std::vector<int>& wa = L.wa; std::vector<int>& wb = L.wb;
  wa.resize(Ls); wb.resize(Ls);
  const double flo = 1.0 - (mc + hv) / kCkm, fhi = 1.0 - (mc - hv) / kCkm;
  int64_t P = 0;
  // lines are frequency-sorted and both window edges scale with the line frequency, so the brackets only move forward:
  // each search gallops from the previous line's bracket (a few cache-local steps instead of 20 misses over 8 MB)
  auto gallop = [&](size_t from, double v, bool upper) -> size_t {
    size_t lo = from, step = 1, hi = from;
    while (hi < C && (upper ? x[hi] <= v : x[hi] < v)) { lo = hi + 1; hi += step; step <<= 1; }
    hi = std::min(hi, C);
    return upper ? (size_t)(std::upper_bound(x + lo, x + hi, v) - x) : (size_t)(std::lower_bound(x + lo, x + hi, v) - x);
  };
  const bool forward = flo > 0.0 && fhi > 0.0;
  size_t prev_a = 0, prev_b = 0;
  for (size_t i = 0; i < Ls; ++i) {
    double xlo = h->l_nu[i] * flo, xhi = h->l_nu[i] * fhi;
    xlo -= std::fabs(xlo) * 1e-12; xhi += std::fabs(xhi) * 1e-12;
    if (forward && i > 0 && h->l_nu[i] >= h->l_nu[i - 1]) {
      prev_a = gallop(prev_a, xlo, false); prev_b = gallop(prev_b, xhi, true);
    } else {
      prev_a = (size_t)(std::lower_bound(x, x + C, xlo) - x); prev_b = (size_t)(std::upper_bound(x, x + C, xhi) - x);
    }
    wa[i] = (int)prev_a; wb[i] = (int)prev_b;
    if (wb[i] < wa[i]) wb[i] = wa[i];
    P += wb[i] - wa[i];
  }
  if (P > (int64_t)0x7fffff00) FAIL("pair list exceeds 2^31 entries; narrow the dV bound or split the spectrum");
  L.P = P;
  // active channels = union of the windows (both ends are non-decreasing in the line index): O(L + A), and the
  // chi-square of the inactive ones from the prefix sums
  std::vector<int>& act_ch = L.act_ch;
  act_ch.clear();
  double y2w_active = 0.0;
  {
    int run_lo = -1, run_hi = -1;
    auto flush = [&]() {
      if (run_hi > run_lo) {
        for (int j = run_lo; j < run_hi; ++j) act_ch.push_back(j);
        y2w_active += h->y2w_prefix[run_hi] - h->y2w_prefix[run_lo];
      }
    };
    for (size_t i = 0; i < Ls; ++i) {
      if (wb[i] <= wa[i]) continue;
      if (run_hi < 0 || wa[i] > run_hi) { flush(); run_lo = wa[i]; run_hi = wb[i]; }
      else run_hi = std::max(run_hi, wb[i]);
    }
    flush();
  }
  L.y2w_active = y2w_active;
  const size_t A = act_ch.size();
  std::vector<double>& ax = L.ax; std::vector<double>& ay = L.ay; std::vector<double>& aw = L.aw; std::vector<double>& ais = L.ais;
  ax.resize(A); ay.resize(A); aw.resize(A); ais.resize(A);
  for (size_t a = 0; a < A; ++a) { const int j = act_ch[a]; ax[a] = x[j]; ay[a] = h->ys[j]; aw[a] = h->ws[j]; ais[a] = h->iss[j]; }
  // ---- group / record / tile layout of the mixed kernel (lte_kernels.cuh) ----
  std::vector<GroupBlk>& gblk = L.gblk;
  std::vector<LineRec>& recs = L.recs;
  std::vector<TileG>& tiles_g = L.tiles;
  gblk.clear(); recs.clear(); tiles_g.clear(); L.grp_c0.clear(); L.grp_c1.clear(); L.grp_rec0.clear(); L.grp_rec1.clear();
  int64_t n_unstaged = 0;
  {
    struct GInfo { size_t a0, a1; size_t rec0, rec1; int lmin, lmax; };
    std::vector<GInfo> ginfo;
    ginfo.reserve(A / 4 + 16);
    size_t g0a = 0;
    size_t ilo = 0, ihi = 0;                  // both only move forward with the group (windows are monotone in line index)
    gblk.reserve(A / 6 + 16); recs.reserve((size_t)P / 6 + 16); L.grp_c0.reserve(A / 6 + 16); L.grp_c1.reserve(A / 6 + 16);
    L.grp_rec0.reserve(A / 6 + 16); L.grp_rec1.reserve(A / 6 + 16);
    while (g0a < A) {
      size_t g1a = g0a + 1;
      while (g1a < A && (g1a - g0a) < (size_t)kGroupCh && (ax[g1a] - ax[g0a]) / ax[g0a] * kCkm <= 1.0) ++g1a;
      // lines whose dv-window intersects the group's channel range [jf, jl]
      const int jf = act_ch[g0a], jl = act_ch[g1a - 1];
      while (ilo < Ls && wb[ilo] <= jf) ++ilo;                                                 // first wb > jf
      if (ihi < ilo) ihi = ilo;
      while (ihi < Ls && wa[ihi] <= jl) ++ihi;                                                 // first wa > jl
      GroupBlk gb;
      std::memset(&gb, 0, sizeof(gb));
      for (int jj = 0; jj < kGroupCh; ++jj) gb.opos[jj] = -1;
      for (size_t a = g0a; a < g1a; ++a) {
        gb.dx[a - g0a] = (float)(ax[a] - ax[g0a]);
        // sigma-scaled data split hi + lo; the model is scaled by the fp32-rounded 1/sigma (an error on m only)
        const double ysc = ay[a] * ais[a];
        const float yh = (float)ysc;
        gb.ysh[a - g0a] = yh; gb.ysl[a - g0a] = (float)(ysc - (double)yh); gb.ns[a - g0a] = -(float)ais[a];
        gb.y2w += ay[a] * ay[a] * aw[a];
        gb.opos[a - g0a] = h->perm[act_ch[a]];
      }
      // padding lanes repeat the last offset (weights 0, opos -1): dx[kGroupCh - 1] is always the group's extent
      for (size_t k = g1a - g0a; k < (size_t)kGroupCh; ++k) gb.dx[k] = gb.dx[g1a - g0a - 1];
      const size_t rec0 = recs.size();
      for (int m = 0; m < M; ++m) {
        size_t cntm = 0;
        for (size_t i = ilo; i < ihi; ++i) {
          if (h->l_mol[i] != m) continue;
          const double f = h->l_nu[i];
          LineRec rc;
          rc.u0 = (float)((f - ax[g0a]) / f * kCkm - mc);
          rc.slope = (float)(kCkm / f);
          rc.line = (int)i; rc.lloc = 0;
          recs.push_back(rc); ++cntm;
        }
        if (cntm > 65535) FAIL("more than 65535 lines overlap one channel group");
        gb.nrec[m] = (unsigned short)cntm;
      }
      gblk.push_back(gb);
      int lmin = std::numeric_limits<int>::max(), lmax = -1;
      for (size_t q = rec0; q < recs.size(); ++q) { lmin = std::min(lmin, recs[q].line); lmax = std::max(lmax, recs[q].line); }
      ginfo.push_back({g0a, g1a, rec0, recs.size(), lmin, lmax});
      L.grp_c0.push_back(jf); L.grp_c1.push_back(jl); L.grp_rec0.push_back((int)rec0); L.grp_rec1.push_back((int)recs.size());
      g0a = g1a;
    }
    size_t gi = 0;
    while (gi < ginfo.size()) {
      size_t gj = gi + 1;
      const double x0 = ax[ginfo[gi].a0];
      const double span_max = 2.0 * kTileMaxRelHalfSpan * x0;
      int lmin = ginfo[gi].lmin, lmax = ginfo[gi].lmax;
      while (gj < ginfo.size() && (gj - gi) < (size_t)kTileMaxGroups &&
             (ax[ginfo[gj].a1 - 1] - x0) <= span_max && (ginfo[gj].rec1 - ginfo[gi].rec0) <= (size_t)kTileMaxRecs &&
             std::max(lmax, ginfo[gj].lmax) - std::min(lmin, ginfo[gj].lmin) + 1 <= kTileMaxLines) {
        lmin = std::min(lmin, ginfo[gj].lmin); lmax = std::max(lmax, ginfo[gj].lmax);
        ++gj;
      }
      TileG t;
      t.g0 = (int)gi; t.ng = (int)(gj - gi);
      t.rec_begin = (int)ginfo[gi].rec0; t.rec_count = (int)(ginfo[gj - 1].rec1 - ginfo[gi].rec0);
      // the single-molecule fast path relies on >= 1 record per group (true by construction: a group is made of
      // channels some line window touches); if it ever were not, send the tile down the general path
      if (M == 1) for (size_t g = gi; g < gj; ++g) if (ginfo[g].rec1 == ginfo[g].rec0) t.rec_count = kTileMaxRecs + 1;
      t.line0 = lmax >= lmin ? lmin : 0; t.nline = lmax >= lmin ? lmax - lmin + 1 : 0; t.pad0 = t.pad1 = 0;
      for (size_t q = ginfo[gi].rec0; q < ginfo[gj - 1].rec1; ++q) recs[q].lloc = (recs[q].line - t.line0) * kWalkersPerBlock;

      const double xl = ax[ginfo[gi].a0], xr = ax[ginfo[gj - 1].a1 - 1];
      t.xc = 0.5 * (xl + xr); t.hs = std::max(0.5 * (xr - xl), 1e-6);
      static const double nodes[4] = {0.9238795325112867, 0.38268343236508984, -0.3826834323650897, -0.9238795325112867};
      for (int n = 0; n < 4; ++n) {
        const double xn = t.xc + t.hs * nodes[n];
        t.jbg[n] = planck_j(xn, kTbg, h->md.eps);
        const double b = beam_size(xn, h->md.dish); t.beam2[n] = b * b;
      }
      {
        const double xe[2] = {t.xc + t.hs, t.xc - t.hs};
        t.jbg_hi = planck_j(xe[0], kTbg, h->md.eps); t.jbg_lo = planck_j(xe[1], kTbg, h->md.eps);
        const double bh = beam_size(xe[0], h->md.dish), bl = beam_size(xe[1], h->md.dish);
        t.beam2_hi = bh * bh; t.beam2_lo = bl * bl;
        t.line_span = 0.0; t.inv_hs = 1.0 / t.hs;
        for (int li = 0; li < t.nline; ++li) t.line_span = std::max(t.line_span, std::fabs(h->l_nu[t.line0 + li] - t.xc));
      }
      for (size_t g = gi; g < gj; ++g) {
        gblk[g].rec_off = (int)(ginfo[g].rec0 - ginfo[gi].rec0);
        gblk[g].tn0 = (float)((ax[ginfo[g].a0] - t.xc) / t.hs);
      }
      if (t.rec_count > kTileMaxRecs || t.nline > kTileMaxLines) n_unstaged++;
      tiles_g.push_back(t);
      gi = gj;
    }
  }
  L.n_unstaged = n_unstaged;
  { LineRec dummy; dummy.u0 = 0.f; dummy.slope = 0.f; dummy.line = 0; dummy.lloc = 0; recs.push_back(dummy); }   // look-ahead slot
  return 0;
}

// Span table of the one-pass channel-stream kernel: for every span of kSpanCh consecutive channels the SEGMENTS it
// holds -- the part of one tile that lies in the span: its groups, their records (contiguous) and the lines those
// records reference, cut into pieces that fit the kernel's staging area.  Groups, tiles and spans all ascend in channel
// index.  soff[s] .. soff[s + 1] are the segments of span s (segs carries one trailing dummy so that it is never
// empty).  Returns false when a single group holds more records than the staging area; `sparse` tells whether the
// one-pass kernel pays off: few active channels per span and (nearly) one segment per span -- a dense forest of lines is
// compute bound and stays with zero-fill + tiles (measured: DESIGN.md, channel stream).
static bool make_span_table(const HostLists& L, size_t C, std::vector<int>& soff, std::vector<SpanSeg>& sorted, bool& sparse) {
  const size_t ns = (C + kSpanCh - 1) / kSpanCh, nt = L.tiles.size(), A = L.act_ch.size();
  std::vector<std::pair<int, SpanSeg>> segs;
  bool span_fits = true;
  for (size_t t = 0; t < nt; ++t) {
    const TileG& T = L.tiles[t];
    const int g_end = T.g0 + T.ng;
    int glo = T.g0;
    const int s_first = L.grp_c0[T.g0] / kSpanCh, s_last = L.grp_c1[g_end - 1] / kSpanCh;
    for (int sp = s_first; sp <= s_last; ++sp) {
      const int64_t c_lo = (int64_t)sp * kSpanCh, c_hi = c_lo + kSpanCh;
      while (glo < g_end && L.grp_c1[glo] < c_lo) ++glo;
      int ghi = glo;
      while (ghi < g_end && L.grp_c0[ghi] < c_hi) ++ghi;
      for (int ga = glo; ga < ghi;) {
        int gb = ga + 1;
        while (gb < ghi && gb - ga < kSegGroups && L.grp_rec1[gb] - L.grp_rec0[ga] <= kSegRecs) ++gb;
        SpanSeg sg;
        sg.tile = (int)t; sg.g_lo = ga; sg.g_n = gb - ga;
        sg.r_lo = L.grp_rec0[ga]; sg.r_n = L.grp_rec1[gb - 1] - sg.r_lo;
        if (sg.r_n > kSegRecs) span_fits = false;      // one group with more records than the staging area holds
        int lmin = std::numeric_limits<int>::max(), lmax = -1;
        for (int q = sg.r_lo; q < sg.r_lo + sg.r_n; ++q) { lmin = std::min(lmin, L.recs[q].line); lmax = std::max(lmax, L.recs[q].line); }
        sg.l_lo = lmax >= lmin ? lmin : 0; sg.l_n = lmax >= lmin ? lmax - lmin + 1 : 0;
        sg.rec_shift = T.rec_begin - sg.r_lo;          // group.rec_off (relative to the tile) -> index into the staged records
        sg.line_shift = T.line0 - sg.l_lo;             // record.lloc / kWalkersPerBlock (relative to the tile) -> staged strength row
        sg.inv_hs = (float)(1.0 / T.hs);
        sg.pad[0] = sg.pad[1] = 0;
        segs.push_back({sp, sg});
        ga = gb;
      }
    }
  }
  soff.assign(ns + 1, 0);
  for (const auto& e : segs) soff[e.first + 1]++;
  for (size_t i = 0; i < ns; ++i) soff[i + 1] += soff[i];
  sorted.assign(segs.size() + 1, SpanSeg{});
  { std::vector<int> cur(soff.begin(), soff.end() - 1); for (const auto& e : segs) sorted[cur[e.first]++] = e.second; }
  size_t nonempty = 0;
  for (size_t i = 0; i < ns; ++i) nonempty += soff[i + 1] > soff[i];
  sparse = 8 * A <= C && 4 * segs.size() <= 5 * std::max<size_t>(nonempty, 1);
  return span_fits;
}

// The narrow list set of the resident sampler (bulk of the proposals; the primary set serves the outliers)
static int build_tight(cha_handle h, double hv) {
  const auto t0 = std::chrono::steady_clock::now();
  HostLists& L = h->host_tight;
  if (make_group_lists(h, hv, L)) return 1;
  TightLists& T = h->tight;
  if (upload(h, T.d_tiles, L.tiles.data(), L.tiles.size() * sizeof(TileG)) ||
      upload(h, T.d_groups, L.gblk.data(), L.gblk.size() * sizeof(GroupBlk)) ||
      upload(h, T.d_recs, L.recs.data(), L.recs.size() * sizeof(LineRec)))
    return 1;
  CK(cudaStreamSynchronize(h->stream));        // host vectors go out of scope
  T.hv = hv; T.n_tiles = (int64_t)L.tiles.size(); T.n_groups = (int64_t)L.gblk.size();
  T.n_recs = (int64_t)L.recs.size() - 1; T.n_pairs = L.P; T.n_unstaged = L.n_unstaged;
  T.chi_const = h->sum_neg_log_w + (h->y2w_prefix[h->xs.size()] - L.y2w_active);
  T.valid = true;
  h->n_rebuild_tight++;
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  h->build_ms_total += ms;
  if (h->debug) fprintf(stderr, "[chalte] narrow lists: class %d hv %.4f pairs %lld tiles %lld in %.2f ms\n", T.cls, hv,
                        (long long)T.n_pairs, (long long)T.n_tiles, ms);
  return 0;
}

static int build_pairs(cha_handle h, double hv, double dv) {
  const auto t_build0 = std::chrono::steady_clock::now();
  const int M = h->md.M;
  const double mc = h->md.mc;
  const size_t C = h->xs.size(), Ls = h->l_nu.size();
  const double* x = h->xs.data();
  HostLists& L = h->host_wide;
  if (make_group_lists(h, hv, L)) return 1;
  const std::vector<int>& wa = L.wa; const std::vector<int>& wb = L.wb; const std::vector<int>& act_ch = L.act_ch;
  const std::vector<double>& ax = L.ax; const std::vector<double>& ay = L.ay; const std::vector<double>& aw = L.aw;
  const int64_t P = L.P;
  const size_t A = act_ch.size();
  const int64_t n_unstaged = L.n_unstaged;
  // walker-independent part of the chi-square (inference.py:160): sum_j -ln w_j, plus y^2 w of the inactive channels
  // (model == 0 exactly there); every kernel forms the residual (y - m)^2 w of the active channels itself
  h->chi_const_fp64 = h->sum_neg_log_w + (h->y2w_prefix[C] - L.y2w_active);
  h->chi_const_mixed = h->chi_const_fp64;
  h->n_tiles_g = (int64_t)L.tiles.size(); h->n_groups = (int64_t)L.gblk.size(); h->n_recs = (int64_t)L.recs.size() - 1;
  h->n_tiles_unstaged = n_unstaged;
  h->tight.valid = false;                       // its classes are relative to this set's half-width
  h->box_lists = false;
  if (upload(h, h->d_tiles_g, L.tiles.data(), L.tiles.size() * sizeof(TileG)) ||
      upload(h, h->d_groups, L.gblk.data(), L.gblk.size() * sizeof(GroupBlk)) ||
      upload(h, h->d_recs, L.recs.data(), L.recs.size() * sizeof(LineRec)))
    return 1;
  const std::vector<TileG>& tiles_g = L.tiles;
  // ---- span table of the one-pass channel-stream kernel (make_span_table) ----
  h->n_spans = 0;
  if (h->perm_identity && n_unstaged == 0 && h->prec == CHA_PREC_MIXED) {
    std::vector<int> soff; std::vector<SpanSeg> segs;
    bool sparse = false;
    const bool span_fits = make_span_table(L, C, soff, segs, sparse);
    if (upload(h, h->d_span_tiles, soff.data(), soff.size() * sizeof(int)) ||
        upload(h, h->d_span_segs, segs.data(), segs.size() * sizeof(SpanSeg))) return 1;
    CK(cudaStreamSynchronize(h->stream));      // host vectors go out of scope
    h->span_sparse = sparse;
    h->n_spans = span_fits ? (int64_t)soff.size() - 1 : 0;
  }
  // ---- per-pair CSR, per-channel constants and tiles of the all-fp64 kernels (reference operation order, full
  //      windows) and of the untiled channel-stream fallback: built only when one of them can run ----
  int64_t n_tiles64 = (int64_t)tiles_g.size();
  if (h->prec == CHA_PREC_FP64 || n_unstaged > 0) {
    std::vector<int> act_of(C, -1);
    for (size_t a = 0; a < A; ++a) act_of[act_ch[a]] = (int)a;
    std::vector<int> off(A * M + 1, 0);
    {
      std::vector<int> cnt(A * M, 0);
      for (size_t i = 0; i < Ls; ++i)
        for (int j = wa[i]; j < wb[i]; ++j) cnt[(size_t)act_of[j] * M + h->l_mol[i]]++;
      for (size_t k = 0; k < A * M; ++k) off[k + 1] = off[k] + cnt[k];
    }
    std::vector<int> cur(off.begin(), off.end() - 1);
    std::vector<int> pline((size_t)P);
    std::vector<double> pu64((size_t)P);
    std::vector<float> pu32((size_t)P);
    for (size_t i = 0; i < Ls; ++i) {
      const double f = h->l_nu[i];
      for (int j = wa[i]; j < wb[i]; ++j) {
        int p = cur[(size_t)act_of[j] * M + h->l_mol[i]]++;
        double u = (f - x[j]) / f * kCkm;                                        // inference.py:51
        pline[p] = (int)i; pu64[p] = u; pu32[p] = (float)(u - mc);
      }
    }
    std::vector<double> ajbg(A), ab2(A);
    std::vector<float> atn(A);
    for (size_t a = 0; a < A; ++a) {
      ajbg[a] = planck_j(ax[a], kTbg, h->md.eps);
      double bsz = beam_size(ax[a], h->md.dish); ab2[a] = bsz * bsz;
    }
    std::vector<TileDev> tiles;
    size_t a0 = 0;
    while (a0 < A) {
      size_t a1 = a0 + 1;
      const double span_max = 2.0 * kTileMaxRelHalfSpan * ax[a0];
      while (a1 < A && (a1 - a0) < (size_t)kTileMaxChan && (ax[a1] - ax[a0]) <= span_max &&
             (off[a1 * M] - off[a0 * M]) < kTileMaxPairs)
        ++a1;
      TileDev t; t.c0 = (int)a0; t.c1 = (int)a1;
      t.xc = 0.5 * (ax[a0] + ax[a1 - 1]);
      t.hs = std::max(0.5 * (ax[a1 - 1] - ax[a0]), 1e-6);
      for (size_t a = a0; a < a1; ++a) atn[a] = (float)((ax[a] - t.xc) / t.hs);
      tiles.push_back(t);
      a0 = a1;
    }
    n_tiles64 = (int64_t)tiles.size();
    if (upload(h, h->d_tiles, tiles.data(), tiles.size() * sizeof(TileDev)) ||
        upload(h, h->d_poff, off.data(), off.size() * 4) || upload(h, h->d_pline, pline.data(), (size_t)P * 4) ||
        upload(h, h->d_pu64, pu64.data(), (size_t)P * 8) || upload(h, h->d_pu32, pu32.data(), (size_t)P * 4) ||
        upload(h, h->d_x, ax.data(), A * 8) || upload(h, h->d_y, ay.data(), A * 8) || upload(h, h->d_w, aw.data(), A * 8) ||
        upload(h, h->d_jbg, ajbg.data(), A * 8) || upload(h, h->d_beam2, ab2.data(), A * 8) ||
        upload(h, h->d_tn, atn.data(), A * 4) || upload(h, h->d_actof, act_of.data(), C * 4))
      return 1;
    CK(cudaStreamSynchronize(h->stream));      // host vectors go out of scope
  }
  CK(cudaStreamSynchronize(h->stream));        // host vectors go out of scope
  h->n_act = (int64_t)A; h->n_pairs = P; h->n_tiles = n_tiles64;
  h->dv_list = dv; h->hv_list = hv;
  h->pairs_dirty = false;
  h->n_rebuild++;
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build0).count();
  h->build_ms_total += ms;
  if (h->debug) fprintf(stderr, "[chalte] lists: hv %.4f dv %.4f pairs %lld tiles %lld in %.2f ms\n", hv, dv, (long long)P,
                        (long long)h->n_tiles_g, ms);
  return 0;
}

static int prepare_static(cha_handle h) {
  if (!h->model_set) FAIL("cha_set_model has not been called");
  if (!h->spec_set) FAIL("cha_set_spectrum has not been called");
  if (h->lines_dirty || h->spec_dirty) h->epoch++;
  if (h->lines_dirty && prepare_lines(h)) return 1;
  if (h->spec_dirty && prepare_spectrum(h)) return 1;
  return 0;
}

// The list must hold, for every walker of the batch, every (line, channel) pair that is inside the reference's
// mask |dv - mc| < 10 dV (inference.py:52) AND closer than kZcut sigma to some component's centre.
//   fp64 path : the full mask window, hv = 10 dV_max (reference semantics, nothing truncated)
//   mixed path: hv = min(10 dV_max, max|vlsr_c - al - mc| + kZcut * dV_max / 2.355)
static int ensure_pairs(cha_handle h, double dv_need, double dabs_need) {
  if (!(dv_need > 0.0) || !std::isfinite(dv_need)) dv_need = h->dv_list > 0 ? h->dv_list : 1e-3;
  if (!(dabs_need >= 0.0) || !std::isfinite(dabs_need)) dabs_need = 0.0;
  double hv_need = 10.0 * dv_need;
  if (h->prec == CHA_PREC_MIXED) hv_need = std::min(hv_need, dabs_need + kZcut * dv_need / kFwhm);
  // Rebuild policy (deterministic in the sequence of needs, so ranks that see the same needs hold the same lists):
  //   grown beyond the list            -> rebuild; the margin widens (2 % -> 5 % -> 12.5 % -> 30 %) while growth
  //                                       rebuilds follow each other closely (an ensemble spreading out of its
  //                                       initial ball), so that a drifting need does not rebuild every call
  //   list > 1.5x wider than needed    -> rebuild tight at once
  //   list > 1.1x wider for 64 calls   -> rebuild tight (steady state after the spread)
  h->calls_since_rebuild++;
  const bool grown = hv_need > h->hv_list || dv_need > h->dv_list;
  const bool shrunk = hv_need < h->hv_list / 1.5 || (h->calls_since_rebuild >= 64 && hv_need * 1.02 < h->hv_list / 1.1);
  if (!(h->pairs_dirty || grown || shrunk)) return 0;
  double margin = 1.02;
  if (!h->pairs_dirty && grown) {
    if (h->calls_since_rebuild < 16) h->grow_margin = std::min(1.3, 1.0 + (h->grow_margin - 1.0) * 2.5);
    else h->grow_margin = 1.02;
    margin = h->grow_margin;
  } else {
    h->grow_margin = 1.02;
  }
  h->calls_since_rebuild = 0;
  h->epoch++;
  return build_pairs(h, hv_need * margin, dv_need * margin);
}

static SpecDev spec_dev(cha_handle h) {
  SpecDev s;
  s.tiles = h->d_tiles.as<TileDev>(); s.pair_off = h->d_poff.as<int>(); s.pair_line = h->d_pline.as<int>();
  s.pair_u64 = h->d_pu64.as<double>(); s.pair_u32 = h->d_pu32.as<float>();
  s.x = h->d_x.as<double>(); s.y = h->d_y.as<double>(); s.w = h->d_w.as<double>();
  s.jbg = h->d_jbg.as<double>(); s.beam2 = h->d_beam2.as<double>(); s.tn = h->d_tn.as<float>();
  return s;
}

static PriorDev prior_dev(cha_handle h) {
  PriorDev p;
  const int nd = h->md.ndim;
  double* base = h->d_prior.as<double>();
  p.lo = base; p.hi = base + nd; p.mu = base + 2 * nd; p.sg = base + 3 * nd;
  p.gauss = h->d_prior_i.as<int>();
  p.vmin_sep = h->vmin_sep; p.vmax_sep = h->vmax_sep;
  return p;
}

// max over the rows that can reach the fused kernel of dV and of |vlsr_c - al - mc|
static void host_need(cha_handle h, const double* theta, int64_t nw, bool with_prior, double* dv, double* dabs) {
  const int nd = h->md.ndim, id = h->md.idx_dv;
  double lo = -INFINITY, hi = INFINITY;
  if (with_prior && h->prior_set) { lo = h->pr_lo[id]; hi = h->pr_hi[id]; }
  double m = 0.0, a = 0.0;
  for (int64_t w = 0; w < nw; ++w) {
    const double* th = theta + w * nd;
    double d = th[id];
    if (!(std::isfinite(d) && d > 0.0 && d > lo && d < hi)) continue;
    if (d > m) m = d;
    for (int c = 0; c < h->md.K; ++c) {
      double x = std::fabs(th[h->md.idx_vlsr[c]] - h->md.al - h->md.mc);
      if (std::isfinite(x) && x > a) a = x;
    }
  }
  *dv = m; *dabs = a;
}

// split: device int naming the first row of the wide side when the batch is served by two list sets (sampler), or
// nullptr; row_offset: row of this chunk's first walker in the whole batch
// the one-pass channel-stream kernel needs the span table (spectrum in ascending channel order, every tile staged),
// an even channel count and a 16-byte aligned output (cp.async.bulk stores of whole rows)
static bool span_stream_ok(cha_handle h, const double* d_out) {
  return h->prec == CHA_PREC_MIXED && (h->span_stream == 2 || (h->span_stream == 1 && h->span_sparse)) && h->n_spans > 0 &&
         h->n_tiles_unstaged == 0 &&
         h->xs.size() % 2 == 0 && ((uintptr_t)d_out & 15) == 0;
}

template <int K>
static void launch_chi2(cha_handle h, const double* d_theta, int nwp, const SpecDev& sp, const int* split, int row_offset,
                        const unsigned long long* void_flag) {
  if (h->prec == CHA_PREC_FP64) {
    dim3 grid((unsigned)h->n_tiles, (unsigned)(nwp / kWalkersPerBlock));
    chi2_fp64_kernel<K><<<grid, kWalkersPerBlock, 0, h->stream>>>(d_theta, nwp, h->md, h->d_ok.as<int>(), sp,
                                                                 h->d_tau.as<double>(), h->d_partial.as<double>());
  } else {
    LinesDev ln;
    ln.Kfac = h->d_lK.as<double>(); ln.El = h->d_lel.as<double>(); ln.nu = h->d_lnu.as<double>();
    ln.mol = h->d_lmol.as<int>(); ln.qinv = h->d_qinv.as<double>(); ln.lK2 = h->d_lK2.as<double>();
    ListsDev wide_set;
    wide_set.tiles = h->d_tiles_g.as<TileG>(); wide_set.groups = h->d_groups.as<GroupBlk>(); wide_set.recs = h->d_recs.as<LineRec>();
    wide_set.n_tiles = (int)h->n_tiles_g; wide_set.hv = (float)h->hv_list;
    RowSplit rs; rs.split = split; rs.row_offset = row_offset; rs.side = 0; rs.void_flag = void_flag;
    if (split && h->tight.n_tiles > 0) {
      // rows below *split: the narrow set; rows from it on: the primary lists -- one launch, tiles of both sets
      ListsDev narrow;
      narrow.tiles = h->tight.d_tiles.as<TileG>(); narrow.groups = h->tight.d_groups.as<GroupBlk>();
      narrow.recs = h->tight.d_recs.as<LineRec>(); narrow.n_tiles = (int)h->tight.n_tiles; narrow.hv = (float)h->tight.hv;
      dim3 grid((unsigned)(narrow.n_tiles + wide_set.n_tiles), (unsigned)(nwp / kWalkersPerBlock));
      chi2_mixed_kernel<K><<<grid, kWalkersPerBlock, 0, h->stream>>>(nwp, h->md, h->d_ok.as<int>(), h->d_wpf.as<float>(),
                                                                    h->d_wpd.as<double>(), narrow, wide_set, ln,
                                                                    h->d_partial.as<double>(), rs);
    } else {
      rs.split = nullptr;
      dim3 grid((unsigned)wide_set.n_tiles, (unsigned)(nwp / kWalkersPerBlock));
      chi2_mixed_kernel<K><<<grid, kWalkersPerBlock, 0, h->stream>>>(nwp, h->md, h->d_ok.as<int>(), h->d_wpf.as<float>(),
                                                                    h->d_wpd.as<double>(), wide_set, wide_set, ln,
                                                                    h->d_partial.as<double>(), rs);
    }
  }
}

template <int K>
static void launch_sim(cha_handle h, const double* d_theta, int nw, int nwp, const SpecDev& sp, double* d_out) {
  const int C = (int)h->xs.size();
  if (h->prec == CHA_PREC_MIXED && h->n_tiles_unstaged == 0) {
    LinesDev ln;
    ln.Kfac = h->d_lK.as<double>(); ln.El = h->d_lel.as<double>(); ln.nu = h->d_lnu.as<double>();
    ln.mol = h->d_lmol.as<int>(); ln.qinv = h->d_qinv.as<double>(); ln.lK2 = h->d_lK2.as<double>();
    if (span_stream_ok(h, d_out)) {
      // one pass, every byte written once: walker tables, then CTA = span of kSpanCh channels x 32 walkers whose rows
      // leave as TMA bulk stores
      if (h->n_tiles_g > 0) {
        sim_gcoef_kernel<K><<<dim3((unsigned)(nwp / kWalkersPerBlock), (unsigned)h->n_tiles_g), kWalkersPerBlock, 0, h->stream>>>(
            nwp, h->md, h->d_ok.as<int>(), h->d_wpd.as<double>(), h->d_tiles_g.as<TileG>(), h->d_gco.as<float>());
        h->n_launch++;
      }
      cudaFuncSetAttribute(simulate_span_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpanDynSmem);
      dim3 grid((unsigned)h->n_spans, (unsigned)((nw + kSpanWalkers - 1) / kSpanWalkers));
      if (!h->capturing) cudaEventRecord(h->ev0, h->stream);          // cha_stat(h, 7): the span kernel of the last call
      simulate_span_kernel<K><<<grid, 256, kSpanDynSmem, h->stream>>>(nw, nwp, h->md, h->d_ok.as<int>(), h->d_wpf.as<float>(),
          h->d_groups.as<GroupBlk>(), h->d_recs.as<LineRec>(), h->d_tau.as<float>(),
          h->d_gco.as<float>(), h->d_span_tiles.as<int>(), h->d_span_segs.as<SpanSeg>(), (size_t)C, d_out);
      if (!h->capturing) { cudaEventRecord(h->ev1, h->stream); h->ev_valid = true; }
      return;
    }
    // spectra not in ascending channel order, and dense line forests: inactive channels are exactly zero (one HBM
    // write stream), then the active channels tile by tile
    if (!h->capturing) cudaEventRecord(h->ev0, h->stream);            // cha_stat(h, 7): zero-fill + tiles of the last call
    cudaMemsetAsync(d_out, 0, (size_t)nw * C * 8, h->stream);
    if (h->n_tiles_g == 0) { if (!h->capturing) { cudaEventRecord(h->ev1, h->stream); h->ev_valid = true; } return; }
    dim3 grid((unsigned)h->n_tiles_g, (unsigned)((nw + kSimWalkers - 1) / kSimWalkers));
    simulate_tiles_kernel<K><<<grid, 256, 0, h->stream>>>(nw, nwp, h->md, h->d_ok.as<int>(), h->d_wpf.as<float>(),
                                                         h->d_wpd.as<double>(), h->d_tiles_g.as<TileG>(),
                                                         h->d_groups.as<GroupBlk>(), h->d_recs.as<LineRec>(), ln,
                                                         (size_t)C, d_out);
    if (!h->capturing) { cudaEventRecord(h->ev1, h->stream); h->ev_valid = true; }
    h->n_launch++;
    return;
  }
  dim3 grid((unsigned)((C + 255) / 256), (unsigned)nw);
  if (h->prec == CHA_PREC_FP64)
    simulate_kernel<K, false><<<grid, 256, 0, h->stream>>>(d_theta, nw, nwp, h->md, h->d_ok.as<int>(), sp, C,
                                                          h->d_actof.as<int>(), h->d_outpos.as<int>(),
                                                          h->d_xall.as<double>(), h->d_tau.p, d_out);
  else
    simulate_kernel<K, true><<<grid, 256, 0, h->stream>>>(d_theta, nw, nwp, h->md, h->d_ok.as<int>(), sp, C,
                                                         h->d_actof.as<int>(), h->d_outpos.as<int>(),
                                                         h->d_xall.as<double>(), h->d_tau.p, d_out);
}

#define DISPATCH_K(FN, ...)                                   \
  switch (h->md.K) {                                          \
    case 1: FN<1>(__VA_ARGS__); break;                        \
    case 2: FN<2>(__VA_ARGS__); break;                        \
    case 3: FN<3>(__VA_ARGS__); break;                        \
    case 4: FN<4>(__VA_ARGS__); break;                        \
    case 5: FN<5>(__VA_ARGS__); break;                        \
    case 6: FN<6>(__VA_ARGS__); break;                        \
    case 7: FN<7>(__VA_ARGS__); break;                        \
    default: FN<8>(__VA_ARGS__); break;                       \
  }

static constexpr int64_t kGraphMaxWalkers = 4096;   // above this the launch sequence is not latency bound
static constexpr int kSortMinWalkers = (int)kGraphMaxWalkers;   // reach-ordered evaluation: only batches that are not graph-replayed

// mode: 0 lnlike, 1 lnprob, 2 lnprior only, 3 simulate (d_out = [nw * C])
// the pair list must already cover the batch (ensure_pairs)
static int eval_device(cha_handle h, const double* d_theta, int64_t nw64, double* d_out, int mode,
                       unsigned long long* d_need_slot = nullptr, unsigned long long* h_need_publish = nullptr,
                       const int* d_split = nullptr, int row_offset = 0, const unsigned long long* void_flag = nullptr,
                       bool may_sort = false) {
  if (nw64 <= 0) return 0;
  const int nw = (int)nw64;
  const int nwp = (nw + kWalkersPerBlock - 1) / kWalkersPerBlock * kWalkersPerBlock;
  const int M = h->md.M, nd = h->md.ndim;
  const int with_prior = (mode == 1 || mode == 2) ? 1 : 0;
  if (with_prior && !h->prior_set) FAIL("cha_set_prior has not been called");
  const size_t Ls = h->l_nu.size();
  const int nqc = h->n_qchunks_max;
  CK(h->d_ok.ensure((size_t)nwp * 4)); CK(h->d_lp.ensure((size_t)nwp * 8));
  CK(h->d_qinv.ensure((size_t)2 * M * nwp * 8)); CK(h->d_qpart.ensure((size_t)M * nqc * nwp * 8));
  if (!h->prior_set) { CK(h->d_prior.ensure(8)); CK(h->d_prior_i.ensure(8)); }
  for (int m = 0; m < M; ++m) {
    if (h->mol[m].q_kind != CHA_Q_SUM) continue;
    const int ns = (int)h->mol[m].sg.size();
    const int nch = (ns + kQChunk - 1) / kQChunk;
    q_state_sum_kernel<<<dim3(nwp / 32, nch), 256, 0, h->stream>>>(d_theta, nw, nd, h->md.idx_tex,
        h->mol[m].d_sg.as<double>(), h->mol[m].d_sE.as<double>(), ns,
        h->d_qpart.as<double>() + (size_t)m * nqc * nwp, nwp);
    h->n_launch++;
  }
  const int K = h->md.K;
  CK(h->d_wpf.ensure((size_t)(2 + K + M * K) * nwp * 4)); CK(h->d_wpd.ensure((size_t)(1 + K) * nwp * 8));
  // Reach-ordered evaluation of a caller's batch (single-component fits on the mixed path, batches too large for graph
  // replay): rows are evaluated in order of their reach class -- how far from the mask centre a walker's own 6-sigma
  // range extends -- so that the walkers of a warp skip the same records (chi2_mixed_kernel, SKIP variant).  Only the
  // slot a row is evaluated at changes: walker_prep_kernel writes its tables there, finalize_kernel writes the result
  // back to the caller's row; log-probs do not depend on the slot, so the output is bit-identical either way.
  // Adaptive: row_reach_class_kernel counts the warps (32 consecutive caller rows) whose classes differ by >= 2; the
  // batch is sorted while more than a quarter of them do, and one call in 32 probes when sorting is off.
  const int* d_dest = nullptr; const int* d_inv = nullptr;
  if (may_sort && mode <= 1 && h->sort_rows != 0 && h->prec == CHA_PREC_MIXED && K == 1 && M == 1 && !d_split && !void_flag &&
      !h->capturing && nw > kSortMinWalkers && h->hv_list > 0.0 && h->n_tiles_unstaged == 0 && h->n_tiles_g > 0) {
    const unsigned long long spread = h->h_sortstat[0], warps = h->h_sortstat[1];     // published by an earlier call
    if (warps > 0) h->sort_on = 4 * spread > warps;
    const bool probe = (h->sort_calls++ % 32) == 0;
    if (h->sort_rows > 0 || h->sort_on || probe) {
      CK(h->d_rcls.ensure((size_t)nwp * 4)); CK(h->d_rdest.ensure((size_t)nwp * 4)); CK(h->d_rinv.ensure((size_t)nwp * 4));
      row_reach_class_kernel<<<(nw + 255) / 256, 256, 0, h->stream>>>(d_theta, nw, h->md, (float)(1.0 / h->hv_list),
                                                                       h->d_rcls.as<int>(), h->d_sortstat.as<unsigned long long>());
      reach_sort_kernel<<<1, kSortThreads, 0, h->stream>>>(nw, h->d_rcls.as<int>(), h->d_rdest.as<int>(), -1, nullptr, nw, nd,
                                                          nullptr, nullptr, nullptr, nullptr, h->d_rinv.as<int>(),
                                                          h->d_sortstat.as<unsigned long long>(), h->h_sortstat);
      h->n_launch += 2; h->n_sorted++;
      d_dest = h->d_rdest.as<int>(); d_inv = h->d_rinv.as<int>();
    }
  }
  walker_prep_kernel<<<nwp / 128, 128, 0, h->stream>>>(d_theta, nw, nwp, h->md, prior_dev(h), with_prior,
      h->d_qdesc.as<QDesc>(), h->d_qpart.as<double>(), nqc, h->d_ok.as<int>(), h->d_lp.as<double>(),
      h->d_qinv.as<double>(), h->d_wpf.as<float>(), h->d_wpd.as<double>(), d_need_slot,
      with_prior && h->prior_set ? h->pr_lo[h->md.idx_dv] : -INFINITY,
      with_prior && h->prior_set ? h->pr_hi[h->md.idx_dv] : INFINITY, d_dest);
  h->n_launch++;
  if (mode == 2) {
    prior_only_kernel<<<(nw + 127) / 128, 128, 0, h->stream>>>(nw, h->d_lp.as<double>(), d_out);
    h->n_launch++;
    CK(cudaGetLastError());
    return 0;
  }
  const bool f64 = h->prec == CHA_PREC_FP64;
  // the line-strength table is consumed by the all-fp64 kernel and by the channel-stream kernel; the fused mixed
  // kernel computes the strengths of each tile's lines itself
  const bool sim_tiled = mode == 3 && !f64 && h->n_tiles_unstaged == 0;
  if (Ls && h->n_tiles && (f64 || (mode == 3 && !sim_tiled))) {
    CK(h->d_tau.ensure(Ls * (size_t)nwp * (f64 ? 8 : 4)));
    const int lpb = 8;
    dim3 g((unsigned)(nwp / kWalkersPerBlock), (unsigned)((Ls + lpb - 1) / lpb));
    if (f64)
      line_tau_kernel<double><<<g, kWalkersPerBlock, 0, h->stream>>>(d_theta, nwp, nd, h->md.idx_tex, h->d_ok.as<int>(),
          h->d_qinv.as<double>(), (int)Ls, h->d_lK.as<double>(), h->d_lel.as<double>(), h->d_lnu.as<double>(),
          h->d_lmol.as<int>(), h->d_tau.as<double>(), lpb);
    else
      line_tau_fast_kernel<<<g, kWalkersPerBlock, 0, h->stream>>>(d_theta, nwp, nd, h->md.idx_tex, h->d_ok.as<int>(),
          h->d_qinv.as<double>(), (int)Ls, h->d_lK.as<double>(), h->d_lel.as<double>(), h->d_lnu.as<double>(),
          h->d_lmol.as<int>(), h->d_tau.as<float>(), lpb);
    h->n_launch++;
  }
  if (mode == 3 && span_stream_ok(h, d_out) && h->n_tiles_g > 0) {
    // walker tables of the one-pass channel-stream kernel: line strengths [line][walker]
    CK(h->d_tau.ensure(std::max<size_t>(Ls, 1) * (size_t)nwp * 4));
    CK(h->d_gco.ensure((size_t)h->n_tiles_g * nwp * 4 * K * 4));
    if (Ls) {
      LinesDev ln;
      ln.Kfac = h->d_lK.as<double>(); ln.El = h->d_lel.as<double>(); ln.nu = h->d_lnu.as<double>();
      ln.mol = h->d_lmol.as<int>(); ln.qinv = h->d_qinv.as<double>(); ln.lK2 = h->d_lK2.as<double>();
      const int lpb = 8;
      sim_line_tau_kernel<<<dim3((unsigned)(nwp / kWalkersPerBlock), (unsigned)((Ls + lpb - 1) / lpb)), kWalkersPerBlock, 0, h->stream>>>(
          nwp, h->d_ok.as<int>(), h->d_wpd.as<double>(), ln, (int)Ls, h->d_tau.as<float>(), lpb);
      h->n_launch++;
    }
  }
  SpecDev sp = spec_dev(h);
  if (mode == 3) {
    if (!(Ls && h->n_tiles)) { CK(h->d_tau.ensure(8)); }
    DISPATCH_K(launch_sim, h, d_theta, nw, nwp, sp, d_out);
    h->n_launch++;
    CK(cudaGetLastError());
    return 0;
  }
  const int64_t nt_used = f64 ? h->n_tiles : h->n_tiles_g;
  if (f64 || !h->tight.valid || h->n_tiles_unstaged || h->tight.n_unstaged || h->tight.n_tiles == 0) d_split = nullptr;   // one list set
  const int64_t nt_tight = d_split ? h->tight.n_tiles : 0;
  CK(h->d_partial.ensure((size_t)std::max<int64_t>(std::max(nt_used, nt_tight), 1) * nwp * 8));
  const bool any_tiles = Ls && h->n_tiles;
  if (any_tiles) {
    if (!h->capturing) CK(cudaEventRecord(h->ev0, h->stream));      // timing events do not belong in a captured graph
    DISPATCH_K(launch_chi2, h, d_theta, nwp, sp, d_split, row_offset, void_flag);
    if (!h->capturing) { CK(cudaEventRecord(h->ev1, h->stream)); h->ev_valid = true; }
    h->n_launch++;
  }
  RowSplit rs; rs.split = d_split; rs.row_offset = row_offset; rs.side = 0; rs.void_flag = nullptr;
  const double cc = f64 ? h->chi_const_fp64 : h->chi_const_mixed;
  finalize_kernel<<<(nw + 31) / 32, 32 * kFinSlices, 0, h->stream>>>(nw, nwp,
      d_split ? (int)nt_tight : (any_tiles ? (int)nt_used : 0), h->d_partial.as<double>(),
      d_split ? h->tight.chi_const : cc, h->d_ok.as<int>(), h->d_lp.as<double>(),
      with_prior, d_out, h_need_publish ? d_need_slot : nullptr, h_need_publish,
      rs, any_tiles ? (int)nt_used : 0, cc, d_inv);
  h->n_launch++;
  CK(cudaGetLastError());
  return 0;
}

static constexpr int64_t kChunkWalkers = 16384;
static constexpr size_t kMaxGraphs = 8;

static void drop_graphs(cha_handle h) {
  for (auto& e : h->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
  h->graphs.clear();
  h->seen_keys.clear();
}

// enqueue() puts the sequence on h->stream.  First sighting of a key: plain launches (this also sizes the workspace);
// second sighting (among the last few keys: a sampler alternates its two colours): captured, instantiated and
// launched as a graph; afterwards: one cudaGraphLaunch.
template <class F>
static int run_graphed(cha_handle h, GraphKey key, F&& enqueue) {
  key.epoch = h->epoch; key.buf_epoch = g_buf_epoch.load(std::memory_order_relaxed);
  for (auto& e : h->graphs)
    if (e.key == key) {
      CK(cudaGraphLaunch(e.exec, h->stream));
      h->n_launch += e.launches; h->n_graph_launch++;
      e.last_use = ++h->graph_clock;
      return 0;
    }
  {
    bool seen = false;
    for (const auto& k : h->seen_keys) if (k == key) { seen = true; break; }
    if (!seen) {
      if (h->seen_keys.size() >= kMaxGraphs) h->seen_keys.erase(h->seen_keys.begin());
      h->seen_keys.push_back(key);
      return enqueue();
    }
  }
  for (size_t i = 0; i < h->graphs.size();) {                                   // graphs of an older configuration
    if (h->graphs[i].key.epoch != key.epoch || h->graphs[i].key.buf_epoch != key.buf_epoch) {
      cudaGraphExecDestroy(h->graphs[i].exec);
      h->graphs.erase(h->graphs.begin() + i);
    } else ++i;
  }
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return enqueue(); }
  const int64_t l0 = h->n_launch;
  h->capturing = true;
  const int rc = enqueue();
  h->capturing = false;
  cudaGraph_t g = nullptr;
  const cudaError_t ec = cudaStreamEndCapture(h->stream, &g);
  const int launches = (int)(h->n_launch - l0);
  h->n_launch = l0;
  bool ok = rc == 0 && ec == cudaSuccess && g != nullptr && h->epoch == key.epoch &&
            g_buf_epoch.load(std::memory_order_relaxed) == key.buf_epoch;
  cudaGraphExec_t exec = nullptr;
  if (ok && cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) ok = false;
  if (g) cudaGraphDestroy(g);
  if (!ok) {
    cudaGetLastError();
    if (exec) cudaGraphExecDestroy(exec);
    h->seen_keys.clear();
    return enqueue();
  }
  if (h->graphs.size() >= kMaxGraphs) {
    size_t lru = 0;
    for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].last_use < h->graphs[lru].last_use) lru = i;
    cudaGraphExecDestroy(h->graphs[lru].exec);
    h->graphs.erase(h->graphs.begin() + lru);
  }
  GraphEntry e; e.key = key; e.exec = exec; e.launches = launches; e.last_use = ++h->graph_clock;
  h->graphs.push_back(e);
  CK(cudaGraphLaunch(exec, h->stream));
  h->n_launch += launches; h->n_graph_launch++;
  return 0;
}

static constexpr int kMaxPend = 256;     // queued calls between two validations: the stream drains at every one of them
static constexpr int kSyncSlot = kMaxPend;                 // need slot of the synchronous path (never a pending call's)
static constexpr int kPoisonIdx = 2 * (kMaxPend + 1);      // index (u64 units) of the sticky skip flag in d_need
static constexpr double kSamplerNeedMargin = 1.15;
static int drain(cha_handle h);
static double hv_needed(cha_handle h, double dv, double dabs);

static int eval_host(cha_handle h, const double* theta, int64_t nw, double* out, int mode) {
  if (!h) return 1;
  if (nw < 0) FAIL("nw < 0");
  if (nw == 0) return 0;
  if (!theta || !out) FAIL("null buffer");
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  if (prepare_static(h)) return 1;
  const int nd = h->md.ndim;
  const int64_t C = (int64_t)h->xs.size();
  const int64_t chunk = mode == 3 ? std::max<int64_t>(1, std::min<int64_t>(kChunkWalkers, (int64_t)(1ll << 28) / std::max<int64_t>(C, 1)))
                                  : kChunkWalkers;
  // A log-prob batch that fits one chunk is launched against the current pair list without first scanning theta on
  // the host: walker_prep_kernel reduces the batch's maxima, they come back with the result, and only a batch the
  // list did not cover is evaluated again after the rebuild (the same contract as cha_log_prob_dev).
  const bool optimistic = mode <= 1 && !h->pairs_dirty && h->h_need && nw <= chunk && h->dv_list > 0.0;
  if (mode != 2 && !optimistic) {
    double dv_need = 0.0, dabs_need = 0.0;
    host_need(h, theta, nw, mode == 1, &dv_need, &dabs_need);
    if (ensure_pairs(h, dv_need, dabs_need)) return 1;
  }
  const size_t out_per = mode == 3 ? (size_t)C : 1;
  if (ensure_pin(h, (size_t)std::min(nw, chunk) * (nd + out_per) * 8)) return 1;
  bool any_qsum = false;
  for (int m = 0; m < h->md.M; ++m) any_qsum = any_qsum || h->mol[m].q_kind == CHA_Q_SUM;
  for (int64_t w0 = 0; w0 < nw; w0 += chunk) {
    const int64_t n = std::min(chunk, nw - w0);
    CK(h->d_theta.ensure((size_t)n * nd * 8));
    CK(h->d_out.ensure((size_t)n * out_per * 8));
    std::memcpy(h->h_pin, theta + w0 * nd, (size_t)n * nd * 8);
    double* stage = h->h_pin + (size_t)n * nd;
    unsigned long long* d_m = optimistic ? h->d_need.as<unsigned long long>() : nullptr;     // slot 0: nothing is pending
    // The pinned staging buffer is mapped into the device's address space (UVA): where a kernel touches each value
    // exactly once it reads theta / writes the result there directly, which removes a copy operation (and its
    // scheduling gap) from each end of the dependent chain.  theta: walker_prep_kernel is its only reader on the
    // mixed path without a state-sum partition function; results: finalize_kernel / prior_only_kernel.
    const bool zc_out = mode <= 2;
    const bool zc_in = zc_out && h->prec == CHA_PREC_MIXED && !any_qsum;
    const double* th_dev = zc_in ? h->h_pin : h->d_theta.as<double>();
    double* out_dev = zc_out ? stage : h->d_out.as<double>();
    auto enqueue = [&]() -> int {
      if (!zc_in) CK(cudaMemcpyAsync(h->d_theta.p, h->h_pin, (size_t)n * nd * 8, cudaMemcpyHostToDevice, h->stream));
      // need slots are zero at rest; finalize_kernel publishes the batch maxima to h_need and re-zeroes the slot
      if (eval_device(h, th_dev, n, out_dev, mode, d_m, d_m ? h->h_need : nullptr, nullptr, 0, nullptr, /*may_sort=*/true)) return 1;
      if (!zc_out) CK(cudaMemcpyAsync(stage, h->d_out.p, (size_t)n * out_per * 8, cudaMemcpyDeviceToHost, h->stream));
      return 0;
    };
    if (mode != 3 && nw <= kGraphMaxWalkers) {
      GraphKey key; key.kind = 1; key.nw = n; key.mode = mode | (optimistic ? 4 : 0);
      if (run_graphed(h, key, enqueue)) return 1;
    } else if (enqueue()) return 1;
    CK(cudaStreamSynchronize(h->stream));
    if (optimistic) {
      double dv, dabs;
      std::memcpy(&dv, h->h_need, 8); std::memcpy(&dabs, h->h_need + 1, 8);
      const bool covered = dv <= h->dv_list && hv_needed(h, dv, dabs) <= h->hv_list;
      if (ensure_pairs(h, dv, dabs)) return 1;          // rebuild policy (growth now, shrinkage for the next call)
      if (!covered) {
        d_m = nullptr;
        if (enqueue()) return 1;
        CK(cudaStreamSynchronize(h->stream));
      }
    }
    std::memcpy(out + w0 * out_per, stage, (size_t)n * out_per * 8);
  }
  if (mode <= 1 && h->ev_valid) cudaEventElapsedTime(&h->last_fused_ms, h->ev0, h->ev1);
  return 0;
}

// device-resident variant: dV_max of the batch is reduced on the device (one 8-byte D2H + sync)
static int device_need(cha_handle h, const double* d_theta, int64_t nw, bool with_prior, double* dv, double* dabs) {
  CK(h->d_scratch.ensure(64));
  unsigned long long* d_m = h->d_scratch.as<unsigned long long>();
  CK(cudaMemsetAsync(d_m, 0, 16, h->stream));
  double lo = -INFINITY, hi = INFINITY;
  if (with_prior && h->prior_set) { lo = h->pr_lo[h->md.idx_dv]; hi = h->pr_hi[h->md.idx_dv]; }
  dv_max_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, h->stream>>>(d_theta, (int)nw, h->md, lo, hi, d_m);
  h->n_launch++;
  if (ensure_pin(h, 64)) return 1;
  CK(cudaMemcpyAsync(h->h_pin, d_m, 16, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::memcpy(dv, h->h_pin, 8); std::memcpy(dabs, h->h_pin + 1, 8);
  return 0;
}

static double hv_needed(cha_handle h, double dv, double dabs) {
  double hv = 10.0 * dv;
  if (h->prec == CHA_PREC_MIXED) hv = std::min(hv, dabs + kZcut * dv / kFwhm);
  return hv;
}

// evaluation on device pointers, list checked BEFORE the launch (one 16-byte D2H + stream sync)
static int log_prob_dev_sync(cha_handle h, const double* d_theta, int64_t nw, double* d_out, int with_prior) {
  if (prepare_static(h)) return 1;
  double need = 0.0, dabs = 0.0;
  if (device_need(h, d_theta, nw, with_prior != 0, &need, &dabs)) return 1;
  if (ensure_pairs(h, need, dabs)) return 1;
  for (int64_t w0 = 0; w0 < nw; w0 += kChunkWalkers) {
    const int64_t n = std::min(kChunkWalkers, nw - w0);
    if (eval_device(h, d_theta + w0 * h->md.ndim, n, d_out + w0, with_prior ? 1 : 0, nullptr, nullptr, nullptr, 0, nullptr, /*may_sort=*/true)) return 1;
  }
  return 0;
}

// optimistic evaluation: no host round trip before the launch.  *slot_out receives the need slot (or -1 when the
// call had to take the synchronous path: first call, list marked dirty, or while re-running)
static int log_prob_dev_opt(cha_handle h, const double* d_theta, int64_t nw, double* d_out, int with_prior, int* slot_out) {
  *slot_out = -1;
  if (prepare_static(h)) return 1;
  if (h->pairs_dirty || h->in_redo || !h->h_need) return log_prob_dev_sync(h, d_theta, nw, d_out, with_prior);
  if ((int)h->pend.size() >= kMaxPend && drain(h)) return 1;
  if (h->pairs_dirty) return log_prob_dev_sync(h, d_theta, nw, d_out, with_prior);   // drain may have asked for a rebuild
  const int slot = (int)h->pend.size();
  // the batch's maxima are reduced by walker_prep_kernel into slot `slot` of d_need and copied to its pinned mirror
  auto enqueue = [&]() -> int {
    // need slots are zero at rest (cha_create, drain); the last chunk's finalize_kernel publishes the maxima to the
    // pinned mirror and re-zeroes the slot: no memset or copy operation in the sequence
    unsigned long long* d_m = h->d_need.as<unsigned long long>() + 2 * slot;
    for (int64_t w0 = 0; w0 < nw; w0 += kChunkWalkers) {
      const int64_t n = std::min(kChunkWalkers, nw - w0);
      const bool last = w0 + n >= nw;
      if (eval_device(h, d_theta + w0 * h->md.ndim, n, d_out + w0, with_prior ? 1 : 0, d_m, last ? h->h_need + 2 * slot : nullptr,
                      nullptr, 0, nullptr, /*may_sort=*/true))
        return 1;
    }
    return 0;
  };
  if (nw <= kGraphMaxWalkers) {
    GraphKey key; key.kind = 0; key.a = d_theta; key.b = d_out; key.nw = nw; key.mode = with_prior ? 1 : 0; key.slot = slot;
    if (run_graphed(h, key, enqueue)) return 1;
  } else if (enqueue()) return 1;
  *slot_out = slot;
  return 0;
}

static int sampler_half_step_impl(cha_handle h, int64_t step, int split, const double* d_all_coords, int64_t store_slot);

// ---- bulk / outlier split of the sampler's batches ------------------------------------------------------------
// The primary (wide) lists must cover the widest proposal of a half-step; the bulk reaches about half as far.  From the
// class histogram of ALL proposals of the last half-step (identical on every rank) the host picks the smallest reach
// class that holds >= 88 % (95 %) of them and keeps a second list set of that half-width; reach_sort_kernel routes every
// proposal by its OWN class.  Called only at synchronisation points (stream idle, nothing pending).
static int tight_class_from_hist(const int* hist, double quantile) {
  long long total = 0;
  for (int c = 0; c < kReachClasses; ++c) total += hist[c];
  if (total < 256) return -2;                                   // no information
  long long cum = 0;
  for (int c = 0; c < kReachClasses; ++c) {
    cum += hist[c];
    if ((double)cum >= quantile * (double)total) return reach_class_upper(c) <= 0.8 ? std::max(c, 1) : -1;
  }
  return -1;
}

static int refresh_tight(cha_handle h, bool from_hist) {
  if (!h->two_lists || h->prec != CHA_PREC_MIXED || h->pairs_dirty || !(h->hv_list > 0.0)) return 0;
  if (from_hist && h->h_hist) {
    // the wider the primary lists are against the bulk, the dearer an outlier: 95 % of the proposals go to the narrow
    // set when the primary lists span the prior box, 88 % when they follow the ensemble
    const int want = tight_class_from_hist(h->h_hist, h->box_lists ? 0.95 : 0.88);
    if (want != -2 && want != h->tight.cls) {
      // a class change costs a list build: adopt it at once when there is no narrow set yet, otherwise only when two
      // consecutive synchronisation points ask for the same one
      h->tight_want_streak = want == h->tight_want ? h->tight_want_streak + 1 : 1;
      h->tight_want = want;
      if (h->tight.cls < 0 || h->tight_want_streak >= 2) { h->tight.cls = want; h->tight.valid = false; h->tight_want_streak = 0; }
    } else {
      h->tight_want_streak = 0;
    }
  }
  if (h->tight.cls < 0) { if (h->tight.valid) h->epoch++; h->tight.valid = false; return 0; }
  if (h->tight.valid) return 0;
  h->epoch++;                                   // captured half-steps hold the set's pointers, sizes and class
  return build_tight(h, reach_class_upper(h->tight.cls) * h->hv_list * (1.0 + 1e-6));
}

// synchronisation point of the optimistic calls: everything from the first call the lists did not cover is run again.
// The first such call takes the synchronous path (it rebuilds the lists); the calls behind it -- void on the device, they
// left at once -- are queued again as ordinary optimistic calls and validated by the next pass of the loop.
static int log_prob_dev_opt(cha_handle h, const double* d_theta, int64_t nw, double* d_out, int with_prior, int* slot_out);
static int drain_impl(cha_handle h);
static int drain(cha_handle h) {
  if (h->pend.empty()) return 0;
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = drain_impl(h);
  h->drain_ms_total += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  h->n_drain++;
  return rc;
}
static int drain_impl(cha_handle h) {
  while (!h->pend.empty()) {
    const auto tw0 = std::chrono::steady_clock::now();
    CK(cudaStreamSynchronize(h->stream));
    if (h->debug) fprintf(stderr, "[chalte] sync point: %zu queued calls, stream drained after %.2f ms\n", h->pend.size(),
                          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count());
    CK(cudaMemsetAsync(h->d_need.p, 0, kMaxPend * 16, h->stream));    // the mirrors are on the host; slots start clean again
    size_t bad = h->pend.size();
    double last_hv = 0.0, win_dv = 0.0, win_dabs = 0.0, win_hv = 0.0;
    for (size_t i = 0; i < h->pend.size(); ++i) {
      double dv, dabs;
      std::memcpy(&dv, h->h_need + 2 * i, 8); std::memcpy(&dabs, h->h_need + 2 * i + 1, 8);
      const double hv = hv_needed(h, dv, dabs);
      win_dv = std::max(win_dv, dv); win_dabs = std::max(win_dabs, dabs); win_hv = std::max(win_hv, hv);
      if (!(dv <= h->pend[i].dv_cover && hv <= h->pend[i].hv_cover)) { bad = i; break; }
      last_hv = hv;
    }
    // what the calls of this window needed, remembered with a slow decay (3 % per synchronisation point): the widest
    // proposal of a half-step is a heavy-tailed quantity, and the lists are sized for its recent maximum, not for the
    // half-step that happens to be at hand when they are rebuilt
    h->dv_hi = std::max(win_dv, 0.97 * h->dv_hi); h->dabs_hi = std::max(win_dabs, 0.97 * h->dabs_hi);
    std::vector<cha_engine::Pend> redo(h->pend.begin() + bad, h->pend.end());
    bool had_sampler = false;
    for (const auto& P : h->pend) had_sampler = had_sampler || P.kind == 1;
    h->pend.clear();
    if (redo.empty()) {
      // list much wider than needed (at once), or moderately wider over 64 optimistic calls: rebuild at the next call.
      // With a narrow set serving the bulk the primary lists only see the outliers: their width costs little and a
      // rebuild a lot, so they are left alone until they are really too wide.
      const bool two = h->tight.valid && had_sampler;     // plain log-prob calls never use the narrow set: for them the
                                                          // primary lists are sized as if it did not exist
      if (h->box_lists && had_sampler) {
        h->slack_calls = 0;                      // lists that span the prior box are not resized by the sampler
      } else {
        const double ref_hv = two ? std::max(last_hv, hv_needed(h, h->dv_hi, h->dabs_hi)) : last_hv;
        const bool slack = ref_hv > 0.0 && ref_hv * 1.02 < h->hv_list / (two ? 2.0 : 1.1);
        h->slack_calls = slack ? h->slack_calls + (int64_t)bad : 0;
        if ((ref_hv > 0.0 && ref_hv < h->hv_list / (two ? 3.0 : 1.5)) || h->slack_calls >= (two ? 512 : 64)) {
          h->pairs_dirty = true; h->slack_calls = 0;
        }
      }
      if (had_sampler && refresh_tight(h, true)) return 1;
      return 0;
    }
    h->slack_calls = 0;
    h->n_events++;
    if (h->debug) fprintf(stderr, "[chalte] call %zu of %zu queued was not covered (hv %.4f): re-running %zu\n", bad,
                          bad + redo.size(), h->hv_list, redo.size());
    CK(cudaMemsetAsync(h->d_need.as<unsigned long long>() + kPoisonIdx, 0, 8, h->stream));   // clear the sticky skip flag
    int rc = 0;
    for (size_t i = 0; i < redo.size() && !rc; ++i) {
      const auto& P = redo[i];
      h->n_redo++;
      h->in_redo = i == 0;                       // the first one synchronously: it is the one that outgrew the lists
      if (P.kind == 0) {
        if (i == 0) rc = log_prob_dev_sync(h, P.d_theta, P.nw, P.d_out, P.with_prior);
        else {
          int slot = -1;
          rc = log_prob_dev_opt(h, P.d_theta, P.nw, P.d_out, P.with_prior, &slot);
          if (!rc && slot >= 0) { cha_engine::Pend Q = P; Q.dv_cover = h->dv_list; Q.hv_cover = h->hv_list; h->pend.push_back(Q); }
        }
      } else {
        rc = sampler_half_step_impl(h, P.step, P.split, P.d_all, P.store_slot);
      }
      h->in_redo = false;
    }
    if (rc) return rc;
  }
  return 0;
}

// evaluation against the current list, no need bookkeeping
static int eval_chunks(cha_handle h, const double* d_theta, int64_t nw, double* d_out, int mode, const int* d_split = nullptr,
                       const unsigned long long* void_flag = nullptr) {
  for (int64_t w0 = 0; w0 < nw; w0 += kChunkWalkers) {
    const int64_t n = std::min(kChunkWalkers, nw - w0);
    if (eval_device(h, d_theta + w0 * h->md.ndim, n, d_out + w0, mode, nullptr, nullptr, d_split, (int)w0, void_flag)) return 1;
  }
  return 0;
}

// The resident sampler sizes the pair list from the proposals of the WHOLE ensemble (proposal_need_kernel: every rank
// recomputes all of them), never from its local share, so every rank holds the same list at the same step whatever
// the sharding.
//   d_all_coords : positions of the whole ensemble supplied by the caller (cha_sampler_half_step), or nullptr:
//                  the engine's own -- refreshed here by an all-gather on the handle's stream when the walkers are
//                  sharded over ranks (cha_comm_init), the resident local array itself on a single rank
//   store_slot   : chain slot the state is appended to after this half-step (-1: none)
static int sampler_half_step_impl(cha_handle h, int64_t step, int split, const double* d_all_coords, int64_t store_slot) {
  const int nd = h->md.ndim;
  const int nl = (int)h->s_nw_local;
  if (prepare_static(h)) return 1;
  double hi_dv = INFINITY;
  if (h->prior_set) hi_dv = h->pr_hi[h->md.idx_dv];
  const bool optimistic = !(h->pairs_dirty || h->in_redo || !h->s_logp_valid) && h->h_dyn;
  if (optimistic && (int)h->pend.size() >= kMaxPend) {
    if (drain(h)) return 1;
    return sampler_half_step_impl(h, step, split, d_all_coords, store_slot);
  }
  const double* caller_all = d_all_coords;
  if (!d_all_coords) {
    if (h->comm) {
      // the one exchange of the complementary-ensemble move: every rank's resident positions into the
      // replicated ensemble array, in stream order before the proposals that read it
      ncclResult_t r = g_nccl.AllGather(h->s_coords.p, h->s_all.p, (size_t)nl * nd, ncclDouble, h->comm, h->stream);
      if (r != ncclSuccess) FAIL(std::string("ncclAllGather: ") + g_nccl.GetErrorString(r));
      h->n_coll++; h->coll_bytes += (int64_t)h->s_nw_global * nd * 8;
      d_all_coords = h->s_all.as<double>();
    } else {
      if (h->s_nw_local != h->s_nw_global) FAIL("walkers are sharded but no communicator is attached (cha_comm_init)");
      d_all_coords = h->s_coords.as<double>();
    }
  }
  const int slot = optimistic ? (int)h->pend.size() : kSyncSlot;
  unsigned long long* need_base = h->d_need.as<unsigned long long>();
  unsigned long long* d_poison = need_base + kPoisonIdx;
  // local walkers of this colour: global ids w0..w0+nl with (id & 1) == split, compacted in order
  const int n_move = (int)(((h->s_w0 + nl + (split ? 0 : 1)) >> 1) - ((h->s_w0 + (split ? 0 : 1)) >> 1));
  const int ncol = (int)((h->s_nw_global + 1) / 2);
  // half-ensembles of >= 1024 proposals are evaluated in order of their reach (lte_sampler.cuh: reach_sort_kernel).
  // The test is on the GLOBAL half-ensemble: with two list sets a proposal's log-prob depends (at the 1e-8 level) on the
  // set it is evaluated against, so whether the split is in use must not depend on how the walkers are sharded.
  int* d_cls = ncol >= 1024 && n_move > 0 ? h->s_cls.as<int>() : nullptr;
  int* d_dest = d_cls ? h->s_dest.as<int>() : nullptr;
  int* d_hist = d_cls && h->two_lists ? h->d_hist.as<int>() : nullptr;
  PriorDev pr = prior_dev(h);
  // Per-launch values travel in a 16-byte device record (step index, need slot) refreshed in stream order before the
  // half-step, so that the kernels' arguments are the same from one half-step to the next and the sequence can be
  // replayed as a CUDA graph.  The synchronous path passes them by value (dyn == nullptr).
  const SamplerDyn* dyn = optimistic ? h->d_dyn.as<SamplerDyn>() : nullptr;
  auto launch_need = [&](float inv_hv_ref) {
    proposal_need_kernel<<<(ncol + 127) / 128, 128, 0, h->stream>>>(
        d_all_coords, (int)h->s_nw_global, h->md, split, h->s_seed, (unsigned long long)step, h->s_a, pr.lo, pr.hi,
        dyn ? need_base : need_base + 2 * slot, dyn, (int)h->s_w0, nl, inv_hv_ref, d_cls, d_hist);
    h->n_launch++;
  };
  if (!optimistic) {
    // ---- synchronous: what this half-step needs is known before it is evaluated, the lists are (re)built for it ----
    unsigned long long* d_m = need_base + 2 * slot;
    CK(cudaMemsetAsync(d_m, 0, 16, h->stream));
    launch_need(h->hv_list > 0.0 ? (float)(1.0 / h->hv_list) : 1.0f);
    if (!h->s_logp_valid) {          // the first half-step also evaluates the current positions of the local walkers
      dv_max_kernel<<<(unsigned)((h->s_nw_global + 255) / 256), 256, 0, h->stream>>>(d_all_coords, (int)h->s_nw_global, h->md,
                                                                                      h->pr_lo[h->md.idx_dv], hi_dv, d_m);
      h->n_launch++;
    }
    CK(cudaMemcpyAsync(h->h_need + 2 * slot, d_m, 16, cudaMemcpyDeviceToHost, h->stream));
    // this path owns need slot kSyncSlot, so the maxima of calls still pending are not disturbed; they are validated
    // first (when this call itself is a re-run from drain() nothing is pending)
    if (!h->pend.empty() && drain(h)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    double dv, dabs;
    std::memcpy(&dv, h->h_need + 2 * slot, 8); std::memcpy(&dabs, h->h_need + 2 * slot + 1, 8);
    CK(cudaMemsetAsync(d_m, 0, 16, h->stream));
    // half-steps queue up without a host round trip; a list that fails to cover one stalls the whole queue until the
    // next synchronisation, so the sampler asks for more than this half-step needs: 15 %, or -- with a narrow set
    // serving the bulk, when the primary lists see only the outliers and their width is cheap -- 40 % and at least
    // 1.25 x the recent maximum
    const double hv_before = h->hv_list;
    const bool wide_is_cheap = h->two_lists && h->tight.cls >= 0;
    const double margin = wide_is_cheap ? 1.4 : kSamplerNeedMargin;
    double dv_ask = wide_is_cheap ? std::max(dv * margin, h->dv_hi * 1.25) : dv * margin;
    double dabs_ask = wide_is_cheap ? std::max(dabs * margin, h->dabs_hi * 1.25) : dabs * margin;
    // Large ensembles whose prior box bounds dV and every vlsr: the primary lists are built ONCE for the box.  No
    // proposal inside the box can outgrow them (those outside are dropped by the prior before any evaluation), so no
    // half-step is ever void and the lists are never resized; only the outliers of a half-step pay for their width,
    // the bulk runs against the narrow set.
    bool want_box = false;
    if (d_cls && h->two_lists && h->prec == CHA_PREC_MIXED && h->prior_set) {
      double dv_box = h->pr_hi[h->md.idx_dv], dabs_box = 0.0;
      const bool dv_finite = std::isfinite(dv_box) && dv_box > 0.0;
      bool vl_finite = true;
      for (int c = 0; c < h->md.K && vl_finite; ++c) {
        const double lo = h->pr_lo[h->md.idx_vlsr[c]], hi = h->pr_hi[h->md.idx_vlsr[c]];
        vl_finite = std::isfinite(lo) && std::isfinite(hi);
        dabs_box = std::max(dabs_box, std::max(std::fabs(lo - h->md.al - h->md.mc), std::fabs(hi - h->md.al - h->md.mc)));
      }
      // a prior that bounds dV but leaves the velocities free (the 4-component layout: only their ORDER is bounded,
      // TMC1_four_component.py:224-233): the box in dV, twice the recent maximum in |vlsr - mask centre|
      if (dv_finite && !vl_finite) dabs_box = 2.0 * std::max(dabs, h->dabs_hi);
      if (dv_finite && hv_needed(h, dv_box, dabs_box) <= 8.0 * std::max(hv_needed(h, dv, dabs), 1e-6)) {
        dv_ask = dv_box; dabs_ask = dabs_box; want_box = true;
      }
    }
    if (ensure_pairs(h, dv_ask, dabs_ask)) return 1;
    h->box_lists = want_box;                    // (any other rebuild of the primary lists clears the flag)
    if (h->hv_list != hv_before) {
      // the reach classes are relative to the primary half-width: the proposals were classified against the old one.
      // Re-classify against the new lists (same kernel, same proposals) and keep the narrow set's absolute width.
      if (h->tight.cls >= 0 && hv_before > 0.0) {
        const double want_hv = reach_class_upper(h->tight.cls) * hv_before;
        int c = 1;
        while (c < kReachClasses - 1 && reach_class_upper(c) * h->hv_list < want_hv) ++c;
        h->tight.cls = reach_class_upper(c) <= 0.8 ? c : -1;
      }
      if (d_cls) {
        if (d_hist) CK(cudaMemsetAsync(d_hist, 0, kReachClasses * 4, h->stream));
        launch_need((float)(1.0 / h->hv_list));
        CK(cudaMemsetAsync(d_m, 0, 16, h->stream));
      }
    }
    if (refresh_tight(h, false)) return 1;
    if (!h->s_logp_valid) {
      // log-probabilities of the local walkers with the ensemble-sized list (cha_sampler_init had only local data)
      if (eval_chunks(h, h->s_coords.as<double>(), nl, h->s_logp.as<double>(), 1)) return 1;
      h->s_logp_valid = true;
    }
  }
  // evaluation order: by reach class; with a valid narrow set the bulk comes first and the outliers start at a block
  // boundary (dead rows in between), each side evaluated against its own list set
  const bool two = d_cls && h->two_lists && h->tight.valid && h->tight.cls >= 0 && h->prec == CHA_PREC_MIXED &&
                   h->n_tiles_unstaged == 0 && h->tight.n_unstaged == 0;
  const int n_rows = two ? (n_move + 127) / 128 * 128 + 128 : n_move;
  int* d_split = two ? h->d_split.as<int>() : nullptr;
  const float inv_hv_ref = h->hv_list > 0.0 ? (float)(1.0 / h->hv_list) : 1.0f;
  auto enqueue = [&]() -> int {
    if (optimistic) launch_need(inv_hv_ref);          // (the synchronous path has run it already)
    if (d_cls) {
      reach_sort_kernel<<<1, kSortThreads, 0, h->stream>>>(n_move, d_cls, d_dest, two ? h->tight.cls : -1, d_split, n_rows, nd,
                                                          h->s_prop.as<double>(), h->s_idx.as<int>(), d_hist, h->h_hist,
                                                          nullptr, nullptr, nullptr);
      h->n_launch++;
    }
    // proposals for local walkers of colour `split` (compacted), partners drawn from the other colour
    stretch_propose_kernel<<<(nl + 127) / 128, 128, 0, h->stream>>>(
        d_all_coords, (int)h->s_nw_global, (int)h->s_w0, nl, nd, split, h->s_seed, (unsigned long long)step, h->s_a,
        h->s_prop.as<double>(), h->s_factor.as<double>(), h->s_idx.as<int>(), dyn, d_dest);
    h->n_launch++;
    if (n_move > 0 && eval_chunks(h, h->s_prop.as<double>(), n_rows, h->s_newlp.as<double>(), 1, d_split,
                                  optimistic ? d_poison : nullptr)) return 1;
    // accept / reject in place; on the optimistic path the kernel first checks on the device that the lists covered
    // the ensemble bound and otherwise leaves the state untouched, raises the sticky flag (every later queued
    // evaluation leaves at once) and the half-step is run again by drain().  It also publishes the proposals' maxima
    // to the host's pinned mirror.  A rank that moves no walker of this colour still runs the check: the flag must go
    // up on every rank at the same half-step.
    if (n_move > 0 || optimistic) {
      ListCover cov;
      cov.need = optimistic ? need_base : nullptr;
      cov.dv_cover = h->dv_list; cov.hv_cover = h->hv_list; cov.mixed = h->prec == CHA_PREC_MIXED ? 1 : 0;
      cov.zc = kZcut; cov.fwhm = kFwhm;
      cov.poison = optimistic ? d_poison : nullptr;
      stretch_accept_kernel<<<std::max(1, (n_rows + 127) / 128), 128, 0, h->stream>>>(
          n_rows, nd, (int)h->s_w0, h->s_idx.as<int>(), h->s_prop.as<double>(), h->s_newlp.as<double>(),
          h->s_factor.as<double>(), h->s_seed, (unsigned long long)step, h->s_coords.as<double>(),
          h->s_logp.as<double>(), h->s_acc.as<unsigned long long>(), cov, dyn, optimistic ? h->h_need : nullptr);
      h->n_launch++;
    }
    CK(cudaGetLastError());
    return 0;
  };
  if (optimistic) {
    h->h_dyn[slot].step = (unsigned long long)step; h->h_dyn[slot].slot = slot; h->h_dyn[slot].pad = 0;
    CK(cudaMemcpyAsync(h->d_dyn.p, &h->h_dyn[slot], sizeof(SamplerDyn), cudaMemcpyHostToDevice, h->stream));
    // the sequence of a half-step is a dozen dependent launches: replayed as one CUDA graph from its second sighting on
    // (one graph per colour; list rebuilds, a new narrow set or a re-initialised sampler start a new epoch)
    GraphKey key; key.kind = 2; key.a = d_all_coords; key.nw = nl; key.mode = split | (two ? 2 : 0);
    if (h->sampler_graphs) { if (run_graphed(h, key, enqueue)) return 1; }
    else if (enqueue()) return 1;
  } else if (enqueue()) return 1;
  if (store_slot >= 0) {
    if (store_slot >= h->s_chain_cap) FAIL("chain slot beyond the reserved store");
    const int n = nl * nd;
    chain_store_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(nl, nd, h->s_coords.as<double>(), h->s_logp.as<double>(),
        h->s_chain_c.as<double>() + (size_t)store_slot * nl * nd, h->s_chain_l.as<double>() + (size_t)store_slot * nl,
        optimistic ? d_poison : nullptr);
    h->n_launch++;
  }
  if (optimistic) {
    cha_engine::Pend P{};
    P.kind = 1; P.step = step; P.split = split; P.d_all = caller_all; P.store_slot = store_slot;
    P.dv_cover = h->dv_list; P.hv_cover = h->hv_list;
    h->pend.push_back(P);
  }
  CK(cudaGetLastError());
  return 0;
}

// =============================================================================================
// extern "C" boundary
// =============================================================================================
extern "C" {

int cha_version(void) { return 100; }

int cha_create(int device_id, cha_handle* out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this engine has no CPU fallback)";
    return 1;
  }
  if (device_id < 0 || device_id >= ndev) { g_create_error = "device_id out of range"; return 1; }
  if ((e = cudaSetDevice(device_id)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return 1; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device_id);
  if (prop.major != 10) {
    g_create_error = "libchalte is built for sm_100a (Blackwell B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return 1;
  }
  cha_engine* h = new cha_engine();
  h->dev = device_id;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
    g_create_error = "stream/event creation failed"; delete h; return 1;
  }
  h->md.ndim = 0; h->md.K = 1; h->md.M = 1;
  if (h->d_need.ensure((kMaxPend + 1) * 16 + 16) != cudaSuccess ||
      cudaMemset(h->d_need.p, 0, (kMaxPend + 1) * 16 + 16) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_need, (kMaxPend + 1) * 16) != cudaSuccess ||
      h->d_hist.ensure(kReachClasses * 4) != cudaSuccess || cudaMemset(h->d_hist.p, 0, kReachClasses * 4) != cudaSuccess ||
      h->d_split.ensure(16) != cudaSuccess || cudaMemset(h->d_split.p, 0, 16) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_hist, kReachClasses * 4) != cudaSuccess ||
      h->d_sortstat.ensure(16) != cudaSuccess || cudaMemset(h->d_sortstat.p, 0, 16) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_sortstat, 16) != cudaSuccess ||
      h->d_dyn.ensure(sizeof(SamplerDyn)) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_dyn, kMaxPend * sizeof(SamplerDyn)) != cudaSuccess) {
    g_create_error = "allocation of the coverage-check buffers failed"; cha_destroy(h); return 1;
  }
  std::memset(h->h_hist, 0, kReachClasses * 4);
  std::memset(h->h_sortstat, 0, 16);
  if (const char* e6 = std::getenv("CHALTE_SORT_ROWS")) h->sort_rows = std::atoi(e6);
  if (const char* e2 = std::getenv("CHALTE_TWO_LISTS")) h->two_lists = std::atoi(e2) != 0;
  if (const char* e3 = std::getenv("CHALTE_DEBUG")) h->debug = std::atoi(e3) != 0;
  if (const char* e4 = std::getenv("CHALTE_SAMPLER_GRAPHS")) h->sampler_graphs = std::atoi(e4) != 0;
  if (const char* e5 = std::getenv("CHALTE_SPAN_STREAM")) h->span_stream = std::atoi(e5);
  *out = h;
  return 0;
}

int cha_destroy(cha_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->dev);
  cudaStreamSynchronize(h->stream);
  drop_graphs(h);
  DevBuf* bufs[] = {&h->d_lnu, &h->d_llogint, &h->d_lel, &h->d_lK, &h->d_lK2, &h->d_lmol, &h->d_qdesc, &h->d_prior, &h->d_prior_i,
                    &h->d_tiles, &h->d_poff, &h->d_pline, &h->d_pu64, &h->d_pu32, &h->d_x, &h->d_y, &h->d_w, &h->d_jbg,
                    &h->d_beam2, &h->d_tn, &h->d_tiles_g, &h->d_groups, &h->d_recs, &h->d_span_tiles, &h->d_span_segs, &h->d_rcls, &h->d_rdest, &h->d_rinv, &h->d_sortstat, &h->d_xall, &h->d_actof, &h->d_outpos, &h->d_theta, &h->d_out, &h->d_ok,
                    &h->d_lp, &h->d_wpf, &h->d_wpd, &h->d_qinv, &h->d_qpart, &h->d_tau, &h->d_gco, &h->d_partial, &h->d_scratch, &h->d_sim,
                    &h->s_coords, &h->s_logp, &h->s_prop, &h->s_newlp, &h->s_factor, &h->s_acc, &h->s_idx, &h->s_cls, &h->s_dest, &h->d_need};
  for (DevBuf* b : bufs) b->release();
  h->s_all.release(); h->s_chain_c.release(); h->s_chain_l.release();
  h->tight.d_tiles.release(); h->tight.d_groups.release(); h->tight.d_recs.release();
  h->d_hist.release(); h->d_split.release();
  if (h->h_hist) cudaFreeHost(h->h_hist);
  if (h->h_sortstat) cudaFreeHost(h->h_sortstat);
  if (h->comm && g_nccl.lib) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
  if (h->h_need) cudaFreeHost(h->h_need);
  if (h->h_dyn) cudaFreeHost(h->h_dyn);
  h->d_dyn.release();
  for (int m = 0; m < kMaxM; ++m) { h->mol[m].d_sg.release(); h->mol[m].d_sE.release(); }
  if (h->h_pin) cudaFreeHost(h->h_pin);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

const char* cha_last_error(cha_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cha_set_molecule(cha_handle h, int mol_id, int64_t n_lines, const double* nu, const double* logint,
                     const double* elower, int q_kind, const double* q_params, int n_q_params,
                     int64_t n_states, const double* state_g, const double* state_E,
                     double ll, double ul, const int64_t* line_idx, int64_t n_sel) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (mol_id < 0 || mol_id >= kMaxM) FAIL("mol_id out of range");
  if (n_lines <= 0 || !nu || !logint || !elower) FAIL("empty catalog");
  if (q_kind < 0 || q_kind > 3 || n_q_params < 0 || n_q_params > 8) FAIL("bad partition-function descriptor");
  if (q_kind == CHA_Q_SUM && (n_states <= 0 || !state_g || !state_E)) FAIL("CHA_Q_SUM needs the state table");
  for (int64_t i = 1; i < n_lines; ++i) if (nu[i] < nu[i - 1]) FAIL("catalog is not frequency-sorted");
  HostMol& m = h->mol[mol_id];
  m.nu.assign(nu, nu + n_lines); m.logint.assign(logint, logint + n_lines); m.elower.assign(elower, elower + n_lines);
  m.q_kind = q_kind; m.n_qp = n_q_params;
  for (int k = 0; k < 8; ++k) m.qp[k] = k < n_q_params ? q_params[k] : 0.0;
  if (q_kind == CHA_Q_SUM) { m.sg.assign(state_g, state_g + n_states); m.sE.assign(state_E, state_E + n_states); }
  else { m.sg.clear(); m.sE.clear(); }
  m.ll = ll; m.ul = ul;
  m.all_lines = (line_idx == nullptr);
  if (line_idx) m.line_idx.assign(line_idx, line_idx + n_sel); else m.line_idx.clear();
  m.set = true;
  h->lines_dirty = true;
  return 0;
}

int cha_set_spectrum(cha_handle h, int64_t n_chan, const double* freq, const double* y, const double* yerr) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (n_chan < 0 || (n_chan && (!freq || !y || !yerr))) FAIL("bad spectrum buffers");
  if (n_chan > (int64_t)0x7ffffff0) FAIL("too many channels");
  h->sx.assign(freq, freq + n_chan); h->sy.assign(y, y + n_chan); h->syerr.assign(yerr, yerr + n_chan);
  h->spec_set = true; h->spec_dirty = true;
  return 0;
}

int cha_set_model(cha_handle h, int ndim, int n_comp, int n_mol, const int* idx_ss, const int* idx_ncol, int idx_tex,
                  const int* idx_vlsr, int idx_dv, double fixed_ss, double dish_size, double aligned_velocity,
                  double mask_centre, double planck_eps) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (ndim < 1 || ndim > kMaxNdim) FAIL("ndim out of range");
  if (n_comp < 1 || n_comp > kMaxK) FAIL("n_comp out of range");
  if (n_mol < 1 || n_mol > kMaxM) FAIL("n_mol out of range");
  auto okidx = [&](int i) { return i >= 0 && i < ndim; };
  if (!okidx(idx_tex) || !okidx(idx_dv)) FAIL("idx_tex/idx_dv out of range");
  ModelDev md{};
  md.ndim = ndim; md.K = n_comp; md.M = n_mol; md.idx_tex = idx_tex; md.idx_dv = idx_dv;
  for (int c = 0; c < n_comp; ++c) {
    if (idx_ss[c] >= ndim || !okidx(idx_vlsr[c])) FAIL("component index out of range");
    if (idx_ss[c] < 0 && !(fixed_ss == fixed_ss)) FAIL("fixed source size requested but fixed_ss is NaN");
    md.idx_ss[c] = idx_ss[c]; md.idx_vlsr[c] = idx_vlsr[c];
    for (int m = 0; m < n_mol; ++m) {
      if (!okidx(idx_ncol[m * n_comp + c])) FAIL("idx_ncol out of range");
      md.idx_ncol[m * n_comp + c] = idx_ncol[m * n_comp + c];
    }
  }
  md.fixed_ss = fixed_ss; md.dish = dish_size; md.al = aligned_velocity; md.mc = mask_centre; md.eps = planck_eps;
  const bool relines = (md.M != h->md.M);
  h->md = md;
  h->epoch++;
  h->model_set = true;
  h->pairs_dirty = true;
  if (relines) h->lines_dirty = true;
  h->prior_set = false;
  return 0;
}

int cha_set_prior(cha_handle h, const double* lo, const double* hi, const double* mu, const double* sigma,
                  const int* gauss, double vlsr_min_sep, double vlsr_max_sep) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (!h->model_set) FAIL("cha_set_model must precede cha_set_prior");
  const int nd = h->md.ndim;
  CK(cudaSetDevice(h->dev));
  h->pr_lo.assign(lo, lo + nd); h->pr_hi.assign(hi, hi + nd); h->pr_mu.assign(mu, mu + nd);
  h->pr_sg.assign(sigma, sigma + nd); h->pr_gauss.assign(gauss, gauss + nd);
  h->vmin_sep = vlsr_min_sep; h->vmax_sep = vlsr_max_sep;
  h->epoch++;
  std::vector<double> pack(4 * nd);
  for (int p = 0; p < nd; ++p) { pack[p] = lo[p]; pack[nd + p] = hi[p]; pack[2 * nd + p] = mu[p]; pack[3 * nd + p] = sigma[p]; }
  if (upload(h, h->d_prior, pack.data(), pack.size() * 8) || upload(h, h->d_prior_i, h->pr_gauss.data(), nd * 4)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  h->prior_set = true;
  return 0;
}

int cha_set_precision(cha_handle h, int prec) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (prec != CHA_PREC_FP64 && prec != CHA_PREC_MIXED) FAIL("unknown precision");
  if (prec != h->prec) { h->pairs_dirty = true; h->epoch++; }   // fp64 keeps the full mask windows, mixed truncates at kZcut sigma
  h->prec = prec;
  return 0;
}

int cha_log_prob(cha_handle h, const double* theta, int64_t nw, double* out) { return eval_host(h, theta, nw, out, 1); }
int cha_log_like(cha_handle h, const double* theta, int64_t nw, double* out) { return eval_host(h, theta, nw, out, 0); }
int cha_log_prior(cha_handle h, const double* theta, int64_t nw, double* out) { return eval_host(h, theta, nw, out, 2); }
int cha_simulate(cha_handle h, const double* theta, int64_t nw, double* out) { return eval_host(h, theta, nw, out, 3); }

int cha_log_prob_dev(cha_handle h, const double* d_theta, int64_t nw, double* d_out, int with_prior) {
  if (!h) return 1;
  if (nw <= 0) return 0;
  CK(cudaSetDevice(h->dev));
  int slot = -1;
  if (log_prob_dev_opt(h, d_theta, nw, d_out, with_prior, &slot)) return 1;
  if (slot >= 0) {
    cha_engine::Pend P{};
    P.kind = 0; P.d_theta = d_theta; P.nw = nw; P.d_out = d_out; P.with_prior = with_prior;
    P.dv_cover = h->dv_list; P.hv_cover = h->hv_list;
    h->pend.push_back(P);
  }
  return 0;
}

int cha_simulate_dev(cha_handle h, const double* d_theta, int64_t nw, double* d_out) {
  if (!h) return 1;
  if (nw <= 0) return 0;
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  if (prepare_static(h)) return 1;
  double need = 0.0, dabs = 0.0;
  if (device_need(h, d_theta, nw, false, &need, &dabs)) return 1;
  if (ensure_pairs(h, need, dabs)) return 1;
  const int64_t C = (int64_t)h->xs.size();
  for (int64_t w0 = 0; w0 < nw; w0 += kChunkWalkers) {
    const int64_t n = std::min(kChunkWalkers, nw - w0);
    if (eval_device(h, d_theta + w0 * h->md.ndim, n, d_out + w0 * C, 3)) return 1;
  }
  return 0;
}

int cha_sync(cha_handle h) {
  if (!h) return 1;
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  if (h->ev_valid) cudaEventElapsedTime(&h->last_fused_ms, h->ev0, h->ev1);    // only once both events were recorded
  return 0;
}

void* cha_stream(cha_handle h) { return h ? (void*)h->stream : nullptr; }

int64_t cha_stat(cha_handle h, int what) {
  if (!h) return -1;
  switch (what) {
    case 0: return h->n_launch;
    case 1: return (int64_t)h->l_nu.size();
    case 2: return h->n_act;
    case 3: return h->n_pairs;
    case 4: return h->prec == CHA_PREC_FP64 ? h->n_tiles : h->n_tiles_g;
    case 8: return h->n_groups;
    case 9: return h->n_recs;
    case 5: return (int64_t)llround(h->dv_list * 1e9);
    case 10: return (int64_t)llround(h->hv_list * 1e9);
    case 6: return h->n_rebuild;
    case 7: return (int64_t)llround((double)h->last_fused_ms * 1e6);
    case 11: return (int64_t)llround(h->build_ms_total * 1e3);     // host microseconds spent building lists
    case 12: return h->n_graph_launch;                              // launch sequences replayed as one CUDA graph
    case 13: return h->n_coll;                                      // all-gathers enqueued
    case 14: return h->coll_bytes;                                  // bytes received in them (this rank)
    case 15: return h->n_redo;                                      // queued calls re-run after a list rebuild
    case 16: return h->tight.valid ? h->tight.n_tiles : 0;          // narrow list set of the sampler: tiles
    case 17: return h->tight.valid ? h->tight.n_pairs : 0;          //   (line, channel) pairs
    case 18: return h->n_rebuild_tight;                             //   builds
    case 19: return h->tight.valid ? (int64_t)llround(h->tight.hv * 1e9) : 0;   //   half-width (km/s x1e9)
    case 20: return (int64_t)llround(h->drain_ms_total * 1e3);      // host microseconds inside synchronisation points
    case 21: return h->n_drain;                                     // synchronisation points that had queued calls
    case 22: return h->n_events;                                    // ... of which found a call the lists had not covered
    case 23: return h->n_sorted;                                    // log-prob batches evaluated in order of reach class
    default: return -1;
  }
}

int cha_count_window_pairs(cha_handle h, const double* theta, int64_t nw, int64_t* out) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (nw <= 0) return 0;
  CK(cudaSetDevice(h->dev));
  if (prepare_static(h)) return 1;
  const int nd = h->md.ndim;
  const int Ls = (int)h->l_nu.size(), C = (int)h->xs.size();
  CK(h->d_theta.ensure((size_t)nw * nd * 8));
  CK(h->d_scratch.ensure((size_t)nw * 8));
  CK(cudaMemcpyAsync(h->d_theta.p, theta, (size_t)nw * nd * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->d_scratch.p, 0, (size_t)nw * 8, h->stream));
  if (Ls && C) {
    for (int64_t w0 = 0; w0 < nw; w0 += 32768) {
      const int n = (int)std::min<int64_t>(32768, nw - w0);
      count_window_pairs_kernel<<<dim3((Ls + 127) / 128, n), 128, 0, h->stream>>>(
          h->d_theta.as<double>() + w0 * nd, n, nd, h->md.idx_dv, h->md.mc, Ls, h->d_lnu.as<double>(), C,
          h->d_xall.as<double>(), h->d_scratch.as<unsigned long long>() + w0);
      h->n_launch++;
    }
  }
  CK(cudaMemcpyAsync(out, h->d_scratch.p, (size_t)nw * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- sampler (kernels in lte_sampler.cuh) -----------------------------------------------------
int cha_sampler_init(cha_handle h, int64_t nw_global, int64_t w0, int64_t nw_local, const double* coords_local,
                     uint64_t seed, double stretch_a) {
  if (!h) return 1;
  if (!h->model_set || !h->prior_set) FAIL("model and prior must be set before the sampler");
  if (nw_global < 2 || (nw_global & 1)) FAIL("nw_global must be even (two equal half-ensembles)");
  if (w0 < 0 || nw_local <= 0 || w0 + nw_local > nw_global) FAIL("bad local walker range");
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  const int nd = h->md.ndim;
  h->s_nw_global = nw_global; h->s_w0 = w0; h->s_nw_local = nw_local; h->s_seed = seed; h->s_a = stretch_a;
  h->s_accepted = 0;
  CK(h->s_coords.ensure((size_t)nw_local * nd * 8)); CK(h->s_logp.ensure((size_t)nw_local * 8));
  const size_t rows = (size_t)nw_local + 384;        // the evaluation batch may be padded to block boundaries (two list sets)
  CK(h->s_prop.ensure(rows * nd * 8)); CK(h->s_newlp.ensure(rows * 8));
  CK(h->s_factor.ensure(rows * 8)); CK(h->s_acc.ensure(16)); CK(h->s_idx.ensure(rows * 4));
  CK(cudaMemsetAsync(h->s_idx.p, 0xff, rows * 4, h->stream));
  CK(cudaMemsetAsync(h->s_prop.p, 0, rows * nd * 8, h->stream));
  h->tight.valid = false; h->tight.cls = -1; h->tight_want = -1; h->tight_want_streak = 0;
  h->dv_hi = 0.0; h->dabs_hi = 0.0; h->box_lists = false;
  CK(h->s_cls.ensure((size_t)nw_local * 4)); CK(h->s_dest.ensure((size_t)nw_local * 4));
  if (h->comm) {
    if (nw_local * h->comm_world != nw_global || w0 != (int64_t)h->comm_rank * nw_local)
      FAIL("with a communicator the walkers must be split into equal contiguous shares in rank order");
    CK(h->s_all.ensure((size_t)nw_global * nd * 8));
  }
  h->s_chain_n = 0;
  CK(cudaMemcpyAsync(h->s_coords.p, coords_local, (size_t)nw_local * nd * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->s_acc.p, 0, 16, h->stream));
  // the initial log-probabilities are computed by the first half-step, which sees the whole ensemble and sizes the
  // pair list from it (identical on every rank); until then cha_sampler_get evaluates them on demand
  h->s_logp_valid = false;
  h->pairs_dirty = true;
  h->epoch++;                    // ensemble geometry and seed are baked into captured half-steps
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int cha_sampler_half_step(cha_handle h, int64_t step, int split, const double* d_all_coords) {
  if (!h) return 1;
  if (!h->s_nw_local) FAIL("sampler not initialised");
  CK(cudaSetDevice(h->dev));
  if (!d_all_coords) FAIL("d_all_coords is NULL (cha_sampler_run gathers the ensemble itself)");
  return sampler_half_step_impl(h, step, split, d_all_coords, -1);
}

// ---- walkers sharded over ranks --------------------------------------------------------------------
int cha_comm_unique_id(unsigned char id[CHA_COMM_ID_BYTES]) {
  static_assert(CHA_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "CHA_COMM_ID_BYTES must equal NCCL_UNIQUE_ID_BYTES");
  if (!id) return 1;
  if (!g_nccl.load()) { g_create_error = g_nccl.err; return 1; }
  ncclUniqueId u;
  ncclResult_t r = g_nccl.GetUniqueId(&u);
  if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return 1; }
  std::memcpy(id, u.internal, CHA_COMM_ID_BYTES);
  return 0;
}

int cha_comm_init(cha_handle h, int rank, int world, const unsigned char id[CHA_COMM_ID_BYTES]) {
  if (!h) return 1;
  if (world < 1 || rank < 0 || rank >= world || !id) FAIL("bad rank/world/id");
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  if (h->comm && g_nccl.lib) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
  h->comm_rank = rank; h->comm_world = world;
  if (world == 1) return 0;                      // one rank: the resident positions are the ensemble, no exchange
  if (!g_nccl.load()) FAIL(g_nccl.err);
  ncclUniqueId u;
  std::memcpy(u.internal, id, CHA_COMM_ID_BYTES);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, u, rank);
  if (r != ncclSuccess) { h->comm = nullptr; FAIL(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
  h->s_nw_local = 0;                             // a sampler initialised before has no gathered-ensemble array
  h->epoch++;
  return 0;
}

int cha_comm_destroy(cha_handle h) {
  if (!h) return 1;
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  if (h->comm && g_nccl.lib) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr; h->comm_rank = 0; h->comm_world = 1;
  return 0;
}

int cha_sampler_run(cha_handle h, int64_t step0, int64_t n_steps, int64_t store_every) {
  if (!h) return 1;
  if (!h->s_nw_local) FAIL("sampler not initialised");
  if (n_steps < 0 || store_every < 0) FAIL("n_steps / store_every < 0");
  CK(cudaSetDevice(h->dev));
  const int nd = h->md.ndim;
  const int64_t nl = h->s_nw_local;
  const int64_t add = store_every > 0 ? n_steps / store_every : 0;
  if (h->s_chain_n + add > h->s_chain_cap) {
    // grow the resident chain (pending half-steps name their slot by index, but they run on the old buffers)
    if (drain(h)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    const int64_t cap = std::max<int64_t>(h->s_chain_n + add, 2 * h->s_chain_cap);
    DevBuf nc, nlp;
    CK(nc.ensure((size_t)cap * nl * nd * 8)); CK(nlp.ensure((size_t)cap * nl * 8));
    if (h->s_chain_n) {
      CK(cudaMemcpyAsync(nc.p, h->s_chain_c.p, (size_t)h->s_chain_n * nl * nd * 8, cudaMemcpyDeviceToDevice, h->stream));
      CK(cudaMemcpyAsync(nlp.p, h->s_chain_l.p, (size_t)h->s_chain_n * nl * 8, cudaMemcpyDeviceToDevice, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    }
    h->s_chain_c.release(); h->s_chain_l.release();
    h->s_chain_c = nc; h->s_chain_l = nlp; h->s_chain_cap = cap;
  }
  for (int64_t s = 0; s < n_steps; ++s)
    for (int split = 0; split < 2; ++split) {
      const bool st = split == 1 && store_every > 0 && (s + 1) % store_every == 0;
      if (sampler_half_step_impl(h, step0 + s, split, nullptr, st ? h->s_chain_n : -1)) return 1;
      if (st) h->s_chain_n++;
    }
  return 0;
}

int64_t cha_sampler_chain_len(cha_handle h) { return h ? h->s_chain_n : -1; }

int cha_sampler_chain_read(cha_handle h, int64_t slot0, int64_t n_slots, double* coords, double* logp) {
  if (!h) return 1;
  if (!h->s_nw_local) FAIL("sampler not initialised");
  if (slot0 < 0 || n_slots < 0 || slot0 + n_slots > h->s_chain_n) FAIL("chain slots out of range");
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  const size_t nl = (size_t)h->s_nw_local, nd = (size_t)h->md.ndim;
  if (n_slots && coords)
    CK(cudaMemcpyAsync(coords, h->s_chain_c.as<double>() + (size_t)slot0 * nl * nd, (size_t)n_slots * nl * nd * 8,
                       cudaMemcpyDeviceToHost, h->stream));
  if (n_slots && logp)
    CK(cudaMemcpyAsync(logp, h->s_chain_l.as<double>() + (size_t)slot0 * nl, (size_t)n_slots * nl * 8,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int cha_sampler_chain_clear(cha_handle h) {
  if (!h) return 1;
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  h->s_chain_n = 0;
  return 0;
}

int cha_sampler_coords_dev(cha_handle h, double** d_coords, double** d_logp) {
  if (!h) return 1;
  if (!h->s_nw_local) FAIL("sampler not initialised");
  if (d_coords) *d_coords = h->s_coords.as<double>();
  if (d_logp) *d_logp = h->s_logp.as<double>();
  return 0;
}

int cha_sampler_get(cha_handle h, double* coords_local, double* logp_local, int64_t* n_accepted) {
  if (!h) return 1;
  if (!h->s_nw_local) FAIL("sampler not initialised");
  CK(cudaSetDevice(h->dev));
  if (drain(h)) return 1;
  const int nd = h->md.ndim;
  if (!h->s_logp_valid && logp_local) {      // before the first step: evaluate, but leave the ensemble-sized pass to it
    if (log_prob_dev_sync(h, h->s_coords.as<double>(), h->s_nw_local, h->s_logp.as<double>(), 1)) return 1;
    h->pairs_dirty = true;
  }
  if (coords_local) CK(cudaMemcpyAsync(coords_local, h->s_coords.p, (size_t)h->s_nw_local * nd * 8, cudaMemcpyDeviceToHost, h->stream));
  if (logp_local) CK(cudaMemcpyAsync(logp_local, h->s_logp.p, (size_t)h->s_nw_local * 8, cudaMemcpyDeviceToHost, h->stream));
  unsigned long long acc = 0;
  CK(cudaMemcpyAsync(&acc, h->s_acc.p, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (n_accepted) *n_accepted = (int64_t)acc;
  return 0;
}

}  // extern "C"

// device Q(T) of one molecule for a single temperature (state sum runs on the device)
static int q_single(cha_handle h, int m, double T, double* out) {
  HostMol& hm = h->mol[m];
  QDesc qd; qd.kind = hm.q_kind; qd.n_params = hm.n_qp;
  for (int k = 0; k < 8; ++k) qd.p[k] = hm.qp[k];
  if (hm.q_kind != CHA_Q_SUM) { *out = q_analytic(qd, T); return 0; }
  const int ns = (int)hm.sg.size();
  const int nch = (ns + kQChunk - 1) / kQChunk;
  if (upload(h, hm.d_sg, hm.sg.data(), hm.sg.size() * 8) || upload(h, hm.d_sE, hm.sE.data(), hm.sE.size() * 8)) return 1;
  CK(h->d_scratch.ensure(64 * 8 + (size_t)nch * 128 * 8));
  double* d_t = h->d_scratch.as<double>();
  double* d_qp = d_t + 64;
  CK(cudaMemcpyAsync(d_t, &T, 8, cudaMemcpyHostToDevice, h->stream));
  q_state_sum_kernel<<<dim3(1, nch), 256, 0, h->stream>>>(d_t, 1, 1, 0, hm.d_sg.as<double>(), hm.d_sE.as<double>(), ns, d_qp, 128);
  h->n_launch++;
  std::vector<double> part((size_t)nch * 128);
  CK(cudaMemcpyAsync(part.data(), d_qp, part.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  double q = 0.0;
  for (int c = 0; c < nch; ++c) q += part[(size_t)c * 128];
  *out = q;
  return 0;
}

extern "C" int cha_stick_spectrum(cha_handle h, int mol_id, double ncol, double tex, double dv, double source_size,
                                  double dish_size, double* out_freq, double* out_tau, double* out_int, int64_t* n_out) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (mol_id < 0 || mol_id >= kMaxM || !h->mol[mol_id].set) FAIL("molecule not set");
  CK(cudaSetDevice(h->dev));
  HostMol& hm = h->mol[mol_id];
  const int64_t N = (int64_t)hm.nu.size();
  double q_ct = 0.0, Q = 0.0;
  if (q_single(h, mol_id, kCT, &q_ct) || q_single(h, mol_id, tex, &Q)) return 1;
  CK(h->d_sim.ensure((size_t)N * 8 * 5));
  double* d = h->d_sim.as<double>();
  CK(cudaMemcpyAsync(d, hm.nu.data(), N * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d + N, hm.logint.data(), N * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d + 2 * N, hm.elower.data(), N * 8, cudaMemcpyHostToDevice, h->stream));
  stick_spectrum_kernel<<<(unsigned)((N + 127) / 128), 128, 0, h->stream>>>((int)N, d, d + N, d + 2 * N, q_ct, Q, ncol, tex, dv,
                                                                          source_size, dish_size, d + 3 * N, d + 4 * N);
  h->n_launch++;
  std::vector<double> tau(N), in(N);
  CK(cudaMemcpyAsync(tau.data(), d + 3 * N, N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(in.data(), d + 4 * N, N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  // trim_array (functions.py:519-534)
  int64_t i0 = N, i1 = N;
  for (int64_t i = 0; i < N; ++i) if (hm.nu[i] > hm.ll) { i0 = i; break; }
  bool empty = false;
  if (i0 == N) { if (N && hm.nu[N - 1] < hm.ll) empty = true; else i0 = 0; }
  if (!empty) { for (int64_t i = 0; i < N; ++i) if (hm.nu[i] > hm.ul) { i1 = i; break; } } else { i0 = i1 = 0; }
  if (i1 < i0) i1 = i0;
  for (int64_t i = i0; i < i1; ++i) { out_freq[i - i0] = hm.nu[i]; out_tau[i - i0] = tau[i]; out_int[i - i0] = in[i]; }
  if (n_out) *n_out = i1 - i0;
  return 0;
}

extern "C" int cha_make_model(cha_handle h, int64_t n_lines, const double* freqs, const double* taus, int64_t n_chan,
                              const double* x, double vlsr, double dv, double tex, double source_size,
                              double aligned_velocity, double dish_size, double mask_centre, double planck_eps, double* out) {
  if (h && !h->pend.empty() && drain(h)) return 1;
  if (!h) return 1;
  if (n_chan <= 0) return 0;
  CK(cudaSetDevice(h->dev));
  CK(h->d_sim.ensure((size_t)(2 * n_lines + 2 * n_chan) * 8 + 64));
  double* d = h->d_sim.as<double>();
  if (n_lines) {
    CK(cudaMemcpyAsync(d, freqs, n_lines * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + n_lines, taus, n_lines * 8, cudaMemcpyHostToDevice, h->stream));
  }
  double* dx = d + 2 * n_lines; double* dout = dx + n_chan;
  CK(cudaMemcpyAsync(dx, x, n_chan * 8, cudaMemcpyHostToDevice, h->stream));
  make_model_kernel<<<(unsigned)((n_chan + 127) / 128), 128, 0, h->stream>>>((int)n_lines, d, d + n_lines, (int)n_chan, dx, vlsr, dv,
                                                                            tex, source_size, aligned_velocity, dish_size,
                                                                            mask_centre, planck_eps, dout);
  h->n_launch++;
  CK(cudaMemcpyAsync(out, dout, n_chan * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
