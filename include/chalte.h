/* chalte.h -- C-ABI of the B200-native LTE likelihood engine (libchalte.so).
 *
 * Drop-in boundary for ONE path of KahaanGandhi/Cha1-MCMC: the emcee walker
 * log-probability.  The reference has no FFI; its de-facto boundary is emcee's
 * contract log_prob_fn(theta, *args) (inference.py:458-459, 467-468) which, with
 * vectorize=True, is log_prob(theta[nw, ndim]) -> float64[nw].  Every entry point
 * below names the reference code it replaces (paths relative to the reference
 * repository root).
 *
 * Conventions
 *   - plain pointers + sizes only; all arrays row-major, float64 unless noted
 *   - every function returns 0 on success, non-zero on error
 *     (cha_last_error(h) gives the message); no C++ exception crosses the ABI
 *   - numerical failure is NOT an error: the walker's lane gets -inf
 *     (inference.py:140-155, 162-164, 204-205, 241-245), never NaN
 *   - caller owns every host buffer; the library copies into HBM at set_* time and
 *     owns device memory until cha_destroy
 *   - one handle = one GPU + one stream; calls on a handle are serialised by the
 *     caller; different handles may be driven from different threads
 *   - there is NO CPU fallback: without a CUDA device cha_create fails
 */
#ifndef CHALTE_H
#define CHALTE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cha_engine* cha_handle;

#define CHA_MAX_COMPONENTS 8   /* velocity components K            */
#define CHA_MAX_MOLECULES  4   /* molecules sharing one channel grid */
#define CHA_MAX_NDIM       64

/* partition-function kinds: which branch of spectral_simulator/functions.py:136-325
 * the catalog takes (resolved ONCE on the host from the file name, shipped as data) */
enum {
  CHA_Q_POLY = 0,  /* Q = sum_n p[n] * T^n                      functions.py:139-162        */
  CHA_Q_LIN  = 1,  /* Q = (p[0]*T + p[1]) * p[2]  or / p[3]     functions.py:173-210        */
  CHA_Q_POW  = 2,  /* Q = p[0]*T^p[1] (+ p[2] when p[3]!=0)     functions.py:164-168,214-257*/
  CHA_Q_SUM  = 3   /* Q = sum_s g[s]*exp(-E[s]/(kcm*T))         functions.py:263-323        */
};

/* precision of the fused profile + chi-square kernel
 *   CHA_PREC_FP64  every operation in fp64 in the reference's operation order over the full 10 dV masks:
 *                  log-likelihoods agree with the reference to ~1e-12 relative.
 *   CHA_PREC_MIXED (default) frequency offsets and sigma-scaled data formed in fp64, the model spectrum in fp32
 *                  (MUFU.EX2 Gaussians, ~2e-7 of a line peak), residuals in fp32 against the sigma-scaled data
 *                  (rounded to fp32 after the scaling; build option CHA_YS_SPLIT=1 keeps a hi + lo split),
 *                  chi-square accumulated in fp64 per 8-channel group.  Error bound held by the tests
 *                  (tests/helpers.py::check_lnlike), with chi2 the chi-square of the row:
 *                      chi2 <= 4 per channel (any reasonable fit):   |d lnlike| <= 1e-3          (BASELINE tolerance)
 *                      rows far from the data:                       |d lnlike| <= 1e-3 + 5e-7 * chi2
 *                  The second case is what a model formed in fp32 can give when 1e-3/|lnlike| is below the fp32
 *                  epsilon (e.g. the hc11n GOTHAM fixture under the HC9N template, lnlike = -37 281); such rows are
 *                  rejected by any sampler long before the difference matters.  Use CHA_PREC_FP64 where it does.
 *                  Model spectra (cha_simulate) are within 1e-5 of the spectrum peak. */
enum {
  CHA_PREC_FP64  = 0,
  CHA_PREC_MIXED = 1
};

/* ---- lifetime ------------------------------------------------------------------ */
int cha_create(int device_id, cha_handle* out);
int cha_destroy(cha_handle h);
const char* cha_last_error(cha_handle h);          /* h may be NULL: last create error */
int cha_version(void);

/* ---- catalog: replaces MolCat.read_catalog precompute (classes.py:90-98) and the
 * per-call trim (functions.py:507-540) + line selection (inference.py:142-144).
 *   nu/logint/elower : the N catalog lines (MHz, log10 nm^2 MHz, cm^-1), frequency-sorted
 *   q_kind/q_params  : partition function branch; state_g = 2*J+1, state_E (cm^-1) for CHA_Q_SUM
 *   ll, ul           : frequency window (MHz)  -> lines with ll < nu <= ul ... (trim_array rule)
 *   line_idx         : indices into the trimmed list (datagrid[3] / covered_trans); NULL = all
 * The device computes sijmu, aij*gup and the line-strength factor K_i once and keeps them in HBM. */
int cha_set_molecule(cha_handle h, int mol_id, int64_t n_lines,
                     const double* nu, const double* logint, const double* elower,
                     int q_kind, const double* q_params, int n_q_params,
                     int64_t n_states, const double* state_g, const double* state_E,
                     double ll, double ul, const int64_t* line_idx, int64_t n_sel);

/* ---- spectrum: replaces datagrid[0..2] (inference.py:129, 151-160) -------------- */
int cha_set_spectrum(cha_handle h, int64_t n_chan, const double* freq, const double* y, const double* yerr);

/* ---- model: theta layout + telescope (inference.py:133-137, 44-61;
 * scripts/MCMC/TMC1_four_component.py:148-181, 189).
 *   idx_ss[K]        : theta index of each component's source size, or -1 -> fixed_ss
 *   idx_ncol[M*K]    : theta index of column density of molecule m, component c ([m*K+c])
 *   mask_centre      : 0 (inference.py:52: the +aligned -aligned cancels) or 5.8 (TMC1:160)
 *   planck_eps       : 1e-10 (inference.py:56-57) or 0 (TMC1:168-169)                  */
int cha_set_model(cha_handle h, int ndim, int n_comp, int n_mol,
                  const int* idx_ss, const int* idx_ncol, int idx_tex, const int* idx_vlsr, int idx_dv,
                  double fixed_ss, double dish_size, double aligned_velocity,
                  double mask_centre, double planck_eps);

/* ---- prior: replaces is_within_bounds + lnprior (inference.py:169-236;
 * TMC1_four_component.py:224-268).  Strict bounds lo < theta < hi (+-inf = none);
 * Gaussian term on parameter p when gauss[p] != 0 with the EFFECTIVE sigma (the host
 * applies the 0.8/0.3*mean_dV overrides of inference.py:200-201);  vlsr ordering
 * vlsr_c < vlsr_{c+1} - min_sep and vlsr_{c+1} < vlsr_c + max_sep when not NaN.        */
int cha_set_prior(cha_handle h, const double* lo, const double* hi,
                  const double* mu, const double* sigma, const int* gauss,
                  double vlsr_min_sep, double vlsr_max_sep);

int cha_set_precision(cha_handle h, int prec);

/* ---- evaluation (host buffers; transfers inside; synchronous) ---------------------
 * theta is staged in a pinned buffer of the handle and out is complete when the call returns.  A batch
 * that the resident line/channel lists do not cover is evaluated again after the rebuild inside the call.
 * Calls repeated with one batch size (<= 4096 walkers) are replayed as one CUDA graph.
 * cha_log_prob  : lnprob   (inference.py:239-246)  out[nw]
 * cha_log_like  : lnlike   (inference.py:127-166)  out[nw]   (no prior, no bounds)
 * cha_log_prior : lnprior  (inference.py:193-236)  out[nw]
 * cha_simulate  : the model spectrum make_model returns (inference.py:44-61),
 *                 out[nw * n_chan] in the caller's channel order                        */
int cha_log_prob(cha_handle h, const double* theta, int64_t nw, double* out);
int cha_log_like(cha_handle h, const double* theta, int64_t nw, double* out);
int cha_log_prior(cha_handle h, const double* theta, int64_t nw, double* out);
int cha_simulate(cha_handle h, const double* theta, int64_t nw, double* out);

/* ---- evaluation on DEVICE pointers (chains resident in HBM; no copies) -----------
 * with_prior: 1 = lnprob, 0 = lnlike.  Runs on the handle's stream and returns without
 * waiting for the device.  The call is OPTIMISTIC: it is launched against the line/channel
 * lists the handle currently holds; whether they covered the batch (its largest dV and
 * |vlsr - aligned - mask_centre|) is checked at cha_sync -- or at any later host-buffer or
 * configuration call -- and whatever was not covered is re-evaluated there after a rebuild.
 * Therefore d_theta and d_out must stay valid and d_theta unmodified until cha_sync returns,
 * and results may only be consumed (by the host or by other streams) after cha_sync.       */
int cha_log_prob_dev(cha_handle h, const double* d_theta, int64_t nw, double* d_out, int with_prior);
int cha_simulate_dev(cha_handle h, const double* d_theta, int64_t nw, double* d_out);
int cha_sync(cha_handle h);
void* cha_stream(cha_handle h);                     /* cudaStream_t */

/* ---- stand-alone pieces of the path the reference exposes as functions -------------
 * cha_stick_spectrum : MolSim(..., gauss=False).run_sim for ONE component (classes.py:336-397) as called
 *                      by predict_intensities (inference.py:249-253) and init_setup (inference.py:324-327):
 *                      freq_sim / tau_sim / int_sim over the catalog lines inside (ll, ul].
 *                      out_* have room for the whole catalog (n_lines); *n_out receives the trimmed count.
 * cha_make_model     : make_model_numba (inference.py:44-61) on caller-supplied line lists:
 *                      Gaussian splat of taus[L] at freqs[L] onto x[C], Planck, 1-exp(-tau), beam dilution.  */
int cha_stick_spectrum(cha_handle h, int mol_id, double ncol, double tex, double dv,
                       double source_size, double dish_size,
                       double* out_freq, double* out_tau, double* out_int, int64_t* n_out);
int cha_make_model(cha_handle h, int64_t n_lines, const double* freqs, const double* taus,
                   int64_t n_chan, const double* x, double vlsr, double dv, double tex, double source_size,
                   double aligned_velocity, double dish_size, double mask_centre, double planck_eps, double* out);

/* ---- on-device ensemble sampler: replaces emcee.EnsembleSampler.run_mcmc with the
 * StretchMove (call sites inference.py:456-473).  Walkers [w0, w0+nw_local) of a global
 * ensemble of nw_global live on this handle; RNG is counter-based and keyed by
 * (seed, step, global walker id) so results do not depend on the sharding.
 * cha_sampler_half_step does: propose for the local walkers of colour `split` from the
 * complementary set (d_all_coords holds ALL nw_global positions, e.g. after an NCCL
 * all-gather), evaluate lnprob, accept/reject in place.                                */
int cha_sampler_init(cha_handle h, int64_t nw_global, int64_t w0, int64_t nw_local,
                     const double* coords_local, uint64_t seed, double stretch_a);
int cha_sampler_half_step(cha_handle h, int64_t step, int split, const double* d_all_coords);
int cha_sampler_coords_dev(cha_handle h, double** d_coords, double** d_logp);   /* local, resident */
int cha_sampler_get(cha_handle h, double* coords_local, double* logp_local, int64_t* n_accepted);

/* ---- walkers sharded over the GPUs of one box: replaces the multiprocessing.Pool hand-off to emcee
 * (inference.py:456-459, 466-468).  One process (rank) per GPU, one handle per rank.  The only exchange of the
 * complementary-ensemble move is one all-gather of positions per half-step; it is enqueued by the engine on the
 * handle's own stream (NCCL over NVLink/NVSwitch), so a run of steps needs no host synchronisation:
 *   rank 0: cha_comm_unique_id(id)  ->  the caller broadcasts the 128 bytes (torch.distributed, MPI, a file ...)
 *   all   : cha_comm_init(h, rank, world, id)          (collective; NCCL is dlopen'ed here, not at load time)
 *   all   : cha_sampler_init(h, nw_global, w0, nw_local, ...) with EQUAL contiguous shares in rank order
 *   all   : cha_sampler_run(h, step0, n_steps, store_every)      (same call sequence on every rank)
 * cha_sampler_run queues n_steps stretch-move steps (two half-steps each: [all-gather] -> proposals of the WHOLE
 * ensemble sized against the resident line/channel lists -> propose -> lnprob -> accept in place) and returns
 * without waiting for the device.  Every store_every-th step (0 = never) the local positions and log-probs are
 * appended to a chain kept in HBM.  A half-step the lists did not cover leaves the state untouched and marks
 * every later queued half-step void on the device; cha_sync (or any configuration / host-buffer call) rebuilds the
 * lists and re-runs them in order -- on every rank alike, because the coverage decision is taken from the
 * proposals of the whole ensemble.  With one rank and no communicator the resident positions serve in place.
 * cha_sampler_chain_read: slots [slot0, slot0 + n_slots) -> coords[n_slots][nw_local][ndim], logp[n_slots][nw_local]. */
#define CHA_COMM_ID_BYTES 128
int cha_comm_unique_id(unsigned char id[CHA_COMM_ID_BYTES]);
int cha_comm_init(cha_handle h, int rank, int world, const unsigned char id[CHA_COMM_ID_BYTES]);
int cha_comm_destroy(cha_handle h);
int cha_sampler_run(cha_handle h, int64_t step0, int64_t n_steps, int64_t store_every);
int64_t cha_sampler_chain_len(cha_handle h);
int cha_sampler_chain_read(cha_handle h, int64_t slot0, int64_t n_slots, double* coords, double* logp);
int cha_sampler_chain_clear(cha_handle h);

/* ---- introspection (benchmark / roofline bookkeeping) ----------------------------
 * what: 0 #kernel launches so far      1 #selected lines (all molecules)
 *       2 #active channels             3 #line-channel pairs in the device pair list
 *       4 #channel tiles               5 dV the pair list was built for (x1e9, rounded)
 *       6 #pair-list rebuilds          7 last fused-kernel time in ns (CUDA events; plain launches only); after
 *         cha_simulate*: the channel-stream kernel (or zero-fill + tiles) of that call
 *       8 #channel groups              9 #line records of the group tiling
 *      10 half-width of the list (km/s x1e9)   11 host microseconds spent building lists
 *      12 #launch sequences replayed as one CUDA graph (batches of <= 4096 walkers: the sequence
 *         walker_prep -> fused kernel -> finalize and its copies is captured on the second identical call;
 *         kernels inside a replayed graph are counted in stat 0 like plain launches)
 *      13 #collectives enqueued (all-gathers of positions)   14 bytes this rank received in them
 *      15 #queued half-steps that had to be re-run after a list rebuild
 *      16-19 the sampler's narrow list set (bulk of the proposals; the primary lists serve the outliers):
 *         #tiles, #line-channel pairs, #builds, half-width (km/s x1e9); 0 while it is not in use
 *      20 host microseconds spent inside synchronisation points (waiting, validating, re-running)
 *      21 #synchronisation points that had queued calls   22 ... of which found a call the lists had not covered
 *      23 #log-prob batches evaluated in order of reach class (single-component fits, > 4096 walkers: when the rows
 *         of a batch reach very differently far from the line centres -- a spread-out ensemble in arbitrary order --
 *         the engine evaluates them grouped by reach and writes every result back to the caller's row; the values
 *         are bit-identical either way.  CHALTE_SORT_ROWS=0 never, 1 always, unset: adaptive) */
int64_t cha_stat(cha_handle h, int what);
/* exact count of Gaussian evaluations the reference's masks admit for theta[nw]:
 * out[w] = sum_i #{j : |dv_ij - mask_centre| < 10 dV_w}  (inference.py:52)              */
int cha_count_window_pairs(cha_handle h, const double* theta, int64_t nw, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* CHALTE_H */
