/* lte_oracle.c -- CPU ORACLE in plain C (test infrastructure, NOT product code).
 *
 * A C restatement of the reference's walker log-probability with the reference's own COST
 * STRUCTURE, so it can serve as the "port" CPU baseline of bench.py:
 *   - per evaluation the optical depth of EVERY catalog line is formed, then trimmed and
 *     indexed (MolSim.run_sim, spectral_simulator/classes.py:336-397; inference.py:141-144)
 *   - per selected line the velocity of EVERY channel is formed and masked
 *     (make_model_numba, inference.py:50-53; TMC1_four_component.py:158-166) : O(K*L*C)
 *   - walkers are farmed out to host threads the way emcee's pool.map farms them to processes
 *     (inference.py:456-459)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
 * It is pinned against the NumPy oracle and the reference-generated goldens (tests/test_oracle.py).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* spectral_simulator/constants.py:2-7 -- same expressions as the reference */
static const double kcm = 0.69503476;
#define CKM (2.998 * 100000.0)
#define CCM (2.998 * 10000000000.0)
#define CM_ (2.998 * 100000000.0)
static const double h_ = 0x1.b85f8c5445f02p-111; /* 6.626 * 10**(-34) as Python evaluates it */
static const double k_ = 0x1.0b1fceca05db0p-76;  /* 1.381 * 10**(-23) */

typedef struct {
  int64_t n;              /* catalog lines */
  const double* nu;       /* MHz */
  const double* logint;
  const double* elower;   /* cm^-1 */
  int q_kind;             /* 0 poly, 1 lin, 2 pow, 3 state sum (same encoding as include/chalte.h) */
  int n_qp;
  double qp[8];
  int64_t n_states;
  const double* sg;       /* 2J+1 */
  const double* sE;
  int64_t i0, i1;         /* trim_array window (functions.py:519-534) */
  int64_t n_sel;          /* selected lines (datagrid[3]) */
  const int64_t* sel;     /* indices into the trimmed list */
  double* aij_gup;        /* precomputed once: classes.py:95-98 (aij*gup; gup cancels) */
} oracle_mol;

typedef struct {
  int ndim, K, M;
  int idx_ss[8], idx_ncol[32], idx_tex, idx_vlsr[8], idx_dv;
  double fixed_ss, dish, al, mc, eps;
  int guard;              /* inference.py:162-164 */
  /* prior */
  int has_prior;
  double lo[64], hi[64], mu[64], sg[64];
  int gauss[64];
  double vmin_sep, vmax_sep;
} oracle_spec;

static double calc_q(const oracle_mol* m, double T) { /* functions.py:136-325 */
  switch (m->q_kind) {
    case 0: { double a = 0, tp = 1; for (int n = 0; n < m->n_qp; ++n) { a += m->qp[n] * tp; tp *= T; } return a; }
    case 1: { double b = m->qp[0] * T + m->qp[1]; return m->qp[3] != 0.0 ? b / m->qp[3] : m->qp[2] * b; }
    case 2: { double v = m->qp[0] * pow(T, m->qp[1]); return m->qp[3] != 0.0 ? v + m->qp[2] : v; }
    default: { double q = 0; for (int64_t s = 0; s < m->n_states; ++s) q += m->sg[s] * exp(-m->sE[s] / (kcm * T)); return q; }
  }
}

/* MolCat precompute, classes.py:90-98 */
int oracle_mol_prepare(oracle_mol* m, double ll, double ul) {
  int64_t N = m->n, i0 = N, i1 = N;
  for (int64_t i = 0; i < N; ++i) if (m->nu[i] > ll) { i0 = i; break; }
  if (i0 == N) { if (N && m->nu[N - 1] < ll) { m->i0 = m->i1 = 0; goto pre; } i0 = 0; }
  for (int64_t i = 0; i < N; ++i) if (m->nu[i] > ul) { i1 = i; break; }
  if (i1 < i0) i1 = i0;
  m->i0 = i0; m->i1 = i1;
pre:;
  double q300 = calc_q(m, 300.0);
  m->aij_gup = (double*)malloc(sizeof(double) * (size_t)(N ? N : 1));
  if (!m->aij_gup) return 1;
  for (int64_t i = 0; i < N; ++i) {
    double el = m->elower[i], f = m->nu[i], eu = el + f / 29979.2458;
    double sijmu = 1.0 / (exp(-(el / 0.695) / 300.0) - exp(-(eu / 0.695) / 300.0)) * (pow(10.0, m->logint[i]) / f) *
                   (1.0 / (4.16231 * 1e-5)) * q300;
    m->aij_gup[i] = 1.16395 * 1e-20 * f * f * f * sijmu;
  }
  return 0;
}
void oracle_mol_free(oracle_mol* m) { free(m->aij_gup); m->aij_gup = NULL; }

static double planck(double x, double T, double eps) { /* inference.py:56-57 */
  return (h_ * x * 1e6 / k_) / (exp((h_ * x * 1e6) / (k_ * T)) - 1.0 + eps);
}

static int within_bounds(const oracle_spec* s, const double* th) { /* inference.py:169-190, TMC1:224-233 */
  for (int p = 0; p < s->ndim; ++p) if (!(s->lo[p] < th[p] && th[p] < s->hi[p])) return 0;
  if (!isnan(s->vmin_sep)) for (int c = 0; c + 1 < s->K; ++c) if (!(th[s->idx_vlsr[c]] < th[s->idx_vlsr[c + 1]] - s->vmin_sep)) return 0;
  if (!isnan(s->vmax_sep)) for (int c = 0; c + 1 < s->K; ++c) if (!(th[s->idx_vlsr[c + 1]] < th[s->idx_vlsr[c]] + s->vmax_sep)) return 0;
  return 1;
}

static double lnprior(const oracle_spec* s, const double* th) { /* inference.py:193-236 */
  if (!within_bounds(s, th)) return -INFINITY;
  double tot = 0;
  for (int p = 0; p < s->ndim; ++p) {
    if (!s->gauss[p]) continue;
    double d = th[p] - s->mu[p];
    tot += log(1.0 / (sqrt(2 * M_PI) * s->sg[p])) - 0.5 * (d * d / (s->sg[p] * s->sg[p]));
  }
  return tot;
}

/* model spectrum for one theta into model[C]; scratch: tau_all (max n over molecules), acc[C] */
static void make_model(const oracle_spec* s, const oracle_mol* mols, int64_t C, const double* x, const double* th,
                       double* model, double* tau_all, double* acc) {
  const double Tex = th[s->idx_tex], dV = th[s->idx_dv];
  memset(model, 0, sizeof(double) * (size_t)C);
  for (int c = 0; c < s->K; ++c) {
    const double vl = th[s->idx_vlsr[c]];
    const double ss = s->idx_ss[c] < 0 ? s->fixed_ss : th[s->idx_ss[c]];
    memset(acc, 0, sizeof(double) * (size_t)C);
    for (int m = 0; m < s->M; ++m) {
      const oracle_mol* mol = &mols[m];
      const double Ncol = th[s->idx_ncol[m * s->K + c]];
      /* MolSim.run_sim over ALL catalog lines: classes.py:347-354 */
      const double Q = calc_q(mol, Tex);
      for (int64_t i = 0; i < mol->n; ++i) {
        double f = mol->nu[i];
        double Nl = Ncol * exp(-mol->elower[i] / (0.695 * Tex)) / Q;
        double lam = CCM / (f * 1e6);
        double num = lam * lam * mol->aij_gup[i] * Nl * (1 - exp(-(h_ * f * 1e6) / (k_ * Tex)));
        double den = 8 * M_PI * (dV * f * 1e6 / CKM);
        tau_all[i] = num / den;
      }
      /* make_model_numba: inference.py:50-53 -- every channel visited for every selected line */
      for (int64_t q = 0; q < mol->n_sel; ++q) {
        const int64_t li = mol->i0 + mol->sel[q];
        const double f = mol->nu[li], t = tau_all[li];
        const double sig = dV / 2.355, hw = dV * 10;
        for (int64_t j = 0; j < C; ++j) {
          double vg = (f - x[j]) / f * CKM + s->al;
          if (fabs(vg - s->al - s->mc) < hw) {
            double z = (vg - vl) / sig;
            acc[j] += t * exp(-0.5 * (z * z));
          }
        }
      }
    }
    for (int64_t j = 0; j < C; ++j) {
      double dJ = planck(x[j], Tex, s->eps) - planck(x[j], 2.7, s->eps);
      double beam = CM_ / (x[j] * 1e6) * 206265 * 1.22 / s->dish;      /* inference.py:35-39 */
      model[j] += dJ * (1 - exp(-acc[j])) * (ss * ss / (beam * beam + ss * ss));
    }
  }
}

/* mode: 0 lnlike, 1 lnprob, 2 lnprior, 3 model spectra (out[nw*C]).  Returns 0 / 1 (alloc failure). */
int oracle_eval(const oracle_spec* s, const oracle_mol* mols, int64_t C, const double* x, const double* y,
                const double* yerr, const double* theta, int64_t nw, int mode, int nthreads, double* out) {
  int64_t nmax = 1;
  for (int m = 0; m < s->M; ++m) if (mols[m].n > nmax) nmax = mols[m].n;
  int fail = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* model = (double*)malloc(sizeof(double) * (size_t)(C ? C : 1));
    double* acc = (double*)malloc(sizeof(double) * (size_t)(C ? C : 1));
    double* tau = (double*)malloc(sizeof(double) * (size_t)nmax);
    if (!model || !acc || !tau) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int64_t w = 0; w < nw; ++w) {
        const double* th = theta + w * s->ndim;
        if (mode == 2) { out[w] = lnprior(s, th); continue; }
        double lp = 0.0;
        if (mode == 1) { lp = lnprior(s, th); if (!isfinite(lp)) { out[w] = -INFINITY; continue; } }
        make_model(s, mols, C, x, th, model, tau, acc);
        if (mode == 3) { memcpy(out + w * C, model, sizeof(double) * (size_t)C); continue; }
        double tot = 0;
        for (int64_t j = 0; j < C; ++j) {                               /* inference.py:157-160 */
          double is2 = 1.0 / (yerr[j] * yerr[j]);
          double r = y[j] - model[j];
          tot += r * r * is2 - log(is2);
        }
        double ll = -0.5 * tot;
        if (s->guard && !isfinite(tot)) ll = -INFINITY;                 /* inference.py:162-164 */
        out[w] = (mode == 1) ? ((s->guard && !isfinite(ll)) ? -INFINITY : lp + ll) : ll;
      }
    }
    free(model); free(acc); free(tau);
  }
  return fail;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
