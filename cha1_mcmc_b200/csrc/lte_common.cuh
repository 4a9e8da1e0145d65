// lte_common.cuh -- numerical contract + small device helpers shared by all kernels.
//
// Constants are the reference's (spectral_simulator/constants.py:2-7), written as the
// exact IEEE-754 doubles its Python *expressions* evaluate to (e.g. 2.998 * 10**10 is
// 29980000000.000004, not 2.998e10).  They are NOT CODATA values and must not be "fixed".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace lte {

// spectral_simulator/constants.py:2-7
constexpr double kKcm = 0x1.63db98979100cp-1;    // kcm = 0.69503476              (functions.py:323)
constexpr double kCkm = 0x1.24c6000000000p+18;   // ckm = 2.998 * 10**5  km/s     (inference.py:51)
constexpr double kCcm = 0x1.bebc9fc000001p+34;   // ccm = 2.998 * 10**10 cm/s     (classes.py:351)
constexpr double kCm  = 0x1.1de95c0000000p+28;   // cm  = 2.998 * 10**8  m/s      (inference.py:35)
constexpr double kH   = 0x1.b85f8c5445f02p-111;  // h   = 6.626 * 10**(-34)       (inference.py:56)
constexpr double kK   = 0x1.0b1fceca05db0p-76;   // k   = 1.381 * 10**(-23)       (inference.py:56)
// literals used on the path
constexpr double kBoltzLit   = 0.695;            // classes.py:95, 349
constexpr double kCT         = 300.0;            // classes.py:19
constexpr double kTbg        = 2.7;              // inference.py:57, classes.py:492
constexpr double kFwhm       = 2.355;            // inference.py:53
constexpr double kBeamConst  = 0x1.eb7da66666666p+17;  // 206265 * 1.22           (inference.py:38)
constexpr double kSijConst   = 0x1.5d28ed37900a0p-15;  // 4.16231 * 10**(-5)      (classes.py:95)
constexpr double kAijConst   = 0x1.b7ba562cada10p-67;  // 1.16395 * 10**(-20)     (classes.py:98)
constexpr double kMHzPerCm   = 29979.2458;       // classes.py:90

constexpr int kMaxK    = 8;    // CHA_MAX_COMPONENTS
constexpr int kMaxM    = 4;    // CHA_MAX_MOLECULES
constexpr int kMaxNdim = 64;   // CHA_MAX_NDIM

// theta layout + telescope, passed by value to every kernel (<= 400 B)
struct ModelDev {
  int ndim, K, M;
  int idx_ss[kMaxK];
  int idx_ncol[kMaxM * kMaxK];
  int idx_tex;
  int idx_vlsr[kMaxK];
  int idx_dv;
  double fixed_ss;
  double dish;
  double al;        // aligned_velocity
  double mc;        // mask centre (0: inference.py:52 ; 5.8: TMC1_four_component.py:160)
  double eps;       // Planck denominator epsilon (1e-10: inference.py:56 ; 0: TMC1:168)
};

// partition-function descriptor of one molecule (functions.py:136-325)
struct QDesc {
  int kind;           // CHA_Q_*
  int n_params;
  double p[8];
  int n_states;       // CHA_Q_SUM only
  const double* g;    // 2J+1
  const double* E;    // cm^-1
};

__host__ __device__ inline double q_analytic(const QDesc& q, double T) {
  switch (q.kind) {
    case 0: {                                   // polynomial, functions.py:139-162
      double acc = 0.0, tp = 1.0;
      for (int n = 0; n < q.n_params; ++n) { acc += q.p[n] * tp; tp *= T; }
      return acc;
    }
    case 1: {                                   // linear rotor fits, functions.py:173-210
      double base = q.p[0] * T + q.p[1];
      if (q.p[3] != 0.0) return base / q.p[3];  // "(a*T+b)/3"
      return q.p[2] * base;                     // "3*(a*T+b)" or plain
    }
    case 2: {                                   // a*T^p (+b), functions.py:214-257
      double v = q.p[0] * pow(T, q.p[1]);
      return (q.p[3] != 0.0) ? v + q.p[2] : v;
    }
    default: return 0.0;
  }
}

// Planck J(x,T) with the reference's epsilon: inference.py:56-57
__host__ __device__ inline double planck_j(double x_mhz, double T, double eps) {
  double hx = kH * x_mhz * 1e6;
  return (hx / kK) / (exp(hx / (kK * T)) - 1.0 + eps);
}

// beam size (arcsec) at x: inference.py:35-38
__host__ __device__ inline double beam_size(double x_mhz, double dish) {
  return kCm / (x_mhz * 1e6) * kBeamConst / dish;
}

}  // namespace lte
