// ubench_pipes.cu -- measures the pipe ceilings the fused profile kernel is bounded by on THIS B200:
// MUFU.EX2 (ex2.approx.ftz.f32), FFMA, DFMA and F2F(f32->f64) issue rates, per SM per clock and chip-wide.
// Output: one JSON line (committed under profiles/ as the roofline denominator for the SFU-bound kernel).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

__global__ void k_ex2(float* out, float seed) {
  float v[UNROLL];
  for (int u = 0; u < UNROLL; ++u) v[u] = seed + threadIdx.x * 1e-3f + u;
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[u]));
  float s = 0; for (int u = 0; u < UNROLL; ++u) s += v[u];
  if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma(float* out, float seed) {
  float v[UNROLL];
  for (int u = 0; u < UNROLL; ++u) v[u] = seed + threadIdx.x * 1e-3f + u;
  float a = seed * 0.5f, b = seed * 0.25f;
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[u]) : "f"(a), "f"(b));
  float s = 0; for (int u = 0; u < UNROLL; ++u) s += v[u];
  if (s == 123.456f) out[0] = s;
}
__global__ void k_dfma(float* out, double seed) {
  double v[UNROLL];
  for (int u = 0; u < UNROLL; ++u) v[u] = seed + threadIdx.x * 1e-3 + u;
  double a = seed * 0.5, b = seed * 0.25;
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[u]) : "d"(a), "d"(b));
  double s = 0; for (int u = 0; u < UNROLL; ++u) s += v[u];
  if (s == 123.456) out[0] = (float)s;
}
__global__ void k_cvt(float* out, float seed) {
  float v[UNROLL]; double d[UNROLL];
  for (int u = 0; u < UNROLL; ++u) { v[u] = seed + threadIdx.x * 1e-3f + u; d[u] = 0; }
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[u]) : "f"(v[u])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(v[u]) : "d"(d[u])); }
  float s = 0; for (int u = 0; u < UNROLL; ++u) s += v[u];
  if (s == 123.456f) out[0] = s;
}
// the production inner loop shape: FFMA, FMUL, MUFU.EX2, FFMA per Gaussian term
__global__ void k_gauss(float* out, float seed) {
  float acc[UNROLL], u0 = seed + threadIdx.x * 1e-3f;
  for (int u = 0; u < UNROLL; ++u) acc[u] = 0.f;
  float a = seed * 0.5f, sc = seed * 0.25f, t0 = seed;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float v, e;
      asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(v) : "f"(u0), "f"(a), "f"(sc));
      asm volatile("mul.f32 %0, %1, %1;" : "=f"(v) : "f"(v));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-v));
      asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[u]) : "f"(t0), "f"(e));
    }
    u0 += 1e-3f;
  }
  float s = 0; for (int u = 0; u < UNROLL; ++u) s += acc[u];
  if (s == 123.456f) out[0] = s;
}

// pipe-sharing probes: N_X MUFU.EX2 and N_D DFMA per loop trip, independent chains
template <int NX, int ND, int NF>
__global__ void k_mix(float* out, float seed) {
  float x[8]; double d[8]; float f[8];
  for (int u = 0; u < 8; ++u) { x[u] = seed + threadIdx.x * 1e-3f + u; d[u] = seed + u; f[u] = seed - u; }
  double a = seed * 0.5, b = seed * 0.25; float fa = seed * 0.5f, fb = seed * 0.25f;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int u = 0; u < NX; ++u) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[u]));
#pragma unroll
    for (int u = 0; u < ND; ++u) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[u]) : "d"(a), "d"(b));
#pragma unroll
    for (int u = 0; u < NF; ++u) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[u]) : "f"(fa), "f"(fb));
  }
  float s = 0; for (int u = 0; u < 8; ++u) s += x[u] + (float)d[u] + f[u];
  if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma2(float* out, float seed) {
  unsigned long long v[UNROLL];
  for (int u = 0; u < UNROLL; ++u) { float q = seed + threadIdx.x * 1e-3f + u; asm("mov.b64 %0, {%1, %1};" : "=l"(v[u]) : "f"(q)); }
  unsigned long long a, b; { float fa = seed * 0.5f, fb = seed * 0.25f; asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(fa)); asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(fb)); }
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[u]) : "l"(a), "l"(b));
  unsigned long long s = 0; for (int u = 0; u < UNROLL; ++u) s += v[u];
  if (s == 123456ull) out[0] = 1.f;
}

template <typename F>
double time_ms(F launch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, 4);
  const int blocks = sms * 8, threads = 256;
  const double n = (double)blocks * threads * ITERS * UNROLL;
  double t_ex2 = time_ms([&] { k_ex2<<<blocks, threads>>>(out, 0.5f); });
  double t_ffma = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 0.5f); });
  double t_dfma = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 0.5); });
  double t_cvt = time_ms([&] { k_cvt<<<blocks, threads>>>(out, 0.5f); });
  double t_g = time_ms([&] { k_gauss<<<blocks, threads>>>(out, 0.5f); });
  double t_f2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 0.5f); });
  // per loop trip: (MUFU, DFMA, FFMA) counts; if pipes are independent the time is the max of the parts
  double t_x8 = time_ms([&] { k_mix<8, 0, 0><<<blocks, threads>>>(out, 0.5f); });
  double t_d8 = time_ms([&] { k_mix<0, 8, 0><<<blocks, threads>>>(out, 0.5f); });
  double t_x2d8 = time_ms([&] { k_mix<2, 8, 0><<<blocks, threads>>>(out, 0.5f); });
  double t_x2d8f8 = time_ms([&] { k_mix<2, 8, 8><<<blocks, threads>>>(out, 0.5f); });
  double t_x1d4f8 = time_ms([&] { k_mix<1, 4, 8><<<blocks, threads>>>(out, 0.5f); });
  double t_f8 = time_ms([&] { k_mix<0, 0, 8><<<blocks, threads>>>(out, 0.5f); });
  const double trips = (double)blocks * threads / 32.0 * ITERS / (sms * 4.0);   // warp-trips per SMSP
  const double cyc = clk_khz * 1e3 * 1e-3;                                       // cycles per ms
  printf("{\"mix_cycles_per_warp_trip\": {\"x8\": %.2f, \"d8\": %.2f, \"f8\": %.2f, \"x2d8\": %.2f, \"x2d8f8\": %.2f, \"x1d4f8\": %.2f}, \"ffma2_per_s\": %.4e}\n",
         t_x8 * cyc / trips, t_d8 * cyc / trips, t_f8 * cyc / trips, t_x2d8 * cyc / trips, t_x2d8f8 * cyc / trips, t_x1d4f8 * cyc / trips,
         n / t_f2 * 1e3);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, "
         "\"ex2_per_s\": %.4e, \"ffma_per_s\": %.4e, \"dfma_per_s\": %.4e, \"cvt_pair_per_s\": %.4e, \"gauss_terms_per_s\": %.4e, "
         "\"ex2_per_sm_clk_at_attr_clock\": %.3f, \"ffma_per_sm_clk_at_attr_clock\": %.3f, \"dfma_per_sm_clk_at_attr_clock\": %.3f}\n",
         p.name, sms, clk_khz, n / t_ex2 * 1e3, n / t_ffma * 1e3, n / t_dfma * 1e3, n / t_cvt * 1e3, n / t_g * 1e3,
         n / t_ex2 * 1e3 / sms / (clk_khz * 1e3), n / t_ffma * 1e3 / sms / (clk_khz * 1e3), n / t_dfma * 1e3 / sms / (clk_khz * 1e3));
  return 0;
}
