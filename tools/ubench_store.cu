// ubench_store.cu -- which store idiom reaches cudaMemset's rate on this B200?  2 GiB zero-fill, best of 5.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_st16(double2* p, size_t n) {           // 16 B per thread, 8 trips, CTA covers 32 KB contiguous
  size_t base = (size_t)blockIdx.x * 2048 + threadIdx.x;
#pragma unroll
  for (int it = 0; it < 8; ++it) { size_t i = base + 256 * it; if (i < n) p[i] = make_double2(0.0, 0.0); }
}
__global__ void k_st16_cs(double2* p, size_t n) {
  size_t base = (size_t)blockIdx.x * 2048 + threadIdx.x;
#pragma unroll
  for (int it = 0; it < 8; ++it) { size_t i = base + 256 * it; if (i < n) __stcs(p + i, make_double2(0.0, 0.0)); }
}
__global__ void k_st16_persist(double2* p, size_t n) {   // persistent grid-stride
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_double2(0.0, 0.0);
}
__global__ void k_st32(double4* p, size_t n) {            // 32 B per thread (two 16 B stores or one 256-bit store if ptxas has it)
  size_t base = (size_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
  for (int it = 0; it < 4; ++it) { size_t i = base + 256 * it; if (i < n) p[i] = make_double4(0.0, 0.0, 0.0, 0.0); }
}
// TMA bulk store: one thread per CTA pushes a zeroed 16 KB shared buffer to consecutive 16 KB chunks
__global__ void k_bulk(char* p, size_t bytes, int chunks_per_cta) {
  __shared__ __align__(128) char z[16384];
  for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) reinterpret_cast<int4*>(z)[i] = make_int4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    size_t off = (size_t)blockIdx.x * chunks_per_cta * 16384;
    for (int c = 0; c < chunks_per_cta; ++c, off += 16384) {
      if (off + 16384 <= bytes) {
        unsigned s = (unsigned)__cvta_generic_to_shared(z);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(s), "r"(16384) : "memory");
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F> double best_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  const size_t bytes = 2ull << 30;
  char* p; cudaMalloc(&p, bytes);
  const size_t n16 = bytes / 16, n32 = bytes / 32;
  double t_ms = best_ms([&] { cudaMemsetAsync(p, 0, bytes); });
  double t_16 = best_ms([&] { k_st16<<<(unsigned)((n16 + 2047) / 2048), 256>>>((double2*)p, n16); });
  double t_cs = best_ms([&] { k_st16_cs<<<(unsigned)((n16 + 2047) / 2048), 256>>>((double2*)p, n16); });
  double t_ps = best_ms([&] { k_st16_persist<<<148 * 8, 256>>>((double2*)p, n16); });
  double t_32 = best_ms([&] { k_st32<<<(unsigned)((n32 + 1023) / 1024), 256>>>((double4*)p, n32); });
  double t_b4 = best_ms([&] { k_bulk<<<(unsigned)(bytes / 16384 / 4), 128>>>(p, bytes, 4); });
  double t_b16 = best_ms([&] { k_bulk<<<(unsigned)(bytes / 16384 / 16), 128>>>(p, bytes, 16); });
  auto gbs = [&](double ms) { return bytes / ms * 1e-6; };
  printf("{\"bytes\": %zu, \"GBps\": {\"cudaMemset\": %.0f, \"st16\": %.0f, \"st16_cs\": %.0f, \"st16_persistent\": %.0f, \"st32\": %.0f, \"tma_bulk_4x16K\": %.0f, \"tma_bulk_16x16K\": %.0f}, \"err\": \"%s\"}\n",
         bytes, gbs(t_ms), gbs(t_16), gbs(t_cs), gbs(t_ps), gbs(t_32), gbs(t_b4), gbs(t_b16), cudaGetErrorString(cudaGetLastError()));
  return 0;
}
