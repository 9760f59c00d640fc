// Micro-benchmark: do MUFU (XU pipe) and IMAD.WIDE (Philox multiplies) overlap, or do they share an issue resource?
// Three kernels with identical loop structure: MUFU only, IMAD.WIDE only, both (independent chains).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE, int NF>  // MODE bit 0: 8 MUFU chains, bit 1: 8 IMAD.WIDE chains, NF extra independent FFMAs per iteration
__global__ void k(float* out, int iters, float seed) {
  float v[8], f[16];
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = seed + threadIdx.x * 1e-3f + i; a[i] = threadIdx.x * 7 + i; b[i] = i * 3 + 1; }
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = seed + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE & 1) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(v[i]) : "f"(v[i]));
      if (MODE & 2) {
        const uint64_t p = (uint64_t)a[i] * 0xD2511F53u;
        a[i] = (uint32_t)(p >> 32) ^ b[i];
        b[i] = (uint32_t)p;
      }
    }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i % 16] = fmaf(f[i % 16], 1.0001f, 0.5f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + (float)(a[i] ^ b[i]);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  if (s == 12345.678f) out[0] = s;
}

template <int MODE, int NF>
void run(const char* name, float* d, int sms, double mhz, int threads) {
  const int iters = 4096, blocks = sms * 4;
  k<MODE, NF><<<blocks, threads>>>(d, 16, 1.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE, NF><<<blocks, threads>>>(d, iters, 1.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warp_iters_per_smsp = (double)blocks * threads / 32 * iters / (sms * 4);
  printf("%-34s %8.3f ms  %7.1f cycles per warp-iteration per scheduler\n", name, ms, ms * 1e-3 * mhz * 1e6 / warp_iters_per_smsp);
}

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 128;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mhz = clk / 1000.0;
  float* d; cudaMalloc(&d, 4);
  const int n = pr.multiProcessorCount;
  printf("threads per CTA %d (4 CTAs per SM)\n", threads);
  run<1, 0>("8 MUFU", d, n, mhz, threads);
  run<2, 0>("8 IMAD.WIDE + 8 LOP3", d, n, mhz, threads);
  run<3, 0>("8 MUFU + 8 IMAD.WIDE + 8 LOP3", d, n, mhz, threads);
  run<0, 32>("32 FFMA", d, n, mhz, threads);
  run<1, 32>("8 MUFU + 32 FFMA", d, n, mhz, threads);
  run<2, 32>("8 IMAD.WIDE + 8 LOP3 + 32 FFMA", d, n, mhz, threads);
  run<3, 32>("8 MUFU + 8 IMAD.WIDE + 32 FFMA", d, n, mhz, threads);
  return 0;
}
