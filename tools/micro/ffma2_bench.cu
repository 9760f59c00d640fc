// Micro-benchmark: packed fp32 FMA (fma.rn.f32x2 = FFMA2 on sm_100) against scalar three-register FFMA, alone and
// next to the MUFU + IMAD.WIDE streams of the hot kernel.  Same loop structure as mix_bench.cu; the FMAs here read
// three registers (the hot kernel's w = cos * rb + A and acc += w * x have no immediate operand).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE, int NF, int PACK>  // MODE bit 0: 8 MUFU, bit 1: 8 IMAD.WIDE; NF fp32 FMAs per iteration, PACK: as NF/2 FFMA2
__global__ void k(float* out, int iters, float seed) {
  float v[8];
  float2 f[8], m[8], c[8];
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = seed + threadIdx.x * 1e-3f + i; a[i] = threadIdx.x * 7 + i; b[i] = i * 3 + 1;
    f[i] = make_float2(seed + i, seed - i);
    m[i] = make_float2(1.0001f + seed * 1e-6f * i, 0.9999f - seed * 1e-6f * i);
    c[i] = make_float2(0.5f * seed + i, 0.25f * seed - i);
  }
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
    for (int i = 0; i < NF / 2; ++i) {
      if (PACK) {
        f[i % 8] = __ffma2_rn(f[i % 8], m[i % 8], c[i % 8]);
      } else {
        f[i % 8].x = fmaf(f[i % 8].x, m[i % 8].x, c[i % 8].x);
        f[i % 8].y = fmaf(f[i % 8].y, m[i % 8].y, c[i % 8].y);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + (float)(a[i] ^ b[i]) + f[i].x + f[i].y;
  if (s == 12345.678f) out[0] = s;
}

template <int MODE, int NF, int PACK>
void run(const char* name, float* d, int sms, double mhz, int threads) {
  const int iters = 4096, blocks = sms * 4;
  k<MODE, NF, PACK><<<blocks, threads>>>(d, 16, 1.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE, NF, PACK><<<blocks, threads>>>(d, iters, 1.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warp_iters_per_smsp = (double)blocks * threads / 32 * iters / (sms * 4);
  printf("%-44s %8.3f ms  %7.1f cycles per warp-iteration per scheduler\n", name, ms, ms * 1e-3 * mhz * 1e6 / warp_iters_per_smsp);
}

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 128;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mhz = clk / 1000.0;
  float* d; cudaMalloc(&d, 4);
  const int n = pr.multiProcessorCount;
  printf("threads per CTA %d (4 CTAs per SM)\n", threads);
  run<0, 32, 0>("32 FFMA (3 registers)", d, n, mhz, threads);
  run<0, 32, 1>("16 FFMA2", d, n, mhz, threads);
  run<1, 32, 0>("8 MUFU + 32 FFMA", d, n, mhz, threads);
  run<1, 32, 1>("8 MUFU + 16 FFMA2", d, n, mhz, threads);
  run<2, 32, 0>("8 IMAD.WIDE + 8 LOP3 + 32 FFMA", d, n, mhz, threads);
  run<2, 32, 1>("8 IMAD.WIDE + 8 LOP3 + 16 FFMA2", d, n, mhz, threads);
  run<3, 32, 0>("8 MUFU + 8 IMAD.WIDE + 8 LOP3 + 32 FFMA", d, n, mhz, threads);
  run<3, 32, 1>("8 MUFU + 8 IMAD.WIDE + 8 LOP3 + 16 FFMA2", d, n, mhz, threads);
  return 0;
}
