// Micro-benchmark: per-SM throughput of the MUFU (XU pipe) operations the noise generator uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

static int g_threads = 512;  // threads per CTA (argv[1]); 4 CTAs per SM

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 5) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 7) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 8) {  // packed half2 tanh: two results per instruction
    unsigned int xi = __float_as_uint(x), yi;
    asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
    y = __uint_as_float(yi);
  }
  if (OP == 9) {
    unsigned int xi = __float_as_uint(x), yi;
    asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
    y = __uint_as_float(yi);
  }
  if (OP == 10) y = fmaf(x, 1.0001f, 0.5f);  // FMA pipe reference
  return y;
}

template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;
}

// the Box-Muller mix: LG2, SQRT, COS, SIN issued round-robin (OP 20), or in runs of 8 of one kind (OP 21)
template <int OP>
__global__ void kmix(float* out, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
    if (OP == 20) {
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        v[i] = op<0>(v[i]); v[i + 1] = op<1>(v[i + 1]); v[i + 2] = op<3>(v[i + 2]); v[i + 3] = op<2>(v[i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = op<0>(v[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = op<1>(v[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = op<3>(v[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = op<2>(v[i]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;
}

template <int OP>
void runmix(const char* name, float* d, int sms, double mhz) {
  const int iters = 2048, threads = g_threads, blocks = sms * 4;
  kmix<OP><<<blocks, threads>>>(d, 16, 1.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kmix<OP><<<blocks, threads>>>(d, iters, 1.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * iters * (OP == 20 ? 8 : 32);
  printf("%-14s %8.3f ms  %7.2f Gop/s  %6.2f lanes/clk/SM (at %.0f MHz)\n", name, ms, ops / ms * 1e-6,
         ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
}

template <int OP>
void run(const char* name, float* d, int sms, double mhz) {
  const int iters = 4096, threads = g_threads, blocks = sms * 4;
  k<OP><<<blocks, threads>>>(d, 16, 1.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(d, iters, 1.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * iters * 8;
  printf("%-14s %8.3f ms  %7.2f Gop/s  %6.2f lanes/clk/SM (at %.0f MHz)\n", name, ms, ops / ms * 1e-6,
         ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
}

int main(int argc, char** argv) {
  if (argc > 1) g_threads = atoi(argv[1]);  // threads per CTA, 4 CTAs per SM: 128 -> 4 warps per scheduler
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mhz = clk / 1000.0;
  float* d; cudaMalloc(&d, 4);
  printf("%s, %d SMs, max clock %.0f MHz\n", pr.name, pr.multiProcessorCount, mhz);
  run<0>("lg2", d, pr.multiProcessorCount, mhz);
  run<1>("sqrt", d, pr.multiProcessorCount, mhz);
  run<2>("sin(+fmul.rz)", d, pr.multiProcessorCount, mhz);
  run<3>("cos(+fmul.rz)", d, pr.multiProcessorCount, mhz);
  run<4>("ex2", d, pr.multiProcessorCount, mhz);
  run<5>("rsqrt", d, pr.multiProcessorCount, mhz);
  run<6>("rcp", d, pr.multiProcessorCount, mhz);
  run<7>("tanh", d, pr.multiProcessorCount, mhz);
  run<8>("tanh.f16x2", d, pr.multiProcessorCount, mhz);
  run<9>("ex2.f16x2", d, pr.multiProcessorCount, mhz);
  run<10>("ffma", d, pr.multiProcessorCount, mhz);
  runmix<20>("mix round-robin", d, pr.multiProcessorCount, mhz);
  runmix<21>("mix runs of 8", d, pr.multiProcessorCount, mhz);
  return 0;
}
