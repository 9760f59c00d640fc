// Throughput of TMEM reads by shape: how many bytes per cycle and SM do tcgen05.ld.32x32b.x32 and
// tcgen05.ld.16x256b.x8 / .x4 deliver (each .x32 / .x8 load hands a warp 4 KB)?  One or two CTAs of 4 warps per SM,
// every warp reads its own TMEM quadrant back to back.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench ldtm_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD32(SHAPE, taddr, v)                                                                                          \
  asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 "                                                                \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),            \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),            \
                 "=r"(v[30]), "=r"(v[31])                                                                              \
               : "r"(taddr)                                                                                            \
               : "memory")

template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, uint32_t* out, long long* cyc) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t0 = tmem_base_s + ((uint32_t)(32 * warp) << 16);
  uint32_t acc = 0;
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[32];
    if (MODE == 0) {          // 32 lanes x 32 columns
      LD32("32x32b.x32", t0 + (uint32_t)((it & 3) * 32), v);
    } else if (MODE == 1) {   // 16 lanes x 64 columns
      LD32("16x256b.x8", t0 + ((uint32_t)(16 * (it & 1)) << 16) + (uint32_t)(((it >> 1) & 1) * 64), v);
    } else {                  // 2 x (16 lanes x 32 columns)
      uint32_t a[16], b[16];
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
                     "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
                   : "r"(t0 + ((uint32_t)(16 * (it & 1)) << 16)) : "memory");
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]),
                     "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
                   : "r"(t0 + ((uint32_t)(16 * (it & 1)) << 16) + 32u) : "memory");
      for (int i = 0; i < 16; ++i) { v[i] = a[i]; v[16 + i] = b[i]; }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
  }
  const long long c1 = clock64();
  out[blockIdx.x * 128 + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(128u) : "memory");
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
  const int iters = 20000, grid = 148 * ctas_per_sm;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, grid * 128 * 4); cudaMalloc(&cyc, grid * 8);
  k<MODE><<<grid, 128>>>(100, out, cyc);
  k<MODE><<<grid, 128>>>(iters, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148 * 4]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
  printf("%-34s %d CTA/SM  %s  %.1f cycles per 4 KB warp load  -> %.1f B / cycle / SM\n", name, ctas_per_sm, cudaGetErrorString(e),
         mean / iters, 4096.0 * 4 * ctas_per_sm / (mean / iters));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int c = 1; c <= 4; c *= 2) {
    run<0>("tcgen05.ld.32x32b.x32", c);
    run<1>("tcgen05.ld.16x256b.x8", c);
    run<2>("2 x tcgen05.ld.16x256b.x4", c);
  }
  return 0;
}
