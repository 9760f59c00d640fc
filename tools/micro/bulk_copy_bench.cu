// Can cp.async.bulk (1-D TMA copies, UBLKCP) carry a ROW GATHER?  Every quad of a warp fetches the row of a random
// neighbour (ROWB bytes, 16-byte aligned) from an L2-resident [R, 128] fp32 matrix into a per-warp shared-memory ring;
// one mbarrier per ring slot (expect_tx = 8 rows), the warp waits for the oldest slot, reads it back with LDS.128 and
// reissues.  Reported: gathered GB/s and copies per SM per 1000 cycles, against the LDG.256 register gather of
// agg_wh_quad_kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_bench bulk_copy_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int ROWB, int NS>
__global__ void __launch_bounds__(128) bulk_gather(const float* __restrict__ x, const int* __restrict__ idx, int iters,
                                                    float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[4][NS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)warp * NS * 8 * ROWB;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[warp][0]);
  if (lane == 0)
    for (int s = 0; s < NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * s) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * 4 + warp;
  const int* my = idx + gw * (int64_t)iters * 8;
  auto issue = [&](int it) {
    const int s = it % NS;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(8u * ROWB) : "memory");
    __syncwarp();
    if (q == 0) {
      const int row = my[it * 8 + quad];
      asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       ring + (uint32_t)(s * 8 + quad) * ROWB),
                   "l"(x + (int64_t)row * 128), "r"((uint32_t)ROWB), "r"(bar0 + 8u * s)
                   : "memory");
    }
  };
  for (int it = 0; it < NS - 1 && it < iters; ++it) issue(it);
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (it + NS - 1 < iters) issue(it + NS - 1);
    const int s = it % NS;
    const uint32_t parity = (uint32_t)(it / NS) & 1u;
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar0 + 8u * s), "r"(parity) : "memory");
    // this lane's share of its quad's row: ROWB / 4 bytes as 16-byte pieces at q * 16 + 64 * m
#pragma unroll
    for (int m = 0; m < ROWB / 64; ++m) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(ring + (uint32_t)(s * 8 + quad) * ROWB + (uint32_t)(q * 16 + 64 * m)));
      acc += v.x + v.y + v.z + v.w;
    }
    __syncwarp();  // the slot may be refilled
  }
  out[blockIdx.x * 128 + threadIdx.x] = acc;
}

// the register gather of the current kernel: 4 x LDG.256 per lane and row, 16 rows (4 rounds) in flight
__global__ void __launch_bounds__(128) ldg_gather(const float* __restrict__ x, const int* __restrict__ idx, int iters,
                                                   float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3;
  const int64_t gw = (int64_t)blockIdx.x * 4 + warp;
  const int* my = idx + gw * (int64_t)iters * 8;
  float acc = 0.f;
  float v[4][32];
  auto load = [&](int it, float* d) {
    const char* src = reinterpret_cast<const char*>(x + (int64_t)my[it * 8 + quad] * 128) + 32 * q;
#pragma unroll
    for (int m = 0; m < 4; ++m)
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(d[8 * m]), "=f"(d[8 * m + 1]), "=f"(d[8 * m + 2]), "=f"(d[8 * m + 3]), "=f"(d[8 * m + 4]),
                     "=f"(d[8 * m + 5]), "=f"(d[8 * m + 6]), "=f"(d[8 * m + 7])
                   : "l"(src + 128 * m));
  };
#pragma unroll
  for (int k = 0; k < 4; ++k) load(k, v[k]);
  for (int it = 0; it < iters; it += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[k][i];
      if (it + 4 + k < iters) load(it + 4 + k, v[k]);
    }
  }
  out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <int ROWB, int NS>
void run(const char* name, const float* x, const int* idx, float* out, int grid, int iters, int ctas_per_sm) {
  const size_t smem = (size_t)4 * NS * 8 * ROWB;
  cudaFuncSetAttribute(bulk_gather<ROWB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  bulk_gather<ROWB, NS><<<grid, 128, smem>>>(x, idx, iters, out);
  cudaEventRecord(a);
  bulk_gather<ROWB, NS><<<grid, 128, smem>>>(x, idx, iters, out);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double copies = (double)grid * 4 * iters * 8;
  printf("%-28s %s  %7.3f ms  %7.1f GB/s  %6.1f copies / SM / 1000 cycles (at 1.9 GHz)  smem %zu KB x %d CTAs/SM\n", name,
         cudaGetErrorString(e), ms, copies * ROWB / ms / 1e6, copies / 148.0 / (ms * 1.9e6) * 1000.0, smem / 1024, ctas_per_sm);
}

int main() {
  const int R = 169343, iters = 2048;
  float* x; int* idx; float* out;
  cudaMalloc(&x, (size_t)R * 128 * 4);
  cudaMemset(x, 0, (size_t)R * 128 * 4);
  const int maxgrid = 148 * 8;
  std::vector<int> h((size_t)maxgrid * 4 * iters * 8);
  srand(1);
  for (auto& v : h) v = (int)(((unsigned)rand() * 2654435761u) % (unsigned)R);
  cudaMalloc(&idx, h.size() * 4);
  cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, (size_t)maxgrid * 128 * 4);
  run<512, 4>("bulk 512 B, ring 4, 2 CTA/SM", x, idx, out, 148 * 2, iters, 2);
  run<512, 4>("bulk 512 B, ring 4, 3 CTA/SM", x, idx, out, 148 * 3, iters, 3);
  run<512, 3>("bulk 512 B, ring 3, 4 CTA/SM", x, idx, out, 148 * 4, iters, 4);
  run<256, 4>("bulk 256 B, ring 4, 4 CTA/SM", x, idx, out, 148 * 4, iters, 4);
  run<256, 6>("bulk 256 B, ring 6, 4 CTA/SM", x, idx, out, 148 * 4, iters, 4);
  run<512, 6>("bulk 512 B, ring 6, 2 CTA/SM", x, idx, out, 148 * 2, iters, 2);
  {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int g = 2; g <= 3; ++g) {
      ldg_gather<<<148 * g, 128>>>(x, idx, iters, out);
      cudaEventRecord(a);
      ldg_gather<<<148 * g, 128>>>(x, idx, iters, out);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, a, b);
      printf("LDG.256 register gather, %d CTA/SM: %s %7.3f ms %7.1f GB/s\n", g, cudaGetErrorString(e), ms,
             (double)148 * g * 4 * iters * 8 * 512 / ms / 1e6);
    }
  }
  return 0;
}
