// Which (TMEM lane, column) does each register of tcgen05.ld.16x256b.x8 hold, and may the lane base be 16?
// TMEM is filled with lane * 1000 + column by tcgen05.st.32x32b (layout: thread t of warp w <-> lane 32 w + t),
// read back with .16x256b.x8 at lane bases 32 w and 32 w + 16, and compared with the expected mapping
//     reg[4 i + 2 h + b] = (lane base + t / 4 + 8 h, column base + 8 i + 2 (t % 4) + b)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_ld_layout tmem_ld_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(int* bad, float* dump) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem0 = tmem_base_s;
  // fill: lane (32 warp + lane), columns 0..127
  for (int c = 0; c < 128; ++c) {
    const float v = (float)((32 * warp + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem0 + ((uint32_t)(32 * warp) << 16) + (uint32_t)c),
                 "r"(__float_as_uint(v))
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  int nbad = 0;
  for (int hf = 0; hf < 2; ++hf)
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      const uint32_t taddr = tmem0 + ((uint32_t)(32 * warp + 16 * hf) << 16) + (uint32_t)(64 * ch);
      asm volatile(
          "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i)
        for (int h = 0; h < 2; ++h)
          for (int b = 0; b < 2; ++b) {
            const int el = 32 * warp + 16 * hf + lane / 4 + 8 * h, ec = 64 * ch + 8 * i + 2 * (lane % 4) + b;
            const float got = __uint_as_float(v[4 * i + 2 * h + b]);
            if (got != (float)(el * 1000 + ec)) ++nbad;
            if (warp == 1 && hf == 1 && ch == 1) dump[lane * 32 + 4 * i + 2 * h + b] = got;
          }
    }
  atomicAdd(bad, nbad);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(128u) : "memory");
}

int main() {
  int* bad;
  float* dump;
  cudaMalloc(&bad, 4);
  cudaMalloc(&dump, 32 * 32 * 4);
  cudaMemset(bad, 0, 4);
  k<<<1, 128>>>(bad, dump);
  cudaError_t e = cudaDeviceSynchronize();
  int h = -1;
  float hd[32 * 32];
  cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hd, dump, sizeof(hd), cudaMemcpyDeviceToHost);
  printf("cuda: %s   mismatches: %d (0 = the mapping in the header holds, lane base 16 included)\n", cudaGetErrorString(e), h);
  for (int t = 0; t < 6; ++t) {
    printf("warp 1 hf 1 ch 1 thread %d:", t);
    for (int r = 0; r < 8; ++r) printf(" %.0f", hd[t * 32 + r]);
    printf("\n");
  }
  return h != 0 || e != cudaSuccess;
}
