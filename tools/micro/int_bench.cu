// Micro-benchmark: per-SM throughput of the integer operations of counter-based generators
// (Philox: 32x32->64 multiplies + 3-input xors; Threefry: adds, rotates, xors).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ void step(uint32_t& a, uint32_t& b, uint32_t k) {
  if (OP == 0) {  // IMAD.WIDE.U32 : both halves used
    const uint64_t p = (uint64_t)a * 0xD2511F53u;
    a = (uint32_t)(p >> 32) ^ b; b = (uint32_t)p;   // + 1 LOP3
  } else if (OP == 1) {  // mul.hi only
    a = __umulhi(a, 0xD2511F53u) ^ b;
  } else if (OP == 2) {  // mul.lo only (IMAD)
    a = a * 0xD2511F53u + b;
  } else if (OP == 3) {  // LOP3 (three-input xor)
    a = a ^ b ^ k;
  } else if (OP == 4) {  // rotate + xor (SHF + LOP3)
    a = __funnelshift_l(a, a, 13) ^ b;
  } else if (OP == 5) {  // add
    a = a + b + k;
  } else if (OP == 6) {  // PRMT
    a = __byte_perm(a, b, 0x7610) + 1;
  } else if (OP == 7) {  // FFMA reference
    a = __float_as_uint(fmaf(__uint_as_float(a), 1.0001f, 0.5f));
  } else if (OP == 8) {  // Philox round pair: 2 IMAD.WIDE + 2 LOP3 on a 4-word state emulation (a,b only)
    const uint64_t p = (uint64_t)a * 0xD2511F53u;
    const uint64_t q = (uint64_t)b * 0xCD9E8D57u;
    a = (uint32_t)(q >> 32) ^ (uint32_t)p ^ k; b = (uint32_t)(p >> 32) ^ (uint32_t)q ^ k;
  }
}

template <int OP>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) step<OP>(a[i], b[i], (uint32_t)it);
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] ^ b[i];
  if (s == 0x12345678u) out[0] = s;
}

template <int OP>
void run(const char* name, uint32_t* d, int sms, double mhz) {
  const int iters = 4096, threads = 512, blocks = sms * 4;
  k<OP><<<blocks, threads>>>(d, 16, 3);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(d, iters, 3);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * iters * 8;
  printf("%-22s %8.3f ms  %6.2f steps/clk/SM (lanes, at %.0f MHz)\n", name, ms, ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double mhz = clk / 1000.0;
  uint32_t* d; cudaMalloc(&d, 4);
  const int n = pr.multiProcessorCount;
  run<0>("imad.wide + lop3", d, n, mhz);
  run<1>("mul.hi + lop3", d, n, mhz);
  run<2>("imad (lo)", d, n, mhz);
  run<3>("lop3 xor3", d, n, mhz);
  run<4>("rotate + xor", d, n, mhz);
  run<5>("iadd3", d, n, mhz);
  run<6>("prmt + iadd", d, n, mhz);
  run<7>("ffma", d, n, mhz);
  run<8>("2 imad.wide + 2 lop3", d, n, mhz);
  return 0;
}
