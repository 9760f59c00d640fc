// Tensor-core noise path of the fused stochastic aggregation: shared pieces (included by spmm.cu; the kernel is in
// spmm_wq.cuh).
//
// STAG_NOISE_NORMAL_HADAMARD: the standard normals of one (edge, sample, 128-channel group) are the
// Walsh-Hadamard mix of 128 independent random bytes,
//     z[c] = kWhInvSd * sum_k H[c,k] * v_k,    H[c,k] = (-1)^popcount(c & k)   (Sylvester, 128 x 128)
//     v_k  = e4m3( (byte_k & 0xCD) | 0x12 )    byte_k = byte k%16 of Philox block 8*group + k/16
// (noise.cuh: wh_* helpers; counter layout (block, eid, sample, offset) as everywhere else).  The masked
// FP8 code is a symmetric scale mixture with excess kurtosis 8e-4, so the sum of 128 of them is normal
// to 5.6e-6 in Kolmogorov distance (exact convolution, oracle/wh_quality.py) -- the same order as the
// 16-bit Box-Muller of STAG_NOISE_NORMAL -- the 128 channels of a group are uncorrelated (H orthogonal)
// with cross-cumulants of the same order, and every partial sum is a multiple of 2^-8 below 2^12: exact
// in fp32, hence bit-identical in the forward pass, the transposed pass and stag_noise_emit.
//
// Why: Box-Muller costs 2 MUFU + ~9 other instructions per normal and pins the streaming kernel to the XU /
// dispatch floor.  Here one byte of Philox output and 128 FP8 MACs on the otherwise idle
// tensor cores make a normal: the random bytes of 128 edges (A operand, one 128-byte row per edge) and H (B operand)
// sit in shared memory in the canonical K-major SWIZZLE_128B layout, one tcgen05.mma.kind::f8f6f4 chain
// (M = 128 edges, N = 128 channels, K = 128) leaves z in TMEM.
//
// Round 1 built two kernel forms around this generator (channel per thread with tcgen05.ld.32x32b; z handed to the
// 16-channels-per-lane stream through a shared-memory scratch): 3.00 and 2.30 ms per launch against 1.68 for
// Box-Muller at the arxiv shape.  Both are gone: spmm_wq.cuh reads z in the consumer's own layout
// (tcgen05.ld.16x256b) and runs the launch in 1.49 ms (profiles/r02_wq_forms.txt).
#pragma once

namespace stag {

__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds32f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// bounded spin: a lost arrival becomes an error, never a hung GPU
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row atoms of 1024 bytes)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace stag
