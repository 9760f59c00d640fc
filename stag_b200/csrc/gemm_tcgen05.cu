// Dense feature transform  out = act(row_scale * (A @ W) + bias)  on the 5th-generation tensor
// cores (tcgen05.mma, accumulator in TMEM), fp32 in / fp32 out at fp32-level accuracy by the
// 3xTF32 split  A*W ~= A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (hi = tf32(x), lo = tf32(x - hi)).
//
// Replaces  rst = th.matmul(rst, weight)  [+ dst-norm, bias, activation]  of the reference's
// GCN.forward (stag/zoo/gcn.py:97-114) and the fc_neigh / fc_self Linear layers of GraphSAGE
// (stag/zoo/graph_sage.py:74-75,91,107).
//
// One CTA (4 warps) owns a 128-row tile of A and ALL output columns (N <= 256):
//   per K block of 32 floats:  threads load the A tile and the W^T tile with 128-bit loads, split
//   them into hi / lo parts and store them into shared memory in the canonical K-major
//   SWIZZLE_128B layout (rows of 128 bytes, 16-byte chunks XOR-ed with row % 8);  one elected
//   thread issues 4 k-steps x 3 tcgen05.mma (kind::tf32, M = 128, N = Npad, K = 8) accumulating
//   into TMEM;  tcgen05.commit -> mbarrier releases the shared tiles.
//   epilogue: each warp reads its 32 TMEM lanes (tcgen05.ld 32x32b.x32), applies row scale, bias,
//   activation and writes 128-byte row segments.
// Two CTAs are resident per SM (96 KB of shared memory each), which overlaps one CTA's loads with
// the other's MMAs; the kernel is bound by the A read / out write stream, not by the tensor pipe.
#include "common.cuh"

namespace stag {

constexpr int GM = 128;       // rows per CTA tile (UMMA M)
constexpr int GK = 32;        // floats per K block = one 128-byte swizzle row
constexpr int GTHREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (1 for swizzled K-major),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B -> 64), [46,48) version = 1,
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32, fp32 accumulate, A and B K-major
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded spin: a lost arrival must become an error, never a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

struct GemmParams {
  const float* a;
  int64_t lda;
  const float* wt;  // [Nout, K] row-major (W transposed)
  int64_t ldw;
  int64_t M;
  int Nout, K, npad;  // npad: Nout rounded up to a multiple of 16 (UMMA N)
  const float* row_scale;
  const float* bias;
  int act;
  float* out;
  int64_t ldo;
  uint32_t tmem_cols;
};

// store one 16-byte chunk (4 floats) of row r, chunk index ch (0..7) of a [rows][32] tile, split hi/lo
__device__ __forceinline__ void put_split(float* hi, float* lo, int r, int ch, float4 v) {
  const int o = r * GK + ((ch ^ (r & 7)) << 2);
  float4 h, l;
  h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
  l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
  *reinterpret_cast<float4*>(hi + o) = h;
  *reinterpret_cast<float4*>(lo + o) = l;
}

__device__ __forceinline__ float4 load_row4(const float* base, int64_t ld, int64_t r, int64_t nrows, int k, int K,
                                            bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < nrows) {
    const float* p = base + r * ld + k;
    if (vec && k + 3 < K) {
      v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      if (k + 0 < K) v.x = __ldg(p + 0);
      if (k + 1 < K) v.y = __ldg(p + 1);
      if (k + 2 < K) v.z = __ldg(p + 2);
      if (k + 3 < K) v.w = __ldg(p + 3);
    }
  }
  return v;
}

__global__ void __launch_bounds__(GTHREADS) gemm_tcgen05_kernel(const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char gsm_raw[];
  // 1024-byte aligned tiles (SWIZZLE_128B atoms are 8 rows x 128 B)
  unsigned char* gsm = reinterpret_cast<unsigned char*>(((uintptr_t)gsm_raw + 1023) & ~(uintptr_t)1023);
  float* a_hi = reinterpret_cast<float*>(gsm);
  float* a_lo = a_hi + GM * GK;
  float* w_hi = a_lo + GM * GK;
  float* w_lo = w_hi + 256 * GK;
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * GM;
  const bool vec_a = (p.lda % 4 == 0) && (((uintptr_t)p.a & 15) == 0);
  const bool vec_w = (p.ldw % 4 == 0) && (((uintptr_t)p.wt & 15) == 0);
  const bool vec_o = (p.ldo % 4 == 0) && (((uintptr_t)p.out & 15) == 0);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc(p.npad);

  const int nkb = (p.K + GK - 1) / GK;
  uint32_t phase = 0;
  // register staging of the NEXT K block: its global loads are in flight while the tensor core works on
  // the current one (A: 8 x 128-bit per thread; W^T: the first 128 rows, wider tiles load the rest late)
  float4 va[8], vw[8];
  auto load_tiles = [&](int kb) {
    const int k0 = kb * GK;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * GTHREADS;
      va[j] = load_row4(p.a, p.lda, m0 + (i >> 3), p.M, k0 + (i & 7) * 4, p.K, vec_a);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * GTHREADS;
      vw[j] = load_row4(p.wt, p.ldw, i < p.npad * 8 ? (i >> 3) : p.Nout, p.Nout, k0 + (i & 7) * 4, p.K, vec_w);
    }
  };
  load_tiles(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int k0 = kb * GK;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * GTHREADS;
      put_split(a_hi, a_lo, i >> 3, i & 7, va[j]);
      if (i < p.npad * 8) put_split(w_hi, w_lo, i >> 3, i & 7, vw[j]);
    }
    for (int i0 = 8 * GTHREADS; i0 < p.npad * 8; i0 += 8 * GTHREADS) {  // W^T rows 128.. (Nout > 128)
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + tid + j * GTHREADS;
        v[j] = load_row4(p.wt, p.ldw, i < p.npad * 8 ? (i >> 3) : p.Nout, p.Nout, k0 + (i & 7) * 4, p.K, vec_w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + tid + j * GTHREADS;
        if (i < p.npad * 8) put_split(w_hi, w_lo, i >> 3, i & 7, v[j]);
      }
    }
    // generic-proxy writes -> visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0 && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t dah = make_desc(smem_u32(a_hi)), dal = make_desc(smem_u32(a_lo));
      const uint64_t dwh = make_desc(smem_u32(w_hi)), dwl = make_desc(smem_u32(w_lo));
#pragma unroll
      for (int ks = 0; ks < GK / 8; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);  // 8 tf32 = 32 bytes = 2 x 16 B along K
        mma_tf32(tmem_d, dah + adv, dwh + adv, idesc, (kb | ks) != 0);
        mma_tf32(tmem_d, dal + adv, dwh + adv, idesc, 1);
        mma_tf32(tmem_d, dah + adv, dwl + adv, idesc, 1);
      }
      // arrives on the mbarrier when all MMAs issued so far have completed (implies the fence)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mma_bar))
                   : "memory");
    }
    if (kb + 1 < nkb) load_tiles(kb + 1);
    // the shared tiles may be overwritten (and, after the last block, TMEM read) once the MMAs are done
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: warp w owns TMEM lanes (= tile rows) 32w .. 32w+31 -------------------------------------
  // TMEM gives a thread one row x 32 columns; a per-warp shared scratch (pitch 33) transposes that so
  // that every store instruction writes four full 128-byte row segments
  float* scratch = a_hi + warp * (32 * 33);  // the operand tiles are free now
  const int64_t row = m0 + warp * 32 + lane;
  const float rs = (p.row_scale && row < p.M) ? __ldg(p.row_scale + row) : 1.0f;
  for (int n0 = 0; n0 < p.npad; n0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    __syncwarp();  // the previous chunk's readers are done with the scratch
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float y = __uint_as_float(v[j]) * rs;
      if (p.bias && n0 + j < p.Nout) y += __ldg(p.bias + n0 + j);
      if (p.act == 1) y = fmaxf(y, 0.f);
      scratch[lane * 33 + j] = y;
    }
    __syncwarp();
    const int cq = (lane & 7) * 4;
#pragma unroll
    for (int r = 0; r < 32; r += 4) {
      const int rr = r + (lane >> 3);
      const int64_t grow = m0 + warp * 32 + rr;
      if (grow < p.M) {
        const float* sp = scratch + rr * 33 + cq;
        float* o = p.out + grow * p.ldo + n0 + cq;
        if (vec_o && n0 + cq + 3 < p.Nout) {
          *reinterpret_cast<float4*>(o) = make_float4(sp[0], sp[1], sp[2], sp[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + cq + j < p.Nout) o[j] = sp[j];
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace stag

using namespace stag;

extern "C" size_t stag_gemm_workspace_bytes(int64_t M, int32_t Nout, int32_t K) {
  (void)M; (void)Nout; (void)K;
  return 0;
}

extern "C" int stag_gemm_tcgen05(const float* a, int64_t lda, const float* wt, int64_t ldw, int64_t M, int32_t Nout,
                                 int32_t K, const float* row_scale, const float* bias, int act, float* out,
                                 int64_t ldo, void* ws, size_t ws_bytes, void* stream_) {
  (void)ws; (void)ws_bytes;
  cudaStream_t stream = (cudaStream_t)stream_;
  STAG_CHECK_ARG(M >= 0 && Nout > 0 && K > 0, "stag_gemm_tcgen05: bad sizes M=%lld N=%d K=%d", (long long)M, Nout, K);
  if (M == 0) return STAG_OK;
  STAG_CHECK_ARG(a && wt && out, "stag_gemm_tcgen05: null argument");
  STAG_CHECK_ARG(lda >= K && ldw >= K && ldo >= Nout, "stag_gemm_tcgen05: leading dimensions too small");
  STAG_CHECK_ARG(act == 0 || act == 1, "stag_gemm_tcgen05: act must be 0 (none) or 1 (relu)");
  if (Nout > 256) {
    set_error("stag_gemm_tcgen05: Nout=%d > 256 output columns per tile are not supported", Nout);
    return STAG_EUNSUPPORTED;
  }
  GemmParams p;
  p.a = a; p.lda = lda; p.wt = wt; p.ldw = ldw; p.M = M; p.Nout = Nout; p.K = K;
  p.npad = (Nout + 15) / 16 * 16;
  p.row_scale = row_scale; p.bias = bias; p.act = act; p.out = out; p.ldo = ldo;
  uint32_t cols = 32;
  while ((int)cols < p.npad) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)(2 * GM * GK + 2 * 256 * GK) * sizeof(float) + 1024;
  STAG_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = (M + GM - 1) / GM;
  STAG_CHECK_ARG(grid < (1ll << 31), "stag_gemm_tcgen05: too many row tiles");
  gemm_tcgen05_kernel<<<(unsigned)grid, GTHREADS, smem, stream>>>(p);
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}
