// Tensor-core noise path of the fused stochastic aggregation (included by spmm.cu).
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
// tensor cores make a normal: H (A operand) and the random bytes (B operand, one 128-byte row per edge) sit in
// shared memory in the canonical K-major SWIZZLE_128B layout, one tcgen05.mma.kind::f8f6f4 chain
// (M = 128 channels, N = 128 edges, K = 128) leaves z in TMEM with lane = channel and column = edge, and the
// four warps read it back (tcgen05.ld 32x32b) with exactly the ownership the segmented reduction wants:
// a thread owns ONE channel and walks the edges of a row range sequentially -- no cross-lane reduction,
// rows of any length (hubs included) without partial sums, summation in stored-edge order.
//
// Work decomposition: the stored edges are cut at row boundaries into units of about TC_UNIT_EDGES edges
// (tc_bounds_kernel); a CTA walks (unit, sample, channel group) triples as one continuous sequence of
// 128-edge tiles through a software pipeline:
//     tile t+2: edge records {gathered-row byte offset, eid | last << 31, A, B} + rows -> per-warp rings (cp.async)
//     tile t+1: Philox -> masked bytes -> B tile (STS.128), 4 MMAs into the other TMEM buffer (one thread)
//     tile t  : per 8-edge chunk: tcgen05.ld z, gathered rows from the per-warp cp.async ring (each warp
//               fetches exactly the 128-byte quarter rows it consumes), w = A + B z, acc += w x;
//               the row is stored when the stream passes its last edge.
#pragma once

namespace stag {

constexpr int TC_T = 128;        // edges per tile = MMA N (64 with 3 CTAs per SM measured slower: 3.1 vs 3.0 ms)
constexpr int TC_THREADS = 128;  // one warpgroup: warp w owns TMEM lanes (= channels) 32w .. 32w+31
constexpr int TC_CE = 8;         // edges per gather chunk
constexpr int TC_NCH = TC_T / TC_CE;
constexpr int TC_RING = 8;       // gather ring slots (chunks) per warp; the chunk loop is unrolled by it
constexpr int TC_LOOK = 6;       // chunks in flight
constexpr int TC_CTAS = 2;       // CTAs per SM (shared memory; TMEM: 2 * TC_T columns each)
constexpr int TC_UNIT_EDGES = 2048;
constexpr uint32_t TC_H_BYTES = 128 * 128, TC_B_BYTES = TC_T * 128;
constexpr uint32_t TC_XRING = TC_RING * TC_CE * 128;  // per warp: 128 bytes (32 channels) per edge
constexpr uint32_t TC_METARING = 3 * TC_T * 8;        // {gathered-row byte offset, eid | last << 31}
constexpr uint32_t TC_ABRING = 3 * TC_T * 8;          // {A, B}
constexpr uint32_t TC_ROWRING = 3 * TC_T * 4;
constexpr uint32_t TC_WARP_BYTES = TC_XRING + TC_METARING + TC_ABRING + TC_ROWRING;
constexpr uint32_t TC_SMEM = TC_H_BYTES + 2 * TC_B_BYTES + 4 * TC_WARP_BYTES + 1024;
static_assert(TC_NCH % TC_RING == 0, "the chunk loop is unrolled by the ring size");

// edge boundaries of the units: ebnd[k] = indptr[first row whose edges start at or after k * TC_UNIT_EDGES]
__global__ void tc_bounds_kernel(const int32_t* __restrict__ indptr, int N, int64_t E, int nunits,
                                 int32_t* __restrict__ ebnd) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nunits) return;
  if (k == nunits) { ebnd[k] = (int32_t)E; return; }
  const int64_t target = (int64_t)k * TC_UNIT_EDGES;
  int lo = 0, hi = N;  // smallest r in [0, N] with indptr[r] >= target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(indptr + mid) >= target) hi = mid; else lo = mid + 1;
  }
  ebnd[k] = __ldg(indptr + lo);
}

// per-call edge records: both degree scalings, the distribution parameters and the variance constant of the
// Hadamard mix folded into (A, B): w * scale = A + B * (raw sum);  .x = byte offset of the gathered row
__global__ void tc_record_kernel(const AggParams p, int4* __restrict__ rec, uint32_t* __restrict__ rowoff) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.E) return;
  const int idx = __ldg(p.indices + j);
  const int ef = __ldg(p.eidf + j);
  const int row = __ldg(p.erow + j);
  float sc = p.gscale ? __ldg(p.gscale + idx) : 1.0f;
  if (p.rscale) sc *= __ldg(p.rscale + row);
  const int64_t pi = p.pshape >= STAG_PARAM_EDGE ? (ef & 0x7fffffff) : 0;
  const float a = sc * __ldg(p.p0 + pi);
  const float b = sc * __ldg(p.p1 + pi) * kWhInvSd;
  rec[j] = make_int4((int)((uint32_t)idx * (uint32_t)p.ldx * 4u), ef, __float_as_int(a), __float_as_int(b));
  rowoff[j] = (uint32_t)row * (uint32_t)p.ldo * 4u;  // byte offset of the output row
}

struct TcTile {
  int e0, n;          // first stored edge, edges in the tile (0: no tile)
  uint32_t smp, blk;  // Philox sample index, this thread's Philox block index
  const char* xb;     // this thread's 16-byte piece of the gathered rows: x + s*x_ss + group + warp quarter
  char* ob;           // this thread's channel of the output rows
};

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

__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src, bool ignore) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tcp.async.ca.shared.global [%0], [%1], 8, p;\n\t}" ::"r"(dst_smem),
      "l"(src), "r"((int)ignore)
      : "memory");
}
__device__ __forceinline__ void tc_ldtm8_issue(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// the registers of every tcgen05.ld issued so far are valid after this (they are operands so that no use is
// scheduled above it)
__device__ __forceinline__ void tc_ldtm8_wait(uint32_t (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
               :
               : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, TC_CTAS) agg_tc_kernel(const AggParams p) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm0 = ((uint32_t)__cvta_generic_to_shared(tc_raw) + 1023u) & ~1023u;
  const uint32_t h_s = sm0, bt_s = sm0 + TC_H_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t warp_s = bt_s + 2 * TC_B_BYTES + (uint32_t)warp * TC_WARP_BYTES;
  const uint32_t xring_s = warp_s, meta_s = warp_s + TC_XRING, ab_s = meta_s + TC_METARING, row_s = ab_s + TC_ABRING;
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(mma_bar);

  // ---- one-off setup: TMEM, barriers, H, cleared byte tiles (stale bytes must stay finite FP8 codes) ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "r"(2u * TC_T)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    tc_mbar_init(bar_s, 1);
    tc_mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // row c of H: 8 chunks of 16 bytes, +1 = 0x38, -1 = 0xB8 (e4m3)
    const int c = tid;
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t wv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = ch * 16 + q * 4 + b;
          word |= ((__popc(c & k) & 1) ? 0xB8u : 0x38u) << (8 * b);
        }
        wv[q] = word;
      }
      sts128(h_s + (uint32_t)c * 128u + (uint32_t)((ch ^ (c & 7)) << 4), make_uint4(wv[0], wv[1], wv[2], wv[3]));
    }
    for (uint32_t o = (uint32_t)tid * 16u; o < 2 * TC_B_BYTES; o += TC_THREADS * 16u)
      sts128(bt_s + o, make_uint4(0u, 0u, 0u, 0u));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem0 = tmem_base_s;
  // kind::f8f6f4, A = B = E4M3 (format 0), fp32 accumulate, both K-major, N = 128, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_T >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t h_desc = tc_desc(h_s);

  // ---- tile cursor (uniform across the CTA) ------------------------------------------------------------
  const int G = p.D >> 7;
  const int64_t total = (int64_t)p.nunits * p.S * G;
  int64_t u = (int64_t)blockIdx.x - (int64_t)gridDim.x;
  int ce = 0, ce_end = 0;
  uint32_t c_smp = 0, c_blk = 0;
  const char* c_xb = reinterpret_cast<const char*>(p.x);
  char* c_ob = reinterpret_cast<char*>(p.out);
  bool drained = false;
  auto next_tile = [&]() {
    TcTile t;
    t.e0 = 0; t.n = 0; t.smp = 0; t.blk = 0; t.xb = reinterpret_cast<const char*>(p.x); t.ob = reinterpret_cast<char*>(p.out);
    while (!drained && ce >= ce_end) {
      u += gridDim.x;
      if (u >= total) { drained = true; break; }
      const int k = (int)(u % p.nunits);
      const int sg = (int)(u / p.nunits);
      const int g = sg % G, s = sg / G;
      ce = __ldg(p.ebnd + k);
      ce_end = __ldg(p.ebnd + k + 1);
      c_smp = (uint32_t)(p.sample_base + s);
      c_blk = (uint32_t)(8 * g + (lane & 7));
      c_xb = reinterpret_cast<const char*>(p.x + (int64_t)s * p.x_ss + g * 128 + warp * 32) + (lane & 7) * 16;
      c_ob = reinterpret_cast<char*>(p.out + (int64_t)s * p.out_ss + g * 128 + warp * 32 + lane);
    }
    if (drained) return t;
    t.e0 = ce;
    t.n = min(TC_T, ce_end - ce);
    t.smp = c_smp; t.blk = c_blk; t.xb = c_xb; t.ob = c_ob;
    ce += TC_T;
    return t;
  };

  // records + rows of a tile -> ring slot (zeros past the end of the tile); joins the next committed group
  auto issue_rec = [&](const TcTile& t, int slot) {
#pragma unroll
    for (int q = 0; q < TC_T / 32; ++q) {
      const int i = lane + 32 * q;
      const bool off = i >= t.n;
      const char* r = reinterpret_cast<const char*>(p.rec + (off ? 0 : t.e0 + i));
      cp_async8(meta_s + (uint32_t)(slot * TC_T + i) * 8u, r, off);
      cp_async8(ab_s + (uint32_t)(slot * TC_T + i) * 8u, r + 8, off);
      cp_async4(row_s + (uint32_t)(slot * TC_T + i) * 4u, p.rowoff + (off ? 0 : t.e0 + i), off);
    }
  };
  // random bytes of a tile: row n of the B operand = the 128 bytes of edge n (8 Philox blocks); this thread makes
  // block lane & 7 of rows 16 i + 4 warp + lane / 8, so that a warp store covers four whole 128-byte rows
  auto make_bytes = [&](const TcTile& t, uint32_t bt, uint32_t metas) {
#pragma unroll 2
    for (int i = 0; i < TC_T / 16; ++i) {
      const int nr = i * 16 + warp * 4 + (lane >> 3);
      if (nr < t.n) {
        const uint32_t eid = (uint32_t)lds32(metas + (uint32_t)nr * 8u + 4u) & 0x7fffffffu;
        uint4 r = philox_rk(t.blk, eid, t.smp, p.key.c3, p);
        r.x = (r.x & kWhAnd) | kWhOr; r.y = (r.y & kWhAnd) | kWhOr;
        r.z = (r.z & kWhAnd) | kWhOr; r.w = (r.w & kWhAnd) | kWhOr;
        sts128(bt + (uint32_t)nr * 128u + (uint32_t)(((lane & 7) ^ (nr & 7)) << 4), r);
      }
    }
  };
  auto issue_mma = [&](int buf) {
    const uint64_t b_desc = tc_desc(bt_s + (uint32_t)buf * TC_B_BYTES);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)  // K = 32 bytes per instruction: + 2 in the (address >> 4) field
      tc_mma_f8(tmem0 + (uint32_t)(buf * TC_T), h_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, ks != 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s + 8u * buf)
                 : "memory");
  };
  // gathered rows of the 8 edges starting at edge `e8` of a tile -> ring slot; a lane copies 16 bytes of
  // edges e8 + lane/8 and e8 + 4 + lane/8
  auto gather = [&](int n, const char* xb, uint32_t metas, int e8, int slot) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ee = e8 + i * 4 + (lane >> 3);
      const uint32_t off = (uint32_t)lds32(metas + (uint32_t)ee * 8u);
      cp_async16(xring_s + (uint32_t)((slot * TC_CE + i * 4 + (lane >> 3)) * 128 + (lane & 7) * 16), xb + off, ee >= n);
    }
  };

  // ---- prologue: the records of the first tile; the loop below starts one tile early (nothing to consume yet), so
  // that the byte tile, the MMAs and the first gathers of tile 0 are made by the same code as every other tile's
  TcTile d0, d1, d2 = next_tile();
  d1 = d2; d1.n = 0;
  issue_rec(d2, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

  float acc = 0.f;
  for (int tau = -1; tau < 0 || d1.n > 0; ++tau) {
    d0 = d1; d1 = d2; d2 = next_tile();
    const int s0 = (tau + 3) % 3, s1 = (tau + 1) % 3, s2 = (tau + 2) % 3;
    const uint32_t meta0 = meta_s + (uint32_t)(s0 * TC_T) * 8u, meta1 = meta_s + (uint32_t)(s1 * TC_T) * 8u;
    const uint32_t ab0 = ab_s + (uint32_t)(s0 * TC_T) * 8u;
    const uint32_t row0 = row_s + (uint32_t)(s0 * TC_T) * 4u;
    // every lane is done with the slot tile tau - 1 lived in, and sees the records of tiles tau and tau + 1 (their
    // copies joined gather groups that were waited for during the previous tile)
    __syncwarp();
    issue_rec(d2, s2);  // committed with the first gather group below
    // last-edge-of-row flags of this tile, one bit per edge: the same four words in every lane
    uint32_t fmask[TC_T / 32];
#pragma unroll
    for (int q = 0; q < TC_T / 32; ++q)
      fmask[q] = __ballot_sync(0xffffffffu, lds32(meta0 + (uint32_t)(lane + 32 * q) * 8u + 4u) < 0);
    const int nb = (tau + 1) & 1;
    if (d1.n > 0) make_bytes(d1, bt_s + (uint32_t)nb * TC_B_BYTES, meta1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // bytes of tile tau + 1 written; TMEM buffer nb read out by every warp (tile tau - 1)
    if (tid == 0 && d1.n > 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_mma(nb);
    }
    if (d0.n > 0) tc_mbar_wait(bar_s + 8u * (tau & 1), (uint32_t)(tau >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem0 + ((uint32_t)(warp * 32) << 16) + (uint32_t)((tau & 1) * TC_T);
    const uint32_t xs = xring_s + (uint32_t)lane * 4u;

#pragma unroll 1
    for (int h = 0; h < TC_NCH / TC_RING; ++h) {
      // the tile the chunks requested from the second half on belong to
      const bool last_h = h == TC_NCH / TC_RING - 1;
      const int la_n = last_h ? d1.n : d0.n;
      const char* la_xb = last_h ? d1.xb : d0.xb;
      const uint32_t la_meta = last_h ? meta1 : meta0;
#pragma unroll
      for (int c = 0; c < TC_RING; ++c) {
        const int jb = (h * TC_RING + c) * TC_CE;  // first edge of the chunk
        // this chunk has landed: the TC_LOOK - 1 chunks requested after it may still be in flight
        cp_async_wait<TC_LOOK - 1>();
        __syncwarp();
        if (jb < d0.n) {
          uint32_t z[TC_CE];
          tc_ldtm8_issue(taddr + (uint32_t)jb, z);
          float xv[TC_CE];
#pragma unroll
          for (int e = 0; e < TC_CE; ++e) xv[e] = lds32f(xs + (uint32_t)((c * TC_CE + e) * 128));
          float4 ab[TC_CE / 2];
#pragma unroll
          for (int e = 0; e < TC_CE / 2; ++e) ab[e] = lds128f(ab0 + (uint32_t)(jb + 2 * e) * 8u);
          const uint32_t fm = ((h ? fmask[2 + (c >> 2)] : fmask[c >> 2]) >> ((c & 3) * 8)) & 0xFFu;
          tc_ldtm8_wait(z);
          float w[TC_CE];
#pragma unroll
          for (int e = 0; e < TC_CE; ++e) {
            const float A = (e & 1) ? ab[e >> 1].z : ab[e >> 1].x, B = (e & 1) ? ab[e >> 1].w : ab[e >> 1].y;
            w[e] = fmaf(__uint_as_float(z[e]), B, A);
          }
          if (fm == 0u) {  // no row ends inside the chunk
#pragma unroll
            for (int e = 0; e < TC_CE; ++e) acc = fmaf(w[e], xv[e], acc);
          } else {
#pragma unroll
            for (int e = 0; e < TC_CE; ++e) {
              acc = fmaf(w[e], xv[e], acc);
              if (fm & (1u << e)) {  // last edge of a row: write it, start the next one
                const uint32_t ro = (uint32_t)lds32(row0 + (uint32_t)(jb + e) * 4u);
                __stcs(reinterpret_cast<float*>(d0.ob + ro), acc);
                acc = 0.f;
              }
            }
          }
        }
        // request the chunk TC_LOOK ahead into the slot consumed TC_RING - TC_LOOK chunks ago
        if (c + TC_LOOK < TC_RING)  // same half
          gather(d0.n, d0.xb, meta0, (h * TC_RING + c + TC_LOOK) * TC_CE, c + TC_LOOK);
        else  // next half: of this tile, or the first half of the next tile
          gather(la_n, la_xb, la_meta, ((last_h ? 0 : (h + 1) * TC_RING) + c + TC_LOOK - TC_RING) * TC_CE,
                 (c + TC_LOOK) % TC_RING);
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();
  __syncwarp();
  zero_empty_rows_tail(p, (int64_t)blockIdx.x * (TC_THREADS / 32) + warp, (int64_t)gridDim.x * (TC_THREADS / 32), lane);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(2u * TC_T) : "memory");
}

// ---- second form: the tensor cores feed the 16-channels-per-lane stream -------------------------------------------
// Same generator, the layout of agg_stream_kernel: a group of 8 lanes walks the edges of one stream item, a lane owns 16
// channels (quads at 4 sl + 32 q) of every edge, so the per-edge costs are paid once per (edge, 16 channels).  The MMA is
// turned around for it: M = 128 TILE ROWS = (16 groups of the CTA) x (8 consecutive edges of each group's item), A = the random
// bytes of those edges, B = H, so TMEM holds lane = (group, edge), column = channel.  The lane that generated the bytes of an
// edge reads that edge's z row back (tcgen05.ld 32x32b.x32, 32 channels per pass) and hands the quads to the 8 lanes of its
// own group through a swizzled per-warp scratch (8 STS.128 + 8 LDS.128 per pass): everything after the MMA is warp-local.
//   round r (8 edges per group):  wait MMA(r) | bytes of round r+1 -> A tile, CTA barrier, MMAs into the other TMEM buffer |
//   4 passes q (channels 32q .. 32q+31): gathers of pass +2 (cp.async ring of 3 pass slots), z quads through the scratch,
//   8 edges x (w = A + B z, acc_q += w x, row end -> 16-byte store of quad q)
constexpr int WH_THREADS = 128, WH_WARPS = 4;
constexpr int WH_NBUF = 4;                                        // record chunks (of 8 edges) per group ring
constexpr uint32_t WH_PASS_BYTES = 4 * 8 * 8 * 16;                // gathered quads of one pass: 4 groups x 8 edges x 8 lanes
constexpr uint32_t WH_XRING = 3 * WH_PASS_BYTES;
constexpr uint32_t WH_REC_BYTES = 4 * (WH_NBUF * 8 * 16 + 16);    // 4 groups, skewed by 16 bytes
constexpr uint32_t WH_ROW_BYTES = 4 * (WH_NBUF * 8 * 4 + 4) + 112;
constexpr uint32_t WH_SCRATCH = 32 * 128;
constexpr uint32_t WH_WARP_BYTES = WH_XRING + WH_REC_BYTES + WH_ROW_BYTES + WH_SCRATCH;
constexpr uint32_t WH_SMEM = 2 * 128 * 128 + WH_WARPS * WH_WARP_BYTES + 1024;
static_assert(WH_WARP_BYTES % 16 == 0, "16-byte aligned rings");

__device__ __forceinline__ void tc_ldtm32(uint32_t taddr, uint32_t (&v)[32]) {
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
}

__global__ void __launch_bounds__(WH_THREADS, 2) agg_wh_stream_kernel(const AggParams p) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int maxn_s[WH_WARPS];
  const uint32_t sm0 = ((uint32_t)__cvta_generic_to_shared(tc_raw) + 1023u) & ~1023u;
  const uint32_t h_s = sm0, at_s = sm0 + 128 * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane >> 3, sl = lane & 7;
  const uint32_t warp_s = at_s + 128 * 128 + (uint32_t)warp * WH_WARP_BYTES;
  const uint32_t xring_s = warp_s;
  const uint32_t rec_g = warp_s + WH_XRING + (uint32_t)sub * (WH_NBUF * 8 * 16 + 16);
  const uint32_t row_g = warp_s + WH_XRING + WH_REC_BYTES + (uint32_t)sub * (WH_NBUF * 8 * 4 + 4);
  const uint32_t scr_s = warp_s + WH_XRING + WH_REC_BYTES + WH_ROW_BYTES;
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(mma_bar);
  constexpr uint32_t rmask = WH_NBUF * 8 - 1;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    tc_mbar_init(bar_s, 1);
    tc_mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const int c = tid;  // row c of H (B operand: N = channel rows, K-major)
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t wv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = ch * 16 + q * 4 + b;
          word |= ((__popc(c & k) & 1) ? 0xB8u : 0x38u) << (8 * b);
        }
        wv[q] = word;
      }
      sts128(h_s + (uint32_t)c * 128u + (uint32_t)((ch ^ (c & 7)) << 4), make_uint4(wv[0], wv[1], wv[2], wv[3]));
    }
    for (uint32_t o = (uint32_t)tid * 16u; o < 128u * 128u; o += WH_THREADS * 16u) sts128(at_s + o, make_uint4(0u, 0u, 0u, 0u));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem0 = tmem_base_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t h_desc = tc_desc(h_s), a_desc = tc_desc(at_s);
  const uint32_t trow = (uint32_t)(32 * warp + lane);  // this lane's tile row = (group 4 warp + sub, edge sl of the round)
  const uint32_t at_row = at_s + trow * 128u;
  uint32_t mma_count = 0;  // MMAs issued so far by this CTA (all threads count): buffer = count & 1, parity = (count >> 1) & 1

  const int G = p.D >> 7;
  const int n_items = p.num_hub_segs + p.num_items;
  const int IG = (n_items + 15) / 16;  // CTA items per (sample, channel group)
  const int64_t total = (int64_t)IG * p.S * G;
  const uint32_t ldxb = (uint32_t)p.ldx * 4u, ldo4 = (uint32_t)p.ldo * 4u;

  for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int64_t outer = item / IG;
    const int gi = (int)(item - outer * IG) * 16 + 4 * warp + sub;
    const int cg = (int)(outer % G), s = (int)(outer / G);
    int e0 = 0, e1 = 0, part_slot = -1;
    if (gi < p.num_hub_segs) {
      int lo = 0, hi = p.num_hubs;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p.hub_seg_ptr + mid) <= gi) lo = mid; else hi = mid;
      }
      const int row = __ldg(p.hub_rows + lo);
      const int k = gi - __ldg(p.hub_seg_ptr + lo);
      e0 = __ldg(p.indptr + row) + k * kHubSegment;
      e1 = min(e0 + kHubSegment, __ldg(p.indptr + row + 1));
      part_slot = gi;
    } else if (gi < n_items) {
      const int4 it = __ldg(reinterpret_cast<const int4*>(p.items) + (gi - p.num_hub_segs));
      e0 = it.z;
      e1 = it.w >= 0 ? it.w : it.z;
    }
    const int nedges = e1 - e0;
    const int rowlim = part_slot < 0 ? nedges : 0;
    int maxn = nedges;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxn = max(maxn, __shfl_xor_sync(0xffffffffu, maxn, o));
    __syncthreads();  // the previous item's readers of maxn_s are done
    if (lane == 0) maxn_s[warp] = maxn;
    __syncthreads();
    maxn = max(max(maxn_s[0], maxn_s[1]), max(maxn_s[2], maxn_s[3]));
    const int rounds = (maxn + 7) >> 3;
    if (rounds == 0) continue;  // uniform: every MMA issued below is waited for exactly once, in order

    const char* xcb = reinterpret_cast<const char*>(p.x + (int64_t)s * p.x_ss + cg * 128 + 4 * sl);
    char* outs = reinterpret_cast<char*>(p.out + (int64_t)s * p.out_ss + cg * 128 + 4 * sl);
    const uint32_t smp = (uint32_t)(p.sample_base + s);
    const int4* recp = p.rec + e0;
    const int32_t* rowp = p.erow + e0;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;

    auto fetch_chunk = [&](int first) {
      const int e = first + sl;
      const uint32_t pos = (uint32_t)e & rmask;
      const bool off = e >= nedges;
      cp_async16(rec_g + pos * 16u, recp + (off ? 0 : e), off);
      cp_async4(row_g + pos * 4u, rowp + (off ? 0 : e), off);
    };
    // random bytes of round r: this lane's edge 8 r + sl of its group -> tile row trow
    auto make_bytes = [&](int r) {
      const int t = 8 * r + sl;
      if (t < nedges) {
        const uint32_t eid = (uint32_t)lds32(rec_g + (((uint32_t)t & rmask) << 4) + 4u) & 0x7fffffffu;
#pragma unroll 2
        for (int j = 0; j < 8; ++j) {
          uint4 v = philox_rk((uint32_t)(8 * cg + j), eid, smp, p.key.c3, p);
          v.x = (v.x & kWhAnd) | kWhOr; v.y = (v.y & kWhAnd) | kWhOr;
          v.z = (v.z & kWhAnd) | kWhOr; v.w = (v.w & kWhAnd) | kWhOr;
          sts128(at_row + (uint32_t)((j ^ (int)(trow & 7u)) << 4), v);
        }
      }
    };
    auto issue_mma = [&]() {  // thread 0
      const uint32_t buf = mma_count & 1u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tc_mma_f8(tmem0 + buf * 128u, a_desc + (uint64_t)(2 * ks), h_desc + (uint64_t)(2 * ks), idesc, ks != 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s + 8u * buf)
                   : "memory");
    };
    // gathered quads of pass q of the round whose neighbour addresses are in xa: slot <- 8 edges x 16 bytes per lane
    const char* xa[8];
    auto round_addresses = [&](int r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t u = (uint32_t)lds32(rec_g + (((uint32_t)(8 * r + j) & rmask) << 4));
        xa[j] = xcb + (uint64_t)u * ldxb;
      }
    };
    auto gather_pass = [&](int r, int q, uint32_t slot_s) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        cp_async16(slot_s + (uint32_t)(((sub * 8 + j) * 8 + sl) * 16), xa[j] + 128 * q, 8 * r + j >= nedges);
    };

    // ---- prologue of the item: records of rounds 0 and 1, bytes + MMAs of round 0, gathers of passes 0 and 1 ----
    fetch_chunk(0);
    fetch_chunk(8);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    make_bytes(0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_mma();
    }
    ++mma_count;
    uint32_t slot = 0;  // ring slot of the pass being consumed; passes +1, +2 live in the next two slots
    round_addresses(0);
    gather_pass(0, 0, xring_s);
    cp_async_commit();
    gather_pass(0, 1, xring_s + WH_PASS_BYTES);
    cp_async_commit();

    for (int r = 0; r < rounds; ++r) {
      fetch_chunk(8 * (r + 2));  // joins the first gather group of this round
      // z of this round is in TMEM buffer (mma_count - 1) & 1
      const uint32_t done = mma_count - 1u;
      tc_mbar_wait(bar_s + 8u * (done & 1u), (done >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem0 + ((uint32_t)(warp * 32) << 16) + (done & 1u) * 128u;
      if (r + 1 < rounds) {
        // the MMAs of this round have read the A tile: next round's bytes, then its MMAs into the other buffer (whose
        // previous contents every warp has read out before it arrives at the barrier)
        make_bytes(r + 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_mma();
        }
        ++mma_count;
      }
      // records of the 8 edges of this round: folded weights, row-end bits, byte offsets of the output rows
      uint32_t efm = 0u, ro[8];
      float A[8], B[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t pos = (uint32_t)(8 * r + j) & rmask;
        const int4 rc = lds128(rec_g + (pos << 4));
        if (rc.y < 0 && 8 * r + j < rowlim) efm |= 1u << j;
        A[j] = __int_as_float(rc.z);
        B[j] = __int_as_float(rc.w);
        ro[j] = (uint32_t)lds32(row_g + (pos << 2)) * ldo4;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // request the pass two ahead: (r, q + 2) or (r + 1, q - 2); its ring slot was read by the whole group in the
        // previous pass, hence the barrier
        __syncwarp();
        if (q == 2) round_addresses(r + 1);
        const uint32_t ahead = slot >= 1 ? slot - 1 : 2;  // (slot + 2) % 3
        gather_pass(q < 2 ? r : r + 1, (q + 2) & 3, xring_s + ahead * WH_PASS_BYTES);
        cp_async_commit();
        cp_async_wait<2>();  // this pass has landed (and the record chunks requested before it)
        __syncwarp();
        // z of channels 32 q .. 32 q + 31 of this lane's tile row -> scratch -> the quads of this lane's 8 edges
        float4 z4[8];
        {
          uint32_t v[32];
          tc_ldtm32(taddr + (uint32_t)(32 * q), v);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            sts128(scr_s + (uint32_t)lane * 128u + (uint32_t)((i ^ (lane & 7)) << 4),
                   make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) z4[j] = lds128f(scr_s + (uint32_t)(sub * 8 + j) * 128u + (uint32_t)((sl ^ j) << 4));
          __syncwarp();  // the scratch may be overwritten by the next pass
        }
        const uint32_t xs = xring_s + slot * WH_PASS_BYTES + (uint32_t)((sub * 64 + sl) * 16);
        float* a4 = acc + 4 * q;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x4 = lds128f(xs + (uint32_t)(j * 128));
          a4[0] = fmaf(fmaf(z4[j].x, B[j], A[j]), x4.x, a4[0]);
          a4[1] = fmaf(fmaf(z4[j].y, B[j], A[j]), x4.y, a4[1]);
          a4[2] = fmaf(fmaf(z4[j].z, B[j], A[j]), x4.z, a4[2]);
          a4[3] = fmaf(fmaf(z4[j].w, B[j], A[j]), x4.w, a4[3]);
          if (efm & (1u << j)) {  // last edge of a row: write this quad of it, start the next row
            __stcs(reinterpret_cast<float4*>(outs + ro[j] + 128 * q), make_float4(a4[0], a4[1], a4[2], a4[3]));
            a4[0] = a4[1] = a4[2] = a4[3] = 0.f;
          }
        }
        slot = slot == 2 ? 0 : slot + 1;
      }
    }
    cp_async_wait<0>();
    __syncwarp();
    if (part_slot >= 0) {  // hub segment: its partial sum, combined by hub_finalize_kernel
      float* o = p.part_acc + ((int64_t)s * p.num_hub_segs + part_slot) * p.dpad + cg * 128 + 4 * sl;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        __stcs(reinterpret_cast<float4*>(o + 32 * q), make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]));
    }
  }
  zero_empty_rows_tail(p, (int64_t)blockIdx.x * WH_WARPS + warp, (int64_t)gridDim.x * WH_WARPS, lane);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(256u) : "memory");
}

static int launch_wh_stream(const AggParams& p, cudaStream_t stream) {
  if (p.E > 0) {
    edge_record_kernel<STAG_NOISE_NORMAL><<<(unsigned)((p.E + 255) / 256), 256, 0, stream>>>(p, const_cast<int4*>(p.rec), 3);
    STAG_LAUNCH_CHECK();
  }
  const int64_t total = (int64_t)((p.num_hub_segs + p.num_items + 15) / 16) * p.S * (p.D >> 7);
  const int64_t cap = (int64_t)num_sms() * 2;
  const int grid = (int)(total < 1 ? 1 : (total < cap ? total : cap));
  STAG_CUDA(cudaFuncSetAttribute(agg_wh_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WH_SMEM));
  agg_wh_stream_kernel<<<grid, WH_THREADS, WH_SMEM, stream>>>(p);
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}

// records, unit boundaries, kernel
static int launch_tc(const AggParams& p_, cudaStream_t stream) {
  AggParams p = p_;
  p.nunits = (int)((p.E + TC_UNIT_EDGES - 1) / TC_UNIT_EDGES);
  if (p.E > 0) {
    tc_record_kernel<<<(unsigned)((p.E + 255) / 256), 256, 0, stream>>>(p, const_cast<int4*>(p.rec),
                                                                        const_cast<uint32_t*>(p.rowoff));
    STAG_LAUNCH_CHECK();
    tc_bounds_kernel<<<(unsigned)((p.nunits + 256) / 256), 256, 0, stream>>>(p.indptr, p.N, p.E, p.nunits,
                                                                            const_cast<int32_t*>(p.ebnd));
    STAG_LAUNCH_CHECK();
  }
  const int64_t total = (int64_t)p.nunits * p.S * (p.D >> 7);
  const int64_t cap = (int64_t)num_sms() * TC_CTAS;
  const int grid = (int)(total < 1 ? 1 : (total < cap ? total : cap));
  STAG_CUDA(cudaFuncSetAttribute(agg_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  agg_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(p);
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}

}  // namespace stag
