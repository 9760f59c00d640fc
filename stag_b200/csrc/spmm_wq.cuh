// Tensor-core noise path, third form (included by spmm.cu after spmm_tc.cuh): four lanes per edge stream, z read from
// TMEM in the consumer's own layout.
//
// tcgen05.ld.16x256b hands thread t of a warp the 32-bit columns 8 i + 2 (t % 4) + {0, 1} (i = 0 .. x-1) of TMEM lanes
// base + t / 4 and base + t / 4 + 8 (base = 32 w or 32 w + 16; tools/micro/tmem_ld_layout.cu checks this on the chip).
// With tile row = EDGE and column = CHANNEL the four threads of a quad t / 4 therefore receive, without any exchange,
// what a lane group of the streaming kernels wants: the z of two consecutive edges of ONE stream, 2 adjacent columns
// out of every 8.  The column -> channel map is ours (row n of the B operand is the Hadamard row of channel(n)):
//     channel(8 k + 2 q + b) = 32 (k / 4) + 8 q + 2 (k % 4) + b
// makes the columns of thread q the octets 32 m + 8 q .. + 7 (m = 0..3) of a 128-channel group: 256-bit gathers
// (LDG.E.ENL2.256), the four threads of a quad cover one whole 128-byte line per load instruction.
//   CTA = 32 quads = 32 stream items walked in lock step, one ROUND = 4 edges per quad = one 128-row tile:
//     bytes of the round (a thread makes Philox block lane % 8 of tile rows of its own warp, first Philox round
//     hoisted) -> A tile (SWIZZLE_128B) -> 4 MMAs (one thread) -> mbarrier -> per (edge pair, channel half):
//     tcgen05.ld.16x256b.x8 (2 edges x 16 channels), w = A + B z and acc += w x as FFMA2 on the register pairs TMEM
//     delivers, rows stored when the stream passes their last edge.  The gathered rows are plain 256-bit loads into a
//     register ring (XB units of 2 edges x 16 channels in flight): no shared-memory staging.
//   NW = 4: a warp owns a TMEM quadrant and both channel halves (32 channels per thread);
//   NW = 8: warps w and w + 4 share quadrant w % 4 and take one channel half each (16 channels per thread, half the
//           registers, twice the warps to hide the TMEM / shared-memory / barrier latencies behind).
//   DB: two TMEM accumulators (256 columns), the MMAs of round r + 1 run under the consumption of round r.
//   CTA items (32 stream items each) are handed out by a work queue; template switches GMODE / LATE_REC (where the
//   gathers are re-issued, when the records of a round are read) are documented at the kernel.
#pragma once

namespace stag {

constexpr int WQ_GROUPS = 32, WQ_EPR = 4;
constexpr int WQ_NBUF = 4;  // record chunks (one round each) per quad ring
constexpr uint32_t WQ_REC_STRIDE = WQ_NBUF * WQ_EPR * 16 + 16;  // skewed by 16 bytes: the 8 quads of a warp read 8 bank groups
constexpr uint32_t WQ_ROW_STRIDE = WQ_NBUF * WQ_EPR * 4 + 4;
__host__ __device__ constexpr uint32_t wq_rec_bytes(int nw) { return (uint32_t)(8 * nw) * WQ_REC_STRIDE; }
__host__ __device__ constexpr uint32_t wq_row_bytes(int nw) { return ((uint32_t)(8 * nw) * WQ_ROW_STRIDE + 15u) & ~15u; }
__host__ __device__ constexpr uint32_t wq_smem(int nw) { return 2 * 128 * 128 + wq_rec_bytes(nw) + wq_row_bytes(nw) + 1024; }

__device__ __forceinline__ void tc_ldtm_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ldtm_wait32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ void tc_ldtm_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ldtm_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void ldg256(const char* src, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(src));
}
__device__ __forceinline__ void stg256cs(char* dst, const float2* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x),
               "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y)
               : "memory");
}

// NW warps; XB units of the register ring (XB divides the 8 / NW * 2 units of a round: the slot of a unit must not
// depend on the round); DB double-buffered TMEM; MINB CTAs per SM
// GMODE: where the registers of a consumed unit get their next gather: 0 after each edge, 1 after the unit (4 LDG.256
//        back to back: 1 % faster), 2 all 16 of a round at its end (as fast: the ring needs ~1 500 cycles of flight time,
//        which the Philox phase of the next round provides; a half-round ring, XB = 2, does not: 1.8 - 2.0 ms)
// LATE_REC: the records / rows / next neighbours of a round are read from the ring AFTER the Philox phase of the next round
//        instead of before it (24 registers fewer live across it: 1.47 -> 1.43 ms)
template <int NW, int XB, bool DB, int MINB, int GMODE = 0, bool LATE_REC = false>
__global__ void __launch_bounds__(32 * NW, MINB) agg_wh_quad_kernel(const AggParams p) {
  constexpr int NCH = 8 / NW;     // channel halves per thread
  constexpr int NU = 2 * NCH;     // units (edge pair, channel half) per round
  constexpr int NBLK = 32 / NW;   // tile rows (Philox blocks) a thread makes per round
  static_assert(NW == 4 || NW == 8, "4 or 8 warps");
  static_assert(NU % XB == 0, "ring slots must be round-invariant");
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int maxn_s[NW];
  __shared__ int arrive_cnt[2];
  constexpr uint32_t TM_COLS = DB ? 256u : 128u;
  const uint32_t sm0 = ((uint32_t)__cvta_generic_to_shared(tc_raw) + 1023u) & ~1023u;
  const uint32_t h_s = sm0, at_s = sm0 + 128 * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3;                    // TMEM quadrant = the 8 quads (stream items) of this warp
  const int ch0 = NW == 8 ? warp >> 2 : 0;    // first channel half of this thread
  const int jq = lane >> 2, q = lane & 3;     // quad of the warp, thread of the quad
  const uint32_t rec_base = at_s + 128 * 128, row_base = rec_base + wq_rec_bytes(NW);
  const uint32_t rec_g = rec_base + (uint32_t)(8 * warp + jq) * WQ_REC_STRIDE;  // every warp keeps its own rings
  const uint32_t row_g = row_base + (uint32_t)(8 * warp + jq) * WQ_ROW_STRIDE;
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(mma_bar);
  constexpr uint32_t rmask = WQ_NBUF * WQ_EPR - 1;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "r"(TM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    arrive_cnt[0] = arrive_cnt[1] = 0;
    tc_mbar_init(bar_s, 1);
    tc_mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int n = tid; n < 128; n += 32 * NW) {
    // row n of the B operand (N = TMEM column n, K-major) = Hadamard row of channel(n)
    const int kk = n >> 3, qq = (n >> 1) & 3, bb = n & 1;
    const int c = 32 * (kk >> 2) + 8 * qq + 2 * (kk & 3) + bb;
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t wv[4];
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = ch * 16 + w4 * 4 + b;
          word |= ((__popc(c & k) & 1) ? 0xB8u : 0x38u) << (8 * b);
        }
        wv[w4] = word;
      }
      sts128(h_s + (uint32_t)n * 128u + (uint32_t)((ch ^ (n & 7)) << 4), make_uint4(wv[0], wv[1], wv[2], wv[3]));
    }
  }
  for (uint32_t o = (uint32_t)tid * 16u; o < 128u * 128u; o += 32u * NW * 16u) sts128(at_s + o, make_uint4(0u, 0u, 0u, 0u));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem0 = tmem_base_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t h_desc = tc_desc(h_s), a_desc = tc_desc(at_s);
  const PhiloxKey key = live_key(p);
  uint32_t mma_count = 0;  // MMA chains issued so far by this CTA (every thread counts): mbarrier = count & 1, parity = (count >> 1) & 1

  const int G = p.D >> 7;
  const int n_items = p.num_hub_segs + p.num_items;
  const int IG = (n_items + WQ_GROUPS - 1) / WQ_GROUPS;  // CTA items per (sample, channel group)
  const int64_t total = (int64_t)IG * p.S * G;
  const uint32_t ldxb = (uint32_t)p.ldx * 4u, ldo4 = (uint32_t)p.ldo * 4u;
  // tile rows this thread makes the bytes of: row rb + lane / 8 + 4 i (i < NBLK) of its quadrant, i.e. edge (row >> 3) of the
  // round of quad (row & 7) = lane / 8 + 4 (i & 1); Philox block lane % 8
  const int rb = NW == 8 ? 16 * (warp >> 2) : 0;
  const uint32_t mk_rec = rec_base + (uint32_t)(8 * warp + (lane >> 3)) * WQ_REC_STRIDE + 4u;
  const uint32_t mk_row = at_s + (uint32_t)(32 * wq + rb + (lane >> 3)) * 128u;

  // CTA items are handed out by a queue (a counter behind the edge records, zeroed by edge_record_kernel): they differ in
  // length (1 .. 47 rounds at the arxiv shape), a static stride left the slowest CTA 10 - 13 % more rounds than the
  // average (tools/item_balance.py).  The next index is fetched while the current item runs.
  __shared__ int item_s;
  int* const queue = reinterpret_cast<int*>(const_cast<int4*>(p.rec) + p.E);
  if (tid == 0) item_s = atomicAdd(queue, 1);
  __syncthreads();
  for (;;) {
    const int64_t item = item_s;
    if (item >= total) break;
    const int64_t outer = item / IG;
    const int gi = (int)(item - outer * IG) * WQ_GROUPS + 8 * wq + jq;
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
    __syncthreads();  // the previous item's readers of maxn_s are done, and everyone has read item_s
    if (lane == 0) maxn_s[warp] = maxn;
    if (tid == 0) item_s = atomicAdd(queue, 1);
    __syncthreads();
    maxn = max(max(maxn_s[0], maxn_s[1]), max(maxn_s[2], maxn_s[3]));  // warps 4..7 walk the items of warps 0..3
    const int rounds = (maxn + WQ_EPR - 1) / WQ_EPR;
    if (rounds == 0) continue;  // uniform: every MMA chain issued below is waited for exactly once, in order

    const char* xcb = reinterpret_cast<const char*>(p.x + (int64_t)s * p.x_ss + cg * 128 + 64 * ch0 + 8 * q);
    char* outs = reinterpret_cast<char*>(p.out + (int64_t)s * p.out_ss + cg * 128 + 64 * ch0 + 8 * q);
    const uint32_t smp = (uint32_t)(p.sample_base + s);
    const uint32_t blk = (uint32_t)(8 * cg + (lane & 7));
    const int4* recp = p.rec + e0;
    const int32_t* rowp = p.erow + e0;
    float2 acc[8 * NCH];  // [channel half][load mm][pair e2]: channels 64 ch + 32 mm + 8 q + 2 e2 + {0, 1}
#pragma unroll
    for (int i = 0; i < 8 * NCH; ++i) acc[i] = make_float2(0.f, 0.f);
    float4 xb[XB][8];  // units in flight: [edge of the pair][octet mm] as two quads

    auto fetch_chunk = [&](int r) {  // records of round r -> ring (zeros past the end of the item)
      const int e = WQ_EPR * r + q;
      const uint32_t pos = (uint32_t)e & rmask;
      const bool off = e >= nedges;
      cp_async16(rec_g + pos * 16u, recp + (off ? 0 : e), off);
      cp_async4(row_g + pos * 4u, rowp + (off ? 0 : e), off);
    };
    auto make_bytes = [&](int r) {
      // (the shared-memory helpers are volatile asm: the loads are batched by hand so that the Philox chains below are
      // independent instruction streams the scheduler can interleave)
      uint32_t eid[NBLK];
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        const int k = (rb + 4 * i) >> 3;  // edge of the round (lane / 8 < 4 does not carry)
        const uint32_t pos = (uint32_t)(WQ_EPR * r + k) & rmask;
        eid[i] = (uint32_t)lds32(mk_rec + (uint32_t)(4 * (i & 1)) * WQ_REC_STRIDE + (pos << 4));
      }
      uint4 v[NBLK];
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        v[i] = philox_rk(blk, eid[i] & 0x7fffffffu, smp, key.c3, p);
        v[i].x = (v[i].x & kWhAnd) | kWhOr; v[i].y = (v[i].y & kWhAnd) | kWhOr;
        v[i].z = (v[i].z & kWhAnd) | kWhOr; v[i].w = (v[i].w & kWhAnd) | kWhOr;
      }
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        const int sw = ((lane >> 3) + 4 * (i & 1)) & 7;  // (tile row) & 7
        sts128(mk_row + (uint32_t)(4 * i) * 128u + (uint32_t)(((lane & 7) ^ sw) << 4), v[i]);
      }
    };
    auto issue_mma = [&]() {  // one thread: lane 0 of the warp that arrived last
      const uint32_t buf = DB ? (mma_count & 1u) : 0u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tc_mma_f8(tmem0 + buf * 128u, a_desc + (uint64_t)(2 * ks), h_desc + (uint64_t)(2 * ks), idesc, ks != 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s + 8u * (mma_count & 1u))
                   : "memory");
    };
    // No CTA barrier: every warp announces its tile rows (and that it has read the accumulator the MMAs will overwrite)
    // on a shared counter; the warp that arrives last issues the MMAs.  Warps only ever block on the MMA mbarrier.
    auto bytes_and_mma = [&](int r) {
      make_bytes(r);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        // one acq_rel read-modify-write instead of fence + atomic + fence: a MEMBAR.CTA costs ~190 cycles with the 8
        // tile stores of the round in flight, and the other 31 lanes of the warp wait for lane 0 at the reconvergence
        uint32_t old;
        const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared(&arrive_cnt[mma_count & 1u]);
        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(cnt_s) : "memory");
        if (old == NW - 1) {
          asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(cnt_s), "r"(0u) : "memory");  // next used two rounds on
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_mma();
        }
      }
      ++mma_count;
    };
    // gathered octets of local channel half chl of the row of neighbour un -> 4 register quads of the ring
    auto load_half = [&](uint32_t un, int chl, float4* dst) {
      const char* src = xcb + (uint64_t)un * ldxb + 256 * chl;
      ldg256(src, dst[0], dst[1]);
      ldg256(src + 128, dst[2], dst[3]);
    };
    // neighbours of the 4 edges of round r (0 past the end of the item: row 0 is read and never accumulated)
    auto neighbours = [&](int r, uint32_t (&un)[4]) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) un[k4] = (uint32_t)lds32(rec_g + (((uint32_t)(WQ_EPR * r + k4) & rmask) << 4));
    };

    // ---- prologue of the item ----
    fetch_chunk(0);
    fetch_chunk(1);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    bytes_and_mma(0);
    uint32_t un1[4];  // neighbours of the next round
    {
      uint32_t un0[4];
      neighbours(0, un0);
#pragma unroll
      for (int u = 0; u < XB; ++u) {
        load_half(un0[2 * (u / NCH)], u % NCH, xb[u]);
        load_half(un0[2 * (u / NCH) + 1], u % NCH, xb[u] + 4);
      }
    }

    for (int r = 0; r < rounds; ++r) {
      cp_async_wait<0>();  // records of round r + 1 (requested during round r - 1)
      __syncwarp();
      fetch_chunk(r + 2);
      cp_async_commit();
      int4 rc4[4];  // records of this round's edges
      uint32_t ro4[4];
      auto read_records = [&]() {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) rc4[k4] = lds128(rec_g + (((uint32_t)(WQ_EPR * r + k4) & rmask) << 4));
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) ro4[k4] = (uint32_t)lds32(row_g + (((uint32_t)(WQ_EPR * r + k4) & rmask) << 2));
        neighbours(r + 1, un1);
      };
      if (!LATE_REC) read_records();
      const uint32_t done = mma_count - 1u;
      tc_mbar_wait(bar_s + 8u * (done & 1u), (done >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem0 + ((uint32_t)(32 * wq) << 16) + (DB ? (done & 1u) * 128u : 0u) + (uint32_t)(64 * ch0);
      if (DB && r + 1 < rounds) bytes_and_mma(r + 1);  // the MMAs of round r have read the A tile; the other accumulator was read out before the barrier
      if (LATE_REC) read_records();
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int hf = u / NCH, chl = u % NCH;
        uint32_t z[32];
        tc_ldtm_16x256b_x8(taddr + ((uint32_t)(16 * hf) << 16) + (uint32_t)(64 * chl), z);
        float4* xu = xb[u % XB];
        tc_ldtm_wait32(z);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int t = WQ_EPR * r + 2 * hf + kk;
          const int4 rc = rc4[2 * hf + kk];
          if (t < nedges) {
            const float2 AA = make_float2(__int_as_float(rc.z), __int_as_float(rc.z));
            const float2 BB = make_float2(__int_as_float(rc.w), __int_as_float(rc.w));
#pragma unroll
            for (int mm = 0; mm < 2; ++mm) {
              const float xv[8] = {xu[4 * kk + 2 * mm].x, xu[4 * kk + 2 * mm].y, xu[4 * kk + 2 * mm].z, xu[4 * kk + 2 * mm].w,
                                   xu[4 * kk + 2 * mm + 1].x, xu[4 * kk + 2 * mm + 1].y, xu[4 * kk + 2 * mm + 1].z, xu[4 * kk + 2 * mm + 1].w};
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                const int zi = 16 * mm + 4 * e2 + 2 * kk;
                const float2 w2 = __ffma2_rn(make_float2(__uint_as_float(z[zi]), __uint_as_float(z[zi + 1])), BB, AA);
                acc[8 * chl + 4 * mm + e2] = __ffma2_rn(w2, make_float2(xv[2 * e2], xv[2 * e2 + 1]), acc[8 * chl + 4 * mm + e2]);
              }
            }
            if (rc.y < 0 && t < rowlim) {  // last edge of a row: write this channel half of it, start the next row
              char* o = outs + ro4[2 * hf + kk] * ldo4 + 256 * chl;
#pragma unroll
              for (int mm = 0; mm < 2; ++mm) {
                stg256cs(o + 128 * mm, acc + 8 * chl + 4 * mm);
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) acc[8 * chl + 4 * mm + e2] = make_float2(0.f, 0.f);
              }
            }
          }
          if (GMODE == 0) {
            // the registers just consumed take the same edge of unit u + XB (of this round or the next)
            const int nu = (u + XB) % NU;
            const uint32_t un = (u + XB) < NU ? (uint32_t)rc4[2 * (nu / NCH) + kk].x : un1[2 * (nu / NCH) + kk];
            load_half(un, nu % NCH, xu + 4 * kk);
          }
        }
        if (GMODE == 1) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int nu = (u + XB) % NU;
            const uint32_t un = (u + XB) < NU ? (uint32_t)rc4[2 * (nu / NCH) + kk].x : un1[2 * (nu / NCH) + kk];
            load_half(un, nu % NCH, xu + 4 * kk);
          }
        }
      }
      if (GMODE == 2) {
        static_assert(GMODE != 2 || XB == NU, "round-end gathers need a whole round in the ring");
#pragma unroll
        for (int u = 0; u < NU; ++u)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) load_half(un1[2 * (u / NCH) + kk], u % NCH, xb[u] + 4 * kk);
      }
      if (!DB && r + 1 < rounds) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        bytes_and_mma(r + 1);
      }
    }
    cp_async_wait<0>();
    __syncwarp();
    if (part_slot >= 0) {  // hub segment: its partial sum, combined by hub_finalize_kernel
      float* o = p.part_acc + ((int64_t)s * p.num_hub_segs + part_slot) * p.dpad + cg * 128 + 64 * ch0 + 8 * q;
#pragma unroll
      for (int m = 0; m < 2 * NCH; ++m) stg256cs(reinterpret_cast<char*>(o + 32 * m), acc + 4 * m);
    }
  }
  zero_empty_rows_tail(p, (int64_t)blockIdx.x * NW + warp, (int64_t)gridDim.x * NW, lane);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(TM_COLS) : "memory");
}

template <int NW, int XB, bool DB, int MINB, int GMODE = 0, bool LATE_REC = false>
static int launch_wh_quad_inst(const AggParams& p, cudaStream_t stream) {
  const int64_t total = (int64_t)((p.num_hub_segs + p.num_items + WQ_GROUPS - 1) / WQ_GROUPS) * p.S * (p.D >> 7);
  const int64_t cap = (int64_t)num_sms() * MINB;
  const int grid = (int)(total < 1 ? 1 : (total < cap ? total : cap));
  // TMEM holds 512 columns per SM: pad the shared-memory request so that no more CTAs than that become resident
  constexpr int max_ctas = DB ? 2 : 4;
  const size_t pad = (size_t)(227 * 1024) / (max_ctas + 1) + 1024;
  const size_t smem = wq_smem(NW) > pad ? wq_smem(NW) : pad;
  STAG_CUDA(cudaFuncSetAttribute(agg_wh_quad_kernel<NW, XB, DB, MINB, GMODE, LATE_REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  agg_wh_quad_kernel<NW, XB, DB, MINB, GMODE, LATE_REC><<<grid, 32 * NW, smem, stream>>>(p);
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}

static int launch_wh_quad(const AggParams& p, cudaStream_t stream) {
  if (p.E > 0) {   // also zeroes the work queue behind the records
    edge_record_kernel<STAG_NOISE_NORMAL><<<(unsigned)((p.E + 255) / 256), 256, 0, stream>>>(p, const_cast<int4*>(p.rec), 3);
    STAG_LAUNCH_CHECK();
  } else {
    STAG_CUDA(cudaMemsetAsync(const_cast<int4*>(p.rec), 0, 16, stream));
  }
  STAG_CHECK_ARG((int64_t)((p.num_hub_segs + p.num_items + WQ_GROUPS - 1) / WQ_GROUPS) * p.S * (p.D >> 7) < (1ll << 30),
                 "stag_spmm: too many work items for one launch");
  // XB = 4 (a whole round of gathered rows in flight per lane) with two TMEM accumulators is the fastest measured
  // form (profiles/r02_wq_forms.txt); STAG_WQ_VARIANT selects the others for A/B runs
  static const char* var = getenv("STAG_WQ_VARIANT");
  const int v = var ? atoi(var) : 0;
  switch (v) {
    case 1: return launch_wh_quad_inst<4, 4, false, 2, 1>(p, stream);
    case 2: return launch_wh_quad_inst<4, 2, false, 3>(p, stream);
    case 3: return launch_wh_quad_inst<8, 2, true, 2>(p, stream);
    case 4: return launch_wh_quad_inst<4, 4, true, 2, 0>(p, stream);   // gathers re-issued per edge (the default until late r02)
    case 5: return launch_wh_quad_inst<4, 4, true, 2, 2>(p, stream);   // all gathers of a round at its end
    case 6: return launch_wh_quad_inst<4, 4, true, 2, 1, false>(p, stream);  // records read before the Philox phase
    case 7: return launch_wh_quad_inst<4, 4, true, 2, 2, true>(p, stream);
    default: return launch_wh_quad_inst<4, 4, true, 2, 1, true>(p, stream);
  }
}

}  // namespace stag
