// On-device CSC / CSR builder: stable LSD radix sort of the COO edge list by
// destination (CSC) or source (CSR), row pointers from the sorted keys, and the
// hub-row schedule used by the aggregation kernels.
//
// Replaces DGL's COO->CSC/CSR conversion that the reference triggers through
// graph.update_all (stag/zoo/gcn.py:95, stag/layers.py:12-15) and
// graph.in_degrees()/out_degrees() (stag/zoo/gcn.py:68,101; stag/layers.py:21).
// Bit-exact target: oracle/ref_index.py (stable argsort).
#include "common.cuh"

namespace stag {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;  // 4096 keys per CTA
constexpr int RS_RADIX = 256;

__global__ void k_init_keys(const int64_t* __restrict__ key64, int64_t E, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) {
    keys[i] = (uint32_t)key64[i];
    vals[i] = (uint32_t)i;
  }
}

// per-CTA digit histogram, stored digit-major: hist[d * nblocks + b]
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint32_t* __restrict__ keys, int64_t E,
                                                           int shift, uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t h[RS_RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; ++r) {
    int64_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < E) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// single-CTA exclusive scan, in place, n arbitrary.  Also returns the total in *total if non-null.
__global__ void __launch_bounds__(1024) k_exclusive_scan(uint32_t* __restrict__ data, int64_t n,
                                                        uint32_t* __restrict__ total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + tid;
    const uint32_t v = i < n ? data[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t warp_off = wid == 0 ? 0u : warp_sums[wid - 1];
    if (i < n) data[i] = carry + warp_off + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
  if (tid == 0 && total) *total = carry_s;
}

// stable scatter of one CTA tile: ranks are assigned in (round, warp, lane) order, which is
// the key order inside the tile.
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, int64_t E, int shift, const uint32_t* __restrict__ hist_scanned,
    int nblocks) {
  __shared__ uint32_t running[RS_RADIX];
  __shared__ uint32_t warp_cnt[RS_WARPS][RS_RADIX];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  running[tid] = hist_scanned[(size_t)tid * nblocks + blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ROUNDS; ++r) {
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) warp_cnt[w][tid] = 0;
    __syncthreads();
    const int64_t i = base + r * RS_THREADS + tid;
    const bool valid = i < E;
    uint32_t key = 0, val = 0, digit = 0xffffffffu;
    if (valid) {
      key = keys_in[i];
      val = vals_in[i];
      digit = (key >> shift) & 0xff;
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warp_cnt[wid][digit] = __popc(peers);
    __syncthreads();
    {
      uint32_t run = running[tid];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = warp_cnt[w][tid];
        warp_cnt[w][tid] = run;
        run += c;
      }
      running[tid] = run;
    }
    __syncthreads();
    if (valid) {
      const uint32_t pos = warp_cnt[wid][digit] + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
  }
}

// indptr from sorted keys: indptr[v] = first position whose key >= v
__global__ void k_indptr_from_sorted(const uint32_t* __restrict__ keys, int64_t E, int64_t N,
                                     int32_t* __restrict__ indptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > E) return;
  const int64_t prev = i == 0 ? -1 : (int64_t)keys[i - 1];
  const int64_t cur = i == E ? N : (int64_t)keys[i];
  for (int64_t v = prev + 1; v <= cur; ++v) indptr[v] = (int32_t)i;
}

__global__ void k_gather_other(const uint32_t* __restrict__ eid_sorted, const uint32_t* __restrict__ key_sorted,
                               const int64_t* __restrict__ other64, int64_t E, int32_t* __restrict__ indices,
                               int32_t* __restrict__ eid, int32_t* __restrict__ erow, int32_t* __restrict__ eidf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) {
    const uint32_t e = eid_sorted[i];
    eid[i] = (int32_t)e;
    indices[i] = (int32_t)other64[e];
    if (erow) erow[i] = (int32_t)key_sorted[i];
    if (eidf) {
      const bool last = i + 1 == E || key_sorted[i + 1] != key_sorted[i];
      eidf[i] = (int32_t)(e | (last ? 0x80000000u : 0u));
    }
  }
}

// stream items by decreasing edge count, so that the items sharing a warp have equal trip counts
__global__ void k_item_keys(const int32_t* __restrict__ items, int64_t n, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    const int e1 = items[4 * k + 3];
    keys[k] = e1 < 0 ? 0u : (uint32_t)min(e1 - items[4 * k + 2], 255);
    vals[k] = (uint32_t)k;
  }
}

__global__ void k_item_permute(const int4* __restrict__ src, const uint32_t* __restrict__ order, int64_t n,
                               int4* __restrict__ dst) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[order[n - 1 - k]];
}

// hub schedule ---------------------------------------------------------------------------
__global__ void k_hub_flags(const int32_t* __restrict__ indptr, int64_t N, uint32_t* __restrict__ is_hub,
                            uint32_t* __restrict__ nseg) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) {
    const int deg = indptr[v + 1] - indptr[v];
    const bool hub = deg > kHubThreshold;
    is_hub[v] = hub ? 1u : 0u;
    nseg[v] = hub ? (uint32_t)((deg + kHubSegment - 1) / kHubSegment) : 0u;
  }
}

__global__ void k_hub_scatter(const int32_t* __restrict__ indptr, int64_t N, const uint32_t* __restrict__ hub_idx,
                              const uint32_t* __restrict__ seg_off, const uint32_t* __restrict__ totals,
                              int32_t* __restrict__ hub_rows, int32_t* __restrict__ hub_seg_ptr) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) {
    const int deg = indptr[v + 1] - indptr[v];
    if (deg > kHubThreshold) {
      hub_rows[hub_idx[v]] = (int32_t)v;
      hub_seg_ptr[hub_idx[v]] = (int32_t)seg_off[v];
    }
  }
  if (v == 0) hub_seg_ptr[totals[0]] = (int32_t)totals[1];
}

// row processing order: rows sorted by decreasing stored-edge count (ties: decreasing row id), so
// that the rows sharing a warp -- and the warps sharing a CTA -- have equal trip counts.  Hub rows
// (handled as segments) all carry the same clamped key.
__global__ void k_degree_keys(const int32_t* __restrict__ indptr, int64_t N, uint32_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) {
    const int deg = indptr[v + 1] - indptr[v];
    keys[v] = (uint32_t)min(deg, kHubThreshold + 1);
    vals[v] = (uint32_t)v;
  }
}

__global__ void k_reverse_order(const uint32_t* __restrict__ sorted_rows, int64_t N, int32_t* __restrict__ row_order) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) row_order[i] = (int32_t)sorted_rows[N - 1 - i];
}

// ---- stream items: consecutive row ranges holding about kRangeEdges stored edges ---------------
// A row starts a new item when it is the first row whose edges begin in a new block of kRangeEdges
// positions, every kRangeRows rows, and around hub rows (a hub row is an item of its own that the
// streaming kernel skips: its edges are processed as hub segments).  item = {row0, row1, e0, e1};
// e1 = -1 marks a hub placeholder.
__device__ __forceinline__ bool range_starts_at(const int32_t* __restrict__ indptr, int64_t v, int re) {
  if (v == 0 || (v % kRangeRows) == 0) return true;
  const int b = indptr[v], a = indptr[v - 1], c = indptr[v + 1];
  return (c - b > kHubThreshold) || (b - a > kHubThreshold) || (b / re != a / re);
}

__global__ void k_range_flags(const int32_t* __restrict__ indptr, int64_t N, int re, uint32_t* __restrict__ flags) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) flags[v] = range_starts_at(indptr, v, re) ? 1u : 0u;
}

__global__ void k_range_scatter(const int32_t* __restrict__ indptr, int64_t N, int re, const uint32_t* __restrict__ pos,
                                int32_t* __restrict__ items) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N && range_starts_at(indptr, v, re)) {
    items[4 * (int64_t)pos[v] + 0] = (int32_t)v;
    items[4 * (int64_t)pos[v] + 2] = indptr[v];
  }
}

__global__ void k_range_close(const int32_t* __restrict__ indptr, int64_t N, const uint32_t* __restrict__ total,
                              int32_t* __restrict__ items) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = total[0];
  if (k >= n) return;
  const int row0 = items[4 * k];
  const int row1 = k + 1 < n ? items[4 * (k + 1)] : (int)N;
  items[4 * k + 1] = row1;
  const bool hub = indptr[row0 + 1] - indptr[row0] > kHubThreshold;
  items[4 * k + 3] = hub ? -1 : indptr[row1];
}


// ---- minibatch-sized graphs: the whole build in ONE launch ---------------------------------------------
// A batch of molecule- or PPI-sized graphs (E, N <= kSmallMax) is built by one CTA in shared memory: bitonic
// sort of (key << 13 | edge id) -- the composite key makes the order total, i.e. the stable order --, degree
// histogram + scan for the row pointers, and the same hub / row-order / stream-item schedules as the general
// path (same kernels' semantics, same tests).  ~35 launches and two host synchronisations become one of each.
constexpr int kSmallMax = 8192;
constexpr int kSmallThreads = 1024;

__device__ __forceinline__ void bitonic_sort_u32(uint32_t* a, int P) {  // P a power of two, ascending
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += kSmallThreads) {
        const int l = i ^ j;
        if (l > i) {
          const uint32_t x = a[i], y = a[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// exclusive scan of v[0..n) in place (n <= kSmallMax), total returned to every thread
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t* v, int n, uint32_t* warp_sums) {
  constexpr int PER = kSmallMax / kSmallThreads;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t loc[PER], sum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = tid * PER + k;
    loc[k] = i < n ? v[i] : 0u;
    sum += loc[k];
  }
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = warp_sums[lane], z = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, z, o);
      if (lane >= o) z += y;
    }
    warp_sums[lane] = z - w;           // exclusive prefix of the warp sums
    if (lane == 31) warp_sums[32] = z;  // total
  }
  __syncthreads();
  uint32_t run = warp_sums[wid] + x - sum;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = tid * PER + k;
    if (i < n) v[i] = run;
    run += loc[k];
  }
  const uint32_t total = warp_sums[32];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(kSmallThreads) k_build_small(
    const int64_t* __restrict__ key64, const int64_t* __restrict__ other64, int E, int N, int32_t* __restrict__ indptr,
    int32_t* __restrict__ indices, int32_t* __restrict__ eid, int32_t* __restrict__ hub_rows,
    int32_t* __restrict__ hub_seg_ptr, int32_t* __restrict__ row_order, int32_t* __restrict__ items,
    int32_t* __restrict__ erow, int32_t* __restrict__ eidf, int32_t* __restrict__ counts) {
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* comp = sm;                     // [kSmallMax] sort buffer
  uint32_t* ip = comp + kSmallMax;         // [kSmallMax + 1 (+3: keeps `its` 16-byte aligned)] degrees -> row pointers
  uint32_t* fa = ip + kSmallMax + 4;       // [kSmallMax] flags / scan
  uint32_t* fb = fa + kSmallMax;           // [kSmallMax] flags / scan
  uint32_t* wsum = fb + kSmallMax;         // [33]
  int4* its = reinterpret_cast<int4*>(wsum + 36);  // stream items before their sort
  const int tid = threadIdx.x;
  auto pow2 = [](int n) { int p = 1; while (p < n) p <<= 1; return p; };

  // edges by (key, edge id)
  const int EP = pow2(E);
  for (int i = tid; i < EP; i += kSmallThreads) comp[i] = i < E ? ((uint32_t)key64[i] << 13) | (uint32_t)i : 0xffffffffu;
  for (int v = tid; v <= N; v += kSmallThreads) ip[v] = 0u;
  __syncthreads();
  bitonic_sort_u32(comp, EP);
  for (int j = tid; j < E; j += kSmallThreads) {
    const uint32_t k = comp[j] >> 13, e = comp[j] & 8191u;
    eid[j] = (int32_t)e;
    indices[j] = (int32_t)other64[e];
    if (erow) erow[j] = (int32_t)k;
    if (eidf) eidf[j] = (int32_t)(e | ((j + 1 == E || (comp[j + 1] >> 13) != k) ? 0x80000000u : 0u));
    atomicAdd(&ip[k], 1u);
  }
  __syncthreads();
  block_exclusive_scan(ip, N + 1 <= kSmallMax ? N + 1 : kSmallMax, wsum);  // ip[v] = first stored edge of row v
  if (N == kSmallMax && tid == 0) ip[N] = (uint32_t)E;
  __syncthreads();
  for (int v = tid; v <= N; v += kSmallThreads) indptr[v] = (int32_t)ip[v];
  auto deg = [&](int v) { return (int)(ip[v + 1] - ip[v]); };

  // rows by decreasing clamped degree, ties by decreasing row
  if (row_order) {
    const int NP = pow2(N);
    for (int v = tid; v < NP; v += kSmallThreads)
      comp[v] = v < N ? ((uint32_t)min(deg(v), kHubThreshold + 1) << 13) | (uint32_t)v : 0xffffffffu;
    __syncthreads();
    bitonic_sort_u32(comp, NP);
    for (int i = tid; i < N; i += kSmallThreads) row_order[i] = (int32_t)(comp[N - 1 - i] & 8191u);
    __syncthreads();
  }

  // hub schedule
  for (int v = tid; v < N; v += kSmallThreads) {
    const int d = deg(v);
    fa[v] = d > kHubThreshold ? 1u : 0u;
    fb[v] = d > kHubThreshold ? (uint32_t)((d + kHubSegment - 1) / kHubSegment) : 0u;
  }
  __syncthreads();
  const uint32_t nhubs = block_exclusive_scan(fa, N, wsum);
  const uint32_t nsegs = block_exclusive_scan(fb, N, wsum);
  for (int v = tid; v < N; v += kSmallThreads)
    if (deg(v) > kHubThreshold) {
      hub_rows[fa[v]] = v;
      hub_seg_ptr[fa[v]] = (int32_t)fb[v];
    }
  if (tid == 0) {
    hub_seg_ptr[nhubs] = (int32_t)nsegs;
    counts[0] = (int32_t)nhubs;
    counts[1] = (int32_t)nsegs;
  }
  __syncthreads();

  // stream items
  uint32_t nitems = 0;
  if (items) {
    auto starts = [&](int v) {
      if (v == 0 || (v % kRangeRows) == 0) return true;
      const int b = (int)ip[v], a = (int)ip[v - 1], c = (int)ip[v + 1];
      return (c - b > kHubThreshold) || (b - a > kHubThreshold) || (b / kRangeEdgesSmall != a / kRangeEdgesSmall);
    };
    for (int v = tid; v < N; v += kSmallThreads) fa[v] = starts(v) ? 1u : 0u;
    __syncthreads();
    nitems = block_exclusive_scan(fa, N, wsum);
    for (int v = tid; v < N; v += kSmallThreads)
      if (starts(v)) {
        its[fa[v]].x = v;
        its[fa[v]].z = (int)ip[v];
      }
    __syncthreads();
    for (int k = tid; k < (int)nitems; k += kSmallThreads) {
      const int row0 = its[k].x;
      const int row1 = k + 1 < (int)nitems ? its[k + 1].x : N;
      its[k].y = row1;
      its[k].w = deg(row0) > kHubThreshold ? -1 : (int)ip[row1];
    }
    __syncthreads();
    const int IP = pow2((int)nitems);
    for (int k = tid; k < IP; k += kSmallThreads) {
      uint32_t key = 0xffffffffu;
      if (k < (int)nitems) key = ((uint32_t)(its[k].w < 0 ? 0 : min(its[k].w - its[k].z, 255)) << 13) | (uint32_t)k;
      comp[k] = key;
    }
    __syncthreads();
    bitonic_sort_u32(comp, IP);
    for (int k = tid; k < (int)nitems; k += kSmallThreads)
      reinterpret_cast<int4*>(items)[k] = its[comp[nitems - 1 - k] & 8191u];
  }
  if (tid == 0) counts[2] = (int32_t)nitems;
}

}  // namespace stag
extern "C" int64_t stag_csx_items_capacity(int64_t num_edges, int64_t num_nodes);
namespace stag {

struct CsxWorkspace {
  uint32_t *keys_a, *vals_a, *keys_b, *vals_b, *hist, *flags_a, *flags_b, *totals, *items_tmp;
  int nblocks;
};

static size_t carve(int64_t E, int64_t N, char* base, CsxWorkspace* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return (uint32_t*)p;
  };
  const int64_t M = E > N ? E : N;  // the same buffers sort the E edges and, later, the N rows
  const int nblocks = (int)((M + RS_TILE - 1) / RS_TILE);
  const size_t e = (size_t)(M > 0 ? M : 1);
  CsxWorkspace tmp;
  tmp.nblocks = nblocks;
  tmp.keys_a = take(e * 4);
  tmp.vals_a = take(e * 4);
  tmp.keys_b = take(e * 4);
  tmp.vals_b = take(e * 4);
  tmp.hist = take((size_t)RS_RADIX * (nblocks > 0 ? nblocks : 1) * 4);
  tmp.flags_a = take((size_t)(N + 1) * 4);
  tmp.flags_b = take((size_t)(N + 1) * 4);
  tmp.totals = take(256);
  tmp.items_tmp = take((size_t)stag_csx_items_capacity(E, N) * 16);
  if (w) *w = tmp;
  return off;
}

}  // namespace stag

using namespace stag;

extern "C" int64_t stag_csx_items_capacity(int64_t num_edges, int64_t num_nodes) {
  // one item per started block of kRangeEdges edge positions and of kRangeRows rows, plus up to three
  // around every hub row
  return num_edges / range_edges_for(num_edges) + num_nodes / kRangeRows + 3 * (num_edges / (kHubThreshold + 1)) + 4;
}

extern "C" size_t stag_csx_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
  return carve(num_edges, num_nodes, nullptr, nullptr);
}

extern "C" int stag_csx_build(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, int by_dst,
                              int32_t* indptr, int32_t* indices, int32_t* eid, int32_t* hub_rows,
                              int32_t* hub_seg_ptr, int32_t* row_order, int32_t* items, int32_t* erow, int32_t* eidf, int32_t* counts_host,
                              void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  STAG_CHECK_ARG(E >= 0 && N >= 0, "stag_csx_build: negative sizes");
  STAG_CHECK_ARG(E < (1ll << 31) && N < (1ll << 31) - 1, "stag_csx_build: E and N must be < 2^31");
  STAG_CHECK_ARG(indptr && counts_host && hub_seg_ptr, "stag_csx_build: null output");
  STAG_CHECK_ARG(E == 0 || (src && dst && indices && eid && hub_rows), "stag_csx_build: null edge buffers");
  if (ws_bytes < stag_csx_workspace_bytes(E, N) || !ws) {
    set_error("stag_csx_build: workspace %zu < required %zu", ws_bytes, stag_csx_workspace_bytes(E, N));
    return STAG_EWORKSPACE;
  }
  CsxWorkspace w;
  carve(E, N, (char*)ws, &w);
  const int64_t* key64 = by_dst ? dst : src;
  const int64_t* other64 = by_dst ? src : dst;

  if (E > 0 && N > 0 && E <= kSmallMax && N <= kSmallMax && (!items || stag_csx_items_capacity(E, N) <= 1024)) {
    // minibatch-sized graph: one launch, one synchronisation
    const size_t smem = (size_t)(4 * kSmallMax + 4 + 36) * 4 + 1024 * sizeof(int4);
    static bool attr_set = false;
    if (!attr_set) {
      STAG_CUDA(cudaFuncSetAttribute(k_build_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    k_build_small<<<1, kSmallThreads, smem, stream>>>(key64, other64, (int)E, (int)N, indptr, indices, eid, hub_rows,
                                                      hub_seg_ptr, row_order, items, erow, eidf, (int32_t*)w.totals);
    STAG_LAUNCH_CHECK();
    STAG_CUDA(cudaMemcpyAsync(counts_host, w.totals, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    STAG_CUDA(cudaStreamSynchronize(stream));
    return STAG_OK;
  }

  uint32_t *kin = w.keys_a, *vin = w.vals_a, *kout = w.keys_b, *vout = w.vals_b;
  if (E > 0) {
    const int tb = 256;
    k_init_keys<<<(unsigned)((E + tb - 1) / tb), tb, 0, stream>>>(key64, E, kin, vin);
    STAG_LAUNCH_CHECK();
    int bits = 1;
    while (bits < 32 && (1ll << bits) < N) ++bits;
    const int passes = (bits + 7) / 8;
    const int eb = (int)((E + RS_TILE - 1) / RS_TILE);
    for (int p = 0; p < passes; ++p) {
      const int shift = 8 * p;
      k_radix_hist<<<eb, RS_THREADS, 0, stream>>>(kin, E, shift, w.hist, eb);
      STAG_LAUNCH_CHECK();
      k_exclusive_scan<<<1, 1024, 0, stream>>>(w.hist, (int64_t)RS_RADIX * eb, nullptr);
      STAG_LAUNCH_CHECK();
      k_radix_scatter<<<eb, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, E, shift, w.hist, eb);
      STAG_LAUNCH_CHECK();
      uint32_t* t;
      t = kin; kin = kout; kout = t;
      t = vin; vin = vout; vout = t;
    }
  }
  {
    const int tb = 256;
    k_indptr_from_sorted<<<(unsigned)((E + 1 + tb - 1) / tb), tb, 0, stream>>>(kin, E, N, indptr);
    STAG_LAUNCH_CHECK();
    if (E > 0) {
      k_gather_other<<<(unsigned)((E + tb - 1) / tb), tb, 0, stream>>>(vin, kin, other64, E, indices, eid, erow, eidf);
      STAG_LAUNCH_CHECK();
    }
  }
  // row processing order: one 8-bit stable radix pass over the clamped degrees
  if (row_order && N > 0) {
    const int tb = 256;
    const unsigned gb = (unsigned)((N + tb - 1) / tb);
    const int nb = (int)((N + RS_TILE - 1) / RS_TILE);
    k_degree_keys<<<gb, tb, 0, stream>>>(indptr, N, w.keys_a, w.vals_a);
    STAG_LAUNCH_CHECK();
    k_radix_hist<<<nb, RS_THREADS, 0, stream>>>(w.keys_a, N, 0, w.hist, nb);
    STAG_LAUNCH_CHECK();
    k_exclusive_scan<<<1, 1024, 0, stream>>>(w.hist, (int64_t)RS_RADIX * nb, nullptr);
    STAG_LAUNCH_CHECK();
    k_radix_scatter<<<nb, RS_THREADS, 0, stream>>>(w.keys_a, w.vals_a, w.keys_b, w.vals_b, N, 0, w.hist, nb);
    STAG_LAUNCH_CHECK();
    k_reverse_order<<<gb, tb, 0, stream>>>(w.vals_b, N, row_order);
    STAG_LAUNCH_CHECK();
  }
  // hub schedule
  int32_t counts[3] = {0, 0, 0};
  if (N > 0 && E > 0) {
    const int tb = 256;
    const unsigned gb = (unsigned)((N + tb - 1) / tb);
    k_hub_flags<<<gb, tb, 0, stream>>>(indptr, N, w.flags_a, w.flags_b);
    STAG_LAUNCH_CHECK();
    k_exclusive_scan<<<1, 1024, 0, stream>>>(w.flags_a, N, w.totals);
    STAG_LAUNCH_CHECK();
    k_exclusive_scan<<<1, 1024, 0, stream>>>(w.flags_b, N, w.totals + 1);
    STAG_LAUNCH_CHECK();
    k_hub_scatter<<<gb, tb, 0, stream>>>(indptr, N, w.flags_a, w.flags_b, w.totals, hub_rows, hub_seg_ptr);
    STAG_LAUNCH_CHECK();
    STAG_CUDA(cudaMemcpyAsync(counts, w.totals, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  } else {
    STAG_CUDA(cudaMemsetAsync(hub_seg_ptr, 0, sizeof(int32_t), stream));
  }
  // stream items
  if (items && N > 0) {
    const int tb = 256;
    const unsigned gb = (unsigned)((N + tb - 1) / tb);
    k_range_flags<<<gb, tb, 0, stream>>>(indptr, N, range_edges_for(E), w.flags_a);
    STAG_LAUNCH_CHECK();
    k_exclusive_scan<<<1, 1024, 0, stream>>>(w.flags_a, N, w.totals + 2);
    STAG_LAUNCH_CHECK();
    k_range_scatter<<<gb, tb, 0, stream>>>(indptr, N, range_edges_for(E), w.flags_a, items);
    STAG_LAUNCH_CHECK();
    k_range_close<<<gb, tb, 0, stream>>>(indptr, N, w.totals + 2, items);  // #items <= N
    STAG_LAUNCH_CHECK();
    STAG_CUDA(cudaMemcpyAsync(counts + 2, w.totals + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    STAG_CUDA(cudaStreamSynchronize(stream));
    const int64_t n = counts[2];
    if (n > 1) {  // order by decreasing edge count (one stable 8-bit radix pass), ties by decreasing row
      const unsigned ib = (unsigned)((n + tb - 1) / tb);
      const int nb = (int)((n + RS_TILE - 1) / RS_TILE);
      k_item_keys<<<ib, tb, 0, stream>>>(items, n, w.keys_a, w.vals_a);
      STAG_LAUNCH_CHECK();
      k_radix_hist<<<nb, RS_THREADS, 0, stream>>>(w.keys_a, n, 0, w.hist, nb);
      STAG_LAUNCH_CHECK();
      k_exclusive_scan<<<1, 1024, 0, stream>>>(w.hist, (int64_t)RS_RADIX * nb, nullptr);
      STAG_LAUNCH_CHECK();
      k_radix_scatter<<<nb, RS_THREADS, 0, stream>>>(w.keys_a, w.vals_a, w.keys_b, w.vals_b, n, 0, w.hist, nb);
      STAG_LAUNCH_CHECK();
      STAG_CUDA(cudaMemcpyAsync(w.items_tmp, items, (size_t)n * 16, cudaMemcpyDeviceToDevice, stream));
      k_item_permute<<<ib, tb, 0, stream>>>((const int4*)w.items_tmp, w.vals_b, n, (int4*)items);
      STAG_LAUNCH_CHECK();
    }
  }
  STAG_CUDA(cudaStreamSynchronize(stream));
  counts_host[0] = counts[0];
  counts_host[1] = counts[1];
  counts_host[2] = counts[2];
  return STAG_OK;
}
