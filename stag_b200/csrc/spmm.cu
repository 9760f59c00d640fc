// Fused stochastic neighbour aggregation for sm_100a.
//
// One kernel family does, in a single pass over a compressed adjacency and without ever
// writing the [E,K] noise tensor to HBM:
//   counter-based Philox noise -> reparameterisation (loc/scale, low/high, probs) -> relu
//   -> message scaling -> segmented reduction (+ in-norm, + degree scalings)
// and, in the backward instantiation, the transposed aggregation (dX) together with the
// SDDMM term reduced straight into the noise-parameter gradients.
//
// Reference path replaced (file:line in /root/reference):
//   StagLayer.rsample_noise        stag/layers.py:115-129
//   relu / _in_norm                stag/layers.py:98-105, 8-36
//   update_all(u_mul_e, sum|mean)  stag/zoo/gcn.py:63,95  stag/zoo/graph_sage.py:57,72,86
//   degree scalings                stag/zoo/gcn.py:67-75,100-108
//   autograd of the above          DGL GSpMM.backward (gspmm on the reverse graph + gsddmm)
//
// Mapping: a row (destination node for CSC, source node for CSR) is owned by a group of
// LPR lanes (LPR = 32 for D >= 128; narrower rows pack 32/LPR rows into a warp); each lane
// owns one channel quad, i.e. one 128-bit feature load and one Philox block per edge.
// Rows longer than kHubThreshold are cut into segments that are scheduled as independent
// work items and combined in a fixed order by a finalize kernel (deterministic).
#include "common.cuh"
#include "noise.cuh"

namespace stag {

constexpr int AGG_THREADS = 256;
constexpr int AGG_WARPS = AGG_THREADS / 32;
constexpr int AGG_UNROLL = 4;

struct AggParams {
  // structure
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eid;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int num_hubs, num_hub_segs;
  int N;
  int64_t E;
  // gathered operand (x[indices[j]]), its per-node scale, row-side scale
  const float* x;
  int64_t ldx, x_ss;
  const float* gscale;
  const float* rscale;
  float* out;
  int64_t ldo, out_ss;
  int D, S, nq;
  int lpr_log2;
  // noise
  int K, pshape, relu, in_norm, sample_base;
  const float* p0;
  const float* p1;
  const float* ext;
  PhiloxKey key;
  float* norm_scale_out;
  // hub partial sums [S][num_hub_segs][nq*4]
  float* part_acc;
  float* part_w;
  // gradient mode
  const float* xrow;
  int64_t ldxr, xr_ss;
  float* dp0;
  float* dp1;
  float* dw_ext;
  float* dp_partial;  // [grid][2][nq*4]
};

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <bool ALIGNED>
__device__ __forceinline__ float4 load4(const float* __restrict__ row, int c, int D) {
  if (ALIGNED) {
    return __ldg(reinterpret_cast<const float4*>(row + c));
  } else {
    float4 v = f4zero();
    if (c + 0 < D) v.x = __ldg(row + c + 0);
    if (c + 1 < D) v.y = __ldg(row + c + 1);
    if (c + 2 < D) v.z = __ldg(row + c + 2);
    if (c + 3 < D) v.w = __ldg(row + c + 3);
    return v;
  }
}

template <bool ALIGNED>
__device__ __forceinline__ void store4(float* __restrict__ row, int c, int D, float4 v) {
  if (ALIGNED) {
    *reinterpret_cast<float4*>(row + c) = v;
  } else {
    if (c + 0 < D) row[c + 0] = v.x;
    if (c + 1 < D) row[c + 1] = v.y;
    if (c + 2 < D) row[c + 2] = v.z;
    if (c + 3 < D) row[c + 3] = v.w;
  }
}

__device__ __forceinline__ float4 bcast4(float v) { return make_float4(v, v, v, v); }

// Parameter quad for (edge e, channel c).  pshape EDGE* are read per edge.
template <bool ALIGNED>
__device__ __forceinline__ float4 edge_param(const float* __restrict__ p, int pshape, int K, int64_t e, int c) {
  if (pshape == STAG_PARAM_EDGE || K == 1) return bcast4(__ldg(p + e));
  return load4<ALIGNED>(p + e * (int64_t)K, c, K);
}

// Noise quad for one (edge, quad, sample).  `raw` returns the un-transformed variate
// (standard normal / uniform) and `pre` the value before relu; both are only needed by
// the gradient instantiation.
template <int KIND, bool ALIGNED>
__device__ __forceinline__ float4 noise_quad(const AggParams& p, int64_t e, int c, int s_local,
                                             const float4& P0, const float4& P1, float4& raw, float4& pre) {
  float4 w;
  if (KIND == STAG_NOISE_NONE) {
    w = bcast4(1.0f);
    raw = w;
    pre = w;
    return w;
  }
  if (KIND == STAG_NOISE_EXTERNAL) {
    const float* base = p.ext + (int64_t)s_local * p.E * p.K;
    if (p.K == 1) w = bcast4(__ldg(base + e));
    else w = load4<ALIGNED>(base + e * (int64_t)p.K, c, p.K);
    raw = w;
  } else {
    const uint32_t q = p.K == 1 ? 0u : (uint32_t)(c >> 2);
    raw = raw_variates<KIND>((uint32_t)e, q, (uint32_t)(p.sample_base + s_local), p.key);
    if (p.K == 1) raw = bcast4(raw.x);
    float4 a = P0, b = P1;
    if (p.pshape >= STAG_PARAM_EDGE) {
      a = edge_param<ALIGNED>(p.p0, p.pshape, p.K, e, c);
      if (KIND != STAG_NOISE_BERNOULLI) b = edge_param<ALIGNED>(p.p1, p.pshape, p.K, e, c);
    }
    w.x = transform<KIND>(raw.x, a.x, b.x);
    w.y = transform<KIND>(raw.y, a.y, b.y);
    w.z = transform<KIND>(raw.z, a.z, b.z);
    w.w = transform<KIND>(raw.w, a.w, b.w);
  }
  pre = w;
  if (p.relu) {
    w.x = fmaxf(w.x, 0.f);
    w.y = fmaxf(w.y, 0.f);
    w.z = fmaxf(w.z, 0.f);
    w.w = fmaxf(w.w, 0.f);
  }
  return w;
}

__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int KIND, bool ALIGNED, bool GRADS>
__global__ void __launch_bounds__(AGG_THREADS) agg_kernel(const AggParams p) {
  extern __shared__ float smem[];  // GRADS: [AGG_WARPS][2][nq*4]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int LPR = 1 << p.lpr_log2;
  const int RPW = 32 >> p.lpr_log2;
  const int sub = lane >> p.lpr_log2;
  const int sl = lane & (LPR - 1);
  const int nq4 = p.nq * 4;
  constexpr bool PARAM_GRADS = GRADS && (KIND == STAG_NOISE_NORMAL || KIND == STAG_NOISE_UNIFORM);

  float* my_sm = nullptr;
  if (PARAM_GRADS) {
    my_sm = smem + (size_t)warp * 2 * nq4;
    for (int i = lane; i < 2 * nq4; i += 32) my_sm[i] = 0.f;
    __syncwarp();
  }

  const int HG = (p.num_hub_segs + RPW - 1) / RPW;
  const int RG = (p.N + RPW - 1) / RPW;
  const int64_t per_sample = (int64_t)HG + RG;
  const int64_t total = per_sample * p.S;
  const int64_t total_warps = (int64_t)gridDim.x * AGG_WARPS;

  for (int64_t item = (int64_t)blockIdx.x * AGG_WARPS + warp; item < total; item += total_warps) {
    const int s = (int)(item / per_sample);
    const int r = (int)(item - (int64_t)s * per_sample);
    // resolve this lane-group's row and edge range
    int row = -1, beg = 0, len = 0, part_slot = -1;
    if (r < HG) {
      const int seg = r * RPW + sub;
      if (seg < p.num_hub_segs) {
        int lo = 0, hi = p.num_hubs;  // last hub with hub_seg_ptr[h] <= seg
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(p.hub_seg_ptr + mid) <= seg) lo = mid; else hi = mid;
        }
        row = __ldg(p.hub_rows + lo);
        const int k = seg - __ldg(p.hub_seg_ptr + lo);
        const int rb = __ldg(p.indptr + row), re = __ldg(p.indptr + row + 1);
        beg = rb + k * kHubSegment;
        len = min(kHubSegment, re - beg);
        part_slot = seg;
      }
    } else {
      const int v = (r - HG) * RPW + sub;
      if (v < p.N) {
        const int rb = __ldg(p.indptr + v), re = __ldg(p.indptr + v + 1);
        if (re - rb <= kHubThreshold) {
          row = v;
          beg = rb;
          len = re - rb;
        }
      }
    }
    int maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    int anyrow = row >= 0;
    anyrow = __any_sync(0xffffffffu, anyrow);
    if (!anyrow) continue;

    const float* xs = p.x + (int64_t)s * p.x_ss;
    const float rs = (row >= 0 && p.rscale) ? __ldg(p.rscale + row) : 1.0f;

    for (int c0 = 0; c0 < nq4; c0 += LPR * 4) {
      const int c = c0 + sl * 4;
      const bool qvalid = c < p.D;
      float4 P0 = f4zero(), P1 = f4zero();
      if (KIND >= STAG_NOISE_NORMAL && p.pshape <= STAG_PARAM_CHANNEL) {
        if (p.pshape == STAG_PARAM_SCALAR || p.K == 1) {
          P0 = bcast4(__ldg(p.p0));
          if (p.p1) P1 = bcast4(__ldg(p.p1));
        } else if (qvalid) {
          P0 = load4<ALIGNED>(p.p0, c, p.D);
          if (p.p1) P1 = load4<ALIGNED>(p.p1, c, p.D);
        }
      }
      float4 acc = f4zero(), wsum = f4zero();
      float4 xr = f4zero(), d0 = f4zero(), d1 = f4zero();
      if (GRADS && row >= 0 && qvalid) {
        xr = load4<ALIGNED>(p.xrow + (int64_t)s * p.xr_ss + (int64_t)row * p.ldxr, c, p.D);
        xr.x *= rs; xr.y *= rs; xr.z *= rs; xr.w *= rs;
      }

      for (int off = 0; off < maxlen; off += LPR) {
        int my_idx = 0, my_eid = 0;
        float my_sc = 0.f;
        if (off + sl < len) {
          my_idx = __ldg(p.indices + beg + off + sl);
          my_eid = __ldg(p.eid + beg + off + sl);
          my_sc = p.gscale ? __ldg(p.gscale + my_idx) : 1.0f;
        }
        const int cntmax = min(LPR, maxlen - off);
        for (int t0 = 0; t0 < cntmax; t0 += AGG_UNROLL) {
          float4 xv[AGG_UNROLL];
          int ee[AGG_UNROLL];
          float sc[AGG_UNROLL];
          bool act[AGG_UNROLL];
#pragma unroll
          for (int k = 0; k < AGG_UNROLL; ++k) {
            const int t = t0 + k;
            const int u = __shfl_sync(0xffffffffu, my_idx, t, LPR);
            ee[k] = __shfl_sync(0xffffffffu, my_eid, t, LPR);
            sc[k] = __shfl_sync(0xffffffffu, my_sc, t, LPR);
            act[k] = (t < LPR) && (off + t < len) && qvalid;
            xv[k] = act[k] ? load4<ALIGNED>(xs + (int64_t)u * p.ldx, c, p.D) : f4zero();
          }
#pragma unroll
          for (int k = 0; k < AGG_UNROLL; ++k) {
            if (t0 + k >= cntmax) break;  // warp-uniform
            float4 raw, pre;
            float4 w = f4zero();
            if (act[k]) w = noise_quad<KIND, ALIGNED>(p, ee[k], c, s, P0, P1, raw, pre);
            else { raw = f4zero(); pre = f4zero(); }
            const float4 g = make_float4(xv[k].x * sc[k], xv[k].y * sc[k], xv[k].z * sc[k], xv[k].w * sc[k]);
            acc.x = fmaf(w.x, g.x, acc.x);
            acc.y = fmaf(w.y, g.y, acc.y);
            acc.z = fmaf(w.z, g.z, acc.z);
            acc.w = fmaf(w.w, g.w, acc.w);
            if (!GRADS) {
              wsum.x += w.x; wsum.y += w.y; wsum.z += w.z; wsum.w += w.w;
            } else {
              float4 dw = make_float4(xr.x * g.x, xr.y * g.y, xr.z * g.z, xr.w * g.w);
              if (KIND == STAG_NOISE_EXTERNAL) {
                if (p.dw_ext) {
                  float* base = p.dw_ext + (int64_t)s * p.E * p.K;
                  if (p.K == 1) {
                    const float tot = group_sum(dw.x + dw.y + dw.z + dw.w, LPR);
                    // several channel chunks accumulate into the same slot, from the same thread
                    if (sl == 0 && (t0 + k < LPR) && (off + t0 + k < len)) {
                      if (c0 == 0) base[ee[k]] = tot; else base[ee[k]] += tot;
                    }
                  } else if (act[k]) {
                    store4<ALIGNED>(base + (int64_t)ee[k] * p.K, c, p.K, dw);
                  }
                }
              } else if (PARAM_GRADS) {
                if (p.relu) {
                  dw.x = pre.x > 0.f ? dw.x : 0.f;
                  dw.y = pre.y > 0.f ? dw.y : 0.f;
                  dw.z = pre.z > 0.f ? dw.z : 0.f;
                  dw.w = pre.w > 0.f ? dw.w : 0.f;
                }
                float4 e0, e1;
                if (KIND == STAG_NOISE_NORMAL) {
                  e0 = dw;
                  e1 = make_float4(dw.x * raw.x, dw.y * raw.y, dw.z * raw.z, dw.w * raw.w);
                } else {
                  e1 = make_float4(dw.x * raw.x, dw.y * raw.y, dw.z * raw.z, dw.w * raw.w);
                  e0 = make_float4(dw.x - e1.x, dw.y - e1.y, dw.z - e1.z, dw.w - e1.w);
                }
                if (p.pshape <= STAG_PARAM_CHANNEL) {
                  d0.x += e0.x; d0.y += e0.y; d0.z += e0.z; d0.w += e0.w;
                  d1.x += e1.x; d1.y += e1.y; d1.z += e1.z; d1.w += e1.w;
                } else if (p.pshape == STAG_PARAM_EDGE || p.K == 1) {
                  const float t0s = group_sum(e0.x + e0.y + e0.z + e0.w, LPR);
                  const float t1s = group_sum(e1.x + e1.y + e1.z + e1.w, LPR);
                  if (sl == 0 && (t0 + k < LPR) && (off + t0 + k < len)) {
                    p.dp0[ee[k]] += t0s;
                    p.dp1[ee[k]] += t1s;
                  }
                } else if (act[k]) {
                  float* q0 = p.dp0 + (int64_t)ee[k] * p.K;
                  float* q1 = p.dp1 + (int64_t)ee[k] * p.K;
                  float4 o0 = load4<ALIGNED>(q0, c, p.K), o1 = load4<ALIGNED>(q1, c, p.K);
                  o0.x += e0.x; o0.y += e0.y; o0.z += e0.z; o0.w += e0.w;
                  o1.x += e1.x; o1.y += e1.y; o1.z += e1.z; o1.w += e1.w;
                  store4<ALIGNED>(q0, c, p.K, o0);
                  store4<ALIGNED>(q1, c, p.K, o1);
                }
              }
            }
          }
        }
      }

      // epilogue for this channel chunk
      if (row >= 0 && qvalid) {
        if (part_slot >= 0) {
          if (p.out) {
            const int64_t o = ((int64_t)s * p.num_hub_segs + part_slot) * nq4 + c;
            *reinterpret_cast<float4*>(p.part_acc + o) = acc;
            if (!GRADS && p.in_norm) *reinterpret_cast<float4*>(p.part_w + o) = wsum;
          }
        } else if (p.out) {
          if (!GRADS && p.in_norm) {
            const float indeg = (float)len;
            float4 sc4;
            sc4.x = wsum.x != 0.f ? indeg / wsum.x : 1.f;
            sc4.y = wsum.y != 0.f ? indeg / wsum.y : 1.f;
            sc4.z = wsum.z != 0.f ? indeg / wsum.z : 1.f;
            sc4.w = wsum.w != 0.f ? indeg / wsum.w : 1.f;
            acc.x *= sc4.x; acc.y *= sc4.y; acc.z *= sc4.z; acc.w *= sc4.w;
            if (p.norm_scale_out) {
              float* ns = p.norm_scale_out + ((int64_t)s * p.N + row) * p.K;
              if (p.K == 1) { if (c == 0) ns[0] = sc4.x; }
              else store4<ALIGNED>(ns, c, p.K, sc4);
            }
          }
          acc.x *= rs; acc.y *= rs; acc.z *= rs; acc.w *= rs;
          store4<ALIGNED>(p.out + (int64_t)s * p.out_ss + (int64_t)row * p.ldo, c, p.D, acc);
        }
      }
      if (PARAM_GRADS && p.pshape <= STAG_PARAM_CHANNEL) {
        // fold the row groups of this warp, then add into the warp's shared slice
        for (int o = LPR; o < 32; o <<= 1) {
          d0.x += __shfl_xor_sync(0xffffffffu, d0.x, o);
          d0.y += __shfl_xor_sync(0xffffffffu, d0.y, o);
          d0.z += __shfl_xor_sync(0xffffffffu, d0.z, o);
          d0.w += __shfl_xor_sync(0xffffffffu, d0.w, o);
          d1.x += __shfl_xor_sync(0xffffffffu, d1.x, o);
          d1.y += __shfl_xor_sync(0xffffffffu, d1.y, o);
          d1.z += __shfl_xor_sync(0xffffffffu, d1.z, o);
          d1.w += __shfl_xor_sync(0xffffffffu, d1.w, o);
        }
        if (sub == 0 && c < nq4) {
          float4* a0 = reinterpret_cast<float4*>(my_sm + c);
          float4* a1 = reinterpret_cast<float4*>(my_sm + nq4 + c);
          float4 v0 = *a0, v1 = *a1;
          v0.x += d0.x; v0.y += d0.y; v0.z += d0.z; v0.w += d0.w;
          v1.x += d1.x; v1.y += d1.y; v1.z += d1.z; v1.w += d1.w;
          *a0 = v0;
          *a1 = v1;
        }
        __syncwarp();
      }
    }
  }

  if (PARAM_GRADS && p.pshape <= STAG_PARAM_CHANNEL) {
    __syncthreads();
    float* dst = p.dp_partial + (size_t)blockIdx.x * 2 * nq4;
    for (int i = threadIdx.x; i < 2 * nq4; i += AGG_THREADS) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < AGG_WARPS; ++w) v += smem[(size_t)w * 2 * nq4 + i];
      dst[i] = v;
    }
  }
}

// Combine the partial sums of hub rows in segment order and finish the row.
__global__ void hub_finalize_kernel(const AggParams p, int grads) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nq4 = p.nq * 4;
  const int64_t total = (int64_t)p.S * p.num_hubs * p.D;
  if (idx >= total) return;
  const int c = (int)(idx % p.D);
  const int h = (int)((idx / p.D) % p.num_hubs);
  const int s = (int)(idx / ((int64_t)p.D * p.num_hubs));
  const int row = p.hub_rows[h];
  const int s0 = p.hub_seg_ptr[h], s1 = p.hub_seg_ptr[h + 1];
  float acc = 0.f, wsum = 0.f;
  for (int seg = s0; seg < s1; ++seg) {
    const int64_t o = ((int64_t)s * p.num_hub_segs + seg) * nq4 + c;
    acc += p.part_acc[o];
    if (!grads && p.in_norm) wsum += p.part_w[o];
  }
  if (!grads && p.in_norm) {
    const float indeg = (float)(p.indptr[row + 1] - p.indptr[row]);
    const float sc = wsum != 0.f ? indeg / wsum : 1.f;
    acc *= sc;
    if (p.norm_scale_out && (p.K != 1 || c == 0))
      p.norm_scale_out[((int64_t)s * p.N + row) * p.K + (p.K == 1 ? 0 : c)] = sc;
  }
  if (p.rscale) acc *= p.rscale[row];
  if (p.out) p.out[(int64_t)s * p.out_ss + (int64_t)row * p.ldo + c] = acc;
}

// Reduce per-CTA parameter-gradient partials in CTA order (deterministic).
__global__ void param_finalize_kernel(const float* __restrict__ partial, int ncta, int nq4, int D, int scalar,
                                      float* __restrict__ dp0, float* __restrict__ dp1) {
  __shared__ float red[2][256];
  const int tid = threadIdx.x;
  if (!scalar) {
    const int c = blockIdx.x * blockDim.x + tid;
    if (c >= D) return;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < ncta; ++k) {
      a += partial[(size_t)k * 2 * nq4 + c];
      b += partial[(size_t)k * 2 * nq4 + nq4 + c];
    }
    dp0[c] = a;
    dp1[c] = b;
  } else {
    float a = 0.f, b = 0.f;
    for (int c = tid; c < D; c += blockDim.x) {
      for (int k = 0; k < ncta; ++k) {
        a += partial[(size_t)k * 2 * nq4 + c];
        b += partial[(size_t)k * 2 * nq4 + nq4 + c];
      }
    }
    red[0][tid] = a;
    red[1][tid] = b;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (tid < o) {
        red[0][tid] += red[0][tid + o];
        red[1][tid] += red[1][tid + o];
      }
      __syncthreads();
    }
    if (tid == 0) {
      dp0[0] = red[0][0];
      dp1[0] = red[1][0];
    }
  }
}

// noise materialisation (compat path + RNG tests): w[s,e,c]
template <int KIND>
__global__ void emit_kernel(const AggParams p, float* __restrict__ w_out, float* __restrict__ eps_out) {
  const int nq = p.nq;
  const int64_t total = (int64_t)p.S * p.E * nq;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % nq);
    const int64_t e = (i / nq) % p.E;
    const int s = (int)(i / ((int64_t)nq * p.E));
    const int c = q * 4;
    float4 P0 = f4zero(), P1 = f4zero();
    if (p.pshape <= STAG_PARAM_CHANNEL) {
      if (p.pshape == STAG_PARAM_SCALAR || p.K == 1) {
        P0 = bcast4(p.p0[0]);
        if (p.p1) P1 = bcast4(p.p1[0]);
      } else {
        P0 = load4<false>(p.p0, c, p.K);
        if (p.p1) P1 = load4<false>(p.p1, c, p.K);
      }
    }
    float4 raw, pre;
    const float4 w = noise_quad<KIND, false>(p, e, c, s, P0, P1, raw, pre);
    const int64_t o = ((int64_t)s * p.E + e) * p.K;
    if (p.K == 1) {
      w_out[o] = w.x;
      if (eps_out) eps_out[o] = raw.x;
    } else {
      store4<false>(w_out + o, c, p.K, w);
      if (eps_out) store4<false>(eps_out + o, c, p.K, raw);
    }
  }
}

__global__ void segment_reduce_kernel(const float* __restrict__ feat, int64_t ldf, const int32_t* __restrict__ ptr,
                                      int B, int D, int mean, float* __restrict__ out, int64_t ldo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * D) return;
  const int c = (int)(idx % D);
  const int b = (int)(idx / D);
  const int n0 = ptr[b], n1 = ptr[b + 1];
  float acc = 0.f;
  for (int n = n0; n < n1; ++n) acc += feat[(int64_t)n * ldf + c];
  if (mean) acc /= (float)max(n1 - n0, 1);
  out[(int64_t)b * ldo + c] = acc;
}

static int lpr_log2_for(int nq) {
  int l = 0;
  while ((1 << l) < nq && l < 5) ++l;
  return l;
}

struct WsLayout {
  size_t part_acc, part_w, dp_partial, total;
};

static WsLayout ws_layout(const StagGraph* g, int D, int S, int grid_max) {
  WsLayout L;
  const size_t nq4 = (size_t)((D + 3) / 4) * 4;
  size_t off = 0;
  L.part_acc = off;
  off += align_up((size_t)S * g->num_hub_segs * nq4 * 4 + 16, 256);
  L.part_w = off;
  off += align_up((size_t)S * g->num_hub_segs * nq4 * 4 + 16, 256);
  L.dp_partial = off;
  off += align_up((size_t)grid_max * 2 * nq4 * 4 + 16, 256);
  L.total = off;
  return L;
}

static int grid_cap() { return num_sms() * 8; }

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int check_noise(const StagNoise* n, int D, const char* who) {
  STAG_CHECK_ARG(n != nullptr, "%s: null noise spec", who);
  STAG_CHECK_ARG(n->kind >= STAG_NOISE_NONE && n->kind <= STAG_NOISE_BERNOULLI, "%s: bad noise kind %d", who, n->kind);
  if (n->kind == STAG_NOISE_NONE) return STAG_OK;
  STAG_CHECK_ARG(n->K == 1 || n->K == D, "%s: noise width K=%d must be 1 or D=%d", who, n->K, D);
  if (n->kind == STAG_NOISE_EXTERNAL) {
    STAG_CHECK_ARG(n->external != nullptr, "%s: EXTERNAL noise needs a tensor", who);
    return STAG_OK;
  }
  STAG_CHECK_ARG(n->param_shape >= STAG_PARAM_SCALAR && n->param_shape <= STAG_PARAM_EDGE_CHANNEL,
                 "%s: bad param_shape %d", who, n->param_shape);
  STAG_CHECK_ARG(n->p0 != nullptr, "%s: null parameter p0", who);
  STAG_CHECK_ARG(n->kind == STAG_NOISE_BERNOULLI || n->p1 != nullptr, "%s: null parameter p1", who);
  return STAG_OK;
}


template <int KIND, bool GRADS>
static int launch_agg_kind(const AggParams& p, bool aligned, int grid, size_t smem, cudaStream_t stream) {
  if (aligned) {
    if (smem > 48 * 1024)
      STAG_CUDA(cudaFuncSetAttribute(agg_kernel<KIND, true, GRADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    agg_kernel<KIND, true, GRADS><<<grid, AGG_THREADS, smem, stream>>>(p);
  } else {
    if (smem > 48 * 1024)
      STAG_CUDA(cudaFuncSetAttribute(agg_kernel<KIND, false, GRADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    agg_kernel<KIND, false, GRADS><<<grid, AGG_THREADS, smem, stream>>>(p);
  }
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}

template <bool GRADS>
static int launch_agg(int kind, const AggParams& p, bool aligned, int grid, size_t smem, cudaStream_t stream) {
  switch (kind) {
    case STAG_NOISE_NONE: return launch_agg_kind<STAG_NOISE_NONE, GRADS>(p, aligned, grid, smem, stream);
    case STAG_NOISE_EXTERNAL: return launch_agg_kind<STAG_NOISE_EXTERNAL, GRADS>(p, aligned, grid, smem, stream);
    case STAG_NOISE_NORMAL: return launch_agg_kind<STAG_NOISE_NORMAL, GRADS>(p, aligned, grid, smem, stream);
    case STAG_NOISE_UNIFORM: return launch_agg_kind<STAG_NOISE_UNIFORM, GRADS>(p, aligned, grid, smem, stream);
    case STAG_NOISE_BERNOULLI: return launch_agg_kind<STAG_NOISE_BERNOULLI, GRADS>(p, aligned, grid, smem, stream);
  }
  set_error("launch_agg: bad kind %d", kind);
  return STAG_EINVAL;
}

static void fill_noise(AggParams& p, const StagNoise* n, int D) {
  p.K = n->kind == STAG_NOISE_NONE ? D : n->K;
  p.pshape = n->param_shape;
  if (p.K == 1 && p.pshape == STAG_PARAM_CHANNEL) p.pshape = STAG_PARAM_SCALAR;
  if (p.K == 1 && p.pshape == STAG_PARAM_EDGE_CHANNEL) p.pshape = STAG_PARAM_EDGE;
  if (n->kind <= STAG_NOISE_EXTERNAL) p.pshape = STAG_PARAM_SCALAR;
  p.relu = n->relu;
  p.in_norm = n->in_norm;
  p.sample_base = n->sample_base;
  p.p0 = n->p0;
  p.p1 = n->p1;
  p.ext = n->external;
  p.key = make_key(n->seed, n->offset);
}

static void fill_graph(AggParams& p, const StagGraph* g) {
  p.indptr = g->indptr;
  p.indices = g->indices;
  p.eid = g->eid;
  p.hub_rows = g->hub_rows;
  p.hub_seg_ptr = g->hub_seg_ptr;
  p.num_hubs = g->num_hubs;
  p.num_hub_segs = g->num_hub_segs;
  p.N = (int)g->num_rows;
  p.E = g->num_edges;
}

static int check_graph(const StagGraph* g, const char* who) {
  STAG_CHECK_ARG(g != nullptr, "%s: null graph", who);
  STAG_CHECK_ARG(g->num_rows >= 0 && g->num_rows < (1ll << 31) && g->num_edges >= 0 && g->num_edges < (1ll << 31),
                 "%s: graph sizes out of range", who);
  STAG_CHECK_ARG(g->indptr != nullptr, "%s: null indptr", who);
  STAG_CHECK_ARG(g->num_edges == 0 || (g->indices && g->eid), "%s: null indices/eid", who);
  STAG_CHECK_ARG(g->num_hubs == 0 || (g->hub_rows && g->hub_seg_ptr), "%s: null hub schedule", who);
  return STAG_OK;
}

static int agg_grid(const AggParams& p) {
  const int RPW = 32 >> p.lpr_log2;
  const int64_t per_sample = (int64_t)(p.num_hub_segs + RPW - 1) / RPW + (p.N + RPW - 1) / RPW;
  const int64_t ctas = (per_sample * p.S + AGG_WARPS - 1) / AGG_WARPS;
  const int64_t cap = grid_cap();
  return (int)(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

}  // namespace stag

using namespace stag;

extern "C" size_t stag_spmm_workspace_bytes(const StagGraph* g, int32_t D, int32_t S) {
  if (!g || D <= 0 || S <= 0) return 0;
  return ws_layout(g, D, S, grid_cap()).total;
}

extern "C" int stag_spmm_fwd(const StagGraph* g, const float* x, int64_t ldx, int64_t x_sample_stride, int32_t D,
                             int32_t S, const StagNoise* noise, const float* src_scale, const float* dst_scale,
                             float* out, int64_t ldo, int64_t out_sample_stride, float* norm_scale_out, void* ws,
                             size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_graph(g, "stag_spmm_fwd");
  if (rc) return rc;
  STAG_CHECK_ARG(D > 0 && S > 0, "stag_spmm_fwd: D=%d S=%d must be positive", D, S);
  STAG_CHECK_ARG(out != nullptr, "stag_spmm_fwd: null output");
  STAG_CHECK_ARG(x != nullptr || g->num_edges == 0, "stag_spmm_fwd: null features");
  STAG_CHECK_ARG(ldx >= D && ldo >= D, "stag_spmm_fwd: row strides smaller than D");
  rc = check_noise(noise, D, "stag_spmm_fwd");
  if (rc) return rc;
  if (g->num_rows == 0) return STAG_OK;
  const WsLayout L = ws_layout(g, D, S, grid_cap());
  if (g->num_hub_segs > 0 && (!ws || ws_bytes < L.total)) {
    set_error("stag_spmm_fwd: workspace %zu < required %zu", ws_bytes, L.total);
    return STAG_EWORKSPACE;
  }
  AggParams p = {};
  fill_graph(p, g);
  fill_noise(p, noise, D);
  p.x = x; p.ldx = ldx; p.x_ss = x_sample_stride;
  p.gscale = src_scale; p.rscale = dst_scale;
  p.out = out; p.ldo = ldo; p.out_ss = out_sample_stride;
  p.D = D; p.S = S; p.nq = (D + 3) / 4;
  p.lpr_log2 = lpr_log2_for(p.nq);
  p.norm_scale_out = norm_scale_out;
  if (g->num_hub_segs > 0) {
    p.part_acc = (float*)((char*)ws + L.part_acc);
    p.part_w = (float*)((char*)ws + L.part_w);
  }
  bool aligned = (D % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && (x_sample_stride % 4 == 0) &&
                 (out_sample_stride % 4 == 0) && aligned16(x) && aligned16(out);
  if (noise->kind == STAG_NOISE_EXTERNAL && noise->K != 1) aligned = aligned && aligned16(noise->external);
  if (noise->kind >= STAG_NOISE_NORMAL && noise->param_shape != STAG_PARAM_SCALAR && noise->K != 1)
    aligned = aligned && aligned16(noise->p0) && (noise->p1 == nullptr || aligned16(noise->p1));
  if (norm_scale_out && noise->K != 1) aligned = aligned && aligned16(norm_scale_out);
  rc = launch_agg<false>(noise->kind, p, aligned, agg_grid(p), 0, stream);
  if (rc) return rc;
  if (g->num_hubs > 0) {
    const int64_t total = (int64_t)S * g->num_hubs * D;
    hub_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, 0);
    STAG_LAUNCH_CHECK();
  }
  return STAG_OK;
}

extern "C" int stag_spmm_bwd(const StagGraph* g, const float* x, int64_t ldx, int64_t x_sample_stride,
                             const float* dout, int64_t ldg, int64_t dout_sample_stride, int32_t D, int32_t S,
                             const StagNoise* noise, const float* src_scale, const float* dst_scale, float* dx,
                             int64_t lddx, int64_t dx_sample_stride, float* dparam0, float* dparam1,
                             float* dw_external, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_graph(g, "stag_spmm_bwd");
  if (rc) return rc;
  STAG_CHECK_ARG(D > 0 && S > 0, "stag_spmm_bwd: D=%d S=%d must be positive", D, S);
  STAG_CHECK_ARG(dout != nullptr || g->num_edges == 0, "stag_spmm_bwd: null upstream gradient");
  STAG_CHECK_ARG(x != nullptr || g->num_edges == 0, "stag_spmm_bwd: null features");
  STAG_CHECK_ARG(ldx >= D && ldg >= D && (dx == nullptr || lddx >= D), "stag_spmm_bwd: row strides smaller than D");
  rc = check_noise(noise, D, "stag_spmm_bwd");
  if (rc) return rc;
  if (noise->in_norm) {
    set_error("stag_spmm_bwd: in_norm has no fused parameter-gradient path (use the emitted-noise path)");
    return STAG_EUNSUPPORTED;
  }
  const bool param_grads = noise->kind == STAG_NOISE_NORMAL || noise->kind == STAG_NOISE_UNIFORM;
  if (param_grads) STAG_CHECK_ARG(dparam0 && dparam1, "stag_spmm_bwd: null parameter-gradient outputs");
  AggParams p = {};
  fill_graph(p, g);
  fill_noise(p, noise, D);
  const bool edge_params = param_grads && p.pshape >= STAG_PARAM_EDGE;
  STAG_CHECK_ARG(!edge_params || S == 1, "stag_spmm_bwd: per-edge parameter gradients require S == 1 (got %d)", S);
  if (g->num_rows == 0) {
    if (param_grads && !edge_params) {
      const size_t n = p.pshape == STAG_PARAM_SCALAR ? 1 : (size_t)D;
      STAG_CUDA(cudaMemsetAsync(dparam0, 0, n * 4, stream));
      STAG_CUDA(cudaMemsetAsync(dparam1, 0, n * 4, stream));
    }
    return STAG_OK;
  }
  const WsLayout L = ws_layout(g, D, S, grid_cap());
  if (!ws || ws_bytes < L.total) {
    set_error("stag_spmm_bwd: workspace %zu < required %zu", ws_bytes, L.total);
    return STAG_EWORKSPACE;
  }
  // transposed roles: gather dout rows scaled by dst_scale, rows are sources scaled by src_scale
  p.x = dout; p.ldx = ldg; p.x_ss = dout_sample_stride;
  p.gscale = dst_scale; p.rscale = src_scale;
  p.out = dx; p.ldo = lddx; p.out_ss = dx_sample_stride;
  p.xrow = x; p.ldxr = ldx; p.xr_ss = x_sample_stride;
  p.D = D; p.S = S; p.nq = (D + 3) / 4;
  p.lpr_log2 = lpr_log2_for(p.nq);
  p.dp0 = dparam0; p.dp1 = dparam1; p.dw_ext = dw_external;
  p.part_acc = (float*)((char*)ws + L.part_acc);
  p.part_w = (float*)((char*)ws + L.part_w);
  p.dp_partial = (float*)((char*)ws + L.dp_partial);
  bool aligned = (D % 4 == 0) && (ldx % 4 == 0) && (ldg % 4 == 0) && (x_sample_stride % 4 == 0) &&
                 (dout_sample_stride % 4 == 0) && aligned16(x) && aligned16(dout);
  if (dx) aligned = aligned && (lddx % 4 == 0) && (dx_sample_stride % 4 == 0) && aligned16(dx);
  if (noise->kind == STAG_NOISE_EXTERNAL && noise->K != 1)
    aligned = aligned && aligned16(noise->external) && (dw_external == nullptr || aligned16(dw_external));
  if (param_grads && noise->param_shape != STAG_PARAM_SCALAR && noise->K != 1)
    aligned = aligned && aligned16(noise->p0) && aligned16(noise->p1) && aligned16(dparam0) && aligned16(dparam1);
  const int grid = agg_grid(p);
  const size_t smem = (param_grads && !edge_params) ? (size_t)AGG_WARPS * 2 * p.nq * 4 * sizeof(float) : 0;
  if (smem > 200 * 1024) {
    set_error("stag_spmm_bwd: D=%d too wide for the shared-memory parameter-gradient staging", D);
    return STAG_EUNSUPPORTED;
  }
  rc = launch_agg<true>(noise->kind, p, aligned, grid, smem, stream);
  if (rc) return rc;
  if (g->num_hubs > 0 && dx) {
    const int64_t total = (int64_t)S * g->num_hubs * D;
    hub_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, 1);
    STAG_LAUNCH_CHECK();
  }
  if (param_grads && !edge_params) {
    const int scalar = p.pshape == STAG_PARAM_SCALAR;
    const int blocks = scalar ? 1 : (D + 255) / 256;
    param_finalize_kernel<<<blocks, 256, 0, stream>>>(p.dp_partial, grid, p.nq * 4, D, scalar, dparam0, dparam1);
    STAG_LAUNCH_CHECK();
  }
  return STAG_OK;
}

extern "C" int stag_noise_emit(const StagNoise* noise, int64_t num_edges, int32_t S, float* w_out, float* eps_out,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  STAG_CHECK_ARG(noise != nullptr && w_out != nullptr, "stag_noise_emit: null argument");
  STAG_CHECK_ARG(noise->kind >= STAG_NOISE_NORMAL && noise->kind <= STAG_NOISE_BERNOULLI,
                 "stag_noise_emit: kind %d is not a generated distribution", noise->kind);
  STAG_CHECK_ARG(noise->K > 0 && S > 0 && num_edges >= 0 && num_edges < (1ll << 31), "stag_noise_emit: bad sizes");
  int rc = check_noise(noise, noise->K, "stag_noise_emit");
  if (rc) return rc;
  if (num_edges == 0) return STAG_OK;
  AggParams p = {};
  fill_noise(p, noise, noise->K);
  p.E = num_edges;
  p.S = S;
  p.D = noise->K;
  p.nq = (noise->K + 3) / 4;
  const int64_t total = (int64_t)S * num_edges * p.nq;
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  switch (noise->kind) {
    case STAG_NOISE_NORMAL: emit_kernel<STAG_NOISE_NORMAL><<<grid, 256, 0, stream>>>(p, w_out, eps_out); break;
    case STAG_NOISE_UNIFORM: emit_kernel<STAG_NOISE_UNIFORM><<<grid, 256, 0, stream>>>(p, w_out, eps_out); break;
    default: emit_kernel<STAG_NOISE_BERNOULLI><<<grid, 256, 0, stream>>>(p, w_out, eps_out); break;
  }
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}

extern "C" int stag_segment_reduce(const float* feat, int64_t ldf, const int32_t* node_ptr, int32_t num_graphs,
                                   int32_t D, int mean, float* out, int64_t ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  STAG_CHECK_ARG(num_graphs >= 0 && D > 0, "stag_segment_reduce: bad sizes");
  if (num_graphs == 0) return STAG_OK;
  STAG_CHECK_ARG(feat && node_ptr && out, "stag_segment_reduce: null argument");
  const int64_t total = (int64_t)num_graphs * D;
  segment_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(feat, ldf, node_ptr, num_graphs, D, mean,
                                                                             out, ldo);
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}
