// Sample-based KL terms of StagLayer.kl_divergence without the noise tensor (included by spmm.cu).
//
// Reference: stag/layers.py:139-141 -- when torch has no analytic KL for (q_a, p_a) (a mixture prior, as in
// scripts/citation_rec_contrastive/gcn/run.py:44-52) the layer falls back to
//     q_a.log_prob(w).sum(-1).mean() - p_a.log_prob(w).sum(-1).mean()
// on the STORED sample w [E,K] of the last forward (597 MB per layer and sample at the arxiv shape).  Here the
// sample is regenerated from (seed, offset, edge id, channel) exactly as the fused aggregation kernels drew it and
// reduced on the fly:
//     sums[0] = sum_{s,e,c} log q(w_sec)        sums[1] = sum_{s,e,c} log p(w_sec)
//     dp0, dp1 = d(sums[0] - sums[1]) / d(p0, p1) of q, reduced to the parameter shape (scalar, [K], [E,1], [E,K]),
//                with the reparameterisation path w(p0, p1) included, as autograd differentiates the reference's
//                expression (q: Normal loc/scale, Uniform low/high; relu masks the path).
// One warp walks the edges e = w, w + W, ...; lane l owns the Philox blocks (octs of channels) l, l + 32, ... of every
// edge, so per-channel accumulators belong to one lane of one warp: a fixed summation order, deterministic results
// (per-warp slices in shared memory -> per-CTA partial rows -> ordered finalize in double precision).
#pragma once

namespace stag {

constexpr int KL_THREADS = 256, KL_WARPS = KL_THREADS / 32;
constexpr int KL_MAX_COMPONENTS = STAG_PRIOR_MAX_COMPONENTS;

struct KlParams {
  int64_t E;
  int S, K, nblk, kpad, pshape, relu, sample_base;
  const float* p0;
  const float* p1;
  PhiloxKey key;
  const uint32_t* ctr_dev;  // optional device-side addend to key.c3
  int M;
  float c_m[KL_MAX_COMPONENTS];    // log weight - log sigma - log sqrt(2 pi)
  float mu_m[KL_MAX_COMPONENTS];
  float is_m[KL_MAX_COMPONENTS];   // 1 / sigma
  float* dp0;
  float* dp1;
  double* cta_sums;   // [grid][4]: log q, log p, scalar d p0, scalar d p1
  float* cta_ch;      // [grid][2][kpad] per-channel gradient partials
};

// log p(w) and d log p / dw of the Normal mixture
__device__ __forceinline__ void kl_prior(const KlParams& p, float w, float& lp, float& g) {
  if (p.M == 1) {
    const float t = (w - p.mu_m[0]) * p.is_m[0];
    lp = fmaf(-0.5f * t, t, p.c_m[0]);
    g = -t * p.is_m[0];
    return;
  }
  float a[KL_MAX_COMPONENTS], t[KL_MAX_COMPONENTS], mx = -3.0e38f;
#pragma unroll
  for (int m = 0; m < KL_MAX_COMPONENTS; ++m)
    if (m < p.M) {
      t[m] = (w - p.mu_m[m]) * p.is_m[m];
      a[m] = fmaf(-0.5f * t[m], t[m], p.c_m[m]);
      mx = fmaxf(mx, a[m]);
    }
  float se = 0.f, sg = 0.f;
#pragma unroll
  for (int m = 0; m < KL_MAX_COMPONENTS; ++m)
    if (m < p.M) {
      const float ex = __expf(a[m] - mx);
      se += ex;
      sg = fmaf(ex, -t[m] * p.is_m[m], sg);
    }
  lp = mx + __logf(se);
  g = sg / se;
}

template <int KIND, bool GRADS>
__global__ void __launch_bounds__(KL_THREADS) noise_kl_kernel(const KlParams p) {
  extern __shared__ float kl_ch[];  // [KL_WARPS][2][kpad], channel-shaped gradients only
  __shared__ double red[KL_WARPS][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PhiloxKey key = p.key;
  if (p.ctr_dev) key.c3 += __ldg(p.ctr_dev);
  const bool ch_grads = GRADS && p.pshape == STAG_PARAM_CHANNEL;
  float* my_ch = kl_ch + (size_t)warp * 2 * p.kpad;
  if (ch_grads)
    for (int i = lane; i < 2 * p.kpad; i += 32) my_ch[i] = 0.f;
  double lq_t = 0.0, lp_t = 0.0, d0_t = 0.0, d1_t = 0.0;
  const int64_t nwarps = (int64_t)gridDim.x * KL_WARPS;
  for (int64_t e = (int64_t)blockIdx.x * KL_WARPS + warp; e < p.E; e += nwarps) {
    float lq_e = 0.f, lp_e = 0.f, d0_e = 0.f, d1_e = 0.f;  // this lane's share of the edge
    for (int j = lane; j < p.nblk; j += 32) {
      const int c = first_chan(0, j);
      float P0[8], P1[8], a0[8], a1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ch = chan(c, i);
        int64_t pi = 0;
        if (p.pshape == STAG_PARAM_CHANNEL) pi = ch < p.K ? ch : 0;
        else if (p.pshape == STAG_PARAM_EDGE) pi = e;
        else if (p.pshape == STAG_PARAM_EDGE_CHANNEL) pi = ch < p.K ? e * p.K + ch : 0;
        P0[i] = __ldg(p.p0 + pi);
        P1[i] = __ldg(p.p1 + pi);
        a0[i] = a1[i] = 0.f;
      }
      for (int s = 0; s < p.S; ++s) {
        float raw[8];
        raw_oct<KIND>((uint32_t)e, (uint32_t)j, (uint32_t)(p.sample_base + s), key, raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (chan(c, i) >= p.K) continue;
          float w, lq, dq0, dq1, dw0, dw1;  // log q, its explicit parameter derivatives, dw/dp0, dw/dp1
          if (KIND == STAG_NOISE_NORMAL) {
            const float inv = 1.0f / P1[i];
            const float wp = fmaf(raw[i], P1[i], P0[i]);
            const bool on = !p.relu || wp > 0.f;
            w = on ? wp : 0.f;
            const float z = (w - P0[i]) * inv;
            lq = -__logf(P1[i]) - 0.9189385332046727f - 0.5f * z * z;
            const float glq = -z * inv;  // d log q / dw
            dw0 = on ? 1.0f : 0.f;
            dw1 = on ? raw[i] : 0.f;
            dq0 = fmaf(glq, dw0, z * inv);
            dq1 = fmaf(glq, dw1, fmaf(z * z, inv, -inv));
          } else {  // Uniform(low = P0, high = P1)
            const float width = P1[i] - P0[i];
            const float wp = fmaf(raw[i], width, P0[i]);
            const bool on = !p.relu || wp > 0.f;
            w = on ? wp : 0.f;
            const bool inside = P0[i] <= w && w < P1[i];
            lq = inside ? -__logf(width) : __int_as_float(0xff800000);
            dw0 = on ? 1.0f - raw[i] : 0.f;
            dw1 = on ? raw[i] : 0.f;
            dq0 = 1.0f / width;
            dq1 = -1.0f / width;
          }
          float lp, g;
          kl_prior(p, w, lp, g);
          lq_e += lq;
          lp_e += lp;
          if (GRADS) {
            a0[i] += dq0 - g * dw0;
            a1[i] += dq1 - g * dw1;
          }
        }
      }
      if (GRADS) {
        if (p.pshape == STAG_PARAM_EDGE_CHANNEL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int ch = chan(c, i);
            if (ch < p.K) {
              p.dp0[e * p.K + ch] = a0[i];
              p.dp1[e * p.K + ch] = a1[i];
            }
          }
        } else if (p.pshape == STAG_PARAM_CHANNEL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int ch = chan(c, i);
            if (ch < p.K) {
              my_ch[ch] += a0[i];
              my_ch[p.kpad + ch] += a1[i];
            }
          }
        } else {
          d0_e += ((a0[0] + a0[1]) + (a0[2] + a0[3])) + ((a0[4] + a0[5]) + (a0[6] + a0[7]));
          d1_e += ((a1[0] + a1[1]) + (a1[2] + a1[3])) + ((a1[4] + a1[5]) + (a1[6] + a1[7]));
        }
      }
    }
    if (GRADS && p.pshape == STAG_PARAM_EDGE) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d0_e += __shfl_xor_sync(0xffffffffu, d0_e, o);
        d1_e += __shfl_xor_sync(0xffffffffu, d1_e, o);
      }
      if (lane == 0) {
        p.dp0[e] = d0_e;
        p.dp1[e] = d1_e;
      }
    } else if (GRADS && p.pshape == STAG_PARAM_SCALAR) {
      d0_t += (double)d0_e;
      d1_t += (double)d1_e;
    }
    lq_t += (double)lq_e;
    lp_t += (double)lp_e;
  }
  // ---- CTA partials, in a fixed order ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lq_t += __shfl_xor_sync(0xffffffffu, lq_t, o);
    lp_t += __shfl_xor_sync(0xffffffffu, lp_t, o);
    d0_t += __shfl_xor_sync(0xffffffffu, d0_t, o);
    d1_t += __shfl_xor_sync(0xffffffffu, d1_t, o);
  }
  if (lane == 0) { red[warp][0] = lq_t; red[warp][1] = lp_t; red[warp][2] = d0_t; red[warp][3] = d1_t; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < KL_WARPS; ++w) t += red[w][threadIdx.x];
    p.cta_sums[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
  }
  if (ch_grads)
    for (int i = threadIdx.x; i < 2 * p.kpad; i += KL_THREADS) {
      float t = 0.f;
      for (int w = 0; w < KL_WARPS; ++w) t += kl_ch[(size_t)w * 2 * p.kpad + i];
      p.cta_ch[(size_t)blockIdx.x * 2 * p.kpad + i] = t;
    }
}

// CTA partials -> results, CTA order (one CTA)
__global__ void noise_kl_finalize(const KlParams p, int ncta, int grads, double* __restrict__ sums) {
  __shared__ double red[4];
  const int tid = threadIdx.x;
  if (tid < 4) {
    double t = 0.0;
    for (int k = 0; k < ncta; ++k) t += p.cta_sums[(size_t)k * 4 + tid];
    red[tid] = t;
    if (tid < 2) sums[tid] = t;
  }
  __syncthreads();
  if (!grads) return;
  if (p.pshape == STAG_PARAM_SCALAR) {
    if (tid == 0) { p.dp0[0] = (float)red[2]; p.dp1[0] = (float)red[3]; }
  } else if (p.pshape == STAG_PARAM_CHANNEL) {
    for (int c = tid; c < p.K; c += blockDim.x) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < ncta; ++k) {
        a += (double)p.cta_ch[(size_t)k * 2 * p.kpad + c];
        b += (double)p.cta_ch[(size_t)k * 2 * p.kpad + p.kpad + c];
      }
      p.dp0[c] = (float)a;
      p.dp1[c] = (float)b;
    }
  }
}

static int kl_grid(int64_t E) {
  const int64_t want = (E + KL_WARPS - 1) / KL_WARPS;
  const int64_t cap = (int64_t)num_sms() * 4;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace stag
