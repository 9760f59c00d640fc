/* TEST INFRASTRUCTURE ONLY -- plain-C (OpenMP) restatement of the reference's stochastic
 * neighbour aggregation, used (a) as the checker at sizes where the torch/numpy oracle is
 * slow and (b) as the timed CPU baseline of bench.py (cpu_baseline.kind = "port").  The
 * product (stag_b200/) never links or calls this.
 *
 * Algorithm restated (reference file:line, /root/reference):
 *   noise [E,K] materialised, w = loc + eps*scale | low + u*(high-low) | (u < p)
 *        stag/layers.py:115-129; torch/distributions/normal.py:82-85, uniform.py:85-88,
 *        bernoulli.py:116-119;  relu stag/layers.py:98-99
 *   in-norm   stag/layers.py:8-36
 *   forward   out[v,c] = ds[v] * sum_{e:(u->v)} w[e,c] * ss[u] * x[u,c], in-edges visited in
 *        edge-id order   stag/zoo/gcn.py:63,67-75,94-96,100-108 (update_all(u_mul_e, sum))
 *   backward  dx[u,c] = ss[u] * sum_{e:(u->v)} w[e,c] * ds[v] * dout[v,c]
 *             dw[e,c] = ss[u]*x[u,c] * ds[v]*dout[v,c]          (DGL GSpMM.backward: gspmm on
 *        the reverse graph + gsddmm)
 * The variates come from the same counter-based generator as the CUDA library
 * (Philox4x32-10, Random123 constants; layout in oracle/ref_philox.py), so the fused GPU
 * path can be checked end to end, not only under external noise.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

enum { K_NONE = 0, K_EXTERNAL = 1, K_NORMAL = 2, K_UNIFORM = 3, K_BERNOULLI = 4 };
enum { P_SCALAR = 0, P_CHANNEL = 1, P_EDGE = 2, P_EDGE_CHANNEL = 3 };

API int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

API void ref_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* stable counting sort of the COO list by destination (CSC) or source (CSR) */
API void ref_csx_build(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, int by_dst,
                       int32_t* indptr, int32_t* indices, int32_t* eid) {
  const int64_t* key = by_dst ? dst : src;
  const int64_t* oth = by_dst ? src : dst;
  memset(indptr, 0, (size_t)(N + 1) * sizeof(int32_t));
  for (int64_t e = 0; e < E; ++e) indptr[key[e] + 1]++;
  for (int64_t v = 0; v < N; ++v) indptr[v + 1] += indptr[v];
  int32_t* cur = (int32_t*)malloc((size_t)(N > 0 ? N : 1) * sizeof(int32_t));
  memcpy(cur, indptr, (size_t)N * sizeof(int32_t));
  for (int64_t e = 0; e < E; ++e) {
    const int32_t p = cur[key[e]]++;
    indices[p] = (int32_t)oth[e];
    eid[p] = (int32_t)e;
  }
  free(cur);
}

#define PHILOX_ROUNDS 7
static inline void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1;
    c[3] = (uint32_t)p0;
    c[0] = n0;
    c[2] = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

static inline float half_f(uint32_t h) { return (float)h; }

static inline void box_muller(uint32_t r, float* z0, float* z1) {
  const float u1 = ((float)(r & 0xffffu) + 0.5f) * 1.52587890625e-05f;
  const float rad = sqrtf(-1.3862943611198906f * log2f(u1));
  const float ang = fmaf(8388608.0f + (float)(r >> 16), 9.58738019107841e-05f, -804.2476806640625f);
  *z0 = rad * cosf(ang);
  *z1 = rad * sinf(ang);
}

/* raw variates of (edge, oct, sample): 8 standard normals or 8 U[0,1) */
static inline void raw8(int kind, uint32_t e, uint32_t q, uint32_t s, uint64_t seed, uint64_t offset, float v[8]) {
  uint32_t c[4] = {q, e, s, (uint32_t)(offset & 0xffffffffu)};
  philox4x32(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32));
  for (int i = 0; i < 4; ++i) {
    if (kind == K_NORMAL) {
      box_muller(c[i], &v[2 * i], &v[2 * i + 1]);
    } else {
      v[2 * i] = half_f(c[i] & 0xffffu) * 1.52587890625e-05f;
      v[2 * i + 1] = half_f(c[i] >> 16) * 1.52587890625e-05f;
    }
  }
}

static inline float param_at(const float* p, int pshape, int64_t e, int c, int K) {
  switch (pshape) {
    case P_SCALAR: return p[0];
    case P_CHANNEL: return p[c];
    case P_EDGE: return p[e];
    default: return p[e * (int64_t)K + c];
  }
}

/* materialise w [E,K] (and optionally the raw variates) for one MC sample */
API void ref_noise(int kind, int64_t E, int K, int sample, uint64_t seed, uint64_t offset, const float* p0,
                   const float* p1, int pshape, int relu, float* w, float* raw) {
  if (K == 1 && pshape == P_CHANNEL) pshape = P_SCALAR;
  if (K == 1 && pshape == P_EDGE_CHANNEL) pshape = P_EDGE;
  /* block q serves channels c0+{0..3} and c0+32+{0..3}, c0 = 64*(q/8) + 4*(q%8) */
  const int rem = K % 64;
  const int nq = 8 * (K / 64) + ((rem + 3) / 4 < 8 ? (rem + 3) / 4 : 8);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e) {
    for (int q = 0; q < nq; ++q) {
      float v[8];
      raw8(kind, (uint32_t)e, (uint32_t)q, (uint32_t)sample, seed, offset, v);
      const int c0 = 64 * (q / 8) + 4 * (q % 8);
      for (int i = 0; i < 8; ++i) {
        const int c = i < 4 ? c0 + i : c0 + 28 + i;
        if (c >= K) continue;
        const float a = param_at(p0, pshape, e, c, K);
        const float b = p1 ? param_at(p1, pshape, e, c, K) : 0.0f;
        float x;
        if (kind == K_NORMAL) x = fmaf(v[i], b, a);
        else if (kind == K_UNIFORM) x = fmaf(v[i], b - a, a);
        else x = v[i] < a ? 1.0f : 0.0f;
        if (relu && x < 0.0f) x = 0.0f;
        w[e * (int64_t)K + c] = x;
        if (raw) raw[e * (int64_t)K + c] = v[i];
      }
    }
  }
}

/* in-norm (stag/layers.py:8-36): w'[e,c] = w[e,c] * (indeg(v)/sum_in w[.,c], or 1 if the sum is 0) */
API void ref_in_norm(const int32_t* indptr, const int32_t* eid, int64_t N, int K, float* w) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t v = 0; v < N; ++v) {
    const int32_t b = indptr[v], e = indptr[v + 1];
    for (int c = 0; c < K; ++c) {
      float s = 0.0f;
      for (int32_t j = b; j < e; ++j) s += w[(int64_t)eid[j] * K + c];
      const float sc = s != 0.0f ? (float)(e - b) / s : 1.0f;
      for (int32_t j = b; j < e; ++j) w[(int64_t)eid[j] * K + c] *= sc;
    }
  }
}

/* rows = destinations (CSC): forward.  With the CSR structure, x := dout and the scales swapped
 * it is the transposed aggregation dX.  w may be NULL (copy_u). */
API void ref_aggregate(const int32_t* indptr, const int32_t* indices, const int32_t* eid, int64_t N, const float* x,
                       int D, const float* w, int K, const float* gather_scale, const float* row_scale,
                       float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t v = 0; v < N; ++v) {
    float* o = out + v * (int64_t)D;
    for (int c = 0; c < D; ++c) o[c] = 0.0f;
    for (int32_t j = indptr[v]; j < indptr[v + 1]; ++j) {
      const int64_t u = indices[j], e = eid[j];
      const float gs = gather_scale ? gather_scale[u] : 1.0f;
      const float* xr = x + u * (int64_t)D;
      if (!w) {
        for (int c = 0; c < D; ++c) o[c] += xr[c] * gs;
      } else if (K == 1) {
        const float ww = w[e];
        for (int c = 0; c < D; ++c) o[c] = fmaf(ww, xr[c] * gs, o[c]);
      } else {
        const float* wr = w + e * (int64_t)K;
        for (int c = 0; c < D; ++c) o[c] = fmaf(wr[c], xr[c] * gs, o[c]);
      }
    }
    if (row_scale) {
      const float rs = row_scale[v];
      for (int c = 0; c < D; ++c) o[c] *= rs;
    }
  }
}

/* SDDMM: dw[e,c] = ss[u] x[u,c] * ds[v] dout[v,c]   (K == D) or summed over c (K == 1) */
API void ref_sddmm(const int64_t* src, const int64_t* dst, int64_t E, const float* x, const float* dout, int D,
                   int K, const float* src_scale, const float* dst_scale, float* dw) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e) {
    const int64_t u = src[e], v = dst[e];
    const float a = src_scale ? src_scale[u] : 1.0f, b = dst_scale ? dst_scale[v] : 1.0f;
    const float* xr = x + u * (int64_t)D;
    const float* gr = dout + v * (int64_t)D;
    if (K == 1) {
      float s = 0.0f;
      for (int c = 0; c < D; ++c) s += (xr[c] * a) * (gr[c] * b);
      dw[e] = s;
    } else {
      for (int c = 0; c < D; ++c) dw[e * (int64_t)K + c] = (xr[c] * a) * (gr[c] * b);
    }
  }
}

/* One un-fused layer pass the way the reference executes it (materialised noise tensor):
 * noise -> forward -> backward dX (+ dW and its reduction to d loc / d scale when vi).
 * Scratch w/dw [E,K] are caller-allocated.  Returns nothing; used for timing and checking. */
API void ref_layer_fwd_bwd(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, const int32_t* csc_indptr,
                           const int32_t* csc_indices, const int32_t* csc_eid, const int32_t* csr_indptr,
                           const int32_t* csr_indices, const int32_t* csr_eid, const float* x, const float* dout,
                           int D, int kind, int sample, uint64_t seed, uint64_t offset, const float* p0,
                           const float* p1, int pshape, const float* src_scale, const float* dst_scale, int vi,
                           float* w, float* raw, float* dw, float* out, float* dx, double* dp0, double* dp1) {
  ref_noise(kind, E, D, sample, seed, offset, p0, p1, pshape, 0, w, vi ? raw : NULL);
  ref_aggregate(csc_indptr, csc_indices, csc_eid, N, x, D, w, D, src_scale, dst_scale, out);
  ref_aggregate(csr_indptr, csr_indices, csr_eid, N, dout, D, w, D, dst_scale, src_scale, dx);
  if (vi) {
    ref_sddmm(src, dst, E, x, dout, D, D, src_scale, dst_scale, dw);
    double a = 0.0, b = 0.0;
#pragma omp parallel for reduction(+ : a, b) schedule(static)
    for (int64_t i = 0; i < E * (int64_t)D; ++i) {
      a += dw[i];
      b += (double)dw[i] * raw[i];
    }
    *dp0 = a;
    *dp1 = b;
  }
}
