// Segmented softmax over the in-edges of every destination node (dgl.nn.edge_softmax as the reference's GAT uses it,
// stag/zoo/gat.py:122:  a = edge_softmax(graph, e),  e [E,H] = noise * leaky_relu(el[u] + er[v])), forward and backward.
// Replaces the scatter-amax / gather / exp / index_add / gather / divide chain (and its autograd graph) by one pass
// per direction over the CSC rows: one warp per destination, a lane per (in-edge, head) pair, per-head max / sum by
// shuffles between the lanes that hold the same head.  logits and outputs are in ORIGINAL edge order ([E,H], like every
// per-edge tensor of the reference); the row's edges are reached through the CSC edge ids.  Deterministic.
#include "common.cuh"

namespace stag {

constexpr int SM_THREADS = 256, SM_WARPS = SM_THREADS / 32;

struct SoftmaxParams {
  const int32_t* indptr;
  const int32_t* eid;
  int64_t N;
  int H;
  const float* in0;  // forward: logits;  backward: a
  const float* in1;  // backward: da
  float* out;        // forward: a;       backward: dlogits
};

// combine `v` over the lanes that hold the same head (lane % H, H a power of two <= 32)
template <bool MAX>
__device__ __forceinline__ float head_reduce(float v, int H) {
  for (int o = 16; o >= H; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = MAX ? fmaxf(v, w) : v + w;
  }
  return v;
}

// H divides 32: element t = lane + 32 k of a row's (edge, head) pairs has head lane % H for every k
template <bool BWD>
__global__ void __launch_bounds__(SM_THREADS) edge_softmax_pow2_kernel(const SoftmaxParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * SM_WARPS;
  const int H = p.H, hsh = __ffs(H) - 1, h = lane & (H - 1);
  for (int64_t v = (int64_t)blockIdx.x * SM_WARPS + (threadIdx.x >> 5); v < p.N; v += nwarps) {
    const int e0 = __ldg(p.indptr + v), e1 = __ldg(p.indptr + v + 1);
    const int total = (e1 - e0) << hsh;
    if (total == 0) continue;
    if (!BWD) {
      float m = -INFINITY;
      for (int t = lane; t < total; t += 32)
        m = fmaxf(m, __ldg(p.in0 + (int64_t)__ldg(p.eid + e0 + (t >> hsh)) * H + h));
      m = head_reduce<true>(m, H);
      float s = 0.f;
      for (int t = lane; t < total; t += 32)
        s += expf(__ldg(p.in0 + (int64_t)__ldg(p.eid + e0 + (t >> hsh)) * H + h) - m);
      s = head_reduce<false>(s, H);
      for (int t = lane; t < total; t += 32) {
        const int64_t o = (int64_t)__ldg(p.eid + e0 + (t >> hsh)) * H + h;
        p.out[o] = expf(__ldg(p.in0 + o) - m) / s;
      }
    } else {
      float d = 0.f;  // sum_e a_e da_e of this head
      for (int t = lane; t < total; t += 32) {
        const int64_t o = (int64_t)__ldg(p.eid + e0 + (t >> hsh)) * H + h;
        d = fmaf(__ldg(p.in0 + o), __ldg(p.in1 + o), d);
      }
      d = head_reduce<false>(d, H);
      for (int t = lane; t < total; t += 32) {
        const int64_t o = (int64_t)__ldg(p.eid + e0 + (t >> hsh)) * H + h;
        p.out[o] = __ldg(p.in0 + o) * (__ldg(p.in1 + o) - d);
      }
    }
  }
}

// any H: one head at a time, lanes over the edges of the row
template <bool BWD>
__global__ void __launch_bounds__(SM_THREADS) edge_softmax_any_kernel(const SoftmaxParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * SM_WARPS;
  const int H = p.H;
  for (int64_t v = (int64_t)blockIdx.x * SM_WARPS + (threadIdx.x >> 5); v < p.N; v += nwarps) {
    const int e0 = __ldg(p.indptr + v), e1 = __ldg(p.indptr + v + 1);
    for (int h = 0; h < H && e1 > e0; ++h) {
      if (!BWD) {
        float m = -INFINITY;
        for (int j = e0 + lane; j < e1; j += 32) m = fmaxf(m, __ldg(p.in0 + (int64_t)__ldg(p.eid + j) * H + h));
        m = head_reduce<true>(m, 1);
        float s = 0.f;
        for (int j = e0 + lane; j < e1; j += 32) s += expf(__ldg(p.in0 + (int64_t)__ldg(p.eid + j) * H + h) - m);
        s = head_reduce<false>(s, 1);
        for (int j = e0 + lane; j < e1; j += 32) {
          const int64_t o = (int64_t)__ldg(p.eid + j) * H + h;
          p.out[o] = expf(__ldg(p.in0 + o) - m) / s;
        }
      } else {
        float d = 0.f;
        for (int j = e0 + lane; j < e1; j += 32) {
          const int64_t o = (int64_t)__ldg(p.eid + j) * H + h;
          d = fmaf(__ldg(p.in0 + o), __ldg(p.in1 + o), d);
        }
        d = head_reduce<false>(d, 1);
        for (int j = e0 + lane; j < e1; j += 32) {
          const int64_t o = (int64_t)__ldg(p.eid + j) * H + h;
          p.out[o] = __ldg(p.in0 + o) * (__ldg(p.in1 + o) - d);
        }
      }
    }
  }
}

static int softmax_launch(const StagGraph* g, const float* in0, const float* in1, int H, float* out, bool bwd,
                          cudaStream_t stream, const char* who) {
  STAG_CHECK_ARG(g != nullptr && g->indptr != nullptr, "%s: null graph", who);
  STAG_CHECK_ARG(H > 0, "%s: H=%d must be positive", who, H);
  if (g->num_rows == 0 || g->num_edges == 0) return STAG_OK;
  STAG_CHECK_ARG(g->eid && in0 && out && (!bwd || in1), "%s: null argument", who);
  SoftmaxParams p;
  p.indptr = g->indptr; p.eid = g->eid; p.N = g->num_rows; p.H = H; p.in0 = in0; p.in1 = in1; p.out = out;
  const int64_t want = (g->num_rows + SM_WARPS - 1) / SM_WARPS;
  const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  const bool pow2 = H <= 32 && (H & (H - 1)) == 0;
  if (pow2) {
    if (bwd) edge_softmax_pow2_kernel<true><<<grid, SM_THREADS, 0, stream>>>(p);
    else edge_softmax_pow2_kernel<false><<<grid, SM_THREADS, 0, stream>>>(p);
  } else {
    if (bwd) edge_softmax_any_kernel<true><<<grid, SM_THREADS, 0, stream>>>(p);
    else edge_softmax_any_kernel<false><<<grid, SM_THREADS, 0, stream>>>(p);
  }
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}

// ---- attention logits fused in (stag/zoo/gat.py:113-122) -----------------------------------------------------------
//   logit[e,h] = w[e,h] * leaky_relu(el[u_e,h] + er[v_e,h]),   a = softmax over the in-edges of v
// The [E,H] logits (two gathers, an add, the leaky relu, the noise product: five torch launches and their autograd
// graph) never exist: the forward recomputes them per pass over the row (el rows come from L1 / L2), the backward emits
//   dw[e,h] = dlogit * leaky(pre),   t[e,h] = dlogit * w * leaky'(pre)  (= d pre, summed per SOURCE by the caller on the CSR),
//   d_er[v,h] = sum over the row of t   -- no atomics anywhere: deterministic, unlike index_add of the gather's autograd.
struct AttnParams {
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eid;
  int64_t N;
  int H;
  const float* el;   // [num_cols, H]
  const float* er;   // [N, H]
  const float* w;    // [E, H] in original edge order, or null
  float slope;
  float* a;          // forward: out;  backward: in
  const float* da;   // backward
  float* dw;         // backward, may be null
  float* t;          // backward: d(el[u] + er[v]) per edge
  float* d_er;       // backward: [N, H]
};

template <bool POW2, bool BWD>
__global__ void __launch_bounds__(SM_THREADS) attention_softmax_kernel(const AttnParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * SM_WARPS;
  const int H = p.H, hsh = POW2 ? __ffs(H) - 1 : 0;
  for (int64_t v = (int64_t)blockIdx.x * SM_WARPS + (threadIdx.x >> 5); v < p.N; v += nwarps) {
    const int e0 = __ldg(p.indptr + v), e1 = __ldg(p.indptr + v + 1);
    if (e1 == e0) {
      if (BWD && lane < H) p.d_er[v * H + lane] = 0.f;
      if (BWD) for (int h = 32 + lane; h < H; h += 32) p.d_er[v * H + h] = 0.f;
      continue;
    }
    // POW2: lanes over the (edge, head) pairs of the row, head = lane % H, one pass;  else: head by head, lanes over the edges
    const int nh = POW2 ? 1 : H;
    for (int hh = 0; hh < nh; ++hh) {
      const int h = POW2 ? (lane & (H - 1)) : hh;
      const int total = POW2 ? (e1 - e0) << hsh : (e1 - e0);
      const int red = POW2 ? H : 1;
      const float erv = __ldg(p.er + v * H + h);
      auto edge_of = [&](int t) { return e0 + (POW2 ? (t >> hsh) : t); };
      auto pre_of = [&](int j) { return __ldg(p.el + (int64_t)__ldg(p.indices + j) * H + h) + erv; };
      auto wt_of = [&](int64_t o) { return p.w ? __ldg(p.w + o) : 1.0f; };
      if (!BWD) {
        float m = -INFINITY;
        for (int t = lane; t < total; t += 32) {
          const int j = edge_of(t);
          const float pre = pre_of(j);
          m = fmaxf(m, (pre > 0.f ? pre : p.slope * pre) * wt_of((int64_t)__ldg(p.eid + j) * H + h));
        }
        m = head_reduce<true>(m, red);
        float sum = 0.f;
        for (int t = lane; t < total; t += 32) {
          const int j = edge_of(t);
          const float pre = pre_of(j);
          sum += expf((pre > 0.f ? pre : p.slope * pre) * wt_of((int64_t)__ldg(p.eid + j) * H + h) - m);
        }
        sum = head_reduce<false>(sum, red);
        for (int t = lane; t < total; t += 32) {
          const int j = edge_of(t);
          const int64_t o = (int64_t)__ldg(p.eid + j) * H + h;
          const float pre = pre_of(j);
          p.a[o] = expf((pre > 0.f ? pre : p.slope * pre) * wt_of(o) - m) / sum;
        }
      } else {
        float d = 0.f;  // sum_e a_e da_e of this head
        for (int t = lane; t < total; t += 32) {
          const int64_t o = (int64_t)__ldg(p.eid + edge_of(t)) * H + h;
          d = fmaf(__ldg(p.a + o), __ldg(p.da + o), d);
        }
        d = head_reduce<false>(d, red);
        float acc = 0.f;
        for (int t = lane; t < total; t += 32) {
          const int j = edge_of(t);
          const int64_t o = (int64_t)__ldg(p.eid + j) * H + h;
          const float dl = __ldg(p.a + o) * (__ldg(p.da + o) - d);
          const float pre = pre_of(j);
          if (p.dw) p.dw[o] = dl * (pre > 0.f ? pre : p.slope * pre);
          const float dp = dl * wt_of(o) * (pre > 0.f ? 1.0f : p.slope);
          p.t[o] = dp;
          acc += dp;
        }
        acc = head_reduce<false>(acc, red);
        if (POW2 ? lane < H : lane == 0) p.d_er[v * H + h] = acc;
      }
    }
  }
}

static int attention_launch(const StagGraph* g, AttnParams p, bool bwd, cudaStream_t stream, const char* who) {
  STAG_CHECK_ARG(g != nullptr && g->indptr != nullptr, "%s: null graph", who);
  STAG_CHECK_ARG(p.H > 0, "%s: H=%d must be positive", who, p.H);
  if (g->num_rows == 0) return STAG_OK;
  STAG_CHECK_ARG(p.er && p.a && (g->num_edges == 0 || (g->eid && g->indices && p.el)) && (!bwd || (p.da && p.t && p.d_er)),
                 "%s: null argument", who);
  p.indptr = g->indptr; p.indices = g->indices; p.eid = g->eid; p.N = g->num_rows;
  const int64_t want = (g->num_rows + SM_WARPS - 1) / SM_WARPS;
  const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  const bool pow2 = p.H <= 32 && (p.H & (p.H - 1)) == 0;
  if (pow2) {
    if (bwd) attention_softmax_kernel<true, true><<<grid, SM_THREADS, 0, stream>>>(p);
    else attention_softmax_kernel<true, false><<<grid, SM_THREADS, 0, stream>>>(p);
  } else {
    if (bwd) attention_softmax_kernel<false, true><<<grid, SM_THREADS, 0, stream>>>(p);
    else attention_softmax_kernel<false, false><<<grid, SM_THREADS, 0, stream>>>(p);
  }
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}

}  // namespace stag

using namespace stag;

extern "C" int stag_edge_softmax(const StagGraph* csc, const float* logits, int32_t H, float* out, void* stream) {
  return softmax_launch(csc, logits, nullptr, H, out, false, (cudaStream_t)stream, "stag_edge_softmax");
}

extern "C" int stag_edge_softmax_bwd(const StagGraph* csc, const float* a, const float* da, int32_t H, float* dlogits,
                                     void* stream) {
  return softmax_launch(csc, a, da, H, dlogits, true, (cudaStream_t)stream, "stag_edge_softmax_bwd");
}

extern "C" int stag_attention_softmax(const StagGraph* csc, const float* el, const float* er, const float* w, float slope,
                                      int32_t H, float* a, void* stream) {
  AttnParams p = {};
  p.H = H; p.el = el; p.er = er; p.w = w; p.slope = slope; p.a = a;
  if (csc && csc->num_edges == 0) return STAG_OK;
  return attention_launch(csc, p, false, (cudaStream_t)stream, "stag_attention_softmax");
}

extern "C" int stag_attention_softmax_bwd(const StagGraph* csc, const float* el, const float* er, const float* w, float slope,
                                          int32_t H, const float* a, const float* da, float* dw, float* dpre, float* d_er,
                                          void* stream) {
  AttnParams p = {};
  p.H = H; p.el = el; p.er = er; p.w = w; p.slope = slope; p.a = const_cast<float*>(a); p.da = da; p.dw = dw; p.t = dpre;
  p.d_er = d_er;
  return attention_launch(csc, p, true, (cudaStream_t)stream, "stag_attention_softmax_bwd");
}
