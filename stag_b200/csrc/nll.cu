// Fused likelihood epilogue: masked mean negative log-likelihood of S Monte-Carlo outputs in one pass,
// together with its gradient with respect to the network output.
//
// Replaces, per MC sample, the chain the reference runs on the last layer's output (stag/models.py:69-72,
// stag/likelihoods.py:13-38 -> torch.distributions):
//     nll = -likelihood.log_prob(feat, y)[mask].mean()
//   Categorical(probs=feat): probs / probs.sum(-1), logits = log(clamp(., eps, 1-eps)), gather at y
//                            (torch/distributions/categorical.py:66-67,128-133; utils.py probs_to_logits)
//   Bernoulli(probs=feat):   -binary_cross_entropy_with_logits(log(pc) - log1p(-pc), y), pc = clamp(feat, eps, 1-eps)
//                            (torch/distributions/bernoulli.py:109-112)
// i.e. about eight elementwise / reduction launches and three [N,C] temporaries per sample.  Here one warp
// owns one (sample, node) row: row sum, the label's term, the per-element gradient; per-CTA partial sums are
// combined in a fixed order by a finalize kernel (deterministic).
#include "common.cuh"

namespace stag {

constexpr int NLL_THREADS = 256, NLL_WARPS = NLL_THREADS / 32;
constexpr float kProbEps = 1.1920928955078125e-07f;  // torch.finfo(float32).eps

struct NllParams {
  const float* probs;
  int64_t ld, ss, N;
  int C, S, kind;
  const int64_t* y_idx;  // categorical: [N]
  const float* y_val;    // bernoulli: [N, C], row stride ldy
  int64_t ldy;
  const uint8_t* mask;   // [N] or null
  float* dprobs;         // [S,N,C] (ld, ss as probs) or null: d(sum of -log_prob over the masked rows of sample s) / d probs
  float* partial;        // [S + 1][grid]: sums of every sample, then the number of masked rows
  float* nll;            // [S]
  float* count;          // [1]: number of terms each mean runs over
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NLL_THREADS) nll_kernel(const NllParams p) {
  __shared__ float cta_sum[NLL_WARPS], cta_cnt[NLL_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * NLL_WARPS;
  for (int s = 0; s < p.S; ++s) {
    float acc = 0.f, rows = 0.f;  // this warp's sum for sample s and its masked rows (lane 0 holds them)
    for (int64_t v = (int64_t)blockIdx.x * NLL_WARPS + warp; v < p.N; v += nwarps) {
      const bool on = p.mask == nullptr || p.mask[v] != 0;
      const float* row = p.probs + (int64_t)s * p.ss + v * p.ld;
      float* drow = p.dprobs ? p.dprobs + (int64_t)s * p.ss + v * p.ld : nullptr;
      if (!on) {
        if (drow)
          for (int c = lane; c < p.C; c += 32) drow[c] = 0.f;
        continue;
      }
      rows += 1.0f;
      if (p.kind == 0) {
        float sum = 0.f;
        for (int c = lane; c < p.C; c += 32) sum += row[c];
        sum = warp_sum(sum);
        const int64_t yv = p.y_idx[v];
        // a label outside [0, C) (an ignore index on an unmasked row) is never dereferenced: the row's loss is NaN and
        // its gradient zero, so the mistake is visible instead of an out-of-bounds read
        const bool bad = yv < 0 || yv >= p.C;
        const float py = bad ? __int_as_float(0x7fc00000) : row[yv];
        const float q = py / sum;
        // torch: Categorical(probs=...) renormalises and clamps; a row that sums to zero (or is NaN) gives NaN there,
        // and fmaxf / fminf would silently turn it into a finite value here
        const float qc = (q != q || !(sum > 0.f)) ? __int_as_float(0x7fc00000) : fminf(fmaxf(q, kProbEps), 1.0f - kProbEps);
        if (lane == 0) acc -= logf(qc);
        if (drow) {
          // -log(clamp(p_y / sum)): the clamp passes the gradient only inside (eps, 1 - eps)
          const bool inside = !bad && q >= kProbEps && q <= 1.0f - kProbEps;
          const float inv_sum = 1.0f / sum;
          for (int c = lane; c < p.C; c += 32) {
            float g = 0.f;
            if (inside) g = inv_sum - (c == yv ? 1.0f / py : 0.f);
            drow[c] = g;
          }
        }
      } else {
        const float* yrow = p.y_val + v * p.ldy;
        float t = 0.f;
        for (int c = lane; c < p.C; c += 32) {
          const float pr = row[c], yy = yrow[c];
          const float pc = fminf(fmaxf(pr, kProbEps), 1.0f - kProbEps);
          // -[y log pc + (1 - y) log(1 - pc)]
          t -= yy * logf(pc) + (1.0f - yy) * log1pf(-pc);
          if (drow) {
            const bool inside = pr >= kProbEps && pr <= 1.0f - kProbEps;
            drow[c] = inside ? (1.0f - yy) / (1.0f - pc) - yy / pc : 0.f;
          }
        }
        t = warp_sum(t);
        if (lane == 0) acc += t;
      }
    }
    if (lane == 0) { cta_sum[warp] = acc; cta_cnt[warp] = rows; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f, n = 0.f;
      for (int w = 0; w < NLL_WARPS; ++w) { t += cta_sum[w]; n += cta_cnt[w]; }
      p.partial[(int64_t)s * gridDim.x + blockIdx.x] = t;
      if (s == 0) p.partial[(int64_t)p.S * gridDim.x + blockIdx.x] = n;  // exact: integers below 2^24 per CTA
    }
    __syncthreads();
  }
}

// nll[s] = (sum of the CTA partials in CTA order) / count;  count = masked rows (x C for Bernoulli)
__global__ void nll_finalize_kernel(const NllParams p, int grid) {
  __shared__ float cnt_s;
  if (threadIdx.x == 0) {
    double cnt = 0.0;
    for (int b = 0; b < grid; ++b) cnt += (double)p.partial[(int64_t)p.S * grid + b];
    if (p.kind == 1) cnt *= p.C;
    cnt_s = (float)cnt;
    p.count[0] = cnt_s;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < p.S; s += blockDim.x) {
    float t = 0.f;
    for (int b = 0; b < grid; ++b) t += p.partial[(int64_t)s * grid + b];
    p.nll[s] = t / cnt_s;  // 0 / 0 = nan for an empty mask, like torch's mean of an empty tensor
  }
}

static int nll_grid(int64_t N) {
  const int64_t want = (N + NLL_WARPS - 1) / NLL_WARPS;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace stag

using namespace stag;

extern "C" size_t stag_nll_workspace_bytes(int64_t N, int32_t S) {
  if (N < 0 || S <= 0) return 0;
  return align_up((size_t)(S + 1) * nll_grid(N) * sizeof(float) + 16, 256);
}

extern "C" int stag_nll(const float* probs, int64_t ld, int64_t sample_stride, int64_t N, int32_t C, int32_t S,
                        int kind, const void* y, int64_t ldy, const uint8_t* mask, float* nll_out, float* count_out,
                        float* dprobs, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  STAG_CHECK_ARG(N >= 0 && C > 0 && S > 0, "stag_nll: bad sizes N=%lld C=%d S=%d", (long long)N, C, S);
  STAG_CHECK_ARG(kind == 0 || kind == 1, "stag_nll: kind must be 0 (categorical) or 1 (bernoulli)");
  STAG_CHECK_ARG(nll_out && count_out, "stag_nll: null output");
  STAG_CHECK_ARG(N == 0 || (probs && y), "stag_nll: null input");
  STAG_CHECK_ARG(ld >= C && (kind == 0 || ldy >= C), "stag_nll: row strides smaller than C");
  const size_t need = stag_nll_workspace_bytes(N, S);
  if (!ws || ws_bytes < need) {
    set_error("stag_nll: workspace %zu < required %zu", ws_bytes, need);
    return STAG_EWORKSPACE;
  }
  NllParams p;
  p.probs = probs; p.ld = ld; p.ss = sample_stride; p.N = N; p.C = C; p.S = S; p.kind = kind;
  p.y_idx = kind == 0 ? (const int64_t*)y : nullptr;
  p.y_val = kind == 1 ? (const float*)y : nullptr;
  p.ldy = ldy; p.mask = mask; p.dprobs = dprobs;
  p.partial = (float*)ws; p.nll = nll_out; p.count = count_out;
  const int grid = nll_grid(N);
  nll_kernel<<<grid, NLL_THREADS, 0, stream>>>(p);
  STAG_LAUNCH_CHECK();
  nll_finalize_kernel<<<1, 128, 0, stream>>>(p, grid);
  STAG_LAUNCH_CHECK();
  return STAG_OK;
}
