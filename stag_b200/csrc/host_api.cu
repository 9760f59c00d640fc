// Host-buffer convenience entry point: the whole path (CSC/CSR build, fused forward,
// fused backward) for a caller that holds everything in host memory and has no torch.
// Temporary device memory is allocated and released inside the call.
#include "common.cuh"

namespace stag {

__global__ void k_gcn_scales(const int32_t* __restrict__ indptr, int64_t N, float* __restrict__ scale) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) {
    const float d = fmaxf((float)(indptr[v + 1] - indptr[v]), 1.0f);
    scale[v] = rsqrtf(d);
  }
}

__global__ void k_sum_samples(const float* __restrict__ in, int S, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += in[(int64_t)s * n + i];
    out[i] = a;
  }
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <class T> T* as() { return (T*)p; }
};

}  // namespace stag

using namespace stag;

extern "C" int stag_aggregate_host(int device, const int64_t* src, const int64_t* dst, int64_t E, int64_t N,
                                   const float* x, const float* dout, int32_t D, int32_t S,
                                   const StagNoise* noise, int gcn_norm_both, float* out, float* dx) {
  STAG_CHECK_ARG(src && dst && x && noise && out, "stag_aggregate_host: null argument");
  STAG_CHECK_ARG(E >= 0 && N > 0 && D > 0 && S > 0, "stag_aggregate_host: bad sizes");
  STAG_CHECK_ARG(noise->kind == STAG_NOISE_NONE || (noise->kind >= STAG_NOISE_NORMAL &&
                 noise->param_shape <= STAG_PARAM_CHANNEL),
                 "stag_aggregate_host: only generated noise with scalar / per-channel parameters");
  STAG_CHECK_ARG((dout == nullptr) == (dx == nullptr), "stag_aggregate_host: dout and dx go together");
  STAG_CHECK_ARG(!(dout && noise->in_norm), "stag_aggregate_host: backward with in_norm is not offered here");
  STAG_CUDA(cudaSetDevice(device));
  cudaStream_t stream;
  STAG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  int rc = STAG_OK;
  {
    DevBuf d_src, d_dst, d_x, d_out, d_dout, d_dxs, d_dx, d_p0, d_p1, d_ws, d_sw;
    DevBuf csc[9], csr[9], d_ss, d_ds;
    const size_t nd = (size_t)N * D * 4;
    const size_t hubcap = (size_t)(E / kHubThreshold + 2) * 4;
#define TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("stag_aggregate_host: %s -> %s", #call, cudaGetErrorString(e__)); rc = STAG_ECUDA; goto done; } } while (0)
#define TRYRC(call) do { rc = (call); if (rc) goto done; } while (0)
    {
      TRY(d_src.alloc(E * 8)); TRY(d_dst.alloc(E * 8)); TRY(d_x.alloc(nd)); TRY(d_out.alloc(nd * S));
      TRY(cudaMemcpyAsync(d_src.p, src, E * 8, cudaMemcpyHostToDevice, stream));
      TRY(cudaMemcpyAsync(d_dst.p, dst, E * 8, cudaMemcpyHostToDevice, stream));
      TRY(cudaMemcpyAsync(d_x.p, x, nd, cudaMemcpyHostToDevice, stream));
      const size_t cws = stag_csx_workspace_bytes(E, N);
      TRY(d_ws.alloc(cws));
      StagGraph G[2];
      for (int k = 0; k < 2; ++k) {
        DevBuf* b = k == 0 ? csc : csr;
        TRY(b[0].alloc((N + 1) * 4)); TRY(b[1].alloc(E * 4)); TRY(b[2].alloc(E * 4));
        TRY(b[3].alloc(hubcap)); TRY(b[4].alloc(hubcap)); TRY(b[5].alloc(N * 4));
        TRY(b[6].alloc((size_t)stag_csx_items_capacity(E, N) * 16)); TRY(b[7].alloc(E * 4)); TRY(b[8].alloc(E * 4));
        int32_t counts[3];
        TRYRC(stag_csx_build((const int64_t*)d_src.p, (const int64_t*)d_dst.p, E, N, k == 0, b[0].as<int32_t>(),
                             b[1].as<int32_t>(), b[2].as<int32_t>(), b[3].as<int32_t>(), b[4].as<int32_t>(),
                             b[5].as<int32_t>(), b[6].as<int32_t>(), b[7].as<int32_t>(), b[8].as<int32_t>(), counts,
                             d_ws.p, cws, stream));
        G[k].num_rows = N; G[k].num_cols = N; G[k].num_edges = E;
        G[k].indptr = b[0].as<int32_t>(); G[k].indices = b[1].as<int32_t>(); G[k].eid = b[2].as<int32_t>();
        G[k].num_hubs = counts[0]; G[k].num_hub_segs = counts[1];
        G[k].hub_rows = b[3].as<int32_t>(); G[k].hub_seg_ptr = b[4].as<int32_t>();
        G[k].row_order = b[5].as<int32_t>();
        G[k].items = b[6].as<int32_t>(); G[k].num_items = counts[2]; G[k].erow = b[7].as<int32_t>(); G[k].eidf = b[8].as<int32_t>();
      }
      const float *ss = nullptr, *ds = nullptr;
      if (gcn_norm_both) {
        TRY(d_ss.alloc(N * 4)); TRY(d_ds.alloc(N * 4));
        const unsigned gb = (unsigned)((N + 255) / 256);
        k_gcn_scales<<<gb, 256, 0, stream>>>(G[1].indptr, N, d_ss.as<float>());  // out-degree
        k_gcn_scales<<<gb, 256, 0, stream>>>(G[0].indptr, N, d_ds.as<float>());  // in-degree
        TRY(cudaGetLastError());
        ss = d_ss.as<float>(); ds = d_ds.as<float>();
      }
      StagNoise nz = *noise;
      if (noise->kind >= STAG_NOISE_NORMAL) {
        const size_t pn = (noise->param_shape == STAG_PARAM_SCALAR || noise->K == 1) ? 1 : (size_t)noise->K;
        TRY(d_p0.alloc(pn * 4));
        TRY(cudaMemcpyAsync(d_p0.p, noise->p0, pn * 4, cudaMemcpyHostToDevice, stream));
        nz.p0 = d_p0.as<float>();
        if (noise->p1) {
          TRY(d_p1.alloc(pn * 4));
          TRY(cudaMemcpyAsync(d_p1.p, noise->p1, pn * 4, cudaMemcpyHostToDevice, stream));
          nz.p1 = d_p1.as<float>();
        }
      }
      size_t sws = stag_spmm_workspace_bytes(&G[0], D, S);
      const size_t sws1 = stag_spmm_workspace_bytes(&G[1], D, S);
      if (sws1 > sws) sws = sws1;
      TRY(d_sw.alloc(sws));
      TRYRC(stag_spmm_fwd(&G[0], d_x.as<float>(), D, 0, D, S, &nz, ss, ds, d_out.as<float>(), D, (int64_t)N * D,
                          nullptr, d_sw.p, sws, stream));
      TRY(cudaMemcpyAsync(out, d_out.p, nd * S, cudaMemcpyDeviceToHost, stream));
      if (dout) {
        TRY(d_dout.alloc(nd * S)); TRY(d_dxs.alloc(nd * S)); TRY(d_dx.alloc(nd));
        TRY(cudaMemcpyAsync(d_dout.p, dout, nd * S, cudaMemcpyHostToDevice, stream));
        StagNoise nb = nz;
        nb.in_norm = 0;
        // dX only: the transposed aggregation is the forward kernel on the CSR graph
        TRYRC(stag_spmm_fwd(&G[1], d_dout.as<float>(), D, (int64_t)N * D, D, S, &nb, ds, ss, d_dxs.as<float>(), D,
                            (int64_t)N * D, nullptr, d_sw.p, sws, stream));
        const int64_t n = (int64_t)N * D;
        k_sum_samples<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_dxs.as<float>(), S, n, d_dx.as<float>());
        TRY(cudaGetLastError());
        TRY(cudaMemcpyAsync(dx, d_dx.p, nd, cudaMemcpyDeviceToHost, stream));
      }
      TRY(cudaStreamSynchronize(stream));
    }
  done:;
#undef TRY
#undef TRYRC
  }
  cudaStreamDestroy(stream);
  return rc;
}
