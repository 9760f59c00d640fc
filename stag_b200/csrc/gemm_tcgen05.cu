// Dense feature transform agg @ W on tcgen05 tensor cores (placeholder until the
// TMEM/TMA kernel lands; the Python host side falls back to torch.matmul/cuBLAS when
// this entry point reports STAG_EUNSUPPORTED).
#include "common.cuh"

extern "C" size_t stag_gemm_workspace_bytes(int64_t M, int32_t Nout, int32_t K) {
  (void)M; (void)Nout; (void)K;
  return 0;
}

extern "C" int stag_gemm_tcgen05(const float* a, int64_t lda, const float* wt, int64_t ldw, int64_t M, int32_t Nout,
                                 int32_t K, const float* row_scale, const float* bias, int act, float* out,
                                 int64_t ldo, void* ws, size_t ws_bytes, void* stream) {
  (void)a; (void)lda; (void)wt; (void)ldw; (void)M; (void)Nout; (void)K; (void)row_scale; (void)bias; (void)act;
  (void)out; (void)ldo; (void)ws; (void)ws_bytes; (void)stream;
  stag::set_error("stag_gemm_tcgen05: not built in this version");
  return STAG_EUNSUPPORTED;
}
