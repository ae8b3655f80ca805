// gemm_tcgen05.cu — placeholder until the tcgen05/TMA kernels land: reports
// every shape as unsupported so gts_gemm_* falls through to the SIMT kernels.
#include "common.cuh"
namespace gts {
bool gemm_nt_tcgen05_supported(const gts_gemm_nt_args*) { return false; }
bool gemm_tn_tcgen05_supported(const float*, int64_t, const float*, int64_t, int32_t, int32_t, int64_t) { return false; }
size_t gemm_tn_tcgen05_ws(int32_t, int32_t, int64_t, int32_t) { return 0; }
int gemm_nt_tcgen05(const gts_gemm_nt_args*, cudaStream_t) { set_error("tcgen05 GEMM not built"); return GTS_ERR_UNSUPPORTED; }
int gemm_tn_tcgen05(const float*, int64_t, const float*, int64_t, float*, int64_t, int32_t, int32_t, int64_t, int32_t, void*, size_t, cudaStream_t) {
  set_error("tcgen05 GEMM not built"); return GTS_ERR_UNSUPPORTED;
}
}  // namespace gts
