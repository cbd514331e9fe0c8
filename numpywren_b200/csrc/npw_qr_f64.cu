// npw_qr_f64.cu — compact-WY Householder QR of a tall tile (kernels.qr_factor → fast_qr,
// kernels.py:86-105,127-130).  Implemented in a later milestone; until then the entry
// point reports NPW_ERR_UNSUPPORTED so callers fail loudly instead of falling back.
#include "npw_common.cuh"

extern "C" {

size_t npw_geqrt_work_bytes(int64_t m, int64_t n) {
  (void)m; (void)n;
  return 0;
}

int npw_geqrt_f64(double* V, int64_t ldv, double* T, int64_t ldt, double* R, int64_t ldr, const double* A, int64_t lda,
                  int64_t m, int64_t n, void* work, npw_stream_t stream) {
  (void)V; (void)ldv; (void)T; (void)ldt; (void)R; (void)ldr; (void)A; (void)lda; (void)m; (void)n; (void)work; (void)stream;
  npw::set_error("npw_geqrt_f64: not implemented yet");
  return NPW_ERR_UNSUPPORTED;
}

}  // extern "C"
