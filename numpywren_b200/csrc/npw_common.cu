// npw_common.cu — error plumbing, launch counter, TMA tensor-map construction.
#include "npw_common.cuh"

#include <stdarg.h>
#include <string.h>

namespace npw {

std::atomic<uint64_t> g_launches{0};

static thread_local char t_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // Resolved through the runtime so the library has no link-time dependency on libcuda.
  static EncodeTiledFn fn = nullptr;
  static std::atomic<int> state{0};
  if (state.load(std::memory_order_acquire) == 2) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  state.store(2, std::memory_order_release);
  return fn;
}

int make_tmap_f64(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * sizeof(double)};
  cuuint32_t box[2] = {16u, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (base=%p rows=%lld cols=%lld ld=%lld)", (int)r, (const void*)base,
              (long long)rows, (long long)cols, (long long)ld);
    return -1;
  }
  return 0;
}

// 3-D int8 tensor map over `planes` digit planes of a row-major rows x k byte matrix (k contiguous):
// box = 128 bytes of k x box_rows rows x 1 plane, 128-byte swizzle (the K-major SWIZZLE_128B operand layout of tcgen05.mma).
int make_tmap_i8_3d(CUtensorMap* map, const int8_t* base, int64_t rows, int64_t k, int64_t planes, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(k) * static_cast<cuuint64_t>(rows)};
  cuuint32_t box[3] = {128u, box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (int8 digits) failed: CUresult %d (base=%p rows=%lld k=%lld planes=%lld)", (int)r,
              (const void*)base, (long long)rows, (long long)k, (long long)planes);
    return -1;
  }
  return 0;
}

}  // namespace npw

extern "C" {

int npw_version(void) { return 100; }  // 0.1.0
const char* npw_last_error(void) { return npw::t_err; }
const char* npw_build_arch(void) { return "sm_100a"; }
uint64_t npw_launch_count(void) { return npw::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
