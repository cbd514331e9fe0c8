// npw_elementwise.cu — HBM-bound tile ops: add_matrices / mul / identity-copy / transpose /
// diagonal shift / triangular masks / synthetic fill.  (kernels.py:16-20, 233-237;
// matrix.py:305-309, 643-661; matrix_utils.py:314-317.)
// All are pure streaming kernels: 16-byte vector accesses, grid = multiple of the SM count.
#include "npw_common.cuh"

namespace npw {
namespace {

constexpr int EW_THREADS = 256;

inline int ew_grid(int64_t work_items, int per_thread = 4) {
  int64_t blocks = (work_items + static_cast<int64_t>(EW_THREADS) * per_thread - 1) / (static_cast<int64_t>(EW_THREADS) * per_thread);
  const int64_t cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

struct PtrPack {
  const double* p[8];
};

template <int COUNT>
__global__ void __launch_bounds__(EW_THREADS) addn_kernel(double* __restrict__ out, PtrPack in, int64_t nelem, int vec) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    const int64_t n2 = nelem >> 1;
    for (int64_t i = tid; i < n2; i += stride) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < COUNT; ++c) {
        const double2 v = reinterpret_cast<const double2*>(in.p[c])[i];
        acc.x += v.x;
        acc.y += v.y;
      }
      reinterpret_cast<double2*>(out)[i] = acc;
    }
    if (tid == 0 && (nelem & 1)) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < COUNT; ++c) acc += in.p[c][nelem - 1];
      out[nelem - 1] = acc;
    }
  } else {
    for (int64_t i = tid; i < nelem; i += stride) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < COUNT; ++c) acc += in.p[c][i];
      out[i] = acc;
    }
  }
}

__global__ void __launch_bounds__(EW_THREADS) mul_kernel(double* __restrict__ out, const double* __restrict__ x,
                                                          const double* __restrict__ y, int64_t nelem) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nelem; i += stride) out[i] = x[i] * y[i];
}

__global__ void __launch_bounds__(EW_THREADS) copy2d_kernel(double* __restrict__ dst, int64_t ldd,
                                                             const double* __restrict__ src, int64_t lds, int64_t rows,
                                                             int64_t cols, int vec) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    const int64_t c2 = cols >> 1;
    const int64_t total = rows * c2;
    for (int64_t i = tid; i < total; i += stride) {
      const int64_t r = i / c2, c = i - r * c2;
      reinterpret_cast<double2*>(dst + r * ldd)[c] = reinterpret_cast<const double2*>(src + r * lds)[c];
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t i = tid; i < total; i += stride) {
      const int64_t r = i / cols, c = i - r * cols;
      dst[r * ldd + c] = src[r * lds + c];
    }
  }
}

// dst[c, r] = src[r, c]; 32x32 tiles through padded shared memory so both sides coalesce.
__global__ void __launch_bounds__(256) transpose_kernel(double* __restrict__ dst, int64_t ldd,
                                                         const double* __restrict__ src, int64_t lds, int64_t rows,
                                                         int64_t cols) {
  __shared__ double tile[32][33];
  const int64_t tiles_c = (cols + 31) / 32;
  const int64_t tiles_r = (rows + 31) / 32;
  const int64_t ntiles = tiles_c * tiles_r;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t tr = t / tiles_c, tc = t - tr * tiles_c;
    const int64_t r0 = tr * 32, c0 = tc * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
      if (r < rows && c < cols) tile[ty + 8 * i][tx] = src[r * lds + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
      if (r < rows && c < cols) dst[c * ldd + r] = tile[tx][ty + 8 * i];
    }
    __syncthreads();
  }
}

__global__ void add_diag_kernel(double* A, int64_t lda, int64_t n, double v) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) A[i * lda + i] += v;
}

__global__ void __launch_bounds__(EW_THREADS) fill2d_kernel(double* A, int64_t lda, int64_t rows, int64_t cols, int mode,
                                                             double value) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    const bool keep = (mode == 1 && c >= r) || (mode == 2 && c <= r);
    if (!keep) A[r * lda + c] = value;
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(EW_THREADS) fill_random_kernel(double* A, int64_t lda, int64_t rows, int64_t cols,
                                                                  uint64_t seed, int64_t row0, int64_t col0) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    const uint64_t key = splitmix64(seed ^ splitmix64(static_cast<uint64_t>(row0 + r) * 0x100000001B3ull + static_cast<uint64_t>(col0 + c)));
    // 53 random bits -> uniform in (-1, 1)
    const double u = static_cast<double>(key >> 11) * (1.0 / 9007199254740992.0);
    A[r * lda + c] = 2.0 * u - 1.0;
  }
}

// Register-only DMMA issue loop: the fp64 tensor pipe's attainable rate (roofline denominator).
__global__ void __launch_bounds__(512) dmma_probe_kernel(double* out, int iters, double a0, double b0) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

template <int COUNT>
int launch_addn(double* out, const PtrPack& pk, int64_t nelem, int vec, cudaStream_t st) {
  addn_kernel<COUNT><<<ew_grid(nelem, 8), EW_THREADS, 0, st>>>(out, pk, nelem, vec);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

}  // namespace

int launch_copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, int trans,
                  cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return NPW_OK;
  if (trans) {
    const int64_t ntiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    const int grid = static_cast<int>(ntiles < 148 * 32 ? ntiles : 148 * 32);
    transpose_kernel<<<grid, 256, 0, st>>>(dst, ldd, src, lds, rows, cols);
  } else {
    const int vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0 && (ldd % 2 == 0) &&
                    (lds % 2 == 0) && (cols % 2 == 0);
    copy2d_kernel<<<ew_grid(rows * cols, 8), EW_THREADS, 0, st>>>(dst, ldd, src, lds, rows, cols, vec);
  }
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

int launch_fill2d(double* A, int64_t lda, int64_t rows, int64_t cols, int mode, double value, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return NPW_OK;
  fill2d_kernel<<<ew_grid(rows * cols, 8), EW_THREADS, 0, st>>>(A, lda, rows, cols, mode, value);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

}  // namespace npw

extern "C" {

int npw_addn_f64(double* out, const double* const* ptrs, int count, int64_t nelem, npw_stream_t stream) {
  if (!out) return -1;
  if (!ptrs) return -2;
  if (count < 1 || count > 8) return -3;
  if (nelem < 0) return -4;
  if (nelem == 0) return NPW_OK;
  npw::PtrPack pk;
  uintptr_t al = reinterpret_cast<uintptr_t>(out);
  for (int c = 0; c < 8; ++c) {
    pk.p[c] = c < count ? ptrs[c] : nullptr;
    if (c < count) {
      if (!ptrs[c]) return -2;
      al |= reinterpret_cast<uintptr_t>(ptrs[c]);
    }
  }
  const int vec = (al & 15u) == 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (count) {
    case 1: return npw::launch_addn<1>(out, pk, nelem, vec, st);
    case 2: return npw::launch_addn<2>(out, pk, nelem, vec, st);
    case 3: return npw::launch_addn<3>(out, pk, nelem, vec, st);
    case 4: return npw::launch_addn<4>(out, pk, nelem, vec, st);
    case 5: return npw::launch_addn<5>(out, pk, nelem, vec, st);
    case 6: return npw::launch_addn<6>(out, pk, nelem, vec, st);
    case 7: return npw::launch_addn<7>(out, pk, nelem, vec, st);
    default: return npw::launch_addn<8>(out, pk, nelem, vec, st);
  }
}

int npw_mul_f64(double* out, const double* x, const double* y, int64_t nelem, npw_stream_t stream) {
  if (!out) return -1;
  if (!x) return -2;
  if (!y) return -3;
  if (nelem < 0) return -4;
  if (nelem == 0) return NPW_OK;
  npw::mul_kernel<<<npw::ew_grid(nelem, 8), npw::EW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(out, x, y, nelem);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

int npw_copy2d_f64(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, int trans,
                   npw_stream_t stream) {
  if (!dst) return -1;
  if (ldd < (trans ? rows : cols)) return -2;
  if (!src) return -3;
  if (lds < cols) return -4;
  if (rows < 0) return -5;
  if (cols < 0) return -6;
  return npw::launch_copy2d(dst, ldd, src, lds, rows, cols, trans, static_cast<cudaStream_t>(stream));
}

int npw_add_diag_f64(double* A, int64_t lda, int64_t rows, int64_t cols, double lambdav, npw_stream_t stream) {
  if (!A) return -1;
  if (lda < cols) return -2;
  const int64_t n = rows < cols ? rows : cols;
  if (n <= 0) return NPW_OK;
  npw::add_diag_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, n, lambdav);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

int npw_fill2d_f64(double* A, int64_t lda, int64_t rows, int64_t cols, int mode, double value, npw_stream_t stream) {
  if (!A) return -1;
  if (lda < cols) return -2;
  if (mode < 0 || mode > 2) return -5;
  return npw::launch_fill2d(A, lda, rows, cols, mode, value, static_cast<cudaStream_t>(stream));
}

size_t npw_fp64_pipe_probe_bytes(int warps_per_sm) {
  if (warps_per_sm < 1 || warps_per_sm > 16) return 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return static_cast<size_t>(sms) * warps_per_sm * 32 * sizeof(double);
}

int npw_fp64_pipe_probe(double* scratch, int iters, int warps_per_sm, double* flops_out, npw_stream_t stream) {
  if (!scratch) return -1;
  if (iters < 1) return -2;
  if (warps_per_sm < 1 || warps_per_sm > 16) return -3;
  int dev = 0, sms = 0;
  NPW_CUDA_CHECK(cudaGetDevice(&dev));
  NPW_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  npw::dmma_probe_kernel<<<sms, warps_per_sm * 32, 0, static_cast<cudaStream_t>(stream)>>>(scratch, iters, 1.0, 1.0);
  NPW_LAUNCH_CHECK();
  if (flops_out) *flops_out = 2.0 * 256.0 * 16.0 * iters * warps_per_sm * sms;
  return NPW_OK;
}

int npw_fill_random_f64(double* A, int64_t lda, int64_t rows, int64_t cols, uint64_t seed, int64_t row0, int64_t col0,
                        npw_stream_t stream) {
  if (!A) return -1;
  if (lda < cols) return -2;
  if (rows <= 0 || cols <= 0) return NPW_OK;
  npw::fill_random_kernel<<<npw::ew_grid(rows * cols, 8), npw::EW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      A, lda, rows, cols, seed, row0, col0);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

}  // extern "C"
