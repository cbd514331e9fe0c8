// npw_qr_f64.cu — compact-WY Householder QR of a tall tile: kernels.qr_factor (kernels.py:127-130) → fast_qr
// (kernels.py:86-105, LAPACK dgeqrt3): A (m x n, m >= n) -> V (unit lower trapezoidal), T (n x n upper), R (n x n upper),
// Q = I - V T V^T.  Same Householder convention as LAPACK dlarfg/dlarft, so V, T, R match the reference up to rounding.
//
// Blocked right-looking algorithm, panel width 32:
//   panel   : ONE cooperative kernel factors the (m-j0) x 32 panel.  Rows are split over the CTAs of the grid; a warp
//             reads a whole panel row as one coalesced 256-byte request (lane = column).  Per column there is a single
//             pass over the rows that (i) applies reflector j, (ii) accumulates the dot products column j+1 needs
//             (norm + w = P^T x) and (iii) the Gram entries V[:,0:j]^T v_j that dlarft needs, all in one 32-lane
//             vector, followed by ONE grid-wide reduction (grid.sync) per column.
//   update  : Wt = C^T V_p (split-K "TN" kernel: k = rows, deterministic two-phase reduction), Wt := Wt T_p,
//             C -= V_p Wt^T through the DMMA GEMM core (NT, k = 32).
//   T       : Gram matrix G = V^T V (same split-K kernel), then T[0:j0, panel] = -T[0:j0,0:j0] G[0:j0,panel] T_pp.
#include <cooperative_groups.h>

#include "npw_common.cuh"

namespace cg = cooperative_groups;

namespace npw {

int launch_gemm(double* C, int64_t ldc, const double* C0, int64_t ldc0, const double* A, int64_t lda, int transA,
                const double* B, int64_t ldb, int transB, int64_t m, int64_t n, int64_t k, double alpha,
                double beta, int lower_only, cudaStream_t stream);
int launch_copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, int trans,
                  cudaStream_t st);
int launch_fill2d(double* A, int64_t lda, int64_t rows, int64_t cols, int mode, double value, cudaStream_t st);

namespace {

constexpr int QW = 32;            // panel width
constexpr int QTHREADS = 256;     // 8 warps
constexpr int QWARPS = QTHREADS / 32;
constexpr int MAX_GRID = 256;

struct PanelArgs {
  double* V;       // m x n working matrix (row-major, ldv); panel columns [j0, j0+w)
  int64_t ldv;
  int m, j0, w;
  double* R;       // n x n output, ldr
  int64_t ldr;
  double* T;       // n x n output, ldt (only the w x w diagonal block of this panel is written)
  int64_t ldt;
  double* tau;     // n
  double* scratch; // 2 x MAX_GRID x 64 doubles: per-CTA partial vectors (double-buffered) + pivot rows
};

// scratch layout per parity buffer: [cta][32] partial sums, then [32] pivot row, starting at parity * SCR_STRIDE
constexpr int SCR_STRIDE = MAX_GRID * QW + QW;

// SMEM = true: the CTA's row chunk of the panel (<= QR_SMEM_ROWS rows x 32 columns) is loaded into shared memory once,
// all 32 column passes run there, and the chunk is written back once (2 global passes instead of 64).
constexpr int QR_SMEM_ROWS = 768;                    // 768 x 32 x 8 B = 192 KB

template <bool SMEM>
__global__ void __launch_bounds__(QTHREADS) qr_panel_kernel(PanelArgs p) {
  extern __shared__ __align__(16) double s_chunk[];
  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x;
  const int cta = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int rows = p.m - p.j0;                       // panel rows (global rows j0 .. m-1)
  const int per = (rows + G - 1) / G;
  const int r_lo = p.j0 + cta * per;                 // this CTA's global row range [r_lo, r_hi)
  const int r_hi = min(p.m, r_lo + per);
  const int w = p.w;
  const bool lane_ok = lane < w;
  double* Vp = p.V + p.j0;                           // column offset of the panel

  __shared__ double s_part[QWARPS][QW];
  __shared__ double s_vec[QW];                       // reduced vector g
  __shared__ double s_prow[QW];
  __shared__ double s_T[QW][QW + 1];
  __shared__ double s_tau[QW];
  __shared__ double s_sc[4];                         // tau_j, scale_j, beta_j

  for (int e = threadIdx.x; e < QW * (QW + 1); e += QTHREADS) (&s_T[0][0])[e] = 0.0;

  auto rowptr = [&](int r) -> double* {
    return SMEM ? s_chunk + static_cast<size_t>(r - r_lo) * QW : Vp + static_cast<int64_t>(r) * p.ldv;
  };
  if (SMEM) {
    for (int r = r_lo + warp; r < r_hi; r += QWARPS)
      s_chunk[static_cast<size_t>(r - r_lo) * QW + lane] = lane_ok ? Vp[static_cast<int64_t>(r) * p.ldv + lane] : 0.0;
    __syncthreads();
  }

  // ---- initial partials for column 0: g[c] = sum_{r > j0} x_r * P[r][c], x = column 0
  {
    double acc = 0.0;
    for (int r = r_lo + warp; r < r_hi; r += QWARPS) {
      if (r <= p.j0) continue;
      const double v = lane_ok ? rowptr(r)[lane] : 0.0;
      const double x = __shfl_sync(0xffffffffu, v, 0);
      acc = fma(x, v, acc);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      double s = 0.0;
      for (int q = 0; q < QWARPS; ++q) s += s_part[q][lane];
      p.scratch[cta * QW + lane] = s;
      if (p.j0 >= r_lo && p.j0 < r_hi)
        p.scratch[MAX_GRID * QW + lane] = lane_ok ? rowptr(p.j0)[lane] : 0.0;
    }
  }

  for (int j = 0; j < w; ++j) {
    const int gj = p.j0 + j;                         // pivot row / column (global)
    const int par = j & 1;
    __threadfence();
    grid.sync();
    // ---- reduce the partial vectors of all CTAs (every CTA redundantly)
    const double* scr = p.scratch + par * SCR_STRIDE;
    {
      double s = 0.0;
      {
        // all partial vectors are fetched with independent loads (fixed order of addition: deterministic result)
        double v[4];
        int q = warp;
        for (; q + 3 * QWARPS < G; q += 4 * QWARPS) {
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = __ldcg(scr + (q + u * QWARPS) * QW + lane);
          s += v[0]; s += v[1]; s += v[2]; s += v[3];
        }
        for (; q < G; q += QWARPS) s += __ldcg(scr + q * QW + lane);
      }
      s_part[warp][lane] = s;
      __syncthreads();
      if (warp == 0) {
        double t = 0.0;
        for (int q = 0; q < QWARPS; ++q) t += s_part[q][lane];
        s_vec[lane] = t;
        s_prow[lane] = __ldcg(scr + MAX_GRID * QW + lane);
      }
      __syncthreads();
    }
    // lanes c >= j of s_vec: dot products for column j (c == j: squared norm below the pivot)
    // lanes c <  j-1 ... of the PREVIOUS column's Gram are folded in below (see z handling)
    // ---- Householder scalars (dlarfg)
    if (threadIdx.x == 0) {
      const double alpha = s_prow[j];
      const double xn2 = s_vec[j];
      double tau = 0.0, scale = 0.0, beta = alpha;
      if (xn2 > 0.0) {
        const double nrm = sqrt(alpha * alpha + xn2);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      s_sc[0] = tau; s_sc[1] = scale; s_sc[2] = beta;
      s_tau[j] = tau;
    }
    __syncthreads();
    const double tau = s_sc[0], scale = s_sc[1], beta = s_sc[2];
    // w_c = v^T P[:, c] = prow[c] + scale * g[c]   (c > j)
    const double wc = (lane > j && lane_ok) ? s_prow[lane] + scale * s_vec[lane] : 0.0;
    const double tw = tau * wc;

    // ---- one pass over this CTA's rows: apply reflector j, store v_j, accumulate for column j+1 and the Gram of v_j
    double acc = 0.0;   // lane c > j: sum_{r > gj+1} x'_r P'[r][c] (x' = updated column j+1); lane i < j: sum v_i[r] v_j[r]
    const int rs = max(r_lo, gj);
    for (int r = rs + warp; r < r_hi; r += QWARPS) {
      double* rowp = rowptr(r);
      double v = lane_ok ? rowp[lane] : 0.0;
      if (r == gj) {
        // pivot row: v_j[gj] = 1; R(gj, c) = P[gj][c] - tau * w_c; diagonal becomes beta
        if (lane > j) v -= tw;
        if (lane == j) v = beta;
        if (lane_ok && lane >= j) {
          p.R[static_cast<int64_t>(gj) * p.ldr + p.j0 + lane] = v;      // row gj of R (panel columns)
          rowp[lane] = (lane == j) ? 1.0 : 0.0;                        // explicit unit-lower V
        }
        // Gram contribution of the pivot row: v_i[gj] * 1 for i < j
        if (lane < j) acc += v;
        continue;
      }
      const double x = __shfl_sync(0xffffffffu, v, j);
      const double vr = x * scale;                    // v_j[r]
      if (lane > j) v = fma(-vr, tw, v);
      if (lane == j) v = vr;
      if (lane_ok && lane >= j) rowp[lane] = v;
      if (lane < j) acc = fma(v, vr, acc);            // V[r][i] * v_j[r]
      if (j + 1 < w && r > gj + 1) {
        const double xn = __shfl_sync(0xffffffffu, v, j + 1);
        if (lane > j) acc = fma(xn, v, acc);
      } else if (j + 1 < w) {
        __shfl_sync(0xffffffffu, v, j + 1);           // keep the warp converged on the shuffle
      }
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      double s = 0.0;
      for (int q = 0; q < QWARPS; ++q) s += s_part[q][lane];
      double* out = p.scratch + (par ^ 1) * SCR_STRIDE;
      out[cta * QW + lane] = s;
      const int gn = gj + 1;                          // next pivot row
      if (j + 1 < w && gn >= r_lo && gn < r_hi && gn < p.m)
        out[MAX_GRID * QW + lane] = lane_ok ? rowptr(gn)[lane] : 0.0;
    }
    // ---- the Gram vector z_i = V[:,i]^T v_{j-1} (i < j-1) reduced this round belongs to column j-1 of T
    if (j > 0 && cta == 0 && threadIdx.x == 0) {
      const int jj = j - 1;
      const double tj = s_tau[jj];
      for (int i = 0; i < jj; ++i) {
        double t = 0.0;
        for (int q = i; q < jj; ++q) t += s_T[i][q] * s_vec[q];   // T[0:jj,0:jj] upper-triangular times z
        s_T[i][jj] = -tj * t;
      }
      s_T[jj][jj] = tj;
    }
    __syncthreads();
  }
  // ---- last column's Gram vector
  __threadfence();
  grid.sync();
  {
    const double* scr = p.scratch + (w & 1) * SCR_STRIDE;
    double s = 0.0;
    for (int q = warp; q < G; q += QWARPS) s += __ldcg(scr + q * QW + lane);
    s_part[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
      double t = 0.0;
      for (int q = 0; q < QWARPS; ++q) t += s_part[q][lane];
      s_vec[lane] = t;
    }
    __syncthreads();
  }
  if (cta == 0) {
    if (threadIdx.x == 0) {
      const int jj = w - 1;
      const double tj = s_tau[jj];
      for (int i = 0; i < jj; ++i) {
        double t = 0.0;
        for (int q = i; q < jj; ++q) t += s_T[i][q] * s_vec[q];
        s_T[i][jj] = -tj * t;
      }
      s_T[jj][jj] = tj;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < w * w; e += QTHREADS) {
      const int i = e / w, c = e - i * w;
      p.T[static_cast<int64_t>(p.j0 + i) * p.ldt + p.j0 + c] = (c >= i) ? s_T[i][c] : 0.0;
    }
    for (int e = threadIdx.x; e < w; e += QTHREADS) p.tau[p.j0 + e] = s_tau[e];
  }
  if (SMEM) {                                        // write the factored chunk back (V below the diagonal, explicit unit/zeros)
    __syncthreads();
    for (int r = r_lo + warp; r < r_hi; r += QWARPS)
      if (lane_ok) Vp[static_cast<int64_t>(r) * p.ldv + lane] = s_chunk[static_cast<size_t>(r - r_lo) * QW + lane];
  }
}

// ------------------------------------------------------------------------------------------------
// Split-K "TN" product for tall operands: P[s] = A[ks:ke, 0:M]^T * B[ks:ke, 0:N], 64x64 output tiles.
// A is rows x M (lda), B is rows x N (ldb); the reduction dimension is the (long) row index.
// ------------------------------------------------------------------------------------------------
constexpr int TT = 64, TKC = 16;

__global__ void __launch_bounds__(256) tn_partial_kernel(double* __restrict__ P, const double* __restrict__ A, int64_t lda,
                                                         const double* __restrict__ B, int64_t ldb, int rows, int M, int N,
                                                         int kchunk) {
  __shared__ double As[TKC][TT + 4];
  __shared__ double Bs[TKC][TT + 4];
  const int tm = blockIdx.y * TT, tn = blockIdx.x * TT, s = blockIdx.z;
  const int k0 = s * kchunk, k1 = min(rows, k0 + kchunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4] = {};
  for (int kb = k0; kb < k1; kb += TKC) {
    for (int e = threadIdx.x; e < TKC * TT; e += 256) {
      const int kk = e / TT, c = e - kk * TT;
      const int r = kb + kk;
      As[kk][c] = (r < k1 && tm + c < M) ? A[static_cast<int64_t>(r) * lda + tm + c] : 0.0;
      Bs[kk][c] = (r < k1 && tn + c < N) ? B[static_cast<int64_t>(r) * ldb + tn + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TKC; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* Ps = P + static_cast<int64_t>(s) * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tm + ty + 16 * i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tn + tx + 16 * j;
      if (c < N) Ps[static_cast<int64_t>(r) * N + c] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256) tn_reduce_kernel(double* __restrict__ C, int64_t ldc, const double* __restrict__ P,
                                                        int M, int N, int nsplit) {
  const int64_t total = static_cast<int64_t>(M) * N;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < nsplit; ++q) s += P[static_cast<int64_t>(q) * total + e];   // fixed order: deterministic
    const int64_t r = e / N, c = e - r * N;
    C[r * ldc + c] = s;
  }
}

inline int tn_nsplit(int64_t rows, int64_t M, int64_t N) {
  const int64_t tiles = ((M + TT - 1) / TT) * ((N + TT - 1) / TT);
  int64_t want = (4 * 148 + tiles - 1) / tiles;           // ~4 CTAs per SM in total
  const int64_t maxs = (rows + 255) / 256;                // at least 256 rows per split
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return static_cast<int>(want);
}

// C (M x N, ldc) = A^T B ; partials must hold nsplit*M*N doubles
int launch_tn(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t rows, int64_t M,
              int64_t N, double* partials, cudaStream_t st) {
  if (M <= 0 || N <= 0) return NPW_OK;
  const int ns = tn_nsplit(rows, M, N);
  const int kchunk = static_cast<int>((rows + ns - 1) / ns);
  dim3 grid(static_cast<unsigned>((N + TT - 1) / TT), static_cast<unsigned>((M + TT - 1) / TT), static_cast<unsigned>(ns));
  tn_partial_kernel<<<grid, 256, 0, st>>>(partials, A, lda, B, ldb, static_cast<int>(rows), static_cast<int>(M),
                                          static_cast<int>(N), kchunk);
  NPW_LAUNCH_CHECK();
  const int64_t total = M * N;
  int rb = static_cast<int>((total + 255) / 256);
  if (rb > 148 * 8) rb = 148 * 8;
  tn_reduce_kernel<<<rb, 256, 0, st>>>(C, ldc, partials, static_cast<int>(M), static_cast<int>(N), ns);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct QrWork {
  size_t scratch, tau, wt, tmp, gram, partials, total;
};

QrWork qr_work_layout(int64_t m, int64_t n) {
  QrWork w;
  size_t off = 0;
  w.scratch = off; off += align256(2 * SCR_STRIDE * sizeof(double));
  w.tau = off; off += align256(static_cast<size_t>(n) * sizeof(double));
  w.wt = off; off += align256(static_cast<size_t>(n) * QW * sizeof(double));         // Wt: (n - j0 - w) x w
  w.tmp = off; off += align256(static_cast<size_t>(n) * QW * sizeof(double));        // Wt T_p  /  G T_pp
  w.gram = off; off += align256(static_cast<size_t>(n) * n * sizeof(double));        // G = V^T V
  const int ns_g = tn_nsplit(m, n, n), ns_w = tn_nsplit(m, n, QW);
  size_t pb = static_cast<size_t>(ns_g) * n * n;
  const size_t pw = static_cast<size_t>(ns_w) * n * QW;
  if (pw > pb) pb = pw;
  // nsplit depends on the row count of each call; bound it by the largest value tn_nsplit can return for these widths
  const size_t pmax = static_cast<size_t>(256) * n * QW;
  if (pmax > pb) pb = pmax;
  w.partials = off; off += align256(pb * sizeof(double));
  w.total = off;
  return w;
}

int g_coop_grid[64] = {};

}  // namespace
}  // namespace npw

extern "C" {

size_t npw_geqrt_work_bytes(int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  return npw::qr_work_layout(m, n).total;
}

int npw_geqrt_f64(double* V, int64_t ldv, double* T, int64_t ldt, double* R, int64_t ldr, const double* A, int64_t lda,
                  int64_t m, int64_t n, void* work, npw_stream_t stream) {
  using namespace npw;
  if (m == 0 || n == 0) return NPW_OK;
  if (!V) return -1;
  if (ldv < n) return -2;
  if (!T) return -3;
  if (ldt < n) return -4;
  if (!R) return -5;
  if (ldr < n) return -6;
  if (!A) return -7;
  if (lda < n) return -8;
  if (m < n || m > INT32_MAX) return -9;
  if (n < 0) return -10;
  if (!work) return -11;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const QrWork wl = qr_work_layout(m, n);
  char* wb = static_cast<char*>(work);
  double* scratch = reinterpret_cast<double*>(wb + wl.scratch);
  double* tau = reinterpret_cast<double*>(wb + wl.tau);
  double* Wt = reinterpret_cast<double*>(wb + wl.wt);
  double* tmp = reinterpret_cast<double*>(wb + wl.tmp);
  double* gram = reinterpret_cast<double*>(wb + wl.gram);
  double* partials = reinterpret_cast<double*>(wb + wl.partials);

  int dev = 0;
  NPW_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && g_coop_grid[dev] == 0) {
    int coop = 0, sms = 0, per_sm = 0;
    NPW_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    NPW_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!coop) {
      set_error("device does not support cooperative launch");
      return NPW_ERR_UNSUPPORTED;
    }
    NPW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qr_panel_kernel<false>, QTHREADS, 0));
    NPW_CUDA_CHECK(cudaFuncSetAttribute(qr_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        QR_SMEM_ROWS * QW * static_cast<int>(sizeof(double))));
    int g = sms * (per_sm > 0 ? 1 : 0);
    if (g > MAX_GRID) g = MAX_GRID;
    if (g < 1) g = 1;
    g_coop_grid[dev] = g;
  }
  const int max_grid = dev < 64 ? g_coop_grid[dev] : 1;

  int rc;
  if (V != A) {
    rc = launch_copy2d(V, ldv, A, lda, m, n, 0, st);
    if (rc) return rc;
  }
  rc = launch_fill2d(R, ldr, n, n, 0, 0.0, st);
  if (rc) return rc;
  rc = launch_fill2d(T, ldt, n, n, 0, 0.0, st);
  if (rc) return rc;

  for (int64_t j0 = 0; j0 < n; j0 += QW) {
    const int w = static_cast<int>(n - j0 < QW ? n - j0 : QW);
    const int64_t rows = m - j0;
    PanelArgs pa;
    pa.V = V; pa.ldv = ldv; pa.m = static_cast<int>(m); pa.j0 = static_cast<int>(j0); pa.w = w;
    pa.R = R; pa.ldr = ldr; pa.T = T; pa.ldt = ldt; pa.tau = tau; pa.scratch = scratch;
    int g = static_cast<int>((rows + 127) / 128);           // >= 128 panel rows per CTA
    if (g > max_grid) g = max_grid;
    if (g < 1) g = 1;
    void* kargs[] = {&pa};
    const int64_t per = (rows + g - 1) / g;                 // rows per CTA, as the kernel computes it
    if (per <= QR_SMEM_ROWS) {
      NPW_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(qr_panel_kernel<true>), dim3(g), dim3(QTHREADS), kargs,
                                                 static_cast<size_t>(per) * QW * sizeof(double), st));
    } else {
      NPW_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(qr_panel_kernel<false>), dim3(g), dim3(QTHREADS), kargs, 0, st));
    }
    count_launch();
    const int64_t nt = n - j0 - w;                          // trailing columns
    if (nt > 0) {
      double* C = V + j0 * ldv + j0 + w;                    // rows x nt, starts at row j0
      const double* Vp = V + j0 * ldv + j0;                 // rows x w (explicit unit-lower after the panel kernel)
      // Wt = C^T V_p  (nt x w)
      rc = launch_tn(Wt, QW, C, ldv, Vp, ldv, rows, nt, w, partials, st);
      if (rc) return rc;
      // tmp = Wt T_p  (applying Q^T = I - V T^T V^T:  C -= V (T^T (V^T C))  <=>  C -= V (Wt T)^T)
      rc = launch_gemm(tmp, QW, nullptr, 0, Wt, QW, 0, T + j0 * ldt + j0, ldt, 0, nt, w, w, 1.0, 0.0, 0, st);
      if (rc) return rc;
      // C -= V_p tmp^T
      rc = launch_gemm(C, ldv, C, ldv, Vp, ldv, 0, tmp, QW, 1, rows, nt, w, -1.0, 1.0, 0, st);
      if (rc) return rc;
      // rows j0 .. j0+w-1 of the trailing columns are final rows of R
      rc = launch_copy2d(R + j0 * ldr + j0 + w, ldr, C, ldv, w, nt, 0, st);
      if (rc) return rc;
    }
  }
  // explicit V: zero strictly above the diagonal (the panel kernel already wrote the unit diagonal)
  rc = launch_fill2d(V, ldv, n, n, 2, 0.0, st);
  if (rc) return rc;
  // off-diagonal blocks of T from the Gram matrix of V
  if (n > QW) {
    rc = launch_tn(gram, n, V, ldv, V, ldv, m, n, n, partials, st);
    if (rc) return rc;
    for (int64_t j0 = QW; j0 < n; j0 += QW) {
      const int w = static_cast<int>(n - j0 < QW ? n - j0 : QW);
      // tmp (j0 x w) = G[0:j0, panel] T_pp
      rc = launch_gemm(tmp, QW, nullptr, 0, gram + j0, n, 0, T + j0 * ldt + j0, ldt, 0, j0, w, w, 1.0, 0.0, 0, st);
      if (rc) return rc;
      // T[0:j0, panel] = -T[0:j0, 0:j0] tmp
      rc = launch_gemm(T + j0, ldt, nullptr, 0, T, ldt, 0, tmp, QW, 0, j0, w, j0, -1.0, 0.0, 0, st);
      if (rc) return rc;
    }
  }
  return NPW_OK;
}

size_t npw_tpqrt_work_bytes(int64_t n) {
  if (n <= 0) return 0;
  return npw::align256(static_cast<size_t>(2 * n) * n * sizeof(double)) + npw_geqrt_work_bytes(2 * n, n);
}

// QR of two stacked upper-triangular factors.  Column j of [triu(R0); triu(R1)] has non-zeros in row j of the top block
// and rows 0..j of the bottom block only, so the general panel factorisation of the 2n x n stack produces exactly
// dtpqrt's reflectors: top half of V = I, bottom half = V2 (upper triangular), same T and R.
int npw_tpqrt_f64(double* V2, int64_t ldv, double* T, int64_t ldt, double* R, int64_t ldr, const double* R0, int64_t ld0,
                  const double* R1, int64_t ld1, int64_t n, void* work, npw_stream_t stream) {
  using namespace npw;
  if (n == 0) return NPW_OK;
  if (!V2) return -1;
  if (ldv < n) return -2;
  if (!T) return -3;
  if (ldt < n) return -4;
  if (!R) return -5;
  if (ldr < n) return -6;
  if (!R0) return -7;
  if (ld0 < n) return -8;
  if (!R1) return -9;
  if (ld1 < n) return -10;
  if (n < 0 || 2 * n > INT32_MAX) return -11;
  if (!work) return -12;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* W = static_cast<double*>(work);                       // the 2n x n stack, leading dimension n
  double* Wlo = W + n * n;
  void* qr_work = static_cast<char*>(work) + align256(static_cast<size_t>(2 * n) * n * sizeof(double));
  int rc = launch_copy2d(W, n, R0, ld0, n, n, 0, st);
  if (rc) return rc;
  rc = launch_copy2d(Wlo, n, R1, ld1, n, n, 0, st);
  if (rc) return rc;
  rc = launch_fill2d(W, n, n, n, 1, 0.0, st);                   // dtpqrt reads the upper triangles only
  if (rc) return rc;
  rc = launch_fill2d(Wlo, n, n, n, 1, 0.0, st);
  if (rc) return rc;
  rc = npw_geqrt_f64(W, n, T, ldt, R, ldr, W, n, 2 * n, n, qr_work, stream);
  if (rc) return rc;
  return launch_copy2d(V2, ldv, Wlo, n, n, n, 0, st);
}

}  // extern "C"
