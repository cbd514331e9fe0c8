// npw_gemm_f64.cu — the fp64 contraction core of the LambdaPACK tile kernels.
//
// Replaces the BLAS dgemm behind kernels.syrk (kernels.py:212-215, `s - x.dot(y.T)`),
// kernels.gemm (kernels.py:239-244) and the level-3 updates inside chol/trsm.
//
// Design (B200 / sm_100a):
//  * tcgen05.mma has no fp64 kind (ptxas rejects kind::f64), so the fp64 tensor pipe is
//    reached with mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4 (all larger PTX shapes decompose
//    into it on sm_100a).  Accumulators therefore live in registers, not TMEM.
//  * "NT" form: C[m,n] = alpha * A[m,k] * B[n,k]^T + beta * C0 — both operands K-contiguous,
//    which is what row-major tiles give for syrk/trsm/potrf updates.
//  * One TMA producer warp streams 128x16 (A) and 128x16 (B) fp64 boxes (128-byte rows,
//    SWIZZLE_128B) through a 6-stage shared-memory ring guarded by full/empty mbarriers.
//  * 8 MMA warps (2 x 4), warp tile 64 x 32 = 8 x 4 DMMA tiles, 128 accumulator registers
//    per lane.  Fragments are read with conflict-free LDS.128: lane (g = lane/4, c = lane%4)
//    reads 16-byte chunk (2c+h) ^ g of row g, i.e. k-slot c of step (h,e) carries
//    k = 4c + 2h + e; A and B use the same slot->k permutation so the product is unchanged.
//  * Grid = ceil(m/128) * ceil(n/128) CTAs, 1 CTA/SM (193 KB smem), rastered in 8-row groups
//    so concurrently resident CTAs share operand panels in the 126 MB L2.
#include <stdlib.h>

#include "npw_common.cuh"

namespace npw {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int MMA_WARPS = 8;
// 2 MMA warpgroups + 1 producer warpgroup (only its first lane issues TMA).  Register
// allocation is per warpgroup: the kernel is compiled for 168 regs/thread (65536 / 384) and
// rebalanced at run time with setmaxnreg: producer 24, MMA warps 240.
constexpr int NTHREADS = (MMA_WARPS + 4) * 32;
constexpr int REGS_PRODUCER = 24;
constexpr int REGS_MMA = 240;
constexpr int STAGES = 6;
constexpr int STAGE_A = BM * BK * 8;
constexpr int STAGE_B = BN * BK * 8;
constexpr int STAGE_BYTES = STAGE_A + STAGE_B;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;
constexpr int RASTER_GROUP = 8;

// tile-rows per raster group; NPW_B200_RASTER overrides the default (tuning knob, read once)
inline int l2_hint_mode() {
  static int v = [] {
    const char* e = getenv("NPW_B200_L2HINT");
    return e ? atoi(e) : 0;
  }();
  return v;
}

inline int raster_group() {
  static int v = [] {
    const char* e = getenv("NPW_B200_RASTER");
    const int x = e ? atoi(e) : RASTER_GROUP;
    return x >= 1 && x <= 64 ? x : RASTER_GROUP;
  }();
  return v;
}

struct GemmArgs {
  double* C;
  const double* C0;
  int64_t ldc, ldc0;
  int m, n, k;
  double alpha, beta;
  int lower_only;
  int grid_m, grid_n;
  int raster;          // tile-rows per raster group
  int l2hint;          // 1: A panel evict_last, C0 / C evict_first (the row panels of a raster group are what L2 should keep)
  int vec_ok;
};

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_nt_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmArgs p) {
  // SWIZZLE_128B needs 1024-byte aligned stage buffers; the kernel has no static shared
  // memory, so the dynamic window starts at the (1024-aligned) base of the CTA's allocation.
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  // grouped raster: RASTER_GROUP tile-rows per group, column-major inside a group
  int tile_m, tile_n;
  {
    const int lin = blockIdx.x;
    const int RG = p.raster;
    const int group_sz = RG * p.grid_n;
    const int gid = lin / group_sz;
    const int first_m = gid * RG;
    const int gm = min(p.grid_m - first_m, RG);
    const int r = lin - gid * group_sz;
    tile_m = first_m + r % gm;
    tile_n = r / gm;
  }
  if (p.lower_only && tile_n * BN > tile_m * BM + (BM - 1)) {
    // tile strictly above the diagonal: not computed.  Out of place, C0 is passed through so that the
    // output tile is fully defined (in place there is nothing to do).
    if (p.C0 != nullptr && p.C0 != p.C) {
      for (int e = threadIdx.x; e < BM * BN; e += NTHREADS) {
        const int r = tile_m * BM + e / BN, c = tile_n * BN + e % BN;
        if (r < p.m && c < p.n) p.C[static_cast<int64_t>(r) * p.ldc + c] = p.C0[static_cast<int64_t>(r) * p.ldc0 + c];
      }
    }
    return;
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MMA_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int kt = (p.k + BK - 1) / BK;

  if (warp >= MMA_WARPS) {
    // ------------------------------------------------------------ TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
    if (warp == MMA_WARPS && lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int it = 0; it < kt; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* st = smem + s * STAGE_BYTES;
        if (p.l2hint) {
          tma_load_2d_hint(st, &tmA, &full[s], it * BK, tile_m * BM, L2_EVICT_LAST);
          tma_load_2d_hint(st + STAGE_A, &tmB, &full[s], it * BK, tile_n * BN, L2_EVICT_NORMAL);
        } else {
          tma_load_2d(st, &tmA, &full[s], it * BK, tile_m * BM);
          tma_load_2d(st + STAGE_A, &tmB, &full[s], it * BK, tile_n * BN);
        }
      }
    }
    return;
  }

  // ---------------------------------------------------------------- MMA warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_MMA));
  const int wm = warp >> 2;  // 0..1  -> rows  wm*64
  const int wn = warp & 3;   // 0..3  -> cols  wn*32
  const int g = lane >> 2;   // 0..7
  const int c = lane & 3;    // 0..3

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // warm L2 with this CTA's C0 tile (read only in the epilogue, ~0.5 ms from now): two 128-byte lines per row
  if (p.beta != 0.0 && p.C0 != nullptr && c < 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = tile_m * BM + wm * 64 + g + i * 8;
      const int col = tile_n * BN + wn * 32 + 16 * c;
      if (row < p.m && col < p.n)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.C0 + static_cast<int64_t>(row) * p.ldc0 + col));
    }
  }

  const uint32_t a_off = (wm * 64 + g) * 128;
  const uint32_t b_off = STAGE_A + (wn * 32 + g) * 128;
  const uint32_t sw[2] = {static_cast<uint32_t>(((2 * c) ^ g) << 4), static_cast<uint32_t>(((2 * c + 1) ^ g) << 4)};

  // Software-pipelined across stage boundaries: the fragments of (stage it+1, h=0) are fetched — including the wait on
  // that stage's full barrier — while the last 64 DMMAs of stage `it` are still being issued, so the tensor pipe does
  // not drain at every k-block.
  auto load_frag = [&](const uint8_t* st, int h, double2 (&a)[8], double2 (&b)[4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const double2*>(st + a_off + i * 1024 + sw[h]);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(st + b_off + j * 1024 + sw[h]);
  };
  double2 a0[8], b0[4], a1[8], b1[4];
  if (kt > 0) {
    mbar_wait(&full[0], 0);
    load_frag(smem, 0, a0, b0);
  }
  for (int it = 0; it < kt; ++it) {
    const int s = it % STAGES;
    const uint8_t* st = smem + s * STAGE_BYTES;
    load_frag(st, 1, a1, b1);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a0[i].x, b0[j].x);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a0[i].y, b0[j].y);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a1[i].x, b1[j].x);
    // every fragment of stage s is in registers (the DMMAs above consumed them): hand the slot back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (it + 1 < kt) {
      const int s2 = (it + 1) % STAGES;
      mbar_wait(&full[s2], ((it + 1) / STAGES) & 1);
      load_frag(smem + s2 * STAGE_BYTES, 0, a0, b0);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a1[i].y, b1[j].y);
  }

  // ----------------------------------------------------------------- epilogue
  const int row0 = tile_m * BM + wm * 64 + g;
  const int col0 = tile_n * BN + wn * 32 + 2 * c;
  const bool has_c0 = (p.beta != 0.0) && (p.C0 != nullptr);
  const bool interior = p.vec_ok && (tile_m * BM + BM <= p.m) && (tile_n * BN + BN <= p.n);
  if (interior) {
    // Fast path (full tile, 16-byte aligned): C0 may alias C, which would make the compiler serialise every load
    // behind the previous store; instead 16 independent 16-byte loads (4 row groups x 4 column groups) are put in
    // flight before anything is stored (the fragment registers are free by now).
#pragma unroll
    for (int ib = 0; ib < 8; ib += 4) {
      double2 sv[4][4];
      if (has_c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sv[i][j] = p.l2hint ? ld_global_hint(p.C0 + static_cast<int64_t>(row0 + (ib + i) * 8) * p.ldc0 + col0 + j * 8, L2_EVICT_FIRST)
                                : *reinterpret_cast<const double2*>(p.C0 + static_cast<int64_t>(row0 + (ib + i) * 8) * p.ldc0 + col0 + j * 8);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double2 o;
          if (has_c0) {
            o.x = fma(p.alpha, acc[ib + i][j][0], p.beta * sv[i][j].x);
            o.y = fma(p.alpha, acc[ib + i][j][1], p.beta * sv[i][j].y);
          } else {
            o.x = p.alpha * acc[ib + i][j][0];
            o.y = p.alpha * acc[ib + i][j][1];
          }
          if (p.l2hint) st_global_hint(p.C + static_cast<int64_t>(row0 + (ib + i) * 8) * p.ldc + col0 + j * 8, o, L2_EVICT_FIRST);
          else *reinterpret_cast<double2*>(p.C + static_cast<int64_t>(row0 + (ib + i) * 8) * p.ldc + col0 + j * 8) = o;
        }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + i * 8;
    if (row >= p.m) continue;
    double* crow = p.C + static_cast<int64_t>(row) * p.ldc;
    const double* srow = has_c0 ? p.C0 + static_cast<int64_t>(row) * p.ldc0 : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + j * 8;
      if (col >= p.n) continue;
      if (p.vec_ok && col + 1 < p.n) {
        double2 sv = make_double2(0.0, 0.0);
        if (has_c0) sv = *reinterpret_cast<const double2*>(srow + col);
        double2 o;
        o.x = fma(p.alpha, acc[i][j][0], p.beta * sv.x);
        o.y = fma(p.alpha, acc[i][j][1], p.beta * sv.y);
        if (!has_c0) { o.x = p.alpha * acc[i][j][0]; o.y = p.alpha * acc[i][j][1]; }
        *reinterpret_cast<double2*>(crow + col) = o;
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (col + e < p.n) {
            const double sv = has_c0 ? srow[col + e] : 0.0;
            crow[col + e] = has_c0 ? fma(p.alpha, acc[i][j][e], p.beta * sv) : p.alpha * acc[i][j][e];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Generic fallback: any transposition, any leading dimension/alignment, small tiles.
// 64x64 CTA tile, 256 threads, 4x4 micro-tile per thread, plain DFMA.  Used for the
// reference's small test tiles (8..64) and for operands TMA cannot describe.
// ------------------------------------------------------------------------------------
constexpr int GB = 64, GK = 16;

template <int TA, int TB>
__global__ void __launch_bounds__(256)
gemm_generic_kernel(double* __restrict__ C, int64_t ldc, const double* C0, int64_t ldc0,
                    const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                    int m, int n, int k, double alpha, double beta, int lower_only) {
  __shared__ double As[GK][GB + 1];
  __shared__ double Bs[GK][GB + 1];
  const int bm = blockIdx.y * GB, bn = blockIdx.x * GB;
  if (lower_only && bn > bm + GB - 1) {
    if (C0 != nullptr && C0 != C) {
      for (int e = threadIdx.x; e < GB * GB; e += 256) {
        const int r = bm + e / GB, c = bn + e % GB;
        if (r < m && c < n) C[static_cast<int64_t>(r) * ldc + c] = C0[static_cast<int64_t>(r) * ldc0 + c];
      }
    }
    return;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += GK) {
    for (int e = threadIdx.x; e < GB * GK; e += 256) {
      // op(A)[bm + r][k0 + kk]
      int r, kk;
      if (TA) { r = e % GB; kk = e / GB; } else { kk = e % GK; r = e / GK; }
      const int gr = bm + r, gk = k0 + kk;
      double v = 0.0;
      if (gr < m && gk < k) v = TA ? A[static_cast<int64_t>(gk) * lda + gr] : A[static_cast<int64_t>(gr) * lda + gk];
      As[kk][r] = v;
    }
    for (int e = threadIdx.x; e < GB * GK; e += 256) {
      // op(B)[k0 + kk][bn + cidx]
      int cidx, kk;
      if (TB) { kk = e % GK; cidx = e / GK; } else { cidx = e % GB; kk = e / GB; }
      const int gc = bn + cidx, gk = k0 + kk;
      double v = 0.0;
      if (gc < n && gk < k) v = TB ? B[static_cast<int64_t>(gc) * ldb + gk] : B[static_cast<int64_t>(gk) * ldb + gc];
      Bs[kk][cidx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
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
  const bool has_c0 = (beta != 0.0) && (C0 != nullptr);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = bm + ty + 16 * i;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = bn + tx + 16 * j;
      if (col >= n) continue;
      const double sv = has_c0 ? C0[static_cast<int64_t>(row) * ldc0 + col] : 0.0;
      C[static_cast<int64_t>(row) * ldc + col] = has_c0 ? fma(alpha, acc[i][j], beta * sv) : alpha * acc[i][j];
    }
  }
}

bool g_attr_set[64] = {};

}  // namespace

int force_generic() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NPW_FORCE_GENERIC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

// Internal launcher shared by syrk/gemm/trsm/potrf/geqrt.
int launch_gemm(double* C, int64_t ldc, const double* C0, int64_t ldc0, const double* A, int64_t lda, int transA,
                const double* B, int64_t ldb, int transB, int64_t m, int64_t n, int64_t k, double alpha,
                double beta, int lower_only, cudaStream_t stream) {
  if (m <= 0 || n <= 0) return NPW_OK;
  if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) {
    set_error("gemm: dimension exceeds int32");
    return NPW_ERR_UNSUPPORTED;
  }
  if (beta == 0.0) C0 = nullptr;
  if (C0 == nullptr) beta = 0.0;
  const bool fast = !force_generic() && !transA && transB && k >= 1 && tma_compatible(A, lda) &&
                    tma_compatible(B, ldb) && (m * n >= 64 * 64);
  if (fast) {
    CUtensorMap tmA, tmB;
    if (make_tmap_f64(&tmA, A, m, k, lda, BM) != 0) return NPW_ERR_CUDA;
    if (make_tmap_f64(&tmB, B, n, k, ldb, BN) != 0) return NPW_ERR_CUDA;
    int dev = 0;
    NPW_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 64 && !g_attr_set[dev]) {
      NPW_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      g_attr_set[dev] = true;
    }
    GemmArgs p;
    p.C = C; p.C0 = C0; p.ldc = ldc; p.ldc0 = ldc0;
    p.m = static_cast<int>(m); p.n = static_cast<int>(n); p.k = static_cast<int>(k);
    p.alpha = alpha; p.beta = beta; p.lower_only = lower_only;
    p.grid_m = static_cast<int>((m + BM - 1) / BM);
    p.grid_n = static_cast<int>((n + BN - 1) / BN);
    p.raster = raster_group();
    p.l2hint = l2_hint_mode();
    p.vec_ok = ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) && (ldc % 2 == 0) &&
               (C0 == nullptr || (((reinterpret_cast<uintptr_t>(C0) & 15u) == 0) && (ldc0 % 2 == 0)));
    gemm_nt_tma_kernel<<<p.grid_m * p.grid_n, NTHREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
    NPW_LAUNCH_CHECK();
    return NPW_OK;
  }
  if (k <= 0) {
    // C = beta*C0 — degenerate, handled by the generic kernel with an empty k loop
  }
  dim3 grid(static_cast<unsigned>((n + GB - 1) / GB), static_cast<unsigned>((m + GB - 1) / GB));
  const int mi = static_cast<int>(m), ni = static_cast<int>(n), ki = static_cast<int>(k);
  if (!transA && !transB)
    gemm_generic_kernel<0, 0><<<grid, 256, 0, stream>>>(C, ldc, C0, ldc0, A, lda, B, ldb, mi, ni, ki, alpha, beta, lower_only);
  else if (!transA && transB)
    gemm_generic_kernel<0, 1><<<grid, 256, 0, stream>>>(C, ldc, C0, ldc0, A, lda, B, ldb, mi, ni, ki, alpha, beta, lower_only);
  else if (transA && !transB)
    gemm_generic_kernel<1, 0><<<grid, 256, 0, stream>>>(C, ldc, C0, ldc0, A, lda, B, ldb, mi, ni, ki, alpha, beta, lower_only);
  else
    gemm_generic_kernel<1, 1><<<grid, 256, 0, stream>>>(C, ldc, C0, ldc0, A, lda, B, ldb, mi, ni, ki, alpha, beta, lower_only);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

}  // namespace npw

extern "C" {

int npw_syrk_f64(double* C_out, int64_t ldc, const double* S, int64_t lds, const double* X, int64_t ldx,
                 const double* Y, int64_t ldy, int64_t m, int64_t n, int64_t k, npw_stream_t stream) {
  if (m == 0 || n == 0) return NPW_OK;  // empty tile: nothing to enqueue
  if (!C_out) return -1;
  if (ldc < n) return -2;
  if (!S) return -3;
  if (lds < n) return -4;
  if (!X && k > 0) return -5;
  if (ldx < k) return -6;
  if (!Y && k > 0) return -7;
  if (ldy < k) return -8;
  if (m < 0) return -9;
  if (n < 0) return -10;
  if (k < 0) return -11;
  return npw::launch_gemm(C_out, ldc, S, lds, X, ldx, 0, Y, ldy, 1, m, n, k, -1.0, 1.0, 0,
                          static_cast<cudaStream_t>(stream));
}

int npw_syrk_lower_f64(double* C_out, int64_t ldc, const double* S, int64_t lds, const double* X, int64_t ldx,
                       const double* Y, int64_t ldy, int64_t m, int64_t n, int64_t k, npw_stream_t stream) {
  if (m == 0 || n == 0) return NPW_OK;
  if (!C_out) return -1;
  if (ldc < n) return -2;
  if (!S) return -3;
  if (lds < n) return -4;
  if (!X && k > 0) return -5;
  if (ldx < k) return -6;
  if (!Y && k > 0) return -7;
  if (ldy < k) return -8;
  if (m < 0) return -9;
  if (n < 0) return -10;
  if (k < 0) return -11;
  return npw::launch_gemm(C_out, ldc, S, lds, X, ldx, 0, Y, ldy, 1, m, n, k, -1.0, 1.0, 1,
                          static_cast<cudaStream_t>(stream));
}

int npw_gemm_f64(double* C, int64_t ldc, const double* C0, int64_t ldc0, const double* A, int64_t lda, int transA,
                 const double* B, int64_t ldb, int transB, int64_t m, int64_t n, int64_t k, double alpha, double beta,
                 npw_stream_t stream) {
  if (m == 0 || n == 0) return NPW_OK;
  if (!C) return -1;
  if (ldc < n) return -2;
  if (beta != 0.0 && C0 && ldc0 < n) return -4;
  if (!A && k > 0) return -5;
  if (lda < (transA ? m : k)) return -6;
  if (!B && k > 0) return -8;
  if (ldb < (transB ? k : n)) return -9;
  if (m < 0) return -11;
  if (n < 0) return -12;
  if (k < 0) return -13;
  return npw::launch_gemm(C, ldc, C0, ldc0, A, lda, transA ? 1 : 0, B, ldb, transB ? 1 : 0, m, n, k, alpha, beta, 0,
                          static_cast<cudaStream_t>(stream));
}

}  // extern "C"
