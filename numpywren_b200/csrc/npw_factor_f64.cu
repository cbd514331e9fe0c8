// npw_factor_f64.cu — tile-local Cholesky (kernels.chol, kernels.py:225-226) and the
// triangular solve kernels.trsm reduces to under the DSL's fixed arguments
// (kernels.py:254-257: dtrsm(1.0, x.T, y, lower=0, side=1) = y * x^{-T}).
//
// Both are blocked so that the O(n^3) work runs through the DMMA GEMM core (npw_gemm_f64.cu); the NB x NB = 128 x 128
// diagonal blocks are handled by ONE kernel family (panel_kernel) built from register-tile phases:
//   chol_phase   the block, as 8x8 register tiles of a 256-thread CTA, is factored by rank-8 updates; its panels stay in
//                shared memory (transposed, padded) together with the inverses of the 8x8 diagonal tiles;
//   solve_phase  a chunk of rows is solved against that block by block forward substitution (X = P L_jj^{-T}, in place);
//   inverse_phase (only for npw_trtri_diag_f64 / the optional invdiag output) L_jj^{-1} by block substitution.
//
//   potrf (two-level right-looking):                      trsm (recursive on the columns of X, in place; optionally two
//                                                         row halves on two streams, NPW_B200_TRSM_SPLIT=1):
//     for each 512-column block column                      X1 = trsm(B1, L11)
//       for each 128-column panel j (ONE launch):           B2 -= X1 * L21^T        (GEMM, large k)
//         every CTA factors A_jj redundantly, then          X2 = trsm(B2, L22)
//         solves its 64 rows of A_[j+1:, j] against it      leaf (<= 128 cols): panel_kernel, no factorisation
//         block column's other columns -= W W^T  (k = 128)
//       trailing matrix -= P P^T  (k = 512, lower CTAs)
#include <stdlib.h>

#include <mutex>

#include "npw_common.cuh"

namespace npw {

int launch_gemm(double* C, int64_t ldc, const double* C0, int64_t ldc0, const double* A, int64_t lda, int transA,
                const double* B, int64_t ldb, int transB, int64_t m, int64_t n, int64_t k, double alpha,
                double beta, int lower_only, cudaStream_t stream);
int launch_copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, int trans,
                  cudaStream_t st);
int launch_fill2d(double* A, int64_t lda, int64_t rows, int64_t cols, int mode, double value, cudaStream_t st);

namespace {

constexpr int NB = NPW_DIAG_NB;  // 128
// One CTA of 256 threads factors a 128x128 block AND inverts the factor.  The matrix lives in REGISTERS as 8x8 tiles:
// thread t owns tile (tx, ty) = (t >> 4, t & 15): rows 8ty.., cols 8tx.. (lower tiles: tx <= ty), so the 16 tiles of a
// block column sit in one half-warp and warps retire as the factorisation moves right.
//
// Cholesky, blocked by 8 columns (2 barriers per sub-panel s):
//   A  the diagonal thread (s,s) factors its 8x8 tile in registers and inverts it (X_ss), publishes X_ss;
//   B  the tiles below it (one half-warp) form L21 = A21 X_ss^T and publish the panel, transposed and padded so that
//      the rank-8 update reads it with conflict-free LDS.128;
//   C  every tile to the right applies the rank-8 update from the published panel (512 DFMA per thread, no shared RMW).
// Inverse X = L^{-1} by block forward substitution (1 barrier per block row q): with L'[s][q] = X_ss L[s][q] (each tile
// pre-multiplied once, in parallel) X[s][t] = -sum_{q<s} L'[s][q] X[q][t]; row block q of X is published, every tile
// below accumulates, the tiles of row q+1 are complete after step q.
// do_factor = 0: the block already holds L (batched trtri).  Blocks smaller than NB are padded with the identity.
constexpr int TS = 8;                    // register tile edge
constexpr int NT = NB / TS;              // 16 tiles per dimension
constexpr int PADR = NB + 2 * NT;        // padded row index space: pad(r) = r + 2*(r>>3)
constexpr int POTF2_THREADS = NT * NT;   // 256
// shared memory (doubles): LT[NT][TS][PADR] | XD[NT][TS*TS] | XR[2][TS][PADR]
constexpr int SM_LT = 0;
constexpr int SM_XD = SM_LT + NT * TS * PADR;
constexpr int SM_XR = SM_XD + NT * TS * TS;
constexpr int SM_END = SM_XR + 2 * TS * PADR;
constexpr int POTF2_SMEM = SM_END * 8;

__device__ __forceinline__ int padr(int r) { return r + 2 * (r >> 3); }

// in-register Cholesky of the lower 8x8 tile c (if do_factor) and its inverse x (lower); returns false on a bad pivot
__device__ __forceinline__ void factor8(double (&c)[TS][TS], double (&x)[TS][TS], int do_factor, int32_t* info, int base) {
  double rdv[TS];
#pragma unroll
  for (int j = 0; j < TS; ++j) {
    if (do_factor) {
      const double ajj = c[j][j];
      double d, rd;
      if (!(ajj > 0.0)) {
        atomicCAS(info, 0, base + j + 1);
        d = rd = nan("");
      } else {
        // this chain runs on ONE thread and sits on the critical path of the whole tile factorisation: no fp64
        // division (a ~30-instruction dependent sequence), only rsqrt plus two Newton corrections
        rd = rsqrt(ajj);
        d = ajj * rd;
        d = fma(0.5 * rd, fma(-d, d, ajj), d);       // sqrt(ajj) to < 1 ulp
        rd = fma(rd, fma(-d, rd, 1.0), rd);          // 1/d to < 1 ulp
      }
      c[j][j] = d;
      rdv[j] = rd;
#pragma unroll
      for (int a = j + 1; a < TS; ++a) c[a][j] *= rd;
#pragma unroll
      for (int b = j + 1; b < TS; ++b)
#pragma unroll
        for (int a = b; a < TS; ++a) c[a][b] = fma(-c[a][j], c[b][j], c[a][b]);
    } else {
      rdv[j] = 1.0 / c[j][j];
    }
  }
#pragma unroll
  for (int j = 0; j < TS; ++j) {
#pragma unroll
    for (int i = 0; i < TS; ++i) x[i][j] = 0.0;
    x[j][j] = rdv[j];
#pragma unroll
    for (int i = j + 1; i < TS; ++i) {
      double s = 0.0;
#pragma unroll
      for (int q = j; q < i; ++q) s = fma(c[i][q], x[q][j], s);
      x[i][j] = -rdv[i] * s;
    }
  }
}

// ---- phase 0: this thread's 8x8 tile of the (lower triangle of the) nbk x nbk block; identity padding up to NB
__device__ __forceinline__ void load_block_tile(double (&c)[TS][TS], const double* Ain, int64_t lda, int nbk) {
  const int tx = threadIdx.x >> 4, ty = threadIdx.x & 15;
  const int r0 = TS * ty, c0 = TS * tx;
  const bool lower = tx <= ty;
#pragma unroll
  for (int a = 0; a < TS; ++a)
#pragma unroll
    for (int b = 0; b < TS; ++b) {
      const int i = r0 + a, k = c0 + b;
      double v = (i == k) ? 1.0 : 0.0;
      if (lower && k <= i && i < nbk && k < nbk) v = Ain[static_cast<int64_t>(i) * lda + k];
      c[a][b] = v;
    }
}

// ---- phase 1: Cholesky of the block held in registers (do_factor) or, for a block that already holds L, just the
// by-products.  Afterwards: c = this thread's tile of L; LT[s][q][pad(row)] = L[row][8s+q] for rows below diagonal
// tile s (the panels, transposed and padded for conflict-free LDS.128); XD[s] = inv(L_ss) of the 8x8 diagonal tiles.
__device__ __forceinline__ void chol_phase(double* sm, double (&c)[TS][TS], int do_factor, int32_t* info, int info_base) {
  const int t = threadIdx.x;
  const int tx = t >> 4, ty = t & 15;
  const int r0 = TS * ty, c0 = TS * tx;
  const bool lower = tx <= ty;
  double* LT = sm + SM_LT;
  double* XD = sm + SM_XD;
  if (!do_factor) {
    // nothing depends on anything: all diagonal tiles are inverted at once, all panels published at once
    if (tx == ty) {
      double x[TS][TS];
      factor8(c, x, 0, info, 0);
#pragma unroll
      for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b) XD[tx * TS * TS + a * TS + b] = x[a][b];
    } else if (lower) {
      double* P = LT + tx * TS * PADR + padr(r0);
#pragma unroll
      for (int q = 0; q < TS; ++q)
#pragma unroll
        for (int a = 0; a < TS; a += 2) *reinterpret_cast<double2*>(P + q * PADR + a) = make_double2(c[a][q], c[a + 1][q]);
    }
    __syncthreads();
    return;
  }
  for (int s = 0; s < NT; ++s) {
    if (tx == s && ty == s) {                        // A: diagonal tile
      double x[TS][TS];
      factor8(c, x, 1, info, info_base + TS * s);
#pragma unroll
      for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b) XD[s * TS * TS + a * TS + b] = x[a][b];
    }
    __syncthreads();
    if (tx == s && ty > s) {                         // B: panel below the diagonal tile
      const double* X = XD + s * TS * TS;            // L21 = A21 X^T : out[a][b] = sum_{q<=b} c[a][q] X[b][q]
#pragma unroll
      for (int b = TS - 1; b >= 0; --b) {
        double xb[TS];
#pragma unroll
        for (int q = 0; q <= b; ++q) xb[q] = X[b * TS + q];
#pragma unroll
        for (int a = 0; a < TS; ++a) {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q <= b; ++q) acc = fma(c[a][q], xb[q], acc);
          c[a][b] = acc;
        }
      }
      double* P = LT + s * TS * PADR + padr(r0);     // publish transposed: P[q][pad(row)]
#pragma unroll
      for (int q = 0; q < TS; ++q)
#pragma unroll
        for (int a = 0; a < TS; a += 2) *reinterpret_cast<double2*>(P + q * PADR + a) = make_double2(c[a][q], c[a + 1][q]);
    }
    __syncthreads();
    if (tx > s && lower) {                           // C: rank-8 update of everything to the right
      const double* Pi = LT + s * TS * PADR + padr(r0);
      const double* Pk = LT + s * TS * PADR + padr(c0);
#pragma unroll
      for (int q = 0; q < TS; ++q) {
        double li[TS], lk[TS];
#pragma unroll
        for (int a = 0; a < TS; a += 2) {
          const double2 v = *reinterpret_cast<const double2*>(Pi + q * PADR + a);
          li[a] = v.x; li[a + 1] = v.y;
          const double2 w = *reinterpret_cast<const double2*>(Pk + q * PADR + a);
          lk[a] = w.x; lk[a + 1] = w.y;
        }
#pragma unroll
        for (int a = 0; a < TS; ++a)
#pragma unroll
          for (int b = 0; b < TS; ++b) c[a][b] = fma(-li[a], lk[b], c[a][b]);
      }
    }
  }
  __syncthreads();
}

// ---- phase 2 (optional): X = L^{-1} by block forward substitution (1 barrier per block row q): with
// L'[s][q] = X_ss L[s][q] (each tile pre-multiplied once, in parallel) X[s][t] = -sum_{q<s} L'[s][q] X[q][t]; row
// block q of X is published, every tile below accumulates, the tiles of row q+1 are complete after step q.
__device__ __forceinline__ void inverse_phase(double* sm, double (&c)[TS][TS], double* inv_out) {
  const int t = threadIdx.x;
  const int tx = t >> 4, ty = t & 15;
  const int r0 = TS * ty, c0 = TS * tx;
  const bool lower = tx <= ty;
  double* LT = sm + SM_LT;
  double* XD = sm + SM_XD;
  double* XR = sm + SM_XR;
  // L'[s][q] = X_ss L[s][q] for the off-diagonal tiles, re-published into LT; diagonal tiles become X_ss
  if (lower && tx < ty) {
    const double* X = XD + ty * TS * TS;             // X_ss of this tile's block row
#pragma unroll
    for (int a = TS - 1; a >= 0; --a) {              // new[a][b] = sum_{r<=a} X[a][r] c[r][b], rows bottom-up in place
      double xa[TS];
#pragma unroll
      for (int r = 0; r <= a; ++r) xa[r] = X[a * TS + r];
#pragma unroll
      for (int b = 0; b < TS; ++b) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r <= a; ++r) acc = fma(xa[r], c[r][b], acc);
        c[a][b] = acc;
      }
    }
    double* P = LT + tx * TS * PADR + padr(r0);
#pragma unroll
    for (int q = 0; q < TS; ++q)
#pragma unroll
      for (int a = 0; a < TS; a += 2) *reinterpret_cast<double2*>(P + q * PADR + a) = make_double2(c[a][q], c[a + 1][q]);
  }
  // the tile registers now become X: diagonal tiles start as X_ss, the others accumulate from zero
#pragma unroll
  for (int a = 0; a < TS; ++a)
#pragma unroll
    for (int b = 0; b < TS; ++b) c[a][b] = (tx == ty) ? XD[ty * TS * TS + a * TS + b] : 0.0;
  __syncthreads();
  for (int q = 0; q < NT; ++q) {
    double* XRq = XR + (q & 1) * TS * PADR;
    if (ty == q && lower) {                          // publish row block q of X (final): XRq[row a][pad(col)]
      // off-diagonal tiles hold +sum L' X: the sign of the substitution is applied here
      const double sg = (tx == ty) ? 1.0 : -1.0;
#pragma unroll
      for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; b += 2) {
          c[a][b] *= sg; c[a][b + 1] *= sg;
          *reinterpret_cast<double2*>(XRq + a * PADR + padr(c0) + b) = make_double2(c[a][b], c[a][b + 1]);
        }
    }
    __syncthreads();
    if (ty > q && tx <= q) {                         // acc[s][t] += L'[s][q] X[q][t]
      const double* Pi = LT + q * TS * PADR + padr(r0);
      const double* Xk = XRq + padr(c0);
#pragma unroll
      for (int r = 0; r < TS; ++r) {
        double li[TS], xk[TS];
#pragma unroll
        for (int a = 0; a < TS; a += 2) {
          const double2 v = *reinterpret_cast<const double2*>(Pi + r * PADR + a);
          li[a] = v.x; li[a + 1] = v.y;
          const double2 w = *reinterpret_cast<const double2*>(Xk + r * PADR + a);
          xk[a] = w.x; xk[a + 1] = w.y;
        }
#pragma unroll
        for (int a = 0; a < TS; ++a)
#pragma unroll
          for (int b = 0; b < TS; ++b) c[a][b] = fma(li[a], xk[b], c[a][b]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < TS; ++a)
#pragma unroll
    for (int b = 0; b < TS; ++b) {
      const int i = r0 + a, k = c0 + b;
      inv_out[i * NB + k] = (lower && k <= i) ? c[a][b] : 0.0;
    }
}

// ---- phase 3: triangular solve of a row chunk against the block whose by-products chol_phase left in shared memory:
// X = P L^{-T} in place, P = `rows` x nbk (rows <= 16 * TR).  Thread (tx, ty) owns the TR x 8 tile (rows TR*ty.., cols
// 8*tx..).  Block forward substitution over the 16 column tiles, one barrier per step s:
//   B  tiles of column s are final up to the diagonal tile's inverse: X_s = p X_ss^T; published transposed (XP[q][pad(row)])
//   C  every tile to the right subtracts X_s L[8tx.., 8s..]^T — a rank-8 update whose L operand is the panel LT[s]
//      the Cholesky phase published (so the factor needs no second pass through memory).
template <int TR>
__device__ __forceinline__ void solve_phase(double* sm, double* P, int64_t ldp, int rows, int nbk) {
  const int t = threadIdx.x;
  const int tx = t >> 4, ty = t & 15;
  const int r0 = TR * ty, c0 = TS * tx;
  const double* LT = sm + SM_LT;
  const double* XD = sm + SM_XD;
  double* XP = sm + SM_XR;
  double p[TR][TS];
  const bool vec = ((reinterpret_cast<uintptr_t>(P) | (static_cast<uint64_t>(ldp) * 8u)) & 15u) == 0 && c0 + TS <= nbk;
#pragma unroll
  for (int a = 0; a < TR; ++a) {
    const int i = r0 + a;
    if (i < rows && vec) {
      const double2* src = reinterpret_cast<const double2*>(P + static_cast<int64_t>(i) * ldp + c0);
#pragma unroll
      for (int b = 0; b < TS; b += 2) {
        const double2 v = src[b >> 1];
        p[a][b] = v.x; p[a][b + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int b = 0; b < TS; ++b) p[a][b] = (i < rows && c0 + b < nbk) ? P[static_cast<int64_t>(i) * ldp + c0 + b] : 0.0;
    }
  }
  for (int s = 0; s < NT; ++s) {
    double* XPs = XP + (s & 1) * TS * PADR;
    if (tx == s) {                                   // B
      const double* X = XD + s * TS * TS;            // out[a][b] = sum_{q<=b} p[a][q] X[b][q]
#pragma unroll
      for (int b = TS - 1; b >= 0; --b) {
        double xb[TS];
#pragma unroll
        for (int q = 0; q <= b; ++q) xb[q] = X[b * TS + q];
#pragma unroll
        for (int a = 0; a < TR; ++a) {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q <= b; ++q) acc = fma(p[a][q], xb[q], acc);
          p[a][b] = acc;
        }
      }
      double* Q = XPs + padr(r0);
#pragma unroll
      for (int q = 0; q < TS; ++q)
#pragma unroll
        for (int a = 0; a < TR; a += 2) *reinterpret_cast<double2*>(Q + q * PADR + a) = make_double2(p[a][q], p[a + 1][q]);
    }
    __syncthreads();
    if (tx > s) {                                    // C
      const double* Pi = XPs + padr(r0);
      const double* Pk = LT + s * TS * PADR + padr(c0);
#pragma unroll
      for (int q = 0; q < TS; ++q) {
        double li[TR], lk[TS];
#pragma unroll
        for (int a = 0; a < TR; a += 2) {
          const double2 v = *reinterpret_cast<const double2*>(Pi + q * PADR + a);
          li[a] = v.x; li[a + 1] = v.y;
        }
#pragma unroll
        for (int b = 0; b < TS; b += 2) {
          const double2 w = *reinterpret_cast<const double2*>(Pk + q * PADR + b);
          lk[b] = w.x; lk[b + 1] = w.y;
        }
#pragma unroll
        for (int a = 0; a < TR; ++a)
#pragma unroll
          for (int b = 0; b < TS; ++b) p[a][b] = fma(-li[a], lk[b], p[a][b]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < TR; ++a) {
    const int i = r0 + a;
    if (i < rows && vec) {
      double2* dst = reinterpret_cast<double2*>(P + static_cast<int64_t>(i) * ldp + c0);
#pragma unroll
      for (int b = 0; b < TS; b += 2) dst[b >> 1] = make_double2(p[a][b], p[a][b + 1]);
    } else if (i < rows) {
#pragma unroll
      for (int b = 0; b < TS; ++b)
        if (c0 + b < nbk) P[static_cast<int64_t>(i) * ldp + c0 + b] = p[a][b];
    }
  }
}

__device__ __forceinline__ void store_block_tile(const double (&c)[TS][TS], double* Lout, int64_t ldl, int nbk) {
  const int tx = threadIdx.x >> 4, ty = threadIdx.x & 15;
  const int r0 = TS * ty, c0 = TS * tx;
  const bool lower = tx <= ty;
#pragma unroll
  for (int a = 0; a < TS; ++a)
#pragma unroll
    for (int b = 0; b < TS; ++b) {
      const int i = r0 + a, k = c0 + b;
      if (i < nbk && k < nbk) Lout[static_cast<int64_t>(i) * ldl + k] = (lower && k <= i) ? c[a][b] : 0.0;
    }
}

// One 128-column panel step of the tile Cholesky in ONE launch (it used to be three: potf2+inverse, panel GEMM, copy):
// every CTA factors the diagonal block itself — redundantly, in registers; the other SMs would idle anyway and it saves a
// grid-wide dependency — and then solves its own chunk of the panel rows below against the factor it just formed.
// CTA 0 has no rows: it stores the factored block to `Ldiag` (scratch: the diagonal block of A itself is still being
// read by the other CTAs; npw_potrf_l_f64's last launch moves it into place).
// do_factor = 0 (kernels.trsm leaf): the block already holds L, every CTA solves rows (no CTA 0 role).
template <int TR>
__global__ void __launch_bounds__(POTF2_THREADS, 1)
panel_kernel(const double* Ajj, int64_t lda, int nbk, double* Ldiag, double* P, int64_t ldp, int rest, int32_t* info,
             int info_base, int do_factor) {
  extern __shared__ __align__(16) double sm[];
  double c[TS][TS];
  load_block_tile(c, Ajj, lda, nbk);
  chol_phase(sm, c, do_factor, info, info_base);
  int chunk = blockIdx.x;
  if (do_factor) {
    if (chunk == 0) {
      store_block_tile(c, Ldiag, NB, NB);            // full NB x NB scratch block (identity padded)
      return;
    }
    --chunk;
  }
  const int row0 = chunk * (NT * TR);
  if (row0 >= rest) return;
  solve_phase<TR>(sm, P + static_cast<int64_t>(row0) * ldp, ldp, min(NT * TR, rest - row0), nbk);
}

// last launch of npw_potrf_l_f64: diagonal blocks from scratch into place, strict upper triangle zeroed
// (np.linalg.cholesky returns zeros there); tiles below the diagonal are left alone.
__global__ void __launch_bounds__(256)
potrf_finalize_kernel(double* L, int64_t ldl, int n, const double* Ldiag) {
  const int bj = blockIdx.y, bk = blockIdx.x;
  if (bk < bj) return;
  const int i0 = bj * NB, k0 = bk * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int a = e / NB, b = e % NB;
    const int i = i0 + a, k = k0 + b;
    if (i >= n || k >= n) continue;
    L[static_cast<int64_t>(i) * ldl + k] = (bk == bj) ? Ldiag[static_cast<int64_t>(bj) * NB * NB + e] : 0.0;
  }
}

bool g_potf2_attr[64] = {};

int set_factor_attrs() {
  int dev = 0;
  NPW_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && !g_potf2_attr[dev]) {
    NPW_CUDA_CHECK(cudaFuncSetAttribute(panel_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    NPW_CUDA_CHECK(cudaFuncSetAttribute(panel_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    g_potf2_attr[dev] = true;
  }
  return NPW_OK;
}

// factor (do_factor) the nbk x nbk block at Ajj and solve the `rest` rows of P (rest x nbk) against it, in place
int launch_panel(const double* Ajj, int64_t lda, int nbk, double* Ldiag, double* P, int64_t ldp, int64_t rest, int32_t* info,
                 int info_base, int do_factor, cudaStream_t st) {
  int rc = set_factor_attrs();
  if (rc) return rc;
  // 64-row chunks (4-row register tiles) spread a tile's panel over ~64 SMs; tall panels use 128-row chunks
  static const int64_t small_rows = [] { const char* e = getenv("NPW_B200_PANEL_SMALL_ROWS"); return e ? atoll(e) : 64ll * 192; }();
  const bool small = rest <= small_rows;
  const int per = small ? NT * 4 : NT * 8;
  const unsigned grid = static_cast<unsigned>((rest + per - 1) / per + (do_factor ? 1 : 0));
  if (grid == 0) return NPW_OK;
  if (small)
    panel_kernel<4><<<grid, POTF2_THREADS, POTF2_SMEM, st>>>(Ajj, lda, nbk, Ldiag, P, ldp, static_cast<int>(rest), info, info_base, do_factor);
  else
    panel_kernel<8><<<grid, POTF2_THREADS, POTF2_SMEM, st>>>(Ajj, lda, nbk, Ldiag, P, ldp, static_cast<int>(rest), info, info_base, do_factor);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

// batched trtri of the diagonal blocks of an existing L: one CTA per block
__global__ void __launch_bounds__(POTF2_THREADS, 1)
trtri_diag_kernel(const double* L, int64_t ldl, int n, double* invdiag) {
  extern __shared__ __align__(16) double sm[];
  const int blk = blockIdx.x;
  const int j0 = blk * NB;
  const int nbk = min(NB, n - j0);
  double c[TS][TS];
  load_block_tile(c, L + static_cast<int64_t>(j0) * ldl + j0, ldl, nbk);
  chol_phase(sm, c, 0, nullptr, 0);
  inverse_phase(sm, c, invdiag + static_cast<int64_t>(blk) * NB * NB);
}

bool g_trtri_attr[64] = {};

inline int64_t nblocks(int64_t n) { return (n + NB - 1) / NB; }

// X[:, j0:j0+nn] <- solve, in place in Bo (m x n, ld ldbo); see file header.
int trsm_rec(double* Bo, int64_t ldbo, const double* L, int64_t ldl, int64_t m, int64_t j0, int64_t nn, cudaStream_t st) {
  if (nn <= NB) {
    // leaf: X = B[:, j0:j0+nn] * L_jj^{-T} in place, by substitution against the diagonal block itself
    return launch_panel(L + j0 * ldl + j0, ldl, static_cast<int>(nn), nullptr, Bo + j0, ldbo, m, nullptr, 0, 0, st);
  }
  int64_t n1 = ((nn / 2 + NB - 1) / NB) * NB;
  if (n1 >= nn) n1 = nn - NB > 0 ? ((nn - 1) / NB) * NB : nn;
  int rc = trsm_rec(Bo, ldbo, L, ldl, m, j0, n1, st);
  if (rc) return rc;
  const int64_t n2 = nn - n1;
  // B2 -= X1 * L21^T ; L21 = L[j0+n1 : j0+nn, j0 : j0+n1]
  rc = launch_gemm(Bo + j0 + n1, ldbo, Bo + j0 + n1, ldbo, Bo + j0, ldbo, 0, L + (j0 + n1) * ldl + j0, ldl, 1, m, n2, n1,
                   -1.0, 1.0, 0, st);
  if (rc) return rc;
  return trsm_rec(Bo, ldbo, L, ldl, m, j0 + n1, n2, st);
}

// ---- fork / join helper: the rows of B are independent problems, so a large solve is cut into two row halves that run
// on two streams — the latency-bound 128-column leaves of one half overlap the GEMM updates of the other (alone on a
// GPU, which is how the panel solves of the multi-GPU Cholesky mostly run, the one-stream solve left ~40 % of the SMs'
// time unused).  A few side streams per device, used round-robin; the mutex only covers the enqueue sequence.
struct ForkJoin {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  std::mutex mu;
};
constexpr int FJ_SLOTS = 2, FJ_LEVELS = 8;
ForkJoin g_fj[16][FJ_LEVELS][FJ_SLOTS];
std::atomic<unsigned> g_fj_next{0};

// the side stream gets the priority of the calling stream: the second half must neither overtake nor lag behind the
// work the engine ordered by priority
ForkJoin* acquire_fork_join(cudaStream_t st) {
  int dev = 0, lo = 0, hi = 0, prio = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 16) return nullptr;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess || cudaStreamGetPriority(st, &prio) != cudaSuccess) return nullptr;
  int level = prio - hi;                                  // 0 = greatest priority
  if (level < 0) level = 0;
  if (level >= FJ_LEVELS) level = FJ_LEVELS - 1;
  ForkJoin* fj = &g_fj[dev][level][g_fj_next.fetch_add(1, std::memory_order_relaxed) % FJ_SLOTS];
  fj->mu.lock();
  if (!fj->side) {
    if (cudaStreamCreateWithPriority(&fj->side, cudaStreamNonBlocking, prio) != cudaSuccess ||
        cudaEventCreateWithFlags(&fj->fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fj->join, cudaEventDisableTiming) != cudaSuccess) {
      fj->side = nullptr;
      fj->mu.unlock();
      return nullptr;
    }
  }
  return fj;
}

}  // namespace
}  // namespace npw

extern "C" {

size_t npw_invdiag_bytes(int64_t n) {
  if (n <= 0) return 0;
  return static_cast<size_t>(npw::nblocks(n)) * npw::NB * npw::NB * sizeof(double);
}

int npw_trtri_diag_f64(double* invdiag, const double* L, int64_t ldl, int64_t n, npw_stream_t stream) {
  if (!invdiag) return -1;
  if (!L) return -2;
  if (ldl < n) return -3;
  if (n < 0 || n > INT32_MAX) return -4;
  if (n == 0) return NPW_OK;
  int dev = 0;
  NPW_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && !npw::g_trtri_attr[dev]) {
    NPW_CUDA_CHECK(cudaFuncSetAttribute(npw::trtri_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, npw::POTF2_SMEM));
    npw::g_trtri_attr[dev] = true;
  }
  npw::trtri_diag_kernel<<<static_cast<unsigned>(npw::nblocks(n)), npw::POTF2_THREADS, npw::POTF2_SMEM,
                           static_cast<cudaStream_t>(stream)>>>(L, ldl, static_cast<int>(n), invdiag);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

size_t npw_trsm_work_bytes(int64_t m, int64_t n) {
  // the solve works in place in B_out; the entry keeps its `work` argument for ABI stability
  if (m <= 0 || n <= 0) return 0;
  return 16;
}

int npw_trsm_rlt_f64(double* B_out, int64_t ldbo, const double* L, int64_t ldl, const double* B, int64_t ldb, int64_t m,
                     int64_t n, const double* invdiag, void* work, npw_stream_t stream) {
  if (!B_out) return -1;
  if (ldbo < n) return -2;
  if (!L) return -3;
  if (ldl < n) return -4;
  if (!B) return -5;
  if (ldb < n) return -6;
  if (m < 0 || m > INT32_MAX) return -7;
  if (n < 0 || n > INT32_MAX) return -8;
  (void)invdiag; (void)work;   // leaves substitute against the diagonal blocks of L directly: no inverses, no scratch
  if (m == 0 || n == 0) return NPW_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B_out != B) {
    int rc = npw::launch_copy2d(B_out, ldbo, B, ldb, m, n, 0, st);
    if (rc) return rc;
  }
  // two row halves on two streams when both halves still fill the GPU's GEMM grid (see ForkJoin)
  // OPT-IN (NPW_B200_TRSM_SPLIT=1): measured 3.64 -> 3.49 ms per 4096^2 solve and +0.1 % on the one-GPU Cholesky, but the
  // extra side streams were never run inside the 8-GPU engine, whose exchange parks wait_signal kernels on per-peer
  // streams and relies on every stream having its own hardware queue (CUDA_DEVICE_MAX_CONNECTIONS=32): not worth a
  // possible false dependency there.
  static const bool split_ok = [] { const char* e = getenv("NPW_B200_TRSM_SPLIT"); return e && atoi(e) != 0; }();
  if (split_ok && m >= 2048 && n >= 1024) {
    if (npw::ForkJoin* fj = npw::acquire_fork_join(st)) {
      const int64_t h = ((m / 2 + npw::NB - 1) / npw::NB) * npw::NB;
      int rc = NPW_OK, rc2 = NPW_OK;
      cudaError_t e = cudaEventRecord(fj->fork, st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(fj->side, fj->fork, 0);
      if (e == cudaSuccess) {
        rc = npw::trsm_rec(B_out, ldbo, L, ldl, h, 0, n, st);
        rc2 = npw::trsm_rec(B_out + h * ldbo, ldbo, L, ldl, m - h, 0, n, fj->side);
        e = cudaEventRecord(fj->join, fj->side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, fj->join, 0);
      }
      fj->mu.unlock();
      if (e != cudaSuccess) {
        npw::set_error("trsm fork/join: %s", cudaGetErrorString(e));
        return NPW_ERR_CUDA;
      }
      return rc ? rc : rc2;
    }
  }
  return npw::trsm_rec(B_out, ldbo, L, ldl, m, 0, n, st);
}

size_t npw_potrf_work_bytes(int64_t n) {
  if (n <= 0) return 0;
  // [ factored diagonal blocks, ceil(n/NB) x NB x NB ] (+ slack kept from the previous layout)
  return static_cast<size_t>(n) * npw::NB * sizeof(double) + npw_invdiag_bytes(n);
}

int npw_potrf_l_f64(double* L_out, int64_t ldl, const double* A, int64_t lda, int64_t n, int32_t* info_dev,
                    double* invdiag_out, void* work, npw_stream_t stream) {
  if (!L_out) return -1;
  if (ldl < n) return -2;
  if (!A) return -3;
  if (lda < n) return -4;
  if (n < 0 || n > INT32_MAX) return -5;
  if (!info_dev) return -6;
  if (!work) return -8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NPW_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int32_t), st));
  if (n == 0) return NPW_OK;
  double* Ldiag = static_cast<double*>(work);
  int rc;
  if (L_out != A) {
    rc = npw::launch_copy2d(L_out, ldl, A, lda, n, n, 0, st);
    if (rc) return rc;
  }
  // Two-level right-looking factorisation.  Outer block columns of W = 512: inside one, each 128-column panel step is
  // ONE launch (diagonal block factored redundantly by every CTA, rows below solved against it) plus a narrow update of
  // the block column's remaining columns (k = 128); the trailing matrix is updated once per outer block column with
  // k = 512, where the DMMA GEMM runs near its peak.
  constexpr int NB = npw::NB;
  constexpr int64_t W = 4 * NB;
  for (int64_t J0 = 0; J0 < n; J0 += W) {
    const int64_t J1 = J0 + W < n ? J0 + W : n;
    for (int64_t j0 = J0; j0 < J1; j0 += NB) {
      const int nbk = static_cast<int>(J1 - j0 < NB ? J1 - j0 : NB);
      const int64_t rest = n - j0 - nbk;
      double* Ajj = L_out + j0 * ldl + j0;
      double* P = L_out + (j0 + nbk) * ldl + j0;       // rest x nbk panel below the diagonal block
      rc = npw::launch_panel(Ajj, ldl, nbk, Ldiag + (j0 / NB) * NB * NB, P, ldl, rest, info_dev, static_cast<int>(j0), 1, st);
      if (rc) return rc;
      const int64_t nc = J1 - (j0 + nbk);              // columns of this outer block column still to be updated
      if (rest > 0 && nc > 0) {
        double* T = L_out + (j0 + nbk) * ldl + j0 + nbk;
        rc = npw::launch_gemm(T, ldl, T, ldl, P, ldl, 0, P, ldl, 1, rest, nc, nbk, -1.0, 1.0, 1, st);
        if (rc) return rc;
      }
    }
    const int64_t rest2 = n - J1;
    if (rest2 > 0) {
      double* Pout = L_out + J1 * ldl + J0;            // rest2 x (J1 - J0) block of finished panels
      double* T2 = L_out + J1 * ldl + J1;
      rc = npw::launch_gemm(T2, ldl, T2, ldl, Pout, ldl, 0, Pout, ldl, 1, rest2, rest2, J1 - J0, -1.0, 1.0, 1, st);
      if (rc) return rc;
    }
  }
  // diagonal blocks into place; np.linalg.cholesky returns zeros above the diagonal
  {
    const unsigned nblk = static_cast<unsigned>(npw::nblocks(n));
    npw::potrf_finalize_kernel<<<dim3(nblk, nblk), 256, 0, st>>>(L_out, ldl, static_cast<int>(n), Ldiag);
    NPW_LAUNCH_CHECK();
  }
  if (invdiag_out) return npw_trtri_diag_f64(invdiag_out, L_out, ldl, n, stream);
  return NPW_OK;
}

}  // extern "C"
