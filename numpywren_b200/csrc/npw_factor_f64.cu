// npw_factor_f64.cu — tile-local Cholesky (kernels.chol, kernels.py:225-226) and the
// triangular solve kernels.trsm reduces to under the DSL's fixed arguments
// (kernels.py:254-257: dtrsm(1.0, x.T, y, lower=0, side=1) = y * x^{-T}).
//
// Both are blocked so that all O(n^3) work runs through the DMMA GEMM core
// (npw_gemm_f64.cu); only NB x NB = 128 x 128 diagonal blocks are handled by a
// single-CTA kernel that factors the block AND inverts the factor in one pass
// (the inverse turns every panel solve into a GEMM, as in MAGMA's trsm):
//
//   potrf (right-looking, NB = 128):            trsm (recursive on the columns of X):
//     L_jj, inv(L_jj) <- potf2_inv(A_jj)           X1 = trsm(B1, L11)
//     W    <- A_[j+1:, j] * inv(L_jj)^T            B2 -= X1 * L21^T        (GEMM, large k)
//     A_[j+1:, j+1:] -= W * W^T  (lower CTAs)      X2 = trsm(B2, L22)
//                                                  leaf (<=128 cols): X = B * inv(L_jj)^T
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

__device__ __forceinline__ void block_chol_inv(double* sm, const double* Ain, int64_t lda, int nbk, double* Lout, int64_t ldl,
                                               double* inv_out, int32_t* info, int info_base, int do_factor, int write_l) {
  const int t = threadIdx.x;
  const int tx = t >> 4, ty = t & 15;
  const int r0 = TS * ty, c0 = TS * tx;
  const bool lower = tx <= ty;
  double* LT = sm + SM_LT;
  double* XD = sm + SM_XD;
  double* XR = sm + SM_XR;
  double c[TS][TS];
#pragma unroll
  for (int a = 0; a < TS; ++a)
#pragma unroll
    for (int b = 0; b < TS; ++b) {
      const int i = r0 + a, k = c0 + b;
      double v = (i == k) ? 1.0 : 0.0;
      if (lower && k <= i && i < nbk && k < nbk) v = Ain[static_cast<int64_t>(i) * lda + k];
      c[a][b] = v;
    }

  // ------------------------------------------------------------------ Cholesky (or, for trtri, just publish L)
  for (int s = 0; s < NT; ++s) {
    if (tx == s && ty == s) {                        // A: diagonal tile
      double x[TS][TS];
      factor8(c, x, do_factor, info, info_base + TS * s);
#pragma unroll
      for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b) XD[s * TS * TS + a * TS + b] = x[a][b];
    }
    __syncthreads();
    if (tx == s && ty > s) {                         // B: panel below the diagonal tile
      if (do_factor) {
        const double* X = XD + s * TS * TS;          // L21 = A21 X^T : out[a][b] = sum_{q<=b} c[a][q] X[b][q]
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
      }
      double* P = LT + s * TS * PADR + padr(r0);     // publish transposed: P[q][pad(row)]
#pragma unroll
      for (int q = 0; q < TS; ++q)
#pragma unroll
        for (int a = 0; a < TS; a += 2) *reinterpret_cast<double2*>(P + q * PADR + a) = make_double2(c[a][q], c[a + 1][q]);
    }
    __syncthreads();
    if (do_factor && tx > s && lower) {              // C: rank-8 update of everything to the right
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
  // ---- L to global (lower; the strict upper of the block is zeroed)
  if (write_l) {
#pragma unroll
    for (int a = 0; a < TS; ++a)
#pragma unroll
      for (int b = 0; b < TS; ++b) {
        const int i = r0 + a, k = c0 + b;
        if (i < nbk && k < nbk) Lout[static_cast<int64_t>(i) * ldl + k] = (lower && k <= i) ? c[a][b] : 0.0;
      }
  }
  if (!inv_out) return;
  __syncthreads();
  // ------------------------------------------------------------------ inverse
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

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* Ablk, int64_t lda, int nbk, double* inv_out, int32_t* info, int info_base, int do_factor,
                 int write_l) {
  extern __shared__ __align__(16) double sm[];
  block_chol_inv(sm, Ablk, lda, nbk, Ablk, lda, inv_out, info, info_base, do_factor, write_l);
}

bool g_potf2_attr[64] = {};

int launch_potf2_inv(double* Ablk, int64_t lda, int nbk, double* inv_out, int32_t* info, int info_base, int do_factor,
                     int write_l, cudaStream_t st) {
  int dev = 0;
  NPW_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && !g_potf2_attr[dev]) {
    NPW_CUDA_CHECK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    g_potf2_attr[dev] = true;
  }
  potf2_inv_kernel<<<1, POTF2_THREADS, POTF2_SMEM, st>>>(Ablk, lda, nbk, inv_out, info, info_base, do_factor, write_l);
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
  block_chol_inv(sm, L + static_cast<int64_t>(j0) * ldl + j0, ldl, nbk, nullptr, 0,
                 invdiag + static_cast<int64_t>(blk) * NB * NB, nullptr, 0, 0, 0);
}

bool g_trtri_attr[64] = {};

inline int64_t nblocks(int64_t n) { return (n + NB - 1) / NB; }

// X[:, j0:j0+nn] <- solve, in place in Bo (m x n, ld ldbo); see file header.
int trsm_rec(double* Bo, int64_t ldbo, const double* L, int64_t ldl, const double* invdiag, double* wpanel, int64_t m,
             int64_t j0, int64_t nn, cudaStream_t st) {
  if (nn <= NB) {
    const double* inv = invdiag + (j0 / NB) * NB * NB;
    // W = B[:, j0:j0+nn] * inv^T   (inv is NB x NB row-major; only its leading nn x nn part is non-trivial)
    int rc = launch_gemm(wpanel, NB, nullptr, 0, Bo + j0, ldbo, 0, inv, NB, 1, m, nn, nn, 1.0, 0.0, 0, st);
    if (rc) return rc;
    return launch_copy2d(Bo + j0, ldbo, wpanel, NB, m, nn, 0, st);
  }
  int64_t n1 = ((nn / 2 + NB - 1) / NB) * NB;
  if (n1 >= nn) n1 = nn - NB > 0 ? ((nn - 1) / NB) * NB : nn;
  int rc = trsm_rec(Bo, ldbo, L, ldl, invdiag, wpanel, m, j0, n1, st);
  if (rc) return rc;
  const int64_t n2 = nn - n1;
  // B2 -= X1 * L21^T ; L21 = L[j0+n1 : j0+nn, j0 : j0+n1]
  rc = launch_gemm(Bo + j0 + n1, ldbo, Bo + j0 + n1, ldbo, Bo + j0, ldbo, 0, L + (j0 + n1) * ldl + j0, ldl, 1, m, n2, n1,
                   -1.0, 1.0, 0, st);
  if (rc) return rc;
  return trsm_rec(Bo, ldbo, L, ldl, invdiag, wpanel, m, j0 + n1, n2, st);
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
  if (m <= 0 || n <= 0) return 0;
  // [ m x NB panel ] + [ invdiag(n) ]
  return static_cast<size_t>(m) * npw::NB * sizeof(double) + npw_invdiag_bytes(n);
}

int npw_trsm_rlt_f64(double* B_out, int64_t ldbo, const double* L, int64_t ldl, const double* B, int64_t ldb, int64_t m,
                     int64_t n, const double* invdiag, void* work, npw_stream_t stream) {
  if (!B_out) return -1;
  if (ldbo < n) return -2;
  if (!L) return -3;
  if (ldl < n) return -4;
  if (!B) return -5;
  if (ldb < n) return -6;
  if (m < 0) return -7;
  if (n < 0) return -8;
  if (!work) return -10;
  if (m == 0 || n == 0) return NPW_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* wpanel = static_cast<double*>(work);
  if (!invdiag) {
    double* inv = wpanel + m * npw::NB;
    int rc = npw_trtri_diag_f64(inv, L, ldl, n, stream);
    if (rc) return rc;
    invdiag = inv;
  }
  if (B_out != B) {
    int rc = npw::launch_copy2d(B_out, ldbo, B, ldb, m, n, 0, st);
    if (rc) return rc;
  }
  return npw::trsm_rec(B_out, ldbo, L, ldl, invdiag, wpanel, m, 0, n, st);
}

size_t npw_potrf_work_bytes(int64_t n) {
  if (n <= 0) return 0;
  // [ n x NB panel ] + [ invdiag(n) (used when the caller does not ask for it) ]
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
  double* wpanel = static_cast<double*>(work);
  double* inv = invdiag_out ? invdiag_out : wpanel + n * npw::NB;
  int rc;
  if (L_out != A) {
    rc = npw::launch_copy2d(L_out, ldl, A, lda, n, n, 0, st);
    if (rc) return rc;
  }
  constexpr int NB = npw::NB;
  for (int64_t j0 = 0; j0 < n; j0 += NB) {
    const int nbk = static_cast<int>(n - j0 < NB ? n - j0 : NB);
    double* Ajj = L_out + j0 * ldl + j0;
    double* invj = inv + (j0 / NB) * NB * NB;
    rc = npw::launch_potf2_inv(Ajj, ldl, nbk, invj, info_dev, static_cast<int>(j0), 1, 1, st);
    if (rc) return rc;
    const int64_t rest = n - j0 - nbk;
    if (rest > 0) {
      double* P = L_out + (j0 + nbk) * ldl + j0;       // rest x nbk panel below the diagonal block
      double* T = L_out + (j0 + nbk) * ldl + j0 + nbk;  // rest x rest trailing matrix
      rc = npw::launch_gemm(wpanel, NB, nullptr, 0, P, ldl, 0, invj, NB, 1, rest, nbk, nbk, 1.0, 0.0, 0, st);
      if (rc) return rc;
      rc = npw::launch_copy2d(P, ldl, wpanel, NB, rest, nbk, 0, st);
      if (rc) return rc;
      rc = npw::launch_gemm(T, ldl, T, ldl, wpanel, NB, 0, wpanel, NB, 1, rest, rest, nbk, -1.0, 1.0, 1, st);
      if (rc) return rc;
    }
  }
  // np.linalg.cholesky returns zeros above the diagonal
  return npw::launch_fill2d(L_out, ldl, n, n, 2, 0.0, st);
}

}  // extern "C"
