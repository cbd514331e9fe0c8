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
constexpr int LDSM = NB + 1;     // padded row stride (doubles) of the shared block
constexpr int POTF2_THREADS = 1024;
constexpr int POTF2_SMEM = NB * LDSM * 8;

// One CTA.  Shared block W[NB][NB+1]:
//   lower triangle incl. diagonal  : A -> L            (W[i][p], p <= i)
//   strictly upper, shifted by one : X^T, X = inv(L)   (X[i][t] at W[t][i+1], t <= i)
// do_factor = 1: Cholesky-factor the block first (potf2); 0: the block already holds L (trtri only).
// Blocks smaller than NB are padded with the identity.
__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* Ablk, int64_t lda, int nbk, double* inv_out, int32_t* info, int info_base, int do_factor,
                 int write_l) {
  extern __shared__ double W[];
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;  // 32 x 32
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;

  for (int e = tid; e < NB * NB; e += POTF2_THREADS) {
    const int i = e / NB, p = e - i * NB;
    if (p <= i) {
      double v = (i == p) ? 1.0 : 0.0;
      if (i < nbk && p < nbk) v = Ablk[static_cast<int64_t>(i) * lda + p];
      W[i * LDSM + p] = v;
    } else {
      // X^T region (t = i, column index p = i'+1): X starts as the identity; X[i'][t] with i' = p-1 >= t
      W[i * LDSM + p] = (p - 1 == i) ? 1.0 : 0.0;
    }
  }
  // last shifted column (p = NB) holds X[NB-1][t]
  for (int t = tid; t < NB; t += POTF2_THREADS) W[t * LDSM + NB] = (t == NB - 1) ? 1.0 : 0.0;
  __syncthreads();

  for (int j = 0; j < NB; ++j) {
    // ---- column j of L
    if (do_factor) {
      const double ajj = W[j * LDSM + j];
      __syncthreads();  // everyone has read a_jj before it is overwritten
      double d;
      if (!(ajj > 0.0)) {
        if (tid == 0 && s_bad == 0) { s_bad = 1; atomicCAS(info, 0, info_base + j + 1); }
        d = nan("");
      } else {
        d = sqrt(ajj);
      }
      if (tid == 0) W[j * LDSM + j] = d;
      const double rd = 1.0 / d;
      for (int i = j + 1 + tid; i < NB; i += POTF2_THREADS) W[i * LDSM + j] = W[i * LDSM + j] * rd;
      __syncthreads();
    }
    const double ljj = W[j * LDSM + j];
    // ---- row j of X is final up to the division by L_jj:  X[j][t] /= L_jj, t <= j  (stored W[t][j+1])
    for (int t = tid; t <= j; t += POTF2_THREADS) W[t * LDSM + j + 1] = W[t * LDSM + j + 1] / ljj;
    __syncthreads();
    // ---- trailing updates (both are rank-1):
    //   A[i][k] -= L[i][j] * L[k][j]      j < k <= i          (only when factoring)
    //   X[i][t] -= L[i][j] * X[j][t]      i > j, t <= j
    if (do_factor) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = ty + 32 * a;
        if (i <= j) continue;
        const double lij = W[i * LDSM + j];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = tx + 32 * b;
          if (k > j && k <= i) W[i * LDSM + k] -= lij * W[k * LDSM + j];
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int t = ty + 32 * b;
      if (t > j) continue;
      const double xjt = W[t * LDSM + j + 1];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = tx + 32 * a;
        if (i > j) W[t * LDSM + i + 1] -= W[i * LDSM + j] * xjt;
      }
    }
    __syncthreads();
  }

  // ---- write back: L (valid part, lower; strict upper of the block zeroed) and inv (full NB x NB, row-major)
  for (int e = tid; e < NB * NB; e += POTF2_THREADS) {
    const int i = e / NB, p = e - i * NB;
    if (write_l && i < nbk && p < nbk) Ablk[static_cast<int64_t>(i) * lda + p] = (p <= i) ? W[i * LDSM + p] : 0.0;
    if (inv_out) inv_out[e] = (p <= i) ? W[p * LDSM + i + 1] : 0.0;
  }
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
  // thin wrapper: same algorithm as potf2_inv_kernel with do_factor = 0, one block per CTA
  extern __shared__ double W[];
  const int blk = blockIdx.x;
  const int j0 = blk * NB;
  const int nbk = min(NB, n - j0);
  const double* Ablk = L + static_cast<int64_t>(j0) * ldl + j0;
  double* inv_out = invdiag + static_cast<int64_t>(blk) * NB * NB;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  for (int e = tid; e < NB * NB; e += POTF2_THREADS) {
    const int i = e / NB, p = e - i * NB;
    if (p <= i) {
      double v = (i == p) ? 1.0 : 0.0;
      if (i < nbk && p < nbk) v = Ablk[static_cast<int64_t>(i) * ldl + p];
      W[i * LDSM + p] = v;
    } else {
      W[i * LDSM + p] = (p - 1 == i) ? 1.0 : 0.0;
    }
  }
  for (int t = tid; t < NB; t += POTF2_THREADS) W[t * LDSM + NB] = (t == NB - 1) ? 1.0 : 0.0;
  __syncthreads();
  for (int j = 0; j < NB; ++j) {
    const double ljj = W[j * LDSM + j];
    for (int t = tid; t <= j; t += POTF2_THREADS) W[t * LDSM + j + 1] = W[t * LDSM + j + 1] / ljj;
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int t = ty + 32 * b;
      if (t > j) continue;
      const double xjt = W[t * LDSM + j + 1];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = tx + 32 * a;
        if (i > j) W[t * LDSM + i + 1] -= W[i * LDSM + j] * xjt;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += POTF2_THREADS) {
    const int i = e / NB, p = e - i * NB;
    inv_out[e] = (p <= i) ? W[p * LDSM + i + 1] : 0.0;
  }
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
