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

// One CTA of 1024 threads factors a 128x128 block AND inverts the factor, with the matrix held in REGISTERS:
// thread (tx, ty) = (tid & 31, tid >> 5) owns the 4x4 tile rows 4ty..4ty+3, cols 4tx..4tx+3 (lower tiles: tx <= ty).
// Right-looking, one __syncthreads per column:
//   loop 1 (Cholesky): the owners of column j publish it (unscaled) in a double-buffered shared vector; every thread
//           derives d = sqrt(a_jj), 1/d itself and applies the rank-1 update to its register tile (no shared RMW);
//   loop 2 (inverse X = L^{-1}, forward substitution by rows): the owners of row j of X publish it; every thread
//           updates X[i][t] -= L[i][j] * X[j][t] / L_jj in registers, reading column j of L from shared memory.
// do_factor = 0: the block already holds L (batched trtri of an existing factor).  Blocks smaller than NB are
// padded with the identity.  Shared: L as W[NB][NB+1] (132 KB) + two small vectors.
__device__ __forceinline__ void block_chol_inv(double* W, const double* Ain, int64_t lda, int nbk, double* Lout, int64_t ldl,
                                               double* inv_out, int32_t* info, int info_base, int do_factor, int write_l,
                                               double* s_vec /* [2][NB] + 2 pivots */, double* s_rinv /* [NB] */) {
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int i0 = 4 * ty, k0 = 4 * tx;
  const bool lower_tile = tx <= ty;
  double c[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + a, k = k0 + b;
      double v = (i == k) ? 1.0 : 0.0;
      if (lower_tile && k <= i && i < nbk && k < nbk) v = Ain[static_cast<int64_t>(i) * lda + k];
      c[a][b] = v;
    }
  int par = 0;
  if (do_factor) {
    // The pivot of column j+1 is final as soon as step j has updated the diagonal tile: its owner computes
    // 1/sqrt one step ahead and publishes it, so sqrt/div are executed by ONE thread per column, not by 1024.
    auto pivot = [&](double ajj, int j) -> double {   // returns L_jj, publishes 1/L_jj
      double d, rd;
      if (!(ajj > 0.0)) {
        atomicCAS(info, 0, info_base + j + 1);
        d = rd = nan("");
      } else {
        rd = rsqrt(ajj);
        d = ajj * rd;
        d = fma(0.5 * rd, fma(-d, d, ajj), d);          // one Newton step: d = sqrt(ajj) to < 1 ulp
        rd = 1.0 / d;
      }
      s_vec[2 * NB + (j & 1)] = rd;
      s_rinv[j] = rd;
      return d;
    };
    if (tid == 0) c[0][0] = pivot(c[0][0], 0);
    for (int j = 0; j < NB; ++j) {
      const int jb = j >> 2, jc = j & 3;
      if (tx == jb && lower_tile) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          double v = c[a][0];
          if (jc == 1) v = c[a][1];
          if (jc == 2) v = c[a][2];
          if (jc == 3) v = c[a][3];
          s_vec[par * NB + i0 + a] = v;
        }
      }
      __syncthreads();
      const double* col = s_vec + par * NB;
      const double rd = s_vec[2 * NB + (j & 1)];
      if (i0 + 3 >= j && lower_tile) {         // tiles entirely above row j are finished
        double li[4], lk[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) li[a] = col[i0 + a] * rd;
#pragma unroll
        for (int b = 0; b < 4; ++b) lk[b] = col[k0 + b] * rd;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int i = i0 + a, k = k0 + b;
            if (k > j && k <= i) c[a][b] = fma(-li[a], lk[b], c[a][b]);
            if (k == j && i > j) c[a][b] = li[a];     // scaled column j of L (the diagonal already holds L_jj)
          }
        const int jn = j + 1;
        if (jn < NB && tx == ty && tx == (jn >> 2)) {
          const int an = jn & 3;
          double ann = c[0][0];
          if (an == 1) ann = c[1][1];
          if (an == 2) ann = c[2][2];
          if (an == 3) ann = c[3][3];
          const double dn = pivot(ann, jn);
          if (an == 0) c[0][0] = dn;
          if (an == 1) c[1][1] = dn;
          if (an == 2) c[2][2] = dn;
          if (an == 3) c[3][3] = dn;
        }
      }
      par ^= 1;
    }
  } else if (tid < NB) {
    s_rinv[tid] = 1.0 / ((tid < nbk) ? Ain[static_cast<int64_t>(tid) * lda + tid] : 1.0);
  }
  // ---- L: registers -> shared (for loop 2) and -> global
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + a, k = k0 + b;
      const double v = (lower_tile && k <= i) ? c[a][b] : 0.0;
      W[i * LDSM + k] = v;
      if (write_l && i < nbk && k < nbk) Lout[static_cast<int64_t>(i) * ldl + k] = v;
      c[a][b] = (i == k) ? 1.0 : 0.0;          // becomes the X tile
    }
  __syncthreads();
  // ---- loop 2: X = L^{-1}
  for (int j = 0; j < NB; ++j) {
    const int jb = j >> 2, ja = j & 3;
    if (ty == jb && lower_tile) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double v = c[0][b];
        if (ja == 1) v = c[1][b];
        if (ja == 2) v = c[2][b];
        if (ja == 3) v = c[3][b];
        s_vec[par * NB + k0 + b] = v;          // unscaled row j of X
      }
    }
    __syncthreads();
    const double* row = s_vec + par * NB;
    const double rd = s_rinv[j];
    if (i0 + 3 >= j && k0 <= j && lower_tile) {
      double xj[4], lj[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) xj[b] = row[k0 + b] * rd;
#pragma unroll
      for (int a = 0; a < 4; ++a) lj[a] = W[(i0 + a) * LDSM + j];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = i0 + a, t = k0 + b;
          if (i > j && t <= j) c[a][b] = fma(-lj[a], xj[b], c[a][b]);
          if (i == j && t <= j) c[a][b] = xj[b];      // scaled row j of X
        }
    }
    par ^= 1;
  }
  if (inv_out) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = i0 + a, t = k0 + b;
        inv_out[i * NB + t] = (lower_tile && t <= i) ? c[a][b] : 0.0;
      }
  }
}

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* Ablk, int64_t lda, int nbk, double* inv_out, int32_t* info, int info_base, int do_factor,
                 int write_l) {
  extern __shared__ double W[];
  __shared__ double s_vec[2 * NB + 2];
  __shared__ double s_rinv[NB];
  block_chol_inv(W, Ablk, lda, nbk, Ablk, lda, inv_out, info, info_base, do_factor, write_l, s_vec, s_rinv);
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
  extern __shared__ double W[];
  __shared__ double s_vec[2 * NB + 2];
  __shared__ double s_rinv[NB];
  const int blk = blockIdx.x;
  const int j0 = blk * NB;
  const int nbk = min(NB, n - j0);
  block_chol_inv(W, L + static_cast<int64_t>(j0) * ldl + j0, ldl, nbk, nullptr, 0,
                 invdiag + static_cast<int64_t>(blk) * NB * NB, nullptr, 0, 0, 0, s_vec, s_rinv);
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
