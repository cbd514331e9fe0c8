/*
 * npw_b200.h — C-ABI of libnpw_b200.so: the fp64 tile kernels of the numpywren
 * LambdaPACK hot path, re-implemented for NVIDIA B200 (sm_100a).
 *
 * Every entry point replaces ONE tile operation of the reference
 * (/root/reference/numpywren/kernels.py) at the seam where the reference's
 * worker calls it: lambdapack.py:344-384 (RemoteCall.compute →
 * `self.compute(*pyarg_list, **self.kwargs)`).  The reference passes NumPy
 * arrays; this ABI passes borrowed DEVICE pointers to row-major (C-order) fp64
 * tiles with an explicit leading dimension (in elements), plus the CUDA stream
 * to enqueue on.  All calls are asynchronous on `stream`, allocate nothing, throw
 * nothing, and are re-entrant across streams and devices (the current device of
 * the calling thread must be the one that owns the pointers).
 *
 * Return value: 0 = enqueued; <0 = bad argument (-k: k-th argument, LAPACK
 * style) ; NPW_ERR_CUDA = a CUDA runtime/driver call failed (see
 * npw_last_error()).  Numerical failure of npw_potrf_l_f64 (non-positive pivot)
 * is reported asynchronously through `info_dev`, like LAPACK's INFO.
 *
 * No torch / Python types appear here: the library is usable from C, ctypes,
 * cgo or JNI alike.  See INTEGRATION.md for the reference-side binding.
 */
#ifndef NPW_B200_H
#define NPW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPW_OK 0
#define NPW_ERR_CUDA (-1000)
#define NPW_ERR_UNSUPPORTED (-1001)

/* Opaque CUDA stream handle (cudaStream_t).  NULL = legacy default stream. */
typedef void* npw_stream_t;

/* Library identity / diagnostics. */
int npw_version(void);                 /* major*10000 + minor*100 + patch */
const char* npw_last_error(void);      /* thread-local text of the last failure */
const char* npw_build_arch(void);      /* "sm_100a" */

/* ------------------------------------------------------------------------
 * kernels.syrk(s, x, y) = s - x.dot(y.T)          (kernels.py:212-215)
 *   C_out[m,n] = S[m,n] - X[m,k] * Y[n,k]^T ; C_out may alias S exactly
 *   (ldc == lds), never X or Y.  The reference's allclose(x,0)/allclose(y,0)
 *   short-circuit returns s unchanged, which this kernel reproduces to within
 *   one rounding of 0.
 * ---------------------------------------------------------------------- */
int npw_syrk_f64(double* C_out, int64_t ldc,
                 const double* S, int64_t lds,
                 const double* X, int64_t ldx,
                 const double* Y, int64_t ldy,
                 int64_t m, int64_t n, int64_t k, npw_stream_t stream);

/* Same update restricted to the 128x128 CTA tiles that touch the lower triangle
 * (row >= col): the scheduler uses it for DIAGONAL tiles S[i,j,j], whose strict
 * upper triangle is never read again (kernels.chol reads the lower triangle
 * only, kernels.py:225-226).  Elements in skipped tiles are copied from S. */
int npw_syrk_lower_f64(double* C_out, int64_t ldc,
                       const double* S, int64_t lds,
                       const double* X, int64_t ldx,
                       const double* Y, int64_t ldy,
                       int64_t m, int64_t n, int64_t k, npw_stream_t stream);

/* ------------------------------------------------------------------------
 * kernels.gemm(A, B, transpose_A=False, transpose_B=False) = op(A).dot(op(B))
 *                                                   (kernels.py:239-244)
 *   General form C = alpha*op(A)*op(B) + beta*C0, op(A) is m x k, op(B) is k x n.
 *   transA/transB are 0/1.  C0 may be NULL when beta == 0; C may alias C0.
 * ---------------------------------------------------------------------- */
int npw_gemm_f64(double* C, int64_t ldc,
                 const double* C0, int64_t ldc0,
                 const double* A, int64_t lda, int transA,
                 const double* B, int64_t ldb, int transB,
                 int64_t m, int64_t n, int64_t k,
                 double alpha, double beta, npw_stream_t stream);

/* ------------------------------------------------------------------------
 * kernels.trsm(x, y) with the only arguments the DSL ever passes
 * (lower=False, right=True; frontend.py:346 drops kwargs) =
 * scipy.linalg.blas.dtrsm(1.0, x.T, y, lower=0, side=1) = y * x^{-T}
 *                                                   (kernels.py:254-257)
 *   B_out[m,n] = B[m,n] * L[n,n]^{-T},  L lower-triangular (strict upper
 *   ignored).  B_out may alias B (the solve works in place in B_out).
 *   `invdiag` and `work` are accepted for ABI stability and ignored: the
 *   128-column leaves substitute against the diagonal blocks of L directly
 *   (npw_trsm_work_bytes returns a token size).
 * ---------------------------------------------------------------------- */
size_t npw_trsm_work_bytes(int64_t m, int64_t n);
int npw_trsm_rlt_f64(double* B_out, int64_t ldbo,
                     const double* L, int64_t ldl,
                     const double* B, int64_t ldb,
                     int64_t m, int64_t n,
                     const double* invdiag, void* work, npw_stream_t stream);

/* Inverses of the NPW_DIAG_NB x NPW_DIAG_NB diagonal blocks of lower-triangular
 * L (n x n): invdiag is ceil(n/NB) consecutive NB x NB row-major blocks. */
#define NPW_DIAG_NB 128
size_t npw_invdiag_bytes(int64_t n);
int npw_trtri_diag_f64(double* invdiag, const double* L, int64_t ldl, int64_t n,
                       npw_stream_t stream);

/* ------------------------------------------------------------------------
 * kernels.chol(x) = np.linalg.cholesky(x)           (kernels.py:225-226)
 *   L_out = lower Cholesky factor of the symmetric matrix whose LOWER triangle
 *   is in A (n x n); the strict upper triangle of L_out is zeroed, like
 *   np.linalg.cholesky.  L_out may alias A.  *info_dev (device int32) is set to
 *   0 on success or to the 1-based index of the first non-positive pivot
 *   (LAPACK dpotrf INFO; the Python shim raises LinAlgError).  `invdiag_out`
 *   (optional, npw_invdiag_bytes(n)) receives the inverted diagonal blocks for
 *   later npw_trsm_rlt_f64 calls.  `work`: npw_potrf_work_bytes(n).
 * ---------------------------------------------------------------------- */
size_t npw_potrf_work_bytes(int64_t n);
int npw_potrf_l_f64(double* L_out, int64_t ldl,
                    const double* A, int64_t lda, int64_t n,
                    int32_t* info_dev, double* invdiag_out, void* work,
                    npw_stream_t stream);

/* ------------------------------------------------------------------------
 * kernels.add_matrices(*args) = zeros + sum(args)   (kernels.py:16-20)
 *   out[i] = sum_{c<count} ptrs[c][i]; 1 <= count <= 8; contiguous nelem.
 *   `ptrs` is a HOST array of device pointers.  out may alias ptrs[0].
 * ---------------------------------------------------------------------- */
int npw_addn_f64(double* out, const double* const* ptrs, int count,
                 int64_t nelem, npw_stream_t stream);

/* kernels.mul(x, y) = x * y (kernels.py:233-234), contiguous nelem. */
int npw_mul_f64(double* out, const double* x, const double* y, int64_t nelem,
                npw_stream_t stream);

/* kernels.identity(x) (kernels.py:236-237) and BigMatrixView transposed reads
 * (matrix.py:643-661): strided 2-D copy / transpose.
 *   trans=0: dst[r,c] = src[r,c]   (rows x cols)
 *   trans=1: dst[c,r] = src[r,c]   (dst is cols x rows) */
int npw_copy2d_f64(double* dst, int64_t ldd, const double* src, int64_t lds,
                   int64_t rows, int64_t cols, int trans, npw_stream_t stream);

/* BigMatrix.get_block's `lambdav` diagonal shift (matrix.py:307-309):
 * A[i,i] += lambdav for i < min(rows, cols). */
int npw_add_diag_f64(double* A, int64_t lda, int64_t rows, int64_t cols,
                     double lambdav, npw_stream_t stream);

/* matrix_utils.constant_zeros parent_fn (matrix_utils.py:314-317) and
 * np.triu/np.tril masks used by the QR kernels (kernels.py:99-104).
 *   mode 0: set all rows x cols to `value`
 *   mode 1: keep upper triangle (j >= i), set the rest to `value`
 *   mode 2: keep lower triangle (j <= i), set the rest to `value` */
int npw_fill2d_f64(double* A, int64_t lda, int64_t rows, int64_t cols,
                   int mode, double value, npw_stream_t stream);

/* ------------------------------------------------------------------------
 * kernels.qr_factor(*blocks) -> (V, T, R)  (kernels.py:127-130 → fast_qr
 * :86-105: LAPACK dgeqrt3 compact-WY QR).  A (m x n, m >= n) is the vertical
 * stack of the blocks.  Outputs: V (m x n, unit lower trapezoidal, explicit
 * ones on the diagonal and zeros above), T (n x n upper triangular),
 * R (n x n upper triangular) with Q = I - V T V^T and A = Q[:, :n] R.
 * V may alias A.  `work`: npw_geqrt_work_bytes(m, n).
 * ---------------------------------------------------------------------- */
size_t npw_geqrt_work_bytes(int64_t m, int64_t n);
int npw_geqrt_f64(double* V, int64_t ldv, double* T, int64_t ldt,
                  double* R, int64_t ldr,
                  const double* A, int64_t lda, int64_t m, int64_t n,
                  void* work, npw_stream_t stream);

/* ------------------------------------------------------------------------
 * kernels.qr_factor_triangular(x0, x1) -> (V, T, R)  (kernels.py:132-134 →
 * fast_qr_triangular :107-124: LAPACK dtpqrt with m = n, l = m).  QR of
 * [triu(R0); triu(R1)] for two n x n factors (only their upper triangles are
 * read).  Outputs: V2 (n x n upper triangular: the bottom half of the
 * reflectors, Q = I - [I; V2] T [I; V2]^T), T (n x n upper, the single
 * compact-WY factor, i.e. dtpqrt with nb = n), R (n x n upper).  The Python shim
 * derives the reference's literal return values from these (qr.py).
 * `work`: npw_tpqrt_work_bytes(n).  No output may alias an input.
 * ---------------------------------------------------------------------- */
size_t npw_tpqrt_work_bytes(int64_t n);
int npw_tpqrt_f64(double* V2, int64_t ldv, double* T, int64_t ldt,
                  double* R, int64_t ldr,
                  const double* R0, int64_t ld0, const double* R1, int64_t ld1,
                  int64_t n, void* work, npw_stream_t stream);

/* ------------------------------------------------------------------------
 * OPTIONAL, off by default (the engine uses it only with NPW_B200_SYRK=i8emu; ran and
 * passed parity on B200 in round 2, tests/test_i8emu_gpu.py):
 * kernels.syrk (kernels.py:212-215) with the product X Y^T emulated on the int8
 * tensor cores (tcgen05.mma kind::i8) — DESIGN.md §8, tools/ozaki_prototype.py.
 *   npw_split_i8_f64: X (rows x k, fp64) -> `ndigits` signed int8 digit planes
 *     (digits[p][row][col], k contiguous, 6 + 7 + 7 ... bits) and per-row exponents,
 *     X = diag(2^e) sum_p 2^-(6+7p) X_p exactly up to the last digit.
 *     digits: npw_i8_digits_bytes(rows, k, ndigits) bytes; exponents: rows int32.
 *   npw_syrk_i8emu_f64: C = S - X Y^T from the digits of X (m x k) and Y (n x k);
 *     m % 128 == 0, n % 64 == 0, k % 128 == 0, 1 <= ndigits <= 8; C may alias S;
 *     lower_only skips the 128 x 64 tiles strictly above the diagonal.
 * ---------------------------------------------------------------------- */
size_t npw_i8_digits_bytes(int64_t rows, int64_t k, int ndigits);
int npw_split_i8_f64(int8_t* digits, int32_t* exponents, const double* X, int64_t ldx,
                     int64_t rows, int64_t k, int ndigits, npw_stream_t stream);
int npw_syrk_i8emu_f64(double* C, int64_t ldc, const double* S, int64_t lds,
                       const int8_t* xdigits, const int32_t* xexp,
                       const int8_t* ydigits, const int32_t* yexp,
                       int64_t m, int64_t n, int64_t k, int ndigits, int lower_only,
                       npw_stream_t stream);

/* Device-side synthetic tile generator used by bench/tests (not a reference
 * op): counter-based uniform(-1,1) fill, reproducible from (seed, row, col). */
int npw_fill_random_f64(double* A, int64_t lda, int64_t rows, int64_t cols,
                        uint64_t seed, int64_t row0, int64_t col0,
                        npw_stream_t stream);

/* Diagnostic used by bench.py for the roofline denominator: a register-only loop of
 * independent DMMA.8x8x4 instructions on every SM (`warps_per_sm` warps x `iters`
 * iterations x 16 accumulator tiles).  *flops_out (host) receives the flops issued.
 * `scratch`: npw_fp64_pipe_probe_bytes(warps_per_sm) bytes of device memory. */
size_t npw_fp64_pipe_probe_bytes(int warps_per_sm);
int npw_fp64_pipe_probe(double* scratch, int iters, int warps_per_sm, double* flops_out,
                        npw_stream_t stream);

/* Number of kernel launches this library has enqueued since load (all threads);
 * bench.py reports its delta over the timed region as `gpu_launches`. */
uint64_t npw_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NPW_B200_H */
