// npw_ozaki_i8.cu — fp64 syrk emulated on the int8 tensor cores (tcgen05.mma kind::i8), DESIGN.md §8.
//
// STATUS: optional path, OFF by default (the engine uses it only with NPW_B200_SYRK=i8emu; bench.py then labels the
// line's dtype "f64 emulated on int8 tensor cores").  First ran on a B200 in round 2: descriptor probe, digit extraction
// and products bit-for-bit against the CPU prototype (tools/ozaki_prototype.py), 4096^3 in 2.1 / 2.6 / 3.1 ms with
// 6 / 7 / 8 digits against 3.83 ms native (profiles/r02c_syrk_i8emu_timing.jsonl), tensor pipe 30 % active
// (profiles/r02c_ozaki_syrk_i8_kernel_ncu.json).  Tests: tests/test_i8emu_gpu.py (incl. a whole Cholesky vs the oracle).
//
// Replaces (optionally) the arithmetic of kernels.syrk (reference kernels.py:212-215, C = S - X Y^T) by
//   1. npw_split_i8_f64 : every row of X (m x k fp64) is scaled by 2^-e_i (e_i = exponent of the row's largest entry)
//                         and cut into s signed int8 digit planes of 6, 7, 7, ... bits:  X = diag(2^e) sum_p 2^-w_p X_p,
//                         w_p = 6 + 7p.  Exact (scaling by powers of two, rint, subtraction).  HBM-bound, done once per
//                         panel tile and reused by all syrks of that tile's block row / column.
//   2. npw_syrk_i8emu_f64: C = S - diag(2^ex) (sum_d 2^-(12+7d) P_d) diag(2^ey),  P_d = sum_{p+q=d} X_p Y_q^T in int32
//                         (exact: |P_d| <= k 2^12 (d+1) < 2^31 for k <= 65536 / (d+1)); pairs with p+q > s-1 are dropped.
//
// Kernel 2 (one CTA per 128 x 64 output tile, 7 warps):
//   warp 6   TMA producer of Y: per 128-byte k-block the s digit tiles of Y (64 rows x 128 B each) into one of two
//            Y buffers;  warp 0: TMA producer of X: the s digit tiles of X (128 rows x 128 B) one by one into a ring of
//            NX slots.  Two independent producers, so that the prefetch of the next k-block's Y digits never queues
//            behind an X slot that is still being read.  SWIZZLE_128B, 3-D tensor maps (k, row, digit plane).
//   warp 1   MMA issuer (one elected lane).  For digit p of X and every q <= s-1-p:  4 x tcgen05.mma (K = 32 bytes)
//            M = 128, N = 64, accumulating into TMEM columns [64 (p+q), 64 (p+q+1)) — all s group accumulators stay
//            resident in TMEM (s x 64 <= 512 columns), so every operand byte is loaded once per k-block.
//            tcgen05.commit releases an X slot after its last pair and a Y buffer after the k-block.
//   warps 2-5 epilogue: tcgen05.ld 16 columns at a time per group, int32 -> fp64, weights, row/column scales, S - (.).
#include "npw_common.cuh"

namespace npw {
namespace {

constexpr int OZ_BM = 128;            // rows of X per CTA (MMA M)
constexpr int OZ_BN = 64;             // rows of Y per CTA (MMA N)
constexpr int OZ_BK = 128;            // int8 elements = bytes per k-block (one 128-byte swizzle row)
constexpr int OZ_UK = 32;             // K of one kind::i8 MMA
constexpr int OZ_NX = 4;              // X ring slots
constexpr int OZ_MAXS = 8;            // 8 x 64 = 512 TMEM columns
constexpr int OZ_THREADS = 224;           // warps: 0 X producer, 1 MMA issuer, 2-5 epilogue, 6 Y producer
constexpr int OZ_XTILE = OZ_BM * OZ_BK;   // 16 KB
constexpr int OZ_YTILE = OZ_BN * OZ_BK;   // 8 KB

__host__ __device__ constexpr int oz_smem_bytes(int s) {
  return OZ_NX * OZ_XTILE + 2 * s * OZ_YTILE + 256;     // tiles + barriers / tmem pointer
}

// ------------------------------------------------------------------------------------------------ digit extraction
// One warp per row.  e = ceil(log2(max |x|)) (0 for an all-zero row), r = x 2^-e in [-1, 1];
// digit 0: q = rint(64 r), r = 64 r - q;  digit p > 0: q = rint(128 r), r = 128 r - q   (|q| <= 64 always).
__global__ void __launch_bounds__(256) split_i8_kernel(int8_t* __restrict__ digits, int32_t* __restrict__ expo,
                                                       const double* __restrict__ X, int64_t ldx, int rows, int k, int s) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const double* x = X + static_cast<int64_t>(warp) * ldx;
  double amax = 0.0;
  for (int c = lane; c < k; c += 32) amax = fmax(amax, fabs(x[c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  int e = 0;
  if (amax > 0.0) {
    int ex;
    const double m = frexp(amax, &ex);          // amax = m 2^ex, m in [0.5, 1)
    e = (m == 0.5) ? ex - 1 : ex;               // ceil(log2(amax))
  }
  if (lane == 0) expo[warp] = e;
  const int64_t plane = static_cast<int64_t>(rows) * k;
  int8_t* d = digits + static_cast<int64_t>(warp) * k;
  for (int c = lane; c < k; c += 32) {
    double r = scalbn(x[c], -e);
    for (int p = 0; p < s; ++p) {
      r *= (p == 0) ? 64.0 : 128.0;
      const double q = rint(r);
      r -= q;
      d[p * plane + c] = static_cast<int8_t>(static_cast<int>(q));
    }
  }
}

// ------------------------------------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: the mbarrier receives one arrival when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (1 for swizzled K-major)
//   bits [32,46) stride byte offset >> 4 = 1024 >> 4 (8 rows x 128 B per swizzle atom)
//   bits [46,48) version = 1 (sm_100)        bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, signed int8 A and B, both K-major, M = 128, N = 64
__device__ __forceinline__ uint32_t umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct OzArgs {
  double* C;
  const double* S;
  int64_t ldc, lds;
  const int32_t* ex;     // row exponents of X (m)
  const int32_t* ey;     // row exponents of Y (n)
  int m, n, k, s;
  int lower_only;
};

// ------------------------------------------------------------------------------------------------ the syrk kernel
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_syrk_i8_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const OzArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int s = p.s;
  uint8_t* xring = smem;                                   // OZ_NX x 16 KB
  uint8_t* ybuf = smem + OZ_NX * OZ_XTILE;                 // 2 x s x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(ybuf + 2 * s * OZ_YTILE);
  uint64_t* x_full = bars;                                 // [OZ_NX]
  uint64_t* x_empty = bars + OZ_NX;                        // [OZ_NX]
  uint64_t* y_full = bars + 2 * OZ_NX;                     // [2]
  uint64_t* y_empty = bars + 2 * OZ_NX + 2;                // [2]
  uint64_t* acc_full = bars + 2 * OZ_NX + 4;               // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * OZ_NX + 5);

  const int tile_m = blockIdx.y, tile_n = blockIdx.x;
  if (p.lower_only && tile_n * OZ_BN > tile_m * OZ_BM + (OZ_BM - 1)) {
    if (p.S != p.C) {
      for (int e = threadIdx.x; e < OZ_BM * OZ_BN; e += OZ_THREADS) {
        const int r = tile_m * OZ_BM + e / OZ_BN, c = tile_n * OZ_BN + e % OZ_BN;
        p.C[static_cast<int64_t>(r) * p.ldc + c] = p.S[static_cast<int64_t>(r) * p.lds + c];
      }
    }
    return;
  }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_cols = s * OZ_BN <= 32 ? 32u : s * OZ_BN <= 64 ? 64u : s * OZ_BN <= 128 ? 128u
                             : s * OZ_BN <= 256 ? 256u : 512u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < OZ_NX; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, tmem_cols);          // one full warp allocates (and frees) the accumulators
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int kblocks = p.k / OZ_BK;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer: X digits
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      int xit = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        for (int pd = 0; pd < s; ++pd, ++xit) {
          const int slot = xit % OZ_NX;
          mbar_wait(&x_empty[slot], ((xit / OZ_NX) & 1) ^ 1);
          mbar_arrive_expect_tx(&x_full[slot], static_cast<uint32_t>(OZ_XTILE));
          tma_load_3d(xring + slot * OZ_XTILE, &tmX, &x_full[slot], kb * OZ_BK, tile_m * OZ_BM, pd);
        }
      }
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------------------------------------ TMA producer: Y digits
    if (lane == 0) {
      tma_prefetch_desc(&tmY);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int yb = kb & 1;
        mbar_wait(&y_empty[yb], ((kb >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&y_full[yb], static_cast<uint32_t>(s * OZ_YTILE));
        for (int q = 0; q < s; ++q)
          tma_load_3d(ybuf + (yb * s + q) * OZ_YTILE, &tmY, &y_full[yb], kb * OZ_BK, tile_n * OZ_BN, q);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_i8(OZ_BM, OZ_BN);
      int xit = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int yb = kb & 1;
        mbar_wait(&y_full[yb], (kb >> 1) & 1);
        for (int pd = 0; pd < s; ++pd, ++xit) {
          const int slot = xit % OZ_NX;
          mbar_wait(&x_full[slot], (xit / OZ_NX) & 1);
          tc_fence_after();
          const uint32_t xaddr = smem_u32(xring + slot * OZ_XTILE);
          for (int q = 0; q + pd < s; ++q) {
            const uint32_t yaddr = smem_u32(ybuf + (yb * s + q) * OZ_YTILE);
            const uint32_t dcol = tmem_base + static_cast<uint32_t>((pd + q) * OZ_BN);
#pragma unroll
            for (int k4 = 0; k4 < OZ_BK / OZ_UK; ++k4) {
              // inside a 128-byte swizzled row the next 32-byte K slice is +32 bytes on the start address
              const uint64_t ad = umma_desc_k128(xaddr + k4 * OZ_UK);
              const uint64_t bd = umma_desc_k128(yaddr + k4 * OZ_UK);
              // group d = pd + q is first written by (kb = 0, pd = 0, q = d, k4 = 0)
              const uint32_t accumulate = (kb > 0 || pd > 0 || k4 > 0) ? 1u : 0u;
              umma_i8(dcol, ad, bd, idesc, accumulate);
            }
          }
          umma_commit(&x_empty[slot]);                       // slot reusable once these MMAs have read it
        }
        umma_commit(&y_empty[yb]);
      }
      umma_commit(acc_full);                                 // all accumulators final
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue (warps 2-5)
    const int quad = warp & 3;                               // TMEM lane quadrant this warp may read
    const int row = tile_m * OZ_BM + quad * 32 + lane;       // TMEM lane = output row
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const double rscale = scalbn(1.0, p.ex[row]);
    const double* srow = p.S + static_cast<int64_t>(row) * p.lds + tile_n * OZ_BN;
    double* crow = p.C + static_cast<int64_t>(row) * p.ldc + tile_n * OZ_BN;
    for (int c0 = 0; c0 < OZ_BN; c0 += 16) {
      double acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.0;
      for (int d = 0; d < s; ++d) {
        int32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(d * OZ_BN + c0), v);
        tmem_wait_ld();
        const double w = scalbn(1.0, -(12 + 7 * d));
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma(static_cast<double>(v[j]), w, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const double cs = scalbn(1.0, p.ey[tile_n * OZ_BN + c0 + j]);
        crow[c0 + j] = srow[c0 + j] - acc[j] * rscale * cs;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

bool g_oz_attr[64] = {};

}  // namespace

int make_tmap_i8_3d(CUtensorMap* map, const int8_t* base, int64_t rows, int64_t k, int64_t planes, uint32_t box_rows);

}  // namespace npw

extern "C" {

size_t npw_i8_digits_bytes(int64_t rows, int64_t k, int digits) {
  if (rows <= 0 || k <= 0 || digits <= 0) return 0;
  return static_cast<size_t>(digits) * rows * k;
}

int npw_split_i8_f64(int8_t* digits, int32_t* exponents, const double* X, int64_t ldx, int64_t rows, int64_t k, int ndigits,
                     npw_stream_t stream) {
  if (!digits) return -1;
  if (!exponents) return -2;
  if (!X) return -3;
  if (ldx < k) return -4;
  if (rows < 0 || rows > INT32_MAX) return -5;
  if (k < 0 || k > INT32_MAX) return -6;
  if (ndigits < 1 || ndigits > npw::OZ_MAXS) return -7;
  if (rows == 0 || k == 0) return NPW_OK;
  const int64_t threads = rows * 32;
  const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
  npw::split_i8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(digits, exponents, X, ldx, static_cast<int>(rows),
                                                                             static_cast<int>(k), ndigits);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

int npw_syrk_i8emu_f64(double* C, int64_t ldc, const double* S, int64_t lds, const int8_t* xdigits, const int32_t* xexp,
                       const int8_t* ydigits, const int32_t* yexp, int64_t m, int64_t n, int64_t k, int ndigits,
                       int lower_only, npw_stream_t stream) {
  using namespace npw;
  if (!C) return -1;
  if (ldc < n) return -2;
  if (!S) return -3;
  if (lds < n) return -4;
  if (!xdigits || !xexp) return -5;
  if (!ydigits || !yexp) return -7;
  if (ndigits < 1 || ndigits > OZ_MAXS) return -12;
  if (m == 0 || n == 0) return NPW_OK;
  if (m % OZ_BM || n % OZ_BN || k % OZ_BK || k <= 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) {
    set_error("syrk_i8emu: needs m %% 128 == 0, n %% 64 == 0, k %% 128 == 0 (got %lld x %lld x %lld)", (long long)m,
              (long long)n, (long long)k);
    return NPW_ERR_UNSUPPORTED;
  }
  // int32 exactness of the largest group: k * 64 * 64 * ndigits < 2^31
  if (k * 4096 * ndigits >= (int64_t(1) << 31)) {
    set_error("syrk_i8emu: k = %lld too large for exact int32 group sums with %d digits", (long long)k, ndigits);
    return NPW_ERR_UNSUPPORTED;
  }
  CUtensorMap tmX, tmY;
  if (make_tmap_i8_3d(&tmX, xdigits, m, k, ndigits, OZ_BM) != 0) return NPW_ERR_CUDA;
  if (make_tmap_i8_3d(&tmY, ydigits, n, k, ndigits, OZ_BN) != 0) return NPW_ERR_CUDA;
  int dev = 0;
  NPW_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && !g_oz_attr[dev]) {
    NPW_CUDA_CHECK(cudaFuncSetAttribute(ozaki_syrk_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        oz_smem_bytes(OZ_MAXS)));
    g_oz_attr[dev] = true;
  }
  OzArgs a;
  a.C = C; a.S = S; a.ldc = ldc; a.lds = lds; a.ex = xexp; a.ey = yexp;
  a.m = static_cast<int>(m); a.n = static_cast<int>(n); a.k = static_cast<int>(k); a.s = ndigits; a.lower_only = lower_only;
  dim3 grid(static_cast<unsigned>(n / OZ_BN), static_cast<unsigned>(m / OZ_BM));
  ozaki_syrk_i8_kernel<<<grid, OZ_THREADS, oz_smem_bytes(ndigits), static_cast<cudaStream_t>(stream)>>>(tmX, tmY, a);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

}  // extern "C"
