// npw_qr_f64.cu — compact-WY Householder QR of a tall tile: kernels.qr_factor (kernels.py:127-130) → fast_qr
// (kernels.py:86-105, LAPACK dgeqrt3): A (m x n, m >= n) -> V (unit lower trapezoidal), T (n x n upper), R (n x n upper),
// Q = I - V T V^T.  Same Householder convention as LAPACK dlarfg/dlarft, so V, T, R match the reference up to rounding.
//
// Blocked right-looking algorithm, panel width 32:
//   panel   : ONE kernel (co-resident grid, <= 1 CTA per SM) factors the (m-j0) x 32 panel.  Rows are split over the
//             CTAs; each CTA keeps its rows in REGISTERS (qr_panel_reg_kernel: 8 warps x 56 rows x 32 lanes, up to 448
//             rows per CTA; tiles with more rows per CTA use the shared-memory / global variants qr_panel_kernel) for all
//             32 column steps.  Per column there is a single pass over the rows that (i) applies reflector j, (ii)
//             accumulates the dot products column j+1 needs (norm + w = P^T x) and (iii) the Gram entries
//             V[:,0:j]^T v_j that dlarft needs, all in one 32-lane vector, followed by ONE all-gather of the per-CTA
//             vectors.  The all-gather is flag-in-data: every double travels as two 64-bit packets {32 payload bits |
//             32-bit sequence number}; readers poll the packets themselves, so a column step costs L2 store->load
//             latencies (two levels: groups of 12 CTAs, then the group sums; one level for <= 12 CTAs) instead of a grid
//             barrier plus a reduction (round 1: 14-25 us per column with cooperative grid.sync; now 4.4 us).  Every
//             CTA adds the vectors in the same fixed order, so all CTAs derive bit-identical Householder scalars and
//             the result is deterministic.  Protocol model check: tests/test_qr_gather_protocol.py.
//   update  : Wt = C^T V_p  split-K "TN" product on the fp64 tensor pipe (DMMA, k = rows), reduced in a fixed order
//             and multiplied by T_p in the same kernel;  C -= V_p (Wt T_p)^T  by a streaming rank-32 DMMA kernel
//             (HBM-bound: 2 x 32 flop per element read and written).
//   T       : Gram matrix G = V^T V (same TN kernel, upper tiles only), then per panel
//             T[0:j0, panel] = -T[0:j0,0:j0] G[0:j0,panel] T_pp  (one small kernel per panel).
#include <stdlib.h>

#include <type_traits>

#include "npw_common.cuh"

namespace npw {

int launch_copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, int trans,
                  cudaStream_t st);
int launch_fill2d(double* A, int64_t lda, int64_t rows, int64_t cols, int mode, double value, cudaStream_t st);

namespace {

constexpr int QW = 32;            // panel width
constexpr int QTHREADS = 512;     // 16 warps
constexpr int QWARPS = QTHREADS / 32;
constexpr int MAX_GRID = 160;     // upper bound on the CTAs of a panel launch (<= SM count)
constexpr int QMAXQ = (MAX_GRID + QWARPS - 1) / QWARPS;   // vectors one reader group adds per step

// packet scratch, per step parity: [MAX_GRID][32 doubles x 2 packets] partial vectors, [32 x 2] the pivot row, then
// [QMAXGROUPS][32 x 2] group sums
constexpr int PK_PER_VEC = 2 * QW;
constexpr int QGSZ = 12;                                   // CTAs per group of the two-level gather (register-resident kernel)
constexpr int QMAXGROUPS = (MAX_GRID + QGSZ - 1) / QGSZ;   // 14
constexpr int PK_GROUP0 = MAX_GRID + 1;                    // group sums live behind the per-CTA vectors and the pivot row
constexpr int PK_STRIDE = (MAX_GRID + 1 + QMAXGROUPS) * PK_PER_VEC;
constexpr size_t PK_BYTES = 2 * static_cast<size_t>(PK_STRIDE) * sizeof(uint64_t);
constexpr long long QR_SPIN_LIMIT = 4000000000ll;          // ~2 s of polling: give up (sets *err) instead of hanging

struct PanelArgs {
  double* V;       // m x n working matrix (row-major, ldv); panel columns [j0, j0+w)
  int64_t ldv;
  int m, j0, w;
  double* R;       // n x n output, ldr
  int64_t ldr;
  double* T;       // n x n output, ldt (only the w x w diagonal block of this panel is written)
  int64_t ldt;
  double* tau;     // n
  uint64_t* packets;   // PK_BYTES, zeroed once per factorisation
  uint32_t seq0;       // first sequence number of this launch (unique within the factorisation, never 0)
  int* err;            // set to 1 if a poll timed out
  long long* prof;     // NPW_QR_PROFILE builds only: [cta][8] accumulated cycles per phase of a column step
};

// A double as two self-validating 64-bit packets (each 8-byte access is single-copy atomic, so no fence or separate
// flag is needed: a packet is either the old one or the complete new one).
__device__ __forceinline__ void ll_store(uint64_t* slot, double v, uint32_t seq) {
  const uint64_t bits = static_cast<uint64_t>(__double_as_longlong(v));
  const uint64_t tag = static_cast<uint64_t>(seq) << 32;
  const uint64_t w0 = (bits & 0xffffffffull) | tag;
  const uint64_t w1 = (bits >> 32) | tag;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ll_load(const uint64_t* slot, uint64_t& w0, uint64_t& w1) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
}
__device__ __forceinline__ bool ll_valid(uint64_t w0, uint64_t w1, uint32_t seq) {
  return static_cast<uint32_t>(w0 >> 32) == seq && static_cast<uint32_t>(w1 >> 32) == seq;
}
__device__ __forceinline__ double ll_value(uint64_t w0, uint64_t w1) {
  return __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
}

// SMEM = true: the CTA's row chunk of the panel (<= QR_SMEM_ROWS rows x 32 columns) is loaded into shared memory once,
// all 32 column passes run there, and the chunk is written back once (2 global passes instead of 64).
constexpr int QR_SMEM_ROWS = 672;                    // 672 x 32 x 8 B = 168 KB (+ ~25 KB static)

template <bool SMEM>
__global__ void __launch_bounds__(QTHREADS, 1) qr_panel_kernel(PanelArgs p) {
  extern __shared__ __align__(16) double s_chunk[];
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

  __shared__ double s_part[QWARPS][QW];              // per-warp partial vectors of a pass
  __shared__ double s_red[QWARPS][QW];               // per-reader-group sums of the gathered vectors
  __shared__ double s_T[QW][QW + 1];
  __shared__ double s_Z[QW][QW];                     // Gram vectors z_j = V[:,0:j]^T v_j (CTA 0 builds T from them)
  __shared__ double s_tau[QW];

  for (int e = threadIdx.x; e < QW * (QW + 1); e += QTHREADS) (&s_T[0][0])[e] = 0.0;

  auto rowptr = [&](int r) -> double* {
    return SMEM ? s_chunk + static_cast<size_t>(r - r_lo) * QW : Vp + static_cast<int64_t>(r) * p.ldv;
  };
  if (SMEM) {
    for (int r = r_lo + warp; r < r_hi; r += QWARPS)
      s_chunk[static_cast<size_t>(r - r_lo) * QW + lane] = lane_ok ? Vp[static_cast<int64_t>(r) * p.ldv + lane] : 0.0;
    __syncthreads();
  }

  // publish this CTA's vector for step `step` (warp 0) and, if it owns row `prow`, that row of the panel (warp 2)
  auto publish = [&](int step, int prow) {
    uint64_t* base = p.packets + static_cast<size_t>(step & 1) * PK_STRIDE;
    const uint32_t seq = p.seq0 + static_cast<uint32_t>(step);
    if (warp == 0) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < QWARPS; ++q) s += s_part[q][lane];
      ll_store(base + static_cast<size_t>(cta) * PK_PER_VEC + 2 * lane, s, seq);
    } else if (warp == 2 && prow >= r_lo && prow < r_hi) {
      ll_store(base + static_cast<size_t>(MAX_GRID) * PK_PER_VEC + 2 * lane, lane_ok ? rowptr(prow)[lane] : 0.0, seq);
    }
  };
  // all-gather + reduction of step `step`: returns g[lane] (sum over all CTAs, fixed order) and the pivot row entry
  auto gather = [&](int step, bool need_pivot, double& g_l, double& prow_l) {
    const uint64_t* base = p.packets + static_cast<size_t>(step & 1) * PK_STRIDE;
    const uint32_t seq = p.seq0 + static_cast<uint32_t>(step);
    uint64_t a0[QMAXQ], a1[QMAXQ], b0 = 0, b1 = 0;
    const long long t0 = clock64();
    bool ok = *static_cast<volatile int*>(p.err) != 0;     // a poll that timed out earlier: do not wait again
    while (!ok) {
#pragma unroll
      for (int u = 0; u < QMAXQ; ++u) {
        const int q = warp + u * QWARPS;
        if (q < G) ll_load(base + static_cast<size_t>(q) * PK_PER_VEC + 2 * lane, a0[u], a1[u]);
      }
      if (need_pivot) ll_load(base + static_cast<size_t>(MAX_GRID) * PK_PER_VEC + 2 * lane, b0, b1);
      ok = !need_pivot || ll_valid(b0, b1, seq);
#pragma unroll
      for (int u = 0; u < QMAXQ; ++u) {
        const int q = warp + u * QWARPS;
        if (q < G) ok = ok && ll_valid(a0[u], a1[u], seq);
      }
      if (!ok && clock64() - t0 > QR_SPIN_LIMIT) {
        *static_cast<volatile int*>(p.err) = 1;
        break;
      }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < QMAXQ; ++u) {
      const int q = warp + u * QWARPS;
      if (q < G) s += ll_value(a0[u], a1[u]);
    }
    s_red[warp][lane] = s;
    prow_l = need_pivot ? ll_value(b0, b1) : 0.0;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < QWARPS; ++q) t += s_red[q][lane];
    g_l = t;
  };

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
    publish(0, p.j0);
  }

  for (int j = 0; j < w; ++j) {
    const int gj = p.j0 + j;                         // pivot row / column (global)
    double g_l, prow_l;
    gather(j, true, g_l, prow_l);
    // lanes c >= j of g: dot products for column j (c == j: squared norm below the pivot);
    // lanes i < j-1: Gram entries of column j-1 (kept for T)
    if (j > 0 && cta == 0 && warp == 1) s_Z[j - 1][lane] = g_l;
    // ---- Householder scalars (dlarfg), redundantly in every thread (identical inputs -> identical results)
    const double alpha = __shfl_sync(0xffffffffu, prow_l, j);
    const double xn2 = __shfl_sync(0xffffffffu, g_l, j);
    double tau = 0.0, scale = 0.0, beta = alpha;
    if (xn2 > 0.0) {
      const double nrm = sqrt(alpha * alpha + xn2);
      beta = alpha >= 0.0 ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    if (threadIdx.x == 0) s_tau[j] = tau;
    // w_c = v^T P[:, c] = prow[c] + scale * g[c]   (c > j)
    const double wc = (lane > j && lane_ok) ? prow_l + scale * g_l : 0.0;
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
      const double xn = __shfl_sync(0xffffffffu, v, (j + 1) & 31);
      if (j + 1 < w && r > gj + 1 && lane > j) acc = fma(xn, v, acc);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    publish(j + 1, (j + 1 < w && gj + 1 < p.m) ? gj + 1 : -1);
  }
  // ---- last column's Gram vector, then T of this panel (dlarft, forward columnwise) by one warp of CTA 0
  {
    double g_l, prow_l;
    gather(w, false, g_l, prow_l);
    if (cta == 0 && warp == 1) {
      s_Z[w - 1][lane] = g_l;
      __syncwarp();
      for (int jj = 0; jj < w; ++jj) {
        const double tj = s_tau[jj];
        double t = 0.0;
        if (lane < jj)
          for (int q = lane; q < jj; ++q) t = fma(s_T[lane][q], s_Z[jj][q], t);   // T[0:jj,0:jj] (upper) times z_jj
        if (lane < jj) s_T[lane][jj] = -tj * t;
        if (lane == jj) s_T[jj][jj] = tj;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (cta == 0) {
    for (int e = threadIdx.x; e < w * w; e += QTHREADS) {
      const int i = e / w, c = e - i * w;
      p.T[static_cast<int64_t>(p.j0 + i) * p.ldt + p.j0 + c] = (c >= i) ? s_T[i][c] : 0.0;
    }
    for (int e = threadIdx.x; e < w; e += QTHREADS) p.tau[p.j0 + e] = s_tau[e];
    // a timed-out poll means the factorisation is garbage: make that impossible to miss
    if (threadIdx.x == 0 && *static_cast<volatile int*>(p.err) != 0)
      p.T[static_cast<int64_t>(p.j0) * p.ldt + p.j0] = __longlong_as_double(0x7ff8000000000000ll);
  }
  if (SMEM) {                                        // write the factored chunk back (V below the diagonal, explicit unit/zeros)
    for (int r = r_lo + warp; r < r_hi; r += QWARPS)
      if (lane_ok) Vp[static_cast<int64_t>(r) * p.ldv + lane] = s_chunk[static_cast<size_t>(r - r_lo) * QW + lane];
  }
}

// ------------------------------------------------------------------------------------------------
// Register-resident variant of the panel kernel (the one a TSQR leaf uses): 8 warps, every warp keeps its RPW rows of
// the panel (lane = column) in registers for all 32 column steps, so a pass is shuffles and DFMAs only — no shared
// memory traffic, fully unrolled and branch-free, hence ILP across the rows of a warp.  A CTA holds up to
// 8 x RPW = 448 rows (65536 rows over 148 CTAs = 443).  Same arithmetic, same order of additions inside a warp's
// partial sums as the shared-memory variant up to the 4-way split of the accumulator; same all-gather protocol.
// ------------------------------------------------------------------------------------------------
constexpr int RTHREADS = 256, RWARPS = 8, RPW = 56;
constexpr int RFEW = 16;                 // short row loops for CTAs with at most RFEW x RWARPS = 128 rows
constexpr int RPIV = QW / RWARPS + 1;   // register rows of a warp that can be (or sit directly below) a pivot row of the launch

// 224 registers x 256 threads leave 8192 of the SM's 65536 registers free ON PURPOSE: the kernel spins on packets from its
// sibling CTAs, so every one of them must become resident — also on an SM where a tiny kernel that itself waits for
// another GPU (the tile exchange's wait_signal) is parked.  With the full 255 registers such an SM could never take its
// CTA, and the 8-GPU TSQR deadlocked until the exchange's signal timeout (profiles/r02i_*).
__global__ void __maxnreg__(224) qr_panel_reg_kernel(PanelArgs p) {
  const int G = gridDim.x;
  const int cta = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int rows = p.m - p.j0;
  const int per = (rows + G - 1) / G;                // <= RWARPS * RPW (host guarantees)
  const int r_lo = p.j0 + cta * per;
  const int r_hi = min(p.m, r_lo + per);
  const int w = p.w;
  const bool lane_ok = lane < w;
  double* Vp = p.V + p.j0;

  __shared__ double s_part[RWARPS][QW];
  __shared__ double s_red[RWARPS][QW];
  __shared__ double s_red2[RWARPS][QW];
  __shared__ double s_prow[QW];
  __shared__ double s_T[QW][QW + 1];
  __shared__ double s_Z[QW][QW];
  __shared__ double s_tau[QW];

  for (int e = threadIdx.x; e < QW * (QW + 1); e += RTHREADS) (&s_T[0][0])[e] = 0.0;

  // row i of this warp: global row r_lo + warp + RWARPS * i
  double v[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = r_lo + warp + RWARPS * i;
    v[i] = (r < r_hi && lane_ok) ? Vp[static_cast<int64_t>(r) * p.ldv + lane] : 0.0;
  }

  auto publish = [&](int step, double s) {            // warp 0: this CTA's vector for step `step`
    uint64_t* base = p.packets + static_cast<size_t>(step & 1) * PK_STRIDE;
    ll_store(base + static_cast<size_t>(cta) * PK_PER_VEC + 2 * lane, s, p.seq0 + static_cast<uint32_t>(step));
  };
  auto publish_row = [&](int step, int prow) {        // the warp that holds row `prow` publishes it (pivot row of `step`)
    if (prow < r_lo || prow >= r_hi || ((prow - r_lo) & (RWARPS - 1)) != warp) return;
    const int ip = (prow - r_lo) / RWARPS;
    double val = 0.0;
#pragma unroll
    for (int i = 0; i < RPIV; ++i) val = (i == ip) ? v[i] : val;   // pivot rows are among the CTA's first 32 + 1 rows
    uint64_t* base = p.packets + static_cast<size_t>(step & 1) * PK_STRIDE;
    ll_store(base + static_cast<size_t>(MAX_GRID) * PK_PER_VEC + 2 * lane, val, p.seq0 + static_cast<uint32_t>(step));
  };
  // Two-level all-gather: the first CTA of every group of QGSZ adds its group's vectors and publishes the group sum;
  // everybody adds the <= QMAXGROUPS group sums.  A thread polls at most three packets per sweep (round 2a read all 148
  // vectors in every CTA: 78 KB per CTA per sweep kept the L2 busy and a sweep took ~1 us).
  const int grp_id = cta / QGSZ;
  const bool leader = cta % QGSZ == 0;
  const int grp_cnt = min(QGSZ, G - grp_id * QGSZ);
  const int ngroups = (G + QGSZ - 1) / QGSZ;
  // poll up to three packets (slot pointers may be null = absent) until all present ones carry `seq`
  auto poll3 = [&](const uint64_t* s0, const uint64_t* s1, const uint64_t* s2, uint32_t seq, double& v0, double& v1, double& v2) {
    uint64_t a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0;
    const long long t0 = clock64();
    bool ok = false;
    unsigned sweeps = 0;
    while (!ok) {
      if (s0) ll_load(s0, a0, a1);
      if (s1) ll_load(s1, b0, b1);
      if (s2) ll_load(s2, c0, c1);
      ok = (!s0 || ll_valid(a0, a1, seq)) && (!s1 || ll_valid(b0, b1, seq)) && (!s2 || ll_valid(c0, c1, seq));
      // every 1024 failed sweeps: give up if this or an earlier poll of the factorisation ran out of time (the error
      // word is deliberately not read on the fast path: it would put an L2 round trip in front of every poll)
      if (!ok && (++sweeps & 1023u) == 0 &&
          (*static_cast<volatile int*>(p.err) != 0 || clock64() - t0 > QR_SPIN_LIMIT)) {
        *static_cast<volatile int*>(p.err) = 1;
        break;
      }
    }
    v0 = s0 ? ll_value(a0, a1) : 0.0;
    v1 = s1 ? ll_value(b0, b1) : 0.0;
    v2 = s2 ? ll_value(c0, c1) : 0.0;
  };
#ifdef NPW_QR_PROFILE
  long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long pt = clock64();
#define QR_STAMP(i) do { const long long _n = clock64(); pf[i] += _n - pt; pt = _n; } while (0)
#else
#define QR_STAMP(i) do { } while (0)
#endif
  auto gather = [&](int step, bool need_pivot, double& g_l, double& prow_l) {
    const uint64_t* base = p.packets + static_cast<size_t>(step & 1) * PK_STRIDE;
    const uint32_t seq = p.seq0 + static_cast<uint32_t>(step);
    QR_STAMP(7);                                            // whatever happened since the last stamp (publish etc.)
    if (ngroups == 1) {
      // a small grid (<= QGSZ CTAs, e.g. the 1024-row merges of a TSQR tree): everybody reads the members directly, one hop
      const int m0 = warp, m1 = warp + RWARPS;
      double v0, v1;
      double pv;
      poll3(m0 < G ? base + static_cast<size_t>(m0) * PK_PER_VEC + 2 * lane : nullptr,
            m1 < G ? base + static_cast<size_t>(m1) * PK_PER_VEC + 2 * lane : nullptr,
            (need_pivot && warp == 0) ? base + static_cast<size_t>(MAX_GRID) * PK_PER_VEC + 2 * lane : nullptr, seq, v0, v1, pv);
      s_red[warp][lane] = v0 + v1;
      if (warp == 0) s_prow[lane] = pv;
      __syncthreads();
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < RWARPS; ++q) t += s_red[q][lane];
      g_l = t;
      prow_l = s_prow[lane];
      return;
    }
    if (leader) {                                           // CTA-uniform
      const int m0 = warp, m1 = warp + RWARPS;
      double v0, v1, v2;
      poll3(m0 < grp_cnt ? base + static_cast<size_t>(grp_id * QGSZ + m0) * PK_PER_VEC + 2 * lane : nullptr,
            m1 < grp_cnt ? base + static_cast<size_t>(grp_id * QGSZ + m1) * PK_PER_VEC + 2 * lane : nullptr, nullptr, seq, v0, v1, v2);
      s_red2[warp][lane] = v0 + v1;
      __syncthreads();
      if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < RWARPS; ++q) t += s_red2[q][lane];
        ll_store(const_cast<uint64_t*>(base) + static_cast<size_t>(PK_GROUP0 + grp_id) * PK_PER_VEC + 2 * lane, t, seq);
      }
    }
    QR_STAMP(0);                                            // leader stage (0 for the others)
    const int g0 = warp, g1 = warp + RWARPS;
    // the pivot row is polled by ONE warp per CTA and handed on through shared memory: with all 8 warps of all 148 CTAs
    // polling the same 512 bytes, the reads queued at one L2 slice in front of the very store they were waiting for
    double v0, v1, pv;
    poll3(g0 < ngroups ? base + static_cast<size_t>(PK_GROUP0 + g0) * PK_PER_VEC + 2 * lane : nullptr,
          g1 < ngroups ? base + static_cast<size_t>(PK_GROUP0 + g1) * PK_PER_VEC + 2 * lane : nullptr,
          (need_pivot && warp == RWARPS - 1) ? base + static_cast<size_t>(MAX_GRID) * PK_PER_VEC + 2 * lane : nullptr, seq, v0, v1, pv);
    QR_STAMP(1);                                            // poll of the group sums
    s_red[warp][lane] = v0 + v1;
    if (warp == RWARPS - 1) s_prow[lane] = pv;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < RWARPS; ++q) t += s_red[q][lane];
    g_l = t;
    prow_l = s_prow[lane];
    QR_STAMP(2);                                            // barrier + sum
  };
  auto block_sum_and_publish = [&](int step, double acc) {
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < RWARPS; ++q) s += s_part[q][lane];
      publish(step, s);
    }
  };

  // ---- initial partials for column 0: g[c] = sum_{r > j0} x_r * P[r][c], x = column 0
  {
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = r_lo + warp + RWARPS * i;
      const double x = __shfl_sync(0xffffffffu, v[i], 0);
      a4[i & 3] = fma(r > p.j0 ? x : 0.0, v[i], a4[i & 3]);
    }
    publish_row(0, p.j0);
    block_sum_and_publish(0, (a4[0] + a4[1]) + (a4[2] + a4[3]));
  }

  for (int j = 0; j < w; ++j) {
    const int gj = p.j0 + j;
    double g_l, prow_l;
    gather(j, true, g_l, prow_l);
    if (j > 0 && cta == 0 && warp == 1) s_Z[j - 1][lane] = g_l;
    const double alpha = __shfl_sync(0xffffffffu, prow_l, j);
    const double xn2 = __shfl_sync(0xffffffffu, g_l, j);
    double tau = 0.0, scale = 0.0, beta = alpha;
    if (xn2 > 0.0) {
      const double nrm = sqrt(alpha * alpha + xn2);
      beta = alpha >= 0.0 ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    if (threadIdx.x == 0) s_tau[j] = tau;
    const double wc = (lane > j && lane_ok) ? prow_l + scale * g_l : 0.0;
    const double tw = tau * wc;
    QR_STAMP(3);                                            // Householder scalars
    const bool right = lane > j, diag = lane == j, left = lane < j;
    const bool next = j + 1 < w;
    // per-lane constants that turn the column roles into arithmetic: new = v * kv + v_j[r] * cv
    //   lane > j: v - v_j[r] tau w_c      lane == j: v_j[r]      lane < j: v (finished columns of V)
    const double kv = diag ? 0.0 : 1.0;
    const double cv = diag ? 1.0 : (right ? -tw : 0.0);

    double a4[4] = {0.0, 0.0, 0.0, 0.0};
    // Each row adds, per lane: lane < j: V[r][i] v_j[r] (Gram); lane > j: x'_r P'[r][c] (dot products of column j+1, rows
    // below the next pivot).  Lane j itself collects a product nobody reads (the gathers use lanes != j only).
    // The row loops exist in two unrolled lengths: all RPW register rows, or the first RFEW when the CTA holds at most
    // RFEW x RWARPS rows (the 1024-row merges of a TSQR tree have 128 rows per CTA: 16 of the 56 register rows are real).
    auto rows_all = [&](auto NRC) {
      constexpr int NR = decltype(NRC)::value;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const double x = __shfl_sync(0xffffffffu, v[i], j);
        const double vr = x * scale;                  // v_j[r]
        const double nv = fma(vr, cv, v[i] * kv);
        v[i] = nv;
        const double xn = __shfl_sync(0xffffffffu, nv, (j + 1) & 31);
        a4[i & 3] = fma(nv, left ? vr : xn, a4[i & 3]);
      }
    };
    auto rows_piv = [&](auto NRC) {
      constexpr int NR = decltype(NRC)::value;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = r_lo + warp + RWARPS * i;
        const double x = __shfl_sync(0xffffffffu, v[i], j);
        const double vr = (i >= RPIV || r > gj) ? x * scale : 0.0;
        const double nv = fma(vr, cv, v[i] * kv);
        v[i] = nv;
        const double xn = __shfl_sync(0xffffffffu, nv, (j + 1) & 31);
        a4[i & 3] = fma(nv, left ? vr : ((i >= RPIV || r > gj + 1) ? xn : 0.0), a4[i & 3]);
      }
    };
    using NFull = std::integral_constant<int, RPW>;
    using NFew = std::integral_constant<int, RFEW>;
    const bool few = per <= RFEW * RWARPS;
    if (r_lo > p.j0 + w) {
      // no pivot row in this CTA during this launch: every row is below the pivots (padding rows are all zero)
      if (few) rows_all(NFew{});
      else rows_all(NFull{});
    } else {
      // This CTA holds pivot rows of this launch.  They are among its first 32 rows (r_lo >= j0), i.e. register rows
      // i < RPIV of some warp: only those need the predicates; rows at or above the pivot get v_j[r] = 0 (nothing
      // changes; column j of a finished row is already 0).
      if (few) rows_piv(NFew{});
      else rows_piv(NFull{});
      // ---- pivot row (one warp of one CTA): R(gj, c) = P[gj][c] - tau w_c, diagonal beta; the row of V becomes e_j
      // (the loop above left its lanes != j untouched)
      if (gj >= r_lo && gj < r_hi && ((gj - r_lo) & (RWARPS - 1)) == warp) {
        const int ip = (gj - r_lo) / RWARPS;
        double pv = 0.0;
#pragma unroll
        for (int i = 0; i < RPIV; ++i) pv = (i == ip) ? v[i] : pv;
        if (left) a4[0] += pv;                        // Gram contribution v_i[gj] * 1
        if (right) pv -= tw;
        if (diag) pv = beta;
        if (lane_ok && !left) p.R[static_cast<int64_t>(gj) * p.ldr + p.j0 + lane] = pv;
        const double nv = diag ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < RPIV; ++i) v[i] = (i == ip && !left) ? nv : v[i];
      }
    }
    QR_STAMP(4);                                            // pass over the rows
    if (next) publish_row(j + 1, gj + 1);
    block_sum_and_publish(j + 1, (a4[0] + a4[1]) + (a4[2] + a4[3]));
    QR_STAMP(5);                                            // block sum + publish
  }
#ifdef NPW_QR_PROFILE
  if (p.prof != nullptr && (threadIdx.x == 0 || threadIdx.x == RTHREADS - 1))
    for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(p.prof) + (cta * 2 + (threadIdx.x ? 1 : 0)) * 8 + i, static_cast<unsigned long long>(pf[i]));
#endif
  {
    double g_l, prow_l;
    gather(w, false, g_l, prow_l);
    if (cta == 0 && warp == 1) {
      s_Z[w - 1][lane] = g_l;
      __syncwarp();
      for (int jj = 0; jj < w; ++jj) {
        const double tj = s_tau[jj];
        double t = 0.0;
        if (lane < jj)
          for (int q = lane; q < jj; ++q) t = fma(s_T[lane][q], s_Z[jj][q], t);
        if (lane < jj) s_T[lane][jj] = -tj * t;
        if (lane == jj) s_T[jj][jj] = tj;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (cta == 0) {
    for (int e = threadIdx.x; e < w * w; e += RTHREADS) {
      const int i = e / w, c = e - i * w;
      p.T[static_cast<int64_t>(p.j0 + i) * p.ldt + p.j0 + c] = (c >= i) ? s_T[i][c] : 0.0;
    }
    for (int e = threadIdx.x; e < w; e += RTHREADS) p.tau[p.j0 + e] = s_tau[e];
    if (threadIdx.x == 0 && *static_cast<volatile int*>(p.err) != 0)
      p.T[static_cast<int64_t>(p.j0) * p.ldt + p.j0] = __longlong_as_double(0x7ff8000000000000ll);
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = r_lo + warp + RWARPS * i;
    if (r < r_hi && lane_ok) Vp[static_cast<int64_t>(r) * p.ldv + lane] = v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Split-K "TN" product for tall operands: P[s] = A[ks:ke, 0:M]^T * B[ks:ke, 0:N].
// A is rows x M (lda), B is rows x N (ldb); the reduction dimension is the (long) row index.
//
// tn_dmma_kernel (the fast path): 64 x 32 output tile per CTA on the fp64 tensor pipe.  A 32-row chunk of both operands
// is staged in shared memory as it lies in memory (row = k index, padded to 68 / 36 doubles so that the m8n8k4 fragment
// loads — lane (g, c) reads element [4 ks + c][8 i + g] — touch 16 different banks per half warp); the next chunk is
// prefetched into registers while the current one feeds the DMMAs.  8 warps, warp tile 16 x 16 (2 x 2 DMMA tiles).
// tn_partial_kernel (generic): plain DFMA, any alignment.
// ------------------------------------------------------------------------------------------------
constexpr int TDM = 64, TDN = 32, TDK = 32;   // measured: 64-row chunks at 2 CTAs/SM 17 % slower, 16-row chunks at 4 CTAs/SM 5 % slower
constexpr int TD_UA = TDK * (TDM / 2) / 256, TD_UB = TDK * (TDN / 2) / 256;   // double2 per thread and chunk
constexpr int TD_LDA = TDM + 4, TD_LDB = TDN + 4;
constexpr int TD_SMEM = 2 * TDK * (TD_LDA + TD_LDB) * static_cast<int>(sizeof(double));   // 53248 B

__global__ void __launch_bounds__(256, 3) tn_dmma_kernel(double* __restrict__ P, const double* __restrict__ A, int64_t lda,
                                                         const double* __restrict__ B, int64_t ldb, int rows, int M, int N,
                                                         int kchunk, int upper_only) {
  extern __shared__ __align__(16) double td_smem[];
  const int tm = blockIdx.y * TDM, tn = blockIdx.x * TDN, sp = blockIdx.z;
  if (upper_only && tn + TDN - 1 < tm) return;              // tile entirely below the diagonal: never read
  double* As = td_smem;                                     // [2][TDK][TD_LDA]
  double* Bs = td_smem + 2 * TDK * TD_LDA;                  // [2][TDK][TD_LDB]
  const int k0 = sp * kchunk, k1 = min(rows, k0 + kchunk);
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int g = lane >> 2, c = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;                  // warp tile: rows 16 wm.., cols 16 wn..
  // staging assignment: A chunk = TDK rows x 32 double2, B chunk = TDK rows x 16 double2
  double2 ra[TD_UA], rb[TD_UB];
  auto fetch = [&](int kb) {
#pragma unroll
    for (int u = 0; u < TD_UA; ++u) {
      const int e = t + 256 * u;
      const int r = kb + (e >> 5), col = tm + 2 * (e & 31);
      double2 v = make_double2(0.0, 0.0);
      if (r < k1) {
        const double* src = A + static_cast<int64_t>(r) * lda + col;
        if (col + 1 < M) v = *reinterpret_cast<const double2*>(src);
        else if (col < M) v.x = src[0];
      }
      ra[u] = v;
    }
#pragma unroll
    for (int u = 0; u < TD_UB; ++u) {
      const int e = t + 256 * u;
      const int r = kb + (e >> 4), col = tn + 2 * (e & 15);
      double2 v = make_double2(0.0, 0.0);
      if (r < k1) {
        const double* src = B + static_cast<int64_t>(r) * ldb + col;
        if (col + 1 < N) v = *reinterpret_cast<const double2*>(src);
        else if (col < N) v.x = src[0];
      }
      rb[u] = v;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int u = 0; u < TD_UA; ++u) {
      const int e = t + 256 * u;
      *reinterpret_cast<double2*>(As + (buf * TDK + (e >> 5)) * TD_LDA + 2 * (e & 31)) = ra[u];
    }
#pragma unroll
    for (int u = 0; u < TD_UB; ++u) {
      const int e = t + 256 * u;
      *reinterpret_cast<double2*>(Bs + (buf * TDK + (e >> 4)) * TD_LDB + 2 * (e & 15)) = rb[u];
    }
  };
  double acc[2][2][2] = {};
  int buf = 0;
  if (k0 < k1) {
    fetch(k0);
    stage(0);
  }
  __syncthreads();
  for (int kb = k0; kb < k1; kb += TDK) {
    const bool more = kb + TDK < k1;
    if (more) fetch(kb + TDK);                              // global loads in flight while this chunk is multiplied
    const double* Ab = As + buf * TDK * TD_LDA + 16 * wm + g;
    const double* Bb = Bs + buf * TDK * TD_LDB + 16 * wn + g;
#pragma unroll
    for (int ks = 0; ks < TDK / 4; ++ks) {
      double a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = Ab[(4 * ks + c) * TD_LDA + 8 * i];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Bb[(4 * ks + c) * TD_LDB + 8 * j];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (more) stage(buf ^ 1);                               // the other buffer was last read before the previous barrier
    __syncthreads();
    buf ^= 1;
  }
  double* Ps = P + static_cast<int64_t>(sp) * M * N;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = tm + 16 * wm + 8 * i + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = tn + 16 * wn + 8 * j + 2 * c;
      if (col < N) Ps[static_cast<int64_t>(r) * N + col] = acc[i][j][0];
      if (col + 1 < N) Ps[static_cast<int64_t>(r) * N + col + 1] = acc[i][j][1];
    }
  }
}

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

// C (M x N, ldc) = sum of the nsplit partial products, added in a fixed order (deterministic).  With Tm != nullptr
// (N <= 32) the sum is multiplied from the right by the N x N matrix Tm (ldt) before it is stored:
// C = (sum_s P[s]) Tm — the "Wt T_p" of the compact-WY update, fused here because Wt is never needed by itself.
__global__ void __launch_bounds__(256) tn_reduce_kernel(double* __restrict__ C, int64_t ldc, const double* __restrict__ P,
                                                        int M, int N, int nsplit, const double* __restrict__ Tm, int64_t ldt) {
  const int64_t total = static_cast<int64_t>(M) * N;
  if (Tm == nullptr) {
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      double s = 0.0;
      for (int q = 0; q < nsplit; ++q) s += P[static_cast<int64_t>(q) * total + e];
      const int64_t r = e / N, c = e - r * N;
      C[r * ldc + c] = s;
    }
    return;
  }
  __shared__ double sT[QW][QW + 1];
  __shared__ double sW[8][QW + 1];
  for (int e = threadIdx.x; e < QW * QW; e += 256) {
    const int i = e / QW, c = e % QW;
    sT[i][c] = (i < N && c < N) ? Tm[static_cast<int64_t>(i) * ldt + c] : 0.0;
  }
  const int rr = threadIdx.x >> 5, c = threadIdx.x & 31;
  for (int r0 = blockIdx.x * 8; r0 < M; r0 += gridDim.x * 8) {
    const int r = r0 + rr;
    double s = 0.0;
    if (r < M && c < N) {
      const int64_t e = static_cast<int64_t>(r) * N + c;
      for (int q = 0; q < nsplit; ++q) s += P[static_cast<int64_t>(q) * total + e];
    }
    __syncthreads();                                         // sT ready (first trip) / sW free again
    sW[rr][c] = s;
    __syncthreads();
    if (r < M && c < N) {
      double o = 0.0;
#pragma unroll 8
      for (int q = 0; q < QW; ++q) o = fma(sW[rr][q], sT[q][c], o);
      C[static_cast<int64_t>(r) * ldc + c] = o;
    }
  }
}

inline bool aligned16(const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (ld % 2 == 0); }

inline int64_t tn_tiles(int64_t M, int64_t N, int upper_only) {
  const int64_t gm = (M + TDM - 1) / TDM, gn = (N + TDN - 1) / TDN;
  if (!upper_only) return gm * gn;
  int64_t cnt = 0;
  for (int64_t i = 0; i < gm; ++i)
    for (int64_t j = 0; j < gn; ++j)
      if (!(j * TDN + TDN - 1 < i * TDM)) ++cnt;
  return cnt;
}

inline int tn_nsplit(int64_t rows, int64_t M, int64_t N, int upper_only = 0) {
  const int64_t tiles = tn_tiles(M, N, upper_only);
  int64_t want = (4 * 148 + tiles - 1) / tiles;           // ~4 CTAs per SM in total
  const int64_t maxs = (rows + 255) / 256;                // at least 256 rows per split
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return static_cast<int>(want);
}

bool g_tn_attr[64] = {};

// C (M x N, ldc) = A^T B [Tm] ; partials must hold nsplit*M*N doubles
int launch_tn(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t rows, int64_t M,
              int64_t N, double* partials, int upper_only, const double* Tm, int64_t ldt, cudaStream_t st) {
  if (M <= 0 || N <= 0) return NPW_OK;
  const int ns = tn_nsplit(rows, M, N, upper_only);
  const int kchunk = static_cast<int>((((rows + ns - 1) / ns) + TDK - 1) / TDK * TDK);
  if (aligned16(A, lda) && aligned16(B, ldb)) {
    int dev = 0;
    NPW_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 64 && !g_tn_attr[dev]) {
      NPW_CUDA_CHECK(cudaFuncSetAttribute(tn_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM));
      g_tn_attr[dev] = true;
    }
    dim3 grid(static_cast<unsigned>((N + TDN - 1) / TDN), static_cast<unsigned>((M + TDM - 1) / TDM), static_cast<unsigned>(ns));
    tn_dmma_kernel<<<grid, 256, TD_SMEM, st>>>(partials, A, lda, B, ldb, static_cast<int>(rows), static_cast<int>(M),
                                               static_cast<int>(N), kchunk, upper_only);
  } else {
    dim3 grid(static_cast<unsigned>((N + TT - 1) / TT), static_cast<unsigned>((M + TT - 1) / TT), static_cast<unsigned>(ns));
    tn_partial_kernel<<<grid, 256, 0, st>>>(partials, A, lda, B, ldb, static_cast<int>(rows), static_cast<int>(M),
                                            static_cast<int>(N), kchunk);
  }
  NPW_LAUNCH_CHECK();
  const int64_t total = M * N;
  int rb = Tm ? static_cast<int>((M + 7) / 8) : static_cast<int>((total + 255) / 256);
  if (rb > 148 * 8) rb = 148 * 8;
  tn_reduce_kernel<<<rb, 256, 0, st>>>(C, ldc, partials, static_cast<int>(M), static_cast<int>(N), ns, Tm, ldt);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

// ------------------------------------------------------------------------------------------------
// Streaming rank-k update (k <= 32) of the trailing columns:  C[r][c] -= sum_q V[r][q] W[c][q]
// (C rows x nt, V rows x k = the panel's reflectors, W nt x k = Wt T_p).  64 x 64 tile per CTA, 8 warps of 32 x 16,
// operands staged k-contiguous with rows padded to 36 doubles (conflict-free fragment loads), DMMA.  Every element of C
// is read and written exactly once and gets 2 x 32 flop: HBM-bound (the tensor pipe would sustain ~1.4x the traffic).
// The C fragments are requested before the operands are staged so that the loads overlap the staging and the math.
// ------------------------------------------------------------------------------------------------
constexpr int RU = 64, RU_N = 32, RU_LD = QW + 4;        // 64-row stripe, 32-column tiles
constexpr int RU_SMEM = (RU + 2 * RU_N) * RU_LD * static_cast<int>(sizeof(double));   // V tile + double-buffered W tile = 36 KB

// A CTA owns a 64-row stripe of C and walks over `tiles_per_cta` column tiles of 32: the V tile is staged once, the W
// tile of the next column tile and the next C fragments are requested before the current tile is multiplied, so the
// HBM stream never waits for the math.  The kernel is memory-bound (the DMMAs of a tile take a third of the time its
// 32 KB of C traffic needs at HBM speed), so it is built for bytes in flight: small register tiles (warp tile 16 x 16),
// ~80 registers, three CTAs per SM, each with the loads of two tiles outstanding.
__global__ void __launch_bounds__(256, 3) rank_update_kernel(double* __restrict__ C, int64_t ldc, const double* __restrict__ V,
                                                             int64_t ldv, const double* __restrict__ W, int64_t ldw, int rows, int nt,
                                                             int k, int vec, int vecv, int tiles_per_cta,
                                                             double* __restrict__ Rrows, int64_t ldr, int rrows) {
  extern __shared__ __align__(16) double ru_smem[];
  double* Vs = ru_smem;                                     // [RU][RU_LD]
  double* Ws = ru_smem + RU * RU_LD;                        // [2][RU_N][RU_LD]
  const int r0 = blockIdx.x * RU;
  const int nct = (nt + RU_N - 1) / RU_N;
  const int ct0 = blockIdx.y * tiles_per_cta;
  const int ct1 = min(nct, ct0 + tiles_per_cta);
  if (ct0 >= ct1) return;
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int g = lane >> 2, c = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;                  // warp tile: rows 16 wm.., cols 16 wn..

  // this lane's C fragments of column tile ct: rows r0 + 16 wm + 8 i + g, cols 32 ct + 16 wn + 8 j + 2 c + {0, 1}
  auto load_c = [&](int ct, double2 (&cf)[2][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + 16 * wm + 8 * i + g;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = ct * RU_N + 16 * wn + 8 * j + 2 * c;
        double2 v = make_double2(0.0, 0.0);
        if (r < rows) {
          const double* src = C + static_cast<int64_t>(r) * ldc + col;
          if (vec && col + 1 < nt) v = *reinterpret_cast<const double2*>(src);
          else {
            if (col < nt) v.x = src[0];
            if (col + 1 < nt) v.y = src[1];
          }
        }
        cf[i][j] = v;
      }
    }
  };
  // an NR x 32 operand tile (row stride ld) as NR / 16 double2 per thread
  auto fetch_one = [&](const double* M, int64_t ld, int row0, int nrows, int aligned, int e) -> double2 {
    const int rr = row0 + (e >> 4), q = 2 * (e & 15);
    double2 v = make_double2(0.0, 0.0);
    if (rr < nrows) {
      const double* src = M + static_cast<int64_t>(rr) * ld + q;
      if (aligned && q + 1 < k) v = *reinterpret_cast<const double2*>(src);
      else {
        if (q < k) v.x = src[0];
        if (q + 1 < k) v.y = src[1];
      }
    }
    return v;
  };
  auto stage_one = [&](double* dst, int e, double2 v) {
    *reinterpret_cast<double2*>(dst + (e >> 4) * RU_LD + 2 * (e & 15)) = v;
  };

  double2 cf[2][2], cfn[2][2], wr[2];
  load_c(ct0, cf);
#pragma unroll
  for (int u = 0; u < 4; ++u) stage_one(Vs, t + 256 * u, fetch_one(V, ldv, r0, rows, vecv, t + 256 * u));
#pragma unroll
  for (int u = 0; u < 2; ++u) stage_one(Ws, t + 256 * u, fetch_one(W, ldw, ct0 * RU_N, nt, 1, t + 256 * u));
  __syncthreads();
  int buf = 0;
  for (int ct = ct0; ct < ct1; ++ct) {
    const bool more = ct + 1 < ct1;
    if (more) {
#pragma unroll
      for (int u = 0; u < 2; ++u) wr[u] = fetch_one(W, ldw, (ct + 1) * RU_N, nt, 1, t + 256 * u);
      load_c(ct + 1, cfn);
    }
    const double* Wb = Ws + buf * RU_N * RU_LD;
    double acc[2][2][2] = {};
#pragma unroll
    for (int ks = 0; ks < QW / 4; ++ks) {
      double a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = Vs[(16 * wm + 8 * i + g) * RU_LD + 4 * ks + c];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Wb[(16 * wn + 8 * j + g) * RU_LD + 4 * ks + c];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + 16 * wm + 8 * i + g;
      if (r >= rows) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = ct * RU_N + 16 * wn + 8 * j + 2 * c;
        double* dst = C + static_cast<int64_t>(r) * ldc + col;
        const double2 o = make_double2(cf[i][j].x - acc[i][j][0], cf[i][j].y - acc[i][j][1]);
        if (vec && col + 1 < nt) *reinterpret_cast<double2*>(dst) = o;
        else {
          if (col < nt) dst[0] = o.x;
          if (col + 1 < nt) dst[1] = o.y;
        }
        if (r < rrows) {                                   // the first `rrows` rows of the updated block are final rows of R
          double* rd = Rrows + static_cast<int64_t>(r) * ldr + col;
          if (col < nt) rd[0] = o.x;
          if (col + 1 < nt) rd[1] = o.y;
        }
      }
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < 2; ++u) stage_one(Ws + (buf ^ 1) * RU_N * RU_LD, t + 256 * u, wr[u]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) cf[i][j] = cfn[i][j];
    }
    __syncthreads();
    buf ^= 1;
  }
}

bool g_ru_attr[64] = {};

int launch_rank_update(double* C, int64_t ldc, const double* V, int64_t ldv, const double* W, int64_t ldw, int64_t rows,
                       int64_t nt, int k, double* Rrows, int64_t ldr, int rrows, cudaStream_t st) {
  if (rows <= 0 || nt <= 0 || k <= 0) return NPW_OK;
  if (k > QW || !aligned16(W, ldw)) {
    set_error("rank update: k = %d > 32 or unaligned W", k);
    return NPW_ERR_UNSUPPORTED;
  }
  int dev = 0;
  NPW_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && !g_ru_attr[dev]) {
    NPW_CUDA_CHECK(cudaFuncSetAttribute(rank_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM));
    g_ru_attr[dev] = true;
  }
  const int64_t stripes = (rows + RU - 1) / RU, nct = (nt + RU_N - 1) / RU_N;
  int64_t groups = (24 * 148 + stripes - 1) / stripes;      // enough CTAs for ~8 waves at 3 per SM
  if (groups > nct) groups = nct;
  if (groups < 1) groups = 1;
  const int tiles_per_cta = static_cast<int>((nct + groups - 1) / groups);
  dim3 grid(static_cast<unsigned>(stripes), static_cast<unsigned>((nct + tiles_per_cta - 1) / tiles_per_cta));
  rank_update_kernel<<<grid, 256, RU_SMEM, st>>>(C, ldc, V, ldv, W, ldw, static_cast<int>(rows), static_cast<int>(nt), k,
                                                 aligned16(C, ldc) ? 1 : 0, aligned16(V, ldv) ? 1 : 0, tiles_per_cta, Rrows, ldr,
                                                 Rrows ? rrows : 0);
  NPW_LAUNCH_CHECK();
  return NPW_OK;
}

// ------------------------------------------------------------------------------------------------
// Off-diagonal block column of T for the panel at j0 (w columns), dlarft's recurrence in block form:
//   T[0:j0, j0:j0+w] = -T[0:j0, 0:j0] (G[0:j0, j0:j0+w] T_pp),   G = V^T V (upper part), T_pp = T[j0:j0+w, j0:j0+w].
// CTA rb computes output rows [32 rb, 32 rb + 32): T is upper triangular, so only k-blocks kb >= rb contribute.
// Four groups of 256 threads take the k-blocks kb = rb + grp, rb + grp + 4, ...; their sums are added in a fixed order.
// ------------------------------------------------------------------------------------------------
constexpr int TO_NG = 4;
constexpr int TO_LD = QW + 4;                               // 36: conflict-free m8n8k4 fragment loads
constexpr int TO_BLK = QW * TO_LD;
constexpr int TO_SMEM = (1 + 3 * TO_NG) * TO_BLK * static_cast<int>(sizeof(double));

// The 32 x 32 x 32 block products run on the fp64 tensor pipe: warp wi of a group owns the 8 x 16 output strip
// (tile row wi / 2, tile columns 2 (wi % 2) and 2 (wi % 2) + 1) — 3 fragment loads per 2 DMMAs instead of 2 shared
// loads per DFMA (round 2a: 16 us per trip, shared-memory bound).
__global__ void __launch_bounds__(256 * TO_NG) t_offdiag_kernel(double* __restrict__ T, int64_t ldt, const double* __restrict__ Gm,
                                                                int64_t ldg, int j0, int w) {
  extern __shared__ __align__(16) double to_smem[];
  double* sTpp = to_smem;                                   // [32][36]
  const int grp = threadIdx.x >> 8;
  const int t = threadIdx.x & 255;
  double* sG = to_smem + TO_BLK * (1 + 3 * grp);            // G block, later this group's partial result
  double* sX = sG + TO_BLK;                                 // G_blk T_pp
  double* sTl = sX + TO_BLK;                                // T[32 rb.., 32 kb..]
  const int rb = blockIdx.x;
  const int wi = t >> 5, lane = t & 31;
  const int g = lane >> 2, c = lane & 3;
  const int tr = wi >> 1, tc0 = (wi & 1) * 2;
  for (int e = threadIdx.x; e < QW * QW; e += 256 * TO_NG) {
    const int i = e >> 5, cc = e & 31;
    sTpp[i * TO_LD + cc] = (i < w && cc < w) ? T[static_cast<int64_t>(j0 + i) * ldt + j0 + cc] : 0.0;
  }
  double acc[2][2] = {};
  const int nkb = (j0 + QW - 1) / QW;
  const int trips = (nkb - rb + TO_NG - 1) / TO_NG;         // same trip count for every group (barriers inside)
  for (int it = 0; it < trips; ++it) {
    const int kb = rb + grp + it * TO_NG;
    __syncthreads();
    for (int e = t; e < QW * QW; e += 256) {
      const int i = e >> 5, cc = e & 31;
      const int gr = kb * QW + i;
      sG[i * TO_LD + cc] = (kb < nkb && gr < j0 && cc < w) ? Gm[static_cast<int64_t>(gr) * ldg + j0 + cc] : 0.0;
      const int trw = rb * QW + i, tcl = kb * QW + cc;
      sTl[i * TO_LD + cc] = (kb < nkb && trw < j0 && tcl < j0) ? T[static_cast<int64_t>(trw) * ldt + tcl] : 0.0;
    }
    __syncthreads();
    double x[2][2] = {};
#pragma unroll
    for (int ks = 0; ks < QW / 4; ++ks) {
      const double a = sG[(8 * tr + g) * TO_LD + 4 * ks + c];
#pragma unroll
      for (int u = 0; u < 2; ++u) dmma884(x[u][0], x[u][1], a, sTpp[(4 * ks + c) * TO_LD + 8 * (tc0 + u) + g]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      sX[(8 * tr + g) * TO_LD + 8 * (tc0 + u) + 2 * c] = x[u][0];
      sX[(8 * tr + g) * TO_LD + 8 * (tc0 + u) + 2 * c + 1] = x[u][1];
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < QW / 4; ++ks) {
      const double a = sTl[(8 * tr + g) * TO_LD + 4 * ks + c];
#pragma unroll
      for (int u = 0; u < 2; ++u) dmma884(acc[u][0], acc[u][1], a, sX[(4 * ks + c) * TO_LD + 8 * (tc0 + u) + g]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    sG[(8 * tr + g) * TO_LD + 8 * (tc0 + u) + 2 * c] = acc[u][0];
    sG[(8 * tr + g) * TO_LD + 8 * (tc0 + u) + 2 * c + 1] = acc[u][1];
  }
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 8 * tr + g, cc = 8 * (tc0 + u) + 2 * c + e;
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < TO_NG; ++q) sum += to_smem[TO_BLK * (1 + 3 * q) + i * TO_LD + cc];
        const int r = rb * QW + i;
        if (r < j0 && cc < w) T[static_cast<int64_t>(r) * ldt + j0 + cc] = -sum;
      }
  }
}

bool g_to_attr[64] = {};

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct QrWork {
  size_t scratch, tau, tmp, gram, partials, total;
};

QrWork qr_work_layout(int64_t m, int64_t n) {
  QrWork w;
  size_t off = 0;
  w.scratch = off; off += align256(PK_BYTES + 256 + MAX_GRID * 2 * 8 * sizeof(long long));   // packets + error word + (profile builds) phase counters
  w.tau = off; off += align256(static_cast<size_t>(n) * sizeof(double));
  w.tmp = off; off += align256(static_cast<size_t>(n) * QW * sizeof(double));        // Wt T_p
  w.gram = off; off += align256(static_cast<size_t>(n) * n * sizeof(double));        // G = V^T V
  const int ns_g = tn_nsplit(m, n, n, 1);
  size_t pb = static_cast<size_t>(ns_g) * n * n;
  // nsplit depends on the row count of each call; bound it by the largest value tn_nsplit can return for these widths
  const size_t pmax = static_cast<size_t>(256) * n * QW;
  if (pmax > pb) pb = pmax;
  w.partials = off; off += align256(pb * sizeof(double));
  w.total = off;
  return w;
}

int g_coop_grid[64] = {};
bool g_qr_no_reg = false;

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
  uint64_t* packets = reinterpret_cast<uint64_t*>(wb + wl.scratch);
  int* err = reinterpret_cast<int*>(wb + wl.scratch + PK_BYTES);
  double* tau = reinterpret_cast<double*>(wb + wl.tau);
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
    NPW_CUDA_CHECK(cudaFuncSetAttribute(qr_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        QR_SMEM_ROWS * QW * static_cast<int>(sizeof(double))));
    NPW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qr_panel_kernel<true>, QTHREADS,
                                                               QR_SMEM_ROWS * QW * sizeof(double)));
    int g = sms * (per_sm > 0 ? 1 : 0);
    if (g > MAX_GRID) g = MAX_GRID;
    if (const char* e = getenv("NPW_B200_QR_NO_REG")) g_qr_no_reg = atoi(e) != 0;   // tuning / test knob: shared-memory variant
    if (const char* e = getenv("NPW_B200_QR_GRID")) {       // tuning knob: CTAs of a panel launch
      const int v = atoi(e);
      if (v >= 1 && v < g) g = v;
    }
    if (g < 1) {
      set_error("qr_panel_kernel does not fit on an SM");
      return NPW_ERR_UNSUPPORTED;
    }
    g_coop_grid[dev] = g;
  }
  const int max_grid = dev < 64 ? g_coop_grid[dev] : 1;

  int rc;
  // the all-gather packets validate themselves by sequence number: start every factorisation from zeroed packets
  // (sequence numbers are unique within a factorisation and never 0)
  NPW_CUDA_CHECK(cudaMemsetAsync(packets, 0, PK_BYTES + 256 + MAX_GRID * 2 * 8 * sizeof(long long), st));
  if (V != A) {
    rc = launch_copy2d(V, ldv, A, lda, m, n, 0, st);
    if (rc) return rc;
  }
  rc = launch_fill2d(R, ldr, n, n, 0, 0.0, st);
  if (rc) return rc;
  rc = launch_fill2d(T, ldt, n, n, 0, 0.0, st);
  if (rc) return rc;

  uint32_t seq = 1;
  for (int64_t j0 = 0; j0 < n; j0 += QW) {
    const int w = static_cast<int>(n - j0 < QW ? n - j0 : QW);
    const int64_t rows = m - j0;
    PanelArgs pa;
    pa.V = V; pa.ldv = ldv; pa.m = static_cast<int>(m); pa.j0 = static_cast<int>(j0); pa.w = w;
    pa.R = R; pa.ldr = ldr; pa.T = T; pa.ldt = ldt; pa.tau = tau; pa.packets = packets; pa.seq0 = seq; pa.err = err;
    pa.prof = reinterpret_cast<long long*>(wb + wl.scratch + PK_BYTES + 256);
    seq += QW + 2;                                          // steps 0 .. w of this launch
    int g = static_cast<int>((rows + 127) / 128);           // >= 128 panel rows per CTA
    if (g > max_grid) g = max_grid;
    if (g < 1) g = 1;
    void* kargs[] = {&pa};
    const int64_t per = (rows + g - 1) / g;                 // rows per CTA, as the kernel computes it
    if (per <= RWARPS * RPW && !g_qr_no_reg) {
      NPW_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(qr_panel_reg_kernel), dim3(g), dim3(RTHREADS), kargs, 0, st));
    } else if (per <= QR_SMEM_ROWS) {
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
      // tmp = (C^T V_p) T_p  (nt x w): applying Q^T = I - V T^T V^T is  C -= V (T^T (V^T C))  <=>  C -= V ((C^T V) T)^T
      rc = launch_tn(tmp, QW, C, ldv, Vp, ldv, rows, nt, w, partials, 0, T + j0 * ldt + j0, ldt, st);
      if (rc) return rc;
      // C -= V_p tmp^T
      // ... and rows j0 .. j0+w-1 of the updated trailing columns are final rows of R (written by the same kernel)
      rc = launch_rank_update(C, ldv, Vp, ldv, tmp, QW, rows, nt, w, R + j0 * ldr + j0 + w, ldr, w, st);
      if (rc) return rc;
    }
  }
  // explicit V: zero strictly above the diagonal (the panel kernel already wrote the unit diagonal)
  rc = launch_fill2d(V, ldv, n, n, 2, 0.0, st);
  if (rc) return rc;
  // off-diagonal blocks of T from the Gram matrix of V (upper tiles)
  if (n > QW) {
    rc = launch_tn(gram, n, V, ldv, V, ldv, m, n, n, partials, 1, nullptr, 0, st);
    if (rc) return rc;
    for (int64_t j0 = QW; j0 < n; j0 += QW) {
      const int w = static_cast<int>(n - j0 < QW ? n - j0 : QW);
      if (dev < 64 && !g_to_attr[dev]) {
        NPW_CUDA_CHECK(cudaFuncSetAttribute(t_offdiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TO_SMEM));
        g_to_attr[dev] = true;
      }
      t_offdiag_kernel<<<static_cast<unsigned>((j0 + QW - 1) / QW), 256 * TO_NG, TO_SMEM, st>>>(T, ldt, gram, n, static_cast<int>(j0), w);
      NPW_LAUNCH_CHECK();
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
