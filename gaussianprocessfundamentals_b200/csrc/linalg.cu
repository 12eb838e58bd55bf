// Dense FP64 linear algebra of the exact-GP likelihood path, batched over independent GPs.
//
//   run_potrf    blocked right-looking Cholesky of K + s2*I (Statistics/CovarianceMatrix.py:247-254, :469-479) with the
//                targets y carried as an extra row, so that L^-1 y (:256-265) and y^T K^-1 y
//                (Metrics/LogLikelihood.py:39) fall out of the factorisation, and sum(log diag L)
//                (Metrics/Metrics.py:152-154) is reduced by the diagonal-block kernel.
//   run_trtri    W = L^-1 by recursive doubling (pure GEMM), run_lauum  K^-1 = W^T W, run_alpha  alpha = W^T z:
//                the pieces of dNLL/dK = 1/2 (K^-1 - alpha alpha^T) that TF's Cholesky gradient produces for
//                Optimizer/Fitter.py:124-158.
//   run_trsv     standalone forward/back substitution (CovarianceMatrix.py:260-262) for get_L_alpha().
#include <algorithm>
#include <type_traits>
#include <vector>
#include <cstdlib>
#include "gemm.cuh"
#include "internal.h"
#include "trace.h"

namespace gpb {

constexpr int GPB_KB_MAX = 8;   // widest outer panel of the factorisation in 128-blocks (GPB_POTRF_KB)

// ---------------------------------------------------------------------------------------------------------------
// GEMM geometries (tile<BM, BN>() fills the job of the calling CTA; BN is always 128 = GPB_NB)
// ---------------------------------------------------------------------------------------------------------------

// trailing update with the panel P = A[r0:, kp*128 : (kp+kb)*128] (kb = 1 or 2 finished 128-blocks):
// A[r0:, r0:] -= P P^T for r0 = (kp+kb)*128, lower tiles, tile columns restricted to [c_lo, c_hi) so that the look-ahead
// driver can split the update across streams.  kb = 2 halves the number of passes over the far trailing matrix and the
// per-tile epilogue cost per flop (k = 256: 31 TFLOP/s against 28 at k = 128).
// sep != 0: the carried row of a matrix whose n is a multiple of 128 is NOT part of the tiles (it would open a tile row of
// its own - 1 useful row in 64 - which costs 14 .. 50 % of the update of the small trailing matrices of a batch);
// carried_row_kernel applies the same product to it.
__device__ __forceinline__ bool carried_row_separate(const GpbMat& d, int sep) {
  return sep && d.aug && (d.n % GPB_NB) == 0;
}

struct GeoSyrk {
  const GpbMat* mats;
  int kp, kb, c_lo, c_hi, sep;
  int grp = 0;   // bulk launches that reach the last tile column: super-tile raster (groups of 8 tile columns)
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    const GpbMat& d = mats[b.z];
    const int nrows = d.n + (carried_row_separate(d, sep) ? 0 : d.aug);
    const int r0 = (kp + kb) * GPB_NB;
    if (r0 >= nrows || r0 > d.n) return false;   // every block of the panel must be a full pivot block
    const int Tm = (nrows - r0 + BM - 1) / BM;
    int ti, tj;
    constexpr int R = BN / BM;
    if (grp && c_hi >= (Tm + R - 1) / R) {
      // The tile columns [c_lo, end) form a lower triangle of their own (rows shifted by R c_lo).  In column-sweep order
      // a wave of 592 tiles streams 590 different A slabs (k = 1024: 0.5 MB each, more than the 126 MB L2) and the next
      // column re-reads them from DRAM: ncu at n = 32768, first bulk update: 31 GB per launch, L2 hit rate 62 %
      // (profiles/r2e_ncu_syrk_bulk_k1024_n32768.txt).  Groups of 8 columns x ~74 row tiles keep a wave's operands in L2.
      if (!tri_map_grouped(b.x, Tm - R * c_lo, R, 8, ti, tj)) return false;
      ti += R * c_lo; tj += c_lo;
    } else if (!tri_map(b.x, Tm, R, c_lo, c_hi, ti, tj)) return false;
    const size_t ld = d.ld;
    const double* P = d.A + (size_t)kp * GPB_NB * ld;
    J.A = P + r0 + ti * BM;
    J.B = P + r0 + tj * BN;
    J.C = d.A + (r0 + ti * BM) + (size_t)(r0 + tj * BN) * ld;
    J.lda = J.ldb = J.ldc = d.ld;
    J.mrem = min(BM, nrows - r0 - ti * BM);
    J.nrem = min(BN, nrows - r0 - tj * BN);
    J.klo = 0; J.khi = kb * GPB_NB;
    J.alpha = -1.0; J.beta = 1.0; J.red = g_red_epilogue;
    return true;
  }
};

// panel solve of step k as a product with the inverted diagonal block: A[r0:, kblk] = A[r0:, kblk] * Wd_k^T
// (in place: a CTA owns all 128 columns of its rows, so BN must be 128)
struct GeoPanel {
  const GpbMat* mats;
  int k, sep;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    static_assert(BN == GPB_NB, "the in-place panel product needs full-width tiles");
    const GpbMat& d = mats[b.z];
    const int nrows = d.n + (carried_row_separate(d, sep) ? 0 : d.aug);
    const int r0 = (k + 1) * GPB_NB;
    if (r0 > d.n) return false;  // block k is not a full pivot block
    const int i0 = r0 + b.x * BM;
    if (i0 >= nrows) return false;
    const size_t ld = d.ld;
    double* P = d.A + (size_t)k * GPB_NB * ld + i0;
    J.A = P; J.C = P;
    J.B = d.Wd + (size_t)k * GPB_NB * GPB_NB;
    J.lda = J.ldc = d.ld; J.ldb = GPB_NB;
    J.mrem = min(BM, nrows - i0);
    J.nrem = GPB_NB;
    J.klo = 0; J.khi = GPB_NB;
    J.alpha = 1.0; J.beta = 0.0; J.red = 0;
    return true;
  }
};

// triangular inverse by recursive doubling, level s: sub-problem p owns the 2s x 2s diagonal block at r0 = 2 s p
//   [W11 0; L21 W22]  ->  W21 = -W22 * (L21 * W11)
// phase T:  T = L21 * W11   (NN, k >= tile column start because W11 is lower triangular)   -> Kinv scratch
struct GeoTrtriT {
  const GpbMat* mats;
  int s, p0;   // sub-problems p0, p0 + 1, ... (grid.y of them)
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    const GpbMat& d = mats[b.z];
    const int r0 = 2 * s * ((int)b.y + p0), rA = r0 + s;
    if (rA >= d.n || *d.info != 0) return false;   // failed factorisation: the inverse stages do no work (grad = NaN)
    const int M = min(s, d.n - rA);
    // column tile slowest and ascending: the k-range (tj*BN .. s) shrinks with tj, so the longest tiles start first and
    // the launch drains with the short ones
    const int tsm = s / BM;
    const int ti = b.x % tsm, tj = b.x / tsm;
    if (ti * BM >= M) return false;
    const size_t ld = d.ld;
    J.A = d.A + (rA + ti * BM) + (size_t)r0 * ld;
    J.B = d.A + r0 + (size_t)(r0 + tj * BN) * ld;
    J.C = d.Kinv + (rA + ti * BM) + (size_t)(r0 + tj * BN) * ld;
    J.lda = J.ldb = J.ldc = d.ld;
    J.mrem = min(BM, M - ti * BM);
    J.nrem = BN;
    J.klo = tj * BN; J.khi = s;
    J.alpha = 1.0; J.beta = 0.0; J.red = 0;
    return true;
  }
};
// phase W:  W21 = -W22 * T   (NN, k <= tile row end because W22 is lower triangular)
struct GeoTrtriW {
  const GpbMat* mats;
  int s, p0;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    const GpbMat& d = mats[b.z];
    const int r0 = 2 * s * ((int)b.y + p0), rA = r0 + s;
    if (rA >= d.n || *d.info != 0) return false;
    const int M = min(s, d.n - rA);
    // row tile slowest and DEscending: the k-range (0 .. (ti+1)*BM) grows with ti, longest tiles first
    const int tsn = s / BN, tsm = s / BM;
    const int ti = tsm - 1 - (int)(b.x / tsn), tj = b.x % tsn;
    if (ti * BM >= M) return false;
    const size_t ld = d.ld;
    J.A = d.A + (rA + ti * BM) + (size_t)rA * ld;
    J.B = d.Kinv + rA + (size_t)(r0 + tj * BN) * ld;
    J.C = d.A + (rA + ti * BM) + (size_t)(r0 + tj * BN) * ld;
    J.lda = J.ldb = J.ldc = d.ld;
    J.mrem = min(BM, M - ti * BM);
    J.nrem = BN;
    J.klo = 0; J.khi = min(M, (ti + 1) * BM);
    J.alpha = -1.0; J.beta = 0.0; J.red = 0;
    return true;
  }
};

// inv(K) = W^T W, lower tiles only, k >= tile row start (W lower triangular), out of place into Kinv
struct GeoLauum {
  const GpbMat* mats;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    const GpbMat& d = mats[b.z];
    if (*d.info != 0) return false;
    const int Tm = (d.n + BM - 1) / BM;
    int ti, tj;
    if (!tri_map_grouped(b.x, Tm, BN / BM, 8, ti, tj)) return false;   // super-tile raster: W exceeds L2
    const size_t ld = d.ld;
    J.A = d.A + (size_t)(ti * BM) * ld;
    J.B = d.A + (size_t)(tj * BN) * ld;
    J.C = d.Kinv + ti * BM + (size_t)(tj * BN) * ld;
    J.lda = J.ldb = J.ldc = d.ld;
    J.mrem = min(BM, d.n - ti * BM);
    J.nrem = min(BN, d.n - tj * BN);
    J.klo = ti * BM; J.khi = d.n;
    J.alpha = 1.0; J.beta = 0.0; J.red = 0;
    return true;
  }
};

struct GeoPlain {
  const double* A; const double* B; double* C;
  int lda, ldb, ldc, M, N, K, akm, bkm;
  double alpha, beta;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    const int i0 = b.x * BM, j0 = b.y * BN;
    J.A = akm ? A + (size_t)i0 * lda : A + i0;
    J.B = bkm ? B + (size_t)j0 * ldb : B + j0;
    J.C = C + i0 + (size_t)j0 * ldc;
    J.lda = lda; J.ldb = ldb; J.ldc = ldc;
    J.mrem = min(BM, M - i0); J.nrem = min(BN, N - j0);
    J.klo = 0; J.khi = K; J.alpha = alpha; J.beta = beta; J.red = (beta == 1.0) ? g_red_epilogue : 0;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// diagonal block: Cholesky of a 128 x 128 block in shared memory (+ rows of the carried y^T that fall inside the
// block), log-det partial, then the block's triangular inverse for the panel product.
// ---------------------------------------------------------------------------------------------------------------
// The block and its inverse are computed TOGETHER by eliminating the augmented matrix [A; I] (256 x 128): the rows of
// the identity are "carried rows" exactly like y^T in the big factorisation, and end up as I * L^-T = inv(L)^T.  The
// elimination runs in 16 micro-panels of 8 columns:
//   (1) panel: thread R (0..255) owns row R of the 8 panel columns.  The 8 x 8 diagonal micro-block is factorised
//       redundantly by every warp (lane j < 8 holds its row j; pivots and multipliers travel by shuffle, so the pivot
//       chain is shfl -> rsqrt -> mul -> shfl -> fma with no barrier and no shared memory), and every thread applies
//       the same eliminations to its own row.  The finished panel goes to shared memory (DMMA operand layout), the
//       L part also to global memory, the inv(L)^T part to a transposing buffer.
//   (2) update: the trailing matrix lives in REGISTERS as DMMA accumulator fragments - 136 lower 8 x 8 tiles of A and
//       136 upper tiles of X = inv(L)^T, 34 per warp, sorted by column so that the live ones form a suffix - and
//       receives the rank-8 update  T(ti, tj) -= P(ti) P(tj)^T  as two DMMA.8x8x4 per tile.  The tiles of the next panel
//       column are then published to shared memory in row-per-thread order.
// Rows of the carried right-hand side that fall inside the block are ordinary non-pivot rows; a non-pivot column (the
// element (n, n) behind y) is eliminated but never used as a pivot.
constexpr int D_P = 264;                 // pitch of a panel buffer [8][D_P]: 256 rows per panel column
constexpr int D_WLD = GPB_NB + 1;        // pitch of the transposing buffer of inv(L)
// register tiles per warp: 272 tiles / NW warps (NW = 8: 34 tiles, 256 threads; NW = 16: 17 tiles, 512 threads)
constexpr int D_SMEM_BYTES = (2 * 8 * D_P + GPB_NB * D_WLD + GPB_NB) * (int)sizeof(double);

// tile L = NW * slot + warp of the column-sorted list: column tj holds 16 - tj lower tiles of A, then tj + 1 upper tiles of X
template <int NW>
__device__ __forceinline__ void diag_tile_of(int slot, int warp, int& ti, int& tj, bool& is_a) {
  // slot is a compile-time constant at every call site: NW * slot / 17 and the warp at which the column index steps
  // fold to immediates, so the map costs a compare and two adds
  const int base = NW * slot, tj0 = base / 17, thr = 17 * (tj0 + 1) - base;
  const int L = base + warp;
  tj = tj0 + ((warp >= thr) ? 1 : 0);
  const int t = L - 17 * tj;
  is_a = t < 16 - tj;
  ti = is_a ? tj + t : t - (16 - tj);
}

template <int NW>
__global__ void __launch_bounds__(32 * NW, 1) diag_kernel_t(const GpbMat* __restrict__ mats, int k) {
  constexpr int D_SLOTS = 272 / NW;
  constexpr int NT = 32 * NW;
  extern __shared__ __align__(16) double dsm[];
  double* Xs = dsm;                  // [8][D_P]  next panel column, row-per-thread order
  double* Ps = Xs + 8 * D_P;         // [8][D_P]  finished panel (DMMA operand)
  double* Ws = Ps + 8 * D_P;         // [128][D_WLD]  W[a][b] = inv(L)[a][b]
  double* piv = Ws + GPB_NB * D_WLD; // [128] pivots L_jj (their logs are taken after the elimination)
  __shared__ int s_info;
  const GpbMat d = mats[blockIdx.x];
  const int nrows = d.n + d.aug;
  const int r0 = k * GPB_NB;
  if (r0 >= d.n) return;
  const int bs = min(GPB_NB, nrows - r0);  // rows / columns held by the block (pivots + carried)
  const int bf = min(GPB_NB, d.n - r0);    // pivots
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const size_t ld = d.ld;
  double* Ag = d.A + r0 + (size_t)r0 * ld;
  if (tid == 0) s_info = 0;

  // ---- the augmented matrix as accumulator tiles ------------------------------------------------------------------
  double acc[D_SLOTS][2];
#pragma unroll
  for (int s = 0; s < D_SLOTS; ++s) {
    int ti, tj; bool is_a;
    diag_tile_of<NW>(s, warp, ti, tj, is_a);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int row = ti * 8 + lr, col = tj * 8 + 2 * lk + e;
      double v;
      if (is_a) v = (row < bs && col < bs && row >= col) ? Ag[row + (size_t)col * ld] : 0.0;
      else v = (row == col) ? 1.0 : 0.0;
      acc[s][e] = v;
    }
  }
  // publish panel column 0
#pragma unroll
  for (int s = 0; s < D_SLOTS; ++s) {
    int ti, tj; bool is_a;
    diag_tile_of<NW>(s, warp, ti, tj, is_a);
    if (tj == 0) {
      const int rb = (is_a ? 0 : GPB_NB) + ti * 8 + lr;
      Xs[(2 * lk) * D_P + rb] = acc[s][0];
      Xs[(2 * lk + 1) * D_P + rb] = acc[s][1];
    }
  }
  __syncthreads();

  const int R = tid;                       // row of [A; I] owned in the panel phase
  const bool a_row = R < GPB_NB;
  const int xi = R - GPB_NB;               // row of X (column of inv(L)) for the lower half
  double lsum = 0.0;
  const int nsteps = (bs + 7) / 8;
  for (int m = 0; m < nsteps; ++m) {
    // ---- (1) panel ------------------------------------------------------------------------------------------------
    const int c0 = 8 * m;
    if (NW == 8 || warp < 8) {
    const bool active = a_row ? (R >= c0) : (xi < c0 + 8);
    double a[8], Dr[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      a[c] = active ? Xs[c * D_P + R] : 0.0;
      Dr[c] = (lane < 8) ? Xs[c * D_P + c0 + lane] : 0.0;
    }
    const int npiv = min(8, max(0, bf - c0));
    // Software-pipelined pivot chain: as soon as column c+1 has received the update of pivot c, its pivot is broadcast
    // and the reciprocal square root of the NEXT pivot is started, so that its latency runs under the updates of the
    // columns c+2.. .  Branch-free (selects) so that the compiler keeps this order: rsqrt of a non-pivot is discarded.
    double dv = __shfl_sync(0xffffffffu, Dr[0], 0);
    double inv = (0 < npiv) ? rsqrt(dv) : 1.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      // 1/sqrt(d) by the hardware seed + Newton (rsqrt, <= 1 ulp), L_cc = d * rsqrt(d): a third of the latency of
      // sqrt followed by a division on the one chain of the factorisation that cannot be parallelised
      const bool is_piv = c < npiv;
      const double lcc = is_piv ? dv * inv : dv;
      if (is_piv && !(dv > 0.0) && tid == 0 && s_info == 0) s_info = r0 + c0 + c + 1;
      if (is_piv && warp == 7 && lane == c) piv[c0 + c] = lcc;
      const double li = Dr[c] * inv;       // lane i > c: L[i][c] of the micro-block
      double xc = a[c] * inv;
      if (a_row && R - c0 == c) xc = lcc;  // the diagonal element itself
      double dv_n = 0.0, inv_n = 1.0;
      if (c + 1 < 8) {
        const double l = __shfl_sync(0xffffffffu, li, c + 1);
        Dr[c + 1] = fma(-li, l, Dr[c + 1]);
        a[c + 1] = fma(-xc, l, a[c + 1]);
        dv_n = __shfl_sync(0xffffffffu, Dr[c + 1], c + 1);
        inv_n = (c + 1 < npiv) ? rsqrt(dv_n) : 1.0;
      }
#pragma unroll
      for (int c2 = c + 2; c2 < 8; ++c2) {
        const double l = __shfl_sync(0xffffffffu, li, c2);   // L[c2][c]
        Dr[c2] = fma(-li, l, Dr[c2]);
        a[c2] = fma(-xc, l, a[c2]);
      }
      a[c] = xc;
      dv = dv_n; inv = inv_n;
    }
    // rows of the micro-block itself: zero above the diagonal (those entries of the symmetric block were never loaded)
    if (a_row && R >= c0 && R < c0 + 8) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c > R - c0) a[c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double v = active ? a[c] : 0.0;
      Ps[c * D_P + R] = v;
      if (a_row) {
        if (R < bs && c0 + c < bs) Ag[R + (size_t)(c0 + c) * ld] = v;     // L (explicit zeros above the diagonal)
      } else {
        Ws[(c0 + c) * D_WLD + xi] = v;                                   // inv(L)[c0 + c][xi] = X[xi][c0 + c]
      }
    }
    }
    __syncthreads();
    // ---- (2) rank-8 update of the live tiles, publication of the next panel column -------------------------------
    if (m + 1 < nsteps) {
      // slots in chunks of 4: one uniform branch per chunk, straight-line inside (operands of dead tiles are zeroed),
      // so that the 16 fragment loads and the 4 independent DMMA chains of a chunk overlap
#pragma unroll
      for (int s0 = 0; s0 < D_SLOTS; s0 += 4) {
        int ti[4], tj[4]; bool is_a[4], act[4];
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (s0 + u < D_SLOTS) {
            diag_tile_of<NW>(s0 + u, warp, ti[u], tj[u], is_a[u]);
            act[u] = tj[u] > m && (is_a[u] || ti[u] <= m);
            any = any || act[u];
          }
        }
        if (any) {
          double a0[4], a1[4], b0[4], b1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (s0 + u < D_SLOTS) {
              const int ra = (is_a[u] ? 0 : GPB_NB) + ti[u] * 8 + lr;
              const int rb = tj[u] * 8 + lr;
              a0[u] = Ps[lk * D_P + ra]; a1[u] = Ps[(4 + lk) * D_P + ra];
              b0[u] = Ps[lk * D_P + rb]; b1[u] = Ps[(4 + lk) * D_P + rb];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (s0 + u < D_SLOTS) dmma884(acc[s0 + u][0], acc[s0 + u][1], act[u] ? -a0[u] : 0.0, b0[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (s0 + u < D_SLOTS) dmma884(acc[s0 + u][0], acc[s0 + u][1], act[u] ? -a1[u] : 0.0, b1[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (s0 + u < D_SLOTS && tj[u] == m + 1) {
            const int rb = (is_a[u] ? 0 : GPB_NB) + ti[u] * 8 + lr;
            Xs[(2 * lk) * D_P + rb] = acc[s0 + u][0];
            Xs[(2 * lk + 1) * D_P + rb] = acc[s0 + u][1];
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- log-det partial, info, inv(L) to global memory -----------------------------------------------------------------
  if (warp < 4) {
    lsum = (tid < bf) ? log(piv[tid]) : 0.0;
    lsum = warp_sum(lsum);
    if (lane == 0) Xs[warp] = lsum;      // Xs is free after the last step
  }
  __syncthreads();
  if (tid == 0) {
    d.part[k] = (Xs[0] + Xs[1]) + (Xs[2] + Xs[3]);
    if (s_info != 0 && *d.info == 0) *d.info = s_info;
  }
  double* Wg = d.Wd + (size_t)k * GPB_NB * GPB_NB;
  for (int idx = tid; idx < GPB_NB * GPB_NB; idx += NT) {
    const int i = idx & (GPB_NB - 1), j = idx >> 7;
    Wg[idx] = (i >= j && i < bf) ? Ws[i * D_WLD + j] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// diagonal block, second design (default): Cholesky in registers + in-place triangular inverse in one shared square.
//
// The [A; I] elimination above is bound by the LATENCY of its instruction stream (8 warps, 36 instructions per DMMA, a
// pivot chain of shuffles).  This version cuts the work and the instruction count:
//   * Only A is eliminated: the 136 lower 8 x 8 tiles live in registers as NEGATED accumulators (N = -C, so the rank-8
//     update is N += P_i P_j^T with unmodified operands), 17 per warp, sorted by column.  A tile is never read again once
//     its column has been published, so "dead" tiles need no predication - whole groups of four are skipped instead.
//   * The 8 x 8 diagonal micro-block is factorised IN-THREAD by every thread that owns a row (36 broadcast loads, then
//     straight-line rsqrt / mul / fma in registers: no shuffles, no barriers inside the pivot chain), and the same
//     eliminations are applied to the thread's own row.
//   * L accumulates in a 128 x 132 shared square (pitch = 4 mod 16: every DMMA fragment load - row- or k-major - is bank
//     conflict free).  W = inv(L) is then computed IN PLACE by recursive doubling (8 -> 16 -> 32 -> 64 -> 128), all DMMA:
//     T = L21 W11 goes to the (free) upper-right block of each sub-problem, W21 = -W22 T overwrites L21.
// Rows of the carried right-hand side inside the block are ordinary non-pivot rows, exactly as above.
// ---------------------------------------------------------------------------------------------------------------
constexpr int D2_P = GPB_NB + 4;
constexpr int D2_SMEM_BYTES = (GPB_NB * D2_P + 2 * GPB_NB + 8) * (int)sizeof(double);

__host__ __device__ constexpr int d2_prefix(int tj) { return tj * 16 - tj * (tj - 1) / 2; }   // tiles left of column tj

// tile L = 8 * slot + warp of the column-sorted lower-triangular tile list (column tj holds ti = tj .. 15)
__device__ __forceinline__ void d2_tile_of(int slot, int warp, int& ti, int& tj) {
  const int L = 8 * slot + warp;
  int c = 0;
#pragma unroll
  for (int q = 1; q < 16; ++q) c += (L >= d2_prefix(q)) ? 1 : 0;
  tj = c;
  ti = c + (L - d2_prefix(c));
}

// one 8 x 8 tile product on the shared square: acc += A(ra.., ka..) * B, fragments as in dmma884
//   a: element (row, k) at S[(ka + k) * P + ra + row]      b: element (k, col) at S[bk + k * bstride_k + col * bstride_c]
__device__ __forceinline__ void d2_mma8(double (&acc)[2], const double* __restrict__ S, int ra, int ka, int boff, int bsk,
                                        int bsc, int lr, int lk) {
  const double a0 = S[(ka + lk) * D2_P + ra + lr], a1 = S[(ka + 4 + lk) * D2_P + ra + lr];
  const double b0 = S[boff + lk * bsk + lr * bsc], b1 = S[boff + (4 + lk) * bsk + lr * bsc];
  dmma884(acc[0], acc[1], a0, b0);
  dmma884(acc[0], acc[1], a1, b1);
}

// phase clocks of one diag2 launch (developer builds only: GPB_NVCC_EXTRA=-DGPB_DIAG_CLOCKS=1; read with
// gpb_debug_diag_clocks): [0] load, [1] sum of panel phases, [2] sum of update phases, [3] log-det / identity rows,
// [4] level 0 of the inverse, [5] levels 8 .. 64, [6] store
#ifdef GPB_DIAG_CLOCKS
__device__ long long g_diag_clocks[8];
#define D2_CLK(slot) do { if (tid == 0) { const long long now_ = clock64(); g_diag_clocks[slot] += now_ - clk_; clk_ = now_; } } while (0)
#else
#define D2_CLK(slot) do { } while (0)
#endif

// 1 / sqrt(d) for the pivot chain: hardware seed (MUFU.RSQ64H, ~2^-20) and ONE third-order step
//   e = 1 - d y^2,  y <- y + y e (1/2 + 3/8 e)      (error ~ 5/16 e^3 < 2^-60)
// - four dependent FP64 operations behind the seed.  The library's rsqrt adds range checks for zero / subnormal / infinite
// arguments; a non-positive pivot yields NaN here as there (the launch reports it through `info`).
__device__ __forceinline__ double d2_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double u = y * e;
  return fma(u, fma(e, 0.375, 0.5), y);
}

__global__ void __launch_bounds__(256, 1) diag2_kernel(const GpbMat* __restrict__ mats, int k) {
  extern __shared__ __align__(16) double dsm[];
  double* S = dsm;                       // [128][D2_P] column-major square: L (lower) -> W = inv(L) (lower)
  double* piv = S + GPB_NB * D2_P;       // [128] pivots L_jj, [128 .. 136) scratch, [136 .. 264) their reciprocals
  __shared__ int s_info;
  const GpbMat d = mats[blockIdx.x];
  const int nrows = d.n + d.aug;
  const int r0 = k * GPB_NB;
  if (r0 >= d.n) return;
  const int bs = min(GPB_NB, nrows - r0);  // rows / columns held by the block (pivots + carried)
  const int bf = min(GPB_NB, d.n - r0);    // pivots
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const size_t ld = d.ld;
  double* Ag = d.A + r0 + (size_t)r0 * ld;
  if (tid == 0) s_info = 0;
  double* rpiv = piv + GPB_NB + 8;       // 1 / L_jj (the rsqrt of the pivot chain, <= 1 ulp)
  if (tid < GPB_NB) { piv[tid] = 1.0; rpiv[tid] = 1.0; }
#ifdef GPB_DIAG_CLOCKS
  long long clk_ = clock64();
  if (tid == 0) for (int q = 0; q < 8; ++q) g_diag_clocks[q] = 0;
#endif

  // ---- the lower triangle as negated accumulator tiles ---------------------------------------------------------------
  constexpr int NS = 17;
  double acc[NS][2];
  int pos[NS];                            // (ti * 8 + lr) | (tj * 8 + lr) << 8 | tj << 16
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    int ti, tj;
    d2_tile_of(s, warp, ti, tj);
    pos[s] = (ti * 8 + lr) | ((tj * 8 + lr) << 8) | (tj << 16);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int row = ti * 8 + lr, col = tj * 8 + 2 * lk + e;
      acc[s][e] = (row < bs && col < bs && row >= col) ? -Ag[row + (size_t)col * ld] : 0.0;
    }
    if (tj == 0) {                        // publish panel column 0
      S[(2 * lk) * D2_P + ti * 8 + lr] = -acc[s][0];
      S[(2 * lk + 1) * D2_P + ti * 8 + lr] = -acc[s][1];
    }
  }
  __syncthreads();
  D2_CLK(0);

  const int R = tid;                       // row owned in the panel phase (threads 0 .. 127)
  const int nsteps = (bs + 7) / 8;
  for (int m = 0; m < nsteps; ++m) {
    const int c0 = 8 * m;
    // ---- (1) panel: in-thread factorisation of the 8 x 8 micro-block, same eliminations on the own row --------------
    const bool panel_thread = R < GPB_NB && R >= c0;
    double D[8][8], a[8];
    if (panel_thread) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        a[c] = S[(c0 + c) * D2_P + R];
#pragma unroll
        for (int i = c; i < 8; ++i) D[i][c] = S[(c0 + c) * D2_P + c0 + i];      // broadcast loads (same address per warp)
      }
    }
    // The rows of the micro-block are written back IN PLACE below, into the very entries every other warp has just read
    // as D: all loads must have completed before the first store (without this barrier a warp that is late with its
    // loads sees a mixture of C and L values - observed as non-reproducible results once 1024 blocks keep every SM busy).
    __syncthreads();
    if (panel_thread) {
      const int npiv = min(8, max(0, bf - c0));
      const int rl = R - c0;               // < 8: this row lies inside the micro-block
      double dvs[8], invs[8];              // pivots and their inverse roots (bookkeeping after the chain, off its path)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        dvs[c] = 1.0; invs[c] = 1.0;
        if (c < npiv) {                    // uniform
          const double dv = D[c][c];
          const double inv = d2_rsqrt(dv);
          dvs[c] = dv; invs[c] = inv;
#pragma unroll
          for (int i = c + 1; i < 8; ++i) D[i][c] *= inv;
          const double xc = a[c] * inv;
#pragma unroll
          for (int j = c + 1; j < 8; ++j) {
#pragma unroll
            for (int i = j; i < 8; ++i) D[i][j] = fma(-D[i][c], D[j][c], D[i][j]);
            a[j] = fma(-xc, D[j][c], a[j]);
          }
          a[c] = xc;
        }
      }
      if (rl == 0) {                       // one thread per micro-block: pivots L_cc = d / sqrt(d), their reciprocals, info
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < npiv) {
            piv[c0 + c] = dvs[c] * invs[c];
            rpiv[c0 + c] = invs[c];
            if (!(dvs[c] > 0.0) && s_info == 0) s_info = r0 + c0 + c + 1;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        // rows of the micro-block itself: explicit zeros above the diagonal (those entries were never loaded)
        const double v = (rl < c) ? 0.0 : a[c];
        S[(c0 + c) * D2_P + R] = v;
        if (R < bs && c0 + c < bs) Ag[R + (size_t)(c0 + c) * ld] = v;
      }
    }
    __syncthreads();
    D2_CLK(1);
    // ---- (2) rank-8 update of the live tiles (groups of four: dead tiles are a prefix and are never read again) -------
    if (m + 1 < nsteps) {
#pragma unroll
      for (int g0 = 0; g0 < NS; g0 += 4) {
        constexpr int G = 4;
        const int last = (g0 + G - 1 < NS) ? g0 + G - 1 : NS - 1;
        if ((pos[last] >> 16) > m) {       // uniform per warp: the group has at least one live tile
          double a0[G], a1[G], b0[G], b1[G];
#pragma unroll
          for (int u = 0; u < G; ++u) {
            if (g0 + u < NS) {
              const int ra = pos[g0 + u] & 255, rb = (pos[g0 + u] >> 8) & 255;
              a0[u] = S[(c0 + lk) * D2_P + ra]; a1[u] = S[(c0 + 4 + lk) * D2_P + ra];
              b0[u] = S[(c0 + lk) * D2_P + rb]; b1[u] = S[(c0 + 4 + lk) * D2_P + rb];
            }
          }
#pragma unroll
          for (int u = 0; u < G; ++u)
            if (g0 + u < NS) dmma884(acc[g0 + u][0], acc[g0 + u][1], a0[u], b0[u]);
#pragma unroll
          for (int u = 0; u < G; ++u)
            if (g0 + u < NS) dmma884(acc[g0 + u][0], acc[g0 + u][1], a1[u], b1[u]);
#pragma unroll
          for (int u = 0; u < G; ++u) {
            if (g0 + u < NS && (pos[g0 + u] >> 16) == m + 1) {     // next panel column: publish
              const int ra = pos[g0 + u] & 255;
              S[(c0 + 8 + 2 * lk) * D2_P + ra] = -acc[g0 + u][0];
              S[(c0 + 8 + 2 * lk + 1) * D2_P + ra] = -acc[g0 + u][1];
            }
          }
        }
      }
    }
    __syncthreads();
    D2_CLK(2);
  }

  // ---- log-det partial, info -------------------------------------------------------------------------------------------
  if (warp < 4) {
    double lsum = (tid < bf) ? log(piv[tid]) : 0.0;
    lsum = warp_sum(lsum);
    if (lane == 0) piv[GPB_NB + warp] = lsum;
  }
  // ---- W = inv(L) in place.  Rows / columns that hold no pivot (a partial last block, the carried row) become identity
  //      rows so that nothing non-finite can leak into the pivot rows through 0 * x products.
  if (bf < GPB_NB) {
    for (int idx = tid; idx < GPB_NB * GPB_NB; idx += 256) {
      const int i = idx & (GPB_NB - 1), j = idx >> 7;
      if (i >= bf || j >= bf) S[j * D2_P + i] = (i == j) ? 1.0 : 0.0;   // also the never-initialised part of the square
    }
  }
  __syncthreads();
  if (tid == 0) {
    d.part[k] = (piv[GPB_NB] + piv[GPB_NB + 1]) + (piv[GPB_NB + 2] + piv[GPB_NB + 3]);
    if (s_info != 0 && *d.info == 0) *d.info = s_info;
  }
  D2_CLK(3);
  // level 0: the sixteen 8 x 8 diagonal micro-blocks; thread t < 128 owns column (t & 7) of block (t >> 3):
  //   W_jj = 1 / L_jj,  W_ij = -(1 / L_ii) sum_{q = j}^{i-1} L_iq W_qj   (the columns of an inverse are independent)
  if (tid < GPB_NB) {
    const int b0 = (tid >> 3) * 8, j = tid & 7;
    double Wc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < i; ++q)
        if (q >= j) t = fma(S[(b0 + q) * D2_P + b0 + i], Wc[q], t);
      const double ri = (b0 + i < bf) ? rpiv[b0 + i] : 1.0;
      Wc[i] = (i == j) ? ri : ((i > j) ? -t * ri : 0.0);
    }
    __syncwarp();      // all loads of L precede the stores of W (lanes of a warp share micro-blocks)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i >= j) S[(b0 + j) * D2_P + b0 + i] = Wc[i];
  }
  __syncthreads();
  D2_CLK(4);
  // levels s = 8 .. 64: [W11 0; L21 W22] -> W21 = -W22 (L21 W11).  A warp computes a 2 x 2 block of 8 x 8 tiles at a time
  // (four independent accumulator chains, every fragment load feeds two DMMAs); index math is shifts (s is a power of 2).
  {
    // s = 8: one tile per sub-problem and warp: W21 = -W22 * (L21 * W11) with all three factors 8 x 8
    const int q0 = 16 * warp, qA = q0 + 8;
    double c2[2] = {0.0, 0.0};
    d2_mma8(c2, S, qA, q0, q0 * D2_P + q0, 1, D2_P, lr, lk);                      // T = L21 W11
    S[(qA + 2 * lk) * D2_P + q0 + lr] = c2[0];
    S[(qA + 2 * lk + 1) * D2_P + q0 + lr] = c2[1];
    __syncwarp();
    double w2[2] = {0.0, 0.0};
    d2_mma8(w2, S, qA, qA, qA * D2_P + q0, 1, D2_P, lr, lk);                      // W22 T
    S[(q0 + 2 * lk) * D2_P + qA + lr] = -w2[0];
    S[(q0 + 2 * lk + 1) * D2_P + qA + lr] = -w2[1];
  }
  __syncthreads();
  for (int ls = 4; ls < 7; ++ls) {          // s = 16, 32, 64
    const int s = 1 << ls;
    const int lb = ls - 4;                  // log2 of the 2 x 2 blocks per side of a sub-problem (s / 16)
    const int nb = (GPB_NB >> (ls + 1)) << (2 * lb);   // sub-problems x blocks per sub-problem: 4, 8, 16
    // phase T: T = L21 W11 into the upper-right block; k >= tile column (W11 lower triangular)
    for (int t = warp; t < nb; t += 8) {
      const int p = t >> (2 * lb), rem = t & ((1 << (2 * lb)) - 1);
      const int bj = rem >> lb, bi = (rem + bj) & ((1 << lb) - 1);     // rotate rows per column: balances the k-ranges
      const int q0 = p << (ls + 1), qA = q0 + s;
      const int ti = 2 * bi, tj = 2 * bj, ts = s >> 3;
      double c[2][2][2] = {};
      for (int kk = tj; kk < ts; ++kk) {
        const int ka = q0 + kk * 8;
        const double a00 = S[(ka + lk) * D2_P + qA + ti * 8 + lr], a01 = S[(ka + 4 + lk) * D2_P + qA + ti * 8 + lr];
        const double a10 = S[(ka + lk) * D2_P + qA + ti * 8 + 8 + lr], a11 = S[(ka + 4 + lk) * D2_P + qA + ti * 8 + 8 + lr];
        const double* bp = S + (q0 + tj * 8 + lr) * D2_P + ka + lk;
        const double b00 = bp[0], b01 = bp[4], b10 = bp[8 * D2_P], b11 = bp[8 * D2_P + 4];
        dmma884(c[0][0][0], c[0][0][1], a00, b00); dmma884(c[1][0][0], c[1][0][1], a10, b00);     // column tj: every kk >= tj
        if (kk > tj) {                      // column tj + 1: the tile (kk = tj, tj + 1) of W11 lies above its diagonal
          dmma884(c[0][1][0], c[0][1][1], a00, b10); dmma884(c[1][1][0], c[1][1][1], a10, b10);
        }
        dmma884(c[0][0][0], c[0][0][1], a01, b01); dmma884(c[1][0][0], c[1][0][1], a11, b01);
        if (kk > tj) {
          dmma884(c[0][1][0], c[0][1][1], a01, b11); dmma884(c[1][1][0], c[1][1][1], a11, b11);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          S[(qA + (tj + v) * 8 + 2 * lk) * D2_P + q0 + (ti + u) * 8 + lr] = c[u][v][0];
          S[(qA + (tj + v) * 8 + 2 * lk + 1) * D2_P + q0 + (ti + u) * 8 + lr] = c[u][v][1];
        }
    }
    __syncthreads();
    // phase W: W21 = -W22 T; k <= tile row (W22 lower triangular)
    for (int t = warp; t < nb; t += 8) {
      const int p = t >> (2 * lb), rem = t & ((1 << (2 * lb)) - 1);
      const int bj = rem >> lb, bi = ((1 << lb) - 1) - ((rem + bj) & ((1 << lb) - 1));
      const int q0 = p << (ls + 1), qA = q0 + s;
      const int ti = 2 * bi, tj = 2 * bj;
      double c[2][2][2] = {};
      for (int kk = 0; kk <= ti + 1; ++kk) {
        const int ka = qA + kk * 8;
        const double a00 = S[(ka + lk) * D2_P + qA + ti * 8 + lr], a01 = S[(ka + 4 + lk) * D2_P + qA + ti * 8 + lr];
        const double a10 = S[(ka + lk) * D2_P + qA + ti * 8 + 8 + lr], a11 = S[(ka + 4 + lk) * D2_P + qA + ti * 8 + 8 + lr];
        const double* bp = S + (qA + tj * 8 + lr) * D2_P + q0 + kk * 8 + lk;
        const double b00 = bp[0], b01 = bp[4], b10 = bp[8 * D2_P], b11 = bp[8 * D2_P + 4];
        if (kk <= ti) {                     // row ti: the tile (ti, kk = ti + 1) of W22 lies above its diagonal
          dmma884(c[0][0][0], c[0][0][1], a00, b00); dmma884(c[0][1][0], c[0][1][1], a00, b10);
        }
        dmma884(c[1][0][0], c[1][0][1], a10, b00); dmma884(c[1][1][0], c[1][1][1], a10, b10);
        if (kk <= ti) {
          dmma884(c[0][0][0], c[0][0][1], a01, b01); dmma884(c[0][1][0], c[0][1][1], a01, b11);
        }
        dmma884(c[1][0][0], c[1][0][1], a11, b01); dmma884(c[1][1][0], c[1][1][1], a11, b11);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          S[(q0 + (tj + v) * 8 + 2 * lk) * D2_P + qA + (ti + u) * 8 + lr] = -c[u][v][0];
          S[(q0 + (tj + v) * 8 + 2 * lk + 1) * D2_P + qA + (ti + u) * 8 + lr] = -c[u][v][1];
        }
    }
    __syncthreads();
  }
  D2_CLK(5);
  double* Wg = d.Wd + (size_t)k * GPB_NB * GPB_NB;
  for (int idx = tid; idx < GPB_NB * GPB_NB; idx += 256) {
    const int i = idx & (GPB_NB - 1), j = idx >> 7;
    Wg[idx] = (i >= j && i < bf) ? S[j * D2_P + i] : 0.0;
  }
  D2_CLK(6);
}

int debug_diag_clocks(long long* out) {
#ifdef GPB_DIAG_CLOCKS
  return cudaMemcpyFromSymbol(out, g_diag_clocks, 8 * sizeof(long long)) == cudaSuccess ? 0 : 1;
#else
  (void)out;
  return -1;
#endif
}

static int diag_design() {      // GPB_DIAG=old: the [A; I] elimination kernel
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPB_DIAG"); v = (e && e[0] == 'o') ? 0 : 1; }
  return v;
}

static int diag_warps() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPB_DIAG_WARPS"); v = (e && atoi(e) == 16) ? 16 : 8; }
  return v;
}
static void launch_diag(const GpbMat* dm, int B, int k, cudaStream_t s) {
  if (diag_design() == 1) diag2_kernel<<<B, 256, D2_SMEM_BYTES, s>>>(dm, k);
  else if (diag_warps() == 16) diag_kernel_t<16><<<B, 512, D_SMEM_BYTES, s>>>(dm, k);
  else diag_kernel_t<8><<<B, 256, D_SMEM_BYTES, s>>>(dm, k);
}

// copy the inverted diagonal blocks into place (level 0 of the recursive-doubling inverse)
__global__ void diag_copy_kernel(const GpbMat* __restrict__ mats, int k0) {
  const GpbMat d = mats[blockIdx.y];
  const int k = blockIdx.x + k0;
  const int r0 = k * GPB_NB;
  if (r0 >= d.n) return;
  const int bf = min(GPB_NB, d.n - r0);
  const double* Wg = d.Wd + (size_t)k * GPB_NB * GPB_NB;
  double* Ag = d.A + r0 + (size_t)r0 * d.ld;
  for (int idx = threadIdx.x; idx < GPB_NB * GPB_NB; idx += blockDim.x) {
    const int i = idx & (GPB_NB - 1), j = idx >> 7;
    if (i < bf && j < bf) Ag[i + (size_t)j * d.ld] = Wg[idx];
  }
}

// The carried right-hand side y^T (row n) of the matrices whose tiles exclude it (carried_row_separate):
//   mode 0 (after the diagonal block kp): row n of block column kp times inv(L_kk)^T, as the panel product does
//   mode 1 (with the panel blocks kp .. kp + kb - 1): row n of the trailing columns [c_lo, c_hi) (128-wide, from
//          (kp + kb) * 128) minus z_panel . P[j, :]^T - including the element (n, n), which accumulates -z^T z
__global__ void __launch_bounds__(256) carried_row_kernel(const GpbMat* __restrict__ mats, int kp, int kb, int c_lo, int c_hi,
                                                          int mode) {
  __shared__ double zs[GPB_KB_MAX * GPB_NB];
  const GpbMat d = mats[blockIdx.z];
  if (!carried_row_separate(d, 1)) return;
  const int n = d.n;
  const size_t ld = d.ld;
  const int tid = threadIdx.x;
  double* rown = d.A + n;                                   // rown[j * ld] = A[n, j]
  if (mode == 0) {
    if ((kp + 1) * GPB_NB > n || blockIdx.x != 0) return;   // block kp is the last one: the diagonal kernel owned the row
    if (tid < GPB_NB) zs[tid] = rown[(size_t)(kp * GPB_NB + tid) * ld];
    __syncthreads();
    if (tid < GPB_NB) {
      const double* Wg = d.Wd + (size_t)kp * GPB_NB * GPB_NB;
      double acc = 0.0;
      for (int j = 0; j <= tid; ++j) acc = fma(zs[j], Wg[tid + j * GPB_NB], acc);
      rown[(size_t)(kp * GPB_NB + tid) * ld] = acc;
    }
    return;
  }
  const int r0 = (kp + kb) * GPB_NB;
  if (r0 > n) return;
  const int kw = kb * GPB_NB;
  for (int q = tid; q < kw; q += 256) zs[q] = rown[(size_t)(kp * GPB_NB + q) * ld];
  __syncthreads();
  const int j = r0 + c_lo * GPB_NB + blockIdx.x * 256 + tid;           // column of row n = row of the panel
  const long long j_hi = (long long)r0 + (long long)c_hi * GPB_NB;
  if (j > n || j >= j_hi) return;
  const double* P = d.A + j + (size_t)kp * GPB_NB * ld;
  double acc = 0.0;
#pragma unroll 4
  for (int q = 0; q < kw; ++q) acc = fma(zs[q], P[(size_t)q * ld], acc);
  rown[(size_t)j * ld] -= acc;
}

// nll = 1/2 z^T z + sum log diag L + 1/2 n log(2 pi)   (Metrics/LogLikelihood.py:39-49,65); also gathers z
__global__ void finalize_kernel(const GpbMat* __restrict__ mats, double log2pi) {
  const GpbMat d = mats[blockIdx.x];
  const int n = d.n;
  const size_t ld = d.ld;
  for (int i = threadIdx.x; i < n; i += blockDim.x) d.zvec[i] = d.A[n + (size_t)i * ld];
  if (threadIdx.x == 0) {
    const int nblk = (n + GPB_NB - 1) / GPB_NB;
    double logdet = 0.0;
    for (int k = 0; k < nblk; ++k) logdet += d.part[k];
    const double quad = -d.A[n + (size_t)n * ld];
    double nll = 0.5 * quad + logdet + 0.5 * ((double)n * log2pi);
    if (*d.info != 0) nll = nan("");
    *d.nll = nll;
    d.terms[0] = quad; d.terms[1] = logdet;
  }
}

// alpha = W^T z : one warp per column of the lower-triangular W
__global__ void __launch_bounds__(256) alpha_kernel(const GpbMat* __restrict__ mats) {
  const GpbMat d = mats[blockIdx.y];
  const int col = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (col >= d.n || *d.info != 0) return;
  const int lane = threadIdx.x & 31;
  const double* w = d.A + (size_t)col * d.ld;
  double s = 0.0;
  for (int i = col + lane; i < d.n; i += 32) s += w[i] * d.zvec[i];
  s = warp_sum(s);
  if (lane == 0) d.alpha[col] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// standalone blocked triangular solves with the stored diagonal-block inverses (get_L_alpha without gradient)
//   forward : v = y, for k: x_k = Wd_k v_k ; v[i>blk k] -= L[i, blk k] x_k          (x = zvec)
//   backward: v = z, for k desc: x_k = Wd_k^T v_k ; v[c<blk k] -= L[blk k, c]^T x_k  (x = alpha)
// Every CTA recomputes x_k (128 x 128 matvec from L2) so that a step is one launch.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trsv_step_kernel(const GpbMat* __restrict__ mats, int k, int transposed) {
  __shared__ double xk[GPB_NB];
  __shared__ double vk[GPB_NB];
  const GpbMat d = mats[blockIdx.y];
  const int r0 = k * GPB_NB;
  if (r0 >= d.n) return;
  const int bf = min(GPB_NB, d.n - r0);
  const int tid = threadIdx.x;
  const double* Wg = d.Wd + (size_t)k * GPB_NB * GPB_NB;
  double* v = d.tmpv;
  double* x = transposed ? d.alpha : d.zvec;
  if (tid < GPB_NB) vk[tid] = (tid < bf) ? v[r0 + tid] : 0.0;
  __syncthreads();
  if (tid < GPB_NB) {
    double s = 0.0;
    if (tid < bf) {
      if (!transposed) { for (int j = 0; j <= tid; ++j) s += Wg[tid + j * GPB_NB] * vk[j]; }
      else { for (int j = tid; j < bf; ++j) s += Wg[j + tid * GPB_NB] * vk[j]; }
    }
    xk[tid] = s;
    if (blockIdx.x == 0 && tid < bf) x[r0 + tid] = s;
  }
  __syncthreads();
  const size_t ld = d.ld;
  if (!transposed) {
    const int i = r0 + GPB_NB + blockIdx.x * 256 + tid;
    if (i < d.n) {
      const double* Lr = d.A + i + (size_t)r0 * ld;
      double s = 0.0;
      for (int j = 0; j < GPB_NB; ++j) s += Lr[(size_t)j * ld] * xk[j];
      v[i] -= s;
    }
  } else {
    const int lane = tid & 31;
    const int c = blockIdx.x * 8 + (tid >> 5);
    if (c < r0) {
      const double* Lc = d.A + r0 + (size_t)c * ld;
      double s = 0.0;
      for (int j = lane; j < bf; j += 32) s += Lc[j] * xk[j];
      s = warp_sum(s);
      if (lane == 0) v[c] -= s;
    }
  }
}

__global__ void trsv_init_kernel(const GpbMat* __restrict__ mats, int transposed) {
  const GpbMat d = mats[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d.n) d.tmpv[i] = transposed ? d.zvec[i] : d.y[i];
}

__global__ void zero_upper_kernel(double* A, int n, int ld) {
  const int j = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < j && i < n) A[i + (size_t)j * ld] = 0.0;
}
__global__ void symmetrize_kernel(double* A, int n, int ld) {
  // A[i,j] (i<j) = A[j,i]; 32x32 tiles through shared memory for coalescing on both sides
  __shared__ double t[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bi > bj) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  // read the lower tile (rows bj*32.., cols bi*32..)
  for (int r = ty; r < 32; r += 8) {
    const int gi = bj * 32 + tx, gj = bi * 32 + r;
    t[r][tx] = (gi < n && gj < n) ? A[gi + (size_t)gj * ld] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int gi = bi * 32 + tx, gj = bj * 32 + r;  // upper element (gi < gj)
    if (gi < n && gj < n && gi < gj) A[gi + (size_t)gj * ld] = t[tx][r];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------------------------
#define GPB_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// tile configuration used by the drivers: GPB_GEMM_CFG=big|half (default half)
static int g_cfg_half = -1;
static bool cfg_half() {
  if (g_cfg_half < 0) {
    const char* e = getenv("GPB_GEMM_CFG");
    g_cfg_half = (e && e[0] == 'b') ? 0 : 1;
  }
  return g_cfg_half == 1;
}

// GPB_TMA=1: the trailing updates of the factorisation stage their operands with bulk copies (gemm_bulk_kernel)
static bool bulk_staging() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPB_TMA"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// persistent = CTAs walk up to TILES_PER_CTA tiles each (equal-cost tiles: the trailing update); otherwise one CTA per
// tile under the hardware scheduler (tiles of very different k-length: triangular inverse, W^T W)
template <class Cfg, bool AKM, bool BKM, class Geo>
static cudaError_t launch_cfg(const Geo& geo, dim3 grid, cudaStream_t s, bool persistent = false, const char* tag = "gemm",
                              int ta = 0) {
  // Tiles per CTA of a persistent (bulk) launch.  Two tiles keep the operand pipeline running across a tile boundary
  // (measured at n = 32768: 1: 425 ms, 2: 415 ms, 4: 416 ms, 8: 423 ms potrf), but CTAs of two tiles quantise a launch of
  // a few waves badly (n = 8192, 1849 tiles: 925 CTAs = 3.1 waves of 296 -> 4 x 2 tile times against 7 x 1), so launches
  // below 8 waves use one tile per CTA (potrf at n = 8192 8.47 -> 8.29 ms).
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return cudaSuccess;
  const long long tiles = (long long)grid.x * grid.y * grid.z;
  const int TILES_PER_CTA = tiles >= 8LL * 296 ? 2 : 1;
  TraceSpan span(tag, s, ta, (int)tiles);
  if constexpr (!AKM && !BKM && std::is_same<Geo, GeoSyrk>::value && !std::is_same<Cfg, CfgQuarter>::value) {
    if (bulk_staging()) {
      gemm_bulk_kernel<Cfg, Geo><<<persistent_ctas(grid, Cfg::MIN_CTAS, persistent ? TILES_PER_CTA : 1), Cfg::THREADS,
                                   Cfg::SMEM_BYTES + G_BULK_EXTRA, s>>>(geo, grid);
      ++g_launches;
      return cudaGetLastError();
    }
  }
  gemm_kernel<Cfg, AKM, BKM, Geo><<<persistent_ctas(grid, Cfg::MIN_CTAS, persistent ? TILES_PER_CTA : 1), Cfg::THREADS,
                                   Cfg::SMEM_BYTES, s>>>(geo, grid);
  ++g_launches;
  return cudaGetLastError();
}

template <class Cfg, bool AKM, bool BKM, class Geo>
static cudaError_t set_smem() {
  return cudaFuncSetAttribute(gemm_kernel<Cfg, AKM, BKM, Geo>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Cfg::SMEM_BYTES);
}
template <class Cfg>
static cudaError_t set_smem_all() {
  GPB_CK((set_smem<Cfg, false, false, GeoSyrk>()));
  GPB_CK((set_smem<Cfg, false, false, GeoPanel>()));
  GPB_CK((set_smem<Cfg, false, true, GeoTrtriT>()));
  GPB_CK((set_smem<Cfg, false, true, GeoTrtriW>()));
  GPB_CK((set_smem<Cfg, true, true, GeoLauum>()));
  GPB_CK((set_smem<Cfg, false, false, GeoPlain>()));
  GPB_CK((set_smem<Cfg, false, true, GeoPlain>()));
  GPB_CK((set_smem<Cfg, true, false, GeoPlain>()));
  GPB_CK((set_smem<Cfg, true, true, GeoPlain>()));
  return cudaSuccess;
}

cudaError_t linalg_init() {
  {
    const char* e = getenv("GPB_RED");
    const int v = (e && e[0] == '0') ? 0 : 1;
    GPB_CK(cudaMemcpyToSymbol(g_red_epilogue, &v, sizeof(int)));
  }
  GPB_CK(set_smem_all<CfgBig>());
  GPB_CK(set_smem_all<CfgHalf>());
  GPB_CK(cudaFuncSetAttribute(gemm_bulk_kernel<CfgBig, GeoSyrk>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgBig::SMEM_BYTES + G_BULK_EXTRA));
  GPB_CK(cudaFuncSetAttribute(gemm_bulk_kernel<CfgHalf, GeoSyrk>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgHalf::SMEM_BYTES + G_BULK_EXTRA));
  GPB_CK((set_smem<CfgQuarter, false, false, GeoSyrk>()));
  GPB_CK((set_smem<CfgQuarter, false, false, GeoPanel>()));
  GPB_CK(cudaFuncSetAttribute(diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM_BYTES));
  GPB_CK(cudaFuncSetAttribute(diag_kernel_t<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM_BYTES));
  GPB_CK(cudaFuncSetAttribute(diag_kernel_t<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM_BYTES));
  return cudaSuccess;
}

static bool grouped_raster() {      // GPB_SYRK_GROUPED=0: column-sweep tile order in the bulk updates
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPB_SYRK_GROUPED"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static bool quarter_tiles() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPB_QUARTER"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Outer panel width of the factorisation in 128-blocks, chosen per outer step from the rows that remain (`rem`).  The
// panels of an outer step hit the far trailing matrix together (k = 128 W: 28 TFLOP/s at W = 1, 31 at 2, 33 at 4) but
// lengthen the serial chain inside the outer panel, so the width shrinks as the trailing matrix does.  Measured on B200
// (potrf ms, fixed W = 1 / 2 / 3 / 4): n = 4096 3.16 / 3.43, 8192 10.08 / 8.44 / 8.50 / 9.02, 16384 - / 51.4 / 50.6 / 50.1,
// 32768 - / 381.5 / 372.1 / 367.6.  GPB_POTRF_KB=1..8 forces one width; GPB_POTRF_T2 / _T4 / _T8 move the thresholds.
// Batches of independent GPs have no critical path to protect: wide panels pay from n = 1024 on (C3 256 x n = 2048:
// potrf 33.7 / 31.0 / 30.6 ms at W = 1 / 2 / 4, C4 1024 x n = 1024: 21.8 / 19.4 / 19.1).
static int kb_forced() {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("GPB_POTRF_KB");
    forced = e ? std::max(1, std::min(GPB_KB_MAX, atoi(e))) : 0;
  }
  return forced;
}
static long long env_rows(const char* name, long long dflt) {
  const char* e = getenv(name);
  return e ? atoll(e) : dflt;
}
static int kb_for(long long rem, int B) {
  if (kb_forced()) return kb_forced();
  if (B >= 8) return rem >= 1024 ? 4 : 1;
  static const long long t2 = env_rows("GPB_POTRF_T2", 5120), t4 = env_rows("GPB_POTRF_T4", 10240),
                         t8 = env_rows("GPB_POTRF_T8", 20480);
  return rem >= t8 ? 8 : rem >= t4 ? 4 : rem >= t2 ? 2 : 1;
}
static int kb_max(int n_max, int B = 1) { return kb_for(n_max, B); }

// Blocked right-looking Cholesky.  Diagonal blocks and panels are 128 wide; two consecutive panels are applied to the
// far trailing matrix together (k = 256).  With look-ahead (one large matrix) the critical path - diagonal block, panel,
// the strip update of the second block column, and the first far column - runs on a high-priority stream so that its
// few CTAs are dispatched ahead of the thousands of queued trailing-update CTAs of the low-priority stream, which does
// the columns the next outer step needs first and then the bulk.
int potrf_outer_blocks(int n_max) { return std::min(2, kb_max(n_max)); }   // the distributed drivers: 1 or 2

// ---------------------------------------------------------------------------------------------------------------
// Triangular inverse overlapped with the factorisation (one large matrix).  The recursive-doubling inverse is a task
// graph: level s, sub-problem p needs only the block columns of L left of its own end, and a block column of L is final
// as soon as its panel is solved.  So while the factorisation's tail is bound by its serial chain (diagonal block ->
// panel -> column) and leaves most SMs idle, the inverse of everything already final runs on a third, low-priority
// stream.  Readiness (fc = number of final columns):
//   T(s, p) : T = L21 W11      needs columns [r0, rA) final and every lower-level task left of rA launched
//   W(s, p) : W21 = -W22 T     needs columns up to the sub-problem's end final, T(s, p) and every lower-level task left
//                              of that end launched
// Tasks are launched in level order on ONE stream, so "launched" implies "ordered before".  A W task overwrites
// L21 inside its own diagonal block only, i.e. rows above fc, which no later step of the factorisation reads.
// ---------------------------------------------------------------------------------------------------------------
struct TrtriSched {
  int n = 0, nblk = 0, nlev = 0, copied = 0;
  // Levels above smax wait for the end of the factorisation.  A tile of level s contracts over up to s elements - 536 us
  // at s = 4096 - and stream priorities act only when a CTA retires: with T(4096) resident on every SM each kernel of the
  // factorisation's serial chain waited for a slot, and the diagonal-block kernel (a whole SM) for two (launch timeline at
  // n = 8192: block columns 48 .. 52 took 1.9 ms instead of 0.3).  GPB_INV_SMAX overrides (0: no limit).
  long long smax = 1024;
  int nextT[24], nextW[24], nsub[24];
  void init(int n_) {
    n = n_; nblk = (n + GPB_NB - 1) / GPB_NB; nlev = 0; copied = 0;
    for (long long s = GPB_NB; s < n && nlev < 24; s *= 2) {
      nsub[nlev] = (int)((n - s + 2 * s - 1) / (2 * s));   // sub-problems with rA = 2 s p + s < n
      nextT[nlev] = nextW[nlev] = 0;
      ++nlev;
    }
    static long long env_smax = -1;
    if (env_smax < 0) { const char* e = getenv("GPB_INV_SMAX"); env_smax = e ? atoll(e) : 1024; }
    smax = env_smax > 0 ? env_smax : (1LL << 40);
  }
  bool lower_done(int li, long long X) const {            // everything below level li that starts left of X is launched
    const long long Xc = X < n ? X : n;
    if (copied < (int)((Xc + GPB_NB - 1) / GPB_NB)) return false;
    for (int lj = 0; lj < li; ++lj) {
      const long long s2 = 2LL * (GPB_NB << lj);
      long long need = (X + s2 - 1) / s2;
      if (need > nsub[lj]) need = nsub[lj];
      if (nextW[lj] < need) return false;
    }
    return true;
  }
  bool finished() const {
    if (copied < nblk) return false;
    for (int li = 0; li < nlev; ++li) if (nextW[li] < nsub[li]) return false;
    return true;
  }
  // Launches everything that became ready now that columns [0, fc) of L are final.  `L` issues the work:
  //   L.copy(first block, count), L.T(level s, first sub-problem, count), L.W(level s, first sub-problem, count)
  template <class Launcher>
  cudaError_t advance(long long fc, Launcher& L) {
    bool progress = true;
    while (progress) {
      progress = false;
      const int newc = fc >= n ? nblk : (int)(fc / GPB_NB);
      if (newc > copied) {
        GPB_CK(L.copy(copied, newc - copied));
        copied = newc; progress = true;
      }
      for (int li = 0; li < nlev; ++li) {
        const long long s = (long long)GPB_NB << li;
        if (s > smax && fc < n) break;
        int cnt = 0;
        while (nextT[li] + cnt < nsub[li]) {
          const long long rA = 2 * s * (nextT[li] + cnt) + s;
          if (rA <= fc && lower_done(li, rA)) ++cnt; else break;
        }
        if (cnt) {
          GPB_CK(L.T((int)s, nextT[li], cnt));
          nextT[li] += cnt; progress = true;
        }
        cnt = 0;
        while (nextW[li] + cnt < nextT[li]) {
          const long long end = 2 * s * (nextW[li] + cnt + 1);
          if ((end < n ? end : n) <= fc && lower_done(li, end)) ++cnt; else break;
        }
        if (cnt) {
          GPB_CK(L.W((int)s, nextW[li], cnt));
          nextW[li] += cnt; progress = true;
        }
      }
    }
    return cudaSuccess;
  }
};

template <class Cfg>
struct TrtriDeviceLauncher {        // the scheduler's tasks as kernel launches on one stream
  const GpbMat* dm;
  cudaStream_t st;
  cudaError_t copy(int k0, int cnt) {
    TraceSpan span("inv_copy", st, k0, cnt);
    diag_copy_kernel<<<dim3(cnt, 1), 1024, 0, st>>>(dm, k0);
    ++g_launches;
    return cudaGetLastError();
  }
  cudaError_t T(int s, int p0, int cnt) {
    const int tiles = (s / Cfg::BM) * (s / Cfg::BN);
    return launch_cfg<Cfg, false, true>(GeoTrtriT{dm, s, p0}, dim3(tiles, cnt, 1), st, false, "inv_T", s);
  }
  cudaError_t W(int s, int p0, int cnt) {
    const int tiles = (s / Cfg::BM) * (s / Cfg::BN);
    return launch_cfg<Cfg, false, true>(GeoTrtriW{dm, s, p0}, dim3(tiles, cnt, 1), st, false, "inv_W", s);
  }
};

struct TrtriRecorder {              // host-only: records the schedule (tests)
  int* out; int cap, count; long long fc;
  cudaError_t put(int kind, int s, int p0, int cnt) {
    for (int q = 0; q < cnt; ++q) {
      if (count < cap) { out[4 * count] = kind; out[4 * count + 1] = s; out[4 * count + 2] = p0 + q; out[4 * count + 3] = (int)fc; }
      ++count;
    }
    return cudaSuccess;
  }
  cudaError_t copy(int k0, int cnt) { return put(0, GPB_NB, k0, cnt); }
  cudaError_t T(int s, int p0, int cnt) { return put(1, s, p0, cnt); }
  cudaError_t W(int s, int p0, int cnt) { return put(2, s, p0, cnt); }
};

int trtri_schedule_host(int n, const long long* fc, int n_fc, int* out, int cap) {
  TrtriSched sched;
  sched.init(n);
  TrtriRecorder rec{out, cap, 0, 0};
  for (int i = 0; i < n_fc; ++i) {
    rec.fc = fc[i];
    sched.advance(fc[i] >= n ? (long long)n : fc[i], rec);
  }
  return sched.finished() ? rec.count : -rec.count - 1;
}



template <class Cfg>
static cudaError_t potrf_impl(const GpbMat* dm, int B, int n_max, int aug, bool lookahead, bool with_trtri,
                              const Exec& ex) {
  constexpr int BM = Cfg::BM, R = Cfg::BN / Cfg::BM;
  const int nrows = n_max + aug;
  const int nblk = (n_max + GPB_NB - 1) / GPB_NB;
  cudaStream_t ms = lookahead ? ex.crit : ex.main;
  if (lookahead) {
    GPB_CK(cudaEventRecord(ex.ev_fork, ex.main));
    GPB_CK(cudaStreamWaitEvent(ex.crit, ex.ev_fork, 0));
    GPB_CK(cudaStreamWaitEvent(ex.mid, ex.ev_fork, 0));
    GPB_CK(cudaStreamWaitEvent(ex.side, ex.ev_fork, 0));
  }
  const bool fuse = lookahead && with_trtri && B == 1;
  TrtriSched sched;
  if (fuse) { sched.init(n_max); GPB_CK(cudaStreamWaitEvent(ex.inv, ex.ev_fork, 0)); }
  int n_adv = 0;
  // block columns [0, kc) of L are final on stream ms at this point: hand everything that became invertible to ex.inv
  auto inverse_progress = [&](int kc) -> cudaError_t {
    if (!fuse) return cudaSuccess;
    GPB_CK(cudaEventRecord(ex.ev_c[n_adv & 1], ms));
    GPB_CK(cudaStreamWaitEvent(ex.inv, ex.ev_c[n_adv & 1], 0));
    ++n_adv;
    const long long fc = (long long)kc * GPB_NB;
    TrtriDeviceLauncher<Cfg> launcher{dm, ex.inv};
    return sched.advance(fc >= n_max ? (long long)n_max : fc, launcher);
  };
  auto full_with_rows = [&](int k) {   // block k (of the largest matrix) is a full pivot block with rows below it
    return k < nblk && (k + 1) * GPB_NB <= n_max && nrows - (k + 1) * GPB_NB > 0;
  };
  auto diag = [&](int k) -> cudaError_t {
    TraceSpan span("diag", ms, k, B);
    launch_diag(dm, B, k, ms);
    ++g_launches;
    return cudaGetLastError();
  };
  // launches of the critical path that would not fill one wave with 64-row tiles use 32-row tiles: half the tile latency
  auto small = [&](long long rows) { return ((rows + BM - 1) / BM) * B <= 148 && quarter_tiles(); };
  // Without look-ahead (batches, small matrices) the carried row of matrices with n = 0 mod 128 is kept out of the GEMM
  // tiles and updated by carried_row_kernel right behind every panel product / trailing update (same stream).
  const int sep = lookahead ? 0 : 1;
  auto carried = [&](int kp, int kb, int c_lo, int c_hi, int mode, cudaStream_t st) -> cudaError_t {
    if (!sep || !aug) return cudaSuccess;
    const int r0 = (kp + (mode ? kb : 1)) * GPB_NB;
    if (r0 > n_max) return cudaSuccess;                                 // the largest matrix decides the grid only
    const long long avail = (long long)n_max - ((long long)r0 + (long long)c_lo * GPB_NB) + 1;
    const long long want = (long long)(c_hi - c_lo) * GPB_NB;
    const int cols = mode ? (int)std::min(avail, want) : GPB_NB;
    if (cols <= 0) return cudaSuccess;
    carried_row_kernel<<<dim3(mode ? (cols + 255) / 256 : 1, 1, B), 256, 0, st>>>(dm, kp, kb, c_lo, c_hi, mode);
    ++g_launches;
    return cudaGetLastError();
  };
  auto panel = [&](int k) -> cudaError_t {
    const int rows = nrows - (k + 1) * GPB_NB;
    if (small(rows)) {
      GPB_CK((launch_cfg<CfgQuarter, false, false>(GeoPanel{dm, k, sep}, dim3((rows + 31) / 32, 1, B), ms, false, "panel_q", k)));
    } else {
      const int Tm = (rows + BM - 1) / BM;
      GPB_CK((launch_cfg<Cfg, false, false>(GeoPanel{dm, k, sep}, dim3(Tm, 1, B), ms, false, "panel", k)));
    }
    return carried(k, 1, 0, 0, 0, ms);
  };
  auto syrk = [&](int kp, int kb, int c_lo, int c_hi, cudaStream_t st, bool bulk, const char* tag) -> cudaError_t {
    const int rows = nrows - (kp + kb) * GPB_NB;
    if (rows <= 0) return cudaSuccess;
    if (!bulk && small((long long)rows * (c_hi - c_lo))) {
      const int Tq = (rows + 31) / 32;
      GPB_CK((launch_cfg<CfgQuarter, false, false>(GeoSyrk{dm, kp, kb, c_lo, c_hi, sep},
                                                   dim3((unsigned)tri_count(Tq, 4, c_lo, c_hi), 1, B), st, false, tag, kp)));
    } else {
      const int Tm = (rows + BM - 1) / BM;
      GPB_CK((launch_cfg<Cfg, false, false>(GeoSyrk{dm, kp, kb, c_lo, c_hi, sep, (bulk && grouped_raster()) ? 1 : 0},
                                            dim3((unsigned)tri_count(Tm, R, c_lo, c_hi), 1, B), st, bulk, tag, kp)));
    }
    return carried(kp, kb, c_lo, c_hi, 1, st);
  };
  // Look-ahead of depth 2 over OUTER steps s (W = 1 or 2 block columns each).  The update of outer step s is split by
  // target columns (128-wide tile columns of the trailing matrix):
  //   A(s) = [0, W)    the next outer panel            critical stream, right behind the factorisation of step s
  //   B(s) = [W, 2W)   the second-next outer panel     medium-priority stream
  //   C(s) = [2W, ..)  the bulk                        low-priority stream
  // A column receives the updates of the outer steps in order (C(s-2), B(s-1), A(s) - bitwise reproducible):
  //   A(s) waits for B(s-1) (event ev_b), B(s) for all of C(s-1) (ev_d), C(s) follows C(s-1) on its stream; B(s) and C(s)
  // wait for the panels of step s (ev_e).  The factorisation of step s+1 therefore depends on C(s-2), not on C(s-1): the
  // critical path runs up to two outer steps ahead of the bulk, and the low-priority stream runs bulk update after bulk
  // update with nothing in between (with depth 1 the small "next columns" launch sat between two bulk updates on the same
  // stream: 48 of 390 us per outer step at n = 8192, one partially filled wave).
  // widths of the outer steps (decided up front: the split of an update follows the NEXT two outer panels)
  std::vector<int> width;
  for (int k = 0; k < nblk;) {
    if (!full_with_rows(k)) { ++k; continue; }
    const int W = B >= 8 ? kb_for(n_max, B) : kb_for((long long)n_max - (long long)k * GPB_NB, B);
    int kb = 1;
    while (kb < W && full_with_rows(k + kb)) ++kb;
    width.push_back(kb);
    k += kb;
  }
  int step = 0;
  for (int k = 0; k < nblk;) {
    GPB_CK(diag(k));
    if (!full_with_rows(k)) { GPB_CK(inverse_progress(k + 1)); ++k; continue; }
    GPB_CK(panel(k));
    const int W = width[step];
    int kb = 1;
    // inside the outer panel the block columns are factorised left-looking: column k + kb receives the kb finished
    // panels in one rank-(128 kb) strip update, then its diagonal block and panel follow
    while (kb < W) {
      GPB_CK(syrk(k, kb, 0, 1, ms, false, "strip"));
      GPB_CK(diag(k + kb));
      GPB_CK(panel(k + kb));
      ++kb;
    }
    GPB_CK(inverse_progress(k + kb));
    const int rows = nrows - (k + kb) * GPB_NB;
    const int Tn = (rows + GPB_NB - 1) / GPB_NB;
    if (!lookahead) {
      GPB_CK(syrk(k, kb, 0, Tn, ms, true, "bulk"));
    } else {
      // columns of the next outer panel / of the second-next one; the last block column (partial or carried row) has no
      // outer step of its own and rides with whatever range reaches it
      const int W1 = step + 1 < (int)width.size() ? width[step + 1] : Tn;
      const int W2 = step + 2 < (int)width.size() ? width[step + 2] : Tn;
      GPB_CK(cudaEventRecord(ex.ev_e[step & 1], ms));
      if (step > 0) GPB_CK(cudaStreamWaitEvent(ms, ex.ev_b[(step - 1) & 1], 0));
      GPB_CK(syrk(k, kb, 0, W1, ms, false, "A"));
      if (Tn > W1) {
        GPB_CK(cudaStreamWaitEvent(ex.mid, ex.ev_e[step & 1], 0));
        if (step > 0) GPB_CK(cudaStreamWaitEvent(ex.mid, ex.ev_d[(step - 1) & 1], 0));
        GPB_CK(syrk(k, kb, W1, W1 + W2, ex.mid, false, "B"));
      }
      GPB_CK(cudaEventRecord(ex.ev_b[step & 1], ex.mid));
      if (Tn > W1 + W2) {
        GPB_CK(cudaStreamWaitEvent(ex.side, ex.ev_e[step & 1], 0));
        GPB_CK(syrk(k, kb, W1 + W2, Tn, ex.side, true, "bulk"));
      }
      GPB_CK(cudaEventRecord(ex.ev_d[step & 1], ex.side));
    }
    k += kb;
    ++step;
  }
  if (fuse) {
    GPB_CK(inverse_progress(nblk));
    if (!sched.finished()) return cudaErrorUnknown;      // the task graph must drain once every column is final
    GPB_CK(cudaEventRecord(ex.ev_join[2], ex.inv));
    GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[2], 0));
  }
  if (lookahead) {
    GPB_CK(cudaEventRecord(ex.ev_join[0], ex.crit));
    GPB_CK(cudaEventRecord(ex.ev_join[1], ex.side));
    GPB_CK(cudaEventRecord(ex.ev_join[3], ex.mid));
    GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[0], 0));
    GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[1], 0));
    GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[3], 0));
  }
  return cudaSuccess;
}

cudaError_t run_potrf(const GpbMat* dm, int B, int n_max, int aug, bool lookahead, bool with_trtri, const Exec& ex) {
  return cfg_half() ? potrf_impl<CfgHalf>(dm, B, n_max, aug, lookahead, with_trtri, ex)
                    : potrf_impl<CfgBig>(dm, B, n_max, aug, lookahead, with_trtri, ex);
}

cudaError_t run_diag(const GpbMat* dm, int B, int k, cudaStream_t s) {
  TraceSpan span("diag", s, k, B);
  launch_diag(dm, B, k, s);
  ++g_launches;
  return cudaGetLastError();
}

cudaError_t run_finalize(const GpbMat* dm, int B, double log2pi, cudaStream_t s) {
  finalize_kernel<<<B, 256, 0, s>>>(dm, log2pi);
  ++g_launches;
  return cudaGetLastError();
}

template <class Cfg>
static cudaError_t trtri_impl(const GpbMat* dm, int B, int n_max, cudaStream_t s) {
  const int nblk = (n_max + GPB_NB - 1) / GPB_NB;
  diag_copy_kernel<<<dim3(nblk, B), 1024, 0, s>>>(dm, 0);
  ++g_launches;
  GPB_CK(cudaGetLastError());
  for (long long sz = GPB_NB; sz < n_max; sz *= 2) {
    const int tiles = (int)(sz / Cfg::BM) * (int)(sz / Cfg::BN);
    const int nsub = (int)((n_max + 2 * sz - 1) / (2 * sz));
    GPB_CK((launch_cfg<Cfg, false, true>(GeoTrtriT{dm, (int)sz, 0}, dim3(tiles, nsub, B), s)));
    GPB_CK((launch_cfg<Cfg, false, true>(GeoTrtriW{dm, (int)sz, 0}, dim3(tiles, nsub, B), s)));
  }
  return cudaSuccess;
}
cudaError_t run_trtri(const GpbMat* dm, int B, int n_max, cudaStream_t s) {
  return cfg_half() ? trtri_impl<CfgHalf>(dm, B, n_max, s) : trtri_impl<CfgBig>(dm, B, n_max, s);
}

cudaError_t run_alpha(const GpbMat* dm, int B, int n_max, cudaStream_t s) {
  alpha_kernel<<<dim3((n_max + 7) / 8, B), 256, 0, s>>>(dm);
  ++g_launches;
  return cudaGetLastError();
}

template <class Cfg>
static cudaError_t lauum_impl(const GpbMat* dm, int B, int n_max, cudaStream_t s) {
  const int Tm = (n_max + Cfg::BM - 1) / Cfg::BM;
  return launch_cfg<Cfg, true, true>(GeoLauum{dm}, dim3((unsigned)tri_count(Tm, Cfg::BN / Cfg::BM, 0, Tm), 1, B), s);
}
cudaError_t run_lauum(const GpbMat* dm, int B, int n_max, cudaStream_t s) {
  return cfg_half() ? lauum_impl<CfgHalf>(dm, B, n_max, s) : lauum_impl<CfgBig>(dm, B, n_max, s);
}

cudaError_t run_trsv(const GpbMat* dm, int B, int n_max, int transposed, cudaStream_t s) {
  const int nblk = (n_max + GPB_NB - 1) / GPB_NB;
  trsv_init_kernel<<<dim3((n_max + 255) / 256, B), 256, 0, s>>>(dm, transposed);
  ++g_launches;
  GPB_CK(cudaGetLastError());
  if (!transposed) {
    for (int k = 0; k < nblk; ++k) {
      const int rest = n_max - (k + 1) * GPB_NB;
      const int gx = rest > 0 ? (rest + 255) / 256 : 1;
      trsv_step_kernel<<<dim3(gx, B), 256, 0, s>>>(dm, k, 0);
      ++g_launches;
      GPB_CK(cudaGetLastError());
    }
  } else {
    for (int k = nblk - 1; k >= 0; --k) {
      const int cols = k * GPB_NB;
      const int gx = cols > 0 ? (cols + 7) / 8 : 1;
      trsv_step_kernel<<<dim3(gx, B), 256, 0, s>>>(dm, k, 1);
      ++g_launches;
      GPB_CK(cudaGetLastError());
    }
  }
  return cudaSuccess;
}

cudaError_t run_zero_upper(double* A, int n, int ld, cudaStream_t s) {
  if (n <= 1) return cudaSuccess;
  zero_upper_kernel<<<dim3((n + 255) / 256, n), 256, 0, s>>>(A, n, ld);
  ++g_launches;
  return cudaGetLastError();
}
cudaError_t run_symmetrize(double* A, int n, int ld, cudaStream_t s) {
  const int t = (n + 31) / 32;
  symmetrize_kernel<<<dim3(t, t), dim3(32, 8), 0, s>>>(A, n, ld);
  ++g_launches;
  return cudaGetLastError();
}

// FP64 pipe probes: kind 0 = DMMA.8x8x4 chains, kind 1 = DFMA chains (register operands only)
__global__ void __launch_bounds__(256) microbench_kernel(int kind, int iters, double* out) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  if (kind == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}
cudaError_t run_microbench(int kind, int iters, int blocks, cudaStream_t s) {
  static double* sink = nullptr;
  if (!sink) GPB_CK(cudaMalloc(&sink, 64));
  microbench_kernel<<<blocks, 256, 0, s>>>(kind, iters, sink);
  ++g_launches;
  return cudaGetLastError();
}

template <class Cfg>
static cudaError_t plain_impl(const GeoPlain& g, int M, int N, cudaStream_t s) {
  dim3 grid((M + Cfg::BM - 1) / Cfg::BM, (N + Cfg::BN - 1) / Cfg::BN, 1);
  if (!g.akm && !g.bkm) return launch_cfg<Cfg, false, false>(g, grid, s);
  if (!g.akm && g.bkm) return launch_cfg<Cfg, false, true>(g, grid, s);
  if (g.akm && !g.bkm) return launch_cfg<Cfg, true, false>(g, grid, s);
  return launch_cfg<Cfg, true, true>(g, grid, s);
}
cudaError_t run_gemm_plain(int akm, int bkm, const double* A, int lda, const double* Bm, int ldb, double* C, int ldc,
                           int M, int N, int K, double alpha, double beta, cudaStream_t s) {
  GeoPlain g{A, Bm, C, lda, ldb, ldc, M, N, K, akm & 1, bkm & 1, alpha, beta};
  // akm bits 1-2 select the tile configuration explicitly (tests / probes): +0 driver default, +2 big, +4 half
  const int sel = akm >> 1;
  const bool half = sel == 0 ? cfg_half() : (sel == 2);
  return half ? plain_impl<CfgHalf>(g, M, N, s) : plain_impl<CfgBig>(g, M, N, s);
}

}  // namespace gpb
