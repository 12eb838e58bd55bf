// FP64 tensor-core GEMM mainloop shared by every dense contraction on the path:
//   Cholesky trailing update / panel solve   (replaces tf.linalg.cholesky, Statistics/CovarianceMatrix.py:250,475)
//   triangular inverse and inv(K) = W^T W   (replaces the Cholesky backward pass TF runs for Optimizer/Fitter.py:124-158)
//
// C[BM x BN tile] = alpha * op(A) * op(B)^T + beta * C, FP64, DMMA.8x8x4 (mma.sync m8n8k4.f64).
// Operands are staged with 16-byte cp.async (zero-filling out-of-range rows / k) into padded, bank-conflict-free
// shared-memory tiles, 3 stages deep.  Two tile configurations share the code:
//   CfgBig   128 x 128, 8 warps (4 x 2), warp tile 32 x 64, one CTA per SM
//   CfgHalf   64 x 128, 4 warps (2 x 2), warp tile 32 x 64, two CTAs per SM (the second CTA hides the first one's
//             barrier / prologue / epilogue bubbles; twice as many CTAs for the single-wave launches of the
//             factorisation's critical path)
//
// An operand is "MN-major" when element (mn, k) lives at P[mn + k*ld] (column-major op(A)=A) and "K-major" when it
// lives at P[k + mn*ld] (column-major op(A)=A^T); both are supported for A and B so that NT / NN / TN products of
// column-major matrices need no transposes in memory.
#pragma once
#include "common.cuh"

namespace gpb {

constexpr int G_BK = 16, G_STAGES = 3;
constexpr int G_LDK = G_BK + 4;    // pitch (doubles) of a K-major tile    [rows][BK+4]

template <int BM_, int BN_, int WARPS_M_, int WARPS_N_, int MIN_CTAS_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, MIN_CTAS = MIN_CTAS_;
  static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;   // warp tile
  static constexpr int FM = WM / 8, FN = WN / 8;               // 8x8 DMMA tiles per warp
  static constexpr int A_TILE = BM * G_LDK > G_BK * (BM + 4) ? BM * G_LDK : G_BK * (BM + 4);
  static constexpr int B_TILE = BN * G_LDK > G_BK * (BN + 4) ? BN * G_LDK : G_BK * (BN + 4);
  static constexpr int SMEM_BYTES = G_STAGES * (A_TILE + B_TILE) * (int)sizeof(double);
};
using CfgBig = GemmCfg<128, 128, 4, 2, 1>;
using CfgHalf = GemmCfg<64, 128, 2, 2, 2>;

struct TileJob {
  const double* A;  // tile-row origin of op(A): MN-major -> &A[i0], K-major -> &A[i0*lda]
  const double* B;  // tile-col origin of op(B)
  double* C;        // &C[i0 + j0*ldc]
  int lda, ldb, ldc;
  int mrem, nrem;   // valid rows / cols of this tile
  int klo, khi;     // contraction range in operand coordinates
  double alpha, beta;
};

// one BK-deep slab of an operand tile with ROWS rows (m or n), by THREADS threads
template <bool KM, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(double* s, const double* g, int ld, int rem, int k0, int khi, int tid) {
  if (!KM) {
    constexpr int PAIRS = ROWS / 2;              // 16-byte chunks per k row
    constexpr int KSTEP = THREADS / PAIRS;       // k rows covered per pass
    const int mn = (tid % PAIRS) * 2;
    const int kb = tid / PAIRS;
    int vm = rem - mn;
    vm = vm < 0 ? 0 : (vm > 2 ? 2 : vm);
#pragma unroll
    for (int q = 0; q < G_BK / KSTEP; ++q) {
      const int k = kb + KSTEP * q;
      const int valid = (k0 + k < khi) ? vm : 0;
      const double* src = valid ? g + mn + (size_t)(k0 + k) * ld : g;
      cp_async16(s + k * (ROWS + 4) + mn, src, valid * 8);
    }
  } else {
    constexpr int RSTEP = THREADS / 8;
    const int k = (tid & 7) * 2;
    const int mb = tid >> 3;
    int vk = khi - (k0 + k);
    vk = vk < 0 ? 0 : (vk > 2 ? 2 : vk);
#pragma unroll
    for (int q = 0; q < ROWS / RSTEP; ++q) {
      const int mn = mb + RSTEP * q;
      const int valid = (mn < rem) ? vk : 0;
      const double* src = valid ? g + (k0 + k) + (size_t)mn * ld : g;
      cp_async16(s + mn * G_LDK + k, src, valid * 8);
    }
  }
}

template <class Cfg, bool AKM, bool BKM>
__device__ __forceinline__ void gemm_tile(const TileJob& J) {
  extern __shared__ __align__(16) double gsm[];
  constexpr int BM = Cfg::BM, BN = Cfg::BN, T = Cfg::THREADS, FM = Cfg::FM, FN = Cfg::FN;
  constexpr int STAGE = Cfg::A_TILE + Cfg::B_TILE;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
  const int lr = lane >> 2, lk = lane & 3;

  double acc[FM][FN][2];
#pragma unroll
  for (int f = 0; f < FM; ++f)
#pragma unroll
    for (int g = 0; g < FN; ++g) { acc[f][g][0] = 0.0; acc[f][g][1] = 0.0; }

  const int nk = (J.khi > J.klo) ? (J.khi - J.klo + G_BK - 1) / G_BK : 0;

  if (J.beta != 0.0) {
    // the accumulate-into tile is needed only by the epilogue: pull it into L2 now (BN columns x BM/16 lines)
    constexpr int SEGS = BM / 16;
#pragma unroll
    for (int q = 0; q < (BN * SEGS + T - 1) / T; ++q) {
      const int idx = tid + T * q;
      const int col = idx / SEGS, seg = (idx % SEGS) * 16;
      if (col < J.nrem && seg < J.mrem) {
        const double* pp = J.C + (size_t)col * J.ldc + seg;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
      }
    }
  }

#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < nk) {
      double* As = gsm + s * STAGE;
      load_tile<AKM, BM, T>(As, J.A, J.lda, J.mrem, J.klo + s * G_BK, J.khi, tid);
      load_tile<BKM, BN, T>(As + Cfg::A_TILE, J.B, J.ldb, J.nrem, J.klo + s * G_BK, J.khi, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      const int nt = kt + G_STAGES - 1;
      if (nt < nk) {
        double* Ns = gsm + (nt % G_STAGES) * STAGE;
        load_tile<AKM, BM, T>(Ns, J.A, J.lda, J.mrem, J.klo + nt * G_BK, J.khi, tid);
        load_tile<BKM, BN, T>(Ns + Cfg::A_TILE, J.B, J.ldb, J.nrem, J.klo + nt * G_BK, J.khi, tid);
      }
      cp_async_commit();
    }
    const double* As = gsm + (kt % G_STAGES) * STAGE;
    const double* Bs = As + Cfg::A_TILE;
#pragma unroll
    for (int kk = 0; kk < G_BK / 4; ++kk) {
      const int kidx = kk * 4 + lk;
      double a[FM], b[FN];
#pragma unroll
      for (int f = 0; f < FM; ++f) {
        const int row = wm * Cfg::WM + f * 8 + lr;
        a[f] = AKM ? As[row * G_LDK + kidx] : As[kidx * (BM + 4) + row];
      }
#pragma unroll
      for (int g = 0; g < FN; ++g) {
        const int col = wn * Cfg::WN + g * 8 + lr;
        b[g] = BKM ? Bs[col * G_LDK + kidx] : Bs[kidx * (BN + 4) + col];
      }
#pragma unroll
      for (int f = 0; f < FM; ++f)
#pragma unroll
        for (int g = 0; g < FN; ++g) dmma884(acc[f][g][0], acc[f][g][1], a[f], b[g]);
    }
  }
  cp_async_wait<0>();

  // epilogue: each quad-row of 8 lanes covers 8 consecutive rows (64 B) of one column.  The C tile is read in
  // batches (all loads of a batch issued before the first store) so that the read-modify-write costs FN/2 memory
  // round trips per tile, not FM*FN*2.
  const double alpha = J.alpha, beta = J.beta;
#pragma unroll
  for (int gp = 0; gp < FN / 2; ++gp) {
    double cv[2][2][FM];
    if (beta != 0.0) {
#pragma unroll
      for (int gg = 0; gg < 2; ++gg)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = wn * Cfg::WN + (2 * gp + gg) * 8 + 2 * lk + e;
          const double* cp = J.C + (size_t)col * J.ldc;
#pragma unroll
          for (int f = 0; f < FM; ++f) {
            const int row = wm * Cfg::WM + f * 8 + lr;
            cv[gg][e][f] = (col < J.nrem && row < J.mrem) ? cp[row] : 0.0;
          }
        }
    }
#pragma unroll
    for (int gg = 0; gg < 2; ++gg)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int g = 2 * gp + gg;
        const int col = wn * Cfg::WN + g * 8 + 2 * lk + e;
        if (col < J.nrem) {
          double* cp = J.C + (size_t)col * J.ldc;
#pragma unroll
          for (int f = 0; f < FM; ++f) {
            const int row = wm * Cfg::WM + f * 8 + lr;
            if (row < J.mrem) {
              double v = alpha * acc[f][g][e];
              if (beta != 0.0) v += beta * cv[gg][e][f];
              cp[row] = v;
            }
          }
        }
      }
  }
}

// Enumeration of the tiles of a lower-triangular region cut into (BM-row x BN-column) tiles, R = BN / BM:
// column tile c (units of BN) owns the row tiles ti >= R*c (units of BM), ti < Tm.  Columns restricted to [c_lo, c_hi).
__host__ __device__ inline long long tri_count(int Tm, int R, int c_lo, int c_hi) {
  const int Tn = (Tm + R - 1) / R;   // column tiles that own at least one row tile
  if (c_hi > Tn) c_hi = Tn;
  if (c_hi <= c_lo) return 0;
  const long long w = c_hi - c_lo, Tp = Tm - (long long)R * c_lo;
  return w * Tp - (long long)R * w * (w - 1) / 2;
}
__device__ __forceinline__ bool tri_map(long long idx, int Tm, int R, int c_lo, int c_hi, int& ti, int& tj) {
  if (idx >= tri_count(Tm, R, c_lo, c_hi)) return false;
  const long long Tp = Tm - (long long)R * c_lo;
  const double b = (double)Tp + 0.5 * R;
  long long c = (long long)floor((b - sqrt(b * b - 2.0 * R * (double)idx)) / (double)R);
  if (c < 0) c = 0;
  while (c > 0 && c * Tp - (long long)R * c * (c - 1) / 2 > idx) --c;
  while ((c + 1) * Tp - (long long)R * (c + 1) * c / 2 <= idx) ++c;
  const long long off = idx - (c * Tp - (long long)R * c * (c - 1) / 2);
  tj = (int)c + c_lo;
  ti = R * tj + (int)off;
  return true;
}

template <class Cfg, bool AKM, bool BKM, class Geo>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) gemm_kernel(const Geo geo) {
  TileJob J;
  if (!geo.template tile<Cfg::BM, Cfg::BN>(J)) return;
  gemm_tile<Cfg, AKM, BKM>(J);
}

}  // namespace gpb
