// FP64 tensor-core GEMM mainloop shared by every dense contraction on the path:
//   Cholesky trailing update / panel solve   (replaces tf.linalg.cholesky, Statistics/CovarianceMatrix.py:250,475)
//   triangular inverse and inv(K) = W^T W   (replaces the Cholesky backward pass TF runs for Optimizer/Fitter.py:124-158)
//
// C[128x128 tile] = alpha * op(A) * op(B)^T + beta * C, FP64, DMMA.8x8x4 (mma.sync m8n8k4.f64).
// Operands are staged with 16-byte cp.async (zero-filling out-of-range rows / k) into padded, bank-conflict-free
// shared-memory tiles, 3 stages deep.  8 warps as 4 (m) x 2 (n); each warp owns a 32 x 64 accumulator block
// (32 DMMA tiles, 64 FP64 accumulators per thread).
//
// An operand is "MN-major" when element (mn, k) lives at P[mn + k*ld] (column-major op(A)=A) and "K-major" when it
// lives at P[k + mn*ld] (column-major op(A)=A^T); both are supported for A and B so that NT / NN / TN products of
// column-major matrices need no transposes in memory.
#pragma once
#include "common.cuh"

namespace gpb {

constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_STAGES = 3, G_THREADS = 256;
constexpr int G_LDMN = G_BM + 4;   // pitch (doubles) of an MN-major tile  [BK][BM+4]
constexpr int G_LDK = G_BK + 4;    // pitch (doubles) of a K-major tile    [BM][BK+4]
constexpr int G_TILE = G_BM * G_LDK;  // 2560 doubles >= BK*G_LDMN = 2112
constexpr int G_SMEM_BYTES = G_STAGES * 2 * G_TILE * (int)sizeof(double);  // 122880

struct TileJob {
  const double* A;  // tile-row origin of op(A): MN-major -> &A[i0], K-major -> &A[i0*lda]
  const double* B;  // tile-col origin of op(B)
  double* C;        // &C[i0 + j0*ldc]
  int lda, ldb, ldc;
  int mrem, nrem;   // valid rows / cols of this tile (<=128)
  int klo, khi;     // contraction range in operand coordinates
  double alpha, beta;
};

template <bool KM>
__device__ __forceinline__ void load_tile(double* s, const double* g, int ld, int rem, int k0, int khi, int tid) {
  if (!KM) {
    const int mn = (tid & 63) * 2;
    const int kb = tid >> 6;
    int vm = rem - mn;
    vm = vm < 0 ? 0 : (vm > 2 ? 2 : vm);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = kb + 4 * q;
      const int valid = (k0 + k < khi) ? vm : 0;
      const double* src = valid ? g + mn + (size_t)(k0 + k) * ld : g;
      cp_async16(s + k * G_LDMN + mn, src, valid * 8);
    }
  } else {
    const int k = (tid & 7) * 2;
    const int mb = tid >> 3;
    int vk = khi - (k0 + k);
    vk = vk < 0 ? 0 : (vk > 2 ? 2 : vk);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int mn = mb + 32 * q;
      const int valid = (mn < rem) ? vk : 0;
      const double* src = valid ? g + (k0 + k) + (size_t)mn * ld : g;
      cp_async16(s + mn * G_LDK + k, src, valid * 8);
    }
  }
}

template <bool AKM, bool BKM>
__device__ __forceinline__ void gemm_tile(const TileJob& J) {
  extern __shared__ __align__(16) double gsm[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int lr = lane >> 2, lk = lane & 3;

  double acc[4][8][2];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int g = 0; g < 8; ++g) { acc[f][g][0] = 0.0; acc[f][g][1] = 0.0; }

  const int nk = (J.khi > J.klo) ? (J.khi - J.klo + G_BK - 1) / G_BK : 0;

  if (J.beta != 0.0) {
    // the accumulate-into tile is needed only by the epilogue: pull it into L2 now (128 columns x 8 lines)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      const int col = idx >> 3, seg = (idx & 7) * 16;
      if (col < J.nrem && seg < J.mrem) {
        const double* pp = J.C + (size_t)col * J.ldc + seg;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
      }
    }
  }

#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < nk) {
      double* As = gsm + s * 2 * G_TILE;
      load_tile<AKM>(As, J.A, J.lda, J.mrem, J.klo + s * G_BK, J.khi, tid);
      load_tile<BKM>(As + G_TILE, J.B, J.ldb, J.nrem, J.klo + s * G_BK, J.khi, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      const int nt = kt + G_STAGES - 1;
      if (nt < nk) {
        double* Ns = gsm + (nt % G_STAGES) * 2 * G_TILE;
        load_tile<AKM>(Ns, J.A, J.lda, J.mrem, J.klo + nt * G_BK, J.khi, tid);
        load_tile<BKM>(Ns + G_TILE, J.B, J.ldb, J.nrem, J.klo + nt * G_BK, J.khi, tid);
      }
      cp_async_commit();
    }
    const double* As = gsm + (kt % G_STAGES) * 2 * G_TILE;
    const double* Bs = As + G_TILE;
#pragma unroll
    for (int kk = 0; kk < G_BK / 4; ++kk) {
      const int kidx = kk * 4 + lk;
      double a[4], b[8];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int row = wm * 32 + f * 8 + lr;
        a[f] = AKM ? As[row * G_LDK + kidx] : As[kidx * G_LDMN + row];
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int col = wn * 64 + g * 8 + lr;
        b[g] = BKM ? Bs[col * G_LDK + kidx] : Bs[kidx * G_LDMN + col];
      }
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int g = 0; g < 8; ++g) dmma884(acc[f][g][0], acc[f][g][1], a[f], b[g]);
    }
  }
  cp_async_wait<0>();

  // epilogue: each quad-row of 8 lanes covers 8 consecutive rows (64 B) of one column.  The C tile is read in
  // batches of 16 values per thread (all loads of a batch issued before the first store) so that the read-modify-write
  // costs 4 memory round trips per tile, not 64.
  const double alpha = J.alpha, beta = J.beta;
#pragma unroll
  for (int gp = 0; gp < 4; ++gp) {
    double cv[2][2][4];
    if (beta != 0.0) {
#pragma unroll
      for (int gg = 0; gg < 2; ++gg)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = wn * 64 + (2 * gp + gg) * 8 + 2 * lk + e;
          const double* cp = J.C + (size_t)col * J.ldc;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int row = wm * 32 + f * 8 + lr;
            cv[gg][e][f] = (col < J.nrem && row < J.mrem) ? cp[row] : 0.0;
          }
        }
    }
#pragma unroll
    for (int gg = 0; gg < 2; ++gg)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int g = 2 * gp + gg;
        const int col = wn * 64 + g * 8 + 2 * lk + e;
        if (col < J.nrem) {
          double* cp = J.C + (size_t)col * J.ldc;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int row = wm * 32 + f * 8 + lr;
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

// linear index -> (ti, tj) over the lower-triangular tile set {tj in [c_lo, c_hi), ti in [tj, T)}
__host__ __device__ inline long long tri_count(int T, int c_lo, int c_hi) {
  if (c_hi > T) c_hi = T;
  if (c_hi <= c_lo) return 0;
  const long long w = c_hi - c_lo, Tp = T - c_lo;
  return w * Tp - w * (w - 1) / 2;
}
__device__ __forceinline__ bool tri_map(long long idx, int T, int c_lo, int c_hi, int& ti, int& tj) {
  if (idx >= tri_count(T, c_lo, c_hi)) return false;
  const double Tp = (double)(T - c_lo);
  const double b = 2.0 * Tp + 1.0;
  long long c = (long long)floor((b - sqrt(b * b - 8.0 * (double)idx)) * 0.5);
  if (c < 0) c = 0;
  const long long Tpi = T - c_lo;
  while (c > 0 && c * Tpi - c * (c - 1) / 2 > idx) --c;
  while ((c + 1) * Tpi - (c + 1) * c / 2 <= idx) ++c;
  const long long off = idx - (c * Tpi - c * (c - 1) / 2);
  tj = (int)c + c_lo;
  ti = tj + (int)off;
  return true;
}

template <bool AKM, bool BKM, class Geo>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_kernel(const Geo geo) {
  TileJob J;
  if (!geo(J)) return;
  gemm_tile<AKM, BKM>(J);
}

}  // namespace gpb
