// Device-side ABI shared by the ahead-of-time kernels of libgpb and by the kernels that are specialised per kernel program
// at run time (jit.cu compiles THIS TEXT, embedded into the library at build time, with NVRTC): the per-GP descriptor and
// the tile enumeration of a lower triangle.  It must stay free of host headers (NVRTC has none): fixed-width integers
// come from the compiler's built-in types.
#pragma once
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>
#include <stdint.h>
#else
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
#endif

#define GPB_NB 128  // block size of every blocked factorisation step (panel width, tile edge)

// Per-matrix descriptor, resident in device memory for the lifetime of a plan.  One entry per GP of a batch
// (a holistic GP is a batch of one; a PartitionedGaussianProcess is a batch of its blocks).
struct GpbMat {
  double* A;         // (n+aug) x ld, column-major, lower triangle: K+s2*I -> L -> inv(L); row n holds y^T -> z^T
  double* Kinv;      // n x ld, column-major: scratch for the triangular inverse, then inv(K) (lower)
  double* Wd;        // nblk x 128 x 128: inverses of the diagonal blocks of L (lower, zero above the diagonal)
  double* part;      // nblk partial sums of log(diag L)
  double* gpart;     // per-CTA partial gradient sums [n_gtiles x (n_hp+1)]
  double* alpha;     // [n]  inv(K) y
  double* zvec;      // [n]  inv(L) y
  double* tmpv;      // [n]  scratch right-hand side of the standalone triangular solves
  const double* X;   // [n x dim] row-major inputs (device)
  const double* y;   // [n] detrended targets (device)
  const double* hp;  // [n_hp] flat hyper-parameters (device)
  const double* noise;  // device scalar s2
  const int32_t* code;  // postfix program (device)
  double* nll;       // device scalar out
  double* grad;      // [n_hp+1] out (last entry: d nll / d s2)
  int* info;         // device scalar out: 0 ok, j>0 first non-positive pivot (1-based)
  double* terms;     // [2] out: y^T K^-1 y and sum(log diag L), the two data-dependent terms of the NLL
  // weights of the two terms in the gradient: d/dtheta [gw_quad * 1/2 y^T K^-1 y + gw_logdet * sum(log diag L)]
  // (1, 1 = the NLL; the rank-3 batch aggregate of Metrics/LogLikelihood.py:62-63 uses 1/B and 1)
  double gw_quad, gw_logdet;
  int n, ld, dim, n_ops, n_hp, aug, cp_mode, n_gtiles;
  // distributed plans (dist.cu): block (I, J) of 128 x 128 is owned by process (I mod own_P, (J / own_W) mod own_Q) -
  // block columns are dealt out in groups of own_W (the outer panel of the factorisation); own_P == 0: all
  int own_P, own_Q, own_p, own_q, own_W;
  // own_compact != 0 (column storage, 1 x Q grids): A holds ONLY the block columns this process owns, packed in ascending
  // order - global column j lives in local column gpb_local_col(j) - instead of the full (n + 1) x ld matrix
  int own_compact;
  // gradient stages of a distributed plan: block column J belongs to rank J mod col_world (col_world == 0: all)
  int col_world, col_rank;
};

// Block columns owned by process column q when they are dealt out in groups of W over Q process columns: how many of
// them lie below block column J, and the t-th one (ascending).
__host__ __device__ inline int gpb_owned_cols_below(int J, int Q, int q, int W) {
  const int g = J / W, r = J % W;
  const int groups = g > q ? (g - q + Q - 1) / Q : 0;       // owned groups q, q + Q, ... that end before group g
  return groups * W + ((g % Q == q) ? r : 0);
}
__host__ __device__ inline int gpb_owned_col_at(int t, int Q, int q, int W) { return ((t / W) * Q + q) * W + t % W; }
// local column (elements) of global column j of a column-storage plan; j must lie in an owned block column
__host__ __device__ inline long long gpb_local_col(long long j, int Q, int q, int W) {
  return (long long)gpb_owned_cols_below((int)(j / GPB_NB), Q, q, W) * GPB_NB + j % GPB_NB;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

namespace gpb {

// Enumeration of the tiles of a lower-triangular region cut into (BM-row x BN-column) tiles, R = BN / BM:
// column tile c (units of BN) owns the row tiles ti >= R*c (units of BM), ti < Tm.  Columns restricted to [c_lo, c_hi).
__host__ __device__ inline long long tri_count(int Tm, int R, int c_lo, int c_hi) {
  const int Tn = (Tm + R - 1) / R;   // column tiles that own at least one row tile
  if (c_hi > Tn) c_hi = Tn;
  if (c_hi <= c_lo) return 0;
  const long long w = c_hi - c_lo, Tp = Tm - (long long)R * c_lo;
  return w * Tp - (long long)R * w * (w - 1) / 2;
}
// 32-bit / single-precision fast path (every launch of the path: Tm <= 16384, so all counts fit in 31 bits): the
// persistent kernel evaluates this twice per tile, so it must stay a few dozen instructions.
__device__ __forceinline__ bool tri_map(unsigned idx, int Tm, int R, int c_lo, int c_hi, int& ti, int& tj) {
  const int Tn = (Tm + R - 1) / R;
  if (c_hi > Tn) c_hi = Tn;
  if (c_hi <= c_lo) return false;
  const int w = c_hi - c_lo, Tp = Tm - R * c_lo;
  const int cnt = w * Tp - R * (w * (w - 1) / 2);
  if (idx >= (unsigned)cnt) return false;
  const float bq = (float)Tp + 0.5f * (float)R;
  int c = (int)floorf((bq - sqrtf(fmaxf(bq * bq - 2.0f * (float)R * (float)idx, 0.0f))) / (float)R);
  c = max(0, min(c, w - 1));
  // prefix(c) = c * Tp - R * c (c - 1) / 2 tiles precede column c
  while (c > 0 && c * Tp - R * (c * (c - 1) / 2) > (int)idx) --c;
  while ((c + 1) * Tp - R * ((c + 1) * c / 2) <= (int)idx) ++c;
  const int off = (int)idx - (c * Tp - R * (c * (c - 1) / 2));
  tj = c + c_lo;
  ti = R * tj + off;
  return true;
}

}  // namespace gpb
