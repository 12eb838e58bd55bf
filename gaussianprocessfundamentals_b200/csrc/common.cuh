// Shared device-side definitions for the gpbasics hot path on sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
  // distributed plans (dist.cu): block (I, J) of 128 x 128 is owned by process (I mod own_P, J mod own_Q); own_P == 0: all
  int own_P, own_Q, own_p, own_q;
  // gradient stages of a distributed plan: block column J belongs to rank J mod col_world (col_world == 0: all)
  int col_world, col_rank;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// FP64 tensor-core MMA (SASS: DMMA.8x8x4).  A: row (lane>>2), k (lane&3).  B: k (lane&3), col (lane>>2).
// C: row (lane>>2), cols 2*(lane&3) + {0,1}.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
