// Covariance assembly and trace gradient SPECIALISED per kernel program.
//
// jit.cu turns a postfix kernel program (program.cuh) into a struct `Prog` of straight-line device code - every stack
// slot, tape entry and adjoint of the interpreter becomes a named scalar, so the whole evaluation lives in registers
// (no local memory, no opcode dispatch) - and compiles THIS TEXT plus that struct with NVRTC.  It replaces the same
// reference code as the interpreter kernels of assemble.cu: Distances.py:4-12, BaseKernels.py get_tf_tensor,
// Operators.py:207-225,306-326,442-476, the dense `noise * eye(n)` of Statistics/CovarianceMatrix.py:197-206 and the
// GradientTape sweep of Optimizer/Fitter.py:124-158.
//
//   struct Prog {
//     static constexpr int N_HP, DIM;
//     // value of the kernel for one pair; h = hyper-parameters, ih = their reciprocals (shared memory)
//     static __device__ double value(const double* h, const double* ih, const double* xi, const double* xj, int gi, int gj);
//     // g[p] += w * d value / d h[p] for every hyper-parameter p (g: registers, constant indices only)
//     static __device__ void grad(const double* h, const double* ih, const double* xi, const double* xj, int gi, int gj,
//                                 double w, double (&g)[N_HP + 1]);
//   };
//
// Tiles are 64 x 64, 256 threads.  A thread owns row (tid & 63) and the columns (tid >> 6) + 4 q: a warp covers 32
// consecutive rows of one column (256 contiguous bytes per access), everything that depends on the row only is
// loop-invariant over the thread's 16 columns, and the generated code is instantiated ONCE per kernel (a second inlined
// copy for a row pair - 16-byte accesses - would double the instruction footprint of a body that is ~200 instructions
// per entry against one store; the instruction cache is the scarcer resource for deep kernel trees).
// The file must stay free of host headers (NVRTC has none).
#pragma once
#if !defined(__CUDACC_RTC__)
#include "device_abi.cuh"
#endif

namespace gpb {

constexpr int S_T = 64;

template <class P>
__device__ __forceinline__ void spec_stage_inputs(const GpbMat& d, int i0, int j0, double* s_hp, double* s_ihp, double* s_xi,
                                                  double* s_xj, int tid) {
  for (int i = tid; i < P::N_HP; i += 256) { const double h = d.hp[i]; s_hp[i] = h; s_ihp[i] = 1.0 / h; }
  for (int i = tid; i < S_T * P::DIM; i += 256) {
    const int gi = i0 + i / P::DIM, gj = j0 + i / P::DIM;
    s_xi[i] = (gi < d.n) ? d.X[(size_t)gi * P::DIM + i % P::DIM] : 0.0;
    s_xj[i] = (gj < d.n) ? d.X[(size_t)gj * P::DIM + i % P::DIM] : 0.0;
  }
}

// K + s2 I into the lower triangle of d.A (+ y^T as the carried row n, 0 at (n, n)); `which` maps blockIdx.z to the GP
template <class P>
__device__ __forceinline__ void assemble_spec_body(const GpbMat* __restrict__ mats, const int* __restrict__ which) {
  __shared__ double s_hp[P::N_HP + 1], s_ihp[P::N_HP + 1];
  __shared__ double s_xi[S_T * P::DIM], s_xj[S_T * P::DIM];
  const GpbMat& d = mats[which ? which[blockIdx.z] : blockIdx.z];
  const int n = d.n;
  const int T = (n + S_T - 1) / S_T;
  int ti, tj;
  if (!tri_map(blockIdx.x, T, 1, 0, T, ti, tj)) return;
  bool own_main = true, own_aug = d.aug && ti == T - 1;
  if (d.own_P) {   // distributed plan: write only the 128-blocks this process owns
    constexpr int R = GPB_NB / S_T;
    own_main = ((ti / R) % d.own_P == d.own_p) && ((tj / R / d.own_W) % d.own_Q == d.own_q);
    own_aug = own_aug && ((n / GPB_NB) % d.own_P == d.own_p) && ((tj / R / d.own_W) % d.own_Q == d.own_q);
    if (!own_main && !own_aug) return;
  }
  const int tid = threadIdx.x;
  const int i0 = ti * S_T, j0 = tj * S_T;
  spec_stage_inputs<P>(d, i0, j0, s_hp, s_ihp, s_xi, s_xj, tid);
  __syncthreads();
  const double noise = d.noise ? *d.noise : 0.0;
  const unsigned ld = (unsigned)d.ld;     // unsigned: 64-bit offsets need no sign extension kept live across the loop
  const int r = tid & 63, cg = tid >> 6;
  const int gi = i0 + r;
  if (own_main && gi < n) {
    const double* xi = s_xi + r * P::DIM;
    double* dst = d.A + gi + (size_t)(d.own_compact ? gpb_local_col(j0, d.own_Q, d.own_q, d.own_W) : (long long)j0) * ld;
#pragma unroll 1
    for (int q = 0; q < S_T / 4; ++q) {
      const int c = cg + 4 * q;
      const int gj = j0 + c;
      if (gj >= n || gj > gi) break;   // columns ascend: everything further right lies above the diagonal too
      double v = P::value(s_hp, s_ihp, xi, s_xj + c * P::DIM, gi, gj);
      if (gi == gj) v += noise;
      dst[(size_t)c * ld] = v;
    }
  }
  if (own_aug && tid < S_T) {
    const int gj = j0 + tid;
    if (gj < n) d.A[n + (size_t)(d.own_compact ? gpb_local_col(gj, d.own_Q, d.own_q, d.own_W) : (long long)gj) * ld] = d.y[gj];
  }
  if (d.aug && ti == T - 1 && tj == T - 1 && tid == 0) {
    const int bn = n / GPB_NB;
    if (!d.own_P || (bn % d.own_P == d.own_p && (bn / d.own_W) % d.own_Q == d.own_q))
      d.A[n + (size_t)(d.own_compact ? gpb_local_col(n, d.own_Q, d.own_q, d.own_W) : (long long)n) * ld] = 0.0;
  }
}

// partial sums of dNLL/dtheta over one 64 x 64 tile of the lower triangle of inv(K): gpart[tile][0 .. N_HP]
template <class P>
__device__ __forceinline__ void grad_spec_body(const GpbMat* __restrict__ mats, const int* __restrict__ which) {
  constexpr int NP = P::N_HP + 1;
  __shared__ double s_hp[NP], s_ihp[NP];
  __shared__ double s_xi[S_T * P::DIM], s_xj[S_T * P::DIM];
  __shared__ double s_ai[S_T], s_aj[S_T];
  __shared__ double s_red[8][NP];
  const GpbMat& d = mats[which ? which[blockIdx.z] : blockIdx.z];
  const int n = d.n;
  const int T = (n + S_T - 1) / S_T;
  int ti, tj;
  if (!tri_map(blockIdx.x, T, 1, 0, T, ti, tj)) return;
  if (*d.info != 0) return;     // failed factorisation: grad_reduce_kernel writes NaN, nothing to sum
  const int tid = threadIdx.x;
  if (d.col_world && (tj / (GPB_NB / S_T)) % d.col_world != d.col_rank) {
    // distributed plan: another rank owns this block column; its partial sums are zero here
    for (int pp = tid; pp < NP; pp += 256) d.gpart[(size_t)blockIdx.x * NP + pp] = 0.0;
    return;
  }
  const int i0 = ti * S_T, j0 = tj * S_T;
  spec_stage_inputs<P>(d, i0, j0, s_hp, s_ihp, s_xi, s_xj, tid);
  if (tid < S_T) {
    s_ai[tid] = (i0 + tid < n) ? d.alpha[i0 + tid] : 0.0;
    s_aj[tid] = (j0 + tid < n) ? d.alpha[j0 + tid] : 0.0;
  }
  __syncthreads();
  double g[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) g[p] = 0.0;
  const unsigned ld = (unsigned)d.ld;     // unsigned: 64-bit offsets need no sign extension kept live across the loop
  const int r = tid & 63, cg = tid >> 6;
  const int gi = i0 + r;
  if (gi < n) {
    const double* xi = s_xi + r * P::DIM;
    const double ai = s_ai[r];
    const double gwl = d.gw_logdet, gwq = d.gw_quad;
    const double* src = d.Kinv + gi + (size_t)j0 * ld;
#pragma unroll 1
    for (int q = 0; q < S_T / 4; ++q) {
      const int c = cg + 4 * q;
      const int gj = j0 + c;
      if (gj >= n || gj > gi) break;
      double w = 0.5 * (gwl * src[(size_t)c * ld] - gwq * (ai * s_aj[c]));
      if (gi == gj) g[P::N_HP] += w; else w *= 2.0;
      P::grad(s_hp, s_ihp, xi, s_xj + c * P::DIM, gi, gj, w, g);
    }
  }
  // deterministic block reduction: warp shuffles, then the 8 warp sums in a fixed order
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const double v = warp_sum(g[p]);
    if (lane == 0) s_red[warp][p] = v;
  }
  __syncthreads();
  if (tid < NP) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += s_red[w8][tid];
    d.gpart[(size_t)blockIdx.x * NP + tid] = s;
  }
}

}  // namespace gpb
