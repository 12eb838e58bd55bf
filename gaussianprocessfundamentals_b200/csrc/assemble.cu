// Fused covariance assembly and fused trace-gradient.
//
//   assemble : one pass that evaluates the whole kernel tree per (i, j) in registers and writes K + s2*I in place
//              (replaces Distances.py:4-12, BaseKernels.py get_tf_tensor, Operators.py get_tf_tensor and the
//              dense `noise * eye(n)` of Statistics/CovarianceMatrix.py:197-206; ~20 n^2 temporaries in the reference).
//   grad     : dNLL/dtheta_p = sum_ij 1/2 (Kinv_ij - alpha_i alpha_j) dK_ij/dtheta_p for all p at once, re-evaluating
//              the kernel tree and its reverse-mode derivatives in registers while streaming the lower triangle of
//              Kinv once (replaces the GradientTape sweep of Optimizer/Fitter.py:124-158).
//
// Tiles are 64 x 64, 256 threads: thread (r = tid & 63, cg = tid >> 6) owns row r and columns cg, cg+4, ... so that
// a warp writes / reads 32 consecutive rows of one column (column-major, coalesced).
#include "gemm.cuh"   // tri_map / tri_count
#include "internal.h"
#include "program.cuh"

namespace gpb {

constexpr int A_T = 64;

struct AsmArgs {
  const int32_t* code; int n_ops; int dim; int cp_mode;
  const double* X; const double* X2;  // X2 == nullptr: symmetric self-covariance
  long long n, m;
  const double* hp; int n_hp;
  const double* noise;   // device scalar or nullptr
  double* K; long long ld;
  int lower_only;
  const double* y; int aug;   // aug: row n receives y^T, (n,n) receives 0
  int own_P, own_Q, own_p, own_q, own_W;   // distributed plans: write only the 128-blocks this process owns (own_P == 0: all)
  int own_compact;                         // column storage: K holds only the own block columns (device_abi.cuh)
};

__device__ __forceinline__ void assemble_tile(const AsmArgs& a, int ti, int tj, int T) {
  extern __shared__ __align__(16) unsigned char asm_smem[];
  bool own_main = true, own_aug = a.aug && ti == T - 1;
  if (a.own_P) {
    constexpr int R = GPB_NB / A_T;
    own_main = ((ti / R) % a.own_P == a.own_p) && ((tj / R / a.own_W) % a.own_Q == a.own_q);
    own_aug = own_aug && ((int)(a.n / GPB_NB) % a.own_P == a.own_p) && ((tj / R / a.own_W) % a.own_Q == a.own_q);
    if (!own_main && !own_aug) return;
  }
  int32_t* s_code = reinterpret_cast<int32_t*>(asm_smem);
  double* s_hp = reinterpret_cast<double*>(asm_smem + ((a.n_ops * GPB_OP_WORDS * 4 + 15) / 16) * 16);
  double* s_ihp = s_hp + ((a.n_hp + 1) / 2) * 2 + 2;       // 1 / hp: the per-entry path has no divisions
  double* s_xi = s_ihp + ((a.n_hp + 1) / 2) * 2 + 2;
  double* s_xj = s_xi + A_T * a.dim;
  const int tid = threadIdx.x;
  for (int i = tid; i < a.n_ops * GPB_OP_WORDS; i += 256) s_code[i] = a.code[i];
  for (int i = tid; i < a.n_hp; i += 256) { const double h = a.hp[i]; s_hp[i] = h; s_ihp[i] = 1.0 / h; }
  const double* Xc = a.X2 ? a.X2 : a.X;
  const long long i0 = (long long)ti * A_T, j0 = (long long)tj * A_T;
  for (int i = tid; i < A_T * a.dim; i += 256) {
    const long long gi = i0 + i / a.dim, gj = j0 + i / a.dim;
    s_xi[i] = (gi < a.n) ? a.X[gi * a.dim + i % a.dim] : 0.0;
    s_xj[i] = (gj < a.m) ? Xc[gj * a.dim + i % a.dim] : 0.0;
  }
  __syncthreads();
  const double noise = (a.noise && !a.X2) ? *a.noise : 0.0;
  const int r = tid & 63, cg = tid >> 6;
  const long long gi = i0 + r;
  GpbPair p;
  p.xi = s_xi + r * a.dim; p.dim = a.dim; p.hp = s_hp; p.ihp = s_ihp; p.cp_mode = a.cp_mode; p.gi = gi;
  if (gi < a.n && own_main) {
#pragma unroll 1
    for (int q = 0; q < A_T / 4; ++q) {
      const int c = cg + 4 * q;
      const long long gj = j0 + c;
      if (gj >= a.m) break;
      if (a.lower_only && gj > gi) continue;
      p.xj = s_xj + c * a.dim; p.gj = gj;
      double v = gpb_eval(s_code, a.n_ops, p);
      if (!a.X2 && gi == gj) v += noise;
      a.K[gi + (a.own_compact ? gpb_local_col(gj, a.own_Q, a.own_q, a.own_W) : gj) * a.ld] = v;
    }
  }
  if (own_aug) {
    // carried right-hand side: row n of the factorisation workspace
    if (tid < A_T) {
      const long long gj = j0 + tid;
      if (gj < a.n) a.K[a.n + (a.own_compact ? gpb_local_col(gj, a.own_Q, a.own_q, a.own_W) : gj) * a.ld] = a.y[gj];
    }
  }
  if (a.aug && ti == T - 1 && tj == T - 1 && tid == 0) {
    // element (n, n) accumulates -z^T z; in a distributed plan it belongs to block (n / 128, n / 128)
    const int bn = (int)(a.n / GPB_NB);
    if (!a.own_P || (bn % a.own_P == a.own_p && (bn / a.own_W) % a.own_Q == a.own_q))
      a.K[a.n + (a.own_compact ? gpb_local_col(a.n, a.own_Q, a.own_q, a.own_W) : a.n) * a.ld] = 0.0;
  }
}

__global__ void __launch_bounds__(256) assemble_batched_kernel(const GpbMat* __restrict__ mats, const int* __restrict__ which) {
  const GpbMat& d = mats[which ? which[blockIdx.z] : blockIdx.z];
  const int T = (d.n + A_T - 1) / A_T;
  int ti, tj;
  if (!tri_map(blockIdx.x, T, 1, 0, T, ti, tj)) return;
  AsmArgs a;
  a.code = d.code; a.n_ops = d.n_ops; a.dim = d.dim; a.cp_mode = d.cp_mode;
  a.X = d.X; a.X2 = nullptr; a.n = d.n; a.m = d.n; a.hp = d.hp; a.n_hp = d.n_hp; a.noise = d.noise;
  a.K = d.A; a.ld = d.ld; a.lower_only = 1; a.y = d.y; a.aug = d.aug;
  a.own_P = d.own_P; a.own_Q = d.own_Q; a.own_p = d.own_p; a.own_q = d.own_q; a.own_W = d.own_W;
  a.own_compact = d.own_compact;
  assemble_tile(a, ti, tj, T);
}

__global__ void __launch_bounds__(256) assemble_rect_kernel(const AsmArgs a) {
  const int Tm = (int)((a.n + A_T - 1) / A_T);
  if (a.lower_only) {
    int ti, tj;
    if (!tri_map(blockIdx.x, Tm, 1, 0, Tm, ti, tj)) return;
    assemble_tile(a, ti, tj, Tm);
  } else {
    assemble_tile(a, blockIdx.x, blockIdx.y, Tm);
  }
}

static size_t asm_smem_bytes(int n_ops, int n_hp, int dim) {
  size_t b = ((size_t)(n_ops * GPB_OP_WORDS * 4 + 15) / 16) * 16;
  b += (size_t)2 * (((n_hp + 1) / 2) * 2 + 2) * 8;
  b += (size_t)2 * A_T * dim * 8;
  return b;
}

// ---------------------------------------------------------------------------------------------------------------
// trace gradient
// ---------------------------------------------------------------------------------------------------------------
struct SmemAcc {
  double* base;  // acc[p * 256 + tid]
  __device__ __forceinline__ void operator()(int p, double v) const { base[p * 256] += v; }
};

__global__ void __launch_bounds__(256) grad_kernel(const GpbMat* __restrict__ mats, const int* __restrict__ which, int acc_stride) {
  extern __shared__ __align__(16) unsigned char g_smem[];
  const GpbMat& d = mats[which ? which[blockIdx.z] : blockIdx.z];
  const int T = (d.n + A_T - 1) / A_T;
  int ti, tj;
  if (!tri_map(blockIdx.x, T, 1, 0, T, ti, tj)) return;
  const int P = d.n_hp + 1;
  if (*d.info != 0) return;     // failed factorisation: grad_reduce_kernel writes NaN, nothing to sum
  if (d.col_world && (tj / (GPB_NB / A_T)) % d.col_world != d.col_rank) {
    // distributed plan: another rank owns this block column; its partial sums are zero here
    for (int pp = threadIdx.x; pp < P; pp += 256) d.gpart[(size_t)blockIdx.x * P + pp] = 0.0;
    return;
  }
  double* s_acc = reinterpret_cast<double*>(g_smem);            // [acc_stride][256]
  double* s_hp = s_acc + (size_t)acc_stride * 256;
  double* s_ihp = s_hp + ((d.n_hp + 1) / 2) * 2 + 2;
  double* s_xi = s_ihp + ((d.n_hp + 1) / 2) * 2 + 2;
  double* s_xj = s_xi + A_T * d.dim;
  double* s_ai = s_xj + A_T * d.dim;
  double* s_aj = s_ai + A_T;
  double* s_red = s_aj + A_T;                                   // [8]
  int32_t* s_code = reinterpret_cast<int32_t*>(s_red + 8);
  const int tid = threadIdx.x;
  for (int i = tid; i < d.n_ops * GPB_OP_WORDS; i += 256) s_code[i] = d.code[i];
  for (int i = tid; i < d.n_hp; i += 256) { const double h = d.hp[i]; s_hp[i] = h; s_ihp[i] = 1.0 / h; }
  for (int p = 0; p < P; ++p) s_acc[p * 256 + tid] = 0.0;
  const int i0 = ti * A_T, j0 = tj * A_T;
  for (int i = tid; i < A_T * d.dim; i += 256) {
    const int gi = i0 + i / d.dim, gj = j0 + i / d.dim;
    s_xi[i] = (gi < d.n) ? d.X[(size_t)gi * d.dim + i % d.dim] : 0.0;
    s_xj[i] = (gj < d.n) ? d.X[(size_t)gj * d.dim + i % d.dim] : 0.0;
  }
  if (tid < A_T) {
    s_ai[tid] = (i0 + tid < d.n) ? d.alpha[i0 + tid] : 0.0;
    s_aj[tid] = (j0 + tid < d.n) ? d.alpha[j0 + tid] : 0.0;
  }
  __syncthreads();
  const int r = tid & 63, cg = tid >> 6;
  const int gi = i0 + r;
  SmemAcc acc{s_acc + tid};
  GpbPair p;
  p.xi = s_xi + r * d.dim; p.dim = d.dim; p.hp = s_hp; p.ihp = s_ihp; p.cp_mode = d.cp_mode; p.gi = gi;
  if (gi < d.n) {
    const double ai = s_ai[r];
    const double gwl = d.gw_logdet, gwq = d.gw_quad;
#pragma unroll 1
    for (int q = 0; q < A_T / 4; ++q) {
      const int c = cg + 4 * q;
      const int gj = j0 + c;
      if (gj >= d.n || gj > gi) continue;
      const double kinv = d.Kinv[gi + (size_t)gj * d.ld];
      double w = 0.5 * (gwl * kinv - gwq * (ai * s_aj[c]));
      if (gi == gj) acc(d.n_hp, w); else w *= 2.0;
      p.xj = s_xj + c * d.dim; p.gj = gj;
      gpb_eval_grad(s_code, d.n_ops, p, w, acc);
    }
  }
  __syncthreads();
  // block reduction: 8 warps, one pass per hyper-parameter
  const int lane = tid & 31, warp = tid >> 5;
  for (int pp = 0; pp < P; ++pp) {
    double v = warp_sum(s_acc[pp * 256 + tid]);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w8 = 0; w8 < 8; ++w8) s += s_red[w8];
      d.gpart[(size_t)blockIdx.x * P + pp] = s;
    }
    __syncthreads();
  }
}

// deterministic second stage: one warp per hyper-parameter sums the per-tile partials
__global__ void __launch_bounds__(256) grad_reduce_kernel(const GpbMat* __restrict__ mats) {
  const GpbMat& d = mats[blockIdx.y];
  const int P = d.n_hp + 1;
  const int pp = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pp >= P) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int t = lane; t < d.n_gtiles; t += 32) s += d.gpart[(size_t)t * P + pp];
  s = warp_sum(s);
  if (lane == 0) d.grad[pp] = (*d.info != 0) ? nan("") : s;
}

int grad_tiles(int n) {
  const int T = (n + A_T - 1) / A_T;
  return (int)tri_count(T, 1, 0, T);
}

static size_t grad_smem_bytes(int n_ops_max, int n_hp_max, int dim) {
  size_t b = (size_t)(n_hp_max + 1) * 256 * 8;
  b += (size_t)2 * (((n_hp_max + 1) / 2) * 2 + 2) * 8;
  b += (size_t)2 * A_T * dim * 8 + (size_t)2 * A_T * 8 + 64;
  b += (size_t)n_ops_max * GPB_OP_WORDS * 4 + 16;
  return b;
}

#define GPB_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t assemble_init() {
  GPB_CK(cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  GPB_CK(cudaFuncSetAttribute(assemble_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  GPB_CK(cudaFuncSetAttribute(assemble_rect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  return cudaSuccess;
}

// the GPs which[0 .. B) of the descriptor array (which == nullptr: the first B) on the interpreter kernels
cudaError_t run_assemble_batched(const GpbMat* dm, const int* which, int B, int n_max, cudaStream_t s) {
  const int T = (n_max + A_T - 1) / A_T;
  const size_t smem = asm_smem_bytes(GPB_MAX_OPS, GPB_MAX_HP, GPB_MAX_DIM);
  assemble_batched_kernel<<<dim3((unsigned)tri_count(T, 1, 0, T), 1, B), 256, smem, s>>>(dm, which);
  ++g_launches;
  return cudaGetLastError();
}

cudaError_t run_assemble_rect(const int32_t* code_dev, int n_ops, int dim, int cp_mode, const double* X,
                              const double* X2, long long n, long long m, const double* hp_dev, int n_hp,
                              const double* noise_dev, double* K, long long ldk, int lower_only, cudaStream_t s) {
  AsmArgs a;
  a.code = code_dev; a.n_ops = n_ops; a.dim = dim; a.cp_mode = cp_mode;
  a.X = X; a.X2 = X2; a.n = n; a.m = m; a.hp = hp_dev; a.n_hp = n_hp; a.noise = noise_dev;
  a.K = K; a.ld = ldk; a.lower_only = lower_only; a.y = nullptr; a.aug = 0;
  a.own_P = a.own_Q = a.own_p = a.own_q = 0; a.own_W = 1; a.own_compact = 0;
  const int Tm = (int)((n + A_T - 1) / A_T), Tn = (int)((m + A_T - 1) / A_T);
  if (Tm == 0 || Tn == 0) return cudaSuccess;
  const size_t smem = asm_smem_bytes(n_ops, n_hp, dim);
  dim3 grid = lower_only ? dim3((unsigned)tri_count(Tm, 1, 0, Tm), 1, 1) : dim3(Tm, Tn, 1);
  assemble_rect_kernel<<<grid, 256, smem, s>>>(a);
  ++g_launches;
  return cudaGetLastError();
}

// per-tile partial sums of the gradient for the GPs which[0 .. B) on the interpreter kernel
cudaError_t run_grad_tiles(const GpbMat* dm, const int* which, int B, int n_max, int n_hp_max, int n_ops_max, int dim,
                           cudaStream_t s) {
  const int tiles = grad_tiles(n_max);
  const size_t smem = grad_smem_bytes(n_ops_max, n_hp_max, dim);
  grad_kernel<<<dim3(tiles, 1, B), 256, smem, s>>>(dm, which, n_hp_max + 1);
  ++g_launches;
  return cudaGetLastError();
}

// second stage for ALL B GPs of the plan (the partial sums may come from interpreter and specialised kernels alike)
cudaError_t run_grad_reduce(const GpbMat* dm, int B, int n_hp_max, cudaStream_t s) {
  grad_reduce_kernel<<<dim3((n_hp_max + 1 + 7) / 8, B), 256, 0, s>>>(dm);
  ++g_launches;
  return cudaGetLastError();
}

}  // namespace gpb
