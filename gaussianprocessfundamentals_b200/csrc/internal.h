// Host-side launchers shared between the translation units of libgpb (not part of the public C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "common.cuh"

namespace gpb {

extern std::atomic<long long> g_launches;  // kernels launched since load (gpb_launch_count)

struct Exec {
  cudaStream_t main;       // the caller's stream
  cudaStream_t crit;       // high priority: diagonal block, panel, next panel column (the critical path)
  cudaStream_t mid;        // medium priority: the update of the second-next outer panel's columns (depth-2 look-ahead)
  cudaStream_t side;       // low priority: the bulk of the trailing update
  cudaStream_t comm;       // high priority: pack / broadcast / unpack of the distributed factorisation's panels
  cudaStream_t inv;        // lowest priority: the triangular inverse of what is already final (overlapped with the tail)
  cudaEvent_t ev_c[2];     // block columns final (recorded on crit, waited on by inv)
  cudaEvent_t ev_fork;     // caller's stream -> crit / side
  cudaEvent_t ev_e[2];     // panel k ready (recorded on crit)
  cudaEvent_t ev_g[2];     // column k+2 updated by panel k (recorded on side; distributed factorisation)
  cudaEvent_t ev_b[2];     // second-next outer panel's columns updated by outer step s (recorded on mid)
  cudaEvent_t ev_d[2];     // bulk update of outer step s complete (recorded on side)
  cudaEvent_t ev_p[2];     // outer panel factorised on its owner (recorded on crit, waited on by comm)
  cudaEvent_t ev_x[2];     // outer panel available on this rank (recorded on comm)
  cudaEvent_t ev_ring[3][3]; // column storage: the updates (crit / mid / side) that read ring slot i are done
  cudaEvent_t ev_join_comm;  // comm -> caller's stream
  cudaEvent_t ev_join[4];  // crit / side / inv / mid -> caller's stream
};

// ---- linalg.cu ------------------------------------------------------------------------------------------------
// with_trtri: also compute W = inv(L) in place, overlapped with the factorisation (look-ahead path, B == 1)
cudaError_t run_potrf(const GpbMat* dmats, int B, int n_max, int aug, bool lookahead, bool with_trtri, const Exec& ex);
cudaError_t run_diag(const GpbMat* dmats, int B, int k, cudaStream_t s);   // diagonal block k: Cholesky + inverse
cudaError_t run_finalize(const GpbMat* dmats, int B, double log2pi, cudaStream_t s);
cudaError_t run_trtri(const GpbMat* dmats, int B, int n_max, cudaStream_t s);
cudaError_t run_alpha(const GpbMat* dmats, int B, int n_max, cudaStream_t s);
cudaError_t run_lauum(const GpbMat* dmats, int B, int n_max, cudaStream_t s);
cudaError_t run_trsv(const GpbMat* dmats, int B, int n_max, int transposed, cudaStream_t s);
cudaError_t run_zero_upper(double* A, int n, int ld, cudaStream_t s);
cudaError_t run_symmetrize(double* A, int n, int ld, cudaStream_t s);
cudaError_t run_gemm_plain(int akm, int bkm, const double* A, int lda, const double* Bm, int ldb, double* C, int ldc,
                           int M, int N, int K, double alpha, double beta, cudaStream_t s);
cudaError_t run_microbench(int kind, int iters, int blocks, cudaStream_t s);
int trtri_schedule_host(int n, const long long* fc, int n_fc, int* out, int cap);   // host-only: the overlapped inverse's schedule
int potrf_outer_blocks(int n_max);   // 128-blocks per outer panel of the factorisation (1 or 2; GPB_POTRF_KB)
cudaError_t linalg_init();
int debug_diag_clocks(long long* out);   // developer builds (-DGPB_DIAG_CLOCKS=1): phase clocks of the last diagonal-block launch

// ---- assemble.cu ----------------------------------------------------------------------------------------------
cudaError_t run_assemble_batched(const GpbMat* dmats, const int* which, int B, int n_max, cudaStream_t s);
cudaError_t run_assemble_rect(const int32_t* code_dev, int n_ops, int dim, int cp_mode, const double* X,
                              const double* X2, long long n, long long m, const double* hp_dev, int n_hp,
                              const double* noise_dev, double* K, long long ldk, int lower_only, cudaStream_t s);
cudaError_t run_grad_tiles(const GpbMat* dmats, const int* which, int B, int n_max, int n_hp_max, int n_ops_max, int dim,
                           cudaStream_t s);
cudaError_t run_grad_reduce(const GpbMat* dmats, int B, int n_hp_max, cudaStream_t s);
int grad_tiles(int n);
cudaError_t assemble_init();

}  // namespace gpb
