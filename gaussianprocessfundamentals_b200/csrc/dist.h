// Process-grid context of the distributed (multi-GPU) factorisation; see dist.cu.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "internal.h"

#define GPB_DIST_MAX_P 8

namespace gpb {

// The exchange steps of the distributed stages go through a transport so that the SAME ownership, packing and launch
// code runs over NCCL (one process per GPU) and over the loop-back world (all ranks of a virtual P x Q grid inside one
// process on one device, each driven by its own host thread - how a single-GPU CI box exercises the distributed path).
struct Transport {
  virtual ~Transport() {}
  virtual bool group_start() = 0;
  virtual bool group_end() = 0;
  // `count` doubles from rank `root` into `recv` of every rank (send == recv: in place); ordered on `s`
  virtual bool broadcast(const double* send, double* recv, size_t count, int root, cudaStream_t s) = 0;
  virtual bool allreduce_sum(double* buf, size_t count, cudaStream_t s) = 0;   // in place
};

struct DistCtx {
  Transport* tr;              // owned
  int rank, world, P, Q, p, q;  // rank = p * Q + q
  int OW;                     // block columns are owned in groups of OW: block (I, J) belongs to (I mod P, (J / OW) mod Q)
};
#define GPB_DIST_MAX_OW 4

const char* dist_last_error();
int dist_col_width(int P);    // ownership width of the block columns on a P x Q grid (GPB_DIST_OW)
int dist_unique_id(unsigned char* id128);
int dist_create(const unsigned char* id128, int rank, int world, int P, int Q, DistCtx** out);
// `world` contexts of one loop-back world (virtual ranks 0 .. world-1 on the current device); out[world]
int dist_create_loopback(int world, int P, int Q, DistCtx** out);
void dist_destroy(DistCtx* d);
void dist_panel_segments(int k, int n_tiles, int P, int* seg_base, int* seg_count, int* seg_first);
size_t dist_stage_bytes(int n);
int dist_owned_cols(int J_lo, int J_hi, int Q, int q, int OW, int* cols, int cap);   // own block columns in [J_lo, J_hi)
cudaError_t run_potrf_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, double* const stage[2], const Exec& ex);
// exchange = false leaves W = inv(L) split by block column in Kinv (run_exchange_lauum_dist must follow)
cudaError_t run_trtri_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, double* scratch, const Exec& ex, bool exchange);
cudaError_t run_exchange_lauum_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, const Exec& ex);
cudaError_t run_lauum_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, cudaStream_t s);
cudaError_t run_grad_allreduce(double* grad, int count, const DistCtx& D, cudaStream_t s);
cudaError_t run_finalize_dist(const GpbMat* dm, double log2pi, cudaStream_t s);
// column storage (1 x Q grids, likelihood only): gdesc = one descriptor per own group of block columns with A shifted so
// that the single-GPU kernels address the packed storage; ring = three outer-panel buffers of dist_ring_bytes each;
// red = 2 + world doubles of scratch
size_t dist_ring_bytes(int n, int OW);
cudaError_t run_potrf_dist_store(const GpbMat* gdesc, const GpbMat& h, const DistCtx& D, double* const ring[3], const Exec& ex);
cudaError_t run_finalize_dist_store(const GpbMat* dm, const DistCtx& D, double* red, double log2pi, cudaStream_t s);

}  // namespace gpb
