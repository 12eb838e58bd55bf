// One large GP over a P x Q process grid (one process per GPU, NCCL over NVLink / NVSwitch): distributed Cholesky,
// triangular inverse, inv(K) and trace gradient.
//
// Replaces, for matrices too large or too slow for one GPU, the same reference calls as the single-GPU plan:
// HolisticCovarianceMatrix.get_L_K / get_L_alpha (Statistics/CovarianceMatrix.py:247-265),
// LogLikelihood.get_metric (Metrics/LogLikelihood.py:30-65) and the gradient of Optimizer/Fitter.py:124-158.
// BASELINE config 5 / SURVEY 8(e), second row.
//
// Factorisation.  2D block-cyclic OWNERSHIP of the 128 x 128 blocks of the lower triangle - block (I, J) is assembled
// and updated by rank (I mod P) * Q + ((J / OW) mod Q), OW = dist_col_width(P) - in two storage layouts:
//   * REPLICATED storage (gpb_plan_create_dist): every rank holds the full (n+1) x ld column-major workspace of the
//     single-GPU plan, its own blocks are live, and each finished panel (+ inverted diagonal block) is broadcast to all
//     ranks, so that after the factorisation every rank holds the complete L - the inverse and gradient stages read all of it;
//   * COLUMN storage (gpb_plan_create_dist_columns, 1 x Q grids, likelihood only): every rank holds only its own block
//     columns plus a ring of three outer-panel buffers (run_potrf_dist_store).
// Drivers:
//   run_potrf_dist_cols  (P == 1, the default grid 1 x N): block columns owned in groups of OW = 2; an outer panel is
//                        factorised on its owner without any exchange and shipped with ONE broadcast on a
//                        communication stream of its own; receivers post it only when the owner can be ready, so the
//                        NCCL kernel does not spin on SMs the bulk update needs; look-ahead of depth 2.
//   run_potrf_dist       (P > 1): single block columns; ncclBroadcast of {L_kk, inv(L_kk)} down the process column, then
//                        one grouped ncclBroadcast per process row of the panel blocks that row owns.
//   run_potrf_dist_store (column storage): the schedule of run_potrf_dist_cols, panels broadcast in place.
// NVSwitch gives every pair of GPUs full bandwidth, so replicating the panel (n^2/2 doubles per rank over the whole
// factorisation, 17 GB at n = 65536, ~25 ms at the measured 700 GB/s) is cheap; what limits the scaling at n = 32768 on
// 8 GPUs is the serial chain factorise -> broadcast -> update of the next owner's columns.
// Panels travel tile-major through two staging buffers (replicated storage).
//
// Streams: the critical path (diagonal block, panel, strip update, next group's columns) runs on the plan's
// high-priority stream, the exchange on the communication stream, the second-next group's columns on the medium and the
// bulk of the trailing update on the low-priority one.  CTAs of the bulk update are one tile each: persistent CTAs
// kept every SM until the update ended and serialised the NCCL kernels behind it.
//
// Gradient stages: see the block comment above GeoDistTriDiag.
#include <dlfcn.h>
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "dist.h"
#include "gemm.cuh"
#include "internal.h"
#include "trace.h"

namespace gpb {

// ---------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time (no link-time dependency: single-GPU users never load it).  Minimal declarations of the
// stable NCCL 2 C API; `ncclFloat64` = 8 and `ncclSum` = 0 in every NCCL 2 release.
// ---------------------------------------------------------------------------------------------------------------
typedef int ncclResult_t;
struct NcclId { char internal[128]; };
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(NcclId*) = nullptr;
  ncclResult_t (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  ncclResult_t (*CommDestroy)(void*) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
constexpr int NCCL_F64 = 8, NCCL_SUM = 0;
static NcclApi g_nccl;
static thread_local std::string g_dist_err;     // per host thread: the loop-back ranks run on one thread each
const char* dist_last_error() { return g_dist_err.c_str(); }

static bool nccl_load() {
  if (g_nccl.handle) return true;
  // prefer the copy already mapped into the process (torch's bundled libnccl), then the system one
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_dist_err = std::string("cannot load libnccl: ") + dlerror(); return false; }
  NcclApi a;
  a.handle = h;
#define GPB_SYM(field, name) *(void**)(&a.field) = dlsym(h, name); if (!a.field) { g_dist_err = "libnccl lacks " name; return false; }
  GPB_SYM(GetUniqueId, "ncclGetUniqueId")
  GPB_SYM(CommInitRank, "ncclCommInitRank")
  GPB_SYM(CommDestroy, "ncclCommDestroy")
  GPB_SYM(Broadcast, "ncclBroadcast")
  GPB_SYM(AllReduce, "ncclAllReduce")
  GPB_SYM(GroupStart, "ncclGroupStart")
  GPB_SYM(GroupEnd, "ncclGroupEnd")
  GPB_SYM(GetErrorString, "ncclGetErrorString")
#undef GPB_SYM
  g_nccl = a;
  return true;
}
static bool nccl_ok(ncclResult_t r, const char* where) {
  if (r == 0) return true;
  g_dist_err = std::string(where) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
  return false;
}
// Ownership width of the block columns.  P == 1 (the default grid 1 x N): groups of 2 block columns - the owner
// factorises a whole 256-wide outer panel without any exchange inside it and ships it with ONE broadcast
// (run_potrf_dist_cols).  P > 1: single block columns (run_potrf_dist).  GPB_DIST_OW=1|2|4 overrides for P == 1.
int dist_col_width(int P) {
  if (P != 1) return 1;
  const char* e = getenv("GPB_DIST_OW");
  const int v = e ? atoi(e) : 2;
  return v < 1 ? 1 : (v > GPB_DIST_MAX_OW ? GPB_DIST_MAX_OW : v);
}

int dist_unique_id(unsigned char* id128) {
  if (!nccl_load()) return 1;
  NcclId id;
  if (!nccl_ok(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return 1;
  memcpy(id128, id.internal, 128);
  return 0;
}

// ---- transport 1: NCCL (one process per GPU) -----------------------------------------------------------------------
struct NcclTransport : Transport {
  void* comm = nullptr;
  ~NcclTransport() override { if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm); }
  bool group_start() override { return nccl_ok(g_nccl.GroupStart(), "ncclGroupStart"); }
  bool group_end() override { return nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd"); }
  bool broadcast(const double* send, double* recv, size_t count, int root, cudaStream_t s) override {
    return nccl_ok(g_nccl.Broadcast(send, recv, count, NCCL_F64, root, comm, s), "ncclBroadcast");
  }
  bool allreduce_sum(double* buf, size_t count, cudaStream_t s) override {
    return nccl_ok(g_nccl.AllReduce(buf, buf, count, NCCL_F64, NCCL_SUM, comm, s), "ncclAllReduce");
  }
};

int dist_create(const unsigned char* id128, int rank, int world, int P, int Q, DistCtx** out) {
  if (!nccl_load()) return 1;
  NcclId id;
  memcpy(id.internal, id128, 128);
  void* comm = nullptr;
  if (!nccl_ok(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank")) return 1;
  NcclTransport* t = new NcclTransport;
  t->comm = comm;
  DistCtx* d = new DistCtx;
  d->tr = t; d->rank = rank; d->world = world; d->P = P; d->Q = Q; d->p = rank / Q; d->q = rank % Q;
  d->OW = dist_col_width(P);
  *out = d;
  return 0;
}

// ---- transport 2: loop-back world (virtual ranks inside one process, one device) ---------------------------------------
// Every virtual rank is driven by its own host thread with its own plan, workspace and streams.  A collective is a host
// rendezvous plus stream-ordered device copies: the root records an event behind its pending work and publishes its
// buffer; each receiver makes its stream wait on that event, copies device-to-device and records a done-event; the root's
// stream then waits on all done-events (its buffer may be overwritten afterwards).  Nothing ever spins on the device -
// the only cross-rank dependencies are CUDA events recorded BEFORE they are waited on - so the virtual ranks cannot
// dead-lock however the hardware interleaves their kernels.  Ranks must issue their collectives in the same order (as
// with NCCL).
struct LoopWorld {
  struct Slot {
    const double* send = nullptr;
    cudaEvent_t ready = nullptr;
    std::vector<cudaEvent_t> done;      // one per receiver that has enqueued its copy
    bool posted = false;
    int finished = 0;                   // ranks that are through with the slot
    std::vector<const double*> contrib; // all-reduce: staged contribution of every rank
    std::vector<cudaEvent_t> contrib_ready;
    double* staging = nullptr;
    int arrived = 0;
  };
  int world;
  std::mutex mu;
  std::condition_variable cv;
  std::map<long long, Slot> slots;      // collective sequence number -> rendezvous state
  std::vector<double*> stagings;        // freed with the world
  explicit LoopWorld(int w) : world(w) {}
  ~LoopWorld() { for (double* p : stagings) cudaFree(p); }
};

__global__ void loop_sum_kernel(double* __restrict__ dst, const double* __restrict__ staging, int world, size_t count) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  double s = 0.0;
  for (int r = 0; r < world; ++r) s += staging[(size_t)r * count + i];   // rank order: every rank gets the same bits
  dst[i] = s;
}

struct LoopTransport : Transport {
  std::shared_ptr<LoopWorld> w;
  int rank = 0;
  long long seq = 0;
  bool fail(cudaError_t e, const char* where) {
    g_dist_err = std::string("loop-back ") + where + ": " + cudaGetErrorString(e);
    return false;
  }
  // a rank that failed (or a caller that did not start every rank) must not hang the others for ever
  template <class Pred>
  bool wait_for(std::unique_lock<std::mutex>& lk, Pred pred) {
    if (w->cv.wait_for(lk, std::chrono::seconds(120), pred)) return true;
    g_dist_err = "loop-back rendezvous timed out: not every virtual rank issued this collective";
    return false;
  }
  bool group_start() override { return true; }
  bool group_end() override { return true; }
  bool broadcast(const double* send, double* recv, size_t count, int root, cudaStream_t s) override {
    const long long id = seq++;
    cudaError_t e;
    std::unique_lock<std::mutex> lk(w->mu);
    LoopWorld::Slot& sl = w->slots[id];
    if (rank == root) {
      if ((e = cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
      if ((e = cudaEventRecord(sl.ready, s)) != cudaSuccess) return fail(e, "record");
      sl.send = send; sl.posted = true;
      w->cv.notify_all();
      if (!wait_for(lk, [&] { return (int)sl.done.size() == w->world - 1; })) return false;
      for (cudaEvent_t ev : sl.done)
        if ((e = cudaStreamWaitEvent(s, ev, 0)) != cudaSuccess) return fail(e, "wait(done)");
      if (send != recv && count)
        if ((e = cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return fail(e, "copy");
    } else {
      if (!wait_for(lk, [&] { return sl.posted; })) return false;
      if ((e = cudaStreamWaitEvent(s, sl.ready, 0)) != cudaSuccess) return fail(e, "wait(ready)");
      if (count)
        if ((e = cudaMemcpyAsync(recv, sl.send, count * sizeof(double), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return fail(e, "copy");
      cudaEvent_t dn;
      if ((e = cudaEventCreateWithFlags(&dn, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
      if ((e = cudaEventRecord(dn, s)) != cudaSuccess) return fail(e, "record");
      sl.done.push_back(dn);
      w->cv.notify_all();
    }
    if (++sl.finished == w->world) {          // events may be destroyed while work is pending: freed on completion
      if (sl.ready) cudaEventDestroy(sl.ready);
      for (cudaEvent_t ev : sl.done) cudaEventDestroy(ev);
      w->slots.erase(id);
    }
    return true;
  }
  bool allreduce_sum(double* buf, size_t count, cudaStream_t s) override {
    const long long id = seq++;
    cudaError_t e;
    std::unique_lock<std::mutex> lk(w->mu);
    LoopWorld::Slot& sl = w->slots[id];
    if (!sl.staging) {
      if ((e = cudaMalloc(&sl.staging, (size_t)w->world * count * sizeof(double))) != cudaSuccess) return fail(e, "staging");
      w->stagings.push_back(sl.staging);
      sl.contrib_ready.assign(w->world, nullptr);
    }
    if ((e = cudaMemcpyAsync(sl.staging + (size_t)rank * count, buf, count * sizeof(double), cudaMemcpyDeviceToDevice, s)) != cudaSuccess)
      return fail(e, "stage");
    if ((e = cudaEventCreateWithFlags(&sl.contrib_ready[rank], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
    if ((e = cudaEventRecord(sl.contrib_ready[rank], s)) != cudaSuccess) return fail(e, "record");
    ++sl.arrived;
    w->cv.notify_all();
    if (!wait_for(lk, [&] { return sl.arrived == w->world; })) return false;
    for (int r = 0; r < w->world; ++r)
      if (r != rank && (e = cudaStreamWaitEvent(s, sl.contrib_ready[r], 0)) != cudaSuccess) return fail(e, "wait(contrib)");
    loop_sum_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(buf, sl.staging, w->world, count);
    ++g_launches;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "sum");
    if (++sl.finished == w->world) {
      for (cudaEvent_t ev : sl.contrib_ready) if (ev) cudaEventDestroy(ev);
      w->slots.erase(id);                      // the staging buffer stays alive until the world is destroyed
    }
    return true;
  }
};

int dist_create_loopback(int world, int P, int Q, DistCtx** out) {
  auto w = std::make_shared<LoopWorld>(world);
  for (int r = 0; r < world; ++r) {
    LoopTransport* t = new LoopTransport;
    t->w = w; t->rank = r;
    DistCtx* d = new DistCtx;
    d->tr = t; d->rank = r; d->world = world; d->P = P; d->Q = Q; d->p = r / Q; d->q = r % Q;
    d->OW = dist_col_width(P);
    out[r] = d;
  }
  return 0;
}

void dist_destroy(DistCtx* d) {
  if (!d) return;
  delete d->tr;
  delete d;
}

// ---------------------------------------------------------------------------------------------------------------
// host-side ownership arithmetic (also exported through the C-ABI for the CPU tests)
// ---------------------------------------------------------------------------------------------------------------
void dist_panel_segments(int k, int n_tiles, int P, int* seg_base, int* seg_count, int* seg_first) {
  // tile t of panel k is block row I = k + t; it belongs to process row I mod P.  Staging order: by process row, each
  // row's tiles ascending.
  int base = 0;
  for (int o = 0; o < P; ++o) {
    const int t0 = ((o - k) % P + P) % P;          // first tile of process row o
    const int cnt = t0 < n_tiles ? (n_tiles - t0 + P - 1) / P : 0;
    seg_base[o] = base; seg_count[o] = cnt; seg_first[o] = t0;
    base += cnt;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------------
struct PanelMap {
  int P;
  int seg_base[GPB_DIST_MAX_P];
  int seg_first[GPB_DIST_MAX_P];
  __device__ __forceinline__ int slot(int k, int t) const {
    const int o = (k + t) % P;
    return seg_base[o] + (t - seg_first[o]) / P;
  }
};

constexpr int TILE_ELEMS = GPB_NB * GPB_NB;

// A -> stage for the panel tiles this rank owns (mode 0), stage -> A for the tiles it does not own (mode 1).
// Tile t = block row k + t of block column k; t0 = first tile handled (0 includes the diagonal block).
__global__ void __launch_bounds__(256) panel_copy_kernel(double* __restrict__ A, long long ld, int nrows, int k, int t0,
                                                         double* __restrict__ stage, PanelMap map, int my_p, int owner_col,
                                                         int my_q, int mode) {
  const int t = t0 + blockIdx.x;
  const int I = k + t;
  const bool mine = (I % map.P == my_p) && (my_q == owner_col);
  if ((mode == 0) != mine) return;
  double* st = stage + (size_t)map.slot(k, t) * TILE_ELEMS;
  const int rows = min(GPB_NB, nrows - I * GPB_NB);
  const int cols = min(GPB_NB, nrows - k * GPB_NB);
  double* a = A + (size_t)I * GPB_NB + (size_t)k * GPB_NB * ld;
  for (int idx = threadIdx.x; idx < TILE_ELEMS; idx += 256) {
    const int i = idx & (GPB_NB - 1), c = idx >> 7;
    if (i < rows && c < cols) {
      if (mode == 0) st[idx] = a[i + (size_t)c * ld];
      else a[i + (size_t)c * ld] = st[idx];
    } else if (mode == 0) {
      st[idx] = 0.0;
    }
  }
}

// The same copy for the 1 x Q path (tile t of panel k <-> slot t), 32 columns of a tile per CTA and 16-byte accesses for
// full tiles: the pack and the unpack sit on the factorisation's critical chain between the panel and the update of the
// next owner (8 GPUs, n = 32768: 120 + 110 us per 66 MB outer panel with the kernel above, a sixth of the chain).
__global__ void __launch_bounds__(256) panel_copy2_kernel(double* __restrict__ A, long long ld, int nrows, int k,
                                                          double* __restrict__ stage, int mode) {
  const int I = k + blockIdx.x;
  const int c0 = blockIdx.y * 32;
  double* st = stage + (size_t)blockIdx.x * TILE_ELEMS + (size_t)c0 * GPB_NB;
  const int rows = min(GPB_NB, nrows - I * GPB_NB);
  const int cols = min(GPB_NB, nrows - k * GPB_NB) - c0;
  double* a = A + (size_t)I * GPB_NB + (size_t)(k * GPB_NB + c0) * ld;
  if (rows == GPB_NB && cols >= 32 && (ld & 1) == 0) {
    double2 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = threadIdx.x + 256 * q, i2 = idx & 63, c = idx >> 6;      // 64 double2 per column, 32 columns
      v[q] = mode == 0 ? *reinterpret_cast<const double2*>(a + 2 * i2 + (size_t)c * ld)
                       : *reinterpret_cast<const double2*>(st + 2 * i2 + c * GPB_NB);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = threadIdx.x + 256 * q, i2 = idx & 63, c = idx >> 6;
      if (mode == 0) *reinterpret_cast<double2*>(st + 2 * i2 + c * GPB_NB) = v[q];
      else *reinterpret_cast<double2*>(a + 2 * i2 + (size_t)c * ld) = v[q];
    }
    return;
  }
  for (int idx = threadIdx.x; idx < 32 * GPB_NB; idx += 256) {
    const int i = idx & (GPB_NB - 1), c = idx >> 7;
    if (i < rows && c < cols) {
      if (mode == 0) st[idx] = a[i + (size_t)c * ld];
      else a[i + (size_t)c * ld] = st[idx];
    } else if (mode == 0) {
      st[idx] = 0.0;
    }
  }
}

// plain copy of 128 x 128 doubles (inverse of the diagonal block in / out of the staging buffer)
__global__ void __launch_bounds__(256) tile_copy_kernel(double* __restrict__ dst, const double* __restrict__ src) {
  for (int idx = threadIdx.x + blockIdx.x * 256; idx < TILE_ELEMS; idx += 256 * gridDim.x) dst[idx] = src[idx];
}

// panel product for the blocks of block column k owned by this process row: A[I, k] = A[I, k] * Wd_k^T
struct GeoDistPanel {
  const GpbMat* mats;
  int k, P, p;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    static_assert(BN == GPB_NB, "the in-place panel product needs full-width tiles");
    const GpbMat& d = mats[0];
    const int nrows = d.n + d.aug;
    const int i0 = (k + 1) * GPB_NB + b.x * BM;
    if (i0 >= nrows) return false;
    if ((i0 / GPB_NB) % P != p) return false;
    double* Pn = d.A + (size_t)k * GPB_NB * d.ld + i0;
    J.A = Pn; J.C = Pn;
    J.B = d.Wd + (size_t)k * GPB_NB * GPB_NB;
    J.lda = J.ldc = d.ld; J.ldb = GPB_NB;
    J.mrem = min(BM, nrows - i0);
    J.nrem = GPB_NB;
    J.klo = 0; J.khi = GPB_NB;
    J.alpha = 1.0; J.beta = 0.0; J.red = 0;
    return true;
  }
};

// block columns owned by process column q (groups of OW, dealt out cyclically over Q): device_abi.cuh
__host__ __device__ inline int owned_cols_below(int J, int Q, int q, int OW) { return gpb_owned_cols_below(J, Q, q, OW); }
__host__ __device__ inline int owned_col_at(int t, int Q, int q, int OW) { return gpb_owned_col_at(t, Q, q, OW); }

// trailing update with panel k of the blocks (I, J) this rank owns, J in [J_lo, J_hi):  A[I, J] -= L[I, k] L[J, k]^T.
// grid.y enumerates this rank's block columns of the range, grid.x the BM-row tiles from the diagonal block down.
struct GeoDistSyrk {
  const GpbMat* mats;
  int k, kb, J_lo, J_hi, P, Q, p, q, OW;   // panel = block columns k .. k + kb - 1
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    static_assert(BN == GPB_NB, "block columns are 128 wide");
    const GpbMat& d = mats[0];
    const int nrows = d.n + d.aug;
    const int Jb = owned_col_at(owned_cols_below(J_lo, Q, q, OW) + (int)b.y, Q, q, OW);   // b.y-th own column >= J_lo
    if (Jb >= J_hi) return false;
    const int row = Jb * GPB_NB + (int)b.x * BM;
    if (row >= nrows) return false;
    if ((row / GPB_NB) % P != p) return false;
    const size_t ld = d.ld;
    const double* Pn = d.A + (size_t)k * GPB_NB * ld;
    J.A = Pn + row;
    J.B = Pn + (size_t)Jb * GPB_NB;
    J.C = d.A + row + (size_t)Jb * GPB_NB * ld;
    J.lda = J.ldb = J.ldc = d.ld;
    J.mrem = min(BM, nrows - row);
    J.nrem = min(BN, nrows - Jb * GPB_NB);
    J.klo = 0; J.khi = kb * GPB_NB;
    J.alpha = -1.0; J.beta = 1.0; J.red = g_red_epilogue;
    return true;
  }
};

// nll from the replicated factor: sum(log diag L), z^T z from the carried row, first non-positive / NaN pivot
__global__ void __launch_bounds__(1024) dist_finalize_kernel(const GpbMat* __restrict__ mats, double log2pi) {
  __shared__ double s_ld[32], s_q[32];
  __shared__ int s_bad[32];
  const GpbMat d = mats[0];
  const int n = d.n;
  const size_t ld = d.ld;
  double lsum = 0.0, qsum = 0.0;
  int bad = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double lii = d.A[i + (size_t)i * ld];
    const double z = d.A[n + (size_t)i * ld];
    d.zvec[i] = z;
    if (!(lii > 0.0)) bad = min(bad, i + 1);
    lsum += log(lii);
    qsum += z * z;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  lsum = warp_sum(lsum); qsum = warp_sum(qsum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  if (lane == 0) { s_ld[warp] = lsum; s_q[warp] = qsum; s_bad[warp] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    int w = 0x7fffffff;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s_ld[i]; b += s_q[i]; w = min(w, s_bad[i]); }
    const int info = (w == 0x7fffffff) ? 0 : w;
    *d.info = info;
    *d.nll = info ? nan("") : 0.5 * b + a + 0.5 * ((double)n * log2pi);
    d.terms[0] = b; d.terms[1] = a;
  }
}

#define GPB_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
#define GPB_TR(x) do { if (!(x)) return cudaErrorUnknown; } while (0)

template <class Cfg, class Geo>
static cudaError_t launch_geo(const Geo& geo, dim3 grid, cudaStream_t s, bool persistent = false, const char* tag = "dgemm",
                              int ta = 0) {
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return cudaSuccess;
  TraceSpan span(tag, s, ta, (int)(grid.x * grid.y * grid.z));
  static bool attr_set = false;
  if (!attr_set) {
    GPB_CK(cudaFuncSetAttribute(gemm_kernel<Cfg, false, false, Geo>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg::SMEM_BYTES));
    attr_set = true;
  }
  gemm_kernel<Cfg, false, false, Geo><<<persistent_ctas(grid, Cfg::MIN_CTAS, persistent ? 1 : 1), Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(geo, grid);
  ++g_launches;
  return cudaGetLastError();
}

static int count_owned_cols(int J_lo, int J_hi, int Q, int q, int OW = 1) {
  if (J_hi <= J_lo) return 0;
  return owned_cols_below(J_hi, Q, q, OW) - owned_cols_below(J_lo, Q, q, OW);
}

// the enumeration GeoDistSyrk uses, on the host (C-ABI gpb_dist_owned_cols; CPU tests)
int dist_owned_cols(int J_lo, int J_hi, int Q, int q, int OW, int* cols, int cap) {
  const int nc = count_owned_cols(J_lo, J_hi, Q, q, OW);
  const int t0 = owned_cols_below(J_lo, Q, q, OW);
  for (int t = 0; t < nc && t < cap; ++t) cols[t] = owned_col_at(t0 + t, Q, q, OW);
  return nc;
}

size_t dist_stage_bytes(int n) {
  const int nrows = n + 1;
  const int nt = (nrows + GPB_NB - 1) / GPB_NB;
  // +1 slot for inv(L_kk), +1 spare, per block column of an outer panel (up to GPB_DIST_MAX_OW travel together)
  return (size_t)GPB_DIST_MAX_OW * (nt + 2) * TILE_ELEMS * sizeof(double);
}

// The distributed factorisation.  dm: device descriptor (one GpbMat), h: its host copy, stage[2]: staging buffers of
// dist_stage_bytes(n) each.
cudaError_t run_potrf_dist_cols(const GpbMat* dm, const GpbMat& h, const DistCtx& D, double* const stage[2], const Exec& ex);

cudaError_t run_potrf_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, double* const stage[2], const Exec& ex) {
  if (D.P == 1) return run_potrf_dist_cols(dm, h, D, stage, ex);
  using Cfg = CfgHalf;
  constexpr int BM = Cfg::BM;
  const int n = h.n, nrows = h.n + h.aug;
  const int nblk = (n + GPB_NB - 1) / GPB_NB;              // pivot block columns
  const int nbr = (nrows + GPB_NB - 1) / GPB_NB;           // block rows (the carried y row may add one)
  const int P = D.P, Q = D.Q, p = D.p, q = D.q;
  cudaStream_t cs = ex.crit, ss = ex.side;
  GPB_CK(cudaEventRecord(ex.ev_fork, ex.main));
  GPB_CK(cudaStreamWaitEvent(cs, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(ss, ex.ev_fork, 0));
  const size_t ld = h.ld;
  // diagonal block k, its panel, and the exchange that leaves panel k (and inv(L_kk)) on every rank; all on the critical stream
  auto factor_block = [&](int k) -> cudaError_t {
    const int qk = k % Q, pk = k % P;
    const int diag_owner = pk * Q + qk;
    const bool full = (k + 1) * GPB_NB <= n;               // full pivot block: a panel exists below it
    const int nt = full ? nbr - k : 1;                     // tiles of the panel incl. the diagonal block
    double* st = stage[k & 1];
    PanelMap map;
    map.P = P;
    int seg_base[GPB_DIST_MAX_P], seg_count[GPB_DIST_MAX_P], seg_first[GPB_DIST_MAX_P];
    dist_panel_segments(k, nt, P, seg_base, seg_count, seg_first);
    for (int o = 0; o < P; ++o) { map.seg_base[o] = seg_base[o]; map.seg_first[o] = seg_first[o]; }
    double* st_wd = st + (size_t)nt * TILE_ELEMS;          // slot of inv(L_kk) behind the panel tiles
    double* Wk = h.Wd + (size_t)k * TILE_ELEMS;

    // ---- diagonal block (its owner), then the panel blocks (their owners in process column qk) ----------------
    if (D.rank == diag_owner) GPB_CK(run_diag(dm, 1, k, cs));
    if (P > 1 && full) {
      // the other process rows of column qk need inv(L_kk) for their panel blocks: ship {L_kk, inv(L_kk)} first
      if (D.rank == diag_owner) {
        panel_copy_kernel<<<1, 256, 0, cs>>>(h.A, (long long)ld, nrows, k, 0, st, map, p, qk, q, 0);
        tile_copy_kernel<<<8, 256, 0, cs>>>(st_wd, Wk);
        g_launches += 2;
      }
      GPB_TR(D.tr->group_start());
      GPB_TR(D.tr->broadcast(st + (size_t)map.seg_base[pk] * TILE_ELEMS, st + (size_t)map.seg_base[pk] * TILE_ELEMS,
                             TILE_ELEMS, diag_owner, cs));
      GPB_TR(D.tr->broadcast(st_wd, st_wd, TILE_ELEMS, diag_owner, cs));
      GPB_TR(D.tr->group_end());
      if (D.rank != diag_owner) {
        panel_copy_kernel<<<1, 256, 0, cs>>>(h.A, (long long)ld, nrows, k, 0, st, map, p, qk, q, 1);
        tile_copy_kernel<<<8, 256, 0, cs>>>(Wk, st_wd);
        g_launches += 2;
      }
    }
    if (full && q == qk) {
      const int Tm = (nrows - (k + 1) * GPB_NB + BM - 1) / BM;
      GPB_CK((launch_geo<Cfg>(GeoDistPanel{dm, k, P, p}, dim3(Tm, 1, 1), cs, false, "panel", k)));
    }
    // ---- exchange: pack own tiles, broadcast every process row's segment, unpack the others' tiles ------------
    const int t0 = (P > 1 && full) ? 1 : 0;                // the diagonal tile already travelled when P > 1
    const bool ship_wd = (t0 == 0);                        // inv(L_kk) is replicated too: the later stages need all of them
    if (nt - t0 > 0) {
      if (q == qk) {
        TraceSpan span("pack", cs, k, nt - t0);
        panel_copy_kernel<<<nt - t0, 256, 0, cs>>>(h.A, (long long)ld, nrows, k, t0, st, map, p, qk, q, 0);
        ++g_launches;
      }
      if (ship_wd && D.rank == diag_owner) { tile_copy_kernel<<<8, 256, 0, cs>>>(st_wd, Wk); ++g_launches; }
      {
        TraceSpan span("bcast", cs, k, nt - t0);
        GPB_TR(D.tr->group_start());
        if (ship_wd) GPB_TR(D.tr->broadcast(st_wd, st_wd, TILE_ELEMS, diag_owner, cs));
        for (int o = 0; o < P; ++o) {
          int first = seg_base[o], cnt = seg_count[o];
          if (t0 == 1 && o == pk) { first += 1; cnt -= 1; }  // skip the diagonal tile (slot 0 of its owner's segment)
          if (cnt <= 0) continue;
          double* b = st + (size_t)first * TILE_ELEMS;
          GPB_TR(D.tr->broadcast(b, b, (size_t)cnt * TILE_ELEMS, o * Q + qk, cs));
        }
        GPB_TR(D.tr->group_end());
      }
      TraceSpan span("unpack", cs, k, nt - t0);
      panel_copy_kernel<<<nt - t0, 256, 0, cs>>>(h.A, (long long)ld, nrows, k, t0, st, map, p, qk, q, 1);
      ++g_launches;
      if (ship_wd && D.rank != diag_owner) { tile_copy_kernel<<<8, 256, 0, cs>>>(Wk, st_wd); ++g_launches; }
    }
    return cudaSuccess;
  };
  const int Jend = nbr;
  auto syrk = [&](int kp, int kb, int Jlo, int Jhi, cudaStream_t s, const char* tag) -> cudaError_t {
    if (Jhi > Jend) Jhi = Jend;
    const int nc = count_owned_cols(Jlo, Jhi, Q, q);
    if (nc == 0) return cudaSuccess;
    const int Tm = (nrows - Jlo * GPB_NB + BM - 1) / BM;
    return launch_geo<Cfg>(GeoDistSyrk{dm, kp, kb, Jlo, Jhi, P, Q, p, q, 1}, dim3(Tm, nc, 1), s, true, tag, kp);
  };
  const bool wide = potrf_outer_blocks(n) > 1;             // 256-wide outer panels (two panels per far update)
  int step = 0;
  for (int k = 0; k < nblk;) {
    GPB_CK(factor_block(k));
    if ((k + 1) * GPB_NB > n) { ++k; continue; }           // last, partial block: no panel below it
    int kb = 1;
    bool waited = false;
    if (wide && (k + 2) * GPB_NB <= n) {
      // block column k+1 received its far update of the previous outer step on the side stream
      if (step > 0) { GPB_CK(cudaStreamWaitEvent(cs, ex.ev_g[(step - 1) & 1], 0)); waited = true; }
      GPB_CK(syrk(k, 1, k + 1, k + 2, cs, "strip"));               // strip: block column k+1 <- panel k
      GPB_CK(factor_block(k + 1));
      kb = 2;
    }
    GPB_CK(cudaEventRecord(ex.ev_e[step & 1], cs));
    // ---- far update with both panels: the next diagonal column on the critical stream; on the side stream first the
    //      columns the next outer step touches, then the rest
    const int J1 = k + kb;
    if (step > 0 && !waited) GPB_CK(cudaStreamWaitEvent(cs, ex.ev_g[(step - 1) & 1], 0));
    GPB_CK(syrk(k, kb, J1, J1 + 1, cs, "col"));
    GPB_CK(cudaStreamWaitEvent(ss, ex.ev_e[step & 1], 0));
    GPB_CK(syrk(k, kb, J1 + 1, J1 + 1 + kb, ss, "next"));
    GPB_CK(cudaEventRecord(ex.ev_g[step & 1], ss));
    GPB_CK(syrk(k, kb, J1 + 1 + kb, Jend, ss, "bulk"));
    k += kb;
    ++step;
  }
  GPB_CK(cudaEventRecord(ex.ev_join[0], cs));
  GPB_CK(cudaEventRecord(ex.ev_join[1], ss));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[0], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[1], 0));
  return cudaSuccess;
}


// ---------------------------------------------------------------------------------------------------------------
// The factorisation on a 1 x Q grid (P == 1, the default).  Block columns are owned in groups of OW; outer step s
// factorises group s = block columns [s OW, (s+1) OW) entirely on its owner (diagonal block, panel, left-looking strip
// update of the next column of the group - no exchange inside the group) and ships all of it with ONE broadcast on a
// communication stream of its own, so that
//   * a panel travels once per OW block columns (one NCCL launch of OW x the bytes instead of OW launches),
//   * the chain per outer step is  [factorise OW columns locally] -> broadcast -> [update the next group's columns on its
//     owner]  instead of OW x (factorise -> broadcast -> update) with a hand-over between ranks at every block column,
//   * the owner's own updates start from its local copy and do not wait for the broadcast.
// Look-ahead of depth 2 as on one GPU: with the panels of step s every rank updates its own columns of group s+1 on
// the critical stream (A), of group s+2 on the medium-priority stream (B) and the rest on the low-priority one (C);
// A(s) waits for B(s-1), B(s) for C(s-1), so the owner of group s+1 starts factorising while bulk updates still run.
// ---------------------------------------------------------------------------------------------------------------
static int post_gate() {      // GPB_DIST_GATE=0 switches the receivers' gate off
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPB_DIST_GATE"); v = e ? atoi(e) : 1; }
  return v;
}

cudaError_t run_potrf_dist_cols(const GpbMat* dm, const GpbMat& h, const DistCtx& D, double* const stage[2], const Exec& ex) {
  using Cfg = CfgHalf;
  constexpr int BM = Cfg::BM;
  const int n = h.n, nrows = h.n + h.aug;
  const int nblk = (n + GPB_NB - 1) / GPB_NB;              // pivot block columns
  const int nbr = (nrows + GPB_NB - 1) / GPB_NB;           // block rows (the carried y row may add one)
  const int Q = D.Q, q = D.q, OW = D.OW;
  cudaStream_t cs = ex.crit, ms = ex.mid, ss = ex.side, xs = ex.comm;
  GPB_CK(cudaEventRecord(ex.ev_fork, ex.main));
  GPB_CK(cudaStreamWaitEvent(cs, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(ms, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(ss, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(xs, ex.ev_fork, 0));
  const size_t ld = h.ld;
  auto syrk = [&](int kp, int kb, int Jlo, int Jhi, cudaStream_t s, const char* tag) -> cudaError_t {
    if (Jhi > nbr) Jhi = nbr;
    const int nc = count_owned_cols(Jlo, Jhi, Q, q, OW);
    if (nc == 0) return cudaSuccess;
    const int Tm = (nrows - Jlo * GPB_NB + BM - 1) / BM;
    return launch_geo<Cfg>(GeoDistSyrk{dm, kp, kb, Jlo, Jhi, 1, Q, 0, q, OW}, dim3(Tm, nc, 1), s, true, tag, kp);
  };
  const int ngroups = (nblk + OW - 1) / OW;
  for (int s = 0; s < ngroups; ++s) {
    const int k0 = s * OW, kb = std::min(OW, nblk - k0), e = s & 1;
    const int owner = s % Q;                               // == (k0 / OW) % Q
    const bool mine = (q == owner);
    double* st = stage[e];
    size_t off[GPB_DIST_MAX_OW + 1];                       // exchange buffer: per block column its tiles, then inv(L_kk)
    off[0] = 0;
    for (int j = 0; j < kb; ++j) off[j + 1] = off[j] + (size_t)(nbr - (k0 + j) + 1) * TILE_ELEMS;
    if (mine) {
      for (int j = 0; j < kb; ++j) {
        const int k = k0 + j, nt = nbr - k;
        if (j > 0) GPB_CK(syrk(k0, j, k, k + 1, cs, "strip"));     // column k <- the j panels of the group before it
        GPB_CK(run_diag(dm, 1, k, cs));
        if ((k + 1) * GPB_NB <= n) {                               // full pivot block: a panel exists below it
          const int Tm = (nrows - (k + 1) * GPB_NB + BM - 1) / BM;
          GPB_CK((launch_geo<Cfg>(GeoDistPanel{dm, k, 1, 0}, dim3(Tm, 1, 1), cs, false, "panel", k)));
        }
        // block column k is final: it is packed on the communication stream while the next column of the group is
        // being factorised (ev_p is re-recorded per column; the wait is enqueued right behind each record)
        GPB_CK(cudaEventRecord(ex.ev_p[e], cs));
        GPB_CK(cudaStreamWaitEvent(xs, ex.ev_p[e], 0));
        TraceSpan span("pack", xs, k, nt);
        panel_copy2_kernel<<<dim3(nt, 4), 256, 0, xs>>>(h.A, (long long)ld, nrows, k, st + off[j], 0);
        tile_copy_kernel<<<8, 256, 0, xs>>>(st + off[j] + (size_t)nt * TILE_ELEMS, h.Wd + (size_t)k * TILE_ELEMS);
        g_launches += 2;
        GPB_CK(cudaGetLastError());
      }
    }
    // A receiver posts its side of the broadcast only when its own bulk update of step s-2 has finished: that is when
    // the owner of this group could start factorising it (its A(s-1) waits for B(s-2), which waits for C(s-3)), give or
    // take one step of load balance.  Posted earlier, the NCCL kernel spins for the owner's data for milliseconds and
    // its CTAs (whole register files) take 16 - 24 SMs away from the bulk update (launch timeline on 2 GPUs,
    // n = 32768: receive spans of 4 - 5 ms per outer step, bulk update at 24 instead of 31 TFLOP/s).  GPB_DIST_GATE=0: off.
    // Gating the sender the same way measured slightly worse (2 GPUs: potrf 213.5 vs 210.4 ms).
    if (s >= 2 && !mine && post_gate() >= 1) GPB_CK(cudaStreamWaitEvent(xs, ex.ev_d[e], 0));
    {
      TraceSpan span("bcast", xs, k0, kb);
      GPB_TR(D.tr->broadcast(st, st, off[kb], owner, xs));
    }
    if (!mine) {
      TraceSpan span("unpack", xs, k0, kb);
      for (int j = 0; j < kb; ++j) {
        const int k = k0 + j, nt = nbr - k;
        panel_copy2_kernel<<<dim3(nt, 4), 256, 0, xs>>>(h.A, (long long)ld, nrows, k, st + off[j], 1);
        tile_copy_kernel<<<8, 256, 0, xs>>>(h.Wd + (size_t)k * TILE_ELEMS, st + off[j] + (size_t)nt * TILE_ELEMS);
        g_launches += 2;
      }
      GPB_CK(cudaGetLastError());
    }
    GPB_CK(cudaEventRecord(ex.ev_x[e], xs));
    // ---- updates of the own columns with the kb panels of this group -------------------------------------------
    const int J1 = k0 + kb;
    if (J1 >= nbr) continue;
    cudaEvent_t ready = mine ? ex.ev_p[e] : ex.ev_x[e];    // the owner's copy is complete before the broadcast
    if (!mine) GPB_CK(cudaStreamWaitEvent(cs, ready, 0));
    if (s > 0) GPB_CK(cudaStreamWaitEvent(cs, ex.ev_b[(s - 1) & 1], 0));
    GPB_CK(syrk(k0, kb, J1, J1 + OW, cs, "A"));
    GPB_CK(cudaStreamWaitEvent(ms, ready, 0));
    if (s > 0) GPB_CK(cudaStreamWaitEvent(ms, ex.ev_d[(s - 1) & 1], 0));
    GPB_CK(syrk(k0, kb, J1 + OW, J1 + 2 * OW, ms, "B"));
    GPB_CK(cudaEventRecord(ex.ev_b[e], ms));
    GPB_CK(cudaStreamWaitEvent(ss, ready, 0));
    GPB_CK(syrk(k0, kb, J1 + 2 * OW, nbr, ss, "bulk"));
    GPB_CK(cudaEventRecord(ex.ev_d[e], ss));
  }
  GPB_CK(cudaEventRecord(ex.ev_join[0], cs));
  GPB_CK(cudaEventRecord(ex.ev_join[1], ss));
  GPB_CK(cudaEventRecord(ex.ev_join[3], ms));
  GPB_CK(cudaEventRecord(ex.ev_join_comm, xs));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[0], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[1], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[3], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join_comm, 0));
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// COLUMN STORAGE (1 x Q grids, likelihood only): every rank keeps ONLY its own block columns (groups of OW, packed in
// ascending order: n^2 * 8 / Q bytes instead of n^2 * 8) plus a ring of three outer-panel buffers.  Same schedule as
// run_potrf_dist_cols, with three differences:
//   * no pack / unpack: the columns of an outer panel are adjacent in the owner's storage, so the owner broadcasts them
//     in place (whole columns, OW * 128 * ld doubles) straight into the receivers' ring slot; the updates read the panel
//     from there (the owner: from its own columns);
//   * the diagonal-block kernel, the panel product and the strip update address block column k as d.A + k * 128 * ld:
//     they get a descriptor per own group whose A is shifted so that this expression lands in the packed storage;
//   * nothing is replicated, so the log-determinant, z^T z and the first bad pivot are reduced over the ranks
//     (run_finalize_dist_store); inverse and gradient stages need the replicated plan.
// A ring slot is reused by the third-next outer step: its broadcast waits for the three updates (A, B, C) that read it.
// ---------------------------------------------------------------------------------------------------------------
struct GeoColSyrk {
  const double* panel;      // element (row 0, first column of the outer panel); column-major, ld
  double* Aloc;             // packed own block columns
  int ld, nrows, kb, J_lo, J_hi, Q, q, OW;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    static_assert(BN == GPB_NB, "block columns are 128 wide");
    const int t = owned_cols_below(J_lo, Q, q, OW) + (int)b.y;          // local index of the b.y-th own column >= J_lo
    const int Jb = owned_col_at(t, Q, q, OW);
    if (Jb >= J_hi) return false;
    const int row = Jb * GPB_NB + (int)b.x * BM;
    if (row >= nrows) return false;
    J.A = panel + row;
    J.B = panel + (size_t)Jb * GPB_NB;
    J.C = Aloc + row + (size_t)t * GPB_NB * ld;
    J.lda = J.ldb = J.ldc = ld;
    J.mrem = min(BM, nrows - row);
    J.nrem = min(BN, nrows - Jb * GPB_NB);
    J.klo = 0; J.khi = kb * GPB_NB;
    J.alpha = -1.0; J.beta = 1.0; J.red = g_red_epilogue;
    return true;
  }
};

size_t dist_ring_bytes(int n, int OW) {
  const size_t ld = (size_t)((n + 1 + 7) / 8 * 8);
  return (size_t)OW * GPB_NB * ld * sizeof(double);
}

cudaError_t run_potrf_dist_store(const GpbMat* gdesc, const GpbMat& h, const DistCtx& D, double* const ring[3], const Exec& ex) {
  using Cfg = CfgHalf;
  constexpr int BM = Cfg::BM;
  const int n = h.n, nrows = h.n + h.aug;
  const int nblk = (n + GPB_NB - 1) / GPB_NB;
  const int nbr = (nrows + GPB_NB - 1) / GPB_NB;
  const int Q = D.Q, q = D.q, OW = D.OW;
  cudaStream_t cs = ex.crit, ms = ex.mid, ss = ex.side, xs = ex.comm;
  GPB_CK(cudaEventRecord(ex.ev_fork, ex.main));
  GPB_CK(cudaStreamWaitEvent(cs, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(ms, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(ss, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(xs, ex.ev_fork, 0));
  const size_t ld = h.ld;
  double* Aloc = h.A;
  auto syrk = [&](const double* panel, int kb, int Jlo, int Jhi, cudaStream_t s, const char* tag, int k0) -> cudaError_t {
    if (Jhi > nbr) Jhi = nbr;
    const int nc = count_owned_cols(Jlo, Jhi, Q, q, OW);
    if (nc == 0) return cudaSuccess;
    const int Tm = (nrows - Jlo * GPB_NB + BM - 1) / BM;
    return launch_geo<Cfg>(GeoColSyrk{panel, Aloc, (int)ld, nrows, kb, Jlo, Jhi, Q, q, OW}, dim3(Tm, nc, 1), s, true, tag, k0);
  };
  const int ngroups = (nblk + OW - 1) / OW;
  for (int s = 0; s < ngroups; ++s) {
    const int k0 = s * OW, kb = std::min(OW, nblk - k0), e = s & 1, slot = s % 3;
    const int owner = s % Q;
    const bool mine = (q == owner);
    const int t0 = (s / Q) * OW;                           // local index of block column k0 on its owner
    double* own_panel = Aloc + (size_t)t0 * GPB_NB * ld;
    if (mine) {
      const GpbMat* gd = gdesc + s / Q;                    // A shifted: gd->A + k * 128 * ld is block column k of this group
      for (int j = 0; j < kb; ++j) {
        const int k = k0 + j;
        if (j > 0)                                         // column k <- the j panels of the group before it (all local)
          GPB_CK((launch_geo<Cfg>(GeoDistSyrk{gd, k0, j, k, k + 1, 1, 1, 0, 0, 1},
                                  dim3((nrows - k * GPB_NB + BM - 1) / BM, 1, 1), cs, true, "strip", k)));
        GPB_CK(run_diag(gd, 1, k, cs));
        if ((k + 1) * GPB_NB <= n) {
          const int Tm = (nrows - (k + 1) * GPB_NB + BM - 1) / BM;
          GPB_CK((launch_geo<Cfg>(GeoDistPanel{gd, k, 1, 0}, dim3(Tm, 1, 1), cs, false, "panel", k)));
        }
      }
      GPB_CK(cudaEventRecord(ex.ev_p[e], cs));
      GPB_CK(cudaStreamWaitEvent(xs, ex.ev_p[e], 0));
    } else {
      if (s >= 3) for (int u = 0; u < 3; ++u) GPB_CK(cudaStreamWaitEvent(xs, ex.ev_ring[slot][u], 0));   // slot free again
      if (s >= 2 && post_gate() >= 1) GPB_CK(cudaStreamWaitEvent(xs, ex.ev_d[e], 0));                    // see run_potrf_dist_cols
    }
    {
      TraceSpan span("bcast", xs, k0, kb);
      double* buf = mine ? own_panel : ring[slot];
      GPB_TR(D.tr->broadcast(buf, buf, (size_t)kb * GPB_NB * ld, owner, xs));
    }
    GPB_CK(cudaEventRecord(ex.ev_x[e], xs));
    const int J1 = k0 + kb;
    if (J1 >= nbr) continue;
    const double* panel = mine ? own_panel : ring[slot];
    cudaEvent_t ready = mine ? ex.ev_p[e] : ex.ev_x[e];
    if (!mine) GPB_CK(cudaStreamWaitEvent(cs, ready, 0));
    if (s > 0) GPB_CK(cudaStreamWaitEvent(cs, ex.ev_b[(s - 1) & 1], 0));
    GPB_CK(syrk(panel, kb, J1, J1 + OW, cs, "A", k0));
    GPB_CK(cudaEventRecord(ex.ev_ring[slot][0], cs));
    GPB_CK(cudaStreamWaitEvent(ms, ready, 0));
    if (s > 0) GPB_CK(cudaStreamWaitEvent(ms, ex.ev_d[(s - 1) & 1], 0));
    GPB_CK(syrk(panel, kb, J1 + OW, J1 + 2 * OW, ms, "B", k0));
    GPB_CK(cudaEventRecord(ex.ev_b[e], ms));
    GPB_CK(cudaEventRecord(ex.ev_ring[slot][1], ms));
    GPB_CK(cudaStreamWaitEvent(ss, ready, 0));
    GPB_CK(syrk(panel, kb, J1 + 2 * OW, nbr, ss, "bulk", k0));
    GPB_CK(cudaEventRecord(ex.ev_d[e], ss));
    GPB_CK(cudaEventRecord(ex.ev_ring[slot][2], ss));
  }
  GPB_CK(cudaEventRecord(ex.ev_join[0], cs));
  GPB_CK(cudaEventRecord(ex.ev_join[1], ss));
  GPB_CK(cudaEventRecord(ex.ev_join[3], ms));
  GPB_CK(cudaEventRecord(ex.ev_join_comm, xs));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[0], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[1], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[3], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join_comm, 0));
  return cudaSuccess;
}

// column storage: partial log-determinant, z^T z and first bad pivot over the OWN pivot columns -> red[0], red[1],
// red[2 + rank]; after the all-reduce (sum) every rank holds all partials and finishes the likelihood itself
__global__ void __launch_bounds__(1024) store_partials_kernel(const GpbMat* __restrict__ mats, double* __restrict__ red, int world,
                                                              int rank) {
  __shared__ double s_ld[32], s_q[32];
  __shared__ int s_bad[32];
  const GpbMat d = mats[0];
  const int n = d.n;
  const size_t ld = d.ld;
  double lsum = 0.0, qsum = 0.0;
  int bad = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (((i / GPB_NB) / d.own_W) % d.own_Q != d.own_q) continue;
    const size_t c = (size_t)gpb_local_col(i, d.own_Q, d.own_q, d.own_W);
    const double lii = d.A[i + c * ld];
    const double z = d.A[n + c * ld];
    if (!(lii > 0.0)) bad = min(bad, i + 1);
    lsum += log(lii);
    qsum += z * z;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  lsum = warp_sum(lsum); qsum = warp_sum(qsum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  if (lane == 0) { s_ld[warp] = lsum; s_q[warp] = qsum; s_bad[warp] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    int w = 0x7fffffff;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s_ld[i]; b += s_q[i]; w = min(w, s_bad[i]); }
    // a rank whose columns hold a bad pivot contributes NaN-free sums: the likelihood is NaN whenever info != 0
    red[0] = (w == 0x7fffffff) ? a : 0.0;
    red[1] = (w == 0x7fffffff) ? b : 0.0;
    for (int r = 0; r < world; ++r) red[2 + r] = 0.0;
    red[2 + rank] = (w == 0x7fffffff) ? 0.0 : (double)w;
  }
}
__global__ void store_finish_kernel(const GpbMat* __restrict__ mats, const double* __restrict__ red, int world, double log2pi) {
  const GpbMat d = mats[0];
  int info = 0;
  for (int r = 0; r < world; ++r) {
    const int w = (int)red[2 + r];
    if (w > 0 && (info == 0 || w < info)) info = w;
  }
  *d.info = info;
  *d.nll = info ? nan("") : 0.5 * red[1] + red[0] + 0.5 * ((double)d.n * log2pi);
  d.terms[0] = red[1]; d.terms[1] = red[0];
}

cudaError_t run_finalize_dist_store(const GpbMat* dm, const DistCtx& D, double* red, double log2pi, cudaStream_t s) {
  store_partials_kernel<<<1, 1024, 0, s>>>(dm, red, D.world, D.rank);
  ++g_launches;
  GPB_CK(cudaGetLastError());
  GPB_TR(D.tr->allreduce_sum(red, (size_t)(2 + D.world), s));
  store_finish_kernel<<<1, 1, 0, s>>>(dm, red, D.world, log2pi);
  ++g_launches;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient stages of a distributed plan.  After the factorisation every rank holds the complete L and all inverted
// diagonal blocks, so W = inv(L) and inv(K) = W^T W split by BLOCK COLUMN with no exchange inside a stage: rank r owns
// the block columns J = r, r + world, ...  (1D, independent of the P x Q grid of the factorisation).
//   trtri : X = W[:, J] for the own block columns by blocked forward substitution, right-looking, out of place in Kinv:
//             step I:  X[I, J] <- -Wd_I * X[I, J]  (J < I; X[I, I] = Wd_I),  X[r, J] += L[r, I] * X[I, J]  for r > I
//           (the second is a rank-128 update over all remaining rows - the same GEMM shape as the Cholesky trailing
//           update), then ONE exchange: every block column is broadcast from its owner's Kinv into everybody's A.
//   lauum : inv(K)[:, J] = W^T W[:, J] for the own block columns into Kinv (inputs: the full W, no exchange).
//   grad  : the trace-gradient kernel over the own block columns, then one ncclAllReduce of n_hp + 1 doubles.
// ---------------------------------------------------------------------------------------------------------------
struct GeoDistTriDiag {     // T_J = -Wd_I * X[I, J]      (out of place: X[I, J] is the B operand)
  const GpbMat* mats;
  double* scratch;          // [nJ][128 x 128]
  int I, world, rank;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    static_assert(BN == GPB_NB, "block columns are 128 wide");
    const GpbMat& d = mats[0];
    const int Jb = rank + (int)b.y * world;
    if (Jb >= I) return false;
    const int rows = min(GPB_NB, d.n - I * GPB_NB);
    const int i0 = (int)b.x * BM;
    if (i0 >= rows) return false;
    J.A = d.Wd + (size_t)I * GPB_NB * GPB_NB + i0;                     // Wd_I rows i0.., MN-major, ld 128
    J.B = d.Kinv + (size_t)I * GPB_NB + (size_t)Jb * GPB_NB * d.ld;    // X[I, J]: element (col, k) at [k + col * ld]
    J.C = scratch + (size_t)b.y * GPB_NB * GPB_NB + i0;
    J.lda = GPB_NB; J.ldb = d.ld; J.ldc = GPB_NB;
    J.mrem = min(BM, rows - i0); J.nrem = GPB_NB;
    J.klo = 0; J.khi = min(rows, i0 + BM);                              // Wd_I is lower triangular
    J.alpha = -1.0; J.beta = 0.0; J.red = 0;
    return true;
  }
};

struct GeoDistTriUpdate {   // X[r, J] += L[r, I .. I+kb) * X[I .. I+kb, J]  for rows r in [row_lo, row_hi), own J < I + kb
  const GpbMat* mats;
  int I, kb, row_lo, row_hi, world, rank;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    static_assert(BN == GPB_NB, "block columns are 128 wide");
    const GpbMat& d = mats[0];
    const int Jb = rank + (int)b.y * world;
    if (Jb > I + kb - 1) return false;
    const int row = row_lo + (int)b.x * BM;
    if (row >= row_hi) return false;
    const size_t ld = d.ld;
    J.A = d.A + row + (size_t)I * GPB_NB * ld;                         // L[r, I..], MN-major
    J.B = d.Kinv + (size_t)I * GPB_NB + (size_t)Jb * GPB_NB * ld;      // X[I.., J], K-major (rows above J's diagonal
    J.C = d.Kinv + row + (size_t)Jb * GPB_NB * ld;                     //  block are the zeros of W)
    J.lda = J.ldb = J.ldc = d.ld;
    J.mrem = min(BM, row_hi - row); J.nrem = GPB_NB;
    J.klo = 0; J.khi = kb * GPB_NB;
    J.alpha = 1.0; J.beta = 1.0; J.red = g_red_epilogue;
    return true;
  }
};

struct GeoDistLauum {       // inv(K)[ti, J] = sum_{k >= ti} W[k, ti]^T W[k, J] for the own block columns J, rows in [row_lo, row_hi)
  const GpbMat* mats;
  int world, rank, row_lo, row_hi;
  template <int BM, int BN>
  __device__ bool tile(TileJob& J, const dim3& b) const {
    static_assert(BN == GPB_NB, "block columns are 128 wide");
    const GpbMat& d = mats[0];
    const int Jb = rank + (int)b.y * world;
    const int col0 = Jb * GPB_NB;
    if (col0 >= d.n) return false;
    const int row = max(col0, row_lo) + (int)b.x * BM;                  // lower tiles: rows from the diagonal block down
    if (row >= d.n || row >= row_hi) return false;
    const size_t ld = d.ld;
    J.A = d.A + (size_t)row * ld;
    J.B = d.A + (size_t)col0 * ld;
    J.C = d.Kinv + row + (size_t)col0 * ld;
    J.lda = J.ldb = J.ldc = d.ld;
    J.mrem = min(BM, d.n - row); J.nrem = min(BN, d.n - col0);
    J.klo = row; J.khi = d.n;
    J.alpha = 1.0; J.beta = 0.0; J.red = 0;
    return true;
  }
};

// X[I, J] <- T_J for the own J < I, X[I, I] <- Wd_I when block column I is own
__global__ void __launch_bounds__(256) tri_diag_store_kernel(const GpbMat* __restrict__ mats, const double* __restrict__ scratch,
                                                             int I, int world, int rank) {
  const GpbMat d = mats[0];
  const int Jb = rank + (int)blockIdx.x * world;
  if (Jb > I) return;
  const int rows = min(GPB_NB, d.n - I * GPB_NB);
  const double* src = (Jb == I) ? d.Wd + (size_t)I * GPB_NB * GPB_NB : scratch + (size_t)blockIdx.x * GPB_NB * GPB_NB;
  const int cols = min(GPB_NB, d.n - Jb * GPB_NB);
  double* dst = d.Kinv + (size_t)I * GPB_NB + (size_t)Jb * GPB_NB * d.ld;
  for (int idx = threadIdx.x; idx < GPB_NB * GPB_NB; idx += 256) {
    const int i = idx & (GPB_NB - 1), c = idx >> 7;
    if (i < rows && c < cols) dst[i + (size_t)c * d.ld] = src[idx];
  }
}

template <class Cfg, bool AKM, bool BKM, class Geo>
static cudaError_t launch_geo2(const Geo& geo, dim3 grid, cudaStream_t s, const char* tag = "dgemm2", int ta = 0) {
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return cudaSuccess;
  TraceSpan span(tag, s, ta, (int)(grid.x * grid.y * grid.z));
  static bool attr_set = false;
  if (!attr_set) {
    GPB_CK(cudaFuncSetAttribute(gemm_kernel<Cfg, AKM, BKM, Geo>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg::SMEM_BYTES));
    attr_set = true;
  }
  gemm_kernel<Cfg, AKM, BKM, Geo><<<persistent_ctas(grid, Cfg::MIN_CTAS, 1), Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(geo, grid);
  ++g_launches;
  return cudaGetLastError();
}

static int own_cols_upto(int J_incl, int world, int rank) {   // own block columns J <= J_incl
  return J_incl < rank ? 0 : (J_incl - rank) / world + 1;
}

cudaError_t run_trtri_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, double* scratch, const Exec& ex, bool exchange) {
  using Cfg = CfgHalf;
  const int n = h.n, world = D.world, rank = D.rank;
  const int nblk = (n + GPB_NB - 1) / GPB_NB;
  const size_t ld = h.ld;
  cudaStream_t s = ex.main, cs = ex.crit, ss = ex.side;
  // zero the own block columns of X (the accumulation starts from 0; rows above the diagonal block stay 0 = W's zeros)
  for (int J = rank; J < nblk; J += world) {
    const int cols = std::min(GPB_NB, n - J * GPB_NB);
    GPB_CK(cudaMemsetAsync(h.Kinv + (size_t)J * GPB_NB * ld, 0, (size_t)cols * ld * sizeof(double), s));
  }
  GPB_CK(cudaEventRecord(ex.ev_fork, s));
  GPB_CK(cudaStreamWaitEvent(cs, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(ss, ex.ev_fork, 0));
  auto diag_solve = [&](int I) -> cudaError_t {            // X[I, J] <- -Wd_I X[I, J] (own J < I), X[I, I] <- Wd_I
    const int nJ_lt = own_cols_upto(I - 1, world, rank);
    const int nJ_le = own_cols_upto(I, world, rank);
    if (nJ_le == 0) return cudaSuccess;
    if (nJ_lt > 0) GPB_CK((launch_geo2<Cfg, false, true>(GeoDistTriDiag{dm, scratch, I, world, rank}, dim3(GPB_NB / Cfg::BM, nJ_lt, 1), cs, "tri_diag", I)));
    tri_diag_store_kernel<<<nJ_le, 256, 0, cs>>>(dm, scratch, I, world, rank);
    ++g_launches;
    return cudaGetLastError();
  };
  auto update = [&](int I, int kb, int row_lo, int row_hi, cudaStream_t st, const char* tag) -> cudaError_t {
    const int nJ = own_cols_upto(I + kb - 1, world, rank);
    if (nJ == 0 || row_hi <= row_lo) return cudaSuccess;
    return launch_geo2<Cfg, false, true>(GeoDistTriUpdate{dm, I, kb, row_lo, row_hi, world, rank},
                                         dim3((row_hi - row_lo + Cfg::BM - 1) / Cfg::BM, nJ, 1), st, tag, I);
  };
  // Look-ahead: the block rows of the NEXT outer step receive this step's update on the critical stream, right behind
  // the two small triangular products, so that the next step's products run while the rest of the update (all remaining
  // rows, low-priority stream) is still in flight.  A row receives the updates of the outer steps in order: the "next"
  // update of step t waits for the rest-update of step t-1 (ev_d), the rest-updates follow each other on their stream.
  const bool wide = potrf_outer_blocks(n) > 1;             // two block rows per far update (k = 256), as in the factorisation
  int step = 0;
  for (int I = 0; I < nblk; ++step) {
    GPB_CK(diag_solve(I));
    int kb = 1;
    if (wide && (I + 2) * GPB_NB <= n) {
      GPB_CK(update(I, 1, (I + 1) * GPB_NB, (I + 2) * GPB_NB, cs, "tri_strip"));   // strip: block row I+1
      GPB_CK(diag_solve(I + 1));
      kb = 2;
    }
    const int e = step & 1;
    GPB_CK(cudaEventRecord(ex.ev_e[e], cs));
    const int r1 = (I + kb) * GPB_NB, r2 = std::min(n, (I + kb + (wide ? 2 : 1)) * GPB_NB);
    if (step > 0) GPB_CK(cudaStreamWaitEvent(cs, ex.ev_d[(step - 1) & 1], 0));
    GPB_CK(update(I, kb, r1, r2, cs, "tri_next"));
    GPB_CK(cudaStreamWaitEvent(ss, ex.ev_e[e], 0));
    GPB_CK(update(I, kb, r2, n, ss, "tri_upd"));
    GPB_CK(cudaEventRecord(ex.ev_d[e], ss));
    I += kb;
  }
  GPB_CK(cudaEventRecord(ex.ev_join[0], cs));
  GPB_CK(cudaEventRecord(ex.ev_join[1], ss));
  GPB_CK(cudaStreamWaitEvent(s, ex.ev_join[0], 0));
  GPB_CK(cudaStreamWaitEvent(s, ex.ev_join[1], 0));
  if (!exchange) return cudaSuccess;   // run_exchange_lauum_dist ships W chunk by chunk under the W^T W product
  // exchange: block column J of W from its owner's Kinv into everybody's A (whole columns: contiguous, in place)
  for (int J0 = 0; J0 < nblk; J0 += 64) {
    TraceSpan span("w_bcast", s, J0, 64);
    GPB_TR(D.tr->group_start());
    for (int J = J0; J < std::min(nblk, J0 + 64); ++J) {
      const int cols = std::min(GPB_NB, n - J * GPB_NB);
      GPB_TR(D.tr->broadcast(h.Kinv + (size_t)J * GPB_NB * ld, h.A + (size_t)J * GPB_NB * ld, (size_t)cols * ld, J % world, s));
    }
    GPB_TR(D.tr->group_end());
  }
  return cudaSuccess;
}

cudaError_t run_lauum_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, cudaStream_t s) {
  using Cfg = CfgHalf;
  const int nblk = (h.n + GPB_NB - 1) / GPB_NB;
  const int nJ = own_cols_upto(nblk - 1, D.world, D.rank);
  const int Tm = (h.n + Cfg::BM - 1) / Cfg::BM;
  return launch_geo2<Cfg, true, true>(GeoDistLauum{dm, D.world, D.rank, 0, nblk * GPB_NB}, dim3(Tm, nJ, 1), s, "lauum", 0);
}

// The exchange of W = inv(L) (every block column from its owner's Kinv into everybody's A) overlapped with
// inv(K) = W^T W.  Block columns travel in ascending chunks on the communication stream; the tiles of inv(K) whose ROWS
// lie in chunk c need the columns of W up to chunk c only (tile (I, J), J <= I, contracts columns I and J of W), so
// their launch follows the arrival of chunk c while chunk c+1 is in flight.  No hazard: the product for chunk c writes
// Kinv[rows of c, own columns <= c], the broadcast of chunk c+1 reads Kinv columns of c+1 and writes A columns of c+1.
// Only the first chunk's transfer is exposed (n = 32768 on 8 GPUs: the exchange was 20 of 78 + 44 ms).
cudaError_t run_exchange_lauum_dist(const GpbMat* dm, const GpbMat& h, const DistCtx& D, const Exec& ex) {
  using Cfg = CfgHalf;
  const int n = h.n, world = D.world, rank = D.rank;
  const int nblk = (n + GPB_NB - 1) / GPB_NB;
  const size_t ld = h.ld;
  cudaStream_t xs = ex.comm, cs2[2] = {ex.mid, ex.side};
  GPB_CK(cudaEventRecord(ex.ev_fork, ex.main));
  GPB_CK(cudaStreamWaitEvent(xs, ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(cs2[0], ex.ev_fork, 0));
  GPB_CK(cudaStreamWaitEvent(cs2[1], ex.ev_fork, 0));
  int chunk = (nblk + 7) / 8;                              // 8 chunks, at least 8 block columns each
  if (chunk < 8) chunk = 8;
  int c = 0;
  for (int J0 = 0; J0 < nblk; J0 += chunk, ++c) {
    const int J1 = std::min(nblk, J0 + chunk);
    {
      TraceSpan span("w_bcast", xs, J0, J1 - J0);
      GPB_TR(D.tr->group_start());
      for (int J = J0; J < J1; ++J) {
        const int cols = std::min(GPB_NB, n - J * GPB_NB);
        GPB_TR(D.tr->broadcast(h.Kinv + (size_t)J * GPB_NB * ld, h.A + (size_t)J * GPB_NB * ld, (size_t)cols * ld, J % world, xs));
      }
      GPB_TR(D.tr->group_end());
    }
    // the products of consecutive chunks go to two streams: a chunk's last, partially filled wave of long tiles (up to
    // n deep) overlaps with the next chunk's first tiles instead of idling the machine once per chunk
    cudaStream_t s = cs2[c & 1];
    GPB_CK(cudaEventRecord(ex.ev_x[c & 1], xs));
    GPB_CK(cudaStreamWaitEvent(s, ex.ev_x[c & 1], 0));
    const int nJ = own_cols_upto(J1 - 1, world, rank);     // own block columns J <= the last row block of the chunk
    if (nJ > 0)
      GPB_CK((launch_geo2<Cfg, true, true>(GeoDistLauum{dm, world, rank, J0 * GPB_NB, J1 * GPB_NB},
                                           dim3((J1 - J0) * GPB_NB / Cfg::BM, nJ, 1), s, "lauum", J0)));
  }
  GPB_CK(cudaEventRecord(ex.ev_join[3], cs2[0]));
  GPB_CK(cudaEventRecord(ex.ev_join[1], cs2[1]));
  GPB_CK(cudaEventRecord(ex.ev_join_comm, xs));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[3], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join[1], 0));
  GPB_CK(cudaStreamWaitEvent(ex.main, ex.ev_join_comm, 0));
  return cudaSuccess;
}

cudaError_t run_grad_allreduce(double* grad, int count, const DistCtx& D, cudaStream_t s) {
  GPB_TR(D.tr->allreduce_sum(grad, (size_t)count, s));
  return cudaSuccess;
}

cudaError_t run_finalize_dist(const GpbMat* dm, double log2pi, cudaStream_t s) {
  dist_finalize_kernel<<<1, 1024, 0, s>>>(dm, log2pi);
  ++g_launches;
  return cudaGetLastError();
}

}  // namespace gpb
