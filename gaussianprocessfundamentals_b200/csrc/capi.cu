// C-ABI of libgpb (see include/gpb.h).  Host-side only: argument checking, workspace layout, launch sequencing.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler is attached

#include "../../include/gpb.h"
#include "dist.h"
#include "internal.h"
#include "jit.h"
#include "program.cuh"
#include "trace.h"

namespace gpb {
std::atomic<long long> g_launches{0};

// ---- launch timeline (trace.h) ---------------------------------------------------------------------------------------
struct TraceRec { const char* tag; int a, b; cudaStream_t s; cudaEvent_t e0, e1; };
static std::mutex g_trace_mu;
static std::vector<TraceRec> g_trace;
static std::atomic<bool> g_trace_on{false};
static cudaEvent_t g_trace_origin = nullptr;
bool trace_active() { return g_trace_on.load(std::memory_order_relaxed); }
int trace_open(const char* tag, cudaStream_t s, int a, int b) {
  TraceRec r{tag, a, b, s, nullptr, nullptr};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
  cudaEventRecord(r.e0, s);
  std::lock_guard<std::mutex> lk(g_trace_mu);
  g_trace.push_back(r);
  return (int)g_trace.size() - 1;
}
void trace_close(int id, cudaStream_t s) {
  cudaEvent_t e1;
  { std::lock_guard<std::mutex> lk(g_trace_mu); e1 = g_trace[id].e1; }
  cudaEventRecord(e1, s);
}
}  // namespace gpb

extern "C" int gpb_trace_begin(void* stream) {
  std::lock_guard<std::mutex> lk(gpb::g_trace_mu);
  for (auto& r : gpb::g_trace) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  gpb::g_trace.clear();
  if (!gpb::g_trace_origin && cudaEventCreate(&gpb::g_trace_origin) != cudaSuccess) return 1;
  if (cudaEventRecord(gpb::g_trace_origin, (cudaStream_t)stream) != cudaSuccess) return 1;
  gpb::g_trace_on = true;
  return 0;
}
// one line per span: "tag a b stream start_us end_us" (times relative to gpb_trace_begin; stream = index by first use)
extern "C" int gpb_trace_end(char* buf, size_t capacity, size_t* needed) {
  gpb::g_trace_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  std::lock_guard<std::mutex> lk(gpb::g_trace_mu);
  std::vector<cudaStream_t> streams;
  std::string out;
  char line[160];
  for (auto& r : gpb::g_trace) {
    int si = -1;
    for (size_t i = 0; i < streams.size(); ++i) if (streams[i] == r.s) si = (int)i;
    if (si < 0) { streams.push_back(r.s); si = (int)streams.size() - 1; }
    float t0 = 0.f, t1 = 0.f;
    cudaEventElapsedTime(&t0, gpb::g_trace_origin, r.e0);
    cudaEventElapsedTime(&t1, gpb::g_trace_origin, r.e1);
    snprintf(line, sizeof line, "%s %d %d %d %.1f %.1f\n", r.tag, r.a, r.b, si, t0 * 1e3, t1 * 1e3);
    out += line;
  }
  if (needed) *needed = out.size() + 1;
  if (buf && capacity) {
    const size_t ncopy = out.size() < capacity - 1 ? out.size() : capacity - 1;
    memcpy(buf, out.data(), ncopy);
    buf[ncopy] = 0;
  }
  return 0;
}

// one NVTX range per stage of an evaluation (host side: brackets the launches of the stage on the timeline)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static thread_local std::string g_err;
static int fail_arg(int k, const char* what) {
  g_err = std::string("invalid argument ") + std::to_string(k) + ": " + what;
  return -k;
}
static int fail_cuda(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return 1000 + (int)e;
}
#define CU(x, where) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail_cuda(e_, where); } while (0)

static std::mutex g_init_mu;
static bool g_inited = false;
static int ensure_init() {      // programs are created from several host threads (run-time compilation in parallel)
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_inited) return 0;
  CU(gpb::linalg_init(), "linalg_init");
  CU(gpb::assemble_init(), "assemble_init");
  g_inited = true;
  return 0;
}

struct gpb_program {
  int n_ops, dim, cp_mode, n_hp, tape;
  std::vector<int32_t> code;
  int32_t* code_dev;
  gpb::JitKernels jit;      // kernels specialised for this program (null: interpreter)
  std::string jit_note;     // why not, when not
};

// structural checks of a postfix program; n_hp, gradient-tape length and stack depth it needs
static int check_program(const int32_t* code, int n_ops, int dim, int cp_mode, int* n_hp_out, int* tape_out, int* depth_out) {
  if (!code) return fail_arg(1, "code is null");
  if (n_ops <= 0 || n_ops > GPB_MAX_OPS) return fail_arg(2, "n_ops out of range");
  if (dim < 1 || dim > GPB_MAX_DIM) return fail_arg(3, "dim out of range");
  if (cp_mode < 0 || cp_mode > 2) return fail_arg(4, "cp_mode out of range");
  int sp = 0, tape = 0, max_sp = 0, n_hp = 0;
  for (int pc = 0; pc < n_ops; ++pc) {
    const int32_t* w = code + pc * GPB_OP_WORDS;
    const int op = w[0];
    if (op >= GPB_OP_SE && op <= GPB_OP_L1) {
      const int nq = gpb_leaf_nhp(op, w[2], dim);
      if (w[1] < 0) return fail_arg(1, "negative hyper-parameter offset");
      if (w[1] + nq > n_hp) n_hp = w[1] + nq;
      tape += nq;
      ++sp;
    } else if (op == GPB_OP_ADD2 || op == GPB_OP_MUL2) {
      if (sp < 2) return fail_arg(1, "stack underflow");
      --sp;
      if (op == GPB_OP_MUL2) tape += 2;
    } else if (op == GPB_OP_CPW) {
      if (sp < 1) return fail_arg(1, "stack underflow");
      if (dim != 1) return fail_arg(3, "change-point operators need 1-d inputs (Operators.py:398)");
      if (w[3] < 2 || w[2] < 0 || w[2] >= w[3]) return fail_arg(1, "bad change-point child index");
      if (w[1] < 0) return fail_arg(1, "negative hyper-parameter offset");
      if (w[1] + w[3] - 1 > n_hp) n_hp = w[1] + w[3] - 1;
      tape += 4;
    } else {
      return fail_arg(1, "unknown opcode");
    }
    if (sp > max_sp) max_sp = sp;
  }
  if (sp != 1) return fail_arg(1, "program does not leave exactly one value");
  if (n_hp > GPB_MAX_HP) return fail_arg(1, "too many hyper-parameters");
  *n_hp_out = n_hp; *tape_out = tape; *depth_out = max_sp;
  return 0;
}

enum { NBUF = 12 };
enum { HOLDS_NONE = 0, HOLDS_K = 1, HOLDS_L = 2, HOLDS_W = 3 };
struct PlanMat {
  int64_t n;
  int ld, nblk, n_hp, n_gtiles;
  size_t off[NBUF], bytes[NBUF];
  size_t off_wd, off_part, off_gpart, off_tmpv;
};

struct gpb_plan {
  int B, want_grad;
  std::vector<PlanMat> mats;
  std::vector<const gpb_program*> progs;
  size_t ws_bytes, off_desc, off_in, in_bytes, off_out, out_bytes;
  size_t off_hp_all, off_noise_all, off_nll_all, off_grad_all, off_info_all, off_terms_all;
  size_t off_data, data_bytes;   // X and y of all GPs, contiguous: one H2D when the caller's buffers share the layout
  int holds;                     // what GPB_BUF_A currently holds (HOLDS_*): stages are checked against it
  int have_kinv;                 // GPB_BUF_KINV holds inv(K) of the current factorisation
  std::vector<double> gw;        // [2 B] gradient weights (quad, logdet) per GP
  // GPs that share a kernel program are assembled / differentiated by one launch of that program's kernels
  struct Group { const gpb_program* prog; std::vector<int> idx; int n_max; size_t off_which; };
  std::vector<Group> groups;
  std::vector<size_t> hp_prefix, grad_prefix;
  char* ws;
  int n_max, n_hp_max, n_ops_max, dim;
  gpb::Exec ex;
  bool own_streams;
  char* h_in;   // pinned mirror of the small input region
  char* h_out;  // pinned mirror of the small output region
  // CUDA graph of the launch sequence of gpb_plan_eval_host (one per stage mask), replayed on a private stream
  cudaStream_t gstream;
  cudaEvent_t g_in, g_out;
  int graph_stages[4];
  cudaGraphExec_t graph_exec[4];
  int n_graphs, graphs_off;
  gpb::DistCtx* dist;       // non-null: ONE GP factorised over a process grid (dist.cu)
  size_t off_stage[2];      // panel staging buffers of a distributed plan
  int col_storage;          // distributed plan that keeps only the own block columns (likelihood only; dist.cu)
  size_t off_ring[3], off_gdesc;   // column storage: ring of outer-panel buffers, descriptors of the own column groups
  int n_own_groups;
  GpbMat h_desc0;           // host copy of the first descriptor (distributed plans)
  std::vector<GpbMat> h_desc;   // host copy of all descriptors
};

struct gpb_dist {
  gpb::DistCtx* ctx;
};

// What the factorisation buffer holds decides which stages may run (a stage on the wrong content would read garbage
// silently): K -> POTRF -> L -> {NLL, BACKSOLVE, TRTRI} ; TRTRI -> W -> {LAUUM} ; LAUUM -> inv(K) -> {GRAD}.
// The single-GPU inverse leaves the carried row z^T intact, so NLL alone stays legal after it; the distributed exchange
// of W overwrites it.
static int advance_state(gpb_plan* p, int stages) {
  int holds = p->holds, kinv = p->have_kinv;
  if (stages & GPB_STAGE_ASSEMBLE) { holds = HOLDS_K; kinv = 0; }
  if (stages & GPB_STAGE_POTRF) {
    if (holds != HOLDS_K) return fail_arg(2, "POTRF needs a freshly assembled matrix (run GPB_STAGE_ASSEMBLE first)");
    holds = HOLDS_L;
  }
  if ((stages & GPB_STAGE_BACKSOLVE) && holds != HOLDS_L)
    return fail_arg(2, "BACKSOLVE needs the Cholesky factor: the buffer holds its inverse (or nothing); re-run ASSEMBLE | POTRF");
  if ((stages & GPB_STAGE_NLL) && holds != HOLDS_L && !(holds == HOLDS_W && !p->dist))
    return fail_arg(2, "NLL needs the Cholesky factor (run ASSEMBLE | POTRF first)");
  if (stages & (GPB_STAGE_INVERSE | GPB_STAGE_TRTRI)) {
    if (holds != HOLDS_L) return fail_arg(2, "TRTRI needs the Cholesky factor (run ASSEMBLE | POTRF first)");
    holds = HOLDS_W;
  }
  if (stages & (GPB_STAGE_INVERSE | GPB_STAGE_LAUUM)) {
    if (holds != HOLDS_W) return fail_arg(2, "LAUUM needs W = inv(L) (run TRTRI first)");
    kinv = 1;
  }
  if ((stages & GPB_STAGE_GRAD) && !kinv) return fail_arg(2, "GRAD needs inv(K) (run INVERSE first)");
  p->holds = holds; p->have_kinv = kinv;
  return 0;
}

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" {

int gpb_version(void) { return 100; }
const char* gpb_last_error(void) { return g_err.c_str(); }
long long gpb_launch_count(void) { return gpb::g_launches.load(); }

int gpb_program_create(const int32_t* code, int n_ops, int dim, int cp_mode, gpb_program_t** out) {
  if (!out) return fail_arg(5, "out is null");
  int n_hp = 0, tape = 0, depth = 0;
  int rc = check_program(code, n_ops, dim, cp_mode, &n_hp, &tape, &depth);
  if (rc) return rc;
  // the value-only interpreter (rectangular assembly: K_s, K_ss, get_K) keeps its operand stack in 8 registers
  if (depth > GPB_MAX_STACK) return fail_arg(1, "operand stack too deep");
  rc = ensure_init();
  if (rc) return rc;
  gpb_program* p = new gpb_program;
  p->n_ops = n_ops; p->dim = dim; p->cp_mode = cp_mode; p->n_hp = n_hp; p->tape = tape;
  p->code.assign(code, code + (size_t)n_ops * GPB_OP_WORDS);
  p->code_dev = nullptr;
  // kernels specialised for this program (assembly + trace gradient of the plans); the interpreter stays the fallback
  std::string why;
  if (gpb::jit_build(code, n_ops, dim, cp_mode, n_hp, p->jit, why)) p->jit_note = why;
  if (!p->jit.grad && tape > GPB_MAX_TAPE) {
    g_err = "gradient tape of " + std::to_string(tape) + " entries exceeds the interpreter's " + std::to_string(GPB_MAX_TAPE) +
            " and no specialised kernel could be built: " + why;
    delete p;
    return -1;
  }
  cudaError_t e = cudaMalloc(&p->code_dev, p->code.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemcpy(p->code_dev, p->code.data(), p->code.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->code_dev) cudaFree(p->code_dev);
    gpb::jit_release(p->jit);
    delete p;
    return fail_cuda(e, "gpb_program_create");
  }
  *out = p;
  return 0;
}

int gpb_program_is_specialised(const gpb_program_t* prog) { return (prog && prog->jit.grad) ? 1 : 0; }

const char* gpb_program_jit_note(const gpb_program_t* prog) { return prog ? prog->jit_note.c_str() : ""; }

int gpb_jit_available(void) { return gpb::jit_enabled(nullptr) ? 1 : 0; }

int gpb_jit_set_nvrtc_path(const char* path) {
  gpb::jit_set_nvrtc_path(path);
  return 0;
}

int gpb_jit_source(const int32_t* code, int n_ops, int dim, int cp_mode, char* buf, size_t capacity, size_t* needed) {
  int n_hp = 0, tape = 0, depth = 0;
  int rc = check_program(code, n_ops, dim, cp_mode, &n_hp, &tape, &depth);
  if (rc) return rc;
  std::string src, err;
  if (gpb::jit_generate(code, n_ops, dim, cp_mode, n_hp, src, err)) { g_err = err; return -1; }
  if (needed) *needed = src.size() + 1;
  if (buf && capacity > 0) {
    const size_t k = src.size() + 1 <= capacity ? src.size() : capacity - 1;
    memcpy(buf, src.data(), k);
    buf[k] = 0;
  }
  return 0;
}

int gpb_jit_cubin(const int32_t* code, int n_ops, int dim, int cp_mode, const char* arch, void* buf, size_t capacity,
                  size_t* needed) {
  if (!arch) return fail_arg(5, "arch is null");
  int n_hp = 0, tape = 0, depth = 0;
  int rc = check_program(code, n_ops, dim, cp_mode, &n_hp, &tape, &depth);
  if (rc) return rc;
  std::string src, err;
  if (gpb::jit_generate(code, n_ops, dim, cp_mode, n_hp, src, err)) { g_err = err; return -1; }
  std::vector<char> cubin;
  std::string log;
  if (gpb::jit_compile(src, arch, cubin, log)) { g_err = log; return 3000; }
  if (needed) *needed = cubin.size();
  if (buf && capacity >= cubin.size()) memcpy(buf, cubin.data(), cubin.size());
  return 0;
}

int gpb_program_num_hp(const gpb_program_t* prog) { return prog ? prog->n_hp : -1; }

void gpb_program_destroy(gpb_program_t* prog) {
  if (!prog) return;
  gpb::jit_release(prog->jit);
  if (prog->code_dev) cudaFree(prog->code_dev);
  delete prog;
}

int gpb_assemble(const gpb_program_t* prog, const double* X, const double* X2, int64_t n, int64_t m,
                 const double* hp, const double* noise, double* K, int64_t ldk, int lower_only, void* stream) {
  if (!prog) return fail_arg(1, "program is null");
  if (!X) return fail_arg(2, "X is null");
  if (n < 0) return fail_arg(4, "n < 0");
  if (m < 0) return fail_arg(5, "m < 0");
  if (!hp && prog->n_hp > 0) return fail_arg(6, "hp is null");
  if (!K) return fail_arg(8, "K is null");
  if (ldk < n) return fail_arg(9, "ldk < n");
  if (lower_only && (X2 || n != m)) return fail_arg(10, "lower_only needs the symmetric case");
  if (n == 0 || m == 0) return 0;
  CU(gpb::run_assemble_rect(prog->code_dev, prog->n_ops, prog->dim, prog->cp_mode, X, X2, n, m, hp, prog->n_hp, noise,
                            K, ldk, lower_only, (cudaStream_t)stream), "gpb_assemble");
  return 0;
}

static int plan_create(int B, const gpb_program_t* const* progs, const int64_t* n, int want_grad, gpb::DistCtx* dist,
                       gpb_plan_t** out, int col_storage = 0) {
  if (B <= 0) return fail_arg(1, "B <= 0");
  if (!progs) return fail_arg(2, "progs is null");
  if (!n) return fail_arg(3, "n is null");
  if (!out) return fail_arg(5, "out is null");
  int rc = ensure_init();
  if (rc) return rc;
  gpb_plan* p = new gpb_plan;
  p->B = B; p->want_grad = want_grad ? 1 : 0; p->ws = nullptr; p->dist = dist;
  p->off_stage[0] = p->off_stage[1] = 0;
  p->col_storage = col_storage; p->off_gdesc = 0; p->n_own_groups = 0;
  p->off_ring[0] = p->off_ring[1] = p->off_ring[2] = 0;
  p->n_max = 0; p->n_hp_max = 0; p->n_ops_max = 0; p->dim = progs[0] ? progs[0]->dim : 1;
  p->mats.resize(B); p->progs.assign(progs, progs + B);
  p->hp_prefix.resize(B + 1); p->grad_prefix.resize(B + 1);
  p->hp_prefix[0] = 0; p->grad_prefix[0] = 0;
  for (int b = 0; b < B; ++b) {
    if (!progs[b]) { delete p; return fail_arg(2, "null program"); }
    if (n[b] < 1 || n[b] > (1 << 20)) { delete p; return fail_arg(3, "n out of range"); }
    if (progs[b]->dim != p->dim) { delete p; return fail_arg(2, "all programs of a plan must share dim"); }
    p->hp_prefix[b + 1] = p->hp_prefix[b] + progs[b]->n_hp;
    p->grad_prefix[b + 1] = p->grad_prefix[b] + progs[b]->n_hp + 1;
    if (n[b] > p->n_max) p->n_max = (int)n[b];
    if (progs[b]->n_hp > p->n_hp_max) p->n_hp_max = progs[b]->n_hp;
    if (progs[b]->n_ops > p->n_ops_max) p->n_ops_max = progs[b]->n_ops;
  }
  size_t off = 0;
  p->off_desc = off; off = al(off + (size_t)B * sizeof(GpbMat));
  p->off_in = off;
  p->off_hp_all = off; off += p->hp_prefix[B] * 8;
  p->off_noise_all = off; off += (size_t)B * 8;
  p->in_bytes = off - p->off_in; off = al(off);
  p->off_out = off;
  p->off_nll_all = off; off += (size_t)B * 8;
  p->off_grad_all = off; off += p->grad_prefix[B] * 8;
  p->off_info_all = off; off += (size_t)B * 4;
  off = (off + 7) & ~(size_t)7;
  p->off_terms_all = off; off += (size_t)B * 16;
  p->out_bytes = off - p->off_out; off = al(off);
  p->holds = HOLDS_NONE; p->have_kinv = 0;
  p->gw.assign((size_t)2 * B, 1.0);
  // the inputs of all GPs form one contiguous region (X_0 | y_0 | X_1 | y_1 | ...)
  p->off_data = off;
  for (int b = 0; b < B; ++b) {
    PlanMat& m = p->mats[b];
    auto put = [&](int which, size_t bytes) { m.off[which] = off; m.bytes[which] = bytes; off = al(off + bytes); };
    put(GPB_BUF_X, (size_t)n[b] * progs[b]->dim * 8);
    put(GPB_BUF_Y, (size_t)n[b] * 8);
  }
  p->data_bytes = off - p->off_data;
  for (int b = 0; b < B; ++b) {
    PlanMat& m = p->mats[b];
    const gpb_program* g = progs[b];
    m.n = n[b];
    m.ld = (int)((n[b] + 1 + 7) / 8 * 8);
    m.nblk = (int)((n[b] + GPB_NB - 1) / GPB_NB);
    m.n_hp = g->n_hp;
    m.n_gtiles = gpb::grad_tiles((int)n[b]);
    auto put = [&](int which, size_t bytes) { m.off[which] = off; m.bytes[which] = bytes; off = al(off + bytes); };
    put(GPB_BUF_ALPHA, (size_t)n[b] * 8);
    put(GPB_BUF_Z, (size_t)n[b] * 8);
    m.off_tmpv = off; off = al(off + (size_t)n[b] * 8);
    m.off_part = off; off = al(off + (size_t)m.nblk * 8);
    m.off_wd = off; off = al(off + (size_t)m.nblk * GPB_NB * GPB_NB * 8);
    m.off_gpart = off; off = al(off + (want_grad ? (size_t)m.n_gtiles * (g->n_hp + 1) * 8 : 0));
    if (col_storage) {
      // only the own block columns (full 128-wide, packed) of the (n + 1)-column workspace
      const int nbr = (int)((n[b] + 1 + GPB_NB - 1) / GPB_NB);
      const int own = gpb_owned_cols_below(nbr, dist->Q, dist->q, dist->OW);
      put(GPB_BUF_A, (size_t)m.ld * ((size_t)own * GPB_NB + GPB_NB) * 8);
    } else {
      put(GPB_BUF_A, (size_t)m.ld * (n[b] + 1) * 8);
    }
    put(GPB_BUF_KINV, want_grad ? (size_t)m.ld * n[b] * 8 : 0);
    m.off[GPB_BUF_HP] = p->off_hp_all + p->hp_prefix[b] * 8; m.bytes[GPB_BUF_HP] = (size_t)g->n_hp * 8;
    m.off[GPB_BUF_NOISE] = p->off_noise_all + (size_t)b * 8; m.bytes[GPB_BUF_NOISE] = 8;
    m.off[GPB_BUF_NLL] = p->off_nll_all + (size_t)b * 8; m.bytes[GPB_BUF_NLL] = 8;
    m.off[GPB_BUF_GRAD] = p->off_grad_all + p->grad_prefix[b] * 8; m.bytes[GPB_BUF_GRAD] = (size_t)(g->n_hp + 1) * 8;
    m.off[GPB_BUF_INFO] = p->off_info_all + (size_t)b * 4; m.bytes[GPB_BUF_INFO] = 4;
    m.off[GPB_BUF_TERMS] = p->off_terms_all + (size_t)b * 16; m.bytes[GPB_BUF_TERMS] = 16;
  }
  if (dist && col_storage) {
    for (int i = 0; i < 3; ++i) { p->off_ring[i] = off; off = al(off + gpb::dist_ring_bytes((int)n[0], dist->OW)); }
    const int nblk0 = (int)((n[0] + GPB_NB - 1) / GPB_NB);
    const int ngroups = (nblk0 + dist->OW - 1) / dist->OW;
    p->n_own_groups = ngroups > dist->q ? (ngroups - dist->q + dist->Q - 1) / dist->Q : 0;
    p->off_gdesc = off; off = al(off + (size_t)(p->n_own_groups + 1) * sizeof(GpbMat));
  } else if (dist) {
    for (int i = 0; i < 2; ++i) { p->off_stage[i] = off; off = al(off + gpb::dist_stage_bytes((int)n[0])); }
  }
  for (int b = 0; b < B; ++b) {
    gpb_plan::Group* g = nullptr;
    for (auto& q : p->groups) if (q.prog == progs[b]) { g = &q; break; }
    if (!g) { p->groups.push_back(gpb_plan::Group{progs[b], {}, 0, 0}); g = &p->groups.back(); }
    g->idx.push_back(b);
    if ((int)n[b] > g->n_max) g->n_max = (int)n[b];
  }
  for (auto& g : p->groups) { g.off_which = off; off = al(off + g.idx.size() * sizeof(int)); }
  p->ws_bytes = off;
  p->h_in = nullptr; p->h_out = nullptr;
  cudaError_t e = cudaMallocHost(&p->h_in, p->in_bytes + 16);
  if (e == cudaSuccess) e = cudaMallocHost(&p->h_out, p->out_bytes + 16);
  p->own_streams = false;
  int prio_lo = 0, prio_hi = 0;
  if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  // four priority levels when the device has them (numerically lower = more urgent; B200: 0 .. -5): critical path >
  // second-next panel's columns > bulk update > overlapped inverse.  With fewer levels neighbours share one.
  const int span = prio_lo - prio_hi;
  const int pr_mid = prio_hi + (span >= 3 ? span / 3 : (span >= 1 ? 1 : 0));
  const int pr_side = span >= 3 ? prio_lo - 1 : prio_lo;
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->ex.crit, cudaStreamNonBlocking, prio_hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->ex.comm, cudaStreamNonBlocking, prio_hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->ex.mid, cudaStreamNonBlocking, pr_mid);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->ex.side, cudaStreamNonBlocking, pr_side);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->ex.inv, cudaStreamNonBlocking, prio_lo);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&p->ex.ev_e[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_g[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_c[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_b[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_d[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_p[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_x[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_join_comm, cudaEventDisableTiming);
  for (int i = 0; i < 9 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&p->ex.ev_ring[i / 3][i % 3], cudaEventDisableTiming);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&p->ex.ev_join[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ex.ev_fork, cudaEventDisableTiming);
  p->n_graphs = 0; p->graphs_off = 0; p->gstream = nullptr;
  {
    const char* ge = getenv("GPB_GRAPH");
    if (ge && ge[0] == '0') p->graphs_off = 1;
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->gstream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->g_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->g_out, cudaEventDisableTiming);
  if (e != cudaSuccess) { delete p; return fail_cuda(e, "gpb_plan_create"); }
  p->own_streams = true;
  *out = p;
  return 0;
}

int gpb_plan_create(int B, const gpb_program_t* const* progs, const int64_t* n, int want_grad, gpb_plan_t** out) {
  return plan_create(B, progs, n, want_grad, nullptr, out);
}

int gpb_plan_create_dist_columns(const gpb_program_t* prog, int64_t n, gpb_dist_t* dist, gpb_plan_t** out) {
  if (!prog) return fail_arg(1, "prog is null");
  if (!dist || !dist->ctx) return fail_arg(3, "dist is null");
  if (dist->ctx->P != 1) return fail_arg(3, "column storage needs a 1 x Q grid");
  const gpb_program_t* progs[1] = {prog};
  return plan_create(1, progs, &n, 0, dist->ctx, out, 1);
}

int gpb_plan_create_dist(const gpb_program_t* prog, int64_t n, int want_grad, gpb_dist_t* dist, gpb_plan_t** out) {
  if (!prog) return fail_arg(1, "program is null");
  if (!dist || !dist->ctx) return fail_arg(4, "dist is null");
  const gpb_program_t* progs[1] = {prog};
  return plan_create(1, progs, &n, want_grad, dist->ctx, out);
}

size_t gpb_plan_workspace_bytes(const gpb_plan_t* plan) { return plan ? plan->ws_bytes : 0; }

int gpb_plan_bind(gpb_plan_t* p, void* workspace) {
  if (!p) return fail_arg(1, "plan is null");
  if (!workspace || ((uintptr_t)workspace & 255)) return fail_arg(2, "workspace null or not 256-byte aligned");
  // graphs captured for a previous workspace hold its addresses
  for (int i = 0; i < p->n_graphs; ++i) cudaGraphExecDestroy(p->graph_exec[i]);
  p->n_graphs = 0;
  p->ws = (char*)workspace;
  std::vector<GpbMat> h(p->B);
  for (int b = 0; b < p->B; ++b) {
    const PlanMat& m = p->mats[b];
    const gpb_program* g = p->progs[b];
    GpbMat& d = h[b];
    memset(&d, 0, sizeof(d));
    char* w = p->ws;
    d.A = (double*)(w + m.off[GPB_BUF_A]);
    d.Kinv = p->want_grad ? (double*)(w + m.off[GPB_BUF_KINV]) : nullptr;
    d.Wd = (double*)(w + m.off_wd);
    d.part = (double*)(w + m.off_part);
    d.gpart = (double*)(w + m.off_gpart);
    d.alpha = (double*)(w + m.off[GPB_BUF_ALPHA]);
    d.zvec = (double*)(w + m.off[GPB_BUF_Z]);
    d.tmpv = (double*)(w + m.off_tmpv);
    d.X = (const double*)(w + m.off[GPB_BUF_X]);
    d.y = (const double*)(w + m.off[GPB_BUF_Y]);
    d.hp = (const double*)(w + m.off[GPB_BUF_HP]);
    d.noise = (const double*)(w + m.off[GPB_BUF_NOISE]);
    d.code = g->code_dev;
    d.nll = (double*)(w + m.off[GPB_BUF_NLL]);
    d.grad = (double*)(w + m.off[GPB_BUF_GRAD]);
    d.info = (int*)(w + m.off[GPB_BUF_INFO]);
    d.terms = (double*)(w + m.off[GPB_BUF_TERMS]);
    d.gw_quad = p->gw[2 * b]; d.gw_logdet = p->gw[2 * b + 1];
    d.n = (int)m.n; d.ld = m.ld; d.dim = g->dim; d.n_ops = g->n_ops; d.n_hp = g->n_hp; d.aug = 1;
    d.cp_mode = g->cp_mode; d.n_gtiles = m.n_gtiles;
    if (p->dist) {
      d.own_P = p->dist->P; d.own_Q = p->dist->Q; d.own_p = p->dist->p; d.own_q = p->dist->q; d.own_W = p->dist->OW;
      d.col_world = p->dist->world; d.col_rank = p->dist->rank;
      d.own_compact = p->col_storage;
    }
  }
  if (p->col_storage) {
    // one descriptor per own group of block columns: A shifted so that d.A + k * 128 * ld (what the diagonal-block kernel
    // and the panel product compute for block column k) lands on the group's packed columns
    std::vector<GpbMat> gd(p->n_own_groups + 1, h[0]);
    const int OW = p->dist->OW, Q = p->dist->Q, q = p->dist->q;
    for (int g = 0; g < p->n_own_groups; ++g) {
      const long long k0 = (long long)(g * Q + q) * OW, t0 = (long long)g * OW;
      gd[g].A = h[0].A + (t0 - k0) * (long long)GPB_NB * h[0].ld;
      gd[g].own_compact = 0;
    }
    cudaError_t ge = cudaMemcpy(p->ws + p->off_gdesc, gd.data(), gd.size() * sizeof(GpbMat), cudaMemcpyHostToDevice);
    if (ge != cudaSuccess) return fail_cuda(ge, "bind (group descriptors)");
  }
  p->h_desc0 = h[0];
  p->h_desc = h;
  p->holds = HOLDS_NONE; p->have_kinv = 0;
  CU(cudaMemcpy(p->ws + p->off_desc, h.data(), (size_t)p->B * sizeof(GpbMat), cudaMemcpyHostToDevice), "gpb_plan_bind");
  CU(cudaMemset(p->ws + p->off_out, 0, p->out_bytes), "gpb_plan_bind");
  for (const auto& g : p->groups)
    CU(cudaMemcpy(p->ws + g.off_which, g.idx.data(), g.idx.size() * sizeof(int), cudaMemcpyHostToDevice), "gpb_plan_bind");
  return 0;
}

int gpb_plan_buffer(const gpb_plan_t* p, int b, int which, void** ptr, size_t* bytes, int64_t* ld) {
  if (!p) return fail_arg(1, "plan is null");
  if (b < 0 || b >= p->B) return fail_arg(2, "b out of range");
  if (which < 0 || which >= NBUF) return fail_arg(3, "unknown buffer");
  if (!p->ws) return fail_arg(1, "plan is not bound");
  const PlanMat& m = p->mats[b];
  if (ptr) *ptr = p->ws + m.off[which];
  if (bytes) *bytes = m.bytes[which];
  if (ld) *ld = m.ld;
  return 0;
}

// assembly / gradient of all GPs of the plan: one launch per kernel program (its specialised kernels, else the interpreter)
static cudaError_t plan_assemble(gpb_plan* p, const GpbMat* dm, cudaStream_t s) {
  for (const auto& g : p->groups) {
    const int* which = (const int*)(p->ws + g.off_which);
    const int T = (g.n_max + 63) / 64;
    cudaError_t e;
    if (g.prog->jit.assemble) e = gpb::jit_launch(g.prog->jit.assemble, (unsigned)(T * (T + 1) / 2), (unsigned)g.idx.size(), dm, which, s);
    else e = gpb::run_assemble_batched(dm, which, (int)g.idx.size(), g.n_max, s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
static cudaError_t plan_grad(gpb_plan* p, const GpbMat* dm, cudaStream_t s) {
  for (const auto& g : p->groups) {
    const int* which = (const int*)(p->ws + g.off_which);
    cudaError_t e;
    if (g.prog->jit.grad) e = gpb::jit_launch(g.prog->jit.grad, (unsigned)gpb::grad_tiles(g.n_max), (unsigned)g.idx.size(), dm, which, s);
    else e = gpb::run_grad_tiles(dm, which, (int)g.idx.size(), g.n_max, g.prog->n_hp, g.prog->n_ops, p->dim, s);
    if (e != cudaSuccess) return e;
  }
  return gpb::run_grad_reduce(dm, p->B, p->n_hp_max, s);
}

int gpb_plan_eval(gpb_plan_t* p, int stages, void* stream) {
  if (!p) return fail_arg(1, "plan is null");
  if (!p->ws) return fail_arg(1, "plan is not bound");
  if ((stages & (GPB_STAGE_INVERSE | GPB_STAGE_TRTRI | GPB_STAGE_LAUUM | GPB_STAGE_GRAD)) && !p->want_grad)
    return fail_arg(2, "plan was created without gradient workspace");
  if (p->col_storage && (stages & ~(GPB_STAGE_ASSEMBLE | GPB_STAGE_POTRF | GPB_STAGE_NLL)))
    return fail_arg(2, "a column-storage plan evaluates the likelihood only (ASSEMBLE | POTRF | NLL): the factor is not replicated");
  { const int rc_state = advance_state(p, stages); if (rc_state) return rc_state; }
  cudaStream_t s = (cudaStream_t)stream;
  const GpbMat* dm = (const GpbMat*)(p->ws + p->off_desc);
  p->ex.main = s;
  if (p->dist) {

    if (stages & GPB_STAGE_ASSEMBLE) {
      NvtxRange r("gpb:assemble");
      CU(cudaMemsetAsync(p->ws + p->off_info_all, 0, 4, s), "reset info");
      CU(plan_assemble(p, dm, s), "assemble");
    }
    if (p->col_storage) {
      if (stages & GPB_STAGE_POTRF) {
        NvtxRange r("gpb:potrf(dist, column storage)");
        double* ring[3] = {(double*)(p->ws + p->off_ring[0]), (double*)(p->ws + p->off_ring[1]), (double*)(p->ws + p->off_ring[2])};
        cudaError_t e = gpb::run_potrf_dist_store((const GpbMat*)(p->ws + p->off_gdesc), p->h_desc0, *p->dist, ring, p->ex);
        if (e != cudaSuccess) {
          if (e == cudaErrorUnknown && gpb::dist_last_error()[0]) { g_err = gpb::dist_last_error(); return 2000; }
          return fail_cuda(e, "potrf_dist_store");
        }
      }
      if (stages & GPB_STAGE_NLL) {
        cudaError_t e = gpb::run_finalize_dist_store(dm, *p->dist, p->h_desc0.tmpv, std::log(M_PI * 2.0), s);
        if (e != cudaSuccess) {
          if (e == cudaErrorUnknown && gpb::dist_last_error()[0]) { g_err = gpb::dist_last_error(); return 2000; }
          return fail_cuda(e, "finalize_dist_store");
        }
      }
      return 0;
    }
    if (stages & GPB_STAGE_POTRF) {
      NvtxRange r("gpb:potrf(dist)");
      double* stage[2] = {(double*)(p->ws + p->off_stage[0]), (double*)(p->ws + p->off_stage[1])};
      cudaError_t e = gpb::run_potrf_dist(dm, p->h_desc0, *p->dist, stage, p->ex);
      if (e != cudaSuccess) {
        if (e == cudaErrorUnknown && gpb::dist_last_error()[0]) { g_err = gpb::dist_last_error(); return 2000; }
        return fail_cuda(e, "potrf_dist");
      }
    }
    if (stages & GPB_STAGE_NLL) CU(gpb::run_finalize_dist(dm, std::log(M_PI * 2.0), s), "finalize_dist");
    // L and the inverted diagonal blocks are complete on every rank: the back substitution runs replicated
    if (stages & GPB_STAGE_BACKSOLVE) CU(gpb::run_trsv(dm, 1, p->n_max, 1, s), "backsolve");
    auto dist_rc = [&](cudaError_t e, const char* where) -> int {
      if (e == cudaSuccess) return 0;
      if (e == cudaErrorUnknown && gpb::dist_last_error()[0]) { g_err = gpb::dist_last_error(); return 2000; }
      return fail_cuda(e, where);
    };
    const bool want_trtri = (stages & (GPB_STAGE_INVERSE | GPB_STAGE_TRTRI)) != 0;
    const bool want_lauum = (stages & (GPB_STAGE_INVERSE | GPB_STAGE_LAUUM)) != 0;
    static int ovl_env = -1;     // GPB_DIST_OVERLAP=0: exchange of W after the inverse, W^T W as one launch
    if (ovl_env < 0) { const char* oe = getenv("GPB_DIST_OVERLAP"); ovl_env = (oe && oe[0] == '0') ? 0 : 1; }
    const bool overlap = want_trtri && want_lauum && ovl_env && s == p->ex.main;
    if (want_trtri) {
      NvtxRange r("gpb:trtri(dist)");
      int rc = dist_rc(gpb::run_trtri_dist(dm, p->h_desc0, *p->dist, (double*)(p->ws + p->off_stage[0]), p->ex, !overlap), "trtri_dist");
      if (rc) return rc;
      if (!overlap) CU(gpb::run_alpha(dm, 1, p->n_max, s), "alpha");
    }
    if (want_lauum) {
      NvtxRange r("gpb:lauum(dist)");
      int rc = overlap ? dist_rc(gpb::run_exchange_lauum_dist(dm, p->h_desc0, *p->dist, p->ex), "exchange+lauum_dist")
                       : dist_rc(gpb::run_lauum_dist(dm, p->h_desc0, *p->dist, s), "lauum_dist");
      if (rc) return rc;
      if (overlap) CU(gpb::run_alpha(dm, 1, p->n_max, s), "alpha");
    }
    if (stages & GPB_STAGE_GRAD) {
      NvtxRange r("gpb:grad(dist)");
      CU(plan_grad(p, dm, s), "grad");
      int rc = dist_rc(gpb::run_grad_allreduce(p->h_desc0.grad, p->mats[0].n_hp + 1, *p->dist, s), "grad allreduce");
      if (rc) return rc;
    }
    return 0;
  }
  if (stages & GPB_STAGE_ASSEMBLE) {
    NvtxRange r("gpb:assemble");
    CU(cudaMemsetAsync(p->ws + p->off_info_all, 0, (size_t)p->B * 4, s), "reset info");
    CU(plan_assemble(p, dm, s), "assemble");
  }
  // one large matrix, factorisation and inverse in the same call: the inverse of the block columns that are already
  // final runs under the factorisation's latency-bound tail (GPB_FUSE_TRTRI=0 keeps the stages apart)
  bool fused_trtri = false;
  if (stages & GPB_STAGE_POTRF) {
    const bool lookahead = (p->B == 1 && p->n_max >= 1024);
    static int fuse_env = -1;
    if (fuse_env < 0) { const char* fe = getenv("GPB_FUSE_TRTRI"); fuse_env = (fe && fe[0] == '0') ? 0 : 1; }
    fused_trtri = lookahead && fuse_env && (stages & (GPB_STAGE_INVERSE | GPB_STAGE_TRTRI)) != 0;
    NvtxRange r(fused_trtri ? "gpb:potrf+trtri" : "gpb:potrf");
    CU(gpb::run_potrf(dm, p->B, p->n_max, 1, lookahead, fused_trtri, p->ex), "potrf");
  }
  if (stages & GPB_STAGE_NLL) {
    NvtxRange r("gpb:nll");
    const double log2pi = std::log(M_PI * 2.0);
    CU(gpb::run_finalize(dm, p->B, log2pi, s), "finalize");
  }
  if (stages & GPB_STAGE_BACKSOLVE) {
    NvtxRange r("gpb:backsolve");
    CU(gpb::run_trsv(dm, p->B, p->n_max, 1, s), "backsolve");
  }
  if (stages & (GPB_STAGE_INVERSE | GPB_STAGE_TRTRI)) {
    NvtxRange r("gpb:trtri");
    if (!fused_trtri) CU(gpb::run_trtri(dm, p->B, p->n_max, s), "trtri");
    CU(gpb::run_alpha(dm, p->B, p->n_max, s), "alpha");
  }
  if (stages & (GPB_STAGE_INVERSE | GPB_STAGE_LAUUM)) {
    NvtxRange r("gpb:lauum");
    CU(gpb::run_lauum(dm, p->B, p->n_max, s), "lauum");
  }
  if (stages & GPB_STAGE_GRAD) {
    NvtxRange r("gpb:grad");
    CU(plan_grad(p, dm, s), "grad");
  }
  return 0;
}

int gpb_plan_eval_host(gpb_plan_t* p, int stages, const double* const* X_host, const double* const* y_host,
                       const double* const* hp_host, const double* noise_host, double* nll_host, double* grad_host,
                       int* info_host, void* stream) {
  if (!p) return fail_arg(1, "plan is null");
  if (!p->ws) return fail_arg(1, "plan is not bound");
  if (!hp_host && p->hp_prefix[p->B] > 0) return fail_arg(5, "hp_host is null");
  if (!noise_host) return fail_arg(6, "noise_host is null");
  cudaStream_t s = (cudaStream_t)stream;
  NvtxRange r_host("gpb:eval_host");
  // Fewer, larger copies: when the caller's X / y buffers are laid out like the plan's data region (one staging buffer,
  // gpb_plan_input_layout) all inputs travel in ONE cudaMemcpyAsync instead of 2 B of them (1024 blocks: 2048 copies
  // cost a sixth of the evaluation).
  bool packed = X_host && y_host && p->B > 1;
  if (packed) {
    const char* base = (const char*)X_host[0] - (p->mats[0].off[GPB_BUF_X] - p->off_data);
    for (int b = 0; b < p->B && packed; ++b) {
      const PlanMat& m = p->mats[b];
      packed = X_host[b] && y_host[b] && (const char*)X_host[b] == base + (m.off[GPB_BUF_X] - p->off_data) &&
               (const char*)y_host[b] == base + (m.off[GPB_BUF_Y] - p->off_data);
    }
    if (packed) CU(cudaMemcpyAsync(p->ws + p->off_data, base, p->data_bytes, cudaMemcpyHostToDevice, s), "H2D X, y (packed)");
  }
  for (int b = 0; b < p->B; ++b) {
    const PlanMat& m = p->mats[b];
    if (!packed && X_host && X_host[b])
      CU(cudaMemcpyAsync(p->ws + m.off[GPB_BUF_X], X_host[b], m.bytes[GPB_BUF_X], cudaMemcpyHostToDevice, s), "H2D X");
    if (!packed && y_host && y_host[b])
      CU(cudaMemcpyAsync(p->ws + m.off[GPB_BUF_Y], y_host[b], m.bytes[GPB_BUF_Y], cudaMemcpyHostToDevice, s), "H2D y");
    if (m.n_hp > 0) memcpy(p->h_in + (p->off_hp_all - p->off_in) + p->hp_prefix[b] * 8, hp_host[b], (size_t)m.n_hp * 8);
  }
  memcpy(p->h_in + (p->off_noise_all - p->off_in), noise_host, (size_t)p->B * 8);
  CU(cudaMemcpyAsync(p->ws + p->off_in, p->h_in, p->in_bytes, cudaMemcpyHostToDevice, s), "H2D hp");
  int rc = 0;
  // The launch sequence (hundreds of kernels on three streams for a large matrix, dozens for a small one) is captured
  // once per stage mask and replayed: the inputs live at fixed addresses inside the workspace.  Distributed plans launch
  // directly (their NCCL calls are ordered with the other ranks by the host loop).
  cudaGraphExec_t exec = nullptr;
  bool exec_cached = false;
  if (!p->dist && !p->graphs_off) {
    for (int i = 0; i < p->n_graphs; ++i)
      if (p->graph_stages[i] == stages) { exec = p->graph_exec[i]; exec_cached = true; }
    if (!exec && p->n_graphs < 4) {
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamBeginCapture(p->gstream, cudaStreamCaptureModeThreadLocal);
      if (ce == cudaSuccess) {
        rc = gpb_plan_eval(p, stages, (void*)p->gstream);
        ce = cudaStreamEndCapture(p->gstream, &graph);
        if (rc == 0 && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
      }
      if (rc != 0 || ce != cudaSuccess || !exec) {
        (void)cudaGetLastError();
        exec = nullptr; rc = 0; p->graphs_off = 1;           // capture is best effort: fall back to direct launches
      } else {
        p->graph_stages[p->n_graphs] = stages; p->graph_exec[p->n_graphs] = exec; ++p->n_graphs;
      }
    }
  }
  if (exec) {
    if (exec_cached) {   // a replay repeats the transitions the capture call made
      const int rc_state = advance_state(p, stages);
      if (rc_state) return rc_state;
    }
    CU(cudaEventRecord(p->g_in, s), "graph fork");
    CU(cudaStreamWaitEvent(p->gstream, p->g_in, 0), "graph fork");
    CU(cudaGraphLaunch(exec, p->gstream), "graph launch");
    CU(cudaEventRecord(p->g_out, p->gstream), "graph join");
    CU(cudaStreamWaitEvent(s, p->g_out, 0), "graph join");
  } else {
    rc = gpb_plan_eval(p, stages, stream);
    if (rc) return rc;
  }
  CU(cudaMemcpyAsync(p->h_out, p->ws + p->off_out, p->out_bytes, cudaMemcpyDeviceToHost, s), "D2H results");
  CU(cudaStreamSynchronize(s), "sync");
  if (nll_host) memcpy(nll_host, p->h_out + (p->off_nll_all - p->off_out), (size_t)p->B * 8);
  if (grad_host) memcpy(grad_host, p->h_out + (p->off_grad_all - p->off_out), p->grad_prefix[p->B] * 8);
  if (info_host) memcpy(info_host, p->h_out + (p->off_info_all - p->off_out), (size_t)p->B * 4);
  return 0;
}

int gpb_plan_set_grad_weights(gpb_plan_t* p, int b, double w_quad, double w_logdet) {
  if (!p) return fail_arg(1, "plan is null");
  if (b < 0 || b >= p->B) return fail_arg(2, "b out of range");
  p->gw[2 * b] = w_quad; p->gw[2 * b + 1] = w_logdet;
  if (p->ws) {
    p->h_desc[b].gw_quad = w_quad; p->h_desc[b].gw_logdet = w_logdet;
    if (b == 0) p->h_desc0 = p->h_desc[0];
    CU(cudaMemcpy(p->ws + p->off_desc + (size_t)b * sizeof(GpbMat), &p->h_desc[b], sizeof(GpbMat), cudaMemcpyHostToDevice),
       "gpb_plan_set_grad_weights");
  }
  return 0;
}

int gpb_plan_last_terms(const gpb_plan_t* p, double* quad_host, double* logdet_host) {
  if (!p) return fail_arg(1, "plan is null");
  if (!p->h_out) return fail_arg(1, "plan has no host mirror");
  const double* t = (const double*)(p->h_out + (p->off_terms_all - p->off_out));
  for (int b = 0; b < p->B; ++b) {
    if (quad_host) quad_host[b] = t[2 * b];
    if (logdet_host) logdet_host[b] = t[2 * b + 1];
  }
  return 0;
}

int gpb_plan_input_layout(const gpb_plan_t* p, int b, size_t* x_offset, size_t* y_offset, size_t* total_bytes) {
  if (!p) return fail_arg(1, "plan is null");
  if (b < 0 || b >= p->B) return fail_arg(2, "b out of range");
  if (x_offset) *x_offset = p->mats[b].off[GPB_BUF_X] - p->off_data;
  if (y_offset) *y_offset = p->mats[b].off[GPB_BUF_Y] - p->off_data;
  if (total_bytes) *total_bytes = p->data_bytes;
  return 0;
}

void gpb_plan_destroy(gpb_plan_t* p) {
  if (!p) return;
  if (p->own_streams) {
    cudaStreamDestroy(p->ex.crit);
    cudaStreamDestroy(p->ex.comm);
    cudaStreamDestroy(p->ex.mid);
    cudaStreamDestroy(p->ex.side);
    cudaStreamDestroy(p->ex.inv);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(p->ex.ev_e[i]); cudaEventDestroy(p->ex.ev_g[i]); cudaEventDestroy(p->ex.ev_c[i]);
      cudaEventDestroy(p->ex.ev_b[i]); cudaEventDestroy(p->ex.ev_d[i]);
      cudaEventDestroy(p->ex.ev_p[i]); cudaEventDestroy(p->ex.ev_x[i]);
    }
    cudaEventDestroy(p->ex.ev_join_comm);
    for (int i = 0; i < 9; ++i) cudaEventDestroy(p->ex.ev_ring[i / 3][i % 3]);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(p->ex.ev_join[i]);
    cudaEventDestroy(p->ex.ev_fork);
    for (int i = 0; i < p->n_graphs; ++i) cudaGraphExecDestroy(p->graph_exec[i]);
    if (p->gstream) { cudaStreamDestroy(p->gstream); cudaEventDestroy(p->g_in); cudaEventDestroy(p->g_out); }
  }
  if (p->h_in) cudaFreeHost(p->h_in);
  if (p->h_out) cudaFreeHost(p->h_out);
  delete p;
}

/* ---- process grid of the distributed factorisation ---------------------------------------------------------- */
int gpb_dist_unique_id(unsigned char* id128) {
  if (!id128) return fail_arg(1, "id is null");
  if (gpb::dist_unique_id(id128)) { g_err = gpb::dist_last_error(); return 2000; }
  return 0;
}

int gpb_dist_init(const unsigned char* id128, int rank, int world, int P, int Q, gpb_dist_t** out) {
  if (!id128) return fail_arg(1, "id is null");
  if (world < 1 || rank < 0 || rank >= world) return fail_arg(2, "rank / world out of range");
  if (P < 1 || Q < 1 || P * Q != world || P > GPB_DIST_MAX_P) return fail_arg(4, "P x Q must equal world, P <= 8");
  if (!out) return fail_arg(6, "out is null");
  int rc = ensure_init();
  if (rc) return rc;
  gpb::DistCtx* ctx = nullptr;
  if (gpb::dist_create(id128, rank, world, P, Q, &ctx)) { g_err = gpb::dist_last_error(); return 2000; }
  gpb_dist* d = new gpb_dist;
  d->ctx = ctx;
  *out = d;
  return 0;
}

int gpb_dist_loopback_create(int world, int P, int Q, gpb_dist_t** out) {
  if (world < 1 || world > 64) return fail_arg(1, "world out of range");
  if (P < 1 || Q < 1 || P * Q != world || P > GPB_DIST_MAX_P) return fail_arg(2, "P x Q must equal world, P <= 8");
  if (!out) return fail_arg(4, "out is null");
  int rc = ensure_init();
  if (rc) return rc;
  std::vector<gpb::DistCtx*> ctx(world, nullptr);
  if (gpb::dist_create_loopback(world, P, Q, ctx.data())) { g_err = gpb::dist_last_error(); return 2000; }
  for (int r = 0; r < world; ++r) {
    out[r] = new gpb_dist;
    out[r]->ctx = ctx[r];
  }
  return 0;
}

void gpb_dist_destroy(gpb_dist_t* d) {
  if (!d) return;
  gpb::dist_destroy(d->ctx);
  delete d;
}

int gpb_dist_owner(int I, int J, int P, int Q) {
  if (I < 0 || J < 0 || P < 1 || Q < 1) return -1;
  return (I % P) * Q + (J % Q);
}

int gpb_dist_col_width(int P) { return P < 1 ? -1 : gpb::dist_col_width(P); }

int gpb_dist_owner_w(int I, int J, int P, int Q, int W) {
  if (I < 0 || J < 0 || P < 1 || Q < 1 || W < 1) return -1;
  return (I % P) * Q + ((J / W) % Q);
}

int gpb_dist_owned_cols(int J_lo, int J_hi, int Q, int q, int W, int* cols, int cap) {
  if (J_lo < 0 || Q < 1 || q < 0 || q >= Q || W < 1 || (cap > 0 && !cols)) return -1;
  return gpb::dist_owned_cols(J_lo, J_hi, Q, q, W, cols, cap);
}

int gpb_dist_panel_segments(int k, int n_tiles, int P, int* seg_base, int* seg_count, int* seg_first) {
  if (k < 0 || n_tiles < 0) return fail_arg(1, "k / n_tiles negative");
  if (P < 1 || P > GPB_DIST_MAX_P) return fail_arg(3, "P out of range");
  if (!seg_base || !seg_count || !seg_first) return fail_arg(4, "null output");
  gpb::dist_panel_segments(k, n_tiles, P, seg_base, seg_count, seg_first);
  return 0;
}

int gpb_trtri_schedule(int n, const long long* final_cols, int n_steps, int* tasks, int capacity) {
  if (n < 1) return fail_arg(1, "n < 1");
  if (!final_cols || n_steps < 1) return fail_arg(2, "no progress sequence");
  if (!tasks && capacity > 0) return fail_arg(4, "tasks is null");
  return gpb::trtri_schedule_host(n, final_cols, n_steps, tasks, capacity);
}

int gpb_gemm(int a_kmajor, int b_kmajor, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
             int M, int N, int K, double alpha, double beta, void* stream) {
  if (!A) return fail_arg(3, "A is null");
  if (!B) return fail_arg(5, "B is null");
  if (!C) return fail_arg(7, "C is null");
  if ((lda & 1) || (ldb & 1)) return fail_arg(4, "leading dimensions of the operands must be even (16-byte cp.async)");
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return fail_arg(3, "operands must be 16-byte aligned");
  if (M <= 0 || N <= 0 || K < 0) return fail_arg(9, "bad shape");
  int rc = ensure_init();
  if (rc) return rc;
  CU(gpb::run_gemm_plain(a_kmajor, b_kmajor, A, lda, B, ldb, C, ldc, M, N, K, alpha, beta, (cudaStream_t)stream), "gpb_gemm");
  return 0;
}

int gpb_debug_diag_clocks(long long* out8) { return gpb::debug_diag_clocks(out8); }

int gpb_microbench(int kind, int iters, int blocks, void* stream) {
  CU(gpb::run_microbench(kind, iters, blocks, (cudaStream_t)stream), "gpb_microbench");
  return 0;
}

int gpb_zero_upper(double* A, int n, int ld, void* stream) {
  if (!A) return fail_arg(1, "A is null");
  CU(gpb::run_zero_upper(A, n, ld, (cudaStream_t)stream), "gpb_zero_upper");
  return 0;
}
int gpb_symmetrize(double* A, int n, int ld, void* stream) {
  if (!A) return fail_arg(1, "A is null");
  CU(gpb::run_symmetrize(A, n, ld, (cudaStream_t)stream), "gpb_symmetrize");
  return 0;
}

}  // extern "C"
