/* libgpb - C-ABI of the B200-native exact-GP likelihood path (drop-in arithmetic for gpbasics 2.0.0).
 *
 * The reference (Bernsai/GaussianProcessFundamentals) has no FFI: its hot path is Python calling TensorFlow.  Every
 * entry point below therefore cites the reference *Python* interface whose arithmetic it replaces
 * (paths relative to main/gpbasics/).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C types only; all matrices FP64; device pointers unless the name ends in _host
 *   - matrices are column-major with a leading dimension (a row-major [n,m] tensor is the same memory as the
 *     column-major m x n transpose; symmetric / triangular results are exposed through transposed views)
 *   - `stream` is a cudaStream_t passed as void*; compute calls are asynchronous on it and never allocate
 *   - return 0 on success, -k if argument k (1-based) is invalid, 1000 + cudaError_t on a CUDA failure;
 *     gpb_last_error() returns a description.  Non positive-definite matrices are reported through `info`
 *     (LAPACK convention: 1-based index of the first non-positive pivot), not through the return code.
 */
#ifndef GPB_H_
#define GPB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpb_program gpb_program_t;
typedef struct gpb_plan gpb_plan_t;
typedef struct gpb_dist gpb_dist_t;

/* op codes of the postfix kernel program, 4 int32 words per op {op, a, b, c}; see csrc/program.cuh */
#define GPB_SE 1
#define GPB_PER 2
#define GPB_LIN 3
#define GPB_MAT32 4
#define GPB_MAT52 5
#define GPB_WN 6
#define GPB_SE_ARD 7
#define GPB_DIST_L2 8
#define GPB_DIST_L1 9
#define GPB_ADD2 16
#define GPB_MUL2 17
#define GPB_CPW 18

/* gpb_plan_eval stage bits */
#define GPB_STAGE_ASSEMBLE 1   /* K + s2 I (+ y as carried row)          */
#define GPB_STAGE_POTRF 2      /* L, z = L^-1 y, log det                  */
#define GPB_STAGE_NLL 4        /* negative log marginal likelihood        */
#define GPB_STAGE_INVERSE 8    /* W = L^-1, alpha = W^T z, Kinv = W^T W  */
#define GPB_STAGE_GRAD 16      /* d nll / d theta, d nll / d s2           */
#define GPB_STAGE_BACKSOLVE 32 /* alpha by substitution, L left intact    */
#define GPB_STAGE_TRTRI 64     /* first half of INVERSE: W = L^-1, alpha  */
#define GPB_STAGE_LAUUM 128    /* second half of INVERSE: Kinv = W^T W (one launch; bench.py's roofline kernel) */
#define GPB_STAGES_LML 7
#define GPB_STAGES_LML_GRAD 31

/* buffers of a plan that can be mapped by the host language (gpb_plan_buffer) */
#define GPB_BUF_A 0      /* (n+1) columns x ld: K+s2I -> L -> L^-1, lower; row n carries y^T -> z^T */
#define GPB_BUF_KINV 1   /* n columns x ld: K^-1, lower                                              */
#define GPB_BUF_ALPHA 2  /* [n]                                                                      */
#define GPB_BUF_Z 3      /* [n]                                                                      */
#define GPB_BUF_X 4      /* [n x dim] row-major                                                      */
#define GPB_BUF_Y 5      /* [n]                                                                      */
#define GPB_BUF_HP 6     /* [n_hp]                                                                   */
#define GPB_BUF_NOISE 7  /* [1]                                                                      */
#define GPB_BUF_NLL 8    /* [1]                                                                      */
#define GPB_BUF_GRAD 9   /* [n_hp + 1], last = d nll / d s2                                          */
#define GPB_BUF_INFO 10  /* [1] int32                                                                */
#define GPB_BUF_TERMS 11 /* [2]: y^T K^-1 y and sum(log diag L) (written by GPB_STAGE_NLL)                 */

int gpb_version(void);
const char* gpb_last_error(void);
/* Number of CUDA kernels launched by this library since load (bench.py's gpu_launches). */
long long gpb_launch_count(void);

/* ---- kernel programs --------------------------------------------------------------------------------------
 * A compiled kernel tree.  Replaces the recursive Kernel.get_tf_tensor walk
 * (KernelBasics/Kernel.py:51, BaseKernels.py:114-134,277-294,440-457, Operators.py:207-225,306-326,442-476).
 * cp_mode: 0 sigmoid, 1 indicator, 2 approx-indicator (global_parameters.py:10-13,44).                       */
int gpb_program_create(const int32_t* code, int n_ops, int dim, int cp_mode, gpb_program_t** out);
int gpb_program_num_hp(const gpb_program_t* prog);
/* Run-time specialisation.  gpb_program_create also turns the program into straight-line CUDA (all stack / tape /
 * adjoint slots of the interpreter become registers), compiles it with NVRTC for the device in use and uses those
 * kernels for the assembly and trace-gradient stages of every plan the program is part of; GPs of a plan that share a
 * program are one launch.  Without NVRTC (or with GPB_JIT=0) the interpreter kernels run instead - same results to
 * rounding - and programs whose gradient tape exceeds the interpreter's 64 entries are refused.
 *   gpb_program_is_specialised : 1 when the program runs on its own kernels
 *   gpb_jit_source / gpb_jit_cubin : host-only (no GPU): the generated source / its CUBIN for `arch` ("sm_100a"), for
 *                                    inspection and tests; `needed` receives the size, buf may be NULL            */
int gpb_program_is_specialised(const gpb_program_t* prog);
const char* gpb_program_jit_note(const gpb_program_t* prog);
int gpb_jit_available(void);
int gpb_jit_set_nvrtc_path(const char* path);
int gpb_jit_source(const int32_t* code, int n_ops, int dim, int cp_mode, char* buf, size_t capacity, size_t* needed);
int gpb_jit_cubin(const int32_t* code, int n_ops, int dim, int cp_mode, const char* arch, void* buf, size_t capacity,
                  size_t* needed);
void gpb_program_destroy(gpb_program_t* prog);

/* ---- covariance assembly ----------------------------------------------------------------------------------
 * K[i + j*ldk] = k(X[i,:], X2[j,:]) (+ noise on the diagonal when X2 == NULL).  X: [n x dim] row-major, X2: [m x dim]
 * row-major or NULL for the symmetric case; hp, noise: device.  lower_only != 0 writes i >= j only.
 * Replaces HolisticCovarianceMatrix.get_K / get_K_noised / get_K_s / get_K_ss
 * (Statistics/CovarianceMatrix.py:187-206,213-221,277-286).                                                    */
int gpb_assemble(const gpb_program_t* prog, const double* X, const double* X2, int64_t n, int64_t m,
                 const double* hp, const double* noise, double* K, int64_t ldk, int lower_only, void* stream);

/* ---- plans: the fused LML (+ gradient) evaluation over a batch of independent GPs --------------------------
 * One plan = B GPs (B = 1: GaussianProcess; B > 1: the blocks of a Blockwise/PartitionedGaussianProcess, or a
 * batch of candidate kernels).  Replaces LogLikelihood.get_metric (Metrics/LogLikelihood.py:30-65) with
 * Metric.get_alpha_cholesky / get_log_determinant_cholesky (Metrics/Metrics.py:138-139,152-154),
 * HolisticCovarianceMatrix.get_L_K / get_L_alpha (Statistics/CovarianceMatrix.py:247-265),
 * SegmentedCovarianceMatrix.get_*_blocks (:316-339,:469-506), BlockwiseLogLikelihood.get_metric
 * (Metrics/LogLikelihood.py:77-104, per-block values; the caller sums) and the GradientTape gradient of
 * Optimizer/Fitter.py:124-132,154-158.                                                                         */
int gpb_plan_create(int B, const gpb_program_t* const* progs, const int64_t* n, int want_grad, gpb_plan_t** out);
size_t gpb_plan_workspace_bytes(const gpb_plan_t* plan);
/* workspace: device memory of gpb_plan_workspace_bytes() bytes, 256-byte aligned, owned by the caller. */
int gpb_plan_bind(gpb_plan_t* plan, void* workspace);
/* device address / byte size of one of the plan's buffers for GP b (all live inside the workspace) */
int gpb_plan_buffer(const gpb_plan_t* plan, int b, int which, void** ptr, size_t* bytes, int64_t* ld);
/* asynchronous, graph-capturable: inputs are read from the GPB_BUF_X/Y/HP/NOISE buffers */
int gpb_plan_eval(gpb_plan_t* plan, int stages, void* stream);
/* host-buffer call (the end-to-end path): copies X, y, hp, noise of every GP to the device, evaluates, copies
 * nll[B], grad (concatenated, n_hp_b + 1 each) and info[B] back and synchronises the stream.
 * X_host / y_host may be NULL to keep the inputs of the previous call resident.                               */
int gpb_plan_eval_host(gpb_plan_t* plan, int stages, const double* const* X_host, const double* const* y_host,
                       const double* const* hp_host, const double* noise_host, double* nll_host, double* grad_host,
                       int* info_host, void* stream);
/* Stage order.  The factorisation buffer holds K after ASSEMBLE, L after POTRF and L^-1 after TRTRI / INVERSE; a stage
 * that needs different content than the buffer holds (POTRF twice, BACKSOLVE after the inverse, GRAD without INVERSE,
 * NLL after the inverse of a distributed plan) is refused with -2 instead of reading garbage.                    */
/* Weights of the two data-dependent terms of the NLL in the gradient GP b reports:
 *   grad = d/dtheta [ w_quad * 1/2 y^T K^-1 y + w_logdet * sum(log diag L) ]          (default 1, 1: the NLL)
 * The rank-3 BatchDataInput likelihood (Metrics/LogLikelihood.py:49,62-63 with Metrics/Metrics.py:152-154) averages the
 * data fit over the batch but sums the log-determinant: its gradient is the sum over b with weights (1/B, 1).     */
int gpb_plan_set_grad_weights(gpb_plan_t* plan, int b, double w_quad, double w_logdet);
/* y^T K^-1 y and sum(log diag L) of every GP as copied back by the latest gpb_plan_eval_host (no device access) */
int gpb_plan_last_terms(const gpb_plan_t* plan, double* quad_host, double* logdet_host);
/* Layout of the plan's input region: X of GP b starts x_offset bytes, y of GP b y_offset bytes into a region of
 * total_bytes.  A caller that keeps its host inputs in ONE (pinned) buffer with this layout and passes pointers into it
 * to gpb_plan_eval_host gets a single host-to-device copy for all GPs instead of two per GP.                     */
int gpb_plan_input_layout(const gpb_plan_t* plan, int b, size_t* x_offset, size_t* y_offset, size_t* total_bytes);
void gpb_plan_destroy(gpb_plan_t* plan);

/* ---- one large GP over a process grid (one process per GPU; NCCL over NVLink / NVSwitch) --------------------
 * Replaces the same reference calls as a B = 1 plan (HolisticCovarianceMatrix.get_L_K / get_L_alpha,
 * Statistics/CovarianceMatrix.py:247-265; LogLikelihood.get_metric, Metrics/LogLikelihood.py:30-65; the gradient of
 * Optimizer/Fitter.py:124-158) when one matrix is evaluated by several GPUs.
 *   factorisation : the 128 x 128 blocks of the lower triangle are owned 2D block-cyclically, block (I, J) by rank
 *                   (I mod P) * Q + ((J / W) mod Q) (W: gpb_dist_col_width); every finished panel and inverted diagonal block is broadcast
 *                   (ncclBroadcast), so all ranks end up with the complete factor L, z = L^-1 y and the same nll / info
 *   gradient      : W = L^-1, K^-1 = W^T W and the trace gradient are split by block column (J mod world), with one
 *                   exchange of W and one ncclAllReduce of the n_hp + 1 gradient entries; every rank gets the same grad
 * libnccl.so.2 is bound at run time by gpb_dist_unique_id / gpb_dist_init only.
 *   rank 0 calls gpb_dist_unique_id and ships the 128 bytes to the other ranks with the host's own transport;
 *   every rank then calls gpb_dist_init (collective), creates the plan with gpb_plan_create_dist, binds a workspace,
 *   fills the SAME X / y / hp / noise into the plan's buffers and calls gpb_plan_eval / gpb_plan_eval_host with any
 *   stage mask (collective: every rank must make the same calls in the same order).                              */
int gpb_dist_unique_id(unsigned char* id128);
int gpb_dist_init(const unsigned char* id128, int rank, int world, int P, int Q, gpb_dist_t** out);
/* The same distributed plan over a LOOP-BACK world: `world` virtual ranks of a P x Q grid inside one process on the
 * current device (out[world] handles, rank r = out[r]).  Each virtual rank gets its own plan (gpb_plan_create_dist), its
 * own workspace and must be driven by its own host thread on its own stream, all ranks making the same calls in the
 * same order - exactly as with NCCL.  The collectives become stream-ordered device copies, everything else (block
 * ownership, panel packing, launch schedule, gradient split) is the code the NCCL path runs: this is how a one-GPU test
 * box exercises the distributed path (the "virtual grid" of SURVEY section 4).                                  */
int gpb_dist_loopback_create(int world, int P, int Q, gpb_dist_t** out);
void gpb_dist_destroy(gpb_dist_t* dist);
int gpb_plan_create_dist(const gpb_program_t* prog, int64_t n, int want_grad, gpb_dist_t* dist, gpb_plan_t** out);
/* The same likelihood (stages ASSEMBLE | POTRF | NLL only) with COLUMN STORAGE on a 1 x Q grid: every rank keeps only the
 * block columns it owns (n^2 * 8 / Q bytes instead of the replicated n^2 * 8) plus a ring of three outer-panel buffers;
 * an outer panel is broadcast in place from its owner's columns into the ring (no pack / unpack), and the log-determinant,
 * z^T z and the first bad pivot are reduced over the ranks.  This is the plan for matrices that do not fit one GPU
 * (n = 131072: 17 GB per rank on 8 GPUs instead of 137 GB); GPB_BUF_A of such a plan is the packed own columns.    */
int gpb_plan_create_dist_columns(const gpb_program_t* prog, int64_t n, gpb_dist_t* dist, gpb_plan_t** out);
/* host arithmetic of the layout (no GPU needed): owner rank of block (I, J); staging order of the n_tiles blocks of
 * panel k (tile t = block row k + t): process row o sends seg_count[o] tiles starting at slot seg_base[o], the first
 * of which is tile seg_first[o], the following ones P apart.                                                    */
int gpb_dist_owner(int I, int J, int P, int Q);
/* Block columns are dealt out in groups of W = gpb_dist_col_width(P): W = 2 on a 1 x Q grid (the owner factorises a
 * 256-wide outer panel without any exchange inside it and ships it with one broadcast), W = 1 when P > 1.  Owner of
 * block (I, J) for a given W: (I mod P) * Q + ((J / W) mod Q).  gpb_dist_owned_cols lists, ascending, the block columns
 * of [J_lo, J_hi) that process column q owns (returns their number; at most cap are written).                    */
int gpb_dist_col_width(int P);
int gpb_dist_owner_w(int I, int J, int P, int Q, int W);
int gpb_dist_owned_cols(int J_lo, int J_hi, int Q, int q, int W, int* cols, int cap);
int gpb_dist_panel_segments(int k, int n_tiles, int P, int* seg_base, int* seg_count, int* seg_first);

/* ---- utilities used by the host mirror and the tests --------------------------------------------------------- */
/* C = alpha * op(A) op(B)^T + beta C on the FP64 tensor-core mainloop (a_kmajor: element (i,k) at A[k + i*lda]) */
int gpb_gemm(int a_kmajor, int b_kmajor, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
             int M, int N, int K, double alpha, double beta, void* stream);
int gpb_zero_upper(double* A, int n, int ld, void* stream);   /* zero the strict upper triangle (column-major) */
int gpb_symmetrize(double* A, int n, int ld, void* stream);   /* copy the lower triangle into the upper one    */

/* Host arithmetic (no GPU): the schedule of the triangular inverse that gpb_plan_eval overlaps with the factorisation
 * of one large matrix.  final_cols[i] = number of columns of L that are final at progress point i (non-decreasing, the
 * last one >= n).  tasks receives 4 ints per task in launch order: kind (0 copy of an inverted diagonal block, 1 phase
 * T = L21 W11, 2 phase W21 = -W22 T), level size s, index (block / sub-problem), final_cols at launch.  Returns the
 * number of tasks (also when it exceeds capacity), or a negative value if the task graph did not drain.              */
int gpb_trtri_schedule(int n, const long long* final_cols, int n_steps, int* tasks, int capacity);

/* FP64 pipe probe used by bench.py to measure the roofline denominator: kind 0 = DMMA.8x8x4 (512 flop per warp
 * instruction), kind 1 = DFMA (64 flop per warp instruction); every thread issues 8 (kind 0) / 16 (kind 1)
 * independent instructions per iteration; `blocks` CTAs of 256 threads.                                         */
int gpb_microbench(int kind, int iters, int blocks, void* stream);
/* Developer builds only (-DGPB_DIAG_CLOCKS=1): SM clock counts of the phases of the last diagonal-block launch; -1 otherwise. */
int gpb_debug_diag_clocks(long long* out8);
/* Developer tool: launch timeline of the multi-stream drivers (the reference has no counterpart; nsys is not available on
 * the pool).  Between gpb_trace_begin and gpb_trace_end every launch of gpb_plan_eval is bracketed by two CUDA events on
 * its stream.  gpb_trace_end synchronises the device and writes one text line per launch:
 * "tag a b stream start_us end_us" (times relative to gpb_trace_begin).  Not usable under stream capture.              */
int gpb_trace_begin(void* stream);
int gpb_trace_end(char* buf, size_t capacity, size_t* needed);

#ifdef __cplusplus
}
#endif
#endif /* GPB_H_ */
