// FP64 tensor-core GEMM mainloop shared by every dense contraction on the path:
//   Cholesky trailing update / panel solve   (replaces tf.linalg.cholesky, Statistics/CovarianceMatrix.py:250,475)
//   triangular inverse and inv(K) = W^T W   (replaces the Cholesky backward pass TF runs for Optimizer/Fitter.py:124-158)
//
// C[BM x BN tile] = alpha * op(A) * op(B)^T + beta * C, FP64, DMMA.8x8x4 (mma.sync m8n8k4.f64).
// Operands are staged with 16-byte cp.async (zero-filling out-of-range rows / k) into padded, bank-conflict-free
// shared-memory tiles, 3 stages deep.  Three tile configurations share the code:
//   CfgBig   128 x 128, 8 warps (4 x 2), warp tile 32 x 64, one CTA per SM
//   CfgQuarter 32 x 128, 4 warps (1 x 4), warp tile 32 x 32, for launches smaller than one wave of CfgHalf tiles
//   CfgHalf   64 x 128, 4 warps (2 x 2), warp tile 32 x 64, two CTAs per SM (the second CTA hides the first one's
//             barrier / prologue / epilogue bubbles) - the default of every driver
//
// An operand is "MN-major" when element (mn, k) lives at P[mn + k*ld] (column-major op(A)=A) and "K-major" when it
// lives at P[k + mn*ld] (column-major op(A)=A^T); both are supported for A and B so that NT / NN / TN products of
// column-major matrices need no transposes in memory.
#pragma once
#include <cstdlib>
#include "common.cuh"

namespace gpb {

constexpr int G_BK = 16, G_STAGES = 3;
#ifndef GPB_EPI
#define GPB_EPI 4
#endif
constexpr int G_EPI = GPB_EPI;     // column groups (of 8) per epilogue batch
constexpr int G_LDK = G_BK + 4;    // pitch (doubles) of a K-major tile    [rows][BK+4]

template <int BM_, int BN_, int WARPS_M_, int WARPS_N_, int MIN_CTAS_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, MIN_CTAS = MIN_CTAS_;
  static constexpr int THREADS = 32 * WARPS_M * WARPS_N;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;   // warp tile
  static constexpr int FM = WM / 8, FN = WN / 8;               // 8x8 DMMA tiles per warp
  static constexpr int A_TILE = BM * G_LDK > G_BK * (BM + 4) ? BM * G_LDK : G_BK * (BM + 4);
  static constexpr int B_TILE = BN * G_LDK > G_BK * (BN + 4) ? BN * G_LDK : G_BK * (BN + 4);
  static constexpr int SMEM_BYTES = G_STAGES * (A_TILE + B_TILE) * (int)sizeof(double);
};
using CfgBig = GemmCfg<128, 128, 4, 2, 1>;
using CfgHalf = GemmCfg<64, 128, 2, 2, 2>;
// 32 x 128, 4 warps (1 x 4), warp tile 32 x 32: for the single-wave launches of the factorisation's critical path (panel
// solve, next-column update) - twice the CTAs of CfgHalf at half the tile latency when the launch has < 148 tiles anyway
using CfgQuarter = GemmCfg<32, 128, 1, 4, 2>;

// 1: accumulate-into epilogues (beta == 1) use RED; set per translation unit at init (GPB_RED=0 switches it off)
static __constant__ int g_red_epilogue = 1;

struct TileJob {
  const double* A;  // tile-row origin of op(A): MN-major -> &A[i0], K-major -> &A[i0*lda]
  const double* B;  // tile-col origin of op(B)
  double* C;        // &C[i0 + j0*ldc]
  int lda, ldb, ldc;
  int mrem, nrem;   // valid rows / cols of this tile
  int klo, khi;     // contraction range in operand coordinates
  double alpha, beta;
  int red;          // beta == 1 and red != 0: C += alpha * acc through fire-and-forget RED.ADD.F64 (no read of C)
};

// one BK-deep slab of an operand tile with ROWS rows (m or n), by THREADS threads
template <bool KM, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(double* s, const double* g, int ld, int rem, int k0, int khi, int tid) {
  if (!KM) {
    constexpr int PAIRS = ROWS / 2;              // 16-byte chunks per k row
    constexpr int KSTEP = THREADS / PAIRS;       // k rows covered per pass
    const int mn = (tid % PAIRS) * 2;
    const int kb = tid / PAIRS;
    int vm = rem - mn;
    vm = vm < 0 ? 0 : (vm > 2 ? 2 : vm);
#pragma unroll
    for (int q = 0; q < G_BK / KSTEP; ++q) {
      const int k = kb + KSTEP * q;
      const int valid = (k0 + k < khi) ? vm : 0;
      const double* src = valid ? g + mn + (size_t)(k0 + k) * ld : g;
      cp_async16(s + k * (ROWS + 4) + mn, src, valid * 8);
    }
  } else {
    constexpr int RSTEP = THREADS / 8;
    const int k = (tid & 7) * 2;
    const int mb = tid >> 3;
    int vk = khi - (k0 + k);
    vk = vk < 0 ? 0 : (vk > 2 ? 2 : vk);
#pragma unroll
    for (int q = 0; q < ROWS / RSTEP; ++q) {
      const int mn = mb + RSTEP * q;
      const int valid = (mn < rem) ? vk : 0;
      const double* src = valid ? g + (k0 + k) + (size_t)mn * ld : g;
      cp_async16(s + mn * G_LDK + k, src, valid * 8);
    }
  }
}

// Epilogue of one thread: acc[f][g][e] is element (f*8, g*8 + e) relative to the thread's first element c0.
// FULL tiles carry no bounds checks.  Accumulate-into tiles (beta == 1, J.red): every element of C is touched by exactly
// one thread of one CTA per launch, so RED.ADD.F64 is deterministic and C never travels through the SM; otherwise a
// read-modify-write in batches of G_EPI column groups (all loads of a batch issued before the first store).
template <class Cfg, bool FULL>
__device__ __forceinline__ void gemm_epilogue(double (&acc)[Cfg::FM][Cfg::FN][2], const TileJob& J, double* c0, int rrem,
                                              int crem) {
  constexpr int FM = Cfg::FM, FN = Cfg::FN;
  const double alpha = J.alpha, beta = J.beta;
  const size_t ldc = J.ldc;
  if (J.red) {
#pragma unroll
    for (int g = 0; g < FN; ++g)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double* cp = c0 + (size_t)(g * 8 + e) * ldc;
        if (FULL || g * 8 + e < crem) {
#pragma unroll
          for (int f = 0; f < FM; ++f)
            if (FULL || f * 8 < rrem) red_add_f64(cp + f * 8, alpha * acc[f][g][e]);
        }
      }
    return;
  }
#pragma unroll
  for (int gp = 0; gp < FN / G_EPI; ++gp) {
    double cv[G_EPI][2][FM];
    if (beta != 0.0) {
#pragma unroll
      for (int gg = 0; gg < G_EPI; ++gg)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc = (G_EPI * gp + gg) * 8 + e;
          const double* cp = c0 + (size_t)cc * ldc;
#pragma unroll
          for (int f = 0; f < FM; ++f) cv[gg][e][f] = (FULL || (cc < crem && f * 8 < rrem)) ? cp[f * 8] : 0.0;
        }
    }
#pragma unroll
    for (int gg = 0; gg < G_EPI; ++gg)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int g = G_EPI * gp + gg;
        const int cc = g * 8 + e;
        double* cp = c0 + (size_t)cc * ldc;
        if (FULL || cc < crem) {
#pragma unroll
          for (int f = 0; f < FM; ++f) {
            if (FULL || f * 8 < rrem) {
              double v = alpha * acc[f][g][e];
              if (beta != 0.0) v += beta * cv[gg][e][f];
              cp[f * 8] = v;
            }
          }
        }
      }
  }
}

// Per-thread constants of the fast (unpredicated) operand loads of full tiles: chunk q of a slab lives at
// g + g_off + q * QROWS * ld in global memory (g = slab origin) and at s_off + q * S_Q in the stage buffer.
template <bool KM, int ROWS, int THREADS>
struct FastLoad {
  static constexpr int PAIRS = ROWS / 2;
  static constexpr int QROWS = KM ? THREADS / 8 : THREADS / PAIRS;   // operand rows (KM) / k rows (MN-major) per pass
  static constexpr int NQ = KM ? ROWS / QROWS : G_BK / QROWS;
  static constexpr int S_Q = KM ? QROWS * G_LDK : QROWS * (ROWS + 4);
  static __device__ __forceinline__ void setup(int tid, int& r0, int& c0, int& s_off) {
    if (!KM) { c0 = (tid % PAIRS) * 2; r0 = tid / PAIRS; s_off = r0 * (ROWS + 4) + c0; }   // r0: k row, c0: mn
    else { c0 = (tid & 7) * 2; r0 = tid >> 3; s_off = r0 * G_LDK + c0; }                    // r0: mn row, c0: k
  }
};

__device__ __forceinline__ void cp_async16_full(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

// ---- bulk-copy (TMA engine) operand staging: the BULK variant of the mainloop -------------------------------------
// One `cp.async.bulk.shared.global` per k-row of an MN-major slab (BM or BN contiguous doubles, 512 B / 1 KB) into the
// same padded stage layout, issued by the lanes of warp 0 and completed on one mbarrier per stage (complete_tx), instead
// of 12 LDGSTS + address adds per thread per slab.  No tensor map is needed (the rows of a column-major panel are
// contiguous), and the padded pitch that keeps the DMMA fragment loads bank-conflict free survives - a tiled
// `cp.async.bulk.tensor` box would be dense (pitch 512 B: 4-way conflicts; SWIZZLE_128B only reaches 2-way for f64).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
constexpr int G_BULK_EXTRA = 64;   // bytes behind the stage buffers: G_STAGES mbarriers

// Tile loop of a CTA.  A CTA walks the virtual grid `vg` (the grid a one-tile-per-CTA launch would use) with stride
// gridDim.x and keeps ONE cp.async pipeline running across tile boundaries: with fewer CTAs than tiles (persistent
// launch) the first operand slabs of tile i+1 are in flight while tile i finishes, so a tile costs its k-steps plus its
// epilogue and no pipeline fill; with gridDim.x = number of tiles it degenerates to one tile per CTA under the hardware
// scheduler.  Full tiles (the overwhelming majority) take a fast path whose per-slab load issue is one address add
// and one LDGSTS per 16-byte chunk and whose epilogue has no bounds checks; edge tiles keep the zero-filling loads.
// Geometry functors map a virtual block index to a TileJob: `bool Geo::tile<BM, BN>(TileJob&, const dim3& b)`.
template <class Cfg, bool AKM, bool BKM, class Geo, bool BULK = false>
__device__ __forceinline__ void gemm_stream(const Geo& geo, const dim3 vg) {
  static_assert(!BULK || (!AKM && !BKM && G_BK == 16), "bulk staging: MN-major operands, 16 k-rows = 16 lanes each");
  extern __shared__ __align__(16) double gsm[];
  constexpr int BM = Cfg::BM, BN = Cfg::BN, T = Cfg::THREADS, FM = Cfg::FM, FN = Cfg::FN;
  constexpr int STAGE = Cfg::A_TILE + Cfg::B_TILE;
  using FA = FastLoad<AKM, BM, T>;
  using FB = FastLoad<BKM, BN, T>;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
  const int lr = lane >> 2, lk = lane & 3;
  const unsigned total = vg.x * vg.y * vg.z;     // < 2^31 for every launch of the path
  const unsigned stride = gridDim.x;
  const bool flat = (vg.y == 1 && vg.z == 1);
  int a_r0, a_c0, a_soff, b_r0, b_c0, b_soff;
  FA::setup(tid, a_r0, a_c0, a_soff);
  FB::setup(tid, b_r0, b_c0, b_soff);
  uint64_t* full = reinterpret_cast<uint64_t*>(gsm + G_STAGES * STAGE);   // BULK: one mbarrier per stage
  unsigned bulk_mask = 0, phase_mask = 0;                                 // stage filled by bulk copies / its phase parity
  if constexpr (BULK) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < G_STAGES; ++s) mbar_init(&full[s], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }

  // next valid tile at or after idx
  auto fetch = [&](unsigned& idx, TileJob& J) -> bool {
    while (idx < total) {
      dim3 b(idx, 0, 0);
      if (!flat) { const unsigned xy = idx / vg.x; b.x = idx - xy * vg.x; b.z = xy / vg.y; b.y = xy - b.z * vg.y; }
      if (geo.template tile<BM, BN>(J, b)) return true;
      idx += stride;
    }
    return false;
  };
  auto nk_of = [](const TileJob& J) { return (J.khi > J.klo) ? (J.khi - J.klo + G_BK - 1) / G_BK : 0; };

  // ---- load cursor: runs up to G_STAGES-1 slabs ahead of the compute cursor, across tiles -------------------------
  TileJob Jl;
  unsigned lidx = blockIdx.x;
  int l_kt = 0, l_nk = 0;
  bool l_ok = false, l_fast = false;
  const double* l_pa = nullptr;   // this thread's first chunk of the next slab (fast path)
  const double* l_pb = nullptr;
  auto next_load_tile = [&]() {
    for (;;) {
      l_ok = fetch(lidx, Jl);
      if (!l_ok) return;
      l_nk = nk_of(Jl);
      if (l_nk > 0) break;
      lidx += stride;
    }
    l_kt = 0;
    l_fast = (Jl.mrem == BM) && (Jl.nrem == BN) && ((Jl.khi - Jl.klo) % G_BK == 0);
    l_pa = AKM ? Jl.A + Jl.klo + a_c0 + (size_t)a_r0 * Jl.lda : Jl.A + a_c0 + (size_t)(Jl.klo + a_r0) * Jl.lda;
    l_pb = BKM ? Jl.B + Jl.klo + b_c0 + (size_t)b_r0 * Jl.ldb : Jl.B + b_c0 + (size_t)(Jl.klo + b_r0) * Jl.ldb;
    if constexpr (BULK) {     // lane q of warp 0 copies k-row q & 15 of A (q < 16) or B: its own row pointer
      l_pa = Jl.A + (size_t)(Jl.klo + (lane & 15)) * Jl.lda;
      l_pb = Jl.B + (size_t)(Jl.klo + (lane & 15)) * Jl.ldb;
    }
  };
  next_load_tile();
  auto issue = [&](int stage) {
    if (l_ok) {
      double* Ns = gsm + stage * STAGE;
      if (BULK && l_fast) {
        bulk_mask |= 1u << stage;
        if (warp == 0) {
          if (lane == 0) mbar_expect_tx(&full[stage], (unsigned)((BM + BN) * G_BK * sizeof(double)));
          __syncwarp();
          if (lane < 16) bulk_g2s(Ns + lane * (BM + 4), l_pa, BM * (unsigned)sizeof(double), &full[stage]);
          else bulk_g2s(Ns + Cfg::A_TILE + (lane - 16) * (BN + 4), l_pb, BN * (unsigned)sizeof(double), &full[stage]);
        }
        l_pa += (size_t)G_BK * Jl.lda;
        l_pb += (size_t)G_BK * Jl.ldb;
      } else if (l_fast) {
#pragma unroll
        for (int q = 0; q < FA::NQ; ++q)
          cp_async16_full(Ns + a_soff + q * FA::S_Q, l_pa + (size_t)(q * FA::QROWS) * Jl.lda);
#pragma unroll
        for (int q = 0; q < FB::NQ; ++q)
          cp_async16_full(Ns + Cfg::A_TILE + b_soff + q * FB::S_Q, l_pb + (size_t)(q * FB::QROWS) * Jl.ldb);
        l_pa += AKM ? (size_t)G_BK : (size_t)G_BK * Jl.lda;
        l_pb += BKM ? (size_t)G_BK : (size_t)G_BK * Jl.ldb;
      } else {
        if constexpr (BULK) bulk_mask &= ~(1u << stage);
        load_tile<AKM, BM, T>(Ns, Jl.A, Jl.lda, Jl.mrem, Jl.klo + l_kt * G_BK, Jl.khi, tid);
        load_tile<BKM, BN, T>(Ns + Cfg::A_TILE, Jl.B, Jl.ldb, Jl.nrem, Jl.klo + l_kt * G_BK, Jl.khi, tid);
      }
      if (++l_kt == l_nk) { lidx += stride; next_load_tile(); }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) issue(s);

  // ---- compute cursor -----------------------------------------------------------------------------------------------
  TileJob J;
  unsigned cidx = blockIdx.x;
  int step = 0;   // slabs consumed so far (ring position)
  while (fetch(cidx, J)) {
    double acc[FM][FN][2];
#pragma unroll
    for (int f = 0; f < FM; ++f)
#pragma unroll
      for (int g = 0; g < FN; ++g) { acc[f][g][0] = 0.0; acc[f][g][1] = 0.0; }
    const int nk = nk_of(J);

    if (J.beta != 0.0 && !J.red) {
      // the accumulate-into tile is needed only by the epilogue: pull it into L2 now (BN columns x BM/16 lines)
      constexpr int SEGS = BM / 16;
#pragma unroll
      for (int q = 0; q < (BN * SEGS + T - 1) / T; ++q) {
        const int idx = tid + T * q;
        const int col = idx / SEGS, seg = (idx % SEGS) * 16;
        if (col < J.nrem && seg < J.mrem) {
          const double* pp = J.C + (size_t)col * J.ldc + seg;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
        }
      }
    }

    for (int kt = 0; kt < nk; ++kt) {
      cp_async_wait<G_STAGES - 2>();
      if constexpr (BULK) {
        const int st = step % G_STAGES;
        if ((bulk_mask >> st) & 1u) { mbar_wait(&full[st], (phase_mask >> st) & 1u); phase_mask ^= 1u << st; }
      }
      __syncthreads();
      issue((step + G_STAGES - 1) % G_STAGES);
      const double* As = gsm + (step % G_STAGES) * STAGE;
      const double* Bs = As + Cfg::A_TILE;
      ++step;
#pragma unroll
      for (int kk = 0; kk < G_BK / 4; ++kk) {
        const int kidx = kk * 4 + lk;
        double a[FM], b[FN];
#pragma unroll
        for (int f = 0; f < FM; ++f) {
          const int row = wm * Cfg::WM + f * 8 + lr;
          a[f] = AKM ? As[row * G_LDK + kidx] : As[kidx * (BM + 4) + row];
        }
#pragma unroll
        for (int g = 0; g < FN; ++g) {
          const int col = wn * Cfg::WN + g * 8 + lr;
          b[g] = BKM ? Bs[col * G_LDK + kidx] : Bs[kidx * (BN + 4) + col];
        }
#pragma unroll
        for (int f = 0; f < FM; ++f)
#pragma unroll
          for (int g = 0; g < FN; ++g) dmma884(acc[f][g][0], acc[f][g][1], a[f], b[g]);
      }
    }

    // ---- epilogue: each quad-row of 8 lanes covers 8 consecutive rows (64 B) of one column -----------------------
    double* c0 = J.C + (wm * Cfg::WM + lr) + (size_t)(wn * Cfg::WN + 2 * lk) * J.ldc;   // this thread's first element
    const int rrem = J.mrem - (wm * Cfg::WM + lr), crem = J.nrem - (wn * Cfg::WN + 2 * lk);
    if ((J.mrem == BM) && (J.nrem == BN)) gemm_epilogue<Cfg, true>(acc, J, c0, rrem, crem);
    else gemm_epilogue<Cfg, false>(acc, J, c0, rrem, crem);
    cidx += stride;
  }
  cp_async_wait<0>();
}

// The same lower-triangular tile set in SUPER-TILE order: tile columns are taken in groups of G; inside a group the
// row tiles run from the group's first diagonal tile to the bottom and the column index runs fastest.  Concurrently
// running CTAs then cover ~(wave / G) row tiles x G column tiles instead of a whole column sweep, which cuts the distinct
// operand columns a wave touches (and so the re-reads of a matrix larger than L2) by ~2.5x for a 296-CTA wave.
__device__ __forceinline__ bool tri_map_grouped(unsigned idx, int Tm, int R, int G, int& ti, int& tj) {
  const int Tn = (Tm + R - 1) / R;
  int rem = (int)idx;
  for (int g0 = 0; g0 < Tn; g0 += G) {
    const int gw = min(G, Tn - g0);                 // columns of this group
    const int row0 = R * g0;                        // first row tile of the group
    const int rows = Tm - row0;
    const int stair_rows = min(rows, R * gw);       // rows in which not every column of the group is reachable yet
    // row r (relative) reaches floor(r / R) + 1 columns (capped at gw)
    int stair = R * gw * (gw + 1) / 2;              // closed form when the staircase is complete
    if (stair_rows < R * gw) { stair = 0; for (int r = 0; r < stair_rows; ++r) stair += min(gw, r / R + 1); }
    const int cnt = stair + (rows - stair_rows) * gw;
    if (rem < cnt) {
      if (rem < stair) {
        int r = 0;
        for (;;) { const int w = min(gw, r / R + 1); if (rem < w) break; rem -= w; ++r; }
        ti = row0 + r; tj = g0 + rem;
      } else {
        rem -= stair;
        const int q = rem / gw;
        ti = row0 + stair_rows + q; tj = g0 + (rem - q * gw);
      }
      return true;
    }
    rem -= cnt;
  }
  return false;
}

template <class Cfg, bool AKM, bool BKM, class Geo>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) gemm_kernel(const Geo geo, const dim3 vgrid) {
  gemm_stream<Cfg, AKM, BKM, Geo>(geo, vgrid);
}
// the same tile loop with bulk-copy operand staging (MN-major operands only; launched with SMEM_BYTES + G_BULK_EXTRA)
template <class Cfg, class Geo>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) gemm_bulk_kernel(const Geo geo, const dim3 vgrid) {
  gemm_stream<Cfg, false, false, Geo, true>(geo, vgrid);
}

// CTAs of a launch over `vgrid` tiles.  max_tiles_per_cta = 1: one CTA per tile.  Larger values let a CTA keep its
// cp.async pipeline running across up to that many tiles (no pipeline fill per tile), but never fewer CTAs than one
// resident wave.  The bound keeps CTA lifetimes short: the kernels of the factorisation's critical path (diagonal block,
// panel, NCCL broadcast) run on a high-priority stream and can only start when a CTA of the bulk update retires -
// fully persistent CTAs serialised them behind the whole update (x1.25 on the distributed factorisation).
inline unsigned persistent_ctas(dim3 vgrid, int min_ctas_per_sm, int max_tiles_per_cta) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  static int env_max = -1;
  if (env_max < 0) {
    const char* e = getenv("GPB_TILES_PER_CTA");
    env_max = e ? atoi(e) : 0;
  }
  if (env_max > 0 && max_tiles_per_cta > 1) max_tiles_per_cta = env_max;
  const long long total = (long long)vgrid.x * vgrid.y * vgrid.z;
  if (max_tiles_per_cta <= 1) return (unsigned)total;
  const long long wave = (long long)sms * min_ctas_per_sm;
  const long long bounded = (total + max_tiles_per_cta - 1) / max_tiles_per_cta;
  long long g = total < wave ? total : (bounded > wave ? bounded : wave);
  return (unsigned)g;
}

}  // namespace gpb
