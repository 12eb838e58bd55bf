// Kernel "program" interpreter: a kernel tree (KernelBasics/Kernel.py:38-140) is flattened by the host
// (gaussianprocessfundamentals_b200/program.py) into postfix code; this header evaluates that code for one
// pair (x_i, x_j) in registers.  It replaces the per-node TensorFlow calls of
//   Auxiliary/Distances.py:4-12                      (pairwise L2 / L1 distance)
//   KernelBasics/BaseKernels.py:114-134,277-294,440-457,702-720,859-880,640-662   (LIN, SE, PER, MAT32, MAT52, WN)
//   KernelBasics/Operators.py:207-225,306-326,410-476 (MUL, ADD, CP)
// and, in gradient mode, the reverse-mode sweep TF's GradientTape does over those ops (Optimizer/Fitter.py:124-132).
//
// The file is plain C++ (host+device) so that the unit tests can compile it with g++ and compare it with the oracle
// without a GPU; the product only ever runs it inside CUDA kernels.
#pragma once
#include <stdint.h>
#include <math.h>
#include "gpb_math.h"   // gpb_sincos / gpb_sin: no large-argument slow path (no local memory)

#if defined(__CUDACC__)
#define GPB_HD __host__ __device__ __forceinline__
#else
#define GPB_HD inline
#endif

// ---- opcodes: every op is 4 int32 words {op, a, b, c} --------------------------------------------------------
enum GpbOp : int32_t {
  GPB_OP_SE = 1,      // a = hp offset of l            b bit0 = scaled (sg at a+1)
  GPB_OP_PER = 2,     // a = hp offset of l, p         b bit0 = scaled (sg at a+2)
  GPB_OP_LIN = 3,     // a = hp offset of c[dim]       b bit0 = scaled (sg at a+dim)
  GPB_OP_MAT32 = 4,   // a = hp offset of l            b bit0 = scaled
  GPB_OP_MAT52 = 5,   // a = hp offset of l            b bit0 = scaled
  GPB_OP_WN = 6,      // no hyper-parameters: 1 where global row index == global column index
  GPB_OP_SE_ARD = 7,  // a = hp offset of l[dim]       b bit0 = scaled (extension; not in the reference)
  GPB_OP_L2 = 8,      // Euclidean distance itself (Auxiliary/Distances.py:4-7), no hyper-parameters
  GPB_OP_L1 = 9,      // Manhattan distance itself (Auxiliary/Distances.py:10-12), no hyper-parameters
  GPB_OP_ADD2 = 16,   // pop b, pop a, push a+b
  GPB_OP_MUL2 = 17,   // pop b, pop a, push a*b
  GPB_OP_CPW = 18     // top *= change-point window weight; a = hp offset of cp_0, b = child index, c = #children
};

// change-point modes (global_parameters.py:10-13)
enum GpbCpMode : int32_t { GPB_CP_SIGMOID = 0, GPB_CP_INDICATOR = 1, GPB_CP_APPROX_INDICATOR = 2 };

#define GPB_OP_WORDS 4
#define GPB_MAX_OPS 256
#define GPB_MAX_STACK 8
#define GPB_MAX_TAPE 64
#define GPB_MAX_DIM 16
#define GPB_MAX_HP 128

struct GpbPair {
  const double* xi;  // [dim]
  const double* xj;  // [dim]
  int dim;
  long long gi, gj;  // global indices (white noise only)
  const double* hp;  // flat hyper-parameter vector
  const double* ihp; // 1 / hp[i], precomputed once per CTA (the per-entry path has no divisions)
  int cp_mode;
};

// ---- distances (Auxiliary/Distances.py; r2 summed directly) ---------------------------------
GPB_HD double gpb_sqdist(const GpbPair& p) {
  if (p.dim == 1) { const double t = p.xi[0] - p.xj[0]; return t * t; }
  double r2 = 0.0;
  for (int d = 0; d < p.dim; ++d) { double t = p.xi[d] - p.xj[d]; r2 += t * t; }
  return r2;
}
GPB_HD double gpb_l1dist(const GpbPair& p) {
  if (p.dim == 1) return fabs(p.xi[0] - p.xj[0]);
  double s = 0.0;
  for (int d = 0; d < p.dim; ++d) s += fabs(p.xi[d] - p.xj[d]);
  return s;
}

// ---- change-point indicator s(x, cp) and d s / d cp  (Operators.py:379-400) -----------------------------------
GPB_HD double gpb_cp_s(double x, double cp, int mode, double* ds) {
  if (mode == GPB_CP_INDICATOR) { *ds = 0.0; return (x < cp) ? 1.0 : 0.0; }
  if (mode == GPB_CP_SIGMOID) {
    double t = tanh((cp - x) / 0.0025);
    *ds = 0.5 * (1.0 - t * t) / 0.0025;
    return 0.5 * (1.0 + t);
  }
  double s = 1.0 / (1.0 + exp(-1.0 * 100.0 * (x - cp)));
  *ds = -100.0 * s * (1.0 - s);
  return s;
}

// window weight of child `i` of a k-child change-point node, and its derivative w.r.t. cp_{i-1} / cp_i
GPB_HD double gpb_cp_weight(const GpbPair& p, int hp_off, int i, int k, double* dprev, double* dcur) {
  double w = 1.0; *dprev = 0.0; *dcur = 0.0;
  double x = p.xi[0], x2 = p.xj[0];
  double wp = 1.0, wc = 1.0;
  if (i > 0) {
    double da, da2;
    double a = gpb_cp_s(x, p.hp[hp_off + i - 1], p.cp_mode, &da);
    double a2 = gpb_cp_s(x2, p.hp[hp_off + i - 1], p.cp_mode, &da2);
    wp = (1.0 - a) * (1.0 - a2);
    *dprev = -(da * (1.0 - a2) + (1.0 - a) * da2);
  }
  if (i < k - 1) {
    double db, db2;
    double b = gpb_cp_s(x, p.hp[hp_off + i], p.cp_mode, &db);
    double b2 = gpb_cp_s(x2, p.hp[hp_off + i], p.cp_mode, &db2);
    wc = b * b2;
    *dcur = db * b2 + b * db2;
  }
  w = wp * wc;
  *dprev *= wc;
  *dcur *= wp;
  return w;
}

// ---- leaf kernels -------------------------------------------------------------------------------------------
// Each returns the (scaled) value; `dk` receives d value / d hp for the leaf's hyper-parameters in list order
// when it is non-null (LIN and SE_ARD write `dim` entries first).
GPB_HD double gpb_leaf(int op, int a, int flags, const GpbPair& p, double* dk) {
  const bool scaled = (flags & 1) != 0;
  const double* h = p.hp + a;
  const double* ih = p.ihp + a;
  double k0 = 0.0;
  int nq = 0;
  switch (op) {
    case GPB_OP_SE: {
      const double il = ih[0];                 // 1 / l
      const double r2 = gpb_sqdist(p);
      const double q = r2 * (il * il);         // r2 / l^2
      k0 = exp(-0.5 * q);
      if (dk) dk[0] = k0 * q * il;             // k r2 / l^3
      nq = 1;
    } break;
    case GPB_OP_PER: {
      const double il = ih[0], ip = ih[1];     // 1 / l, 1 / p
      const double D = gpb_l1dist(p);
      // the one true division of the path: u is the argument of sin and can be O(100), so an extra rounding of D / p
      // would be amplified into the entry (the reference computes pi * (D / p), BaseKernels.py:447)
      const double u = M_PI * (D / h[1]);
      double s, c;
      if (dk) gpb_sincos(u, &s, &c); else { s = gpb_sin(u); c = 0.0; }
      const double sine = s * s;
      const double il2 = il * il;
      k0 = exp((-2.0 * sine) * il2);
      if (dk) {
        dk[0] = k0 * (4.0 * sine) * (il2 * il);
        dk[1] = k0 * (2.0 * M_PI * D * (2.0 * s * c)) * (il2 * (ip * ip));
      }
      nq = 2;
    } break;
    case GPB_OP_LIN: {
      double acc = 0.0;
      for (int d = 0; d < p.dim; ++d) acc += (p.xi[d] - h[d]) * (p.xj[d] - h[d]);
      k0 = acc;
      if (dk) for (int d = 0; d < p.dim; ++d) dk[d] = 2.0 * h[d] - p.xi[d] - p.xj[d];
      nq = p.dim;
    } break;
    case GPB_OP_MAT32: {
      const double il = fabs(ih[0]);           // 1 / |l|
      const double D = gpb_l1dist(p);
      const double f = (sqrt(3.0) * D) * il;
      const double e = exp(-f);
      k0 = (1.0 + f) * e;
      if (dk) dk[0] = (f * f * e * il) * (h[0] < 0.0 ? -1.0 : 1.0);
      nq = 1;
    } break;
    case GPB_OP_MAT52: {
      const double il = fabs(ih[0]);
      const double D = gpb_l1dist(p);
      const double f = (sqrt(5.0) * D) * il;
      const double third = (5.0 * (D * D)) * ((il * il) * (1.0 / 3.0));
      const double e = exp(-f);
      k0 = (1.0 + f + third) * e;
      if (dk) dk[0] = (f * f * (1.0 + f) * (1.0 / 3.0) * e * il) * (h[0] < 0.0 ? -1.0 : 1.0);
      nq = 1;
    } break;
    case GPB_OP_WN: {
      k0 = (p.gi == p.gj) ? 1.0 : 0.0;
      nq = 0;
    } break;
    case GPB_OP_L2: return sqrt(gpb_sqdist(p));
    case GPB_OP_L1: return gpb_l1dist(p);
    case GPB_OP_SE_ARD: {
      double r2 = 0.0;
      for (int d = 0; d < p.dim; ++d) { const double t = (p.xi[d] - p.xj[d]) * ih[d]; r2 += t * t; }
      k0 = exp(-0.5 * r2);
      if (dk) for (int d = 0; d < p.dim; ++d) {
        const double t = (p.xi[d] - p.xj[d]) * ih[d];
        dk[d] = k0 * (t * t) * ih[d];
      }
      nq = p.dim;
    } break;
    default: break;
  }
  if (scaled) {
    double sg = h[nq];
    if (dk) { for (int q = 0; q < nq; ++q) dk[q] *= sg; dk[nq] = k0; }
    return sg * k0;
  }
  return k0;
}

// number of hyper-parameter scalars a leaf consumes
GPB_HD int gpb_leaf_nhp(int op, int flags, int dim) {
  int nq = 0;
  switch (op) {
    case GPB_OP_SE: nq = 1; break;
    case GPB_OP_PER: nq = 2; break;
    case GPB_OP_LIN: nq = dim; break;
    case GPB_OP_MAT32: nq = 1; break;
    case GPB_OP_MAT52: nq = 1; break;
    case GPB_OP_WN: return 0;
    case GPB_OP_L2: return 0;
    case GPB_OP_L1: return 0;
    case GPB_OP_SE_ARD: nq = dim; break;
    default: return 0;
  }
  return nq + ((flags & 1) ? 1 : 0);
}

// ---- value-only evaluation: operand stack kept in 8 named registers (no local memory) -------------------------
GPB_HD double gpb_eval(const int32_t* code, int n_ops, const GpbPair& p) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
  for (int pc = 0; pc < n_ops; ++pc) {
    const int32_t* w = code + pc * GPB_OP_WORDS;
    const int op = w[0];
    if (op < GPB_OP_ADD2) {
      double v = gpb_leaf(op, w[1], w[2], p, nullptr);
      s7 = s6; s6 = s5; s5 = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = v;
    } else if (op == GPB_OP_CPW) {
      double d0, d1;
      s0 = s0 * gpb_cp_weight(p, w[1], w[2], w[3], &d0, &d1);
    } else {
      s0 = (op == GPB_OP_ADD2) ? (s1 + s0) : (s1 * s0);
      s1 = s2; s2 = s3; s3 = s4; s4 = s5; s5 = s6; s6 = s7;
    }
  }
  return s0;
}

// ---- value + reverse-mode gradient -----------------------------------------------------------------------------
// Returns the kernel value; calls acc(hp_index, adj_root * d value / d hp[hp_index]) for every hyper-parameter.
template <class Acc>
GPB_HD double gpb_eval_grad(const int32_t* code, int n_ops, const GpbPair& p, double adj_root, Acc& acc) {
  double st[GPB_MAX_STACK];
  double tape[GPB_MAX_TAPE];
  int sp = 0, tp = 0;
  for (int pc = 0; pc < n_ops; ++pc) {
    const int32_t* w = code + pc * GPB_OP_WORDS;
    const int op = w[0];
    if (op < GPB_OP_ADD2) {
      double dk[GPB_MAX_DIM + 1];
      double v = gpb_leaf(op, w[1], w[2], p, dk);
      const int nq = gpb_leaf_nhp(op, w[2], p.dim);
      for (int q = 0; q < nq; ++q) tape[tp + q] = dk[q];
      tp += nq;
      st[sp++] = v;
    } else if (op == GPB_OP_CPW) {
      double d0, d1;
      double wgt = gpb_cp_weight(p, w[1], w[2], w[3], &d0, &d1);
      tape[tp++] = st[sp - 1];
      tape[tp++] = wgt;
      tape[tp++] = d0;
      tape[tp++] = d1;
      st[sp - 1] *= wgt;
    } else {
      double b = st[--sp];
      double a = st[sp - 1];
      if (op == GPB_OP_MUL2) { tape[tp++] = a; tape[tp++] = b; st[sp - 1] = a * b; }
      else st[sp - 1] = a + b;
    }
  }
  const double value = st[0];
  // backward sweep
  double adj[GPB_MAX_STACK];
  int ap = 0;
  adj[ap++] = adj_root;
  for (int pc = n_ops - 1; pc >= 0; --pc) {
    const int32_t* w = code + pc * GPB_OP_WORDS;
    const int op = w[0];
    if (op < GPB_OP_ADD2) {
      const int nq = gpb_leaf_nhp(op, w[2], p.dim);
      const double g = adj[--ap];
      tp -= nq;
      for (int q = 0; q < nq; ++q) acc(w[1] + q, g * tape[tp + q]);
    } else if (op == GPB_OP_CPW) {
      tp -= 4;
      const double g = adj[ap - 1];
      const double v = tape[tp], wgt = tape[tp + 1];
      if (p.cp_mode != GPB_CP_INDICATOR) {
        if (w[2] > 0) acc(w[1] + w[2] - 1, g * v * tape[tp + 2]);
        if (w[2] < w[3] - 1) acc(w[1] + w[2], g * v * tape[tp + 3]);
      }
      adj[ap - 1] = g * wgt;
    } else if (op == GPB_OP_MUL2) {
      tp -= 2;
      const double g = adj[--ap];
      adj[ap++] = g * tape[tp + 1];  // adjoint of first operand (deeper in the stack)
      adj[ap++] = g * tape[tp];      // adjoint of second operand (its ops come first in reverse order)
    } else {
      const double g = adj[--ap];
      adj[ap++] = g;
      adj[ap++] = g;
    }
  }
  return value;
}
