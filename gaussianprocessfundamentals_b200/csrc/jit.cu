// Run-time specialisation of the covariance assembly and the trace gradient per kernel program.
//
// The interpreter of program.cuh evaluates a kernel tree (KernelBasics/Kernel.py:51; BaseKernels.py:114-134,277-294,
// 440-457; Operators.py:207-225,306-326,442-476) from postfix code per matrix entry: opcode dispatch, an operand stack
// and - for the gradient - a tape and an adjoint stack that are indexed dynamically and therefore live in local memory
// (533 / 789 warp instructions per entry on the benchmark kernel, half of them interpretation).  Here the same postfix
// code is turned into STRAIGHT-LINE device code once per program: every stack slot, tape entry and adjoint becomes a named
// scalar, hyper-parameter offsets and the input dimensionality become constants, the pairwise distances are shared by
// all leaves.  The text is compiled for the device in use with NVRTC (libnvrtc.so.12, bound with dlopen like NCCL: it is
// part of the CUDA toolkit / of PyTorch's wheels) and launched through the driver API.  The numerical formulas are the
// interpreter's, term by term, so both paths agree to rounding (FMA contraction is the only freedom).
// When NVRTC cannot be loaded (GPB_JIT=0, or a host without the toolkit) programs stay on the interpreter kernels.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "internal.h"
#include "jit.h"
#include "program.cuh"

#include "jit_embed.inc"   // kDeviceAbiSrc, kMathSrc, kSpecKernelsSrc: the device headers as text (build.py)

namespace gpb {

// ---------------------------------------------------------------------------------------------------------------
// code generation
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Node {
  int kind;          // GPB_OP_*
  int a = -1, b = -1;  // operand node ids (ADD2 / MUL2: a op b; CPW: a)
  int off = 0, flags = 0, child = 0, nchild = 0;
  int nq = 0;        // hyper-parameters of a leaf (incl. sg)
};

struct Gen {
  int dim, cp_mode;
  bool with_grad;
  std::ostringstream o;
  std::vector<Node> nodes;
  bool need_r2 = false, need_l1 = false;

  std::string xi(int d) const { return "xi[" + std::to_string(d) + "]"; }
  std::string xj(int d) const { return "xj[" + std::to_string(d) + "]"; }
  std::string H(int i) const { return "h[" + std::to_string(i) + "]"; }
  std::string IH(int i) const { return "ih[" + std::to_string(i) + "]"; }

  // s(x, cp) and ds/dcp of one change-point mode (Operators.py:379-400), names: S<tag>, D<tag>
  void cp_s(const std::string& tag, const std::string& x, const std::string& cp) {
    if (cp_mode == GPB_CP_INDICATOR) {
      o << "    const double S" << tag << " = (" << x << " < " << cp << ") ? 1.0 : 0.0;\n";
      if (with_grad) o << "    const double D" << tag << " = 0.0;\n";
    } else if (cp_mode == GPB_CP_SIGMOID) {
      o << "    const double T" << tag << " = tanh((" << cp << " - " << x << ") / 0.0025);\n";
      if (with_grad) o << "    const double D" << tag << " = 0.5 * (1.0 - T" << tag << " * T" << tag << ") / 0.0025;\n";
      o << "    const double S" << tag << " = 0.5 * (1.0 + T" << tag << ");\n";
    } else {
      o << "    const double S" << tag << " = 1.0 / (1.0 + exp(-1.0 * 100.0 * (" << x << " - " << cp << ")));\n";
      if (with_grad) o << "    const double D" << tag << " = -100.0 * S" << tag << " * (1.0 - S" << tag << ");\n";
    }
  }

  void leaf(int k, const Node& nd) {
    const std::string K = std::to_string(k);
    const int a = nd.off;
    const bool scaled = (nd.flags & 1) != 0;
    int nq = 0;
    o << "    // node " << k << ": leaf op " << nd.kind << " at hp[" << a << "]\n";
    switch (nd.kind) {
      case GPB_OP_SE:
        o << "    const double il" << K << " = " << IH(a) << ";\n";
        o << "    const double q" << K << " = r2 * (il" << K << " * il" << K << ");\n";
        o << "    const double k0_" << K << " = exp(-0.5 * q" << K << ");\n";
        if (with_grad) o << "    double d" << K << "_0 = k0_" << K << " * q" << K << " * il" << K << ";\n";
        nq = 1;
        break;
      case GPB_OP_PER:
        o << "    const double il" << K << " = " << IH(a) << ", ip" << K << " = " << IH(a + 1) << ";\n";
        // the one true division of the path (see program.cuh): the reference computes pi * (D / p)
        o << "    const double u" << K << " = 3.14159265358979323846 * (l1 / " << H(a + 1) << ");\n";
        if (with_grad) o << "    double s" << K << ", c" << K << "; gpb_sincos(u" << K << ", &s" << K << ", &c" << K << ");\n";
        else o << "    const double s" << K << " = gpb_sin(u" << K << ");\n";
        o << "    const double sine" << K << " = s" << K << " * s" << K << ";\n";
        o << "    const double il2_" << K << " = il" << K << " * il" << K << ";\n";
        o << "    const double k0_" << K << " = exp((-2.0 * sine" << K << ") * il2_" << K << ");\n";
        if (with_grad) {
          o << "    double d" << K << "_0 = k0_" << K << " * (4.0 * sine" << K << ") * (il2_" << K << " * il" << K << ");\n";
          o << "    double d" << K << "_1 = k0_" << K << " * (2.0 * 3.14159265358979323846 * l1 * (2.0 * s" << K << " * c" << K
            << ")) * (il2_" << K << " * (ip" << K << " * ip" << K << "));\n";
        }
        nq = 2;
        break;
      case GPB_OP_LIN:
        o << "    double k0_" << K << " = 0.0;\n";
        for (int d = 0; d < dim; ++d)
          o << "    k0_" << K << " += (" << xi(d) << " - " << H(a + d) << ") * (" << xj(d) << " - " << H(a + d) << ");\n";
        if (with_grad)
          for (int d = 0; d < dim; ++d)
            o << "    double d" << K << "_" << d << " = 2.0 * " << H(a + d) << " - " << xi(d) << " - " << xj(d) << ";\n";
        nq = dim;
        break;
      case GPB_OP_MAT32:
        o << "    const double il" << K << " = fabs(" << IH(a) << ");\n";
        o << "    const double f" << K << " = (sqrt(3.0) * l1) * il" << K << ";\n";
        o << "    const double e" << K << " = exp(-f" << K << ");\n";
        o << "    const double k0_" << K << " = (1.0 + f" << K << ") * e" << K << ";\n";
        if (with_grad)
          o << "    double d" << K << "_0 = (f" << K << " * f" << K << " * e" << K << " * il" << K << ") * (" << H(a)
            << " < 0.0 ? -1.0 : 1.0);\n";
        nq = 1;
        break;
      case GPB_OP_MAT52:
        o << "    const double il" << K << " = fabs(" << IH(a) << ");\n";
        o << "    const double f" << K << " = (sqrt(5.0) * l1) * il" << K << ";\n";
        o << "    const double th" << K << " = (5.0 * (l1 * l1)) * ((il" << K << " * il" << K << ") * (1.0 / 3.0));\n";
        o << "    const double e" << K << " = exp(-f" << K << ");\n";
        o << "    const double k0_" << K << " = (1.0 + f" << K << " + th" << K << ") * e" << K << ";\n";
        if (with_grad)
          o << "    double d" << K << "_0 = (f" << K << " * f" << K << " * (1.0 + f" << K << ") * (1.0 / 3.0) * e" << K << " * il" << K
            << ") * (" << H(a) << " < 0.0 ? -1.0 : 1.0);\n";
        nq = 1;
        break;
      case GPB_OP_WN:
        o << "    const double k0_" << K << " = (gi == gj) ? 1.0 : 0.0;\n";
        nq = 0;
        break;
      case GPB_OP_L2:
        o << "    const double k0_" << K << " = sqrt(r2);\n";
        nq = 0;
        break;
      case GPB_OP_L1:
        o << "    const double k0_" << K << " = l1;\n";
        nq = 0;
        break;
      case GPB_OP_SE_ARD:
        o << "    double ra" << K << " = 0.0;\n";
        for (int d = 0; d < dim; ++d) {
          o << "    const double t" << K << "_" << d << " = (" << xi(d) << " - " << xj(d) << ") * " << IH(a + d) << ";\n";
          o << "    ra" << K << " += t" << K << "_" << d << " * t" << K << "_" << d << ";\n";
        }
        o << "    const double k0_" << K << " = exp(-0.5 * ra" << K << ");\n";
        if (with_grad)
          for (int d = 0; d < dim; ++d)
            o << "    double d" << K << "_" << d << " = k0_" << K << " * (t" << K << "_" << d << " * t" << K << "_" << d << ") * "
              << IH(a + d) << ";\n";
        nq = dim;
        break;
      default: break;
    }
    if (scaled && nd.kind != GPB_OP_WN && nd.kind != GPB_OP_L2 && nd.kind != GPB_OP_L1) {
      o << "    const double sg" << K << " = " << H(a + nq) << ";\n";
      if (with_grad) {
        for (int q = 0; q < nq; ++q) o << "    d" << K << "_" << q << " *= sg" << K << ";\n";
        o << "    const double d" << K << "_" << nq << " = k0_" << K << ";\n";
      }
      o << "    const double v" << K << " = sg" << K << " * k0_" << K << ";\n";
    } else {
      o << "    const double v" << K << " = k0_" << K << ";\n";
    }
  }

  // window weight of child i of a k-child change-point node (program.cuh gpb_cp_weight): w<K>, dp<K>, dc<K>
  void cpw(int k, const Node& nd) {
    const std::string K = std::to_string(k);
    const int a = nd.off, i = nd.child, kk = nd.nchild;
    o << "    // node " << k << ": change-point window of child " << i << " of " << kk << "\n";
    std::string wp = "1.0", wc = "1.0";
    if (i > 0) {
      cp_s(K + "pa", xi(0), H(a + i - 1));
      cp_s(K + "pb", xj(0), H(a + i - 1));
      o << "    const double wp" << K << " = (1.0 - S" << K << "pa) * (1.0 - S" << K << "pb);\n";
      wp = "wp" + K;
    }
    if (i < kk - 1) {
      cp_s(K + "ca", xi(0), H(a + i));
      cp_s(K + "cb", xj(0), H(a + i));
      o << "    const double wc" << K << " = S" << K << "ca * S" << K << "cb;\n";
      wc = "wc" + K;
    }
    o << "    const double w" << K << " = " << wp << " * " << wc << ";\n";
    if (with_grad && cp_mode != GPB_CP_INDICATOR) {
      if (i > 0)
        o << "    const double dp" << K << " = (-(D" << K << "pa * (1.0 - S" << K << "pb) + (1.0 - S" << K << "pa) * D" << K << "pb)) * "
          << wc << ";\n";
      if (i < kk - 1)
        o << "    const double dc" << K << " = (D" << K << "ca * S" << K << "cb + S" << K << "ca * D" << K << "cb) * " << wp << ";\n";
    }
    o << "    const double v" << K << " = v" << nd.a << " * w" << K << ";\n";
  }
};

bool build_nodes(const int32_t* code, int n_ops, int dim, std::vector<Node>& nodes, int& root, bool& r2, bool& l1, std::string& err) {
  std::vector<int> stack;
  for (int pc = 0; pc < n_ops; ++pc) {
    const int32_t* w = code + pc * GPB_OP_WORDS;
    Node nd;
    nd.kind = w[0];
    if (nd.kind < GPB_OP_ADD2) {
      nd.off = w[1]; nd.flags = w[2];
      nd.nq = gpb_leaf_nhp(nd.kind, nd.flags, dim);
      if (nd.kind == GPB_OP_SE || nd.kind == GPB_OP_L2) r2 = true;
      if (nd.kind == GPB_OP_PER || nd.kind == GPB_OP_MAT32 || nd.kind == GPB_OP_MAT52 || nd.kind == GPB_OP_L1) l1 = true;
      nodes.push_back(nd);
      stack.push_back((int)nodes.size() - 1);
    } else if (nd.kind == GPB_OP_CPW) {
      if (stack.empty()) { err = "stack underflow"; return false; }
      nd.a = stack.back(); nd.off = w[1]; nd.child = w[2]; nd.nchild = w[3];
      nodes.push_back(nd);
      stack.back() = (int)nodes.size() - 1;
    } else if (nd.kind == GPB_OP_ADD2 || nd.kind == GPB_OP_MUL2) {
      if (stack.size() < 2) { err = "stack underflow"; return false; }
      nd.b = stack.back(); stack.pop_back();
      nd.a = stack.back();
      nodes.push_back(nd);
      stack.back() = (int)nodes.size() - 1;
    } else {
      err = "unknown opcode";
      return false;
    }
  }
  if (stack.size() != 1) { err = "program does not leave exactly one value"; return false; }
  root = stack.back();
  return true;
}

void emit_forward(Gen& g) {
  if (g.need_r2) {
    g.o << "    double r2 = 0.0;\n";
    for (int d = 0; d < g.dim; ++d)
      g.o << "    { const double t = " << g.xi(d) << " - " << g.xj(d) << "; r2 += t * t; }\n";
  }
  if (g.need_l1) {
    g.o << "    double l1 = 0.0;\n";
    for (int d = 0; d < g.dim; ++d) g.o << "    l1 += fabs(" << g.xi(d) << " - " << g.xj(d) << ");\n";
  }
  for (int k = 0; k < (int)g.nodes.size(); ++k) {
    const Node& nd = g.nodes[k];
    if (nd.kind < GPB_OP_ADD2) g.leaf(k, nd);
    else if (nd.kind == GPB_OP_CPW) g.cpw(k, nd);
    else g.o << "    const double v" << k << " = v" << nd.a << (nd.kind == GPB_OP_ADD2 ? " + " : " * ") << "v" << nd.b << ";\n";
  }
}

}  // namespace

int jit_generate(const int32_t* code, int n_ops, int dim, int cp_mode, int n_hp, std::string& src, std::string& err) {
  std::vector<Node> nodes;
  int root = -1;
  bool r2 = false, l1 = false;
  if (!build_nodes(code, n_ops, dim, nodes, root, r2, l1, err)) return 1;
  std::ostringstream out;
  out << "// generated by libgpb (jit.cu) from a postfix kernel program of " << n_ops << " ops\n";
  out << "#define GPB_JIT_SPECIALISED 1\n";
  out << kDeviceAbiSrc << "\n" << kMathSrc << "\n" << kSpecKernelsSrc << "\n";
  out << "namespace gpb {\nstruct Prog {\n  static constexpr int N_HP = " << n_hp << ", DIM = " << dim << ";\n";
  const char* sig = "const double* __restrict__ h, const double* __restrict__ ih, const double* __restrict__ xi, "
                    "const double* __restrict__ xj, int gi, int gj";
  {
    Gen g; g.dim = dim; g.cp_mode = cp_mode; g.with_grad = false; g.nodes = nodes; g.need_r2 = r2; g.need_l1 = l1;
    out << "  static __device__ __forceinline__ double value(" << sig << ") {\n";
    emit_forward(g);
    out << g.o.str() << "    return v" << root << ";\n  }\n";
  }
  {
    Gen g; g.dim = dim; g.cp_mode = cp_mode; g.with_grad = true; g.nodes = nodes; g.need_r2 = r2; g.need_l1 = l1;
    out << "  static __device__ __forceinline__ void grad(" << sig << ", double w, double (&g)[N_HP + 1]) {\n";
    emit_forward(g);
    out << g.o.str();
    // reverse sweep: every node has exactly one consumer, so its adjoint is a single assignment
    std::ostringstream b;
    b << "    const double a" << root << " = w;\n";
    for (int k = (int)nodes.size() - 1; k >= 0; --k) {
      const Node& nd = nodes[k];
      if (nd.kind < GPB_OP_ADD2) {
        for (int q = 0; q < nd.nq; ++q) b << "    g[" << nd.off + q << "] += a" << k << " * d" << k << "_" << q << ";\n";
      } else if (nd.kind == GPB_OP_CPW) {
        if (cp_mode != GPB_CP_INDICATOR) {
          if (nd.child > 0) b << "    g[" << nd.off + nd.child - 1 << "] += a" << k << " * v" << nd.a << " * dp" << k << ";\n";
          if (nd.child < nd.nchild - 1) b << "    g[" << nd.off + nd.child << "] += a" << k << " * v" << nd.a << " * dc" << k << ";\n";
        }
        b << "    const double a" << nd.a << " = a" << k << " * w" << k << ";\n";
      } else if (nd.kind == GPB_OP_MUL2) {
        b << "    const double a" << nd.a << " = a" << k << " * v" << nd.b << ";\n";
        b << "    const double a" << nd.b << " = a" << k << " * v" << nd.a << ";\n";
      } else {
        b << "    const double a" << nd.a << " = a" << k << ";\n";
        b << "    const double a" << nd.b << " = a" << k << ";\n";
      }
    }
    out << b.str() << "  }\n";
  }
  out << "};\n}  // namespace gpb\n";
  out << "extern \"C\" __global__ void __launch_bounds__(256, 1) gpb_spec_assemble(const GpbMat* __restrict__ mats, "
         "const int* __restrict__ which) {\n  gpb::assemble_spec_body<gpb::Prog>(mats, which);\n}\n";
  out << "extern \"C\" __global__ void __launch_bounds__(256, 1) gpb_spec_grad(const GpbMat* __restrict__ mats, "
         "const int* __restrict__ which) {\n  gpb::grad_spec_body<gpb::Prog>(mats, which);\n}\n";
  src = out.str();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// NVRTC and the driver API, bound at run time
// ---------------------------------------------------------------------------------------------------------------
namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
  void* handle = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUstream_st* CUstream;
struct DriverApi {
  void* handle = nullptr;
  int (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  int (*ModuleUnload)(CUmodule) = nullptr;
  int (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  int (*GetErrorString)(int, const char**) = nullptr;
};

NvrtcApi g_rtc;
DriverApi g_drv;
std::mutex g_jit_mu;
int g_jit_state = -1;          // -1 unknown, 0 unavailable, 1 ready
std::string g_jit_why;
std::string g_nvrtc_path;      // optional explicit path (gpb_jit_set_nvrtc_path)

bool load_nvrtc(std::string& why) {
  if (g_rtc.handle) return true;
  const char* names[] = {nullptr, "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                         "/usr/local/cuda/lib64/libnvrtc.so"};
  void* h = nullptr;
  std::string tried;
  if (!g_nvrtc_path.empty()) names[0] = g_nvrtc_path.c_str();
  else if (const char* e = getenv("GPB_NVRTC_PATH")) names[0] = e;
  for (const char* nm : names) {
    if (!nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
    tried += std::string(nm) + " ";
  }
  if (!h) { why = "cannot load libnvrtc (tried " + tried + ")"; return false; }
  NvrtcApi a;
  a.handle = h;
#define GPB_SYM(field, name) *(void**)(&a.field) = dlsym(h, name); if (!a.field) { why = "libnvrtc lacks " name; return false; }
  GPB_SYM(CreateProgram, "nvrtcCreateProgram")
  GPB_SYM(CompileProgram, "nvrtcCompileProgram")
  GPB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
  GPB_SYM(GetCUBIN, "nvrtcGetCUBIN")
  GPB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
  GPB_SYM(GetProgramLog, "nvrtcGetProgramLog")
  GPB_SYM(DestroyProgram, "nvrtcDestroyProgram")
  GPB_SYM(GetErrorString, "nvrtcGetErrorString")
#undef GPB_SYM
  g_rtc = a;
  return true;
}

bool load_driver(std::string& why) {
  if (g_drv.handle) return true;
  void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { why = std::string("cannot load libcuda.so.1: ") + dlerror(); return false; }
  DriverApi a;
  a.handle = h;
#define GPB_SYM(field, name) *(void**)(&a.field) = dlsym(h, name); if (!a.field) { why = "libcuda lacks " name; return false; }
  GPB_SYM(ModuleLoadData, "cuModuleLoadData")
  GPB_SYM(ModuleUnload, "cuModuleUnload")
  GPB_SYM(ModuleGetFunction, "cuModuleGetFunction")
  GPB_SYM(LaunchKernel, "cuLaunchKernel")
  GPB_SYM(GetErrorString, "cuGetErrorString")
#undef GPB_SYM
  g_drv = a;
  return true;
}

}  // namespace

void jit_set_nvrtc_path(const char* path) {
  std::lock_guard<std::mutex> lk(g_jit_mu);
  g_nvrtc_path = path ? path : "";
  if (g_jit_state == 0) g_jit_state = -1;   // try again with the new path
}

int jit_compile(const std::string& src, const char* arch, std::vector<char>& cubin, std::string& log) {
  {
    std::lock_guard<std::mutex> lk(g_jit_mu);
    std::string why;
    if (!load_nvrtc(why)) { log = why; return 1; }
  }
  nvrtcProgram prog = nullptr;
  int rc = g_rtc.CreateProgram(&prog, src.c_str(), "gpb_spec.cu", 0, nullptr, nullptr);
  if (rc) { log = std::string("nvrtcCreateProgram: ") + g_rtc.GetErrorString(rc); return 1; }
  const std::string archopt = std::string("--gpu-architecture=") + arch;
  // -lineinfo: the source page of ncu maps to the generated text; fmad stays on as in the ahead-of-time kernels
  const char* opts[] = {archopt.c_str(), "--std=c++17", "-lineinfo", "--fmad=true"};
  rc = g_rtc.CompileProgram(prog, 4, opts);
  size_t lsz = 0;
  g_rtc.GetProgramLogSize(prog, &lsz);
  if (lsz > 1) { log.resize(lsz); g_rtc.GetProgramLog(prog, &log[0]); }
  if (rc) {
    log = std::string("nvrtcCompileProgram: ") + g_rtc.GetErrorString(rc) + "\n" + log;
    g_rtc.DestroyProgram(&prog);
    return 1;
  }
  size_t sz = 0;
  rc = g_rtc.GetCUBINSize(prog, &sz);
  if (rc || sz == 0) { log = "nvrtcGetCUBINSize failed"; g_rtc.DestroyProgram(&prog); return 1; }
  cubin.resize(sz);
  rc = g_rtc.GetCUBIN(prog, cubin.data());
  g_rtc.DestroyProgram(&prog);
  if (rc) { log = "nvrtcGetCUBIN failed"; return 1; }
  return 0;
}

bool jit_enabled(std::string* why) {
  std::lock_guard<std::mutex> lk(g_jit_mu);
  if (g_jit_state < 0) {
    const char* e = getenv("GPB_JIT");
    if (e && e[0] == '0') { g_jit_state = 0; g_jit_why = "disabled by GPB_JIT=0"; }
    else if (!load_nvrtc(g_jit_why) || !load_driver(g_jit_why)) g_jit_state = 0;
    else g_jit_state = 1;
  }
  if (why) *why = g_jit_why;
  return g_jit_state == 1;
}

static std::string drv_err(int rc) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(rc, &s);
  return s ? s : "unknown driver error";
}

int jit_build(const int32_t* code, int n_ops, int dim, int cp_mode, int n_hp, JitKernels& out, std::string& err) {
  out = JitKernels();
  if (!jit_enabled(&err)) return 1;
  std::string src;
  if (jit_generate(code, n_ops, dim, cp_mode, n_hp, src, err)) return 1;
  int dev = 0, major = 0, minor = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  char arch[32];
  // family-specific target as in the ahead-of-time build (sm_100a on B200)
  snprintf(arch, sizeof arch, "sm_%d%d%s", major, minor, major >= 9 ? "a" : "");
  std::vector<char> cubin;
  std::string log;
  if (jit_compile(src, arch, cubin, log)) { err = log; return 1; }
  cudaFree(0);   // the primary context must be current for the driver calls
  CUmodule mod = nullptr;
  int rc = g_drv.ModuleLoadData(&mod, cubin.data());
  if (rc) { err = "cuModuleLoadData: " + drv_err(rc); return 1; }
  CUfunction fa = nullptr, fg = nullptr;
  rc = g_drv.ModuleGetFunction(&fa, mod, "gpb_spec_assemble");
  if (!rc) rc = g_drv.ModuleGetFunction(&fg, mod, "gpb_spec_grad");
  if (rc) { err = "cuModuleGetFunction: " + drv_err(rc); g_drv.ModuleUnload(mod); return 1; }
  out.module = mod; out.assemble = fa; out.grad = fg;
  return 0;
}

void jit_release(JitKernels& k) {
  if (k.module && g_drv.ModuleUnload) g_drv.ModuleUnload((CUmodule)k.module);
  k = JitKernels();
}

cudaError_t jit_launch(void* fn, unsigned gx, unsigned gz, const GpbMat* mats, const int* which, cudaStream_t s) {
  void* args[2] = {(void*)&mats, (void*)&which};
  const int rc = g_drv.LaunchKernel((CUfunction)fn, gx, 1, gz, 256, 1, 1, 0, (CUstream)s, args, nullptr);
  ++g_launches;
  return rc == 0 ? cudaSuccess : cudaErrorLaunchFailure;
}

}  // namespace gpb
