// Run-time specialisation of assembly / gradient kernels per kernel program (jit.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace gpb {

struct JitKernels {
  void* module = nullptr;     // CUmodule
  void* assemble = nullptr;   // CUfunction gpb_spec_assemble(const GpbMat*, const int* which)
  void* grad = nullptr;       // CUfunction gpb_spec_grad(const GpbMat*, const int* which)
};

// host only (no GPU, no NVRTC needed): the CUDA source specialised for the program
int jit_generate(const int32_t* code, int n_ops, int dim, int cp_mode, int n_hp, std::string& src, std::string& err);
// host only (needs libnvrtc, no GPU): source -> CUBIN for `arch` (e.g. "sm_100a"); log receives warnings / errors
int jit_compile(const std::string& src, const char* arch, std::vector<char>& cubin, std::string& log);
// NVRTC + driver API loadable and not switched off (GPB_JIT=0)
bool jit_enabled(std::string* why);
void jit_set_nvrtc_path(const char* path);
// generate + compile for the current device + load; non-zero: err says why (the caller keeps the interpreter)
int jit_build(const int32_t* code, int n_ops, int dim, int cp_mode, int n_hp, JitKernels& out, std::string& err);
void jit_release(JitKernels& k);
// grid (gx, 1, gz) x 256 threads on stream s
cudaError_t jit_launch(void* fn, unsigned gx, unsigned gz, const GpbMat* mats, const int* which, cudaStream_t s);

}  // namespace gpb
