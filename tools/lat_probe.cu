// Dependent-chain latencies (cycles per op) of the FP64 / shuffle / MUFU operations on the pivot chain of the
// diagonal-block kernel.  Developer probe, one warp.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double my_rsqrt(double d) {
  // hardware seed (MUFU.RSQ64H through the single-precision path) + two Newton steps
  float f = (float)d;
  float y0 = rsqrtf(f);
  double y = (double)y0;
  double h = 0.5 * d;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}
__global__ void probe(int iters, double seed, long long* out, double* sink) {
  double x = seed + threadIdx.x * 1e-3;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64(); out[0] = t1 - t0;
  // DMUL chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = x * 1.0000001;
  t1 = clock64(); out[1] = t1 - t0;
  // rsqrt(double) chain
  x = fabs(x) + 1.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = rsqrt(x) + 1.0;
  t1 = clock64(); out[2] = t1 - t0;
  // custom rsqrt chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = my_rsqrt(x) + 1.0;
  t1 = clock64(); out[3] = t1 - t0;
  // 64-bit shuffle chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
  t1 = clock64(); out[4] = t1 - t0;
  // sqrt + div chain
  x = fabs(x) + 1.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = 1.0 / sqrt(x) + 1.0;
  t1 = clock64(); out[5] = t1 - t0;
  // shared-memory round trip
  __shared__ double sm[64];
  sm[threadIdx.x] = x;
  __syncwarp();
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = sm[(threadIdx.x + (int)x) & 31] + 1e-9; }
  t1 = clock64(); out[6] = t1 - t0;
  // double add after the dadd (rsqrt loop has +1.0): DADD chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = x + 1.0000001;
  t1 = clock64(); out[7] = t1 - t0;
  sink[threadIdx.x] = x;
}
int main() {
  long long* out; double* sink;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 512);
  const int iters = 4096;
  probe<<<1, 32>>>(iters, 1.5, out, sink);
  probe<<<1, 32>>>(iters, 1.5, out, sink);
  long long h[8];
  cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
  const char* names[8] = {"dfma", "dmul", "rsqrt(double)+dadd", "seed+2 newton+dadd", "shfl64", "1/sqrt+dadd", "lds+dadd(+cvt)", "dadd"};
  for (int i = 0; i < 8; ++i) printf("%-22s %.1f cycles/op\n", names[i], (double)h[i] / iters);
  return 0;
}
