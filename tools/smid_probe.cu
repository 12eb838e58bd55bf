// Developer probe: which %smid values a full-machine launch sees (are they contiguous 0 .. SMs-1?) and how the block
// scheduler places the CTAs of a grid that is smaller than one wave.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/smid_probe tools/smid_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned* smid_of_block, unsigned* nsmid) {
  unsigned s, n;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
  if (threadIdx.x == 0) { smid_of_block[blockIdx.x] = s; *nsmid = n; }
  // stay resident for a while so that the whole grid is placed before anything retires
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) { }
}
int main() {
  unsigned *d, *dn;
  const int grids[3] = {148, 296, 280};
  cudaMalloc(&d, 4096 * 4); cudaMalloc(&dn, 4);
  for (int gi = 0; gi < 3; ++gi) {
    const int g = grids[gi];
    // 128 threads, 92 KB dynamic smem: two CTAs per SM like the bulk update
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 92 * 1024);
    probe<<<g, 128, 92 * 1024>>>(d, dn);
    unsigned h[4096], n;
    cudaMemcpy(h, d, g * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&n, dn, 4, cudaMemcpyDeviceToHost);
    int cnt[1024] = {0}; unsigned mx = 0;
    for (int i = 0; i < g; ++i) { cnt[h[i]]++; if (h[i] > mx) mx = h[i]; }
    int used = 0, c1 = 0, c2 = 0;
    for (unsigned s = 0; s <= mx; ++s) { if (cnt[s]) ++used; if (cnt[s] == 1) ++c1; if (cnt[s] == 2) ++c2; }
    printf("grid %d: nsmid %u, max smid %u, SMs used %d, with 1 CTA %d, with 2 CTAs %d\n", g, n, mx, used, c1, c2);
    if (gi == 0) { printf("first 32 blocks -> smid:"); for (int i = 0; i < 32; ++i) printf(" %u", h[i]); printf("\n"); }
  }
  return 0;
}
