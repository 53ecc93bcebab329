// Where do the CTAs of a 4-per-SM persistent grid land, and which hardware warp slots do their warps get?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 4) probe(int* smid, int* wslot) {
  extern __shared__ unsigned char sm[];
  unsigned s, w;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
  if ((threadIdx.x & 31) == 0) {
    wslot[blockIdx.x * 4 + (threadIdx.x >> 5)] = (int)w;
    if (threadIdx.x == 0) smid[blockIdx.x] = (int)s;
  }
  sm[threadIdx.x] = (unsigned char)s;
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
}
int main() {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int grid = nsm * 4, *smid, *ws;
  cudaMallocManaged(&smid, grid * 4); cudaMallocManaged(&ws, grid * 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 55296);
  probe<<<grid, 128, 55296>>>(smid, ws);
  cudaDeviceSynchronize();
  int same = 0, mod = 0;
  for (int b = 0; b < grid; ++b) { same += smid[b] == smid[b % nsm]; }
  printf("nsm %d grid %d: blocks b and b %% nsm on the same SM: %d of %d\n", nsm, grid, same, grid);
  for (int b = 0; b < 8; ++b) printf("block %d smid %d warpslots %d %d %d %d | block %d smid %d slots %d %d %d %d\n", b, smid[b], ws[4*b], ws[4*b+1], ws[4*b+2], ws[4*b+3], b + nsm, smid[b+nsm], ws[4*(b+nsm)], ws[4*(b+nsm)+1], ws[4*(b+nsm)+2], ws[4*(b+nsm)+3]);
  // per SM: which blocks
  for (int s = 0; s < 3; ++s) { printf("sm %d:", s); for (int b = 0; b < grid; ++b) if (smid[b] == s) printf(" %d(w%d,%d,%d,%d)", b, ws[4*b], ws[4*b+1], ws[4*b+2], ws[4*b+3]); printf("\n"); }
  return 0;
}
