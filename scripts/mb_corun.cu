// Can an ordinary kernel run beside a grid of 16-CTA clusters (on the SMs the clusters leave empty)?
// A: n clusters x C CTAs, 200 KB dynamic smem, spins ~3 ms.  B: 32 small CTAs on a second stream, records its start time.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void kA(unsigned long long* tA, unsigned long long dur) {
  extern __shared__ unsigned char sm[];
  const unsigned long long t0 = gt();
  if (threadIdx.x == 0) { sm[0] = 1; tA[2 * blockIdx.x] = t0; }
  while (gt() - t0 < dur) __nanosleep(1000);
  if (threadIdx.x == 0) tA[2 * blockIdx.x + 1] = gt();
}
__global__ void kB(unsigned long long* tB) {
  unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) { tB[2 * blockIdx.x] = gt(); tB[2 * blockIdx.x + 1] = smid; }
}
int main() {
  unsigned long long *tA, *tB;
  cudaMallocManaged(&tA, 4096 * 8); cudaMallocManaged(&tB, 4096 * 8);
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(kA, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int C : {1, 8, 16}) {
    for (int smemk : {200, 8}) {
      cudaLaunchConfig_t lc = {};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1; lc.blockDim = dim3(256); lc.dynamicSmemBytes = smemk * 1024; lc.stream = s1;
      int mc = 0; lc.gridDim = dim3(C);
      cudaOccupancyMaxActiveClusters(&mc, kA, &lc);
      int ncl = (C == 1) ? 112 : (smemk == 200 ? mc : 112 / C);
      lc.gridDim = dim3(ncl * C);
      cudaLaunchKernelEx(&lc, kA, tA, 3000000ull);
      kB<<<32, 128, 0, s2>>>(tB);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long a0 = ~0ull, a1 = 0, b0 = ~0ull, b1 = 0;
      for (int i = 0; i < ncl * C; ++i) { if (tA[2 * i] < a0) a0 = tA[2 * i]; if (tA[2 * i + 1] > a1) a1 = tA[2 * i + 1]; }
      for (int i = 0; i < 32; ++i) { if (tB[2 * i] < b0) b0 = tB[2 * i]; if (tB[2 * i] > b1) b1 = tB[2 * i]; }
      printf("C=%2d smem=%3dK maxclusters=%d grid=%d (%s): A ran %.3f ms; B started %.3f .. %.3f ms after A's start\n", C, smemk, mc, ncl * C,
             cudaGetErrorString(e), (a1 - a0) * 1e-6, ((double)b0 - (double)a0) * 1e-6, ((double)b1 - (double)a0) * 1e-6);
    }
  }
  return 0;
}
