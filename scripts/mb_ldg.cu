// mb_ldg.cu — does cp.async.bulk.prefetch.L2 turn later LDG.128 reads into L2 hits?  One warp per CTA reads 8 KB batches
// (16 x LDG.128 per lane) from a large (HBM) region; variants: no prefetch, bulk prefetch D batches ahead, per-line prefetch.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint4 ldw(const void* p) {
  uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
__global__ void k(const unsigned char* base, size_t region, int mode, int ahead, int n, long long* out, unsigned* sink) {
  const int lane = threadIdx.x;
  const unsigned char* b = base + (size_t)blockIdx.x * region;
  unsigned acc = 0;
  const int BATCH = 8192;
  if (mode != 0 && lane == 0) {
    for (int i = 0; i < ahead; ++i) {
      const unsigned char* p = b + ((size_t)i * BATCH) % region;
      if (mode == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(BATCH) : "memory");
    }
  }
  if (mode == 2) for (int i = 0; i < ahead; ++i) for (int j = lane; j < BATCH / 128; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + ((size_t)i * BATCH) % region + j * 128));
  __syncwarp();
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    const unsigned char* pn = b + ((size_t)(i + ahead) * BATCH) % region;
    if (mode == 1 && lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pn), "r"(BATCH) : "memory");
    if (mode == 2) for (int j = lane; j < BATCH / 128; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + j * 128));
    const uint4* p = reinterpret_cast<const uint4*>(b + ((size_t)i * BATCH) % region) + lane;
    uint4 f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = ldw(p + j * 32);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= f[j].x ^ f[j].y ^ f[j].z ^ f[j].w;
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x1234567u) sink[0] = acc;
}
int main() {
  const size_t region = 32 << 20;
  unsigned char* buf; long long* out; unsigned* sink;
  CK(cudaMalloc(&buf, region * 128)); CK(cudaMalloc(&out, 148 * 8)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(buf, 1, region * 128));
  const char* names[] = {"no prefetch", "cp.async.bulk.prefetch.L2", "prefetch.global.L2 per line"};
  for (int ctas : {1, 112})
    for (int mode = 0; mode < 3; ++mode)
      for (int ahead : {4, 16, 64}) {
        if (mode == 0 && ahead != 4) continue;
        const int n = 2000;
        k<<<ctas, 32>>>(buf, region, mode, ahead, n, out, sink);
        CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost));
        double cyc = 0; for (int i = 0; i < ctas; ++i) cyc += h[i]; cyc /= ctas;
        printf("%3d CTAs x 1 warp, %-30s ahead %2d: %6.0f cyc per 8 KB batch (16 LDG.128/lane)\n", ctas, names[mode], ahead, cyc / n);
      }
  // L2-hot reference: tiny region
  for (int ctas : {1, 112}) {
    k<<<ctas, 32>>>(buf, 64 << 10, 0, 4, 2000, out, sink);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < ctas; ++i) cyc += h[i]; cyc /= ctas;
    printf("%3d CTAs x 1 warp, L2-hot region                          : %6.0f cyc per 8 KB batch\n", ctas, cyc / 2000);
  }
  return 0;
}
