// Does cp.async.bulk.prefetch.L2 bring a whole piece into L2, whatever its size?  Prefetch a region with pieces of S bytes,
// then time a streaming read of it (112 CTAs, 128-bit loads) against a cold read and a warm (just read) one.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_pf(const unsigned char* p, size_t bytes, uint32_t piece) {
  const size_t n = bytes / piece;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, GW = (gridDim.x * blockDim.x) >> 5;
  if ((threadIdx.x & 31) == 0)
    for (size_t i = gw; i < n; i += GW) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + i * piece), "r"(piece) : "memory");
}
__global__ void k_read(const uint4* p, size_t n16, unsigned* out) {
  unsigned acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *out = acc;
}
__global__ void k_flush(uint4* p, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 1, 2, 3);
}
int main() {
  const size_t REG = 32ull << 20, FL = 512ull << 20;
  unsigned char *buf, *fl; unsigned* out;
  cudaMalloc(&buf, 256ull << 20); cudaMalloc(&fl, FL); cudaMalloc(&out, 4);
  cudaMemset(buf, 1, 256ull << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto read_ms = [&](const unsigned char* p, size_t bytes) {
    cudaEventRecord(e0); k_read<<<112, 256>>>((const uint4*)p, bytes / 16, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
  };
  auto flush = [&]() { k_flush<<<296, 256>>>((uint4*)fl, FL / 16); cudaDeviceSynchronize(); };
  for (size_t region : {REG, 2 * REG}) {
    flush(); float cold = read_ms(buf, region); float warm = read_ms(buf, region);
    printf("region %zu MB: cold %.1f us (%.0f GB/s), warm %.1f us (%.0f GB/s)\n", region >> 20, cold * 1e3, region / cold * 1e-6, warm * 1e3, region / warm * 1e-6);
    for (uint32_t piece : {4096u, 16384u, 65536u, 262144u, 1048576u}) {
      for (int ctas : {32}) {
        flush();
        cudaEventRecord(e0); k_pf<<<ctas, 128>>>(buf, region, piece); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float pf; cudaEventElapsedTime(&pf, e0, e1);
        cudaDeviceSynchronize();
        // give the prefetches time to land: a dummy kernel of ~100 us
        k_read<<<16, 256>>>((const uint4*)(buf + (128ull << 20)), (8ull << 20) / 16, out); cudaDeviceSynchronize();
        float r = read_ms(buf, region);
        printf("  piece %7u B, %d CTAs: prefetch kernel %.1f us, read after prefetch %.1f us (%.0f GB/s)\n", piece, ctas, pf * 1e3, r * 1e3, region / r * 1e-6);
      }
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
