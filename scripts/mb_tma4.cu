// mb_tma4.cu — is a region that was pulled into L2 with cp.async.bulk.prefetch.L2 served at L2 speed to the K/V ring's TMA copies?
// 112 CTAs x 8 warps stream 64 KB each (58.7 MB in all) through 16 KB slots in 8 KB copies (the ring of mb_tma3.cu):
// cold (after an L2 flush), hot (streamed just before), prefetched (flush, then 32 CTAs bulk-prefetch the region in 64 KB pieces).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(void* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(void* b, uint32_t par) {
  asm volatile("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}\n" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, void* bar, int evict_first) {
  if (evict_first) {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)), "l"(pol) : "memory");
  } else {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
  }
}
constexpr int SLOT = 16384, PER_WARP = 65536;
__global__ void __launch_bounds__(256, 1) k_stream(const unsigned char* base, int evict_first) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 8 * SLOT);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0) return;
  const unsigned char* src = base + ((size_t)blockIdx.x * 8 + warp) * PER_WARP;
  unsigned char* dst = sm + warp * SLOT;
  const int n = PER_WARP / SLOT;
  for (int p = 0; p < 2; ++p) { mbar_expect_tx(&bars[warp * 2 + p], 8192); bulk_load(dst + p * 8192, src + p * 8192, 8192, &bars[warp * 2 + p], evict_first); }
  for (int i = 1; i <= n; ++i)
    for (int p = 0; p < 2; ++p) {
      mbar_wait(&bars[warp * 2 + p], (i - 1) & 1);
      if (i < n) { mbar_expect_tx(&bars[warp * 2 + p], 8192); bulk_load(dst + p * 8192, src + (size_t)i * SLOT + p * 8192, 8192, &bars[warp * 2 + p], evict_first); }
    }
}
__global__ void k_pf(const unsigned char* p, size_t bytes, uint32_t piece) {
  const size_t n = bytes / piece;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, GW = (gridDim.x * blockDim.x) >> 5;
  if ((threadIdx.x & 31) == 0)
    for (size_t i = gw; i < n; i += GW) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + i * piece), "r"(piece) : "memory");
}
__global__ void k_flush(uint4* p, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 1, 2, 3);
}
__global__ void k_spin(long long clk) { const long long t = clock64(); while (clock64() - t < clk) {} }
int main() {
  const int CTAS = 112;
  const size_t REG = (size_t)CTAS * 8 * PER_WARP, FL = 512ull << 20;
  unsigned char *buf, *fl;
  CK(cudaMalloc(&buf, REG)); CK(cudaMalloc(&fl, FL)); CK(cudaMemset(buf, 1, REG));
  CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SLOT + 512));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto stream_us = [&](int ef) {
    CK(cudaEventRecord(e0)); k_stream<<<CTAS, 256, 8 * SLOT + 512>>>(buf, ef); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1e3f;
  };
  auto flush = [&]() { k_flush<<<296, 256>>>((uint4*)fl, FL / 16); CK(cudaDeviceSynchronize()); };
  stream_us(0);
  for (int ef = 0; ef < 2; ++ef) {
    for (int rep = 0; rep < 2; ++rep) {
      flush(); const float cold = stream_us(ef); const float hot = stream_us(ef); const float hot2 = stream_us(ef);
      printf("evict_first %d: %.1f MB  cold %.1f us (%.0f GB/s)  again %.1f us (%.0f GB/s)  third %.1f us\n", ef, REG * 1e-6, cold, REG / cold * 1e-3, hot, REG / hot * 1e-3, hot2);
      for (uint32_t piece : {8192u, 65536u}) {
        flush();
        k_pf<<<32, 128>>>(buf, REG, piece); k_spin<<<1, 32>>>(60000); CK(cudaDeviceSynchronize());
        const float pf = stream_us(ef);
        printf("   after bulk prefetch (%u B pieces, 30 us later): %.1f us (%.0f GB/s)\n", piece, pf, REG / pf * 1e-3);
      }
    }
  }
  return 0;
}
