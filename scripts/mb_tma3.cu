// mb_tma3.cu — the cluster-stream kernel's K/V ring without any compute: 112 (or 148) CTAs x 8 warps, every warp owns one
// 16 KB slot split into `parts` copies (1 x 16 KB, 2 x 8 KB, 4 x 4 KB); lane 0 re-issues a part as soon as it has landed.
// Whole-GPU bytes/s from HBM (a 4 GB stream) and from L2 (a 48 MB region re-read): is 8 KB granularity with 16 KB per warp
// in flight enough to saturate HBM, and how much faster is an L2-resident stream?
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
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
constexpr int SLOT = 16384;
// every warp streams `n` slots' worth (16 KB each) from its own region; region bytes = wrap (power of two per warp)
__global__ void __launch_bounds__(256, 1) k(const unsigned char* base, size_t warp_stride, size_t wrap, int parts, int n, int delay) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 8 * SLOT);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0) return;
  const unsigned char* src = base + ((size_t)blockIdx.x * 8 + warp) * warp_stride;
  unsigned char* dst = sm + warp * SLOT;
  const int pb = SLOT / parts;
  for (int p = 0; p < parts; ++p) { mbar_expect_tx(&bars[warp * 4 + p], pb); bulk_load(dst + p * pb, src + p * pb, pb, &bars[warp * 4 + p]); }
  for (int i = 1; i <= n; ++i) {
    const size_t off = ((size_t)i * SLOT) & (wrap - 1);
    for (int p = 0; p < parts; ++p) {
      mbar_wait(&bars[warp * 4 + p], (i - 1) & 1);
      if (delay) { const long long t = clock64(); while (clock64() - t < delay) {} }  // stands for the compute on the part
      if (i < n) { mbar_expect_tx(&bars[warp * 4 + p], pb); bulk_load(dst + p * pb, src + off + p * pb, pb, &bars[warp * 4 + p]); }
    }
  }
}
int main() {
  unsigned char* buf;
  const size_t TOT = 4ull << 30;
  CK(cudaMalloc(&buf, TOT));
  CK(cudaMemset(buf, 1, TOT));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SLOT + 512));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas : {112, 148}) {
    for (int hot = 0; hot < 2; ++hot) {
      // cold: every warp walks its own 4 MB (x 896 or 1184 warps = 3.5 / 4.6 GB > 4 GB for 148: use 3 MB); hot: 48 KB per warp re-read (43 / 57 MB total)
      const size_t wstride = hot ? 65536 : (ctas == 112 ? (4u << 20) : (2u << 20));
      const size_t wrap = hot ? 32768 : wstride;
      const int n = hot ? 400 : (int)(wstride / SLOT);
      for (int parts : {1, 2, 4}) {
        for (int delay : {0, 600}) {
          k<<<ctas, 256, 8 * SLOT + 512>>>(buf, wstride, wrap, parts, 8, 0);  // warm-up (and L2 fill for the hot case)
          if (hot) k<<<ctas, 256, 8 * SLOT + 512>>>(buf, wstride, wrap, parts, 2, 0);
          CK(cudaEventRecord(e0));
          k<<<ctas, 256, 8 * SLOT + 512>>>(buf, wstride, wrap, parts, n, delay);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          const double bytes = (double)ctas * 8 * n * SLOT;
          printf("%3d CTAs %s parts %d (%5d B) delay %3d clk: %7.1f GB/s  (%.1f B/clk/SM at 1.965 GHz)\n", ctas, hot ? "L2 " : "HBM", parts, SLOT / parts, delay,
                 bytes / ms * 1e-6, bytes / ms * 1e-6 / ctas / 1.965);
        }
      }
    }
  }
  return 0;
}
