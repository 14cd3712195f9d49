// mb_tma.cu — TMA bulk-load latency / throughput per SM: one CTA (or many), chunk size, in-flight depth, L2-hot vs HBM.
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
__device__ __forceinline__ void bulk_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// depth loads of `chunk` bytes kept in flight by one thread; n loads total; region = bytes cycled through (small: L2 hot)
__global__ void k(const unsigned char* base, size_t region, int chunk, int depth, int n, int pf_ahead, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 200 * 1024);
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned char* b = base + (size_t)blockIdx.x * region;
    size_t off = 0, poff = 0;
    for (int i = 0; i < pf_ahead; ++i) { bulk_prefetch(b + poff, chunk); poff += chunk; if (poff + chunk > region) poff = 0; }
    long long t0 = clock64();
    for (int i = 0; i < n + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) mbar_wait(&bars[s], ((i / depth) - 1) & 1);
      if (i < n) {
        if (pf_ahead) { bulk_prefetch(b + poff, chunk); poff += chunk; if (poff + chunk > region) poff = 0; }
        mbar_expect_tx(&bars[s], chunk);
        bulk_load(sm + (size_t)s * chunk, b + off, chunk, &bars[s]);
        off += chunk; if (off + chunk > region) off = 0;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  const size_t total = (size_t)148 * 64 * 1024 * 1024;  // 64 MB per CTA region max
  unsigned char* buf; long long* out; CK(cudaMalloc(&buf, total)); CK(cudaMalloc(&out, 148 * 8)); CK(cudaMemset(buf, 1, total));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  long long h[148];
  struct Cfg { int ctas; size_t region; int chunk, depth, pf; const char* name; };
  Cfg cfgs[] = {
    {1, 64 << 10, 16384, 1, 0, "1 CTA, L2-hot, 16K, depth 1 (latency)"},
    {1, 64 << 20, 16384, 1, 0, "1 CTA, HBM,    16K, depth 1 (latency)"},
    {1, 64 << 10, 4096, 1, 0, "1 CTA, L2-hot,  4K, depth 1 (latency)"},
    {1, 64 << 20, 4096, 1, 0, "1 CTA, HBM,     4K, depth 1 (latency)"},
    {1, 64 << 10, 16384, 4, 0, "1 CTA, L2-hot, 16K, depth 4"},
    {1, 64 << 10, 16384, 8, 0, "1 CTA, L2-hot, 16K, depth 8"},
    {1, 64 << 10, 16384, 12, 0, "1 CTA, L2-hot, 16K, depth 12"},
    {1, 64 << 20, 16384, 4, 0, "1 CTA, HBM,    16K, depth 4"},
    {1, 64 << 20, 16384, 8, 0, "1 CTA, HBM,    16K, depth 8"},
    {1, 64 << 20, 16384, 12, 0, "1 CTA, HBM,    16K, depth 12"},
    {1, 64 << 20, 16384, 8, 16, "1 CTA, HBM + L2 prefetch 16 ahead, 16K, depth 8"},
    {1, 64 << 20, 16384, 8, 64, "1 CTA, HBM + L2 prefetch 64 ahead, 16K, depth 8"},
    {1, 64 << 20, 16384, 4, 64, "1 CTA, HBM + L2 prefetch 64 ahead, 16K, depth 4"},
    {1, 64 << 20, 4096, 8, 0, "1 CTA, HBM,     4K, depth 8"},
    {1, 64 << 20, 4096, 32, 0, "1 CTA, HBM,     4K, depth 32"},
    {112, 64 << 20, 16384, 8, 0, "112 CTAs, HBM,   16K, depth 8"},
    {112, 64 << 20, 16384, 8, 64, "112 CTAs, HBM + L2 prefetch 64 ahead, 16K, depth 8"},
    {112, 64 << 10, 16384, 8, 0, "112 CTAs, L2-hot, 16K, depth 8"},
    {148, 64 << 20, 16384, 8, 0, "148 CTAs, HBM,   16K, depth 8"},
    {148, 64 << 20, 16384, 12, 0, "148 CTAs, HBM,   16K, depth 12"},
  };
  for (auto& c : cfgs) {
    const int n = 2000;
    k<<<c.ctas, 32, 201 * 1024>>>(buf, c.region, c.chunk, c.depth, 200, c.pf, out);  // warm
    k<<<c.ctas, 32, 201 * 1024>>>(buf, c.region, c.chunk, c.depth, n, c.pf, out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, c.ctas * 8, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < c.ctas; ++i) cyc += h[i]; cyc /= c.ctas;
    const double per = cyc / n;
    printf("%-52s: %7.0f cyc/load  -> %6.1f GB/s per SM @1.9GHz (%.2f us per load-slot)\n", c.name, per, c.chunk / per * 1.9, per / 1900.0);
  }
  return 0;
}
