// mb_tma2.cu — what bounds the TMA issue rate of a producer?  Variants of the issue loop (L2-hot 64 KB region, 16 KB loads).
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
constexpr int CH = 16384, DEPTH = 8;
// mode 0: 1 thread, per load {wait, expect, load}.  mode 1: 1 thread, groups of 4 loads share one barrier (1 wait + 1 expect + 4 loads of 4 KB.. here 4 x 16 KB into 4 slots)
// mode 2: lanes 0..3 of one warp, each its own slots.  mode 3: 4 warps (lane 0 each), each its own slots.  mode 4: 1 thread, 32 KB loads
__global__ void k(const unsigned char* base, int mode, int n, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 200 * 1024);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  if (mode == 0 && threadIdx.x == 0) {
    for (int i = 0; i < n + DEPTH; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) mbar_wait(&bars[s], ((i / DEPTH) - 1) & 1);
      if (i < n) { mbar_expect_tx(&bars[s], CH); bulk_load(sm + s * CH, base + (i & 3) * CH, CH, &bars[s]); }
    }
  } else if (mode == 1 && threadIdx.x == 0) {  // 2 groups of 4 slots; one barrier per group
    for (int i = 0; i < n / 4 + 2; ++i) {
      const int gq = i & 1;
      if (i >= 2) mbar_wait(&bars[gq], ((i / 2) - 1) & 1);
      if (i < n / 4) {
        mbar_expect_tx(&bars[gq], 4 * CH);
        for (int j = 0; j < 4; ++j) bulk_load(sm + (gq * 4 + j) * CH, base + j * CH, CH, &bars[gq]);
      }
    }
  } else if (mode == 2 && warp == 0 && lane < 4) {  // 4 lanes: lane owns slots lane, lane+4
    for (int i = 0; i < n / 4 + 2; ++i) {
      const int s = lane + 4 * (i & 1);
      if (i >= 2) mbar_wait(&bars[s], ((i / 2) - 1) & 1);
      if (i < n / 4) { mbar_expect_tx(&bars[s], CH); bulk_load(sm + s * CH, base + lane * CH, CH, &bars[s]); }
    }
  } else if (mode == 3 && lane == 0 && warp < 4) {  // 4 warps
    for (int i = 0; i < n / 4 + 2; ++i) {
      const int s = warp + 4 * (i & 1);
      if (i >= 2) mbar_wait(&bars[s], ((i / 2) - 1) & 1);
      if (i < n / 4) { mbar_expect_tx(&bars[s], CH); bulk_load(sm + s * CH, base + warp * CH, CH, &bars[s]); }
    }
  } else if (mode == 4 && threadIdx.x == 0) {  // 32 KB loads, 4 slots
    for (int i = 0; i < n / 2 + 4; ++i) {
      const int s = i % 4;
      if (i >= 4) mbar_wait(&bars[s], ((i / 4) - 1) & 1);
      if (i < n / 2) { mbar_expect_tx(&bars[s], 2 * CH); bulk_load(sm + s * 2 * CH, base + (i & 1) * 2 * CH, 2 * CH, &bars[s]); }
    }
  } else if (mode == 5 && lane == 0 && warp < 2) {  // 2 warps, 32 KB loads
    for (int i = 0; i < n / 4 + 2; ++i) {
      const int s = warp + 2 * (i & 1);
      if (i >= 2) mbar_wait(&bars[s], ((i / 2) - 1) & 1);
      if (i < n / 4) { mbar_expect_tx(&bars[s], 2 * CH); bulk_load(sm + s * 2 * CH, base + warp * 2 * CH, 2 * CH, &bars[s]); }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
  unsigned char* buf; long long* out; CK(cudaMalloc(&buf, 1 << 20)); CK(cudaMalloc(&out, 148 * 8)); CK(cudaMemset(buf, 1, 1 << 20));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  const char* names[] = {"1 thread {wait,expect,load} x 16K", "1 thread, 4 loads per barrier", "4 lanes of one warp", "4 warps", "1 thread, 32K loads", "2 warps, 32K loads"};
  for (int ctas : {1, 112})
    for (int mode = 0; mode < 6; ++mode) {
      const int n = 4000;
      k<<<ctas, 128, 201 * 1024>>>(buf, mode, 400, out);
      k<<<ctas, 128, 201 * 1024>>>(buf, mode, n, out);
      CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost));
      double cyc = 0; for (int i = 0; i < ctas; ++i) cyc += h[i]; cyc /= ctas;
      printf("%3d CTAs  %-36s: %6.0f cyc per 16 KB -> %6.1f GB/s per SM @1.9GHz\n", ctas, names[mode], cyc / n, 16384.0 / (cyc / n) * 1.9);
    }
  return 0;
}
