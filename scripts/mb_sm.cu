// mb_sm.cu — per-SM rates behind the GEMV phases of the cluster-stream kernel: (a) legacy mma.sync.m16n8k16 bf16 issue rate
// with 8 warps per SM, (b) 128-bit ld.global.nc streaming rate from an L2-resident region with 16 loads in flight per lane.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void mma(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__global__ void k_mma(int n, int chains, float* out, long long* cyc) {
  float c[8][4] = {};
  uint4 a = make_uint4(threadIdx.x, 2, 3, 4);
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) if (j < chains) mma(c[j], a, i, j);
  }
  const long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][3];
  if (s == 1.2345f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__device__ __forceinline__ uint4 ldw(const void* p) {
  uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
struct U8 { uint32_t v[8]; };
__device__ __forceinline__ U8 ldw256(const void* p) {
  U8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}
// 256-bit loads: 8 in flight per lane (same bytes in flight as 16 x 128-bit)
__global__ void k_ldg256(const unsigned char* base, int per_warp, int reps, unsigned* sink, long long* cyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const unsigned char* p0 = base + ((size_t)blockIdx.x * nw + warp) * per_warp + lane * 32;
  unsigned acc = 0;
  const int nb = per_warp / 8192;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    U8 f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = ldw256(p0 + j * 1024);
    for (int b = 1; b <= nb; ++b) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc ^= f[j].v[0] ^ f[j].v[7];
        if (b < nb) f[j] = ldw256(p0 + (size_t)(b * 8 + j) * 1024);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (acc == 0x1234567u) sink[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// each warp streams its own contiguous slice of `per_warp` bytes, `reps` times (L2 resident after the first pass)
__global__ void k_ldg(const unsigned char* base, int per_warp, int reps, unsigned* sink, long long* cyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint4* p0 = reinterpret_cast<const uint4*>(base + ((size_t)blockIdx.x * nw + warp) * per_warp) + lane;
  unsigned acc = 0;
  const int nb = per_warp / 8192;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint4 f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = ldw(p0 + j * 32);
    for (int b = 1; b <= nb; ++b) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        acc ^= f[j].x ^ f[j].w;
        if (b < nb) f[j] = ldw(p0 + (b * 16 + j) * 32);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (acc == 0x1234567u) sink[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; unsigned* sink; unsigned char* buf;
  CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 148 * 8)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&buf, 256 << 20)); CK(cudaMemset(buf, 1, 256 << 20));
  long long h[148];
  for (int ctas : {1, 112})
    for (int warps : {1, 4, 8})
      for (int chains : {1, 4, 8}) {
        const int n = 2000;
        k_mma<<<ctas, warps * 32>>>(n, chains, out, cyc); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
        printf("mma.sync m16n8k16: %3d CTAs x %d warps, %d independent chains: %.1f cycles per mma per warp, %.2f mma/clk/SM\n", ctas, warps, chains,
               (double)h[0] / (n * chains), (double)warps * n * chains / h[0]);
      }
  for (int ctas : {1, 16, 112})
    for (int warps : {8, 16}) {
      const int per_warp = 49152, reps = 50;  // 8 warps x 48 KB = 393 KB per CTA: one layer's slice
      k_ldg<<<ctas, warps * 32>>>(buf, per_warp, reps, sink, cyc); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
      double c = 0; for (int i = 0; i < ctas; ++i) c += h[i]; c /= ctas;
      printf("ld.global.nc 128-bit, L2 resident: %3d CTAs x %2d warps, 16 in flight per lane: %.1f B/clk/SM (%.0f GB/s per SM @1.9GHz)\n", ctas, warps,
             (double)warps * per_warp * reps / c, (double)warps * per_warp * reps / c * 1.9);
    }
  for (int ctas : {1, 16, 112})
    for (int warps : {8}) {
      const int per_warp = 49152, reps = 50;
      k_ldg256<<<ctas, warps * 32>>>(buf, per_warp, reps, sink, cyc); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
      double c = 0; for (int i = 0; i < ctas; ++i) c += h[i]; c /= ctas;
      printf("ld.global.nc 256-bit, L2 resident: %3d CTAs x %2d warps, 8 in flight per lane: %.1f B/clk/SM (%.0f GB/s per SM @1.9GHz)\n", ctas, warps,
             (double)warps * per_warp * reps / c, (double)warps * per_warp * reps / c * 1.9);
    }
  // same data for every CTA (what the 7 clusters do with the weight stream): 112 CTAs, all reading CTA 0's slices
  return 0;
}
