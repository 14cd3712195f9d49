// mb_store.cu — how fast can the SMs WRITE?  148 CTAs x W warps store float4 per lane into a region of S MB (L2-resident for S <= ~60,
// HBM-bound above), either fully contiguous (512 B per warp instruction) or as the GEMM epilogue does (4 row segments of 128 B, rows
// 2 KB apart).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mb_store.bin scripts/mb_store.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
template <int PATTERN>
__global__ void k_store(float4* __restrict__ out, size_t n_f4, int iters) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarp = (size_t)gridDim.x * (blockDim.x >> 5);
  const float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
  for (int it = 0; it < iters; ++it) {
    if (PATTERN == 0) {  // contiguous: 512 B per instruction
      for (size_t i = warp * 32 + lane; i < n_f4; i += nwarp * 32) out[i] = v;
    } else if (PATTERN == 2 || PATTERN == 3) {
      // thread = row (what a tcgen05.ld 32x32b register tile gives without re-staging): a warp covers 32 rows x 128 B with 8 stores of
      // 16 B per lane (PATTERN 2: half sectors) or 4 stores of 32 B per lane (PATTERN 3: st.global.v8, whole sectors)
      const size_t rows = n_f4 / 128;
      for (size_t blk = warp; blk < (rows / 32) * 16; blk += nwarp) {
        float4* p = out + ((blk / 16) * 32 + lane) * 128 + (blk % 16) * 8;
        if (PATTERN == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = v;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + 2 * j), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
    } else {             // rows of 512 floats (2 KB): a warp instruction writes the 128-byte segment `seg` of 4 consecutive rows
      const size_t rows = n_f4 / 128;  // float4 per row = 128
      for (size_t blk = warp; blk < (rows / 4) * 16; blk += nwarp) {
        const size_t r0 = (blk / 16) * 4, seg = blk % 16;
        out[(r0 + (lane >> 3)) * 128 + seg * 8 + (lane & 7)] = v;
      }
    }
  }
}
int main() {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int mb : {32, 256}) {
    const size_t bytes = (size_t)mb << 20, n_f4 = bytes / 16;
    float4* buf; CK(cudaMalloc(&buf, bytes));
    for (int warps : {4, 8}) {
      for (int pat = 0; pat < 4; ++pat) {
        const int iters = mb <= 64 ? 20 : 4;
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaEventRecord(e0));
          if (pat == 0) k_store<0><<<148, warps * 32>>>(buf, n_f4, iters);
          else if (pat == 1) k_store<1><<<148, warps * 32>>>(buf, n_f4, iters);
          else if (pat == 2) k_store<2><<<148, warps * 32>>>(buf, n_f4, iters);
          else k_store<3><<<148, warps * 32>>>(buf, n_f4, iters);
          CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%5d MB region, %2d warps/SM, %s: %7.1f GB/s (%.1f B/clk/SM at 1.965 GHz)\n", mb, warps, pat == 0 ? "contiguous 512 B       " : pat == 1 ? "4 x 128 B row segments" : pat == 2 ? "lane = row, 16 B      " : "lane = row, 32 B (v8) ",
               (double)bytes * iters / ms * 1e-6, (double)bytes * iters / ms * 1e-6 / 148 / 1.965);
      }
    }
    cudaFree(buf);
  }
  return 0;
}
