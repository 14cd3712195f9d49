// mb_ll2.cu — the floor of CTA-to-CTA signalling through L2 on B200, by load / store flavour and by fan-out.
//   ping-pong: CTA 0 stores {value, tag i} into a cell, CTA 1 polls it and answers in another cell; one round trip = two
//   one-way hand-offs.  Flavours: volatile (= relaxed.sys), relaxed.gpu, weak .cg loads, release/acquire.
//   fan-out:   CTA 0 stores one 128-byte line of cells, N CTAs poll it and answer each in its own cell; CTA 0 polls the N answers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mb_ll2.bin scripts/mb_ll2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int F> __device__ __forceinline__ uint2 ld8(const void* p) {
  uint2 r;
  if (F == 0) asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  else if (F == 1) asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  else if (F == 2 || F == 4) asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  else if (F == 5) asm volatile("ld.global.cv.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  else asm volatile("ld.acquire.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
template <int F> __device__ __forceinline__ void st8(void* p, uint32_t a, uint32_t b) {
  if (F == 0 || F == 4 || F == 5) asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
  else if (F == 1) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
  else if (F == 2) asm volatile("st.global.cg.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
  else asm volatile("st.release.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

// grid = 1 + n pollers; cells: [0..15] written by CTA 0 (one line), answers at 16 + 16*k (one line per poller)
template <int F>
__global__ void k_fan(uint2* cells, int iters) {
  extern __shared__ unsigned char big[];  // 200 KB of dynamic shared memory: one CTA per SM, so every hand-off crosses SMs
  const int cta = blockIdx.x, n = gridDim.x - 1;
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  for (int it = 1; it <= iters; ++it) {
    if (cta == 0) {
      if (lane < 16) st8<F>(cells + lane, (uint32_t)lane, (uint32_t)it);
      // wait for the n answers: lane k polls answer k, k + 32, ...
      for (int k = lane; k < n; k += 32) {
        unsigned spins = 0;
        while (ld8<F>(cells + 16 + 16 * k).y != (uint32_t)it && ++spins < (1u << 24)) {}
      }
      __syncwarp();
    } else {
      if (lane < 16) {
        unsigned spins = 0;
        while (ld8<F>(cells + lane).y != (uint32_t)it && ++spins < (1u << 24)) {}
      }
      __syncwarp();
      if (lane == 0) st8<F>(cells + 16 + 16 * (cta - 1), 1u, (uint32_t)it);
    }
  }
}

int main() {
  uint2* cells;
  cudaMalloc(&cells, 1 << 20);
  const int iters = 5000;
  const char* names[6] = {"ld/st.volatile", "ld/st.relaxed.gpu", "ld/st.cg (weak)", "st.release / ld.acquire (gpu)",
                          "st.volatile / ld.cg", "st.volatile / ld.cv"};
  const void* fns[6] = {(const void*)k_fan<0>, (const void*)k_fan<1>, (const void*)k_fan<2>, (const void*)k_fan<3>, (const void*)k_fan<4>, (const void*)k_fan<5>};
  for (int f = 0; f < 6; ++f) cudaFuncSetAttribute(fns[f], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int f = 0; f < 6; ++f) {
    for (int n : {1, 8, 32, 143}) {
      cudaMemset(cells, 0, 1 << 20);
      void* args[] = {&cells, (void*)&iters};
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaError_t err = cudaLaunchCooperativeKernel(fns[f], dim3(1 + n), dim3(32), args, 200 * 1024, 0);
      unsigned check[4] = {0, 0, 0, 0};
      cudaMemcpy(check, cells, 16, cudaMemcpyDeviceToHost);
      cudaEventRecord(e1);
      cudaError_t e2 = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      printf("%-30s 1 -> %3d -> 1: %.3f us per round trip (= 2 hand-offs)  (%s %s; last tag %u)\n", names[f], n, ms * 1000.0 / iters,
             cudaGetErrorString(err), cudaGetErrorString(e2), check[1]);
      fflush(stdout);
    }
  }
  return 0;
}
