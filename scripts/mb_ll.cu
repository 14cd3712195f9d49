// mb_ll.cu — how fast is an all-to-all exchange between ~144 co-resident CTAs through L2 when data and "ready" travel in the
// same 8-byte cell {value, tag} (the LL protocol of the wide decode kernel), against the counter grid barrier?
//   exchange i: every CTA stores its share of a CELLS-cell vector (x REPL replicas), then warp 0 of every CTA polls one
//   replica until all CELLS tags equal i, sums the values (so the next store depends on the gather) and goes on.
//   Two buffers alternate: a CTA can run at most one exchange ahead of the slowest one (it needs everybody's cells of exchange
//   i + 1 to go further), so it never overwrites cells somebody is still polling.  (In the decode kernel the chain of hand-offs
//   inside a layer gives the same guarantee with single buffers.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mb_ll.bin scripts/mb_ll.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint4 ldv16(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stv8(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

// CELLS cells per exchange, written by gridDim.x CTAs (CELLS / gridDim.x each, rounded up), REPL replicas REP_STRIDE cells apart
template <int NL>
__global__ void k_ll(uint2* buf, int cells, int repl, int rep_stride, int iters, float* sink, long long* cyc) {
  const int cta = blockIdx.x, n = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (cells + n - 1) / n;
  float carry = 1.0f;
  __shared__ float s_carry;
  long long t0 = clock64();
  uint2* const buf0 = buf;
  for (int it = 1; it <= iters; ++it) {
    buf = buf0 + (size_t)(it & 1) * (4 << 20) / 8 * 8;  // 4 M cells apart
    // produce: threads 0 .. per*repl-1 store one cell each
    if (tid < per * repl) {
      const int c = cta * per + tid % per, r = tid / per;
      if (c < cells) stv8(buf + (size_t)r * rep_stride + c, __float_as_uint(carry + c), (uint32_t)it);
    }
    // gather: warp 0 reads replica cta % repl, NL 16-byte loads per lane (2 cells each)
    if (warp == 0) {
      const uint4* src = reinterpret_cast<const uint4*>(buf + (size_t)(cta % repl) * rep_stride) + lane;
      uint4 v[NL];
      unsigned spins = 0;
#pragma unroll
      for (int i = 0; i < NL; ++i) v[i] = ldv16(src + i * 32);
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < NL; ++i)
          if (v[i].y != (uint32_t)it || v[i].w != (uint32_t)it) { v[i] = ldv16(src + i * 32); ok = false; }
        if (ok || ++spins > (1u << 22)) break;  // ~seconds: a lost cell must not hang the box
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NL; ++i) s += __uint_as_float(v[i].x) + __uint_as_float(v[i].z);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_carry = s * 1e-9f;
    }
    __syncthreads();
    carry = s_carry;
    __syncthreads();
  }
  if (tid == 0) { cyc[cta] = clock64() - t0; sink[cta] = carry; }
}

__global__ void k_bar(unsigned* counter, int iters, long long* cyc) {
  unsigned target = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const double ghz = prop.clockRate * 1e-6;
  const int iters = 2000;
  uint2* buf; float* sink; long long* cyc; unsigned* counter;
  cudaMalloc(&buf, 96 << 20); cudaMalloc(&sink, 4096); cudaMalloc(&cyc, 8 * 1024); cudaMalloc(&counter, 4);
  std::vector<long long> h(1024);
  for (int grid : {144, 128, 64}) {
    for (int cells : {512, 2048}) {
      for (int repl : {1, 2, 4, 8}) {
        const int per = (cells + grid - 1) / grid;
        if (per * repl > 256) continue;
        cudaMemset(buf, 0, 96 << 20);
        const int rep_stride = 8192;  // cells: 64 KB apart
        void* args[] = {&buf, (void*)&cells, (void*)&repl, (void*)&rep_stride, (void*)&iters, &sink, &cyc};
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaError_t err = (cells == 512) ? cudaLaunchCooperativeKernel((void*)k_ll<8>, dim3(grid), dim3(256), args, 0, 0)
                                         : cudaLaunchCooperativeKernel((void*)k_ll<32>, dim3(grid), dim3(256), args, 0, 0);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("LL exchange: %3d CTAs, %4d cells (%5d B), %d replica(s): %.3f us/exchange  (%s %s)\n", grid, cells, cells * 8, repl,
               ms * 1000.0 / iters, cudaGetErrorString(err), cudaGetErrorString(e2));
        fflush(stdout);
      }
    }
  }
  for (int grid : {144, 128, 64}) {
    cudaMemset(counter, 0, 4);
    void* args[] = {&counter, (void*)&iters, &cyc};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)k_bar, dim3(grid), dim3(256), args, 0, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("counter grid barrier: %3d CTAs: %.3f us/barrier\n", grid, ms * 1000.0 / iters);
    fflush(stdout);
  }
  printf("(clock %.2f GHz)\n", ghz);
  return 0;
}
