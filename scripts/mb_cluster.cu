// mb_cluster.cu — microbenchmark behind DESIGN.md "cluster-stream decode": can C-CTA clusters, each streaming a
// private slice of the (shared) layer weights + private KV bytes through a TMA-fed shared-memory ring, with 4
// cluster barriers + DSMEM all-gathers per layer, sustain the L2->SM ingest the design needs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb_cluster scripts/mb_cluster.cu && /tmp/mb_cluster
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int CHUNK = 32768;
constexpr int STAGES = 5;
constexpr int NCW = 8;  // consumer warps
constexpr int NTHREADS = (NCW + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAITC_%=:\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n @p bra DONEC_%=;\n bra WAITC_%=;\n DONEC_%=:\n}\n" ::"r"(
          smem_u32(b)),
      "r"(parity)
      : "memory");
}
// cluster-wide barrier among the CONSUMER warps only (the TMA producer warp runs ahead freely)
template <int C>
__device__ __forceinline__ void consumer_cluster_sync(uint64_t* cbar, uint32_t& parity) {
  asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32) : "memory");
  if (threadIdx.x < C) mbar_arrive_remote(mapa(smem_u32(cbar), threadIdx.x));
  mbar_wait_cluster(cbar, parity);
  parity ^= 1;
}

__device__ __forceinline__ void st_async_f4(uint32_t addr, float4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar)
               : "memory");
}

struct Smem {
  unsigned char ring[STAGES][CHUNK];
  float gather[4][16][64];  // 4 exchange buffers, up to 16 peers x 64 floats
  uint64_t full[STAGES], empty[STAGES], cbar;
};

// w: shared weights [n_layers][C][wchunks*CHUNK]; kv: private per CTA [n_cta][n_layers][kvchunks*CHUNK]
template <int C>
__global__ void __launch_bounds__(NTHREADS, 1)
mb_stream(const unsigned char* __restrict__ w, const unsigned char* __restrict__ kv, int n_layers, int wchunks, int kvchunks,
          int steps, int do_sync, float* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>(raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cta = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], NCW); }
    mbar_init(&sm.cbar, C);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_arrive(); cluster_wait();
  const int cpl = wchunks + kvchunks;  // chunks per layer
  if (warp == NCW) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int st = 0; st < steps; ++st)
        for (int l = 0; l < n_layers; ++l)
          for (int c = 0; c < cpl; ++c) {
            mbar_wait(&sm.empty[stage], phase ^ 1);
            const unsigned char* src = (c < wchunks)
                ? w + ((size_t)(l * C + rank) * wchunks + c) * CHUNK
                : kv + (((size_t)cta * n_layers + l) * kvchunks + (c - wchunks)) * CHUNK;
            mbar_expect_tx(&sm.full[stage], CHUNK);
            bulk_load(sm.ring[stage], src, CHUNK, &sm.full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
    }
  } else {
    int stage = 0; uint32_t phase = 0;
    uint32_t acc = 0, cpar = 0;
    float4 facc = make_float4(0, 0, 0, 0);
    const int sync_every = cpl / 4 > 0 ? cpl / 4 : 1;
    for (int st = 0; st < steps; ++st)
      for (int l = 0; l < n_layers; ++l) {
        int nsync = 0;
        for (int c = 0; c < cpl; ++c) {
          mbar_wait(&sm.full[stage], phase);
          const uint4* p = reinterpret_cast<const uint4*>(sm.ring[stage]) + warp * 256 + lane;
#pragma unroll
          for (int i = 0; i < 8; ++i) { uint4 v = p[i * 32]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (do_sync && ((c + 1) % sync_every == 0) && nsync < 4) {
            // all-gather: every thread of warps 0..C/… writes 16 B to each peer
            const int k = nsync++;
            if (threadIdx.x < 16 * C) {
              const int peer = threadIdx.x >> 4, q = threadIdx.x & 15;
              const uint32_t a = mapa(smem_u32(&sm.gather[k][rank][q * 4]), peer);
              st_cluster_f4(a, make_float4((float)acc, 1.f, 2.f, 3.f));
            }
            consumer_cluster_sync<C>(&sm.cbar, cpar);
            const float4 g = *reinterpret_cast<const float4*>(&sm.gather[k][threadIdx.x % C][(threadIdx.x >> 4) * 4 % 64]);
            facc.x += g.x; facc.y += g.y;
          }
        }
        if (do_sync) for (; nsync < 4; ++nsync) consumer_cluster_sync<C>(&sm.cbar, cpar);
      }
    if (acc == 0x12345678u || facc.x == 1.2345f) out[threadIdx.x] = acc + facc.y;
  }
  __syncwarp();
  cluster_arrive(); cluster_wait();
}

// sync-only: n exchanges of (16 B from each of 256 threads -> spread over the C peers) + wait.  mode 0: plain DSMEM stores +
// bar.sync + remote mbarrier arrive.release + wait.acquire;  mode 1: st.async (data carries the completion) + wait
template <int C>
__global__ void __launch_bounds__(256, 1) mb_sync(int n, int mode, float* out) {
  __shared__ __align__(16) float buf[2][16][64];
  __shared__ uint64_t cbar, ebar[2];
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&cbar, C); mbar_init(&ebar[0], 1); mbar_init(&ebar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_arrive(); cluster_wait();
  uint32_t cpar = 0, epar[2] = {0, 0};
  float acc = 0.f;
  const int peer = threadIdx.x >> 4, q = threadIdx.x & 15;  // 16 threads per peer (C=16: all 256 threads)
  for (int i = 0; i < n; ++i) {
    const int k = i & 1;
    if (mode == 0) {
      if (peer < C) st_cluster_f4(mapa(smem_u32(&buf[k][rank][q * 4]), peer), make_float4(acc, 1.f, 2.f, 3.f));
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x < C) mbar_arrive_remote(mapa(smem_u32(&cbar), threadIdx.x));
      mbar_wait_cluster(&cbar, cpar); cpar ^= 1;
    } else {
      if (threadIdx.x == 0) mbar_expect_tx(&ebar[k], C * 16 * 16);
      if (peer < C) st_async_f4(mapa(smem_u32(&buf[k][rank][q * 4]), peer), make_float4(acc, 1.f, 2.f, 3.f), mapa(smem_u32(&ebar[k]), peer));
      mbar_wait_cluster(&ebar[k], epar[k]); epar[k] ^= 1;
    }
    acc += buf[k][threadIdx.x % C][q * 4];
  }
  if (acc == 1.2345f) out[0] = acc;
  __syncwarp();
  cluster_arrive(); cluster_wait();
}
template <int C>
void run_sync(int nclusters, int mode, float* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * C); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (C > 8) CK(cudaFuncSetAttribute(mb_sync<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int n = 20000;
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, mb_sync<C>, n, mode, out));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  printf("sync-only C=%2d clusters=%2d mode=%d (%s): %.0f ns per exchange\n", C, nclusters, mode, mode ? "st.async" : "st + arrive.release", best * 1e6 / n);
}

template <int C>
void run(int nclusters, int kvchunks, int do_sync, const unsigned char* w, const unsigned char* kv, float* out, int steps, int n_layers = 24) {
  const int wchunks = 786432 * 8 / C / CHUNK;  // 24 (C=8) or 12 (C=16)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * C); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = sizeof(Smem);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(mb_stream<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  if (C > 8) CK(cudaFuncSetAttribute(mb_stream<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int maxc = 0;
  CK(cudaOccupancyMaxActiveClusters(&maxc, mb_stream<C>, &cfg));
  if (nclusters > maxc) { printf("C=%d: requested %d clusters > max active %d: skip\n", C, nclusters, maxc); return; }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, mb_stream<C>, w, kv, n_layers, wchunks, kvchunks, steps, do_sync, out));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double us_step = best * 1000.0 / steps;
  const double bytes_cta = (double)n_layers * (wchunks + kvchunks) * CHUNK;
  const double ingest = bytes_cta * nclusters * C / (us_step * 1e-6) / 1e12;
  const double hbm = ((double)n_layers * wchunks * CHUNK * C + (double)n_layers * kvchunks * CHUNK * nclusters * C) / (us_step * 1e-6) / 1e12;
  printf("L=%2d C=%2d clusters=%2d (max %2d) kvchunks/layer/cta=%2d sync=%d : %8.1f us/step  per-SM %.1f GB/s  ingest %.2f TB/s  unique(HBM) %.2f TB/s\n",
         n_layers, C, nclusters, maxc, kvchunks, do_sync, us_step, bytes_cta / (us_step * 1e-6) / 1e9, ingest, hbm);
}

int main() {
  const size_t wbytes = (size_t)24 * 786432 * 8;  // 151 MB
  const size_t kvbytes = (size_t)148 * 24 * 16 * CHUNK;  // up to 16 chunks / layer / cta = 1.86 GB
  unsigned char *w, *kv; float* out;
  CK(cudaMalloc(&w, wbytes)); CK(cudaMalloc(&kv, kvbytes)); CK(cudaMalloc(&out, 4096));
  CK(cudaMemset(w, 1, wbytes)); CK(cudaMemset(kv, 2, kvbytes));
  for (int mode = 0; mode < 2; ++mode) { run_sync<8>(1, mode, out); run_sync<8>(15, mode, out); run_sync<16>(1, mode, out); run_sync<16>(7, mode, out); run_sync<4>(1, mode, out); }
  int dev_sms; CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d, smem/CTA %zu\n", dev_sms, sizeof(Smem));
  const int steps = 20;
  for (int sync = 0; sync <= 1; ++sync) {
    run<16>(1, 0, sync, w, kv, out, steps);
    run<16>(1, 0, sync, w, kv, out, steps * 4, 6);   // 38 MB of weights: L2 resident after the first step
    run<16>(7, 0, sync, w, kv, out, steps);
    run<16>(7, 0, sync, w, kv, out, steps * 4, 6);
    run<16>(7, 6, sync, w, kv, out, steps);           // + private KV (B=32 @ S~743: 1.2 GB/step/112 CTAs/24 layers = 14 chunks of 32K)
    run<16>(7, 14, sync, w, kv, out, steps);
    run<8>(15, 0, sync, w, kv, out, steps);
    run<8>(15, 6, sync, w, kv, out, steps);
  }
  return 0;
}
