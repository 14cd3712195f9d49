// mb_gemm.cu — the prefill projections in isolation: k_gemm_tc<128> (one 128 x 128 tile per CTA, two CTAs per SM) against the
// persistent k_gemm_tcp<256> / <128> (operand ring across tiles, two TMEM accumulators, shared-memory-staged epilogue).
// Checks the persistent kernel bit for bit against the non-persistent one (same MMA order over K) and both against a plain
// CUDA-core fp32 reference on sampled rows, then times every projection of a layer at the bench (7,755) and config-5 (115,200)
// row counts with CUDA events.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/mb_gemm.bin scripts/mb_gemm.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../gpt-sovits_b200/csrc/gemm_tc.cuh"

using namespace t2s;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_fill_bf16(bf16* p, size_t n, unsigned seed, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = __float2bfloat16_rn(((int)(h & 0xFFFF) - 32768) * (scale / 32768.f));
  }
}
__global__ void k_fill_f32(float* p, size_t n, unsigned seed, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = ((int)(h & 0xFFFF) - 32768) * (scale / 32768.f);
  }
}
__global__ void k_kvoff(long long* kvoff, int M) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < M) kvoff[r] = kv_row_off(r >> PAGE_SHIFT, r & (PAGE - 1));
}
// reference: one thread per (sampled row, column), fp32 accumulation in k order
__global__ void k_ref(const bf16* A, const bf16* W, const float* bias, int N, int K, const int* rows, int n_rows, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * N) return;
  const int r = rows[i / N], n = i % N;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)r * K + k]) * __bfloat162float(W[(size_t)n * K + k]);
  out[i] = acc + bias[n];
}

struct Bufs {
  bf16 *A, *W, *outb[2], *pool[2];
  float *bias, *resid, *outf[2];
  long long* kvoff;
  int* err;
};

static int g_dbg = 0;
static long long* g_tl = nullptr;
static bool run(int variant, int mode, const Bufs& b, int M, int N, int K, int num_sms, cudaStream_t s) {
  TcEpilogue ep{};
  ep.dbg = g_dbg;
  ep.tl = (variant == 1) ? g_tl : nullptr;
  ep.mode = mode; ep.bias = b.bias; ep.error_flag = b.err;
  const int o = variant ? 1 : 0;
  if (mode == EPI_QKV) { ep.out_f32 = b.outf[o]; ep.kpool = b.pool[o]; ep.vpool = b.pool[o] + KV_V_OFF; ep.kvoff = b.kvoff; ep.layer_off = 0; }
  if (mode == EPI_RESID) { ep.resid = b.resid; ep.out_f32 = b.outf[o]; }
  if (mode == EPI_RELU) ep.out_b16 = b.outb[o];
  if (variant == 0) return launch_gemm_tc<128>(b.A, b.W, M, N, K, ep, s);
  if (variant == 1) return launch_gemm_tcp<256>(b.A, b.W, M, N, K, ep, num_sms, s);
  return launch_gemm_tcp<128>(b.A, b.W, M, N, K, ep, num_sms, s);
}

int main(int argc, char** argv) {
  int only_m = argc > 1 ? atoi(argv[1]) : 0;
  int reps = argc > 2 ? atoi(argv[2]) : 20;
  g_dbg = argc > 3 ? atoi(argv[3]) : 0;  // epilogue ablations (timing only: the comparisons fail by construction)
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int num_sms = prop.multiProcessorCount;
  if (!gemm_tc_init()) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  cudaStream_t s; CK(cudaStreamCreate(&s));
  const int Ms[2] = {7755, 115200};
  struct Shape { const char* name; int mode, N, K; } shapes[4] = {
      {"qkv  (N=1536,K=512)", EPI_QKV, 3 * D, D}, {"wo   (N=512,K=512)", EPI_RESID, D, D},
      {"w1   (N=2048,K=512)", EPI_RELU, FF, D}, {"w2   (N=512,K=2048)", EPI_RESID, D, FF}};
  int fails = 0;
  const bool want_tl = argc > 4 && atoi(argv[4]) != 0;  // print the pipeline timeline of CTAs 0 and 100 of k_gemm_tcp<256>
  if (want_tl) { CK(cudaMalloc(&g_tl, (size_t)148 * 64 * 8)); }
  for (int mi = 0; mi < 2; ++mi) {
    const int M = Ms[mi];
    if (only_m && only_m != M) continue;
    Bufs b{};
    const size_t pool_el = (size_t)((M + PAGE - 1) / PAGE) * KV_PAGE_STRIDE;
    CK(cudaMalloc(&b.A, (size_t)M * FF * 2)); CK(cudaMalloc(&b.W, (size_t)FF * FF * 2));
    for (int o = 0; o < 2; ++o) {
      CK(cudaMalloc(&b.outb[o], (size_t)M * FF * 2)); CK(cudaMalloc(&b.outf[o], (size_t)M * D * 4)); CK(cudaMalloc(&b.pool[o], pool_el * 2));
    }
    CK(cudaMalloc(&b.bias, FF * 4)); CK(cudaMalloc(&b.resid, (size_t)M * D * 4)); CK(cudaMalloc(&b.kvoff, (size_t)M * 8)); CK(cudaMalloc(&b.err, 4));
    CK(cudaMemset(b.err, 0, 4));
    k_fill_bf16<<<1024, 256, 0, s>>>(b.A, (size_t)M * FF, 1u, 1.0f);
    k_fill_bf16<<<1024, 256, 0, s>>>(b.W, (size_t)FF * FF, 2u, 0.05f);
    k_fill_f32<<<64, 256, 0, s>>>(b.bias, FF, 3u, 0.5f);
    k_fill_f32<<<1024, 256, 0, s>>>(b.resid, (size_t)M * D, 4u, 1.0f);
    k_kvoff<<<(M + 255) / 256, 256, 0, s>>>(b.kvoff, M);
    CK(cudaStreamSynchronize(s));
    // sampled rows for the CUDA-core reference
    const int NR = 64;
    std::vector<int> hrows(NR);
    for (int i = 0; i < NR; ++i) hrows[i] = (i < 8) ? i : (i >= NR - 8 ? M - 1 - (NR - 1 - i) : (int)((long long)i * 7919 % M));
    int* drows; float* dref; CK(cudaMalloc(&drows, NR * 4)); CK(cudaMalloc(&dref, (size_t)NR * FF * 4));
    CK(cudaMemcpy(drows, hrows.data(), NR * 4, cudaMemcpyHostToDevice));
    for (int si = 0; si < 4; ++si) {
      const Shape& sh = shapes[si];
      const double flop = 2.0 * M * sh.N * sh.K;
      double ms_v[3] = {0, 0, 0};
      for (int variant = 0; variant < 3; ++variant) {
        if (variant > 0) {  // poison the variant's outputs so that a tile that was never written shows up
          CK(cudaMemsetAsync(b.outf[1], 0xFF, (size_t)M * D * 4, s)); CK(cudaMemsetAsync(b.outb[1], 0xFF, (size_t)M * FF * 2, s));
          CK(cudaMemsetAsync(b.pool[1], 0xFF, pool_el * 2, s));
        } else {
          CK(cudaMemsetAsync(b.pool[0], 0xFF, pool_el * 2, s));
        }
        for (int w = 0; w < 3; ++w) if (!run(variant, sh.mode, b, M, sh.N, sh.K, num_sms, s)) { printf("tensor map failed\n"); return 1; }
        CK(cudaStreamSynchronize(s));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, s));
        for (int r = 0; r < reps; ++r) run(variant, sh.mode, b, M, sh.N, sh.K, num_sms, s);
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        ms_v[variant] = ms / reps;
        if (want_tl && variant == 1) {
          CK(cudaMemset(g_tl, 0, (size_t)148 * 64 * 8));
          run(variant, sh.mode, b, M, sh.N, sh.K, num_sms, s);
          CK(cudaStreamSynchronize(s));
          std::vector<long long> h((size_t)148 * 64);
          CK(cudaMemcpy(h.data(), g_tl, h.size() * 8, cudaMemcpyDeviceToHost));
          for (int cta : {0, 100}) {
            const long long t0 = h[(size_t)cta * 64];
            printf("  timeline CTA %d of %s (us after the producer's first stamp; per tile: producer start | producer done | MMA acc free | first operands | MMAs issued | epilogue start | acc released | epilogue done)\n", cta, sh.name);
            for (int lt = 0; lt < 8 && h[(size_t)cta * 64 + lt * 8]; ++lt) {
              printf("    tile %d:", lt);
              for (int k = 0; k < 8; ++k) printf(" %7.2f", (h[(size_t)cta * 64 + lt * 8 + k] - t0) / 1965.0);
              printf("\n");
            }
          }
        }
        int herr = 0; CK(cudaMemcpy(&herr, b.err, 4, cudaMemcpyDeviceToHost));
        if (herr) { printf("  variant %d: WATCHDOG error flag set\n", variant); ++fails; CK(cudaMemset(b.err, 0, 4)); }
        if (variant > 0) {  // bitwise against variant 0
          size_t nf = 0, nb = 0, np = 0;
          if (sh.mode == EPI_QKV) { nf = (size_t)M * D * 4; np = pool_el * 2; }
          if (sh.mode == EPI_RESID) nf = (size_t)M * D * 4;
          if (sh.mode == EPI_RELU) nb = (size_t)M * sh.N * 2;
          size_t bad = 0;
          auto cmp = [&](const void* x, const void* y, size_t n) {
            if (!n) return;
            std::vector<unsigned char> hx(n), hy(n);
            CK(cudaMemcpy(hx.data(), x, n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hy.data(), y, n, cudaMemcpyDeviceToHost));
            if (memcmp(hx.data(), hy.data(), n) != 0) { for (size_t i = 0; i < n; ++i) bad += hx[i] != hy[i]; }
          };
          cmp(b.outf[0], b.outf[1], nf); cmp(b.outb[0], b.outb[1], nb); cmp(b.pool[0], b.pool[1], np);
          if (bad) { printf("  variant %d %s: %zu bytes differ from k_gemm_tc\n", variant, sh.name, bad); ++fails; }
        }
      }
      // reference check of variant 0's fp32 / bf16 outputs on the sampled rows (bias only; residual / relu / scale applied on the host)
      k_ref<<<(NR * sh.N + 255) / 256, 256, 0, s>>>(b.A, b.W, b.bias, sh.N, sh.K, drows, NR, dref);
      CK(cudaStreamSynchronize(s));
      std::vector<float> href((size_t)NR * sh.N);
      CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      if (sh.mode == EPI_RESID) {
        std::vector<float> ho((size_t)D), hr((size_t)D);
        for (int i = 0; i < NR; ++i) {
          CK(cudaMemcpy(ho.data(), b.outf[1] + (size_t)hrows[i] * D, D * 4, cudaMemcpyDeviceToHost));
          CK(cudaMemcpy(hr.data(), b.resid + (size_t)hrows[i] * D, D * 4, cudaMemcpyDeviceToHost));
          for (int n = 0; n < D; ++n) { double e = fabs((double)ho[n] - ((double)href[(size_t)i * sh.N + n] + hr[n])); if (e > maxerr) maxerr = e; }
        }
      } else if (sh.mode == EPI_QKV) {
        std::vector<float> ho((size_t)D);
        for (int i = 0; i < NR; ++i) {
          CK(cudaMemcpy(ho.data(), b.outf[1] + (size_t)hrows[i] * D, D * 4, cudaMemcpyDeviceToHost));
          for (int n = 0; n < D; ++n) { double e = fabs((double)ho[n] / QSCALE - (double)href[(size_t)i * sh.N + n]); if (e > maxerr) maxerr = e; }
        }
      } else {
        std::vector<bf16> ho((size_t)sh.N);
        for (int i = 0; i < NR; ++i) {
          CK(cudaMemcpy(ho.data(), b.outb[1] + (size_t)hrows[i] * sh.N, sh.N * 2, cudaMemcpyDeviceToHost));
          for (int n = 0; n < sh.N; ++n) {
            const double ref = fmax(0.0, (double)href[(size_t)i * sh.N + n]);
            double e = fabs((double)__bfloat162float(ho[n]) - ref) - 0.004 * fabs(ref);  // bf16 rounding of the output
            if (e > maxerr) maxerr = e;
          }
        }
      }
      if (maxerr > 2e-3) { printf("  %s: max error against the fp32 reference %.5f\n", sh.name, maxerr); ++fails; }
      printf("M=%6d %-22s  k_gemm_tc<128> %8.1f us %6.1f TF/s | k_gemm_tcp<256> %8.1f us %6.1f TF/s | k_gemm_tcp<128> %8.1f us %6.1f TF/s | ref err %.2e\n", M, sh.name,
             ms_v[0] * 1e3, flop / ms_v[0] * 1e-9, ms_v[1] * 1e3, flop / ms_v[1] * 1e-9, ms_v[2] * 1e3, flop / ms_v[2] * 1e-9, maxerr);
      fflush(stdout);
    }
    cudaFree(b.A); cudaFree(b.W); cudaFree(b.bias); cudaFree(b.resid); cudaFree(b.kvoff); cudaFree(b.err); cudaFree(drows); cudaFree(dref);
    for (int o = 0; o < 2; ++o) { cudaFree(b.outb[o]); cudaFree(b.outf[o]); cudaFree(b.pool[o]); }
  }
  printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
