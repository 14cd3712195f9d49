// gemm_tc.cuh — prefill projections on the 5th-generation tensor cores.
//
//   C[M, N] = A[M, K] (bf16, row-major)  x  W[N, K]^T (bf16, row-major)   with a fused epilogue
//
// One CTA computes a 128 x 128 output tile.  Warp roles (256 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D boxes (128 rows x 64 k) of A and W into a 3-stage
//            shared-memory ring, 128-byte swizzle, completion on mbarriers (expect_tx)
//   warp 1   MMA issuer: ONE elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=128, K=16)
//            from shared-memory descriptors; the accumulator lives in TMEM (128 lanes x 128 fp32 columns);
//            tcgen05.commit releases ring slots and finally signals the epilogue
//   warp 2   TMEM allocation / deallocation
//   warps 4-7 epilogue: tcgen05.ld (32 lanes x 32b x 16 columns) -> registers -> bias / ReLU / residual ->
//            global (fp32 or bf16), or the KV-cache scatter for the QKV projection
//
// Replaces the F.linear calls of T2SBlock.process_prompt (t2s_model.py:142,162) and T2SMLP (:81-84).
// Descriptor encodings follow the PTX ISA tables for tcgen05 shared-memory / instruction descriptors
// (K-major operands, SWIZZLE_128B: 8-row x 128-byte atoms, stride-byte-offset 1024).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace t2s {

// 3 stages x 32 KB: TWO CTAs fit on an SM (and 2 x 128 of the 512 TMEM columns), so one CTA's epilogue overlaps the other's
// MMAs - with K = 512 a tile's main loop is only 8 k-blocks, the prologue and the epilogue are most of a CTA's life
constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 3, TC_THREADS = 256;
template <int BN> struct TcCfg {
  static constexpr int STAGE_BYTES = (TC_BM + BN) * TC_BK * 2;  // 32 KB (BN=128) / 24 KB (BN=64) / 20 KB (BN=32)
  // Prefill (BN = 128, thousands of rows: throughput): 3 stages so that two CTAs share an SM.  Large-batch decode (BN <= 64, M <= 256
  // rows: a handful of CTAs streaming the weights once): the tiles are LATENCY bound - with 3 stages a CTA has 72 KB in flight and
  // the 32 k-blocks of linear2 take 11 HBM round trips - so the ring takes all the shared memory of an SM (8-9 stages, one CTA per SM).
  static constexpr int STAGES = BN >= 128 ? TC_STAGES : (BN == 64 ? 8 : 9);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  // instruction descriptor: D = f32, A = B = bf16, both K-major, N = BN, M = 128
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
};

// Epilogues.  Prefill (explicit LayerNorm rows): EPI_QKV, EPI_RESID, EPI_RELU.  Large-batch decode (LayerNorm
// folded in, same data flow as phases.cuh): EPI_D_QKV, EPI_D_O, EPI_D_FFN1, EPI_D_FFN2.
enum { EPI_QKV = 0, EPI_RESID = 1, EPI_RELU = 2, EPI_D_QKV = 3, EPI_D_O = 4, EPI_D_FFN1 = 5, EPI_D_FFN2 = 6 };

struct TcEpilogue {
  int mode;
  const float* bias;     // [N] (decode: c0 = W beta + bias)
  const float* resid;    // EPI_RESID: [M, N] fp32
  float* out_f32;        // EPI_RESID: [M, N]; EPI_QKV / EPI_D_QKV: q [M, 512]; EPI_D_O: y1; EPI_D_FFN2: y2
  bf16* out_b16;         // EPI_RELU / EPI_D_FFN1: h [M, N]; EPI_D_O: yb1; EPI_D_FFN2: yb2
  bf16* kpool;           // QKV
  bf16* vpool;
  const long long* kvoff;  // [M] element offset of each row's cache position
  size_t layer_off;        // layer * kv_layer_stride
  int* error_flag;
  // ---- large-batch decode
  const int* n_rows;       // device scalar: live rows (<= M); NULL: all M rows
  const float2* sp_in;     // partial statistics of the LayerNorm folded into this GEMM ([M][32]); NULL: none
  const float* c1;         // Wg 1
  float2* stat_out;        // EPI_D_QKV: per-row (mean, rstd) for the O-projection's residual
  float2* sp_out;          // EPI_D_O / EPI_D_FFN2: partial statistics of the produced residual sum
  const float* src_f32;    // EPI_D_O: y2 (or x0 rows when layer 0); EPI_D_FFN2: y1
  const float2* stat_in;   // EPI_D_O: stat2 (NULL for layer 0: residual = src_f32 as is)
  const float2* sp_res;    // EPI_D_FFN2: partial statistics of y1 (LayerNorm of the residual)
  const float* g;          // EPI_D_O: previous norm2 gamma/beta; EPI_D_FFN2: this layer's norm1 gamma/beta
  const float* be;
  const float* b2;         // EPI_D_FFN2: linear2.bias
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a lost arrival becomes an error flag instead of a hung GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* error_flag) {
  for (unsigned spins = 0; spins < 400000000u; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  atomicExch(error_flag, 1);
  return false;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// (mean, rstd) of one row from its 32 per-tile partial (sum, sum of squares)
__device__ __forceinline__ float2 row_stats_from_partials(const float2* sp) {
  float s = 0.f, q = 0.f;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) { const float2 p = __ldcg(sp + i); s += p.x; q += p.y; }
  const float mean = s * (1.0f / D);
  const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
  return make_float2(mean, 1.0f / sqrtf(var + LN_EPS));
}
__device__ __forceinline__ void store_bf16x16(bf16* dst, const float (&x)[16]) {
  uint4* o = reinterpret_cast<uint4*>(dst);
  o[0] = make_uint4(pack_bf2(x[0], x[1]), pack_bf2(x[2], x[3]), pack_bf2(x[4], x[5]), pack_bf2(x[6], x[7]));
  o[1] = make_uint4(pack_bf2(x[8], x[9]), pack_bf2(x[10], x[11]), pack_bf2(x[12], x[13]), pack_bf2(x[14], x[15]));
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N, int K,
          TcEpilogue ep) {
  using CF = TcCfg<BN>;
  extern __shared__ unsigned char tc_smem_raw[];
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  if (ep.n_rows) {  // decode: the live row count is a device value; whole CTAs beyond it have nothing to do
    M = min(M, __ldcg(ep.n_rows));
    if (m0 >= M) return;
  }
  // 1024-byte alignment for the 128B-swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int NST = CF::STAGES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * CF::STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 1);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NST), done_bar = smem_u32(bars + 2 * NST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / TC_BK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NST;
        const uint32_t ph = (kb / NST) & 1;
        if (!mbar_wait(empty0 + 8 * s, ph ^ 1, ep.error_flag)) break;
        const uint32_t sa = smem_u32(smem + s * CF::STAGE_BYTES), sb = sa + TC_BM * TC_BK * 2;
        mbar_expect_tx(full0 + 8 * s, CF::STAGE_BYTES);
        tma_load_2d(sa, &map_a, kb * TC_BK, m0, full0 + 8 * s);
        tma_load_2d(sb, &map_w, kb * TC_BK, n0, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NST;
        const uint32_t ph = (kb / NST) & 1;
        if (!mbar_wait(full0 + 8 * s, ph, ep.error_flag)) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * CF::STAGE_BYTES), sb = sa + TC_BM * TC_BK * 2;
        const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)  // advance 32 bytes (16 bf16) inside the swizzle atom: +2 in the address field
          umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, CF::IDESC, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty0 + 8 * s);  // frees the ring slot once these MMAs have read it
      }
      umma_commit(done_bar);  // accumulator complete
    }
  } else if (warp >= 4) {  // ===== epilogue =====
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + wq * 32 + lane;
    const bool row_ok = row < M;
    // per-row inputs are fetched while the MMAs run
    long long kvo = 0;
    if ((ep.mode == EPI_QKV || ep.mode == EPI_D_QKV) && row_ok && n0 >= D) kvo = ep.kvoff[row];
    float2 st = make_float2(0.f, 1.f), st_res = make_float2(0.f, 1.f);
    if (row_ok) {
      if (ep.sp_in) st = row_stats_from_partials(ep.sp_in + (size_t)row * 32);
      if (ep.mode == EPI_D_O && ep.stat_in) st_res = __ldcg(ep.stat_in + row);
      if (ep.mode == EPI_D_FFN2) st_res = row_stats_from_partials(ep.sp_res + (size_t)row * 32);
      if (ep.mode == EPI_D_QKV && ep.stat_out && n0 == 0) ep.stat_out[row] = st;
    }
    mbar_wait(done_bar, 0, ep.error_flag);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!row_ok) continue;
      const int f0 = n0 + c0;
      float x[16];
      if (ep.sp_in) {  // LayerNorm folded in: rstd * (acc - mean * c1) + c0
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = st.y * (__uint_as_float(v[j]) - st.x * ep.c1[f0 + j]) + ep.bias[f0 + j];
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]) + (ep.bias ? ep.bias[f0 + j] : 0.f);
      }
      if (ep.mode == EPI_QKV || ep.mode == EPI_D_QKV) {
        if (f0 < D) {
          float4* o = reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * D + f0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_float4(x[4 * j] * QSCALE, x[4 * j + 1] * QSCALE, x[4 * j + 2] * QSCALE, x[4 * j + 3] * QSCALE);
        } else {
          {  // 16 features = two 16-byte chunks of one head row; chunks are swizzled by position (common.cuh kv_feat)
            const int fk = (f0 < 2 * D) ? f0 - D : f0 - 2 * D;
            bf16* base = (f0 < 2 * D ? ep.kpool : ep.vpool) + ep.layer_off + (size_t)kvo;
            *reinterpret_cast<uint4*>(base + kv_feat(kvo, fk)) = make_uint4(pack_bf2(x[0], x[1]), pack_bf2(x[2], x[3]), pack_bf2(x[4], x[5]), pack_bf2(x[6], x[7]));
            *reinterpret_cast<uint4*>(base + kv_feat(kvo, fk + 8)) = make_uint4(pack_bf2(x[8], x[9]), pack_bf2(x[10], x[11]), pack_bf2(x[12], x[13]), pack_bf2(x[14], x[15]));
          }
        }
      } else if (ep.mode == EPI_RESID) {
        const float4* r4 = reinterpret_cast<const float4*>(ep.resid + (size_t)row * N + f0);
        float4* o = reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * N + f0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 rr = r4[j];
          o[j] = make_float4(rr.x + x[4 * j], rr.y + x[4 * j + 1], rr.z + x[4 * j + 2], rr.w + x[4 * j + 3]);
        }
      } else if (ep.mode == EPI_RELU || ep.mode == EPI_D_FFN1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
        store_bf16x16(ep.out_b16 + (size_t)row * N + f0, x);
      } else {  // EPI_D_O / EPI_D_FFN2: residual sum y -> fp32, bf16 copy, partial statistics of this 16-feature tile
        const float* src = ep.src_f32 + (size_t)row * D + f0;
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float res = __ldcg(src + j);
          if (ep.mode == EPI_D_FFN2) res = (res - st_res.x) * st_res.y * ep.g[f0 + j] + ep.be[f0 + j] + ep.b2[f0 + j];
          else if (ep.stat_in) res = (res - st_res.x) * st_res.y * ep.g[f0 + j] + ep.be[f0 + j];
          x[j] += res;
          sum += x[j]; sq += x[j] * x[j];
        }
        float4* o = reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * D + f0);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        store_bf16x16(ep.out_b16 + (size_t)row * D + f0, x);
        ep.sp_out[(size_t)row * 32 + (f0 >> 4)] = make_float2(sum, sq);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// LayerNorm (or plain copy) of fp32 rows -> fp32 + bf16 copies, one warp per row: feeds the GEMM's A operand
// and its residual (F.layer_norm, t2s_model.py:165-173).
__global__ void k_ln_rows(const float* __restrict__ in, const float* __restrict__ g, const float* __restrict__ b,
                          float* __restrict__ out_f32, bf16* __restrict__ out_b16, int n_rows, int apply_ln) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* src = in + (size_t)row * D;
  float v[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(src + lane * 4 + 128 * j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
  if (apply_ln) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += v[j];
    const float mean = warp_sum(s) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float d = v[j] - mean; sq += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / D) + LN_EPS);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 gg = *reinterpret_cast<const float4*>(g + lane * 4 + 128 * j);
      const float4 bb = *reinterpret_cast<const float4*>(b + lane * 4 + 128 * j);
      v[4 * j] = (v[4 * j] - mean) * rstd * gg.x + bb.x;
      v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * gg.y + bb.y;
      v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * gg.z + bb.z;
      v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * gg.w + bb.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * D + lane * 4 + 128 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    uint2 p = make_uint2(pack_bf2(v[4 * j], v[4 * j + 1]), pack_bf2(v[4 * j + 2], v[4 * j + 3]));
    *reinterpret_cast<uint2*>(out_b16 + (size_t)row * D + lane * 4 + 128 * j) = p;
  }
}

// ---- host side: TMA tensor maps through the driver entry point (no libcuda link dependency) -------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled g_tmap_encode = nullptr;

static inline bool gemm_tc_init() {
  if (g_tmap_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn)
    return false;
  g_tmap_encode = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
  cudaFuncSetAttribute(k_gemm_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM_BYTES);
  cudaFuncSetAttribute(k_gemm_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::SMEM_BYTES);
  cudaFuncSetAttribute(k_gemm_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32>::SMEM_BYTES);
  return true;
}

// 2-D bf16 row-major [rows, cols] tensor, box = box_rows x 64 cols, 128-byte swizzle, OOB rows read as zero
static inline bool make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, int box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_tmap_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// C = A[M,K] * W[N,K]^T with epilogue `ep`; returns false when the tensor maps cannot be built
template <int BN>
static inline bool launch_gemm_tc(const bf16* A, const bf16* W, int M, int N, int K, const TcEpilogue& ep, cudaStream_t s) {
  CUtensorMap ma, mw;
  if (!make_tmap_bf16(&ma, A, (uint64_t)M, (uint64_t)K, TC_BM) || !make_tmap_bf16(&mw, W, (uint64_t)N, (uint64_t)K, BN)) return false;
  dim3 grid((M + TC_BM - 1) / TC_BM, N / BN);
  k_gemm_tc<BN><<<grid, TC_THREADS, TcCfg<BN>::SMEM_BYTES, s>>>(ma, mw, M, N, K, ep);
  return true;
}

}  // namespace t2s
