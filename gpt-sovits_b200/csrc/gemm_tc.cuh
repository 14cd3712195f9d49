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
  int dbg;                 // measurement only (scripts/mb_gemm.cu): 1 = no global stores, 2 = no staging either, 4 = no prefetch loads, 8 = staged epilogue for every mode
  const int* row_map;      // k_gemm_tcp, EPI_RESID: GEMM row r reads its residual from / writes its result to row row_map[r] (NULL: r)
  long long* tl;           // measurement only: clock64 stamps of k_gemm_tcp, [CTA][tile < 8][8] (scripts/mb_gemm.cu)
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

// =====================================================================================================================
// Persistent variant for the prefill projections (thousands of rows): one CTA per SM walks output tiles of 128 rows x BN
// columns (BN = 256: 384 operand rows per k-block for 256 columns, and an M = 128, N = 256 tcgen05.mma reads 96 B of operands
// per clock from shared memory where N = 128 reads 128 B - the full shared-memory bandwidth), with
//   * a shared-memory operand ring that keeps running ACROSS tiles (the producer is up to STAGES k-blocks ahead, so the
//     next tile's first operands land while the current tile's last MMAs run),
//   * TWO accumulators in TMEM (2 x BN of the 512 columns): the MMAs of tile i+1 run while the epilogue drains tile i,
//   * four epilogue warps (one per TMEM lane quarter; tcgen05.ld 32x32b.x32: the thread holds 32 consecutive columns of ITS row)
//     with two exits.  DIRECT (QKV, linear1): bias, q pre-scale / ReLU + bf16 pack / K/V-page swizzle in registers, then 256-bit
//     stores, one whole 32-byte sector per lane.  STAGED (the residual projections out_proj, linear2, bert_proj): every 32 x 32
//     fp32 block goes through a swizzled shared-memory block so that the residual loads and the stores are 128-byte row segments
//     (the non-persistent kernel wrote one 64-byte piece per thread and row: 32 lines per store instruction); the loads of a
//     block (bias, residual) are requested one block ahead.
// Measured (scripts/mb_gemm.cu): the MMAs of a K = 512 tile are issued in 2.1 us, but SHARED MEMORY is the shared
// resource - per 128 x 256 tile 393 KB of TMA fill + 393 KB of operand reads by the MMAs (+ 2 x 128 KB when the tile is staged)
// against 128 B/clk: 3.1 us (4.1 us staged); the main loop alone runs at 3.5 us per tile (1.25-1.4 PFLOP/s).  Staged, every tile
// drained in ~6.3 us; direct, bf16 tiles drain in 3.3-4 us and fp32 tiles in ~6 us (thread-per-row 256-bit stores sustain 15 B/clk
// per SM, 128-byte segments 21-26: scripts/mb_store.cu).  Built, measured and removed: a weight-stationary variant (a CTA keeps one
// 128-column block of W in shared memory and streams only A: 0.81-0.84 PFLOP/s main loop, N = 128 MMAs read 128 B of operands
// per clock, the whole shared-memory bandwidth); eight epilogue warps instead of four (no change); drain warps (TMEM -> staging)
// and store warps (staging -> global) as separate roles with double-buffered staging (no change); bulk-copy stores from the
// staging block (slower).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2.. epilogue (TMEM lane quarter = warp % 4).
// Epilogues: EPI_QKV, EPI_RESID, EPI_RELU (the three of T2SBlock.process_prompt, t2s_model.py:135-174).
constexpr int TCP_EPI_WARPS = 4, TCP_THREADS = 64 + 32 * TCP_EPI_WARPS;  // 8 (two per lane quarter) measured: no faster, and 204 registers spill
template <int BN> struct TcpCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2, W_BYTES = BN * TC_BK * 2;    // one k-block of A (16 KB) / of W
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;                          // 48 KB (BN = 256) / 32 KB (BN = 128)
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int STG_BYTES = TCP_EPI_WARPS * 32 * 128;                     // one swizzled 32 x 32 fp32 block per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  static constexpr uint32_t TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 256-bit global stores (sm_100): one whole 32-byte sector per lane
__device__ __forceinline__ void st_global_256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(TCP_THREADS, 1)
k_gemm_tcp(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N, int K,
           TcEpilogue ep) {
  using CF = TcpCfg<BN>;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int NST = CF::STAGES;
  unsigned char* stg_all = smem + NST * CF::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_all + CF::STG_BYTES);
  // barriers: full[NST] | empty[NST] | acc_full[2] | acc_empty[2]
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NST), accf0 = smem_u32(bars + 2 * NST), acce0 = smem_u32(bars + 2 * NST + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / TC_BK;
  if (ep.n_rows) M = min(M, __ldcg(ep.n_rows));  // large-batch decode: the live row count is a device value (same in every thread)
  const int n_tiles_n = N / BN, n_tiles = ((M + TC_BM - 1) / TC_BM) * n_tiles_n;
  // tile t = (row tile t / n_tiles_n, column tile t % n_tiles_n): CTAs that run side by side share the rows of A through L2

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accf0 + 8 * a, 1); mbar_init(acce0 + 8 * a, TCP_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(CF::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: k-block counter `it` runs over all tiles of this CTA =====
      unsigned it = 0, lt = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x, ++lt) {
        const int m0 = (t / n_tiles_n) * TC_BM, n0 = (t % n_tiles_n) * BN;
        if (ep.tl && lt < 8) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 0] = clock64();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const unsigned s = it % NST, ph = (it / NST) & 1u;
          if (!mbar_wait(empty0 + 8 * s, ph ^ 1u, ep.error_flag)) { ok = false; break; }
          const uint32_t sa = smem_u32(smem + s * CF::STAGE_BYTES);
          mbar_expect_tx(full0 + 8 * s, CF::STAGE_BYTES);
          tma_load_2d(sa, &map_a, kb * TC_BK, m0, full0 + 8 * s);
          tma_load_2d(sa + CF::A_BYTES, &map_w, kb * TC_BK, n0, full0 + 8 * s);
        }
        if (ep.tl && lt < 8) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 1] = clock64();
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer: accumulator (tile count & 1), columns [acc * BN, acc * BN + BN) =====
      unsigned it = 0, lt = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < n_tiles && ok; t += gridDim.x, ++lt) {
        const unsigned acc = lt & 1u, aph = (lt >> 1) & 1u;
        if (!mbar_wait(acce0 + 8 * acc, aph ^ 1u, ep.error_flag)) break;  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ep.tl && lt < 8) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 2] = clock64();
        const uint32_t tacc = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const unsigned s = it % NST, ph = (it / NST) & 1u;
          if (!mbar_wait(full0 + 8 * s, ph, ep.error_flag)) { ok = false; break; }
          if (ep.tl && lt < 8 && kb == 0) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 3] = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + s * CF::STAGE_BYTES), sb = sa + CF::A_BYTES;
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_bf16_ss(tacc, da + 2 * k, db + 2 * k, CF::IDESC, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty0 + 8 * s);
        }
        if (ok) umma_commit(accf0 + 8 * acc);
        if (ep.tl && lt < 8) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 4] = clock64();
      }
    }
  } else {  // ===== epilogue warps: TMEM lane quarter warp % 4, column share (warp - 2) / 4 =====
    const int wq = warp & 3, ch = (warp - 2) >> 2;
    constexpr int CW = BN / (TCP_EPI_WARPS / 4);  // columns per warp
    // staging block: row r = 128 B = eight 16-byte chunks, chunk c stored at position c ^ (r & 7): the row-per-thread writes, the
    // (row, chunk)-per-lane fp32 reads and the (row, chunk pair)-per-lane bf16 reads are all bank-conflict free
    unsigned char* stg = stg_all + (warp - 2) * 32 * 128;
    // fp32 blocks: lane -> (row r4 + 4 i, chunk c4); bf16 blocks: lane -> (row r8 + 8 i, chunks 2 c8, 2 c8 + 1)
    const int r4 = lane >> 3, c4 = lane & 7, r8 = lane >> 2, c8 = lane & 3;
    // The global loads of a column block (the lane's bias pieces; the residual pieces of EPI_RESID) are requested ONE BLOCK AHEAD - for
    // a tile's first block before the wait on its accumulator - so their latency hides behind the previous block's stores.
    float4 rs[8], bA, bB;  // current block: residual pieces; bias of the lane's columns (fp32 mapping: bA; bf16 mapping: bA, bB)
    auto prefetch = [&](int t, int c0, float4 (&r)[8], float4& b0, float4& b1) {
      const int m0 = (t / n_tiles_n) * TC_BM, n0 = (t % n_tiles_n) * BN;
      const int f0 = n0 + c0, rbase = m0 + wq * 32;
      const bool f32map = ep.mode == EPI_RESID || (ep.mode == EPI_QKV && f0 < D);
      if (!ep.bias) { b0 = make_float4(0.f, 0.f, 0.f, 0.f); b1 = b0; }
      else if (f32map) { b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + f0) + c4); b1 = b0; }
      else { b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + f0) + 2 * c8); b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + f0) + 2 * c8 + 1); }
      if (ep.mode == EPI_RESID) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = rbase + r4 + 4 * i;
          const int grow = (ep.row_map && row < M) ? __ldg(ep.row_map + row) : row;
          r[i] = row < M ? __ldcg(reinterpret_cast<const float4*>(ep.resid + (size_t)grow * N + f0) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    const bool staged = ep.mode == EPI_RESID || (ep.dbg & 8);  // dbg 8: every epilogue through the staging block (measurement)
    if (staged && (int)blockIdx.x < n_tiles) prefetch(blockIdx.x, ch * CW, rs, bA, bB);
    unsigned lt = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++lt) {
      const int m0 = (t / n_tiles_n) * TC_BM, n0 = (t % n_tiles_n) * BN;
      const unsigned acc = lt & 1u, aph = (lt >> 1) & 1u;
      const int rbase = m0 + wq * 32;            // first row of this warp's 32-row band
      const int row_t = rbase + lane;            // row held by this thread in the TMEM phase
      long long kvo = 0;
      if (ep.mode == EPI_QKV && n0 >= D && row_t < M) kvo = ep.kvoff[row_t];
      if (!mbar_wait(accf0 + 8 * acc, aph, ep.error_flag)) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (ep.tl && lt < 8 && warp == 2 && lane == 0) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 5] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c0 = ch * CW; c0 < (ch + 1) * CW; c0 += 32) {
        const int f0 = n0 + c0;
        float4 bd[8];  // direct path: the block's 32 bias values (same address in every lane: one transaction each), requested
        if (!staged) {  // in front of the TMEM load so that both latencies overlap
#pragma unroll
          for (int j = 0; j < 8; ++j) bd[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + f0) + j);
        }
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const bool last = c0 + 32 >= (ch + 1) * CW;
        if (last) {  // this warp's part of the accumulator is read: hand it back to the MMA issuer before the stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(acce0 + 8 * acc);
          if (ep.tl && lt < 8 && warp == 2 && lane == 0) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 6] = clock64();
        }
        if (!staged) {
          // ---- direct path (QKV, ReLU): the thread holds 32 consecutive columns of ITS row (128 B of fp32 / 64 B of bf16) and writes
          // them itself with 256-bit stores, one whole sector per lane: 15 B/clk per SM (scripts/mb_store.cu) against 24 for the
          // re-staged 128-byte segments, but no pass through shared memory, which the operand ring and the MMAs saturate
          float x[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = bd[j];
            x[4 * j] = __uint_as_float(v[4 * j]) + b.x; x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
            x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z; x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
          }
          if ((ep.dbg & 1) || row_t >= M) {
            // no stores for this lane (rows past M); the warp reconverges below before the next (warp-aligned) TMEM load
          } else if (ep.mode == EPI_QKV && f0 < D) {
            float* o = ep.out_f32 + (size_t)row_t * D + f0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_global_256(o + 8 * j, __float_as_uint(x[8 * j] * QSCALE), __float_as_uint(x[8 * j + 1] * QSCALE), __float_as_uint(x[8 * j + 2] * QSCALE),
                            __float_as_uint(x[8 * j + 3] * QSCALE), __float_as_uint(x[8 * j + 4] * QSCALE), __float_as_uint(x[8 * j + 5] * QSCALE),
                            __float_as_uint(x[8 * j + 6] * QSCALE), __float_as_uint(x[8 * j + 7] * QSCALE));
          } else {
            uint32_t pk[16];
            if (ep.mode == EPI_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf2(fmaxf(x[2 * j], 0.f), fmaxf(x[2 * j + 1], 0.f));
              bf16* o = ep.out_b16 + (size_t)row_t * N + f0;
              st_global_256(o, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
              st_global_256(o + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
            } else {
              // k or v: the 32 columns are one head row (64 B) of the K/V page; its 16-byte chunk c lives at position c ^ swz
              // (common.cuh kv_feat), so position p holds chunk p ^ swz: bit 0 of swz swaps the chunks inside a 32-byte half, bit 1
              // swaps the halves
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf2(x[2 * j], x[2 * j + 1]);
              const int swz = kv_swz_of(kvo);
              if (swz & 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { uint32_t tmp = pk[j]; pk[j] = pk[4 + j]; pk[4 + j] = tmp; tmp = pk[8 + j]; pk[8 + j] = pk[12 + j]; pk[12 + j] = tmp; }
              }
              if (swz & 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const uint32_t tmp = pk[j]; pk[j] = pk[8 + j]; pk[8 + j] = tmp; }
              }
              const int fk = f0 < 2 * D ? f0 - D : f0 - 2 * D;
              bf16* o = (f0 < 2 * D ? ep.kpool : ep.vpool) + ep.layer_off + (size_t)kvo + (size_t)(fk >> 5) * KV_HEAD_STRIDE;
              st_global_256(o, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
              st_global_256(o + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
            }
          }
          __syncwarp();
          continue;
        }
        float4 rn[8], bAn = bA, bBn = bB;  // next block's loads, in flight during this block's staging and stores
        const bool more = !last || t + (int)gridDim.x < n_tiles;
        if (more && !(ep.dbg & 4)) prefetch(last ? t + gridDim.x : t, last ? ch * CW : c0 + 32, rn, bAn, bBn);
        if (ep.dbg & 2) continue;
        {
          unsigned char* d = stg + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(d + ((j ^ (lane & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncwarp();
        if (ep.dbg & 1) continue;
        if (ep.mode == EPI_RESID || (ep.mode == EPI_QKV && f0 < D)) {
          float* out = ep.out_f32;
          const int ld = ep.mode == EPI_RESID ? N : D;
          const float sc = ep.mode == EPI_QKV ? QSCALE : 1.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = r4 + 4 * i, row = rbase + rl;
            float4 x = *reinterpret_cast<const float4*>(stg + rl * 128 + ((c4 ^ (rl & 7)) << 4));
            x.x += bA.x; x.y += bA.y; x.z += bA.z; x.w += bA.w;
            if (ep.mode == EPI_RESID) { x.x += rs[i].x; x.y += rs[i].y; x.z += rs[i].z; x.w += rs[i].w; }
            else { x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
            const int grow = (ep.row_map && row < M) ? __ldg(ep.row_map + row) : row;
            if (row < M) *(reinterpret_cast<float4*>(out + (size_t)grow * ld + f0) + c4) = x;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = r8 + 8 * i, row = rbase + rl;
            float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + (((2 * c8) ^ (rl & 7)) << 4));
            float4 b = *reinterpret_cast<const float4*>(stg + rl * 128 + (((2 * c8 + 1) ^ (rl & 7)) << 4));
            a.x += bA.x; a.y += bA.y; a.z += bA.z; a.w += bA.w;
            b.x += bB.x; b.y += bB.y; b.z += bB.z; b.w += bB.w;
            if (ep.mode == EPI_RELU) {
              const uint4 o = make_uint4(pack_bf2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)), pack_bf2(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)),
                                         pack_bf2(fmaxf(b.x, 0.f), fmaxf(b.y, 0.f)), pack_bf2(fmaxf(b.z, 0.f), fmaxf(b.w, 0.f)));
              if (row < M) *(reinterpret_cast<uint4*>(ep.out_b16 + (size_t)row * N + f0) + c8) = o;
            } else {  // EPI_QKV, k or v: the 32 columns are one head row (64 B) of the K/V page, 16-byte chunks swizzled by position
              const long long ko = __shfl_sync(0xffffffffu, kvo, rl);
              const int fk = (f0 < 2 * D ? f0 - D : f0 - 2 * D) + c8 * 8;
              bf16* base = (f0 < 2 * D ? ep.kpool : ep.vpool) + ep.layer_off + (size_t)ko;
              const uint4 o = make_uint4(pack_bf2(a.x, a.y), pack_bf2(a.z, a.w), pack_bf2(b.x, b.y), pack_bf2(b.z, b.w));
              if (row < M) *reinterpret_cast<uint4*>(base + kv_feat(ko, fk)) = o;
            }
          }
        }
        __syncwarp();  // the staging block is rewritten by the next column block
        bA = bAn; bB = bBn;
        if (ep.mode == EPI_RESID) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rs[i] = rn[i];
        }
      }
      if (ep.tl && lt < 8 && warp == 2 && lane == 0) ep.tl[((size_t)blockIdx.x * 8 + lt) * 8 + 7] = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CF::TMEM_COLS) : "memory");
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
  cudaFuncSetAttribute(k_gemm_tcp<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcpCfg<256>::SMEM_BYTES);
  cudaFuncSetAttribute(k_gemm_tcp<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcpCfg<128>::SMEM_BYTES);
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

// persistent variant (prefill) with tensor maps the caller built once (the weight maps never change; an activation map changes only
// with the row count): cuTensorMapEncodeTiled twice per projection is measurable when a prefill of a few hundred rows is launch-bound
template <int BN>
static inline void launch_gemm_tcp_maps(const CUtensorMap& ma, const CUtensorMap& mw, int M, int N, int K, const TcEpilogue& ep, int num_sms,
                                        cudaStream_t s) {
  const int tiles = ((M + TC_BM - 1) / TC_BM) * (N / BN);
  k_gemm_tcp<BN><<<tiles < num_sms ? tiles : num_sms, TCP_THREADS, TcpCfg<BN>::SMEM_BYTES, s>>>(ma, mw, M, N, K, ep);
}

// persistent variant (prefill): grid = min(tiles, SMs), tiles of 128 x BN
template <int BN>
static inline bool launch_gemm_tcp(const bf16* A, const bf16* W, int M, int N, int K, const TcEpilogue& ep, int num_sms, cudaStream_t s) {
  CUtensorMap ma, mw;
  if (!make_tmap_bf16(&ma, A, (uint64_t)M, (uint64_t)K, TC_BM) || !make_tmap_bf16(&mw, W, (uint64_t)N, (uint64_t)K, BN)) return false;
  const int tiles = ((M + TC_BM - 1) / TC_BM) * (N / BN);
  k_gemm_tcp<BN><<<tiles < num_sms ? tiles : num_sms, TCP_THREADS, TcpCfg<BN>::SMEM_BYTES, s>>>(ma, mw, M, N, K, ep);
  return true;
}

}  // namespace t2s
