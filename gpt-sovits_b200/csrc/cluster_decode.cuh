// cluster_decode.cuh — "cluster-stream" decode: the whole decode loop with NO grid-wide barrier inside a step.
//
// Reference being replaced: the per-token loop of infer_panel_batch_infer / infer_panel_naive
// (GPT_SoVITS/AR/models/t2s_model.py:701-769 / :878-914): 24 x T2SBlock.decode_next_token (:176-221), ar_predict_layer
// (:706/:884), sample() (AR/models/utils.py:192) and the retirement bookkeeping (:720-763).
//
// Design (measured motivation in DESIGN.md section 4.4): a decode step is a chain of ~100 dependent tiny GEMVs; with
// the layer split over all 148 SMs every link of the chain costs a grid barrier (~1.2 us) plus an L2 round trip.
// Here a thread-block CLUSTER of C = 16 CTAs (one GPC) owns a few sequences end to end:
//   * CTA `rank` of a cluster is attention head `rank` and owns 1/16 of every weight matrix (32 q/k/v features, 32
//     O-proj outputs, 128 FFN hidden units, 32 FFN2 outputs).  Its weights are ONE private, consumption-ordered byte
//     stream (384 KB per layer, packed once by k_pack_stream) that a producer warp pulls through an 8 x 16 KB shared-
//     memory ring with cp.async.bulk (TMA) + mbarriers, running ahead of the math and across step boundaries.
//     Weights are stored in mma.m16n8k16 A-fragment order, so a consumer warp reads a fragment with one LDS.128.
//   * The four hand-offs of a layer (attention out, residual sum 1, FFN hidden, residual sum 2) are all-gathers through
//     DISTRIBUTED SHARED MEMORY: st.async writes 16-byte pieces into every peer's buffer and completes bytes on the
//     peer's mbarrier, so data and "ready" signal travel together (~0.3 us per hand-off instead of a grid barrier).
//   * K/V of earlier positions are read straight from the paged bf16 cache by the head's CTA; the new position's k/v
//     never leave the CTA before attention (they are also appended to the cache for later steps).
//   * Logits go through global memory to one CTA per sequence, which runs the same fused sampler as the other modes
//     (sample_row), then one grid barrier pair per STEP lets CTA 0 retire finished sequences and re-deal rows.
// All clusters read the same weight stream at about the same time, so HBM sees the weights once per step and the other
// clusters hit L2 (126 MB); per-SM ingest measured at 150-170 GB/s, 17 TB/s aggregate over 7 clusters.
// Everything is deterministic: fixed-order reductions, no floating-point atomics.
#pragma once
#include "phases.cuh"

namespace t2s {
namespace cs {

constexpr int C = 16;            // CTAs per cluster (= heads)
constexpr int RMAX = 8;          // sequences per cluster (one MMA n-tile)
constexpr int NCW = 8;           // consumer warps
constexpr int NTC = (NCW + 1) * 32;  // + one TMA producer warp
constexpr int SLOT = 16384;      // ring slot bytes
constexpr int NSLOT = 8;
constexpr int HD = D / C;        // 32: q/k/v features, O-proj outputs, FFN2 outputs per CTA
constexpr int FH = FF / C;       // 128 FFN hidden units per CTA
// per-layer stream of one CTA: vectors 9,344 B | QKV 8 x 12 KB | [K/V pages of the cluster's sequences: from the cache,
// not from this stream] | Wo 2 x 16 KB | W1 8 x 16 KB | W2 8 x 16 KB
constexpr int CH_QKV = 12288, N_QKV = 8, CH_FULL = 16384, N_WO = 2, N_W1 = 8, N_W2 = 8;
// vector chunk (fp32): biases of this CTA's slices, then the four LayerNorm vectors in full (every CTA normalises whole rows)
constexpr int VC_BQ = 0, VC_BK = 32, VC_BV = 64, VC_BO = 96, VC_B1 = 128, VC_B2 = 256, VC_G1 = 288, VC_BE1 = VC_G1 + D,
              VC_G2 = VC_BE1 + D, VC_BE2 = VC_G2 + D, VC_FLOATS = VC_BE2 + D, CH_VEC = VC_FLOATS * 4;  // 9,344 B
constexpr int OFFS_VEC = 0, OFFS_QKV = CH_VEC, OFFS_WO = OFFS_QKV + N_QKV * CH_QKV, OFFS_W1 = OFFS_WO + N_WO * CH_FULL,
              OFFS_W2 = OFFS_W1 + N_W1 * CH_FULL, LAYER_BYTES = OFFS_W2 + N_W2 * CH_FULL;  // 402,560
constexpr int KV_CHUNK_POS = 2 * PAGE;  // positions per K/V ring chunk: K page pair (8 KB) | V page pair (8 KB)
constexpr int HEAD_TILES = 5;    // 16-row tiles of ar_predict_layer per CTA (80 >= 65)
constexpr int CH_HEAD = HEAD_TILES * 4 * 512, N_HEAD = 8, HEAD_BYTES = CH_HEAD * N_HEAD;  // 81,920
constexpr int XS8 = D + 8;       // bf16 row stride of a 512-wide operand (bank-conflict-free B fragments)
constexpr int HS8 = FF + 8;
static_assert(LAYER_BYTES == (3 * D * D + D * D + 2 * FF * D) * 2 / C + CH_VEC && CH_VEC % 16 == 0, "stream size");

struct __align__(128) Smem {
  unsigned char ring[NSLOT][SLOT];
  unsigned char e13[RMAX * HS8 * 2];  // E1: attention out bf16 [RMAX][XS8]; E3: FFN hidden bf16 [RMAX][HS8]; sampler scratch
  bf16 e24[RMAX][XS8];                // E2 / E4: residual sums (pre-LayerNorm) of all 512 features, bf16
  float2 st24[RMAX][C];               // (sum, sum of squares) of every peer's 32-feature slice (fp32, from unrounded values)
  bf16 xn[RMAX][XS8];                 // LayerNorm'ed rows: operand of QKV / FFN1 / head
  float yown[RMAX][HD];               // own slice of the current residual sum, fp32
  float xres[RMAX][HD];               // own slice of the LayerNorm output (the next residual), fp32
  float q[RMAX][HD];
  bf16 knew[RMAX][HD];
  bf16 vnew[RMAX][HD];
  float red[NCW][16][RMAX + 1];
  unsigned char stage[RMAX * FH * 2]; // outgoing slice, bf16 [R][32] or [R][128]
  float am[NCW][RMAX], al[NCW][RMAX], aacc[NCW][RMAX][HD];
  int pt[RMAX][64];                   // page-table rows of this cluster's sequences
  int row_slot[RMAX], row_pos[RMAX];
  long long row_kvoff[RMAX];
  alignas(16) float vec[2][VC_FLOATS];  // the layer's vectors (biases of the own slices, LayerNorm gamma / beta): 2-deep ring of its own
  unsigned long long full[NSLOT], empty[NSLOT], vfull[2], vempty[2], ebar[4], cbar;
  volatile int stop;
  volatile unsigned consumed, vconsumed;
  volatile int step_seq;  // consumers -> producer: steps whose row descriptors (row_pos, pt, n_rows) are in place
  volatile int n_rows;
};

// ---- PTX helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try(void* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* b, uint32_t parity) { while (!mbar_try(b, parity)) {} }
// wait with cluster-scope acquire: remote st.async data / remote arrivals ordered before the phase completion are visible
__device__ __forceinline__ void mbar_wait_cluster(void* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_idx() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t n_clusters() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_async16(uint32_t addr, const uint4& v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async8(uint32_t addr, float a, float b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1,%2}, [%3];" ::"r"(addr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }  // consumer warps only
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- weight stream layout (shared by the packer and the consumer) ---------------------------------------------------
// Byte `off` of CTA `rank`'s layer stream -> (matrix id, source row, source column) of the bf16 element.
// Inside a chunk: fragment (warp w, i) at (w*4 + i) * 512 B; inside a fragment lane*16 B + j*2 B in m16n8k16 A order:
// lane = g*4+t holds rows g | g+8, cols 2t,2t+1 | +8:  j = 0,1:(g,2t..) 2,3:(g+8,2t..) 4,5:(g,2t+8..) 6,7:(g+8,2t+8..)
__host__ __device__ inline void stream_src(int off, int rank, int& mat, int& row, int& col) {
  int c, rel;  // chunk, byte inside the chunk (callers handle off < OFFS_QKV: the vector chunk)
  if (off < OFFS_WO) { mat = 0; c = (off - OFFS_QKV) / CH_QKV; rel = (off - OFFS_QKV) % CH_QKV; }
  else if (off < OFFS_W1) { mat = 1; c = (off - OFFS_WO) / CH_FULL; rel = (off - OFFS_WO) % CH_FULL; }
  else if (off < OFFS_W2) { mat = 2; c = (off - OFFS_W1) / CH_FULL; rel = (off - OFFS_W1) % CH_FULL; }
  else { mat = 3; c = (off - OFFS_W2) / CH_FULL; rel = (off - OFFS_W2) % CH_FULL; }
  const int frag = rel >> 9, w = frag >> 2, i = frag & 3;
  const int lane = (rel >> 4) & 31, j = (rel >> 1) & 7;
  const int g = lane >> 2, t = lane & 3;
  const int fr = g + ((j & 2) ? 8 : 0);                       // row inside the 16-feature tile
  const int kc = 2 * t + (j & 1) + ((j & 4) ? 8 : 0);         // column inside the 16-wide k-block
  int tile_row0, kb;
  if (mat == 0) {         // warp w < 6: tile w = (q|k|v = w>>1, half = w&1) of head `rank`; k-blocks 4c..4c+3
    tile_row0 = (w >> 1) * D + rank * HD + (w & 1) * 16; kb = 4 * c + i;
  } else if (mat == 1) {  // Wo: tile w&1 of the 32 outputs, K quarter w>>1 (8 k-blocks), chunk c = half of it
    tile_row0 = rank * HD + (w & 1) * 16; kb = (w >> 1) * 8 + 4 * c + i;
  } else if (mat == 2) {  // W1: tile w of the 128 hidden units
    tile_row0 = rank * FH + w * 16; kb = 4 * c + i;
  } else {                // W2: tile w&1 of the 32 outputs, K quarter w>>1 (32 k-blocks)
    tile_row0 = rank * HD + (w & 1) * 16; kb = (w >> 1) * 32 + 4 * c + i;
  }
  row = tile_row0 + fr; col = kb * 16 + kc;
}
// head stream: chunk c, warp w < HEAD_TILES: vocabulary tile (rank + 16 w), k-blocks 4c..4c+3
__host__ __device__ inline void head_src(int off, int rank, int& row, int& col) {
  const int c = off / CH_HEAD, rel = off % CH_HEAD;
  const int frag = rel >> 9, w = frag >> 2, i = frag & 3;
  const int lane = (rel >> 4) & 31, j = (rel >> 1) & 7;
  const int g = lane >> 2, t = lane & 3;
  row = (rank + C * w) * 16 + g + ((j & 2) ? 8 : 0);
  col = (4 * c + i) * 16 + 2 * t + (j & 1) + ((j & 4) ? 8 : 0);
}

// wrow: row-major bf16 layer matrices [n_layer][LW] (OFF_* offsets); wvec: [n_layer][LV] fp32; whead_row: [V][D] bf16
__global__ void k_pack_stream(unsigned char* __restrict__ wstream, bf16* __restrict__ hstream, const bf16* __restrict__ wrow,
                              const float* __restrict__ wvec, const bf16* __restrict__ whead_row, int n_layer) {
  const size_t per_layer = (size_t)C * LAYER_BYTES / 2;  // in 2-byte units
  const size_t n_w = (size_t)n_layer * per_layer, n_h = (size_t)C * HEAD_BYTES / 2;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_w + n_h; e += (size_t)gridDim.x * blockDim.x) {
    if (e < n_w) {
      const int layer = (int)(e / per_layer);
      const size_t r2 = e % per_layer;
      const int rank = (int)(r2 / (LAYER_BYTES / 2)), off = (int)(r2 % (LAYER_BYTES / 2)) * 2;
      if (off < OFFS_QKV) {
        if (off & 3) continue;  // vector chunk: one thread per float
        const int i = off >> 2;
        const float* vl = wvec + (size_t)layer * LV;
        float v;
        if (i < VC_BO) v = vl[VO_BQKV + (i >> 5) * D + rank * HD + (i & 31)];
        else if (i < VC_B1) v = vl[VO_BO + rank * HD + (i - VC_BO)];
        else if (i < VC_B2) v = vl[VO_B1 + rank * FH + (i - VC_B1)];
        else if (i < VC_G1) v = vl[VO_B2 + rank * HD + (i - VC_B2)];
        else if (i < VC_BE1) v = vl[VO_G1 + (i - VC_G1)];
        else if (i < VC_G2) v = vl[VO_BE1 + (i - VC_BE1)];
        else if (i < VC_BE2) v = vl[VO_G2 + (i - VC_G2)];
        else v = vl[VO_BE2 + (i - VC_BE2)];
        *reinterpret_cast<float*>(wstream + e * 2) = v;
        continue;
      }
      int mat, row, col;
      stream_src(off, rank, mat, row, col);
      const bf16* src = wrow + (size_t)layer * LW;
      bf16 v;
      if (mat == 0) v = src[OFF_WQKV + (size_t)row * D + col];
      else if (mat == 1) v = src[OFF_WO + (size_t)row * D + col];
      else if (mat == 2) v = src[OFF_W1 + (size_t)row * D + col];
      else v = src[OFF_W2 + (size_t)row * FF + col];
      *reinterpret_cast<bf16*>(wstream + e * 2) = v;
    } else {
      const size_t r2 = e - n_w;
      const int rank = (int)(r2 / (HEAD_BYTES / 2)), off = (int)(r2 % (HEAD_BYTES / 2)) * 2;
      int row, col;
      head_src(off, rank, row, col);
      hstream[r2] = (row < V) ? whead_row[(size_t)row * D + col] : __float2bfloat16_rn(0.f);
    }
  }
}

// ---- consumer-side building blocks -------------------------------------------------------------------------------------
__device__ __forceinline__ bool mbar_test(void* b, uint32_t parity) {  // non-blocking probe
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned ring_slot(unsigned i) { return i & (NSLOT - 1); }
__device__ __forceinline__ unsigned ring_par(unsigned i) { return (i >> 3) & 1u; }

// One matrix = NCH ring chunks.  Every consumer warp w < NWA takes fragments [4w, 4w+4) of each chunk and multiplies them
// with the activation rows `act` (bf16, row stride `astride` elements) at k-blocks kb0 + 4c + i.  acc = 16 features x
// 8 sequences in the m16n8k16 C layout (c0,c1: feature g, sequences 2t,2t+1; c2,c3: feature g+8).
// All NCH "full" barriers are probed once up front (one lane each); chunks that had already landed need no further wait,
// and the fragments of chunk c+1 are fetched from shared memory while the MMAs of chunk c issue.
template <int NCH, int NWA>
__device__ __forceinline__ void gemv_stream(Smem& sm, unsigned& cons, const bf16* act, int astride, int kb0, float (&acc)[4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
  const bf16* arow = act + (size_t)g * astride + kb0 * 16;
  const unsigned base = cons;
  bool rdy = true;
  if (lane < NCH) rdy = mbar_test(&sm.full[ring_slot(base + lane)], ring_par(base + lane));
  const unsigned ready = __ballot_sync(0xffffffffu, rdy);
  uint4 f[4], fn[4];
  auto fetch = [&](int c, uint4 (&dst)[4]) {
    const unsigned slot = ring_slot(base + c);
    if (!((ready >> c) & 1u)) mbar_wait(&sm.full[slot], ring_par(base + c));
    if (warp < NWA) {
      const uint4* fp = reinterpret_cast<const uint4*>(sm.ring[slot]) + (warp * 4) * 32 + lane;
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = fp[i * 32];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[slot]);
  };
  fetch(0, f);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c + 1 < NCH) fetch(c + 1, fn);
    if (warp < NWA) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t* xr = reinterpret_cast<const uint32_t*>(arow + (4 * c + i) * 16);
        mma_bf16_16816(a[i], f[i], xr[t], xr[4 + t]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = fn[i];
  }
  cons = base + NCH;
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = (a[0][j] + a[1][j]) + (a[2][j] + a[3][j]);
}

// All-gather through DSMEM: every CTA sends `PIECES` 16-byte pieces per row (its slice, staged in sm.stage with row
// stride PIECES*16 B) to every peer's buffer `dst` (row stride dst_stride bytes) at byte column rank*PIECES*16.
template <int PIECES>
__device__ __forceinline__ void all_gather(Smem& sm, uint32_t dst, int dst_stride, int R, uint32_t rank, uint32_t ebar) {
  const int total = R * PIECES * C;
  for (int i = threadIdx.x; i < total; i += NCW * 32) {
    const int peer = i & (C - 1), pc = (i >> 4) % PIECES, n = (i >> 4) / PIECES;
    const uint4 v = *reinterpret_cast<const uint4*>(sm.stage + (n * PIECES + pc) * 16);
    st_async16(mapa(dst + n * dst_stride + (rank * PIECES + pc) * 16, peer), v, mapa(ebar, peer));
  }
}

// LayerNorm of the gathered residual rows (e24 + st24) -> xn (bf16, all features) and xres (fp32, own slice from yown).
// gam / bet: the layer's vectors inside the held ring slot (shared memory).
__device__ __forceinline__ void layer_norm_rows(Smem& sm, int R, uint32_t rank, const float* gam, const float* bet) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < R) {
    const int n = warp;
    const float2 p = (lane < C) ? sm.st24[n][lane] : make_float2(0.f, 0.f);
    const float s = warp_sum(p.x), qq = warp_sum(p.y);
    const float mean = s * (1.0f / D);
    const float var = fmaxf(qq * (1.0f / D) - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k0 = h * 256 + lane * 8;
      const uint4 raw = *reinterpret_cast<const uint4*>(&sm.e24[n][k0]);
      const float4 g0 = *reinterpret_cast<const float4*>(gam + k0), g1 = *reinterpret_cast<const float4*>(gam + k0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(bet + k0), b1 = *reinterpret_cast<const float4*>(bet + k0 + 4);
      uint4 o;
      o.x = pack_bf2((bf_lo(raw.x) - mean) * rstd * g0.x + b0.x, (bf_hi(raw.x) - mean) * rstd * g0.y + b0.y);
      o.y = pack_bf2((bf_lo(raw.y) - mean) * rstd * g0.z + b0.z, (bf_hi(raw.y) - mean) * rstd * g0.w + b0.w);
      o.z = pack_bf2((bf_lo(raw.z) - mean) * rstd * g1.x + b1.x, (bf_hi(raw.z) - mean) * rstd * g1.y + b1.y);
      o.w = pack_bf2((bf_lo(raw.w) - mean) * rstd * g1.z + b1.z, (bf_hi(raw.w) - mean) * rstd * g1.w + b1.w);
      *reinterpret_cast<uint4*>(&sm.xn[n][k0]) = o;
    }
    const int f = rank * HD + lane;  // own slice in fp32 from the unrounded residual sum
    sm.xres[n][lane] = (sm.yown[n][lane] - mean) * rstd * gam[f] + bet[f];
  }
}

// Split-K epilogue of Wo / W2: warp w holds the partial of tile (w&1), K quarter (w>>1).  Reduce the four quarters in a
// fixed order, add bias + residual -> yown (fp32) and the outgoing bf16 slice + partial LayerNorm statistics.
__device__ __forceinline__ void residual_epilogue(Smem& sm, const float (&acc)[4], int R, const float* bias, uint32_t rank,
                                                  uint32_t e24_addr, uint32_t st_addr, uint32_t ebar) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int g = lane >> 2, t = lane & 3;
  sm.red[warp][g][2 * t] = acc[0]; sm.red[warp][g][2 * t + 1] = acc[1];
  sm.red[warp][g + 8][2 * t] = acc[2]; sm.red[warp][g + 8][2 * t + 1] = acc[3];
  csync();
  {
    const int n = tid >> 5, fl = tid & 31;  // warp n = sequence n, lane = feature
    const int tl = fl >> 4, fr = fl & 15;
    float y = 0.f;
    if (n < R) {
      const float s = (sm.red[tl][fr][n] + sm.red[tl + 2][fr][n]) + (sm.red[tl + 4][fr][n] + sm.red[tl + 6][fr][n]);
      y = s + bias[fl] + sm.xres[n][fl];
      sm.yown[n][fl] = y;
      reinterpret_cast<bf16*>(sm.stage)[n * HD + fl] = __float2bfloat16_rn(y);
    }
    const float s1 = warp_sum(y), s2 = warp_sum(y * y);
    if (n < R && fl < C) st_async8(mapa(st_addr + (n * C + rank) * 8, fl), s1, s2, mapa(ebar, fl));
  }
  csync();
  all_gather<HD * 2 / 16>(sm, e24_addr, XS8 * 2, R, rank, ebar);
}

// Single-query attention of head `rank` for the cluster's R sequences.  The cached positions [0, pos) arrive through the
// ring as chunks of up to 128 positions: K rows at byte i*64, V rows at 8192 + i*64 (head-major pages are contiguous).
// Position `pos` (this step's token) comes from shared memory.  Warp w takes positions 16w..16w+15 of a chunk; a quad of
// lanes owns one position (4 x 16 B = the head's 32 dims) and keeps an online-softmax state; quads, then warps are merged.
__device__ __forceinline__ void attention_rows(Smem& sm, unsigned& cons, int R) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane >> 2, part = lane & 3;
  for (int n = 0; n < R; ++n) {
    const int pos = sm.row_pos[n];
    float qv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) qv[j] = sm.q[n][part * 8 + j];
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    auto visit = [&](const uint4& kk, const uint4& vv, bool ok) {
      float s = qv[0] * bf_lo(kk.x) + qv[1] * bf_hi(kk.x) + qv[2] * bf_lo(kk.y) + qv[3] * bf_hi(kk.y) +
                qv[4] * bf_lo(kk.z) + qv[5] * bf_hi(kk.z) + qv[6] * bf_lo(kk.w) + qv[7] * bf_hi(kk.w);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (ok) {
        const float mn = fmaxf(m, s);
        const float corr = exp2f(m - mn), p = exp2f(s - mn);
        m = mn;
        l = l * corr + p;
        acc[0] = acc[0] * corr + p * bf_lo(vv.x); acc[1] = acc[1] * corr + p * bf_hi(vv.x);
        acc[2] = acc[2] * corr + p * bf_lo(vv.y); acc[3] = acc[3] * corr + p * bf_hi(vv.y);
        acc[4] = acc[4] * corr + p * bf_lo(vv.z); acc[5] = acc[5] * corr + p * bf_hi(vv.z);
        acc[6] = acc[6] * corr + p * bf_lo(vv.w); acc[7] = acc[7] * corr + p * bf_hi(vv.w);
      }
    };
    // this step's token: warp 0, quad 0
    visit(*reinterpret_cast<const uint4*>(&sm.knew[n][part * 8]), *reinterpret_cast<const uint4*>(&sm.vnew[n][part * 8]),
          warp == 0 && quad == 0);
    for (int p0 = 0; p0 < pos; p0 += KV_CHUNK_POS) {
      const int np = min(KV_CHUNK_POS, pos - p0);
      const unsigned slot = ring_slot(cons);
      mbar_wait(&sm.full[slot], ring_par(cons));
      const unsigned char* kb = sm.ring[slot] + part * 16;
      uint4 kk[2], vv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = warp * 16 + u * 8 + quad;
        if (i < np) {
          kk[u] = *reinterpret_cast<const uint4*>(kb + i * 64);
          vv[u] = *reinterpret_cast<const uint4*>(kb + 8192 + i * 64);
        } else {
          kk[u] = make_uint4(0, 0, 0, 0); vv[u] = make_uint4(0, 0, 0, 0);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
      ++cons;
#pragma unroll
      for (int u = 0; u < 2; ++u) visit(kk[u], vv[u], warp * 16 + u * 8 + quad < np);
    }
    // merge the 8 quads of the warp (fixed xor tree: deterministic)
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, mo);
      const float ca = (m == -INFINITY) ? 0.f : exp2f(m - mn), cb = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
      l = l * ca + lo * cb;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ao = __shfl_xor_sync(0xffffffffu, acc[j], o);
        acc[j] = acc[j] * ca + ao * cb;
      }
      m = mn;
    }
    if (quad == 0) {
      if (part == 0) { sm.am[warp][n] = m; sm.al[warp][n] = l; }
#pragma unroll
      for (int j = 0; j < 8; ++j) sm.aacc[warp][n][part * 8 + j] = acc[j];
    }
  }
  csync();
  {
    const int n = threadIdx.x >> 5, d = threadIdx.x & 31;  // warp n = sequence n, lane = head dim
    if (n < R) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < NCW; ++w) M = fmaxf(M, sm.am[w][n]);
      float L = 0.f, A = 0.f;
#pragma unroll
      for (int w = 0; w < NCW; ++w) {
        const float mw = sm.am[w][n];
        const float sc = (mw == -INFINITY) ? 0.f : exp2f(mw - M);
        L += sm.al[w][n] * sc;
        A += sm.aacc[w][n][d] * sc;
      }
      reinterpret_cast<bf16*>(sm.stage)[n * HD + d] = __float2bfloat16_rn(A / L);
    }
  }
  csync();
}

// grid barrier among the consumer warps of every CTA (one arrival per CTA)
struct GridBar {
  unsigned* counter; int* abort_flag; unsigned target, ncta;
  __device__ __forceinline__ void sync() {
    asm volatile("fence.proxy.async;" ::: "memory");  // this step's K/V appends are read by TMA (async proxy) in later steps
    csync();
    if (threadIdx.x == 0) {
      target += ncta;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v, spins = 0;
      long long t0 = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) break;
        if ((++spins & 0x3FFu) == 0) {
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          if (now - t0 > 4000000000ll || __ldcg(abort_flag) != 0) { atomicExch(abort_flag, 1); break; }
        }
      }
    }
    csync();
  }
};

// =====================================================================================================================
__global__ void __launch_bounds__(NTC, 1)
k_decode_cluster(Ctx c, const unsigned char* __restrict__ wstream, const unsigned char* __restrict__ hstream, int max_new_steps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank(), cid = cluster_idx(), ncl = n_clusters();
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], NCW); }
    for (int k = 0; k < 4; ++k) mbar_init(&sm.ebar[k], 1);
    for (int k = 0; k < 2; ++k) { mbar_init(&sm.vfull[k], 1); mbar_init(&sm.vempty[k], NCW); }
    mbar_init(&sm.cbar, C);
    sm.stop = 0; sm.consumed = 0; sm.vconsumed = 0; sm.step_seq = 0; sm.n_rows = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();

  if (warp == NCW) {
    // ---------------- producer: this CTA's byte stream (weights + the K/V pages of its head) through the ring ------------
    if (lane == 0) {
      unsigned issued = 0, vissued = 0;
      int steps_done = 0;
      bool run = true;
      auto acquire = [&]() -> bool {  // wait for the next ring slot to be free (or for the stop flag)
        const unsigned slot = ring_slot(issued), par = ring_par(issued) ^ 1u;
        while (!mbar_try(&sm.empty[slot], par)) { if (sm.stop) return false; }
        return true;
      };
      auto push = [&](const unsigned char* src, uint32_t bytes) -> bool {
        if (!acquire()) return false;
        const unsigned slot = ring_slot(issued);
        mbar_expect_tx(&sm.full[slot], bytes);
        bulk_load(sm.ring[slot], src, bytes, &sm.full[slot]);
        ++issued;
        return true;
      };
      while (run) {
        for (int layer = 0; layer < c.n_layer && run; ++layer) {
          const unsigned char* base = wstream + ((size_t)layer * C + rank) * LAYER_BYTES;
          {  // the layer's vectors go to their own 2-deep ring (they are needed during the whole layer)
            const unsigned vs = vissued & 1u, vp = ((vissued >> 1) & 1u) ^ 1u;
            while (!mbar_try(&sm.vempty[vs], vp)) { if (sm.stop) { run = false; break; } }
            if (!run) break;
            mbar_expect_tx(&sm.vfull[vs], CH_VEC);
            bulk_load(sm.vec[vs], base + OFFS_VEC, CH_VEC, &sm.vfull[vs]);
            ++vissued;
          }
          for (int ch = 0; ch < N_QKV && run; ++ch) run = push(base + OFFS_QKV + ch * CH_QKV, CH_QKV);
          if (!run) break;
          if (layer == 0) {  // the K/V list of a step is known once the consumers have read the plan
            while (sm.step_seq <= steps_done) { if (sm.stop) { run = false; break; } }
            if (!run) break;
            asm volatile("fence.proxy.async;" ::: "memory");  // K/V rows appended by generic-proxy stores in earlier steps
          }
          const int R = sm.n_rows;
          const bf16* kl = c.kpool + (size_t)layer * c.kv_layer_stride + (size_t)rank * (PAGE * DH);
          const bf16* vl = c.vpool + (size_t)layer * c.kv_layer_stride + (size_t)rank * (PAGE * DH);
          for (int n = 0; n < R && run; ++n) {
            const int pos = sm.row_pos[n];
            for (int p0 = 0; p0 < pos && run; p0 += KV_CHUNK_POS) {
              if (!(run = acquire())) break;
              const unsigned slot = ring_slot(issued);
              const int n0 = min(PAGE, pos - p0), n1 = min(PAGE, max(pos - p0 - PAGE, 0));
              mbar_expect_tx(&sm.full[slot], (uint32_t)(n0 + n1) * (2 * DH * 2));
              const size_t o0 = (size_t)sm.pt[n][p0 >> 6] * (PAGE * D);
              bulk_load(sm.ring[slot], kl + o0, n0 * DH * 2, &sm.full[slot]);
              bulk_load(sm.ring[slot] + 8192, vl + o0, n0 * DH * 2, &sm.full[slot]);
              if (n1 > 0) {
                const size_t o1 = (size_t)sm.pt[n][(p0 >> 6) + 1] * (PAGE * D);
                bulk_load(sm.ring[slot] + 4096, kl + o1, n1 * DH * 2, &sm.full[slot]);
                bulk_load(sm.ring[slot] + 8192 + 4096, vl + o1, n1 * DH * 2, &sm.full[slot]);
              }
              ++issued;
            }
          }
          for (int ch = 0; ch < N_WO + N_W1 + N_W2 && run; ++ch) run = push(base + OFFS_WO + ch * CH_FULL, CH_FULL);
        }
        const unsigned char* hb = hstream + (size_t)rank * HEAD_BYTES;
        for (int ch = 0; ch < N_HEAD && run; ++ch) run = push(hb + (size_t)ch * CH_HEAD, CH_HEAD);
        ++steps_done;
      }
      // drain: every copy that was issued but never consumed must land before the CTA may exit
      for (unsigned i = sm.consumed; i < issued; ++i) mbar_wait(&sm.full[ring_slot(i)], ring_par(i));
      for (unsigned i = sm.vconsumed; i < vissued; ++i) mbar_wait(&sm.vfull[i & 1u], (i >> 1) & 1u);
    }
  } else {
    // ---------------- consumers ------------------------------------------------------------------------------------------
    GridBar gbar{c.bar, c.abort_flag, 0u, gridDim.x};
    unsigned cons = 0, vcons = 0;
    uint32_t epar = 0;  // bit k = parity of exchange barrier k
    uint32_t cpar = 0;
    const uint32_t e13_addr = s32(sm.e13), e24_addr = s32(sm.e24), st_addr = s32(sm.st24);
    const uint32_t eb[4] = {s32(&sm.ebar[0]), s32(&sm.ebar[1]), s32(&sm.ebar[2]), s32(&sm.ebar[3])};
    const int g = lane >> 2, t = lane & 3;
    SampSmem& ss = *reinterpret_cast<SampSmem*>(sm.e13);
    long long* tl = nullptr;  // measurement hook: clock stamps of thread 0 at the markers of one step
    int tk = 0;
#define CS_TL() do { if (tl && tid == 0 && tk < 2 * c.tl_slots) tl[tk++] = clock64(); } while (0)
    for (int it = 0; it < max_new_steps; ++it) {
      const int n_act = ld_cg_i(c.n_active);
      if (n_act == 0 || __ldcg(c.abort_flag) != 0) break;
      tl = (c.timeline && it == c.tl_step) ? c.timeline + (size_t)blockIdx.x * c.tl_slots * 2 : nullptr;
      tk = 0;
      CS_TL();
      const int R = ((int)cid < n_act) ? (n_act - (int)cid + (int)ncl - 1) / (int)ncl : 0;  // rows r = n*ncl + cid
      if (R > 0) {
        // ---- step prologue: row descriptors, page-table rows, layer-0 input
        if (tid < R) {
          const int r = tid * ncl + cid;
          sm.row_slot[tid] = ld_cg_i(c.row_slot + r);
          sm.row_pos[tid] = ld_cg_i(c.row_pos + r);
          sm.row_kvoff[tid] = __ldcg(c.row_kvoff + r);
        }
        csync();
        for (int i = tid; i < R * 64; i += NCW * 32) {
          const int n = i >> 6, pg = i & 63;
          sm.pt[n][pg] = (pg <= (sm.row_pos[n] >> 6)) ? c.page_table[sm.row_slot[n] * c.max_pages + pg] : 0;
        }
        if (warp < R) {
          const int n = warp;
          const float* xr = c.x0 + (size_t)sm.row_slot[n] * D;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k0 = h * 256 + lane * 8;
            const float4 a = ld_cg_f4(xr + k0), b = ld_cg_f4(xr + k0 + 4);
            *reinterpret_cast<uint4*>(&sm.xn[n][k0]) =
                make_uint4(pack_bf2(a.x, a.y), pack_bf2(a.z, a.w), pack_bf2(b.x, b.y), pack_bf2(b.z, b.w));
          }
          sm.xres[n][lane] = ld_cg_f(xr + rank * HD + lane);
        }
        csync();
        if (tid == 0) { sm.n_rows = R; __threadfence_block(); sm.step_seq = sm.step_seq + 1; }  // producer may list this step's K/V
        CS_TL();
        for (int layer = 0; layer < c.n_layer; ++layer) {
          // ---- the layer's vectors (own 2-deep ring; normally long landed)
          const unsigned vslot = vcons & 1u;
          mbar_wait(&sm.vfull[vslot], (vcons >> 1) & 1u);
          ++vcons;
          const float* vec = sm.vec[vslot];
          // ---- QKV for head `rank` (warps 0..5: q lo/hi, k lo/hi, v lo/hi)
          {
            float acc[4];
            gemv_stream<N_QKV, 6>(sm, cons, &sm.xn[0][0], XS8, 0, acc);
            CS_TL();
            if (warp < 6) {
              const int ty = warp >> 1, f0 = (warp & 1) * 16 + g;
              const float b0 = vec[VC_BQ + ty * HD + f0], b1 = vec[VC_BQ + ty * HD + f0 + 8];
              const float v00 = acc[0] + b0, v01 = acc[1] + b0, v10 = acc[2] + b1, v11 = acc[3] + b1;  // (feature, sequence 2t | 2t+1)
              const int n0 = 2 * t, n1 = 2 * t + 1;
              if (ty == 0) {
                sm.q[n0][f0] = v00 * QSCALE; sm.q[n1][f0] = v01 * QSCALE;
                sm.q[n0][f0 + 8] = v10 * QSCALE; sm.q[n1][f0 + 8] = v11 * QSCALE;
              } else {
                bf16 (*dst)[HD] = (ty == 1) ? sm.knew : sm.vnew;
                bf16* pool = ((ty == 1) ? c.kpool : c.vpool) + (size_t)layer * c.kv_layer_stride + (size_t)rank * (PAGE * DH);
                const bf16 h00 = __float2bfloat16_rn(v00), h01 = __float2bfloat16_rn(v01), h10 = __float2bfloat16_rn(v10),
                           h11 = __float2bfloat16_rn(v11);
                dst[n0][f0] = h00; dst[n1][f0] = h01; dst[n0][f0 + 8] = h10; dst[n1][f0 + 8] = h11;
                if (n0 < R) { pool[sm.row_kvoff[n0] + f0] = h00; pool[sm.row_kvoff[n0] + f0 + 8] = h10; }
                if (n1 < R) { pool[sm.row_kvoff[n1] + f0] = h01; pool[sm.row_kvoff[n1] + f0 + 8] = h11; }
              }
            }
            csync();
          }
          // ---- attention, then hand the head's output to every peer (exchange 1)
          CS_TL();
          attention_rows(sm, cons, R);
          CS_TL();
          if (tid == 0) mbar_expect_tx(&sm.ebar[0], (uint32_t)R * D * 2);
          all_gather<HD * 2 / 16>(sm, e13_addr, XS8 * 2, R, rank, eb[0]);
          mbar_wait_cluster(&sm.ebar[0], (epar >> 0) & 1u); epar ^= 1u;
          CS_TL();
          // ---- O-projection (32 outputs, split-K over warp pairs) + bias + residual -> exchange 2
          {
            float acc[4];
            gemv_stream<N_WO, NCW>(sm, cons, reinterpret_cast<const bf16*>(sm.e13), XS8, (warp >> 1) * 8, acc);
            CS_TL();
            if (tid == 0) mbar_expect_tx(&sm.ebar[1], (uint32_t)R * (D * 2 + C * 8));
            residual_epilogue(sm, acc, R, vec + VC_BO, rank, e24_addr, st_addr, eb[1]);
            CS_TL();
          }
          mbar_wait_cluster(&sm.ebar[1], (epar >> 1) & 1u); epar ^= 2u;
          CS_TL();
          layer_norm_rows(sm, R, rank, vec + VC_G1, vec + VC_BE1);
          csync();
          CS_TL();
          // ---- FFN1 (128 hidden units, one tile per warp) + bias + ReLU -> exchange 3
          {
            float acc[4];
            gemv_stream<N_W1, NCW>(sm, cons, &sm.xn[0][0], XS8, 0, acc);
            CS_TL();
            const int f0 = warp * 16 + g;
            const float b0 = vec[VC_B1 + f0], b1 = vec[VC_B1 + f0 + 8];
            bf16* hs = reinterpret_cast<bf16*>(sm.stage);
            hs[(2 * t) * FH + f0] = __float2bfloat16_rn(fmaxf(acc[0] + b0, 0.f));
            hs[(2 * t + 1) * FH + f0] = __float2bfloat16_rn(fmaxf(acc[1] + b0, 0.f));
            hs[(2 * t) * FH + f0 + 8] = __float2bfloat16_rn(fmaxf(acc[2] + b1, 0.f));
            hs[(2 * t + 1) * FH + f0 + 8] = __float2bfloat16_rn(fmaxf(acc[3] + b1, 0.f));
            csync();
            if (tid == 0) mbar_expect_tx(&sm.ebar[2], (uint32_t)R * FF * 2);
            all_gather<FH * 2 / 16>(sm, e13_addr, HS8 * 2, R, rank, eb[2]);
            mbar_wait_cluster(&sm.ebar[2], (epar >> 2) & 1u); epar ^= 4u;
            CS_TL();
          }
          // ---- FFN2 (32 outputs, K = 2048 split over warp pairs) + bias + residual -> exchange 4
          {
            float acc[4];
            gemv_stream<N_W2, NCW>(sm, cons, reinterpret_cast<const bf16*>(sm.e13), HS8, (warp >> 1) * 32, acc);
            CS_TL();
            if (tid == 0) mbar_expect_tx(&sm.ebar[3], (uint32_t)R * (D * 2 + C * 8));
            residual_epilogue(sm, acc, R, vec + VC_B2, rank, e24_addr, st_addr, eb[3]);
          }
          mbar_wait_cluster(&sm.ebar[3], (epar >> 3) & 1u); epar ^= 8u;
          CS_TL();
          layer_norm_rows(sm, R, rank, vec + VC_G2, vec + VC_BE2);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.vempty[vslot]);  // done with the layer's vectors
          csync();
          CS_TL();
        }
        // ---- head: vocabulary tiles rank, rank+16, ... (warps 0..4) -> logits in global memory
        {
          float acc[4];
          gemv_stream<N_HEAD, HEAD_TILES>(sm, cons, &sm.xn[0][0], XS8, 0, acc);
          CS_TL();
          if (warp < HEAD_TILES) {
            const int f0 = ((int)rank + C * warp) * 16 + g;
            const int n0 = 2 * t, n1 = 2 * t + 1;
            if (n0 < R) {
              float* lg = c.logits + (size_t)(n0 * ncl + cid) * VPAD;
              if (f0 < V) lg[f0] = acc[0];
              if (f0 + 8 < V) lg[f0 + 8] = acc[2];
            }
            if (n1 < R) {
              float* lg = c.logits + (size_t)(n1 * ncl + cid) * VPAD;
              if (f0 < V) lg[f0] = acc[1];
              if (f0 + 8 < V) lg[f0 + 8] = acc[3];
            }
          }
        }
        // ---- cluster barrier that also orders the global logits writes (release / acquire at cluster scope)
        csync();
        if (tid < C) mbar_arrive_remote(mapa(s32(&sm.cbar), tid));
        mbar_wait_cluster(&sm.cbar, cpar); cpar ^= 1u;
        CS_TL();
        // ---- sampler: CTA `rank` takes the cluster's sequence `rank`
        if ((int)rank < R) sample_row<1>(c, (int)rank * ncl + cid, ld_cg_i(c.step), ss);
      }
      CS_TL();
      gbar.sync();
      CS_TL();
      if (blockIdx.x == 0) phase_plan<1>(c, reinterpret_cast<int*>(sm.e13));
      gbar.sync();
      CS_TL();
    }
#undef CS_TL
    if (tid == 0) { sm.consumed = cons; sm.vconsumed = vcons; __threadfence_block(); sm.stop = 1; }
  }
  __syncwarp();
  cluster_sync_all();  // no CTA may exit while a peer can still write into its shared memory
}

}  // namespace cs
}  // namespace t2s
