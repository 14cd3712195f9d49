// cluster_decode.cuh — "cluster-stream" decode: the whole decode loop with NO grid-wide barrier inside a step.
//
// Reference being replaced: the per-token loop of infer_panel_batch_infer / infer_panel_naive
// (GPT_SoVITS/AR/models/t2s_model.py:701-769 / :878-914): 24 x T2SBlock.decode_next_token (:176-221), ar_predict_layer
// (:706/:884), sample() (AR/models/utils.py:192) and the retirement bookkeeping (:720-763).
//
// Design (measurements in DESIGN.md): a decode step is a chain of ~100 dependent tiny GEMVs; with a layer split over all
// 148 SMs every link of the chain costs a grid barrier (~1.2 us) plus an L2 round trip.  Here a thread-block CLUSTER of
// C = 16 CTAs (one GPC) owns a few sequences end to end:
//   * CTA `rank` of a cluster is attention head `rank` and owns 1/16 of every weight matrix (32 q/k/v features, 32
//     O-proj outputs, 128 FFN hidden units, 32 FFN2 outputs).  Its weights are ONE private, consumption-ordered byte
//     stream (393 KB per layer, packed once by k_pack_stream, mma.m16n8k16 A-fragment order, contiguous per warp).  A
//     prefetch lane pulls the stream into L2 two layers ahead (cp.async.bulk.prefetch.L2); the consumer warps read their
//     fragments straight from L2 into registers with 128-bit loads, 8 in flight per lane, the first batch of the next
//     matrix issued BEFORE the hand-off wait that precedes it.  (Weights never touch shared memory: a TMA ring doubled
//     the shared-memory traffic and bounded the GEMVs, see DESIGN.md.)
//   * The K/V pages of the head (one contiguous 16 KB block per 128 positions: common.cuh kv_row_off) stream through an
//     8 x 16 KB shared-memory ring fed by TMA bulk copies (one copy per page), running ahead across layers.
//   * The four hand-offs of a layer (attention out, residual sum 1, FFN hidden, residual sum 2) are all-gathers through
//     DISTRIBUTED SHARED MEMORY: st.async writes 16-byte pieces into every peer's buffer and completes bytes on the
//     peer's mbarrier, so data and "ready" signal travel together (~0.3 us per hand-off instead of a grid barrier).
//   * Logits go through global memory to one CTA per sequence, which runs the same fused sampler as the other modes
//     (sample_row); one grid barrier pair per STEP lets CTA 0 retire finished sequences and re-deal rows.
// All clusters read the same weight stream at about the same time, so HBM sees the weights once per step and the other
// clusters hit L2 (126 MB).  Everything is deterministic: fixed-order reductions, no floating-point atomics.
#pragma once
#include "phases.cuh"

namespace t2s {
namespace cs {

constexpr int C = 16;            // CTAs per cluster (= heads)
constexpr int RMAX = 8;          // sequences per cluster (one MMA n-tile)
constexpr int NCW = 8;           // consumer warps
constexpr int SLOT = 16384;      // K/V ring slot: one (page, head) block = K 8 KB | V 8 KB
constexpr int NSLOT = 8;
constexpr int NPW = 2;           // producer warps: K/V TMA ring; weight L2 prefetch + the layer-vector ring
constexpr int NTC = (NCW + NPW) * 32;
constexpr int HD = D / C;        // 32: q/k/v features, O-proj outputs, FFN2 outputs per CTA
constexpr int FH = FF / C;       // 128 FFN hidden units per CTA
// vector chunk (fp32): biases of this CTA's slices, then the four LayerNorm vectors in full (every CTA normalises whole rows)
constexpr int VC_BQ = 0, VC_BK = 32, VC_BV = 64, VC_BO = 96, VC_B1 = 128, VC_B2 = 256, VC_G1 = 288, VC_BE1 = VC_G1 + D,
              VC_G2 = VC_BE1 + D, VC_BE2 = VC_G2 + D, VC_FLOATS = VC_BE2 + D, CH_VEC = VC_FLOATS * 4;  // 9,344 B
// per-layer stream of one CTA: vectors | QKV (6 warps x 32 fragments) | Wo (8 x 8) | W1 (8 x 32) | W2 (8 x 32); a fragment
// is 512 B (32 lanes x 16 B); inside a matrix the fragments of one warp are contiguous (the warp streams them in order)
constexpr int NF_QKV = 32, NW_QKV = 6, NF_WO = 8, NF_W1 = 32, NF_W2 = 32;
constexpr int SZ_QKV = NW_QKV * NF_QKV * 512, SZ_WO = NCW * NF_WO * 512, SZ_W1 = NCW * NF_W1 * 512, SZ_W2 = NCW * NF_W2 * 512;
constexpr int OFFS_VEC = 0, OFFS_QKV = CH_VEC, OFFS_WO = OFFS_QKV + SZ_QKV, OFFS_W1 = OFFS_WO + SZ_WO, OFFS_W2 = OFFS_W1 + SZ_W1,
              LAYER_BYTES = OFFS_W2 + SZ_W2;  // 402,560
constexpr int HEAD_TILES = 5;    // 16-row tiles of ar_predict_layer per CTA (80 >= 65)
constexpr int NF_HEAD = 32, HEAD_BYTES = HEAD_TILES * NF_HEAD * 512;  // 81,920
constexpr int XS8 = D + 8;       // bf16 row stride of the local 512-wide operand xn (bank-conflict-free B fragments)
// Hand-off buffers are laid out per SOURCE CTA so that a source's whole contribution is one contiguous block = one bulk
// copy: block = [optional 64 B of per-row statistics][RMAX rows x (slice + 16 B pad)].  The pad makes the rows of a block
// fall into distinct banks for the MMA B-fragment loads (row stride 80 B / 272 B).
constexpr int ROW_S = HD * 2 + 16;             // 80 B: a 32-feature bf16 slice row
constexpr int ROW_H = FH * 2 + 16;             // 272 B: a 128-feature bf16 slice row (FFN hidden)
constexpr int BLK_A = RMAX * ROW_S;            // 640 B: attention-out block of one source
constexpr int BLK_H = RMAX * ROW_H;            // 2176 B: FFN-hidden block of one source
constexpr int BLK_ST = RMAX * 8;               // 64 B: (sum, sum of squares) per row, in front of a residual block
constexpr int BLK_Y = BLK_ST + RMAX * ROW_S;   // 704 B: residual-sum block of one source
static_assert(LAYER_BYTES == (3 * D * D + D * D + 2 * FF * D) * 2 / C + CH_VEC && CH_VEC % 16 == 0, "stream size");
static_assert(NSLOT == NCW, "ring slot i is read by warp i");
static_assert(SLOT == KV_HEAD_STRIDE * 2, "one ring slot = one (page, head) K|V block");

struct __align__(128) Smem {
  unsigned char ring[NSLOT][SLOT];
  alignas(16) unsigned char e13[C * BLK_H];  // E1: attention out, C blocks of BLK_A; E3: FFN hidden, C blocks of BLK_H; sampler scratch
  alignas(16) unsigned char e24[C * BLK_Y];  // E2 / E4: residual sums (pre-LayerNorm), C blocks of BLK_Y (statistics + bf16 rows)
  bf16 xn[RMAX][XS8];                 // LayerNorm'ed rows: operand of QKV / FFN1 / head
  float yown[RMAX][HD];               // own slice of the current residual sum, fp32
  float xres[RMAX][HD];               // own slice of the LayerNorm output (the next residual), fp32
  float q[RMAX][HD];
  bf16 knew[RMAX][HD];
  bf16 vnew[RMAX][HD];
  float red[NCW][16][RMAX + 1];
  alignas(16) unsigned char stage[2][BLK_H];  // outgoing block (same layout as the destination block), double buffered
  float am[NCW][RMAX], al[NCW][RMAX], aacc[NCW][RMAX][HD];
  int pt[RMAX][32];                   // page-table rows of this cluster's sequences (<= 32 pages of 128 positions)
  int row_slot[RMAX], row_pos[RMAX];
  long long row_kvoff[RMAX];
  alignas(16) float vec[2][VC_FLOATS];  // the layer's vectors (biases of the own slices, LayerNorm gamma / beta): 2-deep ring of its own
  unsigned long long full[NSLOT], empty[NSLOT], vfull[2], vempty[2], ebar[4], cbar;
  volatile int stop;
  volatile unsigned consumed, vconsumed;
  volatile int step_seq;  // consumers -> producers: steps whose row descriptors (row_pos, pt, n_rows) are in place
  volatile int unit_seq;  // consumers -> prefetch lane: (layer | head) units started so far
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
// shared::cta -> (remote) shared::cluster bulk copy; completes `bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }  // consumer warps only
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- weight stream layout (shared by the packer and the consumer) ---------------------------------------------------
// Byte `off` (>= OFFS_QKV) of CTA `rank`'s layer stream -> (matrix id, source row, source column) of the bf16 element.
// Inside a matrix: fragment (warp w, k) at (w*NF + k) * 512 B; inside a fragment lane*16 B + j*2 B in m16n8k16 A order:
// lane = g*4+t holds rows g | g+8, cols 2t,2t+1 | +8:  j = 0,1:(g,2t..) 2,3:(g+8,2t..) 4,5:(g,2t+8..) 6,7:(g+8,2t+8..)
__host__ __device__ inline void frag_elem(int rel, int nf, int& w, int& k, int& fr, int& kc) {
  const int frag = rel >> 9;
  w = frag / nf; k = frag % nf;
  const int lane = (rel >> 4) & 31, j = (rel >> 1) & 7;
  const int g = lane >> 2, t = lane & 3;
  fr = g + ((j & 2) ? 8 : 0);                    // row inside the 16-feature tile
  kc = 2 * t + (j & 1) + ((j & 4) ? 8 : 0);      // column inside the 16-wide k-block
}
__host__ __device__ inline void stream_src(int off, int rank, int& mat, int& row, int& col) {
  int w, k, fr, kc;
  if (off < OFFS_WO) {         // QKV: warp w < 6 = tile (q|k|v = w>>1, half = w&1) of head `rank`, k-block k
    mat = 0; frag_elem(off - OFFS_QKV, NF_QKV, w, k, fr, kc);
    row = (w >> 1) * D + rank * HD + (w & 1) * 16 + fr; col = k * 16 + kc;
  } else if (off < OFFS_W1) {  // Wo: tile w&1 of the 32 outputs, K quarter w>>1 (8 k-blocks)
    mat = 1; frag_elem(off - OFFS_WO, NF_WO, w, k, fr, kc);
    row = rank * HD + (w & 1) * 16 + fr; col = ((w >> 1) * 8 + k) * 16 + kc;
  } else if (off < OFFS_W2) {  // W1: tile w of the 128 hidden units
    mat = 2; frag_elem(off - OFFS_W1, NF_W1, w, k, fr, kc);
    row = rank * FH + w * 16 + fr; col = k * 16 + kc;
  } else {                     // W2: tile w&1 of the 32 outputs, K quarter w>>1 (32 k-blocks)
    mat = 3; frag_elem(off - OFFS_W2, NF_W2, w, k, fr, kc);
    row = rank * HD + (w & 1) * 16 + fr; col = ((w >> 1) * 32 + k) * 16 + kc;
  }
}
// head stream: warp w < HEAD_TILES: vocabulary tile (rank + 16 w), k-block k
__host__ __device__ inline void head_src(int off, int rank, int& row, int& col) {
  int w, k, fr, kc;
  frag_elem(off, NF_HEAD, w, k, fr, kc);
  row = (rank + C * w) * 16 + fr; col = k * 16 + kc;
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
__device__ __forceinline__ unsigned ring_slot(unsigned i) { return i & (NSLOT - 1); }
__device__ __forceinline__ unsigned ring_par(unsigned i) { return (i >> 3) & 1u; }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Weight fragments come straight from L2 (prefetched there two layers ahead) into a rolling buffer of FB 128-bit
// registers per lane: fragment k+FB is requested the moment fragment k has been fed to the tensor core, so FB loads stay
// in flight per lane (8 KB per warp, 64 KB per SM) for the whole matrix.
constexpr int FB = 16;
template <int N>
__device__ __forceinline__ void ldg_batch(const uint4* p, uint4 (&f)[FB]) {  // p = warp's fragment base + lane
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = ld_weight16(p + i * 32);
}
// One matrix slice of this warp = NF fragments (k-blocks kb0 .. kb0+NF-1 of one 16-feature tile); buf = its first
// min(NF, FB) fragments (already in flight / landed).  acc = 16 features x 8 sequences in the m16n8k16 C layout (c0,c1:
// feature g, sequences 2t,2t+1; c2,c3: feature g+8).  act: bf16 activation rows (row stride astride elements).
// The activation element (sequence g, k-block kb) lives at act + g*ROWS + (kb / SKB)*SRCS + (kb % SKB)*16 (elements): SKB
// k-blocks per source block, SRCS elements between source blocks (local operand xn: one "source", ROWS = XS8).
template <int NF, int SKB, int SRCS, int ROWS>
__device__ __forceinline__ void gemv_ldg(const uint4* wp, uint4 (&buf)[FB], const bf16* act, int kb0, float (&acc)[4]) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
  const bf16* arow = act + g * ROWS + (kb0 / SKB) * SRCS;  // kb0 is a multiple of SKB
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(arow + (k / SKB) * SRCS + (k % SKB) * 16);
    mma_bf16_16816(a[k & 3], buf[k % FB], xr[t], xr[4 + t]);
    if (k + FB < NF) buf[k % FB] = ld_weight16(wp + (k + FB) * 32);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = (a[0][j] + a[1][j]) + (a[2][j] + a[3][j]);
}

// All-gather through DSMEM: this CTA's block (staged in `src`, written by generic stores that the callers fenced for the
// async proxy and ordered by a CTA barrier) goes to slot `rank` of every peer's buffer `dst` with one bulk copy per peer;
// the copy completes its bytes on the peer's mbarrier, so data and "ready" travel together.  Two lanes of every warp
// issue (one copy each).
__device__ __forceinline__ void all_gather(uint32_t src, uint32_t dst, uint32_t blk, uint32_t bytes, uint32_t rank, uint32_t ebar) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < C / NCW) {
    const uint32_t peer = warp * (C / NCW) + lane;
    bulk_s2s(mapa(dst + rank * blk, peer), src, bytes, mapa(ebar, peer));
  }
}

// LayerNorm of the gathered residual rows (e24 + st24) -> xn (bf16, all features) and xres (fp32, own slice from yown).
__device__ __forceinline__ void layer_norm_rows(Smem& sm, int R, uint32_t rank, const float* gam, const float* bet) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < R) {
    const int n = warp;
    const float2 p = (lane < C) ? *reinterpret_cast<const float2*>(sm.e24 + lane * BLK_Y + n * 8) : make_float2(0.f, 0.f);
    const float s = warp_sum(p.x), qq = warp_sum(p.y);
    const float mean = s * (1.0f / D);
    const float var = fmaxf(qq * (1.0f / D) - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k0 = h * 256 + lane * 8;
      // features k0..k0+7 = source block k0/32, 16-byte piece (lane & 3) of its row n
      const uint4 raw = *reinterpret_cast<const uint4*>(sm.e24 + (k0 >> 5) * BLK_Y + BLK_ST + n * ROW_S + (lane & 3) * 16);
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
                                                  unsigned char* stage, uint32_t e24_addr, uint32_t ebar) {
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
      *reinterpret_cast<bf16*>(stage + BLK_ST + n * ROW_S + fl * 2) = __float2bfloat16_rn(y);
    }
    const float s1 = warp_sum(y), s2 = warp_sum(y * y);
    if (n < R && fl == 0) *reinterpret_cast<float2*>(stage + n * 8) = make_float2(s1, s2);
  }
  fence_proxy_async_smem();
  csync();
  all_gather(s32(stage), e24_addr, BLK_Y, BLK_ST + R * ROW_S, rank, ebar);
}

// Single-query attention of head `rank` for the cluster's R sequences.  The cached positions [0, pos) arrive through the
// ring, one 128-position page per slot: K rows at byte i*64, V rows at 8192 + i*64.  Position `pos` (this step's token)
// comes from shared memory.  A page is processed by ONE warp (ring chunk i -> warp i % 8 = ring slot i % 8): a quad of lanes
// owns positions quad, quad+8, ... (4 x 16 B = the head's 32 dims each), scores four of them, then folds them into its
// online-softmax state in one update; the quads of a warp, then the warps are merged in a fixed order.
__device__ __forceinline__ float dot8(const float (&q)[8], const uint4& k) {
  return q[0] * bf_lo(k.x) + q[1] * bf_hi(k.x) + q[2] * bf_lo(k.y) + q[3] * bf_hi(k.y) + q[4] * bf_lo(k.z) + q[5] * bf_hi(k.z) +
         q[6] * bf_lo(k.w) + q[7] * bf_hi(k.w);
}
__device__ __forceinline__ void axpy8(float (&acc)[8], float p, const uint4& v) {
  acc[0] += p * bf_lo(v.x); acc[1] += p * bf_hi(v.x); acc[2] += p * bf_lo(v.y); acc[3] += p * bf_hi(v.y);
  acc[4] += p * bf_lo(v.z); acc[5] += p * bf_hi(v.z); acc[6] += p * bf_lo(v.w); acc[7] += p * bf_hi(v.w);
}
__device__ __forceinline__ void attention_rows(Smem& sm, unsigned& cons, int R, unsigned char* stage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane >> 2, part = lane & 3;
  unsigned j = cons;  // ring index of the next page (same order as the producer: sequences, then pages)
  for (int n = 0; n < R; ++n) {
    const int pos = sm.row_pos[n];
    const int npg = (pos + PAGE - 1) >> PAGE_SHIFT;
    // ring chunk i (a whole page) belongs to warp i % 8 = the warp that always reads ring slot i % 8: every warp sees the
    // phases of its slot strictly in order (an mbarrier parity wait must never run a phase ahead), pages are spread evenly
    const int first = (warp - (int)j) & (NCW - 1);
    if (first < npg) {
      float qv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) qv[i] = sm.q[n][part * 8 + i];
      float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int pg = first; pg < npg; pg += NCW) {
        const int np = min(PAGE, pos - pg * PAGE);
        const unsigned slot = ring_slot(j + pg);
        mbar_wait(&sm.full[slot], ring_par(j + pg));
        const unsigned char* kb = sm.ring[slot] + part * 16;
        for (int u0 = 0; u0 * 8 < np; u0 += 4) {  // 4 positions per quad per round: 32 positions per warp round
          uint4 kk[4], vv[4];
          bool ok[4];
          float sc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = (u0 + u) * 8 + quad;
            ok[u] = i < np;
            const int ic = ok[u] ? i : 0;
            kk[u] = *reinterpret_cast<const uint4*>(kb + ic * 64);
            vv[u] = *reinterpret_cast<const uint4*>(kb + 8192 + ic * 64);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) sc[u] = dot8(qv, kk[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 1);
#pragma unroll
          for (int u = 0; u < 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 2);
          float mn = m;
#pragma unroll
          for (int u = 0; u < 4; ++u) if (ok[u]) mn = fmaxf(mn, sc[u]);
          if (ok[0]) {  // ok[u] implies ok[0]
            const float corr = fast_exp2(m - mn);
            float e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = ok[u] ? fast_exp2(sc[u] - mn) : 0.f;
            m = mn;
            l = l * corr + (e[0] + e[1]) + (e[2] + e[3]);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= corr;
#pragma unroll
            for (int u = 0; u < 4; ++u) axpy8(acc, e[u], vv[u]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
      }
      // merge the 8 quads of the warp (fixed xor tree: deterministic)
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
        const float mn = fmaxf(m, mo);
        const float ca = (m == -INFINITY) ? 0.f : fast_exp2(m - mn), cb = (mo == -INFINITY) ? 0.f : fast_exp2(mo - mn);
        l = l * ca + lo * cb;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ao = __shfl_xor_sync(0xffffffffu, acc[i], o);
          acc[i] = acc[i] * ca + ao * cb;
        }
        m = mn;
      }
      if (quad == 0) {
        if (part == 0) { sm.am[warp][n] = m; sm.al[warp][n] = l; }
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.aacc[warp][n][part * 8 + i] = acc[i];
      }
    } else if (lane == 0) {
      sm.am[warp][n] = -INFINITY;  // no page of this sequence for this warp
    }
    j += npg;
  }
  cons = j;
  csync();
  {
    const int n = threadIdx.x >> 5, d = threadIdx.x & 31;  // warp n = sequence n, lane = head dim
    if (n < R) {
      // this step's token (k, v still in shared memory), then the warps' partial states in a fixed order
      const float s_new = warp_sum(sm.q[n][d] * __bfloat162float(sm.knew[n][d]));
      float M = s_new;
#pragma unroll
      for (int w = 0; w < NCW; ++w) M = fmaxf(M, sm.am[w][n]);
      const float e_new = fast_exp2(s_new - M);
      float L = e_new, A = e_new * __bfloat162float(sm.vnew[n][d]);
#pragma unroll
      for (int w = 0; w < NCW; ++w) {
        const float mw = sm.am[w][n];
        if (mw != -INFINITY) {
          const float scl = fast_exp2(mw - M);
          L += sm.al[w][n] * scl;
          A += sm.aacc[w][n][d] * scl;
        }
      }
      *reinterpret_cast<bf16*>(stage + n * ROW_S + d * 2) = __float2bfloat16_rn(A / L);
    }
  }
  fence_proxy_async_smem();
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

__device__ __forceinline__ void l2_prefetch(const unsigned char* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// =====================================================================================================================
__global__ void __launch_bounds__(NTC, 1)
k_decode_cluster(Ctx c, const unsigned char* __restrict__ wstream, const unsigned char* __restrict__ hstream, int max_new_steps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank(), cid = cluster_idx(), ncl = n_clusters();
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }  // a page is consumed by ONE warp
    for (int k = 0; k < 4; ++k) mbar_init(&sm.ebar[k], 1);
    for (int k = 0; k < 2; ++k) { mbar_init(&sm.vfull[k], 1); mbar_init(&sm.vempty[k], NCW); }
    mbar_init(&sm.cbar, C);
    sm.stop = 0; sm.consumed = 0; sm.vconsumed = 0; sm.step_seq = 0; sm.unit_seq = 0; sm.n_rows = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  const int units_per_step = c.n_layer + 1;  // 24 layers + the head

  if (warp == NCW) {
    // ---------------- producer 1: the K/V pages of head `rank` of this cluster's sequences, through the ring --------------
    if (lane == 0) {
      unsigned issued = 0;
      int steps_done = 0;
      bool run = true;
      while (run) {
        while (sm.step_seq <= steps_done) { if (sm.stop) { run = false; break; } }  // rows of the step known?
        if (!run) break;
        asm volatile("fence.proxy.async;" ::: "memory");  // K/V rows appended by generic-proxy stores in earlier steps
        const int R = sm.n_rows;
        for (int layer = 0; layer < c.n_layer && run; ++layer) {
          const bf16* kl = c.kpool + (size_t)layer * c.kv_layer_stride + (size_t)rank * KV_HEAD_STRIDE;
          for (int n = 0; n < R && run; ++n) {
            const int pos = sm.row_pos[n];
            for (int p0 = 0; p0 < pos; p0 += PAGE) {
              const unsigned slot = ring_slot(issued), par = ring_par(issued) ^ 1u;
              while (!mbar_try(&sm.empty[slot], par)) { if (sm.stop) { run = false; break; } }
              if (!run) break;
              const uint32_t bytes = (uint32_t)(PAGE * DH * 2 + min(PAGE, pos - p0) * DH * 2);  // K block + the valid V rows
              mbar_expect_tx(&sm.full[slot], bytes);
              bulk_load(sm.ring[slot], kl + (size_t)sm.pt[n][p0 >> PAGE_SHIFT] * KV_PAGE_STRIDE, bytes, &sm.full[slot]);
              ++issued;
            }
          }
        }
        ++steps_done;
      }
      // drain: every copy that was issued but never consumed must land before the CTA may exit
      for (unsigned i = sm.consumed; i < issued; ++i) mbar_wait(&sm.full[ring_slot(i)], ring_par(i));
    }
  } else if (warp == NCW + 1) {
    // ---------------- producer 2: weight stream -> L2 two units ahead; the layer vectors -> their own 2-deep ring ----------
    if (lane == 0) {
      unsigned vissued = 0;
      bool run = true;
      for (int u = 0; run; ++u) {  // unit u = layer (u % units_per_step), or the head
        while (sm.unit_seq + 2 < u) { if (sm.stop) { run = false; break; } }
        if (!run) break;
        const int layer = u % units_per_step;
        if (layer < c.n_layer) {
          const unsigned char* base = wstream + ((size_t)layer * C + rank) * LAYER_BYTES;
          for (int o = 0; o < LAYER_BYTES; o += 32768) l2_prefetch(base + o, (uint32_t)min(32768, LAYER_BYTES - o));
          const unsigned vs = vissued & 1u, vp = ((vissued >> 1) & 1u) ^ 1u;
          while (!mbar_try(&sm.vempty[vs], vp)) { if (sm.stop) { run = false; break; } }
          if (!run) break;
          mbar_expect_tx(&sm.vfull[vs], CH_VEC);
          bulk_load(sm.vec[vs], base + OFFS_VEC, CH_VEC, &sm.vfull[vs]);
          ++vissued;
        } else {
          const unsigned char* hb = hstream + (size_t)rank * HEAD_BYTES;
          for (int o = 0; o < HEAD_BYTES; o += 32768) l2_prefetch(hb + o, (uint32_t)min(32768, HEAD_BYTES - o));
        }
      }
      for (unsigned i = sm.vconsumed; i < vissued; ++i) mbar_wait(&sm.vfull[i & 1u], (i >> 1) & 1u);
    }
  } else {
    // ---------------- consumers ------------------------------------------------------------------------------------------
    GridBar gbar{c.bar, c.abort_flag, 0u, gridDim.x};
    unsigned cons = 0, vcons = 0;
    uint32_t epar = 0;  // bit k = parity of exchange barrier k
    uint32_t cpar = 0;
    const uint32_t e13_addr = s32(sm.e13), e24_addr = s32(sm.e24);
    unsigned xch = 0;  // hand-offs so far: staging buffer xch & 1
    const uint32_t eb[4] = {s32(&sm.ebar[0]), s32(&sm.ebar[1]), s32(&sm.ebar[2]), s32(&sm.ebar[3])};
    const int g = lane >> 2, t = lane & 3;
    SampSmem& ss = *reinterpret_cast<SampSmem*>(sm.e13);
    long long* tl = nullptr;  // measurement hook: clock stamps of thread 0 at the markers of one step
    int tk = 0;
    int unit = 0;
#define CS_TL() do { if (tl && tid == 0 && tk < 2 * c.tl_slots) tl[tk++] = clock64(); } while (0)
    // this warp's fragment streams inside a layer block (uint4 units), + lane
    const size_t wq = (OFFS_QKV + (size_t)warp * NF_QKV * 512) / 16 + lane, wo = (OFFS_WO + (size_t)warp * NF_WO * 512) / 16 + lane,
                 w1 = (OFFS_W1 + (size_t)warp * NF_W1 * 512) / 16 + lane, w2 = (OFFS_W2 + (size_t)warp * NF_W2 * 512) / 16 + lane;
    for (int it = 0; it < max_new_steps; ++it) {
      const int n_act = ld_cg_i(c.n_active);
      if (n_act == 0 || __ldcg(c.abort_flag) != 0) break;
      tl = (c.timeline && it == c.tl_step) ? c.timeline + (size_t)blockIdx.x * c.tl_slots * 2 : nullptr;
      tk = 0;
      CS_TL();
      const int R = ((int)cid < n_act) ? (n_act - (int)cid + (int)ncl - 1) / (int)ncl : 0;  // rows r = n*ncl + cid
      if (R > 0) {
        uint4 wf[FB];  // first fragment batch of the next matrix, in flight across the hand-off that precedes it
        const uint4* lw = reinterpret_cast<const uint4*>(wstream + (size_t)rank * LAYER_BYTES);
        if (warp < NW_QKV) ldg_batch<FB>(lw + wq, wf);
        // ---- step prologue: row descriptors, page-table rows, layer-0 input
        if (tid < R) {
          const int r = tid * ncl + cid;
          sm.row_slot[tid] = ld_cg_i(c.row_slot + r);
          sm.row_pos[tid] = ld_cg_i(c.row_pos + r);
          sm.row_kvoff[tid] = __ldcg(c.row_kvoff + r);
        }
        csync();
        for (int i = tid; i < R * 32; i += NCW * 32) {
          const int n = i >> 5, pg = i & 31;
          sm.pt[n][pg] = (pg <= (sm.row_pos[n] >> PAGE_SHIFT)) ? c.page_table[sm.row_slot[n] * c.max_pages + pg] : 0;
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
        if (tid == 0) { sm.n_rows = R; __threadfence_block(); sm.step_seq = sm.step_seq + 1; }  // the K/V producer may list this step
        CS_TL();
        for (int layer = 0; layer < c.n_layer; ++layer) {
          if (tid == 0) sm.unit_seq = ++unit;
          // ---- the layer's vectors (own 2-deep ring; normally long landed)
          const unsigned vslot = vcons & 1u;
          mbar_wait(&sm.vfull[vslot], (vcons >> 1) & 1u);
          ++vcons;
          const float* vec = sm.vec[vslot];
          // ---- QKV for head `rank` (warps 0..5: q lo/hi, k lo/hi, v lo/hi)
          {
            float acc[4];
            if (warp < NW_QKV) gemv_ldg<NF_QKV, NF_QKV, 0, XS8>(lw + wq, wf, &sm.xn[0][0], 0, acc);
            ldg_batch<NF_WO>(lw + wo, wf);  // Wo's only batch: in flight during the attention
            CS_TL();
            if (warp < NW_QKV) {
              const int ty = warp >> 1, f0 = (warp & 1) * 16 + g;
              const float b0 = vec[VC_BQ + ty * HD + f0], b1 = vec[VC_BQ + ty * HD + f0 + 8];
              const float v00 = acc[0] + b0, v01 = acc[1] + b0, v10 = acc[2] + b1, v11 = acc[3] + b1;  // (feature, sequence 2t | 2t+1)
              const int n0 = 2 * t, n1 = 2 * t + 1;
              if (ty == 0) {
                sm.q[n0][f0] = v00 * QSCALE; sm.q[n1][f0] = v01 * QSCALE;
                sm.q[n0][f0 + 8] = v10 * QSCALE; sm.q[n1][f0 + 8] = v11 * QSCALE;
              } else {
                bf16 (*dst)[HD] = (ty == 1) ? sm.knew : sm.vnew;
                bf16* pool = ((ty == 1) ? c.kpool : c.vpool) + (size_t)layer * c.kv_layer_stride + (size_t)rank * KV_HEAD_STRIDE;
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
          attention_rows(sm, cons, R, sm.stage[xch & 1]);
          CS_TL();
          if (tid == 0) mbar_expect_tx(&sm.ebar[0], (uint32_t)(C * R * ROW_S));
          all_gather(s32(sm.stage[xch & 1]), e13_addr, BLK_A, R * ROW_S, rank, eb[0]);
          ++xch;
          mbar_wait_cluster(&sm.ebar[0], (epar >> 0) & 1u); epar ^= 1u;
          CS_TL();
          // ---- O-projection (32 outputs, split-K over warp pairs) + bias + residual -> exchange 2
          {
            float acc[4];
            gemv_ldg<NF_WO, 2, BLK_A / 2, ROW_S / 2>(lw + wo, wf, reinterpret_cast<const bf16*>(sm.e13), (warp >> 1) * 8, acc);
            ldg_batch<FB>(lw + w1, wf);  // FFN1's first batch: in flight during the epilogue, hand-off 2 and LayerNorm 1
            CS_TL();
            if (tid == 0) mbar_expect_tx(&sm.ebar[1], (uint32_t)(C * (BLK_ST + R * ROW_S)));
            residual_epilogue(sm, acc, R, vec + VC_BO, rank, sm.stage[xch & 1], e24_addr, eb[1]);
            ++xch;
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
            gemv_ldg<NF_W1, NF_W1, 0, XS8>(lw + w1, wf, &sm.xn[0][0], 0, acc);
            ldg_batch<FB>(lw + w2, wf);  // FFN2's first batch: in flight during hand-off 3
            CS_TL();
            const int f0 = warp * 16 + g;
            const float b0 = vec[VC_B1 + f0], b1 = vec[VC_B1 + f0 + 8];
            bf16* hs = reinterpret_cast<bf16*>(sm.stage[xch & 1]);  // row stride ROW_H bytes
            hs[(2 * t) * (ROW_H / 2) + f0] = __float2bfloat16_rn(fmaxf(acc[0] + b0, 0.f));
            hs[(2 * t + 1) * (ROW_H / 2) + f0] = __float2bfloat16_rn(fmaxf(acc[1] + b0, 0.f));
            hs[(2 * t) * (ROW_H / 2) + f0 + 8] = __float2bfloat16_rn(fmaxf(acc[2] + b1, 0.f));
            hs[(2 * t + 1) * (ROW_H / 2) + f0 + 8] = __float2bfloat16_rn(fmaxf(acc[3] + b1, 0.f));
            fence_proxy_async_smem();
            csync();
            if (tid == 0) mbar_expect_tx(&sm.ebar[2], (uint32_t)(C * R * ROW_H));
            all_gather(s32(sm.stage[xch & 1]), e13_addr, BLK_H, R * ROW_H, rank, eb[2]);
            ++xch;
            mbar_wait_cluster(&sm.ebar[2], (epar >> 2) & 1u); epar ^= 4u;
            CS_TL();
          }
          // ---- FFN2 (32 outputs, K = 2048 split over warp pairs) + bias + residual -> exchange 4
          {
            float acc[4];
            gemv_ldg<NF_W2, 8, BLK_H / 2, ROW_H / 2>(lw + w2, wf, reinterpret_cast<const bf16*>(sm.e13), (warp >> 1) * 32, acc);
            // next unit's first batch (QKV of the next layer, or the head) in flight during hand-off 4 and LayerNorm 2
            if (layer + 1 < c.n_layer) {
              lw += (size_t)C * LAYER_BYTES / 16;
              if (warp < NW_QKV) ldg_batch<FB>(lw + wq, wf);
            } else if (warp < HEAD_TILES) {
              ldg_batch<FB>(reinterpret_cast<const uint4*>(hstream + (size_t)rank * HEAD_BYTES) + (size_t)warp * NF_HEAD * 32 + lane, wf);
            }
            CS_TL();
            if (tid == 0) mbar_expect_tx(&sm.ebar[3], (uint32_t)(C * (BLK_ST + R * ROW_S)));
            residual_epilogue(sm, acc, R, vec + VC_B2, rank, sm.stage[xch & 1], e24_addr, eb[3]);
            ++xch;
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
        if (tid == 0) sm.unit_seq = ++unit;
        {
          float acc[4];
          if (warp < HEAD_TILES) {
            gemv_ldg<NF_HEAD, NF_HEAD, 0, XS8>(reinterpret_cast<const uint4*>(hstream + (size_t)rank * HEAD_BYTES) + (size_t)warp * NF_HEAD * 32 + lane, wf,
                                               &sm.xn[0][0], 0, acc);
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
          CS_TL();
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
