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
//     stream (393 KB per layer, packed once by k_pack_stream, mma.m16n8k16 A-fragment order, contiguous per warp).  The
//     warps pull the stream into L2 two layers ahead (cp.async.bulk.prefetch.L2) and read their fragments straight from
//     L2 into registers with 128-bit loads, 16 in flight per lane, the first batch of the next matrix issued BEFORE the
//     hand-off wait that precedes it.  (Weights never touch shared memory: a TMA ring doubled the shared-memory traffic
//     and bounded the GEMVs, see DESIGN.md.)
//   * The K/V pages of the head (one contiguous 16 KB block per 128 positions: common.cuh kv_row_off) stream through an
//     8 x 16 KB shared-memory ring fed by TMA bulk copies (K half and V half of a page separately, exact sizes, L2
//     evict-first), running ahead across layers.  Ring slot i belongs to warp i: the warp consumes a page on the tensor
//     cores (mma.sync: S = K q^T, online softmax, O += P V) and issues the copies of the next page that maps to its slot as
//     soon as each half is free, so there is no producer warp, no "empty" barrier and all 8 warps keep 254 registers.
//   * The four hand-offs of a layer (attention out, residual sum 1, FFN hidden, residual sum 2) are all-gathers through
//     DISTRIBUTED SHARED MEMORY: every CTA stages its block once and sends it to each peer with one bulk copy
//     (cp.async.bulk.shared::cluster) that completes bytes on the PEER's mbarrier, so data and "ready" signal travel
//     together (0.4-1.2 us per hand-off instead of a grid barrier + L2 round trip).
//   * Logits go through global memory to one CTA per sequence, which runs the same fused sampler as the other modes
//     (sample_row); ONE grid barrier per step; the retirement plan (CTA 0 compacts the active list and re-deals rows) and
//     its second barrier only run in steps in which a sequence stopped.
// All clusters read the same weight stream at about the same time, so HBM sees the weights once per step and the other
// clusters hit L2 (126 MB).  Everything is deterministic: fixed-order reductions, no floating-point atomics.
#pragma once
#include <type_traits>
#include "phases.cuh"

namespace t2s {
namespace cs {

constexpr int C = 16;            // CTAs per cluster (= heads)
constexpr int RMAX = 8;          // sequences per cluster (one MMA n-tile)
constexpr int NCW = 8;           // warps per CTA
constexpr int SLOT = 16384;      // K/V ring slot: one (page, head) block = K 8 KB | V 8 KB
constexpr int NSLOT = 8;
constexpr int NTC = NCW * 32;
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
  int row_npg[RMAX];                  // cached pages per sequence this step
  alignas(8) uint2 pdesc[RMAX * 32];  // page q of a layer (consumption order) -> (offset of this head's K|V block inside a layer's pool, in 16-byte units; valid bytes per half)
  unsigned long long kfull[NSLOT], vfull_kv[NSLOT], vfull[2], ebar[4], cbar;  // K half / V half of a ring slot land separately
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
// L2 eviction policies: the K/V cache is read once per step (evict first), the weight streams are re-read by every cluster
// (evict last), so the 1.2 GB of K/V that flow through L2 every step do not push the 152 MB of weights out
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint4 ld_weight16_keep(const void* p, uint64_t policy) {  // read-only path, no L1, L2 evict-last
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void bulk_load_hint(void* dst, const void* src, uint32_t bytes, void* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar)), "l"(policy) : "memory");
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

// One bulk-prefetch instruction costs the issuing warp ~0.25 us (measured), whatever its size: keep them few and big, and
// issue them where the warp is about to wait anyway (in front of a hand-off wait).
__device__ __forceinline__ void l2_prefetch(const unsigned char* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
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
  const uint64_t keep = l2_policy_evict_last();
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = ld_weight16_keep(p + i * 32, keep);
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
  const uint64_t keep = l2_policy_evict_last();
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(arow + (k / SKB) * SRCS + (k % SKB) * 16);
    mma_bf16_16816(a[k & 3], buf[k % FB], xr[t], xr[4 + t]);
    if (k + FB < NF) buf[k % FB] = ld_weight16_keep(wp + (k + FB) * 32, keep);
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

// Single-query attention of head `rank` for the cluster's R sequences, on the tensor cores.
// The cached positions [0, pos) arrive through the ring, one 128-position page per slot: K rows at byte i*64, V rows at
// 8192 + i*64, the four 16-byte chunks of a row XOR-swizzled by ((i >> 1) & 3) (common.cuh), so ldmatrix is conflict free.
// A page is processed by ONE warp (ring chunk i -> warp i % 8 = ring slot i % 8):
//   scores   S[128 pos x 8] = K[128 x 32] . Qt[32 x 8]   (m16n8k16, A = K tile via ldmatrix; the 8 columns of Qt alternate
//            bf16 hi / lo halves of the fp32 query, so c0 + c1 of every lane is the score at fp32-query precision)
//   softmax  online, state (m, l) per warp; probabilities rounded to bf16 (and summed rounded: l matches the PV operand)
//   output   O[32 dims x 8] = Vt[32 x 128] . P[128 x 8]   (A = V tile via ldmatrix.trans, all columns of P identical)
// This step's token (k, v still in shared memory) joins in the final merge of the warps' partial states.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint4& r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint4& r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// The K/V pages a CTA reads in one step, in consumption order: layer-major, then sequence, then page.  Page q of the step
// has ring index base + q; ring index i lives in slot i % 8, which belongs to warp i % 8.
struct KvStream { unsigned base; int total, ptot; };  // ring index of the step's first page; pages per step; pages per layer
// lane 0 of the owning warp: start the TMA copy of ring index idx (= page `rem` of layer `layer`, rem < ptot) into its slot.
// sm.pdesc[rem] (block offset, valid bytes) was filled by the step prologue: the request sits on the owning warp's critical
// path (lane 0 alone, the other lanes wait), so everything that does not depend on the layer is computed there, once per step.
__device__ __forceinline__ void kv_issue(const Ctx& c, Smem& sm, unsigned idx, int layer, int rem, uint32_t rank, bool v_half) {
  const uint2 d = sm.pdesc[rem];
  const uint32_t bytes = d.y;  // only the valid rows (the last page is partial)
  const unsigned char* src = reinterpret_cast<const unsigned char*>(c.kpool + (size_t)layer * c.kv_layer_stride) + (size_t)d.x * 16 +
                             (v_half ? KV_V_OFF * 2 : 0);
  const unsigned slot = ring_slot(idx);
  void* bar = v_half ? (void*)&sm.vfull_kv[slot] : (void*)&sm.kfull[slot];
  // The K half and the V half of a slot are copied separately: the next page's K is requested as soon as this page's scores
  // are out, its V after the PV product, so each copy overlaps the other half's compute.  (The half's previous contents were
  // read by this warp's ldmatrix, complete before the MMAs that consumed them; the generic -> async proxy fence for rows
  // appended in earlier steps is executed once per step by the caller.)
  mbar_expect_tx(bar, bytes);
  bulk_load_hint(sm.ring[slot] + (v_half ? PAGE * DH * 2 : 0), src, bytes, bar, l2_policy_evict_first());
}
// ring index idx + 8 -> (layer, page-in-layer) of the next page of this warp's slot; false past the end of the step
__device__ __forceinline__ bool kv_next(const Ctx& c, const KvStream& ks, int layer, int page_in_layer, int& nl, int& nr) {
  nl = layer; nr = page_in_layer + NSLOT;
  while (nr >= ks.ptot) { nr -= ks.ptot; ++nl; }
  return nl < c.n_layer;
}

__device__ __forceinline__ void attention_rows(const Ctx& c, Smem& sm, const KvStream& ks, uint32_t rank, int layer, unsigned& cons,
                                               int R, unsigned char* stage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;  // ldmatrix: lane supplies row r8 of 8x8 matrix mi
  // K tile (16 pos x 16 dims) as A: matrices (pos 0-7 | 8-15) x (dims 0-7 | 8-15) in the order a0..a3
  const int k_pos = (mi & 1) * 8 + r8, k_chunk = mi >> 1;   // + tile*16 positions, + kb*2 chunks
  // V tile as A = Vt (16 dims x 16 pos): a0 = (dims 0-7, pos 0-7), a1 = (dims 8-15, pos 0-7), a2 = (dims 0-7, pos 8-15), a3
  const int v_pos = (mi >> 1) * 8 + r8, v_chunk = mi & 1;   // + kb*16 positions, + mt*2 chunks
  unsigned j = cons;  // ring index of the next page (KvStream order: sequences, then pages)
  for (int n = 0; n < R; ++n) {
    const int pos = sm.row_pos[n];
    const int npg = sm.row_npg[n];
    // ring index i (a whole page) belongs to warp i % 8, the owner of ring slot i % 8: every warp sees the phases of its
    // slot strictly in order (an mbarrier parity wait must never run a phase ahead) and pages are spread evenly
    const int first = (warp - (int)j) & (NCW - 1);
    if (first < npg) {
      // B fragments of the query: k = dims 2t,2t+1 (+8), column g: even -> bf16 hi part, odd -> lo part
      uint32_t qb[2][2];
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float q0 = sm.q[n][kb * 16 + h * 8 + 2 * t], q1 = sm.q[n][kb * 16 + h * 8 + 2 * t + 1];
          const float h0 = bf16_round(q0), h1 = bf16_round(q1);
          qb[kb][h] = (g & 1) ? pack_bf2(q0 - h0, q1 - h1) : pack_bf2(h0, h1);
        }
      float m = -INFINITY, l = 0.f;
      float o[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[mt][i] = 0.f;
      // one page; FULL = all 128 positions valid (every page but a sequence's last): the tile and position predicates fold away
      auto page = [&](int pg, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int np = FULL ? PAGE : min(PAGE, pos - pg * PAGE);
        const int ntile = FULL ? 8 : (np + 15) >> 4;
        const unsigned slot = ring_slot(j + pg);
        mbar_wait(&sm.kfull[slot], ring_par(j + pg));
        int nl, nr;
        const bool more = kv_next(c, ks, layer, (int)(j + pg - cons), nl, nr);
        const uint32_t base = s32(sm.ring[slot]);
        float sc[8][2];
        float mx = -INFINITY;
        {
          // all K fragments first (independent ldmatrix), then the MMAs tile-interleaved: consecutive MMAs are independent
          uint4 ka[8][2];
#pragma unroll
          for (int pt = 0; pt < 8; ++pt) {
            const int p = pt * 16 + k_pos;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
              if (pt < ntile) ldsm_x4(base + p * 64 + (((kb * 2 + k_chunk) ^ ((p >> 1) & 3)) << 4), ka[pt][kb]);
          }
          float c[8][4];
#pragma unroll
          for (int pt = 0; pt < 8; ++pt)
#pragma unroll
            for (int i = 0; i < 4; ++i) c[pt][i] = 0.f;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int pt = 0; pt < 8; ++pt)
              if (pt < ntile) mma_bf16_16816(c[pt], ka[pt][kb], qb[kb][0], qb[kb][1]);
#pragma unroll
          for (int pt = 0; pt < 8; ++pt) {
            sc[pt][0] = (pt * 16 + g < np) ? c[pt][0] + c[pt][1] : -INFINITY;
            sc[pt][1] = (pt * 16 + 8 + g < np) ? c[pt][2] + c[pt][3] : -INFINITY;
            mx = fmaxf(mx, fmaxf(sc[pt][0], sc[pt][1]));
          }
        }
        if (lane == 0 && more) kv_issue(c, sm, j + pg + NSLOT, nl, nr, rank, false);  // K half is free: next page's K
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        const float mn = fmaxf(m, mx);  // finite: np >= 1
        const float corr = fast_exp2(m - mn);
        m = mn;
        float ls = 0.f;
        uint32_t pw[8];  // the probabilities of positions (g, g + 8) of tile pt, rounded to bf16 and packed (lo, hi)
#pragma unroll
        for (int pt = 0; pt < 8; ++pt) {
          pw[pt] = pack_bf2(fast_exp2(sc[pt][0] - mn), fast_exp2(sc[pt][1] - mn));  // one cvt.rn.bf16x2 per pair
          ls += __uint_as_float(pw[pt] << 16) + __uint_as_float(pw[pt] & 0xFFFF0000u);  // the row sum of what the MMA multiplies
        }
        l = l * corr + ls;  // this lane's positions only (g, g+8 of every tile); the quads are summed once per sequence
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[mt][i] *= corr;
        mbar_wait(&sm.vfull_kv[slot], ring_par(j + pg));
        {
          // B fragments of P: positions 16kb + (2t, 2t+1 | 2t+8, 2t+9) are held by the quads g' = 2t and 2t+1
          uint32_t pb[8][2];
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {
            const uint32_t w0 = __shfl_sync(0xffffffffu, pw[kb], 8 * t + t), w1 = __shfl_sync(0xffffffffu, pw[kb], 8 * t + 4 + t);
            pb[kb][0] = __byte_perm(w0, w1, 0x5410);  // positions 16kb + 2t, 2t+1   (the two quads' lo halves)
            pb[kb][1] = __byte_perm(w0, w1, 0x7632);  // positions 16kb + 2t+8, 2t+9 (their hi halves)
          }
          float o2[2][4];  // second accumulator chain per output tile (odd position blocks): shorter dependent MMA chains
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) o2[mt][i] = 0.f;
#pragma unroll
          for (int k2 = 0; k2 < 8; k2 += 2) {
            uint4 va[2][2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int p = (k2 + q) * 16 + v_pos;
#pragma unroll
              for (int mt = 0; mt < 2; ++mt)
                if (k2 + q < ntile) ldsm_x4_t(base + 8192 + p * 64 + (((mt * 2 + v_chunk) ^ ((p >> 1) & 3)) << 4), va[q][mt]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              if (k2 < ntile) mma_bf16_16816(o[mt], va[0][mt], pb[k2][0], pb[k2][1]);
              if (k2 + 1 < ntile) mma_bf16_16816(o2[mt], va[1][mt], pb[k2 + 1][0], pb[k2 + 1][1]);
            }
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) o[mt][i] += o2[mt][i];
        }
        __syncwarp();
        if (lane == 0 && more) kv_issue(c, sm, j + pg + NSLOT, nl, nr, rank, true);  // V half is free: next page's V
      };
      for (int pg = first; pg < npg; pg += NCW) {
        if (pos - pg * PAGE >= PAGE) page(pg, std::true_type{});
        else page(pg, std::false_type{});
      }
      // the quads hold disjoint positions: sum l over g (every column of O already covers all positions of the pages)
      l += __shfl_xor_sync(0xffffffffu, l, 4);
      l += __shfl_xor_sync(0xffffffffu, l, 8);
      l += __shfl_xor_sync(0xffffffffu, l, 16);
      if (t == 0) {
        if (g == 0) { sm.am[warp][n] = m; sm.al[warp][n] = l; }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { sm.aacc[warp][n][mt * 16 + g] = o[mt][0]; sm.aacc[warp][n][mt * 16 + 8 + g] = o[mt][2]; }
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
          if (now - t0 > 4000000000ll || __ldcg(abort_flag) != 0) { atomicCAS(abort_flag, 0, ABORT_WATCHDOG); break; }
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
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm.kfull[s], 1); mbar_init(&sm.vfull_kv[s], 1); }
    for (int k = 0; k < 4; ++k) mbar_init(&sm.ebar[k], 1);
    for (int k = 0; k < 2; ++k) mbar_init(&sm.vfull[k], 1);
    mbar_init(&sm.cbar, C);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stale ring bytes are multiplied by zero probabilities in the PV MMAs: they must be finite
  for (int i = tid; i < NSLOT * SLOT / 16; i += NTC) reinterpret_cast<uint4*>(sm.ring)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  cluster_sync_all();
  const int units_per_step = c.n_layer + 1;  // 24 layers + the head

  // weight stream -> L2: unit u = layer (u % units_per_step) or the head; every warp pulls its eighth of unit u, two units
  // ahead of the compute.  The units of the next step are the same bytes, so running ahead is always legal.
  bool helpers = false;  // this step some clusters have no sequence: THEY prefetch the weight stream (see below), the workers do not
  auto prefetch_unit = [&](int u) {
    if (lane != 0 || helpers) return;
    const int layer = u % units_per_step;
    if (layer < c.n_layer) l2_prefetch(wstream + ((size_t)layer * C + rank) * LAYER_BYTES + (size_t)warp * (LAYER_BYTES / NCW), LAYER_BYTES / NCW);
    else l2_prefetch(hstream + (size_t)rank * HEAD_BYTES + (size_t)warp * (HEAD_BYTES / NCW), HEAD_BYTES / NCW);
  };
  // the layer vectors: 2-deep ring fed by TMA; layer number v (counted over all steps) uses buffer v & 1
  auto vec_issue = [&](unsigned v) {
    if (tid != 0) return;
    // (the buffer's previous readers, two layers back, are separated from this copy by several CTA barriers)
    mbar_expect_tx(&sm.vfull[v & 1u], CH_VEC);
    bulk_load(sm.vec[v & 1u], wstream + ((size_t)(v % (unsigned)c.n_layer) * C + rank) * LAYER_BYTES + OFFS_VEC, CH_VEC, &sm.vfull[v & 1u]);
  };
  prefetch_unit(0);
  prefetch_unit(1);
  vec_issue(0);
  {
    GridBar gbar{c.bar, c.abort_flag, 0u, gridDim.x};
    unsigned cons = 0, vcons = 0;  // ring index of the next K/V page; layers consumed so far
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
    // Retirement bookkeeping (phase_plan by CTA 0 + a second grid barrier) only runs in steps in which a sequence stopped
    // (stop flag in c.seg_cnt[step % 3], set by the sampling CTA); otherwise every sequence just advances by one position
    // and each CTA updates its own copy of the row descriptors.
    int step = ld_cg_i(c.step);
    int n_act = 0, R = 0;
    bool fresh = true;  // row descriptors have to be (re)read from global memory: first step, or after a plan
    for (int it = 0; it < max_new_steps; ++it, ++step) {
      if (fresh) {
        n_act = ld_cg_i(c.n_active);
        R = ((int)cid < n_act) ? (n_act - (int)cid + (int)ncl - 1) / (int)ncl : 0;  // rows r = n*ncl + cid
      }
      if (n_act == 0 || __ldcg(c.abort_flag) != 0) break;
      tl = (c.timeline && it == c.tl_step) ? c.timeline + (size_t)blockIdx.x * c.tl_slots * 2 : nullptr;
      tk = 0;
      CS_TL();
      bool stopped = false;
      helpers = n_act < (int)ncl;
      if (R == 0) {
        // ---- a cluster without a sequence: keep the weight stream three units ahead of the workers in L2.  CTA 0 publishes
        // the unit it starts (the word behind the grid-barrier counter); the idle clusters split every CTA slice of a unit between them.  A bulk prefetch
        // issued by a worker sits in front of its own demand loads and hand-off copies (measured +0.4 us per layer at batch 1).
        const int n_idle = (int)ncl - n_act, my_idle = (int)cid - n_act;
        const int unit_end = unit + units_per_step;
        if (tid == 0) {
          int pf = unit;  // units below `unit` belong to earlier steps (already consumed)
          for (;;) {
            const int p = ld_cg_i(reinterpret_cast<const int*>(c.bar) + 1);
            while (pf < p + 3 && pf < unit_end + 2) {
              const int layer = pf % units_per_step;
              const bool is_layer = layer < c.n_layer;
              const unsigned char* base = is_layer ? wstream + ((size_t)layer * C + rank) * LAYER_BYTES : hstream + (size_t)rank * HEAD_BYTES;
              const int total = is_layer ? LAYER_BYTES : HEAD_BYTES;
              const int piece = ((total / n_idle) + 127) & ~127;
              const int o0 = my_idle * piece, o1 = min(total, o0 + piece);
              for (int o = o0; o < o1; o += 65536) l2_prefetch(base + o, (uint32_t)min(65536, o1 - o));
              ++pf;
            }
            if (p >= unit_end || __ldcg(c.abort_flag) != 0) break;
          }
        }
        unit = unit_end;
      }
      if (R > 0) {
        uint4 wf[FB];  // first fragment batch of the next matrix, in flight across the hand-off that precedes it
        const uint4* lw = reinterpret_cast<const uint4*>(wstream + (size_t)rank * LAYER_BYTES);
        if (warp < NW_QKV) ldg_batch<FB>(lw + wq, wf);
        // ---- step prologue: row descriptors, page-table rows, layer-0 input
        if (fresh) {
          if (tid < R) {
            const int r = tid * ncl + cid;
            sm.row_slot[tid] = ld_cg_i(c.row_slot + r);
            sm.row_pos[tid] = ld_cg_i(c.row_pos + r);
            sm.row_kvoff[tid] = __ldcg(c.row_kvoff + r);
          }
          csync();
          for (int i = tid; i < R * 32; i += NCW * 32) {
            const int n = i >> 5, pg = i & 31;
            sm.pt[n][pg] = (pg < c.max_pages) ? c.page_table[sm.row_slot[n] * c.max_pages + pg] : 0;  // whole row: later steps advance locally
          }
        }
        if (tid < R) sm.row_npg[tid] = (sm.row_pos[tid] + PAGE - 1) >> PAGE_SHIFT;
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
        // ---- this step's K/V page stream; every warp starts the copy of the first page that maps to its ring slot
        KvStream ks;
        ks.base = cons; ks.ptot = 0;
        for (int n = 0; n < R; ++n) {
          const int npg = sm.row_npg[n];
          for (int pg = tid; pg < npg; pg += NCW * 32)
            sm.pdesc[ks.ptot + pg] = make_uint2((uint32_t)((((size_t)sm.pt[n][pg] * KV_PAGE_STRIDE + (size_t)rank * KV_HEAD_STRIDE) * 2) >> 4),
                                                (uint32_t)min(PAGE, sm.row_pos[n] - pg * PAGE) * DH * 2);
          ks.ptot += npg;
        }
        ks.total = ks.ptot * c.n_layer;
        csync();
        if (lane == 0 && ks.ptot > 0) {
          asm volatile("fence.proxy.async;" ::: "memory");  // K/V rows appended by ordinary stores in earlier steps -> TMA reads
          int nl = 0, nr = (warp - cons) & (NSLOT - 1);  // the first page of the step that maps to this warp's slot
          const unsigned first = cons + nr;
          while (nr >= ks.ptot) { nr -= ks.ptot; ++nl; }
          if (nl < c.n_layer) { kv_issue(c, sm, first, nl, nr, rank, false); kv_issue(c, sm, first, nl, nr, rank, true); }
        }
        CS_TL();
        for (int layer = 0; layer < c.n_layer; ++layer) {
          ++unit;
          if (helpers && blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile int*>(c.bar + 1) = unit;  // progress for the prefetching clusters
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
                if (n0 < R) { const long long o = sm.row_kvoff[n0]; pool[o + kv_feat(o, f0)] = h00; pool[o + kv_feat(o, f0 + 8)] = h10; }
                if (n1 < R) { const long long o = sm.row_kvoff[n1]; pool[o + kv_feat(o, f0)] = h01; pool[o + kv_feat(o, f0 + 8)] = h11; }
              }
            }
            csync();
          }
          // ---- attention, then hand the head's output to every peer (exchange 1)
          CS_TL();
          attention_rows(c, sm, ks, rank, layer, cons, R, sm.stage[xch & 1]);
          CS_TL();
          if (tid == 0) mbar_expect_tx(&sm.ebar[0], (uint32_t)(C * R * ROW_S));
          all_gather(s32(sm.stage[xch & 1]), e13_addr, BLK_A, R * ROW_S, rank, eb[0]);
          ++xch;
          prefetch_unit(unit + 1);  // weights two units ahead -> L2; issued where the warp is about to wait anyway
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
          vec_issue(vcons);  // the next layer's vectors (its buffer was last read two layers ago)
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
          csync();
          CS_TL();
        }
        // ---- head: vocabulary tiles rank, rank+16, ... (warps 0..4) -> logits in global memory
        ++unit;
        if (helpers && blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile int*>(c.bar + 1) = unit;
        prefetch_unit(unit + 1);
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
        if ((int)rank < R) {
          stopped = sample_row<1>(c, (int)rank * ncl + cid, step, ss, sm.row_slot[rank]);
          if (stopped && tid == 0) c.seg_cnt[step % 3] = 1;
        }
      }
      CS_TL();
      gbar.sync();
      CS_TL();
      if (ld_cg_i(c.seg_cnt + step % 3) != 0) {
        // somebody stopped: compact the active list, re-deal the rows (t2s_model.py:724-745 does this on the host)
        if (blockIdx.x == 0) phase_plan<1>(c, reinterpret_cast<int*>(sm.e13));
        gbar.sync();
        if (blockIdx.x == 0 && tid == 0) c.seg_cnt[step % 3] = 0;  // every CTA read it before the barrier; next use in 3 steps
        fresh = true;
      } else {
        // nobody stopped: every sequence moves on by one position (what phase_plan would have computed)
        // thread 0 publishes row `rank` from the pre-update position; thread `rank` of the same warp overwrites that entry
        // below: read first, then a warp barrier, so independent thread scheduling cannot reorder the two
        const int pub_pos = ((int)rank < R && tid == 0) ? sm.row_pos[rank] : 0;
        __syncwarp();
        if ((int)rank < R && tid == 0) {
          const int slot = sm.row_slot[rank], pos = pub_pos + 1, r = (int)rank * ncl + cid;
          c.seq_len[slot] = pos + 1;
          c.row_pos[r] = pos;  // kept current for the next launch of a sliced decode (t2s_decode with a step budget)
          c.row_kvoff[r] = kv_row_off(sm.pt[rank][pos >> PAGE_SHIFT], pos & (PAGE - 1));
          atomicAdd(c.stats + 0, (unsigned long long)(pos + 1));
        }
        if (blockIdx.x == 0 && tid == 0) {
          *c.step = step + 1;
          atomicAdd(c.stats + 1, 1ull);
          atomicAdd(c.stats + 2, (unsigned long long)n_act);
        }
        if (tid < R) {
          const int pos = sm.row_pos[tid] + 1;
          sm.row_pos[tid] = pos;
          sm.row_kvoff[tid] = kv_row_off(sm.pt[tid][pos >> PAGE_SHIFT], pos & (PAGE - 1));
        }
        csync();
        fresh = false;
      }
      CS_TL();
    }
#undef CS_TL
    if (tid == 0) mbar_wait(&sm.vfull[vcons & 1u], (vcons >> 1) & 1u);  // the one outstanding vector copy must land before exit
  }
  __syncwarp();
  cluster_sync_all();  // no CTA may exit while a peer can still write into its shared memory
}

}  // namespace cs
}  // namespace t2s
