// wide_decode.cuh — "wide" decode for SMALL batches (<= 8 sequences): one decode step spread over 144 SMs, with no grid
// barrier inside the 24 layers.
//
// Reference being replaced: the per-token loop of infer_panel_naive / infer_panel_batch_infer
// (GPT_SoVITS/AR/models/t2s_model.py:878-914 / :701-769): 24 x T2SBlock.decode_next_token (:176-221), ar_predict_layer
// (:884/:706), sample() (AR/models/utils.py:192) and the stop / retirement bookkeeping.
//
// Why another kernel.  At batch 1 a decode step reads 152 MB of weights and ~25 MB of K/V: 27 us at the HBM roofline.  The
// cluster-stream kernel (cluster_decode.cuh) gives a sequence to ONE 16-CTA cluster, so only 16 SMs ingest weights (~62 B/clk
// each from L2: >= 80 us per step) and every layer is a chain of 13 dependent links (259 us per step measured, 10 % of the
// roofline).  Here every SM works on every step:
//   * grid = 16 heads x 9 CTAs (144 of 148 SMs, cooperative launch: all CTAs co-resident).  CTA (h, s) owns: the query
//     projection of head h (all 9 CTAs of a head compute it redundantly: 32 KB of weights, deduplicated by L2, so that
//     nobody waits for anybody before attention), slice s of head h's K/V positions (split-KV), and fixed tiles of the other
//     matrices: k / v projection (s = 0 / 1), one 16-feature tile of W1 (s <= 7), half of the K range of a W2 tile
//     (s = 2..5), a tile of Wo (s = 6, 7), the merge of the head's 9 attention partials (s = 8), 1-2 tiles of the head.
//   * WEIGHTS never wait for activations: each CTA's pieces stream HBM -> shared memory through a 2-layer-deep ring of TMA
//     bulk copies (80 KB per layer and CTA), issued two layers (and steps) ahead; the MMAs read their A fragments from shared
//     memory (fragment order: one conflict-free 128-bit load per lane).
//   * ACTIVATIONS travel between CTAs through L2 in 8-byte cells {value, tag}: data and "ready" are one store (no fence, no
//     atomic).  MEASURED (scripts/mb_ll.cu, mb_ll2.cu, timeline_ws.py): one-way signalling between two SMs through L2 costs
//     ~0.5 us with strong (volatile / relaxed.gpu) accesses, an all-to-all hand-off of 4-16 KB between 144 CTAs 1.2-1.7 us -
//     no better than the counter grid barrier (1.2 us) - and 0.7-1.9 us inside this kernel.  The tag
//     is a counter that grows by one per (step, layer), so buffers are never cleared; each buffer is written once per layer
//     and every reader of layer l's data has provably finished before anybody can write layer l+1's (see "ordering").
//     Vectors that many CTAs read (the residual sums, the FFN hidden) are written in several replicas so that the
//     144 readers do not queue on the same L2 slices.
//   * five hand-offs per layer: attention partials (head-local) -> attention output (all-gather, 1 KB per row) -> residual sum 1
//     -> FFN hidden -> residual sum 2 (two K-halves, summed by the readers in a fixed order).  LayerNorm is computed
//     redundantly by every CTA on the gathered fp32 rows (one warp per row: no block barrier between gather and LayerNorm).
//   * K/V: positions of the slice are read straight from L2 into registers (16 bytes per thread and position: 4 lanes per
//     64-byte head row), requested BEFORE the wait for the layer's input; the slice of layer l+2 is prefetched into L2.
//   * step tail: logits through global memory, grid barrier, sample_row (the fused sampler of phases.cuh) on CTA r for
//     sequence r, grid barrier, and - only in steps in which a sequence stopped - phase_plan + a third barrier.
// Deterministic: fixed-order reductions everywhere, no floating-point atomics.
//
// STATUS (measured on B200): parity-green against the reference goldens (tests/test_gpu_horizons.py,
// mode 6), but NOT faster than the cluster-stream kernel: 283 us per step at batch 1 against 264 us (523 vs 284 at batch 8),
// because the five L2 hand-offs per layer cost 6.3 of the 10.6 us per layer - DSMEM inside a cluster (0.3-0.5 us per hand-off)
// beats L2 between clusters by 3x, which is what the cluster-stream design exploits.  It therefore stays an explicit mode
// (T2S_OPT_DECODE_MODE = 6) and auto mode never picks it.
//
// Ordering argument for the single-buffered hand-offs (X1a partials, X1b attention, X2 sum 1, X3 hidden, X4 sum 2): a CTA
// writes X_k of layer l+1 only after it has read ALL of X4(l) (every CTA starts a layer with that gather), X4(l) is complete
// only when every FFN2 CTA has read X3(l) and X2(l), X3(l) only when every FFN1 CTA has read X2(l), X2(l) only when every Wo
// CTA has read X1b(l), X1b(l) only when every merger has read X1a(l).  Hence all readers of X_k(l) are done before any
// writer of X_k(l+1) starts; and X4(l+1) is written only after X1a(l+1) is complete, i.e. after every CTA has read X4(l).
#pragma once
#include "cluster_decode.cuh"

namespace t2s {
namespace ws {

using cs::bulk_load;
using cs::csync;
using cs::fast_exp2;
using cs::mbar_expect_tx;
using cs::mbar_init;
using cs::mbar_wait;
using cs::s32;

constexpr int NHG = 9;                 // CTAs per head
constexpr int G = NH * NHG;            // 144 CTAs
constexpr int RW = 8;                  // sequences (one MMA n-tile)
constexpr int NTW = 256, NWW = 8;      // threads / warps per CTA (== NT: sample_row / phase_plan use named barrier 1 over NT threads)
constexpr int TILE_B = 32 * 512;       // bytes of a 16-feature x 512-column weight tile in m16n8k16 A-fragment order (32 k-blocks)
constexpr int SLOT_B = 5 * TILE_B;     // one layer's pieces of a CTA: [q 2 tiles | k or v or W2 half-tile or Wo tile: 2 | W1 tile]
constexpr int REG_Q = 0, REG_B = 2 * TILE_B, REG_W1 = 4 * TILE_B;
constexpr int XS8 = D + 8;             // bf16 row stride of the 512-wide operand (conflict-free B fragments)
constexpr int HS8 = FF / 2 + 8;        // bf16 row stride of the FFN2 operand (half of the hidden units)
// ---- packed weights (k_pack_wide): per layer [LNV | CB | QKV by head | Wo | W1 | W2]; after the layers one block for the head
constexpr int LNV_B = 4 * D * 4;       // fp32: norm2 of the PREVIOUS layer (gamma, beta), norm1 of this layer (gamma, beta)
constexpr int CB_FLOATS = 128, CB_B = CB_FLOATS * 4;  // per-CTA biases: [0,32) q, [32,64) k|v, [64,80) Wo tile, [80,96) W1 tile, [96,112) W2 tile (K-half 0 only)
constexpr size_t WL_LNV = 0, WL_CB = LNV_B, WL_QKV = WL_CB + (size_t)G * CB_B, WL_WO = WL_QKV + (size_t)96 * TILE_B,
                 WL_W1 = WL_WO + (size_t)32 * TILE_B, WL_W2 = WL_W1 + (size_t)128 * TILE_B, WL_BYTES = WL_W2 + (size_t)128 * TILE_B;
constexpr size_t WH_TILES = WL_QKV, WH_BYTES = WH_TILES + (size_t)VT * TILE_B;  // head block: [LNV (final norm2) | CB (zeros) | 65 tiles]
// ---- hand-off buffers in global memory, in 8-byte cells {value, tag}
constexpr int R4 = 4, R2 = 4, R3 = 2;  // replicas of the vectors read by (almost) every CTA
constexpr int X1A_SLOT = 2 + DH, X1A_ROW = (NHG + 1) * X1A_SLOT;  // per (head, row): 9 partial states (m, l, o[32]) + (score, -, v[32]) of the new token
constexpr size_t LL_HDR = 16, LL_X4 = LL_HDR, LL_X2 = LL_X4 + (size_t)R4 * 2 * RW * D, LL_X3 = LL_X2 + (size_t)R2 * RW * D,
                 LL_X1B = LL_X3 + (size_t)R3 * RW * (FF / 2), LL_X1A = LL_X1B + (size_t)RW * (D / 2),
                 LL_CELLS = LL_X1A + (size_t)NH * RW * X1A_ROW;
constexpr unsigned TAG_STRIDE = 128;   // tags per step (>= n_layer + 1)

struct __align__(128) Smem {
  unsigned char slot[2][SLOT_B];
  alignas(16) float lnv[2][4 * D];
  alignas(16) float cb[2][CB_FLOATS];
  alignas(16) float y[RW][D];        // fp32 rows: gathered residual sum -> (in place) LayerNorm output = the next residual
  alignas(16) bf16 xn[RW][XS8];      // bf16 GEMM operand: LayerNorm output / layer-0 input / attention output (Wo CTAs)
  union {
    bf16 hh[RW][HS8];                // FFN2 operand: this CTA's half of the hidden units
    struct { float att[NWW][RW][X1A_SLOT]; float red2[NWW][16][RW + 1]; } a;  // attention: per-warp partial states; k|v partial sums
  } u;
  float red[NWW][16][RW + 1];
  float q[RW][DH];
  float kvn[RW][DH];                 // this step's k (s = 0) or v (s = 1) of the head, as stored (bf16-rounded)
  int pt[RW][32];
  int row_slot[RW], row_pos[RW], sl_a[RW], sl_b[RW];
  long long row_kvoff[RW];
  unsigned long long full[4][2];     // [q, B, W1, vec][slot]
};
static_assert(sizeof(Smem) <= 232448, "shared memory budget");
static_assert(sizeof(SampSmem) <= sizeof(float) * RW * D + sizeof(bf16) * RW * XS8, "the sampler scratch aliases y | xn");
static_assert(offsetof(Smem, xn) == offsetof(Smem, y) + sizeof(float) * RW * D, "y and xn are contiguous");

__device__ __forceinline__ uint4 ldv16(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint2 ldv8(const void* p) {
  uint2 r;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_cell(unsigned long long* p, uint32_t v, uint32_t tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v), "r"(tag) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// polling with a watchdog: a lost producer becomes an error (abort flag), never a hung GPU
struct Watch {
  int* abort_flag;
  long long t0;
  unsigned spins;
  bool dead;
  __device__ __forceinline__ void arm() { t0 = 0; spins = 0; }
  __device__ __forceinline__ bool tick() {
    if ((++spins & 0x3FFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ll || __ldcg(abort_flag) != 0) { atomicCAS(abort_flag, 0, ABORT_WATCHDOG); dead = true; }
    }
    return dead;
  }
};
// NL 16-byte loads per lane (2 cells each), vector i*32 + lane of `src`; vectors >= nvec are not read.  Returns when every
// cell carries `tag` (or the watchdog fired).
template <int NL>
__device__ __forceinline__ void poll_vec(const uint4* src, int nvec, uint32_t tag, uint4 (&v)[NL], Watch& w) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NL; ++i) v[i] = (i * 32 + lane < nvec) ? ldv16(src + i * 32 + lane) : make_uint4(0, tag, 0, tag);
  w.arm();
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NL; ++i)
      if (v[i].y != tag || v[i].w != tag) { v[i] = ldv16(src + i * 32 + lane); ok = false; }
    if (ok || w.tick()) break;
  }
}

// ---- packed layout: byte `rel` of a tile region -> (k-block, row inside the 16-feature tile, column inside the k-block)
__host__ __device__ inline void tile_elem(int rel, int& kb, int& fr, int& kc) {
  kb = rel >> 9;
  const int lane = (rel >> 4) & 31, j = (rel >> 1) & 7, g = lane >> 2, t = lane & 3;
  fr = g + ((j & 2) ? 8 : 0);
  kc = 2 * t + (j & 1) + ((j & 4) ? 8 : 0);
}
// roles of CTA (h, s): shared by the packer (biases) and the kernel
struct Role {
  int h, s, kv, w1_tile, w2_tile, w2_half, wo_tile, nht, ht[2];
  bool merger;
};
__host__ __device__ inline Role role_of(int cta) {
  Role r;
  r.h = cta / NHG; r.s = cta % NHG;
  r.kv = (r.s == 0) ? 1 : (r.s == 1) ? 2 : 0;              // 1: k projection, 2: v projection
  r.w1_tile = (r.s <= 7) ? r.h * 8 + r.s : -1;             // 128 tiles of linear1
  const int w2i = (r.s >= 2 && r.s <= 5) ? r.h * 4 + (r.s - 2) : -1;  // 64 = 32 tiles of linear2 x 2 K-halves
  r.w2_tile = w2i >= 0 ? (w2i >> 1) : -1; r.w2_half = w2i >= 0 ? (w2i & 1) : 0;
  r.wo_tile = (r.s == 6 || r.s == 7) ? r.h * 2 + (r.s - 6) : -1;       // 32 tiles of out_proj
  r.merger = (r.s == 8);
  r.nht = 0; r.ht[0] = r.ht[1] = 0;
  if (r.s == 6 || r.s == 7) { r.ht[0] = r.h * 2 + (r.s - 6); r.nht = 1; if (cta == 6) { r.ht[1] = 64; r.nht = 2; } }  // tiles 0..31, + the 65th
  if (r.s == 8) { r.ht[0] = 32 + r.h * 2; r.ht[1] = 33 + r.h * 2; r.nht = 2; }                                          // tiles 32..63
  return r;
}

// wrow: row-major bf16 layer matrices [n_layer][LW]; wvec: [n_layer][LV] fp32; whead_row: [V][D] bf16 -> out (n_layer * WL_BYTES + WH_BYTES)
__global__ void k_pack_wide(unsigned char* __restrict__ out, const bf16* __restrict__ wrow, const float* __restrict__ wvec,
                            const bf16* __restrict__ whead_row, int n_layer) {
  const size_t total = ((size_t)n_layer * WL_BYTES + WH_BYTES) / 2;  // 2-byte units
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t byte = e * 2;
    const int layer = (int)(byte / WL_BYTES);  // == n_layer: the head block
    const size_t off = byte - (size_t)layer * WL_BYTES;
    const bool head = layer >= n_layer;
    if (off < WL_QKV) {  // fp32 vectors: one thread per float
      if (off & 3) continue;
      float v = 0.f;
      if (off < WL_CB) {
        const int i = (int)(off >> 2), which = i / D, f = i % D;
        // norm2 of the previous layer (identity for layer 0: never applied), norm1 of this layer (unused in the head block)
        if (which < 2) v = layer > 0 ? wvec[(size_t)(layer - 1) * LV + (which == 0 ? VO_G2 : VO_BE2) + f] : (which == 0 ? 1.f : 0.f);
        else v = head ? 0.f : wvec[(size_t)layer * LV + (which == 2 ? VO_G1 : VO_BE1) + f];
      } else if (!head) {
        const int i = (int)((off - WL_CB) >> 2), cta = i / CB_FLOATS, k = i % CB_FLOATS;
        const Role r = role_of(cta);
        const float* vl = wvec + (size_t)layer * LV;
        if (k < 32) v = vl[VO_BQKV + r.h * DH + k];
        else if (k < 64) v = r.kv ? vl[VO_BQKV + r.kv * D + r.h * DH + (k - 32)] : 0.f;
        else if (k < 80) v = r.wo_tile >= 0 ? vl[VO_BO + r.wo_tile * 16 + (k - 64)] : 0.f;
        else if (k < 96) v = r.w1_tile >= 0 ? vl[VO_B1 + r.w1_tile * 16 + (k - 80)] : 0.f;
        else if (k < 112) v = (r.w2_tile >= 0 && r.w2_half == 0) ? vl[VO_B2 + r.w2_tile * 16 + (k - 96)] : 0.f;
      }
      *reinterpret_cast<float*>(out + byte) = v;
      continue;
    }
    int kb, fr, kc;
    bf16 v;
    if (head) {
      const size_t rel = off - WH_TILES;
      const int tile = (int)(rel / TILE_B);
      tile_elem((int)(rel % TILE_B), kb, fr, kc);
      const int row = tile * 16 + fr;
      v = row < V ? whead_row[(size_t)row * D + kb * 16 + kc] : __float2bfloat16_rn(0.f);
    } else {
      const bf16* src = wrow + (size_t)layer * LW;
      if (off < WL_WO) {  // head h: q tile 0,1 | k tile 0,1 | v tile 0,1
        const size_t rel = off - WL_QKV;
        const int t6 = (int)(rel / TILE_B), h = t6 / 6, w = t6 % 6;
        tile_elem((int)(rel % TILE_B), kb, fr, kc);
        v = src[OFF_WQKV + (size_t)((w >> 1) * D + h * DH + (w & 1) * 16 + fr) * D + kb * 16 + kc];
      } else if (off < WL_W1) {
        const size_t rel = off - WL_WO;
        tile_elem((int)(rel % TILE_B), kb, fr, kc);
        v = src[OFF_WO + (size_t)((int)(rel / TILE_B) * 16 + fr) * D + kb * 16 + kc];
      } else if (off < WL_W2) {
        const size_t rel = off - WL_W1;
        tile_elem((int)(rel % TILE_B), kb, fr, kc);
        v = src[OFF_W1 + (size_t)((int)(rel / TILE_B) * 16 + fr) * D + kb * 16 + kc];
      } else {  // linear2: a tile is 128 k-blocks = 4 x TILE_B, K-half kh = k-blocks [64 kh, 64 kh + 64)
        const size_t rel = off - WL_W2;
        const int tile = (int)(rel / (4 * TILE_B));
        const int r4 = (int)(rel % (4 * TILE_B));
        tile_elem(r4 % TILE_B, kb, fr, kc);
        kb += (r4 / TILE_B) * 32;
        v = src[OFF_W2 + (size_t)(tile * 16 + fr) * FF + kb * 16 + kc];
      }
    }
    *reinterpret_cast<bf16*>(out + byte) = v;
  }
}

// NK k-blocks of one 16-feature tile: A fragments from shared memory (wt = tile base + first k-block, fragment order), B from the
// bf16 activation rows `act` (row stride ROWS elements, first column of the first k-block).  Four accumulator chains
// (mma.sync latency ~100 cycles).  acc: c0,c1 = feature g, sequences 2t, 2t+1; c2,c3 = feature g + 8.
template <int NK, int ROWS>
__device__ __forceinline__ void mma_tile(const unsigned char* wt, const bf16* act, float (&acc)[4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
  const uint4* wp = reinterpret_cast<const uint4*>(wt) + lane;
  const bf16* arow = act + g * ROWS;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const uint4 f = wp[k * 32];
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(arow + k * 16);
    mma_bf16_16816(a[k & 3], f, xr[t], xr[4 + t]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = (a[0][j] + a[1][j]) + (a[2][j] + a[3][j]);
}
__device__ __forceinline__ void red_store(float (*red)[16][RW + 1], const float (&acc)[4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  red[warp][g][2 * t] = acc[0]; red[warp][g][2 * t + 1] = acc[1];
  red[warp][g + 8][2 * t] = acc[2]; red[warp][g + 8][2 * t + 1] = acc[3];
}

// LayerNorm of one row held by a warp: lane owns values v[i][0..1] = features 64 i + 2 lane + (0, 1).  Writes the fp32 result to
// yrow (the next residual) and its bf16 copy to xrow (the next GEMM operand).
__device__ __forceinline__ void ln_row_store(float (&v)[8][2], const float* gam, const float* bet, float* yrow, bf16* xrow) {
  const int lane = threadIdx.x & 31;
  float s = 0.f, qq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += v[i][0] + v[i][1]; qq += v[i][0] * v[i][0] + v[i][1] * v[i][1]; }
  s = warp_sum(s); qq = warp_sum(qq);
  const float mean = s * (1.0f / D);
  const float var = fmaxf(qq * (1.0f / D) - mean * mean, 0.f);
  const float rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = 64 * i + 2 * lane;
    const float2 gg = *reinterpret_cast<const float2*>(gam + f), bb = *reinterpret_cast<const float2*>(bet + f);
    const float o0 = (v[i][0] - mean) * rstd * gg.x + bb.x, o1 = (v[i][1] - mean) * rstd * gg.y + bb.y;
    *reinterpret_cast<float2*>(yrow + f) = make_float2(o0, o1);
    *reinterpret_cast<uint32_t*>(xrow + f) = pack_bf2(o0, o1);
  }
}

// =====================================================================================================================
__global__ void __launch_bounds__(NTW, 1)
k_decode_wide(Ctx c, const unsigned char* __restrict__ wwide, unsigned long long* __restrict__ ll, int max_new_steps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const Role ro = role_of(cta);
  const int L = c.n_layer, UPS = L + 1;  // units per step: the layers + the head
  if (tid == 0) {
    for (int r = 0; r < 4; ++r)
      for (int s = 0; s < 2; ++s) mbar_init(&sm.full[r][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < RW * XS8 / 2; i += NTW) reinterpret_cast<uint32_t*>(&sm.xn[0][0])[i] = 0u;  // rows >= R stay finite
  for (int i = tid; i < RW * HS8 / 2; i += NTW) reinterpret_cast<uint32_t*>(&sm.u.hh[0][0])[i] = 0u;
  __syncthreads();

  // ---- weight ring: unit u (= step * UPS + layer, the head being "layer" L) lives in slot u & 1; region r of a slot has its
  //      own mbarrier; a region is refilled for unit u + 2 as soon as unit u's MMAs over it are done (thread 0, after the CTA
  //      barrier that follows them).  Weights do not depend on the step, so the ring runs ahead across steps.
  auto region_used = [&](int region, int layer) -> bool {
    if (region == 3) return true;                                  // vectors
    if (layer == L) return region == 0 && ro.nht > 0;              // head tiles sit in the q region
    return region == 0 || ro.s <= 7;                               // q: everybody; B and W1: s <= 7
  };
  auto issue = [&](int u, int region) {  // thread 0 only
    const int layer = u % UPS, slot = u & 1;
    if (!region_used(region, layer)) return;
    const unsigned char* blk = wwide + (size_t)layer * WL_BYTES;
    void* bar = &sm.full[region][slot];
    unsigned char* dst = sm.slot[slot];
    if (region == 3) {
      mbar_expect_tx(bar, LNV_B + CB_B);
      bulk_load(sm.lnv[slot], blk + WL_LNV, LNV_B, bar);
      bulk_load(sm.cb[slot], blk + WL_CB + (size_t)cta * CB_B, CB_B, bar);
    } else if (layer == L) {
      mbar_expect_tx(bar, (uint32_t)ro.nht * TILE_B);
      for (int i = 0; i < ro.nht; ++i) bulk_load(dst + REG_Q + i * TILE_B, blk + WH_TILES + (size_t)ro.ht[i] * TILE_B, TILE_B, bar);
    } else if (region == 0) {
      mbar_expect_tx(bar, 2 * TILE_B);
      bulk_load(dst + REG_Q, blk + WL_QKV + (size_t)ro.h * 6 * TILE_B, 2 * TILE_B, bar);
    } else if (region == 1) {
      if (ro.kv) {
        mbar_expect_tx(bar, 2 * TILE_B);
        bulk_load(dst + REG_B, blk + WL_QKV + ((size_t)ro.h * 6 + 2 * ro.kv) * TILE_B, 2 * TILE_B, bar);
      } else if (ro.w2_tile >= 0) {
        mbar_expect_tx(bar, 2 * TILE_B);
        bulk_load(dst + REG_B, blk + WL_W2 + ((size_t)ro.w2_tile * 4 + 2 * ro.w2_half) * TILE_B, 2 * TILE_B, bar);
      } else {
        mbar_expect_tx(bar, TILE_B);
        bulk_load(dst + REG_B, blk + WL_WO + (size_t)ro.wo_tile * TILE_B, TILE_B, bar);
      }
    } else {
      mbar_expect_tx(bar, TILE_B);
      bulk_load(dst + REG_W1, blk + WL_W1 + (size_t)ro.w1_tile * TILE_B, TILE_B, bar);
    }
  };
  uint32_t par = 0;  // bit (region * 2 + slot): parity of the next completion of that barrier (every thread tracks its own copy)
  auto wait_region = [&](int region, int slot) {
    const uint32_t bit = 1u << (region * 2 + slot);
    mbar_wait(&sm.full[region][slot], (par & bit) ? 1u : 0u);
    par ^= bit;
  };
  if (tid == 0)
    for (int u = 0; u < 2; ++u)
      for (int r = 0; r < 4; ++r) issue(u, r);
  int vec_issued = 1;  // thread 0: last unit whose vectors were requested (they run ONE unit ahead, the matrices two)

  cs::GridBar gbar{c.bar, c.abort_flag, 0u, gridDim.x};
  Watch watch{c.abort_flag, 0, 0u, false};
  const uint32_t ebase = (uint32_t)__ldcg(reinterpret_cast<const unsigned*>(ll));  // tags of earlier launches are all below
  unsigned long long* const X4 = ll + LL_X4;
  unsigned long long* const X2 = ll + LL_X2;
  unsigned long long* const X3 = ll + LL_X3;
  unsigned long long* const X1B = ll + LL_X1B;
  unsigned long long* const X1A = ll + LL_X1A;
  SampSmem& ss = *reinterpret_cast<SampSmem*>(&sm.y[0][0]);
  int step = ld_cg_i(c.step);
  int R = 0;
  bool fresh = true;
  int u = 0, it = 0;
  long long* tl = nullptr;  // measurement hook: clock stamps of thread 0 at the markers of one step
  int tk = 0;
#define WS_TL() do { if (tl && tid == 0 && tk < 2 * c.tl_slots) tl[tk++] = clock64(); } while (0)
  for (; it < max_new_steps; ++it, ++step) {
    if (fresh) R = ld_cg_i(c.n_active);
    if (R == 0 || __ldcg(c.abort_flag) != 0) break;
    bool stopped = false;
    tl = (c.timeline && it == c.tl_step) ? c.timeline + (size_t)blockIdx.x * c.tl_slots * 2 : nullptr;
    tk = 0;
    WS_TL();
    // ---- step prologue: row descriptors, page-table rows, K/V slices, layer-0 input
    if (fresh) {
      if (tid < R) {
        sm.row_slot[tid] = ld_cg_i(c.row_slot + tid);
        sm.row_pos[tid] = ld_cg_i(c.row_pos + tid);
        sm.row_kvoff[tid] = __ldcg(c.row_kvoff + tid);
      }
      csync();
      for (int i = tid; i < R * 32; i += NTW) {
        const int n = i >> 5, pg = i & 31;
        sm.pt[n][pg] = (pg < c.max_pages) ? c.page_table[sm.row_slot[n] * c.max_pages + pg] : 0;
      }
    }
    if (tid < R) {  // slice s of the row's cached positions [0, pos): nine nearly equal parts
      const int pos = sm.row_pos[tid], per = (pos + NHG - 1) / NHG;
      sm.sl_a[tid] = min(pos, ro.s * per);
      sm.sl_b[tid] = min(pos, (ro.s + 1) * per);
    }
    if (warp < R) {
      const float* xr = c.x0 + (size_t)sm.row_slot[warp] * D;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = 64 * i + 2 * lane;
        const float2 a = __ldcg(reinterpret_cast<const float2*>(xr + f));
        *reinterpret_cast<float2*>(&sm.y[warp][f]) = a;
        *reinterpret_cast<uint32_t*>(&sm.xn[warp][f]) = pack_bf2(a.x, a.y);
      }
    }
    csync();
    WS_TL();
    for (int layer = 0; layer <= L; ++layer, ++u) {
      const int slot = u & 1;
      const unsigned char* sb = sm.slot[slot];
      const uint32_t tag = ebase + (uint32_t)it * TAG_STRIDE + (uint32_t)layer + 1u;  // this layer's hand-offs; tag - 1: the previous layer's X4
      wait_region(3, slot);
      const float* lnv = sm.lnv[slot];
      const float* cb = sm.cb[slot];
      // ---- layer input: residual sum 2 of the previous layer (two K-halves) -> LayerNorm 2 -> y (fp32), xn (bf16)
      if (layer > 0 && warp < R) {
        const int n = warp;
        const uint4* src = reinterpret_cast<const uint4*>(X4 + (size_t)(cta % R4) * 2 * RW * D + (size_t)n * D);
        uint4 a[8], b[8];
        poll_vec<8>(src, 256, tag - 1, a, watch);
        poll_vec<8>(src + (size_t)RW * D / 2, 256, tag - 1, b, watch);
        float v[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i][0] = __uint_as_float(a[i].x) + __uint_as_float(b[i].x);
          v[i][1] = __uint_as_float(a[i].z) + __uint_as_float(b[i].z);
        }
        ln_row_store(v, lnv, lnv + D, sm.y[n], sm.xn[n]);
      }
      csync();
      if (tid == 0 && u >= 1) { issue(u + 1, 3); vec_issued = u + 1; }  // the vectors of the next unit: their buffer was last read in unit u - 1
      WS_TL();  // 0: layer input gathered + LayerNorm 2
      if (layer == L) break;
      // ---- q for head h (every CTA), k (s = 0) or v (s = 1): warp = (tile w & 1, K quarter w >> 1)
      {
        wait_region(0, slot);
        if (ro.kv) wait_region(1, slot);
        float acc[4];
        const int tl = warp & 1, kq = warp >> 1;
        mma_tile<8, XS8>(sb + REG_Q + (tl * 32 + kq * 8) * 512, &sm.xn[0][kq * 128], acc);
        red_store(sm.red, acc);
        if (ro.kv) {
          mma_tile<8, XS8>(sb + REG_B + (tl * 32 + kq * 8) * 512, &sm.xn[0][kq * 128], acc);
          red_store(sm.u.a.red2, acc);
        }
      }
      csync();
      if (tid == 0) { issue(u + 2, 0); if (ro.kv) issue(u + 2, 1); }
      {
        const int n = warp, fl = lane, tl = fl >> 4, fr = fl & 15;
        if (n < R) {
          const float sq = (sm.red[tl][fr][n] + sm.red[tl + 2][fr][n]) + (sm.red[tl + 4][fr][n] + sm.red[tl + 6][fr][n]);
          sm.q[n][fl] = (sq + cb[fl]) * QSCALE;
          if (ro.kv) {
            const float sk = (sm.u.a.red2[tl][fr][n] + sm.u.a.red2[tl + 2][fr][n]) + (sm.u.a.red2[tl + 4][fr][n] + sm.u.a.red2[tl + 6][fr][n]);
            const bf16 hv = __float2bfloat16_rn(sk + cb[32 + fl]);
            sm.kvn[n][fl] = __bfloat162float(hv);
            bf16* pool = (ro.kv == 1 ? c.kpool : c.vpool) + (size_t)layer * c.kv_layer_stride + (size_t)ro.h * KV_HEAD_STRIDE;
            const long long o = sm.row_kvoff[n];
            pool[o + kv_feat(o, fl)] = hv;  // appended for the later steps (read by the other CTAs of the head after a grid barrier)
          }
        }
      }
      csync();
      WS_TL();  // 1: q (k | v) projection + epilogue
      // ---- attention over this CTA's slice of the cached positions: thread = (position lane pl, 16-byte quarter j of the head row)
      {
        const int j = tid & 3, pl = tid >> 2;
        const bf16* kbase = c.kpool + (size_t)layer * c.kv_layer_stride + (size_t)ro.h * KV_HEAD_STRIDE;
        const long long pf_off = (long long)((layer + 2 < L) ? 2 : 2 - L) * (long long)c.kv_layer_stride;  // layer + 2 (next step's 0 / 1 at the end)
        for (int n = 0; n < R; ++n) {
          const int a0 = sm.sl_a[n], b0 = sm.sl_b[n];
          float qv[8];
          {
            const float4 q0 = *reinterpret_cast<const float4*>(&sm.q[n][8 * j]), q1 = *reinterpret_cast<const float4*>(&sm.q[n][8 * j + 4]);
            qv[0] = q0.x; qv[1] = q0.y; qv[2] = q0.z; qv[3] = q0.w; qv[4] = q1.x; qv[5] = q1.y; qv[6] = q1.z; qv[7] = q1.w;
          }
          float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = 0.f;
          for (int p0 = a0; p0 < b0; p0 += 4 * 64) {
            uint4 kk[4], vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int p = p0 + 64 * i + pl;
              if (p < b0) {
                const int pin = p & (PAGE - 1);
                const bf16* kr = kbase + kv_row_off(sm.pt[n][p >> PAGE_SHIFT], pin) + ((j ^ ((pin >> 1) & 3)) << 3);
                kk[i] = ld_cg16(kr);
                vv[i] = ld_cg16(kr + KV_V_OFF);
                if (j == 0 && (pin & 1) == 0) { prefetch_l2(kr + pf_off); prefetch_l2(kr + KV_V_OFF + pf_off); }
              } else {
                kk[i] = vv[i] = make_uint4(0, 0, 0, 0);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float sc = qv[0] * bf_lo(kk[i].x) + qv[1] * bf_hi(kk[i].x) + qv[2] * bf_lo(kk[i].y) + qv[3] * bf_hi(kk[i].y) +
                         qv[4] * bf_lo(kk[i].z) + qv[5] * bf_hi(kk[i].z) + qv[6] * bf_lo(kk[i].w) + qv[7] * bf_hi(kk[i].w);
              sc += __shfl_xor_sync(0xffffffffu, sc, 1);
              sc += __shfl_xor_sync(0xffffffffu, sc, 2);
              if (p0 + 64 * i + pl < b0) {
                const float mn = fmaxf(m, sc);
                const float corr = fast_exp2(m - mn), p = fast_exp2(sc - mn);  // m = -inf: corr = 0
                m = mn;
                l = l * corr + p;
                o[0] = o[0] * corr + p * bf_lo(vv[i].x); o[1] = o[1] * corr + p * bf_hi(vv[i].x);
                o[2] = o[2] * corr + p * bf_lo(vv[i].y); o[3] = o[3] * corr + p * bf_hi(vv[i].y);
                o[4] = o[4] * corr + p * bf_lo(vv[i].z); o[5] = o[5] * corr + p * bf_hi(vv[i].z);
                o[6] = o[6] * corr + p * bf_lo(vv[i].w); o[7] = o[7] * corr + p * bf_hi(vv[i].w);
              }
            }
          }
          // merge the 8 position lanes of the warp that share quarter j (lanes j, j + 4, ..., j + 28), fixed order
#pragma unroll
          for (int sh = 4; sh < 32; sh <<= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, sh), l2 = __shfl_xor_sync(0xffffffffu, l, sh);
            const float mn = fmaxf(m, m2);
            const float s1 = (m == -INFINITY) ? 0.f : fast_exp2(m - mn), s2 = (m2 == -INFINITY) ? 0.f : fast_exp2(m2 - mn);
            // both operands of the sum are combined in lane order (lower lane first) so that both partners get the same bits
            const bool lo = (lane & sh) == 0;
            l = lo ? (l * s1 + l2 * s2) : (l2 * s2 + l * s1);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float o2 = __shfl_xor_sync(0xffffffffu, o[e], sh);
              o[e] = lo ? (o[e] * s1 + o2 * s2) : (o2 * s2 + o[e] * s1);
            }
            m = mn;
          }
          if (lane < 4) {
            float* dst = sm.u.a.att[warp][n];
            if (lane == 0) { dst[0] = m; dst[1] = l; }
#pragma unroll
            for (int e = 0; e < 8; ++e) dst[2 + 8 * lane + e] = o[e];
          }
        }
      }
      csync();
      WS_TL();  // 2: attention over the slice
      // ---- CTA-level merge of the 8 warps' states (warp n = sequence n, lane = head dim) -> hand-off X1a
      if (warp < R) {
        const int n = warp, d = lane;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NWW; ++w) M = fmaxf(M, sm.u.a.att[w][n][0]);
        float Ls = 0.f, A = 0.f;
        if (M != -INFINITY) {
#pragma unroll
          for (int w = 0; w < NWW; ++w) {
            const float mw = sm.u.a.att[w][n][0];
            if (mw != -INFINITY) {
              const float scl = fast_exp2(mw - M);
              Ls += sm.u.a.att[w][n][1] * scl;
              A += sm.u.a.att[w][n][2 + d] * scl;
            }
          }
        }
        unsigned long long* dst = X1A + ((size_t)ro.h * RW + n) * X1A_ROW + (size_t)ro.s * X1A_SLOT;
        st_cell(dst + 2 + d, __float_as_uint(A), tag);
        if (d == 0) st_cell(dst + 0, __float_as_uint(M), tag);
        if (d == 1) st_cell(dst + 1, __float_as_uint(Ls), tag);
        if (ro.kv) {  // the new token: its score (s = 0) / its value row (s = 1) go to the extra slot
          unsigned long long* ex = X1A + ((size_t)ro.h * RW + n) * X1A_ROW + (size_t)NHG * X1A_SLOT;
          if (ro.kv == 1) {
            const float sn = warp_sum(sm.q[n][d] * sm.kvn[n][d]);
            if (d == 0) st_cell(ex + 0, __float_as_uint(sn), tag);
            if (d == 1) st_cell(ex + 1, 0u, tag);
          } else {
            st_cell(ex + 2 + d, __float_as_uint(sm.kvn[n][d]), tag);
          }
        }
      }
      // ---- s = 8: merge the head's nine partial states + the new token -> attention output of head h (bf16) -> hand-off X1b
      if (ro.merger && warp < R) {
        const int n = warp, d = lane;
        const unsigned long long* src = X1A + ((size_t)ro.h * RW + n) * X1A_ROW;
        uint4 ml[NHG + 1];
        uint2 ov[NHG + 1];
#pragma unroll
        for (int i = 0; i <= NHG; ++i) { ml[i] = ldv16(src + i * X1A_SLOT); ov[i] = ldv8(src + i * X1A_SLOT + 2 + d); }
        watch.arm();
        for (;;) {
          bool ok = true;
#pragma unroll
          for (int i = 0; i <= NHG; ++i) {
            if (ml[i].y != tag || ml[i].w != tag) { ml[i] = ldv16(src + i * X1A_SLOT); ok = false; }
            if (ov[i].y != tag) { ov[i] = ldv8(src + i * X1A_SLOT + 2 + d); ok = false; }
          }
          if (ok || watch.tick()) break;
        }
        const float snew = __uint_as_float(ml[NHG].x);
        float M = snew;
#pragma unroll
        for (int i = 0; i < NHG; ++i) M = fmaxf(M, __uint_as_float(ml[i].x));
        const float en = fast_exp2(snew - M);
        float Ls = en, A = en * __uint_as_float(ov[NHG].x);
#pragma unroll
        for (int i = 0; i < NHG; ++i) {
          const float mi = __uint_as_float(ml[i].x);
          if (mi != -INFINITY) {
            const float scl = fast_exp2(mi - M);
            Ls += __uint_as_float(ml[i].z) * scl;
            A += __uint_as_float(ov[i].x) * scl;
          }
        }
        const float outv = A / Ls;
        const float nb = __shfl_down_sync(0xffffffffu, outv, 1);
        if ((d & 1) == 0) st_cell(X1B + (size_t)n * (D / 2) + ro.h * (DH / 2) + (d >> 1), pack_bf2(outv, nb), tag);
      }
      WS_TL();  // 3: partial published (s = 8: + the head's merge published)
      // ---- s = 6, 7: out_proj tile + bias + residual -> hand-off X2 (residual sum 1, fp32)
      if (ro.wo_tile >= 0) {
        if (warp < R) {
          const int n = warp;
          uint4 a[4];
          poll_vec<4>(reinterpret_cast<const uint4*>(X1B + (size_t)n * (D / 2)), 128, tag, a, watch);
#pragma unroll
          for (int i = 0; i < 4; ++i)  // vector i*32 + lane = cells 2v, 2v + 1 = bf16 features 4v .. 4v + 3
            *reinterpret_cast<uint2*>(&sm.xn[n][4 * (i * 32 + lane)]) = make_uint2(a[i].x, a[i].z);
        }
        csync();
        wait_region(1, slot);
        float acc[4];
        mma_tile<4, XS8>(sb + REG_B + (warp * 4) * 512, &sm.xn[0][warp * 64], acc);
        red_store(sm.red, acc);
        csync();
        if (tid == 0) issue(u + 2, 1);
        if (tid < 128) {
          const int n = tid >> 4, fl = tid & 15;
          if (n < R) {
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < NWW; ++w) sacc += sm.red[w][fl][n];
            const float yv = sacc + cb[64 + fl] + sm.y[n][ro.wo_tile * 16 + fl];
#pragma unroll
            for (int r = 0; r < R2; ++r) st_cell(X2 + ((size_t)r * RW + n) * D + ro.wo_tile * 16 + fl, __float_as_uint(yv), tag);
          }
        }
      }
      WS_TL();  // 4: (s = 6, 7) out_proj tile published
      // ---- s <= 7: residual sum 1 -> LayerNorm 1 -> linear1 tile + bias + ReLU -> hand-off X3 (bf16 pairs)
      if (ro.s <= 7) {
        if (warp < R) {
          const int n = warp;
          uint4 a[8];
          poll_vec<8>(reinterpret_cast<const uint4*>(X2 + ((size_t)(cta % R2) * RW + n) * D), 256, tag, a, watch);
          float v[8][2];
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[i][0] = __uint_as_float(a[i].x); v[i][1] = __uint_as_float(a[i].z); }
          ln_row_store(v, lnv + 2 * D, lnv + 3 * D, sm.y[n], sm.xn[n]);
        }
        csync();
        WS_TL();  // 5: residual sum 1 gathered + LayerNorm 1
        wait_region(2, slot);
        float acc[4];
        mma_tile<4, XS8>(sb + REG_W1 + (warp * 4) * 512, &sm.xn[0][warp * 64], acc);
        red_store(sm.red, acc);
        csync();
        if (tid == 0) issue(u + 2, 2);
        if (tid < 128) {
          const int n = tid >> 4, fl = tid & 15;
          float hv = 0.f;
          if (n < R) {
#pragma unroll
            for (int w = 0; w < NWW; ++w) hv += sm.red[w][fl][n];
            hv = fmaxf(hv + cb[80 + fl], 0.f);
          }
          const float nb = __shfl_down_sync(0xffffffffu, hv, 1);
          if (n < R && (fl & 1) == 0) {
            const uint32_t pk = pack_bf2(hv, nb);
#pragma unroll
            for (int r = 0; r < R3; ++r) st_cell(X3 + ((size_t)r * RW + n) * (FF / 2) + ro.w1_tile * 8 + (fl >> 1), pk, tag);
          }
        }
      }
      else { WS_TL(); }
      WS_TL();  // 6: linear1 tile published
      // ---- s = 2..5: linear2 tile over one K-half (+ bias + residual on half 0) -> hand-off X4 (fp32 partial of residual sum 2)
      if (ro.w2_tile >= 0) {
        if (warp < R) {
          const int n = warp;
          uint4 a[8];
          poll_vec<8>(reinterpret_cast<const uint4*>(X3 + ((size_t)(cta % R3) * RW + n) * (FF / 2) + (size_t)ro.w2_half * (FF / 4)), 256, tag, a, watch);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint2*>(&sm.u.hh[n][4 * (i * 32 + lane)]) = make_uint2(a[i].x, a[i].z);
        }
        csync();
        WS_TL();  // 7: FFN hidden gathered
        wait_region(1, slot);
        float acc[4];
        mma_tile<8, HS8>(sb + REG_B + (warp * 8) * 512, &sm.u.hh[0][warp * 128], acc);
        red_store(sm.red, acc);
        csync();
        if (tid == 0) issue(u + 2, 1);
        if (tid < 128) {
          const int n = tid >> 4, fl = tid & 15;
          if (n < R) {
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < NWW; ++w) sacc += sm.red[w][fl][n];
            if (ro.w2_half == 0) sacc += cb[96 + fl] + sm.y[n][ro.w2_tile * 16 + fl];
#pragma unroll
            for (int r = 0; r < R4; ++r)
              st_cell(X4 + (((size_t)r * 2 + ro.w2_half) * RW + n) * D + ro.w2_tile * 16 + fl, __float_as_uint(sacc), tag);
          }
        }
      }
      else { WS_TL(); }
      // (the CTA barrier at the top of the next layer - after its gather - separates this layer's readers of y / xn / red from the
      //  next layer's writers: the gather of X4 cannot complete before this CTA's own X4 cells, stored above, have landed)
      csync();
      WS_TL();  // 8: linear2 partial published
    }
    // ---- head: 1-2 vocabulary tiles per CTA (s >= 6) -> logits in global memory
    if (ro.nht > 0) {
      wait_region(0, u & 1);
      const unsigned char* sb = sm.slot[u & 1];
      for (int i = 0; i < ro.nht; ++i) {
        float acc[4];
        mma_tile<4, XS8>(sb + REG_Q + i * TILE_B + (warp * 4) * 512, &sm.xn[0][warp * 64], acc);
        red_store(sm.red, acc);
        csync();
        if (tid < 128) {
          const int n = tid >> 4, fl = tid & 15, f = ro.ht[i] * 16 + fl;
          if (n < R && f < V) {
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < NWW; ++w) sacc += sm.red[w][fl][n];
            c.logits[(size_t)n * VPAD + f] = sacc;
          }
        }
        csync();
      }
      if (tid == 0) issue(u + 2, 0);
    } else if (tid == 0) {
      issue(u + 2, 0);  // the q region of this slot is idle during a head unit
    }
    if (tid == 0) { issue(u + 2, 1); issue(u + 2, 2); }  // B and W1 regions of the head unit's slot: free since unit u - 2
    ++u;
    WS_TL();
    gbar.sync();
    WS_TL();
    // ---- sampler: CTA r takes sequence r (scratch aliases y | xn, rewritten by the next step's prologue)
    if (cta < R) {
      stopped = sample_row<1>(c, cta, step, ss, sm.row_slot[cta]);
      if (stopped && tid == 0) c.seg_cnt[step % 3] = 1;
    }
    WS_TL();
    gbar.sync();
    WS_TL();
    if (ld_cg_i(c.seg_cnt + step % 3) != 0) {
      if (cta == 0) phase_plan<1>(c, reinterpret_cast<int*>(&sm.y[0][0]));
      gbar.sync();
      if (cta == 0 && tid == 0) c.seg_cnt[step % 3] = 0;
      fresh = true;
    } else {
      const int pub_pos = (cta < R && tid == 0) ? sm.row_pos[cta] : 0;
      __syncwarp();
      if (cta < R && tid == 0) {
        const int slotb = sm.row_slot[cta], pos = pub_pos + 1;
        c.seq_len[slotb] = pos + 1;
        c.row_pos[cta] = pos;
        c.row_kvoff[cta] = kv_row_off(sm.pt[cta][pos >> PAGE_SHIFT], pos & (PAGE - 1));
        atomicAdd(c.stats + 0, (unsigned long long)(pos + 1));
      }
      if (cta == 0 && tid == 0) {
        *c.step = step + 1;
        atomicAdd(c.stats + 1, 1ull);
        atomicAdd(c.stats + 2, (unsigned long long)R);
      }
      if (tid < R) {
        const int pos = sm.row_pos[tid] + 1;
        sm.row_pos[tid] = pos;
        sm.row_kvoff[tid] = kv_row_off(sm.pt[tid][pos >> PAGE_SHIFT], pos & (PAGE - 1));
      }
      csync();
      fresh = false;
    }
  }
#undef WS_TL
  // ---- drain: the copies issued for the two units that will not run must land before the CTA may exit; the next launch
  //      continues the tag sequence above everything this one wrote
  if (tid == 0) {
    for (int k = 0; k < 2; ++k)
      for (int r = 0; r < 4; ++r)
        if (region_used(r, (u + k) % UPS) && (r < 3 || u + k <= vec_issued)) wait_region(r, (u + k) & 1);
    if (cta == 0 && it > 0) *reinterpret_cast<unsigned*>(ll) = ebase + (uint32_t)it * TAG_STRIDE;
  }
  __syncthreads();
}

}  // namespace ws
}  // namespace t2s
