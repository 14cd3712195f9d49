// phases.cuh — the decode step as grid-wide "phases".  Every phase is a __device__ function over a
// virtual grid (cta, ncta) so the same code runs (a) as one kernel per phase (CUDA-graph replay) and
// (b) inside the persistent cooperative kernel, separated by grid barriers.
//
// Reference being replaced: T2SBlock.decode_next_token / process_prompt projections and FFN
// (GPT_SoVITS/AR/models/t2s_model.py:135-221), ar_predict_layer (:706/:884), sample()
// (AR/models/utils.py:140-199) and the retirement bookkeeping of infer_panel_batch_infer (:720-763).
#pragma once
#include "common.cuh"

namespace t2s {

// =====================================================================================================
// Skinny projections  Y[rows, features] = X[rows, K] * W[features, K]^T   (+ fused LayerNorm, epilogues)
//
// Work unit = one 16-feature tile over the whole K: its weights are one contiguous block (16 KB for
// K=512, 64 KB for K=2048) stored in m16n8k16 A-fragment order, so a warp loads a fragment with one
// coalesced 128-bit load per lane.  Inside a unit the 8 warps split K (KBW k-blocks each), rows are
// processed in tiles of 32 (4 MMA n-tiles), the warps' partial sums are reduced in shared memory in a
// fixed order (no floating-point atomics anywhere: results are bit-reproducible) and the epilogue is
// fused.  Units are spread over all CTAs, so every SM streams a disjoint slice of the weights; the next
// phase's slice is prefetched into L2 before the grid barrier (prefetch_unit).
//
// Post-LN without an extra phase and without every consumer re-normalising fp32 rows: the PRODUCER of a
// residual sum y (O-proj, FFN2) writes y in fp32, a bf16 copy, and per-16-feature-tile partial
// (sum, sum of squares) of each row.  The CONSUMER stages the raw bf16 rows (half the bytes, no math),
// rebuilds mean / rstd from the 32 partials, and applies LayerNorm in its epilogue:
//     LN(y) W^T + b = rstd * (y Wg^T - mean * c1) + c0,   Wg = W * gamma,  c1 = Wg 1,  c0 = W beta + b
// (gamma is folded into the packed weights once, k_fold_ln in kernels.cuh).
// =====================================================================================================
enum { OUT_QKV = 0, OUT_O = 1, OUT_FFN1 = 2, OUT_FFN2 = 3, OUT_HEAD = 4, OUT_BERT = 5 };

// shared-memory carve-up for a projection with KBW k-blocks per warp (K = KBW*128)
template <int KBW>
struct ProjLayout {
  static constexpr int K = KBW * 128;
  static constexpr int XSK = K + 8;  // bf16 row stride: (K+8)/2 words = 4 mod 32 -> conflict-free B fragments
  static constexpr size_t XS_BYTES = (size_t)RT * XSK * 2;
  static constexpr size_t RED_BYTES = (size_t)NW * 16 * (RT + 1) * 4;
  static constexpr size_t BYTES = XS_BYTES + RED_BYTES + RT * 8 + RT * 8;
  __device__ static bf16* xs(unsigned char* base) { return reinterpret_cast<bf16*>(base); }
  __device__ static float* red(unsigned char* base) { return reinterpret_cast<float*>(base + XS_BYTES); }
  __device__ static long long* kvoff(unsigned char* base) { return reinterpret_cast<long long*>(base + XS_BYTES + RED_BYTES); }
  __device__ static float2* st(unsigned char* base) { return reinterpret_cast<float2*>(base + XS_BYTES + RED_BYTES + RT * 8); }
};
constexpr size_t SMEM_PROJ_MAX = ProjLayout<16>::BYTES;

struct ProjArgs {
  const bf16* w;        // packed (gamma-folded) weights of this matrix
  int n_tiles;          // 16-feature tiles
  int row_groups;       // > 1: units = tiles x row groups (phases with few tiles: more CTAs, less staging each)
  const bf16* in_b16;   // source rows [., K] bf16 (raw residual sums, or attention / FFN hidden)
  const int* in_idx;    // optional row gather (layer-0 input by slot, head rows after prefill)
  const float2* sp;     // partial row statistics of the source [., 32] -> LayerNorm folded in; NULL: no LN
  const float* c1;      // LN: Wg 1
  const float* c0;      // LN: W beta + bias;  no LN: bias (may be NULL)
  const float* res_g;   // OUT_O: previous layer's norm2 (residual recompute)
  const float* res_b;
  const float* g1;      // OUT_FFN1: this layer's norm1 + linear2.bias (y2 initialisation duty)
  const float* be1;
  const float* b2;
  const int* out_idx;   // OUT_BERT: text row -> prompt row
  int layer;
};

// L2 prefetch of one unit's weight block (issued before a grid barrier for the NEXT phase)
__device__ __forceinline__ void prefetch_unit(const bf16* w, int u, int k, int tpc) {
  const size_t bytes = (size_t)16 * tpc * k * 2;
  const unsigned char* base = reinterpret_cast<const unsigned char*>(w) + (size_t)u * bytes;
  for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)NT * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
}

template <int OUT, int KBW>
__device__ __forceinline__ void proj_stage(const Ctx& c, const ProjArgs& a, unsigned char* smem, int r0, int n_rows, int nt) {
  using LY = ProjLayout<KBW>;
  bf16* xs = LY::xs(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NR = RT / NW;                      // rows per warp
  constexpr int CPL = LY::K / 256;                 // 16-byte chunks per lane per row
  constexpr int RG = (16 / CPL) < NR ? (16 / CPL) : NR;  // rows in flight per warp (<= 16 loads per lane)
  // ---- every independent load is issued first (volatile loads keep program order), consumed afterwards:
  //      one L2 round trip instead of three (kv offsets, rows, statistics)
  int ri[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int r = r0 + warp + NW * i;
    ri[i] = (r < n_rows) ? (a.in_idx ? ld_cg_i(a.in_idx + r) : r) : -1;
  }
  long long kvo = 0;
  if (OUT == OUT_QKV && threadIdx.x < RT && r0 + (int)threadIdx.x < n_rows) kvo = ldv_cg_ll(c.row_kvoff + r0 + threadIdx.x);
  float2 part[NR];
  if (a.sp) {
#pragma unroll
    for (int i = 0; i < NR; ++i) part[i] = (ri[i] >= 0) ? ldv_cg_f2(a.sp + (size_t)ri[i] * 32 + lane) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int i0 = 0; i0 < NR; i0 += RG) {
    uint4 v[RG][CPL];
#pragma unroll
    for (int i = 0; i < RG; ++i) {
      const bf16* src = a.in_b16 + (size_t)max(ri[i0 + i], 0) * LY::K;
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[i][j] = (ri[i0 + i] >= 0) ? ldv_cg16(src + (lane + 32 * j) * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < RG; ++i) {
      const int rl = warp + NW * (i0 + i);
#pragma unroll
      for (int j = 0; j < CPL; ++j) *reinterpret_cast<uint4*>(xs + rl * LY::XSK + (lane + 32 * j) * 8) = v[i][j];
    }
  }
  if (OUT == OUT_QKV && threadIdx.x < RT) LY::kvoff(smem)[threadIdx.x] = kvo;
  if (a.sp) {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const float s = warp_sum(part[i].x), q = warp_sum(part[i].y);
      const float mean = s * (1.0f / D);
      const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
      const float rstd = 1.0f / sqrtf(var + LN_EPS);
      if (lane == 0) {
        const int rl = warp + NW * i;
        LY::st(smem)[rl] = make_float2(mean, rstd);
        if (OUT == OUT_QKV && nt == 0 && ri[i] >= 0) c.stat2[r0 + rl] = make_float2(mean, rstd);
      }
    }
  }
}

// 16-lane (one row, 16 features) reduction of (v, v*v): the producer's partial LayerNorm statistics
__device__ __forceinline__ void row_partial_stats(float v, float2* dst, bool write) {
  float s = v, q = v * v;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (write && (threadIdx.x & 15) == 0) *dst = make_float2(s, q);
}

template <int OUT, int KBW>
__device__ void proj_phase(const Ctx& c, const ProjArgs& a, int n_rows, int cta, int ncta, unsigned char* smem) {
  using LY = ProjLayout<KBW>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  constexpr int KB_ROW = KBW * NW;  // k-blocks per feature tile
  bf16* xs = LY::xs(smem);
  float (*red)[16][RT + 1] = reinterpret_cast<float (*)[16][RT + 1]>(LY::red(smem));
  const long long* kvoff = LY::kvoff(smem);
  const float2* st = LY::st(smem);
  const bool ln = a.sp != nullptr;
  // units = feature tiles x row groups; a row group is a multiple of 8 rows (one MMA n-tile)
  const int rgn = (a.row_groups > 1) ? min(a.row_groups, (n_rows + 7) >> 3) : 1;
  const int rows_pg = (((n_rows + rgn - 1) / rgn) + 7) & ~7;
  for (int u = cta; u < a.n_tiles * rgn; u += ncta) {
    const int nt = u % a.n_tiles, rg = u / a.n_tiles;
    const int row_beg = rg * rows_pg, row_end = min(n_rows, row_beg + rows_pg);
    if (row_beg >= row_end) continue;
    // this warp's KBW A fragments (k-blocks warp*KBW .. +KBW-1 of feature tile nt), loaded once per unit
    uint4 af[KBW];
    const uint4* wp = reinterpret_cast<const uint4*>(a.w) + ((size_t)nt * KB_ROW + warp * KBW) * 32 + lane;
#pragma unroll
    for (int i = 0; i < KBW; ++i) af[i] = ld_weight16(wp + i * 32);
    // per-thread epilogue constants: this thread always reduces feature nt*16 + (tid & 15)
    const int fme = nt * 16 + (threadIdx.x & 15);
    const float c0 = a.c0 ? a.c0[fme] : 0.f;
    const float c1 = ln ? a.c1[fme] : 0.f;
    const float resg = (OUT == OUT_O && a.res_g) ? a.res_g[fme] : 1.f;
    const float resb = (OUT == OUT_O && a.res_b) ? a.res_b[fme] : 0.f;
    for (int r0 = row_beg; r0 < row_end; r0 += RT) {
      const int rows_here = min(RT, row_end - r0);
      const int n8 = (rows_here + 7) >> 3;
      __syncthreads();  // previous readers of xs / red / kvoff / st are done
      proj_stage<OUT, KBW>(c, a, smem, r0, row_end, nt);
      __syncthreads();
      if (OUT == OUT_FFN1 && threadIdx.x < rows_here) {
        // duty of FFN1's unit nt: y2[:, 4nt..4nt+3] := LN1(y1) + b2 (FFN2 then adds its product; single writer)
        const int r = r0 + threadIdx.x;
        const float2 s2 = st[threadIdx.x];
        const float4 y = ld_cg_f4(c.y1 + (size_t)r * D + 4 * nt);
        const float4 gg = *reinterpret_cast<const float4*>(a.g1 + 4 * nt);
        const float4 bb = *reinterpret_cast<const float4*>(a.be1 + 4 * nt);
        const float4 b2 = *reinterpret_cast<const float4*>(a.b2 + 4 * nt);
        *reinterpret_cast<float4*>(c.y2 + (size_t)r * D + 4 * nt) =
            make_float4((y.x - s2.x) * s2.y * gg.x + bb.x + b2.x, (y.y - s2.x) * s2.y * gg.y + bb.y + b2.y,
                        (y.z - s2.x) * s2.y * gg.z + bb.z + b2.z, (y.w - s2.x) * s2.y * gg.w + bb.w + b2.w);
      }
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
      for (int i = 0; i < KBW; ++i) {
        const int k0 = (warp * KBW + i) * 16;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          if (n < n8) {
            const uint32_t* xr = reinterpret_cast<const uint32_t*>(xs + (n * 8 + g) * LY::XSK + k0);
            mma_bf16_16816(acc[n], af[i], xr[t], xr[4 + t]);
          }
        }
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        red[warp][g][n * 8 + 2 * t] = acc[n][0];
        red[warp][g][n * 8 + 2 * t + 1] = acc[n][1];
        red[warp][g + 8][n * 8 + 2 * t] = acc[n][2];
        red[warp][g + 8][n * 8 + 2 * t + 1] = acc[n][3];
      }
      __syncthreads();
#pragma unroll
      for (int it = 0; it < 16 * RT / NT; ++it) {
        const int o = threadIdx.x + it * NT;
        const int fl = o & 15, n = o >> 4;  // a 16-lane group = one row, 16 features
        const bool ok = n < rows_here;
        const int r = r0 + n, f = nt * 16 + fl;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w][fl][n];
        float val = s + c0;
        if (ln) { const float2 s2 = st[n]; val = s2.y * (s - s2.x * c1) + c0; }
        if (OUT == OUT_QKV) {
          if (ok) {
            if (f < D) {
              c.q[(size_t)r * D + f] = val * QSCALE;
            } else {
              const size_t off = (size_t)a.layer * c.kv_layer_stride + (size_t)kvoff[n];
              if (f < 2 * D) c.kpool[off + kv_feat(kvoff[n], f - D)] = __float2bfloat16_rn(val);
              else c.vpool[off + kv_feat(kvoff[n], f - 2 * D)] = __float2bfloat16_rn(val);
            }
          }
        } else if (OUT == OUT_O) {
          float y = 0.f;
          if (ok) {
            float res;
            if (a.layer == 0) {
              const int ri = c.x0_by_slot ? ld_cg_i(c.row_slot + r) : r;
              res = ld_cg_f(c.x0 + (size_t)ri * D + f);
            } else {
              const float2 s2 = __ldcg(c.stat2 + r);
              res = (ld_cg_f(c.y2 + (size_t)r * D + f) - s2.x) * s2.y * resg + resb;
            }
            y = res + val;
            c.y1[(size_t)r * D + f] = y;
            c.yb1[(size_t)r * D + f] = __float2bfloat16_rn(y);
          }
          row_partial_stats(y, c.sp1 + (size_t)r * 32 + nt, ok);
        } else if (OUT == OUT_FFN1) {
          if (ok) c.h[(size_t)r * FF + f] = __float2bfloat16_rn(fmaxf(val, 0.f));
        } else if (OUT == OUT_FFN2) {
          float y = 0.f;
          if (ok) {
            float* p = c.y2 + (size_t)r * D + f;
            y = __ldcg(p) + s;  // y2 was initialised to LN1(y1) + b2 by FFN1's duty; single writer
            *p = y;
            c.yb2[(size_t)r * D + f] = __float2bfloat16_rn(y);
          }
          row_partial_stats(y, c.sp2 + (size_t)r * 32 + nt, ok);
        } else if (OUT == OUT_HEAD) {
          if (ok && f < V) c.logits[(size_t)r * VPAD + f] = val;
        } else if (OUT == OUT_BERT) {
          if (ok) {
            const size_t o2 = (size_t)a.out_idx[r] * D + f;
            const float y = __ldcg(c.x0 + o2) + s;
            c.x0[o2] = y;
            c.x0b[o2] = __float2bfloat16_rn(y);
          }
        }
      }
    }
  }
}

// ---- the five per-layer projection phases + head, with their argument wiring ------------------------
__device__ __forceinline__ void phase_qkv(const Ctx& c, int layer, int n_rows, int cta, int ncta, unsigned char* sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_WQKV;
  a.n_tiles = 3 * D / 16; a.layer = layer;
  a.c0 = c.wvec + (size_t)layer * LV + VO_C0_QKV;  // = in_proj_bias for layer 0
  if (layer == 0) {
    a.in_b16 = c.x0b; a.in_idx = c.x0_by_slot ? c.row_slot : nullptr;
  } else {
    a.in_b16 = c.yb2; a.sp = c.sp2;
    a.c1 = c.wvec + (size_t)layer * LV + VO_C1_QKV;
  }
  proj_phase<OUT_QKV, 4>(c, a, n_rows, cta, ncta, sm);
}
__device__ __forceinline__ void phase_oproj(const Ctx& c, int layer, int n_rows, int cta, int ncta, unsigned char* sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_WO;
  a.n_tiles = D / 16; a.layer = layer; a.row_groups = 4;
  a.in_b16 = c.attn;
  a.c0 = c.wvec + (size_t)layer * LV + VO_BO;
  if (layer > 0) {
    a.res_g = c.wvec + (size_t)(layer - 1) * LV + VO_G2;
    a.res_b = c.wvec + (size_t)(layer - 1) * LV + VO_BE2;
  }
  proj_phase<OUT_O, 4>(c, a, n_rows, cta, ncta, sm);
}
__device__ __forceinline__ void phase_ffn1(const Ctx& c, int layer, int n_rows, int cta, int ncta, unsigned char* sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_W1;
  a.n_tiles = FF / 16; a.layer = layer;
  a.in_b16 = c.yb1; a.sp = c.sp1;
  a.c1 = c.wvec + (size_t)layer * LV + VO_C1_FFN1;
  a.c0 = c.wvec + (size_t)layer * LV + VO_C0_FFN1;
  a.g1 = c.wvec + (size_t)layer * LV + VO_G1;
  a.be1 = c.wvec + (size_t)layer * LV + VO_BE1;
  a.b2 = c.wvec + (size_t)layer * LV + VO_B2;
  proj_phase<OUT_FFN1, 4>(c, a, n_rows, cta, ncta, sm);
}
__device__ __forceinline__ void phase_ffn2(const Ctx& c, int layer, int n_rows, int cta, int ncta, unsigned char* sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_W2;
  a.n_tiles = D / 16; a.layer = layer; a.row_groups = 4;
  a.in_b16 = c.h;
  proj_phase<OUT_FFN2, 16>(c, a, n_rows, cta, ncta, sm);
}
__device__ __forceinline__ void phase_head(const Ctx& c, int n_rows, int cta, int ncta, unsigned char* sm) {
  ProjArgs a{};
  a.w = c.whead;
  a.n_tiles = VT; a.layer = c.n_layer;
  a.in_b16 = c.yb2; a.sp = c.sp2; a.in_idx = c.head_rows;
  a.c1 = c.head_c1; a.c0 = c.head_c0;
  proj_phase<OUT_HEAD, 4>(c, a, n_rows, cta, ncta, sm);
}

// L2 prefetch of the weights this CTA will need in the phase AFTER the coming one (called right before a
// grid barrier so HBM latency overlaps the barrier and the next phase).  `next` = the phase about to be
// entered after the barrier is `cur + 1`; we prefetch for cur + 2's projection when cur + 1 has no weights.
enum { PF_WO = 0, PF_W1 = 1, PF_W2 = 2, PF_WQKV_NEXT = 3, PF_HEAD = 4 };
__device__ __forceinline__ void prefetch_phase(const Ctx& c, int what, int layer, int cta) {
  const bf16* wl = c.wmat + (size_t)layer * LW;
  if (what == PF_WO) { if (cta < D / 16) prefetch_unit(wl + OFF_WO, cta, D, 1); }
  else if (what == PF_W1) { if (cta < FF / 16) prefetch_unit(wl + OFF_W1, cta, D, 1); }
  else if (what == PF_W2) { if (cta < D / 16) prefetch_unit(wl + OFF_W2, cta, FF, 1); }
  else if (what == PF_WQKV_NEXT) {
    if (layer + 1 < c.n_layer) { if (cta < 3 * D / 16) prefetch_unit(wl + LW + OFF_WQKV, cta, D, 1); }
    else if (cta < VT) prefetch_unit(c.whead, cta, D, 1);
  }
}

// =====================================================================================================
// Decode attention: one query per active sequence against its paged bf16 KV cache, all 16 heads at
// once.  The flattened list of (sequence, CH-position chunk) items is split evenly over the CTAs
// (split-KV); a CTA keeps an online-softmax state in registers while it stays on one sequence, the
// partial states of a sequence are merged by the last CTA to finish it.
//
// Lane mapping: lane l loads the 16-byte chunk (head l/4, dims 8*(l%4)..+7) and the one of head 8+l/4 of a position;
// pages are head-major (kv_row_off), so a quad of lanes reads one 64 B head row.
// =====================================================================================================
struct AttnSmem {
  int pre[MAX_B + 1];
  float m[NW][NH];
  float l[NW][NH];
  float acc[NW][D];
  int scratch[NW + 2];
};

__device__ __forceinline__ void attn_merge_write(const Ctx& c, AttnSmem& sm, int r, int seg, int count) {
  // combine the 8 warps' states for feature d, d+256; then either finish or publish a partial
  const int tid = threadIdx.x;
  float M[2], L[2], A[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int d = tid + 256 * k, hd = d >> 5;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) mm = fmaxf(mm, sm.m[w][hd]);
    float ll = 0.f, aa = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float sc = exp2f(sm.m[w][hd] - mm);
      ll += sm.l[w][hd] * sc;
      aa += sm.acc[w][d] * sc;
    }
    M[k] = mm; L[k] = ll; A[k] = aa;
  }
  if (count == 1) {
#pragma unroll
    for (int k = 0; k < 2; ++k) c.attn[(size_t)r * D + tid + 256 * k] = __float2bfloat16_rn(A[k] / L[k]);
    return;
  }
  float* p = c.part + (size_t)seg * PART_STRIDE;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int d = tid + 256 * k;
    p[2 * NH + d] = A[k];
    if ((d & 31) == 0) { p[d >> 5] = M[k]; p[NH + (d >> 5)] = L[k]; }
  }
}

// K/V rows of U consecutive positions starting at p0 (same page: p0 is a multiple of U <= 4)
template <int U>
__device__ __forceinline__ void attn_load(const bf16* kbase, const bf16* vbase, const int* pt, int p0, int pend, int lane,
                                          uint4 (&ka)[U], uint4 (&kb)[U], uint4 (&va)[U], uint4 (&vb)[U]) {
  const int page = pt[p0 >> PAGE_SHIFT];
  // lane l: 16-byte chunk of head l/4 (dims 8*(l%4)..+7) and of head 8 + l/4 (head-major pages, chunks swizzled by position)
  const int pin = p0 & (PAGE - 1);
  const size_t rowoff = (size_t)kv_row_off(page, pin) + (size_t)(lane >> 2) * KV_HEAD_STRIDE;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (p0 + u < pend) {
      const int ch = ((lane & 3) ^ (((pin + u) >> 1) & 3)) * 8;
      const bf16* kr = kbase + rowoff + (size_t)u * DH + ch;
      const bf16* vr = vbase + rowoff + (size_t)u * DH + ch;
      ka[u] = ld_cg16(kr);
      kb[u] = ld_cg16(kr + 8 * KV_HEAD_STRIDE);
      va[u] = ld_cg16(vr);
      vb[u] = ld_cg16(vr + 8 * KV_HEAD_STRIDE);
    } else {
      ka[u] = kb[u] = va[u] = vb[u] = make_uint4(0, 0, 0, 0);
    }
  }
}

template <int U>
__device__ __forceinline__ void attn_accumulate(int p0, int pend, const uint4 (&ka)[U], const uint4 (&kb)[U],
                                                const uint4 (&va)[U], const uint4 (&vb)[U], const float (&qa)[8],
                                                const float (&qb)[8], float (&m)[2], float (&l)[2], float (&accA)[8],
                                                float (&accB)[8]) {
  float sA[U], sB[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float a = qa[0] * bf_lo(ka[u].x) + qa[1] * bf_hi(ka[u].x) + qa[2] * bf_lo(ka[u].y) + qa[3] * bf_hi(ka[u].y) +
              qa[4] * bf_lo(ka[u].z) + qa[5] * bf_hi(ka[u].z) + qa[6] * bf_lo(ka[u].w) + qa[7] * bf_hi(ka[u].w);
    float b = qb[0] * bf_lo(kb[u].x) + qb[1] * bf_hi(kb[u].x) + qb[2] * bf_lo(kb[u].y) + qb[3] * bf_hi(kb[u].y) +
              qb[4] * bf_lo(kb[u].z) + qb[5] * bf_hi(kb[u].z) + qb[6] * bf_lo(kb[u].w) + qb[7] * bf_hi(kb[u].w);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    const bool ok = (p0 + u < pend);
    sA[u] = ok ? a : -INFINITY;
    sB[u] = ok ? b : -INFINITY;
  }
  float mA = m[0], mB = m[1];
#pragma unroll
  for (int u = 0; u < U; ++u) { mA = fmaxf(mA, sA[u]); mB = fmaxf(mB, sB[u]); }
  const float cA = exp2f(m[0] - mA), cB = exp2f(m[1] - mB);  // position p0 is valid, so mA/mB are finite
  m[0] = mA; m[1] = mB;
  l[0] *= cA; l[1] *= cB;
#pragma unroll
  for (int i = 0; i < 8; ++i) { accA[i] *= cA; accB[i] *= cB; }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float pA = exp2f(sA[u] - mA), pB = exp2f(sB[u] - mB);
    l[0] += pA; l[1] += pB;
    accA[0] += pA * bf_lo(va[u].x); accA[1] += pA * bf_hi(va[u].x);
    accA[2] += pA * bf_lo(va[u].y); accA[3] += pA * bf_hi(va[u].y);
    accA[4] += pA * bf_lo(va[u].z); accA[5] += pA * bf_hi(va[u].z);
    accA[6] += pA * bf_lo(va[u].w); accA[7] += pA * bf_hi(va[u].w);
    accB[0] += pB * bf_lo(vb[u].x); accB[1] += pB * bf_hi(vb[u].x);
    accB[2] += pB * bf_lo(vb[u].y); accB[3] += pB * bf_hi(vb[u].y);
    accB[4] += pB * bf_lo(vb[u].z); accB[5] += pB * bf_hi(vb[u].z);
    accB[6] += pB * bf_lo(vb[u].w); accB[7] += pB * bf_hi(vb[u].w);
  }
}

// One warp walks positions pbeg + U*warp, + U*NW, ... with the NEXT group's loads in flight while the current
// group is reduced (two register buffers, loop unrolled by two so the buffers are never copied).
template <int U>
__device__ __forceinline__ void attn_segment(const Ctx& c, int layer, int slot, int pbeg, int pend,
                                             const float (&qa)[8], const float (&qb)[8],
                                             float (&m)[2], float (&l)[2], float (&accA)[8], float (&accB)[8]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* kbase = c.kpool + (size_t)layer * c.kv_layer_stride;
  const bf16* vbase = c.vpool + (size_t)layer * c.kv_layer_stride;
  const int* pt = c.page_table + slot * c.max_pages;
  constexpr int STEP = U * NW;
  int p0 = pbeg + U * warp;
  if (p0 >= pend) return;
  uint4 ka0[U], kb0[U], va0[U], vb0[U], ka1[U], kb1[U], va1[U], vb1[U];
  attn_load<U>(kbase, vbase, pt, p0, pend, lane, ka0, kb0, va0, vb0);
  for (;;) {
    const int p1 = p0 + STEP;
    if (p1 < pend) attn_load<U>(kbase, vbase, pt, p1, pend, lane, ka1, kb1, va1, vb1);
    attn_accumulate<U>(p0, pend, ka0, kb0, va0, vb0, qa, qb, m, l, accA, accB);
    if (p1 >= pend) break;
    const int p2 = p1 + STEP;
    if (p2 < pend) attn_load<U>(kbase, vbase, pt, p2, pend, lane, ka0, kb0, va0, vb0);
    attn_accumulate<U>(p1, pend, ka1, kb1, va1, vb1, qa, qb, m, l, accA, accB);
    if (p2 >= pend) break;
    p0 = p2;
  }
}

// The split-KV work assignment is the same for all layers of a step, so phase_plan computes it once:
// every CTA gets at most two descriptors {row, pbeg, pend, (j << 16) | count}: its j-th of `count`
// position ranges of `row`.  A row's partials live in c.part[first_cta .. first_cta + count).
__device__ void phase_attn_decode(const Ctx& c, int layer, int n_rows, int cta, int ncta, AttnSmem& sm,
                                  const int4* cached_desc = nullptr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = 0; e < 2; ++e) {
    const int4 ds = cached_desc ? cached_desc[e] : __ldcg(reinterpret_cast<const int4*>(c.attn_desc) + cta * 2 + e);
    if (ds.x < 0) break;
    const int r = ds.x & 0xFFFF, slot = ds.x >> 16;
    const int pbeg = ds.y, pend = ds.z, j = ds.w >> 16, count = ds.w & 0xFFFF;
    float qa[8], qb[8];
    {
      const float* qr = c.q + (size_t)r * D;
      float4 x0 = ld_cg_f4(qr + lane * 8), x1 = ld_cg_f4(qr + lane * 8 + 4);
      float4 y0 = ld_cg_f4(qr + 256 + lane * 8), y1 = ld_cg_f4(qr + 256 + lane * 8 + 4);
      qa[0] = x0.x; qa[1] = x0.y; qa[2] = x0.z; qa[3] = x0.w; qa[4] = x1.x; qa[5] = x1.y; qa[6] = x1.z; qa[7] = x1.w;
      qb[0] = y0.x; qb[1] = y0.y; qb[2] = y0.z; qb[3] = y0.w; qb[4] = y1.x; qb[5] = y1.y; qb[6] = y1.z; qb[7] = y1.w;
    }
      float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, accA[8], accB[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { accA[i] = 0.f; accB[i] = 0.f; }
    if (pend - pbeg > 16) attn_segment<4>(c, layer, slot, pbeg, pend, qa, qb, m, l, accA, accB);
    else attn_segment<1>(c, layer, slot, pbeg, pend, qa, qb, m, l, accA, accB);
      __syncthreads();  // previous entry's merge readers are done with sm.m/l/acc
    if ((lane & 3) == 0) {
      sm.m[warp][lane >> 2] = m[0]; sm.l[warp][lane >> 2] = l[0];
      sm.m[warp][8 + (lane >> 2)] = m[1]; sm.l[warp][8 + (lane >> 2)] = l[1];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { sm.acc[warp][lane * 8 + i] = accA[i]; sm.acc[warp][256 + lane * 8 + i] = accB[i]; }
    __syncthreads();
    attn_merge_write(c, sm, r, cta, count);
      if (count > 1) {
      __syncthreads();
      if (tid == 0) {
        // release our partial, acquire the others' (acq_rel RMW at gpu scope)
        int old;
        asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(c.seg_cnt + r) : "memory");
        sm.scratch[NW] = (old == count - 1);
      }
      __syncthreads();
          if (sm.scratch[NW]) {  // last CTA of this sequence: merge all partials (count <= 16)
        const float* pbase = c.part + (size_t)(cta - j) * PART_STRIDE;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int d = tid + 256 * k, hd = d >> 5;
          float mv[16], lv[16], av[16];
#pragma unroll
          for (int s = 0; s < 16; ++s) {  // every load in flight at once
            const bool ok = s < count;
            const float* ps = pbase + (size_t)s * PART_STRIDE;
            mv[s] = ok ? ld_cg_f(ps + hd) : -INFINITY;
            lv[s] = ok ? ld_cg_f(ps + NH + hd) : 0.f;
            av[s] = ok ? ld_cg_f(ps + 2 * NH + d) : 0.f;
          }
          float mm = -INFINITY;
#pragma unroll
          for (int s = 0; s < 16; ++s) mm = fmaxf(mm, mv[s]);
          float ll = 0.f, aa = 0.f;
#pragma unroll
          for (int s = 0; s < 16; ++s) {  // fixed order: deterministic
            const float sc = exp2f(mv[s] - mm);
            ll += lv[s] * sc;
            aa += av[s] * sc;
          }
          c.attn[(size_t)r * D + d] = __float2bfloat16_rn(aa / ll);
        }
        if (tid == 0) c.seg_cnt[r] = 0;
      }
        }
  }
}

// =====================================================================================================
// Sampler: one CTA per active sequence.  Restates utils.py:147-199 in the reference's order:
// repetition penalty (in place) -> argmax of the penalised logits (EOS test) -> top-p on the
// un-tempered logits -> /temperature -> top-k by pivot (ties kept) -> softmax -> argmax(p/q), q~Exp(1)
// from Philox4x32-10(counter = (i/4, step, slot, 0), key = seed).  Then the stop rule and the next
// step's input embedding (t2s_model.py:718-769 / :893-914).
// =====================================================================================================
constexpr int SV = 5;  // values per thread: 5*256 >= 1025

// CTA-wide sync of the NT sampler threads: SB = 0 -> the whole CTA (__syncthreads); SB = 1 -> named barrier 1 over
// the first NT threads (cluster-stream kernel: the CTA has an extra TMA producer warp that must not take part)
template <int SB>
__device__ __forceinline__ void cta_sync() {
  if (SB == 0) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

struct SampSmem {
  float redf[NW];
  int redi[NW];
  unsigned hist[256];
  int bcast[4];
  unsigned long long sortbuf[2048];  // top-p path only: (key << 32 | index)
  float cum[2048 / 8];
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint32_t float_key(float f) {  // order-preserving float -> uint
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// block-wide (max value, smallest index among maxima)
template <int SB>
__device__ __forceinline__ int block_argmax(float v, int idx, SampSmem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  cta_sync<SB>();
  if (lane == 0) { sm.redf[warp] = v; sm.redi[warp] = idx; }
  cta_sync<SB>();
  float bv = sm.redf[0];
  int bi = sm.redi[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) {
    const float ov = sm.redf[w];
    const int oi = sm.redi[w];
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  return bi;
}
template <int SB>
__device__ __forceinline__ float block_max(float v, SampSmem& sm) {
  v = warp_max(v);
  cta_sync<SB>();
  if ((threadIdx.x & 31) == 0) sm.redf[threadIdx.x >> 5] = v;
  cta_sync<SB>();
  float r = sm.redf[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) r = fmaxf(r, sm.redf[w]);
  return r;
}
template <int SB>
__device__ __forceinline__ float block_sum(float v, SampSmem& sm) {  // fixed order => deterministic
  v = warp_sum(v);
  cta_sync<SB>();
  if ((threadIdx.x & 31) == 0) sm.redf[threadIdx.x >> 5] = v;
  cta_sync<SB>();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) r += sm.redf[w];
  return r;
}

// k-th largest key among the thread-distributed values (8-bit radix select, 4 passes)
template <int SB>
__device__ __forceinline__ uint32_t block_kth_largest(const uint32_t (&key)[SV], const bool (&valid)[SV], int k,
                                                      SampSmem& sm) {
  uint32_t prefix = 0, mask = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    cta_sync<SB>();
    sm.hist[tid] = 0;  // NT == 256 bins
    cta_sync<SB>();
    // warp-aggregated histogram: logits crowd into a few bins of the leading byte, and same-address shared-memory atomics
    // serialise; lanes with the same bin elect a leader that adds their count once
#pragma unroll
    for (int j = 0; j < SV; ++j) {
      const bool act = valid[j] && (key[j] & mask) == prefix;
      const unsigned bin = (key[j] >> shift) & 0xFFu;
      const unsigned am = __ballot_sync(0xffffffffu, act);
      if (act) {
        const unsigned peers = __match_any_sync(am, bin);
        if (lane == __ffs(peers) - 1) atomicAdd(&sm.hist[bin], (unsigned)__popc(peers));
      }
    }
    cta_sync<SB>();
    if (warp == 0) {
      // lane l owns bins 255-8l .. 248-8l (descending); find the bin where the running count reaches k
      unsigned cnt[8], loc = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { cnt[j] = sm.hist[255 - 8 * lane - j]; loc += cnt[j]; }
      unsigned inc = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      const unsigned exc = inc - loc;
      if (exc < (unsigned)k && inc >= (unsigned)k) {
        unsigned run = exc;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (run < (unsigned)k && run + cnt[j] >= (unsigned)k) { sm.bcast[0] = 255 - 8 * lane - j; sm.bcast[1] = k - run; sm.bcast[2] = (int)cnt[j]; }
          run += cnt[j];
        }
      }
    }
    cta_sync<SB>();
    prefix |= (uint32_t)sm.bcast[0] << shift;
    mask |= 0xFFu << shift;
    k = sm.bcast[1];
    const int left = sm.bcast[2];  // keys that share the prefix found so far
    if (pass < 3 && left <= 32) {
      // Few candidates left (typical after two passes: logits that agree in their leading 16 bits): gather them and let
      // every warp rank them with shuffles - the k-th largest VALUE is the candidate with (#greater) < k <= (#greater-or-equal).
      cta_sync<SB>();
      if (tid == 0) sm.bcast[3] = 0;
      cta_sync<SB>();
#pragma unroll
      for (int j = 0; j < SV; ++j)
        if (valid[j] && (key[j] & mask) == prefix) sm.hist[atomicAdd(reinterpret_cast<unsigned*>(&sm.bcast[3]), 1u)] = key[j];
      cta_sync<SB>();
      const uint32_t ck = lane < left ? sm.hist[lane] : 0u;
      int gt = 0, ge = 0;
      for (int i = 0; i < left; ++i) {
        const uint32_t o = __shfl_sync(0xffffffffu, ck, i);
        gt += o > ck; ge += o >= ck;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, lane < left && gt < k && k <= ge);
      return __shfl_sync(0xffffffffu, ck, __ffs(hit) - 1);
    }
  }
  return prefix;
}

// top-p (rare path): full descending bitonic sort of (key, index), softmax over the sorted values,
// inclusive prefix sum, remove where cum > top_p except the first (utils.py:169-179).
template <int SB>
__device__ void block_top_p(float (&x)[SV], const bool (&valid)[SV], int width, float top_p, SampSmem& sm) {
  const int tid = threadIdx.x;
  for (int i = tid; i < 2048; i += NT) sm.sortbuf[i] = 0ull;  // key 0 sorts last
  cta_sync<SB>();
#pragma unroll
  for (int j = 0; j < SV; ++j) {
    const int i = tid + NT * j;
    // ties: smaller index first in descending order (stable sort of -x) => store ~index in the low word
    if (valid[j]) sm.sortbuf[i] = ((unsigned long long)float_key(x[j]) << 32) | (uint32_t)(0xFFFFFFFFu - i);
  }
  cta_sync<SB>();
  for (int size = 2; size <= 2048; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < 1024; t += NT) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = sm.sortbuf[lo], b = sm.sortbuf[hi];
        if ((a < b) == desc) { sm.sortbuf[lo] = b; sm.sortbuf[hi] = a; }
      }
      cta_sync<SB>();
    }
  }
  // softmax over sorted values: max is element 0
  auto key_to_float = [](uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); };
  const float mx = key_to_float((uint32_t)(sm.sortbuf[0] >> 32));
  // thread t owns sorted positions 8t..8t+7 (2048/256)
  float e[8], loc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = 8 * tid + j;
    e[j] = (i < width) ? expf(key_to_float((uint32_t)(sm.sortbuf[i] >> 32)) - mx) : 0.f;
    loc += e[j];
  }
  const float total = block_sum<SB>(loc, sm);
  loc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { e[j] = e[j] / total; loc += e[j]; }
  cta_sync<SB>();
  sm.cum[tid] = loc;
  cta_sync<SB>();
  float run = 0.f;
  for (int i = 0; i < tid; ++i) run += sm.cum[i];
  cta_sync<SB>();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = 8 * tid + j;
    run += e[j];
    // cum_probs > top_p is removed, position 0 always kept (utils.py:172-173); tag by clearing the key
    if ((i < width) && (i > 0) && (run > top_p)) sm.sortbuf[i] &= 0xFFFFFFFFull;
  }
  cta_sync<SB>();
  // each thread looks for its own elements: build a removal bitmap in cum[] reinterpret (1025 bits)
  unsigned* bits = reinterpret_cast<unsigned*>(sm.cum);
  for (int i = tid; i < 64; i += NT) bits[i] = 0u;
  cta_sync<SB>();
  for (int i = tid; i < width; i += NT) {
    const unsigned long long v = sm.sortbuf[i];
    if ((v >> 32) == 0) {
      const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(v & 0xFFFFFFFFull);
      atomicOr(&bits[idx >> 5], 1u << (idx & 31));
    }
  }
  cta_sync<SB>();
#pragma unroll
  for (int j = 0; j < SV; ++j) {
    const int i = tid + NT * j;
    if (valid[j] && ((bits[i >> 5] >> (i & 31)) & 1u)) x[j] = -INFINITY;
  }
  cta_sync<SB>();
}

template <int SB>
// slot_hint >= 0: the caller already knows the row's slot (saves a dependent global load in front of everything else).
// `gstep` is the session's GLOBAL step; the sequence's own step (the reference's idx) is gstep - slot_step0[slot]: utterances
// admitted into a resident session (t2s_admit) start later than the first request's.
__device__ bool sample_row(const Ctx& c, int r, int gstep, SampSmem& sm, int slot_hint = -1) {  // returns: the sequence stopped at this step
  const int tid = threadIdx.x;
  {
    const int slot = slot_hint >= 0 ? slot_hint : ld_cg_i(c.active + r);
    float v[SV];
    bool valid[SV];
    uint32_t sw[SV];  // the "seen" words of the repetition penalty: requested together with the logits (one L2 round trip)
    const uint32_t* seen = c.seen + (size_t)slot * SEEN_WORDS;
    const int step0 = ld_cg_i(c.slot_step0 + slot), P = ld_cg_i(c.slot_P + slot), uid = ld_cg_i(c.slot_uid + slot);  // in flight with the logits
#pragma unroll
    for (int j = 0; j < SV; ++j) {
      const int i = tid + NT * j;
      valid[j] = i < V;
      v[j] = valid[j] ? ld_cg_f(c.logits + (size_t)r * VPAD + i) : -INFINITY;
      sw[j] = (valid[j] && c.rep_pen != 1.0f) ? __ldcg(seen + (i >> 5)) : 0u;
    }
    const int step = gstep - step0;
    const int width = (step < c.eos_window) ? (V - 1) : V;
    // test hooks: rows indexed by slot, or by utterance id (a recycled slot serves several utterances)
    const int hstride = c.hook_rows > 0 ? c.hook_rows : c.B0;
    const int hrow = c.hook_rows > 0 ? ((uid >= 0 && uid < c.hook_rows) ? uid : -1) : slot;
#pragma unroll
    for (int j = 0; j < SV; ++j) {
      const int i = tid + NT * j;
      if (c.logits_rec && step < c.n_logits_rec && i < V && hrow >= 0) c.logits_rec[((size_t)step * hstride + hrow) * V + i] = v[j];
      if (i >= width) { valid[j] = false; v[j] = -INFINITY; }
    }
    // repetition penalty over every distinct previous token (prompt + generated), utils.py:159-167
    if (c.rep_pen != 1.0f) {
#pragma unroll
      for (int j = 0; j < SV; ++j) {
        const int i = tid + NT * j;
        if (valid[j] && ((sw[j] >> (i & 31)) & 1u)) v[j] = (v[j] < 0.f) ? v[j] * c.rep_pen : v[j] / c.rep_pen;
      }
    }
    // argmax of the penalised logits: the reference's EOS test (t2s_model.py:721/:901)
    int greedy;
    {
      float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < SV; ++j) if (valid[j] && (v[j] > bv)) { bv = v[j]; bi = tid + NT * j; }
      greedy = block_argmax<SB>(bv, bi, sm);
    }
    float x[SV];
#pragma unroll
    for (int j = 0; j < SV; ++j) x[j] = v[j];
    if (c.top_p < 1.0f) block_top_p<SB>(x, valid, width, c.top_p, sm);
    const float temp = fmaxf(c.temperature, 1e-5f);
#pragma unroll
    for (int j = 0; j < SV; ++j) x[j] = x[j] / temp;
    {
      const int k = min(c.top_k, width);
      uint32_t key[SV];
#pragma unroll
      for (int j = 0; j < SV; ++j) key[j] = float_key(x[j]);
      const uint32_t pivot = block_kth_largest<SB>(key, valid, k, sm);
#pragma unroll
      for (int j = 0; j < SV; ++j) if (valid[j] && key[j] < pivot) x[j] = -INFINITY;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SV; ++j) if (valid[j]) mx = fmaxf(mx, x[j]);
    mx = block_max<SB>(mx, sm);
    float e[SV], loc = 0.f;
#pragma unroll
    for (int j = 0; j < SV; ++j) { e[j] = valid[j] ? expf(x[j] - mx) : 0.f; loc += e[j]; }
    const float total = block_sum<SB>(loc, sm);
    int tok;
    {
      float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < SV; ++j) {
        const int i = tid + NT * j;
        // tokens outside the top-k / top-p set have probability 0 and can never win the race: no random number needed
        // (the winner is the same as if every element had drawn one)
        if (!valid[j] || e[j] == 0.f) continue;
        uint32_t w[4];
        philox4x32_10((uint32_t)(i >> 2), (uint32_t)step, (uint32_t)uid, 0u, c.seed_lo, c.seed_hi, w);
        const uint32_t word = w[i & 3];
        const float u = ((float)(word >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23
        const float q = -logf(u);
        const float sc = (e[j] / total) / q;
        if (sc > bv) { bv = sc; bi = i; }
      }
      tok = block_argmax<SB>(bv, bi, sm);
    }
    // ---- bookkeeping (t2s_model.py:718-769)
    int emit = tok;
    if (c.forced && step < c.n_forced && hrow >= 0) {
      emit = __ldcg(c.forced + (size_t)hrow * c.n_forced + step);
      if (emit < 0 || emit >= V) {  // a forced id outside the embedding table: flag it and keep the reads in range
        if (tid == 0) atomicExch(c.abort_flag, ABORT_BAD_ID);
        emit = 0;
      }
    }
    bool stop = (emit == V - 1) || (greedy == V - 1);
    if (c.early_stop != -1 && (step + 1) > c.early_stop) stop = true;
    if (step == c.max_steps - 1) stop = true;
    if (tid == 0) {
      c.sampled[(size_t)slot * c.max_steps + step] = tok;
      if (c.greedy_rec) c.greedy_rec[(size_t)slot * c.max_steps + step] = greedy;
      c.gen[(size_t)slot * c.max_steps + step] = emit;
      if (emit >= 0 && emit < V) c.seen[(size_t)slot * SEEN_WORDS + (emit >> 5)] |= 1u << (emit & 31);
      if (stop) { c.done[slot] = 1; c.out_idx[slot] = step; }
    }
    if (!stop) {
      // next input: emb(y[:, -1]) * x_scale(=1) + alpha * pe[P + idx]  (t2s_model.py:766-769)
      const bf16* er = c.emb_audio + (size_t)emit * D;
      const float* pr = c.pe + (size_t)min(P + step, c.pe_len - 1) * D;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int d = tid + NT * k;
        const float xv = __bfloat162float(er[d]) + c.alpha_audio * pr[d];
        c.x0[(size_t)slot * D + d] = xv;
        c.x0b[(size_t)slot * D + d] = __float2bfloat16_rn(xv);
      }
    }
    cta_sync<SB>();
    return stop;
  }
}

__device__ void phase_sample(const Ctx& c, int n_active, int cta, int ncta, SampSmem& sm) {
  const int step = ld_cg_i(c.step);  // global step
  for (int r = cta; r < n_active; r += ncta) sample_row<0>(c, r, step, sm);
}

// Retirement: compact the active list on device (no host round trip), publish next step's rows.
// Runs on one CTA.  (t2s_model.py:724-745 does this with .tolist() + 48 index_selects.)
template <int SB = 0>
__device__ void phase_plan(const Ctx& c, int* smem_i) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = ld_cg_i(c.n_active);
  int slot = -1, keep = 0;
  if (tid < n) { slot = ld_cg_i(c.active + tid); keep = ld_cg_i(c.done + slot) ? 0 : 1; }
  int inc = keep;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) smem_i[warp] = inc;
  cta_sync<SB>();
  int woff = 0, total = 0;
  for (int w = 0; w < NW; ++w) { if (w < warp) woff += smem_i[w]; total += smem_i[w]; }
  unsigned long long kvpos = 0;
  if (keep) {
    const int p = woff + inc - 1;
    const int pos = ld_cg_i(c.seq_len + slot);
    c.active[p] = slot;
    c.row_slot[p] = slot;
    c.row_pos[p] = pos;
    c.row_kvoff[p] = kv_row_off(c.page_table[slot * c.max_pages + (pos >> PAGE_SHIFT)], pos & (PAGE - 1));
    c.seq_len[slot] = pos + 1;
    kvpos = (unsigned long long)(pos + 1);
  }
  if (kvpos) atomicAdd(c.stats + 0, kvpos);
  if (tid == 0) {
    *c.n_active = total;
    *c.n_rows = total;
    if (n > 0) *c.step = ld_cg_i(c.step) + 1;  // graph modes replay whole chunks: steps after the last sequence retired do not count
    if (total > 0) { atomicAdd(c.stats + 1, 1ull); atomicAdd(c.stats + 2, (unsigned long long)total); }
  }
  // ---- split-KV work assignment for the next step's attention (valid for all layers)
  int* np_s = smem_i + 16;        // [MAX_B] positions per new row
  int* slot_s = smem_i + 16 + MAX_B;  // [MAX_B] slot of each new row
  if (keep) { np_s[woff + inc - 1] = (int)kvpos; slot_s[woff + inc - 1] = slot; }
  const int ncta = c.attn_ctas;
  int4* desc = reinterpret_cast<int4*>(c.attn_desc);
  for (int i = tid; i < 2 * ncta; i += NT) desc[i] = make_int4(-1, 0, 0, 0);
  cta_sync<SB>();
  const int n2 = total;
  if (n2 == 0) return;
  if (n2 > ncta) {  // more rows than CTAs: whole rows, round-robin (n2 <= MAX_B < 2 * ncta)
    if (tid < n2) desc[(tid % ncta) * 2 + tid / ncta] = make_int4(tid | (slot_s[tid] << 16), 0, np_s[tid], 1);
    return;
  }
  // Np
  int v = (tid < n2) ? np_s[tid] : 0, tot = v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  cta_sync<SB>();
  if (lane == 0) smem_i[warp] = tot;
  cta_sync<SB>();
  long long Np = 0;
  for (int w = 0; w < NW; ++w) Np += smem_i[w];
  // CTAs per row: 1 + share of the spare CTAs, capped by 16 partials and by the number of 8-position items
  int k = 0;
  if (tid < n2) {
    k = 1 + (int)(((long long)(ncta - n2) * v) / Np);
    k = min(k, 16);
    k = min(k, (v + 7) >> 3);
  }
  int kin = k;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int x = __shfl_up_sync(0xffffffffu, kin, o);
    if (lane >= o) kin += x;
  }
  cta_sync<SB>();
  if (lane == 31) smem_i[warp] = kin;
  cta_sync<SB>();
  int koff = 0;
  for (int w = 0; w < warp; ++w) koff += smem_i[w];
  const int c0 = koff + kin - k;  // first CTA of this row
  if (tid < n2) {
    const int q = (((v + k - 1) / k) + 7) & ~7;  // positions per CTA, multiple of 8
    const int count = (v + q - 1) / q;
    for (int jj = 0; jj < count; ++jj)
      desc[(c0 + jj) * 2] = make_int4(tid | (slot_s[tid] << 16), jj * q, min(v, (jj + 1) * q), (jj << 16) | count);
  }
}

}  // namespace t2s
