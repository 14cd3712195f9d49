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
// Skinny projections  Y[rows, features] = act(X[rows, K]) * W[features, K]^T
//
// Work unit = (16-feature tile, 512-wide K slice): its weights are one contiguous 16 KB block stored
// in m16n8k16 A-fragment order, so a warp loads a fragment with one coalesced 128-bit load per lane.
// Inside a unit the 8 warps split K (4 k-blocks each), rows are processed in tiles of 32 (4 MMA
// n-tiles), partial sums are reduced across warps in shared memory, and the epilogue is fused.
// Units are spread over all CTAs, so every SM streams a disjoint slice of the weights.
// =====================================================================================================
enum { IN_X0 = 0, IN_LN = 1, IN_BF16 = 2 };
enum { OUT_QKV = 0, OUT_O = 1, OUT_FFN1 = 2, OUT_FFN2 = 3, OUT_HEAD = 4, OUT_BERT = 5 };

struct ProjSmem {
  bf16 xs[RT * XS];
  float red[NW][16][RT + 1];
};

struct ProjArgs {
  const bf16* w;       // packed weights of this matrix
  int n_tiles;         // 16-feature tiles
  int k_slices;        // 512-wide K slices per feature row
  int k_seq;           // slices accumulated sequentially inside one unit (k_slices/k_seq units run in parallel)
  const float* in_f32; // IN_X0 / IN_LN source rows [.,D]
  const bf16* in_b16;  // IN_BF16 source rows
  int in_stride;       // IN_BF16 row stride (elements)
  const int* in_idx;   // optional row gather for the fp32 sources
  const float* ln_g;   // IN_LN
  const float* ln_b;
  // epilogue
  const float* bias;
  const float* res_g;  // OUT_O: LN params of the previous layer's norm2 (residual recompute)
  const float* res_b;
  const float* b2;     // OUT_FFN1: linear2.bias for the y2 initialisation
  const int* out_idx;  // OUT_BERT: text row -> prompt row
  int layer;
};

template <int IN, int OUT>
__device__ __forceinline__ void proj_stage(const Ctx& c, const ProjArgs& a, ProjSmem& sm, int r0, int n_rows,
                                           int nt, int ks, const float (&g)[16], const float (&be)[16]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < RT / NW; ++i) {
    const int rl = warp + NW * i;
    const int r = r0 + rl;
    uint32_t* dst = reinterpret_cast<uint32_t*>(sm.xs + rl * XS);
    if (r >= n_rows) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dst[(lane * 4 + 128 * j) / 2] = 0u;
        dst[(lane * 4 + 128 * j) / 2 + 1] = 0u;
      }
      continue;
    }
    if (IN == IN_BF16) {
      const bf16* src = a.in_b16 + (size_t)r * a.in_stride + ks * 512;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 v = ld_cg16(src + (lane + 32 * j) * 8);
        *reinterpret_cast<uint4*>(sm.xs + rl * XS + (lane + 32 * j) * 8) = v;
      }
    } else {
      const int ri = a.in_idx ? ld_cg_i(a.in_idx + r) : r;
      const float* src = a.in_f32 + (size_t)ri * D;
      float v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 t = ld_cg_f4(src + lane * 4 + 128 * j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
      if (IN == IN_LN) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) s += v[j];
        const float mean = warp_sum(s) * (1.0f / D);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { float d = v[j] - mean; sq += d * d; }
        const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / D) + LN_EPS);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (v[j] - mean) * rstd * g[j] + be[j];
        if (OUT == OUT_QKV && nt == 0 && lane == 0) c.stat2[r] = make_float2(mean, rstd);
        if (OUT == OUT_FFN1 && lane == (nt & 31)) {
          // y2 := LN1(y1) + b2 on this unit's 4-feature slice; FFN2 adds its split-K partials atomically
          const int j = nt >> 5;
          const float4 bb = *reinterpret_cast<const float4*>(a.b2 + 4 * nt);
          float4 o;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            if (jj == j) o = make_float4(v[4 * jj] + bb.x, v[4 * jj + 1] + bb.y, v[4 * jj + 2] + bb.z, v[4 * jj + 3] + bb.w);
          *reinterpret_cast<float4*>(c.y2 + (size_t)r * D + 4 * nt) = o;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dst[(lane * 4 + 128 * j) / 2] = pack_bf2(v[4 * j], v[4 * j + 1]);
        dst[(lane * 4 + 128 * j) / 2 + 1] = pack_bf2(v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
}

template <int OUT>
__device__ __forceinline__ void proj_epilogue(const Ctx& c, const ProjArgs& a, int r, int f, float acc) {
  if (OUT == OUT_QKV) {
    const float val = acc + a.bias[f];
    if (f < D) {
      c.q[(size_t)r * D + f] = val * QSCALE;
    } else {
      const int slot = ld_cg_i(c.row_slot + r), pos = ld_cg_i(c.row_pos + r);
      const int page = c.page_table[slot * c.max_pages + (pos >> 6)];
      const size_t off = (size_t)a.layer * c.kv_layer_stride + ((size_t)page * PAGE + (pos & (PAGE - 1))) * D;
      if (f < 2 * D) c.kpool[off + (f - D)] = __float2bfloat16_rn(val);
      else c.vpool[off + (f - 2 * D)] = __float2bfloat16_rn(val);
    }
  } else if (OUT == OUT_O) {
    float res;
    if (a.layer == 0) {
      const int ri = c.x0_by_slot ? ld_cg_i(c.row_slot + r) : r;
      res = ld_cg_f(c.x0 + (size_t)ri * D + f);
    } else {
      const float2 st = __ldcg(c.stat2 + r);
      res = (ld_cg_f(c.y2 + (size_t)r * D + f) - st.x) * st.y * a.res_g[f] + a.res_b[f];
    }
    c.y1[(size_t)r * D + f] = res + a.bias[f] + acc;
  } else if (OUT == OUT_FFN1) {
    c.h[(size_t)r * FF + f] = __float2bfloat16_rn(fmaxf(acc + a.bias[f], 0.f));
  } else if (OUT == OUT_FFN2) {
    float* p = c.y2 + (size_t)r * D + f;
    if (a.k_seq == a.k_slices) *p = __ldcg(p) + acc;  // single writer: deterministic
    else atomicAdd(p, acc);                           // split-K across CTAs
  } else if (OUT == OUT_HEAD) {
    if (f < V) c.logits[(size_t)r * VPAD + f] = acc;
  } else if (OUT == OUT_BERT) {
    float* p = c.x0 + (size_t)a.out_idx[r] * D + f;
    if (a.k_seq == a.k_slices) *p = __ldcg(p) + acc;
    else atomicAdd(p, acc);
  }
}

template <int IN, int OUT>
__device__ void proj_phase(const Ctx& c, const ProjArgs& a, int n_rows, int cta, int ncta, ProjSmem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_units = a.n_tiles * (a.k_slices / a.k_seq);
  const int kb_per_row = a.k_slices * 32;
  float lg[16], lb[16];
  if (IN == IN_LN) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 x = *reinterpret_cast<const float4*>(a.ln_g + lane * 4 + 128 * j);
      float4 y = *reinterpret_cast<const float4*>(a.ln_b + lane * 4 + 128 * j);
      lg[4 * j] = x.x; lg[4 * j + 1] = x.y; lg[4 * j + 2] = x.z; lg[4 * j + 3] = x.w;
      lb[4 * j] = y.x; lb[4 * j + 1] = y.y; lb[4 * j + 2] = y.z; lb[4 * j + 3] = y.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) { lg[j] = 1.f; lb[j] = 0.f; }
  }
  for (int u = cta; u < n_units; u += ncta) {
    const int nt = u % a.n_tiles, kp = u / a.n_tiles;
    for (int r0 = 0; r0 < n_rows; r0 += RT) {
      const int rows_here = min(RT, n_rows - r0);
      const int n8 = (rows_here + 7) >> 3;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int kq = 0; kq < a.k_seq; ++kq) {
        const int ks = kp * a.k_seq + kq;
        // this warp's 4 A fragments (k-blocks ks*32 + warp*4 .. +3 of feature tile nt)
        uint4 af[4];
        const uint4* wp = reinterpret_cast<const uint4*>(a.w) + ((size_t)nt * kb_per_row + ks * 32 + warp * 4) * 32 + lane;
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = ld_weight16(wp + i * 32);
        __syncthreads();  // previous readers of xs / red are done
        proj_stage<IN, OUT>(c, a, sm, r0, n_rows, nt, ks, lg, lb);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k0 = (warp * 4 + i) * 16;
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            if (n < n8) {
              const uint32_t* xr = reinterpret_cast<const uint32_t*>(sm.xs + (n * 8 + g) * XS + k0);
              mma_bf16_16816(acc[n], af[i], xr[t], xr[4 + t]);
            }
          }
        }
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        sm.red[warp][g][n * 8 + 2 * t] = acc[n][0];
        sm.red[warp][g][n * 8 + 2 * t + 1] = acc[n][1];
        sm.red[warp][g + 8][n * 8 + 2 * t] = acc[n][2];
        sm.red[warp][g + 8][n * 8 + 2 * t + 1] = acc[n][3];
      }
      __syncthreads();
#pragma unroll
      for (int o = threadIdx.x; o < 16 * RT; o += NT) {
        const int fl = o & 15, n = o >> 4;
        if (n < rows_here) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < NW; ++w) s += sm.red[w][fl][n];
          proj_epilogue<OUT>(c, a, r0 + n, nt * 16 + fl, s);
        }
      }
    }
  }
}

// ---- the five per-layer projection phases + head, with their argument wiring ------------------------
__device__ __forceinline__ void phase_qkv(const Ctx& c, int layer, int n_rows, int cta, int ncta, ProjSmem& sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_WQKV;
  a.n_tiles = 3 * D / 16; a.k_slices = 1; a.k_seq = 1; a.layer = layer;
  a.bias = c.wvec + (size_t)layer * LV + VO_BQKV;
  if (layer == 0) {
    a.in_f32 = c.x0; a.in_idx = c.x0_by_slot ? c.row_slot : nullptr;
    proj_phase<IN_X0, OUT_QKV>(c, a, n_rows, cta, ncta, sm);
  } else {
    a.in_f32 = c.y2;
    a.ln_g = c.wvec + (size_t)(layer - 1) * LV + VO_G2;
    a.ln_b = c.wvec + (size_t)(layer - 1) * LV + VO_BE2;
    proj_phase<IN_LN, OUT_QKV>(c, a, n_rows, cta, ncta, sm);
  }
}
__device__ __forceinline__ void phase_oproj(const Ctx& c, int layer, int n_rows, int cta, int ncta, ProjSmem& sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_WO;
  a.n_tiles = D / 16; a.k_slices = 1; a.k_seq = 1; a.layer = layer;
  a.in_b16 = c.attn; a.in_stride = D;
  a.bias = c.wvec + (size_t)layer * LV + VO_BO;
  if (layer > 0) {
    a.res_g = c.wvec + (size_t)(layer - 1) * LV + VO_G2;
    a.res_b = c.wvec + (size_t)(layer - 1) * LV + VO_BE2;
  }
  proj_phase<IN_BF16, OUT_O>(c, a, n_rows, cta, ncta, sm);
}
__device__ __forceinline__ void phase_ffn1(const Ctx& c, int layer, int n_rows, int cta, int ncta, ProjSmem& sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_W1;
  a.n_tiles = FF / 16; a.k_slices = 1; a.k_seq = 1; a.layer = layer;
  a.in_f32 = c.y1;
  a.ln_g = c.wvec + (size_t)layer * LV + VO_G1;
  a.ln_b = c.wvec + (size_t)layer * LV + VO_BE1;
  a.bias = c.wvec + (size_t)layer * LV + VO_B1;
  a.b2 = c.wvec + (size_t)layer * LV + VO_B2;
  proj_phase<IN_LN, OUT_FFN1>(c, a, n_rows, cta, ncta, sm);
}
__device__ __forceinline__ void phase_ffn2(const Ctx& c, int layer, int n_rows, int cta, int ncta, ProjSmem& sm) {
  ProjArgs a{};
  a.w = c.wmat + (size_t)layer * LW + OFF_W2;
  a.n_tiles = D / 16; a.k_slices = FF / 512; a.k_seq = c.deterministic ? FF / 512 : 1; a.layer = layer;
  a.in_b16 = c.h; a.in_stride = FF;
  proj_phase<IN_BF16, OUT_FFN2>(c, a, n_rows, cta, ncta, sm);
}
__device__ __forceinline__ void phase_head(const Ctx& c, int n_rows, int cta, int ncta, ProjSmem& sm) {
  ProjArgs a{};
  a.w = c.whead;
  a.n_tiles = VT; a.k_slices = 1; a.k_seq = 1; a.layer = c.n_layer;
  a.in_f32 = c.y2; a.in_idx = c.head_rows;
  a.ln_g = c.wvec + (size_t)(c.n_layer - 1) * LV + VO_G2;
  a.ln_b = c.wvec + (size_t)(c.n_layer - 1) * LV + VO_BE2;
  proj_phase<IN_LN, OUT_HEAD>(c, a, n_rows, cta, ncta, sm);
}

// =====================================================================================================
// Decode attention: one query per active sequence against its paged bf16 KV cache, all 16 heads at
// once.  The flattened list of (sequence, CH-position chunk) items is split evenly over the CTAs
// (split-KV); a CTA keeps an online-softmax state in registers while it stays on one sequence, the
// partial states of a sequence are merged by the last CTA to finish it.
//
// Lane mapping: a K (or V) row is 1 KB = 64 x 16 B.  Lane l loads chunk l (head l/4, dims 8*(l%4)..+7)
// and chunk 32+l (head 8+l/4): two fully coalesced 512 B requests per row per warp.
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

template <int U>
__device__ __forceinline__ void attn_segment(const Ctx& c, int layer, int slot, int pbeg, int pend,
                                             const float (&qa)[8], const float (&qb)[8],
                                             float (&m)[2], float (&l)[2], float (&accA)[8], float (&accB)[8]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* kbase = c.kpool + (size_t)layer * c.kv_layer_stride;
  const bf16* vbase = c.vpool + (size_t)layer * c.kv_layer_stride;
  const int* pt = c.page_table + slot * c.max_pages;
  for (int p0 = pbeg + U * warp; p0 < pend; p0 += U * NW) {
    const int page = pt[p0 >> 6];
    const size_t rowoff = ((size_t)page * PAGE + (p0 & (PAGE - 1))) * D;
    uint4 ka[U], kb[U], va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p0 + u < pend) {
        const bf16* kr = kbase + rowoff + (size_t)u * D;
        const bf16* vr = vbase + rowoff + (size_t)u * D;
        ka[u] = ld_cg16(kr + lane * 8);
        kb[u] = ld_cg16(kr + 256 + lane * 8);
        va[u] = ld_cg16(vr + lane * 8);
        vb[u] = ld_cg16(vr + 256 + lane * 8);
      } else {
        ka[u] = kb[u] = va[u] = vb[u] = make_uint4(0, 0, 0, 0);
      }
    }
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
}

__device__ void phase_attn_decode(const Ctx& c, int layer, int n_rows, int cta, int ncta, AttnSmem& sm) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- 1. chunk size from the total number of cached positions
  int np = (tid < n_rows) ? ld_cg_i(c.row_pos + tid) + 1 : 0;
  int tot = np;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (lane == 0) sm.scratch[warp] = tot;
  __syncthreads();
  int total_pos = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) total_pos += sm.scratch[w];
  int CH = PAGE;
  while (CH > 8 && (total_pos + CH - 1) / CH < ncta) CH >>= 1;
  // ---- 2. exclusive scan of per-row item counts
  const int items = (np + CH - 1) / CH;
  int inc = items;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();  // scratch reuse
  if (lane == 31) sm.scratch[warp] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += sm.scratch[w];
  sm.pre[tid + 1] = woff + inc;  // tid < 256 = MAX_B
  if (tid == 0) sm.pre[0] = 0;
  __syncthreads();
  const int T = sm.pre[n_rows];
  const int per = (T + ncta - 1) / ncta;
  const int i0 = cta * per, i1 = min(T, i0 + per);
  if (i0 >= i1) return;
  // ---- 3. first row whose items intersect [i0, i1)
  int lo = 0, hi = n_rows - 1;
  while (lo < hi) {  // largest r with pre[r] <= i0
    const int mid = (lo + hi + 1) >> 1;
    if (sm.pre[mid] <= i0) lo = mid; else hi = mid - 1;
  }
  for (int r = lo; r < n_rows && sm.pre[r] < i1; ++r) {
    const int a0 = max(i0, sm.pre[r]), a1 = min(i1, sm.pre[r + 1]);
    if (a0 >= a1) continue;  // rows with zero items cannot occur (np >= 1), kept for safety
    const int npos = ld_cg_i(c.row_pos + r) + 1;
    const int pbeg = (a0 - sm.pre[r]) * CH, pend = min(npos, (a1 - sm.pre[r]) * CH);
    const int slot = ld_cg_i(c.row_slot + r);
    const int first_cta = sm.pre[r] / per, last_cta = (sm.pre[r + 1] - 1) / per;
    const int count = last_cta - first_cta + 1;
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
    if (pend - pbeg > 32) attn_segment<4>(c, layer, slot, pbeg, pend, qa, qb, m, l, accA, accB);
    else attn_segment<1>(c, layer, slot, pbeg, pend, qa, qb, m, l, accA, accB);
    __syncthreads();  // previous segment's merge readers are done with sm.m/l/acc
    if ((lane & 3) == 0) {
      sm.m[warp][lane >> 2] = m[0]; sm.l[warp][lane >> 2] = l[0];
      sm.m[warp][8 + (lane >> 2)] = m[1]; sm.l[warp][8 + (lane >> 2)] = l[1];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { sm.acc[warp][lane * 8 + i] = accA[i]; sm.acc[warp][256 + lane * 8 + i] = accB[i]; }
    __syncthreads();
    attn_merge_write(c, sm, r, r + cta, count);
    if (count > 1) {
      __threadfence();
      __syncthreads();
      if (tid == 0) sm.scratch[NW] = (atomicAdd(c.seg_cnt + r, 1) == count - 1);
      __syncthreads();
      if (sm.scratch[NW]) {  // last CTA of this sequence: merge all partials
        __threadfence();
        const float* pbase = c.part + (size_t)(r + first_cta) * PART_STRIDE;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int d = tid + 256 * k, hd = d >> 5;
          float mm = -INFINITY;
          for (int s = 0; s < count; ++s) mm = fmaxf(mm, ld_cg_f(pbase + (size_t)s * PART_STRIDE + hd));
          float ll = 0.f, aa = 0.f;
          for (int s = 0; s < count; ++s) {
            const float* ps = pbase + (size_t)s * PART_STRIDE;
            const float sc = exp2f(ld_cg_f(ps + hd) - mm);
            ll += ld_cg_f(ps + NH + hd) * sc;
            aa += ld_cg_f(ps + 2 * NH + d) * sc;
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
__device__ __forceinline__ int block_argmax(float v, int idx, SampSmem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if (lane == 0) { sm.redf[warp] = v; sm.redi[warp] = idx; }
  __syncthreads();
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
__device__ __forceinline__ float block_max(float v, SampSmem& sm) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.redf[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sm.redf[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) r = fmaxf(r, sm.redf[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, SampSmem& sm) {  // fixed order => deterministic
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.redf[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) r += sm.redf[w];
  return r;
}

// k-th largest key among the thread-distributed values (8-bit radix select, 4 passes)
__device__ __forceinline__ uint32_t block_kth_largest(const uint32_t (&key)[SV], const bool (&valid)[SV], int k,
                                                      SampSmem& sm) {
  uint32_t prefix = 0, mask = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    __syncthreads();
    sm.hist[tid] = 0;  // NT == 256 bins
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SV; ++j)
      if (valid[j] && (key[j] & mask) == prefix) atomicAdd(&sm.hist[(key[j] >> shift) & 0xFFu], 1u);
    __syncthreads();
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
          if (run < (unsigned)k && run + cnt[j] >= (unsigned)k) { sm.bcast[0] = 255 - 8 * lane - j; sm.bcast[1] = k - run; }
          run += cnt[j];
        }
      }
    }
    __syncthreads();
    prefix |= (uint32_t)sm.bcast[0] << shift;
    mask |= 0xFFu << shift;
    k = sm.bcast[1];
  }
  return prefix;
}

// top-p (rare path): full descending bitonic sort of (key, index), softmax over the sorted values,
// inclusive prefix sum, remove where cum > top_p except the first (utils.py:169-179).
__device__ void block_top_p(float (&x)[SV], const bool (&valid)[SV], int width, float top_p, SampSmem& sm) {
  const int tid = threadIdx.x;
  for (int i = tid; i < 2048; i += NT) sm.sortbuf[i] = 0ull;  // key 0 sorts last
  __syncthreads();
#pragma unroll
  for (int j = 0; j < SV; ++j) {
    const int i = tid + NT * j;
    // ties: smaller index first in descending order (stable sort of -x) => store ~index in the low word
    if (valid[j]) sm.sortbuf[i] = ((unsigned long long)float_key(x[j]) << 32) | (uint32_t)(0xFFFFFFFFu - i);
  }
  __syncthreads();
  for (int size = 2; size <= 2048; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < 1024; t += NT) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = sm.sortbuf[lo], b = sm.sortbuf[hi];
        if ((a < b) == desc) { sm.sortbuf[lo] = b; sm.sortbuf[hi] = a; }
      }
      __syncthreads();
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
  const float total = block_sum(loc, sm);
  loc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { e[j] = e[j] / total; loc += e[j]; }
  __syncthreads();
  sm.cum[tid] = loc;
  __syncthreads();
  float run = 0.f;
  for (int i = 0; i < tid; ++i) run += sm.cum[i];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = 8 * tid + j;
    run += e[j];
    // cum_probs > top_p is removed, position 0 always kept (utils.py:172-173); tag by clearing the key
    if ((i < width) && (i > 0) && (run > top_p)) sm.sortbuf[i] &= 0xFFFFFFFFull;
  }
  __syncthreads();
  // each thread looks for its own elements: build a removal bitmap in cum[] reinterpret (1025 bits)
  unsigned* bits = reinterpret_cast<unsigned*>(sm.cum);
  for (int i = tid; i < 64; i += NT) bits[i] = 0u;
  __syncthreads();
  for (int i = tid; i < width; i += NT) {
    const unsigned long long v = sm.sortbuf[i];
    if ((v >> 32) == 0) {
      const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(v & 0xFFFFFFFFull);
      atomicOr(&bits[idx >> 5], 1u << (idx & 31));
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < SV; ++j) {
    const int i = tid + NT * j;
    if (valid[j] && ((bits[i >> 5] >> (i & 31)) & 1u)) x[j] = -INFINITY;
  }
  __syncthreads();
}

__device__ void phase_sample(const Ctx& c, int n_active, int cta, int ncta, SampSmem& sm) {
  const int tid = threadIdx.x;
  const int step = ld_cg_i(c.step);
  const int width = (step < c.eos_window) ? (V - 1) : V;
  for (int r = cta; r < n_active; r += ncta) {
    const int slot = ld_cg_i(c.active + r);
    float v[SV];
    bool valid[SV];
#pragma unroll
    for (int j = 0; j < SV; ++j) {
      const int i = tid + NT * j;
      valid[j] = i < width;
      v[j] = valid[j] ? ld_cg_f(c.logits + (size_t)r * VPAD + i) : -INFINITY;
      if (c.logits_rec && step < c.n_logits_rec && i < V)
        c.logits_rec[((size_t)step * c.B0 + slot) * V + i] = ld_cg_f(c.logits + (size_t)r * VPAD + i);
    }
    // repetition penalty over every distinct previous token (prompt + generated), utils.py:159-167
    if (c.rep_pen != 1.0f) {
      const uint32_t* seen = c.seen + (size_t)slot * SEEN_WORDS;
#pragma unroll
      for (int j = 0; j < SV; ++j) {
        const int i = tid + NT * j;
        if (valid[j] && ((__ldcg(seen + (i >> 5)) >> (i & 31)) & 1u)) v[j] = (v[j] < 0.f) ? v[j] * c.rep_pen : v[j] / c.rep_pen;
      }
    }
    // argmax of the penalised logits: the reference's EOS test (t2s_model.py:721/:901)
    int greedy;
    {
      float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < SV; ++j) if (valid[j] && (v[j] > bv)) { bv = v[j]; bi = tid + NT * j; }
      greedy = block_argmax(bv, bi, sm);
    }
    float x[SV];
#pragma unroll
    for (int j = 0; j < SV; ++j) x[j] = v[j];
    if (c.top_p < 1.0f) block_top_p(x, valid, width, c.top_p, sm);
    const float temp = fmaxf(c.temperature, 1e-5f);
#pragma unroll
    for (int j = 0; j < SV; ++j) x[j] = x[j] / temp;
    {
      const int k = min(c.top_k, width);
      uint32_t key[SV];
#pragma unroll
      for (int j = 0; j < SV; ++j) key[j] = float_key(x[j]);
      const uint32_t pivot = block_kth_largest(key, valid, k, sm);
#pragma unroll
      for (int j = 0; j < SV; ++j) if (valid[j] && key[j] < pivot) x[j] = -INFINITY;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SV; ++j) if (valid[j]) mx = fmaxf(mx, x[j]);
    mx = block_max(mx, sm);
    float e[SV], loc = 0.f;
#pragma unroll
    for (int j = 0; j < SV; ++j) { e[j] = valid[j] ? expf(x[j] - mx) : 0.f; loc += e[j]; }
    const float total = block_sum(loc, sm);
    int tok;
    {
      float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < SV; ++j) {
        const int i = tid + NT * j;
        if (!valid[j]) continue;
        uint32_t w[4];
        philox4x32_10((uint32_t)(i >> 2), (uint32_t)step, (uint32_t)slot, 0u, c.seed_lo, c.seed_hi, w);
        const uint32_t word = w[i & 3];
        const float u = ((float)(word >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23
        const float q = -logf(u);
        const float sc = (e[j] / total) / q;
        if (sc > bv) { bv = sc; bi = i; }
      }
      tok = block_argmax(bv, bi, sm);
    }
    // ---- bookkeeping (t2s_model.py:718-769)
    int emit = tok;
    if (c.forced && step < c.n_forced) emit = __ldcg(c.forced + (size_t)slot * c.n_forced + step);
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
      const float* pr = c.pe + (size_t)min(c.P + step, c.pe_len - 1) * D;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int d = tid + NT * k;
        c.x0[(size_t)slot * D + d] = __bfloat162float(er[d]) + c.alpha_audio * pr[d];
      }
    }
    __syncthreads();
  }
}

// Retirement: compact the active list on device (no host round trip), publish next step's rows.
// Runs on one CTA.  (t2s_model.py:724-745 does this with .tolist() + 48 index_selects.)
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
  __syncthreads();
  int woff = 0, total = 0;
  for (int w = 0; w < NW; ++w) { if (w < warp) woff += smem_i[w]; total += smem_i[w]; }
  unsigned long long kvpos = 0;
  if (keep) {
    const int p = woff + inc - 1;
    const int pos = ld_cg_i(c.seq_len + slot);
    c.active[p] = slot;
    c.row_slot[p] = slot;
    c.row_pos[p] = pos;
    c.seq_len[slot] = pos + 1;
    kvpos = (unsigned long long)(pos + 1);
  }
  if (kvpos) atomicAdd(c.stats + 0, kvpos);
  if (tid == 0) {
    *c.n_active = total;
    *c.n_rows = total;
    *c.step = ld_cg_i(c.step) + 1;
    if (total > 0) { atomicAdd(c.stats + 1, 1ull); atomicAdd(c.stats + 2, (unsigned long long)total); }
  }
}

}  // namespace t2s
