// kernels.cuh — __global__ entry points: per-phase wrappers (graph mode), the persistent cooperative
// decode kernel, prefill-only kernels (input embedding, prefix-LM attention) and weight packing.
#pragma once
#include "phases.cuh"

namespace t2s {

constexpr size_t SMEM_PROJ = SMEM_PROJ_MAX;
constexpr size_t SMEM_ATTN = sizeof(AttnSmem);
constexpr size_t SMEM_SAMP = sizeof(SampSmem);
constexpr size_t SMEM_MAX = SMEM_PROJ > SMEM_ATTN ? (SMEM_PROJ > SMEM_SAMP ? SMEM_PROJ : SMEM_SAMP)
                                                  : (SMEM_ATTN > SMEM_SAMP ? SMEM_ATTN : SMEM_SAMP);

// ---- one kernel per phase (graph mode and prefill) ---------------------------------------------------
enum { PH_QKV = 0, PH_ATTN, PH_OPROJ, PH_FFN1, PH_FFN2, PH_HEAD, PH_SAMPLE, PH_PLAN };

template <int PH>
__global__ void __launch_bounds__(NT, 1) k_phase(Ctx c, int layer) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int cta = blockIdx.x, ncta = gridDim.x;
  if (PH == PH_QKV) {
    phase_qkv(c, layer, ld_cg_i(c.n_rows), cta, ncta, smem);
  } else if (PH == PH_ATTN) {
    phase_attn_decode(c, layer, ld_cg_i(c.n_rows), cta, ncta, *reinterpret_cast<AttnSmem*>(smem));
  } else if (PH == PH_OPROJ) {
    phase_oproj(c, layer, ld_cg_i(c.n_rows), cta, ncta, smem);
  } else if (PH == PH_FFN1) {
    phase_ffn1(c, layer, ld_cg_i(c.n_rows), cta, ncta, smem);
  } else if (PH == PH_FFN2) {
    phase_ffn2(c, layer, ld_cg_i(c.n_rows), cta, ncta, smem);
  } else if (PH == PH_HEAD) {
    phase_head(c, ld_cg_i(c.n_active), cta, ncta, smem);
  } else if (PH == PH_SAMPLE) {
    phase_sample(c, ld_cg_i(c.n_active), cta, ncta, *reinterpret_cast<SampSmem*>(smem));
  } else if (PH == PH_PLAN) {
    if (cta == 0) phase_plan(c, reinterpret_cast<int*>(smem));
  }
}

// ---- persistent cooperative decode: every remaining step of the request in ONE launch ----------------
// Grid = one CTA per SM (cooperative launch guarantees co-residency).  Phases are separated by the
// GridBarrier; the loop exits when the active list is empty (all CTAs read the same count after a
// barrier) or after max_new_steps.
__global__ void __launch_bounds__(NT, 1) k_decode_persistent(Ctx c, int max_new_steps) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int cta = blockIdx.x, ncta = gridDim.x;
  GridBarrier bar;
  bar.init(c.bar, c.abort_flag, (unsigned)ncta);
  AttnSmem& as = *reinterpret_cast<AttnSmem*>(smem);
  SampSmem& ss = *reinterpret_cast<SampSmem*>(smem);
  __shared__ int4 s_desc[2];  // this CTA's split-KV descriptors: fixed for all 24 layers of a step
  for (int it = 0; it < max_new_steps; ++it) {
    const int n = ld_cg_i(c.n_active);
    if (n == 0 || __ldcg(c.abort_flag) != 0) break;
    if (threadIdx.x < 2) s_desc[threadIdx.x] = __ldcg(reinterpret_cast<const int4*>(c.attn_desc) + cta * 2 + threadIdx.x);
    __syncthreads();
    if (c.timeline) {
      const bool on = (it == c.tl_step);
      bar.tl = on ? c.timeline + (size_t)cta * c.tl_slots * 2 : nullptr;
      bar.tl_k = 0; bar.tl_n = c.tl_slots;
    }
    for (int layer = 0; layer < c.n_layer; ++layer) {
      phase_qkv(c, layer, n, cta, ncta, smem);
      prefetch_phase(c, PF_WO, layer, cta);  // attention has no weights: fetch the O-projection slice early
      bar.sync();
      phase_attn_decode(c, layer, n, cta, ncta, as, s_desc);
      prefetch_phase(c, PF_W1, layer, cta);
      bar.sync();
      phase_oproj(c, layer, n, cta, ncta, smem);
      prefetch_phase(c, PF_W2, layer, cta);
      bar.sync();
      phase_ffn1(c, layer, n, cta, ncta, smem);
      prefetch_phase(c, PF_WQKV_NEXT, layer, cta);  // next layer's QKV slice, or the head's after the last layer
      bar.sync();
      phase_ffn2(c, layer, n, cta, ncta, smem);
      bar.sync();
    }
    phase_head(c, n, cta, ncta, smem);
    bar.sync();
    phase_sample(c, n, cta, ncta, ss);
    if (cta < 3 * D / 16) prefetch_unit(c.wmat + OFF_WQKV, cta, D, 1);  // layer 0 of the next step
    bar.sync();
    if (cta == 0) phase_plan(c, reinterpret_cast<int*>(smem));
    bar.sync();
  }
}

// Grid-barrier latency microbenchmark (measurement hook): n barriers back to back.
__global__ void __launch_bounds__(NT, 1) k_barrier_bench(unsigned* counter, int* abort_flag, int n) {
  GridBarrier bar;
  bar.init(counter, abort_flag, gridDim.x);
  for (int i = 0; i < n; ++i) bar.sync();
}

// ---- session initialisation ----------------------------------------------------------------------------
// One CTA per utterance of the request: local index b = blockIdx.x, session slot = slots[b] (fresh session: b; t2s_admit: free
// slots, released ones first).  seen-bitmap from the prompt
// (previous_tokens = y includes the prompt, t2s_model.py:714), per-slot counters.  fresh = 1: a new session (global counters
// reset, identity active list); fresh = 0: t2s_admit adds the utterances to the resident session at global step `step0`
// (k_admit appends them to the active list after their prefill).
__global__ void k_init_session(Ctx c, const long long* prompt, long long prompt_row_stride, const int* s0, const int* slots,
                               const int* uids, int step0, int fresh) {
  const int b = blockIdx.x, tid = threadIdx.x, slot = slots[b];
  __shared__ unsigned bits[SEEN_WORDS];
  if (tid < SEEN_WORDS) bits[tid] = 0u;
  __syncthreads();
  for (int i = tid; i < c.P; i += blockDim.x) {
    const long long t = prompt[(long long)b * prompt_row_stride + i];
    if (t >= 0 && t < V) atomicOr(&bits[t >> 5], 1u << (t & 31));
    else atomicExch(c.abort_flag, ABORT_BAD_ID);  // nn.Embedding(1025) raises IndexError here (t2s_model.py:636)
  }
  __syncthreads();
  if (tid < SEEN_WORDS) c.seen[(size_t)slot * SEEN_WORDS + tid] = bits[tid];
  if (tid == 0) {
    c.done[slot] = 0;
    c.out_idx[slot] = -1;
    c.seq_len[slot] = s0[b];
    c.slot_step0[slot] = step0;
    c.slot_P[slot] = c.P;
    c.slot_uid[slot] = uids[b];
    const_cast<const long long**>(c.slot_prompt)[slot] = prompt + (long long)b * prompt_row_stride;
    c.seg_cnt[slot] = 0;
    if (fresh) {
      c.active[slot] = slot;
      if (b == 0) {
        *c.n_active = gridDim.x;
        *c.step = 0;
        c.stats[0] = c.stats[1] = c.stats[2] = 0ull;
      }
    }
  }
}

// t2s_admit, after the new utterances' prefill and step-0 sample: append those that did not stop at once to the active list
// and give them the row descriptors phase_plan would have produced (their first decode step is the session's next step).
// One CTA; n_new <= MAX_B.
__global__ void k_admit(Ctx c, const int* slots, int n_new) {
  __shared__ int keep_s[MAX_B];
  __shared__ int base_s;
  const int tid = threadIdx.x;
  for (int i = tid; i < n_new; i += blockDim.x) keep_s[i] = ld_cg_i(c.done + slots[i]) ? 0 : 1;
  __syncthreads();
  if (tid == 0) {
    const int n_old = ld_cg_i(c.n_active);
    int p = n_old;
    unsigned long long kv = 0;
    for (int i = 0; i < n_new; ++i) {
      if (!keep_s[i]) { keep_s[i] = -1; continue; }
      keep_s[i] = p++;
    }
    base_s = p - n_old;
    *c.n_active = p;
    *c.n_rows = p;
    (void)kv;
    if (p > n_old) {
      if (n_old == 0) atomicAdd(c.stats + 1, 1ull);  // the previous plan had scheduled no step: this one is scheduled now
      atomicAdd(c.stats + 2, (unsigned long long)(p - n_old));
    }
  }
  __syncthreads();
  for (int i = tid; i < n_new; i += blockDim.x) {
    const int p = keep_s[i];
    if (p < 0) continue;
    const int slot = slots[i], pos = ld_cg_i(c.seq_len + slot);
    c.active[p] = slot;
    c.row_slot[p] = slot;
    c.row_pos[p] = pos;
    c.row_kvoff[p] = kv_row_off(c.page_table[slot * c.max_pages + (pos >> PAGE_SHIFT)], pos & (PAGE - 1));
    c.seq_len[slot] = pos + 1;
    atomicAdd(c.stats + 0, (unsigned long long)(pos + 1));
  }
}

// ---- prefill input embedding (t2s_model.py:611-622, 636-641; embedding.py:74-78) ----------------------
// Row r = (slot, j).  Text rows:  emb_text[ph] + bert_proj.bias + alpha_t*pe[j]   (+ bert_proj GEMM, added
// afterwards by the OUT_BERT projection phase);  audio rows: emb_audio[tok] + alpha_a*pe[j - L].
__global__ void k_embed_rows(Ctx c, int n_rows, const long long* phoneme_ids, const int* text_off,
                             const int* text_len, const long long* prompt, long long prompt_row_stride, const int* slot_local) {
  const int r = blockIdx.x;
  if (r >= n_rows) return;
  const int slot = slot_local[c.row_slot[r]], j = c.row_pos[r], L = text_len[slot];  // request-local utterance index
  float* out = c.x0 + (size_t)r * D;
  if (j < L) {
    long long ph = phoneme_ids[text_off[slot] + j];
    if (ph < 0 || ph >= c.phoneme_vocab) {  // e.g. v2 symbol ids fed to a 512-symbol v1 checkpoint: the reference raises IndexError
      if (threadIdx.x == 0) atomicExch(c.abort_flag, ABORT_BAD_ID);
      ph = 0;
    }
    const bf16* er = c.emb_text + (size_t)ph * D;
    const float* pr = c.pe + (size_t)j * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      const float v = __bfloat162float(er[d]) + c.bbert[d] + c.alpha_text * pr[d];
      out[d] = v;
      c.x0b[(size_t)r * D + d] = __float2bfloat16_rn(v);  // text rows: overwritten with the final value by k_bert_proj
    }
  } else {
    long long tok = prompt[(long long)slot * prompt_row_stride + (j - L)];
    if (tok < 0 || tok >= V) {
      if (threadIdx.x == 0) atomicExch(c.abort_flag, ABORT_BAD_ID);
      tok = 0;
    }
    const bf16* er = c.emb_audio + (size_t)tok * D;
    const float* pr = c.pe + (size_t)(j - L) * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      const float v = __bfloat162float(er[d]) + c.alpha_audio * pr[d];
      out[d] = v;
      c.x0b[(size_t)r * D + d] = __float2bfloat16_rn(v);
    }
  }
}

// BERT features arrive feature-major per utterance ([1024, L_i], TTS.py:1215); gather them token-major
// in bf16 for the bert_proj GEMM.  One CTA per text row.
template <typename T>
__global__ void k_bert_rows(bf16* out, const void* const* bert, const long long* stride_c,
                            const long long* stride_t, const int* trow_slot, const int* trow_j) {
  const int r = blockIdx.x;
  const int slot = trow_slot[r], j = trow_j[r];
  const T* src = reinterpret_cast<const T*>(bert[slot]);
  const long long sc = stride_c[slot], st = stride_t[slot];
  for (int ch = threadIdx.x; ch < BERT; ch += blockDim.x)
    out[(size_t)r * BERT + ch] = __float2bfloat16_rn((float)src[ch * sc + j * st]);
}

__global__ void __launch_bounds__(NT, 1) k_bert_proj(Ctx c, const bf16* bert_rows, const int* trow_row, int n_text_rows) {
  extern __shared__ __align__(16) unsigned char smem[];
  ProjArgs a{};
  a.w = c.wbert; a.n_tiles = D / 16;
  a.in_b16 = bert_rows; a.out_idx = trow_row;
  proj_phase<OUT_BERT, 8>(c, a, n_text_rows, blockIdx.x, gridDim.x, smem);
}

// ---- prefill attention with the prefix-LM mask (t2s_model.py:644-683 + SDPA :157) ------------------------
// Row j of a sequence with L text positions sees keys [0, L) if j < L (text: bidirectional over text,
// no audio) and [0, j] otherwise (audio: all text + causal audio).  Left padding never materialises.
// CTA = (64-query tile of one sequence, head); one thread per query, K/V tiles staged in shared memory
// as fp32 and read by broadcast.  Reads the bf16 K/V just written to the cache pages, so prefill and
// decode see identical (rounded) keys and values.
struct QTile { int slot, q0, row0, n_q; };  // row0 = global row index of query q0

__global__ void __launch_bounds__(64) k_prefill_attn(Ctx c, int layer, const QTile* tiles, const int* text_len) {
  const QTile qt = tiles[blockIdx.x];
  const int head = blockIdx.y, tid = threadIdx.x;
  const int L = text_len[qt.slot];
  __shared__ float Ks[64][DH];
  __shared__ float Vs[64][DH];
  const bool has_q = tid < qt.n_q;
  const int j = qt.q0 + tid;
  const int nvis = has_q ? ((j < L) ? L : j + 1) : 0;
  const int last = qt.q0 + qt.n_q - 1;
  const int kv_end = (qt.q0 < L) ? max(L, last + 1) : last + 1;
  float q[DH], acc[DH], m = -INFINITY, l = 0.f;
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; acc[d] = 0.f; }
  if (has_q) {
    const float* qr = c.q + (size_t)(qt.row0 + tid) * D + head * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      float4 t = *reinterpret_cast<const float4*>(qr + d);
      q[d] = t.x; q[d + 1] = t.y; q[d + 2] = t.z; q[d + 3] = t.w;
    }
  }
  const bf16* kbase = c.kpool + (size_t)layer * c.kv_layer_stride;
  const bf16* vbase = c.vpool + (size_t)layer * c.kv_layer_stride;
  const int* pt = c.page_table + qt.slot * c.max_pages;
  for (int kt = 0; kt < kv_end; kt += 64) {
    __syncthreads();
    {
      const int p = kt + tid;
      if (p < kv_end) {
        const size_t off = (size_t)kv_row_off(pt[p >> PAGE_SHIFT], p & (PAGE - 1)) + (size_t)head * KV_HEAD_STRIDE;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int pc = (ch ^ ((p >> 1) & 3)) * 8;  // chunks of a head row are swizzled by position (common.cuh kv_feat)
          const uint4 kk = *reinterpret_cast<const uint4*>(kbase + off + pc);
          const uint4 vv = *reinterpret_cast<const uint4*>(vbase + off + pc);
          float* kd = &Ks[tid][ch * 8];
          float* vd = &Vs[tid][ch * 8];
          kd[0] = bf_lo(kk.x); kd[1] = bf_hi(kk.x); kd[2] = bf_lo(kk.y); kd[3] = bf_hi(kk.y);
          kd[4] = bf_lo(kk.z); kd[5] = bf_hi(kk.z); kd[6] = bf_lo(kk.w); kd[7] = bf_hi(kk.w);
          vd[0] = bf_lo(vv.x); vd[1] = bf_hi(vv.x); vd[2] = bf_lo(vv.y); vd[3] = bf_hi(vv.y);
          vd[4] = bf_lo(vv.z); vd[5] = bf_hi(vv.z); vd[6] = bf_lo(vv.w); vd[7] = bf_hi(vv.w);
        }
      }
    }
    __syncthreads();
    const int jmax = min(64, nvis - kt);
    for (int jj = 0; jj < jmax; ++jj) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[jj][d]);
        s += q[d] * kk.x + q[d + 1] * kk.y + q[d + 2] * kk.z + q[d + 3] * kk.w;
      }
      if (s > m) {
        const float corr = exp2f(m - s);
        l *= corr;
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] *= corr;
        m = s;
      }
      const float p = exp2f(s - m);
      l += p;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[jj][d]);
        acc[d] += p * vv.x; acc[d + 1] += p * vv.y; acc[d + 2] += p * vv.z; acc[d + 3] += p * vv.w;
      }
    }
  }
  if (has_q) {
    const float inv = 1.0f / l;
    bf16* o = c.attn + (size_t)(qt.row0 + tid) * D + head * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 2)
      *reinterpret_cast<uint32_t*>(o + d) = pack_bf2(acc[d] * inv, acc[d + 1] * inv);
  }
}

// ---- the same attention on the tensor cores (mma.sync.m16n8k16, flash-attention style) ---------------------------------
// CTA = (64-query tile of one sequence, head), 4 warps x 16 queries.  K/V tiles of 64 keys (4 KB + 4 KB: contiguous in the
// head-major pages) are staged with cp.async, double buffered, keys past the visible range zero-filled.
//   S = Q K^T : A = Q (registers; bf16 hi + lo halves of the fp32 pre-scaled query, so the scores keep fp32-query
//               precision), B = K rows straight from shared memory (chunks XOR-swizzled by position: conflict free)
//   P -> A operand of O += P V directly from the S accumulator registers; B = V through ldmatrix.trans
// Online softmax per query row; the prefix-LM mask is applied on the scores (text rows: keys < L; audio rows: keys <= own).
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(128) k_prefill_attn_tc(Ctx c, int layer, const QTile* tiles, const int* text_len) {
  const QTile qt = tiles[blockIdx.x];
  const int head = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int L = text_len[qt.slot];
  __shared__ __align__(128) unsigned char kv[2][8192];  // [buffer][K: 64 keys x 64 B | V: 64 keys x 64 B]
  const int last = qt.q0 + qt.n_q - 1;
  const int kv_end = (qt.q0 < L) ? max(L, last + 1) : last + 1;
  const int ntile = (kv_end + 63) >> 6;
  const bf16* kbase = c.kpool + (size_t)layer * c.kv_layer_stride + (size_t)head * KV_HEAD_STRIDE;
  const int* pt = c.page_table + qt.slot * c.max_pages;
  auto stage = [&](int kt, int buf) {  // 512 pieces of 16 B: K 0..255, V 256..511
    const int p0 = kt * 64;
    const bf16* src = kbase + (size_t)pt[p0 >> PAGE_SHIFT] * KV_PAGE_STRIDE + (size_t)(p0 & (PAGE - 1)) * DH;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(kv[buf]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pc = tid + 128 * i;             // piece
      const int key = (pc & 255) >> 2;          // 4 pieces per key
      const bf16* sp = src + (pc >= 256 ? KV_V_OFF : 0) + (size_t)(pc & 255) * 8;
      cp_async16_zfill(dst + pc * 16, sp, p0 + key < kv_end);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- this warp's queries: rows i0 = q0 + 16*warp + g and i0 + 8
  const int li0 = warp * 16 + g, li1 = li0 + 8;
  const bool ok0 = li0 < qt.n_q, ok1 = li1 < qt.n_q;
  const int i0 = qt.q0 + li0, i1 = qt.q0 + li1;
  const int nv0 = ok0 ? ((i0 < L) ? L : i0 + 1) : kv_end, nv1 = ok1 ? ((i1 < L) ? L : i1 + 1) : kv_end;
  uint32_t qh[2][4], ql[2][4];  // A fragments: a0 (row g, dims 2t..), a1 (row g+8), a2 (row g, dims 2t+8..), a3 (row g+8)
  {
    const float* q0p = c.q + (size_t)(qt.row0 + li0) * D + head * DH;
    const float* q1p = c.q + (size_t)(qt.row0 + li1) * D + head * DH;
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int d = kb * 16 + h * 8 + 2 * t;
        const float2 a = ok0 ? *reinterpret_cast<const float2*>(q0p + d) : make_float2(0.f, 0.f);
        const float2 b = ok1 ? *reinterpret_cast<const float2*>(q1p + d) : make_float2(0.f, 0.f);
        const float ah0 = __bfloat162float(__float2bfloat16_rn(a.x)), ah1 = __bfloat162float(__float2bfloat16_rn(a.y));
        const float bh0 = __bfloat162float(__float2bfloat16_rn(b.x)), bh1 = __bfloat162float(__float2bfloat16_rn(b.y));
        qh[kb][2 * h] = pack_bf2(ah0, ah1);     ql[kb][2 * h] = pack_bf2(a.x - ah0, a.y - ah1);
        qh[kb][2 * h + 1] = pack_bf2(bh0, bh1); ql[kb][2 * h + 1] = pack_bf2(b.x - bh0, b.y - bh1);
      }
  }
  float o[4][4];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[dt][i] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int mi = lane >> 3, r8 = lane & 7;  // ldmatrix: lane supplies row r8 of 8x8 matrix mi
  int nv_min = min(nv0, nv1);              // keys every row of this warp sees (warp-uniform): tiles below it need no mask
#pragma unroll
  for (int o_ = 16; o_ > 0; o_ >>= 1) nv_min = min(nv_min, __shfl_xor_sync(0xffffffffu, nv_min, o_));
  stage(0, 0);
  for (int kt = 0; kt < ntile; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntile) { stage(kt + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const unsigned char* kt_s = kv[buf];
    const uint32_t vt_s = (uint32_t)__cvta_generic_to_shared(kv[buf]) + 4096;
    // ---- scores: the B fragments of 8 keys x 32 dims (K rows as stored: b0 | b1 of k-block 0, b0 | b1 of k-block 1) come from ONE
    //      ldmatrix.x4 (lane -> key r8 of the group, 16-byte chunk mi of its swizzled 64-byte row)
    float s[8][4];
    const uint32_t kt_a = (uint32_t)__cvta_generic_to_shared(kt_s);
#pragma unroll
    for (int n8 = 0; n8 < 8; ++n8) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[n8][i] = 0.f;
      const int key = n8 * 8 + r8;
      uint32_t b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                   : "r"(kt_a + key * 64 + ((mi ^ ((key >> 1) & 3)) << 4)));
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        mma_bf16_16816(s[n8], make_uint4(qh[kb][0], qh[kb][1], qh[kb][2], qh[kb][3]), b[2 * kb], b[2 * kb + 1]);
        mma_bf16_16816(s[n8], make_uint4(ql[kb][0], ql[kb][1], ql[kb][2], ql[kb][3]), b[2 * kb], b[2 * kb + 1]);
      }
    }
    // ---- mask + online softmax (rows g and g+8; a row is spread over the 4 lanes of a quad).  A tile that every row of the warp
    //      sees completely (all of it in front of the diagonal / inside the text block) needs no mask
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (kt * 64 + 64 <= nv_min) {
#pragma unroll
      for (int n8 = 0; n8 < 8; ++n8) {
        mx0 = fmaxf(mx0, fmaxf(s[n8][0], s[n8][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n8][2], s[n8][3]));
      }
    } else {
#pragma unroll
      for (int n8 = 0; n8 < 8; ++n8) {
        const int j = kt * 64 + n8 * 8 + 2 * t;
        if (j >= nv0) s[n8][0] = -INFINITY;
        if (j + 1 >= nv0) s[n8][1] = -INFINITY;
        if (j >= nv1) s[n8][2] = -INFINITY;
        if (j + 1 >= nv1) s[n8][3] = -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[n8][0], s[n8][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n8][2], s[n8][3]));
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    // a tile can be entirely masked for a row (text rows past the text keys): keep the state untouched then
    const float c0 = (mn0 == -INFINITY) ? 1.f : ex2f(m0 - mn0), c1 = (mn1 == -INFINITY) ? 1.f : ex2f(m1 - mn1);
    const float e0 = (mn0 == -INFINITY) ? 0.f : mn0, e1 = (mn1 == -INFINITY) ? 0.f : mn1;
    m0 = mn0; m1 = mn1;
    float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
    for (int n8 = 0; n8 < 8; ++n8) {
      s[n8][0] = ex2f(s[n8][0] - e0); s[n8][1] = ex2f(s[n8][1] - e0);
      s[n8][2] = ex2f(s[n8][2] - e1); s[n8][3] = ex2f(s[n8][3] - e1);
      ls0 += s[n8][0] + s[n8][1]; ls1 += s[n8][2] + s[n8][3];
    }
    l0 = l0 * c0 + ls0; l1 = l1 * c1 + ls1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
    // ---- O += P V : P straight from the score registers, V through ldmatrix.trans
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const uint4 a = make_uint4(pack_bf2(s[2 * kb][0], s[2 * kb][1]), pack_bf2(s[2 * kb][2], s[2 * kb][3]),
                                 pack_bf2(s[2 * kb + 1][0], s[2 * kb + 1][1]), pack_bf2(s[2 * kb + 1][2], s[2 * kb + 1][3]));
      const int key = kb * 16 + (mi & 1) * 8 + r8;
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {  // dim-tile pairs (0,1), (2,3)
        uint32_t r0, r1, r2, r3;
        const uint32_t addr = vt_s + key * 64 + (((dp * 2 + (mi >> 1)) ^ ((key >> 1) & 3)) << 4);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
        mma_bf16_16816(o[2 * dp], a, r0, r1);
        mma_bf16_16816(o[2 * dp + 1], a, r2, r3);
      }
    }
    __syncthreads();  // everybody is done with kv[buf] before the stage of tile kt + 2 overwrites it
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    if (ok0) *reinterpret_cast<uint32_t*>(c.attn + (size_t)(qt.row0 + li0) * D + head * DH + dt * 8 + 2 * t) = pack_bf2(o[dt][0] * inv0, o[dt][1] * inv0);
    if (ok1) *reinterpret_cast<uint32_t*>(c.attn + (size_t)(qt.row0 + li1) * D + head * DH + dt * 8 + 2 * t) = pack_bf2(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

// ---- the first op after the path: SynthesizerTrn.decode's quantizer.decode(codes) + nearest x2 upsample --------------
// (module/models.py:989-991; core_vq.py:195-197 dequantize = embedding lookup, :286-290 "b n d -> b d n", :359-365 sum over the
// n_q = 1 layers).  out[d][up*t + j] = codebook[codes[t]][d]: a gather along the code axis written channels-first, so a
// 32 x 32 tile goes through shared memory to keep both the codebook reads (along d) and the writes (along t) coalesced.
// grid (ceil(T/32), ceil(dim/32)), block (32, 8).  Codes outside [0, n_codes) (EOS never belongs to the kept tokens) raise *bad.
__global__ void k_codes_to_latent(const long long* __restrict__ codes, int T, const float* __restrict__ cb, int n_codes, int dim,
                                  int up, float* __restrict__ out, int* bad) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32, lane = threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int t = t0 + i;
    float v = 0.f;
    if (t < T) {
      const long long cde = codes[t];
      if (cde < 0 || cde >= n_codes) {
        if (lane == 0) atomicExch(bad, 1);
      } else if (d0 + lane < dim) {
        v = cb[(size_t)cde * dim + d0 + lane];
      }
    }
    tile[i][lane] = v;
  }
  __syncthreads();
  const size_t row = (size_t)T * up;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int d = d0 + i;
    if (d >= dim) break;
    for (int j = lane; j < 32 * up; j += 32) {
      const int tl = j / up;
      if (t0 + tl < T) out[(size_t)d * row + (size_t)t0 * up + j] = tile[tl][i];
    }
  }
}

// ---- results: prompt ++ kept tokens, original batch order (t2s_model.py:733,753,779) -----------------
__global__ void k_finalize(Ctx c, long long* out, long long row_stride, int* idx_out) {
  const int b = blockIdx.x;
  const int idx = c.out_idx[b];
  const int n = idx < 0 ? 0 : idx;
  const int P = c.slot_P[b];
  const long long* prompt = c.slot_prompt[b];
  long long* o = out + (long long)b * row_stride;
  for (int i = threadIdx.x; i < P; i += blockDim.x) o[i] = prompt[i];
  for (long long i = threadIdx.x; P + i < row_stride; i += blockDim.x)
    o[P + i] = (i < n) ? (long long)c.gen[(size_t)b * c.max_steps + i] : -1ll;
  if (threadIdx.x == 0) idx_out[b] = idx;
}

// Large-batch decode (tensor-core projections read their A operand through TMA, which cannot gather): put the
// layer-0 input of every live row, written per slot by the sampler, in row order.
__global__ void k_gather_x0(Ctx c, float* __restrict__ x0_rows, bf16* __restrict__ x0b_rows) {
  const int r = blockIdx.x;
  if (r >= ld_cg_i(c.n_rows)) return;
  const int slot = ld_cg_i(c.row_slot + r);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    x0_rows[(size_t)r * D + d] = c.x0[(size_t)slot * D + d];
    x0b_rows[(size_t)r * D + d] = c.x0b[(size_t)slot * D + d];
  }
}

// ---- LayerNorm folding and the prefill -> decode hand-off ---------------------------------------------------
// Wg[f][k] = W[f][k] * gamma[k] (bf16), c1[f] = sum_k Wg[f][k], c0[f] = bias[f] + sum_k W[f][k] * beta[k].
// gamma == NULL: no LayerNorm in front of this matrix (layer-0 QKV): Wg = W, c1 = 0, c0 = bias.
// One warp per output feature; rows >= N (head padding) are zero.
__global__ void k_fold_ln(const bf16* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ bias, bf16* __restrict__ wg, float* __restrict__ c1,
                          float* __restrict__ c0, int N, int n_pad) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= n_pad) return;
  float s1 = 0.f, s0 = 0.f;
  for (int k = lane; k < D; k += 32) {
    float wv = 0.f, gv = 0.f;
    if (f < N) {
      wv = __bfloat162float(w[(size_t)f * D + k]);
      const bf16 r = __float2bfloat16_rn(gamma ? wv * gamma[k] : wv);
      gv = __bfloat162float(r);
      wg[(size_t)f * D + k] = r;
      if (beta) s0 += wv * beta[k];
    } else {
      wg[(size_t)f * D + k] = __float2bfloat16_rn(0.f);
    }
    s1 += gv;
  }
  s1 = warp_sum(s1); s0 = warp_sum(s0);
  if (lane == 0) {
    c1[f] = gamma ? s1 : 0.f;
    c0[f] = (f < N && bias ? bias[f] : 0.f) + s0;
  }
}

// After a tensor-core prefill the residual sums exist in fp32 only: give the B last rows (the head's input) the
// bf16 copy + partial statistics the decode-side projections expect.  One warp per row.
__global__ void k_rows_stats(const float* __restrict__ y, const int* __restrict__ rows, int n, bf16* __restrict__ yb,
                             float2* __restrict__ sp) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int r = rows ? rows[i] : i;
  const float* src = y + (size_t)r * D + lane * 16;  // lane = 16-feature tile
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float v = src[j];
    s += v; q += v * v;
    yb[(size_t)r * D + lane * 16 + j] = __float2bfloat16_rn(v);
  }
  sp[(size_t)r * 32 + lane] = make_float2(s, q);
}

// ---- weight packing -----------------------------------------------------------------------------------
__device__ __forceinline__ float load_as_float(const void* p, int dtype, size_t i) {
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  if (dtype == 1) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}
// src [N][K] row-major -> dst in m16n8k16 A-fragment order: ((tile*KB + kb)*32 + lane)*8 + j where
// lane = g*4+t holds (row g | g+8, col 2t+{0,1} | +8): j = 0,1:(g,2t) 2,3:(g+8,2t) 4,5:(g,2t+8) 6,7:(g+8,2t+8)
__global__ void k_pack_matrix(bf16* dst, const void* src, int dtype, int N, int K, int n_tiles) {
  const size_t total = (size_t)n_tiles * 16 * K;
  const int KB = K / 16;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int j = o & 7, lane = (o >> 3) & 31;
    const size_t blk = o >> 8;
    const int kb = blk % KB, tile = blk / KB;
    const int g = lane >> 2, t = lane & 3;
    const int row = tile * 16 + g + ((j & 2) ? 8 : 0);
    const int col = kb * 16 + 2 * t + (j & 1) + ((j & 4) ? 8 : 0);
    const float v = (row < N) ? load_as_float(src, dtype, (size_t)row * K + col) : 0.f;
    dst[o] = __float2bfloat16_rn(v);
  }
}
__global__ void k_convert_bf16(bf16* dst, const void* src, int dtype, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(load_as_float(src, dtype, i));
}
__global__ void k_convert_f32(float* dst, const void* src, int dtype, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = load_as_float(src, dtype, i);
}

}  // namespace t2s
