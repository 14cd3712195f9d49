// common.cuh — constants, the device context, and small device helpers shared by every kernel.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace t2s {

// ---- architecture constants of the s1 family (GPT_SoVITS/configs/s1longer-v2.yaml:19-31) -------
constexpr int D = 512;          // hidden_dim
constexpr int NH = 16;          // heads
constexpr int DH = 32;          // head_dim
constexpr int FF = 2048;        // 4*hidden (t2s_model.py:304)
constexpr int V = 1025;         // vocab incl. EOS
constexpr int VT = 65;          // 16-row feature tiles of the head (1040 rows, zero padded)
constexpr int VPAD = VT * 16;
constexpr int BERT = 1024;
constexpr int PAGE = 128;       // KV positions per page
constexpr int PAGE_SHIFT = 7;
constexpr int MAX_B = 256;      // max utterances per call
constexpr int NT = 256;         // threads per CTA in every phase kernel
constexpr int NW = NT / 32;
constexpr int RT = 32;          // rows per projection tile (4 n-tiles of the m16n8k16 MMA)
constexpr int XS = D + 8;       // bf16 row stride of the staged activation tile (bank-conflict free)
constexpr int SEEN_WORDS = 33;  // ceil(1025/32)
constexpr float LN_EPS = 1e-5f; // transformer.py:194
// q is pre-scaled by log2(e)/sqrt(head_dim) so softmax uses exp2
constexpr float QSCALE = 1.4426950408889634f * 0.17677669529663687f;

// packed per-layer matrix arena (bf16 elements) and vector arena (fp32 elements)
constexpr size_t OFF_WQKV = 0;
constexpr size_t OFF_WO = OFF_WQKV + (size_t)3 * D * D;
constexpr size_t OFF_W1 = OFF_WO + (size_t)D * D;
constexpr size_t OFF_W2 = OFF_W1 + (size_t)FF * D;
constexpr size_t LW = OFF_W2 + (size_t)D * FF;  // 3,145,728
constexpr int VO_BQKV = 0, VO_BO = 1536, VO_B1 = 2048, VO_B2 = 4096, VO_G1 = 4608, VO_BE1 = 5120,
              VO_G2 = 5632, VO_BE2 = 6144,
              // LayerNorm folded into the consumer GEMMs (k_fold_ln): c1 = Wg 1, c0 = W beta + bias
              VO_C1_QKV = 6656, VO_C0_QKV = 8192, VO_C1_FFN1 = 9728, VO_C0_FFN1 = 11776, LV = 13824;

// KV cache layout inside one layer's pool: [page][head][K: 128 positions x 32 dims | V: 128 x 32] bf16, i.e. everything one
// head needs of one 128-position page is ONE contiguous 16 KB block (K 8 KB, then V 8 KB): a single TMA bulk copy fetches
// it.  K and V share the pool: vpool = kpool + KV_V_OFF.  A row's `kvoff` is the element offset of (its page, head 0, its
// position, dim 0); feature f = head*32 + dim lives kv_feat(f) elements further.
constexpr int KV_HEAD_STRIDE = 2 * PAGE * DH;       // elements between heads inside a page (K block + V block)
constexpr int KV_PAGE_STRIDE = NH * KV_HEAD_STRIDE; // elements per page (all heads, K and V)
constexpr int KV_V_OFF = PAGE * DH;                 // V block behind the K block
__host__ __device__ constexpr long long kv_row_off(long long page, int pos_in_page) { return page * KV_PAGE_STRIDE + (long long)pos_in_page * DH; }
// A position's 64-byte head row is stored with its four 16-byte chunks XOR-swizzled by ((position >> 1) & 3): eight
// consecutive positions x one logical chunk then fall into eight distinct shared-memory bank groups, so ldmatrix reads of a
// page that a bulk copy dropped into shared memory verbatim are conflict free (cluster_decode.cuh).
// kv_feat(kvoff, f): element offset of feature f = head*32 + dim relative to the row's kvoff (which encodes the position).
__host__ __device__ constexpr int kv_swz_of(long long kvoff) { return (int)((kvoff >> 6) & 3); }  // (pos_in_page >> 1) & 3
__host__ __device__ constexpr int kv_feat(long long kvoff, int f) {
  return (f >> 5) * KV_HEAD_STRIDE + (((((f >> 3) & 3) ^ kv_swz_of(kvoff))) << 3) + (f & 7);
}

// abort_flag values: a lost CTA (grid-barrier watchdog) | an input / forced token id outside its embedding table
constexpr int ABORT_WATCHDOG = 1, ABORT_BAD_ID = 2;

constexpr int PART_STRIDE = 2 * NH + D;  // per attention partial: m[16], l[16], acc[512]

typedef __nv_bfloat16 bf16;

// Everything a kernel needs, passed by value (fits the 4 KB parameter space).
struct Ctx {
  // model
  const bf16* wmat;       // [n_layer][LW], each matrix in MMA-fragment tile order (see pack.cuh)
  const float* wvec;      // [n_layer][LV]
  const bf16* whead;      // ar_predict_layer (gamma-folded), packed, VT tiles
  const float* head_c1;   // [VPAD]
  const float* head_c0;
  const bf16* wbert;      // bert_proj.weight packed (32 tiles x 64 k-blocks)
  const float* bbert;     // bert_proj.bias
  const bf16* emb_audio;  // [V][D] row-major bf16
  const bf16* emb_text;   // [phoneme_vocab][D]
  const float* pe;        // [pe_len][D]
  float alpha_audio, alpha_text;
  int n_layer;
  int pe_len;
  int phoneme_vocab;  // rows of emb_text: ids are validated on device (abort_flag = ABORT_BAD_ID), like nn.Embedding's index check
  // KV cache: one pool [n_layer][n_pages][NH][K|V][PAGE][DH] bf16 (see kv_row_off; vpool = kpool + KV_V_OFF), page table [B][max_pages]
  bf16* kpool;
  bf16* vpool;
  size_t kv_layer_stride;
  const int* page_table;
  int max_pages;
  // rows of the current pass (prefill: every prompt position; decode: one per active sequence)
  int* n_rows;
  int* row_slot;
  int* row_pos;
  long long* row_kvoff;   // element offset of the row's KV position inside one layer's pool
  const int* head_rows;  // prefill: last row of each slot; NULL in decode (identity)
  int x0_by_slot;        // layer-0 input indexed by slot (decode) or by row (prefill)
  // activations
  float* x0;
  bf16* x0b;     // bf16 copy of x0 (GEMM operand)
  float* q;
  bf16* attn;
  float* y1;     // residual sums (pre-LayerNorm), fp32 ...
  bf16* yb1;     // ... their bf16 copies (next GEMM's operand) ...
  float2* sp1;   // ... and per-16-feature-tile partial (sum, sum of squares) [rows][32]
  bf16* h;
  float* y2;
  bf16* yb2;
  float2* sp2;
  float2* stat2;
  float* logits;  // [MAX_B][VPAD]
  // decode attention split-KV scratch
  float* part;
  int* seg_cnt;
  int* attn_desc;   // [attn_ctas][2] int4 {row, pbeg, pend, (j << 16) | count}, written by phase_plan
  int attn_ctas;    // grid size of the decode attention phase (fixed for the session)
  // session (indexed by slot = original batch index)
  int B0, P, max_steps, eos_window, early_stop, top_k;  // B0 = slot capacity of the session (stride of the per-step hooks); P: this request's prompt length
  int slot_base;  // first utterance of this session inside the caller's batch (t2s_generate splits large batches): keeps the Philox streams per utterance
  float top_p, temperature, rep_pen;
  uint32_t seed_lo, seed_hi;
  int* step;
  int* seq_len;
  int* active;
  int* n_active;
  int* done;
  int* out_idx;
  // per-slot session state (continuous batching: t2s_admit adds utterances to a resident session at a later global step)
  int* slot_step0;                      // global step at which the slot's step 0 was sampled (0 for the first request)
  int* slot_P;                          // prompt length of the slot's request
  int* slot_uid;                        // utterance id keying the slot's Philox stream (default: slot + slot_base)
  const long long* const* slot_prompt;  // the slot's prompt row (caller-owned, alive for the session)
  int* gen;      // [B0][max_steps]
  int* sampled;  // [B0][max_steps]
  int* greedy_rec;  // optional [B0][max_steps]: argmax of the penalised logits (test hook)
  const int* forced;
  int n_forced;
  float* logits_rec;
  int n_logits_rec;
  int hook_rows;  // 0: the two hooks above are indexed by session slot ([.][B0]); n > 0: by utterance id, n rows (T2S_OPT_HOOKS_BY_UTTERANCE)
  uint32_t* seen;  // [B0][SEEN_WORDS]
  // persistent-kernel grid barrier + watchdog, statistics
  unsigned* bar;
  int* abort_flag;
  unsigned long long* stats;  // [0] kv positions, [1] steps, [2] sequence-steps
  long long* timeline;        // measurement hook: [ncta][tl_slots][2] (arrive, release) SM clocks of one step
  int tl_step, tl_slots;
};

// ---- loads / stores ------------------------------------------------------------------------------
// Weights are immutable for the kernel's lifetime: read-only path, do not pollute L1.
__device__ __forceinline__ uint4 ld_weight16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// Data produced by other CTAs inside the same (persistent) kernel: L2-coherent loads that bypass L1.
__device__ __forceinline__ uint4 ld_cg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ld_cg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_cg_f(const float* p) { return __ldcg(p); }
__device__ __forceinline__ int ld_cg_i(const int* p) { return __ldcg(p); }

// volatile variants: keep their program order, so a batch of independent loads is issued back to back
// (one L2 round trip) instead of being sunk to their first use by the compiler
__device__ __forceinline__ uint4 ldv_cg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldv_cg_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ long long ldv_cg_ll(const long long* p) {
  long long r;
  asm volatile("ld.global.cg.s64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);  // .x = a (low half), .y = b
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// ---- grid barrier for the persistent kernel ---------------------------------------------------------
// Monotonic arrival counter (zeroed by the host before launch).  Thread 0 of every CTA arrives with
// red.release.gpu and spins with ld.acquire.gpu (measured 4% faster per step than relaxed polling + fence); a watchdog turns a lost CTA into an error instead
// of a hung GPU.
struct GridBarrier {
  unsigned* counter;
  int* abort_flag;
  unsigned target;
  unsigned ncta;
  long long* tl;  // timeline row of this CTA (or nullptr)
  int tl_k, tl_n;
  __device__ __forceinline__ void init(unsigned* c, int* a, unsigned n) {
    counter = c; abort_flag = a; target = 0; ncta = n; tl = nullptr; tl_k = 0; tl_n = 0;
  }
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      if (tl && tl_k < tl_n) tl[2 * tl_k] = clock64();
      target += ncta;
      // release: cumulative over the CTA's writes ordered before it by the bar.sync above
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      unsigned spins = 0;
      long long t0 = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) break;
        if ((++spins & 0x3FFu) == 0) {
          long long now = clock64();
          if (t0 == 0) t0 = now;
          if (now - t0 > 4000000000ll || __ldcg(abort_flag) != 0) {  // ~2 s at 1.9 GHz
            atomicCAS(abort_flag, 0, ABORT_WATCHDOG);
            break;
          }
        }
      }
      if (tl && tl_k < tl_n) { tl[2 * tl_k + 1] = clock64(); ++tl_k; }
    }
    __syncthreads();
  }
};

}  // namespace t2s
