// engine.cu — host side of libt2s_b200.so: the C ABI of include/t2s_b200.h.
// Owns the packed weights, the paged KV pool and the session state; launches the kernels of
// kernels.cuh.  No torch, no CPU compute path: everything numerical happens in the CUDA kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/t2s_b200.h"
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "cluster_decode.cuh"
#include "wide_decode.cuh"

using namespace t2s;

static thread_local std::string g_err;
// The persistent decode kernels (cluster-stream, wide) synchronise their CTAs with hand-rolled grid barriers and therefore need
// every CTA co-resident.  Two engines of one process decoding at the same time on one GPU could each get a partial set of
// SMs and spin until the watchdog fires, so decode launches of a process are serialised (a launch is held until its kernel
// has finished: t2s_decode synchronises anyway).  Other PROCESSES on the same GPU are the operator's business.
static std::mutex g_decode_mu;

static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct t2s_engine {
  t2s_model_config cfg{};
  int device = 0, num_sms = 0;
  // weights
  DevBuf wmat, wvec, whead, wbert, bbert, emb_audio, emb_text, pe, wrow;  // wrow: row-major bf16 copies for the TMA GEMM
  DevBuf wstream, hstream;  // cluster-stream decode: per-(layer, CTA rank) consumption-ordered weight streams
  int max_clusters = 0;     // co-resident 16-CTA clusters of k_decode_cluster (0: unavailable)
  DevBuf wwide, llbuf;      // wide decode (wide_decode.cuh): packed weights by (head, tile); hand-off cells {value, tag}
  bool wide_ok = false;     // 144 CTAs of k_decode_wide can be co-resident
  DevBuf wrow_g, wrow_head, wrow_head_g, head_c;  // gamma-folded row-major copies (Wqkv, W1 per layer; head) + head c1/c0
  bool weights_final = false;
  std::vector<CUtensorMap> tcp_wmaps;  // persistent prefill GEMM: the four weight maps of every layer (built on first use)
  float alpha_audio = 1.f, alpha_text = 1.f;
  std::vector<char> loaded;  // per (tensor, layer)
  // kv pool
  DevBuf kpool, vpool;
  size_t pool_pages = 0;
  // session buffers
  DevBuf ints, ints2, kvoff, attn_desc, x0_rows, x0_slots, x0b_rows, x0b_slots, yb1, yb2, sp1, sp2, q, attn, y1, h, y2, stat2, logits, part, seg_cnt, gen, sampled, seen, misc, bert_rows;
  DevBuf in_ids, in_prompt, in_bert, in_bert_ptrs, out_tokens, out_idx, latent_flag;
  Ctx cp{}, cd{};  // prefill / decode contexts
  bool session = false;
  int B = 0, P = 0, T = 0, n_text = 0, max_steps = 0;  // B: slots in use (t2s_admit grows it)
  int cap = 0, maxP = 0, sess_max_pages = 0;           // session geometry: slot capacity, longest prompt, K/V pages per slot
  int session_slots = 0, session_positions = 0;        // options: slots / positions per slot to reserve at t2s_prefill
  size_t next_page = 0;
  std::vector<int> h_page_table, h_slot_text_len, h_slot_local, slot_pages;  // host mirrors per slot (pages handed to the slot so far)
  std::vector<char> slot_free;     // released by t2s_release_slots: t2s_admit reuses them (their K/V pages too) before fresh ones
  std::vector<int> pending_uids;   // t2s_set_utterance_ids: Philox ids of the next request's utterances
  DevBuf slot_aux;                 // device: [text length by slot | request-local index by slot]
  DevBuf page_tab, drows, slot_prompt;                 // session-persistent: page table, decode row descriptors, per-slot prompt pointers
  std::vector<int> h_text_len, h_s0;
  int n_qtiles = 0;
  // device int layout inside `ints`
  int *d_row_slot = nullptr, *d_row_pos = nullptr, *d_head_rows = nullptr, *d_text_off = nullptr, *d_text_len = nullptr,
      *d_s0 = nullptr, *d_trow_slot = nullptr, *d_trow_j = nullptr, *d_trow_row = nullptr, *d_page_table = nullptr;
  QTile* d_qtiles = nullptr;
  // hooks / options
  const int* forced = nullptr;
  int n_forced = 0;
  float* logits_rec = nullptr;
  int n_logits_rec = 0;
  int hook_rows = 0;  // T2S_OPT_HOOKS_BY_UTTERANCE
  long long* timeline = nullptr;
  int tl_step = 0, tl_slots = 0;
  int decode_mode = 5, prefill_gemm = 0, num_ctas = 0, check_steps = 16, tc_decode_min_batch = 160;
  int graph_mode = -1;
  int slot_base = 0;  // set by t2s_generate while it runs a large batch in chunks
  bool tc_ok = false;
  DevBuf xf, xb;  // tcgen05 prefill path: LayerNorm'ed rows (fp32 residual + bf16 GEMM operand)
  // graph cache (decode_mode 0)
  cudaGraphExec_t graph_exec = nullptr;
  Ctx graph_ctx{};
  int nodes_per_step = 0;
  // stats
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  t2s_stats st{};
  long long launches = 0;
  int* h_pinned = nullptr;  // [4] pinned: n_active, step, abort
};

static size_t dtype_size(int dt) { return dt == T2S_F32 ? 4 : 2; }

extern "C" const char* t2s_last_error(void) { return g_err.c_str(); }

extern "C" int t2s_create(const t2s_model_config* cfg, t2s_engine** out) {
  if (!cfg || !out) return fail("t2s_create: null argument");
  if (cfg->d_model != D || cfg->n_head != NH || cfg->d_ff != FF || cfg->vocab != V || cfg->bert_dim != BERT ||
      cfg->eos != V - 1)
    return fail("t2s_create: this build is specialised for d_model=512, n_head=16, d_ff=2048, vocab=1025, "
                "bert_dim=1024, eos=1024 (got %d/%d/%d/%d/%d/%d)",
                cfg->d_model, cfg->n_head, cfg->d_ff, cfg->vocab, cfg->bert_dim, cfg->eos);
  if (cfg->n_layer < 1 || cfg->n_layer > 64) return fail("t2s_create: n_layer out of range");
  if (cfg->max_batch < 1 || cfg->max_batch > MAX_B) return fail("t2s_create: max_batch must be in [1,%d]", MAX_B);
  if (cfg->pe_len < 16) return fail("t2s_create: pe_len too small");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("t2s_create: no CUDA device; this library has no CPU fallback");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail("t2s_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major,
                prop.minor);
  t2s_engine* e = new t2s_engine();
  e->cfg = *cfg;
  e->device = dev;
  e->num_sms = prop.multiProcessorCount;
  const int L = cfg->n_layer;
  int rc = 0;
  rc |= e->wmat.ensure((size_t)L * LW * 2);
  rc |= e->wvec.ensure((size_t)L * LV * 4);
  rc |= e->whead.ensure((size_t)VPAD * D * 2);
  rc |= e->wbert.ensure((size_t)D * BERT * 2);
  rc |= e->bbert.ensure(D * 4);
  rc |= e->emb_audio.ensure((size_t)V * D * 2);
  rc |= e->emb_text.ensure((size_t)cfg->phoneme_vocab * D * 2);
  rc |= e->pe.ensure((size_t)cfg->pe_len * D * 4);
  rc |= e->wrow.ensure(((size_t)L * LW + (size_t)D * BERT) * 2);
  rc |= e->wrow_g.ensure((size_t)L * (3 * D + FF) * D * 2);
  rc |= e->wrow_head.ensure((size_t)VPAD * D * 2);
  rc |= e->wrow_head_g.ensure((size_t)VPAD * D * 2);
  rc |= e->head_c.ensure((size_t)2 * VPAD * 4);
  rc |= e->wstream.ensure((size_t)L * cs::C * cs::LAYER_BYTES);
  rc |= e->hstream.ensure((size_t)cs::C * cs::HEAD_BYTES);
  rc |= e->wwide.ensure((size_t)L * ws::WL_BYTES + ws::WH_BYTES);
  rc |= e->llbuf.ensure(ws::LL_CELLS * 8);
  rc |= e->x0b_slots.ensure((size_t)MAX_B * D * 2);
  rc |= e->logits.ensure((size_t)MAX_B * VPAD * 4);
  rc |= e->part.ensure((size_t)(MAX_B + 1024) * PART_STRIDE * 4);
  rc |= e->seg_cnt.ensure(MAX_B * 4);
  rc |= e->attn_desc.ensure((size_t)1024 * 2 * 16);
  rc |= e->misc.ensure(256);
  rc |= e->x0_slots.ensure((size_t)MAX_B * D * 4);
  rc |= e->seen.ensure((size_t)MAX_B * SEEN_WORDS * 4);
  if (rc) { t2s_destroy(e); return 1; }
  cudaMemset(e->seg_cnt.p, 0, MAX_B * 4);
  cudaMemset(e->misc.p, 0, 256);
  e->loaded.assign((size_t)T2S_W_COUNT * (L + 1), 0);
  if (cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess ||
      cudaMallocHost(&e->h_pinned, 64) != cudaSuccess) {
    t2s_destroy(e);
    return fail("t2s_create: event / pinned allocation failed");
  }
  // opt in to > 48 KB dynamic shared memory for the projection / persistent kernels
  cudaFuncSetAttribute(k_phase<PH_QKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_phase<PH_OPROJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_phase<PH_FFN1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_phase<PH_FFN2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_phase<PH_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_phase<PH_ATTN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_phase<PH_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_phase<PH_PLAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_bert_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  cudaFuncSetAttribute(k_decode_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
  e->tc_ok = gemm_tc_init();
  {
    // cluster-stream decode kernel: 16-CTA clusters (non-portable size), ~200 KB dynamic shared memory per CTA
    cudaError_t ce = cudaFuncSetAttribute(cs::k_decode_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cs::Smem));
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(cs::k_decode_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (ce == cudaSuccess) {
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(cs::C); lc.blockDim = dim3(cs::NTC); lc.dynamicSmemBytes = sizeof(cs::Smem);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs::C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      int mc = 0;
      if (cudaOccupancyMaxActiveClusters(&mc, cs::k_decode_cluster, &lc) == cudaSuccess) e->max_clusters = mc;
    }
    cudaGetLastError();  // a failure here only disables decode mode 4
  }
  {
    // wide decode kernel: 144 CTAs (16 heads x 9), one per SM, cooperative launch
    int per_sm = 0;
    if (cudaFuncSetAttribute(ws::k_decode_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ws::Smem)) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ws::k_decode_wide, ws::NTW, sizeof(ws::Smem)) == cudaSuccess)
      e->wide_ok = per_sm >= 1 && e->num_sms >= ws::G && prop.cooperativeLaunch;
    cudaGetLastError();  // a failure here only disables decode mode 6
    cudaMemset(e->llbuf.p, 0, ws::LL_CELLS * 8);
  }
  e->prefill_gemm = e->tc_ok ? 1 : 0;  // tcgen05/TMEM + TMA GEMMs for prefill unless the driver entry point is missing
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    t2s_destroy(e);
    return fail("t2s_create: kernel attribute setup failed: %s", cudaGetErrorString(err));
  }
  e->st.num_sms = e->num_sms;
  e->st.weight_bytes_per_step = ((int64_t)L * LW + (int64_t)V * D) * 2 + (int64_t)L * LV * 4;
  e->st.kv_bytes_per_position = (int64_t)L * 2 * D * 2;
  *out = e;
  return 0;
}

extern "C" void t2s_destroy(t2s_engine* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&e->wmat, &e->wvec, &e->whead, &e->wbert, &e->bbert, &e->emb_audio, &e->emb_text, &e->pe, &e->wrow,
                    &e->wrow_g, &e->wrow_head, &e->wrow_head_g, &e->head_c, &e->wstream, &e->hstream, &e->wwide, &e->llbuf, &e->page_tab, &e->drows, &e->slot_prompt, &e->slot_aux, &e->x0b_rows, &e->x0b_slots, &e->yb1, &e->yb2, &e->sp1, &e->sp2,
                    &e->kpool, &e->vpool, &e->ints, &e->ints2, &e->kvoff, &e->attn_desc, &e->x0_rows, &e->x0_slots, &e->q, &e->attn, &e->y1, &e->h,
                    &e->y2, &e->stat2, &e->logits, &e->part, &e->seg_cnt, &e->gen, &e->sampled, &e->seen, &e->misc,
                    &e->bert_rows, &e->xf, &e->xb, &e->in_ids, &e->in_prompt, &e->in_bert, &e->in_bert_ptrs, &e->out_tokens, &e->out_idx, &e->latent_flag};
  for (DevBuf* b : bufs) b->release();
  if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  delete e;
}

// ---- weights ---------------------------------------------------------------------------------------------
extern "C" int t2s_load_tensor(t2s_engine* e, int32_t id, int32_t layer, const void* data, int32_t dtype,
                               int64_t numel, int32_t on_device, void* stream_) {
  if (!e || !data) return fail("t2s_load_tensor: null argument");
  if (id < 0 || id >= T2S_W_COUNT) return fail("t2s_load_tensor: bad tensor id %d", id);
  if (dtype < 0 || dtype > 2) return fail("t2s_load_tensor: bad dtype %d", dtype);
  cudaStream_t s = (cudaStream_t)stream_;
  const bool per_layer = id >= T2S_W_IN_PROJ_W;
  if (per_layer && (layer < 0 || layer >= e->cfg.n_layer)) return fail("t2s_load_tensor: layer %d out of range", layer);
  struct Spec { int64_t n; int rows, cols; };
  const int PV = e->cfg.phoneme_vocab;
  Spec sp;
  switch (id) {
    case T2S_W_BERT_PROJ_W: sp = {(int64_t)D * BERT, D, BERT}; break;
    case T2S_W_BERT_PROJ_B: sp = {D, 0, 0}; break;
    case T2S_W_TEXT_EMB: sp = {(int64_t)PV * D, PV, D}; break;
    case T2S_W_TEXT_ALPHA: case T2S_W_AUDIO_ALPHA: sp = {1, 0, 0}; break;
    case T2S_W_AUDIO_EMB: sp = {(int64_t)V * D, V, D}; break;
    case T2S_W_PE: sp = {(int64_t)e->cfg.pe_len * D, 0, 0}; break;
    case T2S_W_PREDICT: sp = {(int64_t)V * D, V, D}; break;
    case T2S_W_IN_PROJ_W: sp = {(int64_t)3 * D * D, 3 * D, D}; break;
    case T2S_W_IN_PROJ_B: sp = {3 * D, 0, 0}; break;
    case T2S_W_OUT_PROJ_W: sp = {(int64_t)D * D, D, D}; break;
    case T2S_W_LIN1_W: sp = {(int64_t)FF * D, FF, D}; break;
    case T2S_W_LIN1_B: sp = {FF, 0, 0}; break;
    case T2S_W_LIN2_W: sp = {(int64_t)D * FF, D, FF}; break;
    default: sp = {D, 0, 0}; break;  // out_proj bias, lin2 bias, norms
  }
  if (numel != sp.n) return fail("t2s_load_tensor: tensor %d expects %lld elements, got %lld", id, (long long)sp.n, (long long)numel);
  // stage on device
  const void* src = data;
  DevBuf tmp;
  if (!on_device) {
    if (tmp.ensure((size_t)numel * dtype_size(dtype))) return 1;
    CK(cudaMemcpyAsync(tmp.p, data, (size_t)numel * dtype_size(dtype), cudaMemcpyHostToDevice, s));
    src = tmp.p;
  }
  const int blocks = 592, threads = 256;
  auto pack = [&](bf16* dst, int N, int K, int tiles) {
    k_pack_matrix<<<blocks, threads, 0, s>>>(dst, src, dtype, N, K, tiles);
    e->launches++;
  };
  auto cvt_f32 = [&](float* dst, size_t n) {
    k_convert_f32<<<blocks, threads, 0, s>>>(dst, src, dtype, n);
    e->launches++;
  };
  auto cvt_b16 = [&](bf16* dst, size_t n) {
    k_convert_bf16<<<blocks, threads, 0, s>>>(dst, src, dtype, n);
    e->launches++;
  };
  bf16* wl = e->wmat.as<bf16>() + (size_t)layer * LW;
  bf16* wr = e->wrow.as<bf16>() + (size_t)layer * LW;  // row-major copies (TMA GEMM operand B)
  float* vl = e->wvec.as<float>() + (size_t)layer * LV;
  switch (id) {
    case T2S_W_BERT_PROJ_W:
      pack(e->wbert.as<bf16>(), D, BERT, D / 16);
      cvt_b16(e->wrow.as<bf16>() + (size_t)e->cfg.n_layer * LW, (size_t)D * BERT);
      break;
    case T2S_W_BERT_PROJ_B: cvt_f32(e->bbert.as<float>(), D); break;
    case T2S_W_TEXT_EMB: cvt_b16(e->emb_text.as<bf16>(), (size_t)PV * D); break;
    case T2S_W_AUDIO_EMB: cvt_b16(e->emb_audio.as<bf16>(), (size_t)V * D); break;
    case T2S_W_TEXT_ALPHA: case T2S_W_AUDIO_ALPHA: {
      cvt_f32(e->misc.as<float>() + 32, 1);
      float v = 0.f;
      CK(cudaMemcpyAsync(&v, e->misc.as<float>() + 32, 4, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      (id == T2S_W_TEXT_ALPHA ? e->alpha_text : e->alpha_audio) = v;
      break;
    }
    case T2S_W_PE: cvt_f32(e->pe.as<float>(), (size_t)e->cfg.pe_len * D); break;
    case T2S_W_PREDICT: cvt_b16(e->wrow_head.as<bf16>(), (size_t)V * D); break;  // packed by finalize_weights (LN fold)
    case T2S_W_IN_PROJ_W: cvt_b16(wr + OFF_WQKV, (size_t)3 * D * D); break;  // packed by finalize_weights
    case T2S_W_OUT_PROJ_W: pack(wl + OFF_WO, D, D, D / 16); cvt_b16(wr + OFF_WO, (size_t)D * D); break;
    case T2S_W_LIN1_W: cvt_b16(wr + OFF_W1, (size_t)FF * D); break;  // packed by finalize_weights
    case T2S_W_LIN2_W: pack(wl + OFF_W2, D, FF, D / 16); cvt_b16(wr + OFF_W2, (size_t)D * FF); break;
    case T2S_W_IN_PROJ_B: cvt_f32(vl + VO_BQKV, 3 * D); break;
    case T2S_W_OUT_PROJ_B: cvt_f32(vl + VO_BO, D); break;
    case T2S_W_LIN1_B: cvt_f32(vl + VO_B1, FF); break;
    case T2S_W_LIN2_B: cvt_f32(vl + VO_B2, D); break;
    case T2S_W_NORM1_W: cvt_f32(vl + VO_G1, D); break;
    case T2S_W_NORM1_B: cvt_f32(vl + VO_BE1, D); break;
    case T2S_W_NORM2_W: cvt_f32(vl + VO_G2, D); break;
    case T2S_W_NORM2_B: cvt_f32(vl + VO_BE2, D); break;
    default: break;
  }
  CK(cudaGetLastError());
  if (!on_device) { CK(cudaStreamSynchronize(s)); tmp.release(); }
  e->loaded[(size_t)id * (e->cfg.n_layer + 1) + (per_layer ? layer : 0)] = 1;
  e->weights_final = false;
  e->tcp_wmaps.clear();
  return 0;
}

static int check_loaded(t2s_engine* e) {
  const int L = e->cfg.n_layer;
  for (int id = 0; id < T2S_W_COUNT; ++id) {
    const bool per_layer = id >= T2S_W_IN_PROJ_W;
    for (int l = 0; l < (per_layer ? L : 1); ++l)
      if (!e->loaded[(size_t)id * (L + 1) + l]) return fail("weights incomplete: tensor id %d layer %d was never loaded", id, l);
  }
  return 0;
}

// Folds every LayerNorm that precedes a projection into that projection's packed weights (phases.cuh) and
// packs the folded matrices into MMA-fragment order.  Runs once after (re)loading weights.
static int finalize_weights(t2s_engine* e, cudaStream_t s) {
  if (e->weights_final) return 0;
  const int L = e->cfg.n_layer;
  const size_t GL = (size_t)(3 * D + FF) * D;  // folded row-major elements per layer
  for (int l = 0; l < L; ++l) {
    float* vl = e->wvec.as<float>() + (size_t)l * LV;
    const float* vp = e->wvec.as<float>() + (size_t)(l > 0 ? l - 1 : 0) * LV;
    const bf16* wr = e->wrow.as<bf16>() + (size_t)l * LW;
    bf16* wg = e->wrow_g.as<bf16>() + (size_t)l * GL;
    bf16* wl = e->wmat.as<bf16>() + (size_t)l * LW;
    k_fold_ln<<<(3 * D + 7) / 8, 256, 0, s>>>(wr + OFF_WQKV, l > 0 ? vp + VO_G2 : nullptr, l > 0 ? vp + VO_BE2 : nullptr,
                                            vl + VO_BQKV, wg, vl + VO_C1_QKV, vl + VO_C0_QKV, 3 * D, 3 * D);
    k_pack_matrix<<<592, 256, 0, s>>>(wl + OFF_WQKV, wg, T2S_BF16, 3 * D, D, 3 * D / 16);
    k_fold_ln<<<(FF + 7) / 8, 256, 0, s>>>(wr + OFF_W1, vl + VO_G1, vl + VO_BE1, vl + VO_B1, wg + (size_t)3 * D * D,
                                         vl + VO_C1_FFN1, vl + VO_C0_FFN1, FF, FF);
    k_pack_matrix<<<592, 256, 0, s>>>(wl + OFF_W1, wg + (size_t)3 * D * D, T2S_BF16, FF, D, FF / 16);
    e->launches += 4;
  }
  const float* vlast = e->wvec.as<float>() + (size_t)(L - 1) * LV;
  k_fold_ln<<<(VPAD + 7) / 8, 256, 0, s>>>(e->wrow_head.as<bf16>(), vlast + VO_G2, vlast + VO_BE2, nullptr,
                                         e->wrow_head_g.as<bf16>(), e->head_c.as<float>(), e->head_c.as<float>() + VPAD, V, VPAD);
  k_pack_matrix<<<592, 256, 0, s>>>(e->whead.as<bf16>(), e->wrow_head_g.as<bf16>(), T2S_BF16, VPAD, D, VT);
  cs::k_pack_stream<<<1184, 256, 0, s>>>(e->wstream.as<unsigned char>(), e->hstream.as<bf16>(), e->wrow.as<bf16>(), e->wvec.as<float>(), e->wrow_head.as<bf16>(), L);
  ws::k_pack_wide<<<1184, 256, 0, s>>>(e->wwide.as<unsigned char>(), e->wrow.as<bf16>(), e->wvec.as<float>(), e->wrow_head.as<bf16>(), L);
  e->launches += 4;
  CK(cudaGetLastError());
  e->weights_final = true;
  return 0;
}

// ---- options / hooks ----------------------------------------------------------------------------------
extern "C" int t2s_set_option(t2s_engine* e, int32_t opt, int64_t v) {
  if (!e) return fail("t2s_set_option: null engine");
  switch (opt) {
    case T2S_OPT_DECODE_MODE: if (v < 0 || v > 6) return fail("decode mode must be 0 .. 6"); e->decode_mode = (int)v; break;
    case T2S_OPT_PREFILL_GEMM:
      if (v < 0 || v > 2) return fail("prefill gemm must be 0, 1 or 2");
      if (v >= 1 && !e->tc_ok) return fail("tcgen05 GEMM unavailable: cuTensorMapEncodeTiled entry point not found");
      e->prefill_gemm = (int)v; break;
    case T2S_OPT_NUM_CTAS: if (v != 0 && (v < MAX_B / 2 || v > 1024)) return fail("num_ctas must be 0 or in [128, 1024]"); e->num_ctas = (int)v; break;
    case T2S_OPT_TC_DECODE_MIN_BATCH: if (v < 0 || v > 100000) return fail("tc_decode_min_batch out of range"); e->tc_decode_min_batch = (int)v; break;
    case T2S_OPT_SESSION_SLOTS: if (v < 0 || v > MAX_B) return fail("session_slots must be in [0,%d]", MAX_B); e->session_slots = (int)v; break;
    case T2S_OPT_HOOKS_BY_UTTERANCE: if (v < 0 || v > 1000000) return fail("hooks_by_utterance out of range"); e->hook_rows = (int)v; break;
    case T2S_OPT_SESSION_POSITIONS: if (v < 0 || v > 4000) return fail("session_positions must be in [0,4000]"); e->session_positions = (int)v; break;
    case T2S_OPT_CHECK_STEPS: if (v < 1 || v > 4096) return fail("check_steps out of range"); e->check_steps = (int)v; break;
    default: return fail("unknown option %d", opt);
  }
  return 0;
}
extern "C" int t2s_set_forced_tokens(t2s_engine* e, const int32_t* forced, int32_t n) {
  if (!e) return fail("null engine");
  e->forced = forced; e->n_forced = forced ? n : 0;
  return 0;
}
extern "C" int t2s_set_logits_capture(t2s_engine* e, float* buf, int32_t n) {
  if (!e) return fail("null engine");
  e->logits_rec = buf; e->n_logits_rec = buf ? n : 0;
  return 0;
}

// ---- launch helpers --------------------------------------------------------------------------------------
template <int PH>
static void launch_phase(t2s_engine* e, const Ctx& c, int layer, int grid, cudaStream_t s) {
  k_phase<PH><<<grid, NT, SMEM_MAX, s>>>(c, layer);
  e->launches++;
}

static void launch_decode_step(t2s_engine* e, const Ctx& c, cudaStream_t s) {
  const int g = e->num_sms;
  for (int l = 0; l < c.n_layer; ++l) {
    launch_phase<PH_QKV>(e, c, l, g, s);
    launch_phase<PH_ATTN>(e, c, l, g, s);
    launch_phase<PH_OPROJ>(e, c, l, g, s);
    launch_phase<PH_FFN1>(e, c, l, g, s);
    launch_phase<PH_FFN2>(e, c, l, g, s);
  }
  launch_phase<PH_HEAD>(e, c, 0, g, s);
  launch_phase<PH_SAMPLE>(e, c, 0, std::min(std::max(c.B0, 1), g), s);
  launch_phase<PH_PLAN>(e, c, 0, 1, s);
}

// Large-batch decode step: projections on the tensor cores (k_gemm_tc<64>, LayerNorm folded into the epilogue,
// A operand = the bf16 activations the previous kernel's epilogue wrote, fetched by TMA), attention / head /
// sampler / plan as phase kernels.  Captured into a CUDA graph by t2s_decode.
// (Round 2 also ran the step as a 256-row "prefill" - k_ln_rows + the persistent k_gemm_tcp<128> with the three prefill epilogues, 7
// launches per layer: parity green and SLOWER, 1702 / 1958 / 2714 us per step at batch 64 / 128 / 256 against 1598 / 1917 / 2403 for this
// form: at M <= 256 every projection is a fixed ~10-15 us latency chain whatever the kernel; removed.)
static bool launch_decode_step_tc(t2s_engine* e, const Ctx& c, cudaStream_t s) {
  const int g = e->num_sms, B0 = c.B0;
  float* x0r = e->x0_rows.as<float>();
  bf16* x0br = e->x0b_rows.as<bf16>();
  const size_t GL = (size_t)(3 * D + FF) * D;
  k_gather_x0<<<B0, 128, 0, s>>>(c, x0r, x0br);
  e->launches++;
  bool ok = true;
  for (int l = 0; l < c.n_layer && ok; ++l) {
    const float* vl = c.wvec + (size_t)l * LV;
    const float* vp = c.wvec + (size_t)(l > 0 ? l - 1 : 0) * LV;
    const bf16* wr = e->wrow.as<bf16>() + (size_t)l * LW;
    const bf16* wg = e->wrow_g.as<bf16>() + (size_t)l * GL;
    TcEpilogue ep{};
    ep.error_flag = c.abort_flag; ep.n_rows = c.n_rows;
    ep.mode = EPI_D_QKV; ep.bias = vl + VO_C0_QKV; ep.c1 = vl + VO_C1_QKV; ep.sp_in = l > 0 ? c.sp2 : nullptr;
    ep.stat_out = l > 0 ? c.stat2 : nullptr; ep.out_f32 = c.q; ep.kpool = c.kpool; ep.vpool = c.vpool;
    ep.kvoff = c.row_kvoff; ep.layer_off = (size_t)l * c.kv_layer_stride;
    ok = ok && launch_gemm_tc<64>(l == 0 ? x0br : c.yb2, wg, B0, 3 * D, D, ep, s);
    launch_phase<PH_ATTN>(e, c, l, g, s);
    ep = TcEpilogue{};
    ep.error_flag = c.abort_flag; ep.n_rows = c.n_rows;
    ep.mode = EPI_D_O; ep.bias = vl + VO_BO; ep.src_f32 = l == 0 ? x0r : c.y2; ep.stat_in = l == 0 ? nullptr : c.stat2;
    ep.g = vp + VO_G2; ep.be = vp + VO_BE2; ep.out_f32 = c.y1; ep.out_b16 = c.yb1; ep.sp_out = c.sp1;
    ok = ok && launch_gemm_tc<32>(c.attn, wr + OFF_WO, B0, D, D, ep, s);  // 512 outputs: 32-wide tiles put 2 x 16 CTAs on the weights
    ep = TcEpilogue{};
    ep.error_flag = c.abort_flag; ep.n_rows = c.n_rows;
    ep.mode = EPI_D_FFN1; ep.bias = vl + VO_C0_FFN1; ep.c1 = vl + VO_C1_FFN1; ep.sp_in = c.sp1; ep.out_b16 = c.h;
    ok = ok && launch_gemm_tc<64>(c.yb1, wg + (size_t)3 * D * D, B0, FF, D, ep, s);
    ep = TcEpilogue{};
    ep.error_flag = c.abort_flag; ep.n_rows = c.n_rows;
    ep.mode = EPI_D_FFN2; ep.src_f32 = c.y1; ep.sp_res = c.sp1; ep.g = vl + VO_G1; ep.be = vl + VO_BE1; ep.b2 = vl + VO_B2;
    ep.out_f32 = c.y2; ep.out_b16 = c.yb2; ep.sp_out = c.sp2;
    ok = ok && launch_gemm_tc<32>(c.h, wr + OFF_W2, B0, D, FF, ep, s);
    e->launches += 4;
  }
  launch_phase<PH_HEAD>(e, c, 0, g, s);
  launch_phase<PH_SAMPLE>(e, c, 0, std::min(std::max(c.B0, 1), g), s);
  launch_phase<PH_PLAN>(e, c, 0, 1, s);
  return ok;
}

// ---- prefill ---------------------------------------------------------------------------------------------
static int read_state(t2s_engine* e, cudaStream_t s, int* n_active, int* step, int* aborted);

// One implementation for t2s_prefill (a new session: slots [0, B)) and t2s_admit (B more utterances join the resident session
// in slots [e->B, e->B + B) at the session's current global step: continuous batching).  The request-local arrays (row
// descriptors of the prompt rows, text offsets, q-tiles) live in per-call buffers; what the decode loop needs across calls
// (page table, decode row descriptors, per-slot prompt pointers, token buffers) lives in session buffers sized for the
// session's slot capacity.
static int prefill_impl(t2s_engine* e, const t2s_request* rq, cudaStream_t s, bool admit) {
  const char* who = admit ? "t2s_admit" : "t2s_prefill";
  if (check_loaded(e)) return 1;
  if (finalize_weights(e, s)) return 1;
  const int B = rq->batch, P = rq->prompt_len;
  if (B < 1 || B > e->cfg.max_batch) return fail("%s: batch %d outside [1,%d]", who, B, e->cfg.max_batch);
  if (P < 0) return fail("%s: negative prompt_len", who);
  if (P > 0 && !rq->prompt) return fail("%s: prompt_len > 0 but prompt is NULL", who);
  if (rq->top_k < 1) return fail("%s: top_k must be >= 1 (the reference's torch.topk fails otherwise)", who);
  if (rq->max_steps < 1) return fail("%s: max_steps must be >= 1", who);
  if (!(rq->repetition_penalty > 0.f)) return fail("%s: repetition_penalty must be > 0", who);
  if (rq->bert_dtype < 0 || rq->bert_dtype > 2) return fail("%s: bad bert_dtype", who);
  if (!rq->phoneme_ids || !rq->phoneme_lens || !rq->bert || !rq->bert_stride_c || !rq->bert_stride_t)
    return fail("%s: null input pointer", who);
  // early_stop_num: -1 = off; any other value n stops once more than n tokens were sampled (t2s_model.py:747/:897:
  // `early_stop_num != -1 and (len - prefix) > early_stop_num`), so n < -1 behaves like 0: stop at the first step
  const int early_stop = rq->early_stop_num == -1 ? -1 : std::max(rq->early_stop_num, 0);
  int steps_cap = rq->max_steps;
  if (early_stop >= 0) steps_cap = std::min(steps_cap, early_stop + 1);
  int slot0 = 0, step0 = 0;
  if (admit) {
    if (!e->session) return fail("t2s_admit: no resident session (call t2s_prefill first)");
    if (e->forced || e->logits_rec) {
      // the per-step hooks are indexed [step][slot capacity]: fine, they were sized by the caller for the capacity
    }
    const Ctx& d = e->cd;
    if (rq->top_k != d.top_k || rq->top_p != d.top_p || rq->temperature != d.temperature || rq->repetition_penalty != d.rep_pen ||
        early_stop != d.early_stop || rq->eos_suppress_steps != d.eos_window || rq->max_steps != d.max_steps)
      return fail("t2s_admit: sampling parameters / stop rules are per session and must equal those of t2s_prefill");
    slot0 = e->B;
    int n_free = 0;
    for (int i = 0; i < e->B; ++i) n_free += e->slot_free[i] ? 1 : 0;
    if (B > n_free + (e->cap - e->B))
      return fail("t2s_admit: %d more utterances do not fit the session's %d slots (%d in use, %d of them released; reserve slots with "
                  "T2S_OPT_SESSION_SLOTS before t2s_prefill, free finished ones with t2s_release_slots)", B, e->cap, e->B, n_free);
    int n_active = 0, aborted = 0;
    if (read_state(e, s, &n_active, &step0, &aborted)) return 1;
    if (aborted) return fail("t2s_admit: the session is in an error state");
    // the running sequences have completed global step step0 - 1 (their next sample is step0): the new ones' step 0 is sampled
    // now AS step step0 - 1, so that old and new sequences take the session's next step together
    step0 -= 1;
    if (step0 < 0) return fail("t2s_admit: the session has not completed its first step");
    if (n_active + B > e->max_clusters * cs::RMAX)
      return fail("t2s_admit: %d active + %d new sequences exceed the %d the cluster-stream decode kernel holds", n_active, B, e->max_clusters * cs::RMAX);
  }
  std::vector<int> text_len(B), text_off(B), s0(B), row0(B);
  int n_text = 0, T = 0, need_pages = 0;
  for (int b = 0; b < B; ++b) {
    const int L = rq->phoneme_lens[b];
    if (L < 1) return fail("%s: utterance %d has %d phonemes", who, b, L);
    text_len[b] = L; text_off[b] = n_text; n_text += L;
    s0[b] = L + P; row0[b] = T; T += L + P;
    if (std::max(L, P + steps_cap) >= e->cfg.pe_len)
      return fail("%s: position %d exceeds the %d-entry positional table (embedding.py:52)", who, std::max(L, P + steps_cap), e->cfg.pe_len);
    need_pages = std::max(need_pages, (s0[b] + steps_cap + PAGE - 1) / PAGE);
  }
  // ---- session geometry: slot capacity and pages per slot are fixed by t2s_prefill
  if (!admit) {
    e->cap = std::min(e->cfg.max_batch, std::max(B, e->session_slots));
    e->sess_max_pages = std::max(need_pages, (e->session_positions + PAGE - 1) / PAGE);
    e->next_page = 0;
    e->h_page_table.assign((size_t)e->cap * e->sess_max_pages, 0);
  } else if (need_pages > e->sess_max_pages) {
    return fail("t2s_admit: an utterance needs %d K/V pages, the session was opened with %d per slot (T2S_OPT_SESSION_POSITIONS)", need_pages, e->sess_max_pages);
  }
  const int max_pages = e->sess_max_pages, cap = e->cap;
  if (!admit) {
    e->h_slot_text_len.assign(cap, 0); e->h_slot_local.assign(cap, 0); e->slot_pages.assign(cap, 0); e->slot_free.assign(cap, 0);
  }
  // the request's utterances -> session slots: a fresh session takes 0..B-1, an admission released slots first (lowest first)
  std::vector<int> new_slots(B);
  {
    int nb = 0;
    if (admit)
      for (int i = 0; i < e->B && nb < B; ++i)
        if (e->slot_free[i]) { new_slots[nb++] = i; e->slot_free[i] = 0; }
    for (int i = admit ? e->B : 0; nb < B; ++i) new_slots[nb++] = i;
  }
  std::vector<int> uids(B);
  if (!e->pending_uids.empty() && (int)e->pending_uids.size() != B)
    return fail("%s: t2s_set_utterance_ids gave %zu ids for a batch of %d", who, e->pending_uids.size(), B);
  for (int b = 0; b < B; ++b) uids[b] = e->pending_uids.empty() ? new_slots[b] + e->slot_base : e->pending_uids[b];
  e->pending_uids.clear();
  // ---- KV pool + page table (pages handed out contiguously per slot).  A fresh session sizes the pool for its CAPACITY when slots
  //      were reserved (the pool cannot grow under a resident session: its contents are the session), else for what it uses.
  for (int b = 0; b < B; ++b) {
    const int np = (s0[b] + steps_cap + PAGE - 1) / PAGE, slot = new_slots[b];
    for (int i = e->slot_pages[slot]; i < np; ++i) e->h_page_table[(size_t)slot * max_pages + i] = (int)e->next_page++;  // a reused slot keeps its pages
    e->slot_pages[slot] = std::max(e->slot_pages[slot], np);
    e->h_slot_text_len[slot] = text_len[b];
    e->h_slot_local[slot] = b;
  }
  {
    // a session opened with a reservation keeps max_pages per slot for EVERY slot - also when the first request fills all of them:
    // a recycled slot may take an utterance that is longer than its previous occupant
    const bool reserved = e->session_slots > 0 || e->session_positions > 0;
    const size_t pages = admit ? e->next_page : std::max(e->next_page, (size_t)(reserved ? (size_t)cap * max_pages : 0));
    if (pages > e->pool_pages) {
      if (admit) return fail("t2s_admit: K/V pool exhausted (%zu pages needed, %zu reserved)", pages, e->pool_pages);
      const size_t bytes = (size_t)e->cfg.n_layer * pages * KV_PAGE_STRIDE * 2;  // K and V interleaved per (page, head)
      CK(cudaStreamSynchronize(s));
      if (e->kpool.ensure(bytes)) return 1;
      e->pool_pages = pages;
    }
  }
  const std::vector<int>& page_table = e->h_page_table;
  // ---- host-built index arrays of THIS request's prompt rows
  std::vector<int> row_slot(T), row_pos(T), head_rows(B), trow_slot(n_text), trow_j(n_text), trow_row(n_text);
  std::vector<QTile> qtiles;
  for (int b = 0, r = 0, tr = 0; b < B; ++b) {
    for (int j = 0; j < s0[b]; ++j, ++r) {
      row_slot[r] = new_slots[b]; row_pos[r] = j;
      if (j < text_len[b]) { trow_slot[tr] = b; trow_j[tr] = j; trow_row[tr] = r; ++tr; }
    }
    head_rows[b] = r - 1;
    for (int q0 = 0; q0 < s0[b]; q0 += 64) qtiles.push_back(QTile{new_slots[b], q0, row0[b] + q0, std::min(64, s0[b] - q0)});
  }
  const size_t R = (size_t)std::max(T, MAX_B);
  const size_t n_ints = (size_t)2 * R + B * 6 + (size_t)3 * n_text + qtiles.size() * 4 + 64;
  int rc = 0;
  rc |= e->ints.ensure(n_ints * 4);
  rc |= e->kvoff.ensure(R * 8);
  // scratch rows that the large-batch decode graph (mode 3) points at directly: a reallocation invalidates the captured graph
  void* const scratch_before[4] = {e->x0_rows.p, e->x0b_rows.p, e->xf.p, e->xb.p};
  rc |= e->x0_rows.ensure(R * D * 4);
  rc |= e->x0b_rows.ensure(R * D * 2);
  rc |= e->yb1.ensure(R * D * 2);
  rc |= e->yb2.ensure(R * D * 2);
  rc |= e->sp1.ensure(R * 32 * 8);
  rc |= e->sp2.ensure(R * 32 * 8);
  rc |= e->q.ensure(R * D * 4);
  rc |= e->attn.ensure(R * D * 2);
  rc |= e->y1.ensure(R * D * 4);
  rc |= e->h.ensure(R * FF * 2);
  rc |= e->y2.ensure(R * D * 4);
  rc |= e->stat2.ensure(R * 8);
  rc |= e->bert_rows.ensure((size_t)n_text * BERT * 2);
  rc |= e->xf.ensure((e->prefill_gemm ? R : (size_t)MAX_B) * D * 4);
  rc |= e->xb.ensure((e->prefill_gemm ? R : (size_t)MAX_B) * D * 2);
  if (e->graph_exec && (scratch_before[0] != e->x0_rows.p || scratch_before[1] != e->x0b_rows.p || scratch_before[2] != e->xf.p ||
                        scratch_before[3] != e->xb.p)) {
    cudaGraphExecDestroy(e->graph_exec);
    e->graph_exec = nullptr;
  }
  if (!admit) {  // session buffers (a resident session's must not move)
    rc |= e->ints2.ensure((size_t)(MAX_B * 7 + 16) * 4);
    rc |= e->gen.ensure((size_t)cap * rq->max_steps * 4);
    rc |= e->sampled.ensure((size_t)cap * rq->max_steps * 4);
    rc |= e->page_tab.ensure((size_t)cap * max_pages * 4);
    rc |= e->drows.ensure((size_t)MAX_B * 16);
    rc |= e->slot_prompt.ensure((size_t)MAX_B * 8);
    rc |= e->slot_aux.ensure((size_t)MAX_B * 8);
  }
  if (rc) return 1;
  // pack the int arrays into one upload
  std::vector<int> hi(n_ints, 0);
  size_t o = 0;
  auto put = [&](const int* src, size_t n) { size_t at = o; if (n) memcpy(&hi[o], src, n * 4); o += n; return at; };
  const size_t o_row_slot = put(row_slot.data(), T); o = R;
  const size_t o_row_pos = put(row_pos.data(), T); o = 2 * R;
  const size_t o_head = put(head_rows.data(), B);
  const size_t o_toff = put(text_off.data(), B);
  const size_t o_tlen = put(text_len.data(), B);
  const size_t o_s0 = put(s0.data(), B);
  const size_t o_new = put(new_slots.data(), B);
  const size_t o_uid = put(uids.data(), B);
  const size_t o_ts = put(trow_slot.data(), n_text);
  const size_t o_tj = put(trow_j.data(), n_text);
  const size_t o_tr = put(trow_row.data(), n_text);
  o = (o + 3) & ~(size_t)3;
  const size_t o_qt = put(reinterpret_cast<const int*>(qtiles.data()), qtiles.size() * 4);
  CK(cudaMemcpyAsync(e->ints.p, hi.data(), o * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->page_tab.p, page_table.data(), (size_t)cap * max_pages * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->slot_aux.p, e->h_slot_text_len.data(), (size_t)cap * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->slot_aux.as<int>() + MAX_B, e->h_slot_local.data(), (size_t)cap * 4, cudaMemcpyHostToDevice, s));
  std::vector<long long> kvoff(T);
  for (int r = 0; r < T; ++r)
    kvoff[r] = kv_row_off(page_table[(size_t)row_slot[r] * max_pages + (row_pos[r] >> PAGE_SHIFT)], row_pos[r] & (PAGE - 1));
  CK(cudaMemcpyAsync(e->kvoff.p, kvoff.data(), (size_t)T * 8, cudaMemcpyHostToDevice, s));
  int* di = e->ints.as<int>();
  e->d_row_slot = di + o_row_slot; e->d_row_pos = di + o_row_pos; e->d_head_rows = di + o_head;
  e->d_text_off = di + o_toff; e->d_text_len = di + o_tlen; e->d_s0 = di + o_s0;
  e->d_trow_slot = di + o_ts; e->d_trow_j = di + o_tj; e->d_trow_row = di + o_tr; e->d_page_table = e->page_tab.as<int>();
  e->d_qtiles = reinterpret_cast<QTile*>(di + o_qt);
  e->n_qtiles = (int)qtiles.size();
  // ---- inputs: device pointers, or host buffers copied here (end-to-end form)
  const long long* d_ids = reinterpret_cast<const long long*>(rq->phoneme_ids);
  const long long* d_prompt = reinterpret_cast<const long long*>(rq->prompt);
  long long prompt_stride = rq->prompt_row_stride;
  std::vector<const void*> bert_ptrs(B);
  std::vector<long long> bsc(B), bst(B);
  const size_t es = dtype_size(rq->bert_dtype);
  if (rq->inputs_on_host) {
    if (admit) return fail("t2s_admit: host inputs are not supported (the session keeps pointers to the prompt rows)");
    if (e->in_ids.ensure((size_t)n_text * 8) || e->in_bert.ensure((size_t)n_text * BERT * es) ||
        e->in_prompt.ensure((size_t)std::max(1, B * P) * 8))
      return 1;
    CK(cudaMemcpyAsync(e->in_ids.p, rq->phoneme_ids, (size_t)n_text * 8, cudaMemcpyHostToDevice, s));
    d_ids = e->in_ids.as<long long>();
    if (P > 0) {
      if (rq->prompt_row_stride == 0) {
        CK(cudaMemcpyAsync(e->in_prompt.p, rq->prompt, (size_t)P * 8, cudaMemcpyHostToDevice, s));
        prompt_stride = 0;
      } else {
        CK(cudaMemcpy2DAsync(e->in_prompt.p, (size_t)P * 8, rq->prompt, (size_t)rq->prompt_row_stride * 8, (size_t)P * 8, B,
                             cudaMemcpyHostToDevice, s));
        prompt_stride = P;
      }
      d_prompt = e->in_prompt.as<long long>();
    }
    for (int b = 0; b < B; ++b) {
      if (rq->bert_stride_t[b] != 1 || rq->bert_stride_c[b] != text_len[b])
        return fail("t2s_prefill: host BERT features must be contiguous [1024, L]");
      char* dst = e->in_bert.as<char>() + (size_t)text_off[b] * BERT * es;
      CK(cudaMemcpyAsync(dst, rq->bert[b], (size_t)text_len[b] * BERT * es, cudaMemcpyHostToDevice, s));
      bert_ptrs[b] = dst; bsc[b] = text_len[b]; bst[b] = 1;
    }
  } else {
    for (int b = 0; b < B; ++b) { bert_ptrs[b] = rq->bert[b]; bsc[b] = rq->bert_stride_c[b]; bst[b] = rq->bert_stride_t[b]; }
  }
  if (e->in_bert_ptrs.ensure((size_t)B * 24)) return 1;
  CK(cudaMemcpyAsync(e->in_bert_ptrs.p, bert_ptrs.data(), (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->in_bert_ptrs.as<char>() + (size_t)B * 8, bsc.data(), (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->in_bert_ptrs.as<char>() + (size_t)B * 16, bst.data(), (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));  // host staging vectors go out of scope below; uploads are tiny
  if (!admit) { e->P = P; e->max_steps = rq->max_steps; }
  e->maxP = admit ? std::max(e->maxP, P) : P;
  e->T = T; e->n_text = n_text;
  e->h_text_len = text_len; e->h_s0 = s0;
  // ---- contexts
  int* i2 = e->ints2.as<int>();
  Ctx c{};
  if (admit) {
    c = e->cd;  // the session's model pointers, sampling parameters, token buffers, counters
  } else {
    c.wmat = e->wmat.as<bf16>(); c.wvec = e->wvec.as<float>(); c.whead = e->whead.as<bf16>(); c.wbert = e->wbert.as<bf16>();
    c.head_c1 = e->head_c.as<float>(); c.head_c0 = e->head_c.as<float>() + VPAD;
    c.bbert = e->bbert.as<float>(); c.emb_audio = e->emb_audio.as<bf16>(); c.emb_text = e->emb_text.as<bf16>();
    c.pe = e->pe.as<float>(); c.alpha_audio = e->alpha_audio; c.alpha_text = e->alpha_text;
    c.n_layer = e->cfg.n_layer; c.pe_len = e->cfg.pe_len; c.phoneme_vocab = e->cfg.phoneme_vocab;
    c.kpool = e->kpool.as<bf16>(); c.vpool = c.kpool + KV_V_OFF;
    c.kv_layer_stride = e->pool_pages * (size_t)KV_PAGE_STRIDE;
    c.page_table = e->d_page_table; c.max_pages = max_pages;
    c.n_rows = i2 + 0; c.n_active = i2 + 1; c.step = i2 + 2; c.abort_flag = i2 + 3;
    c.bar = reinterpret_cast<unsigned*>(i2 + 4);
    c.stats = reinterpret_cast<unsigned long long*>(i2 + 8);  // 3 x u64, 8-byte aligned
    c.seq_len = i2 + 16; c.active = i2 + 16 + MAX_B; c.done = i2 + 16 + 2 * MAX_B; c.out_idx = i2 + 16 + 3 * MAX_B;
    c.slot_step0 = i2 + 16 + 4 * MAX_B; c.slot_P = i2 + 16 + 5 * MAX_B; c.slot_uid = i2 + 16 + 6 * MAX_B;
    c.slot_prompt = e->slot_prompt.as<const long long*>();
    c.logits = e->logits.as<float>();
    c.part = e->part.as<float>(); c.seg_cnt = e->seg_cnt.as<int>();
    c.attn_desc = e->attn_desc.as<int>();
    c.attn_ctas = ((e->decode_mode == 1 || e->decode_mode == 5) && e->num_ctas > 0) ? std::min(e->num_ctas, e->num_sms) : e->num_sms;
    c.B0 = cap; c.max_steps = rq->max_steps; c.eos_window = rq->eos_suppress_steps;
    c.early_stop = early_stop; c.top_k = rq->top_k;
    c.top_p = rq->top_p; c.temperature = rq->temperature; c.rep_pen = rq->repetition_penalty;
    c.seed_lo = (uint32_t)(rq->seed & 0xFFFFFFFFull); c.seed_hi = (uint32_t)(rq->seed >> 32);
    c.slot_base = e->slot_base;
    c.gen = e->gen.as<int>(); c.sampled = e->sampled.as<int>();
    c.forced = e->forced; c.n_forced = e->n_forced; c.logits_rec = e->logits_rec; c.n_logits_rec = e->n_logits_rec;
    c.hook_rows = e->hook_rows;
    c.seen = e->seen.as<uint32_t>();
    c.timeline = e->timeline; c.tl_step = e->tl_step; c.tl_slots = e->tl_slots;
  }
  c.P = P;  // this request's prompt length (k_init_session stores it per slot)
  // activations of this call's rows (may have been re-allocated)
  c.q = e->q.as<float>(); c.attn = e->attn.as<bf16>(); c.y1 = e->y1.as<float>(); c.h = e->h.as<bf16>();
  c.yb1 = e->yb1.as<bf16>(); c.yb2 = e->yb2.as<bf16>(); c.sp1 = e->sp1.as<float2>(); c.sp2 = e->sp2.as<float2>();
  c.y2 = e->y2.as<float>(); c.stat2 = e->stat2.as<float2>();
  // decode context: rows of the ACTIVE sequences (own arrays: an admit's prefill must not disturb them)
  Ctx cd = c;
  cd.row_slot = e->drows.as<int>(); cd.row_pos = cd.row_slot + MAX_B; cd.row_kvoff = reinterpret_cast<long long*>(cd.row_slot + 2 * MAX_B);
  cd.x0 = e->x0_slots.as<float>(); cd.x0b = e->x0b_slots.as<bf16>(); cd.x0_by_slot = 1; cd.head_rows = nullptr;
  // prefill context: this request's prompt rows; with admit its "active list" is the list of new slots and its row count a scratch word
  Ctx cpx = c;
  cpx.row_slot = e->d_row_slot; cpx.row_pos = e->d_row_pos; cpx.row_kvoff = e->kvoff.as<long long>();
  cpx.x0 = e->x0_rows.as<float>(); cpx.x0b = e->x0b_rows.as<bf16>(); cpx.x0_by_slot = 0; cpx.head_rows = e->d_head_rows;
  Ctx cs0 = cd;  // step-0 sampler of the new utterances: logits row b <-> slot new_slots[b]
  if (admit) {
    int* scratch = reinterpret_cast<int*>(e->misc.as<char>() + 192);  // [n_rows, n_active] of the admitted request
    cpx.n_rows = scratch; cpx.n_active = scratch + 1;
    cs0.n_rows = scratch; cs0.n_active = scratch + 1; cs0.active = di + o_new; cs0.step = scratch + 2;
    const int three[3] = {T, B, step0};
    CK(cudaMemcpyAsync(scratch, three, 12, cudaMemcpyHostToDevice, s));
  }
  e->cd = cd; e->cp = cpx;
  const Ctx& cp = e->cp;
  // ---- launch
  CK(cudaEventRecord(e->ev0, s));
  const int g = e->num_sms;
  if (!admit) CK(cudaMemsetAsync(cd.abort_flag, 0, 4, s));
  Ctx ci = cd; ci.P = P;
  k_init_session<<<B, 128, 0, s>>>(ci, d_prompt, prompt_stride, e->d_s0, di + o_new, di + o_uid, step0, admit ? 0 : 1);
  if (!admit) {
    CK(cudaMemcpyAsync(cp.n_rows, &e->T, 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(e->llbuf.p, 0, 64, s));  // wide decode: the tag sequence of the hand-off cells restarts (cells are compared for equality)
  }
  k_embed_rows<<<T, 128, 0, s>>>(cp, T, d_ids, e->d_text_off, e->d_text_len, d_prompt, prompt_stride, e->slot_aux.as<int>() + MAX_B);
  const void* const* dptr = reinterpret_cast<const void* const*>(e->in_bert_ptrs.p);
  const long long* dsc = reinterpret_cast<const long long*>(e->in_bert_ptrs.as<char>() + (size_t)B * 8);
  const long long* dst_ = reinterpret_cast<const long long*>(e->in_bert_ptrs.as<char>() + (size_t)B * 16);
  if (rq->bert_dtype == T2S_F32)
    k_bert_rows<float><<<n_text, 256, 0, s>>>(e->bert_rows.as<bf16>(), dptr, dsc, dst_, e->d_trow_slot, e->d_trow_j);
  else if (rq->bert_dtype == T2S_F16)
    k_bert_rows<__half><<<n_text, 256, 0, s>>>(e->bert_rows.as<bf16>(), dptr, dsc, dst_, e->d_trow_slot, e->d_trow_j);
  else
    k_bert_rows<bf16><<<n_text, 256, 0, s>>>(e->bert_rows.as<bf16>(), dptr, dsc, dst_, e->d_trow_slot, e->d_trow_j);
  if (e->prefill_gemm == 1) {
    // bert_proj (t2s_model.py:613) on the persistent tcgen05 GEMM: [text rows, 1024] x [512, 1024]^T, added in place to the text
    // rows of x0 (which hold embedding + bias + positional term) through the row map; the bf16 copy x0b is not read on this path
    TcEpilogue ep{};
    ep.error_flag = cp.abort_flag; ep.mode = EPI_RESID; ep.resid = cp.x0; ep.out_f32 = cp.x0; ep.row_map = e->d_trow_row;
    if (!launch_gemm_tcp<256>(e->bert_rows.as<bf16>(), e->wrow.as<bf16>() + (size_t)e->cfg.n_layer * LW, n_text, D, BERT, ep, g, s))
      return fail("%s: cuTensorMapEncodeTiled failed", who);
  } else {
    k_bert_proj<<<g, NT, SMEM_MAX, s>>>(cp, e->bert_rows.as<bf16>(), e->d_trow_row, n_text);
  }
  e->launches += 4;
  const int* d_text_len_by_slot = e->slot_aux.as<int>();  // the attention kernels index text_len by session slot
  if (e->prefill_gemm) {
    // tcgen05/TMEM + TMA GEMMs (gemm_tc.cuh); LayerNorm rows are materialised once per sub-layer
    float* xf = e->xf.as<float>();
    bf16* xb = e->xb.as<bf16>();
    const int ln_blocks = (T + 7) / 8;
    bool ok = true;
    // 1: persistent kernel (128 x 256 tiles, operand ring across tiles, two TMEM accumulators); 2: one 128 x 128 tile per CTA (round 1)
    const bool persistent = e->prefill_gemm == 1;
    const int T_ = T, nsm = e->num_sms;
    // tensor maps of the persistent kernel: weights once per engine, activations (xb, attn: [T, 512]; h: [T, 2048]) once per call
    CUtensorMap map_xb, map_attn, map_h;
    if (persistent) {
      if (e->tcp_wmaps.empty()) {
        e->tcp_wmaps.resize((size_t)4 * cp.n_layer);
        for (int l = 0; l < cp.n_layer && ok; ++l) {
          const bf16* wr = e->wrow.as<bf16>() + (size_t)l * LW;
          ok = ok && make_tmap_bf16(&e->tcp_wmaps[4 * l + 0], wr + OFF_WQKV, 3 * D, D, 256) && make_tmap_bf16(&e->tcp_wmaps[4 * l + 1], wr + OFF_WO, D, D, 256) &&
               make_tmap_bf16(&e->tcp_wmaps[4 * l + 2], wr + OFF_W1, FF, D, 256) && make_tmap_bf16(&e->tcp_wmaps[4 * l + 3], wr + OFF_W2, D, FF, 256);
        }
        if (!ok) e->tcp_wmaps.clear();
      }
      ok = ok && make_tmap_bf16(&map_xb, xb, (uint64_t)T, D, TC_BM) && make_tmap_bf16(&map_attn, cp.attn, (uint64_t)T, D, TC_BM) &&
           make_tmap_bf16(&map_h, cp.h, (uint64_t)T, FF, TC_BM);
    }
    auto gemm = [&](const bf16* A, const bf16* W, int N, int K, const TcEpilogue& ep, const CUtensorMap* ma, int layer, int which) {
      if (!persistent) return launch_gemm_tc<128>(A, W, T_, N, K, ep, s);
      launch_gemm_tcp_maps<256>(*ma, e->tcp_wmaps[(size_t)4 * layer + which], T_, N, K, ep, nsm, s);
      return true;
    };
    for (int l = 0; l < cp.n_layer && ok; ++l) {
      const float* vl = cp.wvec + (size_t)l * LV;
      const bf16* wr = e->wrow.as<bf16>() + (size_t)l * LW;
      const float* resid;
      if (l == 0) {
        k_ln_rows<<<ln_blocks, 256, 0, s>>>(cp.x0, nullptr, nullptr, nullptr, xb, T, 0);
        resid = cp.x0;
      } else {
        const float* vp = cp.wvec + (size_t)(l - 1) * LV;
        k_ln_rows<<<ln_blocks, 256, 0, s>>>(cp.y2, vp + VO_G2, vp + VO_BE2, xf, xb, T, 1);
        resid = xf;
      }
      TcEpilogue ep{};
      ep.error_flag = cp.abort_flag;
      ep.mode = EPI_QKV; ep.bias = vl + VO_BQKV; ep.out_f32 = cp.q; ep.kpool = cp.kpool; ep.vpool = cp.vpool;
      ep.kvoff = cp.row_kvoff; ep.layer_off = (size_t)l * cp.kv_layer_stride;
      ok = ok && gemm(xb, wr + OFF_WQKV, 3 * D, D, ep, &map_xb, l, 0);
      k_prefill_attn_tc<<<dim3(e->n_qtiles, NH), 128, 0, s>>>(cp, l, e->d_qtiles, d_text_len_by_slot);
      ep = TcEpilogue{};
      ep.error_flag = cp.abort_flag;
      ep.mode = EPI_RESID; ep.bias = vl + VO_BO; ep.resid = resid; ep.out_f32 = cp.y1;
      ok = ok && gemm(cp.attn, wr + OFF_WO, D, D, ep, &map_attn, l, 1);
      k_ln_rows<<<ln_blocks, 256, 0, s>>>(cp.y1, vl + VO_G1, vl + VO_BE1, xf, xb, T, 1);
      ep.mode = EPI_RELU; ep.bias = vl + VO_B1; ep.resid = nullptr; ep.out_f32 = nullptr; ep.out_b16 = cp.h;
      ok = ok && gemm(xb, wr + OFF_W1, FF, D, ep, &map_xb, l, 2);
      ep.mode = EPI_RESID; ep.bias = vl + VO_B2; ep.resid = xf; ep.out_f32 = cp.y2; ep.out_b16 = nullptr;
      ok = ok && gemm(cp.h, wr + OFF_W2, D, FF, ep, &map_h, l, 3);
      e->launches += 7;
    }
    if (!ok) return fail("%s: cuTensorMapEncodeTiled failed", who);
    k_rows_stats<<<(B + 7) / 8, 256, 0, s>>>(cp.y2, e->d_head_rows, B, cp.yb2, cp.sp2);
    e->launches++;
  } else {
    for (int l = 0; l < cp.n_layer; ++l) {
      launch_phase<PH_QKV>(e, cp, l, g, s);
      k_prefill_attn<<<dim3(e->n_qtiles, NH), 64, 0, s>>>(cp, l, e->d_qtiles, d_text_len_by_slot);
      e->launches++;
      launch_phase<PH_OPROJ>(e, cp, l, g, s);
      launch_phase<PH_FFN1>(e, cp, l, g, s);
      launch_phase<PH_FFN2>(e, cp, l, g, s);
    }
  }
  launch_phase<PH_HEAD>(e, cp, 0, g, s);            // rows = the request's B utterances, gathered through head_rows
  launch_phase<PH_SAMPLE>(e, cs0, 0, std::min(B, g), s);  // step 0 sample of the new utterances; writes the decode-side x0
  if (admit) {
    k_admit<<<1, 256, 0, s>>>(e->cd, di + o_new, B);
    e->launches++;
  } else {
    launch_phase<PH_PLAN>(e, e->cd, 0, 1, s);
  }
  CK(cudaEventRecord(e->ev1, s));
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(e->h_pinned + 8, e->cd.abort_flag, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (e->h_pinned[8] == ABORT_BAD_ID)
    return fail("%s: index out of range: a phoneme id lies outside [0,%d) or a prompt / forced token outside [0,%d) "
                "(the reference's nn.Embedding raises IndexError here)", who, e->cfg.phoneme_vocab, V);
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  if (admit) {
    e->st.prefill_ms += ms;
    e->st.prefill_rows += T;
    for (int b = 0; b < B; ++b) e->B = std::max(e->B, new_slots[b] + 1);
  } else {
    e->st.prefill_ms = ms;
    e->st.decode_ms = 0.0;
    e->st.decode_steps = 0;
    e->st.decode_kv_positions = 0;
    e->st.decode_tokens = 0;
    e->st.prefill_rows = T;
    e->B = B;
    e->session = true;
  }
  return 0;
}

extern "C" int t2s_prefill(t2s_engine* e, const t2s_request* rq, void* stream_) {
  if (!e || !rq) return fail("t2s_prefill: null argument");
  e->session = false;
  return prefill_impl(e, rq, (cudaStream_t)stream_, false);
}

extern "C" int t2s_admit(t2s_engine* e, const t2s_request* rq, void* stream_) {
  if (!e || !rq) return fail("t2s_admit: null argument");
  return prefill_impl(e, rq, (cudaStream_t)stream_, true);
}

extern "C" int t2s_release_slots(t2s_engine* e, const int32_t* slots, int32_t n, void* stream_) {
  if (!e || (n > 0 && !slots)) return fail("t2s_release_slots: null argument");
  if (!e->session) return fail("t2s_release_slots: no resident session");
  cudaStream_t s = (cudaStream_t)stream_;
  std::vector<int> done(e->B);
  CK(cudaMemcpyAsync(done.data(), e->cd.done, (size_t)e->B * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int i = 0; i < n; ++i) {
    const int sl = slots[i];
    if (sl < 0 || sl >= e->B) return fail("t2s_release_slots: slot %d is not in use (0..%d)", sl, e->B - 1);
    if (!done[sl]) return fail("t2s_release_slots: slot %d is still decoding", sl);
    if (e->slot_free[sl]) return fail("t2s_release_slots: slot %d was already released", sl);
  }
  for (int i = 0; i < n; ++i) e->slot_free[slots[i]] = 1;
  return 0;
}

extern "C" int t2s_set_utterance_ids(t2s_engine* e, const int32_t* ids, int32_t n) {
  if (!e || (n > 0 && !ids)) return fail("t2s_set_utterance_ids: null argument");
  e->pending_uids.assign(ids, ids + std::max(n, 0));
  return 0;
}

// ---- decode ----------------------------------------------------------------------------------------------
static int read_state(t2s_engine* e, cudaStream_t s, int* n_active, int* step, int* aborted) {
  CK(cudaMemcpyAsync(e->h_pinned, e->cd.n_rows, 16, cudaMemcpyDeviceToHost, s));  // n_rows, n_active, step, abort
  CK(cudaStreamSynchronize(s));
  *n_active = e->h_pinned[1]; *step = e->h_pinned[2]; *aborted = e->h_pinned[3];
  return 0;
}

extern "C" int t2s_decode(t2s_engine* e, int32_t max_new_steps, void* stream_, int32_t* steps_run) {
  if (!e) return fail("t2s_decode: null engine");
  if (!e->session) return fail("t2s_decode: no session (call t2s_prefill first)");
  cudaStream_t s = (cudaStream_t)stream_;
  int n_active = 0, step0 = 0, aborted = 0;
  if (read_state(e, s, &n_active, &step0, &aborted)) return 1;
  int budget = max_new_steps < 0 ? e->max_steps : max_new_steps;
  std::unique_lock<std::mutex> decode_lock(g_decode_mu);  // held until the kernel has finished (read_state below synchronises)
  CK(cudaEventRecord(e->ev0, s));
  if (n_active > 0 && budget > 0) {
    int mode = e->decode_mode;
    if (mode == 5) {  // auto: the cluster-stream kernel whenever the batch fits its 16-CTA clusters, else the grid-wide phases
      const bool fits = e->max_clusters >= 1 && n_active <= e->max_clusters * cs::RMAX && e->cd.max_pages <= 32;
      mode = fits ? 4 : 1;
    }
    if (mode == 1 && e->tc_ok && e->tc_decode_min_batch > 0 && e->B >= e->tc_decode_min_batch) mode = 3;
    if (mode == 3 && !e->tc_ok) return fail("t2s_decode: tcgen05 decode needs the TMA descriptor entry point");
    if (mode == 6) {
      if (!e->wide_ok) return fail("t2s_decode: wide decode unavailable (needs %d co-resident CTAs with %zu bytes of shared memory)", ws::G, sizeof(ws::Smem));
      if (n_active > ws::RW) return fail("t2s_decode: wide decode holds at most %d sequences (got %d)", ws::RW, n_active);
      if (e->cd.max_pages > 32) return fail("t2s_decode: wide decode supports at most 32 KV pages per sequence");
      if (e->cfg.n_layer + 1 >= (int)ws::TAG_STRIDE) return fail("t2s_decode: wide decode supports at most %d layers", (int)ws::TAG_STRIDE - 2);
    }
    if (mode == 4) {
      if (e->max_clusters < 1) return fail("t2s_decode: cluster-stream decode unavailable (no co-resident 16-CTA cluster)");
      if (n_active > e->max_clusters * cs::RMAX) return fail("t2s_decode: cluster-stream decode holds at most %d sequences (got %d)", e->max_clusters * cs::RMAX, n_active);
      if (e->cd.max_pages > 32) return fail("t2s_decode: cluster-stream decode supports at most 32 KV pages per sequence");
    }
    e->st.decode_mode = mode;
    if (mode == 4) {
      const int ncl = e->max_clusters;  // always the full set: clusters without a sequence pull the weight stream into L2 for the others
      CK(cudaMemsetAsync(e->cd.bar, 0, 8, s));  // grid-barrier counter + the progress word of the prefetching clusters
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(ncl * cs::C); lc.blockDim = dim3(cs::NTC); lc.dynamicSmemBytes = sizeof(cs::Smem); lc.stream = s;
      // Cluster launch.  Every CTA is co-resident (grid <= cudaOccupancyMaxActiveClusters, one request at a time per engine),
      // which is what the per-step grid barrier needs; the cooperative attribute is not combined with cluster dimensions
      // (profilers reject that launch).
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs::C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      CK(cudaLaunchKernelEx(&lc, cs::k_decode_cluster, e->cd, (const unsigned char*)e->wstream.p, (const unsigned char*)e->hstream.p, budget));
      e->launches++;
    } else if (mode == 6) {
      CK(cudaMemsetAsync(e->cd.bar, 0, 8, s));
      Ctx c = e->cd;
      const unsigned char* ww = e->wwide.as<unsigned char>();
      unsigned long long* llp = e->llbuf.as<unsigned long long>();
      int steps = budget;
      void* args[] = {&c, &ww, &llp, &steps};
      CK(cudaLaunchCooperativeKernel((const void*)ws::k_decode_wide, dim3(ws::G), dim3(ws::NTW), args, sizeof(ws::Smem), s));
      e->launches++;
    } else if (mode == 1) {
      int grid = e->cd.attn_ctas;
      int per_sm = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode_persistent, NT, SMEM_MAX));
      if (per_sm < 1) return fail("t2s_decode: persistent kernel does not fit on an SM");
      if (grid > per_sm * e->num_sms) return fail("t2s_decode: %d CTAs cannot be co-resident", grid);
      CK(cudaMemsetAsync(e->cd.bar, 0, 4, s));
      Ctx c = e->cd;
      int steps = budget;
      void* args[] = {&c, &steps};
      CK(cudaLaunchCooperativeKernel((const void*)k_decode_persistent, dim3(grid), dim3(NT), args, SMEM_MAX, s));
      e->launches++;
    } else {
      // one graph = one decode step (122 kernel nodes); re-captured only when the context changes
      if (mode != 2 && (!e->graph_exec || e->graph_mode != mode || memcmp(&e->graph_ctx, &e->cd, sizeof(Ctx)) != 0)) {
        if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
        cudaStream_t cs;
        CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        const long long before = e->launches;
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        bool cap_ok = true;
        if (mode == 3) cap_ok = launch_decode_step_tc(e, e->cd, cs);
        else launch_decode_step(e, e->cd, cs);
        cudaGraph_t graph;
        cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        e->nodes_per_step = (int)(e->launches - before);
        e->launches = before;
        if (ce != cudaSuccess || !cap_ok) { cudaStreamDestroy(cs); return fail("graph capture failed: %s", cap_ok ? cudaGetErrorString(ce) : "tensor map encode"); }
        ce = cudaGraphInstantiate(&e->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        if (ce != cudaSuccess) return fail("graph instantiate failed: %s", cudaGetErrorString(ce));
        e->graph_ctx = e->cd;
        e->graph_mode = mode;
      }
      int done_steps = 0;
      while (done_steps < budget && n_active > 0) {
        const int chunk = std::min(e->check_steps, budget - done_steps);
        if (mode == 2) {  // plain stream launches, no graph (profiling aid)
          for (int i = 0; i < chunk; ++i) launch_decode_step(e, e->cd, s);
        } else {
          for (int i = 0; i < chunk; ++i) CK(cudaGraphLaunch(e->graph_exec, s));
          e->launches += (long long)chunk * e->nodes_per_step;
        }
        done_steps += chunk;
        CK(cudaMemcpyAsync(e->h_pinned, e->cd.n_rows, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        n_active = e->h_pinned[1];
      }
    }
  }
  CK(cudaEventRecord(e->ev1, s));
  CK(cudaGetLastError());
  int step1 = 0;
  if (read_state(e, s, &n_active, &step1, &aborted)) return 1;
  decode_lock.unlock();
  if (aborted == ABORT_BAD_ID) return fail("t2s_decode: index out of range: a forced token lies outside [0,%d)", V);
  if (aborted) return fail("t2s_decode: grid barrier watchdog fired (a CTA never arrived); results are invalid");
  unsigned long long stats1[3];
  CK(cudaMemcpyAsync(stats1, e->cd.stats, 24, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  // Session-cumulative: plan() accounts a decode step (its attended KV positions and active sequences)
  // when it schedules it, starting with the plan that follows the prefill's step-0 sample.
  e->st.decode_ms += ms;
  e->st.decode_steps = step1 - 1;
  e->st.decode_kv_positions = (int64_t)stats1[0];
  e->st.decode_tokens = (int64_t)stats1[2];
  if (steps_run) *steps_run = step1 - step0;
  return 0;
}

// ---- results ---------------------------------------------------------------------------------------------
extern "C" int t2s_result(t2s_engine* e, int64_t* tokens_out, int64_t row_stride, int32_t tokens_on_host,
                          int32_t* idx_out, void* stream_) {
  if (!e || !tokens_out || !idx_out) return fail("t2s_result: null argument");
  if (!e->session) return fail("t2s_result: no session");
  cudaStream_t s = (cudaStream_t)stream_;
  const int width = e->maxP + e->max_steps;
  if (row_stride < width) return fail("t2s_result: row_stride %lld < P + max_steps = %d", (long long)row_stride, width);
  if (e->out_idx.ensure((size_t)e->B * 4)) return 1;
  long long* dst = reinterpret_cast<long long*>(tokens_out);
  long long stride = row_stride;
  if (tokens_on_host) {
    if (e->out_tokens.ensure((size_t)e->B * width * 8)) return 1;
    dst = e->out_tokens.as<long long>();
    stride = width;
  }
  k_finalize<<<e->B, 256, 0, s>>>(e->cd, dst, stride, e->out_idx.as<int>());
  e->launches++;
  CK(cudaGetLastError());
  if (tokens_on_host)
    CK(cudaMemcpy2DAsync(tokens_out, (size_t)row_stride * 8, dst, (size_t)width * 8, (size_t)width * 8, e->B,
                         cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(idx_out, e->out_idx.p, (size_t)e->B * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int t2s_codes_to_latent(t2s_engine* e, const int64_t* codes, int32_t n, const float* codebook,
                                   int32_t codebook_size, int32_t dim, int32_t upsample, float* out, void* stream_) {
  if (!e) return fail("t2s_codes_to_latent: null engine");
  if (n < 0 || codebook_size < 1 || dim < 1 || upsample < 1 || upsample > 8)
    return fail("t2s_codes_to_latent: n=%d codebook_size=%d dim=%d upsample=%d out of range", n, codebook_size, dim, upsample);
  if (n == 0) return 0;  // an empty utterance: nothing to write (out may be a zero-sized allocation)
  if (!codes || !codebook || !out) return fail("t2s_codes_to_latent: null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  if (e->latent_flag.ensure(8)) return 1;
  CK(cudaMemsetAsync(e->latent_flag.p, 0, 4, s));
  dim3 grid((n + 31) / 32, (dim + 31) / 32), block(32, 8);
  k_codes_to_latent<<<grid, block, 0, s>>>(reinterpret_cast<const long long*>(codes), n, codebook, codebook_size, dim, upsample,
                                          out, e->latent_flag.as<int>());
  e->launches++;
  CK(cudaGetLastError());
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, e->latent_flag.p, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (bad) return fail("t2s_codes_to_latent: a code lies outside [0, %d) (index out of range in the codebook lookup)", codebook_size);
  return 0;
}

extern "C" int t2s_generate(t2s_engine* e, const t2s_request* rq, int64_t* tokens_out, int64_t row_stride,
                            int32_t tokens_on_host, int32_t* idx_out, void* stream) {
  if (!e || !rq) return fail("t2s_generate: null argument");
  // Auto mode: a batch that does not fit the cluster-stream kernel (8 sequences per co-resident cluster) is run as equal
  // chunks that do, one after the other (sequences never interact: t2s_model.py:583-779 has no cross-sequence op).  Measured:
  // 2 x 32 sequences take 2 x 403 us per step against 1001 us for 64 on the grid-wide phase kernels, 5 x 52 take 5 x 449 us
  // against 2863 us for 256 with the tcgen05 projections of mode 3 (which the explicit modes and the test hooks still use).
  const int fit = e->max_clusters * cs::RMAX;
  const bool hooks = e->forced || e->logits_rec || e->timeline;
  const bool whole = e->decode_mode != 5 || fit < 1 || rq->batch <= fit || hooks;
  if (whole) {
    if (t2s_prefill(e, rq, stream)) return 1;
    int32_t n = 0;
    if (t2s_decode(e, -1, stream, &n)) return 1;
    return t2s_result(e, tokens_out, row_stride, tokens_on_host, idx_out, stream);
  }
  if (rq->batch > e->cfg.max_batch) return fail("t2s_generate: batch %d outside [1,%d]", rq->batch, e->cfg.max_batch);
  const int B = rq->batch, nchunk = (B + fit - 1) / fit, per = (B + nchunk - 1) / nchunk;
  t2s_stats acc = e->st;
  acc.prefill_ms = 0; acc.decode_ms = 0; acc.decode_steps = 0; acc.decode_tokens = 0; acc.decode_kv_positions = 0; acc.prefill_rows = 0;
  int rc = 0;
  int64_t id_off = 0;
  // t2s_set_utterance_ids covers the whole request: every chunk gets its own slice
  const std::vector<int> all_uids = e->pending_uids;
  e->pending_uids.clear();
  if (!all_uids.empty() && (int)all_uids.size() != B)
    return fail("t2s_generate: t2s_set_utterance_ids gave %zu ids for a batch of %d", all_uids.size(), B);
  for (int b0 = 0; b0 < B && !rc; b0 += per) {
    const int n = std::min(per, B - b0);
    if (!all_uids.empty()) e->pending_uids.assign(all_uids.begin() + b0, all_uids.begin() + b0 + n);
    t2s_request sub = *rq;
    sub.batch = n;
    sub.phoneme_ids = rq->phoneme_ids + id_off;
    sub.phoneme_lens = rq->phoneme_lens + b0;
    sub.bert = rq->bert + b0;
    sub.bert_stride_c = rq->bert_stride_c + b0;
    sub.bert_stride_t = rq->bert_stride_t + b0;
    sub.prompt = rq->prompt ? rq->prompt + (int64_t)b0 * rq->prompt_row_stride : nullptr;
    for (int b = 0; b < n; ++b) id_off += rq->phoneme_lens[b0 + b];
    e->slot_base = b0;
    int32_t steps = 0;
    rc = t2s_prefill(e, &sub, stream) || t2s_decode(e, -1, stream, &steps) ||
         t2s_result(e, tokens_out + (int64_t)b0 * row_stride, row_stride, tokens_on_host, idx_out + b0, stream);
    acc.prefill_ms += e->st.prefill_ms; acc.decode_ms += e->st.decode_ms; acc.decode_steps += e->st.decode_steps;
    acc.decode_tokens += e->st.decode_tokens; acc.decode_kv_positions += e->st.decode_kv_positions; acc.prefill_rows += e->st.prefill_rows;
    acc.decode_mode = e->st.decode_mode;
  }
  e->slot_base = 0;
  if (!rc) e->st = acc;
  return rc;
}

extern "C" int t2s_get_sampled(t2s_engine* e, int32_t* out, int32_t n_steps, void* stream_) {
  if (!e || !out) return fail("t2s_get_sampled: null argument");
  if (!e->session) return fail("t2s_get_sampled: no session");
  cudaStream_t s = (cudaStream_t)stream_;
  const int n = std::min(n_steps, e->max_steps);
  CK(cudaMemcpy2DAsync(out, (size_t)n_steps * 4, e->sampled.p, (size_t)e->max_steps * 4, (size_t)n * 4, e->B,
                       cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int t2s_sampler_test(t2s_engine* e, const float* logits, int32_t n, int32_t width, const int32_t* prev,
                                int32_t m, int32_t top_k, float top_p, float temperature, float repetition_penalty,
                                uint64_t seed, int32_t step, int32_t* tok_out, int32_t* greedy_out, void* stream_) {
  if (!e || !logits || !tok_out || !greedy_out) return fail("t2s_sampler_test: null argument");
  if (n < 1 || n > MAX_B) return fail("t2s_sampler_test: n out of range");
  if (width != V && width != V - 1) return fail("t2s_sampler_test: width must be 1024 or 1025");
  if (top_k < 1 || step < 0) return fail("t2s_sampler_test: bad top_k / step");
  cudaStream_t s = (cudaStream_t)stream_;
  e->session = false;  // clobbers the session state
  const int ms = step + 2;
  if (e->ints2.ensure((size_t)(MAX_B * 7 + 16) * 4) || e->gen.ensure((size_t)n * ms * 4 * 3)) return 1;
  std::vector<float> lg((size_t)n * VPAD, 0.f);
  for (int r = 0; r < n; ++r) memcpy(&lg[(size_t)r * VPAD], logits + (size_t)r * V, V * 4);
  std::vector<uint32_t> seen((size_t)n * SEEN_WORDS, 0u);
  for (int r = 0; r < n; ++r)
    for (int j = 0; j < m; ++j) {
      const int t = prev ? prev[(size_t)r * m + j] : -1;
      if (t >= 0 && t < V) seen[(size_t)r * SEEN_WORDS + (t >> 5)] |= 1u << (t & 31);
    }
  std::vector<int> i2(16 + 7 * MAX_B, 0);  // ... + slot_step0 (0) + slot_P (0) + slot_uid (0: the test's Philox stream is utterance 0)
  i2[0] = n; i2[1] = n; i2[2] = step;
  for (int r = 0; r < n; ++r) i2[16 + MAX_B + r] = r;  // active = identity
  CK(cudaMemcpyAsync(e->logits.p, lg.data(), lg.size() * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->seen.p, seen.data(), seen.size() * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->ints2.p, i2.data(), i2.size() * 4, cudaMemcpyHostToDevice, s));
  Ctx c{};
  int* d2 = e->ints2.as<int>();
  c.n_rows = d2; c.n_active = d2 + 1; c.step = d2 + 2; c.abort_flag = d2 + 3;
  c.seq_len = d2 + 16; c.active = d2 + 16 + MAX_B; c.done = d2 + 16 + 2 * MAX_B; c.out_idx = d2 + 16 + 3 * MAX_B;
  c.slot_step0 = d2 + 16 + 4 * MAX_B; c.slot_P = d2 + 16 + 5 * MAX_B; c.slot_uid = d2 + 16 + 6 * MAX_B;
  c.logits = e->logits.as<float>(); c.seen = e->seen.as<uint32_t>();
  c.gen = e->gen.as<int>(); c.sampled = c.gen + (size_t)n * ms; c.greedy_rec = c.gen + (size_t)2 * n * ms;
  c.B0 = n; c.P = 0; c.max_steps = ms; c.eos_window = (width == V) ? 0 : step + 1; c.early_stop = -1; c.top_k = top_k;
  c.top_p = top_p; c.temperature = temperature; c.rep_pen = repetition_penalty;
  c.seed_lo = (uint32_t)(seed & 0xFFFFFFFFull); c.seed_hi = (uint32_t)(seed >> 32);
  c.emb_audio = e->emb_audio.as<bf16>(); c.pe = e->pe.as<float>(); c.pe_len = e->cfg.pe_len; c.alpha_audio = 0.f;
  c.x0 = e->x0_slots.as<float>(); c.x0b = e->x0b_slots.as<bf16>();
  launch_phase<PH_SAMPLE>(e, c, 0, std::min(n, e->num_sms), s);
  CK(cudaGetLastError());
  CK(cudaMemcpy2DAsync(tok_out, 4, c.sampled + step, (size_t)ms * 4, 4, n, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpy2DAsync(greedy_out, 4, c.greedy_rec + step, (size_t)ms * 4, 4, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int t2s_set_timeline(t2s_engine* e, long long* buf, int32_t step, int32_t slots) {
  if (!e) return fail("null engine");
  e->timeline = buf; e->tl_step = step; e->tl_slots = buf ? slots : 0;
  return 0;
}

extern "C" int t2s_bench_barrier(t2s_engine* e, int32_t n_barriers, int32_t n_ctas, float* ms_out, void* stream_) {
  if (!e || !ms_out || n_barriers < 1) return fail("t2s_bench_barrier: bad argument");
  cudaStream_t s = (cudaStream_t)stream_;
  if (e->ints2.ensure((size_t)(MAX_B * 7 + 16) * 4)) return 1;
  int* i2 = e->ints2.as<int>();
  CK(cudaMemsetAsync(i2, 0, 64, s));
  unsigned* bar = reinterpret_cast<unsigned*>(i2 + 4);
  int* ab = i2 + 3;
  int grid = n_ctas > 0 ? n_ctas : e->num_sms;
  void* args[] = {&bar, &ab, &n_barriers};
  CK(cudaEventRecord(e->ev0, s));
  CK(cudaLaunchCooperativeKernel((const void*)k_barrier_bench, dim3(grid), dim3(NT), args, 0, s));
  CK(cudaEventRecord(e->ev1, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(ms_out, e->ev0, e->ev1));
  e->launches++;
  return 0;
}

extern "C" int t2s_get_stats(t2s_engine* e, t2s_stats* out) {
  if (!e || !out) return fail("t2s_get_stats: null argument");
  e->st.kernel_launches = e->launches;
  *out = e->st;
  return 0;
}
