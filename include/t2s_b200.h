/* t2s_b200.h — C ABI of libt2s_b200.so: B200 (sm_100a) text-to-semantic autoregressive decode.
 *
 * This is the drop-in boundary for the hot path of GPT-SoVITS
 * (reference: GPT_SoVITS/AR/models/t2s_model.py).  Plain pointers and sizes only; no torch types.
 * Each entry point names the reference interface it replaces.  The Python binding a maintainer
 * would add is gpt-sovits_b200/_lib.py (ctypes); see INTEGRATION.md.
 *
 * Threading: one request at a time per engine (the reference serves one request at a time:
 * api_v2.py:496 uvicorn workers=1).  All calls return 0 on success, non-zero on error;
 * t2s_last_error() gives the message of the calling thread's last failure.
 */
#ifndef T2S_B200_H
#define T2S_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t2s_engine t2s_engine;

/* element types of caller-provided tensors */
enum { T2S_F32 = 0, T2S_F16 = 1, T2S_BF16 = 2 };

/* Model hyper-parameters = config["model"] of Text2SemanticDecoder.__init__ (t2s_model.py:260-273).
 * This build is specialised for the s1 architecture family d_model=512, n_head=16 (head_dim 32),
 * d_ff=2048, vocab=1025 (EOS=1024), bert_dim=1024; n_layer and phoneme_vocab are free. */
typedef struct {
  int32_t n_layer;        /* 24 for s1-v2 */
  int32_t d_model;        /* 512 */
  int32_t n_head;         /* 16 */
  int32_t d_ff;           /* 2048 (= 4*hidden, t2s_model.py:304) */
  int32_t vocab;          /* 1025 */
  int32_t phoneme_vocab;  /* 732 for v2 */
  int32_t bert_dim;       /* 1024 */
  int32_t eos;            /* 1024 */
  int32_t pe_len;         /* 4000 (embedding.py:52) */
  int32_t max_batch;      /* upper bound on utterances per call (<= 256) */
} t2s_model_config;

/* Tensor ids = the reference state_dict keys (SURVEY.md section 8b "Weight source"). */
enum {
  T2S_W_BERT_PROJ_W = 0,   /* bert_proj.weight [512,1024] */
  T2S_W_BERT_PROJ_B,       /* bert_proj.bias [512] */
  T2S_W_TEXT_EMB,          /* ar_text_embedding.word_embeddings.weight [phoneme_vocab,512] */
  T2S_W_TEXT_ALPHA,        /* ar_text_position.alpha [1] */
  T2S_W_AUDIO_EMB,         /* ar_audio_embedding.word_embeddings.weight [1025,512] */
  T2S_W_AUDIO_ALPHA,       /* ar_audio_position.alpha [1] */
  T2S_W_PE,                /* SinePositionalEmbedding.pe [pe_len,512] (embedding.py:54-72) */
  T2S_W_PREDICT,           /* ar_predict_layer.weight [1025,512] */
  T2S_W_IN_PROJ_W,         /* h.layers.{i}.self_attn.in_proj_weight [1536,512] */
  T2S_W_IN_PROJ_B,         /* ...in_proj_bias [1536] */
  T2S_W_OUT_PROJ_W,        /* ...self_attn.out_proj.weight [512,512] */
  T2S_W_OUT_PROJ_B,        /* ...out_proj.bias [512] */
  T2S_W_LIN1_W,            /* ...linear1.weight [2048,512] */
  T2S_W_LIN1_B,            /* ...linear1.bias [2048] */
  T2S_W_LIN2_W,            /* ...linear2.weight [512,2048] */
  T2S_W_LIN2_B,            /* ...linear2.bias [512] */
  T2S_W_NORM1_W, T2S_W_NORM1_B, /* ...norm1.{weight,bias} [512] */
  T2S_W_NORM2_W, T2S_W_NORM2_B, /* ...norm2.{weight,bias} [512] */
  T2S_W_COUNT
};

/* Replaces Text2SemanticDecoder.__init__ (t2s_model.py:260-353): allocates the packed bf16 weight
 * arena on the current CUDA device.  Fails (no CPU fallback) when no sm_100 device is present. */
int t2s_create(const t2s_model_config* cfg, t2s_engine** out);
void t2s_destroy(t2s_engine* e);
const char* t2s_last_error(void);

/* Replaces load_state_dict + .to(device) + .half() (TTS.py:585-599): copies one parameter tensor
 * (contiguous, row-major, `dtype`, on host or device) into the engine's packed layout (matrices are
 * rounded to bf16 and re-tiled for the tensor-core fragments; vectors are kept fp32).
 * `layer` is ignored for non-layer tensors. */
int t2s_load_tensor(t2s_engine* e, int32_t tensor_id, int32_t layer, const void* data, int32_t dtype,
                    int64_t numel, int32_t on_device, void* stream);

/* One call of infer_panel / infer_panel_batch_infer (t2s_model.py:583-595 / :814-826 signature;
 * argument meaning is the reference's).  Pointers are device pointers unless inputs_on_host. */
typedef struct {
  int32_t batch;                 /* number of utterances B (len(x)) */
  const int64_t* phoneme_ids;    /* all_phoneme_ids concatenated [sum L_i] */
  const int32_t* phoneme_lens;   /* HOST array [B]: L_i (all_phoneme_lens) */
  const void* const* bert;       /* HOST array [B] of pointers: all_bert_features[i], element (c, t)
                                    at bert[i] + c*bert_stride_c[i] + t*bert_stride_t[i] */
  const int64_t* bert_stride_c;  /* HOST [B], in elements */
  const int64_t* bert_stride_t;  /* HOST [B], in elements */
  int32_t bert_dtype;            /* T2S_F32 / T2S_F16 / T2S_BF16 */
  const int64_t* prompt;         /* prompts [B,P] (NULL when prompt_len == 0: reference-free) */
  int64_t prompt_row_stride;     /* elements between rows; 0 for TTS.run's .expand view (TTS.py:1210) */
  int32_t prompt_len;            /* P */
  int32_t top_k;                 /* >= 1 */
  float top_p;                   /* >= 1.0 disables (utils.py:169) */
  float temperature;             /* clamped to >= 1e-5 (utils.py:181) */
  float repetition_penalty;      /* 1.0 disables (utils.py:159) */
  int32_t early_stop_num;        /* -1 disables (t2s_model.py:747/:897) */
  int32_t eos_suppress_steps;    /* EOS column removed while idx < this: 11 = infer_panel_naive (:888),
                                    1 = infer_panel_batch_infer (:708-710) */
  int32_t max_steps;             /* 1500 in the reference (:701/:878) */
  uint64_t seed;                 /* Philox key for the exponential-race sampler */
  int32_t inputs_on_host;        /* phoneme_ids / bert[i] / prompt are HOST pointers (copied H2D here) */
} t2s_request;

/* Embedding + prefix-LM prefill (t2s_model.py:611-683 + process_prompt :703/:880) and the first
 * sampled token (idx 0).  Leaves the session resident in the engine. */
int t2s_prefill(t2s_engine* e, const t2s_request* req, void* stream);

/* Continuous batching (the reference has no counterpart: TTS.run hands whole batches to infer_panel,
 * TTS.py:1215): adds the request's utterances to the RESIDENT session at its current step - their prompt rows are prefilled
 * and their step 0 is sampled here, the following t2s_decode calls advance old and new sequences together, each with its own
 * idx / EOS window / early stop / Philox stream (sequences never interact, so an admitted utterance gets exactly the tokens it
 * would get in a request of its own).  Slots and K/V pages must have been reserved before t2s_prefill
 * (T2S_OPT_SESSION_SLOTS, T2S_OPT_SESSION_POSITIONS); sampling parameters and stop rules are per session and must match;
 * the prompt length may differ.  Device inputs only; the prompt rows must stay valid for the session.  New slots follow the
 * existing ones in t2s_result's order. */
int t2s_admit(t2s_engine* e, const t2s_request* req, void* stream);
/* Returns the slots of FINISHED utterances (idx >= 0 in t2s_result; fetch their tokens first) to the session: the next
 * t2s_admit reuses them - and their K/V pages - before it takes fresh slots, so a resident session serves an unbounded
 * stream of utterances with a fixed number of slots.  A slot that is still decoding is an error. */
int t2s_release_slots(t2s_engine* e, const int32_t* slots, int32_t n, void* stream);
/* The utterances of the NEXT t2s_prefill / t2s_admit get these ids (host array, one per utterance) as the key of their Philox
 * streams instead of their slot index: sampling then does not depend on which slot, chunk, session or GPU an utterance lands in
 * (gpt-sovits_b200/shard.py passes the position in the caller's batch, StreamingSession the submission order). */
int t2s_set_utterance_ids(t2s_engine* e, const int32_t* ids, int32_t n);

/* The decode loop (t2s_model.py:701-769 / :878-914) for at most max_new_steps further steps
 * (-1: until every sequence has stopped).  Synchronises `stream`.  steps_run = steps executed. */
int t2s_decode(t2s_engine* e, int32_t max_new_steps, void* stream, int32_t* steps_run);

/* Results in ORIGINAL batch order (t2s_model.py:699,735,779): tokens_out[b, 0:P+idx_b] =
 * prompt ++ kept tokens (the token sampled at the stopping step is dropped, :733/:918),
 * idx_out[b] = idx_b (number of kept tokens).  tokens_out is [B, row_stride] int64 on device (or on
 * host when tokens_on_host), row_stride >= P + max_steps; idx_out is a HOST array [B].
 * unfinished sequences (decode stopped early by max_new_steps) report idx = -1. */
int t2s_result(t2s_engine* e, int64_t* tokens_out, int64_t row_stride, int32_t tokens_on_host,
               int32_t* idx_out, void* stream);

/* prefill + decode(-1) + result in one call: the form bench.py's end-to-end leg uses with host buffers. */
int t2s_generate(t2s_engine* e, const t2s_request* req, int64_t* tokens_out, int64_t row_stride,
                 int32_t tokens_on_host, int32_t* idx_out, void* stream);

/* The first op after the path: SynthesizerTrn.decode's `quantized = self.quantizer.decode(codes)`
 * followed by `F.interpolate(quantized, size=2*T, mode="nearest")` for the 25 Hz models (module/models.py:989-991;
 * ResidualVectorQuantization.decode core_vq.py:359-365 with n_q = 1, VectorQuantization.decode :286-290,
 * EuclideanCodebook.dequantize = F.embedding).  codes: n int64 semantic tokens ON DEVICE (e.g. a row of t2s_result's
 * tokens_out, offset to its last idx_b entries as TTS.py slices `item[-idx:]`); codebook: [codebook_size, dim] fp32 on device
 * (`quantizer.vq.layers.0._codebook.embed`); out: [dim, upsample*n] fp32 on device, out[d][upsample*t + j] =
 * codebook[codes[t]][d].  A code outside [0, codebook_size) is an error (the reference's embedding lookup raises). */
int t2s_codes_to_latent(t2s_engine* e, const int64_t* codes, int32_t n, const float* codebook, int32_t codebook_size,
                        int32_t dim, int32_t upsample, float* out, void* stream);

/* ---- test / measurement hooks (not part of the reference surface) ------------------------- */

/* Teacher forcing: step s of slot b emits forced[b*n_steps + s] instead of the sampled token (the
 * sampled one is still recorded).  Device pointer, must stay valid for the session; NULL clears. */
int t2s_set_forced_tokens(t2s_engine* e, const int32_t* forced, int32_t n_steps);
/* Records the raw (pre-penalty) fp32 logits of steps < n_steps into buf[s][b][1025] (device). */
int t2s_set_logits_capture(t2s_engine* e, float* buf, int32_t n_steps);
/* Copies the raw sampled tokens [B, max_steps] int32 (before forcing) to a HOST buffer. */
int t2s_get_sampled(t2s_engine* e, int32_t* out, int32_t n_steps, void* stream);

/* Runs ONLY the fused sampling kernel on caller-provided logits (host, [n][1025] fp32; columns >= width
 * ignored) with previous tokens prev (host, [n][m] int32, -1 = padding) at decode step `step`:
 * tok_out[n] = sampled token, greedy_out[n] = argmax of the penalised logits.  Used by the tests to pin
 * the kernel against the reference's logits_to_probs known-answer rows (utils.py:147-199). */
int t2s_sampler_test(t2s_engine* e, const float* logits, int32_t n, int32_t width, const int32_t* prev, int32_t m,
                     int32_t top_k, float top_p, float temperature, float repetition_penalty, uint64_t seed,
                     int32_t step, int32_t* tok_out, int32_t* greedy_out, void* stream);

/* Measurement hook: during persistent-kernel iteration `step` of the next t2s_decode, thread 0 of every CTA
 * records its SM clock when it arrives at / is released from each grid barrier (mode 1: buf[cta][slots][2]) or at
 * the phase markers of the cluster-stream kernel (mode 4: buf[cta][2*slots] consecutive stamps), device int64.
 * Set before t2s_prefill; NULL clears. */
int t2s_set_timeline(t2s_engine* e, long long* buf, int32_t step, int32_t slots);

/* Measurement hook: latency of n_barriers back-to-back grid barriers of the persistent kernel. */
int t2s_bench_barrier(t2s_engine* e, int32_t n_barriers, int32_t n_ctas, float* ms_out, void* stream);

enum {
  T2S_OPT_DECODE_MODE = 0,   /* 5 (default): auto = 4 when the batch fits the cluster-stream kernel (<= 8 sequences per
                                co-resident 16-CTA cluster: 56 on a B200); t2s_generate runs larger batches
                                as equal chunks that fit, one after the other; with test hooks set: else 1;
                                4: cluster-stream kernel: thread-block clusters own sequences end to end, no grid barrier
                                   inside a step (all remaining steps of the request in ONE launch);
                                0: one kernel per phase over all SMs, CUDA-graph replay; 1: the same phases in one
                                persistent cooperative kernel with grid barriers; 2: plain stream launches (profiling aid);
                                3: CUDA graph with the projections on tcgen05 tensor cores (large batches);
                                mode 1 switches to 3 by itself when batch >= T2S_OPT_TC_DECODE_MIN_BATCH;
                                6: "wide" small-batch kernel (<= 8 sequences): every layer spread over 144 SMs, weights
                                   through a TMA shared-memory ring, hand-offs through L2 {value, tag} cells; parity-green
                                   but measured slower than mode 4, so auto never picks it */
  T2S_OPT_PREFILL_GEMM = 1,  /* 0: warp-MMA row-tile projections; 1 (default): persistent tcgen05/TMEM + TMA GEMM; 2: one tile per CTA */
  T2S_OPT_NUM_CTAS = 2,      /* persistent grid size (0 = one CTA per SM) */
  T2S_OPT_CHECK_STEPS = 3,   /* graph mode: host checks the active count every this many steps */
  T2S_OPT_TC_DECODE_MIN_BATCH = 4, /* batch size from which decode projections run on tcgen05 (default 160, the measured crossover; 0: never) */
  T2S_OPT_SESSION_SLOTS = 5,     /* continuous batching: slots (utterances) the NEXT t2s_prefill reserves for the session, >= its own
                                    batch (0 = exactly its batch: no t2s_admit possible); the K/V pool is sized for all of them */
  T2S_OPT_SESSION_POSITIONS = 6, /* K/V positions reserved per slot (0 = what the first request's longest utterance needs:
                                    phonemes + prompt + step cap); an admitted utterance must fit */
  T2S_OPT_HOOKS_BY_UTTERANCE = 7 /* test hooks: 0 = forced tokens / captured logits are indexed by session slot; n > 0 = by utterance id
                                    (t2s_set_utterance_ids), n rows - a recycled slot serves several utterances */
};
int t2s_set_option(t2s_engine* e, int32_t option, int64_t value);

typedef struct {
  double prefill_ms;          /* device time of the last t2s_prefill (CUDA events on `stream`) */
  double decode_ms;           /* device time of t2s_decode, cumulative over the current session */
  int64_t decode_steps;       /* decode steps (after the prefill step) executed in the session */
  int64_t decode_tokens;      /* sequence-steps (sum of active sequences over those steps) */
  int64_t decode_kv_positions;/* sum over those sequence-steps of attended KV positions */
  int64_t kernel_launches;    /* kernels this library launched since t2s_create */
  int64_t prefill_rows;       /* prompt rows (text + audio) of the last prefill */
  int64_t weight_bytes_per_step; /* bf16 weight bytes one decode step streams */
  int64_t kv_bytes_per_position; /* bytes of K+V per cached position (all layers) */
  int32_t num_sms;
  int32_t decode_mode;
} t2s_stats;
int t2s_get_stats(t2s_engine* e, t2s_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* T2S_B200_H */
