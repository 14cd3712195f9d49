"""TEST INFRASTRUCTURE — CPU restatement (numpy fp32) of the reference text-to-semantic decode path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this; the product path (gpt-sovits_b200/) never does and fails loudly without its CUDA library.

What is restated (reference file:line, all under GPT_SoVITS/AR/):
  * input embedding            models/t2s_model.py:611-622, 636-641 / :827-829, 842-848
                               modules/embedding.py:36-78 (x*x_scale(=1) + alpha*pe[:T])
  * prefix-LM mask             models/t2s_model.py:644-683 / :860-876 (text rows see all text and no
                               audio; audio rows see all text + causal audio).  Left padding is inert
                                so sequences are kept ragged, not padded.
  * T2SBlock.process_prompt    models/t2s_model.py:135-174   (post-LN, eps 1e-5, ReLU FFN 512-2048-512)
  * T2SBlock.decode_next_token models/t2s_model.py:176-221   (KV append + single-query attention)
  * ar_predict_layer           models/t2s_model.py:313, :706/:884 (no bias, last position only)
  * the decode loops           infer_panel_naive :814-918, infer_panel_batch_infer :583-779
                               (EOS column dropped for idx<11 / idx==0; stop on sample==EOS or
                               argmax(penalised logits)==EOS, early_stop_num, 1500-step cap; the token
                               sampled at the stopping step is dropped; idx = number of kept tokens)
  * sampler                    see oracle/sampler_oracle.py

Parity pin: the reference has no tests or golden vectors for this path, so the
oracle is pinned against the reference ITSELF, run in the build container on identical
bf16-representable synthetic weights: tests/golden/*.npz (made by oracle/make_goldens.py) hold the
reference's per-step logits / tokens / (y, idx) outputs and tests/test_oracle.py checks this
restatement against them (and, where /root/reference exists, against the live reference).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import sampler_oracle as so

EOS_WINDOW_NAIVE = 11  # t2s_model.py:888  (idx < 11)
EOS_WINDOW_BATCH = 1  # t2s_model.py:708-710 (idx == 0)
MAX_STEPS = 1500  # t2s_model.py:701/:878


def _np(t) -> np.ndarray:
    if isinstance(t, np.ndarray):
        return t.astype(np.float32, copy=False)
    return t.detach().cpu().float().numpy()


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    mu = x.mean(axis=-1, keepdims=True, dtype=np.float32)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=np.float32)
    return (xc / np.sqrt(var + np.float32(eps)) * g + b).astype(np.float32)


class T2SOracle:
    """Holds fp32 copies of the parameters (names = reference state_dict keys) and runs the path."""

    def __init__(self, state_dict: Dict[str, "np.ndarray"], pe, n_head: int = 16, eos: int = 1024):
        sd = {k: _np(v) for k, v in state_dict.items()}
        self.sd = sd
        self.pe = _np(pe)  # [4000, 512]
        self.H = n_head
        self.EOS = eos
        self.d = sd["ar_predict_layer.weight"].shape[1]
        self.L = 0
        while f"h.layers.{self.L}.linear1.weight" in sd:
            self.L += 1
        self.dh = self.d // self.H
        self.alpha_t = np.float32(sd["ar_text_position.alpha"].reshape(-1)[0])
        self.alpha_a = np.float32(sd["ar_audio_position.alpha"].reshape(-1)[0])
        self.layers = []
        for i in range(self.L):
            p = f"h.layers.{i}."
            self.layers.append(dict(
                wqkv=sd[p + "self_attn.in_proj_weight"], bqkv=sd[p + "self_attn.in_proj_bias"],
                wo=sd[p + "self_attn.out_proj.weight"], bo=sd[p + "self_attn.out_proj.bias"],
                w1=sd[p + "linear1.weight"], b1=sd[p + "linear1.bias"],
                w2=sd[p + "linear2.weight"], b2=sd[p + "linear2.bias"],
                g1=sd[p + "norm1.weight"], be1=sd[p + "norm1.bias"],
                g2=sd[p + "norm2.weight"], be2=sd[p + "norm2.bias"],
            ))
        self.wp = sd["ar_predict_layer.weight"]

    # ---- input embedding -------------------------------------------------------------------
    def embed_text(self, ids: np.ndarray, bert: np.ndarray) -> np.ndarray:
        """ids [L] int, bert [1024, L] -> [L, d]  (t2s_model.py:614-616)."""
        x = self.sd["ar_text_embedding.word_embeddings.weight"][ids]
        x = x + (bert.T.astype(np.float32) @ self.sd["bert_proj.weight"].T + self.sd["bert_proj.bias"])
        return (x + self.alpha_t * self.pe[: x.shape[0]]).astype(np.float32)

    def embed_audio(self, tokens: np.ndarray, pos0: int = 0) -> np.ndarray:
        """tokens [n] -> [n, d] with PE rows pos0..pos0+n-1 (t2s_model.py:636-640, :766-769)."""
        y = self.sd["ar_audio_embedding.word_embeddings.weight"][tokens]
        return (y + self.alpha_a * self.pe[pos0 : pos0 + y.shape[0]]).astype(np.float32)

    # ---- transformer -----------------------------------------------------------------------
    def _ffn_ln(self, lw, x, attn):
        x = layer_norm(x + attn, lw["g1"], lw["be1"])
        h = np.maximum(x @ lw["w1"].T + lw["b1"], 0)
        return layer_norm(x + (h @ lw["w2"].T + lw["b2"]), lw["g2"], lw["be2"])

    def prefill_one(self, xy: np.ndarray, n_text: int, cap: int):
        """One sequence.  xy [S0, d]; rows < n_text are text.  Returns (last hidden [d], K, V) with
        K,V [L, H, cap, dh] filled on [:S0]."""
        S0 = xy.shape[0]
        H, dh = self.H, self.dh
        K = np.zeros((self.L, H, cap, dh), np.float32)
        V = np.zeros((self.L, H, cap, dh), np.float32)
        # masked[i, j] True = not visible (t2s_model.py:652-664)
        i = np.arange(S0)[:, None]
        j = np.arange(S0)[None, :]
        masked = np.where(i < n_text, j >= n_text, j > i)
        x = xy
        scale = np.float32(1.0 / np.sqrt(dh))
        for li, lw in enumerate(self.layers):
            qkv = x @ lw["wqkv"].T + lw["bqkv"]
            q, k, v = np.split(qkv, 3, axis=-1)
            q = q.reshape(S0, H, dh).transpose(1, 0, 2)
            k = k.reshape(S0, H, dh).transpose(1, 0, 2)
            v = v.reshape(S0, H, dh).transpose(1, 0, 2)
            K[li, :, :S0] = k
            V[li, :, :S0] = v
            s = (q @ k.transpose(0, 2, 1)) * scale
            s = np.where(masked[None], np.float32(-np.inf), s)
            s = s - s.max(axis=-1, keepdims=True)
            p = np.exp(s)
            p = p / p.sum(axis=-1, keepdims=True, dtype=np.float32)
            a = (p @ v).transpose(1, 0, 2).reshape(S0, self.d)
            attn = a @ lw["wo"].T + lw["bo"]
            x = self._ffn_ln(lw, x, attn.astype(np.float32))
        return x[-1], K, V

    def decode_step(self, x: np.ndarray, K: np.ndarray, V: np.ndarray, sel: np.ndarray, lens: np.ndarray):
        """Batched single-token step over the active slots `sel`.  x [n, d]; K,V [L, B, H, cap, dh];
        lens [n] = cached positions per row BEFORE this token.  Appends k,v at lens[r] and attends
        to [0, lens[r]]  (t2s_model.py:184-203; the cache is preallocated instead of torch.cat'ed)."""
        n = x.shape[0]
        H, dh = self.H, self.dh
        scale = np.float32(1.0 / np.sqrt(dh))
        smax = int(lens.max()) + 1
        vis = np.arange(smax)[None, :] <= lens[:, None]  # [n, smax]
        full = n == K.shape[1]
        for li, lw in enumerate(self.layers):
            qkv = x @ lw["wqkv"].T + lw["bqkv"]
            q, k, v = np.split(qkv, 3, axis=-1)
            K[li, sel, :, lens] = k.reshape(n, H, dh)
            V[li, sel, :, lens] = v.reshape(n, H, dh)
            Kl = K[li, :, :, :smax] if full else K[li, sel, :, :smax]
            Vl = V[li, :, :, :smax] if full else V[li, sel, :, :smax]
            qh = q.reshape(n, H, 1, dh)
            s = (qh @ Kl.transpose(0, 1, 3, 2))[:, :, 0, :] * scale  # [n,H,smax]
            s = np.where(vis[:, None, :], s, np.float32(-np.inf))
            s = s - s.max(axis=-1, keepdims=True)
            p = np.exp(s)
            p = p / p.sum(axis=-1, keepdims=True, dtype=np.float32)
            a = (p[:, :, None, :] @ Vl)[:, :, 0, :].reshape(n, self.d)
            attn = a @ lw["wo"].T + lw["bo"]
            x = self._ffn_ln(lw, x, attn.astype(np.float32))
        return x

    # ---- the decode loop (shared by the naive and the batched entry points) -----------------
    def generate(
        self,
        phoneme_ids: Sequence[np.ndarray],
        bert: Sequence[np.ndarray],
        prompt: Optional[np.ndarray],  # [B, P] or None (reference-free)
        top_k: int = 15,
        top_p: float = 1.0,
        temperature: float = 1.0,
        repetition_penalty: float = 1.35,
        early_stop_num: int = -1,
        eos_window: int = EOS_WINDOW_BATCH,
        max_steps: int = MAX_STEPS,
        seed: int = 0,
        forced: Optional[np.ndarray] = None,  # [B, n] teacher-forced tokens (test hook)
        record_logits: bool = False,
        noise_fn: Optional[Callable[[int, int, int], np.ndarray]] = None,
    ):
        """Returns dict(tokens=[per-slot int64 arrays prompt+kept], idx=[...], logits=[steps][n,1025]
        (raw, before penalty; rows in active order), active=[steps] slot lists, sampled=[B,steps],
        margins=[B,steps])."""
        s = OracleSession(self, phoneme_ids, bert, prompt, top_k=top_k, top_p=top_p, temperature=temperature,
                          repetition_penalty=repetition_penalty, early_stop_num=early_stop_num, eos_window=eos_window,
                          max_steps=max_steps, seed=seed, forced=forced, record_logits=record_logits, noise_fn=noise_fn)
        s.run()
        return s.result()

    # ---- reference-shaped entry points ---------------------------------------------------------
    def infer_panel_naive(self, x, x_lens, prompts, bert_feature, top_k=-100, top_p=100,
                          early_stop_num=-1, temperature=1.0, repetition_penalty=1.35, **kw):
        """x [1,L], prompts [1,P]|None, bert [1,1024,L] -> (y [1, P+idx], idx)  (t2s_model.py:814-918).
        Reference-free (prompts None) returns idx 0 (:916-917)."""
        r = self.generate([np.asarray(x)[0]], [np.asarray(bert_feature)[0]],
                          None if prompts is None else np.asarray(prompts),
                          top_k=top_k, top_p=top_p, temperature=temperature,
                          repetition_penalty=repetition_penalty, early_stop_num=early_stop_num,
                          eos_window=EOS_WINDOW_NAIVE, **kw)
        return r["tokens"][0][None, :], (0 if prompts is None else r["idx"][0])

    def infer_panel_naive_batched(self, x, x_lens, prompts, bert_feature, top_k=-100, top_p=100,
                                  early_stop_num=-1, temperature=1.0, repetition_penalty=1.35, **kw):
        """The reference's Python loop of infer_panel_naive over the items (t2s_model.py:781-812): lists in, lists out;
        `prompts` [B,P] or None (every item reference-free: int tokens, idx 0)."""
        ys, idxs = [], []
        for i in range(len(x)):
            y, idx = self.infer_panel_naive(np.asarray(x[i])[None], None,
                                            None if prompts is None else np.asarray(prompts)[i][None],
                                            np.asarray(bert_feature[i])[None], top_k=top_k, top_p=top_p,
                                            early_stop_num=early_stop_num, temperature=temperature,
                                            repetition_penalty=repetition_penalty, **kw)
            ys.append(y[0])
            idxs.append(idx)
        return ys, idxs

    def infer_panel_batch_infer(self, x, x_lens, prompts, bert_feature, top_k=-100, top_p=100,
                                early_stop_num=-1, temperature=1.0, repetition_penalty=1.35, **kw):
        """Lists as in TTS.run -> (List[y_i], List[idx_i]) in original order (t2s_model.py:583-779)."""
        kw.pop("max_len", None)  # left padding is inert, see module docstring
        r = self.generate([np.asarray(t) for t in x], [np.asarray(t) for t in bert_feature],
                          np.asarray(prompts), top_k=top_k, top_p=top_p, temperature=temperature,
                          repetition_penalty=repetition_penalty, early_stop_num=early_stop_num,
                          eos_window=EOS_WINDOW_BATCH, **kw)
        return r["tokens"], r["idx"]


class OracleSession:
    """One infer_panel* call as a resumable object: prefill at construction, then ``run(n)`` executes at most n steps of the
    loop (t2s_model.py:701-769 / :878-914).  ``T2SOracle.generate`` is ``OracleSession(...).run()``; bench.py's CPU arms time
    ``run(n)`` slices of one resident session so that a bounded sample measures DECODE steps, not the prefill."""

    def __init__(self, o: "T2SOracle", phoneme_ids, bert, prompt, top_k=15, top_p=1.0, temperature=1.0,
                 repetition_penalty=1.35, early_stop_num=-1, eos_window=EOS_WINDOW_BATCH, max_steps=MAX_STEPS, seed=0,
                 forced=None, record_logits=False, noise_fn=None):
        import time
        self.t_start = time.perf_counter()
        self.o = o
        self.kw = dict(top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty)
        self.early_stop_num, self.eos_window, self.max_steps = early_stop_num, eos_window, max_steps
        self.forced, self.record_logits = forced, record_logits
        B = len(phoneme_ids)
        P = 0 if prompt is None else int(prompt.shape[1])
        self.B, self.P = B, P
        self.noise_fn = noise_fn or (lambda slot, step, n: so.exp_noise(seed, slot, step, n))
        S0 = [len(phoneme_ids[b]) + P for b in range(B)]
        cap = max(S0) + (max_steps if early_stop_num == -1 else min(max_steps, early_stop_num + 1)) + 1
        self.K = np.zeros((o.L, B, o.H, cap, o.dh), np.float32)
        self.V = np.zeros_like(self.K)
        self.hid = np.zeros((B, o.d), np.float32)
        for b in range(B):
            x = o.embed_text(np.asarray(phoneme_ids[b]), np.asarray(bert[b], dtype=np.float32))
            xy = x if P == 0 else np.concatenate([x, o.embed_audio(np.asarray(prompt[b]), 0)], axis=0)
            h, Kb, Vb = o.prefill_one(xy, len(phoneme_ids[b]), cap)
            self.hid[b], self.K[:, b], self.V[:, b] = h, Kb, Vb
        self.hist: List[List[int]] = [list(map(int, prompt[b])) if P else [] for b in range(B)]
        self.gen: List[List[int]] = [[] for _ in range(B)]
        self.idx_out = [None] * B
        self.active = list(range(B))
        self.lens = np.array(S0, dtype=np.int64)
        self.step = 0
        self.out = dict(logits=[], active=[], sampled=np.full((B, max_steps), -1, np.int64),
                        margins=np.ones((B, max_steps), np.float32))
        self.out["t_prefill"] = time.perf_counter() - self.t_start

    def run(self, n_steps: int = -1) -> int:
        """Executes up to n_steps loop iterations (all remaining if < 0); returns how many ran."""
        o, out = self.o, self.out
        ran = 0
        while self.active and self.step < self.max_steps and (n_steps < 0 or ran < n_steps):
            step, active = self.step, self.active
            logits = (self.hid[active] @ o.wp.T).astype(np.float32)  # [n, 1025]
            if self.record_logits:
                out["logits"].append(logits.copy())
            out["active"].append(list(active))
            width = o.EOS if step < self.eos_window else o.EOS + 1
            still = []
            for r, b in enumerate(active):
                row = logits[r, :width].copy()
                q = self.noise_fn(b, step, width)
                tok, greedy, _, margin = so.sample_row(
                    row, self.hist[b] + self.gen[b], q, self.kw["temperature"], self.kw["top_k"], self.kw["top_p"],
                    self.kw["repetition_penalty"])
                out["sampled"][b, step] = tok
                out["margins"][b, step] = margin
                if self.forced is not None and step < self.forced.shape[1]:
                    tok = int(self.forced[b, step])
                self.gen[b].append(tok)
                stop = tok == o.EOS or greedy == o.EOS
                if self.early_stop_num != -1 and (step + 1) > self.early_stop_num:
                    stop = True
                if step == self.max_steps - 1:
                    stop = True
                if stop:
                    self.idx_out[b] = step
                else:
                    still.append(b)
            self.active = still
            self.step += 1
            ran += 1
            if not still:
                break
            x = np.stack([o.embed_audio(np.array([self.gen[b][-1]]), self.P + step)[0] for b in still])
            sel = np.array(still)
            self.hid[sel] = o.decode_step(x, self.K, self.V, sel, self.lens[sel])
            self.lens[sel] += 1
        return ran

    def result(self):
        import time
        out = self.out
        out["t_total"] = time.perf_counter() - self.t_start
        out["tokens"] = [np.array(self.hist[b] + self.gen[b][: (self.idx_out[b] if self.idx_out[b] is not None else len(self.gen[b]))],
                                  dtype=np.int64) for b in range(self.B)]
        out["idx"] = [int(i) if i is not None else -1 for i in self.idx_out]
        out["generated"] = self.gen
        return out


def codes_to_latent(codes: np.ndarray, codebook: np.ndarray, upsample: int = 2) -> np.ndarray:
    """The first op after the path: ``quantizer.decode(codes)`` then nearest-neighbour upsampling by ``upsample``
    (module/models.py:989-991; core_vq.py:359-365 sum over the single residual layer, :286-290 lookup + "b n d -> b d n",
    :195-197 dequantize = F.embedding).  codes int64 [T] -> fp32 [1, dim, upsample*T]."""
    codes = np.asarray(codes).reshape(-1)
    if codes.size and (codes.min() < 0 or codes.max() >= codebook.shape[0]):
        raise IndexError("index out of range in self")  # F.embedding's error
    q = codebook[codes].astype(np.float32).T            # [dim, T]
    # F.interpolate(mode="nearest") to an integer multiple: output position p reads input floor(p / upsample)
    return np.repeat(q, upsample, axis=1)[None]
