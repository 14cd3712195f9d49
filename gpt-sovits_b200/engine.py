"""T2SEngine — Python handle on the C-ABI engine (include/t2s_b200.h).

PyTorch is used here only for device memory, streams and tensor plumbing; every numerical step runs
inside libt2s_b200.so.  Argument meaning follows the reference's infer_panel family
(GPT_SoVITS/AR/models/t2s_model.py:583-595, :814-826).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .synthetic import BERT_DIM, PE_LEN, S1V2_CONFIG, sine_pe

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}

MAX_STEPS = 1500  # t2s_model.py:701/:878
EOS_WINDOW_NAIVE = 11  # t2s_model.py:888
EOS_WINDOW_BATCH = 1  # t2s_model.py:708-710

_GLOBAL_KEYS = {
    "bert_proj.weight": _lib.W_BERT_PROJ_W,
    "bert_proj.bias": _lib.W_BERT_PROJ_B,
    "ar_text_embedding.word_embeddings.weight": _lib.W_TEXT_EMB,
    "ar_text_position.alpha": _lib.W_TEXT_ALPHA,
    "ar_audio_embedding.word_embeddings.weight": _lib.W_AUDIO_EMB,
    "ar_audio_position.alpha": _lib.W_AUDIO_ALPHA,
    "ar_predict_layer.weight": _lib.W_PREDICT,
}
_LAYER_KEYS = {
    "self_attn.in_proj_weight": _lib.W_IN_PROJ_W,
    "self_attn.in_proj_bias": _lib.W_IN_PROJ_B,
    "self_attn.out_proj.weight": _lib.W_OUT_PROJ_W,
    "self_attn.out_proj.bias": _lib.W_OUT_PROJ_B,
    "linear1.weight": _lib.W_LIN1_W,
    "linear1.bias": _lib.W_LIN1_B,
    "linear2.weight": _lib.W_LIN2_W,
    "linear2.bias": _lib.W_LIN2_B,
    "norm1.weight": _lib.W_NORM1_W,
    "norm1.bias": _lib.W_NORM1_B,
    "norm2.weight": _lib.W_NORM2_W,
    "norm2.bias": _lib.W_NORM2_B,
}


@dataclass
class InferResult:
    tokens: torch.Tensor  # [B, P + max_steps] int64 (device, or pinned host for host I/O); -1 beyond P+idx
    idx: List[int]
    prompt_len: int
    logits: Optional[torch.Tensor] = None  # [n, B, 1025] raw logits of the first n steps (capture hook)
    sampled: Optional[torch.Tensor] = None  # [B, n] raw sampled tokens before teacher forcing
    stats: Dict[str, float] = field(default_factory=dict)
    prompt_lens: Optional[List[int]] = None  # per slot, when utterances with other prompts were admitted into the session

    def sequences(self) -> List[torch.Tensor]:
        """prompt ++ kept tokens per utterance, original order (t2s_model.py:733,779)."""
        P = self.prompt_lens or [self.prompt_len] * len(self.idx)
        return [self.tokens[b, : P[b] + max(i, 0)] for b, i in enumerate(self.idx)]


class T2SEngine:
    def __init__(self, config: Optional[dict] = None, device: str | torch.device = "cuda:0", max_batch: int = 256):
        self._h = None
        if not torch.cuda.is_available():
            raise RuntimeError("gpt-sovits_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("gpt-sovits_b200 runs on CUDA devices only (got %s)" % self.device)
        m = (config or S1V2_CONFIG)["model"]
        self.config = m
        self.n_layer = int(m["n_layer"])
        self.vocab = int(m["vocab_size"])
        self.eos = int(m["EOS"])
        cfg = _lib.ModelConfig(
            n_layer=self.n_layer, d_model=int(m["hidden_dim"]), n_head=int(m["head"]), d_ff=4 * int(m["hidden_dim"]),
            vocab=self.vocab, phoneme_vocab=int(m["phoneme_vocab_size"]), bert_dim=BERT_DIM, eos=self.eos,
            pe_len=PE_LEN, max_batch=max_batch)
        if int(m["embedding_dim"]) != int(m["hidden_dim"]):
            raise RuntimeError("embedding_dim != hidden_dim is not supported")
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.t2s_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._keep = []  # tensors the engine holds raw pointers to during a session

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, "_h", None):
            with torch.cuda.device(self.device):
                self.lib.t2s_destroy(self._h)
            self._h = None

    # ---- weights -----------------------------------------------------------------------------------
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _load(self, tid: int, layer: int, t: torch.Tensor) -> None:
        t = t.detach()
        if t.dtype not in _DT:
            t = t.float()
        t = t.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.t2s_load_tensor(self._h, tid, layer, C.c_void_p(t.data_ptr()), _DT[t.dtype],
                                                t.numel(), 1 if t.is_cuda else 0, self._stream()))
            if t.is_cuda:
                torch.cuda.current_stream(self.device).synchronize()

    def load_state_dict(self, sd: Dict[str, torch.Tensor], pe: Optional[torch.Tensor] = None) -> None:
        """`sd` uses the reference's key names, optionally with the checkpoint's ``model.`` prefix
        (SURVEY.md section 8b).  `pe` = SinePositionalEmbedding.pe ([4000,512]); regenerated if None."""
        sd = {(k[6:] if k.startswith("model.") else k): v for k, v in sd.items()}
        for key, tid in _GLOBAL_KEYS.items():
            if key not in sd:
                raise KeyError("state_dict is missing " + key)
            self._load(tid, 0, sd[key])
        for i in range(self.n_layer):
            for key, tid in _LAYER_KEYS.items():
                full = f"h.layers.{i}.{key}"
                if full not in sd:
                    raise KeyError("state_dict is missing " + full)
                self._load(tid, i, sd[full])
        if pe is None:
            pe = sine_pe(PE_LEN, int(self.config["hidden_dim"]))
        pe = pe.reshape(-1, pe.shape[-1])[:PE_LEN].float()
        self._load(_lib.W_PE, 0, pe)

    def set_option(self, opt: int, value: int) -> None:
        _lib.check(self.lib.t2s_set_option(self._h, opt, int(value)))

    def _request(self, phoneme_ids, bert, prompt, top_k, top_p, temperature, repetition_penalty, early_stop_num,
                 eos_suppress_steps, max_steps, seed, host_io):
        """Validates one infer_panel-shaped call and packs it into the C-ABI request (include/t2s_b200.h t2s_request).
        Returns (request, tensors to keep alive, batch, prompt length)."""
        B = len(phoneme_ids)
        if B == 0:
            raise ValueError("empty batch")
        if len(bert) != B:
            raise ValueError("len(bert) != len(phoneme_ids)")
        dev = self.device
        want_cuda = not host_io
        lens = [int(t.shape[0]) for t in phoneme_ids]
        for i, (p, f) in enumerate(zip(phoneme_ids, bert)):
            if f.shape[0] != BERT_DIM or f.shape[1] != lens[i]:
                raise ValueError(f"bert[{i}] has shape {tuple(f.shape)}, expected [{BERT_DIM}, {lens[i]}]")
            if p.is_cuda != want_cuda or f.is_cuda != want_cuda:
                raise ValueError("inputs must all be on %s" % ("the CUDA device" if want_cuda else "the host (host_io)"))
        ids = torch.cat([t.reshape(-1).to(torch.int64) for t in phoneme_ids]).contiguous()
        bdt = bert[0].dtype
        if bdt not in _DT:
            bert = [f.float() for f in bert]
            bdt = torch.float32
        bert = [f if f.dtype == bdt else f.to(bdt) for f in bert]
        if host_io:
            bert = [f.contiguous() for f in bert]
        P = 0
        prow = 0
        if prompt is not None:
            if prompt.dim() != 2 or prompt.shape[0] != B:
                raise ValueError("prompt must be [B, P]")
            if prompt.is_cuda != want_cuda:
                raise ValueError("prompt is on the wrong device")
            if prompt.dtype != torch.int64:
                prompt = prompt.to(torch.int64)
            P = int(prompt.shape[1])
            if P > 0 and prompt.stride(1) != 1:
                prompt = prompt.contiguous()
            prow = int(prompt.stride(0)) if B > 1 else max(int(prompt.stride(0)), 0)
            if P == 0:
                prompt = None
        if seed is None:
            # keeps TTS.run's set_seed() meaningful (TTS.py:194-214): drawn from torch's seeded generator
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        lens_c = (C.c_int32 * B)(*lens)
        ptrs = (C.c_void_p * B)(*[f.data_ptr() for f in bert])
        sc = (C.c_int64 * B)(*[int(f.stride(0)) for f in bert])
        st = (C.c_int64 * B)(*[int(f.stride(1)) for f in bert])
        rq = _lib.Request(
            batch=B, phoneme_ids=ids.data_ptr(), phoneme_lens=lens_c, bert=ptrs, bert_stride_c=sc, bert_stride_t=st,
            bert_dtype=_DT[bdt], prompt=(prompt.data_ptr() if prompt is not None else None), prompt_row_stride=prow,
            prompt_len=P, top_k=int(top_k), top_p=float(top_p), temperature=float(temperature),
            repetition_penalty=float(repetition_penalty), early_stop_num=int(early_stop_num),
            eos_suppress_steps=int(eos_suppress_steps), max_steps=int(max_steps), seed=int(seed) & (2 ** 64 - 1),
            inputs_on_host=1 if host_io else 0)
        rq._keep = (lens_c, ptrs, sc, st)
        return rq, (ids, bert, prompt), B, P

    # ---- one infer_panel call ------------------------------------------------------------------------
    def infer(
        self,
        phoneme_ids: Sequence[torch.Tensor],
        bert: Sequence[torch.Tensor],
        prompt: Optional[torch.Tensor],
        top_k: int = 15,
        top_p: float = 1.0,
        temperature: float = 1.0,
        repetition_penalty: float = 1.35,
        early_stop_num: int = -1,
        eos_suppress_steps: int = EOS_WINDOW_BATCH,
        max_steps: int = MAX_STEPS,
        seed: Optional[int] = None,
        forced: Optional[torch.Tensor] = None,
        capture_logits: int = 0,
        max_new_steps: int = -1,
        host_io: bool = False,
        reserve_slots: int = 0,
        reserve_positions: int = 0,
        utt_ids: Optional[Sequence[int]] = None,
        hooks_by_utterance: int = 0,
    ) -> InferResult:
        """phoneme_ids: B tensors [L_i] int64; bert: B tensors [1024, L_i]; prompt: [B, P] int64 or None.
        host_io=True takes CPU tensors and returns CPU tokens: the H2D / D2H copies happen inside the
        C-ABI call (bench.py's end-to-end leg).  reserve_slots / reserve_positions > 0 open the session with room for
        later ``admit()`` calls (continuous batching): that many slots in total, that many K/V positions per slot.
        utt_ids: one integer per utterance keying its Philox stream (default: its index in this call), so that sampling does not
        depend on how a caller shards or batches the utterances.  hooks_by_utterance = n > 0: the rows of ``forced`` / the captured
        logits are indexed by utterance id (n rows) instead of by slot, so they follow an utterance through recycled slots."""
        rq, keep, B, P = self._request(phoneme_ids, bert, prompt, top_k, top_p, temperature, repetition_penalty, early_stop_num,
                                       eos_suppress_steps, max_steps, seed, host_io)
        ids, bert, prompt = keep
        dev = self.device
        cap = max(B, int(reserve_slots))
        if hooks_by_utterance > 0:
            cap = int(hooks_by_utterance)  # rows of the hook buffers
        self.set_option(_lib.OPT_HOOKS_BY_UTTERANCE, int(hooks_by_utterance))
        self.set_option(_lib.OPT_SESSION_SLOTS, int(reserve_slots))
        self.set_option(_lib.OPT_SESSION_POSITIONS, int(reserve_positions))
        self._slot_P = [P] * B
        self._free: List[int] = []
        self._set_utt_ids(utt_ids, B)
        self._session_kw = dict(top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty,
                                early_stop_num=early_stop_num, eos_suppress_steps=eos_suppress_steps, max_steps=max_steps, seed=seed)
        self._max_steps = int(max_steps)
        width = P + int(max_steps)
        with torch.cuda.device(dev):
            stream = self._stream()
            forced_dev = None
            n_forced = 0
            if forced is not None:
                forced_dev = forced.to(device=dev, dtype=torch.int32).contiguous()
                n_forced = int(forced_dev.shape[1])
            _lib.check(self.lib.t2s_set_forced_tokens(
                self._h, C.c_void_p(forced_dev.data_ptr()) if forced_dev is not None else None, n_forced))
            logits_buf = None
            if capture_logits > 0:
                logits_buf = torch.full((capture_logits, cap, self.vocab), float("nan"), device=dev, dtype=torch.float32)
            _lib.check(self.lib.t2s_set_logits_capture(
                self._h, C.c_void_p(logits_buf.data_ptr()) if logits_buf is not None else None, int(capture_logits)))
            self._keep = [ids, bert, prompt, forced_dev, logits_buf]
            idx = (C.c_int32 * B)()
            if host_io:
                tokens = torch.empty((B, width), dtype=torch.int64).pin_memory()
            else:
                tokens = torch.empty((B, width), dtype=torch.int64, device=dev)
            if max_new_steps < 0 and forced is None and capture_logits == 0:
                _lib.check(self.lib.t2s_generate(self._h, C.byref(rq), C.c_void_p(tokens.data_ptr()), width,
                                                 1 if host_io else 0, idx, stream))
            else:
                _lib.check(self.lib.t2s_prefill(self._h, C.byref(rq), stream))
                n = C.c_int32(0)
                _lib.check(self.lib.t2s_decode(self._h, int(max_new_steps), stream, C.byref(n)))
                _lib.check(self.lib.t2s_result(self._h, C.c_void_p(tokens.data_ptr()), width, 1 if host_io else 0, idx, stream))
            # (with reserved slots logits_buf is [n, capacity, V]: the rows of admitted slots fill in as they decode)
            res = InferResult(tokens=tokens, idx=[int(v) for v in idx], prompt_len=P, logits=logits_buf, stats=self.stats())
            if forced is not None or capture_logits > 0:
                n_s = max(n_forced, capture_logits, 1)
                samp = torch.empty((B, n_s), dtype=torch.int32)
                _lib.check(self.lib.t2s_get_sampled(self._h, C.c_void_p(samp.data_ptr()), n_s, stream))
                res.sampled = samp
        return res

    def _set_utt_ids(self, utt_ids, B: int) -> None:
        if utt_ids is None:
            _lib.check(self.lib.t2s_set_utterance_ids(self._h, None, 0))
            return
        if len(utt_ids) != B:
            raise ValueError("utt_ids must have one entry per utterance")
        arr = (C.c_int32 * B)(*[int(v) & 0x7FFFFFFF for v in utt_ids])
        _lib.check(self.lib.t2s_set_utterance_ids(self._h, arr, B))

    def release(self, slots: Sequence[int]) -> None:
        """Hands the slots of FINISHED utterances back to the resident session (after their tokens were fetched): the next
        ``admit()`` reuses them and their K/V pages (C ABI: t2s_release_slots)."""
        arr = (C.c_int32 * len(slots))(*[int(v) for v in slots])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.t2s_release_slots(self._h, arr, len(slots), self._stream()))
        self._free = sorted(set(self._free) | set(int(v) for v in slots))

    def admit(self, phoneme_ids: Sequence[torch.Tensor], bert: Sequence[torch.Tensor], prompt: Optional[torch.Tensor],
              utt_ids: Optional[Sequence[int]] = None) -> List[int]:
        """Continuous batching: adds utterances to the resident session (opened by ``infer(..., max_new_steps=k,
        reserve_slots=n)``) at its current step; returns their slot indices (= their rows in ``result()``): released slots
        first, lowest first, then fresh ones.  Sampling parameters and stop rules are the session's; the prompt may differ from
        the first request's (C ABI: t2s_admit)."""
        kw = self._session_kw
        rq, keep, B, P = self._request(phoneme_ids, bert, prompt, kw["top_k"], kw["top_p"], kw["temperature"],
                                       kw["repetition_penalty"], kw["early_stop_num"], kw["eos_suppress_steps"], kw["max_steps"],
                                       kw["seed"] if kw["seed"] is not None else 0, False)
        self._set_utt_ids(utt_ids, B)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.t2s_admit(self._h, C.byref(rq), self._stream()))
        self._keep.append(keep)
        slots = []
        for _ in range(B):  # the C side's order: released slots first (ascending), then fresh ones
            if self._free:
                sl = self._free.pop(0)
                self._slot_P[sl] = P
            else:
                sl = len(self._slot_P)
                self._slot_P.append(P)
            slots.append(sl)
        return slots

    def session_result(self) -> InferResult:
        """``result()`` for every slot of the resident session, admitted ones included."""
        return self.result(len(self._slot_P), max(self._slot_P), self._max_steps, prompt_lens=list(self._slot_P))

    def sampled(self, n_steps: int) -> torch.Tensor:
        """The raw sampled tokens (before teacher forcing) of every slot of the resident session, [slots, n_steps] int32 on the
        host, indexed by each sequence's OWN step."""
        samp = torch.empty((len(self._slot_P), n_steps), dtype=torch.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.t2s_get_sampled(self._h, C.c_void_p(samp.data_ptr()), n_steps, self._stream()))
        return samp

    def decode_more(self, max_new_steps: int = -1) -> int:
        """Continue the resident session (after ``infer(..., max_new_steps=k)``) for at most ``max_new_steps`` further
        steps; returns the steps executed.  With ``result()`` this is the streaming form: sequences that have stopped can be
        handed to the vocoder while the others keep decoding (the reference's return_fragment mode, TTS.py:1049-1053)."""
        n = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.t2s_decode(self._h, int(max_new_steps), self._stream(), C.byref(n)))
        return int(n.value)

    def result(self, batch: int, prompt_len: int, max_steps: int = MAX_STEPS, prompt_lens: Optional[List[int]] = None) -> InferResult:
        """Tokens / idx of the resident session as they stand (idx = -1: still decoding)."""
        width = prompt_len + int(max_steps)
        idx = (C.c_int32 * batch)()
        with torch.cuda.device(self.device):
            tokens = torch.empty((batch, width), dtype=torch.int64, device=self.device)
            _lib.check(self.lib.t2s_result(self._h, C.c_void_p(tokens.data_ptr()), width, 0, idx, self._stream()))
        return InferResult(tokens=tokens, idx=[int(v) for v in idx], prompt_len=prompt_len, stats=self.stats(), prompt_lens=prompt_lens)

    def codes_to_latent(self, codes: torch.Tensor, codebook: torch.Tensor, upsample: int = 2) -> torch.Tensor:
        """``F.interpolate(quantizer.decode(codes), size=upsample*T, mode="nearest")`` of SynthesizerTrn.decode
        (module/models.py:989-991) on device: ``codes`` int64 ``[T]`` (or the reference's ``[1, 1, T]``), ``codebook`` fp32
        ``[codebook_size, dim]`` -> fp32 ``[1, dim, upsample*T]``."""
        if codes.dtype != torch.int64 or codebook.dtype != torch.float32 or codebook.dim() != 2:
            raise TypeError("codes_to_latent: codes must be int64 and codebook a 2-d fp32 tensor")
        if codes.dim() == 3 and (codes.shape[0] != 1 or codes.shape[1] != 1):
            raise ValueError("codes_to_latent: codes must be [T] or [1, 1, T] (n_q = 1, one utterance)")
        codes = codes.reshape(-1).to(self.device).contiguous()
        codebook = codebook.to(self.device).contiguous()
        n, dim = int(codes.numel()), int(codebook.shape[1])
        with torch.cuda.device(self.device):
            out = torch.empty((1, dim, upsample * n), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.t2s_codes_to_latent(self._h, C.c_void_p(codes.data_ptr()), n, C.c_void_p(codebook.data_ptr()),
                                                    int(codebook.shape[0]), dim, int(upsample), C.c_void_p(out.data_ptr()),
                                                    self._stream()))
        return out

    def stats(self) -> Dict[str, float]:
        s = _lib.Stats()
        _lib.check(self.lib.t2s_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in _lib.Stats._fields_}
