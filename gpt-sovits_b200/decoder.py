"""Host-side mirror of the reference operator interface for the T2S decode path.

Same names, argument meaning, return convention and error behaviour as
``Text2SemanticDecoder.infer_panel`` / ``infer_panel_naive`` / ``infer_panel_naive_batched`` /
``infer_panel_batch_infer`` (GPT_SoVITS/AR/models/t2s_model.py:583-935), so that
``TTS_infer_pack/TTS.py:1215`` and ``api.py:935`` / ``inference_webui.py:878`` keep working unchanged:

    from gpt_sovits_b200 import patch_reference
    patch_reference()          # class-level patch; every (re)built model instance resolves to it

``TTS.run`` rebinds ``model.infer_panel`` on the *instance* at every request to one of the batched
variants (TTS.py:1042-1047) and rebuilds the model object on weight switch / error
(TTS.py:585-599, 1352-1363), hence the patch is applied to the class and the packed weights are cached
per instance, keyed by a fingerprint of the parameters (data_ptr, _version, dtype).

No CPU fallback: on a non-CUDA model the patched methods raise.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .engine import EOS_WINDOW_BATCH, EOS_WINDOW_NAIVE, MAX_STEPS, T2SEngine

_ENGINE_ATTR = "_b200_engine"
_FPRINT_ATTR = "_b200_fingerprint"


def _fingerprint(module) -> Tuple:
    return tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in module.parameters())


def engine_for(module) -> T2SEngine:
    """Engine holding `module`'s current weights (re-packed when any parameter changed: .half(), .to(),
    load_state_dict all mutate the parameters in place)."""
    dev = next(module.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("gpt-sovits_b200: the model is on %s; the B200 path has no CPU fallback" % dev)
    fp = _fingerprint(module)
    eng = getattr(module, _ENGINE_ATTR, None)
    if eng is None or getattr(module, _FPRINT_ATTR, None) != fp:
        if eng is not None:
            eng.close()
        cfg = {"model": {
            "hidden_dim": module.model_dim, "embedding_dim": module.embedding_dim, "head": module.num_head,
            "n_layer": module.num_layers, "vocab_size": module.vocab_size,
            "phoneme_vocab_size": module.phoneme_vocab_size, "dropout": 0.0, "EOS": module.EOS}}
        eng = T2SEngine(cfg, device=dev)
        sd = {k: v for k, v in module.state_dict().items()
              if not k.startswith(("ar_accuracy_metric", "loss_fct"))}
        pe = getattr(module.ar_audio_position, "pe", None)
        eng.load_state_dict(sd, pe=None if pe is None else pe.detach().float().cpu()[0])
        object.__setattr__(module, _ENGINE_ATTR, eng)
        object.__setattr__(module, _FPRINT_ATTR, fp)
    return eng


def _check_top_k(top_k: int) -> None:
    if top_k < 1:
        # the reference dies inside torch.topk for top_k <= 0 (utils.py:184); keep it an error
        raise RuntimeError(f"selected index k out of range (top_k={top_k})")


def infer_panel_batch_infer(self, x: List[torch.LongTensor], x_lens: torch.LongTensor, prompts: torch.LongTensor,
                            bert_feature: List[torch.Tensor], top_k: int = -100, top_p: int = 100,
                            early_stop_num: int = -1, temperature: float = 1.0, repetition_penalty: float = 1.35,
                            **kwargs):
    """t2s_model.py:583-779.  `max_len` (kwargs) is accepted and ignored: left padding is inert
    (SURVEY.md section 8a item 3), sequences stay ragged on the device."""
    if prompts is None:
        print("Warning: Prompt free is not supported batch_infer! switch to naive_infer")
        return infer_panel_naive_batched(self, x, x_lens, prompts, bert_feature, top_k=top_k, top_p=top_p,
                                         early_stop_num=early_stop_num, temperature=temperature, **kwargs)
    _check_top_k(top_k)
    eng = engine_for(self)
    r = eng.infer(list(x), list(bert_feature), prompts, top_k=top_k, top_p=top_p, temperature=temperature,
                  repetition_penalty=repetition_penalty, early_stop_num=early_stop_num,
                  eos_suppress_steps=EOS_WINDOW_BATCH, max_steps=MAX_STEPS)
    return r.sequences(), r.idx


def infer_panel_naive_batched(self, x: List[torch.LongTensor], x_lens: torch.LongTensor, prompts: torch.LongTensor,
                              bert_feature: List[torch.Tensor], top_k: int = -100, top_p: int = 100,
                              early_stop_num: int = -1, temperature: float = 1.0, repetition_penalty: float = 1.35,
                              **kwargs):
    """t2s_model.py:781-812: the reference loops infer_panel_naive over the items; sequences never
    interact, so the same per-item semantics (11-step EOS window) run here as ONE ragged batch."""
    _check_top_k(top_k)
    eng = engine_for(self)
    r = eng.infer(list(x), list(bert_feature), prompts, top_k=top_k, top_p=top_p, temperature=temperature,
                  repetition_penalty=repetition_penalty, early_stop_num=early_stop_num,
                  eos_suppress_steps=EOS_WINDOW_NAIVE, max_steps=MAX_STEPS)
    seqs = r.sequences()
    if prompts is None:  # reference-free: int32 tokens, idx 0 (t2s_model.py:849-856, :916-917)
        return [s.to(torch.int32) for s in seqs], [0] * len(seqs)
    return seqs, r.idx


def infer_panel_naive(self, x: torch.LongTensor, x_lens: torch.LongTensor, prompts: torch.LongTensor,
                      bert_feature: torch.Tensor, top_k: int = -100, top_p: int = 100, early_stop_num: int = -1,
                      temperature: float = 1.0, repetition_penalty: float = 1.35, **kwargs):
    """t2s_model.py:814-918: x [1,L], prompts [1,P] | None, bert_feature [1,1024,L] -> (y [1,P+idx], idx)."""
    _check_top_k(top_k)
    bsz = x.shape[0]
    if bsz != 1:  # before any engine work (the reference would run but only ever looks at row 0's stop condition)
        raise RuntimeError("infer_panel_naive: batch size must be 1 (use infer_panel_batch_infer)")
    eng = engine_for(self)
    r = eng.infer([x[i] for i in range(bsz)], [bert_feature[i] for i in range(bsz)], prompts, top_k=top_k,
                  top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty,
                  early_stop_num=early_stop_num, eos_suppress_steps=EOS_WINDOW_NAIVE, max_steps=MAX_STEPS)
    y = r.sequences()[0].unsqueeze(0)
    if prompts is None:
        return y.to(torch.int32), 0
    return y, r.idx[0]


def infer_panel(self, x, x_lens, prompts, bert_feature, top_k: int = -100, top_p: int = 100, early_stop_num: int = -1,
                temperature: float = 1.0, repetition_penalty: float = 1.35, **kwargs):
    """t2s_model.py:920-935 (alias of infer_panel_naive)."""
    return infer_panel_naive(self, x, x_lens, prompts, bert_feature, top_k, top_p, early_stop_num, temperature,
                             repetition_penalty, **kwargs)


def infer_panel_stream(self, x: List[torch.LongTensor], x_lens: torch.LongTensor, prompts: torch.LongTensor,
                       bert_feature: List[torch.Tensor], top_k: int = -100, top_p: int = 100, early_stop_num: int = -1,
                       temperature: float = 1.0, repetition_penalty: float = 1.35, slice_steps: int = 25, **kwargs):
    """Generator form of infer_panel_batch_infer for the reference's ``return_fragment`` mode (TTS.py:1049-1053, 1319-1329;
    api_v2.py:348-365 streams the audio of each fragment as soon as it exists): yields ``(i, y_i, idx_i)`` - original batch
    index, prompt ++ kept tokens, idx exactly as infer_panel_batch_infer returns them - in the order in which the sequences
    RETIRE, while the rest of the batch keeps decoding, so SoVITS can start on the first finished sentence instead of waiting
    for the slowest one.  Not part of the class-level patch (the reference has no such method); a caller opts in."""
    from .batching import StreamingSession
    if prompts is None:
        raise RuntimeError("infer_panel_stream needs prompts (the reference's batched path has no reference-free mode)")
    _check_top_k(top_k)
    eng = engine_for(self)
    sess = StreamingSession(eng, slots=len(x), slice_steps=slice_steps, top_k=top_k, top_p=top_p, temperature=temperature,
                            repetition_penalty=repetition_penalty, early_stop_num=early_stop_num,
                            eos_suppress_steps=EOS_WINDOW_BATCH, max_steps=MAX_STEPS)
    sess.submit(list(x), list(bert_feature), prompts)
    yield from sess


_PATCHED: Dict[type, Dict[str, object]] = {}
_METHODS = {"infer_panel": infer_panel, "infer_panel_naive": infer_panel_naive,
            "infer_panel_naive_batched": infer_panel_naive_batched, "infer_panel_batch_infer": infer_panel_batch_infer}


def patch_reference(decoder_cls: Optional[type] = None) -> type:
    """Class-level patch of Text2SemanticDecoder (or any class with the same attributes)."""
    if decoder_cls is None:
        from AR.models.t2s_model import Text2SemanticDecoder as decoder_cls  # the caller put GPT_SoVITS on sys.path
    if decoder_cls in _PATCHED:
        return decoder_cls
    _PATCHED[decoder_cls] = {n: decoder_cls.__dict__.get(n) for n in _METHODS}
    for n, f in _METHODS.items():
        setattr(decoder_cls, n, f)
    return decoder_cls


def unpatch_reference(decoder_cls: type) -> None:
    saved = _PATCHED.pop(decoder_cls, None)
    if saved:
        for n, f in saved.items():
            if f is None:
                delattr(decoder_cls, n)
            else:
                setattr(decoder_cls, n, f)
