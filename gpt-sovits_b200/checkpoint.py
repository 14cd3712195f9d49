"""s1 checkpoint -> engine, without instantiating the Lightning module.

The reference loads text-to-semantic weights in ``TTS.init_t2s_weights`` (GPT_SoVITS/TTS_infer_pack/TTS.py:585-599):
``torch.load(path)`` gives ``{"weight": state_dict, "config": dict, "info": str}`` (written by ``my_save``,
GPT_SoVITS/process_ckpt.py:12-17, from s1_train.py:62-81); the state_dict keys carry the Lightning wrapper's ``model.``
prefix and the tensors are fp16; ``Text2SemanticLightningModule(config, ...)`` is built only to receive them.  This module
reads the same file, validates it against what the kernels are specialised for, and hands the tensors to the C-ABI engine
(they are packed to bf16 there), so a serving process needs neither pytorch_lightning nor the nn.Module.

Host-side plumbing only: no arithmetic happens here.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .engine import _GLOBAL_KEYS, _LAYER_KEYS, T2SEngine

# what csrc/ is specialised for (t2s_create rejects anything else): s1 family, hidden 512 / 16 heads / FFN 2048 / vocab 1025
_FIXED = {"hidden_dim": 512, "embedding_dim": 512, "head": 16, "vocab_size": 1025, "EOS": 1024}


def _shapes(n_layer: int, d: int, vocab: int, phonemes: int) -> Dict[str, Tuple[int, ...]]:
    ff = 4 * d  # t2s_model.py:304 (dim_feedforward = hidden_dim * 4; the yaml's linear_units is ignored)
    s: Dict[str, Tuple[int, ...]] = {
        "bert_proj.weight": (d, 1024), "bert_proj.bias": (d,),
        "ar_text_embedding.word_embeddings.weight": (phonemes, d), "ar_text_position.alpha": (1,),
        "ar_audio_embedding.word_embeddings.weight": (vocab, d), "ar_audio_position.alpha": (1,),
        "ar_predict_layer.weight": (vocab, d),
    }
    per_layer = {
        "self_attn.in_proj_weight": (3 * d, d), "self_attn.in_proj_bias": (3 * d,),
        "self_attn.out_proj.weight": (d, d), "self_attn.out_proj.bias": (d,),
        "linear1.weight": (ff, d), "linear1.bias": (ff,), "linear2.weight": (d, ff), "linear2.bias": (d,),
        "norm1.weight": (d,), "norm1.bias": (d,), "norm2.weight": (d,), "norm2.bias": (d,),
    }
    for i in range(n_layer):
        for k, v in per_layer.items():
            s[f"h.layers.{i}.{k}"] = v
    return s


def read_checkpoint(path: str) -> Tuple[dict, Dict[str, torch.Tensor]]:
    """-> (config, state_dict without the ``model.`` prefix), validated.  Raises ValueError with the first problem found
    (the reference would fail inside load_state_dict with a size-mismatch message, TTS.py:594)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(ckpt, dict) or "weight" not in ckpt or "config" not in ckpt:
        raise ValueError(f"{path}: not an s1 checkpoint (expected a dict with 'weight' and 'config', process_ckpt.py:12-17)")
    config = ckpt["config"]
    model = dict(config.get("model", {}))
    for k, want in _FIXED.items():
        if int(model.get(k, -1)) != want:
            raise ValueError(f"{path}: config['model'][{k!r}] = {model.get(k)!r}; this build supports {want}")
    n_layer = int(model.get("n_layer", 0))
    phonemes = int(model.get("phoneme_vocab_size", 0))
    if not 1 <= n_layer <= 64 or phonemes < 1:
        raise ValueError(f"{path}: n_layer = {n_layer}, phoneme_vocab_size = {phonemes}")
    sd = {(k[6:] if k.startswith("model.") else k): v for k, v in ckpt["weight"].items()}
    for key, shape in _shapes(n_layer, 512, 1025, phonemes).items():
        if key not in sd:
            raise ValueError(f"{path}: weight {key!r} is missing")
        if tuple(sd[key].shape) != shape and sd[key].numel() != 1:
            raise ValueError(f"{path}: weight {key!r} has shape {tuple(sd[key].shape)}, expected {shape}")
    assert set(_GLOBAL_KEYS) <= set(sd) and all(f"h.layers.0.{k}" in sd for k in _LAYER_KEYS)
    return config, sd


def engine_from_checkpoint(path: str, device: str | torch.device = "cuda:0", max_batch: int = 256,
                           pe: Optional[torch.Tensor] = None) -> Tuple[T2SEngine, dict]:
    """What ``TTS.init_t2s_weights`` does, for this engine: -> (engine with the checkpoint's weights loaded, config).
    ``config["data"]["max_sec"]`` is what TTS.run turns into ``early_stop_num`` (TTS.py:592,1224)."""
    config, sd = read_checkpoint(path)
    eng = T2SEngine({"model": config["model"]}, device=device, max_batch=max_batch)
    eng.load_state_dict(sd, pe=pe)
    return eng, config
