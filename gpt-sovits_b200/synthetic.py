"""Deterministic synthetic weights and inputs of the s1-v2 text-to-semantic architecture.

There are no pretrained checkpoints and no network, so every parity test and benchmark runs on
random-init weights (BASELINE.json prescribes this).  PyTorch's default init makes the 24-layer
post-LN stack collapse to an input-independent output ("Pitfall 1"), which would
hide KV-cache / mask / position bugs, so the "sensitive init" recommended there is used:

  * normal weights with std g/sqrt(fan_in): Q,K rows of in_proj g=gqk (2.0); V rows, linear1,
    ar_predict_layer g=1; out_proj / linear2 g=gres (0.2);
  * embeddings, biases, LayerNorm, bert_proj at PyTorch-default distributions;
  * every tensor is rounded to bf16 and handed out as fp32, so the reference (fp32 compute) and the
    CUDA engine (bf16 storage) see bit-identical parameter values.

numpy's legacy RandomState is used because its stream is guaranteed stable across versions and
machines; the goldens under tests/golden/ were produced from exactly these tensors.

Parameter names/shapes follow the reference state_dict (GPT_SoVITS/AR/models/t2s_model.py:260-353).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

S1V2_CONFIG = {
    "model": {
        "hidden_dim": 512,
        "embedding_dim": 512,
        "head": 16,
        "n_layer": 24,
        "vocab_size": 1025,
        "phoneme_vocab_size": 732,
        "dropout": 0.0,
        "EOS": 1024,
    }
}
BERT_DIM = 1024
PE_LEN = 4000


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def sine_pe(n: int = PE_LEN, dim: int = 512) -> torch.Tensor:
    """Sinusoidal table exactly as GPT_SoVITS/AR/modules/embedding.py:54-72 builds it (fp32 torch ops)."""
    pe = torch.zeros(n, dim)
    position = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def make_state_dict(
    seed: int = 0,
    config: Optional[dict] = None,
    gqk: float = 2.0,
    gres: float = 0.2,
    eos_scale: float = 1.0,
    eos_bias_dir: float = 0.0,
    rounding: str = "bf16",
) -> Dict[str, torch.Tensor]:
    """fp32 tensors holding bf16-representable values, keyed like ``Text2SemanticDecoder.state_dict()``.

    ``eos_scale`` multiplies the EOS row of ``ar_predict_layer.weight`` (0 forbids EOS for throughput
    runs: its logit is then exactly 0 and never wins at logit sigma ~ 1).  ``rounding`` = "bf16" (default: values the
    engine stores exactly) or "fp16" (what a real s1 checkpoint holds, TTS.py:598-599: fp16-representable values that
    the engine has to round to bf16)."""
    cfg = (config or S1V2_CONFIG)["model"]
    d, L = cfg["hidden_dim"], cfg["n_layer"]
    ff = 4 * d  # t2s_model.py:304 (dim_feedforward = hidden*4; the yaml's linear_units is ignored)
    V, PV = cfg["vocab_size"], cfg["phoneme_vocab_size"]
    rs = np.random.RandomState(seed)

    def normal(shape, std):
        return torch.from_numpy((rs.standard_normal(size=shape) * std).astype(np.float32))

    def uniform(shape, bound):
        return torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))

    sd: Dict[str, torch.Tensor] = {}
    sd["bert_proj.weight"] = uniform((d, BERT_DIM), 1.0 / math.sqrt(BERT_DIM))
    sd["bert_proj.bias"] = uniform((d,), 1.0 / math.sqrt(BERT_DIM))
    sd["ar_text_embedding.word_embeddings.weight"] = normal((PV, d), 1.0)
    sd["ar_text_position.alpha"] = torch.ones(1)
    sd["ar_audio_embedding.word_embeddings.weight"] = normal((V, d), 1.0)
    sd["ar_audio_position.alpha"] = torch.ones(1)
    for i in range(L):
        p = f"h.layers.{i}."
        w = normal((3 * d, d), 1.0 / math.sqrt(d))
        w[: 2 * d] *= gqk
        sd[p + "self_attn.in_proj_weight"] = w
        sd[p + "self_attn.in_proj_bias"] = torch.zeros(3 * d)
        sd[p + "self_attn.out_proj.weight"] = normal((d, d), gres / math.sqrt(d))
        sd[p + "self_attn.out_proj.bias"] = torch.zeros(d)
        sd[p + "linear1.weight"] = normal((ff, d), 1.0 / math.sqrt(d))
        sd[p + "linear1.bias"] = uniform((ff,), 1.0 / math.sqrt(d))
        sd[p + "linear2.weight"] = normal((d, ff), gres / math.sqrt(ff))
        sd[p + "linear2.bias"] = uniform((d,), 1.0 / math.sqrt(ff))
        sd[p + "norm1.weight"] = torch.ones(d)
        sd[p + "norm1.bias"] = torch.zeros(d)
        sd[p + "norm2.weight"] = torch.ones(d)
        sd[p + "norm2.bias"] = torch.zeros(d)
    wp = normal((V, d), 1.0 / math.sqrt(d))
    wp[cfg["EOS"]] *= eos_scale
    if eos_bias_dir != 0.0:
        # push the EOS row along the mean hidden direction so EOS fires "naturally" (config 3)
        wp[cfg["EOS"]] += eos_bias_dir / math.sqrt(d)
    sd["ar_predict_layer.weight"] = wp
    if rounding == "fp16":
        return {k: v.to(torch.float16).to(torch.float32).contiguous() for k, v in sd.items()}
    assert rounding == "bf16", rounding
    return {k: bf16_round(v).contiguous() for k, v in sd.items()}


def make_inputs(
    batch: int,
    phoneme_lens: List[int],
    prompt_len: int,
    seed: int = 0,
    bert_zero: bool = False,
    phoneme_vocab: int = 732,
    prompt_vocab: int = 1024,
) -> Tuple[List[torch.Tensor], torch.Tensor, Optional[torch.Tensor], List[torch.Tensor]]:
    """(all_phoneme_ids, all_phoneme_lens, prompt[B,P]|None, all_bert_features) in TTS.run's layout
    (GPT_SoVITS/TTS_infer_pack/TTS.py:1215-1227).  BERT features are bf16-representable fp32."""
    assert len(phoneme_lens) == batch
    rs = np.random.RandomState(1000 + seed)
    ids = [torch.from_numpy(rs.randint(0, phoneme_vocab, size=(n,)).astype(np.int64)) for n in phoneme_lens]
    if bert_zero:
        bert = [torch.zeros(BERT_DIM, n) for n in phoneme_lens]
    else:
        bert = [bf16_round(torch.from_numpy(rs.standard_normal(size=(BERT_DIM, n)).astype(np.float32)))
                for n in phoneme_lens]
    if prompt_len > 0:
        # TTS.run hands every item the same reference-audio prompt (an .expand view, TTS.py:1210-1212)
        row = torch.from_numpy(rs.randint(0, prompt_vocab, size=(prompt_len,)).astype(np.int64))
        prompt = row.unsqueeze(0).expand(batch, -1)
    else:
        prompt = None
    lens = torch.tensor(phoneme_lens, dtype=torch.int64)
    return ids, lens, prompt, bert


def config_lens(batch: int, lo: int, hi: int, seed: int = 0) -> List[int]:
    """Phoneme lengths ~ U{lo..hi}, deterministic (BASELINE.json configs 2-5)."""
    rs = np.random.RandomState(2000 + seed)
    return [int(v) for v in rs.randint(lo, hi + 1, size=(batch,))]
