"""TEST INFRASTRUCTURE — CPU restatement (numpy fp32) of the reference sampler.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this; the product path (gpt-sovits_b200/) never does.

Restates GPT_SoVITS/AR/models/utils.py:
  * logits_to_probs  :147-189  (repetition penalty -> top-p -> temperature -> top-k -> softmax)
  * multinomial_sample_one_no_sync :140-144 (argmax(probs / q), q ~ Exp(1))
  * sample :192-199

The reference draws q from torch's global generator, whose stream is not a stable contract
(SURVEY.md section 8c "Sampling parity"); as BASELINE.json's north_star states, seeded sampling is
checked against this fp32 replay, which draws q from the same counter-based Philox4x32-10 stream the
CUDA kernel uses: counter = (element//4, step, slot, 0), key = (seed_lo, seed_hi), word element%4,
u = ((w >> 9) + 0.5) * 2^-23 in (0,1), q = -log(u).

Pinned against the live reference by tests/golden/sampler_kat.npz (oracle/make_goldens.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10 (Salmon et al., Random123).  c* are uint32 arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0  # 32x32 -> 64, exact in uint64
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def exp_noise(seed: int, slot: int, step: int, n: int) -> np.ndarray:
    """q[n] ~ Exp(1) as fp32, element i from Philox counter (i//4, step, slot, 0), word i%4."""
    n4 = (n + 3) // 4
    blk = np.arange(n4, dtype=np.uint32)
    w = philox4x32_10(blk, np.full(n4, step, np.uint32), np.full(n4, slot, np.uint32),
                      np.zeros(n4, np.uint32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=1).reshape(-1)[:n]
    u = ((words >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    return (-np.log(u)).astype(np.float32)


def apply_repetition_penalty(logits: np.ndarray, previous_tokens, repetition_penalty: float) -> None:
    """In place, like the reference's scatter_ (utils.py:159-167): every distinct previous token is
    penalised once (gather happens before scatter, duplicates write identical values)."""
    if previous_tokens is None or repetition_penalty == 1.0:
        return
    prev = np.unique(np.asarray(previous_tokens, dtype=np.int64))
    if prev.size == 0:
        return
    rp = np.float32(repetition_penalty)
    score = logits[prev]
    logits[prev] = np.where(score < 0, score * rp, score / rp).astype(np.float32)


def _softmax(x: np.ndarray) -> np.ndarray:
    m = np.max(x)
    e = np.exp((x - m).astype(np.float32)).astype(np.float32)
    return (e / np.sum(e, dtype=np.float32)).astype(np.float32)


def logits_to_probs(
    logits: np.ndarray,
    previous_tokens=None,
    temperature: float = 1.0,
    top_k: Optional[int] = None,
    top_p: Optional[float] = None,
    repetition_penalty: float = 1.0,
) -> np.ndarray:
    """One row.  `logits` (fp32, 1-D) is penalised IN PLACE (the caller's later argmax sees it,
    t2s_model.py:721/:901); returns probs."""
    assert logits.dtype == np.float32 and logits.ndim == 1
    apply_repetition_penalty(logits, previous_tokens, repetition_penalty)
    x = logits.copy()
    if top_p is not None and top_p < 1.0:
        order = np.argsort(-x, kind="stable")  # descending
        sp = _softmax(x[order])
        cum = np.cumsum(sp, dtype=np.float32)
        remove_sorted = cum > np.float32(top_p)
        remove_sorted[0] = False  # keep at least one option (utils.py:173) -- no HF-style right shift
        remove = np.zeros_like(remove_sorted)
        remove[order] = remove_sorted
        x = np.where(remove, np.float32(-np.inf), x)
    x = (x / np.float32(max(temperature, 1e-5))).astype(np.float32)
    if top_k is not None:
        k = min(int(top_k), x.shape[0])
        pivot = np.sort(x)[-k]  # k-th largest, duplicates counted (torch.topk semantics)
        x = np.where(x < pivot, np.float32(-np.inf), x)
    return _softmax(x)


def sample_row(
    logits: np.ndarray,
    previous_tokens,
    q: np.ndarray,
    temperature: float = 1.0,
    top_k: Optional[int] = None,
    top_p: Optional[float] = None,
    repetition_penalty: float = 1.0,
) -> Tuple[int, int, np.ndarray, float]:
    """Returns (sampled token, argmax of the penalised logits, probs, relative race margin).

    margin = (best - second best) / best of probs/q; a replay comparison is only meaningful where it
    exceeds the fp32 rounding noise of exp/log on two different libms."""
    probs = logits_to_probs(logits, previous_tokens, temperature, top_k, top_p, repetition_penalty)
    greedy = int(np.argmax(logits))
    score = (probs / q[: probs.shape[0]]).astype(np.float32)
    tok = int(np.argmax(score))
    if score.shape[0] > 1:
        part = np.partition(score, -2)
        best, second = float(part[-1]), float(part[-2])
        margin = (best - second) / best if best > 0 else 0.0
    else:
        margin = 1.0
    return tok, greedy, probs, margin
