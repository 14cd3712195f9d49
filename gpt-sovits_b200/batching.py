"""Request batching either side of the T2S decode path.

* ``bucket_batches`` / ``recovery_order`` — the index logic of ``TTS.to_batch`` / ``TTS.recovery_order``
  (GPT_SoVITS/TTS_infer_pack/TTS.py:842-879, 957-973): sentences sorted by length and cut into batches whose median length is
  at least ``threshold`` x their mean, so that a batch wastes little on padding.  The B200 path keeps sequences ragged (padding
  never materialises), but callers that mirror ``TTS.run`` still want the same batches in the same order.
* ``StreamingSession`` — continuous batching + fragment return on top of ``T2SEngine``: the decode loop runs in slices
  (``t2s_decode`` with a step budget); after every slice the utterances that have stopped are handed out while the others keep
  decoding (the reference's ``return_fragment`` mode yields per batch only, TTS.py:1049-1053, 1319-1329), and waiting
  utterances are admitted into the free slots of the resident session (``t2s_admit``).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import EOS_WINDOW_BATCH, MAX_STEPS, T2SEngine


def bucket_batches(lengths: Sequence[int], batch_size: int = 5, threshold: float = 0.75, split_bucket: bool = True) -> List[List[int]]:
    """Indices of ``lengths`` grouped into batches exactly as ``TTS.to_batch`` groups sentences (TTS.py:859-886).

    split_bucket: sort by length (stable), then from the current position take up to ``batch_size`` items and shrink the
    window from the right until (length of the window's middle item) / (mean length) >= threshold or one item is left.
    Otherwise: consecutive groups of ``batch_size`` in input order.  Arithmetic is float32 like the reference's."""
    n = len(lengths)
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    if not split_bucket:
        return [list(range(i, min(i + batch_size, n))) for i in range(0, n, batch_size)]
    order = sorted(range(n), key=lambda i: int(lengths[i]))  # stable: ties keep input order
    lens = np.asarray([lengths[i] for i in order], dtype=np.int64)
    out: List[List[int]] = []
    pos = 0
    while pos < n:
        end = min(pos + batch_size, n)
        while True:
            window = lens[pos:end].astype(np.float32)
            score = window[(end - pos) // 2] / (window.mean() + 1e-8)
            if score >= threshold or end - pos == 1:
                break
            end -= 1
        out.append([order[i] for i in range(pos, end)])
        pos = end
    return out


def recovery_order(batched: Sequence[Sequence], batch_index_list: Sequence[Sequence[int]]) -> list:
    """Per-batch results back in the callers' original order (TTS.py:957-973)."""
    n = sum(len(b) for b in batch_index_list)
    out = [None] * n
    for items, idxs in zip(batched, batch_index_list):
        if len(items) != len(idxs):
            raise ValueError("a batch has %d results for %d indices" % (len(items), len(idxs)))
        for item, i in zip(items, idxs):
            out[i] = item
    return out


class StreamingSession:
    """One resident decode session that hands utterances out as they retire and takes new ones in as slots free up.

        sess = StreamingSession(engine, slots=32, slice_steps=25, top_k=15, ...)
        sess.submit(ids, bert, prompt)          # any number of times, also while iterating
        for key, tokens, idx in sess:           # tokens = prompt ++ kept tokens (int64, device), idx as the reference returns it
            ...

    ``key`` is the position of the utterance in submission order.  ``slots`` bounds the utterances decoding at the same time: a
    finished utterance's slot (and K/V pages) is released when it is handed out and reused by the next admission, so one resident
    session serves any number of utterances.  Utterances never interact and every Philox stream is keyed by (seed, key, own step),
    so an utterance gets the same tokens whatever slot it lands in and whenever it is admitted."""

    def __init__(self, engine: T2SEngine, slots: int = 32, positions: int = 0, slice_steps: int = 25, admit_min: int = 1, top_k: int = 15,
                 top_p: float = 1.0, temperature: float = 1.0, repetition_penalty: float = 1.35, early_stop_num: int = -1,
                 eos_suppress_steps: int = EOS_WINDOW_BATCH, max_steps: int = MAX_STEPS, seed: Optional[int] = None,
                 forced: Optional[torch.Tensor] = None, capture_logits: int = 0, hooks_by_key: int = 0):
        if slots < 1 or slice_steps < 1:
            raise ValueError("slots and slice_steps must be >= 1")
        self.eng, self.slots, self.positions, self.slice_steps = engine, int(slots), int(positions), int(slice_steps)
        # an admission is a prefill (launch-bound for a handful of rows: ~2 ms): waiting utterances are taken in once at least
        # `admit_min` slots are free (or as many as are waiting, or nothing is decoding), not one by one
        self.admit_min = max(1, int(admit_min))
        self.kw = dict(top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty,
                       early_stop_num=early_stop_num, eos_suppress_steps=eos_suppress_steps, max_steps=max_steps, seed=seed)
        # test hooks; hooks_by_key = n > 0: rows of `forced` / of the captured logits are indexed by utterance key (n keys), not by slot
        self.hooks = dict(forced=forced, capture_logits=capture_logits, hooks_by_utterance=int(hooks_by_key))
        self.waiting: List[Tuple[int, torch.Tensor, torch.Tensor, Optional[torch.Tensor]]] = []  # (key, ids, bert, prompt row)
        self.slot_key: Dict[int, int] = {}  # session slot -> key of the utterance decoding in it
        self.n_submitted = 0
        self.started = False
        self.first_logits = None
        self.active = 0

    def submit(self, phoneme_ids: Sequence[torch.Tensor], bert: Sequence[torch.Tensor], prompt: Optional[torch.Tensor]) -> List[int]:
        """Queues utterances (``infer_panel`` layout: lists + prompt [B, P] or None); returns their keys."""
        keys = []
        for b in range(len(phoneme_ids)):
            self.waiting.append((self.n_submitted, phoneme_ids[b], bert[b], None if prompt is None else prompt[b]))
            keys.append(self.n_submitted)
            self.n_submitted += 1
        return keys

    def _take(self, n: int):
        """Up to n waiting utterances that share the first one's prompt length (one request = one prompt length)."""
        if not self.waiting or n < 1:
            return None
        P0 = -1 if self.waiting[0][3] is None else int(self.waiting[0][3].shape[0])
        take, rest = [], []
        for w in self.waiting:
            P = -1 if w[3] is None else int(w[3].shape[0])
            (take if P == P0 and len(take) < n else rest).append(w)
        self.waiting = rest
        prompt = None if P0 < 0 else torch.stack([w[3] for w in take])
        return [w[0] for w in take], [w[1] for w in take], [w[2] for w in take], prompt

    def _admit_waiting(self) -> None:
        while self.waiting and len(self.slot_key) < self.slots:
            free = self.slots - len(self.slot_key)
            if self.slot_key and free < min(self.admit_min, len(self.waiting)):
                break
            got = self._take(free)
            if got is None:
                break
            keys, ids, bert, prompt = got
            if not self.started:
                r = self.eng.infer(ids, bert, prompt, max_new_steps=0, reserve_slots=self.slots, reserve_positions=self.positions,
                                   utt_ids=keys, **self.kw, **self.hooks)
                self.first_logits = r.logits
                self.started = True
                slots = list(range(len(keys)))
            else:
                slots = self.eng.admit(ids, bert, prompt, utt_ids=keys)
            for sl, k in zip(slots, keys):
                self.slot_key[sl] = k

    def __iter__(self) -> Iterator[Tuple[int, torch.Tensor, int]]:
        while True:
            self._admit_waiting()
            if not self.slot_key:
                return
            res = self.eng.session_result()
            finished = [sl for sl in self.slot_key if res.idx[sl] >= 0]
            if finished:
                seqs = res.sequences()
            for sl in finished:
                yield self.slot_key.pop(sl), seqs[sl].clone(), res.idx[sl]
            if finished:
                self.eng.release(finished)
                continue  # admit into the freed slots before the next slice
            self.eng.decode_more(self.slice_steps)
