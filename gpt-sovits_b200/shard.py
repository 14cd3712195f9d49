"""Multi-GPU utterance sharding for the T2S decode path.

Sequences never communicate (nothing in t2s_model.py:583-779 mixes batch rows), so the path shards by
utterance: one process per GPU, each with a full replica of the 152 MB bf16 weights and its own decode
stream; there is NO collective on the data path.  Only the final (token list, idx) pairs are gathered
and put back in the caller's order — the host-side counterpart of the reference's batch_idx_map
(t2s_model.py:699,735) and TTS.run's recovery_order (TTS.py:957-973).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_utterances(costs: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of utterance indices to `world` ranks.

    `costs[i]` estimates utterance i's decode work (e.g. phoneme count: the trained model emits ~2
    semantic tokens per phoneme, SURVEY.md section 8d "Length realism").  Ties keep input order so the
    partition is deterministic on every rank."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += int(costs[i]) + 1
    for p in parts:
        p.sort()
    return parts


def sharded_infer(
    infer_fn: Callable[[List[int]], Tuple[List[torch.Tensor], List[int]]],
    costs: Sequence[int],
    rank: Optional[int] = None,
    world: Optional[int] = None,
    group=None,
) -> Tuple[List[torch.Tensor], List[int]]:
    """Runs `infer_fn(my_indices)` on this rank's share and returns the full `(y_list, idx_list)` in the
    ORIGINAL utterance order on every rank.  `infer_fn` returns, for its indices in order, the token
    tensors (prompt ++ kept tokens) and idx values — e.g. a closure over
    ``Text2SemanticDecoder.infer_panel_batch_infer`` or ``T2SEngine.infer``.

    The only communication is one all_gather_object of the (short) token lists at the end."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    parts = partition_utterances(costs, world)
    mine = parts[rank]
    if mine:
        ys, idxs = infer_fn(mine)
        if len(ys) != len(mine) or len(idxs) != len(mine):
            raise RuntimeError("infer_fn returned %d/%d results for %d utterances" % (len(ys), len(idxs), len(mine)))
        payload = [(i, y.detach().cpu().tolist(), int(k)) for i, y, k in zip(mine, ys, idxs)]
    else:
        payload = []
    if world > 1:
        gathered: List[Optional[list]] = [None] * world
        dist.all_gather_object(gathered, payload, group=group)
    else:
        gathered = [payload]
    n = len(costs)
    y_list: List[Optional[torch.Tensor]] = [None] * n
    idx_list: List[Optional[int]] = [None] * n
    for part in gathered:
        for i, toks, k in part:
            y_list[i] = torch.tensor(toks, dtype=torch.int64)
            idx_list[i] = k
    missing = [i for i in range(n) if y_list[i] is None]
    if missing:
        raise RuntimeError(f"utterances {missing} were not produced by any rank")
    return y_list, idx_list  # type: ignore[return-value]
