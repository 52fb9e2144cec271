"""Host-side KV-cache index arithmetic of the rolling / sink window.

Integer restatement of wan/modules/causal_model.py:195-236 (CausalWanSelfAttention.forward, cache
branch).  The reference reads `global_end_index` / `local_end_index` back from the device with
4-6 `.item()` syncs per layer per forward; here the values are mirrored on the host (they are a
pure function of the call sequence) and the device tensors in the cache dict are still updated
so any other consumer of the dict sees the reference's state.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class CachePlan:
    roll: bool
    roll_src: int
    roll_dst: int
    roll_len: int
    write_start: int
    write_end: int
    attn_start: int
    attn_end: int
    global_end: int
    local_end: int


def plan_cache_update(global_end: int, local_end: int, current_start: int, num_new: int, cache_size: int,
                      local_attn_size: int, sink_tokens: int, max_attention_size: int) -> CachePlan:
    """One self-attention call's effect on one layer's cache.

    causal_model.py:202      current_end = current_start + num_new
    causal_model.py:207-208  roll iff local attention, the chunk is new (current_end > global_end) and it
                             does not fit (num_new + local_end > cache_size)
    causal_model.py:212-221  evict the oldest tokens after the sink, shift the rest left, append
    causal_model.py:226-229  otherwise write at local_end + (current_end - global_end) - num_new
                             (re-denoising the same chunk overwrites it in place)
    causal_model.py:230-234  attend to the last max_attention_size tokens ending at the new local end
    """
    current_end = current_start + num_new
    if local_attn_size != -1 and current_end > global_end and num_new + local_end > cache_size:
        evicted = num_new + local_end - cache_size
        rolled = local_end - evicted - sink_tokens
        new_local_end = local_end + current_end - global_end - evicted
        roll, src, dst, rlen = True, sink_tokens + evicted, sink_tokens, rolled
    else:
        new_local_end = local_end + current_end - global_end
        roll, src, dst, rlen = False, 0, 0, 0
    if new_local_end - num_new < 0 or new_local_end > cache_size:
        raise ValueError(
            f"KV cache write [{new_local_end - num_new}, {new_local_end}) outside the cache of {cache_size} tokens "
            f"(global_end={global_end}, local_end={local_end}, current_start={current_start}, num_new={num_new})")
    return CachePlan(roll, src, dst, rlen, new_local_end - num_new, new_local_end,
                     max(0, new_local_end - max_attention_size), new_local_end, current_end, new_local_end)


class IndexMirror:
    """Host copy of (global_end_index, local_end_index) for every layer's cache dict.

    The pipeline resets a cache by *rebinding* fresh index tensors (pipeline/causal_inference.py:125-132),
    so the mirror is keyed on tensor identity: a tensor we have not written ourselves is read back once
    (one batched device->host copy for all layers), everything after that is host arithmetic.  An in-place edit by
    the caller (`.zero_()`, `.fill_(0)` -- the reference would see it through `.item()`) bumps the tensors'
    `_version` counters, which the entry records: a changed version is re-read from the device like a rebind.
    """

    def __init__(self):
        self._known: Dict[int, tuple] = {}   # id(global tensor) -> (global, local, g, l, global._version, local._version)

    def read(self, kv_cache: List[dict]) -> List[Tuple[int, int]]:
        out: List[Tuple[int, int]] = [None] * len(kv_cache)  # type: ignore
        unknown = []
        for i, c in enumerate(kv_cache):
            g, l = c["global_end_index"], c["local_end_index"]
            hit = self._known.get(id(g))
            if hit is not None and hit[0] is g and hit[1] is l and hit[4] == g._version and hit[5] == l._version:
                out[i] = (hit[2], hit[3])
            else:
                unknown.append(i)
        if unknown:
            if len(self._known) > 4096:   # stale entries of rebound tensors
                self._known.clear()
            vals = torch.stack([torch.cat([kv_cache[i]["global_end_index"].reshape(1),
                                           kv_cache[i]["local_end_index"].reshape(1)]) for i in unknown]).cpu()
            for row, i in zip(vals.tolist(), unknown):
                c = kv_cache[i]
                self._remember(c, int(row[0]), int(row[1]))
                out[i] = (int(row[0]), int(row[1]))
        return out

    def write(self, kv_cache: List[dict], values: List[Tuple[int, int]]) -> None:
        """Store the new indices into the dict's device tensors (causal_model.py:235-236) and the mirror."""
        groups: Dict[int, List[torch.Tensor]] = {}
        for i, (c, (g, l)) in enumerate(zip(kv_cache, values)):
            groups.setdefault(g, []).append(c["global_end_index"])
            groups.setdefault(l, []).append(c["local_end_index"])
        for value, tensors in groups.items():
            src = torch.full((1,), value, dtype=tensors[0].dtype, device=tensors[0].device)
            torch._foreach_copy_(tensors, [src] * len(tensors))
        for c, (g, l) in zip(kv_cache, values):   # after the copies: they bump the version counters
            self._remember(c, g, l)

    def _remember(self, c: dict, g: int, l: int) -> None:
        gt, lt = c["global_end_index"], c["local_end_index"]
        self._known[id(gt)] = (gt, lt, g, l, gt._version, lt._version)
