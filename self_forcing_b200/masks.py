"""Block-mask tables of the three mask builders of CausalWanModel, computed analytically.

The reference builds FlexAttention `BlockMask`s for its training-side forward
(wan/modules/causal_model.py:518-574 block-wise causal, :576-662 teacher forcing, :664-723 i2v)
by evaluating a `mask_mod` on a dense [Lp, Lp] grid (Lp = length padded to x128; 4 s / 29 s on CPU
for the 21-frame masks).  The rollout itself never needs them (the KV window *is* the mask), but
the tables are part of the parity contract, so they are produced here bit-exactly from interval
arithmetic in O(blocks^2 * 128) without materialising the dense mask:

    every query row q is allowed a small set of disjoint half-open KV intervals (+ the diagonal),
    a 128x128 tile is *present* if any row's interval set touches it (or it is a diagonal tile),
    and *full* if every row's intervals cover all 128 columns.

Output follows torch's BlockMask convention (`_dense_to_ordered`): for each query block the
partial-tile count and the row of KV block indices (non-empty first, ascending, then the rest
ascending), and the same for full tiles.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

BLOCK = 128


def _pad(n: int) -> int:
    return (n + BLOCK - 1) // BLOCK * BLOCK


def _ordered(dense: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(count, indices) per row like torch.nn.attention.flex_attention._dense_to_ordered."""
    num = dense.sum(axis=1).astype(np.int32)
    idx = np.argsort(-dense.astype(np.int8), axis=1, kind="stable").astype(np.int32)
    return num, idx


def _tables(intervals: List[Tuple[np.ndarray, np.ndarray]], n: int) -> Dict[str, np.ndarray]:
    """intervals: list of (lo[n], hi[n]) -- row q may attend kv in the union of [lo_i[q], hi_i[q])
    (disjoint per row) plus kv == q."""
    nb = n // BLOCK
    starts = (np.arange(nb) * BLOCK)[None, None, :]          # [1, 1, nb]
    ends = starts + BLOCK
    touch = np.zeros((nb, BLOCK, nb), dtype=bool)
    covered = np.zeros((nb, BLOCK, nb), dtype=np.int32)
    for lo, hi in intervals:
        lo3 = lo.reshape(nb, BLOCK, 1)
        hi3 = hi.reshape(nb, BLOCK, 1)
        ov = np.minimum(hi3, ends) - np.maximum(lo3, starts)
        ov = np.maximum(ov, 0)
        touch |= ov > 0
        covered += ov
    # the diagonal element (q == kv) adds one column to rows whose intervals miss it
    q = np.arange(n).reshape(nb, BLOCK, 1)
    diag_in = np.zeros((nb, BLOCK, 1), dtype=bool)
    for lo, hi in intervals:
        diag_in |= (lo.reshape(nb, BLOCK, 1) <= q) & (q < hi.reshape(nb, BLOCK, 1))
    eye = np.eye(nb, dtype=bool)[:, None, :]                  # tile (qb, kb) holds the diagonal iff qb == kb
    covered = covered + (eye & ~diag_in)
    touch = touch | eye
    any_ = touch.any(axis=1)
    all_ = (covered >= BLOCK).all(axis=1)
    partial = any_ & ~all_
    kv_num, kv_idx = _ordered(partial)
    full_num, full_idx = _ordered(all_)
    return dict(kv_num_blocks=kv_num, kv_indices=kv_idx, full_kv_num_blocks=full_num, full_kv_indices=full_idx,
                sparsity=100.0 * (1.0 - any_.sum() / any_.size))


def _chunk_ends(num_frames: int, frame_seqlen: int, num_frame_per_block: int, lone_first_frame: bool) -> np.ndarray:
    total = num_frames * frame_seqlen
    ends = np.zeros(_pad(total), dtype=np.int64)
    blk = frame_seqlen * num_frame_per_block
    first = 0
    if lone_first_frame:
        ends[:frame_seqlen] = frame_seqlen
        first = frame_seqlen
    for s in range(first, total, blk):
        ends[s:s + blk] = s + blk      # numpy clips the slice at the padded length like torch
    return ends


def blockwise_causal_tables(num_frames: int = 21, frame_seqlen: int = 1560, num_frame_per_block: int = 1,
                            local_attn_size: int = -1, lone_first_frame: bool = False) -> Dict[str, np.ndarray]:
    """causal_model.py:518-574 (and the i2v variant :664-723 with lone_first_frame=True):
    kv < ends[q] (and kv >= ends[q] - local_attn_size*frame_seqlen), or kv == q."""
    ends = _chunk_ends(num_frames, frame_seqlen, num_frame_per_block, lone_first_frame)
    lo = np.zeros_like(ends) if local_attn_size == -1 else np.maximum(ends - local_attn_size * frame_seqlen, 0)
    return _tables([(lo, ends)], ends.shape[0])


def teacher_forcing_tables(num_frames: int = 21, frame_seqlen: int = 1560,
                           num_frame_per_block: int = 1) -> Dict[str, np.ndarray]:
    """causal_model.py:576-662: sequence = [clean frames | noisy frames]; clean rows see clean tokens up
    to the end of their chunk, noisy rows see their own noisy chunk plus the clean chunks before it."""
    half = num_frames * frame_seqlen
    total = 2 * half
    n = _pad(total)
    blk = frame_seqlen * num_frame_per_block
    lo1 = np.zeros(n, dtype=np.int64)
    hi1 = np.zeros(n, dtype=np.int64)   # clean rows: [0, chunk end) ; noisy rows: [0, clean context end)
    lo2 = np.zeros(n, dtype=np.int64)
    hi2 = np.zeros(n, dtype=np.int64)   # noisy rows: [chunk start, chunk end)
    for s in range(0, half, blk):
        hi1[s:s + blk] = s + blk
    for bi, s in enumerate(range(half, total, blk)):
        hi1[s:s + blk] = bi * blk
        lo2[s:s + blk] = s
        hi2[s:s + blk] = s + blk
    return _tables([(lo1, hi1), (lo2, hi2)], n)
