"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's training-time (cache-free) forward.

    CausalWanModel._forward_train                   wan/modules/causal_model.py:895-1069
    CausalWanSelfAttention.forward, kv_cache=None   wan/modules/causal_model.py:119-193  (FlexAttention + BlockMask)
    mask builders                                   wan/modules/causal_model.py:518-723  (restated in causal_wan_oracle)

All frames of a video go through the network at once; causality comes from the block mask: a query token of chunk c
sees the keys before the end of its chunk (optionally only the last `local_attn_size` frames), and under teacher
forcing the sequence is [clean video | noisy video], a noisy chunk seeing the CLEAN chunks before it plus itself.
RoPE restarts at frame 0 for the noisy half (:127-136).  The head runs on the noisy half only (:1058-1062).

Attention arithmetic: the reference calls torch's `flex_attention`; its eager ("math") form multiplies Q and K in
float32, scales and masks the scores in float32, takes a float32 softmax and multiplies the probabilities, cast back to
the input dtype, with V (torch 2.11 torch/_higher_order_ops/flex_attention.py: math_attention).  That is restated here.  The
reference wraps flex_attention in torch.compile(mode="max-autotune-no-cudagraphs") (:24-25), which Inductor cannot
lower on a CPU; the pinning run (oracle/make_golden.py) rebinds it to the uncompiled torch op -- the third-party kernel
is replaced, the reference code is untouched.

Parity status: PINNED against the unmodified reference (tests/golden/train_forward_tiny.pt, tests/test_train_forward.py).
Only tests/ may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import causal_wan_oracle as O

Tensor = torch.Tensor


def masked_attention(q: Tensor, k: Tensor, v: Tensor, mask: Tensor) -> Tensor:
    """q, k, v [B, L, H, D]; mask [Lp, Lp] bool over the length padded to x128 (True = attend).  Restates the call site
    (:137-160) literally -- zero-pad to x128, flex_attention's eager arithmetic, cut the padding off -- because the
    result is a VIEW whose batch stride is the padded length, and the CPU bf16 GEMM of the following `o` projection
    rounds differently for that layout than for a compact copy (bit-exact pinning needs the same layout)."""
    B, L, H, D = q.shape
    pad = mask.shape[0] - L
    assert pad > 0, "the reference slices [:-padded_length]: a length that is a multiple of 128 would return nothing"
    qh, kh, vh = (torch.cat([t, t.new_zeros(B, pad, H, D)], dim=1).transpose(2, 1) for t in (q, k, v))
    scores = qh.to(torch.float32) @ kh.to(torch.float32).transpose(-2, -1)
    scores = scores * (1.0 / (D ** 0.5))
    scores = torch.where(mask.to(scores.device)[None, None], scores, torch.full_like(scores, float("-inf")))
    probs = torch._safe_softmax(scores, dim=-1)
    out = probs.to(q.dtype) @ vh                                    # [B, H, Lp, D]
    out = out.transpose(1, 2).contiguous().transpose(1, 2)          # flex_attention returns the query's memory layout
    return out[:, :, :-pad].transpose(2, 1)                          # [B, L, H, D] view, batch stride Lp * H * D


def attention_mask(num_frames: int, frame_tokens: int, num_frame_per_block: int, local_attn_size: int = -1,
                   independent_first_frame: bool = False, teacher_forcing: bool = False) -> Tensor:
    """Dense [Lp, Lp] mask of the forward (the reference's mask_mod over the length padded to x128)."""
    if teacher_forcing:
        if independent_first_frame:
            raise NotImplementedError("reference raises for teacher forcing with an independent first frame (:941-942)")
        return O.teacher_forcing_mask(num_frames, frame_tokens, num_frame_per_block)
    return O.blockwise_causal_mask(num_frames, frame_tokens, num_frame_per_block, local_attn_size, independent_first_frame)


def train_forward(p: Dict[str, Tensor], cfg: O.OracleConfig, x: Tensor, t: Tensor, context: Tensor,
                  num_frame_per_block: int = 1, clean_x: Optional[Tensor] = None, aug_t: Optional[Tensor] = None,
                  independent_first_frame: bool = False) -> Tensor:
    """x [B, 16, F, H, W], t [B, F] (per-frame timesteps), context [B, <=512, text_dim]; clean_x like x switches
    teacher forcing on (aug_t = timesteps of the clean half, zeros by default).  -> flow [B, 16, F, H, W]."""
    B, _, nf, Hh, Ww = x.shape
    pt, ph, pw = cfg.patch_size
    grid = (nf // pt, Hh // ph, Ww // pw)
    ft = grid[1] * grid[2]
    angles = O.rope_angle_table(cfg.head_dim)
    H, D = cfg.num_heads, cfg.head_dim

    def embed(u):
        return F.conv3d(u, p["patch_embedding.weight"], p["patch_embedding.bias"], stride=cfg.patch_size).flatten(2).transpose(1, 2)

    def time_embed(tt):
        e = O.sinusoid_embed(cfg.freq_dim, tt.flatten()).type_as(tok)
        e = O._lin(p, "time_embedding.2", F.silu(O._lin(p, "time_embedding.0", e)))
        e0 = O._lin(p, "time_projection.1", F.silu(e)).unflatten(1, (6, cfg.dim)).unflatten(0, tt.shape)
        return e, e0

    tok = embed(x)                                                                   # [B, L, C]
    L = tok.shape[1]
    e, e0 = time_embed(t)
    if context.shape[1] < cfg.text_len:
        context = torch.cat([context, context.new_zeros(B, cfg.text_len - context.shape[1], context.shape[2])], 1)
    ctx = O._lin(p, "text_embedding.2", F.gelu(O._lin(p, "text_embedding.0", context), approximate="tanh"))
    tf = clean_x is not None
    if tf:
        tok = torch.cat([embed(clean_x), tok], dim=1)                                # :1020
        if aug_t is None:
            aug_t = torch.zeros_like(t)
        _, e0_clean = time_embed(aug_t)
        e0 = torch.cat([e0_clean, e0], dim=1)                                        # :1027
    mask = attention_mask(nf, ft, num_frame_per_block, cfg.local_attn_size, independent_first_frame, tf)
    nfr = e0.shape[1]                                                                # frames of modulation (F or 2F)

    def per_frame(u):
        return u.unflatten(1, (nfr, ft))

    for i in range(cfg.num_layers):
        pre = f"blocks.{i}."
        em = (p[pre + "modulation"].unsqueeze(1) + e0).chunk(6, dim=2)
        h = (per_frame(O.layer_norm(tok, None, None, cfg.eps)) * (1 + em[1]) + em[0]).flatten(1, 2)
        sa = pre + "self_attn."
        q = O.rms_norm(O._lin(p, sa + "q", h), p[sa + "norm_q.weight"], cfg.eps).view(B, -1, H, D)
        k = O.rms_norm(O._lin(p, sa + "k", h), p[sa + "norm_k.weight"], cfg.eps).view(B, -1, H, D)
        v = O._lin(p, sa + "v", h).view(B, -1, H, D)
        if tf:   # the clean and the noisy half carry the same positions (:127-136)
            q = torch.cat([O.rope_rotate(c, grid, angles, 0).type_as(v) for c in q.chunk(2, dim=1)], dim=1)
            k = torch.cat([O.rope_rotate(c, grid, angles, 0).type_as(v) for c in k.chunk(2, dim=1)], dim=1)
        else:
            q, k = O.rope_rotate(q, grid, angles, 0).type_as(v), O.rope_rotate(k, grid, angles, 0).type_as(v)
        y = O._lin(p, sa + "o", masked_attention(q, k, v, mask).flatten(2))
        tok = tok + (per_frame(y) * em[2]).flatten(1, 2)
        tok = tok + O.cross_attention(p, pre + "cross_attn.",
                                      O.layer_norm(tok, p[pre + "norm3.weight"], p[pre + "norm3.bias"], cfg.eps), ctx, None, cfg)
        h = (per_frame(O.layer_norm(tok, None, None, cfg.eps)) * (1 + em[4]) + em[3]).flatten(1, 2)
        y = O._lin(p, pre + "ffn.2", F.gelu(O._lin(p, pre + "ffn.0", h), approximate="tanh"))
        tok = tok + (per_frame(y) * em[5]).flatten(1, 2)

    if tf:
        tok = tok[:, tok.shape[1] // 2:]                                             # :1058-1059
    eh = (p["head.modulation"].unsqueeze(1) + e.unflatten(0, t.shape).unsqueeze(2)).chunk(2, dim=2)
    y = O.layer_norm(tok, None, None, cfg.eps).unflatten(1, (nf, ft)) * (1 + eh[1]) + eh[0]
    y = O._lin(p, "head.head", y)
    c = cfg.out_dim
    y = y.reshape(B, grid[0], grid[1], grid[2], pt, ph, pw, c)
    return y.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, c, grid[0] * pt, grid[1] * ph, grid[2] * pw)
