"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Self-Forcing rollout hot path.

This is the *oracle* for the B200 kernels: a plain, functional PyTorch restatement (no
nn.Module graph, weights in a flat dict keyed like the reference `state_dict`) of

    CausalInferencePipeline.inference      pipeline/causal_inference.py:47-276
    WanDiffusionWrapper.forward            utils/wan_wrapper.py:253-349
    FlowMatchScheduler tables / add_noise  utils/scheduler.py:118-176
    CausalWanModel._forward_inference      wan/modules/causal_model.py:725-893
    CausalWanAttentionBlock / SelfAttention / Head   causal_model.py:59-367
    WanT2VCrossAttention, WanRMSNorm, WanLayerNorm, rope, sinusoid   wan/modules/model.py:15-194
    block-mask tables                      causal_model.py:518-723

It follows the reference's arithmetic and *rounding points* (every intermediate lives in the
weights' dtype, normally bf16; norms, softmax and the sampler use the same wider types the
reference uses), so on the same seeded inputs it reproduces the reference to the last bit
on CPU for everything except RoPE, which is restated in real float64 arithmetic instead of
complex128 (identical after the bf16 rounding that follows).

Parity status: PINNED against the unmodified reference executed in the build container
(`tests/test_oracle_vs_reference.py`, golden vectors in `tests/golden/` made by
`oracle/make_golden.py`).  The reference itself ships no golden vectors or numeric tests for
this path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product (`self_forcing_b200/`) never does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration and synthetic weights
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Hyper-parameters of CausalWanModel (causal_model.py:382-398)."""
    dim: int = 1536
    ffn_dim: int = 8960
    num_heads: int = 12
    num_layers: int = 30
    in_dim: int = 16
    out_dim: int = 16
    text_dim: int = 4096
    text_len: int = 512
    freq_dim: int = 256
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    eps: float = 1e-6
    local_attn_size: int = -1
    sink_size: int = 0
    # the reference hard-codes 1560 tokens/frame in max_attention_size (causal_model.py:77)
    frame_tokens_for_window: int = 1560

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads

    @property
    def max_attention_size(self) -> int:
        return 32760 if self.local_attn_size == -1 else self.local_attn_size * self.frame_tokens_for_window

    def reference_kwargs(self) -> dict:
        return dict(model_type="t2v", patch_size=tuple(self.patch_size), text_len=self.text_len,
                    in_dim=self.in_dim, dim=self.dim, ffn_dim=self.ffn_dim, freq_dim=self.freq_dim,
                    text_dim=self.text_dim, out_dim=self.out_dim, num_heads=self.num_heads,
                    num_layers=self.num_layers, local_attn_size=self.local_attn_size,
                    sink_size=self.sink_size, qk_norm=True, cross_attn_norm=True, eps=self.eps)


WAN_1_3B = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30)
WAN_TINY = dict(dim=1536, ffn_dim=512, num_heads=12, num_layers=2)


def parameter_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Names/shapes of the CausalWanModel t2v state_dict (causal_model.py:457-503)."""
    C, Fd = cfg.dim, cfg.ffn_dim
    pt, ph, pw = cfg.patch_size
    s: Dict[str, Tuple[int, ...]] = {
        "patch_embedding.weight": (C, cfg.in_dim, pt, ph, pw),
        "patch_embedding.bias": (C,),
        "text_embedding.0.weight": (C, cfg.text_dim), "text_embedding.0.bias": (C,),
        "text_embedding.2.weight": (C, C), "text_embedding.2.bias": (C,),
        "time_embedding.0.weight": (C, cfg.freq_dim), "time_embedding.0.bias": (C,),
        "time_embedding.2.weight": (C, C), "time_embedding.2.bias": (C,),
        "time_projection.1.weight": (6 * C, C), "time_projection.1.bias": (6 * C,),
        "head.head.weight": (cfg.out_dim * pt * ph * pw, C),
        "head.head.bias": (cfg.out_dim * pt * ph * pw,),
        "head.modulation": (1, 2, C),
    }
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        for attn in ("self_attn", "cross_attn"):
            for lin in ("q", "k", "v", "o"):
                s[p + f"{attn}.{lin}.weight"] = (C, C)
                s[p + f"{attn}.{lin}.bias"] = (C,)
            s[p + f"{attn}.norm_q.weight"] = (C,)
            s[p + f"{attn}.norm_k.weight"] = (C,)
        s[p + "norm3.weight"] = (C,)
        s[p + "norm3.bias"] = (C,)
        s[p + "ffn.0.weight"] = (Fd, C)
        s[p + "ffn.0.bias"] = (Fd,)
        s[p + "ffn.2.weight"] = (C, Fd)
        s[p + "ffn.2.bias"] = (C,)
        s[p + "modulation"] = (1, 6, C)
    return s


def make_random_params(cfg: OracleConfig, seed: int = 0, dtype=torch.bfloat16,
                       device="cpu") -> Dict[str, Tensor]:
    """Seeded synthetic weights of the named architecture (no checkpoint / network).

    Matmul weights ~ N(0, 2/(fan_in+fan_out)) (the variance of the reference's xavier init,
    causal_model.py:1111-1116); biases and the head weight ~ N(0, .02); norm affines
    1 + N(0, .02); modulation ~ N(0, 1/dim) (causal_model.py:282).  The reference zero-inits
    biases / head weight, which would make parity vacuous.  Generated
    on CPU one tensor at a time from a per-name seed so any subset is reproducible.
    """
    out: Dict[str, Tensor] = {}
    for idx, (name, shape) in enumerate(parameter_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        if name.endswith("modulation"):
            w = torch.randn(shape, generator=g) / math.sqrt(cfg.dim)
        elif name.endswith(".bias") or name == "head.head.weight":
            w = torch.randn(shape, generator=g) * 0.02
        elif "norm" in name:
            w = 1.0 + torch.randn(shape, generator=g) * 0.02
        else:
            fan_out = shape[0]
            fan_in = int(math.prod(shape[1:]))
            w = torch.randn(shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))
        out[name] = w.to(dtype).to(device)
    return out


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def sinusoid_embed(freq_dim: int, t_flat: Tensor) -> Tensor:
    """model.py:15-25 -- float64 table cat(cos, sin)(t * 10000^(-i/half))."""
    half = freq_dim // 2
    pos = t_flat.to(torch.float64)
    inv = torch.pow(10000, -torch.arange(half).to(pos).div(half))
    ang = torch.outer(pos, inv)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def rope_angle_table(head_dim: int, max_pos: int = 1024, theta: float = 10000.0) -> Tensor:
    """Angles (float64) [max_pos, head_dim/2] of the 3-axis table (model.py:29-36,
    causal_model.py:482-488): the head's complex pairs are split (frame, h, w) with widths
    d - 4*(d//6), 2*(d//6), 2*(d//6) real dims, each its own theta ladder."""
    d = head_dim
    parts = []
    for dim in (d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)):
        inv = 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim))
        parts.append(torch.outer(torch.arange(max_pos), inv))  # reference: int64 outer f64
    return torch.cat(parts, dim=1)


def rope_rotate(x: Tensor, grid: Tuple[int, int, int], angles: Tensor, start_frame: int) -> Tensor:
    """causal_model.py:28-56 -- rotate interleaved pairs (2i, 2i+1) of every head by the angle
    of (frame + start_frame, h, w); float64 math, result cast back to x.dtype."""
    B, L, H, D = x.shape
    f, h, w = grid
    c = D // 2
    n_f, n_h, n_w = c - 2 * (c // 3), c // 3, c // 3
    a_f, a_h, a_w = angles.to(x.device).split([n_f, n_h, n_w], dim=1)
    ang = torch.cat([
        a_f[start_frame:start_frame + f].view(f, 1, 1, n_f).expand(f, h, w, n_f),
        a_h[:h].view(1, h, 1, n_h).expand(f, h, w, n_h),
        a_w[:w].view(1, 1, w, n_w).expand(f, h, w, n_w)], dim=-1).reshape(f * h * w, 1, c)
    cos, sin = torch.cos(ang), torch.sin(ang)
    n = f * h * w
    xr = x[:, :n].to(torch.float64).reshape(B, n, H, c, 2)
    re, im = xr[..., 0], xr[..., 1]
    out = torch.stack([re * cos - im * sin, re * sin + im * cos], dim=-1).reshape(B, n, H, D)
    out = out.to(x.dtype)
    if n < L:
        out = torch.cat([out, x[:, n:]], dim=1)
    return out


def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """model.py:78-86 -- fp32 normalise over the last dim, cast back, then * weight."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return y.type_as(x) * weight


def layer_norm(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float) -> Tensor:
    """model.py:89-99 -- nn.LayerNorm on the stored dtype (fp32 statistics inside ATen)."""
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps).type_as(x)


def dense_attention(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """attention.py:187-202 (and FA2 at :136-150): softmax(q k^T / sqrt(d)) v, no mask.
    q [B, Lq, H, D], k/v [B, Lk, H, D] -> [B, Lq, H, D]."""
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return o.transpose(1, 2).contiguous()


# --------------------------------------------------------------------------------------
# KV-cache index arithmetic (integers; must be bit-exact)
# --------------------------------------------------------------------------------------
@dataclass
class CachePlan:
    """What one self-attention call does to one layer's rolling cache (causal_model.py:195-236)."""
    roll: bool
    roll_src: int      # first token of the slice that is moved left
    roll_dst: int      # where it lands (= sink tokens)
    roll_len: int
    write_start: int   # new K/V rows go to cache[write_start:write_end]
    write_end: int
    attn_start: int    # attention reads cache[attn_start:attn_end]
    attn_end: int
    global_end: int    # values stored back into the dict
    local_end: int


def plan_cache_update(global_end: int, local_end: int, current_start: int, num_new: int,
                      cache_size: int, local_attn_size: int, sink_tokens: int,
                      max_attention_size: int) -> CachePlan:
    current_end = current_start + num_new                                     # :202
    if local_attn_size != -1 and current_end > global_end and num_new + local_end > cache_size:  # :207-208
        evicted = num_new + local_end - cache_size                            # :212
        rolled = local_end - evicted - sink_tokens                            # :213
        new_local_end = local_end + current_end - global_end - evicted        # :219-220
        roll = True
        src, dst, rlen = sink_tokens + evicted, sink_tokens, rolled
    else:
        new_local_end = local_end + current_end - global_end                  # :226
        roll, src, dst, rlen = False, 0, 0, 0
    write_start = new_local_end - num_new                                     # :221 / :227
    attn_start = max(0, new_local_end - max_attention_size)                   # :232
    return CachePlan(roll, src, dst, rlen, write_start, new_local_end, attn_start, new_local_end,
                     current_end, new_local_end)


# --------------------------------------------------------------------------------------
# network
# --------------------------------------------------------------------------------------
def _lin(p: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def self_attention(p, pre: str, x: Tensor, grid, angles: Tensor, cache: dict, current_start: int,
                   cfg: OracleConfig) -> Tensor:
    """CausalWanSelfAttention.forward, cache branch (causal_model.py:106-117,194-241)."""
    B, L, _ = x.shape
    H, D = cfg.num_heads, cfg.head_dim
    q = rms_norm(_lin(p, pre + "q", x), p[pre + "norm_q.weight"], cfg.eps).view(B, L, H, D)
    k = rms_norm(_lin(p, pre + "k", x), p[pre + "norm_k.weight"], cfg.eps).view(B, L, H, D)
    v = _lin(p, pre + "v", x).view(B, L, H, D)
    frame_tokens = grid[1] * grid[2]
    start_frame = current_start // frame_tokens
    q = rope_rotate(q, grid, angles, start_frame).type_as(v)
    k = rope_rotate(k, grid, angles, start_frame).type_as(v)
    plan = plan_cache_update(int(cache["global_end_index"].item()), int(cache["local_end_index"].item()),
                             current_start, L, cache["k"].shape[1], cfg.local_attn_size,
                             cfg.sink_size * frame_tokens, cfg.max_attention_size)
    if plan.roll:
        for name in ("k", "v"):
            src = cache[name][:, plan.roll_src:plan.roll_src + plan.roll_len].clone()
            cache[name][:, plan.roll_dst:plan.roll_dst + plan.roll_len] = src
    cache["k"][:, plan.write_start:plan.write_end] = k
    cache["v"][:, plan.write_start:plan.write_end] = v
    o = dense_attention(q, cache["k"][:, plan.attn_start:plan.attn_end],
                        cache["v"][:, plan.attn_start:plan.attn_end])
    cache["global_end_index"].fill_(plan.global_end)
    cache["local_end_index"].fill_(plan.local_end)
    return _lin(p, pre + "o", o.flatten(2))


def cross_attention(p, pre: str, x: Tensor, context: Tensor, cache: Optional[dict],
                    cfg: OracleConfig) -> Tensor:
    """WanT2VCrossAttention.forward (model.py:161-194); K/V of the text computed once."""
    B = x.shape[0]
    H, D = cfg.num_heads, cfg.head_dim
    q = rms_norm(_lin(p, pre + "q", x), p[pre + "norm_q.weight"], cfg.eps).view(B, -1, H, D)
    if cache is not None and cache["is_init"]:
        k, v = cache["k"], cache["v"]
    else:
        k = rms_norm(_lin(p, pre + "k", context), p[pre + "norm_k.weight"], cfg.eps).view(B, -1, H, D)
        v = _lin(p, pre + "v", context).view(B, -1, H, D)
        if cache is not None:
            cache["is_init"] = True
            cache["k"], cache["v"] = k, v
    o = dense_attention(q, k, v)
    return _lin(p, pre + "o", o.flatten(2))


def block_forward(p, i: int, x: Tensor, e0: Tensor, grid, angles, context, kv_cache, crossattn_cache,
                  current_start: int, cfg: OracleConfig) -> Tensor:
    """CausalWanAttentionBlock.forward (causal_model.py:307-336).  e0 [B, F, 6, C]."""
    pre = f"blocks.{i}."
    B, L, C = x.shape
    nf = e0.shape[1]
    ft = L // nf
    e = (p[pre + "modulation"].unsqueeze(1) + e0).chunk(6, dim=2)   # 6 x [B, F, 1, C]

    def per_frame(t: Tensor) -> Tensor:
        return t.unflatten(1, (nf, ft))

    h = (per_frame(layer_norm(x, None, None, cfg.eps)) * (1 + e[1]) + e[0]).flatten(1, 2)
    y = self_attention(p, pre + "self_attn.", h, grid, angles, kv_cache, current_start, cfg)
    x = x + (per_frame(y) * e[2]).flatten(1, 2)
    x = x + cross_attention(p, pre + "cross_attn.",
                            layer_norm(x, p[pre + "norm3.weight"], p[pre + "norm3.bias"], cfg.eps),
                            context, crossattn_cache, cfg)
    h = (per_frame(layer_norm(x, None, None, cfg.eps)) * (1 + e[4]) + e[3]).flatten(1, 2)
    y = _lin(p, pre + "ffn.2", F.gelu(_lin(p, pre + "ffn.0", h), approximate="tanh"))
    x = x + (per_frame(y) * e[5]).flatten(1, 2)
    return x


def model_forward(p: Dict[str, Tensor], cfg: OracleConfig, x: Tensor, t: Tensor, context: Tensor,
                  kv_cache: List[dict], crossattn_cache: List[dict], current_start: int,
                  angles: Optional[Tensor] = None) -> Tensor:
    """CausalWanModel._forward_inference (causal_model.py:725-893).

    x [B, 16, F, H, W], t [B, F], context [B, <=512, 4096] -> flow [B, 16, F, H, W]."""
    B, _, nf, Hh, Ww = x.shape
    pt, ph, pw = cfg.patch_size
    grid = (nf // pt, Hh // ph, Ww // pw)
    if angles is None:
        angles = rope_angle_table(cfg.head_dim)
    tok = F.conv3d(x, p["patch_embedding.weight"], p["patch_embedding.bias"], stride=cfg.patch_size)
    tok = tok.flatten(2).transpose(1, 2)                                           # [B, L, C]

    e = sinusoid_embed(cfg.freq_dim, t.flatten()).type_as(tok)                    # :829-830
    e = _lin(p, "time_embedding.2", F.silu(_lin(p, "time_embedding.0", e)))       # [B*F, C]
    e0 = _lin(p, "time_projection.1", F.silu(e)).unflatten(1, (6, cfg.dim)).unflatten(0, t.shape)

    if context.shape[1] < cfg.text_len:                                           # :837-842
        context = torch.cat([context, context.new_zeros(B, cfg.text_len - context.shape[1], context.shape[2])], 1)
    ctx = _lin(p, "text_embedding.2", F.gelu(_lin(p, "text_embedding.0", context), approximate="tanh"))

    for i in range(cfg.num_layers):
        tok = block_forward(p, i, tok, e0, grid, angles, ctx, kv_cache[i], crossattn_cache[i],
                            current_start, cfg)

    # CausalHead (causal_model.py:356-367): modulated by the *un-projected* time embedding
    eh = (p["head.modulation"].unsqueeze(1) + e.unflatten(0, t.shape).unsqueeze(2)).chunk(2, dim=2)
    ft = tok.shape[1] // t.shape[1]
    y = layer_norm(tok, None, None, cfg.eps).unflatten(1, (t.shape[1], ft)) * (1 + eh[1]) + eh[0]
    y = _lin(p, "head.head", y)                                                   # [B, F, ft, 64]
    # unpatchify (causal_model.py:1081-1104): (f h w) (p q r c) -> c (f p) (h q) (w r)
    c = cfg.out_dim
    y = y.reshape(B, grid[0], grid[1], grid[2], pt, ph, pw, c)
    y = y.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, c, grid[0] * pt, grid[1] * ph, grid[2] * pw)
    return y


def bidirectional_forward(p: Dict[str, Tensor], cfg: OracleConfig, x: Tensor, t: Tensor, context: Tensor) -> Tensor:
    """WanModel._forward for t2v (wan/modules/model.py:637-771; blocks :275-354, head :439-466) for samples of
    exactly seq_len tokens: one timestep per SAMPLE (`e0 [B, 6, C]` :697-700), every token attends to the whole
    sequence, RoPE from frame 0 (rope_apply :40-67).  That is the cached forward above with a single timestep group
    and an empty cache, so it is restated as exactly that.  x [B, 16, F, H, W], t [B] -> [B, 16, F, H, W]."""
    B, _, nf, Hh, Ww = x.shape
    ft = (Hh // cfg.patch_size[1]) * (Ww // cfg.patch_size[2])
    kv = new_kv_cache(cfg, B, ft, x.dtype, x.device, cache_tokens=nf * ft)
    ca = new_crossattn_cache(cfg, B, x.dtype, x.device)
    for c in ca:                       # WanT2VCrossAttention without a cache recomputes K/V (model.py:175-180)
        c["is_init"] = False
    return model_forward(p, cfg, x, t.reshape(B, 1), context, kv, ca, 0)


# --------------------------------------------------------------------------------------
# scheduler / wrapper / rollout
# --------------------------------------------------------------------------------------
class OracleScheduler:
    """FlowMatchScheduler(shift, sigma_min=0, extra_one_step=True).set_timesteps(1000, training=True)
    (scheduler.py:118-141 as constructed at wan_wrapper.py:171-174)."""

    def __init__(self, shift: float, num_train_timesteps: int = 1000):
        s = torch.linspace(1.0, 0.0, num_train_timesteps + 1)[:-1]
        self.sigmas = shift * s / (1 + (shift - 1) * s)
        self.timesteps = self.sigmas * num_train_timesteps

    def nearest_index(self, timestep: Tensor) -> Tensor:
        ts = self.timesteps.to(timestep.device)
        return torch.argmin((ts.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)

    def add_noise(self, x0: Tensor, noise: Tensor, timestep: Tensor) -> Tensor:
        """scheduler.py:159-176 -- (1 - sigma) x0 + sigma noise with fp32 sigma, cast to noise dtype."""
        sigma = self.sigmas.to(noise.device)[self.nearest_index(timestep)].reshape(-1, 1, 1, 1)
        return ((1 - sigma) * x0 + sigma * noise).type_as(noise)


def warp_denoising_steps(sched: OracleScheduler, steps: Sequence[int]) -> Tensor:
    """causal_inference.py:27-31 -- index the shifted timestep table (+[0]) with 1000 - step."""
    ts = torch.cat((sched.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
    return ts[1000 - torch.tensor(list(steps), dtype=torch.long)]


def flow_to_x0(sched: OracleScheduler, flow: Tensor, xt: Tensor, timestep: Tensor) -> Tensor:
    """wan_wrapper.py:204-228 -- x0 = x_t - sigma_t * flow in float64, cast back."""
    sig = sched.sigmas.double().to(flow.device)
    ts = sched.timesteps.double().to(flow.device)
    idx = torch.argmin((ts.unsqueeze(0) - timestep.double().unsqueeze(1)).abs(), dim=1)
    return (xt.double() - sig[idx].reshape(-1, 1, 1, 1) * flow.double()).to(flow.dtype)


class OracleWrapper:
    """WanDiffusionWrapper.forward for the cached causal path (wan_wrapper.py:253-349)."""

    def __init__(self, params: Dict[str, Tensor], cfg: OracleConfig, timestep_shift: float):
        self.params, self.cfg = params, cfg
        self.scheduler = OracleScheduler(timestep_shift)
        self.angles = rope_angle_table(cfg.head_dim)

    def __call__(self, noisy: Tensor, prompt_embeds: Tensor, timestep: Tensor, kv_cache, crossattn_cache,
                 current_start: int) -> Tuple[Tensor, Tensor]:
        flow = model_forward(self.params, self.cfg, noisy.permute(0, 2, 1, 3, 4), timestep, prompt_embeds,
                             kv_cache, crossattn_cache, current_start, self.angles).permute(0, 2, 1, 3, 4)
        x0 = flow_to_x0(self.scheduler, flow.flatten(0, 1), noisy.flatten(0, 1),
                        timestep.flatten(0, 1)).unflatten(0, flow.shape[:2])
        return flow, x0


def new_kv_cache(cfg: OracleConfig, batch: int, frame_tokens: int, dtype, device, cache_tokens=None):
    """causal_inference.py:278-298 (cache size local_attn_size*fs, else 32760 by default)."""
    if cache_tokens is None:
        cache_tokens = cfg.local_attn_size * frame_tokens if cfg.local_attn_size != -1 else 32760
    return [dict(k=torch.zeros(batch, cache_tokens, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
                 v=torch.zeros(batch, cache_tokens, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
                 global_end_index=torch.tensor([0], dtype=torch.long, device=device),
                 local_end_index=torch.tensor([0], dtype=torch.long, device=device))
            for _ in range(cfg.num_layers)]


def new_crossattn_cache(cfg: OracleConfig, batch: int, dtype, device):
    """causal_inference.py:300-312."""
    return [dict(k=torch.zeros(batch, cfg.text_len, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
                 v=torch.zeros(batch, cfg.text_len, cfg.num_heads, cfg.head_dim, dtype=dtype, device=device),
                 is_init=False) for _ in range(cfg.num_layers)]


@dataclass
class RolloutTrace:
    latents: Tensor
    per_chunk: List[Tensor] = field(default_factory=list)
    index_trace: List[Tuple[int, int]] = field(default_factory=list)   # layer-0 (global, local) per forward


def rollout(wrapper: OracleWrapper, noise: Tensor, prompt_embeds: Tensor, denoising_steps: Tensor,
            num_frame_per_block: int, context_noise: float = 0, independent_first_frame: bool = False,
            kv_cache=None, crossattn_cache=None, cache_tokens=None, max_chunks: Optional[int] = None,
            noise_fn=None, initial_latent: Optional[Tensor] = None) -> RolloutTrace:
    """CausalInferencePipeline.inference without T5/VAE (causal_inference.py:72-246), t2v only.

    noise [B, F, 16, H, W].  `noise_fn(like)` supplies the re-noise sample (default
    torch.randn_like, consuming the global generator in the reference's order :208)."""
    cfg = wrapper.cfg
    B, nfr, _, Hh, Ww = noise.shape
    ft = (Hh // cfg.patch_size[1]) * (Ww // cfg.patch_size[2])
    if noise_fn is None:
        noise_fn = torch.randn_like
    n_in = 0 if initial_latent is None else initial_latent.shape[1]
    if independent_first_frame and initial_latent is None:
        assert (nfr - 1) % num_frame_per_block == 0
        chunks = [1] + [num_frame_per_block] * ((nfr - 1) // num_frame_per_block)
    else:
        assert nfr % num_frame_per_block == 0
        chunks = [num_frame_per_block] * (nfr // num_frame_per_block)
    if kv_cache is None:
        kv_cache = new_kv_cache(cfg, B, ft, noise.dtype, noise.device, cache_tokens)
    if crossattn_cache is None:
        crossattn_cache = new_crossattn_cache(cfg, B, noise.dtype, noise.device)
    out = torch.zeros(B, n_in + nfr, *noise.shape[2:], dtype=noise.dtype, device=noise.device)
    trace = RolloutTrace(latents=out)

    def note():
        trace.index_trace.append((int(kv_cache[0]["global_end_index"].item()),
                                  int(kv_cache[0]["local_end_index"].item())))

    start = 0
    if initial_latent is not None:
        # causal_inference.py:135-169: the conditioning frames are written to the output and pushed through the model
        # at timestep 0 only to fill the KV cache (one lone first frame if independent_first_frame, then whole blocks)
        groups = []
        if independent_first_frame:
            assert (n_in - 1) % num_frame_per_block == 0
            groups.append(1)
            groups += [num_frame_per_block] * ((n_in - 1) // num_frame_per_block)
        else:
            assert n_in % num_frame_per_block == 0
            groups = [num_frame_per_block] * (n_in // num_frame_per_block)
        for n in groups:
            ref = initial_latent[:, start:start + n]
            out[:, start:start + n] = ref
            wrapper(ref, prompt_embeds, torch.zeros([B, n], device=noise.device, dtype=torch.int64), kv_cache,
                    crossattn_cache, start * ft)
            note()
            start += n
    for ci, n in enumerate(chunks):
        if max_chunks is not None and ci >= max_chunks:
            break
        x = noise[:, start - n_in:start - n_in + n]
        for si, ts in enumerate(denoising_steps):
            timestep = torch.ones([B, n], device=noise.device, dtype=torch.int64) * ts   # :191-194
            _, x0 = wrapper(x, prompt_embeds, timestep, kv_cache, crossattn_cache, start * ft)
            note()
            if si < len(denoising_steps) - 1:
                nxt = denoising_steps[si + 1] * torch.ones([B * n], device=noise.device, dtype=torch.long)
                flat = x0.flatten(0, 1)
                x = wrapper.scheduler.add_noise(flat, noise_fn(flat), nxt).unflatten(0, x0.shape[:2])
        out[:, start:start + n] = x0
        trace.per_chunk.append(x0.clone())
        ctx_t = torch.ones_like(timestep) * context_noise                                # :227
        wrapper(x0, prompt_embeds, ctx_t, kv_cache, crossattn_cache, start * ft)         # cache refresh
        note()
        start += n
    return trace


# --------------------------------------------------------------------------------------
# block masks (training-side consumers; the tables must be bit-exact)
# --------------------------------------------------------------------------------------
def _pad128(n: int) -> int:
    return (n + 127) // 128 * 128


def blockwise_causal_ends(num_frames: int, frame_tokens: int, num_frame_per_block: int,
                          independent_first_frame: bool = False) -> torch.Tensor:
    """`ends[q]` of causal_model.py:533-546 (and the i2v variant :680-696), padded to x128
    with zeros like the reference."""
    total = num_frames * frame_tokens
    ends = torch.zeros(_pad128(total), dtype=torch.long)
    blk = frame_tokens * num_frame_per_block
    first = 0
    if independent_first_frame:
        ends[:frame_tokens] = frame_tokens
        first = frame_tokens
    for s in range(first, total, blk):
        ends[s:s + blk] = s + blk
    return ends


def blockwise_causal_mask(num_frames: int, frame_tokens: int, num_frame_per_block: int,
                          local_attn_size: int = -1, independent_first_frame: bool = False) -> torch.Tensor:
    """Dense boolean [Lp, Lp] version of the mask_mod at causal_model.py:548-552 / :698-703."""
    ends = blockwise_causal_ends(num_frames, frame_tokens, num_frame_per_block, independent_first_frame)
    n = ends.numel()
    q = torch.arange(n).view(n, 1)
    kv = torch.arange(n).view(1, n)
    e = ends.view(n, 1)
    m = kv < e
    if local_attn_size != -1:
        m = m & (kv >= e - local_attn_size * frame_tokens)
    return m | (q == kv)


def teacher_forcing_mask(num_frames: int, frame_tokens: int, num_frame_per_block: int) -> torch.Tensor:
    """Dense boolean version of causal_model.py:592-645 (clean half followed by noisy half)."""
    half = num_frames * frame_tokens
    total = 2 * half
    n = _pad128(total)
    blk = frame_tokens * num_frame_per_block
    ctx_end = torch.zeros(n, dtype=torch.long)
    nn_start = torch.zeros(n, dtype=torch.long)
    nn_end = torch.zeros(n, dtype=torch.long)
    nc_end = torch.zeros(n, dtype=torch.long)
    for s in range(0, half, blk):
        ctx_end[s:s + blk] = s + blk
    for bi, s in enumerate(range(half, total, blk)):
        nn_start[s:s + blk] = s
        nn_end[s:s + blk] = s + blk
        nc_end[s:s + blk] = bi * blk
    q = torch.arange(n).view(n, 1)
    kv = torch.arange(n).view(1, n)
    clean = (q < half) & (kv < ctx_end.view(n, 1))
    c1 = (kv < nn_end.view(n, 1)) & (kv >= nn_start.view(n, 1))
    c2 = kv < nc_end.view(n, 1)            # noise_context_starts is all-zero (:601,632)
    noisy = (q >= half) & (c1 | c2)
    return (q == kv) | clean | noisy


def block_table(mask: torch.Tensor, block: int = 128) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per 128x128 tile: (any, all) -- what create_block_mask reduces a mask_mod to
    (partial blocks = any & ~all, full blocks = all)."""
    n = mask.shape[0] // block
    t = mask.view(n, block, n, block).permute(0, 2, 1, 3).reshape(n, n, -1)
    return t.any(-1), t.all(-1)
