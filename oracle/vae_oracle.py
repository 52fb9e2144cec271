"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Wan VAE *decoder* as the rollout uses it right after the hot path
(SURVEY.md section 8f rank 1): `WanVAEWrapper.decode_to_pixel` (utils/wan_wrapper.py:94-117) ->
`WanVAE_.decode` / `cached_decode` (wan/modules/vae.py:545-593) -> `Decoder3d.forward` (:423-472), one latent frame
at a time with a two-frame feature cache in front of every causal 3-D convolution.

Functional: parameters live in a dict keyed by the reference's `state_dict` names (`conv2.weight`,
`decoder.conv1.weight`, `decoder.middle.0.residual.0.gamma`, `decoder.upsamples.3.time_conv.weight`, ...), activations are
channels-first like the reference, every op runs in the activations' dtype (bf16 on the reference's GPU path:
`inference.py:73` casts the whole pipeline).  Each function cites the reference lines it follows; nothing is copied.

Pinned by tests/golden/vae_decode_tiny.pt, produced by the unmodified reference `WanVAE_` (oracle/make_golden.py).
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

CACHE_FRAMES = 2   # vae.py:14

LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]     # wan_wrapper.py:61-68
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


@dataclass(frozen=True)
class VaeConfig:
    """`_video_vae` defaults (vae.py:617-624) as seen by the decoder."""
    dim: int = 96
    z_dim: int = 16
    dim_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temporal_upsample: Tuple[bool, ...] = (True, True, False)    # reversed temperal_downsample (:499)

    def stage_plan(self):
        """[(kind, name, in_dim, out_dim)] of `decoder.upsamples` in module order (vae.py:389-415)."""
        dims = [self.dim * u for u in (self.dim_mult[-1],) + tuple(reversed(self.dim_mult))]
        plan, n = [], 0
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                cin //= 2
            for _ in range(self.num_res_blocks + 1):
                plan.append(("res", f"decoder.upsamples.{n}", cin, cout))
                cin = cout
                n += 1
            if i != len(self.dim_mult) - 1:
                plan.append(("up3d" if self.temporal_upsample[i] else "up2d", f"decoder.upsamples.{n}", cout, cout // 2))
                n += 1
        return plan, dims


def decoder_parameter_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    plan, dims = cfg.stage_plan()
    s: Dict[str, Tuple[int, ...]] = {}

    def conv3(name, cin, cout, k=(3, 3, 3)):
        s[name + ".weight"], s[name + ".bias"] = (cout, cin) + tuple(k), (cout,)

    def res(name, cin, cout):
        s[name + ".residual.0.gamma"] = (cin, 1, 1, 1)
        conv3(name + ".residual.2", cin, cout)
        s[name + ".residual.3.gamma"] = (cout, 1, 1, 1)
        conv3(name + ".residual.6", cout, cout)
        if cin != cout:
            conv3(name + ".shortcut", cin, cout, (1, 1, 1))

    conv3("conv2", cfg.z_dim, cfg.z_dim, (1, 1, 1))
    conv3("decoder.conv1", cfg.z_dim, dims[0])
    res("decoder.middle.0", dims[0], dims[0])
    s["decoder.middle.1.norm.gamma"] = (dims[0], 1, 1)
    s["decoder.middle.1.to_qkv.weight"], s["decoder.middle.1.to_qkv.bias"] = (3 * dims[0], dims[0], 1, 1), (3 * dims[0],)
    s["decoder.middle.1.proj.weight"], s["decoder.middle.1.proj.bias"] = (dims[0], dims[0], 1, 1), (dims[0],)
    res("decoder.middle.2", dims[0], dims[0])
    for kind, name, cin, cout in plan:
        if kind == "res":
            res(name, cin, cout)
        else:
            s[name + ".resample.1.weight"], s[name + ".resample.1.bias"] = (cout, cin, 3, 3), (cout,)
            if kind == "up3d":
                conv3(name + ".time_conv", cin, 2 * cin, (3, 1, 1))
    s["decoder.head.0.gamma"] = (dims[-1], 1, 1, 1)
    conv3("decoder.head.2", dims[-1], 3)
    return s


def make_random_vae_params(cfg: VaeConfig, seed: int = 0, dtype=torch.bfloat16) -> Dict[str, Tensor]:
    """Synthetic decoder weights: fan-in scaled normal convolutions, small biases, gammas near one (there is no
    network for the real checkpoint; the attention output projection is NOT zero-initialised so that the block matters)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in decoder_parameter_shapes(cfg).items():
        if name.endswith("gamma"):
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / fan_in ** 0.5
        out[name] = t.to(dtype)
    return out


# ---- layers ------------------------------------------------------------------------------------------------------
def causal_conv3d(p, name: str, x: Tensor, cache: Optional[Tensor] = None) -> Tensor:
    """vae.py:17-36: all temporal padding goes in front (2 * pad_t frames); cached frames replace that many zeros."""
    w, b = p[name + ".weight"], p[name + ".bias"]
    kt, kh, kw = w.shape[2:]
    pad_t = kt - 1
    if cache is not None and pad_t > 0:
        x = torch.cat([cache, x], dim=2)
        pad_t -= cache.shape[2]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, pad_t, 0))
    return F.conv3d(x, w, b)


def rms_norm(x: Tensor, gamma: Tensor) -> Tensor:
    """vae.py:39-54: unit L2 norm over channels, times sqrt(C), times gamma (bias is the float 0.)."""
    return F.normalize(x, dim=1) * x.shape[1] ** 0.5 * gamma + 0.0


def _next_cache(x: Tensor, old) -> Tensor:
    """Last two input frames; if the call brought only one, the newest cached frame is kept in front (:207-216)."""
    keep = x[:, :, -CACHE_FRAMES:].clone()
    if keep.shape[2] < 2 and isinstance(old, Tensor):
        keep = torch.cat([old[:, :, -1:], keep], dim=2)
    return keep


def _cached_conv(p, name: str, x: Tensor, cache: List, idx: List[int]) -> Tensor:
    i = idx[0]
    keep = _next_cache(x, cache[i])
    y = causal_conv3d(p, name, x, cache[i])
    cache[i] = keep
    idx[0] += 1
    return y


def residual_block(p, name: str, x: Tensor, cache: List, idx: List[int]) -> Tensor:
    """vae.py:186-220; the 1x1x1 shortcut never touches the cache."""
    h = causal_conv3d(p, name + ".shortcut", x) if (name + ".shortcut.weight") in p else x
    y = F.silu(rms_norm(x, p[name + ".residual.0.gamma"]))
    y = _cached_conv(p, name + ".residual.2", y, cache, idx)
    y = F.silu(rms_norm(y, p[name + ".residual.3.gamma"]))
    y = _cached_conv(p, name + ".residual.6", y, cache, idx)
    return y + h


def attention_block(p, name: str, x: Tensor) -> Tensor:
    """vae.py:223-262: per frame, one head of width C over the H*W positions."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm(y, p[name + ".norm.gamma"])
    qkv = F.conv2d(y, p[name + ".to_qkv.weight"], p[name + ".to_qkv.bias"])
    q, k, v = qkv.reshape(b * t, 1, 3 * c, h * w).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
    y = F.scaled_dot_product_attention(q, k, v)
    y = y.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
    y = F.conv2d(y, p[name + ".proj.weight"], p[name + ".proj.bias"])
    return y.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x


def upsample(p, name: str, kind: str, x: Tensor, cache: List, idx: List[int]) -> Tensor:
    """vae.py:101-147 (upsample2d / upsample3d): optional temporal doubling through `time_conv` (skipped for the very
    first frame of a video, marked 'Rep'), then nearest 2x in H, W and a 3x3 Conv2d that halves the channels."""
    b, c, t, h, w = x.shape
    if kind == "up3d":
        i = idx[0]
        if cache[i] is None:
            cache[i] = "Rep"
            idx[0] += 1
        else:
            keep = x[:, :, -CACHE_FRAMES:].clone()
            if keep.shape[2] < 2:
                front = torch.zeros_like(keep) if isinstance(cache[i], str) else cache[i][:, :, -1:]
                keep = torch.cat([front, keep], dim=2)
            y = causal_conv3d(p, name + ".time_conv", x, None if isinstance(cache[i], str) else cache[i])
            cache[i] = keep
            idx[0] += 1
            # channel halves become alternating frames (:141-144)
            x = y.reshape(b, 2, c, t, h, w).permute(0, 2, 3, 1, 4, 5).reshape(b, c, 2 * t, h, w)
            t = 2 * t
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest").type_as(y)       # :57-63
    y = F.conv2d(y, p[name + ".resample.1.weight"], p[name + ".resample.1.bias"], padding=1)
    return y.reshape(b, t, c // 2, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)


def decoder_forward(p, cfg: VaeConfig, x: Tensor, cache: List, idx: List[int]) -> Tensor:
    """vae.py:423-472."""
    plan, _ = cfg.stage_plan()
    x = _cached_conv(p, "decoder.conv1", x, cache, idx)
    x = residual_block(p, "decoder.middle.0", x, cache, idx)
    x = attention_block(p, "decoder.middle.1", x)
    x = residual_block(p, "decoder.middle.2", x, cache, idx)
    for kind, name, _, _ in plan:
        x = residual_block(p, name, x, cache, idx) if kind == "res" else upsample(p, name, kind, x, cache, idx)
    x = F.silu(rms_norm(x, p["decoder.head.0.gamma"]))
    return _cached_conv(p, "decoder.head.2", x, cache, idx)


def cache_slots(cfg: VaeConfig) -> int:
    """Number of feature-cache slots the decoder walks through per call (vae.py:475-481 counts every CausalConv3d,
    shortcuts included; the walk itself only visits conv1, two per residual block, the time convs and the head)."""
    plan, _ = cfg.stage_plan()
    return 1 + 2 * 2 + sum(2 if k == "res" else (1 if k == "up3d" else 0) for k, *_ in plan) + 1


def decode(p, cfg: VaeConfig, z: Tensor, cache: Optional[List] = None) -> Tensor:
    """vae.py:545-593.  z [B, 16, F, h, w] -> [B, 3, 1 + 4 (F - 1), 8 h, 8 w].  With `cache` (a list of
    `cache_slots` entries, initially None) the call continues a video like `cached_decode`."""
    dt = z.dtype
    mean = torch.tensor(LATENT_MEAN, dtype=torch.float32).to(dt).view(1, cfg.z_dim, 1, 1, 1)
    inv_std = (1.0 / torch.tensor(LATENT_STD, dtype=torch.float32).to(dt)).view(1, cfg.z_dim, 1, 1, 1)   # wan_wrapper.py:101-102
    z = z / inv_std + mean
    x = causal_conv3d(p, "conv2", z)
    if cache is None:
        cache = [None] * cache_slots(cfg)
    outs = []
    for i in range(x.shape[2]):
        outs.append(decoder_forward(p, cfg, x[:, :, i:i + 1], cache, [0]))
    return torch.cat(outs, dim=2)


def decode_to_pixel(p, cfg: VaeConfig, latent: Tensor, cache: Optional[List] = None) -> Tensor:
    """wan_wrapper.py:94-117: latent [B, F, 16, h, w] -> float32 [B, T, 3, H, W] clamped to [-1, 1]."""
    out = [decode(p, cfg, u.unsqueeze(0), cache).float().clamp_(-1, 1).squeeze(0) for u in latent.permute(0, 2, 1, 3, 4)]
    return torch.stack(out, dim=0).permute(0, 2, 1, 3, 4)
