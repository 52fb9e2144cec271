"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the UMT5 text encoder, the step right before the rollout
(SURVEY.md section 8f rank 4): `WanTextEncoder.forward` (utils/wan_wrapper.py:18-52) -> `T5Encoder.forward`
(wan/modules/t5.py:303-312) -> 24 x `T5SelfAttention` (:170-175) with a per-layer relative position bias
(`shared_pos=False`, :221-264), un-scaled attention with fp32 softmax (:86-120) and a gated-GELU feed-forward whose
GELU is evaluated op by op (:46-50, :136-141).

Functional: parameters in a dict keyed by the reference's `state_dict` names (`token_embedding.weight`,
`blocks.N.attn.q.weight`, `blocks.N.pos_embedding.embedding.weight`, `blocks.N.ffn.gate.0.weight`, `norm.weight`, ...).
Every op runs in the activations' dtype like the reference (bf16 after `inference.py:73` casts the pipeline).

Pinned by tests/golden/t5_tiny.pt (oracle/make_golden.py runs the unmodified reference `T5Encoder`).
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


@dataclass(frozen=True)
class T5Config:
    """`umt5_xxl` (t5.py:456-469), encoder side."""
    vocab: int = 256384
    dim: int = 4096
    dim_attn: int = 4096
    dim_ffn: int = 10240
    num_heads: int = 64
    num_layers: int = 24
    num_buckets: int = 32
    max_dist: int = 128
    eps: float = 1e-6


def parameter_shapes(cfg: T5Config) -> Dict[str, Tuple[int, ...]]:
    s = {"token_embedding.weight": (cfg.vocab, cfg.dim), "norm.weight": (cfg.dim,)}
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        s[b + "norm1.weight"] = s[b + "norm2.weight"] = (cfg.dim,)
        for n in "qkv":
            s[b + f"attn.{n}.weight"] = (cfg.dim_attn, cfg.dim)
        s[b + "attn.o.weight"] = (cfg.dim, cfg.dim_attn)
        s[b + "pos_embedding.embedding.weight"] = (cfg.num_buckets, cfg.num_heads)
        s[b + "ffn.gate.0.weight"] = s[b + "ffn.fc1.weight"] = (cfg.dim_ffn, cfg.dim)
        s[b + "ffn.fc2.weight"] = (cfg.dim, cfg.dim_ffn)
    return s


def make_random_t5_params(cfg: T5Config, seed: int = 0, dtype=torch.bfloat16) -> Dict[str, Tensor]:
    """Synthetic weights with the reference's initialisation scales (t5.py:27-43), norm weights near one."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in parameter_shapes(cfg).items():
        if name.endswith("norm.weight") or "norm1" in name or "norm2" in name:
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif name == "token_embedding.weight":
            t = torch.randn(shape, generator=g)
        elif "pos_embedding" in name:
            t = torch.randn(shape, generator=g) * 0.5
        else:
            t = torch.randn(shape, generator=g) * shape[1] ** -0.5
        out[name] = t.to(dtype)
    return out


def layer_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """t5.py:61-66: RMS norm, statistics and scaling in fp32, cast to the weight dtype, then the weight."""
    y = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        y = y.type_as(weight)
    return weight * y


def relative_buckets(lq: int, lk: int, num_buckets: int, max_dist: int) -> Tensor:
    """t5.py:245-264, bidirectional: half the buckets per sign, exact up to num_buckets/4, logarithmic beyond."""
    rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
    nb = num_buckets // 2
    out = (rel > 0).long() * nb
    rel = rel.abs()
    exact = nb // 2
    large = exact + (torch.log(rel.float() / exact) / math.log(max_dist / exact) * (nb - exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(rel < exact, rel, large)


def gelu_chain(x: Tensor) -> Tensor:
    """t5.py:46-50: the tanh GELU written out with tensor ops (each one rounds in the activations' dtype)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def attention(p, pre: str, x: Tensor, mask: Optional[Tensor], bias: Tensor, heads: int) -> Tensor:
    """t5.py:86-120 (self-attention): no 1/sqrt(d) scaling, additive position bias, masked keys at finfo.min, softmax in
    fp32."""
    b, L, _ = x.shape
    q = F.linear(x, p[pre + "q.weight"]).view(b, L, heads, -1)
    k = F.linear(x, p[pre + "k.weight"]).view(b, L, heads, -1)
    v = F.linear(x, p[pre + "v.weight"]).view(b, L, heads, -1)
    ab = x.new_zeros(b, heads, L, L)
    ab += bias
    if mask is not None:
        ab.masked_fill_(mask.view(b, 1, 1, -1) == 0, torch.finfo(x.dtype).min)
    a = torch.einsum("binc,bjnc->bnij", q, k) + ab
    a = F.softmax(a.float(), dim=-1).type_as(a)
    y = torch.einsum("bnij,bjnc->binc", a, v).reshape(b, L, -1)
    return F.linear(y, p[pre + "o.weight"])


def encoder_forward(p: Dict[str, Tensor], cfg: T5Config, ids: Tensor, mask: Optional[Tensor]) -> Tensor:
    """t5.py:303-312 (dropout is the identity in eval mode).  ids / mask [B, L] -> [B, L, dim]."""
    x = F.embedding(ids, p["token_embedding.weight"])
    L = x.shape[1]
    buckets = relative_buckets(L, L, cfg.num_buckets, cfg.max_dist).to(ids.device)
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        bias = F.embedding(buckets, p[b + "pos_embedding.embedding.weight"]).permute(2, 0, 1).unsqueeze(0).contiguous()
        x = x + attention(p, b + "attn.", layer_norm(x, p[b + "norm1.weight"], cfg.eps), mask, bias, cfg.num_heads)
        h = layer_norm(x, p[b + "norm2.weight"], cfg.eps)
        h = F.linear(h, p[b + "ffn.fc1.weight"]) * gelu_chain(F.linear(h, p[b + "ffn.gate.0.weight"]))
        x = x + F.linear(h, p[b + "ffn.fc2.weight"])
    return layer_norm(x, p["norm.weight"], cfg.eps)


def text_encoder(p: Dict[str, Tensor], cfg: T5Config, ids: Tensor, mask: Tensor) -> Tensor:
    """`WanTextEncoder.forward` after the tokenizer (wan_wrapper.py:38-52): rows beyond each prompt's length are zeroed."""
    ctx = encoder_forward(p, cfg, ids, mask)
    for u, n in zip(ctx, mask.gt(0).sum(dim=1).long()):
        u[n:] = 0.0
    return ctx
