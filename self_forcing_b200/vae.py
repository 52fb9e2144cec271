"""Wan VAE decoder on the B200 kernels -- the step right after the rollout.

Mirrors `WanVAEWrapper.decode_to_pixel(latent, use_cache=False)` (utils/wan_wrapper.py:94-117) and underneath it
`WanVAE_.decode` / `cached_decode` (wan/modules/vae.py:545-593) -> `Decoder3d.forward` (:423-472): one latent frame at
a time, every causal 3-D convolution preceded by a two-frame feature cache, the first frame of a video skipping the
temporal upsampling ("Rep", :104-109).  `load_state_dict` takes the reference's keys (`conv2.*`, `decoder.*`; `encoder.*`
and `conv1.*` are ignored -- the encoder is not on the path).

B200 layout: activations are channels-last `[T, H, W, C]` bf16, so
  * a convolution is a GEMM over gathered voxel rows (`sfb_causal_conv3d_cl`: bias, the ResidualBlock's `x + h` and
    the time_conv's channel-halves -> alternate-frames shuffle are all fused into the tcgen05 GEMM epilogue),
  * the per-voxel channel RMS norm + SiLU is a row kernel (`sfb_vae_norm_silu`),
  * nearest-neighbour 2x upsampling never materialises: the gather of the following Conv2d reads (h/2, w/2),
  * the single-head attention of the middle block is two GEMMs around a row softmax.
Weights are repacked once at load: `[Cout, Cin, kt, kh, kw]` -> `[Cout (padded to 8), kt*kh*kw*Cin]`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

EPI_BIAS, EPI_RESIDUAL, EPI_F32 = 0, 2, 4
CACHE_FRAMES = 2      # vae.py:14

LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]     # wan_wrapper.py:61-68
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


def _decoder_plan(dim: int, dim_mult, num_res_blocks: int, temporal_upsample):
    """Module order of `decoder.upsamples` (vae.py:389-415): [(kind, name, in_dim, out_dim)]."""
    dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    plan, n = [], 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin //= 2
        for _ in range(num_res_blocks + 1):
            plan.append(("res", f"decoder.upsamples.{n}", cin, cout))
            cin = cout
            n += 1
        if i != len(dim_mult) - 1:
            plan.append(("up3d" if temporal_upsample[i] else "up2d", f"decoder.upsamples.{n}", cout, cout // 2))
            n += 1
    return plan, dims


class B200VAEDecoder:
    def __init__(self, dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                 temperal_downsample=(False, True, True), ops=None, device=None):
        if z_dim != 16:
            raise NotImplementedError("B200 VAE decoder: z_dim must be 16 (the latent-in kernel is specialised)")
        self.dim, self.z_dim = dim, z_dim
        self.plan, self.dims = _decoder_plan(dim, list(dim_mult), num_res_blocks, list(temperal_downsample)[::-1])
        self._ops = ops
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.raw: Dict[str, torch.Tensor] = {}       # reference-named tensors as loaded
        self.w: Dict[str, torch.Tensor] = {}         # packed GEMM operands / gammas / biases
        self.geom: Dict[str, Tuple[int, int]] = {}   # conv name -> (kt, ks)
        self.cache: Optional[List] = None            # persistent feature cache of `cached_decode`
        # 3x3 convolutions as implicit GEMM (shifted TMA boxes, nothing staged); False = gather + GEMM for all
        self.implicit_conv = True

    @property
    def ops(self):
        if self._ops is None:
            from .ops import CudaOps
            self._ops = CudaOps()
        return self._ops

    # ---- weights ---------------------------------------------------------------------------
    def expected_keys(self) -> List[str]:
        keys = ["conv2.weight", "conv2.bias", "decoder.conv1.weight", "decoder.conv1.bias"]

        def res(name, cin, cout):
            keys.extend([name + ".residual.0.gamma", name + ".residual.2.weight", name + ".residual.2.bias",
                         name + ".residual.3.gamma", name + ".residual.6.weight", name + ".residual.6.bias"])
            if cin != cout:
                keys.extend([name + ".shortcut.weight", name + ".shortcut.bias"])

        res("decoder.middle.0", self.dims[0], self.dims[0])
        keys.extend(["decoder.middle.1.norm.gamma", "decoder.middle.1.to_qkv.weight", "decoder.middle.1.to_qkv.bias",
                     "decoder.middle.1.proj.weight", "decoder.middle.1.proj.bias"])
        res("decoder.middle.2", self.dims[0], self.dims[0])
        for kind, name, cin, cout in self.plan:
            if kind == "res":
                res(name, cin, cout)
            else:
                keys.extend([name + ".resample.1.weight", name + ".resample.1.bias"])
                if kind == "up3d":
                    keys.extend([name + ".time_conv.weight", name + ".time_conv.bias"])
        keys.extend(["decoder.head.0.gamma", "decoder.head.2.weight", "decoder.head.2.bias"])
        return keys

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        want = self.expected_keys()
        missing = [k for k in want if k not in sd]
        unexpected = [k for k in sd if k not in want and not k.startswith(("encoder.", "conv1."))]
        if missing or (strict and unexpected):
            raise KeyError(f"VAE decoder state_dict: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected {unexpected[:5]}")
        dev = self.device
        self.raw = {k: sd[k].detach().to(device=dev, dtype=torch.bfloat16) for k in want}
        self.w, self.geom = {}, {}
        for k, t in self.raw.items():
            if k.endswith("gamma"):
                self.w[k] = t.reshape(-1).contiguous()
            elif k.endswith(".weight"):
                name = k[:-len(".weight")]
                if t.dim() == 4:                       # Conv2d -> kt = 1
                    t = t.unsqueeze(2)
                cout, cin, kt, kh, kw = t.shape
                packed = t.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh * kw * cin)
                bias = self.raw[name + ".bias"]
                pad = (-cout) % 8                      # the GEMM writes 16-byte column groups
                if pad:
                    packed = torch.cat([packed, packed.new_zeros(pad, packed.shape[1])])
                    bias = torch.cat([bias, bias.new_zeros(pad)])
                self.w[k], self.w[name + ".bias"] = packed.contiguous(), bias.contiguous()
                self.geom[name] = (kt, kh)
        self.mean = torch.tensor(LATENT_MEAN, dtype=torch.float32).to(device=dev, dtype=torch.bfloat16)
        self.inv_std = (1.0 / torch.tensor(LATENT_STD, dtype=torch.float32).to(torch.bfloat16)).to(dev)   # wan_wrapper.py:101-102
        return missing, unexpected

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.raw)

    def to(self, device):
        self.device = torch.device(device)
        if self.raw:
            self.load_state_dict(self.raw, strict=True)
        return self

    def cache_slots(self) -> int:
        return 1 + 4 + sum(2 if k == "res" else (1 if k == "up3d" else 0) for k, *_ in self.plan) + 1

    def clear_cache(self) -> None:
        self.cache = None

    # ---- layers (x is [T, H, W, C]) --------------------------------------------------------
    def _conv(self, name: str, x: torch.Tensor, t_zero_pad: int, residual=None, upsample=False) -> torch.Tensor:
        kt, ks = self.geom[name]
        w, b = self.w[name + ".weight"], self.w[name + ".bias"]
        t_out = x.shape[0] + t_zero_pad - (kt - 1)
        implicit = self.implicit_conv and ks == 3 and x.shape[3] % 16 == 0
        if implicit and upsample:                      # the TMA boxes need the upsampled frames in memory
            up = torch.empty(x.shape[0], 2 * x.shape[1], 2 * x.shape[2], x.shape[3], dtype=x.dtype, device=x.device)
            self.ops.upsample2x(x, up)
            x, upsample = up, False
        H, W = (2 * x.shape[1], 2 * x.shape[2]) if upsample else (x.shape[1], x.shape[2])
        y = torch.empty(t_out, H, W, w.shape[0], dtype=x.dtype, device=x.device)
        self.ops.causal_conv3d(x, t_zero_pad, w, b, kt, ks, y.view(-1, w.shape[0]), upsample=upsample,
                               residual=None if residual is None else residual.reshape(-1, w.shape[0]), implicit=implicit)
        return y

    def _cached_conv(self, name: str, x: torch.Tensor, cache: List, idx: List[int], residual=None) -> torch.Tensor:
        """vae.py:204-219: the previous call's last two input frames stand in front of the new ones."""
        i = idx[0]
        old = cache[i]
        keep = x[-CACHE_FRAMES:].clone()
        if keep.shape[0] < 2 and isinstance(old, torch.Tensor):
            keep = torch.cat([old[-1:], keep], dim=0)
        if isinstance(old, torch.Tensor):
            xin, pad = torch.cat([old, x], dim=0), CACHE_FRAMES - old.shape[0]
        else:
            xin, pad = x, CACHE_FRAMES
        y = self._conv(name, xin, pad, residual=residual)
        cache[i] = keep
        idx[0] += 1
        return y

    def _norm(self, x: torch.Tensor, gamma: torch.Tensor, silu: bool) -> torch.Tensor:
        y = torch.empty_like(x)
        C = x.shape[-1]
        self.ops.vae_norm_silu(x.view(-1, C), gamma, y.view(-1, C), silu)
        return y

    def _res(self, name: str, x: torch.Tensor, cache: List, idx: List[int]) -> torch.Tensor:
        """vae.py:186-220."""
        h = self._conv(name + ".shortcut", x, 0) if (name + ".shortcut") in self.geom else x
        y = self._norm(x, self.w[name + ".residual.0.gamma"], True)
        y = self._cached_conv(name + ".residual.2", y, cache, idx)
        y = self._norm(y, self.w[name + ".residual.3.gamma"], True)
        return self._cached_conv(name + ".residual.6", y, cache, idx, residual=h)

    def _attn(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """vae.py:223-262: per frame, one head of width C over the H*W positions."""
        T, H, W, C = x.shape
        ops = self.ops
        out = torch.empty_like(x)
        y = self._norm(x, self.w[name + ".norm.gamma"], False)
        hw = H * W
        qkv = torch.empty(hw, 3 * C, dtype=x.dtype, device=x.device)
        scores = torch.empty(hw, hw, dtype=torch.float32, device=x.device)     # logits stay fp32 (SFB_EPI_F32)
        probs = torch.empty(hw, hw, dtype=x.dtype, device=x.device)
        vt = torch.empty(C, hw, dtype=x.dtype, device=x.device)
        o = torch.empty(hw, C, dtype=x.dtype, device=x.device)
        for t in range(T):
            ops.gemm(y[t].view(hw, C), self.w[name + ".to_qkv.weight"], self.w[name + ".to_qkv.bias"], qkv)
            ops.gemm(qkv[:, :C], qkv[:, C:2 * C], None, scores, epilogue=EPI_F32)
            ops.softmax_rows(scores, probs, 1.0 / C ** 0.5)
            ops.transpose(qkv[:, 2 * C:], vt)
            ops.gemm(probs, vt, None, o)
            ops.gemm(o, self.w[name + ".proj.weight"], self.w[name + ".proj.bias"], out[t].view(hw, C),
                     epilogue=EPI_RESIDUAL, residual=x[t].view(hw, C))
        return out

    def _upsample(self, name: str, kind: str, x: torch.Tensor, cache: List, idx: List[int]) -> torch.Tensor:
        """vae.py:101-147."""
        T, H, W, C = x.shape
        if kind == "up3d":
            i = idx[0]
            old = cache[i]
            if old is None:
                cache[i] = "Rep"                       # first frame of a video: no temporal upsampling
            else:
                keep = x[-CACHE_FRAMES:].clone()
                if keep.shape[0] < 2:
                    front = torch.zeros_like(keep) if isinstance(old, str) else old[-1:]
                    keep = torch.cat([front, keep], dim=0)
                if isinstance(old, str):
                    xin, pad = x, CACHE_FRAMES
                else:
                    xin, pad = torch.cat([old, x], dim=0), CACHE_FRAMES - old.shape[0]
                cache[i] = keep
                # time_conv (3,1,1): C -> 2C; the two channel halves are frames 2t and 2t+1 (vae.py:141-144), written
                # by the GEMM as two output segments
                tw, tb = self.w[name + ".time_conv.weight"], self.w[name + ".time_conv.bias"]
                doubled = torch.empty(2 * T, H, W, C, dtype=x.dtype, device=x.device)
                for t in range(T):
                    lo = max(0, t - pad)
                    self.ops.causal_conv3d(xin[lo:t + 3 - pad], max(0, pad - t), tw, tb, 3, 1,
                                           doubled[2 * t].view(-1, C), doubled[2 * t + 1].view(-1, C), seg_cols=C)
                x = doubled
            idx[0] += 1
        return self._conv(name + ".resample.1", x, 0, upsample=True)

    def _decoder(self, x: torch.Tensor, cache: List, idx: List[int]) -> torch.Tensor:
        """vae.py:423-472 for one latent frame x [1, h, w, 16] -> [T, 8h, 8w, 8] (3 channels used)."""
        x = self._cached_conv("decoder.conv1", x, cache, idx)
        x = self._res("decoder.middle.0", x, cache, idx)
        x = self._attn("decoder.middle.1", x)
        x = self._res("decoder.middle.2", x, cache, idx)
        for kind, name, _, _ in self.plan:
            x = self._res(name, x, cache, idx) if kind == "res" else self._upsample(name, kind, x, cache, idx)
        x = self._norm(x, self.w["decoder.head.0.gamma"], True)
        return self._cached_conv("decoder.head.2", x, cache, idx)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, use_cache: bool = False) -> torch.Tensor:
        """z [16, F, h, w] bf16 (one sample) -> fp32 pixels [1 + 4 (F - 1) (or 4 F when continuing), 3, 8h, 8w] in
        [-1, 1].  use_cache=True continues the video of the previous call (`cached_decode`, vae.py:571-593)."""
        if not self.w:
            raise RuntimeError("B200VAEDecoder: load_state_dict first")
        if z.dtype != torch.bfloat16 and getattr(self.ops, "requires_bf16", True):
            raise TypeError("B200 VAE decoder runs on bfloat16 latents")
        if use_cache:
            if self.cache is None:
                self.cache = [None] * self.cache_slots()
            cache = self.cache
        else:
            cache = [None] * self.cache_slots()
        _, F_, h, w = z.shape
        frames = []
        for i in range(F_):
            zi = z[:, i].contiguous()
            x = torch.empty(1, h, w, 16, dtype=z.dtype, device=z.device)
            self.ops.vae_latent_in(zi, self.mean, self.inv_std, self.w["conv2.weight"], self.w["conv2.bias"],
                                   x.view(h * w, 16))
            y = self._decoder(x, cache, [0])
            T, H, W, ld = y.shape
            px = torch.empty(T, 3, H, W, dtype=torch.float32, device=z.device)
            self.ops.vae_pixel_out(y.view(T * H * W, ld), px)
            frames.append(px)
        return torch.cat(frames, dim=0)


class B200VAEDecoderWrapper(torch.nn.Module):
    """The demo's streaming decoder (demo_utils/vae_block3.py:130-185): `forward(z, *feat_cache) -> (pixels, feat_cache)`
    with the 32 feature-cache tensors owned by the caller (channels-first `[1, C, 2, H, W]`, all zeros to start with --
    demo_utils/constant.py:5-38 -- in which case even the first frame is temporally upsampled)."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, device=None, ops=None):
        super().__init__()
        self.model = B200VAEDecoder(ops=ops, device=device)
        if state_dict is not None:
            self.model.load_state_dict(state_dict, strict=False)

    def forward(self, z: torch.Tensor, *feat_cache):
        """z [1, F, 16, h, w] -> (fp32 pixels [1, T, 3, 8h, 8w] in [-1, 1], list of updated cache tensors)."""
        assert z.shape[0] == 1, "the streaming decoder works on one video"
        dec = self.model
        if len(feat_cache) != dec.cache_slots():
            raise ValueError(f"expected {dec.cache_slots()} feature-cache entries, got {len(feat_cache)}")
        # caller layout [1, C, T, H, W] <-> channels-last frames [T, H, W, C]
        dec.cache = [c if c is None or isinstance(c, str) else c[0].permute(1, 2, 3, 0).contiguous() for c in feat_cache]
        pixels = dec.decode(z[0].permute(1, 0, 2, 3), use_cache=True)
        out_cache = [c if c is None or isinstance(c, str) else c.permute(3, 0, 1, 2).unsqueeze(0) for c in dec.cache]
        dec.cache = None
        return pixels.unsqueeze(0), out_cache


class B200VAEWrapper(torch.nn.Module):
    """`WanVAEWrapper` (utils/wan_wrapper.py:58-117), decode side."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, vae_path: Optional[str] = None, device=None,
                 ops=None):
        super().__init__()
        self.model = B200VAEDecoder(ops=ops, device=device)
        if state_dict is None and vae_path is not None:
            state_dict = torch.load(vae_path, map_location="cpu")
        if state_dict is not None:
            self.model.load_state_dict(state_dict, strict=False)

    def encode_to_latent(self, pixel):
        raise NotImplementedError("the VAE encoder is not on the B200 path (training-side only)")

    def decode_to_pixel(self, latent: torch.Tensor, use_cache: bool = False) -> torch.Tensor:
        """latent [B, F, 16, h, w] -> fp32 [B, T, 3, 8h, 8w] clamped to [-1, 1]."""
        if use_cache:
            assert latent.shape[0] == 1, "Batch size must be 1 when using cache"
        else:
            self.model.clear_cache()
        out = [self.model.decode(u.permute(1, 0, 2, 3), use_cache=use_cache) for u in latent]
        return torch.stack(out, dim=0)


# ---- synthetic weights / FLOP count for benchmarks (there is no network for the real checkpoint) ------------------
def random_decoder_weights(dec: "B200VAEDecoder", seed: int = 0):
    """(state_dict, conv shapes): fan-in scaled normal weights for every key the decoder expects."""
    g = torch.Generator().manual_seed(seed)
    dims = dec.dims
    shapes = {"conv2": (16, 16, 1, 1, 1), "decoder.conv1": (dims[0], 16, 3, 3, 3), "decoder.head.2": (3, dims[-1], 3, 3, 3),
              "decoder.middle.1.to_qkv": (3 * dims[0], dims[0], 1, 1), "decoder.middle.1.proj": (dims[0], dims[0], 1, 1)}

    def res(name, cin, cout):
        shapes[name + ".residual.2"] = (cout, cin, 3, 3, 3)
        shapes[name + ".residual.6"] = (cout, cout, 3, 3, 3)
        if cin != cout:
            shapes[name + ".shortcut"] = (cout, cin, 1, 1, 1)

    res("decoder.middle.0", dims[0], dims[0])
    res("decoder.middle.2", dims[0], dims[0])
    for kind, name, cin, cout in dec.plan:
        if kind == "res":
            res(name, cin, cout)
        else:
            shapes[name + ".resample.1"] = (cout, cin, 3, 3)
            if kind == "up3d":
                shapes[name + ".time_conv"] = (2 * cin, cin, 3, 1, 1)
    sd = {}
    for k in dec.expected_keys():
        if k.endswith("gamma"):
            sd[k] = None   # filled below from the conv it feeds
        elif k.endswith(".weight"):
            s = shapes[k[:-7]]
            fan = 1
            for d in s[1:]:
                fan *= d
            sd[k] = (torch.randn(s, generator=g) / fan ** 0.5).to(torch.bfloat16)
        else:
            sd[k] = (0.02 * torch.randn(shapes[k[:-5]][0], generator=g)).to(torch.bfloat16)
    for k in list(sd):
        if sd[k] is None:
            base = k[:-len(".gamma")]
            if base.endswith("residual.0"):
                c = shapes[base[:-1] + "2"][1]
            elif base.endswith("residual.3"):
                c = shapes[base[:-1] + "6"][1]
            elif base.endswith("norm"):
                c = dims[0]
            else:
                c = dims[-1]
            sd[k] = (1.0 + 0.05 * torch.randn(c, generator=g)).to(torch.bfloat16)
    return sd, shapes



def decode_flops(dec, shapes, frames, h, w):
    """2 * MACs of every convolution and of the middle attention for `frames` latent frames (the first yields 1 pixel
    frame, the others 4)."""
    def conv(name, voxels):
        s = shapes[name]
        k = 1
        for d in s[1:]:
            k *= d
        return 2.0 * voxels * s[0] * k

    total = 0.0
    for f in range(frames):
        T, H, W = 1, h, w
        fl = conv("decoder.conv1", T * H * W)
        for name in ("decoder.middle.0", "decoder.middle.2"):
            fl += conv(name + ".residual.2", T * H * W) + conv(name + ".residual.6", T * H * W)
        C = dec.dims[0]
        fl += conv("decoder.middle.1.to_qkv", H * W) + conv("decoder.middle.1.proj", H * W) + 4.0 * (H * W) ** 2 * C
        for kind, name, cin, cout in dec.plan:
            if kind == "res":
                fl += conv(name + ".residual.2", T * H * W) + conv(name + ".residual.6", T * H * W)
                if name + ".shortcut" in shapes:
                    fl += conv(name + ".shortcut", T * H * W)
            else:
                if kind == "up3d" and f > 0:
                    fl += conv(name + ".time_conv", T * H * W)
                    T *= 2
                H, W = 2 * H, 2 * W
                fl += conv(name + ".resample.1", T * H * W)
        fl += conv("decoder.head.2", T * H * W)
        total += fl
    return total
