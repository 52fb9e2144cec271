"""SURVEY.md section 8f rank 1 on ONE GPU: Wan2.1 VAE decode of a full video -- 21 latent frames 60x104 -> 81 frames
480x832 -- through B200VAEWrapper (random-init decoder weights of the real architecture, synthetic latents).
Prints one JSON line: seconds per video, frames/s, algorithmic TFLOP of the convolutions / attention and TFLOP/s.

    python tools/vae_bench.py [--frames 21] [--hw 60 104] [--runs 2]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_forcing_b200.ops import CudaOps            # noqa: E402
from self_forcing_b200.vae import B200VAEWrapper     # noqa: E402


def random_decoder_weights(dec, seed=0):
    """Fan-in scaled normal weights for every key the decoder expects (shapes from the reference architecture)."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=21)
    ap.add_argument("--hw", type=int, nargs=2, default=[60, 104])
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--gather", action="store_true", help="gather + GEMM for every convolution (no implicit GEMM)")
    ap.add_argument("--breakdown", action="store_true", help="one more decode with per-launch CUDA events")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = CudaOps()
    wrap = B200VAEWrapper(device=dev, ops=ops)
    sd, shapes = random_decoder_weights(wrap.model)
    wrap.model.load_state_dict(sd)
    wrap.model.implicit_conv = not a.gather
    h, w = a.hw
    lat = torch.randn(1, a.frames, 16, h, w, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).to(dev)
    out = wrap.decode_to_pixel(lat[:, :2])                      # warm-up: both frame kinds
    torch.cuda.synchronize()
    before = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.runs):
        out = wrap.decode_to_pixel(lat)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / a.runs
    fl = decode_flops(wrap.model, shapes, a.frames, h, w)
    breakdown = None
    if a.breakdown:
        ops.start_profile()
        wrap.decode_to_pixel(lat)
        agg = {}
        for name, tag, ms in ops.stop_profile():
            key = name if name != "causal_conv3d" else "conv " + "x".join(str(v) for v in tag[1:])
            n, t = agg.get(key, (0, 0.0))
            agg[key] = (n + 1, t + ms)
        total = sum(t for _, t in agg.values())
        breakdown = {k: dict(launches=n, ms=round(t, 2), share=round(t / total, 3))
                     for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]}
    print(json.dumps(dict(
        workload=f"wan2.1 vae decode, {a.frames} latent frames {h}x{w} -> {out.shape[1]} frames {8 * h}x{8 * w}, "
                 "random-init decoder, channels-last, " + ("gather + tcgen05 GEMM" if a.gather else "implicit-GEMM tcgen05 convolutions"),
        seconds_per_video=sec, frames_per_s=out.shape[1] / sec, tflop=fl / 1e12, tflops=fl / sec / 1e12,
        launches=(ops.launches - before) // a.runs, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
        implicit_conv=not a.gather, finite=bool(torch.isfinite(out).all()), out_shape=list(out.shape),
        breakdown=breakdown)), flush=True)


if __name__ == "__main__":
    main()
