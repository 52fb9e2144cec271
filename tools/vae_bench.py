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
from self_forcing_b200.vae import B200VAEWrapper, decode_flops, random_decoder_weights     # noqa: E402


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
        frac_of_sustained_peak=fl / sec / 1e12 / 1397.2,
        launches=(ops.launches - before) // a.runs, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
        implicit_conv=not a.gather, finite=bool(torch.isfinite(out).all()), out_shape=list(out.shape),
        breakdown=breakdown)), flush=True)


if __name__ == "__main__":
    main()
