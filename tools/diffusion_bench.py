"""SURVEY.md section 8f rank 2 on ONE GPU: the 50-step guided sampler (CFG + UniPC) of the 1.3B model at full size
-- 21 latent frames (81 pixel frames) 480x832, chunks of 3 frames, 50 steps per chunk, guidance 3, shift 5
(configs/self_forcing_dmd.yaml:18 guidance_scale; causal_diffusion_inference.py:66 sampling_steps).  Random-init
weights, synthetic embeddings.  Prints one JSON line: seconds per video, frames/s, ms per guided step (one batch-2
forward + one fused sampler-step kernel), model TFLOP/s.

    python tools/diffusion_bench.py [--frames 21] [--steps 50] [--runs 1]
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_forcing_b200.diffusion_pipeline import CausalDiffusionInferencePipeline   # noqa: E402
from self_forcing_b200.ops import CudaOps                                           # noqa: E402
from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper            # noqa: E402


class _NoVAE:
    def decode_to_pixel(self, x, use_cache=False):
        return x


def forward_flops(L: int, S: int, batch: int) -> float:
    """NL * [12 L C^2 + 4 L C FFN + 4 L S C + 4 L T C] per sample."""
    NL, C, FFN, T = 30, 1536, 8960, 512
    return batch * NL * (12.0 * L * C * C + 4.0 * L * C * FFN + 4.0 * L * S * C + 4.0 * L * T * C)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=21)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--chunk-frames", type=int, default=3)
    ap.add_argument("--runs", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = CudaOps()
    gen = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B), timestep_shift=5.0, device=dev, init_seed=0, ops=ops)
    pe = torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    neg = torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16).to(dev)
    noise = torch.randn(1, a.frames, 16, 60, 104, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(dev)
    args = types.SimpleNamespace(num_train_timestep=1000, timestep_shift=5.0, guidance_scale=3.0, negative_prompt="neg",
                                 num_frame_per_block=a.chunk_frames, independent_first_frame=False, model_kwargs={},
                                 sampling_steps=a.steps)
    enc = lambda text_prompts: {"prompt_embeds": neg if text_prompts[0] == "neg" else pe}   # noqa: E731
    pipe = CausalDiffusionInferencePipeline(args, dev, generator=gen, text_encoder=enc, vae=_NoVAE())
    # warm-up: the first chunk of a video (graphs are captured on the second step of every chunk anyway)
    pipe.inference(noise[:, :a.chunk_frames], ["synthetic"], None, None, None, return_latents=True)
    torch.cuda.synchronize()
    before = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.runs):
        _, lat = pipe.inference(noise, ["synthetic"], None, None, None, return_latents=True)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / a.runs
    chunks = a.frames // a.chunk_frames
    L = a.chunk_frames * 1560
    flops = sum((a.steps + 1) * forward_flops(L, (c + 1) * L, 2) for c in range(chunks))
    pixel_frames = (a.frames - 1) * 4 + 1
    print(json.dumps(dict(
        workload=f"wan2.1-t2v-1.3b causal 50-step sampler: CFG 3.0 + UniPC, {a.frames} latent frames 60x104, "
                 f"{a.chunk_frames} frames/chunk, {a.steps} steps/chunk, cond+uncond as one batch-2 forward",
        seconds_per_video=sec, frames_per_s=pixel_frames / sec, latent_frames_per_s=a.frames / sec,
        ms_per_guided_step=sec * 1e3 / (chunks * (a.steps + 1)), forwards_batch2=chunks * (a.steps + 1),
        model_tflop=flops / 1e12, model_tflops=flops / sec / 1e12, frac_of_sustained_peak=flops / sec / 1e12 / 1397.2,
        launches=(ops.launches - before) // a.runs, finite=bool(torch.isfinite(lat.float()).all()),
        cuda_graphs=bool(gen.model.use_cuda_graphs))), flush=True)


if __name__ == "__main__":
    main()
