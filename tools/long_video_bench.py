"""Long-video (rolling KV window) rollout on one B200: the mode where the roll / sink index arithmetic of
wan/modules/causal_model.py:203-236 is exercised at scale.

Wan2.1-T2V-1.3B architecture (random init), chunks of 3 latent frames, local_attn_size = 21 frames with a 3-frame sink
(KV cache = 32760 tokens, the same memory as the 21-frame headline rollout), `--frames` latent frames in total
(default 201 = 801 pixel frames).  Reports frames/s, the steady-state time per chunk and how many CUDA graphs the
whole video needed (one per forward kind in the steady state).

    python tools/long_video_bench.py [--frames 201] [--local 21] [--sink 3] [--runs 2]
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_forcing_b200.ops import CudaOps                                       # noqa: E402
from self_forcing_b200.pipeline import CausalInferencePipeline                  # noqa: E402
from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper        # noqa: E402


class _NoVAE:
    def decode_to_pixel(self, x, use_cache=False):
        return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=201)
    ap.add_argument("--local", type=int, default=21)
    ap.add_argument("--sink", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=3)
    ap.add_argument("--runs", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = CudaOps()
    gen = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B), timestep_shift=5.0, device=dev, init_seed=0, ops=ops,
                               local_attn_size=a.local, sink_size=a.sink)
    pe = torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    noise = torch.randn(1, a.frames, 16, 60, 104, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(dev)
    pargs = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                  num_frame_per_block=a.chunk, independent_first_frame=False, context_noise=0, model_kwargs={})
    pipe = CausalInferencePipeline(pargs, dev, generator=gen, text_encoder=lambda text_prompts: {"prompt_embeds": pe}, vae=_NoVAE())
    torch.manual_seed(7)
    recs = []
    for run in range(a.runs):
        launches0 = ops.launches
        _, lat = pipe.inference(noise, ["synthetic"], return_latents=True, profile=True)
        prof = pipe.last_profile
        blocks = prof["block_ms"]
        steady = blocks[len(blocks) // 2:]
        recs.append(dict(run=run, diffusion_ms=prof["diffusion_ms"], chunks=len(blocks),
                         first_chunk_ms=blocks[0], steady_chunk_ms=sum(steady) / len(steady),
                         pixel_frames=(a.frames - 1) * 4 + 1,
                         frames_per_s=((a.frames - 1) * 4 + 1) / (prof["diffusion_ms"] / 1e3),
                         launches=ops.launches - launches0,
                         graphs=sum(1 for g in gen.model._graphs.values() if g != "seen"),
                         index=[int(pipe.kv_cache1[0]["global_end_index"]), int(pipe.kv_cache1[0]["local_end_index"])],
                         finite=bool(torch.isfinite(lat.float()).all())))
    out = dict(workload=f"wan2.1-t2v-1.3b self-forcing rolling-window rollout: {a.frames} latent frames, chunks of {a.chunk}, "
                        f"local_attn_size {a.local} frames, sink {a.sink} frames, 4 steps + refresh per chunk, batch 1",
               kv_cache_tokens=pipe.kv_cache1[0]["k"].shape[1], runs=recs,
               note="run 0 captures the CUDA graphs (one per chunk position while the window fills, then one per forward kind "
                    "for every later chunk); run 1 replays only")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
