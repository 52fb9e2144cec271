"""Diagnostic: how much of a rollout is spent BETWEEN the per-forward CUDA graphs (host prelude, sampler kernels, output
clones)?  Brackets every graph replay of one rollout with CUDA events and compares their sum with the rollout's time."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from self_forcing_b200.ops import CudaOps  # noqa: E402
from self_forcing_b200.pipeline import CausalInferencePipeline  # noqa: E402
from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = CudaOps()
    gen = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B), timestep_shift=bench.SHIFT, device=dev, init_seed=0, ops=ops)
    pe = torch.randn(1, bench.T_CTX, 4096, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    noise = torch.randn(1, bench.LAT_FRAMES, 16, bench.LAT_H, bench.LAT_W, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(dev)
    pargs = types.SimpleNamespace(denoising_step_list=bench.DENOISE_STEPS, warp_denoising_step=True, num_frame_per_block=3,
                                  independent_first_frame=False, context_noise=0, model_kwargs={}, skip_refresh_tail=False)
    pipe = CausalInferencePipeline(pargs, dev, generator=gen, text_encoder=lambda text_prompts: {"prompt_embeds": pe}, vae=bench._NoVAE())
    for _ in range(3):
        pipe.inference(noise, ["synthetic"], return_latents=True)
    torch.cuda.synchronize()
    pairs = []
    orig = torch.cuda.CUDAGraph.replay

    def timed_replay(self):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(self)
        b.record()
        pairs.append((a, b))

    res = {}
    for tag, fn in (("plain", orig), ("bracketed", timed_replay)):
        torch.cuda.CUDAGraph.replay = fn
        pairs.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            pipe.inference(noise, ["synthetic"], return_latents=True)
        e1.record()
        torch.cuda.synchronize()
        res[tag + "_ms_per_rollout"] = e0.elapsed_time(e1) / 2
        if pairs:
            res["replays_per_rollout"] = len(pairs) // 2
            res["sum_of_replays_ms_per_rollout"] = sum(a.elapsed_time(b) for a, b in pairs) / 2
    torch.cuda.CUDAGraph.replay = orig
    res["between_graphs_ms_per_rollout"] = res["bracketed_ms_per_rollout"] - res["sum_of_replays_ms_per_rollout"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
