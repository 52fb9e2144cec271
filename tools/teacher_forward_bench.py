"""BASELINE config 5 on ONE GPU: the Wan2.1 14B bidirectional teacher forward (random-init weights of that
architecture, x [16, 21, 60, 104], one timestep) through B200WanModel.  Prints ms per forward and model TFLOP/s
(1676 TFLOP algorithmic).  The 8-GPU Ulysses variant is `torchrun ... --ulysses`."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_forcing_b200.model import B200WanModel          # noqa: E402
from self_forcing_b200.wrapper import WAN_T2V_14B         # noqa: E402


def main():
    ulysses = "--ulysses" in sys.argv
    layers = int(os.environ.get("TEACHER_LAYERS", "40"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sp = None
    if ulysses:
        import torch.distributed as dist
        from self_forcing_b200.ulysses import UlyssesGroup
        dist.init_process_group("nccl", device_id=dev)
        sp = UlyssesGroup(device=dev)
    cfg = dict(WAN_T2V_14B, num_layers=layers)
    with torch.device(dev):
        model = B200WanModel(**cfg).to(torch.bfloat16)
    model.init_weights(0)
    if sp is not None:
        model.enable_ulysses(sp)
    x = torch.randn(1, 16, 21, 60, 104, device=dev).bfloat16()
    ctx = torch.randn(1, 512, 4096, device=dev).bfloat16()
    t = torch.tensor([500.0], device=dev)
    L, C, FFN, T = 32760, 5120, 13824, 512
    flops = layers * (12.0 * L * C * C + 4.0 * L * C * FFN + 4.0 * L * L * C + 4.0 * L * T * C + 4.0 * T * C * C)
    out = model(x, t=t, context=ctx, seq_len=L)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 2
    for _ in range(n):
        out = model(x, t=t, context=ctx, seq_len=L)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    if sp is None or sp.rank == 0:
        print(json.dumps(dict(workload="wan2.1-14b bidirectional forward, 21x60x104 latents", layers=layers,
                              gpus=1 if sp is None else sp.world, ms_per_forward=ms, tflop=flops / 1e12,
                              tflops=flops / ms / 1e9, finite=bool(torch.isfinite(out.float()).all()))), flush=True)
    if sp is not None:
        os._exit(0)


if __name__ == "__main__":
    main()
