"""SURVEY.md section 8f rank 4 on ONE GPU: the UMT5-XXL text encoder (24 layers, dim 4096, ffn 10240, 64 heads) on one
512-token prompt through B200T5Encoder, random-init weights of that architecture (vocabulary cut to 32k rows: the
embedding is a table lookup).  Prints ms per prompt and the projections' TFLOP/s.

    python tools/t5_bench.py [--layers 24] [--runs 3]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_forcing_b200.ops import CudaOps          # noqa: E402
from self_forcing_b200.t5 import B200T5Encoder     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--runs", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = CudaOps()
    vocab, C, Fd, H, L = 32768, 4096, 10240, 64, 512
    enc = B200T5Encoder(vocab=vocab, dim=C, dim_attn=C, dim_ffn=Fd, num_heads=H, num_layers=a.layers, ops=ops, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g, device=dev) * scale).to(torch.bfloat16)
    sd = {"token_embedding.weight": rnd(vocab, C), "norm.weight": rnd(C, scale=0.05) + 1}
    for i in range(a.layers):
        b = f"blocks.{i}."
        sd[b + "norm1.weight"], sd[b + "norm2.weight"] = rnd(C, scale=0.05) + 1, rnd(C, scale=0.05) + 1
        for n in "qkvo":
            sd[b + f"attn.{n}.weight"] = rnd(C, C, scale=C ** -0.5)
        sd[b + "pos_embedding.embedding.weight"] = rnd(32, H, scale=0.5)
        sd[b + "ffn.gate.0.weight"], sd[b + "ffn.fc1.weight"] = rnd(Fd, C, scale=C ** -0.5), rnd(Fd, C, scale=C ** -0.5)
        sd[b + "ffn.fc2.weight"] = rnd(C, Fd, scale=Fd ** -0.5)
    enc.load_state_dict(sd)
    del sd
    ids = torch.randint(1, vocab, (1, L), device=dev)
    mask = torch.ones(1, L, dtype=torch.long, device=dev)
    mask[0, 300:] = 0
    for _ in range(3):      # eager, capture, first replay
        out = enc(ids, mask)
    torch.cuda.synchronize()
    before = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.runs):
        out = enc(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.runs
    flops = a.layers * (2.0 * L * C * (4 * C + 3 * Fd) + 4.0 * L * L * C)
    print(json.dumps(dict(workload=f"umt5-xxl encoder, {a.layers} layers, one 512-token prompt", ms_per_prompt=ms,
                          tflop=flops / 1e12, tflops=flops / ms / 1e9, launches=(ops.launches - before) // a.runs,
                          finite=bool(torch.isfinite(out.float()).all()))), flush=True)


if __name__ == "__main__":
    main()
