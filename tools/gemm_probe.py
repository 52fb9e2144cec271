"""GEMM probe at the rollout's projection shapes: cold (L2 flushed) and hot (back-to-back, operands L2-resident) device
times of the product kernel vs cuBLAS; with SFB_GEMM_TIMING=1 the pair kernel also prints its per-CTA clock64 timeline.

    python tools/gemm_probe.py [name-substring ...]        e.g.  SFB_GEMM_TIMING=1 python tools/gemm_probe.py o_proj
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_forcing_b200.ops import CudaOps  # noqa: E402

BF = torch.bfloat16
SHAPES = [  # name, M, N, K, epilogue
    ("qkv", 4680, 4608, 1536, 0), ("o_proj", 4680, 1536, 1536, 3), ("cross_q", 4680, 1536, 1536, 0),
    ("cross_o", 4680, 1536, 1536, 2), ("ffn1", 4680, 8960, 1536, 1), ("ffn2", 4680, 1536, 8960, 3),
    ("frame_qkv", 1560, 4608, 1536, 0), ("frame_o_proj", 1560, 1536, 1536, 3), ("frame_ffn1", 1560, 8960, 1536, 1),
    ("frame_ffn2", 1560, 1536, 8960, 3), ("ulysses4_ffn2", 1170, 1536, 8960, 3),
]


def event_ms(fn, reps, flush=None):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def hot_ms(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    ops = CudaOps()
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    block_n = int(os.environ.get("SFB_PROBE_BLOCK_N", "0"))
    timing = os.environ.get("SFB_GEMM_TIMING") is not None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for name, M, N, K, epi in SHAPES:
        if only and not any(o in name for o in only):
            continue
        x = torch.randn(M, K, device="cuda").to(BF)
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(BF)
        b = torch.randn(N, device="cuda").to(BF)
        res = torch.randn(M, N, device="cuda").to(BF)
        gate = torch.randn(3, N, device="cuda").to(BF)
        y = torch.empty(M, N, device="cuda", dtype=BF)
        kw = dict(epilogue=epi, block_n=block_n)
        if epi in (2, 3):
            kw["residual"] = res
        if epi == 3:
            kw.update(gate=gate, gate_stride=N, rows_per_gate=max(1, M // 3))
        fn = lambda: ops.gemm(x, w, b, y, **kw)   # noqa: E731
        if timing:
            sys.stderr.write(f"--- {name} (second launch: L2 hot)\n")
            fn(); fn()
            torch.cuda.synchronize()
            continue
        ref = lambda: torch.nn.functional.linear(x, w, b)   # noqa: E731
        fl = 2.0 * M * N * K
        for _ in range(3):
            fn(); ref()
        rec = dict(shape=name, M=M, N=N, K=K, epi=epi, block_n=block_n,
                   cold_us=event_ms(fn, 10, flush) * 1e3, hot_us=hot_ms(fn) * 1e3,
                   cublas_cold_us=event_ms(ref, 10, flush) * 1e3, cublas_hot_us=hot_ms(ref) * 1e3)
        rec.update(cold_tflops=fl / rec["cold_us"] / 1e6, hot_tflops=fl / rec["hot_us"] / 1e6,
                   cublas_cold_tflops=fl / rec["cublas_cold_us"] / 1e6, cublas_hot_tflops=fl / rec["cublas_hot_us"] / 1e6)
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)
        out.append(rec)
    if out:
        tag = os.environ.get("SFB_PROBE_TAG", "")
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"gemm_probe{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
