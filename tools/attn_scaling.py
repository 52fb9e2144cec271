"""Diagnostic: attention time vs number of CTAs (SMs) used, to separate per-SM efficiency from chip-wide limits
(power / clocks / L2).  Prints ms and TFLOP/s for grid caps given on the command line."""
import math, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from self_forcing_b200.ops import CudaOps
    ops = CudaOps()
    Lq, S, H = 4680, int(sys.argv[2]), 12
    q = torch.randn(1, Lq, H, 128, device="cuda").bfloat16(); k = torch.randn(1, S, H, 128, device="cuda").bfloat16()
    v = torch.randn(1, S, H, 128, device="cuda").bfloat16(); o = torch.empty_like(q)
    reps = int(sys.argv[3])
    for _ in range(3): ops.attention(q, k, v, o, 1 / math.sqrt(128))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): ops.attention(q, k, v, o, 1 / math.sqrt(128))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps(dict(grid=os.environ.get("SFB_ATTN_GRID", "all"), S=S, reps=reps, ms=ms, tflops=4.0 * Lq * S * H * 128 / ms / 1e9)))
else:
    for S in (32760,):
        for reps in (1, 200):
            for g in ("148", "111", "74", "37"):
                env = dict(os.environ, SFB_ATTN_GRID=g)
                subprocess.run([sys.executable, __file__, "child", str(S), str(reps)], env=env)
