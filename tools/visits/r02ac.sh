#!/bin/bash
# A/B: whole-item tile mode vs split mode at S = 9360 (74 KV tiles) -- isolated launches, then the rollout
OUT=gpurun_out; mkdir -p $OUT
python - <<'PY'
import math, os, sys, json, subprocess
code = r'''
import math, sys, torch
sys.path.insert(0, ".")
from self_forcing_b200.ops import CudaOps
ops = CudaOps()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for S in (14040, 18720, 23400):
    q = torch.randn(1, 4680, 12, 128, device="cuda").bfloat16(); k = torch.randn(1, S, 12, 128, device="cuda").bfloat16(); v = torch.randn_like(k); o = torch.empty_like(q)
    for _ in range(3): ops.attention(q, k, v, o, 1 / math.sqrt(128))
    ts = []
    for _ in range(7):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); ops.attention(q, k, v, o, 1 / math.sqrt(128)); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); print(S, round(ts[3] * 1e3, 1), "us")
'''
for env in ({}, {"SFB_ATTN_MIN_SPLIT": "120"}, {"SFB_ATTN_MIN_SPLIT": "160"}, {"SFB_ATTN_MIN_SPLIT": "200"}):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=120)
    print(env or "default(48)", r.stdout.strip().replace("\n", " | "), r.stderr[-200:] if r.returncode else "")
PY
