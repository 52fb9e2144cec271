"""Diagnostic: how much of a forward is inter-kernel gap?  Times one full-depth cached forward (chunk 3 of the rollout,
KV window 18720) launched eagerly vs replayed as a CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper

dev = torch.device("cuda")
w = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B), timestep_shift=5.0, device=dev, init_seed=0)
m = w.model
kv = m.allocate_kv_cache(1, 32760, torch.bfloat16, dev)
ca = [dict(k=torch.zeros(1, 512, 12, 128, dtype=torch.bfloat16, device=dev), v=torch.zeros(1, 512, 12, 128, dtype=torch.bfloat16, device=dev), is_init=False) for _ in range(30)]
pe = torch.randn(1, 512, 4096, device=dev).bfloat16()
x = torch.randn(1, 3, 16, 60, 104, device=dev).bfloat16()
t = torch.full((1, 3), 937.5, device=dev)
L = 4680
for c in range(4):   # fill the cache up to chunk 3
    w(x, {"prompt_embeds": pe}, t, kv_cache=kv, crossattn_cache=ca, current_start=c * L)
torch.cuda.synchronize()

def fwd():
    return w(x, {"prompt_embeds": pe}, t, kv_cache=kv, crossattn_cache=ca, current_start=3 * L)

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

eager = timeit(fwd)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fwd(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        out = fwd()
torch.cuda.synchronize()
graph = timeit(g.replay)
eager2 = timeit(fwd)
print(f"forward (S=18720): eager {eager:.3f} ms, eager again {eager2:.3f} ms, cuda-graph replay {graph:.3f} ms, launches/forward {m.ops.launches // 10}")
