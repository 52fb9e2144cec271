"""Full-size parity probe on ONE GPU (not part of the test suite: the oracle needs seconds per forward here).

Product (CUDA kernels through the C ABI) against the oracle run in PyTorch eager mode on the same GPU, at BASELINE's
full size -- Wan2.1-T2V-1.3B architecture (30 layers), one chunk of 3 latent frames (L = 4680 tokens):

  * forward 1: chunk 0 at t = 1000 (S = 4680), forward 2: its clean-context refresh, forward 3: chunk 1 at t = 937.5
    attending to chunk 0 (S = 9360) -- flow / x0 / layer-0 and last-layer K/V rel-L2 after each;
  * `--rollout`: the whole 35-forward rollout on both sides with the same re-noise stream, latents rel-L2.

The randomly initialised 30-layer network amplifies bf16 rounding noise, so `--fp32-yardstick` also runs the oracle in
fp32 (TF32 off) for the three forwards and reports how far the oracle's own bf16 run is from it: the product should be
no further from the fp32 result than that.

    python tools/fullsize_parity.py [--layers 30] [--rollout] [--fp32-yardstick]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle import causal_wan_oracle as O                                       # noqa: E402
from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper        # noqa: E402


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def three_forwards(call, new_caches, x, pe, fs):
    """[(flow, x0, k_first, v_last)] after each of the three forwards; `call(x, ts, kv, ca, start)`."""
    kv, ca = new_caches()
    out = []
    for t, lo, start in ((1000.0, 0, 0), (0.0, 0, 0), (937.5, 3, 3)):
        xin = x[:, lo:lo + 3]
        ts = torch.full((1, 3), t, device=x.device)
        flow, x0 = call(xin, ts, kv, ca, start * fs)
        out.append(tuple(t.float().cpu().clone() for t in (flow, x0, kv[0]["k"][:, :9360], kv[-1]["v"][:, :9360])))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--rollout", action="store_true")
    ap.add_argument("--fp32-yardstick", action="store_true")
    ap.add_argument("--device", default="cuda", help="'cpu' = dry run of this script through the torch test double")
    a = ap.parse_args()
    ops = None
    if a.device == "cpu":
        from _torch_ops import TorchOps
        dev, ops = torch.device("cpu"), TorchOps()
    else:
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    gen = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B, num_layers=a.layers), timestep_shift=5.0, device=dev, init_seed=0,
                               ops=ops)
    gen.model.use_cuda_graphs = False
    cfg = O.OracleConfig(dim=1536, ffn_dim=8960, num_heads=12, num_layers=a.layers)
    params = dict(gen.model.state_dict())
    ow = O.OracleWrapper(params, cfg, 5.0)
    fs = 1560
    pe = torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    x = torch.randn(1, 6, 16, 60, 104, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(dev)

    def caches(dtype=torch.bfloat16):
        return (O.new_kv_cache(cfg, 1, fs, dtype, dev, cache_tokens=9360), O.new_crossattn_cache(cfg, 1, dtype, dev))

    with torch.no_grad():
        prod = three_forwards(lambda xi, ts, kv, ca, st: gen(xi, {"prompt_embeds": pe}, ts, kv_cache=kv, crossattn_cache=ca,
                                                             current_start=st), caches, x, pe, fs)
        orac = three_forwards(lambda xi, ts, kv, ca, st: ow(xi, pe, ts, kv, ca, st), caches, x, pe, fs)
    names = ("flow", "x0", "k_layer0", "v_last_layer")
    report = {"layers": a.layers, "forwards": []}
    for i, (p_, o_) in enumerate(zip(prod, orac)):
        report["forwards"].append({n: rel_l2(pv, ov) for n, pv, ov in zip(names, p_, o_)})
    if a.fp32_yardstick:
        ow32 = O.OracleWrapper({k: v.float() for k, v in params.items()}, cfg, 5.0)
        with torch.no_grad():
            exact = three_forwards(lambda xi, ts, kv, ca, st: ow32(xi.float(), pe.float(), ts, kv, ca, st),
                                   lambda: caches(torch.float32), x, pe, fs)
        report["vs_fp32"] = [{"product_" + n: rel_l2(pv, ev) for n, pv, ev in zip(names, p_, e_)} |
                             {"oracle_bf16_" + n: rel_l2(ov, ev) for n, ov, ev in zip(names, o_, e_)}
                             for p_, o_, e_ in zip(prod, orac, exact)]
    if a.rollout:
        import types
        from helpers import _IdentityVAE, _TextEncoder
        from self_forcing_b200.pipeline import CausalInferencePipeline
        noise = torch.randn(1, 21, 16, 60, 104, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).to(dev)
        args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True, num_frame_per_block=3,
                                     independent_first_frame=False, context_noise=0, model_kwargs={})
        pipe = CausalInferencePipeline(args, dev, generator=gen, text_encoder=_TextEncoder(pe), vae=_IdentityVAE())
        steps = O.warp_denoising_steps(ow.scheduler, [1000, 750, 500, 250])
        with torch.no_grad():
            torch.manual_seed(11)
            _, lat = pipe.inference(noise, ["synthetic"], return_latents=True)
            torch.manual_seed(11)
            tr = O.rollout(ow, noise, pe, steps, 3)
        report["rollout_latents_rel_l2"] = rel_l2(lat, tr.latents)
        report["rollout_index"] = [int(pipe.kv_cache1[0]["global_end_index"]), tr.index_trace[-1][0]]
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
