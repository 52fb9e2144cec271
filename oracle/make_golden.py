"""Generates tests/golden/* by running the UNMODIFIED reference (imported from /root/reference
through oracle/ref_shim.py) on seeded synthetic inputs.  Only runnable in the build container.

    python -m oracle.make_golden

Fixtures (all inputs are re-creatable from seeds, so only outputs are stored):
  scheduler_tables.pt   FlowMatchScheduler sigmas/timesteps for shift 5 and 8 + warped step lists
  rollout_tiny.pt       CausalInferencePipeline.inference latents, 2-layer model at full width:
                          a) 3 frames, 1 frame/block, independent_first_frame, shift 8 (configs/tiny_test.yaml)
                          b) 6 frames, 3 frames/block (chunk-wise), shift 5 (self_forcing_dmd.yaml sampler)
  model_rolling.pt      CausalWanModel.forward with a rolling + sink KV cache on a small grid:
                          per-forward flow, final K/V caches, (global_end, local_end) trace
  block_masks.pt        BlockMask tables (kv_num_blocks, kv_indices, full_*) of the three mask builders
"""
from __future__ import annotations

import contextlib
import io
import os
import types

import torch

from oracle import causal_wan_oracle as O
from oracle import ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class SeededNoise:
    """Device-independent stand-in for torch.randn_like: draws on CPU from a seeded generator."""

    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def __call__(self, like: torch.Tensor, **kw) -> torch.Tensor:
        return torch.randn(like.shape, generator=self.g, dtype=torch.float32).to(like.dtype).to(like.device)


@contextlib.contextmanager
def patched_randn_like(seed: int):
    orig = torch.randn_like
    torch.randn_like = SeededNoise(seed)
    try:
        yield
    finally:
        torch.randn_like = orig


def synthetic_inputs(batch: int, frames: int, H: int = 60, W: int = 104, text_dim: int = 4096):
    pe = torch.randn(batch, 512, text_dim, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    noise = torch.randn(batch, frames, 16, H, W, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    return pe, noise


class _TextEncoder(torch.nn.Module):
    def __init__(self, pe):
        super().__init__()
        self.pe = pe

    def forward(self, text_prompts):
        return {"prompt_embeds": self.pe}


class _IdentityVAE(torch.nn.Module):
    def decode_to_pixel(self, x, use_cache=False):
        return x


ROLLOUT_CASES = {
    "tiny_test_yaml": dict(frames=3, num_frame_per_block=1, independent_first_frame=True, shift=8.0),
    "chunkwise": dict(frames=6, num_frame_per_block=3, independent_first_frame=False, shift=5.0),
    # video continuation: 3 given latent frames are cached at t = 0 (causal_inference.py:135-169), 3 more are generated
    "continuation": dict(frames=3, num_frame_per_block=3, independent_first_frame=False, shift=5.0, initial_frames=3),
}


# the rolling KV window driven through the PIPELINE (the model-level fixture below uses a small grid): 4 chunks of one frame,
# a 2-frame local window and a 1-frame attention sink -> the cache rolls on chunks 3 and 4 (causal_model.py:207-221)
ROLLING_ROLLOUT_CASES = {
    "rolling_window": dict(frames=4, num_frame_per_block=1, independent_first_frame=False, shift=5.0, local_attn_size=2,
                           sink_size=1),
}


def rollout_cfg(case: dict) -> O.OracleConfig:
    return O.OracleConfig(**O.WAN_TINY, local_attn_size=case.get("local_attn_size", -1), sink_size=case.get("sink_size", 0))


def initial_latent_for(case: dict, H: int = 60, W: int = 104):
    n = case.get("initial_frames", 0)
    if not n:
        return None
    return torch.randn(1, n, 16, H, W, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16)


def reference_rollout(ref, case: dict, params, cfg: O.OracleConfig):
    w = ref_shim.make_reference_wrapper(ref, cfg.reference_kwargs(), case["shift"])
    w.model.load_state_dict(params, strict=False)
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=case["num_frame_per_block"],
                                 independent_first_frame=case["independent_first_frame"], context_noise=0,
                                 model_kwargs={})
    pe, noise = synthetic_inputs(1, case["frames"])
    with contextlib.redirect_stdout(io.StringIO()):
        pipe = ref.CausalInferencePipeline(args, "cpu", generator=w, text_encoder=_TextEncoder(pe), vae=_IdentityVAE())
        pipe.num_transformer_blocks = cfg.num_layers
        with torch.no_grad(), patched_randn_like(3):
            _, lat = pipe.inference(noise, ["synthetic"], return_latents=True, initial_latent=initial_latent_for(case))
    idx = (int(pipe.kv_cache1[0]["global_end_index"]), int(pipe.kv_cache1[0]["local_end_index"]))
    return lat, idx


ROLLING = dict(dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_dim=512, local_attn_size=3, sink_size=1,
               frame_hw=(8, 12), chunks=6, forwards_per_chunk=2)


def rolling_model_inputs():
    r = ROLLING
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 16, r["chunks"], *r["frame_hw"], generator=g).to(torch.bfloat16)
    ctx = torch.randn(1, 512, r["text_dim"], generator=g).to(torch.bfloat16)
    return x, ctx


def rolling_cfg() -> O.OracleConfig:
    r = ROLLING
    return O.OracleConfig(dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                          text_dim=r["text_dim"], local_attn_size=r["local_attn_size"], sink_size=r["sink_size"])


def reference_rolling(ref):
    r = ROLLING
    cfg = rolling_cfg()
    params = O.make_random_params(cfg, seed=5)
    torch.manual_seed(0)
    model = ref.CausalWanModel(**cfg.reference_kwargs())
    model.load_state_dict(params, strict=False)
    model = model.to(torch.bfloat16).eval()
    x, ctx = rolling_model_inputs()
    ft = (r["frame_hw"][0] // 2) * (r["frame_hw"][1] // 2)
    kv = O.new_kv_cache(cfg, 1, ft, torch.bfloat16, "cpu", cache_tokens=r["local_attn_size"] * ft)
    ca = O.new_crossattn_cache(cfg, 1, torch.bfloat16, "cpu")
    flows, trace = [], []
    with torch.no_grad():
        for c in range(r["chunks"]):
            for k in range(r["forwards_per_chunk"]):
                t = torch.full((1, 1), 1000.0 - 300.0 * k)
                f = model(x[:, :, c:c + 1], t=t, context=ctx, seq_len=32760, kv_cache=kv, crossattn_cache=ca,
                          current_start=c * ft, cache_start=None)
                flows.append(f.clone())
                trace.append((int(kv[0]["global_end_index"]), int(kv[0]["local_end_index"])))
    return dict(flows=torch.stack(flows), trace=trace, k=[c["k"].clone() for c in kv], v=[c["v"].clone() for c in kv])


BIDIR = dict(dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_dim=512, frames=2, frame_hw=(8, 12), batch=2)


def bidirectional_inputs():
    r = BIDIR
    g = torch.Generator().manual_seed(21)
    x = torch.randn(r["batch"], 16, r["frames"], *r["frame_hw"], generator=g).to(torch.bfloat16)
    ctx = torch.randn(r["batch"], 512, r["text_dim"], generator=g).to(torch.bfloat16)
    t = torch.tensor([937.5, 250.0])
    return x, t, ctx


def bidirectional_cfg() -> O.OracleConfig:
    r = BIDIR
    return O.OracleConfig(dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                          text_dim=r["text_dim"])


def reference_bidirectional(ref):
    """The unmodified bidirectional WanModel (wan/modules/model.py:497-771) on a tiny config, two samples with
    different timesteps."""
    cfg = bidirectional_cfg()
    params = O.make_random_params(cfg, seed=9)
    torch.manual_seed(0)
    kw = cfg.reference_kwargs()
    kw.pop("local_attn_size"), kw.pop("sink_size")
    model = ref.WanModel(**kw)
    missing = model.load_state_dict(params, strict=False)
    assert not [k for k in missing.missing_keys if not k.startswith("pose_proj")], missing
    model = model.to(torch.bfloat16).eval()
    x, t, ctx = bidirectional_inputs()
    L = x.shape[2] * (x.shape[3] // 2) * (x.shape[4] // 2)
    with torch.no_grad():
        out = model(list(x), t=t, context=list(ctx), seq_len=L)
    return dict(flow=out.clone(), seq_len=L)


def _attention_with_k_lens(q, k, v, q_lens=None, k_lens=None, **kw):
    """flash_attn_varlen_func's semantics at the reference's call site (wan/modules/attention.py:88-150) restated with
    SDPA: sample b attends to its first k_lens[b] keys only.  The reference's own CPU fallback (`attention.attention`,
    :189-202) prints "Padding mask is disabled" and ignores k_lens, so a CPU run of the PADDED path needs this stand-in
    for the third-party kernel (flash_attn 2.8.3, not runnable without CUDA); everything else stays the unmodified
    reference."""
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    mask = None
    if k_lens is not None:
        mask = (torch.arange(Lk)[None, :] < k_lens.to(torch.long)[:, None])[:, None, None, :].expand(B, 1, Lq, Lk)
    out = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2).to(torch.bfloat16), k.transpose(1, 2).to(torch.bfloat16),
                                                           v.transpose(1, 2).to(torch.bfloat16), attn_mask=mask)
    return out.transpose(1, 2).contiguous()


def reference_bidirectional_padded(ref):
    """The same two samples pushed through the unmodified WanModel with seq_len LARGER than the samples (the model pads
    every sample with zero tokens to seq_len and passes k_lens = seq_lens to attention, model.py:684-693,150-156)."""
    import wan.modules.model as ref_model
    cfg = bidirectional_cfg()
    params = O.make_random_params(cfg, seed=9)
    kw = cfg.reference_kwargs()
    kw.pop("local_attn_size"), kw.pop("sink_size")
    model = ref.WanModel(**kw)
    model.load_state_dict(params, strict=False)
    model = model.to(torch.bfloat16).eval()
    x, t, ctx = bidirectional_inputs()
    L = x.shape[2] * (x.shape[3] // 2) * (x.shape[4] // 2)
    saved = ref_model.flash_attention
    ref_model.flash_attention = _attention_with_k_lens
    try:
        with torch.no_grad():
            out = model(list(x), t=t, context=[c[:300] for c in ctx], seq_len=L + 16)
    finally:
        ref_model.flash_attention = saved
    return dict(flow=out.clone(), seq_len=L + 16, context_rows=300)


# Training-time (cache-free) forward with block masks: tiny width, several mask shapes.
TRAIN = dict(dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_dim=512, frame_hw=(8, 12), batch=2)
TRAIN_CASES = {
    "causal_chunks_of_2": dict(frames=4, num_frame_per_block=2, local_attn_size=-1, independent_first_frame=False, tf=False),
    "causal_local_window": dict(frames=5, num_frame_per_block=1, local_attn_size=2, independent_first_frame=False, tf=False),
    "causal_lone_first_frame": dict(frames=5, num_frame_per_block=2, local_attn_size=-1, independent_first_frame=True, tf=False),
    "teacher_forcing": dict(frames=4, num_frame_per_block=2, local_attn_size=-1, independent_first_frame=False, tf=True),
}


def train_cfg(case: dict) -> O.OracleConfig:
    r = TRAIN
    return O.OracleConfig(dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                          text_dim=r["text_dim"], local_attn_size=case["local_attn_size"])


def train_inputs(case: dict):
    r = TRAIN
    g = torch.Generator().manual_seed(31)
    x = torch.randn(r["batch"], 16, case["frames"], *r["frame_hw"], generator=g).to(torch.bfloat16)
    clean = torch.randn(r["batch"], 16, case["frames"], *r["frame_hw"], generator=g).to(torch.bfloat16)
    ctx = torch.randn(r["batch"], 512, r["text_dim"], generator=g).to(torch.bfloat16)
    t = torch.randint(0, 1000, (r["batch"], case["frames"]), generator=g).float()
    aug = torch.randint(0, 100, (r["batch"], case["frames"]), generator=g).float()
    return x, clean, t, aug, ctx


def reference_train_forward(ref):
    """The unmodified CausalWanModel._forward_train (causal_model.py:895-1069) -- FlexAttention block masks, no cache.
    The reference compiles flex_attention with Inductor (:24-25), which cannot be lowered on a CPU: the module attribute
    is rebound to torch's uncompiled flex_attention (same op, eager arithmetic)."""
    from torch.nn.attention.flex_attention import flex_attention as eager_flex
    cm = ref.causal_model
    saved = cm.flex_attention
    cm.flex_attention = eager_flex
    out = {}
    try:
        for name, case in TRAIN_CASES.items():
            cfg = train_cfg(case)
            model = ref.CausalWanModel(**cfg.reference_kwargs())
            model.load_state_dict(O.make_random_params(cfg, seed=13), strict=False)
            model = model.to(torch.bfloat16).eval()
            model.num_frame_per_block = case["num_frame_per_block"]
            model.independent_first_frame = case["independent_first_frame"]
            x, clean, t, aug, ctx = train_inputs(case)
            kw = dict(clean_x=clean, aug_t=aug) if case["tf"] else {}
            with torch.no_grad():
                flow = model(x, t=t, context=list(ctx), seq_len=x.shape[2] * 24, **kw)
            out[name] = dict(flow=flow.clone(), case=case)
            print(name, flow.shape, float(flow.float().abs().mean()))
    finally:
        cm.flex_attention = saved
    return out


# 50-step sampler: CFG + UniPC on the same cached forward, separate pos / neg caches.
# `sampling_steps` is the reference pipeline's own attribute (hard-wired to 50 in its constructor,
# causal_diffusion_inference.py:66); the fixture lowers it on the instance so the CPU run stays short -- the solver
# still goes through warm-up (order 1), order-2 predictor + corrector and the lower-order final step.
DIFFUSION_CASES = {
    "cfg_unipc": dict(frames=2, num_frame_per_block=1, independent_first_frame=False, shift=5.0, sampling_steps=6,
                      guidance_scale=3.0),
}
NEGATIVE_PROMPT = "synthetic negative"
UNIPC_TRACE = dict(shape=(1, 2, 16, 8, 8), steps=50, shift=5.0)


class _TextEncoder2(torch.nn.Module):
    """prompt_embeds for the prompt, a second tensor for the negative prompt."""

    def __init__(self, pe, pe_neg):
        super().__init__()
        self.pe, self.pe_neg = pe, pe_neg

    def forward(self, text_prompts):
        return {"prompt_embeds": self.pe_neg if text_prompts[0] == NEGATIVE_PROMPT else self.pe}


def negative_embeds(batch: int = 1, text_dim: int = 4096):
    return torch.randn(batch, 512, text_dim, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16)


def diffusion_args(case: dict, **extra):
    return types.SimpleNamespace(num_train_timestep=1000, timestep_shift=case["shift"],
                                 guidance_scale=case["guidance_scale"], negative_prompt=NEGATIVE_PROMPT,
                                 num_frame_per_block=case["num_frame_per_block"],
                                 independent_first_frame=case["independent_first_frame"], model_kwargs={}, **extra)


def reference_diffusion(ref, case: dict, params, cfg: O.OracleConfig):
    w = ref_shim.make_reference_wrapper(ref, cfg.reference_kwargs(), case["shift"])
    w.model.load_state_dict(params, strict=False)
    pe, noise = synthetic_inputs(1, case["frames"])
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        pipe = ref.CausalDiffusionInferencePipeline(diffusion_args(case), "cpu", generator=w,
                                                    text_encoder=_TextEncoder2(pe, negative_embeds()),
                                                    vae=_IdentityVAE(), image_encoder=object())
        pipe.num_transformer_blocks = cfg.num_layers
        pipe.sampling_steps = case["sampling_steps"]
        with torch.no_grad():
            _, lat = pipe.inference(noise, ["synthetic"], None, None, None, return_latents=True)
    idx = tuple(int(c[0][k]) for c in (pipe.kv_cache_pos, pipe.kv_cache_neg) for k in ("global_end_index", "local_end_index"))
    return lat, idx


def unipc_trace_flow(sample: torch.Tensor, step: int) -> torch.Tensor:
    """Deterministic stand-in for the model: a seeded draw mixed with the current sample."""
    g = torch.Generator().manual_seed(100 + step)
    return (torch.randn(sample.shape, generator=g) + 0.5 * sample.float()).to(sample.dtype)


def reference_unipc_trace(ref, dtype):
    u = UNIPC_TRACE
    s = ref.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(u["steps"], device="cpu", shift=u["shift"])
    x = torch.randn(u["shape"], generator=torch.Generator().manual_seed(99)).to(dtype)
    xs = []
    for i, t in enumerate(s.timesteps):
        x = s.step(unipc_trace_flow(x, i), t, x, return_dict=False)[0]
        xs.append(x.clone())
    return dict(timesteps=s.timesteps.clone(), sigmas=s.sigmas.clone(), samples=torch.stack(xs))


# VAE decode: full-width decoder (96 base channels), tiny latent grid
VAE_CASE = dict(frames=3, hw=(4, 6), batch=1, seed=21)


def vae_latents(case=VAE_CASE, frames=None, seed_offset=0):
    f = frames or case["frames"]
    g = torch.Generator().manual_seed(31 + seed_offset)
    return torch.randn(case["batch"], f, 16, *case["hw"], generator=g).to(torch.bfloat16)


def reference_vae_decode(case=VAE_CASE):
    from . import vae_oracle as V
    rv = ref_shim.load_reference_vae()
    cfg = V.VaeConfig()
    params = V.make_random_vae_params(cfg, seed=case["seed"])
    model = rv.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                       temperal_downsample=[False, True, True], dropout=0.0)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "conv1.")) for k in missing), (missing, unexpected)
    model = model.to(torch.bfloat16).eval()
    lat = vae_latents(case)
    scale = [torch.tensor(V.LATENT_MEAN).to(torch.bfloat16), 1.0 / torch.tensor(V.LATENT_STD).to(torch.bfloat16)]
    with torch.no_grad():
        zs = lat.permute(0, 2, 1, 3, 4)
        full = model.decode(zs, scale)                                     # clears the cache before and after
        first = model.cached_decode(zs[:, :, :2], scale)                   # streaming: 2 frames, then 1 more
        second = model.cached_decode(zs[:, :, 2:], scale)
        model.clear_cache()
        # the same weights evaluated in fp32: the yardstick for "how far apart may two bf16 evaluations be"
        exact = model.float().decode(zs.float(), [torch.tensor(V.LATENT_MEAN), 1.0 / torch.tensor(V.LATENT_STD)])
    out = dict(case=case, pixels=full, streamed=torch.cat([first, second], dim=2), pixels_fp32=exact)
    # the demo's streaming decoder (demo_utils/vae_block3.py): explicit feature-cache list, all-zero tensors to start
    # with (demo_utils/constant.py ZERO_VAE_CACHE), so even the first frame goes through the temporal upsampling
    blk = ref_shim.load_reference_vae_block3()
    stream = blk.VAEDecoderWrapper()
    missing, unexpected = stream.load_state_dict(params, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    stream = stream.to(torch.bfloat16).eval()
    h, w = case["hw"]
    zero_cache = [torch.zeros(1, c, 2, h * s_, w * s_, dtype=torch.bfloat16) for c, s_ in vae_stream_cache_layout(cfg)]
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        a, cache = stream(lat[:, :2], *zero_cache)
        b, cache = stream(lat[:, 2:], *cache)
    out["block3_pixels"] = torch.cat([a, b], dim=1)
    out["block3_cache_sum"] = torch.stack([c.float().abs().sum() for c in cache])
    return out


def vae_stream_cache_layout(cfg):
    """(channels, spatial scale) of the 32 feature-cache slots in walk order (demo_utils/constant.py:5-38)."""
    plan, dims = cfg.stage_plan()
    slots = [(cfg.z_dim, 1)] + [(dims[0], 1)] * 4
    scale = 1
    for kind, _, cin, cout in plan:
        if kind == "res":
            slots += [(cin, scale), (cout, scale)]
        else:
            if kind == "up3d":
                slots.append((cin, scale))
            scale *= 2
    return slots + [(dims[-1], scale)]


# UMT5 text encoder: narrow / shallow instance of the same architecture, two prompts of
# different lengths padded to 24 tokens
T5_CASE = dict(vocab=200, dim=256, dim_attn=256, dim_ffn=640, num_heads=4, num_layers=3, num_buckets=32, seq=24,
               lengths=(24, 9), seed=41)


def t5_case_inputs(case=T5_CASE):
    g = torch.Generator().manual_seed(case["seed"] + 1)
    ids = torch.randint(1, case["vocab"], (len(case["lengths"]), case["seq"]), generator=g)
    mask = torch.zeros_like(ids)
    for i, n in enumerate(case["lengths"]):
        mask[i, :n] = 1
        ids[i, n:] = 0
    return ids, mask


def t5_case_cfg(case=T5_CASE):
    from . import t5_oracle as T
    return T.T5Config(**{k: case[k] for k in ("vocab", "dim", "dim_attn", "dim_ffn", "num_heads", "num_layers", "num_buckets")})


def reference_t5(case=T5_CASE):
    from . import t5_oracle as T
    t5 = ref_shim.load_reference_t5()
    cfg = t5_case_cfg(case)
    params = T.make_random_t5_params(cfg, seed=case["seed"])
    out = {"case": case}
    for name, dtype in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        m = t5.T5Encoder(vocab=cfg.vocab, dim=cfg.dim, dim_attn=cfg.dim_attn, dim_ffn=cfg.dim_ffn, num_heads=cfg.num_heads,
                         num_layers=cfg.num_layers, num_buckets=cfg.num_buckets, shared_pos=False, dropout=0.1)
        m.load_state_dict(params, strict=True)
        m = m.to(dtype).eval()
        ids, mask = t5_case_inputs(case)
        with torch.no_grad():
            ctx = m(ids, mask)
            for u, n in zip(ctx, mask.gt(0).sum(dim=1).long()):     # wan_wrapper.py:47-48
                u[n:] = 0.0
        out["context_" + name] = ctx
    return out


MASK_CASES = {
    "causal_6f_2blk": ("causal", dict(num_frames=6, frame_seqlen=200, num_frame_per_block=2, local_attn_size=-1)),
    "causal_6f_local2": ("causal", dict(num_frames=6, frame_seqlen=200, num_frame_per_block=1, local_attn_size=2)),
    "causal_21f_3blk_1560": ("causal", dict(num_frames=21, frame_seqlen=1560, num_frame_per_block=3, local_attn_size=-1)),
    "i2v_7f_3blk": ("i2v", dict(num_frames=7, frame_seqlen=200, num_frame_per_block=3, local_attn_size=-1)),
    "tf_4f_2blk": ("tf", dict(num_frames=4, frame_seqlen=200, num_frame_per_block=2)),
}


def reference_masks(ref):
    out = {}
    M = ref.CausalWanModel
    for name, (kind, kw) in MASK_CASES.items():
        with contextlib.redirect_stdout(io.StringIO()):
            if kind == "causal":
                bm = M._prepare_blockwise_causal_attn_mask("cpu", **kw)
            elif kind == "i2v":
                bm = M._prepare_blockwise_causal_attn_mask_i2v("cpu", **kw)
            else:
                bm = M._prepare_teacher_forcing_mask("cpu", **kw)
        out[name] = dict(kv_num_blocks=bm.kv_num_blocks[0, 0].to(torch.int32),
                         full_kv_num_blocks=bm.full_kv_num_blocks[0, 0].to(torch.int32),
                         kv_indices=bm.kv_indices[0, 0].to(torch.int16),
                         full_kv_indices=bm.full_kv_indices[0, 0].to(torch.int16),
                         sparsity=float(bm.sparsity()))
    return out


def main():
    ref = ref_shim.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    sched = {}
    for shift in (5.0, 8.0):
        s = ref.FlowMatchScheduler(shift=shift, sigma_min=0.0, extra_one_step=True)
        s.set_timesteps(1000, training=True)
        ts = torch.cat((s.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
        sched[shift] = dict(sigmas=s.sigmas.clone(), timesteps=s.timesteps.clone(),
                            warped=ts[1000 - torch.tensor([1000, 750, 500, 250])].clone())
        x0 = torch.randn(3, 16, 8, 8, generator=torch.Generator().manual_seed(7)).to(torch.bfloat16)
        nz = torch.randn(3, 16, 8, 8, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
        tt = sched[shift]["warped"][1:].clone()
        sched[shift]["add_noise_out"] = s.add_noise(x0, nz, tt)
    torch.save(sched, os.path.join(GOLDEN, "scheduler_tables.pt"))

    cfg = O.OracleConfig(**O.WAN_TINY)
    params = O.make_random_params(cfg, seed=0)
    roll = {}
    for name, case in ROLLOUT_CASES.items():
        lat, idx = reference_rollout(ref, case, params, cfg)
        roll[name] = dict(latents=lat, final_index=idx, case=case)
        print(name, lat.shape, idx, float(lat.float().std()))
    torch.save(roll, os.path.join(GOLDEN, "rollout_tiny.pt"))
    rolling = {}
    for name, case in ROLLING_ROLLOUT_CASES.items():
        lat, idx = reference_rollout(ref, case, params, rollout_cfg(case))
        rolling[name] = dict(latents=lat, final_index=idx, case=case)
        print(name, lat.shape, idx)
    torch.save(rolling, os.path.join(GOLDEN, "rollout_rolling.pt"))

    torch.save(reference_rolling(ref), os.path.join(GOLDEN, "model_rolling.pt"))
    torch.save(reference_masks(ref), os.path.join(GOLDEN, "block_masks.pt"))
    torch.save(reference_bidirectional(ref), os.path.join(GOLDEN, "bidirectional_tiny.pt"))
    torch.save(reference_bidirectional_padded(ref), os.path.join(GOLDEN, "bidirectional_padded.pt"))
    torch.save(reference_train_forward(ref), os.path.join(GOLDEN, "train_forward_tiny.pt"))

    diff = {"unipc_trace_bf16": reference_unipc_trace(ref, torch.bfloat16),
            "unipc_trace_fp32": reference_unipc_trace(ref, torch.float32)}
    for name, case in DIFFUSION_CASES.items():
        lat, idx = reference_diffusion(ref, case, params, cfg)
        diff[name] = dict(latents=lat, final_index=idx, case=case)
        print(name, lat.shape, idx, float(lat.float().std()))
    torch.save(diff, os.path.join(GOLDEN, "diffusion_tiny.pt"))
    torch.save(reference_vae_decode(), os.path.join(GOLDEN, "vae_decode_tiny.pt"))
    torch.save(reference_t5(), os.path.join(GOLDEN, "t5_tiny.pt"))
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
