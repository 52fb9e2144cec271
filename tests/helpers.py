"""Shared test helpers (synthetic inputs identical to oracle/make_golden.py)."""
from __future__ import annotations

import os
import types

import torch

from oracle import causal_wan_oracle as O
# the fixture definitions live next to the script that generated them; re-exported here for the test modules
from oracle.make_golden import (DIFFUSION_CASES, NEGATIVE_PROMPT, ROLLING_ROLLOUT_CASES, ROLLOUT_CASES,  # noqa: F401
                                SeededNoise, _IdentityVAE, _TextEncoder, _TextEncoder2, diffusion_args,  # noqa: F401
                                initial_latent_for, negative_embeds, patched_randn_like, synthetic_inputs)  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name: str):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_l2(a, b) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def pipeline_args(case: dict, **extra):
    extra = {k: v for k, v in extra.items() if k not in ("initial_frames", "local_attn_size", "sink_size")}
    return types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=case["num_frame_per_block"],
                                 independent_first_frame=case["independent_first_frame"], context_noise=0,
                                 model_kwargs={}, **extra)


def make_product_pipeline(case: dict, device, ops=None, num_layers=2, ffn_dim=512, dtype=torch.bfloat16, seed=0,
                          hw=(60, 104), **extra):
    """B200DiffusionWrapper + product CausalInferencePipeline for a tiny-depth, full-width model."""
    from self_forcing_b200.pipeline import CausalInferencePipeline
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper
    window = dict(local_attn_size=case.get("local_attn_size", -1), sink_size=case.get("sink_size", 0))
    cfg = O.OracleConfig(dim=1536, ffn_dim=ffn_dim, num_heads=12, num_layers=num_layers, **window)
    params = O.make_random_params(cfg, seed=seed, dtype=dtype)
    w = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B, ffn_dim=ffn_dim, num_layers=num_layers),
                             timestep_shift=case["shift"], device=device, ops=ops, dtype=dtype, **window)
    w.model.load_state_dict(params, strict=True)
    pe, noise = synthetic_inputs(1, case["frames"], *hw)
    pe, noise = pe.to(device=device, dtype=dtype), noise.to(device=device, dtype=dtype)
    pipe = CausalInferencePipeline(pipeline_args(case, **extra), device, generator=w, text_encoder=_TextEncoder(pe),
                                   vae=_IdentityVAE())
    return pipe, cfg, params, pe, noise


def make_product_diffusion_pipeline(case: dict, device, ops=None, num_layers=2, ffn_dim=512, dtype=torch.bfloat16, seed=0,
                                    scalar_rounding="bf16", hw=(60, 104)):
    """B200DiffusionWrapper + product CausalDiffusionInferencePipeline (CFG + UniPC) for a tiny-depth model.
    scalar_rounding "bf16" = how the CPU run of the reference treats the solver's 0-dim tensor scalars, i.e. what
    the golden vectors contain (self_forcing_b200/unipc.py)."""
    from self_forcing_b200.diffusion_pipeline import CausalDiffusionInferencePipeline
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper
    cfg = O.OracleConfig(dim=1536, ffn_dim=ffn_dim, num_heads=12, num_layers=num_layers)
    params = O.make_random_params(cfg, seed=seed, dtype=dtype)
    w = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B, ffn_dim=ffn_dim, num_layers=num_layers),
                             timestep_shift=case["shift"], device=device, ops=ops, dtype=dtype)
    w.model.load_state_dict(params, strict=True)
    pe, noise = synthetic_inputs(1, case["frames"], *hw)
    pe, noise = pe.to(device=device, dtype=dtype), noise.to(device=device, dtype=dtype)
    neg = negative_embeds().to(device=device, dtype=dtype)
    args = diffusion_args(case, sampling_steps=case["sampling_steps"], unipc_scalar_rounding=scalar_rounding)
    pipe = CausalDiffusionInferencePipeline(args, device, generator=w, text_encoder=_TextEncoder2(pe, neg),
                                            vae=_IdentityVAE())
    return pipe, cfg, params, pe, neg, noise
