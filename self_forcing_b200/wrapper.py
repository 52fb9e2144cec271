"""B200DiffusionWrapper -- drop-in for `WanDiffusionWrapper` (utils/wan_wrapper.py:120-371) on the
causal KV-cached path.

Keeps the surface `pipeline/causal_inference.py` touches:
`.model` (with `.local_attn_size`, `.num_frame_per_block`), `.scheduler` / `.get_scheduler()`,
`.seq_len`, and `__call__(noisy_image_or_video=[B,F,16,H,W], conditional_dict={"prompt_embeds"},
timestep=[B,F], kv_cache=, crossattn_cache=, current_start=) -> (flow_pred, pred_x0)`.
The flow -> x0 conversion (wan_wrapper.py:204-228, float64) is fused into the model's last kernel.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional

import torch
from torch import nn

from .model import B200CausalWanModel, B200WanModel
from .scheduler import FlowMatchScheduler

WAN_T2V_1_3B = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=1536, ffn_dim=8960,
                    freq_dim=256, text_dim=4096, out_dim=16, num_heads=12, num_layers=30, qk_norm=True,
                    cross_attn_norm=True, eps=1e-6)
WAN_T2V_14B = dict(WAN_T2V_1_3B, dim=5120, ffn_dim=13824, num_heads=40, num_layers=40)


class B200DiffusionWrapper(nn.Module):
    def __init__(self, model_name: str = "Wan2.1-T2V-1.3B", model_path: Optional[str] = None,
                 timestep_shift: float = 8.0, is_causal: bool = True, local_attn_size: int = -1, sink_size: int = 0,
                 model_config: Optional[dict] = None, device=None, init_seed: Optional[int] = None, ops=None,
                 dtype=torch.bfloat16, **unsupported):
        """model_config given (or init_seed set) -> random-init weights of that architecture on `device`;
        otherwise `config.json` + safetensors / .pth weights are read from model_path (HF layout used by
        CausalWanModel.from_pretrained at wan_wrapper.py:139-145)."""
        super().__init__()
        lora = {k: v for k, v in unsupported.items() if k.startswith("lora") and v}
        if lora.get("lora_rank"):
            raise NotImplementedError("merge LoRA weights offline (scripts/merge_lora.py) before loading")
        cfg = dict(model_config) if model_config is not None else None
        state = None
        if cfg is None:
            path = model_path or f"wan_models/{model_name}/"
            cfg, state = _read_checkpoint_dir(path)
        if is_causal:
            cfg.update(local_attn_size=local_attn_size, sink_size=sink_size)
            self.model = B200CausalWanModel(**cfg, ops=ops)
        else:   # bidirectional teacher (wan_wrapper.py:146 `WanModel.from_pretrained`); one timestep per sample
            self.model = B200WanModel(**cfg, ops=ops)
        if device is not None:
            self.model.to(device)
        self.model.to(dtype)
        if state is not None:
            load_checkpoint_strict(self.model, state)
        elif init_seed is not None:
            self.model.init_weights(init_seed)
        self.model.eval()
        self.uniform_timestep = not is_causal      # wan_wrapper.py:170
        self.scheduler = FlowMatchScheduler(shift=timestep_shift, sigma_min=0.0, extra_one_step=True, ops=ops)
        self.scheduler.set_timesteps(1000, training=True)
        self.model.set_sampler_tables(self.scheduler.timesteps, self.scheduler.sigmas)
        self.seq_len = 32760   # [1, 21, 16, 60, 104]

    def get_scheduler(self) -> FlowMatchScheduler:
        return self.scheduler

    def forward(self, noisy_image_or_video: torch.Tensor, conditional_dict: dict, timestep: torch.Tensor,
                kv_cache: Optional[List[dict]] = None, crossattn_cache: Optional[List[dict]] = None,
                current_start: Optional[int] = None, classify_mode: bool = False,
                concat_time_embeddings: bool = False, clean_x=None, aug_t=None, cache_start: Optional[int] = None,
                add_condition=None, clip_feature=None, y=None, refresh_only: bool = False):
        if classify_mode:
            raise NotImplementedError("the GAN classify branch is a training-only head (out of scope)")
        if kv_cache is None and not isinstance(self.model, B200WanModel):
            # cache-free forward of the causal model (wan_wrapper.py:301-337: training-time rollouts, teacher forcing with
            # clean_x / aug_t): block-masked attention over the whole video, forward only
            flow, x0 = self.model(noisy_image_or_video.permute(0, 2, 1, 3, 4), t=timestep,
                                  context=conditional_dict["prompt_embeds"], seq_len=self.seq_len,
                                  clean_x=None if clean_x is None else clean_x.permute(0, 2, 1, 3, 4), aug_t=aug_t,
                                  return_x0=True)
            return flow.permute(0, 2, 1, 3, 4), x0
        if clean_x is not None:
            raise NotImplementedError("teacher forcing applies to the causal model")
        if kv_cache is None:
            # bidirectional forward (wan_wrapper.py:329-337): [B, F] timesteps must be uniform per sample (:282-283)
            t0 = timestep[:, 0]
            if not bool((timestep == t0[:, None]).all()):
                raise ValueError("the bidirectional model takes one timestep per sample")
            flow, x0 = self.model(noisy_image_or_video.permute(0, 2, 1, 3, 4), t=t0,
                                  context=conditional_dict["prompt_embeds"], seq_len=self.seq_len, return_x0=True)
            return flow.permute(0, 2, 1, 3, 4), x0
        cond = conditional_dict
        if add_condition is None:
            add_condition = cond.get("add_condition")
        if clip_feature is None:
            clip_feature = cond.get("clip_feature")
        if y is None:
            y = cond.get("y")
        out = self.model(noisy_image_or_video.permute(0, 2, 1, 3, 4), t=timestep, context=cond["prompt_embeds"],
                         seq_len=self.seq_len, kv_cache=kv_cache, crossattn_cache=crossattn_cache,
                         current_start=current_start, cache_start=cache_start, add_condition=add_condition,
                         clip_fea=clip_feature, y=y, return_x0=not refresh_only, skip_output=refresh_only)
        if refresh_only:
            return None, None
        flow, x0 = out
        return flow.permute(0, 2, 1, 3, 4), x0


def _read_checkpoint_dir(path: str):
    cfg_file = os.path.join(path, "config.json")
    if not os.path.isfile(cfg_file):
        raise FileNotFoundError(f"{cfg_file} not found (pass model_config=... for random-init weights)")
    raw = json.load(open(cfg_file))
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers", "qk_norm", "cross_attn_norm", "eps")
    cfg = {k: raw[k] for k in keys if k in raw}
    if "patch_size" in cfg:
        cfg["patch_size"] = tuple(cfg["patch_size"])
    state = {}
    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if files:
        from safetensors.torch import load_file
        for f in files:
            state.update(load_file(os.path.join(path, f)))
    else:
        # a Wan model directory also holds the T5 / VAE / CLIP checkpoints: pick the DiT weights by name, never by
        # listdir order
        pth = sorted(f for f in os.listdir(path) if f.endswith((".pth", ".pt", ".bin")))
        dit = [f for f in pth if f.startswith("diffusion_pytorch_model")]
        if not dit:
            raise FileNotFoundError(f"no diffusion_pytorch_model*.safetensors/.pth/.pt/.bin under {path} "
                                    f"(found {pth or 'nothing'})")
        for f in dit:
            state.update(torch.load(os.path.join(path, f), map_location="cpu"))
    return cfg, state


def unwrap_checkpoint(state: dict) -> dict:
    """Self-Forcing training checkpoints are `{generator | generator_ema: {"model.<key>": tensor}}`
    (trainer/distillation.py:203-228, loaded at inference.py:69-71); HF checkpoints are flat.  Returns the flat
    `<key>: tensor` dict of the DiT, with the fork's `pose_proj.*` dropped."""
    for outer in ("generator_ema", "generator"):
        if outer in state and isinstance(state[outer], dict):
            state = state[outer]
            break
    if state and all(k.startswith("model.") for k in state):
        state = {k[len("model."):]: v for k, v in state.items()}
    return {k: v for k, v in state.items() if not k.startswith("pose_proj.")}


def load_checkpoint_strict(model: nn.Module, state: dict) -> None:
    """The model's parameters are created with torch.empty, so a checkpoint whose keys do not match would leave it
    running on uninitialised memory: every parameter must be found, unknown keys are an error too."""
    flat = unwrap_checkpoint(state)
    want = set(model.state_dict().keys())
    missing, unexpected = sorted(want - set(flat)), sorted(set(flat) - want)
    if missing or unexpected:
        raise KeyError(f"checkpoint does not match the model: {len(missing)} missing (e.g. {missing[:3]}), "
                       f"{len(unexpected)} unexpected (e.g. {unexpected[:3]})")
    model.load_state_dict(flat, strict=True)
