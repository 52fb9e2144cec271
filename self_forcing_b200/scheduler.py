"""Flow-matching schedule of the few-step sampler.

Own implementation of the pieces of `FlowMatchScheduler` the rollout touches
(utils/scheduler.py:106-176 in the reference, constructed at utils/wan_wrapper.py:171-174 with
shift=timestep_shift, sigma_min=0, extra_one_step=True, set_timesteps(1000, training=True)):
the shifted sigma / timestep tables and `add_noise`, which runs as one fused sm_100a kernel
(`sfb_add_noise`) instead of ~5 elementwise launches + an argmin.
"""
from __future__ import annotations

from typing import Dict

import torch


class FlowMatchScheduler:
    def __init__(self, shift: float = 3.0, sigma_min: float = 0.0, sigma_max: float = 1.0,
                 num_train_timesteps: int = 1000, extra_one_step: bool = True, ops=None):
        self.shift, self.sigma_min, self.sigma_max = shift, sigma_min, sigma_max
        self.num_train_timesteps, self.extra_one_step = num_train_timesteps, extra_one_step
        self._ops = ops
        self._dev: Dict[str, tuple] = {}
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int = 1000, denoising_strength: float = 1.0, training: bool = False):
        """sigmas = shift*s / (1 + (shift-1)*s) over s = linspace(sigma_start, sigma_min, n(+1))[:n];
        timesteps = sigmas * num_train_timesteps   (scheduler.py:118-133)."""
        start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            s = torch.linspace(start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            s = torch.linspace(start, self.sigma_min, num_inference_steps)
        self.sigmas = self.shift * s / (1 + (self.shift - 1) * s)
        self.timesteps = self.sigmas * self.num_train_timesteps
        self._dev.clear()

    def tables_on(self, device) -> tuple:
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (self.timesteps.to(device=device, dtype=torch.float32).contiguous(),
                              self.sigmas.to(device=device, dtype=torch.float32).contiguous())
        return self._dev[key]

    @property
    def ops(self):
        if self._ops is None:
            from .ops import CudaOps
            self._ops = CudaOps()
        return self._ops

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
        """(1 - sigma) * x0 + sigma * noise with sigma of the nearest table timestep, fp32 math, result in
        noise.dtype (scheduler.py:159-176).  Shapes [N, C, H, W], [N, C, H, W], [N] (or [B, T])."""
        if timestep.ndim == 2:
            timestep = timestep.flatten(0, 1)
        if getattr(self.ops, "requires_bf16", True) and (noise.dtype != torch.bfloat16 or
                                                         original_samples.dtype != torch.bfloat16):
            raise TypeError("B200 add_noise runs on bfloat16 latents")
        ts, sg = self.tables_on(noise.device)
        out = torch.empty_like(noise, memory_format=torch.contiguous_format)
        self.ops.add_noise(original_samples.contiguous(), noise.contiguous(), timestep.contiguous(), ts, sg, out)
        return out
