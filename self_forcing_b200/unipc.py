"""Flow-matching UniPC multistep solver of the 50-step sampler -- host side.

Mirrors the interface of the reference's `FlowUniPCMultistepScheduler` (wan/utils/fm_solvers_unipc.py:20-800) as the
pipeline uses it (pipeline/causal_diffusion_inference.py:519-527, :423-428): `set_timesteps(n, device=, shift=)`,
`.timesteps` (int64), `.sigmas` (fp32, CPU), `step(model_output, timestep, sample, return_dict=False)[0]`.

Own design: the solver's scalar coefficients depend on the sigma table only, so the host computes them (fp32, the
same arithmetic as fm_solvers_unipc.py:404-456 / :549-603) and ONE kernel (`sfb_cfg_unipc_step`, csrc/sampler.cu)
does all tensor work of a step -- classifier-free guidance, flow -> x0, UniC corrector, UniP predictor -- with the
reference's bf16 rounding after every tensor op.  `step(..., model_output_uncond=, guidance_scale=)` is the fused
entry the pipeline calls; without those arguments it is the reference's plain `step`.

Supported configuration = what the reference pipeline constructs: flow_prediction, predict_x0, solver_order <= 2,
bh1 / bh2, lower_order_final, final sigma 0, no thresholding, no dynamic shifting.  Anything else raises.

`scalar_rounding`: the reference multiplies bf16 tensors by fp32 0-dim CPU tensors.  On a CUDA device torch keeps such
a scalar in fp32 ("fp32", the default here); on a CPU device torch first casts it to the tensor dtype ("bf16"), which
is what the CPU-generated golden vectors contain.  Division by a 0-dim tensor keeps fp32 in both cases.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch


def _bf16(v: float) -> float:
    return float(torch.tensor(v, dtype=torch.float32).to(torch.bfloat16))


class SchedulerOutput:
    def __init__(self, prev_sample: torch.Tensor):
        self.prev_sample = prev_sample


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, prediction_type: str = "flow_prediction",
                 shift: Optional[float] = 1.0, use_dynamic_shifting: bool = False, thresholding: bool = False,
                 predict_x0: bool = True, solver_type: str = "bh2", lower_order_final: bool = True,
                 disable_corrector: Sequence[int] = (), final_sigmas_type: str = "zero", ops=None,
                 scalar_rounding: str = "fp32"):
        if prediction_type != "flow_prediction" or not predict_x0 or thresholding or use_dynamic_shifting:
            raise NotImplementedError("B200 UniPC: only flow_prediction / predict_x0 without thresholding or dynamic "
                                      "shifting (the configuration of causal_diffusion_inference.py:521-524)")
        if solver_order not in (1, 2):
            raise NotImplementedError("B200 UniPC: solver_order must be 1 or 2")
        if solver_type in ("midpoint", "heun", "logrho"):   # fm_solvers_unipc.py:97-99
            solver_type = "bh2"
        if solver_type not in ("bh1", "bh2"):
            raise NotImplementedError(f"{solver_type} is not implemented")
        if final_sigmas_type != "zero":
            raise NotImplementedError("B200 UniPC: final_sigmas_type must be 'zero'")
        if scalar_rounding not in ("fp32", "bf16"):
            raise ValueError("scalar_rounding must be 'fp32' or 'bf16'")
        self.num_train_timesteps, self.solver_order, self.solver_type = num_train_timesteps, solver_order, solver_type
        self.shift, self.lower_order_final = shift, lower_order_final
        self.disable_corrector = list(disable_corrector)
        self.scalar_rounding = scalar_rounding
        self._ops = ops
        # training schedule before set_timesteps (fm_solvers_unipc.py:104-131)
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sig = torch.from_numpy(1.0 - alphas).to(torch.float32)
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigmas = sig
        self.timesteps = sig * num_train_timesteps
        self.sigma_min, self.sigma_max = float(sig[-1]), float(sig[0])
        self.num_inference_steps = None
        self._reset()

    def _reset(self) -> None:
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.solver_order   # x0 predictions, newest last
        self.lower_order_nums = 0
        self.last_sample: Optional[torch.Tensor] = None
        self.this_order = 0
        self._step_index: Optional[int] = None
        self._begin_index: Optional[int] = None

    @property
    def ops(self):
        if self._ops is None:
            from .ops import CudaOps
            self._ops = CudaOps()
        return self._ops

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0) -> None:
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, sigmas=None, mu=None,
                      shift: Optional[float] = None) -> None:
        """fm_solvers_unipc.py:160-228: float64 linspace over [sigma_max, sigma_min), shifted; timesteps truncated to
        int64; sigmas cast to fp32 with a final 0 appended (kept on the host)."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        else:
            sigmas = np.asarray(sigmas, dtype=np.float64)
        if shift is None:
            shift = self.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = torch.from_numpy(sigmas * self.num_train_timesteps).to(device=device, dtype=torch.int64)
        self._timesteps_host = [int(v) for v in (sigmas * self.num_train_timesteps).astype(np.int64)]
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0.0]]).astype(np.float32))
        self.num_inference_steps = len(sigmas)
        self._reset()

    def scale_model_input(self, sample: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return sample

    def __len__(self) -> int:
        return self.num_train_timesteps

    # ------------------------------------------------------------------------------------
    def _lambda(self, i: int) -> torch.Tensor:
        s = self.sigmas[i]
        return torch.log(1 - s) - torch.log(s)

    def _coefficients(self, i_from: int, i_to: int, i_older: int, order: int, corrector: bool):
        """Scalars of one B(h) update from sigma[i_from] to sigma[i_to] (fp32 0-dim arithmetic like the reference):
        (sigma_t/sigma_s, alpha_t*h_phi_1, alpha_t*B_h, 1/rk, rho_0, rho_last)."""
        sig_t, sig_s = self.sigmas[i_to], self.sigmas[i_from]
        alpha_t = 1 - sig_t
        h = self._lambda(i_to) - self._lambda(i_from)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        B_h = hh if self.solver_type == "bh1" else torch.expm1(hh)
        inv_rk, rho0, rho_last = 0.0, 0.0, 0.5
        if order == 2:
            rk = (self._lambda(i_older) - self._lambda(i_from)) / h
            inv_rk = float(torch.tensor(1.0) / rk)
            if corrector:
                # R rho = b with R = [[1, 1], [rk, 1]], b_i = h*phi_{i+1}(h) * i! / B(h)  (:571-603)
                p2 = h_phi_1 / hh - 1
                p3 = p2 / hh - 1 / 2
                R = torch.stack([torch.ones(2), torch.stack([rk, torch.tensor(1.0)])])
                rho = torch.linalg.solve(R, torch.stack([p2 / B_h, p3 * 2 / B_h])).to(torch.bfloat16)
                rho0, rho_last = float(rho[0]), float(rho[1])
        coef = [float(sig_t / sig_s), float(alpha_t * h_phi_1), float(alpha_t * B_h)]
        if self.scalar_rounding == "bf16":
            coef = [_bf16(c) for c in coef]
        return coef + [inv_rk, rho0, rho_last]

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True,
             generator=None, model_output_uncond: Optional[torch.Tensor] = None, guidance_scale: float = 1.0):
        """fm_solvers_unipc.py:655-739.  With `model_output_uncond` the guidance
        `uncond + guidance_scale * (cond - uncond)` (causal_diffusion_inference.py:420-421) is fused into the step."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if getattr(self.ops, "requires_bf16", True) and (sample.dtype != torch.bfloat16 or model_output.dtype != torch.bfloat16):
            raise TypeError("B200 UniPC step runs on bfloat16 latents")
        if self._step_index is None:
            if self._begin_index is not None:
                self._step_index = self._begin_index
            else:                                              # index_for_timestep (:628-641)
                hits = [i for i, v in enumerate(self._timesteps_host) if v == int(timestep)]
                if not hits:
                    raise ValueError(f"timestep {int(timestep)} is not in the schedule")
                self._step_index = hits[1 if len(hits) > 1 else 0]
        i = self._step_index
        n_steps = self.num_inference_steps
        corr_order = 0
        if i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None:
            corr_order = self.this_order
        if self.lower_order_final:
            pred_order = min(self.solver_order, n_steps - i)
        else:
            pred_order = self.solver_order
        pred_order = min(pred_order, self.lower_order_nums + 1)
        assert pred_order > 0

        sigma = float(self.sigmas[i])
        coef = [float(guidance_scale), _bf16(sigma) if self.scalar_rounding == "bf16" else sigma]
        coef += self._coefficients(i - 1, i, i - 2, corr_order, True) if corr_order else [0.0] * 6
        coef += self._coefficients(i, i + 1, i - 1, pred_order, False)[:4]

        m0 = self.model_outputs[-1]
        m1 = self.model_outputs[-2] if self.solver_order == 2 else None
        sample = sample.contiguous()
        # ring of history buffers: the oldest prediction's storage receives the new one (the kernel reads every element
        # before writing it), the previous corrected sample's storage receives the new corrected sample
        def fresh():
            return torch.empty_like(sample, memory_format=torch.contiguous_format)

        oldest = m1 if self.solver_order == 2 else m0
        m_out = oldest if oldest is not None else fresh()
        sample_out = self.last_sample if self.last_sample is not None else fresh()
        prev = fresh()
        self.ops.cfg_unipc_step(model_output.contiguous(),
                                None if model_output_uncond is None else model_output_uncond.contiguous(),
                                sample, self.last_sample if corr_order else None,
                                m0 if (corr_order or pred_order == 2) else None, m1 if corr_order == 2 else None,
                                m_out, sample_out, prev, coef, corr_order, pred_order)
        if self.solver_order == 2:
            self.model_outputs = [m0, m_out]
        else:
            self.model_outputs = [m_out]
        self.last_sample = sample_out
        self.this_order = pred_order
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return (prev,) if not return_dict else SchedulerOutput(prev)
