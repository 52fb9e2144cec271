"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the 50-step sampler that sits on the same model forward
(SURVEY.md section 8f rank 2): classifier-free guidance + the flow-matching UniPC multistep solver, and the
chunk-wise driver with separate positive / negative KV caches.

Follows (cited, not copied):
  * `FlowUniPCMultistepScheduler` -- wan/utils/fm_solvers_unipc.py: tables `:160-228`, flow -> x0 `:279-348`,
    predictor UniP-B(h) `:350-484`, corrector UniC-B(h) `:486-626`, `step` bookkeeping `:655-739`; as configured by
    the pipeline (`pipeline/causal_diffusion_inference.py:519-527`): flow_prediction, predict_x0, solver_order 2,
    bh2, lower_order_final, final sigma 0, shift applied in `set_timesteps`.
  * `CausalDiffusionInferencePipeline.inference` -- pipeline/causal_diffusion_inference.py:174-457 (t2v: no image,
    no pose input).

The reference evaluates every tensor expression op by op in the latents' dtype (bf16 on the GPU path) with the
scalar coefficients held as fp32 0-dim CPU tensors; the restatement keeps that order so that it matches the
reference bit for bit on the same host.  Only tests/, smoke() and bench.py's CPU legs may import this module.

Pinned by tests/golden/diffusion_tiny.pt (oracle/make_golden.py runs the unmodified reference classes).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import causal_wan_oracle as O


class OracleUniPC:
    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2):
        self.n_train, self.order_max = num_train_timesteps, solver_order
        # range of the unshifted training schedule (fm_solvers_unipc.py:111-131 with shift 1)
        self.sigma_max = float(np.float32(1.0 - 1.0 / num_train_timesteps))
        self.sigma_min = 0.0
        self.sigmas: Optional[Tensor] = None
        self.timesteps: Optional[Tensor] = None

    def set_timesteps(self, num_steps: int, shift: float) -> None:
        """`:182-211`: linspace in float64 numpy, shift, timesteps truncated to int64, sigmas cast to fp32 + final 0."""
        s = np.linspace(self.sigma_max, self.sigma_min, num_steps + 1).copy()[:-1]
        s = shift * s / (1 + (shift - 1) * s)
        self.timesteps = torch.from_numpy(s * self.n_train).to(torch.int64)
        self.sigmas = torch.from_numpy(np.concatenate([s, [0]]).astype(np.float32))
        self.hist: List[Optional[Tensor]] = [None] * self.order_max    # x0 predictions, newest last
        self.warm = 0            # lower_order_nums
        self.last_sample: Optional[Tensor] = None
        self.index: Optional[int] = None
        self.cur_order = 0       # order chosen by the previous step (used by the corrector)

    # ---- scalar pieces, all fp32 0-dim tensors like the reference ----
    def _lam(self, i: int) -> Tensor:
        sg = self.sigmas[i]
        return torch.log(1 - sg) - torch.log(sg)

    def _update(self, x: Tensor, m0: Tensor, older: Optional[Tensor], i_from: int, i_to: int, i_older: int,
                order: int, new: Optional[Tensor]) -> Tensor:
        """Shared shape of UniP (`new is None`, `:404-484`) and UniC (`new` = x0 prediction at the target point,
        `:549-626`): exponential-integrator step from sigma[i_from] to sigma[i_to] with a first-difference term."""
        sig_t, sig_s = self.sigmas[i_to], self.sigmas[i_from]
        alpha_t = 1 - sig_t
        h = self._lam(i_to) - self._lam(i_from)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        B_h = torch.expm1(hh)
        rks, diffs = [], []
        if order == 2:
            rk = (self._lam(i_older) - self._lam(i_from)) / h
            rks.append(rk)
            diffs.append((older - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks, device=x.device)
        # b_i = h*phi_{i+1}(h) * i! / B(h), R_ij = rk_j^(i-1)
        rows, b = [], []
        h_phi_k, fact = h_phi_1 / hh - 1, 1
        for i in range(1, order + 1):
            rows.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        R, b = torch.stack(rows), torch.tensor(b, device=x.device)
        base = sig_t / sig_s * x - alpha_t * h_phi_1 * m0
        if new is None:                                            # predictor
            if not diffs:
                return (base - alpha_t * B_h * 0).to(x.dtype)
            rho = torch.tensor([0.5], dtype=x.dtype, device=x.device)   # order 2 shortcut `:461-462`
            res = torch.einsum("k,bkc...->bc...", rho, torch.stack(diffs, dim=1))
            return (base - alpha_t * B_h * res).to(x.dtype)
        rho = torch.tensor([0.5], dtype=x.dtype, device=x.device) if order == 1 else torch.linalg.solve(R, b).to(x.dtype)
        res = torch.einsum("k,bkc...->bc...", rho[:-1], torch.stack(diffs, dim=1)) if diffs else 0
        return (base - alpha_t * B_h * (res + rho[-1] * (new - m0))).to(x.dtype)

    def step(self, flow: Tensor, timestep, sample: Tensor) -> Tensor:
        """`:655-739`."""
        if self.index is None:
            hits = (self.timesteps == timestep).nonzero()
            self.index = int(hits[1 if len(hits) > 1 else 0])
        i = self.index
        x0 = sample - self.sigmas[i] * flow                        # `:320-323`
        older = self.hist[-2] if self.order_max > 1 else None
        if i > 0 and self.last_sample is not None:
            sample = self._update(self.last_sample, self.hist[-1], older, i - 1, i, i - 2, self.cur_order, x0)
        self.hist = self.hist[1:] + [x0]
        self.cur_order = min(self.order_max, len(self.timesteps) - i, self.warm + 1)
        self.last_sample = sample
        nxt = self._update(sample, x0, self.hist[-2] if self.order_max > 1 else None, i, i + 1, i - 1, self.cur_order, None)
        self.warm = min(self.warm + 1, self.order_max)
        self.index += 1
        return nxt


@dataclass
class DiffusionTrace:
    latents: Tensor
    index_trace: List[Tuple[int, int, int, int]] = field(default_factory=list)   # (pos g, pos l, neg g, neg l)


def diffusion_rollout(wrapper: O.OracleWrapper, noise: Tensor, cond: Tensor, uncond: Tensor, guidance_scale: float,
                      num_frame_per_block: int, sampling_steps: int = 50, shift: float = 5.0,
                      independent_first_frame: bool = False, initial_latent: Optional[Tensor] = None,
                      cache_tokens=None) -> DiffusionTrace:
    """`CausalDiffusionInferencePipeline.inference` without T5 / VAE / image / pose inputs."""
    cfg = wrapper.cfg
    B, nfr, _, Hh, Ww = noise.shape
    ft = (Hh // cfg.patch_size[1]) * (Ww // cfg.patch_size[2])
    n_in = 0 if initial_latent is None else initial_latent.shape[1]
    if independent_first_frame and initial_latent is None:
        chunks = [1] + [num_frame_per_block] * ((nfr - 1) // num_frame_per_block)
    else:
        assert nfr % num_frame_per_block == 0
        chunks = [num_frame_per_block] * (nfr // num_frame_per_block)
    kv = [O.new_kv_cache(cfg, B, ft, noise.dtype, noise.device, cache_tokens) for _ in range(2)]
    ca = [O.new_crossattn_cache(cfg, B, noise.dtype, noise.device) for _ in range(2)]
    ctx = [cond, uncond]
    out = torch.zeros(B, n_in + nfr, *noise.shape[2:], dtype=noise.dtype, device=noise.device)
    trace = DiffusionTrace(latents=out)

    def both(x, t, start):
        flows = [wrapper(x, ctx[s], t, kv[s], ca[s], start * ft)[0] for s in range(2)]
        trace.index_trace.append(tuple(int(kv[s][0][key]) for s in range(2) for key in ("global_end_index", "local_end_index")))
        return flows

    start = 0
    if initial_latent is not None:                                 # `:256-312`
        groups = []
        if independent_first_frame:
            groups.append(1)
            groups += [num_frame_per_block] * ((n_in - 1) // num_frame_per_block)
        else:
            groups = [num_frame_per_block] * (n_in // num_frame_per_block)
        for n in groups:
            ref = initial_latent[:, start:start + n]
            out[:, start:start + n] = ref
            both(ref, torch.zeros([B, 1], dtype=torch.int64), start)
            start += n
    for n in chunks:                                               # `:371-451`
        x = noise[:, start - n_in:start - n_in + n]
        solver = OracleUniPC()
        solver.set_timesteps(sampling_steps, shift)
        for t in solver.timesteps:
            timestep = t * torch.ones([B, n], dtype=torch.float32)
            fc, fu = both(x, timestep, start)
            x = solver.step(fu + guidance_scale * (fc - fu), t, x)
        out[:, start:start + n] = x
        both(x, timestep * 0, start)                               # clean-context refresh `:431-448`
        start += n
    return trace
