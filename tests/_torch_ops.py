"""TEST DOUBLE (CPU): the op interface of self_forcing_b200.ops.CudaOps implemented with plain
PyTorch in the reference's dtype/rounding order.  It exists so that the *host orchestration*
(buffer wiring, cache slots, modulation indexing, drop-in surface) can be checked against the
oracle on a machine without a GPU.  It is never imported by the product."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import causal_wan_oracle as O

EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_GATE_RES, EPI_F32 = 0, 1, 2, 3, 4


class TorchOps:
    requires_bf16 = False

    def __init__(self):
        self.launches = 0
        self.log = []

    @staticmethod
    def _stats_records(y):
        """(mean, M2) per 128-column chunk of every row, like the GEMM epilogue writes them."""
        c = y.float().reshape(y.shape[0], -1, 128)
        mean = c.mean(dim=2)
        return torch.stack([mean, ((c - mean[..., None]) ** 2).sum(dim=2)], dim=2)

    @staticmethod
    def _stats_mean_var(rec):
        mean = rec[..., 0].mean(dim=1)
        m2 = (rec[..., 1] + 128.0 * (rec[..., 0] - mean[:, None]) ** 2).sum(dim=1)
        return mean, m2 / (128.0 * rec.shape[1])

    def gemm(self, x, w, bias, out, *, epilogue=EPI_BIAS, residual=None, gate=None, gate_stride=0,
             rows_per_gate=1, gate_row_offset=0, outs=None, seg_cols=0, block_n=0, stats_out=None, ln_stats=None,
             ln_sc=None, ln_eps=0.0):
        self.launches += 1
        self.log.append("gemm")
        if ln_stats is not None:      # LayerNorm of x folded into the epilogue (include/sfb200.h: sfb_gemm_bf16_stats)
            assert epilogue == EPI_BIAS and bias is None
            mean, var = self._stats_mean_var(ln_stats)
            acc = x.float() @ w.float().t()
            y = (torch.rsqrt(var + ln_eps)[:, None] * (acc - mean[:, None] * ln_sc[:, 0][None]) + ln_sc[:, 1][None]).to(out.dtype)
            out.copy_(y)
            if stats_out is not None:
                stats_out.copy_(self._stats_records(y))
            return
        if epilogue == EPI_F32:
            out.copy_(F.linear(x.float(), w.float(), None if bias is None else bias.float()))
            return
        y = F.linear(x, w, bias)
        if epilogue == EPI_GELU:
            y = F.gelu(y, approximate="tanh")
        elif epilogue == EPI_RESIDUAL:
            y = residual + y
        elif epilogue == EPI_GATE_RES:
            M = x.shape[0]
            g = gate[(torch.arange(M) + gate_row_offset) // rows_per_gate]   # gate is a [groups, C] strided view
            y = residual + y * g
        if stats_out is not None:
            stats_out.copy_(self._stats_records(y))
        if outs is not None:
            for i, o in enumerate(outs):
                o.copy_(y[:, i * seg_cols:(i + 1) * seg_cols])
        else:
            out.copy_(y)

    def attention(self, q, k, v, out, scale, q_stats=None, q_eps=0.0):
        self.launches += 1
        self.log.append("attention")
        if q_stats is not None:       # RMSNorm row factor of q applied to the scores (sfb_attention_fwd_qnorm)
            self.log.append("attention_qnorm")
            B, Lq, H, D = q.shape
            ms = (q_stats[..., 1] + 128.0 * q_stats[..., 0] ** 2).sum(dim=1) / (128.0 * q_stats.shape[1])
            rstd = torch.rsqrt(ms + q_eps).reshape(B, 1, Lq, 1)
            s_ = torch.einsum("blhd,bshd->bhls", q.float(), k.float()) * scale * rstd
            out.copy_(torch.einsum("bhls,bshd->blhd", torch.softmax(s_, dim=-1).to(v.dtype).float(), v.float()).to(out.dtype))
            return
        out.copy_(O.dense_attention(q, k, v))

    def modulation_table(self, mod, e, out, e_row_stride, e_group_stride):
        self.launches += 1
        NL, G, C = mod.shape
        R = out.shape[1]
        ef = e.reshape(-1)
        for r in range(R):
            for g in range(G):
                base = r * e_row_stride + g * e_group_stride
                out[:, r, g] = mod[:, g] + ef[base:base + C]

    def ln_modulate(self, x, y, shift, scale, mod_stride, rows_per_mod, eps, row_offset=0, stats=None):
        self.launches += 1
        if stats is not None:     # the records must describe the rows handed in (sfb_ln_modulate_stats)
            self.log.append("ln_modulate_stats")
            mean, var = self._stats_mean_var(stats)
            xf = x.float()
            assert torch.allclose(mean, xf.mean(dim=1), rtol=1e-4, atol=1e-4) and torch.allclose(var, xf.var(dim=1, unbiased=False), rtol=1e-3, atol=1e-4)
        n = F.layer_norm(x, (x.shape[1],), None, None, eps)
        idx = (torch.arange(x.shape[0]) + row_offset) // rows_per_mod
        y.copy_(n * (1 + scale[idx]) + shift[idx])

    def ln_affine(self, x, y, weight, bias, eps):
        self.launches += 1
        y.copy_(F.layer_norm(x, (x.shape[1],), weight, bias, eps))

    def rmsnorm(self, x, y, weight, eps):
        self.launches += 1
        y.copy_(O.rms_norm(x, weight, eps))

    def qk_norm_rope(self, q_in, k_in, v_in, wq, wk, eps, cos_tab, sin_tab, B, L, head_dim, grid, start_frame,
                     q_out, k_out, v_out, start_frame_dev=None, stats=None, q_chunk0=0, k_chunk0=0):
        self.launches += 1
        if stats is not None:     # the records must describe the rows handed in (sfb_qk_norm_rope_stats)
            self.log.append("qk_norm_rope_stats")
            nc = q_in.shape[1] // 128
            for x, c0 in ((q_in, q_chunk0), (k_in, k_chunk0)):
                rec = stats[:, c0:c0 + nc]
                ss = (rec[..., 1] + 128.0 * rec[..., 0] ** 2).sum(dim=1)
                assert torch.allclose(ss, x.float().pow(2).sum(dim=1), rtol=1e-3, atol=1e-3)
        if start_frame_dev is not None:
            start_frame = int(start_frame_dev)
        C = q_in.shape[1]
        H = C // head_dim
        ang = O.rope_angle_table(head_dim)
        q = O.rms_norm(q_in, wq, eps).view(B, L, H, head_dim)
        k = O.rms_norm(k_in, wk, eps).view(B, L, H, head_dim)
        q_out.copy_(O.rope_rotate(q, grid, ang, start_frame).reshape(B, L, C))
        k_out.copy_(O.rope_rotate(k, grid, ang, start_frame))
        if v_in is not None:
            v_out.copy_(v_in.view(B, L, H, head_dim))

    # ---- Ulysses head-parallel path: the CUDA ops store into peer memory; this double exchanges with
    # torch.distributed collectives (gloo) so the host orchestration can be tested with world_size > 1 on CPU
    def qk_norm_rope_sp(self, q_in, k_in, v_in, wq, wk, eps, cos_tab, sin_tab, head_dim, grid, start_frame,
                        token_offset, sp, q_buf, k_slot_local, v_slot_local, k_slot_ptrs, v_slot_ptrs):
        import torch.distributed as dist
        self.launches += 1
        rows, C = q_in.shape
        H = C // head_dim
        Hg = H // sp.world
        ang = O.rope_angle_table(head_dim)
        L = grid[0] * grid[1] * grid[2]

        def rot(x):   # rotate as if the rows sat at their chunk-global positions
            full = x.new_zeros(1, L, H, head_dim)
            full[0, token_offset:token_offset + rows] = x.view(rows, H, head_dim)
            return O.rope_rotate(full, grid, ang, start_frame)[0, token_offset:token_offset + rows]

        q = rot(O.rms_norm(q_in, wq, eps))
        k = rot(O.rms_norm(k_in, wk, eps))
        v = v_in.view(rows, H, head_dim)
        for name, t, dst in (("q", q, q_buf.local.view(L, Hg, head_dim)), ("k", k, k_slot_local), ("v", v, v_slot_local)):
            gathered = [torch.empty_like(t) for _ in range(sp.world)]
            dist.all_gather(gathered, t.contiguous(), group=sp.group)
            full = torch.cat(gathered, dim=0)                      # [L, H, D] in token order
            dst.copy_(full[:, sp.rank * Hg:(sp.rank + 1) * Hg].reshape(dst.shape))

    def attention_sp(self, q, k, v, scale, sp, out_buf, rows_per_rank):
        import torch.distributed as dist
        self.launches += 1
        o = O.dense_attention(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0))[0]      # [L, Hg, D]
        gathered = [torch.empty_like(o) for _ in range(sp.world)]
        dist.all_gather(gathered, o.contiguous(), group=sp.group)
        full = torch.cat(gathered, dim=1)                          # [L, H, D]
        lo = sp.rank * rows_per_rank
        out_buf.local.copy_(full[lo:lo + rows_per_rank].reshape(rows_per_rank, -1))

    def peer_barrier(self, sp):
        import torch.distributed as dist
        dist.barrier(group=sp.group)

    def kv_roll(self, tensors, table, dst_row, src_row, n_rows):
        for t in tensors:
            t[:, dst_row:dst_row + n_rows] = t[:, src_row:src_row + n_rows].clone()

    def patchify(self, x, out):
        self.launches += 1
        B, Cin, F_, H, W = x.shape
        t = x.reshape(B, Cin, F_, H // 2, 2, W // 2, 2).permute(0, 2, 3, 5, 1, 4, 6)
        out.copy_(t.reshape(B * F_ * (H // 2) * (W // 2), Cin * 4))

    def sinusoid(self, t, out, freq_dim):
        self.launches += 1
        out.copy_(O.sinusoid_embed(freq_dim, t).to(out.dtype))

    def skinny_linear(self, x, w, bias, y, silu_in):
        self.launches += 1
        y.copy_(F.linear(F.silu(x) if silu_in else x, w, bias))

    def head_finish(self, head_out, xt, timestep, timesteps, sigmas, flow, x0):
        self.launches += 1
        B, F_, Cout, H, W = xt.shape
        y = head_out.reshape(B, F_, H // 2, W // 2, 1, 2, 2, Cout)
        y = y.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, Cout, F_, H, W).permute(0, 2, 1, 3, 4)
        flow.copy_(y)
        if x0 is not None:
            ts, sg = timesteps.double(), sigmas.double()
            idx = torch.argmin((ts.unsqueeze(0) - timestep.flatten().double().unsqueeze(1)).abs(), dim=1)
            s = sg[idx].reshape(B, F_, 1, 1, 1)
            x0.copy_((xt.double() - s * flow.double()).to(flow.dtype))

    def add_noise(self, x0, noise, timestep, timesteps, sigmas, out):
        self.launches += 1
        idx = torch.argmin((timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)
        s = sigmas[idx].reshape(-1, 1, 1, 1)
        out.copy_(((1 - s) * x0 + s * noise).type_as(noise))

    def cfg_unipc_step(self, flow_cond, flow_uncond, sample, last_sample, m0, m1, m_out, sample_out, prev_out, coef,
                       corrector_order, predictor_order):
        """Op-by-op tensor chain of the reference (causal_diffusion_inference.py:420-421,
        fm_solvers_unipc.py:320-323, :606-624, :465-482); python-float scalars stay fp32 against bf16 tensors."""
        self.launches += 1
        g, sigma, cx, cm0, cb, cirk, crho0, crho1, px, pm0, pb, pirk = [float(c) for c in coef]
        flow = flow_cond if flow_uncond is None else flow_uncond + g * (flow_cond - flow_uncond)
        mt = sample - sigma * flow
        xc = sample.clone()
        if corrector_order > 0:
            base = cx * last_sample - cm0 * m0
            inner = crho1 * (mt - m0)
            if corrector_order == 2:
                inner = crho0 * ((m1 - m0) * cirk) + inner
            xc = base - cb * inner
        nxt = px * xc - pm0 * mt
        if predictor_order == 2:
            nxt = nxt - pb * (0.5 * ((m0 - mt) * pirk))
        m_out.copy_(mt)
        sample_out.copy_(xc)
        prev_out.copy_(nxt)

    # ---- Wan VAE decoder ops (channels-last in, channels-first torch inside) -----------------------------------
    def vae_latent_in(self, z_frame, mean, inv_std, w, bias, out):
        self.launches += 1
        x = z_frame / inv_std.view(16, 1, 1) + mean.view(16, 1, 1)
        y = F.conv3d(x.unsqueeze(0).unsqueeze(2), w.view(16, 16, 1, 1, 1), bias)          # [1, 16, 1, h, w]
        out.copy_(y[0, :, 0].permute(1, 2, 0).reshape(-1, 16))

    def vae_norm_silu(self, x, gamma, y, silu):
        self.launches += 1
        n = F.normalize(x, dim=1) * x.shape[1] ** 0.5 * gamma
        y.copy_(F.silu(n) if silu else n)

    def upsample2x(self, x, y):
        self.launches += 1
        y.copy_(x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))

    def causal_conv3d(self, x, t_zero_pad, w, bias, kt, ks, y0, y1=None, *, upsample=False, residual=None, seg_cols=0,
                      implicit=False):
        self.launches += 1
        t_in, H, W, Cin = x.shape
        cout = w.shape[0]
        xc = x.permute(3, 0, 1, 2).unsqueeze(0)                                          # [1, Cin, T, H, W]
        if upsample:
            xc = F.interpolate(xc[0].permute(1, 0, 2, 3).float(), scale_factor=(2.0, 2.0), mode="nearest").type_as(x)
            xc = xc.permute(1, 0, 2, 3).unsqueeze(0)
        xc = F.pad(xc, (ks // 2, ks // 2, ks // 2, ks // 2, t_zero_pad, 0))
        w5 = w.view(cout, kt, ks, ks, Cin).permute(0, 4, 1, 2, 3)
        y = F.conv3d(xc, w5, bias)[0].permute(1, 2, 3, 0).reshape(-1, cout)
        if residual is not None:
            y = residual + y
        if y1 is None:
            y0.copy_(y)
        else:
            y0.copy_(y[:, :seg_cols])
            y1.copy_(y[:, seg_cols:2 * seg_cols])

    def softmax_rows(self, s, p, scale):
        self.launches += 1
        p.copy_(torch.softmax(s * scale, dim=1).to(p.dtype))

    def transpose(self, x, out):
        self.launches += 1
        out.copy_(x.t())

    def vae_pixel_out(self, y, out):
        self.launches += 1
        T, _, H, W = out.shape
        out.copy_(y[:, :3].float().clamp(-1, 1).reshape(T, H, W, 3).permute(0, 3, 1, 2))

    # ---- UMT5 text encoder ops ----------------------------------------------------------------------------------
    def t5_rmsnorm(self, x, w, y, eps):
        self.launches += 1
        n = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
        y.copy_(w * (n.type_as(w) if w.dtype in (torch.float16, torch.bfloat16) else n))

    def softmax_bias_rows(self, s, bias, key_mask, p):
        self.launches += 1
        ab = bias.clone()
        if key_mask is not None:
            ab.masked_fill_(key_mask.view(1, -1) == 0, torch.finfo(s.dtype).min)
        p.copy_(F.softmax((s + ab).float(), dim=-1).to(p.dtype))

    def t5_gated_gelu(self, fc1, gate, out):
        self.launches += 1
        import math
        g = 0.5 * gate * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (gate + 0.044715 * torch.pow(gate, 3.0))))
        out.copy_(fc1 * g)
