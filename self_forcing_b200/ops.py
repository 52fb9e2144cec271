"""Tensor-level wrappers over the C ABI (include/sfb200.h): torch supplies device memory and the
current stream, everything else happens in libsfb200.so.  No fallbacks."""
from __future__ import annotations

import ctypes
import functools
from typing import Optional, Sequence

import torch

from . import _lib

EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_GATE_RES, EPI_F32 = 0, 1, 2, 3, 4
STATS_CHUNK = 128   # columns per row-statistics record (include/sfb200.h: SFB_STATS_CHUNK)
_T_DTYPE = {torch.float32: 0, torch.int64: 1, torch.float64: 2, torch.bfloat16: 3}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _check_2d(t: torch.Tensor, name: str) -> None:
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.bfloat16 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA bf16 matrix with unit column stride, got "
                         f"{tuple(t.shape)} {t.dtype} {t.device} strides {t.stride()}")


def _op(fn):
    """Counts the launch and, while profiling is on, brackets it with CUDA events on the launching stream."""
    name = fn.__name__

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        self.launches += 1
        if self._prof is None or (self._prof_only is not None and name not in self._prof_only):
            return fn(self, *a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(self, *a, **k)
        e1.record()
        self._prof.append((name, self._prof_tag(name, a, k), e0, e1))
        return out
    return wrapped


class CudaOps:
    """The op set the host model is written against.  Every method enqueues one kernel of libsfb200.so on the
    current CUDA stream (attention: plus the small partial-merge kernel when a long KV window is split across SMs);
    nothing synchronises with the host, so a whole forward can be captured into a CUDA graph."""

    requires_bf16 = True
    supports_cuda_graphs = True   # every op only enqueues kernels on the current stream (no host sync, no allocation)
    supports_row_stats = True     # gemm(stats_out= / ln_stats=) and attention(q_stats=): norms folded into the neighbouring kernels

    def __init__(self):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SfbError("CUDA device required: the B200 path has no CPU fallback")
        self.launches = 0
        self._prof = None
        self._prof_only = None
        self._attn_ws = {}    # device index -> uint8 scratch for the split-KV attention schedule
        self._conv_ws = {}    # device index -> staging buffer of the VAE convolutions' gathered operand
        self.conv_workspace_bytes = 1 << 30

    # -- optional per-launch device timing (bench.py roofline leg) ------------------------------
    def start_profile(self, only=None):
        """only: optional set of op names to bracket (the others run untouched)."""
        self._prof = []
        self._prof_only = set(only) if only else None

    def stop_profile(self):
        """-> list of (op name, tag, milliseconds) for every launch since start_profile()."""
        torch.cuda.synchronize()
        out = [(n, tag, e0.elapsed_time(e1)) for n, tag, e0, e1 in (self._prof or [])]
        self._prof = None
        return out

    @staticmethod
    def _prof_tag(name, a, k):
        if name == "attention":
            q, kk = a[0], a[1]
            return ("attention", q.shape[0], q.shape[1], kk.shape[1], q.shape[2])
        if name == "attention_sp":
            q, kk = a[0], a[1]
            return ("attention", 1, q.shape[0], kk.shape[0], q.shape[1])
        if name == "gemm":
            x, w = a[0], a[1]
            return ("gemm", x.shape[0], w.shape[0], x.shape[1], k.get("epilogue", 0))
        if name == "causal_conv3d":
            x, w = a[0], a[2]
            return ("conv", x.shape[1], x.shape[2], x.shape[3], w.shape[0], f"k{a[4]}{a[5]}", "up" if k.get("upsample") else "",
                    "implicit" if k.get("implicit") else "gather")
        return (name,)

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    # -- dense projections ---------------------------------------------------------------
    @_op
    def gemm(self, x, w, bias, out, *, epilogue=EPI_BIAS, residual=None, gate=None, gate_stride=0,
             rows_per_gate=1, gate_row_offset=0, outs: Optional[Sequence[torch.Tensor]] = None, seg_cols=0,
             block_n=0, stats_out=None, ln_stats=None, ln_sc=None, ln_eps: float = 0.0):
        """stats_out: fp32 [M, N / 128, 2] receives (mean, M2) records of the output rows; ln_stats [M, K / 128, 2] + ln_sc
        [N, 2]: the LayerNorm in front of the Linear folded into the epilogue (include/sfb200.h: sfb_gemm_bf16_stats)."""
        _check_2d(x, "x"); _check_2d(w, "w")
        M, K = x.shape
        N = w.shape[0]
        assert w.shape[1] == K
        segs = list(outs) if outs is not None else [out]
        for s in segs:
            if epilogue == EPI_F32:      # fp32 result (one segment): strides are counted in floats
                assert s.dtype == torch.float32 and s.dim() == 2 and s.stride(1) == 1 and len(segs) == 1
            else:
                _check_2d(s, "out")
        while len(segs) < 3:
            segs.append(None)
        if residual is not None:
            _check_2d(residual, "residual")
        if stats_out is not None or ln_stats is not None:
            for s_, shape in ((stats_out, (M, N // STATS_CHUNK, 2)), (ln_stats, (M, K // STATS_CHUNK, 2)), (ln_sc, (N, 2))):
                assert s_ is None or (s_.dtype == torch.float32 and s_.is_contiguous() and tuple(s_.shape) == shape), \
                    f"statistics tensor {None if s_ is None else tuple(s_.shape)} != {shape} fp32 contiguous"
            _lib.check(self.lib.sfb_gemm_bf16_stats(
                x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), M, N, K, epilogue,
                _ptr(segs[0]), segs[0].stride(0), _ptr(segs[1]), segs[1].stride(0) if segs[1] is not None else 0,
                _ptr(segs[2]), segs[2].stride(0) if segs[2] is not None else 0, seg_cols,
                _ptr(residual), residual.stride(0) if residual is not None else 0,
                _ptr(gate), gate_stride, rows_per_gate, gate_row_offset, block_n, _ptr(stats_out), _ptr(ln_stats),
                _ptr(ln_sc), ln_eps, self._stream()), "sfb_gemm_bf16_stats")
            return
        _lib.check(self.lib.sfb_gemm_bf16(
            x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), M, N, K, epilogue,
            _ptr(segs[0]), segs[0].stride(0), _ptr(segs[1]), segs[1].stride(0) if segs[1] is not None else 0,
            _ptr(segs[2]), segs[2].stride(0) if segs[2] is not None else 0, seg_cols,
            _ptr(residual), residual.stride(0) if residual is not None else 0,
            _ptr(gate), gate_stride, rows_per_gate, gate_row_offset, block_n, 0, 0,
            self._stream()), "sfb_gemm_bf16")

    # -- attention ------------------------------------------------------------------------
    @_op
    def attention(self, q, k, v, out, scale: float, q_stats=None, q_eps: float = 0.0):
        """q/out [B, Lq, H, D] views, k/v [B, S, H, D] views (the cache window).  q_stats fp32 [B * Lq, chunks, 2]: q is the
        un-normalised projection and its full-width RMSNorm factor is applied inside the softmax (sfb_attention_fwd_qnorm)."""
        B, Lq, H, D = q.shape
        S = k.shape[1]
        for t in (q, k, v, out):
            assert t.stride(3) == 1 and t.stride(2) == D and t.dtype == torch.bfloat16
        assert k.stride() == v.stride() and k.shape == v.shape
        ws = self._attn_workspace(q.device)
        if q_stats is not None:
            assert q_stats.dtype == torch.float32 and q_stats.is_contiguous() and q_stats.dim() == 3 \
                and q_stats.shape[0] == B * Lq and q_stats.shape[2] == 2
            _lib.check(self.lib.sfb_attention_fwd_qnorm(
                q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(1), k.stride(0),
                out.data_ptr(), out.stride(1), out.stride(0), B, Lq, S, H, D, scale, q_stats.data_ptr(), q_stats.shape[1],
                q_eps, ws.data_ptr(), ws.numel(), self._stream()), "sfb_attention_fwd_qnorm")
            return
        _lib.check(self.lib.sfb_attention_fwd(
            q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(1), k.stride(0),
            out.data_ptr(), out.stride(1), out.stride(0), B, Lq, S, H, D, scale, ws.data_ptr(), ws.numel(),
            self._stream()), "sfb_attention_fwd")

    # -- normalisation / modulation -------------------------------------------------------
    @_op
    def modulation_table(self, mod, e, out, e_row_stride: int, e_group_stride: int):
        NL, G, C = mod.shape
        R = out.shape[1]
        assert out.shape == (NL, R, G, C) and out.is_contiguous() and mod.is_contiguous()
        _lib.check(self.lib.sfb_modulation_table(mod.data_ptr(), e.data_ptr(), out.data_ptr(), NL, R, G, C,
                                                 e_row_stride, e_group_stride, self._stream()),
                   "sfb_modulation_table")

    @_op
    def ln_modulate(self, x, y, shift, scale, mod_stride: int, rows_per_mod: int, eps: float, row_offset: int = 0,
                    stats=None):
        """stats: fp32 [rows, C / 128, 2] statistics records of x from the GEMM that wrote it -> streaming kernel."""
        _check_2d(x, "x"); _check_2d(y, "y")
        if stats is not None:
            assert stats.dtype == torch.float32 and stats.is_contiguous() and stats.dim() == 3 \
                and stats.shape[0] == x.shape[0] and stats.shape[2] == 2
            _lib.check(self.lib.sfb_ln_modulate_stats(
                x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[0], x.shape[1], eps, shift.data_ptr(),
                scale.data_ptr(), mod_stride, rows_per_mod, row_offset, stats.data_ptr(), stats.shape[1], self._stream()),
                "sfb_ln_modulate_stats")
            return
        _lib.check(self.lib.sfb_ln_modulate(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[0],
                                            x.shape[1], eps, shift.data_ptr(), scale.data_ptr(), mod_stride,
                                            rows_per_mod, row_offset, self._stream()), "sfb_ln_modulate")

    @_op
    def ln_affine(self, x, y, weight, bias, eps: float):
        _check_2d(x, "x"); _check_2d(y, "y")
        _lib.check(self.lib.sfb_ln_affine(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[0],
                                          x.shape[1], eps, weight.data_ptr(), bias.data_ptr(), self._stream()),
                   "sfb_ln_affine")

    @_op
    def rmsnorm(self, x, y, weight, eps: float):
        _check_2d(x, "x"); _check_2d(y, "y")
        _lib.check(self.lib.sfb_rmsnorm(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[0],
                                        x.shape[1], eps, weight.data_ptr(), self._stream()), "sfb_rmsnorm")

    @_op
    def qk_norm_rope(self, q_in, k_in, v_in, wq, wk, eps, cos_tab, sin_tab, B, L, head_dim, grid, start_frame,
                     q_out, k_out, v_out, start_frame_dev=None, stats=None, q_chunk0=0, k_chunk0=0):
        """q_in/k_in/v_in [B*L, C]; q_out [B, L, C]-like view; k_out/v_out [B, L, H, D] cache-slot views.
        start_frame_dev: optional int32 device scalar that overrides start_frame at run time (graph replay).
        stats: fp32 [B*L, chunks, 2] statistics records of the QKV projection's output rows (gemm(stats_out=)); q's
        records start at chunk q_chunk0, k's at k_chunk0 -- the rows are then streamed (sfb_qk_norm_rope_stats)."""
        if start_frame_dev is not None:
            assert start_frame_dev.dtype == torch.int32 and start_frame_dev.numel() == 1
        _check_2d(q_in, "q_in"); _check_2d(k_in, "k_in")
        C = q_in.shape[1]
        F_, Hh, Ww = grid
        assert cos_tab.dtype == torch.float32 and cos_tab.is_contiguous() and sin_tab.is_contiguous()
        assert k_out.stride() == v_out.stride()
        if stats is not None:
            assert stats.dtype == torch.float32 and stats.is_contiguous() and stats.dim() == 3 and stats.shape[0] == B * L \
                and stats.shape[2] == 2
            _lib.check(self.lib.sfb_qk_norm_rope_stats(
                q_in.data_ptr(), q_in.stride(0), k_in.data_ptr(), k_in.stride(0), _ptr(v_in),
                v_in.stride(0) if v_in is not None else 0, wq.data_ptr(), wk.data_ptr(), eps, stats.data_ptr(),
                stats.shape[1], q_chunk0, k_chunk0, cos_tab.data_ptr(), sin_tab.data_ptr(), cos_tab.shape[0], B, L, C,
                head_dim, F_, Hh, Ww, start_frame, _ptr(start_frame_dev), q_out.data_ptr(), q_out.stride(1),
                q_out.stride(0), k_out.data_ptr(), v_out.data_ptr(), k_out.stride(1), k_out.stride(0), self._stream()),
                "sfb_qk_norm_rope_stats")
            return
        _lib.check(self.lib.sfb_qk_norm_rope(
            q_in.data_ptr(), q_in.stride(0), k_in.data_ptr(), k_in.stride(0), _ptr(v_in),
            v_in.stride(0) if v_in is not None else 0, wq.data_ptr(), wk.data_ptr(), eps, cos_tab.data_ptr(),
            sin_tab.data_ptr(), cos_tab.shape[0], B, L, C, head_dim, F_, Hh, Ww, start_frame, _ptr(start_frame_dev),
            q_out.data_ptr(), q_out.stride(1), q_out.stride(0), k_out.data_ptr(), v_out.data_ptr(),
            k_out.stride(1), k_out.stride(0), self._stream()), "sfb_qk_norm_rope")

    # -- Ulysses head-parallel path (self_forcing_b200/ulysses.py) ---------------------------
    def _attn_workspace(self, device):
        ws = self._attn_ws.get(device.index)
        if ws is None:
            with torch.cuda.device(device):
                ws = torch.empty(int(self.lib.sfb_attention_workspace_bytes()), dtype=torch.uint8, device=device)
            self._attn_ws[device.index] = ws
        return ws

    @_op
    def qk_norm_rope_sp(self, q_in, k_in, v_in, wq, wk, eps, cos_tab, sin_tab, head_dim, grid, start_frame,
                        token_offset: int, sp, q_buf, k_slot_local, v_slot_local, k_slot_ptrs, v_slot_ptrs):
        """This rank's token rows (all heads) -> head group g's rows of rank g's q buffer / KV-cache slot, stored
        through peer-mapped pointers (the forward all-to-all).  q_buf: PeerTensor [L, H/P * D]."""
        _check_2d(q_in, "q_in"); _check_2d(k_in, "k_in"); _check_2d(v_in, "v_in")
        C = q_in.shape[1]
        F_, Hh, Ww = grid
        group_cols = C // sp.world
        _lib.check(self.lib.sfb_qk_norm_rope_sp(
            q_in.data_ptr(), q_in.stride(0), k_in.data_ptr(), k_in.stride(0), v_in.data_ptr(), v_in.stride(0),
            wq.data_ptr(), wk.data_ptr(), eps, cos_tab.data_ptr(), sin_tab.data_ptr(), cos_tab.shape[0],
            q_in.shape[0], C, head_dim, F_, Hh, Ww, start_frame, token_offset, sp.world,
            _lib.ptr_array(q_buf.ptrs), group_cols, _lib.ptr_array(k_slot_ptrs), _lib.ptr_array(v_slot_ptrs),
            group_cols, self._stream()), "sfb_qk_norm_rope_sp")

    @_op
    def attention_sp(self, q, k, v, scale: float, sp, out_buf, rows_per_rank: int):
        """q [L, H/P, D] (this rank's head group, all tokens), k / v [S, H/P, D] cache window; output rows of rank d's
        tokens are stored into rank d's out_buf (PeerTensor [L/P, C]) at this head group's columns (reverse
        all-to-all)."""
        Lq, Hg, D = q.shape
        S = k.shape[0]
        for t in (q, k, v):
            assert t.stride(2) == 1 and t.stride(1) == D and t.dtype == torch.bfloat16
        assert k.stride() == v.stride() and out_buf.local.stride(1) == 1
        ws = self._attn_workspace(q.device)
        col0 = sp.rank * Hg * D
        _lib.check(self.lib.sfb_attention_fwd_sp(
            q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), _lib.ptr_array(out_buf.ptrs_at(col0)),
            sp.world, rows_per_rank, out_buf.local.stride(0), Lq, S, Hg, D, scale, ws.data_ptr(), ws.numel(),
            self._stream()), "sfb_attention_fwd_sp")

    @_op
    def peer_barrier(self, sp):
        _lib.check(self.lib.sfb_peer_barrier(_lib.ptr_array(sp.flags.ptrs), sp.rank, sp.world, self._stream()),
                   "sfb_peer_barrier")

    # -- rolling KV window -----------------------------------------------------------------
    @_op
    def kv_roll(self, tensors, table, dst_row: int, src_row: int, n_rows: int):
        """Shift rows [src_row, src_row + n_rows) of every [B, S, H, D] cache tensor in `tensors` to dst_row (< src_row)
        -- the eviction of the rolling window (causal_model.py:212-221) for all layers' K and V in one call.
        `table` = int64 device tensor holding the tensors' data pointers (built once per cache, see model.py)."""
        t0 = tensors[0]
        for t in tensors:
            assert t.shape == t0.shape and t.stride() == t0.stride() and t.dtype == t0.dtype
        assert t0.stride(3) == 1 and t0.stride(2) == t0.shape[3] and t0.stride(1) == t0.shape[2] * t0.shape[3]
        assert table.dtype == torch.int64 and table.numel() == len(tensors) and table.is_cuda
        es = t0.element_size()
        _lib.check(self.lib.sfb_kv_roll(table.data_ptr(), len(tensors), t0.shape[0], t0.stride(0) * es, t0.stride(1) * es,
                                        dst_row, src_row, n_rows, self._stream()), "sfb_kv_roll")

    # -- embeddings -----------------------------------------------------------------------
    @_op
    def patchify(self, x, out):
        """x [B, Cin, F, H, W] (any strides) -> out [B*F*(H/2)*(W/2), Cin*4]."""
        B, Cin, F_, H, W = x.shape
        assert out.is_contiguous() and x.dtype == torch.bfloat16
        sb, sc, sf, sy, sx = x.stride()
        _lib.check(self.lib.sfb_patchify(x.data_ptr(), sb, sc, sf, sy, sx, out.data_ptr(), B, Cin, F_, H, W,
                                         self._stream()), "sfb_patchify")

    @_op
    def sinusoid(self, t, out, freq_dim: int):
        assert t.is_contiguous() and t.dtype in _T_DTYPE
        _lib.check(self.lib.sfb_sinusoid(t.data_ptr(), _T_DTYPE[t.dtype], out.data_ptr(), t.numel(), freq_dim,
                                         self._stream()), "sfb_sinusoid")

    @_op
    def skinny_linear(self, x, w, bias, y, silu_in: bool):
        _check_2d(x, "x"); _check_2d(w, "w"); _check_2d(y, "y")
        _lib.check(self.lib.sfb_skinny_linear(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
                                              y.data_ptr(), y.stride(0), x.shape[0], w.shape[0], x.shape[1],
                                              1 if silu_in else 0, self._stream()), "sfb_skinny_linear")

    # -- sampler --------------------------------------------------------------------------
    @_op
    def head_finish(self, head_out, xt, timestep, timesteps, sigmas, flow, x0):
        """head_out [B*L, 4*Cout]; xt [B, F, Cout, H, W] view; timestep [B, F]; flow/x0 contiguous outputs."""
        B, F_, Cout, H, W = xt.shape
        assert flow.is_contiguous() and (x0 is None or x0.is_contiguous())
        assert timestep.is_contiguous() and timestep.dtype in _T_DTYPE
        sb, sf, sc, sy, sx = xt.stride()
        _lib.check(self.lib.sfb_head_finish(
            head_out.data_ptr(), head_out.stride(0), xt.data_ptr(), sb, sf, sc, sy, sx, timestep.data_ptr(),
            _T_DTYPE[timestep.dtype], _ptr(timesteps), _ptr(sigmas), 0 if timesteps is None else timesteps.numel(),
            flow.data_ptr(), _ptr(x0), B, F_, Cout, H, W, self._stream()), "sfb_head_finish")

    @_op
    def add_noise(self, x0, noise, timestep, timesteps, sigmas, out):
        """x0/noise/out [N, C, H, W] contiguous, timestep [N]."""
        assert x0.is_contiguous() and noise.is_contiguous() and out.is_contiguous()
        assert timestep.is_contiguous() and timestep.dtype in _T_DTYPE
        n = x0.shape[0]
        _lib.check(self.lib.sfb_add_noise(x0.data_ptr(), noise.data_ptr(), timestep.data_ptr(),
                                          _T_DTYPE[timestep.dtype], timesteps.data_ptr(), sigmas.data_ptr(),
                                          timesteps.numel(), out.data_ptr(), n, x0[0].numel(), self._stream()),
                   "sfb_add_noise")

    @_op
    def cfg_unipc_step(self, flow_cond, flow_uncond, sample, last_sample, m0, m1, m_out, sample_out, prev_out, coef,
                       corrector_order: int, predictor_order: int):
        """One fused CFG + UniPC step (sfb200.h).  Tensors bf16 contiguous with the same numel; `flow_uncond`,
        `last_sample`, `m0`, `m1` may be None where the step does not use them; coef = 12 python floats."""
        tensors = [t for t in (flow_cond, flow_uncond, sample, last_sample, m0, m1, m_out, sample_out, prev_out)
                   if t is not None]
        n = sample.numel()
        for t in tensors:
            assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == n
        ptr = _ptr
        host = (ctypes.c_float * 12)(*[float(c) for c in coef])
        _lib.check(self.lib.sfb_cfg_unipc_step(ptr(flow_cond), ptr(flow_uncond), ptr(sample), ptr(last_sample), ptr(m0),
                                               ptr(m1), ptr(m_out), ptr(sample_out), ptr(prev_out), n, host,
                                               int(corrector_order), int(predictor_order), self._stream()),
                   "sfb_cfg_unipc_step")

    # -- Wan VAE decoder (channels-last frames [T, H, W, C]) --------------------------------
    @_op
    def vae_latent_in(self, z_frame, mean, inv_std, w, bias, out):
        """z_frame [16, h, w] (any channel stride, contiguous h*w), out [h*w, 16]."""
        assert z_frame.dtype == torch.bfloat16 and z_frame.stride(2) == 1 and z_frame.stride(1) == z_frame.shape[2]
        assert out.is_contiguous() and w.is_contiguous()
        _lib.check(self.lib.sfb_vae_latent_in(z_frame.data_ptr(), z_frame.stride(0), mean.data_ptr(), inv_std.data_ptr(),
                                              w.data_ptr(), bias.data_ptr(), out.data_ptr(), out.shape[0], self._stream()),
                   "sfb_vae_latent_in")

    @_op
    def vae_norm_silu(self, x, gamma, y, silu: bool):
        """x, y [rows, C] (unit column stride), gamma [C]."""
        _check_2d(x, "x"); _check_2d(y, "y")
        _lib.check(self.lib.sfb_vae_norm_silu(x.data_ptr(), x.stride(0), gamma.data_ptr(), y.data_ptr(), y.stride(0),
                                              x.shape[0], x.shape[1], int(silu), self._stream()), "sfb_vae_norm_silu")

    @_op
    def causal_conv3d(self, x, t_zero_pad: int, w, bias, kt: int, ks: int, y0, y1=None, *, upsample=False,
                      residual=None, seg_cols=0, implicit=False):
        """x [t_in, H, W, Cin] contiguous (cached frames first); w packed [Cout, kt*ks*ks*Cin]; y0 (/y1) [rows, seg]
        with a common row stride; residual [rows, Cout]."""
        assert x.is_contiguous() and w.is_contiguous() and x.dtype == torch.bfloat16
        t_in, H, W, Cin = x.shape
        _check_2d(y0, "y0")
        if y1 is not None:
            _check_2d(y1, "y1")
            assert y1.stride(0) == y0.stride(0)
        if residual is not None:
            _check_2d(residual, "residual")
        ws = None if implicit else self._conv_ws.get(x.device.index)
        if ws is None and not implicit:
            with torch.cuda.device(x.device):
                ws = torch.empty(self.conv_workspace_bytes, dtype=torch.uint8, device=x.device)
            self._conv_ws[x.device.index] = ws
        _lib.check(self.lib.sfb_causal_conv3d_cl(
            x.data_ptr(), t_in, H, W, Cin, t_zero_pad, int(upsample), w.data_ptr(), _ptr(bias), w.shape[0], kt, ks,
            _ptr(residual), residual.stride(0) if residual is not None else 0, y0.data_ptr(), _ptr(y1), y0.stride(0),
            seg_cols, _ptr(ws), ws.numel() if ws is not None else 0, self._stream()), "sfb_causal_conv3d_cl")

    @_op
    def upsample2x(self, x, y):
        """x [T, H, W, C] -> y [T, 2H, 2W, C], nearest."""
        assert x.is_contiguous() and y.is_contiguous() and x.dtype == torch.bfloat16
        T, H, W, C = x.shape
        assert y.shape == (T, 2 * H, 2 * W, C)
        _lib.check(self.lib.sfb_upsample2x_cl(x.data_ptr(), y.data_ptr(), T, H, W, C, self._stream()), "sfb_upsample2x_cl")

    @_op
    def softmax_rows(self, s, p, scale: float):
        """s fp32 [rows, cols] -> p bf16 [rows, cols] = softmax(scale * s)."""
        _check_2d(p, "p")
        assert s.dtype == torch.float32 and s.stride(1) == 1 and s.shape == p.shape
        _lib.check(self.lib.sfb_softmax_rows(s.data_ptr(), s.stride(0), p.data_ptr(), p.stride(0), s.shape[0], s.shape[1],
                                             scale, self._stream()), "sfb_softmax_rows")

    @_op
    def transpose(self, x, out):
        _check_2d(x, "x"); _check_2d(out, "out")
        assert out.shape == (x.shape[1], x.shape[0])
        _lib.check(self.lib.sfb_transpose_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0],
                                               x.shape[1], self._stream()), "sfb_transpose_bf16")

    # -- UMT5 text encoder ------------------------------------------------------------------
    @_op
    def t5_rmsnorm(self, x, w, y, eps: float):
        _check_2d(x, "x"); _check_2d(y, "y")
        _lib.check(self.lib.sfb_t5_rmsnorm(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[0], x.shape[1], eps,
                                           w.data_ptr(), self._stream()), "sfb_t5_rmsnorm")

    @_op
    def softmax_bias_rows(self, s, bias, key_mask, p):
        """s, bias, p bf16 [rows, cols]; key_mask int32 [cols] or None."""
        _check_2d(s, "s"); _check_2d(bias, "bias"); _check_2d(p, "p")
        assert key_mask is None or (key_mask.dtype == torch.int32 and key_mask.is_contiguous())
        _lib.check(self.lib.sfb_softmax_bias_rows(s.data_ptr(), s.stride(0), bias.data_ptr(), bias.stride(0), _ptr(key_mask),
                                                  p.data_ptr(), p.stride(0), s.shape[0], s.shape[1], self._stream()),
                   "sfb_softmax_bias_rows")

    @_op
    def t5_gated_gelu(self, fc1, gate, out):
        _check_2d(fc1, "fc1"); _check_2d(gate, "gate"); _check_2d(out, "out")
        _lib.check(self.lib.sfb_t5_gated_gelu(fc1.data_ptr(), fc1.stride(0), gate.data_ptr(), gate.stride(0), out.data_ptr(),
                                              out.stride(0), fc1.shape[0], fc1.shape[1], self._stream()), "sfb_t5_gated_gelu")

    @_op
    def vae_pixel_out(self, y, out):
        """y [T*H*W, ld >= 3] bf16, out fp32 [T, 3, H, W] contiguous."""
        _check_2d(y, "y")
        assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[1] == 3
        T = out.shape[0]
        hw = out.shape[2] * out.shape[3]
        assert y.shape[0] == T * hw
        _lib.check(self.lib.sfb_vae_pixel_out(y.data_ptr(), y.stride(0), out.data_ptr(), T, hw, self._stream()),
                   "sfb_vae_pixel_out")
