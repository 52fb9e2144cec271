"""B200CausalWanModel -- drop-in for `CausalWanModel` on the KV-cached inference path.

Same constructor arguments, `state_dict` keys and `forward(x, t=, context=, seq_len=, kv_cache=,
crossattn_cache=, current_start=, cache_start=)` surface as wan/modules/causal_model.py:370-1128
(reference), but `_forward_inference` (:725-893) is re-implemented as a schedule of hand-written
sm_100a kernels (libsfb200.so, include/sfb200.h) with no PyTorch compute on the path:

  per forward   patchify -> GEMM(+bias)                 patch embed            (ref :775-781)
                sinusoid -> 3 skinny linears            time MLPs              (ref :829-832)
                modulation tables for all layers        e = modulation + e0    (ref :310, :365)
                [first call per prompt] text MLP + cross K/V into crossattn_cache (ref :837-842, model.py:175-180)
  per block     LN+modulate -> QKV GEMM -> QK-RMSNorm+RoPE+KV-append -> attention over the cache window
                -> O GEMM (+gate+residual) -> LN affine -> Q GEMM -> RMSNorm -> cross attention
                -> O GEMM (+residual) -> LN+modulate -> FFN1 GEMM (+GELU) -> FFN2 GEMM (+gate+residual)
  head          LN+modulate -> GEMM -> unpatchify (+ flow -> x0)               (ref :356-367, :1081-1104)

The torch.nn modules below only *hold parameters* under the reference's names; nothing calls their
forward.  The training branch of the reference (`kv_cache is None`, FlexAttention) is out of scope
and raises.

Execution: every forward is split into a host prelude (cache plan + index mirror: integers only) and
`_device_forward` (all kernels, no host sync); `_device_forward` is replayed as ONE CUDA graph per static signature
(captured on its second occurrence).  `enable_ulysses` switches self-attention to head-parallel execution over several
GPUs (ulysses.py).  `B200WanModel` (bottom of this file) is the bidirectional teacher forward on the same schedule.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
from torch import nn

from .cache import IndexMirror, plan_cache_update
from .ops import EPI_GATE_RES, EPI_GELU, EPI_RESIDUAL, STATS_CHUNK
from .ulysses import UlyssesGroup, shard_rows


def rope_tables(head_dim: int, max_pos: int = 1024, theta: float = 10000.0):
    """cos / sin tables fp32 [max_pos, head_dim/2] of the reference's complex128 `freqs`
    (wan/modules/model.py:29-36 concatenated at causal_model.py:482-488): three theta ladders
    for the (frame, height, width) axes over d-4(d//6), 2(d//6), 2(d//6) real dims."""
    d = head_dim
    parts = []
    for dim in (d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)):
        inv = 1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim)
        parts.append(torch.outer(torch.arange(max_pos, dtype=torch.float64), inv))
    ang = torch.cat(parts, dim=1)
    return torch.cos(ang).float().contiguous(), torch.sin(ang).float().contiguous()


class _ParamHolder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - never used
        raise RuntimeError("parameter holder: the B200 path does not call torch module forwards")


class _Linear(_ParamHolder):
    def __init__(self, in_f: int, out_f: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f))
        self.bias = nn.Parameter(torch.empty(out_f))


class _Scale(_ParamHolder):
    def __init__(self, dim: int, bias: bool = False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(dim))
        if bias:
            self.bias = nn.Parameter(torch.empty(dim))


class _Attn(_ParamHolder):
    def __init__(self, dim: int):
        super().__init__()
        self.q, self.k, self.v, self.o = (_Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = _Scale(dim), _Scale(dim)


class _Block(_ParamHolder):
    def __init__(self, dim: int, ffn_dim: int):
        super().__init__()
        self.self_attn = _Attn(dim)
        self.norm3 = _Scale(dim, bias=True)
        self.cross_attn = _Attn(dim)
        self.ffn = nn.ModuleList([_Linear(dim, ffn_dim), _ParamHolder(), _Linear(ffn_dim, dim)])
        self.modulation = nn.Parameter(torch.empty(1, 6, dim))


class _Head(_ParamHolder):
    def __init__(self, dim: int, out_features: int):
        super().__init__()
        self.head = _Linear(dim, out_features)
        self.modulation = nn.Parameter(torch.empty(1, 2, dim))


class _PatchEmbed(_ParamHolder):
    def __init__(self, in_dim: int, dim: int, patch):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(dim, in_dim, *patch))
        self.bias = nn.Parameter(torch.empty(dim))


class B200CausalWanModel(nn.Module):
    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, local_attn_size=-1,
                 sink_size=0, qk_norm=True, cross_attn_norm=True, eps=1e-6, ops=None):
        super().__init__()
        if model_type != "t2v":
            raise NotImplementedError("only the t2v model type runs the cached path (SURVEY.md section 9)")
        if tuple(patch_size) != (1, 2, 2) or not qk_norm or not cross_attn_norm:
            raise NotImplementedError("B200 path implements patch (1,2,2), qk_norm and cross_attn_norm")
        assert dim % num_heads == 0 and (dim // num_heads) % 2 == 0
        self.model_type, self.patch_size, self.text_len = model_type, tuple(patch_size), text_len
        self.in_dim, self.dim, self.ffn_dim, self.freq_dim = in_dim, dim, ffn_dim, freq_dim
        self.text_dim, self.out_dim, self.num_heads, self.num_layers = text_dim, out_dim, num_heads, num_layers
        self.local_attn_size, self.sink_size, self.qk_norm = local_attn_size, sink_size, qk_norm
        self.cross_attn_norm, self.eps = cross_attn_norm, eps
        self.head_dim = dim // num_heads
        # reference hard-codes 1560 tokens/frame here (causal_model.py:77)
        self.max_attention_size = 32760 if local_attn_size == -1 else local_attn_size * 1560
        self.num_frame_per_block = 1
        self.independent_first_frame = False
        self.block_mask = None

        self.patch_embedding = _PatchEmbed(in_dim, dim, self.patch_size)
        self.text_embedding = nn.ModuleList([_Linear(text_dim, dim), _ParamHolder(), _Linear(dim, dim)])
        self.time_embedding = nn.ModuleList([_Linear(freq_dim, dim), _ParamHolder(), _Linear(dim, dim)])
        self.time_projection = nn.ModuleList([_ParamHolder(), _Linear(dim, dim * 6)])
        self.blocks = nn.ModuleList([_Block(dim, ffn_dim) for _ in range(num_layers)])
        self.head = _Head(dim, out_dim * math.prod(self.patch_size))

        self._ops = ops
        self._packed: Optional[Dict[str, object]] = None
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self._mirror = IndexMirror()
        self._sampler_tables = None   # (timesteps fp32, sigmas fp32) on device, set by the wrapper
        self._rope = None
        self._sp: Optional[UlyssesGroup] = None   # Ulysses head-parallel group (one long video over several GPUs)
        self._sp_kv: Dict[int, tuple] = {}        # data_ptr of a head-sharded cache tensor -> (pool, element offset)
        self._sp_buf: Dict[int, tuple] = {}       # chunk length -> (q buffer, attention-output buffer) peer tensors
        self._ptr_tables: Dict[tuple, torch.Tensor] = {}   # cache tensors' data pointers -> device pointer table (kv_roll)
        # Cross-attention with its two norms folded away (norm3 into the q projection's epilogue, norm_q into the softmax
        # scale and the cached K): on by default where the ops provide row statistics (the CTA-pair GEMM: dim % 256 == 0).
        self.fold_cross_norms = os.environ.get("SFB_NO_FOLD", "0") != "1"   # (the variable is a diagnostic A/B knob)
        self._ck_fold: Dict[tuple, torch.Tensor] = {}      # (layer, data_ptr of the cross-attention K cache) -> K * norm_q.weight
        # CUDA graphs: a cached forward is ~430 launches issued from Python; replaying the whole forward as one graph
        # removes the launch gaps (measured 29.6 -> 28.0 ms at S = 18720).  One graph per distinct static signature
        # (shapes, cache pointers, cache plan), captured on its second occurrence.
        self.use_cuda_graphs = True
        self.max_cuda_graphs = 64
        self._graphs: Dict[tuple, object] = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())
        self._register_load_state_dict_pre_hook(self._drop_foreign_keys)

    # ------------------------------------------------------------------ parameters
    @staticmethod
    def _drop_foreign_keys(state_dict, prefix, *args):
        # upstream / fork checkpoints carry `pose_proj.*` (causal_model.py:500-503); not on this path
        for k in [k for k in state_dict if k.startswith(prefix + "pose_proj.")]:
            del state_dict[k]

    @torch.no_grad()
    def init_weights(self, seed: int = 0) -> None:
        """Synthetic random init of the named architecture (no checkpoint): xavier-variance matrices as
        causal_model.py:1111-1125, but N(0,.02) biases / head weight and 1+N(0,.02) norm affines instead
        of the reference's zeros/ones so that every kernel does non-trivial work."""
        g = torch.Generator(device=self.patch_embedding.weight.device).manual_seed(seed)
        for name, p in self.named_parameters():
            shape = p.shape
            if name.endswith("modulation"):
                w = torch.randn(shape, generator=g, device=p.device) / math.sqrt(self.dim)
            elif name.endswith(".bias") or name == "head.head.weight":
                w = torch.randn(shape, generator=g, device=p.device) * 0.02
            elif "norm" in name:
                w = 1.0 + torch.randn(shape, generator=g, device=p.device) * 0.02
            else:
                fan_in = int(math.prod(shape[1:]))
                w = torch.randn(shape, generator=g, device=p.device) * math.sqrt(2.0 / (fan_in + shape[0]))
            p.copy_(w.to(p.dtype))
        self.invalidate_packed()

    def invalidate_packed(self) -> None:
        self._packed = None
        self._graphs.clear()
        self._ck_fold.clear()

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        self._graphs.clear()
        self._ws.clear()
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------ Ulysses head parallelism
    def enable_ulysses(self, sp: UlyssesGroup) -> None:
        """Run every forward head-parallel over the ranks of `sp` (self_forcing_b200/ulysses.py).  All ranks must
        call forward with identical inputs; KV caches must come from allocate_kv_cache()."""
        if self.num_heads % sp.world:
            raise ValueError(f"{self.num_heads} heads do not split over {sp.world} ranks")
        self._sp = sp if sp.world > 1 else None
        self._sp_kv.clear()
        self._sp_buf.clear()

    def allocate_kv_cache(self, batch_size: int, tokens: int, dtype, device, layers: Optional[int] = None) -> List[dict]:
        """Per-layer cache dicts like CausalInferencePipeline._initialize_kv_cache (pipeline/causal_inference.py:278-298).
        Under Ulysses the K/V tensors are head-sharded [1, S, H/P, D] views of one peer-mapped pool, because the
        other ranks' qk_norm_rope kernels store this rank's head group straight into it."""
        sp = self._sp
        heads = self.num_heads // (sp.world if sp is not None else 1)
        nl = self.num_layers if layers is None else layers
        if sp is None:
            pool = torch.zeros(nl, 2, batch_size, tokens, heads, self.head_dim, dtype=dtype, device=device)
            kv = [(pool[i, 0], pool[i, 1]) for i in range(nl)]
        else:
            if batch_size != 1:
                raise ValueError("Ulysses mode runs one video (batch 1) per group")
            peer = sp.alloc((nl, 2, tokens, heads, self.head_dim), dtype)
            peer.local.zero_()
            per = tokens * heads * self.head_dim
            kv = []
            for i in range(nl):
                k, v = peer.local[i, 0].unsqueeze(0), peer.local[i, 1].unsqueeze(0)
                self._sp_kv[k.data_ptr()] = (peer, (2 * i) * per)
                self._sp_kv[v.data_ptr()] = (peer, (2 * i + 1) * per)
                kv.append((k, v))
            sp.sync_host()
        return [{"k": k, "v": v,
                 "global_end_index": torch.tensor([0], dtype=torch.long, device=device),
                 "local_end_index": torch.tensor([0], dtype=torch.long, device=device)} for k, v in kv]

    def _sp_buffers(self, L: int, dtype):
        buf = self._sp_buf.get(L)
        if buf is None:
            sp = self._sp
            hg_cols = (self.num_heads // sp.world) * self.head_dim
            buf = (sp.alloc((L, hg_cols), dtype), sp.alloc((L // sp.world, self.dim), dtype))
            sp.sync_host()
            self._sp_buf[L] = buf
        return buf

    def _pointer_table(self, tensors, dev) -> torch.Tensor:
        """int64 device tensor of data pointers (argument of the multi-tensor roll kernel), cached per tensor set so
        that a replayed CUDA graph keeps reading a live buffer."""
        key = tuple(t.data_ptr() for t in tensors)
        tab = self._ptr_tables.get(key)
        if tab is None:
            if len(self._ptr_tables) > 64:
                self._ptr_tables.clear()
                self._graphs.clear()       # captured graphs may reference the dropped tables
            tab = torch.tensor(key, dtype=torch.int64, device=dev) if dev.type == "cuda" else torch.zeros(len(key), dtype=torch.int64)
            self._ptr_tables[key] = tab
        return tab

    def set_sampler_tables(self, timesteps: torch.Tensor, sigmas: torch.Tensor) -> None:
        """FlowMatchScheduler tables (utils/scheduler.py:118-141) used by the fused flow->x0 epilogue."""
        self._sampler_tables = (timesteps.detach().float().contiguous(), sigmas.detach().float().contiguous())

    @property
    def ops(self):
        if self._ops is None:
            from .ops import CudaOps
            self._ops = CudaOps()
        return self._ops

    @torch.no_grad()
    def _pack(self) -> Dict[str, object]:
        """Concatenated / stacked weight layouts the kernels read (built once per weight load)."""
        dev = self.patch_embedding.weight.device
        dt = self.patch_embedding.weight.dtype
        if dt != torch.bfloat16 and getattr(self.ops, "requires_bf16", True):
            raise TypeError(f"B200CausalWanModel runs in bfloat16 (call .to(torch.bfloat16)); got {dt}")
        pk: Dict[str, object] = {}
        pk["patch_w"] = self.patch_embedding.weight.detach().flatten(1).contiguous()
        blocks = []
        for b in self.blocks:
            sa, ca = b.self_attn, b.cross_attn
            blocks.append(dict(
                wqkv=torch.cat([sa.q.weight, sa.k.weight, sa.v.weight]).detach().contiguous(),
                bqkv=torch.cat([sa.q.bias, sa.k.bias, sa.v.bias]).detach().contiguous(),
                wkv_c=torch.cat([ca.k.weight, ca.v.weight]).detach().contiguous(),
                bkv_c=torch.cat([ca.k.bias, ca.v.bias]).detach().contiguous()))
        pk["blocks"] = blocks
        pk["fold"] = bool(self.fold_cross_norms and getattr(self.ops, "supports_row_stats", False) and self.dim % 256 == 0)
        if pk["fold"]:
            # norm3 (affine LayerNorm) folded into the cross-attention q projection (include/sfb200.h: sfb_gemm_bf16_stats):
            #   Linear(LN(x) * w3 + b3) = rstd * (x @ (W * w3)^T - mean * colsum(W * w3)) + (b + W @ b3)
            # and norm_q's weight folded into the cached text K (its row factor goes into the softmax scale).
            for b, pb in zip(self.blocks, blocks):
                ca = b.cross_attn
                wf = (ca.q.weight.detach().float() * b.norm3.weight.detach().float()[None, :]).to(dt)
                pb["wq_c_fold"] = wf.contiguous()
                pb["sc_c"] = torch.stack([wf.float().sum(dim=1),
                                          ca.q.bias.detach().float() + ca.q.weight.detach().float() @ b.norm3.bias.detach().float()],
                                         dim=1).contiguous()
                pb["gk_fold"] = (ca.norm_k.weight.detach().float() * ca.norm_q.weight.detach().float()).to(dt).contiguous()
        pk["mod"] = torch.cat([b.modulation.detach() for b in self.blocks]).contiguous()       # [NL, 6, C]
        pk["head_mod"] = self.head.modulation.detach().contiguous()                              # [1, 2, C]
        cos, sin = rope_tables(self.head_dim)
        pk["cos"], pk["sin"] = cos.to(dev), sin.to(dev)
        if self._sampler_tables is not None:
            self._sampler_tables = tuple(t.to(dev) for t in self._sampler_tables)
        self._packed = pk
        return pk

    def _workspace(self, B: int, L: int, F_: int, dev, L_full: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """L = token rows this rank works on per sample, L_full = tokens of the whole chunk (== L unless Ulysses)."""
        L_full = L if L_full is None else L_full
        key = (B, L, F_, str(dev), L_full)
        ws = self._ws.get(key)
        if ws is None:
            C, bf = self.dim, self.patch_embedding.weight.dtype
            R = B * L

            def e(*shape):
                return torch.empty(*shape, dtype=bf, device=dev)

            ws = dict(tok=e(B * L_full, self.in_dim * 4), head_full=e(B * L_full, self.out_dim * 4), x=e(R, C), h=e(R, C), q_lin=e(R, C), k_lin=e(R, C), v_lin=e(R, C),
                      q=e(R, C), attn=e(R, C), ffn=e(R, self.ffn_dim), sin=e(B * F_, self.freq_dim),
                      e1=e(B * F_, C), e=e(B * F_, C), e0=e(B * F_, 6 * C), mod=e(self.num_layers, B * F_, 6, C),
                      head_mod=e(1, B * F_, 2, C), head_out=e(R, self.out_dim * 4),
                      ctx_h=e(B * self.text_len, C), ctx=e(B * self.text_len, C), ctx_k=e(B * self.text_len, C))
            if C % STATS_CHUNK == 0:   # row-statistics records of x (after the self-attention residual) and of the cross q
                ws["x_stats"] = torch.empty(R, C // STATS_CHUNK, 2, dtype=torch.float32, device=dev)
                ws["q_stats"] = torch.empty(R, C // STATS_CHUNK, 2, dtype=torch.float32, device=dev)
                ws["qkv_stats"] = torch.empty(R, 3 * C // STATS_CHUNK, 2, dtype=torch.float32, device=dev)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ forward
    def forward(self, *args, **kwargs):
        """Dispatch like the reference (causal_model.py:1071-1079): with a KV cache the cached inference forward, without
        one the cache-free training-time forward (block-masked attention over the whole video; forward only)."""
        if kwargs.get("kv_cache", None) is None:
            return self._forward_train(*args, **kwargs)
        return self._forward_inference(*args, **kwargs)

    # ------------------------------------------------------------------ training-time forward (no cache)
    def attention_windows(self, num_frames: int, frame_tokens: int, teacher_forcing: bool):
        """The reference's FlexAttention block masks (causal_model.py:518-723) as dense rectangles: a list of
        (q_lo, q_hi, kv_lo, kv_hi) -- query rows [q_lo, q_hi) attend keys [kv_lo, kv_hi) -- in execution order.
        Block-wise causal: a chunk sees everything before its own end (only the last `local_attn_size` frames if set;
        a lone first frame if independent_first_frame).  Teacher forcing: sequence = [clean | noisy]; clean chunks are
        block-causal among themselves, noisy chunk c sees the clean chunks before it plus itself -- two intervals,
        made ONE rectangle by staging the noisy chunk's K/V over the clean chunk's rows (kv_lo < 0 marks that
        staging step: copy rows [-kv_lo - 1 ...) -- see _forward_train), walked from the last chunk to the first so that
        every clean chunk is overwritten only after all its readers are done.  masks.py holds the same intervals as
        BlockMask tables; tests check both against the reference's masks."""
        blk = frame_tokens * self.num_frame_per_block
        total = num_frames * frame_tokens
        chunks = []                      # (first row, last row + 1, ends[] value of the rows: :533-546 / :680-696)
        s = 0
        if self.independent_first_frame and not teacher_forcing:
            chunks.append((0, frame_tokens, frame_tokens))
            s = frame_tokens
        while s < total:
            chunks.append((s, min(s + blk, total), s + blk))     # a ragged last chunk keeps its nominal end
            s += blk
        wins = []
        for lo, hi, end in chunks:
            kv_lo = 0 if (self.local_attn_size == -1 or teacher_forcing) else max(0, end - self.local_attn_size * frame_tokens)
            # kv < ends[q]: a ragged last chunk keeps its nominal end, which under teacher forcing reaches into the
            # noisy half that follows the clean rows (the reference's mask does exactly that, :611-613,632-635)
            wins.append((lo, hi, kv_lo, min(end, total * (2 if teacher_forcing else 1))))
        if teacher_forcing:
            for lo, hi, _ in reversed(chunks):
                wins.append((total + lo, total + hi, -(lo + 1), hi))        # stage noisy K/V rows over clean rows [lo, hi), then [0, hi)
        return wins

    @torch.no_grad()
    def _forward_train(self, x, t, context, seq_len, clean_x=None, aug_t=None, clip_fea=None, y=None, add_condition=None,
                       return_x0: bool = False, **unsupported):
        """CausalWanModel._forward_train (causal_model.py:895-1069), forward only: all frames of the video at once, per-frame
        timesteps t [B, F], causality from the block mask; clean_x [B, 16, F, H, W] switches teacher forcing on (the
        sequence becomes [clean | noisy], aug_t = timesteps of the clean half, RoPE restarts for the noisy half, the head
        runs on the noisy half).  Same kernels as the cached forward; the mask is executed as dense attention calls over
        the rectangles of attention_windows().  Returns flow [B, 16, F, H, W] (and x0 [B, F, 16, H, W] with return_x0)."""
        if add_condition is not None or clip_fea is not None or y is not None:
            raise NotImplementedError("pose / image conditioning is not part of the t2v path")
        if any(v is not None for v in unsupported.values()):
            raise NotImplementedError(f"unsupported arguments {sorted(unsupported)}")
        if self._sp is not None:
            raise NotImplementedError("the training-time forward is not head-parallel")
        if isinstance(x, (list, tuple)):
            x = torch.stack(list(x))
        if isinstance(clean_x, (list, tuple)):
            clean_x = torch.stack(list(clean_x))
        if isinstance(context, (list, tuple)):
            context = torch.stack([torch.cat([u, u.new_zeros(self.text_len - u.size(0), u.size(1))]) for u in context])
        tf = clean_x is not None
        if tf and self.independent_first_frame:
            raise NotImplementedError("teacher forcing with an independent first frame (the reference raises too, :941-942)")
        ops, pk = self.ops, (self._packed or self._pack())
        B, Cin, F_, H, W = x.shape
        Hh, Ww = H // 2, W // 2
        fs, L = Hh * Ww, F_ * Hh * Ww
        assert L <= seq_len and t.shape == (B, F_), f"timestep shape {tuple(t.shape)} != {(B, F_)}"
        C, NL, D, NH = self.dim, self.num_layers, self.head_dim, self.num_heads
        dev, bf = x.device, self.patch_embedding.weight.dtype
        halves = 2 if tf else 1
        Lt, Fm = halves * L, halves * F_            # tokens / modulation frames per sample
        R = B * Lt
        key = ("train", B, Lt, Fm, str(dev))
        ws = self._ws.get(key)
        if ws is None:
            def e(*shape):
                return torch.empty(*shape, dtype=bf, device=dev)
            ws = dict(tok=e(L, self.in_dim * 4), x=e(R, C), h=e(R, C), q_lin=e(R, C), k_lin=e(R, C), q=e(R, C), attn=e(R, C),
                      ffn=e(R, self.ffn_dim), sin=e(B * Fm, self.freq_dim), e1=e(B * Fm, C), e=e(B * Fm, C),
                      e0=e(B * Fm, 6 * C), mod=e(NL, B * Fm, 6, C), head_mod=e(1, B * Fm, 2, C), head_out=e(B * L, self.out_dim * 4),
                      ctx_h=e(B * self.text_len, C), ctx=e(B * self.text_len, C), ctx_k=e(B * self.text_len, C),
                      k=e(B, Lt, NH, D), v=e(B, Lt, NH, D), ck=e(B, self.text_len, NH, D), cv=e(B, self.text_len, NH, D))
            self._ws[key] = ws
        X = ws["x"].view(B, Lt, C)

        # ---- embeddings: [clean | noisy] token rows per sample (causal_model.py:966-976,1011-1020) ---------------
        for b in range(B):
            for hf, src in enumerate((clean_x, x) if tf else (x,)):
                ops.patchify(src[b:b + 1], ws["tok"])
                ops.gemm(ws["tok"], pk["patch_w"], self.patch_embedding.bias, X[b, hf * L:(hf + 1) * L])
        tt = (torch.cat([aug_t if aug_t is not None else torch.zeros_like(t), t], dim=1) if tf else t).contiguous()
        ops.sinusoid(tt.reshape(-1), ws["sin"], self.freq_dim)
        te, tp = self.time_embedding, self.time_projection
        ops.skinny_linear(ws["sin"], te[0].weight, te[0].bias, ws["e1"], silu_in=False)
        ops.skinny_linear(ws["e1"], te[2].weight, te[2].bias, ws["e"], silu_in=True)
        ops.skinny_linear(ws["e"], tp[1].weight, tp[1].bias, ws["e0"], silu_in=True)
        ops.modulation_table(pk["mod"], ws["e0"], ws["mod"], e_row_stride=6 * C, e_group_stride=C)
        ops.modulation_table(pk["head_mod"], ws["e"], ws["head_mod"], e_row_stride=C, e_group_stride=0)
        if context.shape[1] < self.text_len:
            context = torch.cat([context, context.new_zeros(B, self.text_len - context.shape[1], context.shape[2])], dim=1)
        tx = self.text_embedding
        ops.gemm(context.reshape(B * self.text_len, self.text_dim).contiguous(), tx[0].weight, tx[0].bias, ws["ctx_h"], epilogue=EPI_GELU)
        ops.gemm(ws["ctx_h"], tx[2].weight, tx[2].bias, ws["ctx"])

        windows = self.attention_windows(F_, fs, tf)
        kv_table = self._pointer_table([ws["k"], ws["v"]], dev)
        scale = 1.0 / math.sqrt(D)
        mstride = 6 * C
        Q4, A4 = ws["q"].view(B, Lt, NH, D), ws["attn"].view(B, Lt, NH, D)
        for i, blk in enumerate(self.blocks):
            pb, m = pk["blocks"][i], ws["mod"][i]
            sa, ca = blk.self_attn, blk.cross_attn
            ops.ln_modulate(ws["x"], ws["h"], shift=m[:, 0], scale=m[:, 1], mod_stride=mstride, rows_per_mod=fs, eps=self.eps)
            ops.gemm(ws["h"], pb["wqkv"], pb["bqkv"], None, seg_cols=C, outs=[ws["q_lin"], ws["k_lin"], ws["v"].view(R, C)])
            # RoPE positions restart at frame 0 for every half of every sample (causal_model.py:127-136)
            for b in range(B):
                for hf in range(halves):
                    r0 = b * Lt + hf * L
                    ops.qk_norm_rope(ws["q_lin"][r0:r0 + L], ws["k_lin"][r0:r0 + L], None, sa.norm_q.weight, sa.norm_k.weight,
                                     self.eps, pk["cos"], pk["sin"], 1, L, D, (F_, Hh, Ww), 0,
                                     q_out=ws["q"][r0:r0 + L].view(1, L, C), k_out=ws["k"][b:b + 1, hf * L:(hf + 1) * L],
                                     v_out=ws["v"][b:b + 1, hf * L:(hf + 1) * L])
            for q_lo, q_hi, kv_lo, kv_hi in windows:
                if kv_lo < 0:       # teacher forcing: stage this noisy chunk's K / V over its clean twin's rows
                    dst = -kv_lo - 1
                    ops.kv_roll([ws["k"], ws["v"]], kv_table, dst, L + dst, q_hi - q_lo)
                    kv_lo = 0
                ops.attention(Q4[:, q_lo:q_hi], ws["k"][:, kv_lo:kv_hi], ws["v"][:, kv_lo:kv_hi], A4[:, q_lo:q_hi], scale)
            ops.gemm(ws["attn"], sa.o.weight, sa.o.bias, ws["x"], epilogue=EPI_GATE_RES, residual=ws["x"],
                     gate=m[:, 2], gate_stride=mstride, rows_per_gate=fs)
            # cross attention: text K / V recomputed per call (no cache on this path, model.py:175-180)
            ops.gemm(ws["ctx"], pb["wkv_c"], pb["bkv_c"], None, seg_cols=C, outs=[ws["ctx_k"], ws["cv"].view(B * self.text_len, C)])
            ops.rmsnorm(ws["ctx_k"], ws["ck"].view(B * self.text_len, C), ca.norm_k.weight, self.eps)
            ops.ln_affine(ws["x"], ws["h"], blk.norm3.weight, blk.norm3.bias, self.eps)
            ops.gemm(ws["h"], ca.q.weight, ca.q.bias, ws["q_lin"])
            ops.rmsnorm(ws["q_lin"], ws["q"], ca.norm_q.weight, self.eps)
            ops.attention(Q4, ws["ck"], ws["cv"], A4, scale)
            ops.gemm(ws["attn"], ca.o.weight, ca.o.bias, ws["x"], epilogue=EPI_RESIDUAL, residual=ws["x"])
            ops.ln_modulate(ws["x"], ws["h"], shift=m[:, 3], scale=m[:, 4], mod_stride=mstride, rows_per_mod=fs, eps=self.eps)
            ops.gemm(ws["h"], blk.ffn[0].weight, blk.ffn[0].bias, ws["ffn"], epilogue=EPI_GELU)
            ops.gemm(ws["ffn"], blk.ffn[2].weight, blk.ffn[2].bias, ws["x"], epilogue=EPI_GATE_RES, residual=ws["x"],
                     gate=m[:, 5], gate_stride=mstride, rows_per_gate=fs)

        # ---- head on the noisy half (causal_model.py:1058-1062), modulated by the un-projected embedding of t -------
        hm = ws["head_mod"][0]
        for b in range(B):
            r0 = b * Lt + (halves - 1) * L
            ops.ln_modulate(ws["x"][r0:r0 + L], ws["h"][r0:r0 + L], shift=hm[:, 0], scale=hm[:, 1], mod_stride=2 * C,
                            rows_per_mod=fs, eps=self.eps, row_offset=r0)
            ops.gemm(ws["h"][r0:r0 + L], self.head.head.weight, self.head.head.bias, ws["head_out"][b * L:(b + 1) * L])
        flow = torch.empty(B, F_, self.out_dim, H, W, dtype=bf, device=dev)
        if return_x0:
            if self._sampler_tables is None:
                raise RuntimeError("return_x0 needs set_sampler_tables() (done by B200DiffusionWrapper)")
            x0 = torch.empty_like(flow)
            ops.head_finish(ws["head_out"], x.permute(0, 2, 1, 3, 4), t.contiguous(), self._sampler_tables[0],
                            self._sampler_tables[1], flow, x0)
            return flow.permute(0, 2, 1, 3, 4), x0
        ops.head_finish(ws["head_out"], x.permute(0, 2, 1, 3, 4), t.contiguous(), None, None, flow, None)
        return flow.permute(0, 2, 1, 3, 4)

    @torch.no_grad()
    def _forward_inference(self, x, t, context, seq_len, clip_fea=None, y=None, add_condition=None,
                           kv_cache: List[dict] = None, crossattn_cache: List[dict] = None,
                           current_start: int = 0, cache_start: Optional[int] = None, return_x0: bool = False,
                           skip_output: bool = False):
        """x [B, 16, F, H, W] (or a list of [16, F, H, W]), t [B, F], context [B, <=512, 4096] (or list)
        -> flow [B, 16, F, H, W]  (and x0 [B, F, 16, H, W] when return_x0)."""
        if add_condition is not None or clip_fea is not None or y is not None:
            raise NotImplementedError("pose / image conditioning is not part of the t2v rollout hot path")
        if isinstance(x, (list, tuple)):
            x = torch.stack(list(x))
        if isinstance(context, (list, tuple)):
            context = torch.stack([torch.cat([u, u.new_zeros(self.text_len - u.size(0), u.size(1))]) for u in context])
        ops = self.ops
        pk = self._packed or self._pack()
        B, Cin, F_, H, W = x.shape
        Hh, Ww = H // 2, W // 2
        fs = Hh * Ww
        L = F_ * fs
        assert L <= seq_len
        # one timestep per group of frames: [B, F] per-frame (causal rollout) ... [B, 1] per sample (bidirectional teacher)
        assert Cin == self.in_dim and t.dim() == 2 and t.shape[0] == B and F_ % t.shape[1] == 0, \
            f"timestep shape {tuple(t.shape)} does not divide the {F_} frames of a batch of {B}"
        Fm = t.shape[1]
        mod_rows = (F_ // Fm) * fs        # token rows that share one adaLN modulation vector
        C, NL, D, NH = self.dim, self.num_layers, self.head_dim, self.num_heads
        dev = x.device
        sp = self._sp
        if sp is not None:
            # Ulysses: this rank owns token rows [off, off + Lr) of the chunk for everything but self-attention
            if B != 1:
                raise ValueError("Ulysses mode runs one video (batch 1) per group")
            off, Lr = shard_rows(L, sp.world, sp.rank)
            NHg = NH // sp.world
            q_peer, attn_peer = self._sp_buffers(L, self.patch_embedding.weight.dtype)
        else:
            off, Lr, NHg = 0, L, NH
        ws = self._workspace(B, Lr, Fm, dev, L)
        R = B * Lr

        # ---- host prelude: what this call has to do (no device work yet) ----------------------
        need_ctx = any(not c["is_init"] for c in crossattn_cache[:NL])
        frame_tokens = fs
        start_frame = current_start // frame_tokens
        if start_frame + F_ > 1024:      # the reference's RoPE table has 1024 positions per axis (causal_model.py:483-488)
            raise ValueError(f"frame index {start_frame + F_ - 1} beyond the 1024-entry RoPE table")
        sink_tokens = self.sink_size * frame_tokens
        idx = self._mirror.read(kv_cache[:NL])   # KV-cache plans (host integers; ref causal_model.py:195-236)
        plans = [plan_cache_update(g, l, current_start, L, kv_cache[i]["k"].shape[1], self.local_attn_size,
                                   sink_tokens, self.max_attention_size) for i, (g, l) in enumerate(idx)]
        if return_x0 and self._sampler_tables is None:
            raise RuntimeError("return_x0 needs set_sampler_tables() (done by B200DiffusionWrapper)")

        # cross-attention norms folded into the neighbouring kernels: per layer, if the folded K of this cache exists (or is
        # about to be written by this call)
        fold_ok = pk["fold"] and R > 128
        for i in range(NL):       # entries this call will fill: right shape, and a buffer for the folded K (allocations stay
            cc = crossattn_cache[i]    # out of the device work so that it can be captured as a CUDA graph)
            if cc["is_init"]:
                continue
            ck = cc["k"]
            if ck.shape != (B, self.text_len, NH, D) or not ck.is_contiguous() or not cc["v"].is_contiguous():
                ck = torch.empty(B, self.text_len, NH, D, dtype=self.patch_embedding.weight.dtype, device=dev)
                cc["k"], cc["v"] = ck, torch.empty_like(ck)
            if fold_ok and (i, ck.data_ptr()) not in self._ck_fold:
                if len(self._ck_fold) >= 4 * NL:      # callers that keep re-allocating their caches
                    self._ck_fold.clear()
                    self._graphs.clear()
                self._ck_fold[(i, ck.data_ptr())] = torch.empty_like(ck)
        fold = tuple(fold_ok and (i, crossattn_cache[i]["k"].data_ptr()) in self._ck_fold for i in range(NL))
        env = dict(fold=fold, ops=ops, pk=pk, ws=ws, sp=sp, dev=dev, mod_rows=mod_rows, Fm=Fm, B=B, F_=F_, H=H, W=W, Hh=Hh, Ww=Ww, fs=fs, L=L, Lr=Lr, off=off, R=R,
                   C=C, NL=NL, D=D, NH=NH, NHg=NHg, kv_cache=kv_cache, crossattn_cache=crossattn_cache, plans=plans,
                   need_ctx=need_ctx, start_frame=start_frame, current_start=current_start, skip_output=skip_output,
                   return_x0=return_x0)
        if sp is not None:
            env.update(q_peer=q_peer, attn_peer=attn_peer)
        result = self._run_device_work(x, t, context, env)
        self._mirror.write(kv_cache[:NL], [(p.global_end, p.local_end) for p in plans])
        if skip_output:
            return None
        flow, x0 = result
        out = flow.permute(0, 2, 1, 3, 4)   # reference returns [B, C, F, H, W]
        return (out, x0) if return_x0 else out

    # ------------------------------------------------------------------ CUDA-graph dispatch
    def _run_device_work(self, x, t, context, env):
        """Runs (or replays) all kernels of one forward.  `env` = the prelude's local variables."""
        ops, sp, dev = env["ops"], env["sp"], env["dev"]
        eligible = (self.use_cuda_graphs and dev.type == "cuda"
                    and getattr(ops, "supports_cuda_graphs", False) and getattr(ops, "_prof", None) is None)
        if not eligible:
            return self._device_forward(x, t, context, env)
        kv, ca, NL = env["kv_cache"], env["crossattn_cache"], env["NL"]
        # Static signature of the device work.  The chunk position enters the kernels only through the RoPE frame offset,
        # which is passed in device memory under replay, and the cache plan only through its device-visible fields -- so
        # in the steady state of a rolling-window video (same roll, same write slot, same window every chunk) ONE graph
        # per forward kind serves every chunk.
        # The first forward of a prompt also fills the cross-attention cache from the text context: its own signature,
        # with the context as a third static input.
        need_ctx = env["need_ctx"]
        fills = tuple(not c["is_init"] for c in ca[:NL]) if need_ctx else None
        key = (tuple(x.shape), tuple(t.shape), t.dtype, env["return_x0"], env["skip_output"], env["fold"],
               fills, (tuple(context.shape), context.dtype) if need_ctx else None,
               tuple((p.roll, p.roll_src, p.roll_dst, p.roll_len, p.write_start, p.write_end, p.attn_start, p.attn_end)
                     for p in env["plans"]),
               tuple(c["k"].data_ptr() for c in kv[:NL]), tuple(c["v"].data_ptr() for c in kv[:NL]),
               tuple(c["k"].data_ptr() for c in ca[:NL]), tuple(c["v"].data_ptr() for c in ca[:NL]))
        ent = self._graphs.get(key)
        if ent is None:                       # first occurrence: run eagerly (also warms every lazy initialisation)
            # bounded: each captured graph owns ~430 nodes plus a private pool -- evict the least recently used
            while len(self._graphs) >= self.max_cuda_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = "seen"
            return self._device_forward(x, t, context, env)
        self._graphs[key] = self._graphs.pop(key)   # mark as most recently used (dicts keep insertion order)
        if sp is not None:
            env = dict(env, defer_gather=True)   # keep the NCCL collective out of the captured graph
        if ent == "seen":                     # second occurrence: capture
            xs = torch.empty(x.shape, dtype=x.dtype, device=dev)
            ts = torch.empty(t.shape, dtype=t.dtype, device=dev)
            cs = torch.empty_like(context) if need_ctx else None
            sf = torch.zeros(1, dtype=torch.int32, device=dev) if sp is None else None
            graph = torch.cuda.CUDAGraph()
            before = ops.launches
            with torch.cuda.graph(graph):
                outs = self._device_forward(xs, ts, cs if need_ctx else context, dict(env, start_frame_dev=sf, replay_sets_flags=True))
            ent = (graph, xs, ts, outs, ops.launches - before, sf, env["start_frame"], cs)
            ops.launches = before
            self._graphs[key] = ent
        graph, xs, ts, outs, n_launches, sf, sf_captured, cs = ent
        if sf is not None:
            sf.fill_(env["start_frame"])
        elif sf_captured != env["start_frame"]:   # Ulysses: the frame offset is baked into the captured kernels
            self._graphs[key] = "seen"
            return self._device_forward(x, t, context, env)
        xs.copy_(x)
        ts.copy_(t)
        if cs is not None:
            cs.copy_(context)
        graph.replay()
        ops.launches += n_launches
        if need_ctx:
            for c in ca[:NL]:
                c["is_init"] = True
        if outs is None:
            return None
        if outs == "gather":
            return self._head_tail(xs, ts, env)
        return tuple(o.clone() if o is not None else None for o in outs)

    def _device_forward(self, x, t, context, env):
        """Every kernel of one cached forward (static given `env`): capturable as one CUDA graph."""
        ops, pk, ws, sp = env["ops"], env["pk"], env["ws"], env["sp"]
        B, F_, H, W, Hh, Ww, fs, L, Lr, off, R = (env[k] for k in ("B", "F_", "H", "W", "Hh", "Ww", "fs", "L", "Lr", "off", "R"))
        C, NL, D, NH, NHg, dev = (env[k] for k in ("C", "NL", "D", "NH", "NHg", "dev"))
        kv_cache, crossattn_cache, plans = env["kv_cache"], env["crossattn_cache"], env["plans"]
        need_ctx, start_frame, skip_output, return_x0 = env["need_ctx"], env["start_frame"], env["skip_output"], env["return_x0"]
        mod_rows, fold = env["mod_rows"], env["fold"]
        if sp is not None:
            q_peer, attn_peer = env["q_peer"], env["attn_peer"]

        # ---- embeddings -------------------------------------------------------------------
        ops.patchify(x, ws["tok"])
        ops.gemm(ws["tok"][off:off + Lr] if sp is not None else ws["tok"], pk["patch_w"], self.patch_embedding.bias,
                 ws["x"])
        tflat = t.reshape(-1).contiguous()
        ops.sinusoid(tflat, ws["sin"], self.freq_dim)
        te, tp = self.time_embedding, self.time_projection
        ops.skinny_linear(ws["sin"], te[0].weight, te[0].bias, ws["e1"], silu_in=False)
        ops.skinny_linear(ws["e1"], te[2].weight, te[2].bias, ws["e"], silu_in=True)
        ops.skinny_linear(ws["e"], tp[1].weight, tp[1].bias, ws["e0"], silu_in=True)
        ops.modulation_table(pk["mod"], ws["e0"], ws["mod"], e_row_stride=6 * C, e_group_stride=C)
        ops.modulation_table(pk["head_mod"], ws["e"], ws["head_mod"], e_row_stride=C, e_group_stride=0)

        # ---- text context: only consumed to fill the cross-attention cache ------------------
        if need_ctx:
            if context.shape[1] < self.text_len:
                context = torch.cat([context, context.new_zeros(B, self.text_len - context.shape[1],
                                                                context.shape[2])], dim=1)
            ctx_in = context.reshape(B * self.text_len, self.text_dim)
            if not ctx_in.is_contiguous():
                ctx_in = ctx_in.contiguous()
            tx = self.text_embedding
            ops.gemm(ctx_in, tx[0].weight, tx[0].bias, ws["ctx_h"], epilogue=EPI_GELU)
            ops.gemm(ws["ctx_h"], tx[2].weight, tx[2].bias, ws["ctx"])

        scale = 1.0 / math.sqrt(D)
        mod = ws["mod"]
        mstride = 6 * C   # elements between consecutive (b, f) rows of one layer's table

        # ---- rolling window: evict the oldest chunk of every layer's cache before anything is appended ----------
        # (reference causal_model.py:212-221 shifts inside each layer's attention call; nothing in between reads the
        # caches, so all layers move in ONE kernel per phase)
        rolls: Dict[tuple, List[int]] = {}
        for i in range(NL):
            if plans[i].roll and plans[i].roll_len > 0:
                kc = kv_cache[i]["k"]
                rolls.setdefault((plans[i].roll_dst, plans[i].roll_src, plans[i].roll_len, tuple(kc.shape), tuple(kc.stride())), []).append(i)
        for (dst, src, n, _, _), layers in rolls.items():
            tensors = [kv_cache[i][name] for i in layers for name in ("k", "v")]
            ops.kv_roll(tensors, self._pointer_table(tensors, dev), dst, src, n)

        # Row statistics instead of reduction passes: every GEMM that writes the residual stream x
        # also writes its records, and the adaLN LayerNorms stream their rows (SFB_NO_STREAM_LN=1: A/B knob) ...
        stream_ln = pk["fold"] and R > 128 and os.environ.get("SFB_NO_STREAM_LN", "0") != "1"
        x_stats_out = dict(stats_out=ws["x_stats"]) if stream_ln else {}
        x_stats_in = dict(stats=ws["x_stats"]) if stream_ln else {}
        # ... and the QKV projection writes the q / k records that let qk_norm_rope stream as well
        stream_rope = pk["fold"] and R > 128 and D == 128 and sp is None
        qkv_stats = dict(stats_out=ws["qkv_stats"]) if stream_rope else {}
        rope_stats = dict(stats=ws["qkv_stats"], q_chunk0=0, k_chunk0=C // STATS_CHUNK) if stream_rope else {}
        for i, blk in enumerate(self.blocks):
            pb = pk["blocks"][i]
            cache, plan = kv_cache[i], plans[i]
            m = mod[i]   # [B*F, 6, C]
            sa, ca = blk.self_attn, blk.cross_attn
            # -- self attention --
            # (x_stats always describes the current x from here on: every GEMM that writes x writes its records too;
            # the patch embedding in front of layer 0 is a one-CTA GEMM without statistics)
            ops.ln_modulate(ws["x"], ws["h"], shift=m[:, 0], scale=m[:, 1], mod_stride=mstride, rows_per_mod=mod_rows,
                            eps=self.eps, row_offset=off, **(x_stats_in if i > 0 else {}))
            kc, vc = cache["k"], cache["v"]
            if kc.shape[0] != B or kc.shape[2] != NHg or kc.shape[3] != D:
                raise ValueError(f"kv_cache[{i}]['k'] shape {tuple(kc.shape)} does not match B={B}, H={NHg}, D={D}")
            k_slot = kc[:, plan.write_start:plan.write_end]
            v_slot = vc[:, plan.write_start:plan.write_end]
            if sp is not None:
                # head-parallel self-attention: the all-to-alls are peer stores inside the two kernels below
                ops.gemm(ws["h"], pb["wqkv"], pb["bqkv"], None, seg_cols=C,
                         outs=[ws["q_lin"], ws["k_lin"], ws["v_lin"]])
                try:
                    (k_pool, k_base), (v_pool, v_base) = self._sp_kv[kc.data_ptr()], self._sp_kv[vc.data_ptr()]
                except KeyError:
                    raise ValueError("Ulysses mode needs KV caches from model.allocate_kv_cache()") from None
                slot_off = plan.write_start * NHg * D
                ops.qk_norm_rope_sp(ws["q_lin"], ws["k_lin"], ws["v_lin"], sa.norm_q.weight, sa.norm_k.weight, self.eps,
                                    pk["cos"], pk["sin"], D, (F_, Hh, Ww), start_frame, off, sp, q_peer,
                                    k_slot[0], v_slot[0], k_pool.ptrs_at(k_base + slot_off),
                                    v_pool.ptrs_at(v_base + slot_off))
                sp.barrier(ops)          # every rank's q / k / v rows have landed
                if skip_output and i == NL - 1:
                    break
                ops.attention_sp(q_peer.local.view(L, NHg, D), kc[0, plan.attn_start:plan.attn_end],
                                 vc[0, plan.attn_start:plan.attn_end], scale, sp, attn_peer, Lr)
                sp.barrier(ops)          # every head group's output columns have landed
                ops.gemm(attn_peer.local, sa.o.weight, sa.o.bias, ws["x"], epilogue=EPI_GATE_RES, residual=ws["x"],
                         gate=m[:, 2], gate_stride=mstride, rows_per_gate=mod_rows, gate_row_offset=off,
                         **(x_stats_out if (fold[i] or stream_ln) else {}))
            elif B == 1:   # V projection lands directly in its cache slot
                ops.gemm(ws["h"], pb["wqkv"], pb["bqkv"], None, seg_cols=C,
                         outs=[ws["q_lin"], ws["k_lin"], v_slot.view(L, C)], **qkv_stats)   # .view: a non-viewable cache layout must raise
                v_src = None
            else:
                ops.gemm(ws["h"], pb["wqkv"], pb["bqkv"], None, seg_cols=C,
                         outs=[ws["q_lin"], ws["k_lin"], ws["v_lin"]], **qkv_stats)
                v_src = ws["v_lin"]
            q4 = ws["q"].view(B, Lr, NH, D)
            if sp is None:
                ops.qk_norm_rope(ws["q_lin"], ws["k_lin"], v_src, sa.norm_q.weight, sa.norm_k.weight, self.eps,
                                 pk["cos"], pk["sin"], B, L, D, (F_, Hh, Ww), start_frame,
                                 q_out=ws["q"].view(B, L, C), k_out=k_slot, v_out=v_slot,
                                 start_frame_dev=env.get("start_frame_dev"), **rope_stats)
                if skip_output and i == NL - 1:
                    break   # cache-refresh pass: nothing after the last layer's K/V append is consumed
                ops.attention(q4, kc[:, plan.attn_start:plan.attn_end], vc[:, plan.attn_start:plan.attn_end],
                              ws["attn"].view(B, L, NH, D), scale)
                ops.gemm(ws["attn"], sa.o.weight, sa.o.bias, ws["x"], epilogue=EPI_GATE_RES, residual=ws["x"],
                         gate=m[:, 2], gate_stride=mstride, rows_per_gate=mod_rows,
                         **(x_stats_out if (fold[i] or stream_ln) else {}))
            # -- cross attention --
            cc = crossattn_cache[i]
            if not cc["is_init"]:
                ck, cv = cc["k"], cc["v"]       # shape-checked (and the folded-K buffer allocated) by the host prelude
                ops.gemm(ws["ctx"], pb["wkv_c"], pb["bkv_c"], None, seg_cols=C,
                         outs=[ws["ctx_k"], cv.view(B * self.text_len, C)])
                ops.rmsnorm(ws["ctx_k"], ck.view(B * self.text_len, C), ca.norm_k.weight, self.eps)
                if fold[i]:    # K with norm_q's weight multiplied in, kept beside the reference-visible cache entry
                    ops.rmsnorm(ws["ctx_k"], self._ck_fold[(i, ck.data_ptr())].view(B * self.text_len, C), pb["gk_fold"], self.eps)
                if not env.get("replay_sets_flags"):
                    cc["is_init"] = True
            if fold[i]:
                # norm3 -> q -> norm_q -> attention as TWO launches: the q projection reads x directly (LayerNorm applied in
                # its epilogue from the row statistics the o projection just wrote) and emits the statistics of q; the
                # attention applies norm_q's row factor inside the softmax
                ops.gemm(ws["x"], pb["wq_c_fold"], None, ws["q_lin"], stats_out=ws["q_stats"], ln_stats=ws["x_stats"],
                         ln_sc=pb["sc_c"], ln_eps=self.eps)
                ops.attention(ws["q_lin"].view(B, Lr, NH, D), self._ck_fold[(i, cc["k"].data_ptr())], cc["v"],
                              ws["attn"].view(B, Lr, NH, D), scale, q_stats=ws["q_stats"], q_eps=self.eps)
            else:
                ops.ln_affine(ws["x"], ws["h"], blk.norm3.weight, blk.norm3.bias, self.eps)
                ops.gemm(ws["h"], ca.q.weight, ca.q.bias, ws["q_lin"])
                ops.rmsnorm(ws["q_lin"], ws["q"], ca.norm_q.weight, self.eps)
                ops.attention(q4, cc["k"], cc["v"], ws["attn"].view(B, Lr, NH, D), scale)
            ops.gemm(ws["attn"], ca.o.weight, ca.o.bias, ws["x"], epilogue=EPI_RESIDUAL, residual=ws["x"], **x_stats_out)
            # -- feed forward --
            ops.ln_modulate(ws["x"], ws["h"], shift=m[:, 3], scale=m[:, 4], mod_stride=mstride, rows_per_mod=mod_rows,
                            eps=self.eps, row_offset=off, **x_stats_in)
            ops.gemm(ws["h"], blk.ffn[0].weight, blk.ffn[0].bias, ws["ffn"], epilogue=EPI_GELU)
            ops.gemm(ws["ffn"], blk.ffn[2].weight, blk.ffn[2].bias, ws["x"], epilogue=EPI_GATE_RES,
                     residual=ws["x"], gate=m[:, 5], gate_stride=mstride, rows_per_gate=mod_rows, gate_row_offset=off,
                     **x_stats_out)

        if skip_output:
            return None

        # ---- head ------------------------------------------------------------------------------
        hm = ws["head_mod"][0]   # [B*F, 2, C]
        ops.ln_modulate(ws["x"], ws["h"], shift=hm[:, 0], scale=hm[:, 1], mod_stride=2 * C, rows_per_mod=mod_rows,
                        eps=self.eps, row_offset=off, **(x_stats_in if NL > 0 else {}))
        ops.gemm(ws["h"], self.head.head.weight, self.head.head.bias, ws["head_out"])
        if env.get("defer_gather"):
            return "gather"          # Ulysses under graph capture: the NCCL gather + unpatchify run outside the graph
        return self._head_tail(x, t, env)

    def _head_tail(self, x, t, env):
        """[Ulysses: all-gather the token shards of the head output (xdit_context_parallel.py:142)] -> unpatchify
        (+ flow -> x0)."""
        ops, ws, sp, dev = env["ops"], env["ws"], env["sp"], env["dev"]
        B, F_, H, W = env["B"], env["F_"], env["H"], env["W"]
        head_out = ws["head_out"]
        if sp is not None:
            sp.all_gather_rows(ws["head_out"], ws["head_full"])
            head_out = ws["head_full"]
        flow = torch.empty(B, F_, self.out_dim, H, W, dtype=ws["x"].dtype, device=dev)
        x0 = None
        xt = x.permute(0, 2, 1, 3, 4)   # [B, F, C, H, W] view of the input
        if env["Fm"] != F_:             # the flow -> x0 kernel wants one timestep per frame
            t = t.repeat_interleave(F_ // env["Fm"], dim=1)
        if env["return_x0"]:
            x0 = torch.empty_like(flow)
            ops.head_finish(head_out, xt, t.contiguous(), self._sampler_tables[0], self._sampler_tables[1],
                            flow, x0)
        else:
            ops.head_finish(head_out, xt, t.contiguous(), None, None, flow, None)
        return flow, x0


class B200WanModel(B200CausalWanModel):
    """Drop-in for the bidirectional `WanModel` forward (wan/modules/model.py:497-771, t2v, no GAN/classify branch):
    the teacher / ODE-pair generator of BASELINE config 5.

    A bidirectional forward is the cached causal forward with ONE timestep per sample and an empty cache: every
    token attends to all tokens written by this very call, RoPE starts at frame 0, adaLN vectors are per sample
    (`e [B, 6, C]`, model.py:697-700) and the text K/V are recomputed per call.  So this class only supplies scratch
    K/V (one buffer shared by all layers -- nothing persists) and forwards to the same kernel schedule; under
    `enable_ulysses` it is the head-parallel forward of wan/distributed/xdit_context_parallel.py:66-192.

    Samples shorter than `seq_len`: the reference zero-pads every sample to seq_len tokens and hands k_lens = seq_lens
    to flash_attn (model.py:684-693,150-156), so the valid tokens of a sample only ever see that sample's own valid
    tokens and every other op is row-wise -- the padded rows change nothing that is returned (unpatchify drops them,
    :745-771).  The B200 forward therefore runs the exact-length problem: no padded rows are computed at all.  Samples
    of different shapes in one call run one after the other (their token grids, hence RoPE, differ)."""

    def __init__(self, *args, **kwargs):
        kwargs.pop("local_attn_size", None)
        kwargs.pop("sink_size", None)
        super().__init__(*args, **kwargs)
        self._scratch: Dict[tuple, tuple] = {}

    def _apply(self, fn, *a, **k):
        self._scratch.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, return_x0: bool = False, **unsupported):
        if any(v is not None and v is not False for v in unsupported.values()):
            raise NotImplementedError(f"B200WanModel: unsupported arguments {sorted(unsupported)} (t2v forward only)")
        if isinstance(x, (list, tuple)) and len({tuple(u.shape) for u in x}) > 1:
            # mixed shapes: one exact-length forward per sample; the reference stacks the outputs (model.py:771), which
            # only works if they happen to share a shape -- same behaviour here
            ctx = context if isinstance(context, (list, tuple)) else list(context)
            outs = [self.forward([u], t[i:i + 1], [ctx[i]], seq_len, return_x0=return_x0) for i, u in enumerate(x)]
            if return_x0:
                return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
            return torch.cat(outs)
        if isinstance(x, (list, tuple)):
            x = torch.stack(list(x))
        B, _, F_, H, W = x.shape
        L = F_ * (H // 2) * (W // 2)
        if L > seq_len:
            raise ValueError(f"sample of {L} tokens exceeds seq_len {seq_len}")      # reference: assert seq_lens.max() <= seq_len
        if self._sp is not None and L % self._sp.world:
            raise NotImplementedError(f"Ulysses mode needs the {L} tokens of a sample to split evenly over {self._sp.world} ranks")
        dev, dt = x.device, self.patch_embedding.weight.dtype
        key = (B, L, str(dev))
        if key not in self._scratch:
            kv = self.allocate_kv_cache(B, L, dt, dev, layers=1)[0]
            ck = torch.zeros(B, self.text_len, self.num_heads, self.head_dim, dtype=dt, device=dev)
            self._scratch[key] = (kv["k"], kv["v"], ck, torch.zeros_like(ck))
        k, v, ck, cv = self._scratch[key]
        zero = torch.zeros(1, dtype=torch.long, device=dev)
        kv_cache = [dict(k=k, v=v, global_end_index=zero.clone(), local_end_index=zero.clone())
                    for _ in range(self.num_layers)]
        ca_cache = [dict(k=ck, v=cv, is_init=False) for _ in range(self.num_layers)]
        if y is not None or clip_fea is not None:
            raise NotImplementedError("image conditioning (i2v) is not supported")
        return self._forward_inference(x, t.reshape(B, 1), context, seq_len, kv_cache=kv_cache, crossattn_cache=ca_cache,
                                       current_start=0, return_x0=return_x0)
