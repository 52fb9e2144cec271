"""GPU parity checks of every kernel behind the C ABI, shared by the pytest suite (`-m gpu`) and by
`tools/gpu_report.py` (which runs each check in its own process under a timeout so one hanging
kernel cannot hide the others' results).

Each check builds seeded inputs, calls the kernel through `self_forcing_b200.ops.CudaOps` (ctypes ->
libsfb200.so) and compares with a plain PyTorch fp32 restatement of the same op that rounds to bf16
at the reference's rounding points.  Returns a dict of error metrics and raises AssertionError on a
parity failure.  Tolerances: bf16 outputs may differ by one rounding (rel-L2 <= 3e-3); integer /
copy outputs must be bit-exact.
"""
from __future__ import annotations

import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from _torch_ops import TorchOps  # noqa: E402
from oracle import causal_wan_oracle as O  # noqa: E402

BF = torch.bfloat16


def _ops():
    from self_forcing_b200.ops import CudaOps
    return CudaOps()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _randn(*shape, seed=0, scale=1.0, dtype=BF):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


def _finish(name, metrics, tol):
    torch.cuda.synchronize()
    bad = {k: v for k, v in metrics.items() if k.startswith("err") and not (v <= tol)}
    metrics["tol"] = tol
    assert not bad, f"{name}: parity failure {bad} (tol {tol}); all metrics {metrics}"
    return metrics


# --------------------------------------------------------------------------------------
def check_gemm(M=300, N=256, K=192, epilogue=0, block_n=0, rows_per_gate=100, seed=0, gate_row_offset=0):
    ops = _ops()
    x = _randn(M, K, seed=seed)
    w = _randn(N, K, seed=seed + 1, scale=1.0 / math.sqrt(K))
    b = _randn(N, seed=seed + 2, scale=0.5)
    res = _randn(M, N, seed=seed + 3)
    groups = (M + gate_row_offset + rows_per_gate - 1) // rows_per_gate
    gate_tab = _randn(groups, 3, N, seed=seed + 4)        # strided gate rows like the modulation table
    gate = gate_tab[:, 1]
    out = torch.full((M, N), float("nan"), device="cuda", dtype=BF)
    kw = {}
    if epilogue in (2, 3):
        kw["residual"] = res
    if epilogue == 3:
        kw.update(gate=gate, gate_stride=gate_tab.stride(0), rows_per_gate=rows_per_gate, gate_row_offset=gate_row_offset)
    ops.gemm(x, w, b, out, epilogue=epilogue, block_n=block_n, **kw)
    acc = x.float() @ w.float().t() + b.float()
    y = acc.to(BF)
    if epilogue == 1:
        ref = F.gelu(y.float(), approximate="tanh").to(BF)
    elif epilogue == 2:
        ref = (res.float() + y.float()).to(BF)
    elif epilogue == 3:
        g = gate[(torch.arange(M, device="cuda") + gate_row_offset) // rows_per_gate]
        ref = (res.float() + (y.float() * g.float()).to(BF).float()).to(BF)
    else:
        ref = y
    m = dict(err_rel_l2=rel_l2(out, ref), max_abs=float((out.float() - ref.float()).abs().max()),
             nan=int(torch.isnan(out.float()).sum()))
    assert m["nan"] == 0, f"gemm produced NaN / left output unwritten: {m}"
    return _finish(f"gemm M{M} N{N} K{K} epi{epilogue} bn{block_n}", m, 3e-3)


def _stats_ref(y):
    """(mean, M2) per 128-column chunk of every row (include/sfb200.h: statistics records)."""
    c = y.float().reshape(y.shape[0], -1, 128)
    mean = c.mean(dim=2)
    return torch.stack([mean, ((c - mean[..., None]) ** 2).sum(dim=2)], dim=2)


def check_gemm_stats(M=700, N=512, K=256, epilogue=3, rows_per_gate=130, seed=0, mean_shift=0.0, block_n=0):
    """Row statistics of the OUTPUT written by the pair kernel's epilogue (any epilogue); a large common offset in the
    rows (mean_shift) must not cost accuracy in M2 (shifted sums + Chan's merge)."""
    ops = _ops()
    x = _randn(M, K, seed=seed)
    w = _randn(N, K, seed=seed + 1, scale=1.0 / math.sqrt(K))
    b = _randn(N, seed=seed + 2, scale=0.5)
    res = (_randn(M, N, seed=seed + 3).float() + mean_shift).to(BF)
    groups = (M + rows_per_gate - 1) // rows_per_gate
    gate = _randn(groups, N, seed=seed + 4)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=BF)
    stats = torch.full((M, N // 128, 2), float("nan"), device="cuda")
    kw = dict(residual=res) if epilogue in (2, 3) else {}
    if epilogue == 3:
        kw.update(gate=gate, gate_stride=N, rows_per_gate=rows_per_gate)
    ops.gemm(x, w, b, out, epilogue=epilogue, stats_out=stats, block_n=block_n, **kw)
    ref = _stats_ref(out)                       # statistics of the bf16 values the kernel stored
    var = ref[..., 1].sum(1) / N
    m = dict(err_mean=float((stats[..., 0] - ref[..., 0]).abs().max() / (1.0 + abs(mean_shift))),
             err_m2=float(((stats[..., 1] - ref[..., 1]).abs() / (ref[..., 1] + 1e-3)).max()),
             nan=int(torch.isnan(stats).sum() + torch.isnan(out.float()).sum()), var_mean=float(var.mean()))
    assert m["nan"] == 0, f"unwritten statistics / output: {m}"
    return _finish(f"gemm_stats M{M} N{N} K{K} epi{epilogue} shift{mean_shift}", m, 2e-3)


def check_gemm_lnfold(M=700, N=512, K=256, seed=0, mean_shift=0.0):
    """norm3 folded into the cross-attention q projection: producer GEMM writes x and its statistics, consumer GEMM applies
    the LayerNorm in its epilogue and writes the statistics of q.  Reference: fp32 LayerNorm (affine) -> Linear."""
    ops = _ops()
    eps = 1e-6
    a = _randn(M, K, seed=seed)
    wp = _randn(K, K, seed=seed + 1, scale=1.0 / math.sqrt(K))
    res = (_randn(M, K, seed=seed + 2).float() + mean_shift).to(BF)
    x = torch.empty(M, K, device="cuda", dtype=BF)
    x_stats = torch.full((M, K // 128, 2), float("nan"), device="cuda")
    ops.gemm(a, wp, None, x, epilogue=2, residual=res, stats_out=x_stats)            # x = res + a @ wp^T, with statistics
    wq = _randn(N, K, seed=seed + 3, scale=1.0 / math.sqrt(K))
    bq = _randn(N, seed=seed + 4, scale=0.5)
    w3 = (1.0 + 0.1 * _randn(K, seed=seed + 5).float()).to(BF)
    b3 = _randn(K, seed=seed + 6, scale=0.1)
    wf = (wq.float() * w3.float()[None, :]).to(BF)
    sc = torch.stack([wf.float().sum(1), bq.float() + wq.float() @ b3.float()], dim=1).contiguous()
    q = torch.full((M, N), float("nan"), device="cuda", dtype=BF)
    q_stats = torch.full((M, N // 128, 2), float("nan"), device="cuda")
    ops.gemm(x, wf, None, q, stats_out=q_stats, ln_stats=x_stats, ln_sc=sc, ln_eps=eps)
    h = F.layer_norm(x.float(), (K,), w3.float(), b3.float(), eps)
    ref = h @ wq.float().t() + bq.float()                                              # fp32 reference, no intermediate rounding
    ref_rounded = (h.to(BF).float() @ wq.float().t() + bq.float()).to(BF)              # what the reference's op chain produces
    qs_ref = _stats_ref(q)
    m = dict(err_rel_l2=rel_l2(q, ref), ref_chain_rel_l2=rel_l2(ref_rounded, ref),
             err_qstats=float(((q_stats - qs_ref).abs() / (qs_ref.abs() + 1e-2)).max()) * 1e-1,
             nan=int(torch.isnan(q.float()).sum() + torch.isnan(q_stats).sum()))
    assert m["nan"] == 0, f"unwritten output: {m}"
    # no further from the fp32 result than the reference's own bf16 op chain (x 1.5), and within the GEMM tolerance
    assert m["err_rel_l2"] <= max(1.5 * m["ref_chain_rel_l2"], 1e-3), m
    return _finish(f"gemm_lnfold M{M} N{N} K{K} shift{mean_shift}", m, 4e-3)


def check_attention_qnorm(B=1, Lq=300, S=512, H=2, seed=0):
    """WanRMSNorm of q folded into the softmax scale + norm weight folded into K, against norm -> attention."""
    ops = _ops()
    D, eps = 128, 1e-6
    C = H * D
    q_lin = _randn(B * Lq, C, seed=seed, scale=2.0)
    gq = (1.0 + 0.1 * _randn(C, seed=seed + 1).float()).to(BF)
    k = _randn(B, S, H, D, seed=seed + 2)
    v = _randn(B, S, H, D, seed=seed + 3)
    scale = 1.0 / math.sqrt(D)
    q_stats = _stats_ref(q_lin).contiguous()
    kf = (k.float() * gq.float().view(1, 1, H, D)).to(BF)
    out = torch.full((B, Lq, H, D), float("nan"), device="cuda", dtype=BF)
    ops.attention(q_lin.view(B, Lq, H, D), kf, v, out, scale, q_stats=q_stats, q_eps=eps)
    qn = (q_lin.float() * torch.rsqrt(q_lin.float().pow(2).mean(dim=1, keepdim=True) + eps) * gq.float())
    ref = _attn_ref(qn.view(B, Lq, H, D), k, v, scale)
    m = dict(err_rel_l2=rel_l2(out, ref), nan=int(torch.isnan(out.float()).sum()))
    assert m["nan"] == 0, f"attention produced NaN / left rows unwritten: {m}"
    return _finish(f"attention_qnorm B{B} Lq{Lq} S{S} H{H}", m, 6e-3)



def check_gemm_segments(M=300, C=256, K=128, seed=0, block_n=0):
    """QKV-style call: one GEMM, three destinations with different row strides (V into a cache slot)."""
    ops = _ops()
    x = _randn(M, K, seed=seed)
    w = _randn(3 * C, K, seed=seed + 1, scale=1.0 / math.sqrt(K))
    b = _randn(3 * C, seed=seed + 2, scale=0.5)
    q = torch.zeros(M, C, device="cuda", dtype=BF)
    k = torch.zeros(M, C + 64, device="cuda", dtype=BF)[:, :C]
    cache = torch.zeros(M + 50, C, device="cuda", dtype=BF)
    ops.gemm(x, w, b, None, outs=[q, k, cache[20:20 + M]], seg_cols=C, block_n=block_n)
    ref = (x.float() @ w.float().t() + b.float()).to(BF)
    m = dict(err_q=rel_l2(q, ref[:, :C]), err_k=rel_l2(k, ref[:, C:2 * C]), err_v=rel_l2(cache[20:20 + M], ref[:, 2 * C:]),
             err_untouched=float(cache[:20].float().abs().max() + cache[20 + M:].float().abs().max()))
    return _finish("gemm segments", m, 3e-3)


def _attn_ref(q, k, v, scale):
    qf, kf, vf = (t.float().transpose(1, 2) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    return (p @ vf).transpose(1, 2)


def check_attention(B=1, Lq=300, S=520, H=2, seed=0, cache_rows=None, window_start=0, fused_q=False):
    ops = _ops()
    D = 128
    cache_rows = cache_rows or (S + window_start + 7)
    if fused_q:   # q lives inside a wider buffer (row stride > H*D)
        qbuf = _randn(B, Lq, H * D + 256, seed=seed)
        q = qbuf[:, :, :H * D].unflatten(2, (H, D))
    else:
        q = _randn(B, Lq, H, D, seed=seed)
    kc = _randn(B, cache_rows, H, D, seed=seed + 1)
    vc = _randn(B, cache_rows, H, D, seed=seed + 2)
    k = kc[:, window_start:window_start + S]
    v = vc[:, window_start:window_start + S]
    out = torch.full((B, Lq, H, D), float("nan"), device="cuda", dtype=BF)
    scale = 1.0 / math.sqrt(D)
    ops.attention(q, k, v, out, scale)
    ref = _attn_ref(q, k, v, scale)
    m = dict(err_rel_l2=rel_l2(out, ref), max_abs=float((out.float() - ref).abs().max()),
             nan=int(torch.isnan(out.float()).sum()))
    assert m["nan"] == 0, f"attention produced NaN / left rows unwritten: {m}"
    # per-head error helps localise descriptor / layout mistakes
    m["per_head"] = [round(rel_l2(out[:, :, h], ref[:, :, h]), 5) for h in range(H)]
    return _finish(f"attention B{B} Lq{Lq} S{S} H{H}", m, 6e-3)


def check_attention_sharp(Lq=256, S=640, H=1, seed=3):
    """Large-magnitude scores (row max grows late) exercise the lazy O rescale."""
    ops = _ops()
    D = 128
    q = _randn(1, Lq, H, D, seed=seed, scale=3.0)
    k = _randn(1, S, H, D, seed=seed + 1, scale=3.0)
    k[:, S - 100:] *= 2.0        # later tiles dominate -> reference max jumps by far more than 2^8
    v = _randn(1, S, H, D, seed=seed + 2)
    out = torch.empty(1, Lq, H, D, device="cuda", dtype=BF)
    ops.attention(q, k, v, out, 1.0 / math.sqrt(D))
    ref = _attn_ref(q, k, v, 1.0 / math.sqrt(D))
    m = dict(err_rel_l2=rel_l2(out, ref), nan=int(torch.isnan(out.float()).sum()))
    assert m["nan"] == 0
    return _finish("attention sharp", m, 8e-3)


# --------------------------------------------------------------------------------------
def _against_double(name, run, outputs, tol=3e-3, exact=()):
    """run(ops, outs) is executed with CudaOps and with the TorchOps double; outputs compared."""
    cu = {k: v.clone() for k, v in outputs.items()}
    rf = {k: v.clone() for k, v in outputs.items()}
    run(_ops(), cu)
    run(TorchOps(), rf)
    torch.cuda.synchronize()
    m = {}
    for k in outputs:
        if k in exact:
            m[f"err_{k}_mismatch"] = float((cu[k] != rf[k]).sum())
        else:
            m[f"err_{k}"] = rel_l2(cu[k], rf[k])
            m[f"max_{k}"] = float((cu[k].float() - rf[k].float()).abs().max())
    bad = {k: v for k, v in m.items() if k.startswith("err") and not (v <= (0 if k.endswith("mismatch") else tol))}
    assert not bad, f"{name}: {bad}; {m}"
    return m


def check_ln_modulate(rows=777, C=1536, rpm=200, seed=0):
    x = _randn(rows, C, seed=seed, scale=2.0) + 0.3
    tab = _randn((rows + rpm - 1) // rpm, 6, C, seed=seed + 1, scale=0.5)

    def run(ops, o):
        ops.ln_modulate(x, o["y"], shift=tab[:, 3], scale=tab[:, 4], mod_stride=6 * C, rows_per_mod=rpm, eps=1e-6)
    return _against_double("ln_modulate", run, dict(y=torch.zeros(rows, C, device="cuda", dtype=BF)))


def check_ln_modulate_stats(rows=777, C=1536, rpm=200, seed=0, mean_shift=0.3, row_offset=0):
    """Streaming form fed by statistics records against the resident-row kernel (only the summation order of the row
    statistics differs) and against the torch double."""
    ops = _ops()
    x = (_randn(rows, C, seed=seed, scale=2.0).float() + mean_shift).to(BF)
    tab = _randn((rows + row_offset + rpm - 1) // rpm, 6, C, seed=seed + 1, scale=0.5)
    stats = _stats_ref(x).contiguous()
    kw = dict(shift=tab[:, 3], scale=tab[:, 4], mod_stride=6 * C, rows_per_mod=rpm, eps=1e-6, row_offset=row_offset)
    y0 = torch.zeros(rows, C, device="cuda", dtype=BF)
    y1 = torch.full((rows, C), float("nan"), device="cuda", dtype=BF)
    ops.ln_modulate(x, y0, **kw)
    ops.ln_modulate(x, y1, stats=stats, **kw)
    yr = torch.zeros(rows, C, device="cuda", dtype=BF)
    TorchOps().ln_modulate(x, yr, **kw)
    m = dict(err_vs_resident=rel_l2(y1, y0), err_vs_double=rel_l2(y1, yr), mismatch_frac=float((y1 != y0).float().mean()),
             nan=int(torch.isnan(y1.float()).sum()))
    assert m["nan"] == 0 and m["mismatch_frac"] <= 5e-3, m
    return _finish("ln_modulate_stats", m, 3e-3)


def check_ln_affine(rows=333, C=1536, seed=0):
    x = _randn(rows, C, seed=seed, scale=2.0) - 0.2
    w, b = _randn(C, seed=seed + 1) * 0.1 + 1, _randn(C, seed=seed + 2) * 0.1

    def run(ops, o):
        ops.ln_affine(x, o["y"], w, b, 1e-6)
    return _against_double("ln_affine", run, dict(y=torch.zeros(rows, C, device="cuda", dtype=BF)))


def check_rmsnorm(rows=333, C=1536, seed=0):
    x = _randn(rows, C, seed=seed, scale=2.0)
    w = _randn(C, seed=seed + 1) * 0.1 + 1

    def run(ops, o):
        ops.rmsnorm(x, o["y"], w, 1e-6)
    return _against_double("rmsnorm", run, dict(y=torch.zeros(rows, C, device="cuda", dtype=BF)))


def check_qk_norm_rope(B=2, F_=2, Hh=5, Ww=7, C=1536, start_frame=3, seed=0, with_v=True):
    D = 128
    H = C // D
    L = F_ * Hh * Ww
    qkv = _randn(B * L, 3 * C, seed=seed)
    wq, wk = _randn(C, seed=seed + 1) * 0.1 + 1, _randn(C, seed=seed + 2) * 0.1 + 1
    from self_forcing_b200.model import rope_tables
    cos, sin = (t.cuda() for t in rope_tables(D))
    cache_rows = L + 11

    def run(ops, o):
        ops.qk_norm_rope(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:] if with_v else None, wq, wk, 1e-6, cos, sin, B, L,
                         D, (F_, Hh, Ww), start_frame, q_out=o["q"], k_out=o["kc"][:, 5:5 + L], v_out=o["vc"][:, 5:5 + L])
    outs = dict(q=torch.zeros(B, L, C, device="cuda", dtype=BF),
                kc=torch.zeros(B, cache_rows, H, D, device="cuda", dtype=BF),
                vc=torch.zeros(B, cache_rows, H, D, device="cuda", dtype=BF))
    return _against_double("qk_norm_rope", run, outs, exact=("vc",))


def check_qk_norm_rope_stats(B=2, F_=2, Hh=5, Ww=7, C=1536, start_frame=3, seed=0, with_v=True):
    """Streaming form fed by the QKV projection's statistics records against the resident-row kernel: the only difference
    is the summation order of the row's sum of squares -> at most a last-bit difference of the normalisation factor."""
    ops = _ops()
    D = 128
    H = C // D
    L = F_ * Hh * Ww
    qkv = _randn(B * L, 3 * C, seed=seed)
    wq, wk = _randn(C, seed=seed + 1) * 0.1 + 1, _randn(C, seed=seed + 2) * 0.1 + 1
    from self_forcing_b200.model import rope_tables
    cos, sin = (t.cuda() for t in rope_tables(D))
    stats = _stats_ref(qkv).contiguous()
    outs = []
    for use_stats in (False, True):
        q = torch.zeros(B, L, C, device="cuda", dtype=BF)
        kc = torch.zeros(B, L + 11, H, D, device="cuda", dtype=BF)
        vc = torch.zeros_like(kc)
        kw = dict(stats=stats, q_chunk0=0, k_chunk0=C // 128) if use_stats else {}
        ops.qk_norm_rope(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:] if with_v else None, wq, wk, 1e-6, cos, sin, B, L, D,
                         (F_, Hh, Ww), start_frame, q_out=q, k_out=kc[:, 5:5 + L], v_out=vc[:, 5:5 + L], **kw)
        outs.append((q, kc, vc))
    (q0, k0, v0), (q1, k1, v1) = outs
    m = dict(err_q=rel_l2(q1, q0), err_k=rel_l2(k1, k0), err_v_mismatch=float((v1 != v0).sum()),
             q_mismatch_frac=float((q1 != q0).float().mean()), k_mismatch_frac=float((k1 != k0).float().mean()))
    assert m["err_v_mismatch"] == 0 and m["q_mismatch_frac"] <= 2e-3 and m["k_mismatch_frac"] <= 2e-3, m
    return _finish("qk_norm_rope_stats", m, 1e-4)


class _FakeGroup:
    """Stands in for UlyssesGroup in single-GPU kernel checks: `world` ranks emulated in one process."""
    def __init__(self, world, rank):
        self.world, self.rank = world, rank


def check_qk_norm_rope_sp(P=2, F_=2, Hh=6, Ww=7, C=1536, start_frame=3, seed=0):
    """The scatter form run once per emulated rank (token slice) into P per-group destinations on ONE GPU must equal
    the plain kernel's output split by head group: bit-exact."""
    from self_forcing_b200.model import rope_tables
    from self_forcing_b200.ulysses import PeerTensor
    ops = _ops()
    D = 128
    H = C // D
    Hg = H // P
    L = F_ * Hh * Ww
    Lr = L // P
    qkv = _randn(L, 3 * C, seed=seed)
    wq, wk = _randn(C, seed=seed + 1) * 0.1 + 1, _randn(C, seed=seed + 2) * 0.1 + 1
    cos, sin = (t.cuda() for t in rope_tables(D))
    q_ref = torch.zeros(1, L, C, device="cuda", dtype=BF)
    k_ref = torch.zeros(1, L + 9, H, D, device="cuda", dtype=BF)
    v_ref = torch.zeros_like(k_ref)
    ops.qk_norm_rope(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], wq, wk, 1e-6, cos, sin, 1, L, D, (F_, Hh, Ww), start_frame,
                     q_out=q_ref, k_out=k_ref[:, 4:4 + L], v_out=v_ref[:, 4:4 + L])
    # per-group destinations ("rank g's" q buffer and head-sharded cache)
    q_dst = [torch.zeros(L, Hg * D, device="cuda", dtype=BF) for _ in range(P)]
    k_dst = [torch.zeros(L + 9, Hg, D, device="cuda", dtype=BF) for _ in range(P)]
    v_dst = [torch.zeros(L + 9, Hg, D, device="cuda", dtype=BF) for _ in range(P)]
    for r in range(P):
        rows = slice(r * Lr, (r + 1) * Lr)
        q_buf = PeerTensor(q_dst[r], [t.data_ptr() for t in q_dst])
        ops.qk_norm_rope_sp(qkv[rows, :C], qkv[rows, C:2 * C], qkv[rows, 2 * C:], wq, wk, 1e-6, cos, sin, D, (F_, Hh, Ww),
                            start_frame, r * Lr, _FakeGroup(P, r), q_buf, None, None,
                            [t[4:].data_ptr() for t in k_dst], [t[4:].data_ptr() for t in v_dst])
    torch.cuda.synchronize()
    m = {}
    for g in range(P):
        cols = slice(g * Hg * D, (g + 1) * Hg * D)
        m[f"err_q{g}_mismatch"] = float((q_dst[g] != q_ref[0][:, cols]).sum())
        m[f"err_k{g}_mismatch"] = float((k_dst[g] != k_ref[0][:, g * Hg:(g + 1) * Hg]).sum())
        m[f"err_v{g}_mismatch"] = float((v_dst[g] != v_ref[0][:, g * Hg:(g + 1) * Hg]).sum())
    return _finish("qk_norm_rope_sp", m, 0.0)


def check_attention_sp(P=2, Lq=4680, S=9360, Hg=3, seed=0):
    """Head-parallel form: output rows scattered to P per-rank buffers at this head group's columns must equal the
    plain kernel's output (same schedule -> bit-exact)."""
    from self_forcing_b200.ulysses import PeerTensor
    ops = _ops()
    D = 128
    q = _randn(Lq, Hg, D, seed=seed)
    k = _randn(S, Hg, D, seed=seed + 1)
    v = _randn(S, Hg, D, seed=seed + 2)
    ref = torch.zeros(1, Lq, Hg, D, device="cuda", dtype=BF)
    ops.attention(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), ref, 1 / math.sqrt(D))
    rank = 1                                       # pretend to be rank 1 of P: columns [rank*Hg*D, ...)
    C = P * Hg * D
    Lr = Lq // P
    outs = [torch.full((Lr, C), float("nan"), device="cuda", dtype=BF) for _ in range(P)]
    ops.attention_sp(q, k, v, 1 / math.sqrt(D), _FakeGroup(P, rank), PeerTensor(outs[rank], [t.data_ptr() for t in outs]), Lr)
    torch.cuda.synchronize()
    m = {}
    for d in range(P):
        got = outs[d][:, rank * Hg * D:(rank + 1) * Hg * D]
        m[f"err_rows{d}_mismatch"] = float((got != ref[0, d * Lr:(d + 1) * Lr].reshape(Lr, Hg * D)).sum())
        other = torch.cat([outs[d][:, :rank * Hg * D], outs[d][:, (rank + 1) * Hg * D:]], dim=1)
        m[f"err_untouched{d}"] = float((~torch.isnan(other.float())).sum())
    return _finish("attention_sp", m, 0.0)


def check_peer_barrier_single():
    """world = 1 barrier (self-signal) returns; multi-GPU behaviour is covered by tools/ulysses_gpu_check.py."""
    from self_forcing_b200.ulysses import PeerTensor
    ops = _ops()
    flags = torch.zeros(2, device="cuda", dtype=torch.int32)

    class G(_FakeGroup):
        pass
    g = G(1, 0)
    g.flags = PeerTensor(flags, [flags.data_ptr()])
    for _ in range(3):
        ops.peer_barrier(g)
    torch.cuda.synchronize()
    return _finish("peer_barrier", dict(err_flag=float(abs(int(flags[0]) - 3)), err_count=float(abs(int(flags[1]) - 3))), 0.0)


def check_ln_row_offset(rows=300, C=1536, seed=0):
    x = _randn(rows, C, seed=seed)
    tab = _randn(5, 6, C, seed=seed + 1, scale=0.3)

    def run(ops, o):
        ops.ln_modulate(x, o["y"], shift=tab[:, 0], scale=tab[:, 1], mod_stride=6 * C, rows_per_mod=130, eps=1e-6,
                        row_offset=77)
    return _against_double("ln_modulate row_offset", run, dict(y=torch.zeros(rows, C, device="cuda", dtype=BF)))


def check_kv_roll(n_tensors=6, B=2, S=700, H=12, D=128, dst=100, src=230, n=470, seed=0):
    """Rolling-window eviction kernel (all layers' K and V in one call, several phases because n > src - dst) against
    the reference's clone-and-assign (causal_model.py:212-221); rows outside the moved range must be untouched."""
    ops = _ops()
    pool = _randn(n_tensors, B, S + 3, H, D, seed=seed)
    tensors = [pool[i, :, :S] for i in range(n_tensors)]          # non-contiguous batch stride, like views of one pool
    ref = [t.clone() for t in tensors]
    for r in ref:
        r[:, dst:dst + n] = r[:, src:src + n].clone()
    table = torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device="cuda")
    ops.kv_roll(tensors, table, dst, src, n)
    torch.cuda.synchronize()
    m = dict(err_mismatch=float(sum((a != b).sum() for a, b in zip(tensors, ref))))
    return _finish("kv_roll", m, 0.0)


def check_patchify(B=2, F_=3, H=12, W=20, seed=0):
    x = _randn(B, F_, 16, H, W, seed=seed).permute(0, 2, 1, 3, 4)      # the wrapper's permuted view

    def run(ops, o):
        ops.patchify(x, o["t"])
    return _against_double("patchify", run, dict(t=torch.zeros(B * F_ * (H // 2) * (W // 2), 64, device="cuda", dtype=BF)),
                           exact=("t",))


def check_sinusoid(seed=0):
    m = {}
    for t in (torch.tensor([1000.0, 937.5, 833.3333129882812, 625.0, 0.0, 3.0], device="cuda"),
              torch.tensor([1000, 937, 0, 17], device="cuda", dtype=torch.int64)):
        def run(ops, o):
            ops.sinusoid(t, o["s"], 256)
        r = _against_double("sinusoid", run, dict(s=torch.zeros(t.numel(), 256, device="cuda", dtype=BF)), tol=2e-3)
        m.update({f"{k}_{t.dtype}": v for k, v in r.items()})
    return m


def check_skinny_linear(M=3, N=1536, K=256, silu=False, seed=0):
    x = _randn(M, K, seed=seed)
    w = _randn(N, K, seed=seed + 1, scale=1 / math.sqrt(K))
    b = _randn(N, seed=seed + 2, scale=0.1)

    def run(ops, o):
        ops.skinny_linear(x, w, b, o["y"], silu_in=silu)
    return _against_double(f"skinny M{M} N{N} K{K}", run, dict(y=torch.zeros(M, N, device="cuda", dtype=BF)))


def check_modulation_table(NL=3, R=4, C=1536, seed=0):
    mod = _randn(NL, 6, C, seed=seed)
    e0 = _randn(R, 6 * C, seed=seed + 1)
    hm = _randn(1, 2, C, seed=seed + 2)
    e = _randn(R, C, seed=seed + 3)

    def run(ops, o):
        ops.modulation_table(mod, e0, o["a"], e_row_stride=6 * C, e_group_stride=C)
        ops.modulation_table(hm, e, o["b"], e_row_stride=C, e_group_stride=0)
    return _against_double("modulation_table", run, dict(a=torch.zeros(NL, R, 6, C, device="cuda", dtype=BF),
                                                         b=torch.zeros(1, R, 2, C, device="cuda", dtype=BF)),
                           exact=("a", "b"))


def check_head_finish(B=2, F_=3, H=12, W=20, shift=5.0, seed=0):
    sched = O.OracleScheduler(shift)
    ts, sg = sched.timesteps.cuda(), sched.sigmas.cuda()
    L = F_ * (H // 2) * (W // 2)
    head_out = _randn(B * L, 64, seed=seed)
    big = _randn(B, F_ + 2, 16, H, W, seed=seed + 1)
    xt = big[:, 1:1 + F_]                                   # non-contiguous slice like noise[:, a:b]
    m = {}
    for timestep in (torch.tensor([[1000.0, 937.5, 833.3333129882812]] * B, device="cuda"),
                     torch.tensor([[0, 625, 250]] * B, device="cuda", dtype=torch.int64)):
        def run(ops, o):
            ops.head_finish(head_out, xt, timestep, ts, sg, o["flow"], o["x0"])
        r = _against_double("head_finish", run, dict(flow=torch.zeros(B, F_, 16, H, W, device="cuda", dtype=BF),
                                                     x0=torch.zeros(B, F_, 16, H, W, device="cuda", dtype=BF)),
                            exact=("flow", "x0"))
        m.update({f"{k}_{timestep.dtype}": v for k, v in r.items()})
    return m


def check_add_noise(N=3, shift=5.0, seed=0):
    sched = O.OracleScheduler(shift)
    ts, sg = sched.timesteps.cuda(), sched.sigmas.cuda()
    x0 = _randn(N, 16, 12, 20, seed=seed)
    nz = _randn(N, 16, 12, 20, seed=seed + 1)
    m = {}
    for timestep in (torch.tensor([937.5, 833.3333129882812, 625.0], device="cuda")[:N],
                     torch.tensor([937, 0, 250], device="cuda", dtype=torch.int64)[:N]):
        def run(ops, o):
            ops.add_noise(x0, nz, timestep, ts, sg, o["y"])
        r = _against_double("add_noise", run, dict(y=torch.zeros_like(x0)), exact=("y",))
        m.update({f"{k}_{timestep.dtype}": v for k, v in r.items()})
    return m


def check_cfg_unipc_step(corr=2, pred=2, cfg=True, n=3 * 16 * 60 * 104, seed=0):
    """The fused sampler step against the op-by-op torch chain on the same GPU (bit-exact: same rounding points)."""
    t = [_randn(n, seed=seed + i) for i in range(6)]
    fc, fu, x, xl, m0, m1 = t
    coef = [3.0, 0.9375, 0.9957, -0.0043, 0.0042, -1.37, 0.3125, 0.4375, 0.9915, -0.0085, 0.0084, -0.93]

    def run(ops, o):
        ops.cfg_unipc_step(fc, fu if cfg else None, x, xl if corr else None, m0 if (corr or pred == 2) else None,
                           m1 if corr == 2 else None, o["m"], o["xc"], o["xn"], coef, corr, pred)
    z = torch.zeros(n, device="cuda", dtype=BF)
    return _against_double("cfg_unipc_step", run, dict(m=z, xc=z.clone(), xn=z.clone()), exact=("m", "xc", "xn"))


def check_cfg_unipc_alias(n=2 * 16 * 60 * 104, seed=3):
    """m_out aliasing m1 and sample_out aliasing last_sample (the scheduler's history ring) give the same result."""
    fc, fu, x, xl, m0, m1 = [_randn(n, seed=seed + i) for i in range(6)]
    coef = [7.5, 0.5, 0.98, -0.02, 0.019, -1.1, 0.25, 0.5, 0.97, -0.03, 0.029, -0.9]
    ops = _ops()
    m, xc, xn = (torch.zeros(n, device="cuda", dtype=BF) for _ in range(3))
    ops.cfg_unipc_step(fc, fu, x, xl, m0, m1, m, xc, xn, coef, 2, 2)
    m1a, xla, xn2 = m1.clone(), xl.clone(), torch.zeros_like(xn)
    ops.cfg_unipc_step(fc, fu, x, xla, m0, m1a, m1a, xla, xn2, coef, 2, 2)
    torch.cuda.synchronize()
    assert torch.equal(m, m1a) and torch.equal(xc, xla) and torch.equal(xn, xn2)
    return dict(err_alias_mismatch=0.0)


def check_unipc_reference_semantics(steps=50, shift=5.0):
    """50 solver steps on the GPU: host coefficients + fused kernel against the oracle's op-by-op chain run on CUDA
    tensors with 0-dim CPU scalar tensors, i.e. exactly what the reference executes on a GPU.  Also reports which
    scalar-rounding mode of unipc.py matches it."""
    from oracle import unipc_oracle as U
    from oracle.make_golden import UNIPC_TRACE, unipc_trace_flow
    from self_forcing_b200.unipc import FlowUniPCMultistepScheduler
    x0 = torch.randn(UNIPC_TRACE["shape"], generator=torch.Generator().manual_seed(99)).to(BF).cuda()
    out = {}
    ref = U.OracleUniPC()
    ref.set_timesteps(steps, shift)
    xr, ref_xs = x0, []
    for i, t in enumerate(ref.timesteps):
        xr = ref.step(unipc_trace_flow(xr.cpu(), i).cuda(), t, xr)
        ref_xs.append(xr)
    for mode in ("fp32", "bf16"):
        s = FlowUniPCMultistepScheduler(shift=1, ops=_ops(), scalar_rounding=mode)
        s.set_timesteps(steps, device="cuda", shift=shift)
        x, worst, n_bad = x0, 0.0, 0
        for i, t in enumerate(s.timesteps):
            # feed both chains the SAME flow (derived from the reference chain's sample) so that errors do not compound
            flow = unipc_trace_flow((ref_xs[i - 1] if i else x0).cpu(), i).cuda()
            x = s.step(flow, t, ref_xs[i - 1] if i else x0, return_dict=False)[0]
            worst = max(worst, rel_l2(x, ref_xs[i]))
            n_bad += int((x != ref_xs[i]).sum())
        out[f"worst_rel_l2_{mode}"] = worst
        out[f"mismatches_{mode}"] = float(n_bad)
    torch.cuda.synchronize()
    assert out["worst_rel_l2_fp32"] <= 2e-3, out
    return out


# ---- VAE decoder kernels ---------------------------------
def check_vae_latent_in(h=12, w=20, seed=0):
    z = _randn(16, h, w, seed=seed, scale=2.0)
    mean, inv = _randn(16, seed=seed + 1), (_randn(16, seed=seed + 2).abs() + 0.3)
    wt, b = _randn(16, 16, seed=seed + 3, scale=0.25), _randn(16, seed=seed + 4, scale=0.1)

    def run(ops, o):
        ops.vae_latent_in(z, mean, inv, wt, b, o["y"])
    return _against_double("vae_latent_in", run, dict(y=torch.zeros(h * w, 16, device="cuda", dtype=BF)))


def check_vae_norm_silu(rows=1000, C=96, silu=True, seed=0):
    x = _randn(rows, C, seed=seed, scale=1.5)
    x[3] = 0                                                     # an all-zero voxel exercises the eps clamp
    g = _randn(C, seed=seed + 1) * 0.1 + 1

    def run(ops, o):
        ops.vae_norm_silu(x, g, o["y"], silu)
    return _against_double("vae_norm_silu", run, dict(y=torch.zeros(rows, C, device="cuda", dtype=BF)))


def check_causal_conv3d(t_in=1, H=8, W=12, Cin=16, Cout=384, kt=3, ks=3, pad=2, upsample=False, residual=False,
                        segments=False, ws_bytes=None, seed=0, implicit=False):
    """Gather + tcgen05 GEMM against F.conv3d on the same GPU (fp32 accumulation on both sides)."""
    x = _randn(t_in, H, W, Cin, seed=seed)
    K = kt * ks * ks * Cin
    w = _randn(Cout, K, seed=seed + 1, scale=K ** -0.5)
    b = _randn(Cout, seed=seed + 2, scale=0.1)
    t_out = t_in + pad - (kt - 1)
    Ho, Wo = (2 * H, 2 * W) if upsample else (H, W)
    rows = t_out * Ho * Wo
    res = _randn(rows, Cout, seed=seed + 3) if residual else None
    cu = _ops()
    saved = cu.conv_workspace_bytes
    if ws_bytes is not None:
        cu.conv_workspace_bytes = ws_bytes
    if segments:
        y0, y1 = (torch.zeros(rows, Cout // 2, device="cuda", dtype=BF) for _ in range(2))
    else:
        y0, y1 = torch.zeros(rows, Cout, device="cuda", dtype=BF), None
    cu.causal_conv3d(x, pad, w, b, kt, ks, y0, y1, upsample=upsample, residual=res, seg_cols=Cout // 2 if segments else 0,
                     implicit=implicit)
    got = y0 if y1 is None else torch.cat([y0, y1], dim=1)
    # reference: the same convolution in fp32 (F.conv3d through the test double), rounded where the reference rounds:
    # bf16(conv + bias), then bf16(residual + that)
    exact = torch.zeros(rows, Cout, device="cuda", dtype=torch.float32)
    torch.backends.cudnn.allow_tf32 = False
    TorchOps().causal_conv3d(x.float(), pad, w.float(), b.float(), kt, ks, exact, upsample=upsample)
    ref = exact.to(BF)
    if residual:
        ref = res + ref
    torch.cuda.synchronize()
    cu.conv_workspace_bytes = saved
    return _finish("causal_conv3d", dict(err=rel_l2(got, ref), max=float((got.float() - ref.float()).abs().max()),
                                         err_vs_fp32=rel_l2(got, exact + (res.float() if residual else 0))), 3e-3)


def check_upsample2x(T=2, H=5, W=7, C=96, seed=0):
    x = _randn(T, H, W, C, seed=seed)

    def run(ops, o):
        ops.upsample2x(x, o["y"])
    return _against_double("upsample2x", run, dict(y=torch.zeros(T, 2 * H, 2 * W, C, device="cuda", dtype=BF)), exact=("y",))


def check_gemm_f32(M=300, N=312, K=384, seed=0):
    x, w = _randn(M, K, seed=seed), _randn(N, K, seed=seed + 1)
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    _ops().gemm(x, w, None, out, epilogue=4)
    ref = x.float() @ w.float().t()
    torch.cuda.synchronize()
    return _finish("gemm_f32", dict(err=rel_l2(out, ref)), 1e-5)


def check_softmax_transpose(rows=300, cols=312, seed=0):
    s = _randn(rows, cols, seed=seed, dtype=torch.float32) * 30
    p = torch.zeros(rows, cols, device="cuda", dtype=BF)
    _ops().softmax_rows(s, p, 0.051)
    ref = torch.softmax(s * 0.051, dim=1)
    x = _randn(rows, cols, seed=seed + 1)
    xt = torch.zeros(cols, rows, device="cuda", dtype=BF)
    _ops().transpose(x[:, 8:], xt[:cols - 8])                     # strided source view
    torch.cuda.synchronize()
    assert torch.equal(xt[:cols - 8], x[:, 8:].t())
    return _finish("softmax_rows", dict(err=rel_l2(p, ref), err_rowsum=float((p.float().sum(1) - 1).abs().max())), 1e-2)


def check_vae_pixel_out(T=3, H=16, W=24, seed=0):
    y = _randn(T * H * W, 8, seed=seed, scale=0.8)

    def run(ops, o):
        ops.vae_pixel_out(y, o["px"])
    return _against_double("vae_pixel_out", run, dict(px=torch.zeros(T, 3, H, W, device="cuda")), exact=("px",))


def check_vae_decoder(implicit=False):
    """Whole decoder on the B200 against the pixels of the unmodified reference (CPU, bf16) and its fp32 run."""
    from helpers import golden
    from oracle import vae_oracle as V
    from oracle.make_golden import VAE_CASE, vae_latents
    from self_forcing_b200.vae import B200VAEWrapper
    g = golden("vae_decode_tiny.pt")
    w = B200VAEWrapper(state_dict=V.make_random_vae_params(V.VaeConfig(), seed=VAE_CASE["seed"]), device="cuda", ops=_ops())
    w.model.implicit_conv = implicit
    lat = vae_latents().cuda()
    out = w.decode_to_pixel(lat)
    a = w.decode_to_pixel(lat[:, :2], use_cache=True)
    b = w.decode_to_pixel(lat[:, 2:], use_cache=True)
    torch.cuda.synchronize()
    ref = g["pixels"].float().clamp(-1, 1).permute(0, 2, 1, 3, 4)
    exact = g["pixels_fp32"].clamp(-1, 1).permute(0, 2, 1, 3, 4)
    m = dict(err_vs_ref_bf16=rel_l2(out.cpu(), ref), vs_fp32=rel_l2(out.cpu(), exact), ref_noise_floor=rel_l2(ref, exact),
             err_stream_mismatch=float((torch.cat([a, b], dim=1) != out).sum()))
    assert m["vs_fp32"] <= 1.25 * m["ref_noise_floor"], m
    assert m["err_stream_mismatch"] == 0, m
    return _finish("vae_decoder", m, 2.5e-2)


# --------------------------------------------------------------------------------------
def _tiny_setup(num_layers=2, ffn_dim=512, shift=5.0, seed=0):
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper
    cfg = O.OracleConfig(dim=1536, ffn_dim=ffn_dim, num_heads=12, num_layers=num_layers)
    params = O.make_random_params(cfg, seed=seed)
    w = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B, ffn_dim=ffn_dim, num_layers=num_layers),
                             timestep_shift=shift, device="cuda")
    w.model.load_state_dict(params, strict=True)
    gpu_params = {k: v.cuda() for k, v in params.items()}
    return cfg, gpu_params, w


def check_model_forward(F_=1, H=60, W=104, seed=0):
    """Two consecutive cached forwards of a 2-layer model vs the oracle run on the same GPU."""
    cfg, params, w = _tiny_setup()
    ow = O.OracleWrapper(params, cfg, 5.0)
    fs = (H // 2) * (W // 2)
    pe = _randn(1, 512, 4096, seed=1)
    x = _randn(1, 2 * F_, 16, H, W, seed=2)
    m = {}
    kv_a, ca_a = O.new_kv_cache(cfg, 1, fs, BF, "cuda", cache_tokens=4 * F_ * fs), O.new_crossattn_cache(cfg, 1, BF, "cuda")
    kv_b, ca_b = O.new_kv_cache(cfg, 1, fs, BF, "cuda", cache_tokens=4 * F_ * fs), O.new_crossattn_cache(cfg, 1, BF, "cuda")
    for step, (t, start) in enumerate([(1000.0, 0), (0.0, 0), (937.5, F_)]):
        xin = x[:, :F_] if start == 0 else x[:, F_:]
        ts = torch.full((1, F_), t, device="cuda")
        f1, x1 = w(xin, {"prompt_embeds": pe}, ts, kv_cache=kv_a, crossattn_cache=ca_a, current_start=start * fs)
        with torch.no_grad():
            f2, x2 = ow(xin, pe, ts, kv_b, ca_b, start * fs)
        m[f"err_flow_{step}"] = rel_l2(f1, f2)
        m[f"err_x0_{step}"] = rel_l2(x1, x2)
        m[f"err_k_{step}"] = rel_l2(kv_a[1]["k"], kv_b[1]["k"])
        m[f"err_v_{step}"] = rel_l2(kv_a[1]["v"], kv_b[1]["v"])
        m[f"idx_{step}"] = (int(kv_a[0]["global_end_index"]), int(kv_a[0]["local_end_index"]),
                            int(kv_b[0]["global_end_index"]), int(kv_b[0]["local_end_index"]))
        assert m[f"idx_{step}"][:2] == m[f"idx_{step}"][2:], m
    return _finish("model_forward", m, 1e-2)


# ---- UMT5 text encoder kernels ----------------------
def check_t5_rmsnorm(rows=300, C=4096, seed=0):
    x, w = _randn(rows, C, seed=seed, scale=3.0), _randn(C, seed=seed + 1) * 0.1 + 1

    def run(ops, o):
        ops.t5_rmsnorm(x, w, o["y"], 1e-6)
    return _against_double("t5_rmsnorm", run, dict(y=torch.zeros(rows, C, device="cuda", dtype=BF)))


def check_softmax_bias_rows(rows=200, cols=512, seed=0):
    s, b = _randn(rows, cols, seed=seed, scale=4.0), _randn(rows, cols, seed=seed + 1)
    mask = torch.ones(cols, device="cuda", dtype=torch.int32)
    mask[300:] = 0

    def run(ops, o):
        ops.softmax_bias_rows(s, b, mask, o["p"])
        ops.softmax_bias_rows(s, b, None, o["q"])
    m = _against_double("softmax_bias_rows", run, dict(p=torch.zeros(rows, cols, device="cuda", dtype=BF),
                                                       q=torch.zeros(rows, cols, device="cuda", dtype=BF)), tol=5e-3)
    return m


def check_t5_gated_gelu(rows=300, cols=640, seed=0):
    a, g = _randn(rows, cols, seed=seed), _randn(rows, cols, seed=seed + 1, scale=2.0)

    def run(ops, o):
        ops.t5_gated_gelu(a, g, o["y"])
    return _against_double("t5_gated_gelu", run, dict(y=torch.zeros(rows, cols, device="cuda", dtype=BF)), tol=5e-3)


def check_t5_encoder():
    """Whole (tiny) encoder on the B200 against the unmodified reference's output (CPU, bf16) and its fp32 run."""
    from helpers import golden
    from oracle import t5_oracle as T
    from oracle.make_golden import T5_CASE, t5_case_cfg, t5_case_inputs
    from self_forcing_b200.t5 import B200T5Encoder
    g = golden("t5_tiny.pt")
    cfg = t5_case_cfg()
    enc = B200T5Encoder(vocab=cfg.vocab, dim=cfg.dim, dim_attn=cfg.dim_attn, dim_ffn=cfg.dim_ffn, num_heads=cfg.num_heads,
                        num_layers=cfg.num_layers, ops=_ops(), device="cuda")
    enc.load_state_dict(T.make_random_t5_params(cfg, seed=T5_CASE["seed"]))
    ids, mask = t5_case_inputs()
    out = enc(ids.cuda(), mask.cuda()).cpu()
    # second call captures the forward as a CUDA graph, third and fourth replay it (the fourth with other inputs in between)
    out2 = enc(ids.cuda(), mask.cuda()).cpu()
    out3 = enc(ids.cuda(), mask.cuda()).cpu()
    enc(ids.flip(1).cuda().contiguous(), mask.cuda())
    out4 = enc(ids.cuda(), mask.cuda()).cpu()
    assert torch.equal(out, out2) and torch.equal(out, out3) and torch.equal(out, out4), "graph replay differs from the eager forward"
    for u, n in zip(out, T5_CASE["lengths"]):
        u[n:] = 0
    torch.cuda.synchronize()
    m = dict(err_vs_ref_bf16=rel_l2(out, g["context_bf16"]), vs_fp32=rel_l2(out, g["context_fp32"]),
             ref_noise_floor=rel_l2(g["context_bf16"], g["context_fp32"]))
    assert m["vs_fp32"] <= 1.25 * m["ref_noise_floor"], m
    return _finish("t5_encoder", m, 3e-2)


# Checks of code paths that exist but have not yet been measured / validated on hardware: NOT part of the pytest suite;
# run them with `python tools/gpu_report.py --pending` and move them into ALL once green.
PENDING = {
}

ALL = {
    # CTA-pair kernel (gemm2_tcgen05.cu): 512 = one pair per cluster, 515 = two pairs sharing A by TMA multicast
    **{f"gemm_c{np_}_{name}": (lambda kw=kw, np_=np_: check_gemm(block_n=(512, 515)[np_ - 1], **kw)) for np_ in (1, 2) for name, kw in {
        "small": dict(M=300, N=512, K=192, seed=40),
        "tail": dict(M=130, N=256, K=64, seed=41),
        "one_cta_empty_residual": dict(M=700, N=256, K=1536, epilogue=2, seed=42),
        "odd_n_blocks_gelu": dict(M=515, N=768, K=320, epilogue=1, seed=43),
        "gate_two_vectors": dict(M=1000, N=768, K=320, epilogue=3, rows_per_gate=100, seed=44),
        "gate_many_vectors": dict(M=1000, N=512, K=320, epilogue=3, rows_per_gate=30, gate_row_offset=77, seed=45),
        "qkv_full": dict(M=4680, N=4608, K=1536, seed=46),
        "o_proj_full": dict(M=4680, N=1536, K=1536, epilogue=3, rows_per_gate=1560, seed=47),
        "ffn1_full": dict(M=4680, N=8960, K=1536, epilogue=1, seed=48),
        "ffn2_full": dict(M=4680, N=1536, K=8960, epilogue=3, rows_per_gate=1560, seed=49),
        "persistent_many_tiles": dict(M=4680, N=8960, K=256, epilogue=2, seed=50),
    }.items()},
    "gemm_c1_segments": lambda: check_gemm_segments(block_n=512, seed=51),
    "gemm_c2_segments": lambda: check_gemm_segments(M=700, C=512, K=256, block_n=515, seed=52),
    "gemm_small": lambda: check_gemm(),
    "gemm_bn64": lambda: check_gemm(M=200, N=64, K=1536, block_n=64),
    "gemm_bn128": lambda: check_gemm(M=300, N=384, K=256, block_n=128),
    "gemm_bn256": lambda: check_gemm(M=515, N=512, K=320, block_n=256),
    "gemm_k64": lambda: check_gemm(M=700, N=1536, K=64),
    "gemm_gelu": lambda: check_gemm(M=300, N=512, K=256, epilogue=1),
    "gemm_residual": lambda: check_gemm(M=300, N=256, K=256, epilogue=2),
    "gemm_gate": lambda: check_gemm(M=300, N=256, K=256, epilogue=3, rows_per_gate=70),
    "gemm_qkv_full": lambda: check_gemm(M=4680, N=4608, K=1536),
    "gemm_ffn2_full": lambda: check_gemm(M=4680, N=1536, K=8960, epilogue=3, rows_per_gate=1560),
    "gemm_segments": check_gemm_segments,
    "gemm_segments_1cta": lambda: check_gemm_segments(C=128),
    "gemm_pair_small": lambda: check_gemm(M=300, N=512, K=192, block_n=512),
    "gemm_pair_tail": lambda: check_gemm(M=130, N=256, K=64, block_n=512),
    "gemm_pair_one_cta_empty": lambda: check_gemm(M=700, N=256, K=1536, block_n=512, epilogue=2),
    "gemm_pair_persistent_gelu": lambda: check_gemm(M=4680, N=8960, K=256, epilogue=1, block_n=512),
    "gemm_pair_gate": lambda: check_gemm(M=1000, N=768, K=320, epilogue=3, rows_per_gate=70, block_n=512),
    "gemm_pair_o_proj": lambda: check_gemm(M=4680, N=1536, K=1536, epilogue=3, rows_per_gate=1560, block_n=512),
    "gemm_stats_gate": check_gemm_stats,
    "gemm_stats_bias_shifted": lambda: check_gemm_stats(M=300, N=256, K=128, epilogue=2, seed=3, mean_shift=40.0),
    "gemm_stats_o_proj": lambda: check_gemm_stats(M=4680, N=1536, K=1536, epilogue=3, rows_per_gate=1560, seed=4),
    "gemm_stats_ffn2_cluster": lambda: check_gemm_stats(M=4680, N=1536, K=8960, epilogue=3, rows_per_gate=1560, seed=5),
    "gemm_lnfold": check_gemm_lnfold,
    "gemm_lnfold_shifted": lambda: check_gemm_lnfold(M=300, N=256, K=512, seed=7, mean_shift=25.0),
    "gemm_lnfold_cross_q": lambda: check_gemm_lnfold(M=4680, N=1536, K=1536, seed=8),
    "attn_qnorm": check_attention_qnorm,
    "attn_qnorm_cross": lambda: check_attention_qnorm(Lq=4680, S=512, H=12, seed=2),
    "attn_qnorm_batch_split": lambda: check_attention_qnorm(B=2, Lq=1300, S=7000, H=4, seed=3),
    "attn_small": lambda: check_attention(),
    "attn_one_tile": lambda: check_attention(Lq=128, S=128, H=1),
    "attn_tail": lambda: check_attention(Lq=72, S=72, H=3),
    "attn_window": lambda: check_attention(Lq=260, S=1000, H=2, window_start=37),
    "attn_batch": lambda: check_attention(B=2, Lq=200, S=384, H=2),
    "attn_fusedq": lambda: check_attention(Lq=300, S=520, H=2, fused_q=True),
    "attn_cross": lambda: check_attention(Lq=1560, S=512, H=12),
    "attn_sharp": check_attention_sharp,
    "attn_chunk": lambda: check_attention(Lq=4680, S=4680, H=12),
    "attn_split_tail": lambda: check_attention(Lq=4000, S=5000, H=12, seed=5),
    "attn_split_batch": lambda: check_attention(B=2, Lq=1560, S=9360, H=12, seed=6),
    "attn_split_3way": lambda: check_attention(Lq=1300, S=40000, H=32, seed=7),
    "attn_whole_items": lambda: check_attention(Lq=1560, S=4680, H=12, seed=8),
    "attn_half_even_kv": lambda: check_attention(B=2, Lq=400, S=1024, H=3, seed=11),          # lone 4th query tile, 8 KV tiles
    "attn_half_two_kv": lambda: check_attention(Lq=100, S=200, H=5, seed=12),                 # half items only, one step each
    "attn_half_cross_full": lambda: check_attention(Lq=4680, S=512, H=12, seed=13),           # cross-attention of a chunk
    "attn_half_split_long": lambda: check_attention(Lq=4680, S=18720, H=12, seed=14),         # split schedule with half items
    "ln_modulate": check_ln_modulate,
    "ln_modulate_stats": check_ln_modulate_stats,
    "ln_modulate_stats_chunk": lambda: check_ln_modulate_stats(rows=4680, rpm=1560, seed=3, mean_shift=-1.5),
    "ln_modulate_stats_offset": lambda: check_ln_modulate_stats(rows=300, rpm=130, seed=4, mean_shift=20.0, row_offset=77),
    "ln_affine": check_ln_affine,
    "rmsnorm": check_rmsnorm,
    "qk_norm_rope": check_qk_norm_rope,
    "qk_norm_rope_nov": lambda: check_qk_norm_rope(B=1, with_v=False),
    "qk_norm_rope_stats": check_qk_norm_rope_stats,
    "qk_norm_rope_stats_chunk": lambda: check_qk_norm_rope_stats(B=1, F_=3, Hh=30, Ww=52, start_frame=6, seed=5, with_v=False),
    "qk_norm_rope_sp": check_qk_norm_rope_sp,
    "qk_norm_rope_sp4": lambda: check_qk_norm_rope_sp(P=4, Hh=8),
    "attn_sp": check_attention_sp,
    "attn_sp4": lambda: check_attention_sp(P=4, Lq=4680, S=4680, Hg=3),
    "attn_few_items": lambda: check_attention(Lq=1560, S=32760, H=12, seed=9),
    "attn_very_few_items": lambda: check_attention(Lq=300, S=20000, H=2, seed=10),
    "peer_barrier": check_peer_barrier_single,
    "ln_row_offset": check_ln_row_offset,
    "ln_row_offset_c5120": lambda: check_ln_row_offset(rows=200, C=5120, seed=3),
    "qk_norm_rope_c5120": lambda: check_qk_norm_rope(B=1, F_=2, Hh=4, Ww=6, C=5120, start_frame=0, seed=4),
    "gemm_gate_row_offset": lambda: check_gemm(M=700, N=512, K=256, epilogue=3, rows_per_gate=130, block_n=512, gate_row_offset=77),
    "gemm_gate_row_offset_1cta": lambda: check_gemm(M=300, N=384, K=256, epilogue=3, rows_per_gate=70, block_n=128, gate_row_offset=33),
    "kv_roll": check_kv_roll,
    "kv_roll_single_phase": lambda: check_kv_roll(n_tensors=2, B=1, S=3120, dst=1560, src=3000, n=120, seed=1),
    "kv_roll_framewise_1p3b": lambda: check_kv_roll(n_tensors=4, B=1, S=6 * 1560, dst=1560, src=2 * 1560, n=4 * 1560, seed=2),
    "patchify": check_patchify,
    "sinusoid": check_sinusoid,
    "skinny_linear": check_skinny_linear,
    "skinny_linear_silu": lambda: check_skinny_linear(M=6, N=9216, K=1536, silu=True),
    "modulation_table": check_modulation_table,
    "head_finish": check_head_finish,
    "add_noise": check_add_noise,
    "cfg_unipc_step": check_cfg_unipc_step,
    "cfg_unipc_first_step": lambda: check_cfg_unipc_step(corr=0, pred=1, seed=10),
    "cfg_unipc_warmup": lambda: check_cfg_unipc_step(corr=1, pred=2, seed=11),
    "cfg_unipc_last_step": lambda: check_cfg_unipc_step(corr=2, pred=1, seed=12),
    "unipc_no_guidance_ragged": lambda: check_cfg_unipc_step(corr=2, pred=2, cfg=False, n=8 * 1001 + 5, seed=13),
    "cfg_unipc_alias": check_cfg_unipc_alias,
    "unipc_reference_semantics": check_unipc_reference_semantics,
    "vae_latent_in": check_vae_latent_in,
    "vae_norm_silu_c96": check_vae_norm_silu,
    "vae_norm_c192": lambda: check_vae_norm_silu(rows=333, C=192, silu=False, seed=1),
    "vae_norm_silu_c384": lambda: check_vae_norm_silu(rows=777, C=384, seed=2),
    "conv3d_first_frame_k432": check_causal_conv3d,
    "conv3d_cached_residual": lambda: check_causal_conv3d(t_in=4, H=10, W=14, Cin=96, Cout=96, pad=0, residual=True, seed=1),
    "conv3d_one_cached_frame": lambda: check_causal_conv3d(t_in=2, H=6, W=10, Cin=192, Cout=384, pad=1, seed=2),
    "conv2d_upsample": lambda: check_causal_conv3d(t_in=2, H=6, W=10, Cin=192, Cout=96, kt=1, pad=0, upsample=True, seed=3),
    "time_conv_two_frames": lambda: check_causal_conv3d(t_in=3, H=6, W=10, Cin=384, Cout=768, ks=1, pad=0, segments=True, seed=4),
    "conv_shortcut_1x1x1": lambda: check_causal_conv3d(t_in=2, H=6, W=10, Cin=192, Cout=384, kt=1, ks=1, pad=0, seed=5),
    "conv3d_head_8_channels": lambda: check_causal_conv3d(t_in=3, H=16, W=24, Cin=96, Cout=8, pad=0, seed=6),
    "conv3d_chunked_workspace": lambda: check_causal_conv3d(t_in=3, H=16, W=24, Cin=96, Cout=96, pad=0, ws_bytes=200 * 2592 * 2, seed=7),
    "conv3d_implicit_first_frame": lambda: check_causal_conv3d(implicit=True, seed=20),
    "conv3d_implicit_cached_residual": lambda: check_causal_conv3d(t_in=4, H=10, W=14, Cin=96, Cout=96, pad=0, residual=True, implicit=True, seed=21),
    "conv3d_implicit_one_cached_frame": lambda: check_causal_conv3d(t_in=2, H=6, W=10, Cin=192, Cout=384, pad=1, implicit=True, seed=22),
    "conv2d_implicit": lambda: check_causal_conv3d(t_in=2, H=12, W=20, Cin=192, Cout=96, kt=1, pad=0, implicit=True, seed=23),
    "conv3d_implicit_head_8_channels": lambda: check_causal_conv3d(t_in=3, H=16, W=24, Cin=96, Cout=8, pad=0, implicit=True, seed=24),
    "conv3d_implicit_wide_rows": lambda: check_causal_conv3d(t_in=3, H=9, W=136, Cin=96, Cout=192, pad=0, residual=True, implicit=True, seed=25),
    "conv3d_implicit_many_tiles": lambda: check_causal_conv3d(t_in=5, H=40, W=72, Cin=384, Cout=384, pad=0, implicit=True, seed=26),
    # >= 256 * 4 * SMs voxels with Cout <= 128: the 256-voxel (two accumulator sub-tiles) variant, ragged H and W
    "conv3d_implicit_dual_tile": lambda: check_causal_conv3d(t_in=5, H=205, W=250, Cin=96, Cout=96, pad=0, residual=True, implicit=True, seed=27),
    "conv2d_implicit_dual_tile": lambda: check_causal_conv3d(t_in=2, H=300, W=264, Cin=192, Cout=96, kt=1, pad=0, implicit=True, seed=28),
    "upsample2x": check_upsample2x,
    "vae_decoder_implicit": lambda: check_vae_decoder(implicit=True),
    "gemm_f32_logits": check_gemm_f32,
    "softmax_rows_transpose": check_softmax_transpose,
    "vae_pixel_out": check_vae_pixel_out,
    "vae_decoder_gather": check_vae_decoder,
    "model_forward": check_model_forward,
    "conv3d_implicit_n192": lambda: check_causal_conv3d(t_in=4, H=24, W=40, Cin=192, Cout=192, pad=0, residual=True,
                                                        implicit=True, seed=33),
    # UMT5 text encoder (validated on B200 in round 2: profiles/r02a_pending_report.json)
    "t5_rmsnorm_c4096": check_t5_rmsnorm,
    "t5_rmsnorm_c256": lambda: check_t5_rmsnorm(rows=77, C=256, seed=3),
    "softmax_bias_rows": check_softmax_bias_rows,
    "t5_gated_gelu": check_t5_gated_gelu,
    "t5_encoder": check_t5_encoder,
}


if __name__ == "__main__":
    import json
    name = sys.argv[1]
    try:
        res = (ALL.get(name) or PENDING[name])()
        print("RESULT " + json.dumps(dict(name=name, ok=True, metrics=res)))
    except AssertionError as e:
        print("RESULT " + json.dumps(dict(name=name, ok=False, error=str(e)[:2000])))
        sys.exit(1)
