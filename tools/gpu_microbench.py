"""Device-time microbenchmarks of the two tensor-core kernels at the 1.3B rollout shapes.
Writes gpurun_out/microbench.json.  CUDA events on the launching stream, L2 flushed between reps."""
from __future__ import annotations

import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from self_forcing_b200.ops import CudaOps  # noqa: E402

BF = torch.bfloat16


def timeit(fn, reps=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ops = CudaOps()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = []
    only = sys.argv[1:]
    M = 4680
    # (name, N, K, epilogue, block_n): 0 = the dispatcher's choice, 512 / 515 = pair tiles with 1 / 2 pairs per cluster
    gemms = [("qkv", 4608, 1536, 0, 0), ("qkv_1cta256", 4608, 1536, 0, 256), ("o_proj", 1536, 1536, 3, 0),
             ("o_proj_1cta128", 1536, 1536, 3, 128), ("cross_q", 1536, 1536, 0, 0), ("cross_o", 1536, 1536, 2, 0),
             ("ffn1", 8960, 1536, 1, 0), ("ffn2", 1536, 8960, 3, 0), ("ffn2_np1", 1536, 8960, 3, 512),
             ("ffn2_np2", 1536, 8960, 3, 515), ("head", 64, 1536, 0, 0)]
    for name, N, K, epi, bn in gemms:
        if only and not any(o in "gemm_" + name for o in only):
            continue
        x = torch.randn(M, K, device="cuda").to(BF)
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(BF)
        b = torch.randn(N, device="cuda").to(BF)
        res = torch.randn(M, N, device="cuda").to(BF)
        gate = torch.randn(3, N, device="cuda").to(BF)
        y = torch.empty(M, N, device="cuda", dtype=BF)
        kw = dict(epilogue=epi, block_n=bn)
        if epi in (2, 3):
            kw["residual"] = res
        if epi == 3:
            kw.update(gate=gate, gate_stride=N, rows_per_gate=1560)
        med, best = timeit(lambda: ops.gemm(x, w, b, y, **kw), flush=flush)
        tmed, _ = timeit(lambda: torch.nn.functional.linear(x, w, b), flush=flush)
        fl = 2.0 * M * N * K
        rec = dict(kernel="gemm_" + name, M=M, N=N, K=K, ms=med, ms_best=best, tflops=fl / med / 1e9,
                   cublas_ms=tmed, cublas_tflops=fl / tmed / 1e9)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    # the cross-attention q projection with norm3 folded in (+ statistics of q) and the o projection writing statistics
    for name, epi, use_ln in [("cross_q_fold", 0, True), ("o_proj_stats", 3, False), ("cross_q_stats_only", 0, False)]:
        if only and not any(o in "gemm_" + name for o in only):
            continue
        N = K = 1536
        x = torch.randn(M, K, device="cuda").to(BF)
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(BF)
        b = torch.randn(N, device="cuda").to(BF)
        res = torch.randn(M, N, device="cuda").to(BF)
        gate = torch.randn(3, N, device="cuda").to(BF)
        y = torch.empty(M, N, device="cuda", dtype=BF)
        st_in = torch.rand(M, K // 128, 2, device="cuda") + 0.5
        st_out = torch.empty(M, N // 128, 2, device="cuda")
        sc = torch.randn(N, 2, device="cuda")
        kw = dict(epilogue=epi, stats_out=st_out)
        if epi == 3:
            kw.update(residual=res, gate=gate, gate_stride=N, rows_per_gate=1560)
        if use_ln:
            kw.update(ln_stats=st_in, ln_sc=sc, ln_eps=1e-6)
        bias = None if use_ln else b
        med, best = timeit(lambda: ops.gemm(x, w, bias, y, **kw), flush=flush)
        fl = 2.0 * M * N * K
        rec = dict(kernel="gemm_" + name, M=M, N=N, K=K, ms=med, ms_best=best, tflops=fl / med / 1e9)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    for name, Lq, S, H in [("self_S4680", 4680, 4680, 12), ("self_S18720", 4680, 18720, 12),
                           ("self_S32760", 4680, 32760, 12), ("cross_S512", 4680, 512, 12),
                           ("frame_S32760", 1560, 32760, 12)]:
        if only and not any(o in "attn_" + name for o in only):
            continue
        q = torch.randn(1, Lq, H, 128, device="cuda").to(BF)
        k = torch.randn(1, S, H, 128, device="cuda").to(BF)
        v = torch.randn(1, S, H, 128, device="cuda").to(BF)
        o = torch.empty_like(q)
        med, best = timeit(lambda: ops.attention(q, k, v, o, 1 / math.sqrt(128)), flush=flush, reps=5)
        fl = 4.0 * Lq * S * H * 128
        rec = dict(kernel="attn_" + name, Lq=Lq, S=S, H=H, ms=med, ms_best=best, tflops=fl / med / 1e9)
        try:
            tmed, _ = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(
                q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)), flush=flush, reps=5)
            rec.update(sdpa_ms=tmed, sdpa_tflops=fl / tmed / 1e9)
        except Exception as e:  # pragma: no cover
            rec["sdpa_error"] = str(e)[:100]
        try:
            from flash_attn import flash_attn_func
            tmed, _ = timeit(lambda: flash_attn_func(q, k, v), flush=flush, reps=5)
            rec.update(fa2_ms=tmed, fa2_tflops=fl / tmed / 1e9)
        except Exception as e:  # pragma: no cover
            rec["fa2_error"] = str(e)[:100]
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if not only or any(o in "elementwise" for o in only):
        from self_forcing_b200.model import rope_tables
        L, C = 4680, 1536
        x = torch.randn(L, C, device="cuda").to(BF); y = torch.empty_like(x)
        qkv = torch.randn(L, 3 * C, device="cuda").to(BF)
        tab = torch.randn(3, 6, C, device="cuda").to(BF)
        w = torch.randn(C, device="cuda").to(BF); b = torch.randn(C, device="cuda").to(BF)
        cos, sin = (t.cuda() for t in rope_tables(128))
        qo = torch.empty(1, L, C, device="cuda", dtype=BF)
        kc = torch.empty(1, L, 12, 128, device="cuda", dtype=BF); vc = torch.empty_like(kc)
        cases = [
            ("ln_modulate", 2 * L * C * 2, lambda: ops.ln_modulate(x, y, shift=tab[:, 0], scale=tab[:, 1], mod_stride=6 * C, rows_per_mod=1560, eps=1e-6)),
            ("ln_affine", 2 * L * C * 2, lambda: ops.ln_affine(x, y, w, b, 1e-6)),
            ("rmsnorm", 2 * L * C * 2, lambda: ops.rmsnorm(x, y, w, 1e-6)),
            ("qk_norm_rope", 6 * L * C * 2, lambda: ops.qk_norm_rope(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], w, w, 1e-6, cos, sin, 1, L, 128, (3, 30, 52), 0, q_out=qo, k_out=kc, v_out=vc)),
        ]
        x_stats = torch.rand(L, 12, 2, device="cuda") + 0.5
        cases.append(("ln_modulate_stats", 2 * L * C * 2, lambda: ops.ln_modulate(x, y, shift=tab[:, 0], scale=tab[:, 1], mod_stride=6 * C,
                                                                                rows_per_mod=1560, eps=1e-6, stats=x_stats)))
        qkv_stats = torch.rand(L, 36, 2, device="cuda") + 0.5
        cases.append(("qk_norm_rope_stats", 6 * L * C * 2, lambda: ops.qk_norm_rope(
            qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], w, w, 1e-6, cos, sin, 1, L, 128, (3, 30, 52), 0, q_out=qo, k_out=kc,
            v_out=vc, stats=qkv_stats, q_chunk0=0, k_chunk0=12)))
        for name, nbytes, fn in cases:
            for _ in range(3): fn()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50): fn()
            e.record(); torch.cuda.synchronize()
            us = a.elapsed_time(e) / 50 * 1e3
            med, _ = timeit(fn, flush=flush)
            rec = dict(kernel="ew_" + name, us_back_to_back=us, gbs_back_to_back=nbytes / us / 1e3, us_cold=med * 1e3,
                       gbs_cold=nbytes / med / 1e6)
            print(json.dumps(rec), flush=True)
            out.append(rec)
    tag = os.environ.get("SFB_MICROBENCH_TAG", "")
    with open(os.path.join(ROOT, "gpurun_out", f"microbench{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
