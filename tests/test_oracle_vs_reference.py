"""Live pinning of the oracle against the unmodified reference (build container only: the reference
checkout does not travel to the GPU box, where these tests skip)."""
import contextlib
import io

import pytest
import torch

from helpers import ROLLOUT_CASES, patched_randn_like, synthetic_inputs
from oracle import causal_wan_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present")


def test_single_forward_bit_exact():
    ref = ref_shim.load_reference()
    cfg = O.OracleConfig(**O.WAN_TINY)
    params = O.make_random_params(cfg, seed=3)
    w = ref_shim.make_reference_wrapper(ref, cfg.reference_kwargs(), 5.0)
    w.model.load_state_dict(params, strict=False)
    ow = O.OracleWrapper(params, cfg, 5.0)
    pe, noise = synthetic_inputs(1, 1)
    t = torch.full((1, 1), 937.5)
    kv_a, ca_a = O.new_kv_cache(cfg, 1, 1560, torch.bfloat16, "cpu", 3120), O.new_crossattn_cache(cfg, 1, torch.bfloat16, "cpu")
    kv_b, ca_b = O.new_kv_cache(cfg, 1, 1560, torch.bfloat16, "cpu", 3120), O.new_crossattn_cache(cfg, 1, torch.bfloat16, "cpu")
    with torch.no_grad():
        f1, x1 = w(noise, {"prompt_embeds": pe}, t, kv_cache=kv_a, crossattn_cache=ca_a, current_start=0)
        f2, x2 = ow(noise, pe, t, kv_b, ca_b, 0)
    assert torch.equal(f1, f2) and torch.equal(x1, x2)
    for a, b in zip(kv_a, kv_b):
        assert torch.equal(a["k"], b["k"]) and torch.equal(a["v"], b["v"])
        assert int(a["global_end_index"]) == int(b["global_end_index"]) == 1560


def test_reference_sdpa_vs_fp32_softmax():
    """flash_attn / SDPA arithmetic lives outside the reference; cross-check the restated attention against
    an explicit fp32 softmax (third-party arithmetic (i))."""
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, 40, 2, 128, generator=g).to(torch.bfloat16) for _ in range(3))
    o = O.dense_attention(q, k, v).float()
    s = torch.einsum("blhd,bshd->bhls", q.float(), k.float()) / 128 ** 0.5
    ref = torch.einsum("bhls,bshd->blhd", torch.softmax(s, -1), v.float())
    assert float((o - ref).norm() / ref.norm()) < 1e-2


def test_bidirectional_reference_equals_golden_and_oracle():
    """Container only: re-run the unmodified bidirectional WanModel and compare with the committed golden + oracle."""
    import torch
    from helpers import golden
    from oracle import causal_wan_oracle as O
    from oracle import make_golden as G
    ref = ref_shim.load_reference()
    fresh = G.reference_bidirectional(ref)
    g = golden("bidirectional_tiny.pt")
    assert torch.equal(fresh["flow"], g["flow"])
    x, t, ctx = G.bidirectional_inputs()
    cfg = G.bidirectional_cfg()
    with torch.no_grad():
        assert torch.equal(O.bidirectional_forward(O.make_random_params(cfg, seed=9), cfg, x, t, ctx), g["flow"])
