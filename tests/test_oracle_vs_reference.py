"""Live pinning of the oracle against the unmodified reference (build container only: the reference
checkout does not travel to the GPU box, where these tests skip)."""
import pytest
import torch

from helpers import synthetic_inputs
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


@pytest.mark.parametrize("steps,shift,order,dtype", [(50, 8.0, 2, torch.bfloat16), (9, 3.0, 2, torch.float32),
                                                      (12, 5.0, 1, torch.bfloat16), (3, 1.0, 2, torch.float32)])
def test_unipc_oracle_equals_live_reference(steps, shift, order, dtype):
    """Beyond the committed trace: other step counts, shifts, solver order 1, both dtypes -- bit for bit."""
    from oracle import unipc_oracle as U
    from oracle.make_golden import unipc_trace_flow
    ref = ref_shim.load_reference()
    r = ref.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, solver_order=order)
    r.set_timesteps(steps, device="cpu", shift=shift)
    o = U.OracleUniPC(solver_order=order)
    o.set_timesteps(steps, shift)
    assert torch.equal(r.timesteps, o.timesteps) and torch.equal(r.sigmas, o.sigmas)
    xr = xo = torch.randn(1, 2, 16, 6, 8, generator=torch.Generator().manual_seed(steps)).to(dtype)
    for i, t in enumerate(r.timesteps):
        xr = r.step(unipc_trace_flow(xr, i), t, xr, return_dict=False)[0]
        xo = o.step(unipc_trace_flow(xo, i), t, xo)
        assert torch.equal(xr, xo), i


@pytest.mark.parametrize("frames,hw,seed", [(1, (2, 4), 1), (4, (6, 4), 2)])
def test_vae_oracle_equals_live_reference(frames, hw, seed):
    """Other grids / frame counts than the committed fixture, including a single-frame video (no temporal upsampling at
    all) and a different weight draw."""
    from oracle import vae_oracle as V
    rv = ref_shim.load_reference_vae()
    cfg = V.VaeConfig()
    params = V.make_random_vae_params(cfg, seed=seed)
    model = rv.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                       temperal_downsample=[False, True, True], dropout=0.0)
    model.load_state_dict(params, strict=False)
    model = model.to(torch.bfloat16).eval()
    z = torch.randn(1, 16, frames, *hw, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16)
    scale = [torch.tensor(V.LATENT_MEAN).to(torch.bfloat16), 1.0 / torch.tensor(V.LATENT_STD).to(torch.bfloat16)]
    with torch.no_grad():
        want = model.decode(z, scale)
        got = V.decode(params, cfg, z)
    assert want.shape == (1, 3, 1 + 4 * (frames - 1), 8 * hw[0], 8 * hw[1])
    assert torch.equal(got, want)
