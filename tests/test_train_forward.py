"""the training-time (cache-free) forward with block masks, forward only.

* the oracle (oracle/train_oracle.py) against the unmodified reference `_forward_train` (FlexAttention) -- golden
  made by oracle/make_golden.py, bit-exact on the generating host;
* the product's dense attention windows against the reference's masks (integer work: exact);
* the product forward (host schedule with the CPU test double; CUDA kernels under -m gpu) against the same golden.
"""
import pytest
import torch

from _torch_ops import TorchOps
from helpers import golden, rel_l2
from oracle import causal_wan_oracle as O
from oracle import train_oracle as T
from oracle.make_golden import TRAIN, TRAIN_CASES, train_cfg, train_inputs

TOL = 1e-2


def _run_oracle(case):
    cfg = train_cfg(case)
    x, clean, t, aug, ctx = train_inputs(case)
    with torch.no_grad():
        return T.train_forward(O.make_random_params(cfg, seed=13), cfg, x, t, ctx, case["num_frame_per_block"],
                               clean_x=clean if case["tf"] else None, aug_t=aug if case["tf"] else None,
                               independent_first_frame=case["independent_first_frame"])


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_train_oracle_matches_reference_golden(name):
    g = golden("train_forward_tiny.pt")[name]
    out = _run_oracle(g["case"])
    assert out.shape == g["flow"].shape and rel_l2(out, g["flow"]) <= 2e-3      # bit-exact where the golden was made


def _product_model(case, device, ops=None):
    from self_forcing_b200.model import B200CausalWanModel
    r = TRAIN
    m = B200CausalWanModel(dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                           text_dim=r["text_dim"], local_attn_size=case["local_attn_size"], ops=ops).to(device).to(torch.bfloat16)
    m.load_state_dict(O.make_random_params(train_cfg(case), seed=13), strict=True)
    m.num_frame_per_block = case["num_frame_per_block"]
    m.independent_first_frame = case["independent_first_frame"]
    return m


@pytest.mark.parametrize("name,frames,fs", [(n, f, s) for n in TRAIN_CASES for f, s in ((TRAIN_CASES[n]["frames"], 24), (21, 40))])
def test_attention_windows_equal_the_reference_masks(name, frames, fs):
    """The rectangles the product executes, replayed on a boolean grid (with the staging copy of the teacher-forcing
    walk), must reproduce the reference's mask_mod exactly for every query row."""
    case = dict(TRAIN_CASES[name], frames=frames)
    if case["independent_first_frame"] and (frames - 1) % case["num_frame_per_block"]:
        frames += 1
    m = _product_model(case, "cpu", ops=TorchOps())
    total = frames * fs
    n = total * (2 if case["tf"] else 1)
    ref = T.attention_mask(frames, fs, case["num_frame_per_block"], case["local_attn_size"], case["independent_first_frame"],
                           case["tf"])[:n, :n]
    key_id = torch.arange(n)                  # which logical key sits in each K/V row (the staging copies move keys)
    seen = torch.zeros(n, n, dtype=torch.bool)
    done = torch.zeros(n, dtype=torch.bool)
    for q_lo, q_hi, kv_lo, kv_hi in m.attention_windows(frames, fs, case["tf"]):
        if kv_lo < 0:
            dst = -kv_lo - 1
            key_id[dst:dst + (q_hi - q_lo)] = key_id[total + dst:total + dst + (q_hi - q_lo)].clone()
            kv_lo = 0
        assert not done[q_lo:q_hi].any()      # every query row is computed exactly once
        done[q_lo:q_hi] = True
        seen[q_lo:q_hi, key_id[kv_lo:kv_hi]] = True
    assert done.all() and torch.equal(seen, ref)


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_train_forward_host_schedule_matches_reference_golden(name):
    g = golden("train_forward_tiny.pt")[name]
    case = g["case"]
    m = _product_model(case, "cpu", ops=TorchOps())
    x, clean, t, aug, ctx = train_inputs(case)
    kw = dict(clean_x=clean, aug_t=aug) if case["tf"] else {}
    out = m(x, t=t, context=list(ctx), seq_len=x.shape[2] * 24, **kw)
    assert out.shape == g["flow"].shape and rel_l2(out, g["flow"]) <= TOL


def test_wrapper_routes_cache_free_calls_to_the_train_forward():
    """WanDiffusionWrapper.forward without kv_cache on the causal model (wan_wrapper.py:301-337) -> (flow, x0)."""
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper
    r, case = TRAIN, TRAIN_CASES["teacher_forcing"]
    cfgd = dict(WAN_T2V_1_3B, dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"], text_dim=r["text_dim"])
    w = B200DiffusionWrapper(model_config=cfgd, timestep_shift=5.0, ops=TorchOps())
    w.model.load_state_dict(O.make_random_params(train_cfg(case), seed=13), strict=True)
    w.model.num_frame_per_block = case["num_frame_per_block"]
    x, clean, t, aug, ctx = train_inputs(case)
    flow, x0 = w(x.permute(0, 2, 1, 3, 4), {"prompt_embeds": ctx}, t, clean_x=clean.permute(0, 2, 1, 3, 4), aug_t=aug)
    g = golden("train_forward_tiny.pt")["teacher_forcing"]
    assert rel_l2(flow.permute(0, 2, 1, 3, 4), g["flow"]) <= TOL
    sched = O.OracleScheduler(5.0)
    ref_x0 = O.flow_to_x0(sched, flow.flatten(0, 1), x.permute(0, 2, 1, 3, 4).flatten(0, 1), t.flatten(0, 1)).unflatten(0, flow.shape[:2])
    assert torch.equal(x0, ref_x0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_train_forward_on_gpu_matches_reference_golden(name):
    g = golden("train_forward_tiny.pt")[name]
    case = g["case"]
    m = _product_model(case, "cuda")
    x, clean, t, aug, ctx = (v.cuda() for v in train_inputs(case))
    kw = dict(clean_x=clean, aug_t=aug) if case["tf"] else {}
    out = m(x, t=t, context=list(ctx), seq_len=x.shape[2] * 24, **kw)
    assert rel_l2(out.cpu(), g["flow"]) <= TOL
    assert torch.equal(out, m(x, t=t, context=ctx, seq_len=x.shape[2] * 24, **kw))     # workspace reuse, deterministic


@pytest.mark.gpu
def test_train_forward_full_width_vs_oracle_on_gpu():
    """1.3B width (C = 1536, 12 heads, 1560 tokens / frame), 2 layers, 6 frames in chunks of 3 with teacher forcing:
    product (CUDA kernels, 18720-token sequence) vs the training oracle run on the same GPU."""
    from self_forcing_b200.model import B200CausalWanModel
    cfg = O.OracleConfig(dim=1536, ffn_dim=512, num_heads=12, num_layers=2)
    params = {k: v.cuda() for k, v in O.make_random_params(cfg, seed=3).items()}
    m = B200CausalWanModel(dim=1536, ffn_dim=512, num_heads=12, num_layers=2).to("cuda").to(torch.bfloat16)
    m.load_state_dict(params, strict=True)
    m.num_frame_per_block = 3
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(1, 16, 6, 60, 104, generator=g, device="cuda").to(torch.bfloat16)
    clean = torch.randn(1, 16, 6, 60, 104, generator=g, device="cuda").to(torch.bfloat16)
    ctx = torch.randn(1, 512, 4096, generator=g, device="cuda").to(torch.bfloat16)
    t = torch.tensor([[937.5] * 3 + [625.0] * 3], device="cuda")
    for kw in ({}, dict(clean_x=clean, aug_t=torch.full_like(t, 20.0))):
        out = m(x, t=t, context=ctx, seq_len=32760, **kw)
        with torch.no_grad():
            ref = T.train_forward(params, cfg, x, t, ctx, 3, **kw)
        err = rel_l2(out, ref)
        print("train forward full width", "tf" if kw else "causal", err)
        assert err <= TOL, err
