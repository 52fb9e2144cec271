"""The CPU oracle against the golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

from helpers import ROLLOUT_CASES, golden, initial_latent_for, patched_randn_like, rel_l2, synthetic_inputs
from oracle import causal_wan_oracle as O
from oracle.make_golden import MASK_CASES, ROLLING, rolling_cfg, rolling_model_inputs

# bf16 CPU matmuls are bit-reproducible on the machine that made the fixtures; another host ISA
# (AMX vs AVX-512 accumulation order) may flip last bits, hence a tolerance instead of equality.
TOL = 1e-2


def test_scheduler_tables_exact():
    g = golden("scheduler_tables.pt")
    for shift, ref in g.items():
        s = O.OracleScheduler(shift)
        assert torch.equal(s.sigmas, ref["sigmas"])
        assert torch.equal(s.timesteps, ref["timesteps"])
        assert torch.equal(O.warp_denoising_steps(s, [1000, 750, 500, 250]), ref["warped"])
        x0 = torch.randn(3, 16, 8, 8, generator=torch.Generator().manual_seed(7)).to(torch.bfloat16)
        nz = torch.randn(3, 16, 8, 8, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
        assert torch.equal(s.add_noise(x0, nz, ref["warped"][1:].clone()), ref["add_noise_out"])
    # known answers probed from the reference
    s5 = O.OracleScheduler(5.0)
    assert O.warp_denoising_steps(s5, [1000, 750, 500, 250]).tolist() == pytest.approx([1000, 937.5, 833.3333, 625], rel=1e-6)
    assert float(s5.sigmas[999]) == pytest.approx(0.0049800803, rel=1e-6)
    assert float(O.OracleScheduler(8.0).sigmas[999]) == pytest.approx(0.0079443902, rel=1e-6)


@pytest.mark.parametrize("name", list(ROLLOUT_CASES))
def test_rollout_matches_reference_golden(name):
    g = golden("rollout_tiny.pt")[name]
    case = g["case"]
    cfg = O.OracleConfig(**O.WAN_TINY)
    params = O.make_random_params(cfg, seed=0)
    ow = O.OracleWrapper(params, cfg, case["shift"])
    pe, noise = synthetic_inputs(1, case["frames"])
    steps = O.warp_denoising_steps(ow.scheduler, [1000, 750, 500, 250])
    with torch.no_grad(), patched_randn_like(3):
        tr = O.rollout(ow, noise, pe, steps, case["num_frame_per_block"],
                       independent_first_frame=case["independent_first_frame"], initial_latent=initial_latent_for(case))
    assert tr.index_trace[-1] == tuple(g["final_index"])          # integers: bit-exact
    assert rel_l2(tr.latents, g["latents"]) <= TOL


def test_rolling_window_rollout_matches_reference_golden():
    """The rolling KV window driven through the pipeline: 2-frame local window + 1-frame sink, four one-frame chunks,
    so the cache rolls on chunks 3 and 4 while the global index keeps growing."""
    from oracle.make_golden import ROLLING_ROLLOUT_CASES, rollout_cfg
    g = golden("rollout_rolling.pt")["rolling_window"]
    case = ROLLING_ROLLOUT_CASES["rolling_window"]
    assert g["case"] == case
    cfg = rollout_cfg(case)
    ow = O.OracleWrapper(O.make_random_params(O.OracleConfig(**O.WAN_TINY), seed=0), cfg, case["shift"])
    pe, noise = synthetic_inputs(1, case["frames"])
    steps = O.warp_denoising_steps(ow.scheduler, [1000, 750, 500, 250])
    with torch.no_grad(), patched_randn_like(3):
        tr = O.rollout(ow, noise, pe, steps, case["num_frame_per_block"])
    assert tr.index_trace[4::5] == [(1560, 1560), (3120, 3120), (4680, 3120), (6240, 3120)]   # after each chunk's refresh
    assert tr.index_trace[-1] == tuple(g["final_index"]) == (6240, 3120)
    assert rel_l2(tr.latents, g["latents"]) <= TOL


def test_rolling_sink_cache_model_level():
    g = golden("model_rolling.pt")
    r = ROLLING
    cfg = rolling_cfg()
    params = O.make_random_params(cfg, seed=5)
    x, ctx = rolling_model_inputs()
    ft = (r["frame_hw"][0] // 2) * (r["frame_hw"][1] // 2)
    kv = O.new_kv_cache(cfg, 1, ft, torch.bfloat16, "cpu", cache_tokens=r["local_attn_size"] * ft)
    ca = O.new_crossattn_cache(cfg, 1, torch.bfloat16, "cpu")
    flows, trace = [], []
    with torch.no_grad():
        for c in range(r["chunks"]):
            for k in range(r["forwards_per_chunk"]):
                t = torch.full((1, 1), 1000.0 - 300.0 * k)
                flows.append(O.model_forward(params, cfg, x[:, :, c:c + 1], t, ctx, kv, ca, c * ft))
                trace.append((int(kv[0]["global_end_index"]), int(kv[0]["local_end_index"])))
    assert trace == [tuple(t) for t in g["trace"]]
    assert rel_l2(torch.stack(flows), g["flows"]) <= TOL
    for i in range(cfg.num_layers):
        assert rel_l2(kv[i]["k"], g["k"][i]) <= TOL and rel_l2(kv[i]["v"], g["v"][i]) <= TOL


@pytest.mark.parametrize("name", list(MASK_CASES))
def test_block_mask_tables(name):
    kind, kw = MASK_CASES[name]
    ref = golden("block_masks.pt")[name]
    if kind == "tf":
        m = O.teacher_forcing_mask(kw["num_frames"], kw["frame_seqlen"], kw["num_frame_per_block"])
    else:
        m = O.blockwise_causal_mask(kw["num_frames"], kw["frame_seqlen"], kw["num_frame_per_block"],
                                    kw.get("local_attn_size", -1), independent_first_frame=(kind == "i2v"))
    any_, all_ = O.block_table(m)
    partial = any_ & ~all_
    assert torch.equal(partial.sum(1).int(), ref["kv_num_blocks"])
    assert torch.equal(all_.sum(1).int(), ref["full_kv_num_blocks"])
    for row in range(partial.shape[0]):
        n = int(ref["kv_num_blocks"][row])
        assert torch.nonzero(partial[row]).flatten().tolist() == ref["kv_indices"][row, :n].tolist()
        n = int(ref["full_kv_num_blocks"][row])
        assert torch.nonzero(all_[row]).flatten().tolist() == ref["full_kv_indices"][row, :n].tolist()


def test_cache_plan_known_answers():
    """SURVEY.md section 8c KAT (ii): local window 6 frames x 4 tokens (cache 24), chunks of 12 tokens, two
    forwards per chunk, sink 0 and 1."""
    for sink in (0, 1):
        g = l = 0
        trace = []
        for chunk in range(5):
            for _ in range(2):
                p = O.plan_cache_update(g, l, chunk * 12, 12, 24, 6, sink * 4, 6 * 1560)
                g, l = p.global_end, p.local_end
                trace.append((g, l))
                if chunk == 2 and sink == 1 and p.roll:
                    assert (p.roll_src, p.roll_dst, p.roll_len) == (16, 4, 8)
        assert trace == [(12, 12), (12, 12), (24, 24), (24, 24), (36, 24), (36, 24), (48, 24), (48, 24), (60, 24), (60, 24)]


def test_bidirectional_forward_matches_reference_golden():
    """BASELINE config 5 (teacher forward): the oracle's bidirectional restatement equals the unmodified reference
    WanModel bit for bit."""
    from oracle.make_golden import bidirectional_cfg, bidirectional_inputs
    g = golden("bidirectional_tiny.pt")
    x, t, ctx = bidirectional_inputs()
    cfg = bidirectional_cfg()
    with torch.no_grad():
        out = O.bidirectional_forward(O.make_random_params(cfg, seed=9), cfg, x, t, ctx)
    assert torch.equal(out, g["flow"])


def test_bidirectional_padded_samples_match_reference_golden():
    """Samples shorter than seq_len (model.py:684-693 zero-pads them and attention gets k_lens): the valid tokens of a
    sample only see that sample's own tokens, so the exact-length restatement must reproduce the reference's padded
    run (golden made with flash_attn's k_lens semantics restated over SDPA, oracle/make_golden.py)."""
    from oracle.make_golden import bidirectional_cfg, bidirectional_inputs
    g = golden("bidirectional_padded.pt")
    x, t, ctx = bidirectional_inputs()
    cfg = bidirectional_cfg()
    ctx = torch.cat([ctx[:, :g["context_rows"]], torch.zeros_like(ctx[:, g["context_rows"]:])], dim=1)   # model.py:704-709
    with torch.no_grad():
        out = O.bidirectional_forward(O.make_random_params(cfg, seed=9), cfg, x, t, ctx)
    assert rel_l2(out, g["flow"]) <= 2e-3      # masked vs unmasked SDPA kernels round differently
