"""Host-side logic of the product on CPU: cache index arithmetic, scheduler tables, mask tables and the
whole model/wrapper/pipeline orchestration driven through the TorchOps test double (tests/_torch_ops.py)
and compared with the oracle.  No CUDA kernel runs here."""
import os
import numpy as np
import pytest
import torch

from _torch_ops import TorchOps
from helpers import ROLLOUT_CASES, golden, initial_latent_for, make_product_pipeline, patched_randn_like, rel_l2
from oracle import causal_wan_oracle as O
from oracle.make_golden import MASK_CASES, ROLLING, rolling_cfg, rolling_model_inputs
from self_forcing_b200 import masks
from self_forcing_b200.cache import IndexMirror, plan_cache_update
from self_forcing_b200.scheduler import FlowMatchScheduler


def test_cache_plan_matches_oracle_exhaustively():
    """Product and oracle index arithmetic are separate restatements of causal_model.py:195-236: sweep them
    against each other over rolling / sink / re-denoise sequences."""
    for cache, local, sink, new in [(24, 6, 0, 12), (24, 6, 1, 12), (30, 5, 2, 6), (32760, -1, 0, 4680), (72, 3, 1, 24)]:
        g = l = 0
        for chunk in range(9):
            if local == -1 and (chunk + 1) * new > cache:
                break
            for rep in range(3):
                a = plan_cache_update(g, l, chunk * new, new, cache, local, sink * 4, 7 * 4)
                b = O.plan_cache_update(g, l, chunk * new, new, cache, local, sink * 4, 7 * 4)
                assert (a.roll, a.roll_src, a.roll_dst, a.roll_len, a.write_start, a.write_end, a.attn_start,
                        a.attn_end, a.global_end, a.local_end) == \
                       (b.roll, b.roll_src, b.roll_dst, b.roll_len, b.write_start, b.write_end, b.attn_start,
                        b.attn_end, b.global_end, b.local_end)
                g, l = a.global_end, a.local_end


def test_cache_plan_overflow_raises():
    with pytest.raises(ValueError):
        plan_cache_update(32760, 32760, 32760, 4680, 32760, -1, 0, 32760)


def test_index_mirror_tracks_rebinding():
    kv = [dict(global_end_index=torch.tensor([0]), local_end_index=torch.tensor([0])) for _ in range(3)]
    m = IndexMirror()
    assert m.read(kv) == [(0, 0)] * 3
    m.write(kv, [(12, 12)] * 3)
    assert int(kv[1]["global_end_index"]) == 12 and m.read(kv) == [(12, 12)] * 3
    kv[1]["global_end_index"] = torch.tensor([5])       # the pipeline resets by rebinding
    kv[1]["local_end_index"] = torch.tensor([4])
    assert m.read(kv) == [(12, 12), (5, 4), (12, 12)]


def test_scheduler_tables_match_golden():
    g = golden("scheduler_tables.pt")
    for shift, ref in g.items():
        s = FlowMatchScheduler(shift=shift, sigma_min=0.0, extra_one_step=True)
        s.set_timesteps(1000, training=True)
        assert torch.equal(s.sigmas, ref["sigmas"]) and torch.equal(s.timesteps, ref["timesteps"])


def test_scheduler_add_noise_through_double():
    g = golden("scheduler_tables.pt")[5.0]
    s = FlowMatchScheduler(shift=5.0, ops=TorchOps())
    s.set_timesteps(1000, training=True)
    x0 = torch.randn(3, 16, 8, 8, generator=torch.Generator().manual_seed(7)).to(torch.bfloat16)
    nz = torch.randn(3, 16, 8, 8, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    assert torch.equal(s.add_noise(x0, nz, g["warped"][1:].clone()), g["add_noise_out"])


@pytest.mark.parametrize("name", list(MASK_CASES))
def test_mask_tables_bit_exact(name):
    kind, kw = MASK_CASES[name]
    ref = golden("block_masks.pt")[name]
    r = masks.teacher_forcing_tables(**kw) if kind == "tf" else \
        masks.blockwise_causal_tables(lone_first_frame=(kind == "i2v"), **kw)
    for k in ("kv_num_blocks", "full_kv_num_blocks", "kv_indices", "full_kv_indices"):
        assert np.array_equal(r[k], ref[k].numpy().astype(np.int32)), k
    assert r["sparsity"] == pytest.approx(ref["sparsity"], abs=1e-4)


def test_mask_known_answer_21_frames():
    """causal 21 frames, 3 per block -> sparsity 42.52 %, full_kv_num_blocks[0] = 36."""
    r = masks.blockwise_causal_tables(21, 1560, 3)
    assert round(r["sparsity"], 2) == 42.52 and int(r["full_kv_num_blocks"][0]) == 36


@pytest.mark.parametrize("name", list(ROLLOUT_CASES))
def test_pipeline_orchestration_fp32_exact(name):
    """fp32 on both sides removes rounding noise: any wiring mistake (wrong modulation row, cache slot,
    frame offset, chunk order) shows up as a large error; a correct schedule reproduces the oracle."""
    case = ROLLOUT_CASES[name]
    ops = TorchOps()
    pipe, cfg, params, pe, noise = make_product_pipeline(case, "cpu", ops=ops, dtype=torch.float32, hw=(16, 24))
    init = initial_latent_for(case, 16, 24)
    init = None if init is None else init.float()
    with patched_randn_like(3):
        _, lat = pipe.inference(noise, ["synthetic"], return_latents=True, initial_latent=init)
    ow = O.OracleWrapper(params, cfg, case["shift"])
    steps = O.warp_denoising_steps(ow.scheduler, [1000, 750, 500, 250])
    with torch.no_grad(), patched_randn_like(3):
        tr = O.rollout(ow, noise, pe, steps, case["num_frame_per_block"],
                       independent_first_frame=case["independent_first_frame"], initial_latent=init)
    assert rel_l2(lat, tr.latents) < 1e-5
    assert (int(pipe.kv_cache1[0]["global_end_index"]), int(pipe.kv_cache1[-1]["local_end_index"])) == tr.index_trace[-1]
    assert ops.launches > 0


def test_pipeline_orchestration_batch2_fp32():
    """Two different prompts / noise samples in one batch (BASELINE config 4 runs B > 1 per GPU when memory allows):
    per-sample cache rows, per-sample text K/V and [B, F] timesteps must stay separate -- fp32 against the oracle."""
    case = ROLLOUT_CASES["chunkwise"]
    pipe, cfg, params, _, _ = make_product_pipeline(case, "cpu", ops=TorchOps(), dtype=torch.float32, hw=(16, 24))
    g = torch.Generator().manual_seed(12)
    pe = torch.randn(2, 512, 4096, generator=g)
    noise = torch.randn(2, case["frames"], 16, 16, 24, generator=g)
    from helpers import _TextEncoder
    pipe.text_encoder = _TextEncoder(pe)
    with patched_randn_like(3):
        _, lat = pipe.inference(noise, ["a", "b"], return_latents=True)
    ow = O.OracleWrapper(params, cfg, case["shift"])
    steps = O.warp_denoising_steps(ow.scheduler, [1000, 750, 500, 250])
    with torch.no_grad(), patched_randn_like(3):
        tr = O.rollout(ow, noise, pe, steps, case["num_frame_per_block"])
    assert rel_l2(lat, tr.latents) < 1e-5
    assert rel_l2(lat[0], lat[1]) > 0.5                          # the two samples really differ
    assert pipe.kv_cache1[0]["k"].shape[0] == 2


def test_pipeline_bf16_vs_reference_golden():
    case = ROLLOUT_CASES["tiny_test_yaml"]
    g = golden("rollout_tiny.pt")["tiny_test_yaml"]
    pipe, *_ , noise = make_product_pipeline(case, "cpu", ops=TorchOps())
    with patched_randn_like(3):
        _, lat = pipe.inference(noise, ["synthetic"], return_latents=True)
    assert rel_l2(lat, g["latents"]) <= 1e-2
    assert (int(pipe.kv_cache1[0]["global_end_index"]), int(pipe.kv_cache1[0]["local_end_index"])) == tuple(g["final_index"])


def test_rolling_window_pipeline_vs_reference_golden():
    """Host cache plan + index mirror under a rolling local window (2 frames, sink 1) through the whole pipeline."""
    from helpers import ROLLING_ROLLOUT_CASES
    g = golden("rollout_rolling.pt")["rolling_window"]
    case = ROLLING_ROLLOUT_CASES["rolling_window"]
    pipe, *_, noise = make_product_pipeline(case, "cpu", ops=TorchOps())
    assert pipe.local_attn_size == 2
    with patched_randn_like(3):
        _, lat = pipe.inference(noise, ["synthetic"], return_latents=True)
    assert pipe.kv_cache1[0]["k"].shape[1] == 2 * 1560
    assert (int(pipe.kv_cache1[0]["global_end_index"]), int(pipe.kv_cache1[-1]["local_end_index"])) == tuple(g["final_index"])
    assert rel_l2(lat, g["latents"]) <= 1e-2


def test_skip_refresh_tail_is_output_neutral():
    case = dict(ROLLOUT_CASES["tiny_test_yaml"], frames=2, independent_first_frame=False)
    outs = []
    for skip in (False, True):
        pipe, *_, noise = make_product_pipeline(case, "cpu", ops=TorchOps(), dtype=torch.float32, hw=(16, 24),
                                                skip_refresh_tail=skip)
        with patched_randn_like(3):
            outs.append(pipe.inference(noise, ["synthetic"], return_latents=True)[1])
    assert torch.equal(outs[0], outs[1])


def test_rolling_sink_model_matches_golden():
    from self_forcing_b200.model import B200CausalWanModel
    g = golden("model_rolling.pt")
    r = ROLLING
    cfg = rolling_cfg()
    params = O.make_random_params(cfg, seed=5)
    model = B200CausalWanModel(dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                               text_dim=r["text_dim"], local_attn_size=r["local_attn_size"], sink_size=r["sink_size"],
                               ops=TorchOps()).to(torch.bfloat16)
    model.load_state_dict(params, strict=True)
    x, ctx = rolling_model_inputs()
    ft = (r["frame_hw"][0] // 2) * (r["frame_hw"][1] // 2)
    kv = O.new_kv_cache(cfg, 1, ft, torch.bfloat16, "cpu", cache_tokens=r["local_attn_size"] * ft)
    ca = O.new_crossattn_cache(cfg, 1, torch.bfloat16, "cpu")
    flows, trace = [], []
    for c in range(r["chunks"]):
        for k in range(r["forwards_per_chunk"]):
            t = torch.full((1, 1), 1000.0 - 300.0 * k)
            flows.append(model(x[:, :, c:c + 1], t=t, context=ctx, seq_len=32760, kv_cache=kv, crossattn_cache=ca,
                               current_start=c * ft).clone())
            trace.append((int(kv[0]["global_end_index"]), int(kv[0]["local_end_index"])))
    assert trace == [tuple(t) for t in g["trace"]]
    assert rel_l2(torch.stack(flows), g["flows"]) <= 1e-2
    for i in range(cfg.num_layers):
        assert rel_l2(kv[i]["k"], g["k"][i]) <= 1e-2 and rel_l2(kv[i]["v"], g["v"][i]) <= 1e-2


def test_state_dict_is_reference_compatible():
    from self_forcing_b200.model import B200CausalWanModel
    cfg = O.OracleConfig(dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_dim=512)
    m = B200CausalWanModel(dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_dim=512, ops=TorchOps())
    assert set(m.state_dict().keys()) == set(O.parameter_shapes(cfg).keys())
    sd = O.make_random_params(cfg)
    sd["pose_proj.weight"] = torch.zeros(256, 5120)      # fork-specific extra keys are tolerated
    sd["pose_proj.bias"] = torch.zeros(256)
    m.load_state_dict(sd, strict=True)


def test_cache_free_call_dispatches_to_the_training_forward():
    """forward() without a kv_cache is the training-time forward (reference causal_model.py:1071-1079; parity in
    tests/test_train_forward.py); pose / image conditioning still raises."""
    from self_forcing_b200.model import B200CausalWanModel
    m = B200CausalWanModel(dim=256, ffn_dim=256, num_heads=2, num_layers=1, text_dim=512, ops=TorchOps()).to(torch.bfloat16)
    m.init_weights(0)
    x = torch.zeros(1, 16, 1, 8, 8, dtype=torch.bfloat16)
    out = m(x, t=torch.zeros(1, 1), context=torch.zeros(1, 512, 512, dtype=torch.bfloat16), seq_len=100)
    assert out.shape == x.shape
    with pytest.raises(NotImplementedError):
        m(x, t=torch.zeros(1, 1), context=torch.zeros(1, 512, 512, dtype=torch.bfloat16), seq_len=100, y=[x[0]])


# (the attention work schedule is checked against the kernel's own host/device functions in tests/test_attention_schedule.py)


def test_reference_arm_runs_on_rank0_only():
    """Under torchrun the CPU reference arm is rank 0's job; other ranks exit 0 without work or output."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_gpu_eager_leg_logic(monkeypatch):
    """bench.py's `gpu_eager_baseline` leg (the oracle in eager mode with the product's own weights) on a 2-layer
    model: the product's state_dict must be directly usable as the oracle's parameter dict."""
    import bench
    monkeypatch.setattr(bench, "NL", 2)
    monkeypatch.setattr(bench, "FFN", 512)
    case = dict(frames=2, num_frame_per_block=1, independent_first_frame=False, shift=bench.SHIFT)
    pipe, cfg, params, pe, noise = make_product_pipeline(case, "cpu", ops=TorchOps())
    out = bench.gpu_eager_leg(pipe.generator, 1, pe, noise, product_fps=10.0)
    assert out["finite"] and out["value"] > 0 and out["kind"] == "port"
    assert out["product_speedup"] == pytest.approx(10.0 / out["value"])


def test_bench_flop_model_matches_survey():
    """bench.py's algorithmic FLOP count is the SURVEY.md section 8d figure (990.3 TFLOP chunk-wise, 943.2 frame-wise)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert abs(b.rollout_flops(3) / 1e12 - 990.3) < 0.5
    assert abs(b.rollout_flops(1) / 1e12 - 943.2) < 0.5


def test_bidirectional_model_matches_reference_golden():
    """B200WanModel (host orchestration, CPU test double for the kernels): per-sample timesteps, scratch K/V shared by
    all layers, text K/V recomputed per call -> equals the reference WanModel golden."""
    from oracle.make_golden import BIDIR, bidirectional_cfg, bidirectional_inputs
    from self_forcing_b200.model import B200WanModel
    g = golden("bidirectional_tiny.pt")
    r = BIDIR
    model = B200WanModel(dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                         text_dim=r["text_dim"], ops=TorchOps()).to(torch.bfloat16)
    model.load_state_dict(O.make_random_params(bidirectional_cfg(), seed=9), strict=True)
    x, t, ctx = bidirectional_inputs()
    out = model(list(x), t=t, context=list(ctx), seq_len=g["seq_len"])
    assert out.shape == g["flow"].shape and rel_l2(out, g["flow"]) <= 1e-2
    out2 = model(x, t=t, context=ctx, seq_len=g["seq_len"])        # scratch buffers are reusable
    assert torch.equal(out, out2)
    # samples shorter than seq_len (the reference pads them and masks the padded keys): exact-length forward
    gp = golden("bidirectional_padded.pt")
    outp = model(list(x), t=t, context=[c[:gp["context_rows"]] for c in ctx], seq_len=gp["seq_len"])
    assert outp.shape == gp["flow"].shape and rel_l2(outp, gp["flow"]) <= 1e-2
    with pytest.raises(ValueError):
        model(x, t=t, context=ctx, seq_len=g["seq_len"] - 8)      # a sample longer than seq_len (reference asserts)
    # a one-frame sample under the same seq_len (its own token grid, hence its own RoPE positions)
    a = model([x[1][:, :1]], t=t[1:], context=[ctx[1]], seq_len=g["seq_len"])
    assert a.shape == (1, 16, 1, *x.shape[3:])


def test_bidirectional_wrapper_surface():
    """WanDiffusionWrapper(is_causal=False) surface (utils/wan_wrapper.py:253-349 without kv_cache): [B, F] uniform
    timesteps -> (flow_pred, pred_x0) with x0 = x_t - sigma_t * flow in float64."""
    from oracle.make_golden import BIDIR, bidirectional_cfg, bidirectional_inputs
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper
    g = golden("bidirectional_tiny.pt")
    r = BIDIR
    cfgd = dict(WAN_T2V_1_3B, dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                text_dim=r["text_dim"])
    w = B200DiffusionWrapper(model_config=cfgd, timestep_shift=5.0, is_causal=False, ops=TorchOps())
    w.model.load_state_dict(O.make_random_params(bidirectional_cfg(), seed=9), strict=True)
    w.seq_len = g["seq_len"]
    x, t, ctx = bidirectional_inputs()
    xt = x.permute(0, 2, 1, 3, 4)                                     # the wrapper takes [B, F, C, H, W]
    tt = t[:, None].expand(-1, xt.shape[1]).contiguous()
    flow, x0 = w(xt, {"prompt_embeds": ctx}, tt)
    assert rel_l2(flow.permute(0, 2, 1, 3, 4), g["flow"]) <= 1e-2
    sched = O.OracleScheduler(5.0)
    ref_x0 = O.flow_to_x0(sched, flow.flatten(0, 1), xt.flatten(0, 1), tt.flatten(0, 1)).unflatten(0, flow.shape[:2])
    assert torch.equal(x0, ref_x0)
    with pytest.raises(ValueError):
        w(xt, {"prompt_embeds": ctx}, torch.tensor([[937.5, 250.0], [250.0, 250.0]]))


def test_wrapper_loads_a_checkpoint_directory(tmp_path):
    """HF-style directory (config.json + weights) like CausalWanModel.from_pretrained (utils/wan_wrapper.py:139-145);
    fork-specific `pose_proj.*` keys are tolerated."""
    import json
    from self_forcing_b200.wrapper import B200DiffusionWrapper
    cfg = O.OracleConfig(dim=256, ffn_dim=256, num_heads=2, num_layers=1, text_dim=512)
    sd = O.make_random_params(cfg, seed=3)
    sd["pose_proj.weight"] = torch.zeros(256, 5120, dtype=torch.bfloat16)
    json.dump(dict(model_type="t2v", patch_size=[1, 2, 2], text_len=512, in_dim=16, dim=256, ffn_dim=256, freq_dim=256,
                   text_dim=512, out_dim=16, num_heads=2, num_layers=1, qk_norm=True, cross_attn_norm=True, eps=1e-6,
                   _class_name="CausalWanModel"), open(tmp_path / "config.json", "w"))
    torch.save(sd, tmp_path / "diffusion_pytorch_model.pth")
    w = B200DiffusionWrapper(model_path=str(tmp_path), timestep_shift=5.0, ops=TorchOps())
    got = w.model.state_dict()
    assert all(torch.equal(got[k], v) for k, v in sd.items() if not k.startswith("pose_proj"))
    assert w.model.num_layers == 1 and w.scheduler.shift == 5.0 and w.seq_len == 32760
    with pytest.raises(FileNotFoundError):
        B200DiffusionWrapper(model_path=str(tmp_path / "missing"), ops=TorchOps())


def test_cross_attention_norm_fold_matches_the_unfolded_schedule():
    """norm3 folded into the cross-attention q projection and norm_q into the softmax scale / cached K (model.py: `fold`):
    in fp32 the folded schedule is the unfolded one up to rounding, on a multi-chunk cached rollout."""
    from self_forcing_b200.model import B200CausalWanModel

    class StatsOps(TorchOps):
        supports_row_stats = True

    def run(ops):
        torch.manual_seed(0)
        m = B200CausalWanModel(dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_dim=512, ops=ops)
        m.init_weights(3)
        fs, F_ = 8 * 20, 2                  # 160 tokens per frame -> 320 rows per forward (> 128: fold eligible)
        kv = m.allocate_kv_cache(1, 3 * F_ * fs, torch.float32, torch.device("cpu"))
        ca = [dict(k=torch.zeros(1, 512, 2, 128), v=torch.zeros(1, 512, 2, 128), is_init=False) for _ in range(2)]
        ctx = torch.randn(1, 512, 512)
        outs = []
        for chunk in range(3):
            x = torch.randn(1, 16, F_, 16, 40)
            outs.append(m(x, t=torch.full((1, F_), 500.0), context=ctx, seq_len=32760, kv_cache=kv, crossattn_cache=ca,
                          current_start=chunk * F_ * fs))
        return torch.stack(outs), ops.log

    ref, log_ref = run(TorchOps())
    got, log_fold = run(StatsOps())
    assert log_fold.count("gemm") == log_ref.count("gemm")
    assert log_fold.count("attention_qnorm") == 2 * 3 and log_ref.count("attention_qnorm") == 0
    assert rel_l2(got, ref) <= 2e-5


def test_cross_attention_fold_falls_back_for_caches_filled_elsewhere():
    """The folded cross-attention needs a private copy of the text K with norm_q's weight multiplied in.  A cache entry
    that is already initialised but has no such copy (filled by another model instance) must take the unfolded schedule
    -- same result -- and the copy table stays bounded when callers keep re-allocating their caches."""
    from self_forcing_b200.model import B200CausalWanModel

    class StatsOps(TorchOps):
        supports_row_stats = True

    def make():
        m = B200CausalWanModel(dim=256, ffn_dim=256, num_heads=2, num_layers=2, text_dim=512, ops=StatsOps())
        m.init_weights(5)
        return m

    torch.manual_seed(1)
    fs, F_ = 8 * 20, 1
    ctx = torch.randn(1, 512, 512)
    x = torch.randn(1, 16, F_, 16, 40)
    t = torch.full((1, F_), 300.0)

    def caches(m):
        kv = m.allocate_kv_cache(1, 2 * fs, torch.float32, torch.device("cpu"))
        ca = [dict(k=torch.zeros(1, 512, 2, 128), v=torch.zeros(1, 512, 2, 128), is_init=False) for _ in range(2)]
        return kv, ca

    a = make()
    kv_a, ca = caches(a)
    ref = a(x, t=t, context=ctx, seq_len=32760, kv_cache=kv_a, crossattn_cache=ca, current_start=0)
    assert a.ops.log.count("attention_qnorm") == 2 and all(c["is_init"] for c in ca)
    b = make()                                    # same weights, but it never saw these cache entries being filled
    kv_b, _ = caches(b)
    out = b(x, t=t, context=ctx, seq_len=32760, kv_cache=kv_b, crossattn_cache=ca, current_start=0)
    assert b.ops.log.count("attention_qnorm") == 0            # unfolded schedule for both layers
    assert rel_l2(out, ref) <= 2e-5
    # a caller that re-allocates its cross-attention cache on every call cannot grow the table without bound
    for _ in range(12):
        kv_c, ca_c = caches(a)
        a(x, t=t, context=ctx, seq_len=32760, kv_cache=kv_c, crossattn_cache=ca_c, current_start=0)
    assert len(a._ck_fold) <= 4 * a.num_layers
