"""The 50-step sampler: CFG + flow-matching UniPC on the cached causal forward.
CPU side: the oracle against the vectors produced by the unmodified reference, then the product's host logic
(scheduler coefficients, history ring, batched pos/neg caches) against the same vectors through the torch test double."""
import pytest
import torch

from _torch_ops import TorchOps
from helpers import golden, make_product_diffusion_pipeline, negative_embeds, rel_l2, synthetic_inputs
from oracle import causal_wan_oracle as O
from oracle import unipc_oracle as U
from oracle.make_golden import UNIPC_TRACE, unipc_trace_flow
from self_forcing_b200.unipc import FlowUniPCMultistepScheduler

TOL = 1e-2


def _trace_start(dtype):
    return torch.randn(UNIPC_TRACE["shape"], generator=torch.Generator().manual_seed(99)).to(dtype)


@pytest.mark.parametrize("name,dtype", [("unipc_trace_bf16", torch.bfloat16), ("unipc_trace_fp32", torch.float32)])
def test_oracle_unipc_matches_reference_trace(name, dtype):
    ref = golden("diffusion_tiny.pt")[name]
    s = U.OracleUniPC()
    s.set_timesteps(UNIPC_TRACE["steps"], UNIPC_TRACE["shift"])
    assert torch.equal(s.timesteps, ref["timesteps"])            # int64 table: bit-exact
    assert torch.equal(s.sigmas, ref["sigmas"])
    x = _trace_start(dtype)
    for i, t in enumerate(s.timesteps):
        x = s.step(unipc_trace_flow(x, i), t, x)
        assert torch.equal(x, ref["samples"][i]), i              # same host, same op order: identical


def test_unipc_table_known_answers():
    """50 steps, shift 5 (probed from the reference): truncated int64 timesteps, final sigma 0."""
    s = U.OracleUniPC()
    s.set_timesteps(50, 5.0)
    assert s.timesteps[:6].tolist() == [999, 995, 991, 987, 982, 978]
    assert s.timesteps[-3:].tolist() == [241, 172, 92]
    assert float(s.sigmas[-1]) == 0.0 and len(s.sigmas) == 51
    assert float(s.sigmas[0]) == pytest.approx(0.9998, abs=1e-4)


def test_oracle_diffusion_rollout_matches_reference_golden():
    g = golden("diffusion_tiny.pt")["cfg_unipc"]
    case = g["case"]
    cfg = O.OracleConfig(**O.WAN_TINY)
    ow = O.OracleWrapper(O.make_random_params(cfg, seed=0), cfg, case["shift"])
    pe, noise = synthetic_inputs(1, case["frames"])
    with torch.no_grad():
        tr = U.diffusion_rollout(ow, noise, pe, negative_embeds(), case["guidance_scale"], case["num_frame_per_block"],
                                 case["sampling_steps"], case["shift"])
    assert tr.index_trace[-1] == tuple(g["final_index"])
    assert len(tr.index_trace) == case["frames"] * (case["sampling_steps"] + 1)
    assert rel_l2(tr.latents, g["latents"]) <= TOL


def test_host_scheduler_reproduces_reference_trace():
    """Host coefficients + history ring + the op-by-op tensor chain = the reference's 50 steps, bit for bit
    (scalar_rounding "bf16": the CPU reference casts its 0-dim scalar tensors to the tensor dtype)."""
    ref = golden("diffusion_tiny.pt")["unipc_trace_bf16"]
    s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, ops=TorchOps(),
                                    scalar_rounding="bf16")
    s.set_timesteps(UNIPC_TRACE["steps"], device="cpu", shift=UNIPC_TRACE["shift"])
    assert torch.equal(s.timesteps, ref["timesteps"]) and torch.equal(s.sigmas, ref["sigmas"])
    x = _trace_start(torch.bfloat16)
    for i, t in enumerate(s.timesteps):
        x = s.step(unipc_trace_flow(x, i), t, x, return_dict=False)[0]
        assert torch.equal(x, ref["samples"][i]), i
    assert s.step_index == 50 and s.this_order == 1              # lower-order final step


def test_host_scheduler_fp32_scalars_stay_close():
    """The CUDA semantics (scalars kept in fp32) differ from the CPU golden only by the bf16 rounding of ~6 scalars
    per step."""
    ref = golden("diffusion_tiny.pt")["unipc_trace_bf16"]
    s = FlowUniPCMultistepScheduler(shift=1, ops=TorchOps())
    s.set_timesteps(UNIPC_TRACE["steps"], device="cpu", shift=UNIPC_TRACE["shift"])
    x = _trace_start(torch.bfloat16)
    for i, t in enumerate(s.timesteps):
        x = s.step(unipc_trace_flow(x, i), t, x, return_dict=False)[0]
    assert 0 < rel_l2(x, ref["samples"][-1]) <= 3e-2


def test_scheduler_plain_step_and_order1():
    """solver_order 1 and a guidance-free step go through the same kernel entry."""
    s = FlowUniPCMultistepScheduler(shift=1, solver_order=1, ops=TorchOps(), scalar_rounding="bf16")
    s.set_timesteps(4, device="cpu", shift=3.0)
    o = U.OracleUniPC(solver_order=1)
    o.set_timesteps(4, 3.0)
    x = xo = _trace_start(torch.bfloat16)
    for i, t in enumerate(s.timesteps):
        out = s.step(unipc_trace_flow(x, i), t, x)
        x = out.prev_sample
        xo = o.step(unipc_trace_flow(xo, i), t, xo)
        assert torch.equal(x, xo), i


def test_scheduler_rejects_unsupported_configurations():
    with pytest.raises(NotImplementedError):
        FlowUniPCMultistepScheduler(prediction_type="epsilon")
    with pytest.raises(NotImplementedError):
        FlowUniPCMultistepScheduler(solver_order=3)
    with pytest.raises(NotImplementedError):
        FlowUniPCMultistepScheduler(use_dynamic_shifting=True)
    s = FlowUniPCMultistepScheduler(ops=TorchOps())
    with pytest.raises(ValueError):
        s.step(torch.zeros(8, dtype=torch.bfloat16), 999, torch.zeros(8, dtype=torch.bfloat16))   # set_timesteps missing


def test_product_diffusion_pipeline_matches_reference_golden():
    g = golden("diffusion_tiny.pt")["cfg_unipc"]
    case = g["case"]
    pipe, cfg, *_ , noise = make_product_diffusion_pipeline(case, "cpu", ops=TorchOps())
    _, lat = pipe.inference(noise, ["synthetic"], None, None, None, return_latents=True)
    idx = tuple(int(c[0][k]) for c in (pipe.kv_cache_pos, pipe.kv_cache_neg) for k in ("global_end_index", "local_end_index"))
    assert idx == tuple(g["final_index"])                          # integers: bit-exact
    assert rel_l2(lat, g["latents"]) <= TOL
    # the positive and negative prompts fill different halves of the batched cache
    assert pipe.kv_cache[0]["k"].shape[0] == 2 and pipe.kv_cache_pos[0]["k"].shape[0] == 1
    assert not torch.equal(pipe.kv_cache_pos[1]["k"][:, :3120], pipe.kv_cache_neg[1]["k"][:, :3120])
    assert float(pipe.kv_cache_pos[0]["k"][:, 3120:].abs().max()) == 0.0
    with pytest.raises(NotImplementedError):
        pipe.inference(noise, ["synthetic"], object(), None, None)


@pytest.mark.parametrize("indep,frames,init_frames,nfpb", [(False, 2, 2, 2), (True, 2, 1, 1), (True, 3, 0, 1)])
def test_product_diffusion_pipeline_wiring_fp32(indep, frames, init_frames, nfpb):
    """fp32 on both sides removes rounding noise, so the conditioning-frame paths (video continuation, image-to-video
    with a lone first frame, and a first chunk of one frame) must reproduce the oracle driver: positive / negative cache
    halves, frame offsets, the solver restarted per chunk, the clean-context refresh."""
    case = dict(frames=frames, num_frame_per_block=nfpb, independent_first_frame=indep, shift=5.0, sampling_steps=4,
                guidance_scale=3.0)
    pipe, cfg, params, pe, neg, noise = make_product_diffusion_pipeline(case, "cpu", ops=TorchOps(), dtype=torch.float32,
                                                                        scalar_rounding="fp32", hw=(16, 24))
    init = None
    if init_frames:
        init = torch.randn(1, init_frames, 16, 16, 24, generator=torch.Generator().manual_seed(4))
    _, lat = pipe.inference(noise, ["synthetic"], None, None, None, initial_latent=init, return_latents=True)
    ow = O.OracleWrapper(params, cfg, case["shift"])
    with torch.no_grad():
        tr = U.diffusion_rollout(ow, noise, pe, neg, case["guidance_scale"], nfpb, case["sampling_steps"], case["shift"],
                                 independent_first_frame=indep, initial_latent=init)
    assert lat.shape == tr.latents.shape == (1, frames + init_frames, 16, 16, 24)
    assert rel_l2(lat, tr.latents) < 1e-4
    idx = tuple(int(c[0][k]) for c in (pipe.kv_cache_pos, pipe.kv_cache_neg) for k in ("global_end_index", "local_end_index"))
    assert idx == tr.index_trace[-1]
    # a second call resets the caches by rebinding their index tensors and reproduces the result
    _, lat2 = pipe.inference(noise, ["synthetic"], None, None, None, initial_latent=init, return_latents=True)
    assert torch.equal(lat, lat2)


@pytest.mark.parametrize("steps,shift,order", [(50, 5.0, 2), (7, 3.0, 2), (5, 8.0, 1)])
def test_unipc_is_exact_for_a_perfect_model(steps, shift, order):
    """Size-independent property: if the model always predicts the true clean sample (flow = (x - x0) / sigma), every
    first-difference term of UniP / UniC vanishes and the solver must walk the exact path x_i = (1 - sigma_i) x0 +
    sigma_i eps, ending at x0 -- for any step count, shift and solver order.  fp32 through the host scheduler."""
    g = torch.Generator().manual_seed(5)
    x0, eps = torch.randn(1, 2, 16, 6, 10, generator=g), torch.randn(1, 2, 16, 6, 10, generator=g)
    s = FlowUniPCMultistepScheduler(shift=1, solver_order=order, ops=TorchOps())
    s.set_timesteps(steps, device="cpu", shift=shift)
    sig = s.sigmas
    x = (1 - sig[0]) * x0 + sig[0] * eps
    for i, t in enumerate(s.timesteps):
        flow = (x - x0) / sig[i]
        x = s.step(flow, t, x, return_dict=False)[0]
        expect = (1 - sig[i + 1]) * x0 + sig[i + 1] * eps
        assert rel_l2(x, expect) < 2e-5, (i, rel_l2(x, expect))
    assert rel_l2(x, x0) < 2e-5


def test_guidance_scale_one_equals_conditional_flow():
    """flow = uncond + 1.0 * (cond - uncond) is the conditional flow up to one rounding: the fused guided step with
    g = 1 and the plain step on the conditional flow agree."""
    g = torch.Generator().manual_seed(6)
    x, fc, fu = (torch.randn(1, 1, 16, 6, 10, generator=g) for _ in range(3))
    outs = []
    for guided in (True, False):
        s = FlowUniPCMultistepScheduler(shift=1, ops=TorchOps())
        s.set_timesteps(4, device="cpu", shift=5.0)
        kw = dict(model_output_uncond=fu, guidance_scale=1.0) if guided else {}
        outs.append(s.step(fc, s.timesteps[0], x, return_dict=False, **kw)[0])
    assert rel_l2(outs[0], outs[1]) < 1e-6
