"""Drop-in boundary: the reference's *unmodified* pipeline/causal_inference.py drives our
B200DiffusionWrapper (constructor injection).  Build container only; kernels are
replaced by the CPU test double, so this checks the Python surface, not the CUDA code."""
import contextlib
import io

import pytest
import torch

from _torch_ops import TorchOps
from helpers import ROLLOUT_CASES, _IdentityVAE, _TextEncoder, golden, patched_randn_like, pipeline_args, rel_l2, synthetic_inputs
from oracle import causal_wan_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present")


def test_reference_pipeline_runs_on_our_wrapper():
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper
    ref = ref_shim.load_reference()
    case = ROLLOUT_CASES["tiny_test_yaml"]
    g = golden("rollout_tiny.pt")["tiny_test_yaml"]
    cfg = O.OracleConfig(**O.WAN_TINY)
    w = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B, ffn_dim=512, num_layers=2), timestep_shift=case["shift"],
                             ops=TorchOps())
    w.model.load_state_dict(O.make_random_params(cfg, seed=0), strict=True)
    pe, noise = synthetic_inputs(1, case["frames"])
    with contextlib.redirect_stdout(io.StringIO()):
        pipe = ref.CausalInferencePipeline(pipeline_args(case), "cpu", generator=w, text_encoder=_TextEncoder(pe),
                                           vae=_IdentityVAE())
        pipe.num_transformer_blocks = 2       # the reference hard-codes 30 (causal_inference.py:33)
        with torch.no_grad(), patched_randn_like(3):
            _, lat = pipe.inference(noise, ["synthetic"], return_latents=True)
    assert rel_l2(lat, g["latents"]) <= 1e-2
    assert (int(pipe.kv_cache1[0]["global_end_index"]), int(pipe.kv_cache1[0]["local_end_index"])) == tuple(g["final_index"])
    # second inference() re-uses the caches through the reference's rebinding reset (:123-132)
    with contextlib.redirect_stdout(io.StringIO()), torch.no_grad(), patched_randn_like(3):
        _, lat2 = pipe.inference(noise, ["synthetic"], return_latents=True)
    assert torch.equal(lat, lat2)
