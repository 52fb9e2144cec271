"""Ulysses head-parallel rollout on real GPUs (peer-memory all-to-all inside qk_norm_rope_sp / attention_fwd_sp):
launches tools/ulysses_gpu_check.py under torchrun with one rank per GPU.  Needs >= 2 GPUs on the box; skips otherwise
(the host logic is covered on CPU by tests/test_ulysses_gloo.py with gloo, world_size 2)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_ulysses_rollout_matches_reference_golden_on_gpus(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "ulysses_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    # the ranks share one stdout pipe: two records can land on one line, so decode the JSON object after every marker
    dec = json.JSONDecoder()
    lines = [dec.raw_decode(chunk.lstrip())[0] for chunk in r.stdout.split("ULYSSES_CHECK ")[1:]]
    assert r.returncode == 0 and len(lines) == world, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    for res in lines:
        assert res["ok"] and res["rel_l2"] <= 1e-2 and res["identical_across_ranks"] and res["repeatable"], res
        assert tuple(res["index"]) == tuple(res["golden_index"]), res      # cache indices: bit-exact
        assert res["heads_per_rank"] == 12 // world
