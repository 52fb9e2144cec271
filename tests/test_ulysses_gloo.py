"""Ulysses head-parallel host logic with world_size 2 on CPU (gloo): token sharding, per-rank modulation / gate /
RoPE offsets, head-sharded KV caches and index traces, the final gather.  The kernels are replaced by the CPU test
double (tests/_torch_ops.py), whose *_sp ops exchange with gloo collectives instead of peer stores; the result must
equal the single-process rollout golden made by the unmodified reference."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(4)
    try:
        from _torch_ops import TorchOps
        from helpers import golden, make_product_pipeline, patched_randn_like, rel_l2
        from self_forcing_b200.ulysses import UlyssesGroup, shard_rows
        assert shard_rows(4680, world, rank) == (rank * 4680 // world, 4680 // world)
        g = golden("rollout_tiny.pt")["chunkwise"]
        pipe, *_, noise = make_product_pipeline(g["case"], "cpu", ops=TorchOps())
        sp = UlyssesGroup(device="cpu")
        pipe.generator.model.enable_ulysses(sp)
        with patched_randn_like(3):
            _, lat = pipe.inference(noise, ["synthetic"], return_latents=True)
        kc = pipe.kv_cache1[0]["k"]
        res = dict(err=rel_l2(lat, g["latents"]), heads=kc.shape[2],
                   index=(int(pipe.kv_cache1[0]["global_end_index"]), int(pipe.kv_cache1[-1]["local_end_index"])),
                   final_index=tuple(g["final_index"]), tail_zero=float(kc[:, 9360:].float().abs().max()),
                   head_nonzero=float(kc[:, :9360].float().abs().max()))
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
        torch.save(lat, os.path.join(out_dir, f"lat{rank}.pt"))

        # bidirectional teacher forward (BASELINE config 5) head-parallel: sample 0 of the reference golden
        from oracle import causal_wan_oracle as O
        from oracle.make_golden import BIDIR, bidirectional_cfg, bidirectional_inputs
        from self_forcing_b200.model import B200WanModel
        gb = golden("bidirectional_tiny.pt")
        r = BIDIR
        teacher = B200WanModel(dim=r["dim"], ffn_dim=r["ffn_dim"], num_heads=r["num_heads"], num_layers=r["num_layers"],
                               text_dim=r["text_dim"], ops=TorchOps()).to(torch.bfloat16)
        teacher.load_state_dict(O.make_random_params(bidirectional_cfg(), seed=9), strict=True)
        teacher.enable_ulysses(sp)
        xb, tb, cb = bidirectional_inputs()
        out = teacher(xb[:1], t=tb[:1], context=cb[:1], seq_len=gb["seq_len"])
        torch.save(dict(err=rel_l2(out, gb["flow"][:1])), os.path.join(out_dir, f"bidir{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_ulysses_rollout_world2_matches_reference_golden(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    lats = [torch.load(tmp_path / f"lat{r}.pt") for r in range(world)]
    assert torch.equal(lats[0], lats[1])                       # every rank ends with the full, identical video
    for r in res:
        assert r["heads"] == 12 // world                       # caches are head-sharded
        assert r["index"] == r["final_index"]                  # index arithmetic unchanged (bit-exact)
        assert r["tail_zero"] == 0.0 and r["head_nonzero"] > 0
        assert r["err"] <= 1e-2, r["err"]
    for rk in range(world):
        assert torch.load(tmp_path / f"bidir{rk}.pt")["err"] <= 1e-2
