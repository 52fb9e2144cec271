"""Ulysses head-parallel execution of ONE long video across the GPUs of an NVSwitch box.

Follows the pattern of the reference's sequence-parallel wrapper for the bidirectional WanModel
(wan/distributed/xdit_context_parallel.py:66-192: tokens chunked on dim 1 per rank `:131-133`, rank-offset RoPE
`:52-58`, `xFuserLongContextAttention` = all-to-all -> attention over all tokens for a head group -> all-to-all
back `:179-184`, final all-gather `:142`), applied here to the KV-cached causal rollout:

  * every rank holds L / P consecutive tokens of the chunk (all heads) for the token-wise work -- LayerNorm +
    modulation, the projection GEMMs, the FFN, cross-attention against the (replicated) text K/V;
  * self-attention runs head-parallel: rank g owns head group g (H / P heads) for ALL tokens, and the rolling KV
    cache is stored head-sharded ([S, H/P, 128] per rank);
  * the two exchanges per block are not NCCL calls: `sfb_qk_norm_rope_sp` stores each head group's rotated q / k / v
    rows straight into the owning rank's q buffer and KV-cache slot, and `sfb_attention_fwd_sp` stores each output
    row straight into the token owner's buffer -- plain st.global on peer-mapped pointers over NVLink, inside the
    kernels that produce the data.  `sfb_peer_barrier` (flag words in peer memory) separates the phases.
  * the [L/P, 64] head outputs are all-gathered (NCCL, once per forward) before the unpatchify / flow->x0 kernel.

torch supplies the plumbing only: symmetric device memory (`torch.distributed._symmetric_memory`) for the
peer-mapped buffers and the process group.  H = 12 for the 1.3B model, so P is 2 or 4 (8 GPUs = 2 videos x 4).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass
class PeerTensor:
    """Same-shape buffer on every rank: `local` is this rank's tensor, `ptrs[p]` the device address of rank p's
    buffer as mapped into this process (empty on CPU, where the test double exchanges with collectives)."""
    local: torch.Tensor
    ptrs: List[int] = field(default_factory=list)
    keep: object = None   # keeps the symmetric-memory handle / opened IPC storages alive

    def ptrs_at(self, elem_offset: int) -> List[int]:
        step = elem_offset * self.local.element_size()
        return [p + step for p in self.ptrs]


class UlyssesGroup:
    def __init__(self, group: Optional[dist.ProcessGroup] = None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("UlyssesGroup needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        if not 1 <= self.world <= 8:
            raise ValueError("Ulysses group size must be 1..8 (one NVSwitch box)")
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.flags = self.alloc((self.world + 1,), torch.int32)   # [p] written by rank p, [world] = own call counter
        self.flags.local.zero_()
        self.sync_host()

    @property
    def on_cuda(self) -> bool:
        return self.device.type == "cuda"

    def sync_host(self) -> None:
        """Host-level barrier (setup / teardown only -- never on the data path)."""
        if self.on_cuda:
            torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def alloc(self, shape: Sequence[int], dtype) -> PeerTensor:
        """Collective: every rank allocates the same shape; returns the local tensor plus every rank's address."""
        if not self.on_cuda:
            return PeerTensor(torch.zeros(*shape, dtype=dtype))
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(*shape, dtype=dtype, device=self.device)
        hdl = symm.rendezvous(t, self.group)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != t.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned an unexpected pointer table")
        return PeerTensor(t, ptrs, hdl)

    def barrier(self, ops) -> None:
        """Device-side barrier on the current stream: all peer stores issued before it on any rank are visible to
        kernels launched after it on every rank."""
        ops.peer_barrier(self)

    def all_gather_rows(self, local: torch.Tensor, out: torch.Tensor) -> None:
        dist.all_gather_into_tensor(out, local, group=self.group)


def shard_rows(total_rows: int, world: int, rank: int):
    """Contiguous token slice of one rank (xdit_context_parallel.py:131-133: torch.chunk on dim 1)."""
    if total_rows % world:
        raise ValueError(f"{total_rows} tokens do not split evenly over {world} ranks")
    n = total_rows // world
    return rank * n, n
