"""Multi-GPU plumbing: one process per GPU, contiguous env shards, one all-gather per step.

Envs are independent, so the path shards embarrassingly: rank g owns envs
[g*N/G, (g+1)*N/G) with its own handle and Philox stream ids (first_env_id = shard start, which
mirrors the reference's per-worker seeding ``workerseed = seed + 10000*rank``,
/root/reference/src/trpo.py:341).  The only collective is the all-gather of the packed
[N/G, obs_dim+2] fp32 record (obs, reward, done) -- NCCL over NVLink on GPUs, gloo in the CPU
tests of this host logic.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [start, end) of rank; the first (num_envs % world_size) ranks get one more."""
    if world_size <= 0 or not (0 <= rank < world_size) or num_envs < world_size:
        raise ValueError("need 0 <= rank < world_size <= num_envs")
    base, rem = divmod(num_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def mixed_clip_ids(start: int, end: int, nclip: int) -> torch.Tensor:
    """Clip id per env for mixed batches: global env index modulo number of clips (BASELINE config 5)."""
    return (torch.arange(start, end, dtype=torch.int64) % nclip).to(torch.int32)


class RecordGather:
    """All-gather of the per-step record into a pre-allocated [N_global, width] buffer."""

    def __init__(self, local_rec: torch.Tensor, num_envs_global: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local = local_rec
        width = local_rec.shape[1]
        self.equal = num_envs_global % self.world == 0
        self.out = torch.empty(num_envs_global, width, dtype=local_rec.dtype, device=local_rec.device)
        if not self.equal:  # ragged shards: pad every rank to the largest shard, gather, then compact
            self.sizes = [shard_range(num_envs_global, self.world, r) for r in range(self.world)]
            self.maxn = max(e - s for s, e in self.sizes)
            self.pad = torch.zeros(self.maxn, width, dtype=local_rec.dtype, device=local_rec.device)
            self.buf = torch.empty(self.world * self.maxn, width, dtype=local_rec.dtype, device=local_rec.device)

    def __call__(self) -> torch.Tensor:
        if self.world == 1:
            self.out.copy_(self.local)
        elif self.equal:
            dist.all_gather_into_tensor(self.out, self.local, group=self.group)
        else:
            self.pad[: self.local.shape[0]].copy_(self.local)
            dist.all_gather_into_tensor(self.buf, self.pad, group=self.group)
            for r, (s, e) in enumerate(self.sizes):
                self.out[s:e].copy_(self.buf[r * self.maxn: r * self.maxn + (e - s)])
        return self.out
