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
    """All-gather of the per-step record into pre-allocated [N_global, width] buffers.

    ``g(local)`` is the blocking form (gather on the current stream, used by the CPU/gloo tests and simple callers).
    ``launch(local)`` / ``wait()`` is the overlapped form used by the rollout loops: the collective runs on a side
    stream, gated by an event recorded after the step kernel, and writes the gathered record of step t into
    ``out_buffers[t % depth]`` while the compute stream already runs the following steps (the env step never waits
    for the slowest rank's previous kernel; SURVEY.md section 5 "Distributed communication backend").  ``wait()`` makes
    the current stream wait for the oldest outstanding gather and returns its buffer.  ``depth`` outstanding gathers
    let the ranks drift ``depth - 1`` steps apart: the step time of the heaviest env varies from step to step, and a
    deeper pipeline averages that variation out instead of paying the slowest rank every step.  The caller keeps
    the gathered tensors' source buffers alive for as long (``BatchedSim(rec_depth=...)``).
    """

    def __init__(self, local_rec: torch.Tensor, num_envs_global: int, group=None, depth: int = 2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local = local_rec
        width = local_rec.shape[1]
        self.equal = num_envs_global % self.world == 0
        kw = dict(dtype=local_rec.dtype, device=local_rec.device)
        self.depth = max(2, int(depth))
        self.out_buffers = [torch.empty(num_envs_global, width, **kw) for _ in range(self.depth)]
        self.out = self.out_buffers[0]
        if not self.equal:  # ragged shards: pad every rank to the largest shard, gather, then compact
            self.sizes = [shard_range(num_envs_global, self.world, r) for r in range(self.world)]
            self.maxn = max(e - s for s, e in self.sizes)
            self.pad = torch.zeros(self.maxn, width, **kw)
            self.buf = torch.empty(self.world * self.maxn, width, **kw)
        self.cuda = local_rec.is_cuda
        self.side = torch.cuda.Stream(device=local_rec.device) if self.cuda else None
        self._pending = []      # (event, buffer) of launched gathers, oldest first
        self._t = 0

    def _gather(self, out: torch.Tensor, local: torch.Tensor) -> None:
        if self.world == 1:
            out.copy_(local)
        elif self.equal:
            dist.all_gather_into_tensor(out, local, group=self.group)
        else:
            self.pad[: local.shape[0]].copy_(local)
            dist.all_gather_into_tensor(self.buf, self.pad, group=self.group)
            for r, (s, e) in enumerate(self.sizes):
                out[s:e].copy_(self.buf[r * self.maxn: r * self.maxn + (e - s)])

    def __call__(self, local: torch.Tensor = None) -> torch.Tensor:
        self._gather(self.out, self.local if local is None else local)
        return self.out

    def launch(self, local: torch.Tensor = None) -> None:
        """Start gathering ``local`` (default: the tensor given at construction) without blocking the current
        stream.  At most ``depth`` gathers may be outstanding: call ``wait()`` before the next one."""
        local = self.local if local is None else local
        if len(self._pending) >= self.depth:
            raise RuntimeError(f"RecordGather: {self.depth} gathers already outstanding; call wait() first")
        out = self.out_buffers[self._t % self.depth]
        self._t += 1
        if not self.cuda:
            self._gather(out, local)
            self._pending.append((None, out))
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(local.device))      # after the step kernel that wrote `local`
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            self._gather(out, local)
            done = torch.cuda.Event()
            done.record(self.side)
        self._pending.append((done, out))

    def wait(self) -> torch.Tensor:
        """Make the current stream wait for the oldest outstanding gather; returns its [N_global, width] buffer."""
        done, out = self._pending.pop(0)
        if done is not None:
            torch.cuda.current_stream(out.device).wait_event(done)
        self.out = out
        return out

    def drain(self) -> None:
        while self._pending:
            self.wait()
