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


class _DevArray:
    """A raw device allocation exposed through ``__cuda_array_interface__`` so that torch can view it."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class PeerRecordGather:
    """The all-gather of the step record FUSED into the step kernel over NVLink peer memory (include/dmb.h,
    ``dmb_set_peer_gather``): every rank owns ``depth`` gathered ``[N_global, width]`` buffers and arrival flags in
    IPC-shareable device memory; the step kernel of every rank stores each env's record row straight into all ranks'
    buffers (posted peer stores from its epilogue) and its last CTA bumps every rank's flag.  No collective kernel
    exists, so nothing competes with the persistent step kernel for SMs and no rank waits for another one until it
    *consumes* a gathered record (``wait()``: one polling thread on the caller's stream).

    Protocol (all ranks run the same step sequence): ``arm()`` before ``sim.step`` of step t selects slot t % depth;
    ``wait()`` returns the gathered buffer of the oldest step not yet waited for -- as a polling kernel on the current
    stream, or (``in_next_step=True``) folded into the next step kernel.  A slot is overwritten ``depth`` steps later
    by every rank; a reader of step t - L that runs between the kernels of step t and t + 1 has finished before its
    rank signals step t + 1, and a producer stores step i only after it has seen every rank's signal of step i - L,
    so no slot is overwritten while it is read if ``depth >= 2 L + 1`` (L = 2: depth 5; default 6).  One node,
    <= 8 ranks (one NVLink domain).
    """

    def __init__(self, sim, num_envs_global: int, first_env: int, group=None, depth: int = 6):
        import ctypes as C
        from . import lib as _lib
        if not dist.is_initialized():
            raise RuntimeError("PeerRecordGather needs an initialised process group (handle exchange)")
        self.sim, self.L, self.group = sim, sim.L, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("the fused peer gather serves one NVLink domain (<= 8 ranks)")
        self.depth, self.width, self.N = max(2, int(depth)), sim.obs_dim + 2, int(num_envs_global)
        self.first_env = int(first_env)
        dev = sim.device.index
        nbytes = self.N * self.width * 4
        self._own, handles = [], []
        for d in range(self.depth):                      # gathered buffer + one flag word per slot
            ptrs = []
            for size in (nbytes, 256):
                p, h = C.c_void_p(), C.create_string_buffer(64)
                _lib.check(self.L.dmb_peer_alloc(dev, size, C.byref(p), h), None, "dmb_peer_alloc")
                ptrs.append(p.value); handles.append(h.raw)
            self._own.append(tuple(ptrs))
        allh = [None] * self.world
        dist.all_gather_object(allh, (dev, handles), group=group)
        self._opened = []
        self.peer_rec = [[0] * self.world for _ in range(self.depth)]
        self.peer_flag = [[0] * self.world for _ in range(self.depth)]
        for r, (_, hs) in enumerate(allh):
            for d in range(self.depth):
                if r == self.rank:
                    self.peer_rec[d][r], self.peer_flag[d][r] = self._own[d]
                    continue
                for k, table in ((0, self.peer_rec), (1, self.peer_flag)):
                    p = C.c_void_p()
                    _lib.check(self.L.dmb_peer_open(dev, hs[2 * d + k], C.byref(p)), None, "dmb_peer_open")
                    table[d][r] = p.value
                    self._opened.append(p.value)
        with torch.cuda.device(sim.device):
            self.out_buffers = [torch.as_tensor(_DevArray(self._own[d][0], (self.N, self.width), "<f4"), device=sim.device)
                                for d in range(self.depth)]
            self.flags = [torch.as_tensor(_DevArray(self._own[d][1], (1,), "<i4"), device=sim.device) for d in range(self.depth)]
        self._rec_arr = [(C.c_void_p * self.world)(*self.peer_rec[d]) for d in range(self.depth)]
        self._flag_arr = [(C.c_void_p * self.world)(*self.peer_flag[d]) for d in range(self.depth)]
        self._t = 0            # steps armed so far
        self._waited = 0       # steps waited for so far
        self.out = self.out_buffers[0]
        dist.barrier(group=group)

    def arm(self) -> None:
        """Point the next ``sim.step`` at slot t % depth of every rank."""
        import ctypes as C
        from . import lib as _lib
        if self._t - self._waited >= self.depth:
            raise RuntimeError(f"PeerRecordGather: {self.depth} steps outstanding; call wait() first")
        d = self._t % self.depth
        _lib.check(self.L.dmb_set_peer_gather(self.sim.handle, self.world, self._rec_arr[d], self._flag_arr[d], self.first_env),
                   self.sim.handle, "dmb_set_peer_gather")
        self._t += 1

    def wait(self, in_next_step: bool = False) -> torch.Tensor:
        """Make the current stream wait until every rank's rows of the oldest outstanding step have arrived.
        ``in_next_step``: fold the wait into the next ``sim.step`` (one thread of the step kernel polls at its start;
        no extra launch) -- the returned buffer is complete once that step has completed."""
        import ctypes as C
        from . import lib as _lib
        t = self._waited
        d = t % self.depth
        target = self.world * (t // self.depth + 1)          # flags only ever count up
        if in_next_step:
            _lib.check(self.L.dmb_set_peer_wait(self.sim.handle, C.c_void_p(self._own[d][1]), target), self.sim.handle,
                       "dmb_set_peer_wait")
            self._waited += 1
            self.out = self.out_buffers[d]
            return self.out
        with torch.cuda.device(self.sim.device):
            _lib.check(self.L.dmb_peer_wait(self.sim.handle, C.c_void_p(self._own[d][1]), target,
                                            C.c_void_p(torch.cuda.current_stream(self.sim.device).cuda_stream)),
                       self.sim.handle, "dmb_peer_wait")
        self._waited += 1
        self.out = self.out_buffers[d]
        return self.out

    def drain(self) -> None:
        while self._waited < self._t:
            self.wait()

    def close(self) -> None:
        if getattr(self, "_own", None) is None:
            return
        import ctypes as C
        self.L.dmb_set_peer_gather(self.sim.handle, 0, None, None, 0)
        torch.cuda.synchronize(self.sim.device)
        dist.barrier(group=self.group)                   # nobody stores into a buffer that is about to go away
        for p in self._opened:
            self.L.dmb_peer_close(C.c_void_p(p))
        for rec, flag in self._own:
            self.L.dmb_peer_free(C.c_void_p(rec)); self.L.dmb_peer_free(C.c_void_p(flag))
        self._own = None

