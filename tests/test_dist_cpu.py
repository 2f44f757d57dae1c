"""Host-side sharding + all-gather logic on CPU with gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common  # noqa: F401
from deepmimic_mujoco_b200.dist import RecordGather, mixed_clip_ids, shard_range


def test_shard_range_partitions():
    for n, w in ((4096, 1), (4096, 8), (65536, 8), (10, 3), (7, 7)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [e - s for s, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(3, 4, 0)
    assert mixed_clip_ids(4, 10, 3).tolist() == [1, 2, 0, 1, 2, 0]


def _worker(rank, world, port, n_global, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(n_global, world, rank)
    local = torch.arange(s, e, dtype=torch.float32)[:, None] * torch.ones(1, 58) + torch.arange(58) * 1e-3
    g = RecordGather(local, n_global)
    g4 = RecordGather(local, n_global, depth=4)
    for k in range(4):
        g4.launch(local * (k + 1))
    deep_ok = all(bool(torch.equal(g4.wait(), (torch.arange(n_global, dtype=torch.float32)[:, None] * torch.ones(1, 58) + torch.arange(58) * 1e-3) * (k + 1))) for k in range(4))
    out = g()
    expect = torch.arange(n_global, dtype=torch.float32)[:, None] * torch.ones(1, 58) + torch.arange(58) * 1e-3
    ok = bool(torch.equal(out, expect))
    # overlapped form (double buffer): two launches may be outstanding, results come back oldest first
    g.launch(local); g.launch(local * 2)
    with pytest.raises(RuntimeError):
        g.launch(local)
    a = g.wait().clone(); b = g.wait()
    ok = ok and bool(torch.equal(a, expect)) and bool(torch.equal(b, expect * 2)) and a.data_ptr() != b.data_ptr()
    q.put((rank, ok and deep_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [16, 17])
def test_all_gather_world2(n_global):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
