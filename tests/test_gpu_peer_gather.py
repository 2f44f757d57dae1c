"""The all-gather of the step record fused into the step kernel over peer memory (include/dmb.h dmb_set_peer_gather,
dist.PeerRecordGather).  With one rank the kernel stores into its own IPC-allocated buffer and signals its own flag,
which exercises the whole device path (epilogue stores, last-CTA signal, polling wait / wait folded into the next step)
on a single GPU; with >= 2 visible GPUs the multi-process check of tools/gpu_peer_gather_check.py runs as well
(bit-identical to ncclAllGather on every step)."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist

import common

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_fused_gather_single_rank_matches_local_record():
    from deepmimic_mujoco_b200.dist import PeerRecordGather
    from deepmimic_mujoco_b200.env import DPVecEnv
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1)
    try:
        n = 300
        env = DPVecEnv(n, motions=("walk",), seed=5, reward_mode=4, auto_reset=True)
        sim = env.sim
        env.reset()
        peer = PeerRecordGather(sim, n, 0, depth=6)
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        recs = []
        for t in range(14):
            peer.arm()
            if t >= 2:
                got = peer.wait(in_next_step=True)         # step t - 2, complete once this step's kernel is
            obs, rew, done, _ = env.step(torch.rand(n, 28, device="cuda", generator=g) - 0.5)
            recs.append(sim.rec.clone())
            assert torch.equal(sim.rec[:, :56], obs) and torch.equal(sim.rec[:, 56], rew)
            if t >= 2:
                torch.cuda.synchronize()
                assert torch.equal(got, recs[t - 2]), t
        for t in (12, 13):                                  # the polling-kernel form drains the rest
            assert torch.equal(peer.wait(), recs[t])
        with pytest.raises(RuntimeError):
            for _ in range(7):
                peer.arm()                                  # more outstanding steps than buffers
        peer.close()
        env.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_fused_gather_two_ranks_bit_identical_to_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(common.ROOT, "tools", "gpu_peer_gather_check.py"), "500"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=common.ROOT)
    assert p.returncode == 0 and "OK, bit-identical" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])
