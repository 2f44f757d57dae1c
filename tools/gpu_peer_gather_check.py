#!/usr/bin/env python3
"""Correctness of the fused all-gather (run under torchrun on >= 2 GPUs of one node): for a rollout with per-rank
different random actions, the record gathered through NVLink peer stores (dist.PeerRecordGather) must equal, bit for
bit, an ncclAllGather of every rank's own record, for every step, with a consumer lag of 0 and of 2 steps."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.dist import PeerRecordGather, mixed_clip_ids  # noqa: E402
from deepmimic_mujoco_b200.env import DPVecEnv  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
first = rank * E
env = DPVecEnv(E, motions=("walk", "spinkick"), device=dev, seed=3, first_env_id=first, reward_mode=4, auto_reset=True,
               clip_ids=mixed_clip_ids(first, first + E, 2))
sim = env.sim
env.reset()
peer = PeerRecordGather(sim, E * world, first, depth=6)
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
ref = torch.empty(E * world, sim.obs_dim + 2, device=dev)
bad = 0
for lag, steps in ((0, 12), (2, 40), (-2, 40)):      # -2: lag 2 with the wait folded into the step kernel
    refs = []
    fold = lag < 0
    lag = abs(lag)
    base = peer._t
    for t in range(steps):
        peer.arm()
        if fold and t >= lag:
            got_fold = peer.wait(in_next_step=True)      # step t - lag, complete once this step's kernel is
        env.step(torch.rand(E, sim.nu, device=dev, generator=g) - 0.5)
        if fold and t >= lag:
            torch.cuda.synchronize()
            bad += int(not torch.equal(got_fold, refs[t - lag]))
        r = torch.empty_like(ref)
        dist.all_gather_into_tensor(r, sim.rec.contiguous())
        refs.append(r)
        if t >= lag and not fold:
            got = peer.wait()
            torch.cuda.synchronize()
            if not torch.equal(got, refs[t - lag]):
                bad += 1
                if rank == 0:
                    print(f"lag {lag} step {t}: mismatch, max |diff| {(got - refs[t - lag]).abs().max().item():.3e}")
    while peer._waited < peer._t:
        t = peer._waited
        got = peer.wait()
        torch.cuda.synchronize()
        bad += int(not torch.equal(got, refs[t - base]))
    dist.barrier()
tot = torch.tensor([bad], device=dev)
dist.all_reduce(tot)
if rank == 0:
    print(f"peer gather check: world {world}, {E} envs per rank: {'OK, bit-identical to ncclAllGather on every step' if int(tot) == 0 else f'{int(tot)} MISMATCHES'}")
peer.close()
env.close()
dist.destroy_process_group()
sys.exit(0 if int(tot) == 0 else 1)
