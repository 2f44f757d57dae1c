#!/usr/bin/env python3
"""Where does the heterogeneous batch lose time: inside CTAs (lockstep waits) or between CTAs (tail)?
With DMB_NO_SORT=1 the schedule is index order, so CTA c gets envs [14c, 14c+14).  Three batches with the same
multiset of env states: (a) random order, (b) sorted by measured cost, (c) each CTA 14 copies of one env."""
import os, sys
import numpy as np, torch
os.environ["DMB_NO_SORT"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv
E = 4096
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=False)
sim = env.sim
env2 = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
env2.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
for t in range(60):
    env2.step(torch.rand(E, 28, device="cuda", generator=g) - 0.5)
keys = ("qpos", "qvel", "warm", "idx_curr", "idx_init", "ep_len", "ep_ret", "reset_count")
state = {k: getattr(env2.sim, k).clone() for k in keys}
act = torch.rand(E, 28, device="cuda", generator=g) - 0.5
def load(perm):
    for k in keys: getattr(sim, k).copy_(state[k][perm])
def timed(perm, reps=5):
    ts = []
    for _ in range(reps):
        load(perm); a = act[perm].contiguous()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
ident = torch.arange(E, device="cuda")
load(ident); d = sim.forward_debug(act); cost = torch.tensor(d["nefc"] * d["iter"], device="cuda")
print("mean nefc*iter %.1f  mean nefc %.2f" % (cost.float().mean().item(), d["nefc"].mean()))
print("(a) random order          : %.3f ms" % timed(torch.randperm(E, device="cuda", generator=g)))
srt = torch.argsort(cost, descending=True)
print("(b) sorted by true cost   : %.3f ms" % timed(srt))
print("(b') sorted ascending     : %.3f ms" % timed(torch.flip(srt, [0])))
rep = srt[::14].repeat_interleave(14)[:E]
print("(c) 14 copies per CTA     : %.3f ms  (same per-CTA cost profile as (b), no intra-CTA spread)" % timed(rep))
