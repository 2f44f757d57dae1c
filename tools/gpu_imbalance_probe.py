#!/usr/bin/env python3
"""How much does work imbalance between the envs of a CTA cost?  Compares the normal heterogeneous
rollout with a homogeneous one (every env is a copy of one env, same action) at matched work."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv
E = 4096
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
sim = env.sim
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
def timed(act):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); env.step(act); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for t in range(40):
    env.step(torch.rand(E, 28, device="cuda", generator=g) - 0.5)
het = []
for t in range(40):
    act = torch.rand(E, 28, device="cuda", generator=g) - 0.5
    warm = sim.warm.clone(); d = sim.forward_debug(act); sim.warm.copy_(warm)
    het.append((timed(act), float((d["nefc"] * d["iter"]).mean()), float(d["nefc"].mean())))
print("heterogeneous: ms/step %.3f  mean nefc*iter %.1f  mean nefc %.2f" % tuple(np.mean(het, axis=0)))
# homogeneous: broadcast one env (try a few source envs with different work)
state = {k: getattr(sim, k).clone() for k in ("qpos", "qvel", "warm", "idx_curr", "idx_init", "ep_len", "ep_ret", "reset_count")}
for src in range(6):
    for k, v in state.items():
        getattr(sim, k).copy_(v[src:src + 1].expand_as(v))
    res = []
    for t in range(6):
        act = (torch.rand(1, 28, device="cuda", generator=g) - 0.5).expand(E, 28).contiguous()
        warm = sim.warm.clone(); d = sim.forward_debug(act); sim.warm.copy_(warm)
        res.append((timed(act), float((d["nefc"] * d["iter"]).mean()), float(d["nefc"].mean())))
    r = np.array(res)
    print("homogeneous src %d: " % src + "  ".join("%.3f ms (nefc*it %.0f, nefc %.0f)" % tuple(x) for x in r))
