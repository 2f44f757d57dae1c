#!/usr/bin/env python3
"""Workload statistics of the bench rollout on the GPU: per-step kernel time and the
distribution of ncon / nefc / PGS iterations across envs as episodes desynchronise."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
sim = env.sim
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
zero = torch.zeros(E, sim.nu, device="cuda")
for t in range(241):
    act = torch.rand(E, sim.nu, device="cuda", generator=g) - 0.5
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); obs, rew, done, info = env.step(act); e1.record(); torch.cuda.synchronize()
    if t % 20 == 0 or t < 6:
        warm = sim.warm.clone()
        d = sim.forward_debug(act)      # one mj_forward at the current state (first RK stage of the next step)
        sim.warm.copy_(warm)
        print(f"step {t:3d}: {e0.elapsed_time(e1):6.3f} ms  done {int(done.sum()):4d}  ncon mean {d['ncon'].mean():5.2f} max {d['ncon'].max():2d}  "
              f"nefc mean {d['nefc'].mean():5.2f} p90 {np.percentile(d['nefc'],90):4.0f} max {d['nefc'].max():2d}  iters mean {d['iter'].mean():5.1f}  "
              f"mean ep_len {sim.ep_len.float().mean().item():5.1f} flags {int((sim.flags!=0).sum())}")
