#!/usr/bin/env python3
"""What does the L2 flush between timed steps cost the step kernel, and which cold data is responsible?
Times dmb_step (CUDA events around the call) in four situations:
  warm        no flush between steps
  cold        192 MiB memset before every step (bench.py's timed loop)
  cold+code   flush, then one step of a 28-env helper sim (re-fetches the kernel's code, the model tables and
              the mocap tables -- but not the 4096 envs' state) before the timed step
  cold+state  flush, then a read of the state tensors (qpos/qvel/warm/actions) before the timed step
Run on the GPU box: python tools/gpu_cold_start_probe.py [envs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
helper = DPVecEnv(28, motions=("walk",), seed=1, reward_mode=4, auto_reset=True)
env.reset(); helper.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.rand(16, E, 28, device="cuda", generator=g) - 0.5
hact = torch.rand(28, 28, device="cuda", generator=g) - 0.5
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
for t in range(60):
    env.step(acts[t % 16])
torch.cuda.synchronize()


def run(mode, K=100):
    tot = 0.0
    for t in range(K):
        if mode != "warm":
            flush.zero_()
        if mode == "cold+code":
            helper.step(hact)
        if mode == "cold+state":
            s = env.sim
            _ = s.qpos.sum() + s.qvel.sum() + s.warm.sum() + acts[t % 16].sum()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        env.step(acts[t % 16])
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / K * 1e3


for mode in ("warm", "cold", "cold+code", "cold+state", "warm", "cold"):
    print(f"{mode:11s} {run(mode):7.1f} us/step  ({E} envs, launch {env.sim.launch_info()})")
