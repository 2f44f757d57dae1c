#!/usr/bin/env python3
"""Small rollout for profiling one configuration: python tools/gpu_small_rollout.py <envs> <steps>"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv
E = int(sys.argv[1]); T = int(sys.argv[2])
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.rand(16, E, env.sim.nu, device="cuda", generator=g) - 0.5
for t in range(T):
    env.step(acts[t % 16])
torch.cuda.synchronize()
