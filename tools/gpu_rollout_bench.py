#!/usr/bin/env python3
"""Device-resident rollout throughput: fused policy inference + env step (+ GAE), 4096 envs."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv
from deepmimic_mujoco_b200.policy import MlpPolicy
from deepmimic_mujoco_b200.rollout import SegmentGenerator, add_vtarg_and_adv
E, T = 4096, 64
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4)
pi = MlpPolicy(seed=0)
gen = SegmentGenerator(pi, env, horizon=T)
seg = next(gen); pi.ob_rms.update(seg["ob"])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    seg = next(gen)
    add_vtarg_and_adv(seg, 0.995, 0.97)
    pi.ob_rms.update(seg["ob"])
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"rollout with policy in the loop: {4*T*E/dt/1e6:.2f} M env-steps/s ({dt/(4*T)*1e3:.3f} ms per step of {E} envs); "
      f"mean ep len {seg['ep_lens'].float().mean().item():.1f}")
# policy kernel alone
ob = seg["ob"][0].contiguous()
for _ in range(5): pi.act(True, ob)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): pi.act(True, ob)
e1.record(); torch.cuda.synchronize()
print(f"policy kernel: {e0.elapsed_time(e1)/50*1e3:.1f} us per call of {E} rows")
