#!/usr/bin/env python3
"""Average cycles per forward-evaluation phase and warp (needs a -DDMB_PHASE_TIMERS=1 build selected with DMB_LIB).
Run on the GPU box:  DMB_LIB=$PWD/variants/libdmb200_timers.so python tools/gpu_phase_timers.py [envs]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
sim = env.sim
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.rand(16, E, sim.nu, device="cuda", generator=g) - 0.5
buf = np.zeros(32, dtype=np.uint64)
for t in range(100):
    env.step(acts[t % 16])
torch.cuda.synchronize()
sim.L.dmb_phase_cycles(buf.ctypes.data_as(C.c_void_p))
T = 50
for t in range(T):
    env.step(acts[t % 16])
torch.cuda.synchronize()
sim.L.dmb_phase_cycles(buf.ctypes.data_as(C.c_void_p))
names = {1: "wait @stage barrier", 16: "kinematics", 2: "comPos", 11: "crb: composite inertias", 12: "crb: M entries",
         13: "factor: L'DL elimination", 3: "factor: scaling", 17: "smooth forces (RNE)", 4: "LT solve of qfrc_smooth",
         5: "geom poses + collision", 6: "count rows + J rows + row scalars", 7: "half solve (unshared path)",
         18: "wait @barrier A (rows assembled)", 19: "half-solve tasks (shared)", 20: "wait @barrier B", 21: "Gram tasks (shared)",
         8: "wait @barrier C", 9: "wait @solve barrier",
         14: "solve: warmstart + residual", 15: "solve: PGS sweeps", 10: "solve: Y'f + L solve",
         25: "step prologue / RK4 update between stages", 22: "RK4 final update", 23: "reward (FK + features)", 24: "done / reset / obs / state store"}
per = buf.astype(np.float64) / (T * E * 4)
order = [25, 1, 16, 2, 11, 12, 13, 3, 17, 4, 5, 6, 7, 18, 19, 20, 21, 8, 9, 14, 15, 10, 22, 23, 24]
tot = sum(per[i] for i in order)
print(f"cycles per stage and warp (step-level buckets 22-25 divided by 4): {tot:9.0f}  ({tot*4/1.965e3:7.1f} us per env-step at 1.965 GHz)")
for i in order:
    print(f"  {names[i]:30s} {per[i]:9.0f}  {100*per[i]/tot:5.1f}%")
