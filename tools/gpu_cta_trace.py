#!/usr/bin/env python3
"""Per-CTA timeline of the step kernel (DMB_TRACE=1): when each CTA finishes its scheduler rounds, how long the
SMs sit idle at the tail.  Run on the GPU box:  DMB_TRACE=1 python tools/gpu_cta_trace.py [envs]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ.setdefault("DMB_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
sim = env.sim
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
buf = np.zeros((256, 8), dtype=np.int64)
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda") if os.environ.get("TRACE_FLUSH") == "1" else None
for t in range(140):
    act = torch.rand(E, sim.nu, device="cuda", generator=g) - 0.5
    if flush is not None:
        flush.zero_()      # cold L2, as between the timed steps of bench.py
    env.step(act)
    if t >= 100 and t % 10 == 0:
        torch.cuda.synchronize()
        n = sim.L.dmb_get_trace(sim.handle, buf.ctypes.data_as(C.c_void_p), 256)
        tr = buf[:n].astype(np.float64)
        t0 = tr[:, 0].min()
        rounds = (tr[:, 1:6] > 0).sum(axis=1)
        end = tr[:, 6] - t0                      # the CTA's last warp
        diag = buf[:n, 7]
        sweeps, rows, nscr = diag & 0xffff, (diag >> 16) & 0xff, diag >> 24
        r1 = tr[:, 1] - t0
        start = tr[:, 0] - t0
        print(f"step {t}: kernel {end.max()/1e3:7.1f} us | CTA end mean {end.mean()/1e3:7.1f} min {end.min()/1e3:7.1f} max {end.max()/1e3:7.1f} "
              f"| idle tail {100*(1-end.mean()/end.max()):4.1f}% | round-1 end mean {r1.mean()/1e3:6.1f} min {r1.min()/1e3:6.1f} max {r1.max()/1e3:6.1f} "
              f"| rounds/CTA {np.bincount(rounds)} | start spread {start.max()/1e3:5.1f} us")
        slow = np.argsort(end)[-5:][::-1]
        print("   slowest CTAs (us, most sweeps of one env, most rows, scratch envs):",
              [(round(end[i] / 1e3, 1), int(sweeps[i]), int(rows[i]), int(nscr[i])) for i in slow],
              f"| corr(end, sweeps) {np.corrcoef(end, sweeps)[0, 1]:.2f} corr(end, rows) {np.corrcoef(end, rows)[0, 1]:.2f}",
              f"| CTAs with a scratch env: {int((nscr > 0).sum())}, their mean end {end[nscr > 0].mean() / 1e3 if (nscr > 0).any() else 0:.1f} us")
        if t == 130:
            order = np.argsort(r1)
            r2 = end - r1
            print("round-1 duration deciles (us):", np.round(np.percentile(r1, [0, 10, 25, 50, 75, 90, 100]) / 1e3, 1))
            if (rounds >= 2).any():
                print("round-2 duration deciles (us):", np.round(np.percentile(r2[rounds >= 2], [0, 10, 25, 50, 75, 90, 100]) / 1e3, 1))
                print("corr(round1, round2) =", np.corrcoef(r1[rounds >= 2], r2[rounds >= 2])[0, 1])
