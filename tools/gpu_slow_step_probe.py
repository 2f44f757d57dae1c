#!/usr/bin/env python3
"""Which env makes a step slow?  DMB_TRACE=1: every env reports (rows, PGS sweeps, time until its warp was done) in
the upper bits of `flags`; per step we print the kernel span and the slowest envs, and save the pre-step state of the
slowest env of the slowest steps for offline analysis with the oracle.  Run on the GPU box:
DMB_TRACE=1 python tools/gpu_slow_step_probe.py [envs] [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["DMB_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 160
env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
sim = env.sim
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1234)
pool = torch.rand(16, E, sim.nu, device="cuda", generator=g) - 0.5
buf = np.zeros((256, 8), dtype=np.int64)
saved = []
for t in range(T):
    q0, v0, w0 = sim.qpos.clone(), sim.qvel.clone(), sim.warm.clone()
    env.step(pool[t % 16])
    torch.cuda.synchronize()
    n = sim.L.dmb_get_trace(sim.handle, buf.ctypes.data_as(C.c_void_p), 256)
    tr = buf[:n].astype(np.float64)
    span = (tr[:, 6].max() - tr[:, 0].min()) / 1e3
    cta_end = (tr[:, 6] - tr[:, 0]) / 1e3
    f = sim.flags.cpu().numpy()
    rows, sweeps, tw = (f >> 8) & 63, (f >> 14) & 255, ((f >> 22) & 1023) * 2.0
    top = np.argsort(tw)[-3:][::-1]
    if t >= 30:
        print(f"step {t:3d} span {span:6.1f} us | CTA end p50 {np.median(cta_end):6.1f} max {cta_end.max():6.1f} | slowest envs "
              + " ".join(f"[env {i} {tw[i]:.0f}us rows {rows[i]} sweeps {sweeps[i]} flags {f[i] & 7}]" for i in top)
              + f" | envs > 24 rows: {int((rows > 24).sum())} > 32: {int((rows > 32).sum())} | done {int(sim.done.sum())}")
        if span > 600 and len(saved) < 12:
            i = int(top[0])
            saved.append(dict(step=t, env=i, span=span, qpos=q0[i].double().cpu().numpy(), qvel=v0[i].double().cpu().numpy(),
                              warm=w0[i].double().cpu().numpy(), action=pool[t % 16][i].double().cpu().numpy(),
                              rows=int(rows[i]), sweeps=int(sweeps[i]), tw=float(tw[i])))
if saved:
    np.savez(os.path.join(ROOT, "gpurun_out", "slow_step_envs.npz"), **{k: np.array([s[k] for s in saved]) for k in saved[0]})
