#!/usr/bin/env python3
"""Event-timed step time next to the in-kernel CTA span (DMB_TRACE=1), for fresh vs cycled action tensors.
Run on the GPU box: DMB_TRACE=1 python tools/gpu_time_probe.py [envs]"""
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
g = torch.Generator(device="cuda"); g.manual_seed(1234)
pool = torch.rand(16, E, sim.nu, device="cuda", generator=g) - 0.5
buf = np.zeros((256, 8), dtype=np.int64)
tracing = os.environ.get("DMB_TRACE") == "1"
for mode in ("cycled", "fresh", "cycled", "fresh"):
    ev, span, mean_cta = [], [], []
    for t in range(120):
        act = pool[t % 16] if mode == "cycled" else torch.rand(E, sim.nu, device="cuda", generator=g) - 0.5
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(act); e1.record()
        torch.cuda.synchronize()
        if t >= 40:
            ev.append(e0.elapsed_time(e1) * 1e3)
            if tracing:
                n = sim.L.dmb_get_trace(sim.handle, buf.ctypes.data_as(C.c_void_p), 256)
                tr = buf[:n].astype(np.float64)
                span.append((tr[:, 6].max() - tr[:, 0].min()) / 1e3)
                mean_cta.append((tr[:, 6] - tr[:, 0]).mean() / 1e3)
    nefc = None
    print(f"{mode:7s} events {np.mean(ev):7.1f} us (min {np.min(ev):6.1f} max {np.max(ev):6.1f})"
          + (f" | in-kernel span {np.mean(span):7.1f} (max {np.max(span):6.1f}) mean CTA {np.mean(mean_cta):6.1f}" if tracing else ""))
