#!/usr/bin/env python3
"""Offline look at the full-size parity outliers saved by tests/test_gpu_parity_full.py (gpurun_out/parity_outliers_*.npz):
is the ORACLE itself sensitive to fp32-sized input perturbations on these states (a branch / contact flip), i.e. is the
single-step map ill-conditioned there, or does the CUDA result lie outside what perturbations of the inputs can explain?
Runs in the build container (CPU only)."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import oracle.pyoracle as po  # noqa: E402

o = po.Oracle(common.model())
rng = np.random.default_rng(0)
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_outliers_*.npz"))):
    z = np.load(f)
    print("==", os.path.basename(f), "outliers:", len(z["idx"]))
    for k in range(len(z["idx"])):
        q, v, w, a = z["qpos"][k], z["qvel"][k], z["warm"][k], z["action"][k]
        def step(qq, vv):
            o.set_state(qq, vv, a, w); o.step()
            return o.qpos.copy(), o.qvel.copy(), o.d.ncon, o.d.nefc, o.d.solver_iter
        q0, v0, nc0, ne0, it0 = step(q, v)
        dev_gpu = np.abs(z["gpu_qvel"][k] - v0).max()
        spread, ncs = 0.0, set()
        for _ in range(24):
            qp = q * (1 + rng.uniform(-1, 1, q.shape) * 6e-8) + rng.uniform(-1, 1, q.shape) * 1e-8
            vp = v * (1 + rng.uniform(-1, 1, v.shape) * 6e-8)
            q1, v1, nc1, ne1, it1 = step(qp, vp)
            spread = max(spread, np.abs(v1 - v0).max()); ncs.add((nc1, ne1))
        print(f"  env {int(z['idx'][k]):6d}: |gpu - oracle| qvel {dev_gpu:9.2e} | oracle spread under 1-ulp input noise {spread:9.2e} "
              f"| ncon/nefc {nc0}/{ne0} iter {it0} variants {sorted(ncs)} | max|qvel| {np.abs(v).max():.1f}")
