#!/usr/bin/env python3
"""Generate tests/golden/mocap_interp.npz: interpolated mocap reference poses computed with the
REFERENCE's own quaternion routines (/root/reference/src/transformations.py: quaternion_slerp
1270-1308, quaternion_from_euler 1100-1154, euler_from_quaternion 1089-1097) on the REFERENCE
loader's tables (tests/golden/mocap_<clip>.npz, produced by make_mocap_golden.py).

The reference ships the interpolation primitives (imported at dp_env_v3.py:18, unused) and the
loop / root-offset logic (MocapDM.play, mocap_v2.py:151-182) but never combines them; the combination
pinned here is the one documented in deepmimic_mujoco_b200/mocap.py::sample_tables:
  frame coordinate u = t / clip_dt, cycle = floor(u / (F-1)), k = floor(u - cycle (F-1)), alpha = frac;
  linear: root position (+ cycle * last frame's root xy), 1-DoF joints, all velocities;
  slerp:  root quaternion, 3-DoF joints via quaternion_from_euler('rxyz') -> slerp -> euler_from_quaternion.
Run in the build container only (needs /root/reference).
"""
import os
import sys
import warnings

import numpy as np

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
TRIPLES = (7, 10, 13, 17, 21, 25, 28, 32)   # qpos address of each 3-DoF joint
SINGLES = (16, 20, 24, 31)


def main():
    sys.path.insert(0, REF_SRC)
    warnings.simplefilter("ignore")
    import transformations as T
    rng = np.random.default_rng(7)
    out = {}
    for clip in ("walk", "spinkick", "dance_b", "run", "backflip"):
        g = np.load(os.path.join(HERE, "mocap_%s.npz" % clip))
        cfg, vel = g["data_config"], np.nan_to_num(g["data_vel"], nan=0.0, posinf=0.0, neginf=0.0)
        F = cfg.shape[0]
        us = np.concatenate([rng.uniform(0, 3.2 * (F - 1), size=40), [0.0, 1.0, F - 2.0, F - 1.0, F - 0.5, 2.0 * (F - 1) + 0.25]])
        qs, vs = [], []
        for u in us:
            cycle = int(np.floor(u / (F - 1)))
            uu = u - cycle * (F - 1)
            k = min(int(uu), F - 2)
            a = uu - k
            c0, c1 = cfg[k], cfg[k + 1]
            q = c0 + a * (c1 - c0)
            q[0:2] += cycle * cfg[F - 1, 0:2]
            # root quaternion: table order is wxyz, transformations uses xyzw; slerp is component-order agnostic
            q[3:7] = T.quaternion_slerp(c0[3:7], c1[3:7], a)
            for adr in TRIPLES:
                q0 = T.quaternion_from_euler(c0[adr], c0[adr + 1], c0[adr + 2], "rxyz")
                q1 = T.quaternion_from_euler(c1[adr], c1[adr + 1], c1[adr + 2], "rxyz")
                q[adr:adr + 3] = T.euler_from_quaternion(T.quaternion_slerp(q0, q1, a), "rxyz")
            qs.append(q)
            vs.append(vel[k] + a * (vel[k + 1] - vel[k]))
        out[clip + "_u"] = us
        out[clip + "_qpos"] = np.array(qs)
        out[clip + "_qvel"] = np.array(vs)
    np.savez_compressed(os.path.join(HERE, "mocap_interp.npz"), **out)
    print("wrote mocap_interp.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
