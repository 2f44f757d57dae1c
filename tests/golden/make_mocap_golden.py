#!/usr/bin/env python3
"""Generate tests/golden/mocap_<clip>.npz by running the REFERENCE loader
(/root/reference/src/mujoco/mocap_v2.py MocapDM, mocap_util.py, transformations.py) here.

pyquaternion (un-vendored, unpinned, absent) is replaced by the minimal shim below, which
restates exactly the API the reference uses: Quaternion(w,x,y,z),
Quaternion(matrix=R), ``*``, ``.conjugate``, ``.angle``, ``.axis``, ``.elements``.
Everything else executed is the reference's own code.  Run in the build container only.
Also records transformations.py doctest known-answers used as KATs.
"""
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


class Quaternion:
    def __init__(self, *args, matrix=None):
        if matrix is not None:
            R = np.asarray(matrix, dtype=float)
            assert np.allclose(R @ R.T, np.eye(3), rtol=1e-5, atol=1e-8) and np.isclose(np.linalg.det(R), 1.0)
            t = np.trace(R)
            assert t > 0  # both alignment matrices have trace 1
            s = np.sqrt(t + 1.0) * 2
            self.q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
        else:
            self.q = np.array(args, dtype=float)

    def __mul__(self, o):
        a, b = self.q, o.q
        return Quaternion(a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3],
                          a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
                          a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1],
                          a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0])

    @property
    def conjugate(self):
        return Quaternion(self.q[0], -self.q[1], -self.q[2], -self.q[3])

    @property
    def elements(self):
        return self.q

    def _unit(self):
        n = np.linalg.norm(self.q)
        return self.q / n if n > 0 and abs(1.0 - n) >= 1e-14 else self.q

    @property
    def angle(self):
        q = self._unit()
        th = 2.0 * np.arctan2(np.linalg.norm(q[1:]), q[0])
        r = ((th + np.pi) % (2 * np.pi)) - np.pi
        return np.pi if r == -np.pi else r

    @property
    def axis(self):
        q = self._unit()
        n = np.linalg.norm(q[1:])
        return np.zeros(3) if n < 1e-17 else q[1:] / n


def main():
    shim = types.ModuleType("pyquaternion")
    shim.Quaternion = Quaternion
    sys.modules["pyquaternion"] = shim
    sys.path.insert(0, REF_SRC)
    import warnings
    warnings.simplefilter("ignore")
    from mujoco.mocap_v2 import MocapDM  # the reference's local package named `mujoco`
    import transformations as T
    for clip in ("walk", "spinkick", "dance_b", "run", "backflip"):
        m = MocapDM()
        m.load_mocap(os.path.join(REF_SRC, "mujoco/motions/humanoid3d_%s.txt" % clip))
        np.savez_compressed(os.path.join(HERE, "mocap_%s.npz" % clip), dt=np.float64(m.dt),
                            data=np.asarray(m.data), data_config=np.asarray(m.data_config),
                            data_vel=np.asarray(m.data_vel), durations=np.asarray(m.durations))
        print(clip, np.asarray(m.data_config).shape, np.asarray(m.data_vel).shape, m.dt)
    # the other ten clips the reference ships: a digest instead of the full tables (every 5th frame + the last one of
    # data_config / data_vel, and float64 column sums over ALL frames), tests/golden/mocap_digest.npz
    digest = {}
    for clip in ("cartwheel", "crawl", "dance_a", "getup_facedown", "getup_faceup", "jump", "kick", "punch", "roll", "spin"):
        m = MocapDM()
        m.load_mocap(os.path.join(REF_SRC, "mujoco/motions/humanoid3d_%s.txt" % clip))
        cfg, vel = np.asarray(m.data_config), np.asarray(m.data_vel)
        rows = np.unique(np.r_[np.arange(0, len(cfg), 5), len(cfg) - 1])
        digest.update({clip + "/dt": np.float64(m.dt), clip + "/rows": rows, clip + "/cfg_rows": cfg[rows],
                       clip + "/vel_rows": vel[rows], clip + "/cfg_colsum": cfg.sum(0), clip + "/vel_colsum": vel.sum(0),
                       clip + "/cfg_abssum": np.abs(cfg).sum(0), clip + "/nframes": np.int64(len(cfg))})
        print(clip, cfg.shape, m.dt)
    np.savez_compressed(os.path.join(HERE, "mocap_digest.npz"), **digest)
    # known answers from transformations.py doctests (lines 1092-1093, 1106-1107, 1231-1232)
    kat = dict(
        euler_from_quaternion=np.array(T.euler_from_quaternion([0.06146124, 0, 0, 0.99810947])),
        quaternion_from_euler=np.array(T.quaternion_from_euler(1, 2, 3, 'ryxz')),
        quaternion_multiply=np.array(T.quaternion_multiply([1, -2, 3, 4], [-5, 6, 7, 8])),
    )
    rng = np.random.default_rng(0)
    qs = rng.normal(size=(64, 4))
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    kat["rand_quat_wxyz"] = qs
    kat["rand_euler_rxyz"] = np.array([T.euler_from_quaternion([q[1], q[2], q[3], q[0]], axes='rxyz') for q in qs])
    np.savez_compressed(os.path.join(HERE, "transformations_kat.npz"), **kat)
    print("KATs:", {k: np.round(v, 6) for k, v in kat.items() if v.size < 8})


if __name__ == "__main__":
    main()
