#!/usr/bin/env python3
"""(build container, CPU) How discriminating is the time-to-fall pin?  The oracle is run
under the reference's protocol with ONE model quantity perturbed at a time; the statistics the tests assert (mean,
spread, quartiles, Kolmogorov-Smirnov against the first 100 episodes of the reference's MuJoCo-produced monitor log)
are printed with a pass / FAIL verdict per variant.  usage: python tools/fall_time_sensitivity.py [episodes]"""
import copy
import os
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import oracle.pyoracle as po  # noqa: E402

EPISODES = int(sys.argv[1]) if len(sys.argv) > 1 else 400
mt = common.tables()
ref = common.ref_fall_lengths(100)


def run(model, seed=7, z_min=0.7, act_sd=1.0, noise=0.01):
    o = po.Oracle(model)
    rng = np.random.default_rng(seed)
    lens = []
    for ep in range(EPISODES):
        o.set_state(mt.qpos0 + rng.uniform(-noise, noise, mt.nq), rng.uniform(-noise, noise, mt.nv))
        for t in range(600):
            o.d.arr("ctrl")[:mt.nu] = act_sd * rng.normal(size=mt.nu)
            o.step()
            z = o.d.arr("com")[2]
            if z < z_min or z > 2.0:
                break
        lens.append(t + 1)
    return np.asarray(lens, dtype=np.float64)


def verdict(lens):
    ks = stats.ks_2samp(lens, ref)
    dq = np.abs(np.percentile(lens, [25, 50, 75]) - np.percentile(ref, [25, 50, 75])).max()
    ok = abs(lens.mean() - ref.mean()) < 3.0 and 0.75 < lens.std() / ref.std() < 1.25 and ks.pvalue > 0.01 and dq <= 3.0
    return (f"mean {lens.mean():6.2f} ({lens.mean() - ref.mean():+5.2f})  sd ratio {lens.std() / ref.std():4.2f}  "
            f"max quartile diff {dq:4.1f}  KS D {ks.statistic:5.3f} p {ks.pvalue:8.2g}  {'pass' if ok else 'FAIL'}")


def variant(**kw):
    m = copy.deepcopy(common.model())
    for k, f in kw.items():
        a = getattr(m, k)
        if hasattr(a, "__len__"):
            for i in range(len(a)):
                if hasattr(a[i], "__len__"):
                    for j in range(len(a[i])):
                        a[i][j] = f(a[i][j], i, j)
                else:
                    a[i] = f(a[i], i, 0)
        else:
            setattr(m, k, f(a, 0, 0))
    return m


def without_own_narrow_phase_pairs():
    """The model with the 15 candidate pairs removed whose narrow phase is this repo's own routine (capsule-box,
    box-box: DESIGN.md section 2 deviation 1) -- if a statistic does not move, it cannot depend on how those
    routines differ from MuJoCo's."""
    m = copy.deepcopy(common.model())
    own = lambda a, b: m.geom_type[a] == 6 and m.geom_type[b] in (3, 6)
    keep = [(m.pair_geom1[i], m.pair_geom2[i]) for i in range(m.npair)
            if not (own(m.pair_geom1[i], m.pair_geom2[i]) or own(m.pair_geom2[i], m.pair_geom1[i]))]
    for i, (a, b) in enumerate(keep):
        m.pair_geom1[i], m.pair_geom2[i] = a, b
    m.npair = len(keep)
    return m


CASES = [
    ("gravity x 0.8", dict(gravity=lambda v, i, j: 0.8 * v)),
    ("gravity x 1.25", dict(gravity=lambda v, i, j: 1.25 * v)),
    ("actuator gear x 0.5", dict(act_gear=lambda v, i, j: 0.5 * v)),
    ("actuator gear x 2", dict(act_gear=lambda v, i, j: 2.0 * v)),
    ("joint damping x 0", dict(dof_damping=lambda v, i, j: 0.0)),
    ("joint damping x 4", dict(dof_damping=lambda v, i, j: 4.0 * v)),
    ("body mass + inertia x 1.5", dict(body_mass=lambda v, i, j: 1.5 * v, body_inertia=lambda v, i, j: 1.5 * v)),
    ("armature + 0.05", dict(dof_armature=lambda v, i, j: v + (0.05 if i >= 6 else 0.0))),
    ("timestep x 0.5 (0.0083 s)", dict(timestep=lambda v, i, j: 0.5 * v)),
    ("ctrlrange x 2 (+-1.0)", dict(act_ctrlrange=lambda v, i, j: 2.0 * v)),
    ("no joint limits", dict(jnt_limited=lambda v, i, j: 0)),
    ("no contacts (0 pairs)", dict(npair=lambda v, i, j: 0)),
]


def main():
    print(f"reference (first 100 episodes): mean {ref.mean():.2f} sd {ref.std():.2f} quartiles {np.percentile(ref, [25, 50, 75])}")
    print(f"{EPISODES} oracle episodes per variant")
    print(f"{'as shipped':34s}", verdict(run(common.model())))
    print(f"{'as shipped, another seed':34s}", verdict(run(common.model(), seed=11)))
    for name, kw in CASES:
        print(f"{name:34s}", verdict(run(variant(**kw))))
    print(f"{'no capsule-box / box-box pairs':34s}", verdict(run(without_own_narrow_phase_pairs())))
    print(f"{'termination height 0.6':34s}", verdict(run(common.model(), z_min=0.6)))
    print(f"{'termination height 0.8':34s}", verdict(run(common.model(), z_min=0.8)))
    print(f"{'action sd 0.25 (not the init policy)':34s}", verdict(run(common.model(), act_sd=0.25)))


if __name__ == "__main__":
    main()
