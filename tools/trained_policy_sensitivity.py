#!/usr/bin/env python3
"""(build container, CPU) How discriminating is the trained-policy pin?  The policy the
reference's TRPO run trained in MuJoCo 2.0 (tests/golden/ref_trained_policy.npz) is run in the oracle under the
reference's protocol with ONE model quantity perturbed at a time; the survival statistics are printed next to the
episodes the reference's own monitor recorded around the time the checkpoint was written.
usage: python tools/trained_policy_sensitivity.py [episodes]"""
import os
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import common  # noqa: E402
import oracle.pyoracle as po  # noqa: E402

EPISODES = int(sys.argv[1]) if len(sys.argv) > 1 else 300
mt = common.tables()
pol = common.RefTrainedPolicy()
ref = pol.monitor_window(50)


def run(model, seed=5, z_min=0.7, noise=0.01, stochastic=True, obs_scale=1.0):
    o = po.Oracle(model)
    rng = np.random.default_rng(seed)
    lens = []
    for ep in range(EPISODES):
        o.set_state(mt.qpos0 + rng.uniform(-noise, noise, mt.nq), rng.uniform(-noise, noise, mt.nv))
        for t in range(3000):
            ob = np.concatenate([o.qpos[7:], o.qvel[6:]]) * obs_scale
            a = pol.mean_action(ob[None])[0]
            o.d.arr("ctrl")[:mt.nu] = a + pol.act_std * rng.normal(size=mt.nu) if stochastic else a
            o.step()
            z = o.d.arr("com")[2]
            if z < z_min or z > 2.0:
                break
        lens.append(t + 1)
    return np.asarray(lens, dtype=np.float64)


def verdict(lens):
    ks = stats.ks_2samp(lens, ref)
    ok = common.trained_policy_verdict(lens, ref)
    return (f"mean {lens.mean():7.1f} ({lens.mean() / ref.mean() - 1:+5.0%})  median {np.median(lens):6.1f} "
            f"({np.median(lens) / np.median(ref) - 1:+5.0%})  KS D {ks.statistic:5.3f} p {ks.pvalue:8.2g}  "
            f"{'pass' if ok else 'FAIL'}")


if __name__ == "__main__":
    from fall_time_sensitivity import CASES, variant, without_own_narrow_phase_pairs
    MILD = [                                  # perturbations the random-policy pin (section 2 (ix)) cannot resolve
        ("gravity x 0.9", dict(gravity=lambda v, i, j: 0.9 * v)),
        ("gravity x 1.1", dict(gravity=lambda v, i, j: 1.1 * v)),
        ("actuator gear x 0.8", dict(act_gear=lambda v, i, j: 0.8 * v)),
        ("actuator gear x 1.25", dict(act_gear=lambda v, i, j: 1.25 * v)),
        ("body mass + inertia x 0.8", dict(body_mass=lambda v, i, j: 0.8 * v, body_inertia=lambda v, i, j: 0.8 * v)),
        ("body mass + inertia x 1.25", dict(body_mass=lambda v, i, j: 1.25 * v, body_inertia=lambda v, i, j: 1.25 * v)),
        ("body inertia x 2 (mass kept)", dict(body_inertia=lambda v, i, j: 2.0 * v)),
        ("joint damping x 2", dict(dof_damping=lambda v, i, j: 2.0 * v)),
        ("joint damping x 0.5", dict(dof_damping=lambda v, i, j: 0.5 * v)),
        ("armature + 0.02", dict(dof_armature=lambda v, i, j: v + (0.02 if i >= 6 else 0.0))),
        ("timestep x 0.8", dict(timestep=lambda v, i, j: 0.8 * v)),
        ("timestep x 1.25", dict(timestep=lambda v, i, j: 1.25 * v)),
    ]
    INTERNALS = [                             # parameters of the restated constraint model / solver (no MJCF attribute)
        ("PGS iterations 20", dict(iterations=lambda v, i, j: 20)),
        ("PGS iterations 200", dict(iterations=lambda v, i, j: 200)),
        ("PGS tolerance 0 (never exits early)", dict(tolerance=lambda v, i, j: 0.0)),
        ("invweight0 x 0.5 (stiffer rows)", dict(body_invweight0=lambda v, i, j: 0.5 * v, dof_invweight0=lambda v, i, j: 0.5 * v)),
        ("invweight0 x 2 (softer rows)", dict(body_invweight0=lambda v, i, j: 2 * v, dof_invweight0=lambda v, i, j: 2 * v)),
        ("solimp dmin/dmax 0.99", dict(solimp=lambda v, i, j: 0.99 if i < 2 else v)),
        ("solimp dmin/dmax 0.8", dict(solimp=lambda v, i, j: 0.8 if i < 2 else v)),
        ("solref timeconst 0.1", dict(solref=lambda v, i, j: 0.1 if i == 0 else v)),
        ("contact margin 0.01", dict(margin=lambda v, i, j: 0.01)),
        ("contact margin 0", dict(margin=lambda v, i, j: 0.0)),
    ]
    print(f"reference monitor, {len(ref)} episodes around the checkpoint: mean {ref.mean():.1f} sd {ref.std():.1f} "
          f"quartiles {np.percentile(ref, [25, 50, 75])}")
    print(f"{EPISODES} oracle episodes per variant")
    print(f"{'as shipped':34s}", verdict(run(common.model())))
    print(f"{'as shipped, another seed':34s}", verdict(run(common.model(), seed=11)))
    for name, kw in CASES + MILD + INTERNALS:
        print(f"{name:34s}", verdict(run(variant(**kw))), flush=True)
    print(f"{'no capsule-box / box-box pairs':34s}", verdict(run(without_own_narrow_phase_pairs())))
    print(f"{'termination height 0.6':34s}", verdict(run(common.model(), z_min=0.6)))
    print(f"{'termination height 0.8':34s}", verdict(run(common.model(), z_min=0.8)))
    print(f"{'deterministic policy (mode)':34s}", verdict(run(common.model(), stochastic=False)))
    print(f"{'observation x 0.9':34s}", verdict(run(common.model(), obs_scale=0.9)))
