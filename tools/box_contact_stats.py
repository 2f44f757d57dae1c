#!/usr/bin/env python3
"""(build container, CPU) How much of the workload touches DESIGN.md section 2 deviation 1?  The capsule-box and box-box
narrow phases are this repo's own restatements, not MuJoCo's mjc_CapsuleBox / mjc_BoxBox; every other pair type on
this model (plane-sphere/capsule/box, sphere-sphere, sphere-capsule, capsule-capsule, sphere-box) is a closed form.
This tool runs the float64 oracle through the benchmark workloads (mocap RSI, U(-0.5, 0.5) actions, CoM termination,
auto-reset: BASELINE configs 2 / 3 / 5 clips) and through the reference's own protocol (standing pose, the policy the
reference trained in MuJoCo) and counts, over the contacts of the last RK4 stage of every env step, how many belong to
which pair class and in what fraction of env steps an own-algorithm contact is active.
usage: python tools/box_contact_stats.py [envs=32] [steps=400]"""
import ctypes as C
import os
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import oracle.pyoracle as po  # noqa: E402
from deepmimic_mujoco_b200.model_blob import default_config  # noqa: E402
from deepmimic_mujoco_b200.refaux import compute_ref_aux  # noqa: E402
from deepmimic_mujoco_b200.sim import load_motions, make_mocap_struct  # noqa: E402

NAMES = {0: "plane", 2: "sphere", 3: "capsule", 6: "box"}
OWN = {("box", "capsule"), ("box", "box")}


def classify(m, d, counts):
    """Adds the contacts of d to counts; returns True if a capsule-box / box-box contact is among them."""
    own = False
    for k in range(d.ncon):
        c = d.contact[k]
        pair = tuple(sorted((NAMES[m.geom_type[c.geom1]], NAMES[m.geom_type[c.geom2]])))
        counts[pair] += 1
        own |= pair in OWN
    return own


def report(title, steps, own_steps, counts):
    total = sum(counts.values())
    print(f"\n{title}: {steps} env steps, {total} contacts ({total / max(steps, 1):.2f} per step)")
    for pair, n in counts.most_common():
        tag = "   <- own algorithm (deviation 1)" if pair in OWN else ""
        print(f"  {pair[0]:>7s} - {pair[1]:<7s} {n:8d}  {100.0 * n / max(total, 1):6.2f} % of contacts{tag}")
    print(f"  env steps with an own-algorithm contact active: {own_steps} = {100.0 * own_steps / max(steps, 1):.3f} %")


def benchmark_workload(motion, n_envs, n_steps, seed=0):
    m, L = common.model(), po.lib()
    cfg = default_config(reward_mode=4, reset_mode=0, auto_reset=1)
    mcs, keep = make_mocap_struct(load_motions([motion]), compute_ref_aux([motion]))
    rng = np.random.default_rng(seed)
    counts, steps, own_steps = Counter(), 0, 0
    obs, rew = np.zeros(256), C.c_double()
    for i in range(n_envs):
        e = po.DmoEnv()
        L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), seed, i, 0)
        L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 0)
        for t in range(n_steps):
            a = np.ascontiguousarray(rng.uniform(-0.5, 0.5, 28))
            done = L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(a), po.dptr(obs), C.byref(rew))
            if not done:                       # after an auto-reset e.d no longer holds the step's contacts
                steps += 1
                own_steps += classify(m, e.d, counts)
    report(f"{motion}: RSI + U(-0.5, 0.5) actions + CoM termination + auto-reset", steps, own_steps, counts)


def trained_policy_workload(n_episodes, seed=0):
    m, mt = common.model(), common.tables()
    o, pol, rng = po.Oracle(m), common.RefTrainedPolicy(), np.random.default_rng(seed)
    counts, steps, own_steps = Counter(), 0, 0
    for ep in range(n_episodes):
        o.set_state(mt.qpos0 + rng.uniform(-0.01, 0.01, mt.nq), rng.uniform(-0.01, 0.01, mt.nv))
        for t in range(3000):
            ob = np.concatenate([o.qpos[7:], o.qvel[6:]])
            o.d.arr("ctrl")[:mt.nu] = pol.mean_action(ob[None])[0] + pol.act_std * rng.normal(size=mt.nu)
            o.step()
            steps += 1
            own_steps += classify(m, o.d, counts)
            z = o.d.arr("com")[2]
            if z < 0.7 or z > 2.0:
                break
    report("standing pose + the reference's MuJoCo-trained policy (section 2 (x) protocol)", steps, own_steps, counts)


if __name__ == "__main__":
    n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    for motion in ("walk", "spinkick", "dance_b"):
        benchmark_workload(motion, n_envs, n_steps)
    trained_policy_workload(max(8, n_envs // 2))
