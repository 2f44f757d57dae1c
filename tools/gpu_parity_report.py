#!/usr/bin/env python3
"""Stage-by-stage CUDA vs float64-oracle parity report (run on the GPU box).

Prints, per scenario, the worst absolute / relative error of every stage output of one
mj_forward evaluation and of one full RK4 env step.  Diagnostic tool; tests/test_gpu_parity.py
holds the asserted tolerances.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import oracle.pyoracle as po  # noqa: E402
from deepmimic_mujoco_b200.sim import BatchedSim  # noqa: E402


def oracle_forward(o, mt, qpos, qvel, ctrl, warm):
    o.set_state(qpos, qvel, ctrl, warm)
    o.forward()
    d = o.d
    nv, nb, nM = mt.nv, mt.nbody, mt.nM
    out = dict(xpos=d.arr("xpos")[:nb].copy(), xquat=d.arr("xquat")[:nb].copy(), xipos=d.arr("xipos")[:nb].copy(),
               com=d.arr("com").copy(), qM=d.arr("qM")[:nM].copy(), qLD=d.arr("qLD")[:nM].copy(),
               qfrc_bias=d.arr("qfrc_bias")[:nv].copy(), qfrc_smooth=d.arr("qfrc_smooth")[:nv].copy(),
               qacc_smooth=d.arr("qacc_smooth")[:nv].copy(), ncon=d.ncon, nefc=d.nefc, iter=d.solver_iter,
               efc_pos=d.arr("efc_pos")[:d.nefc].copy(), efc_R=d.arr("efc_R")[:d.nefc].copy(),
               efc_aref=d.arr("efc_aref")[:d.nefc].copy(), efc_b=d.arr("efc_b")[:d.nefc].copy(),
               efc_force=d.arr("efc_force")[:d.nefc].copy(),
               efc_AR_diag=np.array([d.efc_AR[r][r] for r in range(d.nefc)]),
               qacc=d.arr("qacc")[:nv].copy(), cvel=d.arr("cvel")[:nb].copy(), z_com=d.arr("com")[2],
               contact=np.array([[c.dist, *c.pos, *c.frame, c.geom1, c.geom2, c.dim] for c in d.contact[:d.ncon]]).reshape(-1, 16),
               warm=d.arr("qacc_warmstart")[:nv].copy(), flags=d.flags)
    return out


def report(name, sim, o, mt, qpos, qvel, ctrl, warm=None, verbose=True):
    n = qpos.shape[0]
    warm = np.zeros((n, mt.nv)) if warm is None else warm
    sim.set_state(qpos, qvel, warm)
    g = sim.forward_debug(torch.tensor(ctrl, dtype=torch.float32))
    torch.cuda.synchronize()
    keys = ["xpos", "xquat", "xipos", "com", "qM", "qLD", "qfrc_bias", "qfrc_smooth", "qacc_smooth", "cvel",
            "efc_pos", "efc_R", "efc_aref", "efc_b", "efc_AR_diag", "efc_force", "qacc"]
    worst = {k: (0.0, 0.0) for k in keys}
    mism = 0
    for i in range(n):
        r = oracle_forward(o, mt, qpos[i], qvel[i], ctrl[i], warm[i])
        if r["ncon"] != g["ncon"][i] or r["nefc"] != g["nefc"][i]:
            mism += 1
            print(f"  [{name} env {i}] ncon/nefc mismatch: oracle {r['ncon']}/{r['nefc']} gpu {g['ncon'][i]}/{g['nefc'][i]}")
            continue
        for k in keys:
            a = np.asarray(r[k], dtype=np.float64).ravel()
            b = np.asarray(g[k][i], dtype=np.float64).ravel()[: a.size]
            if a.size == 0:
                continue
            err = np.abs(a - b).max()
            rel = err / max(1.0, np.abs(a).max())
            if err > worst[k][0]:
                worst[k] = (err, rel)
        if r["ncon"]:
            cerr = np.abs(r["contact"] - g["contact"][i][: r["ncon"]]).max()
            worst.setdefault("contact", (0.0, 0.0))
            if cerr > worst["contact"][0]:
                worst["contact"] = (cerr, cerr)
            if cerr > 1e-3 and verbose:
                np.set_printoptions(precision=5, suppress=True, linewidth=200)
                print(f"  [{name} env {i}] contact mismatch {cerr:.3e}")
                for c in range(r["ncon"]):
                    a, b = r["contact"][c], g["contact"][i][c]
                    print("     oracle g", int(a[13]), int(a[14]), "dist", a[0], "pos", a[1:4], "n", a[4:7])
                    print("     gpu    g", int(b[13]), int(b[14]), "dist", b[0], "pos", b[1:4], "n", b[4:7])
    print(f"== {name}: n={n} mismatched={mism}  ncon range {g['ncon'].min()}..{g['ncon'].max()}  nefc {g['nefc'].min()}..{g['nefc'].max()} iters {g['iter'].min()}..{g['iter'].max()}")
    if verbose:
        for k, (e, rel) in worst.items():
            print(f"   {k:12s} abs {e:.3e}  rel {rel:.3e}")
    return worst


def step_report(name, sim, o, mt, qpos, qvel, ctrl, warm=None):
    n = qpos.shape[0]
    warm = np.zeros((n, mt.nv)) if warm is None else warm
    sim.set_state(qpos, qvel, warm)
    act = torch.tensor(ctrl, dtype=torch.float32, device=sim.device)
    obs, rew, done = sim.step(act)
    torch.cuda.synchronize()
    gq, gv, gw = sim.get_state()
    eq = ev = 0.0
    dmis = 0
    for i in range(n):
        o.set_state(qpos[i], qvel[i], ctrl[i], warm[i])
        o.step()
        zc = o.d.arr("com")[2]
        od = (zc < 0.7) or (zc > 2.0)
        eq = max(eq, np.abs(o.qpos - gq[i]).max())
        ev = max(ev, (np.abs(o.qvel - gv[i]) / np.maximum(1.0, np.abs(o.qvel))).max())
        dmis += int(od != bool(done[i].item()))
    print(f"== step {name}: n={n} |dqpos| {eq:.3e}  |dqvel|rel {ev:.3e}  done mismatches {dmis}  flags {sim.flags.unique().tolist()}")
    return eq, ev


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    mt = common.tables()
    n = 64
    sim = BatchedSim(n, motions=("walk",), seed=1)
    print("launch:", sim.launch_info())
    o = po.Oracle(common.model())
    zeros = np.zeros((n, mt.nu))
    ctrl = common.f32(rng.uniform(-0.6, 0.6, size=(n, mt.nu)))
    q, v = common.airborne_states(rng, n)
    report("airborne", sim, o, mt, q, v, ctrl)
    step_report("airborne", sim, o, mt, q, v, ctrl)
    if os.environ.get("ONLY_AIR"):
        return
    q, v = common.standing_states(rng, n)
    report("standing", sim, o, mt, q, v, ctrl)
    step_report("standing", sim, o, mt, q, v, ctrl)
    idx = rng.integers(0, 39, size=n)
    q, v = common.mocap_states("walk", idx)
    report("mocap-walk", sim, o, mt, q, v, zeros)
    step_report("mocap-walk", sim, o, mt, q, v, ctrl)
    q, v, w = common.rollout_states(rng, n)
    report("rollout", sim, o, mt, q, v, ctrl, w)
    step_report("rollout", sim, o, mt, q, v, ctrl, w)


if __name__ == "__main__":
    main()
