"""Time-based mocap phase with interpolation (phase_mode 1) and the DeepMimic 197-d state (obs_mode 1):
host restatement and float64 oracle against golden vectors produced with the REFERENCE's own
transformations.quaternion_slerp / quaternion_from_euler / euler_from_quaternion
(tests/golden/make_interp_golden.py) -- CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import common
import oracle.pyoracle as po
from deepmimic_mujoco_b200 import mocap
from deepmimic_mujoco_b200.model_blob import default_config
from deepmimic_mujoco_b200.refaux import compute_ref_aux
from deepmimic_mujoco_b200.sim import load_motions, make_mocap_struct

CLIPS = ("walk", "spinkick", "dance_b", "run", "backflip")


def _golden():
    return np.load(os.path.join(common.GOLDEN, "mocap_interp.npz"))


@pytest.mark.parametrize("name", CLIPS)
def test_host_sample_matches_reference_routines(name):
    g, c = _golden(), common.clip(name)
    vel = np.nan_to_num(c.data_vel)
    for u, q, v in zip(g[name + "_u"], g[name + "_qpos"], g[name + "_qvel"]):
        qq, vv, _, _, _ = mocap.sample_tables(c.data_config, vel, float(u))
        assert np.abs(qq - q).max() < 1e-12 and np.abs(vv - v).max() < 1e-11


def _oracle_sample(m, mcs, clip, u):
    q, v, ph = np.zeros(40), np.zeros(40), C.c_double()
    po.lib().dmo_mocap_sample(C.byref(m), C.byref(mcs), clip, float(u), po.dptr(q), po.dptr(v), C.byref(ph))
    return q[: m.nq].copy(), v[: m.nv].copy(), ph.value


@pytest.mark.parametrize("name", CLIPS)
def test_oracle_sample_matches_reference_routines(name):
    """dmo_mocap_sample consumes fp32-rounded tables (as the CUDA path stores them): exact against the host
    restatement on the same rounded tables, and within the rounding against the reference-routine golden."""
    g, m = _golden(), common.model()
    mc = load_motions([name])
    mcs, keep = make_mocap_struct(mc)
    cfg32, vel32 = common.f32(keep[0]), common.f32(keep[1])
    for u, q, v in zip(g[name + "_u"], g[name + "_qpos"], g[name + "_qvel"]):
        oq, ov, ph = _oracle_sample(m, mcs, 0, u)
        hq, hv, _, _, _ = mocap.sample_tables(cfg32, vel32, float(u))
        assert np.abs(oq - hq).max() < 1e-12 and np.abs(ov - hv).max() < 1e-11
        assert np.abs(oq - q).max() < 3e-6 and np.abs(ov - v).max() < 2e-6 * max(1.0, np.abs(v).max())
        assert 0.0 <= ph < 1.0


def test_sample_properties():
    m = common.model()
    c = common.clip("walk")
    mcs, keep = make_mocap_struct(load_motions(["walk"]))
    F = len(c)
    cfg32 = common.f32(c.data_config)
    for k in (0, 3, 17, F - 2):   # integer frame coordinates reproduce the table rows (quaternions normalised)
        q, v, ph = _oracle_sample(m, mcs, 0, float(k))
        ref = cfg32[k].copy(); ref[3:7] /= np.linalg.norm(ref[3:7])
        assert np.abs(q - ref).max() < 1e-9
        assert abs(ph - k / (F - 1)) < 1e-12
    # loop wrap: one cycle later the pose repeats with the root shifted by the last frame's root xy
    for x in (0.0, 0.4, 5.25, F - 1.5):
        q0, v0, p0 = _oracle_sample(m, mcs, 0, x)
        q1, v1, p1 = _oracle_sample(m, mcs, 0, x + 2 * (F - 1))
        d = q1 - q0
        assert np.abs(d[:2] - 2 * cfg32[F - 1][:2]).max() < 1e-9 and np.abs(d[2:]).max() < 1e-9
        assert np.abs(v1 - v0).max() < 1e-9 and abs(p1 - p0) < 1e-9
    # slerp end points and midpoint against the closed form
    a, b = np.array([1.0, 0, 0, 0]), np.array([np.cos(0.4), 0, np.sin(0.4), 0])
    assert np.allclose(mocap.quat_slerp(a, b, 0.5), [np.cos(0.2), 0, np.sin(0.2), 0], atol=1e-15)
    assert np.allclose(mocap.quat_slerp(a, -b, 0.25), [np.cos(0.1), 0, np.sin(0.1), 0], atol=1e-15)   # shortest path


def _env(m, cfg, mcs, clip=0, seed=3, env_id=0):
    e = po.DmoEnv()
    L = po.lib()
    L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), seed, env_id, clip)
    L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 0)
    return e


@pytest.mark.parametrize("reward_mode", [1, 2, 3, 4])
def test_phase_modes_agree_when_clip_rate_is_one(reward_mode):
    """With clip_dt == timestep the time-based reference lands exactly on table frames: the phase_mode 1 reward
    must equal the phase_mode 0 reward evaluated on that frame (ties the two lookups together)."""
    m = common.model()
    aux = compute_ref_aux(["dance_b"])
    mcs, keep = make_mocap_struct(load_motions(["dance_b"]), aux)
    mcs.clip_dt[0] = m.timestep
    F = mcs.clip_len[0]
    L = po.lib()
    rng = np.random.default_rng(1)
    c1 = default_config(reward_mode=reward_mode, phase_mode=1)
    c0 = default_config(reward_mode=reward_mode, phase_mode=0)
    e1, e0 = _env(m, c1, mcs), _env(m, c0, mcs)
    assert e1.idx_init == e0.idx_init
    obs, r1, r0 = np.zeros(256), C.c_double(), C.c_double()
    for t in range(12):
        a = rng.uniform(-0.5, 0.5, 28)
        # phase_mode 1 compares the post-step state with the frame of the post-step time (idx_init + t + 1);
        # point phase_mode 0 at the same frame (modes 2/3 advance before evaluating, modes 1/4 after)
        tgt = (e1.idx_init + t + 1) % (F - 1)
        e0.idx_curr = tgt - 1 if reward_mode in (2, 3) else tgt
        if e0.idx_curr < 0:
            e0.idx_curr += F
        np.ctypeslib.as_array(e0.d.qpos)[:] = np.ctypeslib.as_array(e1.d.qpos)
        np.ctypeslib.as_array(e0.d.qvel)[:] = np.ctypeslib.as_array(e1.d.qvel)
        np.ctypeslib.as_array(e0.d.qacc_warmstart)[:] = np.ctypeslib.as_array(e1.d.qacc_warmstart)
        L.dmo_env_step(C.byref(m), C.byref(c1), C.byref(mcs), C.byref(e1), po.dptr(a), po.dptr(obs), C.byref(r1))
        L.dmo_env_step(C.byref(m), C.byref(c0), C.byref(mcs), C.byref(e0), po.dptr(a), po.dptr(obs), C.byref(r0))
        cyc = (e1.idx_init + t + 1) // (F - 1)
        if cyc == 0:   # after a wrap the time-based root carries the accumulated xy offset by design
            assert abs(r1.value - r0.value) < 1e-6, (t, r1.value, r0.value)
        assert e1.idx_curr == tgt
        assert 0.0 < r1.value + (3.0 if reward_mode in (2, 3) else 0.0)


def test_phase_mode1_rollout_walk_half_rate():
    """walk is a 30 Hz clip stepped at 60 Hz: the reference frame advances every other env step (App. F #6 fixed)."""
    m = common.model()
    mcs, keep = make_mocap_struct(load_motions(["walk"]), compute_ref_aux(["walk"]))
    cfg = default_config(reward_mode=4, phase_mode=1)
    e = _env(m, cfg, mcs, seed=9)
    L = po.lib()
    obs, r = np.zeros(256), C.c_double()
    rate = m.timestep / mcs.clip_dt[0]
    F = mcs.clip_len[0]
    for t in range(30):
        a = np.zeros(28)
        L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(a), po.dptr(obs), C.byref(r))
        u = e.idx_init + (t + 1) * rate
        assert e.idx_curr == min(int(u % (F - 1)), F - 2)
        assert 0.0 < r.value <= 1.0


def _yaw_quat(phi):
    return np.array([np.cos(phi / 2), 0.0, 0.0, np.sin(phi / 2)])


def test_dm_state_layout_and_heading_invariance():
    m = common.model()
    assert m.npart == 15 and list(m.part_geom[:15]) == [1, 2, 3, 10, 11, 12, 4, 5, 6, 13, 14, 15, 7, 8, 9]
    mcs, keep = make_mocap_struct(load_motions(["spinkick"]))
    cfg = default_config(obs_mode=1)
    L = po.lib()
    e = _env(m, cfg, mcs, seed=2)
    rng = np.random.default_rng(4)
    q, v = common.standing_states(rng, 1, noise=0.3, vel=1.0)
    q, v = q[0].copy(), v[0].copy()
    np.ctypeslib.as_array(e.d.qpos)[: m.nq] = q
    np.ctypeslib.as_array(e.d.qvel)[: m.nv] = v
    o1 = np.zeros(256)
    n = L.dmo_env_obs_dm(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(o1))
    assert n == 197
    assert abs(o1[0] - e.idx_curr / mcs.clip_len[0]) < 1e-12 and abs(o1[1] - q[2]) < 1e-12
    quats = o1[2:2 + 105].reshape(15, 7)[:, 3:]
    assert np.abs(np.linalg.norm(quats, axis=1) - 1).max() < 1e-12 and (quats[:, 0] >= 0).all()
    pos = o1[2:2 + 105].reshape(15, 7)[:, :3]
    assert np.abs(pos[0] - [0, 0, 0.07]).max() < 0.05          # root part = sphere 7 cm above the root origin
    # rotate the whole character about z and translate it in xy: the heading-frame state must not change
    phi = 1.234
    q2, v2 = q.copy(), v.copy()
    c, s = np.cos(phi), np.sin(phi)
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    q2[:3] = Rz @ q[:3] + [3.0, -2.0, 0.0]
    q2[3:7] = mocap.qmul(_yaw_quat(phi), q[3:7])
    v2[:3] = Rz @ v[:3]           # world-frame linear velocity rotates; body-frame angular velocity does not
    np.ctypeslib.as_array(e.d.qpos)[: m.nq] = q2
    np.ctypeslib.as_array(e.d.qvel)[: m.nv] = v2
    o2 = np.zeros(256)
    L.dmo_env_obs_dm(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(o2))
    assert np.abs(o1[:197] - o2[:197]).max() < 1e-9
    # velocities: finite-difference check of the part-centre linear velocity through a tiny position step
    h = 1e-6
    d = po.DmoData()
    d.arr("qpos")[: m.nq] = q
    L.dmo_kinematics(C.byref(m), C.byref(d))
    g0 = d.arr("geom_xpos")[:16].copy()
    qn = q.copy()
    qn[:3] += h * v[:3]
    qn[7:] += h * v[6:]
    w = v[3:6]                      # body-frame angular velocity: q <- q * exp(h w / 2)
    dq = np.concatenate([[1.0], 0.5 * h * w])
    qn[3:7] = mocap.qmul(q[3:7], dq); qn[3:7] /= np.linalg.norm(qn[3:7])
    d.arr("qpos")[: m.nq] = qn
    L.dmo_kinematics(C.byref(m), C.byref(d))
    vfd = (d.arr("geom_xpos")[:16] - g0) / h
    R = np.array(d.arr("xmat")[1]).reshape(3, 3)
    hd = np.arctan2(R[1, 0], R[0, 0]); ch, sh = np.cos(hd), np.sin(hd)
    vel = o1[2 + 105:197].reshape(15, 6)
    for p in range(15):
        vw = vfd[m.part_geom[p]]
        exp = np.array([ch * vw[0] + sh * vw[1], -sh * vw[0] + ch * vw[1], vw[2]])
        assert np.abs(vel[p, :3] - exp).max() < 1e-4 * max(1.0, np.abs(exp).max())


def test_env_step_emits_dm_state_after_reset():
    m = common.model()
    mcs, keep = make_mocap_struct(load_motions(["walk"]))
    cfg = default_config(obs_mode=1, auto_reset=1, phase_mode=1)
    L = po.lib()
    e = _env(m, cfg, mcs, seed=5)
    obs, r = np.zeros(256), C.c_double()
    rng = np.random.default_rng(0)
    seen_reset = False
    for t in range(80):
        a = rng.uniform(-0.5, 0.5, 28)
        done = L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(a), po.dptr(obs), C.byref(r))
        assert np.isfinite(obs[:197]).all() and 0.0 <= obs[0] < 1.0
        if done:
            seen_reset = True
            assert e.ep_len == 0 and abs(obs[0] - (e.idx_init % 38) / 38.0) < 1e-12   # post-reset phase
            assert abs(obs[1] - e.d.qpos[2]) < 1e-12
    assert seen_reset
