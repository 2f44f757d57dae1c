"""Pins for the CPU float64 oracle (PARITY UNPINNED against MuJoCo itself -- these are the
independent checks we can make): inertia vs an independent numpy CRB, exact solves, energy and
momentum balances, contact sanity, PGS invariants, Philox known answer -- and the time-to-fall distribution of the
reference's own MuJoCo-produced episode log.  CPU only."""
import ctypes as C

import numpy as np

import common
import oracle.pyoracle as po
from deepmimic_mujoco_b200 import mjcf


def make():
    return po.Oracle(common.model())


def energy(o, mt):
    M = o.full_M()
    v = o.qvel
    pe = 9.81 * sum(mt.body_mass[b] * o.d.arr("xipos")[b][2] for b in range(mt.nbody))
    return 0.5 * v @ M @ v + pe


def test_mass_matrix_matches_independent_crb():
    mt, o = common.tables(), make()
    rng = np.random.default_rng(0)
    q, v = common.airborne_states(rng, 5)
    for i in range(5):
        o.set_state(q[i], v[i]); o.forward()
        Mn = mjcf.np_mass_matrix(mt, q[i])
        assert np.abs(o.full_M() - Mn).max() < 1e-12
        x = rng.normal(size=mt.nv)
        assert np.abs(Mn @ o.solve_M(x) - x).max() < 1e-11


def test_energy_rate_equals_nonconservative_power():
    """d/dt (KE + PE) = qvel . (passive + actuator + constraint): validates M, the RNE bias
    (Coriolis + gravity) and the integrator together."""
    mt = common.tables()
    rng = np.random.default_rng(1)
    m = common.model(); m.timestep = 1e-4
    o = po.Oracle(m)
    q, v = common.airborne_states(rng, 4, frac=0.3)
    for i in range(4):
        o.set_state(q[i], v[i], ctrl=rng.uniform(-0.5, 0.5, mt.nu)); o.forward()
        e0 = energy(o, mt)
        pw = lambda: o.qvel @ (o.d.arr("qfrc_passive")[:mt.nv] + o.d.arr("qfrc_actuator")[:mt.nv] + o.d.arr("qfrc_constraint")[:mt.nv])
        p0 = pw()
        o.step(); o.forward()
        e1, p1 = energy(o, mt), pw()
        assert abs((e1 - e0) / 1e-4 - 0.5 * (p0 + p1)) < 2e-3 * max(1.0, abs(p0))


def test_free_flight_momentum():
    """No contacts: CoM accelerates at exactly g regardless of internal torques/damping."""
    mt, o = common.tables(), make()
    rng = np.random.default_rng(2)
    q, v = common.airborne_states(rng, 3, frac=0.3)
    for i in range(3):
        o.set_state(q[i], v[i], ctrl=rng.uniform(-0.5, 0.5, mt.nu)); o.forward()
        if o.d.ncon:
            continue
        # CoM acceleration = (1/M) sum_b m_b J_b qacc + velocity-product terms; check via finite differences
        c0 = o.d.arr("com").copy()
        m2 = common.model(); m2.timestep = 1e-3
        o2 = po.Oracle(m2); o2.set_state(q[i], v[i], ctrl=o.d.arr("ctrl")[:mt.nu].copy())
        coms = []
        for _ in range(3):
            o2.forward(); coms.append(o2.d.arr("com").copy()); o2.step()
        acc = (coms[2] - 2 * coms[1] + coms[0]) / 1e-6
        if o2.d.ncon == 0:
            assert np.abs(acc - [0, 0, -9.81]).max() < 5e-2


def _momenta(o, mt):
    """Whole-body linear momentum and angular momentum about the CoM from the oracle's body data."""
    d = o.d
    com = d.arr("com").copy()
    P, L = np.zeros(3), np.zeros(3)
    for b in range(1, mt.nbody):
        R = np.array(d.arr("xmat")[b]).reshape(3, 3)
        I = mt.body_inertia[b]
        Ib = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]])
        w, vl = np.array(d.arr("cvel")[b][:3]), np.array(d.arr("cvel")[b][3:])
        r = np.array(d.arr("xipos")[b]) - com
        vb = vl + np.cross(w, r)
        P += mt.body_mass[b] * vb
        L += R @ Ib @ R.T @ w + mt.body_mass[b] * np.cross(r, vb)
    return P, L


def test_free_flight_conserves_angular_momentum():
    """Airborne, arbitrary motor torques, joint damping, joint limits and self-contacts are all internal: the
    angular momentum about the CoM is conserved and the linear momentum changes by m g t, up to the integration
    error, which must shrink ~4x when the step is halved (mj_integratePos advances the free-joint quaternion with
    the combined angular velocity: 2nd order, see test_rk4_convergence_order).  Pins inertia tensors,
    Coriolis/centrifugal terms and the quaternion integrator independently of MuJoCo."""
    mt = common.tables()
    rng = np.random.default_rng(5)
    q, v = common.airborne_states(rng, 4, frac=0.4, vel=3.0)
    ctrl = rng.uniform(-0.5, 0.5, (4, mt.nu))
    T = 0.025

    def drift(i, dt):
        m = common.model(); m.timestep = dt
        o = po.Oracle(m)
        o.set_state(q[i], v[i], ctrl=ctrl[i]); o.forward()
        P0, L0 = _momenta(o, mt)
        for _ in range(int(round(T / dt))):
            o.step()
        o.forward()
        assert all(c.geom1 != 0 for c in o.d.contact[: o.d.ncon])       # never touched the floor
        P1, L1 = _momenta(o, mt)
        eL = np.abs(L1 - L0).max() / max(1.0, np.abs(L0).max())
        eP = np.abs(P1 - P0 - mt.body_mass.sum() * np.array([0, 0, -9.81]) * T).max() / max(1.0, np.abs(P0).max())
        return eL, eP

    for i in range(4):
        eL1, eP1 = drift(i, 1e-3)
        eL2, eP2 = drift(i, 5e-4)
        assert eL1 < 2e-6 and eP1 < 2e-7, (i, eL1, eP1)
        assert eL2 < eL1 / 3 + 1e-10 and eP2 < eP1 / 3 + 1e-11, (i, eL1, eL2, eP1, eP2)


def test_standing_contacts_support_weight():
    mt, o = common.tables(), make()
    qpos = mt.qpos0.copy(); qpos[2] -= 0.0225  # soles 2.5 mm into the floor
    o.set_state(qpos, np.zeros(mt.nv)); o.forward()
    d = o.d
    assert d.ncon == 8 and d.nefc == 32          # 2 feet x 4 corners x 4 pyramid edges
    f = d.arr("efc_force")[:32]
    assert np.all(f >= 0)
    # total normal force (sum of all pyramid edge forces, normal component 1 each) is of the order of the weight
    assert 0.2 * 45 * 9.81 < f.sum() < 5 * 45 * 9.81
    for c in d.contact[:8]:
        assert c.geom1 == 0 and c.dim == 3 and abs(c.frame[2] - 1) < 1e-12 and c.dist < 0
    # floor frame: t1 = +y, t2 = -x (mju_makeFrame with n = +z)
    assert np.allclose(np.array(d.contact[0].frame), [0, 0, 1, 0, 1, 0, -1, 0, 0])


def test_pgs_invariants_and_warmstart_fixed_point():
    mt, o = common.tables(), make()
    rng = np.random.default_rng(3)
    q, v, w = common.rollout_states(rng, 6)
    for i in range(6):
        o.set_state(q[i], v[i], warm=w[i]); o.forward()
        d = o.d
        n = d.nefc
        if n == 0:
            continue
        f = d.arr("efc_force")[:n]
        assert np.all(f >= 0)
        AR = np.array([[d.efc_AR[r][c] for c in range(n)] for r in range(n)])
        assert np.allclose(AR, AR.T) and np.all(np.linalg.eigvalsh(AR) > 0)
        # qacc = qacc_smooth + M^-1 J' f
        J = d.arr("efc_J")[:n, :mt.nv]
        qa = d.arr("qacc_smooth")[:mt.nv] + np.linalg.solve(o.full_M(), J.T @ f)
        assert np.abs(qa - d.arr("qacc")[:mt.nv]).max() < 1e-8 * max(1, np.abs(qa).max())
        # A = J M^-1 J' + R
        A = J @ np.linalg.solve(o.full_M(), J.T) + np.diag(d.arr("efc_R")[:n])
        assert np.abs(A - AR).max() < 1e-9 * max(1.0, np.abs(A).max())


def test_limit_rows_and_order():
    mt, o = common.tables(), make()
    qpos = mt.qpos0.copy(); qpos[2] = 3.0
    qpos[7 + 9] = -0.2     # right_elbow below its lower limit 0
    qpos[7 + 17] = 0.3     # right_knee above its upper limit 0
    o.set_state(qpos, np.zeros(mt.nv)); o.forward()
    d = o.d
    assert d.nefc >= 2 and d.efc_type[0] == 0 and d.efc_type[1] == 0
    J = d.arr("efc_J")
    assert J[0, 6 + 9] == 1.0 and abs(d.efc_pos[0] + 0.2) < 1e-12      # lower side: J = +1, pos = q - lo
    assert J[1, 6 + 17] == -1.0 and abs(d.efc_pos[1] + 0.3) < 1e-12    # upper side: J = -1, pos = hi - q
    assert d.efc_force[0] > 0 and d.efc_force[1] > 0


def test_rk4_convergence_order():
    """Halving h on a contact-free trajectory shrinks the global error at least ~4x.  (The hinge
    coordinates are 4th order; the free-joint quaternion is advanced as q0 * exp(h * sum b_i w_i),
    MuJoCo's mj_integratePos, which is 2nd order when the angular velocity direction changes.)"""
    mt = common.tables()
    rng = np.random.default_rng(5)
    q, v = common.airborne_states(rng, 1, frac=0.2, vel=1.0)
    def run(h, n):
        m = common.model(); m.timestep = h
        o = po.Oracle(m); o.set_state(q[0], v[0])
        for _ in range(n):
            o.step()
        assert o.d.ncon == 0
        return o.qvel.copy()
    ref = run(0.0005, 256)
    e1 = np.abs(run(0.032, 4) - ref).max()
    e2 = np.abs(run(0.016, 8) - ref).max()
    assert e2 < 1e-7 and e1 / e2 > 3.0, (e1, e2)


def test_philox_known_answer():
    L = po.lib()
    out = (C.c_uint32 * 4)()
    L.dmo_philox(0, 0, 0, 0, out)   # Random123 kat_vectors: philox4x32-10, zero counter and key
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


def test_fall_time_distribution_matches_reference_monitor_log():
    """The one MuJoCo-PRODUCED artefact in the reference: the monitor rows of its training run.  Under the reference's
    protocol (trpo.py:27-80: standing pose +- 0.01, dp_env_v3.py:158-164; N(0,1) actions of the initial Gaussian
    policy, clamped to the ctrlrange; done when the CoM height leaves [0.7, 2.0]) the time to fall is a statistic of
    the whole pipeline -- inertia, gravity, actuator gears and clamp, joint limits, damping, foot contacts, RK4, the
    termination rule.  The oracle's distribution must be the reference's: mean within 3 steps (8 %; measured -0.8 with
    2000 episodes, reference sample error 0.8), spread within 25 %, two-sample Kolmogorov-Smirnov not rejecting at 1 %
    (measured D = 0.08, p = 0.6)."""
    from scipy import stats
    mt, o = common.tables(), make()
    ref = common.ref_fall_lengths(100)
    rng = np.random.default_rng(7)
    lens = []
    for ep in range(400):
        o.set_state(mt.qpos0 + rng.uniform(-0.01, 0.01, mt.nq), rng.uniform(-0.01, 0.01, mt.nv))
        for t in range(400):
            o.d.arr("ctrl")[:mt.nu] = rng.normal(size=mt.nu)
            o.step()
            z = o.d.arr("com")[2]
            if z < 0.7 or z > 2.0:
                break
        lens.append(t + 1)
    lens = np.asarray(lens, dtype=np.float64)
    assert abs(lens.mean() - ref.mean()) < 3.0, (lens.mean(), ref.mean())
    assert 0.75 < lens.std() / ref.std() < 1.25, (lens.std(), ref.std())
    ks = stats.ks_2samp(lens, ref)
    assert ks.pvalue > 0.01, ks
    # and the quartiles individually (reference: 29.75 / 33.5 / 39)
    assert np.abs(np.percentile(lens, [25, 50, 75]) - np.percentile(ref, [25, 50, 75])).max() <= 3.0



def test_trained_policy_survival_matches_reference_monitor_log():
    """The reference ships the policy its TRPO run trained IN MuJoCo 2.0 (checkpoint_tmp/.../trpo-walk-0, read without
    TensorFlow by tests/golden/make_ref_policy_golden.py; its logstd reproduces the entropy logged by update 1899 to
    1e-5, which dates the file) and the monitor rows of the episodes that policy played.  A random policy falls after
    ~35 steps; this one keeps the MuJoCo humanoid up for ~290 (mean of the 100 episodes around the save, median 246,
    heavy tail).  The same weights under the same protocol (trpo.py:27-80: reset_model_init, stochastic actions,
    done on CoM height) must keep the ORACLE's humanoid up for as long: mean, median and lower quartile within 20 %,
    Kolmogorov-Smirnov not rejecting at 1 % (measured with 4500 episodes over 3 seeds: mean 278 = -4 %, median 238 =
    -3 %, KS p 0.4-0.8 on 200-episode samples).  tools/trained_policy_sensitivity.py shows what this resolves that
    the random-policy test cannot: actuator gear x 0.8 / x 1.25 (+32 % / -36 %), timestep x 0.8 / x 1.25."""
    mt, o = common.tables(), make()
    pol = common.RefTrainedPolicy()
    ref = pol.monitor_window(50)
    assert len(ref) == 100 and 250 < ref.mean() < 330
    rng = np.random.default_rng(5)
    lens = []
    for ep in range(250):
        o.set_state(mt.qpos0 + rng.uniform(-0.01, 0.01, mt.nq), rng.uniform(-0.01, 0.01, mt.nv))
        for t in range(3000):
            ob = np.concatenate([o.qpos[7:], o.qvel[6:]])               # _get_obs, dp_env_v3.py:62-65
            o.d.arr("ctrl")[:mt.nu] = pol.mean_action(ob[None])[0] + pol.act_std * rng.normal(size=mt.nu)
            o.step()
            z = o.d.arr("com")[2]
            if z < 0.7 or z > 2.0:
                break
        lens.append(t + 1)
    lens = np.asarray(lens, dtype=np.float64)
    assert lens.mean() > 5 * common.ref_fall_lengths(100).mean()        # nothing like the random policy's 35 steps
    assert common.trained_policy_verdict(lens, ref), (lens.mean(), np.percentile(lens, [25, 50, 75]), ref.mean(),
                                                      np.percentile(ref, [25, 50, 75]))
    import torch                                                         # the torch form the GPU test uses
    x = rng.normal(size=(64, 56)) * 0.5
    assert np.abs(pol.torch_mean_action(torch.as_tensor(x, dtype=torch.float32)).numpy() - pol.mean_action(x)).max() < 1e-4
    wide = pol.monitor_window(100)                                       # and against the wider window of the log
    assert abs(lens.mean() / wide.mean() - 1.0) < 0.2 and abs(np.median(lens) / np.median(wide) - 1.0) < 0.2


def test_ref_aux_matches_independent_numpy():
    """Reference-pose features (end effectors in the heading frame, CoM velocity): the oracle's C routine
    vs the product's host numpy implementation (independent code paths)."""
    from deepmimic_mujoco_b200.refaux import compute_ref_aux
    mt = common.tables()
    aux = compute_ref_aux(["walk"])
    c = common.clip("walk")
    L = po.lib(); m = common.model()
    out = np.zeros(24)
    for f in (0, 7, 20, 38):
        q = np.ascontiguousarray(c.data_config[f]); v = np.ascontiguousarray(np.nan_to_num(c.data_vel[f]))
        L.dmo_ref_aux(C.byref(m), po.dptr(q), po.dptr(v), po.dptr(out))
        assert np.abs(out - aux[f]).max() < 1e-10


def test_env_reset_modes_and_reward_bounds():
    from deepmimic_mujoco_b200.model_blob import default_config
    from deepmimic_mujoco_b200.refaux import compute_ref_aux
    from deepmimic_mujoco_b200.sim import load_motions, make_mocap_struct
    mt = common.tables()
    L = po.lib(); m = common.model()
    cfg = default_config(reward_mode=4, auto_reset=1)
    aux = compute_ref_aux(["walk"])
    mcs, keep = make_mocap_struct(load_motions(["walk"]), aux)
    e = po.DmoEnv()
    L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 3, 0, 0)
    L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 0)
    assert 0 <= e.idx_init < 39 and e.idx_curr == e.idx_init
    q = np.ctypeslib.as_array(e.d.qpos)[:35]
    assert np.abs(q - np.float32(common.clip("walk").data_config[e.idx_init])).max() == 0
    obs = np.zeros(56); rew = C.c_double()
    rng = np.random.default_rng(0)
    # a step from the exact reference pose with zero action scores a high imitation reward
    a = np.zeros(28)
    L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(a), po.dptr(obs), C.byref(rew))
    assert 0.3 < rew.value <= 1.0
    ndone = 0
    for t in range(300):
        a = rng.uniform(-0.5, 0.5, 28)
        ndone += L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(a), po.dptr(obs), C.byref(rew))
        assert 0.0 <= rew.value <= 1.0 and np.all(np.isfinite(obs))
    assert ndone >= 3
    e2 = po.DmoEnv()
    L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e2), 3, 0, 0)
    L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e2), 1)
    q2 = np.ctypeslib.as_array(e2.d.qpos)[:35]
    assert 1e-4 < np.abs(q2 - mt.qpos0).max() <= 0.01 + 1e-7
