"""CUDA path vs the float64 oracle, through the C-ABI (run on the B200 box: pytest -m gpu).

Tolerances: single step from identical fp32-representable inputs,
|dqpos| <= 1e-4, |dqvel| / max(1,|qvel|) <= 1e-4, reward <= 1e-5 abs,
FK quantities <= 1e-5, `done` exact away from the threshold.  Measured on B200 (profiles/r2_parity_measured.json):
|dqpos| <= 1.4e-7, rel |dqvel| <= 4.4e-6, reward <= 5e-7 on every case of this file.  PARITY UNPINNED against MuJoCo
itself (no binary, no golden vectors) -- the oracle is the restatement being matched.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import common

pytestmark = pytest.mark.gpu

N = 64


@pytest.fixture(scope="module")
def ctx():
    import oracle.pyoracle as po
    from deepmimic_mujoco_b200.sim import BatchedSim
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sim = BatchedSim(N, motions=("walk",), seed=11)
    return sim, po.Oracle(common.model()), common.tables(), po


def oracle_forward(o, mt, q, v, c, w):
    o.set_state(q, v, c, w)
    o.forward()
    return o.d


# Envs whose contact normal is ill-conditioned in fp32 (deep sphere/capsule-vs-foot-box penetration: the normal is
# (c - clamp(c)) / |c - clamp(c)| with |.| ~ 1e-4): rows, forces and qacc are compared with the bound scaled by this factor.
# Measured worst case (profiles/r2_parity_measured.json): 2.5 x the normal bound on aref / b, 2 x on qacc.
LOOSE_SCALE = 5.0


def compare_forward(ctx, q, v, ctrl, warm=None, tol_scale=1.0, label="forward"):
    sim, o, mt, po = ctx
    n = q.shape[0]
    warm = np.zeros((n, mt.nv)) if warm is None else warm
    sim.set_state(q, v, warm)
    g = sim.forward_debug(torch.tensor(ctrl, dtype=torch.float32))
    for i in range(n):
        d = oracle_forward(o, mt, q[i], v[i], ctrl[i], warm[i])
        assert d.ncon == g["ncon"][i] and d.nefc == g["nefc"][i], (i, d.ncon, g["ncon"][i], d.nefc, g["nefc"][i])
        nb, nv, ne = mt.nbody, mt.nv, d.nefc
        env_scale = 1.0
        assert np.abs(d.arr("xpos")[:nb] - g["xpos"][i]).max() < 1e-5
        assert np.abs(d.arr("xquat")[:nb] - g["xquat"][i]).max() < 1e-5
        assert np.abs(d.arr("xipos")[:nb] - g["xipos"][i]).max() < 1e-5
        assert np.abs(d.arr("com") - g["com"][i]).max() < 1e-5
        qM = d.arr("qM")[:mt.nM]
        assert np.abs(qM - g["qM"][i]).max() < 1e-5 * max(1.0, np.abs(qM).max())
        bias = d.arr("qfrc_bias")[:nv]
        assert np.abs(bias - g["qfrc_bias"][i]).max() < 2e-5 * max(1.0, np.abs(bias).max())
        qs = d.arr("qacc_smooth")[:nv]
        assert np.abs(qs - g["qacc_smooth"][i]).max() < 5e-5 * max(1.0, np.abs(qs).max())
        if d.ncon:
            oc = np.array([[c.dist, *c.pos, *c.frame, c.geom1, c.geom2, c.dim] for c in d.contact[:d.ncon]])
            gc = g["contact"][i][: d.ncon]
            # dist, pos, normal, geom ids, dim always; tangents only where they matter (condim 3)
            cols = [0, 1, 2, 3, 4, 5, 6, 13, 14, 15]
            # a sphere centre that sits (almost) on a box face makes the normal = (c - clamp(c)) / |.|
            # ill-conditioned in fp32: deep capsule/sphere-vs-foot-box penetrations get a loose bound
            boxy = np.isin(oc[:, 13], (12, 15)) | np.isin(oc[:, 14], (12, 15))
            loose = boxy & (oc[:, 15] == 1)
            assert np.abs(oc[:, [0, 13, 14, 15]] - gc[:, [0, 13, 14, 15]]).max() < 2e-5
            if (~loose).any():
                assert np.abs(oc[~loose][:, cols] - gc[~loose][:, cols]).max() < 2e-5
            if loose.any():
                # measured: <= 2e-3 on the normal (profiles/r2_parity_measured.json); the rows / forces / qacc that
                # inherit this normal are still compared below, with the bound scaled by the same factor
                assert np.abs(oc[loose][:, cols] - gc[loose][:, cols]).max() < 5e-3
                common.record(label, "loose_contact_normal", np.abs(oc[loose][:, cols] - gc[loose][:, cols]).max())
                env_scale = LOOSE_SCALE
            fr = oc[:, 15] > 1
            if fr.any():
                assert np.abs(oc[fr, 7:13] - gc[fr, 7:13]).max() < 1e-4
        if ne:
            # aref/b carry k * (pos - margin) with k ~ 1e3: fp32 position round-off (~3e-7 at |x| ~ 3 m)
            # shows up as ~5e-4 absolute
            for k, rel, ab in (("efc_pos", 1e-5, 0.0), ("efc_R", 1e-4, 0.0), ("efc_aref", 1e-4, 1e-3), ("efc_b", 1e-4, 1e-3)):
                a = d.arr(k)[:ne]
                err = np.abs(a - g[k][i][:ne]).max()
                common.record(label + ("/loose" if env_scale > 1 else ""), k + "_over_bound", err / (ab + rel * max(1.0, np.abs(a).max())))
                assert err < (ab + rel * max(1.0, np.abs(a).max())) * tol_scale * env_scale, (k, i)
            f = d.arr("efc_force")[:ne]
            err = np.abs(f - g["efc_force"][i][:ne]).max() / max(1.0, np.abs(f).max())
            common.record(label + ("/loose" if env_scale > 1 else ""), "efc_force_rel", err)
            assert err < 1e-3 * tol_scale * env_scale, i
        qa = d.arr("qacc")[:nv]
        err = np.abs(qa - g["qacc"][i]).max() / max(1.0, np.abs(qa).max())
        common.record(label + ("/loose" if env_scale > 1 else ""), "qacc_rel", err)
        assert err < 2e-4 * tol_scale * env_scale, i


def compare_step(ctx, q, v, ctrl, warm=None, vtol=1e-4, label="step"):
    sim, o, mt, po = ctx
    n = q.shape[0]
    warm = np.zeros((n, mt.nv)) if warm is None else warm
    sim.set_state(q, v, warm)
    act = torch.tensor(ctrl, dtype=torch.float32, device=sim.device)
    obs, rew, done = sim.step(act)
    gq, gv, gw = sim.get_state()
    obs = obs.double().cpu().numpy(); done = done.cpu().numpy(); rew = rew.cpu().numpy()
    for i in range(n):
        o.set_state(q[i], v[i], ctrl[i], warm[i])
        o.step()
        common.record(label, "qpos_abs", np.abs(o.qpos - gq[i]).max())
        common.record(label, "qvel_rel", (np.abs(o.qvel - gv[i]) / np.maximum(1.0, np.abs(o.qvel))).max())
        assert np.abs(o.qpos - gq[i]).max() < 1e-4
        assert (np.abs(o.qvel - gv[i]) / np.maximum(1.0, np.abs(o.qvel))).max() < vtol
        zc = o.d.arr("com")[2]
        if min(abs(zc - 0.7), abs(zc - 2.0)) > 1e-4:
            assert bool(done[i]) == bool(zc < 0.7 or zc > 2.0)
        assert np.abs(obs[i] - np.concatenate([gq[i][7:], gv[i][6:]])).max() == 0.0
        assert rew[i] == 1.0


def test_forward_airborne(ctx):
    rng = np.random.default_rng(0)
    q, v = common.airborne_states(rng, N)
    compare_forward(ctx, q, v, common.f32(rng.uniform(-0.6, 0.6, (N, 28))), label="forward_airborne")


def test_forward_standing_contacts(ctx):
    rng = np.random.default_rng(1)
    q, v = common.standing_states(rng, N)
    compare_forward(ctx, q, v, common.f32(rng.uniform(-0.6, 0.6, (N, 28))), label="forward_standing")


def test_forward_rollout_states_with_warmstart(ctx):
    rng = np.random.default_rng(2)
    q, v, w = common.rollout_states(rng, N)
    compare_forward(ctx, q, v, common.f32(rng.uniform(-0.6, 0.6, (N, 28))), w, label="forward_rollout_warm")


@pytest.mark.parametrize("clip", ["walk", "spinkick", "dance_b"])
def test_step_from_mocap_frames(ctx, clip):
    rng = np.random.default_rng(3)
    c = common.clip(clip)
    idx = rng.integers(0, len(c), size=N)
    q, v = common.mocap_states(clip, idx)
    compare_step(ctx, q, v, common.f32(rng.uniform(-0.5, 0.5, (N, 28))), label="step_mocap_" + clip)


def test_step_standing_and_rollout(ctx):
    rng = np.random.default_rng(4)
    q, v = common.standing_states(rng, N)
    compare_step(ctx, q, v, common.f32(rng.uniform(-0.7, 0.7, (N, 28))), label="step_standing")
    q, v, w = common.rollout_states(rng, N)
    compare_step(ctx, q, v, common.f32(rng.normal(size=(N, 28))), w, label="step_rollout")


def test_step_airborne_self_contacts(ctx):
    rng = np.random.default_rng(5)
    q, v = common.airborne_states(rng, N, frac=0.35)
    compare_step(ctx, q, v, common.f32(rng.uniform(-0.5, 0.5, (N, 28))), label="step_airborne_selfcontact")


def _env_pair(po, reward_mode, ctrl_mode, motions=("walk",), n=32, seed=5, auto_reset=1, reset_mode=0, term_mode=0,
              phase_mode=0, obs_mode=0):
    from deepmimic_mujoco_b200.model_blob import default_config
    from deepmimic_mujoco_b200.refaux import compute_ref_aux
    from deepmimic_mujoco_b200.sim import BatchedSim, load_motions, make_mocap_struct
    cfg = default_config(reward_mode=reward_mode, ctrl_mode=ctrl_mode, auto_reset=auto_reset, reset_mode=reset_mode,
                         term_mode=term_mode, phase_mode=phase_mode, obs_mode=obs_mode)
    aux = compute_ref_aux(motions) if reward_mode == 4 else None
    clip_ids = torch.arange(n, dtype=torch.int32) % len(motions)
    sim = BatchedSim(n, motions=motions, seed=seed, config=cfg, ref_aux=aux, clip_ids=clip_ids)
    mcs, keep = make_mocap_struct(load_motions(list(motions)), aux)
    return sim, cfg, mcs, keep, clip_ids


@pytest.mark.parametrize("reward_mode,ctrl_mode,term_mode,phase_mode,obs_mode",
                         [(0, 0, 0, 0, 0), (1, 0, 0, 0, 0), (2, 0, 0, 0, 0), (3, 0, 0, 0, 0), (4, 0, 0, 0, 0), (4, 1, 0, 0, 0),
                          (1, 2, 0, 0, 0), (4, 0, 1, 0, 0),
                          # time-based phase with lerp/slerp interpolation; DeepMimic 197-d state
                          (4, 0, 0, 1, 0), (1, 0, 0, 1, 0), (2, 0, 0, 1, 0), (3, 0, 1, 1, 1), (4, 0, 0, 0, 1), (0, 0, 1, 1, 1)])
def test_env_step_rewards_pd_and_reset(ctx, reward_mode, ctrl_mode, term_mode, phase_mode, obs_mode):
    """Full env step (PD -> RK4 -> reward -> done -> auto reset) for a few consecutive steps vs the
    oracle env; RSI frame indices must be bit-identical (same Philox stream)."""
    _, _, mt, po = ctx
    n = 32
    motions = ("walk", "dance_b", "spinkick")
    sim, cfg, mcs, keep, clip_ids = _env_pair(po, reward_mode, ctrl_mode, motions, n, term_mode=term_mode,
                                              phase_mode=phase_mode, obs_mode=obs_mode)
    odim = 197 if obs_mode == 1 else 56
    assert sim.obs_dim == odim and tuple(sim.obs.shape) == (n, odim) and tuple(sim.rec.shape) == (n, odim + 2)
    L = po.lib()
    m = common.model()
    envs = [po.DmoEnv() for _ in range(n)]
    for i, e in enumerate(envs):
        L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 5, i, int(clip_ids[i]))
        L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 0)
    obs_r = sim.reset().double().cpu().numpy()
    gq, gv, _ = sim.get_state()
    obs_o = np.zeros(256); rew_o = C.c_double()
    for i, e in enumerate(envs):
        assert e.idx_init == int(sim.idx_init[i]) and e.idx_curr == int(sim.idx_curr[i])
        assert np.abs(np.ctypeslib.as_array(e.d.qpos)[: mt.nq] - gq[i]).max() == 0.0
        if obs_mode == 1:   # post-reset observation, and dmb_get_obs of the stored state
            L.dmo_env_obs_dm(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(obs_o))
            assert np.abs(obs_o[:odim] - obs_r[i]).max() < 5e-5 * max(1.0, np.abs(obs_o[:odim]).max()), i
    if obs_mode == 1:
        assert np.abs(sim.get_obs().double().cpu().numpy() - obs_r).max() < 1e-6
    rng = np.random.default_rng(9)
    ndone = 0
    for t in range(6 if term_mode == 0 else 14):
        scale = 1.0 if ctrl_mode == 0 else 1.0
        act = common.f32(rng.uniform(-0.5, 0.5, (n, 28)) * scale)
        # identical inputs for both: copy the GPU state (fp32) into the oracle envs each step
        gq, gv, gw = sim.get_state()
        for i, e in enumerate(envs):
            np.ctypeslib.as_array(e.d.qpos)[: mt.nq] = gq[i]
            np.ctypeslib.as_array(e.d.qvel)[: mt.nv] = gv[i]
            np.ctypeslib.as_array(e.d.qacc_warmstart)[: mt.nv] = gw[i]
        obs, rew, done = sim.step(torch.tensor(act, dtype=torch.float32, device=sim.device))
        obs = obs.double().cpu().numpy(); rew = rew.double().cpu().numpy(); done = done.cpu().numpy()
        for i, e in enumerate(envs):
            a = np.ascontiguousarray(act[i])
            od = L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(a), po.dptr(obs_o), C.byref(rew_o))
            zc = e.d.com[2] if not od else None
            common.record(f"env_step_r{reward_mode}", "reward_abs", abs(rew_o.value - rew[i]))
            assert abs(rew_o.value - rew[i]) < 1e-5, (t, i, rew_o.value, rew[i])
            assert bool(od) == bool(done[i]), (t, i)
            ndone += int(od)
            assert e.idx_curr == int(sim.idx_curr[i]) and e.idx_init == int(sim.idx_init[i])
            assert np.abs(obs_o[:odim] - obs[i]).max() < 2e-4 * max(1.0, np.abs(obs_o[:odim]).max()), (t, i)
            if obs_mode == 1:
                assert abs(obs_o[0] - obs[i][0]) < 1e-6   # phase
        rec = sim.rec.double().cpu().numpy()
        assert np.array_equal(rec[:, :odim], obs) and np.array_equal(rec[:, odim], rew) and np.array_equal(rec[:, odim + 1], done.astype(float))
    if term_mode == 1:
        assert ndone > 0   # spinkick / dance frames put hands or knees on the floor quickly
    sim.close()


def test_reset_modes_and_determinism(ctx):
    _, _, mt, po = ctx
    sim, cfg, mcs, keep, _ = _env_pair(po, 0, 0, ("walk",), 32, seed=77, reset_mode=1)
    o1 = sim.reset().clone()
    q1, v1, _ = sim.get_state()
    assert np.abs(q1 - mt.qpos0).max() <= 0.01 + 1e-6 and np.abs(v1).max() <= 0.01 + 1e-6
    assert np.abs(q1 - mt.qpos0).max() > 1e-4
    L = po.lib(); m = common.model()
    e = po.DmoEnv()
    L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 77, 3, 0)
    L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 1)
    assert np.abs(np.ctypeslib.as_array(e.d.qpos)[: mt.nq] - q1[3]).max() < 1e-6
    assert np.abs(np.ctypeslib.as_array(e.d.qvel)[: mt.nv] - v1[3]).max() < 1e-6
    # masked reset only touches the masked envs; a fresh sim with the same seed reproduces the first reset
    mask = torch.zeros(32, dtype=torch.uint8); mask[5] = 1
    before = sim.qpos.clone()
    sim.reset(mask=mask.cuda())
    changed = (sim.qpos != before).any(dim=1).cpu().numpy()
    assert changed[5] and changed.sum() == 1
    sim2, *_ = _env_pair(po, 0, 0, ("walk",), 32, seed=77, reset_mode=1)
    assert torch.equal(sim2.reset(), o1)
    sim.close(); sim2.close()


def test_gym_surface_dpenv():
    from deepmimic_mujoco_b200.env import DPEnv
    env = DPEnv(motion="walk", seed=0)
    assert env.observation_space.shape == (56,) and env.action_space.shape == (28,)
    assert np.all(env.action_space.low == -0.5) and np.all(env.action_space.high == 0.5)
    env.seed(0)
    ob = env.reset()
    assert ob.shape == (56,) and ob.dtype == np.float64
    assert 0 <= env.idx_curr < 39
    assert np.allclose(ob[:28], np.float32(env.mocap.data_config[env.idx_init][7:]), atol=1e-6)
    ob2 = env.reset_model_init()
    assert np.abs(ob2[:28]).max() <= 0.0101
    n = 0
    for t in range(300):
        ob, rew, done, info = env.step(env.action_space.sample())
        n += 1
        assert rew == 1.0 and isinstance(done, bool) and info == {}
        if done:
            break
    assert done and 5 < n < 300   # an unactuated humanoid falls (progress.csv:2-4: ~35 steps)
    assert abs(env.dt - 0.0166 * 6) < 1e-12
    env.close()


def test_full_size_properties():
    """BASELINE config-2 size (4096 envs, walk): size-independent properties of a rollout."""
    from deepmimic_mujoco_b200.env import DPVecEnv
    env = DPVecEnv(4096, motions=("walk",), seed=3, reward_mode=4, auto_reset=True)
    obs = env.reset()
    assert obs.shape == (4096, 56) and torch.isfinite(obs).all()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    ndone = 0
    for t in range(40):
        act = torch.rand(4096, 28, device="cuda", generator=g) - 0.5
        obs, rew, done, info = env.step(act)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert (rew >= 0).all() and (rew <= 1.0 + 1e-6).all()
        qn = env.sim.qpos[:, 3:7].norm(dim=1)
        assert (qn - 1).abs().max() < 1e-4          # free-joint quaternion stays normalised
        assert (env.sim.flags & 4).sum() == 0       # no non-finite states
        ndone += int(done.sum())
        # auto-reset envs restart their episode counters and sit on a mocap frame
        d = done.bool()
        assert (env.sim.ep_len[d] == 0).all() and (env.sim.ep_len[~d] > 0).all()
    assert ndone > 0
    # determinism: same seed, same actions -> identical trajectories
    env2 = DPVecEnv(4096, motions=("walk",), seed=3, reward_mode=4, auto_reset=True)
    env2.reset()
    g.manual_seed(0)
    for t in range(40):
        act = torch.rand(4096, 28, device="cuda", generator=g) - 0.5
        o2, r2, d2, _ = env2.step(act)
    assert torch.equal(o2, obs) and torch.equal(r2, rew)
    env.close(); env2.close()


def test_config3_spinkick_16384_and_config5_mixed():
    """BASELINE configs[2] (16384-env spinkick, early termination + auto-reset) and configs[4]
    (mixed walk/dance_b/spinkick with per-env phase) at full per-GPU size: size-independent properties."""
    from deepmimic_mujoco_b200.dist import mixed_clip_ids
    from deepmimic_mujoco_b200.env import DPVecEnv
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    env = DPVecEnv(16384, motions=("spinkick",), seed=1, reward_mode=4, auto_reset=True)
    obs = env.reset()
    assert (env.sim.idx_init >= 0).all() and (env.sim.idx_init < 78).all()
    assert env.sim.idx_init.float().std() > 10          # RSI spreads over the clip
    tot_done = 0
    for t in range(12):
        obs, rew, done, info = env.step(torch.rand(16384, 28, device="cuda", generator=g) - 0.5)
        tot_done += int(done.sum())
        assert torch.isfinite(obs).all() and (rew >= 0).all() and (rew <= 1 + 1e-6).all()
    assert tot_done > 0 and int((env.sim.flags & 4).sum()) == 0
    env.close()
    clip_ids = mixed_clip_ids(0, 4096, 3)
    env = DPVecEnv(4096, motions=("walk", "dance_b", "spinkick"), seed=2, reward_mode=4, auto_reset=True, clip_ids=clip_ids)
    env.reset()
    lens = torch.tensor([39, 153, 78], device="cuda")[clip_ids.cuda().long()]
    for t in range(8):
        obs, rew, done, info = env.step(torch.rand(4096, 28, device="cuda", generator=g) - 0.5)
        assert (env.sim.idx_curr >= 0).all() and (env.sim.idx_curr < lens).all()   # phase stays inside each env's own clip
        assert torch.isfinite(obs).all()
    # phase advanced by one frame per step for envs that did not reset (dp_env_v3.py:101-102)
    alive = env.sim.ep_len == 8
    assert alive.any()
    assert ((env.sim.idx_init[alive] + 8) % lens[alive] == env.sim.idx_curr[alive]).all()
    env.close()


def test_capacity_overflow_matches_oracle(ctx):
    """A humanoid lying on the floor produces more contacts / rows than the per-env capacities
    (max_con 16, max_efc 40): the CUDA path must drop exactly the contacts the oracle drops and raise
    the same flags."""
    sim, o, mt, po = ctx
    rng = np.random.default_rng(21)
    q = np.tile(mt.qpos0, (N, 1)); v = np.zeros((N, mt.nv))
    for i in range(N):
        ang = np.pi / 2 + rng.uniform(-0.2, 0.2)
        q[i, 3:7] = [np.cos(ang / 2), 0, np.sin(ang / 2), 0]      # pitched onto its back / front
        q[i, 2] = rng.uniform(0.08, 0.14)
        q[i, 7:] = rng.uniform(-0.15, 0.15, mt.nq - 7)
    q, v = common.f32(q), common.f32(v)
    ctrl = np.zeros((N, mt.nu))
    sim.set_state(q, v)
    g = sim.forward_debug(torch.tensor(ctrl, dtype=torch.float32))
    flags = sim.flags.cpu().numpy()
    nover = 0
    for i in range(N):
        d = oracle_forward(o, mt, q[i], v[i], ctrl[i], np.zeros(mt.nv))
        assert d.ncon == g["ncon"][i] and d.nefc == g["nefc"][i], (i, d.ncon, g["ncon"][i], d.nefc, g["nefc"][i])
        assert (d.flags & 3) == (flags[i] & 3), (i, d.flags, flags[i])
        assert d.ncon <= 16 and d.nefc <= 40
        nover += int((d.flags & 3) != 0)
        if d.ncon:
            oc = np.array([[c.dist, c.geom1, c.geom2, c.dim] for c in d.contact[:d.ncon]])
            gc = g["contact"][i][: d.ncon][:, [0, 13, 14, 15]]
            assert np.abs(oc - gc).max() < 2e-5
    assert nover > 0, "test state should exceed the capacities at least once"


def test_nonfinite_state_guard():
    from deepmimic_mujoco_b200.sim import BatchedSim
    mt = common.tables()
    sim = BatchedSim(8, motions=("walk",), seed=0)
    q = np.tile(mt.qpos0, (8, 1)); v = np.zeros((8, mt.nv))
    v[3, 10] = 1e12                      # |qvel| >= 1e10 is MuJoCo's mj_checkVel "bad state"
    q[5, 1] = np.nan
    sim.set_state(q, v)
    obs, rew, done = sim.step(torch.zeros(8, 28, device=sim.device))
    done = done.cpu().numpy(); flags = sim.flags.cpu().numpy(); rew = rew.cpu().numpy()
    assert done[3] == 1 and done[5] == 1 and (flags[3] & 4) and (flags[5] & 4) and rew[3] == 0 and rew[5] == 0
    assert done[[0, 1, 2, 4, 6, 7]].sum() == 0 and (flags[[0, 1, 2, 4, 6, 7]] & 4).sum() == 0
    gq, gv, _ = sim.get_state()
    assert np.isfinite(gq).all() and np.isfinite(gv).all()
    assert np.abs(gq[3] - np.float32(mt.qpos0)).max() == 0 and np.abs(gv[3]).max() == 0   # parked at the reference pose
    assert torch.isfinite(obs).all()
    sim.close()


def test_odd_batch_sizes():
    """N not a multiple of the envs per CTA / smaller than one CTA; results independent of N and of the schedule."""
    from deepmimic_mujoco_b200.sim import BatchedSim
    mt = common.tables()
    rng = np.random.default_rng(5)
    q, v = common.standing_states(rng, 37)
    act = common.f32(rng.uniform(-0.5, 0.5, (37, 28)))
    outs = []
    for n in (37, 5, 1):
        sim = BatchedSim(n, motions=("walk",), seed=0)
        sim.set_state(q[:n], v[:n])
        sim.step(torch.tensor(act[:n], dtype=torch.float32, device=sim.device))
        outs.append(sim.get_state()[:2])
        sim.close()
    for n, (gq, gv) in zip((5, 1), outs[1:]):
        assert np.array_equal(gq, outs[0][0][:n]) and np.array_equal(gv, outs[0][1][:n])


@pytest.mark.parametrize("name", ["walk", "spinkick", "dance_b", "run", "backflip"])
def test_mocap_sample_kernel(name):
    """dmb_mocap_sample (in-kernel lerp / slerp, fp32) vs the oracle and vs the golden vectors computed with the
    reference's transformations.quaternion_slerp / euler_from_quaternion (tests/golden/make_interp_golden.py)."""
    import os
    import oracle.pyoracle as po
    from deepmimic_mujoco_b200.sim import BatchedSim, load_motions, make_mocap_struct
    g = np.load(os.path.join(common.GOLDEN, "mocap_interp.npz"))
    sim = BatchedSim(4, motions=("walk", name), seed=1)
    mcs, keep = make_mocap_struct(load_motions(["walk", name]))
    m = common.model()
    us, gq, gv = g[name + "_u"], g[name + "_qpos"], g[name + "_qvel"]
    dt = float(sim.mocap.clip_dt[1])
    q, v, ph = sim.mocap_sample(us * dt, clip_ids=torch.ones(len(us), dtype=torch.int32))
    q, v, ph = q.double().cpu().numpy(), v.double().cpu().numpy(), ph.double().cpu().numpy()
    L = po.lib()
    oq, ov, oph = np.zeros(40), np.zeros(40), C.c_double()
    for i, u in enumerate(us):
        L.dmo_mocap_sample(C.byref(m), C.byref(mcs), 1, float(us[i] * dt / dt), po.dptr(oq), po.dptr(ov), C.byref(oph))
        # Euler angles near +-pi may differ by 2 pi between fp32 and fp64: compare on the circle
        dq = q[i] - oq[:35]; dq[7:] = (dq[7:] + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(dq).max() < 2e-5, (i, u, np.abs(dq).max())
        assert np.abs(v[i] - ov[:34]).max() < 1e-5 * max(1.0, np.abs(ov[:34]).max())
        assert abs(ph[i] - oph.value) < 1e-6
        dg = q[i] - gq[i]; dg[7:] = (dg[7:] + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(dg).max() < 2e-5 and np.abs(v[i] - gv[i]).max() < 1e-5 * max(1.0, np.abs(gv[i]).max())
    sim.close()


def test_gym_surface_dm_state_and_phase():
    from deepmimic_mujoco_b200.env import DPEnv
    env = DPEnv(motion="walk", seed=3, reward_mode=4, phase_mode=1, obs_mode=1)
    assert env.observation_space.shape == (197,)
    ob = env.reset()
    assert ob.shape == (197,) and np.isfinite(ob).all()
    k0 = env.idx_init
    rate = env._sim.tables.timestep / env.mocap_dt
    for t in range(5):
        ob, r, done, _ = env.step(env.action_space.sample())
        assert ob.shape == (197,) and 0.0 < r <= 1.0
        assert env.idx_curr == min(int((k0 + (t + 1) * rate) % 38), 37)
    env.close()
