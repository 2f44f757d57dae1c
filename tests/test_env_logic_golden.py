"""Env logic of the oracle env (dmo_env_*; the CUDA env is compared with it in tests/test_gpu_parity.py) against golden
sequences produced by the REFERENCE class dp_env_v3.DPEnv itself, run unmodified over a mujoco-py-shaped adapter whose
simulator calls are the oracle's (tests/golden/make_env_logic_golden.py).  Physics is common to both sides, so every
difference would be env logic: observation slicing (dp_env_v3.py:62-65), the constant reward (:117), the dormant pose
reward and its frame counter (:89-104), CoM-height termination on the stale xipos (:134-139), reference-state
initialisation from the mocap tables (:148-156), the standing-pose reset (:158-164).  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import common
import oracle.pyoracle as po
from deepmimic_mujoco_b200.model_blob import default_config
from deepmimic_mujoco_b200.sim import load_motions, make_mocap_struct

G = np.load(os.path.join(common.GOLDEN, "env_logic_golden.npz"))


def _env(motion, reward_mode):
    m, L = common.model(), po.lib()
    cfg = default_config(reward_mode=reward_mode, reset_mode=0, auto_reset=0)
    mcs, keep = make_mocap_struct(load_motions([motion]))
    e = po.DmoEnv()
    L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 1, 0, 0)
    return m, L, cfg, mcs, keep, e


@pytest.mark.parametrize("motion", ["walk", "dance_b"])
@pytest.mark.parametrize("reward_mode", [0, 1])
def test_step_obs_reward_done_against_the_reference_class(motion, reward_mode):
    g = lambda k: G[f"{motion}/{k}"]
    m, L, cfg, mcs, keep, e = _env(motion, reward_mode)
    mt = common.tables()
    obs, rew = np.zeros(256), C.c_double()
    T = len(g("action"))
    ndone = 0
    for t in range(T):
        np.ctypeslib.as_array(e.d.qpos)[: mt.nq] = g("pre_qpos")[t]
        np.ctypeslib.as_array(e.d.qvel)[: mt.nv] = g("pre_qvel")[t]
        np.ctypeslib.as_array(e.d.qacc_warmstart)[: mt.nv] = g("pre_warm")[t]
        e.idx_curr = int(g("idx_before")[t])
        a = np.ascontiguousarray(g("action")[t])
        done = L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(a), po.dptr(obs), C.byref(rew))
        # same physics underneath: the post-step state is the golden's to round-off
        assert np.abs(np.ctypeslib.as_array(e.d.qpos)[: mt.nq] - g("qpos")[t]).max() < 1e-12
        assert np.abs(np.ctypeslib.as_array(e.d.qvel)[: mt.nv] - g("qvel")[t]).max() < 1e-10
        assert np.abs(obs[:56] - g("obs")[t]).max() < 1e-10                      # _get_obs: qpos[7:] || qvel[6:]
        assert abs(e.zcom_last - g("zcom")[t]) < 1e-12                           # sum(m xipos) / sum(m), stale stage
        assert bool(done) == bool(g("done")[t]), t                               # is_done
        if reward_mode == 0:
            assert rew.value == 1.0 == g("reward")[t]                            # reward_alive
        else:                                                                    # calc_config_reward (fp32 tables here)
            assert abs(rew.value - g("cfg_reward")[t]) < 1e-5 * max(1e-3, g("cfg_reward")[t]) + 1e-9
            assert e.idx_curr == int(g("idx_after")[t])                          # (idx + 1) % len after the reward
        ndone += int(done)
    assert ndone == int(g("done").sum()) >= 2


@pytest.mark.parametrize("motion", ["walk", "dance_b"])
def test_reference_state_init_and_standing_reset_against_the_reference_class(motion):
    g = lambda k: G[f"{motion}/{k}"]
    c = common.clip(motion)
    vel = np.nan_to_num(c.data_vel)
    for i, idx in enumerate(g("reset_idx")):            # reset_model: state <- data_config[idx_init], data_vel[idx_init]
        assert 0 <= idx < len(c)
        assert np.abs(c.data_config[idx] - g("reset_qpos")[i]).max() < 1e-12
        assert np.abs(vel[idx] - g("reset_qvel")[i]).max() < 1e-10
        assert np.abs(np.r_[c.data_config[idx][7:], vel[idx][6:]] - g("reset_obs")[i]).max() < 1e-10
    # ... and what the oracle env does for the same frame: the fp32-rounded table row, frame counter = idx_init
    m, L, cfg, mcs, keep, e = _env(motion, 0)
    mt = common.tables()
    seen = set()
    for k in range(40):
        L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 0)
        e.reset_count += 1
        assert e.idx_curr == e.idx_init and 0 <= e.idx_init < len(c)
        seen.add(e.idx_init)
        q = np.ctypeslib.as_array(e.d.qpos)[: mt.nq]
        ref = common.f32(c.data_config[e.idx_init])
        ref[3:7] /= np.linalg.norm(ref[3:7])
        assert np.abs(q - ref).max() < 1e-6 and np.abs(np.ctypeslib.as_array(e.d.qvel)[: mt.nv] - common.f32(vel[e.idx_init])).max() < 1e-6
    assert len(seen) > 15                                 # uniform over the frames (random.randint in the reference)
    # reset_model_init: init_qpos / init_qvel + U(-0.01, 0.01) noise (np_random of the env; first draw of seed 3)
    assert np.abs(g("init_qpos") - (mt.qpos0 + g("init_noise"))).max() < 1e-15
    assert np.abs(g("init_qvel")).max() <= 0.01 and np.abs(g("init_obs")[:28] - g("init_qpos")[7:]).max() == 0.0
    L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 1)
    q = np.ctypeslib.as_array(e.d.qpos)[: mt.nq]
    d = q - mt.qpos0
    assert 0 < np.abs(d).max() <= 0.01 + 1e-7 and np.abs(np.ctypeslib.as_array(e.d.qvel)[: mt.nv]).max() <= 0.01 + 1e-7
