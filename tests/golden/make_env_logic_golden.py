#!/usr/bin/env python3
"""Generate tests/golden/env_logic_golden.npz by running the REFERENCE env class itself --
/root/reference/src/dp_env_v3.py ``DPEnv`` (step 106-132, _get_obs 62-65, is_done 134-139, calc_config_reward 89-104,
reset_model 148-156, reference_state_init 67-71, reset_model_init 158-164), unmodified -- here, with the reference's own
mocap loader underneath.  Its three absent third-party dependencies are replaced by adapters:

* ``pyquaternion``: the shim of make_mocap_golden.py;
* ``gym``: ``gym.envs.mujoco.mujoco_env.MujocoEnv`` restated from gym 0.10-0.15 (``__init__`` keeps model / sim / data,
  ``init_qpos`` / ``init_qvel``, ``np_random``; ``set_state`` = write qpos / qvel + ``sim.forward()``;
  ``do_simulation`` = ``sim.data.ctrl[:] = ctrl`` + ``n_frames`` x ``sim.step()``; ``reset`` = ``sim.reset()`` +
  ``reset_model()``; ``dt``), ``gym.utils.EzPickle`` (no-op), ``gym.spaces.Box``;
* ``mujoco_py``: an ``MjSim``-shaped object whose ``step()`` / ``forward()`` call THIS REPO'S float64 oracle
  (oracle/dm_oracle.c) and mirror ``data.qpos / qvel / ctrl / xipos`` and ``model.body_mass / nq / nv``.

So everything above the simulator call is the reference's code and everything below it is the oracle's physics: the
golden pins the ENV LOGIC of the oracle env (``dmo_env_*``: observation slicing, constant reward, the dormant pose
reward and its frame counter, CoM-height termination on the stale last-stage ``xipos``, reference-state
initialisation from the mocap tables, the standing-pose reset) against the reference class, state for state.  The
physics itself is NOT pinned by this file.  Run in the build container only."""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
import common  # noqa: E402
import oracle.pyoracle as po  # noqa: E402
from make_mocap_golden import Quaternion  # noqa: E402


class OracleSim:
    """mujoco_py.MjSim surface used by dp_env_v3 / gym MujocoEnv, on top of the oracle."""

    def __init__(self):
        mt = common.tables()
        self.o = po.Oracle(common.model())
        self.mt = mt
        self.model = types.SimpleNamespace(body_mass=np.asarray(mt.body_mass, dtype=np.float64).copy(), nq=mt.nq, nv=mt.nv,
                                           nu=mt.nu, opt=types.SimpleNamespace(timestep=mt.timestep),
                                           actuator_ctrlrange=np.asarray(mt.act_ctrlrange, dtype=np.float64).copy())
        self.data = types.SimpleNamespace(qpos=mt.qpos0.astype(np.float64).copy(), qvel=np.zeros(mt.nv), ctrl=np.zeros(mt.nu),
                                          xipos=np.zeros((mt.nbody, 3)), time=0.0)
        self.reset()

    def _push(self):
        self.o.d.arr("qpos")[: self.mt.nq] = self.data.qpos
        self.o.d.arr("qvel")[: self.mt.nv] = self.data.qvel
        self.o.d.arr("ctrl")[: self.mt.nu] = self.data.ctrl

    def _pull(self):
        self.data.qpos[:] = self.o.qpos
        self.data.qvel[:] = self.o.qvel
        self.data.xipos[:] = self.o.d.arr("xipos")[: self.mt.nbody]

    def reset(self):                                      # mj_resetData: qpos0, zero velocity / ctrl / warmstart
        self.o.set_state(self.mt.qpos0, np.zeros(self.mt.nv))
        self.data.ctrl[:] = 0.0
        self.data.time = 0.0
        self._pull()

    def forward(self):                                    # mj_forward (leaves qacc_warmstart = qacc)
        self._push()
        self.o.forward()
        self.o.d.arr("qacc_warmstart")[: self.mt.nv] = self.o.d.arr("qacc")[: self.mt.nv]
        self._pull()

    def step(self):                                       # mj_step; data.xipos stays that of the last RK4 stage
        self._push()
        self.o.step()
        self.data.time += self.mt.timestep
        self._pull()


def install_shims():
    """pyquaternion, mujoco_py and gym adapters (gym.core / gym.spaces come from tests/monitor_contract.py, so that the
    reference's bench.Monitor can wrap the env as well)."""
    from monitor_contract import install_gym_shim
    gym = install_gym_shim()
    Box = gym.spaces.Box
    pq = types.ModuleType("pyquaternion"); pq.Quaternion = Quaternion
    mj = types.ModuleType("mujoco_py"); mj.load_model_from_xml = mj.MjSim = mj.MjViewer = None

    class MujocoEnv(gym.Env):
        def __init__(self, model_path, frame_skip):
            assert os.path.exists(model_path), model_path
            self.frame_skip = frame_skip
            self.sim = OracleSim()
            self.model, self.data = self.sim.model, self.sim.data
            self.viewer = None
            self.metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": int(np.round(1.0 / self.dt))}
            self.init_qpos = self.sim.data.qpos.ravel().copy()
            self.init_qvel = self.sim.data.qvel.ravel().copy()
            self.np_random = np.random.RandomState(0)
            observation, _reward, done, _info = self.step(np.zeros(self.model.nu))
            assert not done
            self.obs_dim = observation.size
            b = self.model.actuator_ctrlrange.copy()
            self.action_space = Box(b[:, 0], b[:, 1], dtype=np.float32)
            self.observation_space = Box(-np.inf * np.ones(self.obs_dim), np.inf * np.ones(self.obs_dim), dtype=np.float64)

        def seed(self, seed=None):
            self.np_random = np.random.RandomState(seed)
            return [seed]

        def reset(self):
            self.sim.reset()
            return self.reset_model()

        def set_state(self, qpos, qvel):
            assert qpos.shape == (self.model.nq,) and qvel.shape == (self.model.nv,)
            self.sim.data.qpos[:] = qpos
            self.sim.data.qvel[:] = qvel
            self.sim.forward()

        @property
        def dt(self):
            return self.model.opt.timestep * self.frame_skip

        def do_simulation(self, ctrl, n_frames):
            self.sim.data.ctrl[:] = ctrl
            for _ in range(n_frames):
                self.sim.step()

        def render(self, *a, **k):
            return None

        def close(self):
            pass

    class EzPickle:
        def __init__(self, *a, **k):
            pass

    envs = types.ModuleType("gym.envs"); mjm = types.ModuleType("gym.envs.mujoco"); me = types.ModuleType("gym.envs.mujoco.mujoco_env")
    utils = types.ModuleType("gym.utils")
    me.MujocoEnv, utils.EzPickle = MujocoEnv, EzPickle
    gym.envs, gym.utils, envs.mujoco, mjm.mujoco_env = envs, utils, mjm, me
    sys.modules.update({"pyquaternion": pq, "mujoco_py": mj, "gym.envs": envs, "gym.envs.mujoco": mjm,
                        "gym.envs.mujoco.mujoco_env": me, "gym.utils": utils})


def main():
    import warnings
    warnings.simplefilter("ignore")
    install_shims()
    os.chdir(REF_SRC)                                     # config.py builds its paths from getcwd() at import time
    sys.path.insert(0, REF_SRC)
    from config import Config
    import dp_env_v3                                      # the reference module, unmodified
    out = {}
    for motion in ("walk", "dance_b"):
        Config.mocap_path = "%s%s/humanoid3d_%s.txt" % (Config.curr_path, Config.motion_folder, motion)
        random.seed(11)                                   # reference_state_init draws idx_init from python `random`
        env = dp_env_v3.DPEnv()
        assert env.mocap_data_len == {"walk": 39, "dance_b": 153}[motion] and env.obs_dim == 56
        rng = np.random.default_rng(5)
        T = 64
        rec = {k: [] for k in ("reset_idx", "reset_obs", "reset_qpos", "reset_qvel", "t_reset", "action", "obs", "reward",
                               "done", "qpos", "qvel", "cfg_reward", "idx_after", "pre_qpos", "pre_qvel", "pre_warm", "idx_before", "zcom")}
        ob = env.reset()
        t = 0
        while t < T:
            if t == 0 or done:
                if t:
                    ob = env.reset()
                rec["reset_idx"].append(env.idx_init); rec["reset_obs"].append(ob.copy()); rec["t_reset"].append(t)
                rec["reset_qpos"].append(env.sim.data.qpos.copy()); rec["reset_qvel"].append(env.sim.data.qvel.copy())
            a = rng.uniform(-0.5, 0.5, 28) * (3.0 if t % 7 == 0 else 1.0)       # some actions beyond the ctrlrange
            rec["pre_qpos"].append(env.sim.data.qpos.copy()); rec["pre_qvel"].append(env.sim.data.qvel.copy())
            rec["pre_warm"].append(env.sim.o.d.arr("qacc_warmstart")[:34].copy()); rec["idx_before"].append(env.idx_curr)
            ob, rew, done, info = env.step(a)
            assert info == {}
            mass = env.model.body_mass[:, None]
            rec["zcom"].append(float((np.sum(mass * env.sim.data.xipos, 0) / np.sum(mass))[2]))
            rec["action"].append(a); rec["obs"].append(ob.copy()); rec["reward"].append(rew); rec["done"].append(done)
            rec["qpos"].append(env.sim.data.qpos.copy()); rec["qvel"].append(env.sim.data.qvel.copy())
            rec["cfg_reward"].append(env.calc_config_reward())                    # dormant in step(): dp_env_v3.py:119,127
            rec["idx_after"].append(env.idx_curr)
            t += 1
        for k, v in rec.items():
            out[f"{motion}/{k}"] = np.asarray(v)
        # the standing-pose reset of trpo.py:78-79 (env.reset() then env.env.reset_model_init())
        env.seed(3)
        ob0 = env.reset_model_init()
        out[f"{motion}/init_obs"], out[f"{motion}/init_qpos"], out[f"{motion}/init_qvel"] = ob0, env.sim.data.qpos.copy(), env.sim.data.qvel.copy()
        out[f"{motion}/init_noise"] = np.random.RandomState(3).uniform(-0.01, 0.01, size=35)   # first draw of that seed
        print(motion, "episodes", len(rec["reset_idx"]), "idx_init", rec["reset_idx"], "done steps", int(np.sum(rec["done"])),
              "cfg reward range", float(np.min(rec["cfg_reward"])), float(np.max(rec["cfg_reward"])))
    np.savez_compressed(os.path.join(HERE, "env_logic_golden.npz"), **out)


if __name__ == "__main__":
    main()
