"""Drop-in env surfaces over BatchedSim.

* :class:`DPEnv` keeps the gym.Env surface of the reference env
  (/root/reference/src/dp_env_v3.py:34-171 ``DPEnv``): no-arg constructor, ``observation_space`` (56,)
  float64, ``action_space`` Box(-0.5, 0.5, (28,), float32), ``reset() -> ob``,
  ``step(ac) -> (ob, rew, done, info)``, ``reset_model_init()``, ``seed``, ``close``, ``render``,
  ``set_state``, ``dt``, ``mocap.data_config/data_vel``, ``idx_curr`` -- so it drops in under
  /root/reference/src/trpo.py:27-80 (``traj_segment_generator``) and bench/monitor.py.  It is a
  num_envs=1 view of the batched CUDA path (numpy in/out); there is no CPU path.
* :class:`DPVecEnv` is the batched surface modelled on the reference's unused VecEnv ABC
  (/root/reference/src/utils/vec_env/__init__.py:26-131): ``reset``, ``step_async``, ``step_wait``,
  ``step`` with auto-reset of done envs inside ``step`` (dummy_vec_env.py:51-54 semantics); all
  tensors are CUDA float32, env-major.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .model_blob import default_config
from .sim import BatchedSim

try:  # the reference requires real gym spaces when gym is importable (mlp_policy_trpo.py:25)
    from gym import spaces as _spaces  # type: ignore
    _Box = _spaces.Box
    try:
        import gym as _gym  # type: ignore
        _EnvBase = _gym.Env
    except Exception:  # pragma: no cover
        _EnvBase = object
except Exception:
    _EnvBase = object

    class _Box:  # minimal stand-in with the attributes the reference touches
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, dtype=self.dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box{self.shape}"


class Config:
    """Mirror of /root/reference/src/config.py:3-17 defaults (motion is selected by name here)."""
    motion = "dance_b"
    env_name = "dp_env_v3"


class DPVecEnv:
    """N batched envs on one GPU with vec_env semantics (auto-reset inside step)."""

    def __init__(self, num_envs: int, motions: Sequence[str] = ("walk",), device=None, seed: int = 0,
                 first_env_id: int = 0, reward_mode: int = 0, ctrl_mode: int = 0, reset_mode: int = 0,
                 auto_reset: bool = True, clip_ids: Optional[torch.Tensor] = None, phase_mode: int = 0,
                 obs_mode: int = 0, rec_depth: int = 4, **cfg_kw):
        """phase_mode 1: time-based mocap phase with lerp/slerp interpolation (instead of one frame per step);
        obs_mode 1: the 197-d DeepMimic state instead of qpos[7:] || qvel[6:]."""
        cfg = default_config(reward_mode=reward_mode, ctrl_mode=ctrl_mode, reset_mode=reset_mode,
                             auto_reset=int(auto_reset), phase_mode=phase_mode, obs_mode=obs_mode, **cfg_kw)
        ref_aux = None
        if reward_mode == 4:
            from .refaux import compute_ref_aux
            ref_aux = compute_ref_aux(motions)
        self.sim = BatchedSim(num_envs, motions=motions, device=device, seed=seed, first_env_id=first_env_id,
                              config=cfg, clip_ids=clip_ids, ref_aux=ref_aux, rec_depth=rec_depth)
        self.num_envs = num_envs
        self.observation_space = _Box(-np.inf, np.inf, (self.sim.obs_dim,), np.float64)
        lo, hi = self.sim.tables.act_ctrlrange[:, 0], self.sim.tables.act_ctrlrange[:, 1]
        self.action_space = _Box(lo.astype(np.float32), hi.astype(np.float32), (self.sim.nu,), np.float32)
        self._pending = None

    def reset(self) -> torch.Tensor:
        return self.sim.reset()

    def step_async(self, actions: torch.Tensor, rec_host=None) -> None:
        self._pending = (actions, rec_host)

    def step_wait(self):
        obs, rew, done = self.sim.step(self._pending[0], rec_host=self._pending[1])
        self._pending = None
        infos = {"episode_return": self.sim.last_ret, "episode_length": self.sim.last_len, "flags": self.sim.flags}
        return obs, rew, done, infos

    def step(self, actions: torch.Tensor, rec_host=None):
        """``rec_host`` (a ``sim.alloc_host((N, obs_dim + 2))`` array): the kernel writes the packed (obs, reward, done)
        record of this step straight into pinned host memory -- no device-to-host copy to enqueue afterwards."""
        self.step_async(actions, rec_host)
        return self.step_wait()

    def step_host(self, act=None, rec=None):
        """``step`` for a host-resident policy: actions are read from, and the packed (obs, reward, done) record is
        written to, pinned host arrays by the step kernel itself (``BatchedSim.step_host``; buffers from
        ``self.sim.enable_host_io()`` / ``self.sim.alloc_host``).  Returns the record HostArray; synchronise the
        stream before reading it.  Finished-episode statistics stay on the device (``self.sim.last_ret/last_len``)."""
        return self.sim.step_host(act, rec)

    def state_dict(self):
        s = self.sim
        return {k: getattr(s, k).clone() for k in ("qpos", "qvel", "warm", "clip", "idx_init", "idx_curr",
                                                    "reset_count", "ep_len", "ep_ret", "flags")}

    def load_state_dict(self, sd):
        for k, v in sd.items():
            getattr(self.sim, k).copy_(v)

    def close(self):
        self.sim.close()


class _StateView(np.ndarray):
    """``env.sim.data.qpos`` / ``qvel``: a float64 host copy of one state row whose item assignment writes
    through to the device tensor, so the mujoco-py idiom ``sim.data.qpos[:] = x`` (dp_env_v3.py:192-197 style
    scripts) changes the simulation state instead of a temporary."""

    def __new__(cls, values, push):
        obj = np.asarray(values, dtype=np.float64).view(cls)
        obj._push = push
        return obj

    def __array_finalize__(self, obj):
        self._push = getattr(obj, "_push", None)

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        root = self
        while isinstance(root.base, _StateView):     # a slice of the view writes through as well
            root = root.base
        if root._push is not None:
            root._push(np.asarray(root))


class _SimData:
    """``env.sim.data`` (the mjData fields the reference's scripts touch)."""

    def __init__(self, env):
        self._env = env

    @property
    def qpos(self):
        s = self._env._sim
        def push(a):
            s.qpos[0, : s.nq] = torch.as_tensor(a, dtype=torch.float32, device=s.device)
        return _StateView(s.qpos[0, : s.nq].double().cpu().numpy(), push)

    @property
    def qvel(self):
        s = self._env._sim
        def push(a):
            s.qvel[0, : s.nv] = torch.as_tensor(a, dtype=torch.float32, device=s.device)
        return _StateView(s.qvel[0, : s.nv].double().cpu().numpy(), push)

    @property
    def ctrl(self):
        return self._env._hact.array[0].astype(np.float64)


class _SimView:
    """``env.sim``: ``data`` plus ``forward()`` (mj_forward at the current state) as mujoco-py's MjSim."""

    def __init__(self, env):
        self.data = _SimData(env)
        self._env = env

    def forward(self):
        self._env._forward()


class _MocapView:
    """``env.mocap`` as poked by the reference scripts (dp_env_v3.py:192-197, env_torque_test.py)."""

    def __init__(self, clip):
        self.data = clip.data
        self.data_config = list(clip.data_config)
        self.data_vel = list(clip.data_vel)
        self.dt = clip.dt


class DPEnv(_EnvBase):
    """gym-surface single env (num_envs = 1 on the CUDA path)."""

    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def __init__(self, motion: Optional[str] = None, device=None, seed: int = 0, reward_mode: int = 0,
                 ctrl_mode: int = 0, frame_skip: int = 6, phase_mode: int = 0, obs_mode: int = 0):
        from .mocap import load_clip
        from .sim import motion_path
        self.motion = motion or Config.motion
        cfg = default_config(reward_mode=reward_mode, ctrl_mode=ctrl_mode, reset_mode=0, auto_reset=0,
                             phase_mode=phase_mode, obs_mode=obs_mode)
        ref_aux = None
        if reward_mode == 4:
            from .refaux import compute_ref_aux
            ref_aux = compute_ref_aux([self.motion])
        self._seed = seed
        self._sim = BatchedSim(1, motions=(self.motion,), device=device, seed=seed, config=cfg, ref_aux=ref_aux)
        clip = load_clip(motion_path(self.motion), name=self.motion)
        self.mocap = _MocapView(clip)
        self.mocap_dt = clip.dt
        self.mocap_data_len = len(clip)
        t = self._sim.tables
        self.frame_skip = frame_skip            # accepted and ignored, as in the reference (App. F #1)
        self.init_qpos, self.init_qvel = t.qpos0.copy(), np.zeros(t.nv)
        self.action_space = _Box(t.act_ctrlrange[:, 0].astype(np.float32), t.act_ctrlrange[:, 1].astype(np.float32),
                                 (t.nu,), np.float32)
        self.np_random = np.random.RandomState(seed)
        self._act = torch.zeros(1, t.nu, dtype=torch.float32, device=self._sim.device)
        # one launch per step: the kernel reads the action from / writes the record to pinned host memory
        self._hact, self._hrec = self._sim.enable_host_io()
        # gym MujocoEnv.__init__ side effect: one probe step with a random action
        ob, _, done, _ = self.step(self.action_space.sample())
        assert not done
        self.observation_space = _Box(-np.inf, np.inf, (ob.size,), np.float64)
        self.sim = _SimView(self)  # env.sim.data.qpos / qvel (write-through), env.sim.forward()

    # --- gym.Env ------------------------------------------------------------------------
    @property
    def unwrapped(self):
        return self

    @property
    def env(self):  # trpo.py:79 calls env.env.reset_model_init() through the Monitor wrapper
        return self

    @property
    def dt(self):
        return self._sim.tables.timestep * self.frame_skip

    @property
    def idx_curr(self):
        return int(self._sim.idx_curr[0].item())

    @property
    def idx_init(self):
        return int(self._sim.idx_init[0].item())

    @property
    def qpos(self):
        return self._sim.qpos[0, : self._sim.nq].double().cpu().numpy()

    @property
    def qvel(self):
        return self._sim.qvel[0, : self._sim.nv].double().cpu().numpy()

    def seed(self, seed=None):
        self._seed = 0 if seed is None else int(seed)
        self.np_random = np.random.RandomState(self._seed)
        self.action_space.seed(self._seed)
        return [self._seed]

    def _get_obs(self):
        return self._sim.get_obs()[0].double().cpu().numpy()

    def _forward(self):
        """gym ``MujocoEnv.set_state`` ends with ``sim.forward()``: one mj_forward (data.ctrl = the last action) that
        leaves qacc_warmstart = qacc for the next step (the batched auto-reset path starts from 0 instead)."""
        self._act.copy_(self._hact.tensor)
        self._sim.forward_debug(self._act)

    def set_state(self, qpos, qvel):
        self._sim.set_state(np.asarray(qpos)[None], np.asarray(qvel)[None])
        self._forward()                                           # data.ctrl still holds the last action

    def reset(self):
        ob = self._sim.reset(mode=0)[0].double().cpu().numpy()   # reset_model(): mocap RSI
        self._hact.array[:] = 0.0                                 # sim.reset() clears data.ctrl
        self._forward()
        return ob

    def reset_model(self):
        return self.reset()

    def reset_model_init(self):
        c = 0.01
        self.set_state(self.init_qpos + self.np_random.uniform(low=-c, high=c, size=self.init_qpos.size),
                       self.init_qvel + self.np_random.uniform(low=-c, high=c, size=self.init_qvel.size))
        return self._get_obs()

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, -1)
        if a.shape[1] != self._sim.nu:
            raise ValueError(f"action must have {self._sim.nu} entries")
        self._hact.array[:] = a
        self._sim.step_host(self._hact, self._hrec)
        torch.cuda.current_stream(self._sim.device).synchronize()
        out = self._hrec.array[0].astype(np.float64)
        return out[:-2], float(out[-2]), bool(out[-1] != 0.0), {}

    def render(self, mode="human"):
        return None

    def close(self):
        self._sim.close()
