"""Test infrastructure: the gym shim and a restatement of the reference's episode monitor.

* ``install_gym_shim()`` registers a minimal ``gym`` / ``gym.core`` / ``gym.spaces`` (Env, Wrapper, Box) -- the
  subset /root/reference/src/bench/monitor.py:3-4,12-16 imports -- so that the REFERENCE Monitor can be imported in
  the build container (gym is absent there, like pyquaternion; same approach as tests/golden/make_mocap_golden.py).
* ``MonitorContract`` restates /root/reference/src/bench/monitor.py:12-92 (Monitor + ResultsWriter): needs_reset
  protocol (34-49,51-56), per-episode (r, l, t) rows with r rounded to 6 digits (58-76), the CSV header (100-118).
  It exists because /root/reference does not travel to the GPU box; tests/test_monitor_cpu.py pins it against rows
  produced by the reference's own class (tests/golden/monitor_golden.json, made by make_monitor_golden.py), and
  against the reference class itself when /root/reference is present.
Nothing here is on the product path.
"""
import csv
import json
import sys
import time
import types

import numpy as np


def install_gym_shim():
    if "gym" in sys.modules:
        return sys.modules["gym"]
    gym = types.ModuleType("gym")
    core = types.ModuleType("gym.core")
    spaces = types.ModuleType("gym.spaces")

    class Env:
        metadata = {"render.modes": []}
        reward_range = (-float("inf"), float("inf"))
        spec = None
        action_space = None
        observation_space = None

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.action_space = env.action_space
            self.observation_space = env.observation_space
            self.reward_range = getattr(env, "reward_range", Env.reward_range)
            self.metadata = getattr(env, "metadata", Env.metadata)

        @property
        def spec(self):
            return self.env.spec

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def seed(self, seed=None):
            return self.env.seed(seed)

        def close(self):
            return self.env.close()

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(np.shape(low) if shape is None else shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    gym.Env, gym.Wrapper, gym.core, gym.spaces = Env, Wrapper, core, spaces
    core.Env, core.Wrapper = Env, Wrapper
    spaces.Box = Box
    sys.modules["gym"], sys.modules["gym.core"], sys.modules["gym.spaces"] = gym, core, spaces
    return gym


class MonitorContract:
    """bench/monitor.py:12-92 restated (Monitor(env, filename, allow_early_resets=False))."""
    EXT = "monitor.csv"

    def __init__(self, env, filename=None, allow_early_resets=False):
        self.env = env
        self.action_space, self.observation_space = env.action_space, env.observation_space
        self.tstart = time.time()
        self.f = None
        if filename is not None:                                   # ResultsWriter, monitor.py:100-118
            if not filename.endswith(self.EXT):
                filename = filename + "." + self.EXT
            self.f = open(filename, "wt")
            self.f.write("# {} \n".format(json.dumps({"t_start": self.tstart, "env_id": env.spec and env.spec.id})))
            self.logger = csv.DictWriter(self.f, fieldnames=("r", "l", "t"))
            self.logger.writeheader()
            self.f.flush()
        self.allow_early_resets = allow_early_resets
        self.rewards = None
        self.needs_reset = True
        self.episode_rewards, self.episode_lengths, self.episode_times = [], [], []
        self.total_steps = 0

    def reset(self, **kwargs):                                      # monitor.py:34-49
        if not self.allow_early_resets and not self.needs_reset:
            raise RuntimeError("Tried to reset an environment before done. If you want to allow early resets, "
                               "wrap your env with Monitor(env, path, allow_early_resets=True)")
        self.rewards = []
        self.needs_reset = False
        return self.env.reset(**kwargs)

    def step(self, action):                                         # monitor.py:51-56
        if self.needs_reset:
            raise RuntimeError("Tried to step environment that needs reset")
        ob, rew, done, info = self.env.step(action)
        self.update(ob, rew, done, info)
        return ob, rew, done, info

    def update(self, ob, rew, done, info):                          # monitor.py:58-76
        self.rewards.append(rew)
        if done:
            self.needs_reset = True
            eprew, eplen = sum(self.rewards), len(self.rewards)
            epinfo = {"r": round(eprew, 6), "l": eplen, "t": round(time.time() - self.tstart, 6)}
            self.episode_rewards.append(eprew)
            self.episode_lengths.append(eplen)
            self.episode_times.append(time.time() - self.tstart)
            if self.f is not None:
                self.logger.writerow(epinfo)
                self.f.flush()
            if isinstance(info, dict):
                info["episode"] = epinfo
        self.total_steps += 1

    def close(self):
        if self.f is not None:
            self.f.close()

    def get_total_steps(self):
        return self.total_steps

    def get_episode_rewards(self):
        return self.episode_rewards

    def get_episode_lengths(self):
        return self.episode_lengths


class ScriptedEnv:
    """Deterministic stand-in env: rewards and episode ends follow a fixed script (seeded), obs = step counter."""
    spec = None
    metadata = {}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, seed=0, nsteps=400):
        rng = np.random.RandomState(seed)
        self.rew = np.round(rng.uniform(-1, 2, nsteps), 4).tolist()
        self.done = (rng.uniform(size=nsteps) < 0.07).tolist()
        gym = install_gym_shim()
        self.action_space = gym.spaces.Box(-0.5, 0.5, (28,), np.float32)
        self.observation_space = gym.spaces.Box(-np.inf, np.inf, (56,), np.float64)
        self.t = 0
        self.nreset = 0

    @property
    def unwrapped(self):
        return self

    def reset(self):
        self.nreset += 1
        return np.full(56, float(self.t))

    def reset_model_init(self):
        return np.full(56, -float(self.t))

    def step(self, action):
        r, d = self.rew[self.t], self.done[self.t]
        self.t += 1
        return np.full(56, float(self.t)), r, d, {}


def drive_like_trpo(env, nsteps, policy=None):
    """The loop body of traj_segment_generator (/root/reference/src/trpo.py:47-80) with a random policy:
    ``ob = env.reset()``; per step ``ob, rew, new, _ = env.step(ac)``; on ``new``: ``env.reset()`` then
    ``ob = env.env.reset_model_init()`` (the double reset of trpo.py:78-79).  Returns (ep_rets, ep_lens, obs)."""
    ac = env.action_space.sample()
    ob = env.reset()
    cur_ret, cur_len, ep_rets, ep_lens, obs = 0.0, 0, [], [], []
    for t in range(nsteps):
        ac = env.action_space.sample() if policy is None else policy(ob)
        obs.append(ob)
        ob, rew, new, _ = env.step(ac)
        cur_ret += rew
        cur_len += 1
        if new:
            ep_rets.append(cur_ret); ep_lens.append(cur_len)
            cur_ret, cur_len = 0.0, 0
            env.reset()
            ob = env.env.reset_model_init()
    return ep_rets, ep_lens, obs
