#!/usr/bin/env python3
"""(build container, CPU) Dry run of the LOGIC of tests/test_z_gpu_reference_log.py -- the three GPU tests that pin the
CUDA path against the reference's MuJoCo-produced monitor log and its MuJoCo-trained policy -- with stand-ins for the
two CUDA-only classes: DPVecEnv backed by the float64 oracle (one Oracle per env, reset_model_init on done, vec_env
auto-reset semantics) and MlpPolicy.act evaluated with torch on the CPU.  The real tf_checkpoint.policy_arrays,
MlpPolicy.load_arrays and rollout.evaluate run unchanged.  Written because those tests were added after round 2's GPU
minutes were spent: this checks their bookkeeping (first-episode lengths, censoring, acceptance rule) before their first
run on hardware.  usage: python tools/dry_run_gpu_reference_tests.py [envs=192]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import oracle.pyoracle as po  # noqa: E402
import deepmimic_mujoco_b200.env as envmod  # noqa: E402
import deepmimic_mujoco_b200.policy as polmod  # noqa: E402


class OracleVecEnv:
    def __init__(self, n, motions=("walk",), seed=0, reward_mode=0, reset_mode=1, auto_reset=True):
        assert reward_mode == 0 and reset_mode == 1 and auto_reset
        self.num_envs, self.mt, self.rng = n, common.tables(), np.random.default_rng(seed)
        self.o = [po.Oracle(common.model()) for _ in range(n)]
        self.ep_len, self.last_len = np.zeros(n, np.int32), np.zeros(n, np.int32)

    def _reset(self, i):
        mt = self.mt
        self.o[i].set_state(mt.qpos0 + self.rng.uniform(-0.01, 0.01, mt.nq), self.rng.uniform(-0.01, 0.01, mt.nv))
        self.ep_len[i] = 0

    def _obs(self):
        return torch.tensor(np.array([np.concatenate([o.qpos[7:], o.qvel[6:]]) for o in self.o]), dtype=torch.float32)

    def reset(self):
        for i in range(self.num_envs):
            self._reset(i)
        return self._obs()

    def step(self, act):
        a, done = act.numpy().astype(np.float64), np.zeros(self.num_envs, np.uint8)
        for i, o in enumerate(self.o):
            o.d.arr("ctrl")[:28] = a[i]
            o.step()
            self.ep_len[i] += 1
            z = o.d.arr("com")[2]
            if z < 0.7 or z > 2.0:
                done[i], self.last_len[i] = 1, self.ep_len[i]
                self._reset(i)
        info = {"episode_length": torch.tensor(self.last_len.copy()), "episode_return": torch.tensor(self.last_len.astype(np.float32))}
        return self._obs(), torch.ones(self.num_envs), torch.tensor(done), info

    def close(self):
        pass


class TorchPolicy(polmod.MlpPolicy):
    def __init__(self, seed=0, **kw):
        z = lambda *s: torch.zeros(*s)
        self.params = dict(vw1=z(56, 100), vb1=z(100), vw2=z(100, 100), vb2=z(100), vw3=z(100, 1), vb3=z(1), pw1=z(56, 100),
                           pb1=z(100), pw2=z(100, 100), pb2=z(100), pw3=z(100, 28), pb3=z(28), logstd=z(28))
        self.ob_rms = polmod.RunningMeanStd((56,), "cpu")
        self.g = torch.Generator(); self.g.manual_seed(seed)

    def act(self, stochastic, ob, out_ac=None, out_vpred=None, out_mean=None, first_row=0):
        p = self.params
        h = torch.clamp((ob - self.ob_rms.mean) / self.ob_rms.std, -5, 5)
        m = torch.tanh(torch.tanh(h @ p["pw1"] + p["pb1"]) @ p["pw2"] + p["pb2"]) @ p["pw3"] + p["pb3"]
        if out_mean is not None:
            out_mean.copy_(m)
        ac = m + torch.exp(p["logstd"]) * torch.randn(m.shape, generator=self.g) if stochastic else m
        return ac, torch.zeros(ob.shape[0])


def main():
    envmod.DPVecEnv, polmod.MlpPolicy = OracleVecEnv, TorchPolicy
    import test_z_gpu_reference_log as t
    t.DEVICE, t.N_ENVS = "cpu", int(sys.argv[1]) if len(sys.argv) > 1 else 192
    for name in ("test_fall_time_distribution_matches_reference_monitor_log",
                 "test_trained_policy_survival_matches_reference_monitor_log",
                 "test_reference_checkpoint_through_the_fused_policy_kernel_and_evaluate"):
        t0 = time.time()
        getattr(t, name)()
        print(f"{name}: ok ({t.N_ENVS} oracle-backed envs, {time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
