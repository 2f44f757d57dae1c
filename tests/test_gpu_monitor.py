"""The reference's own callers driven over the drop-in env on the GPU:
the episode monitor of /root/reference/src/bench/monitor.py:12-92 (restated in tests/monitor_contract.py and pinned
against the reference class by tests/test_monitor_cpu.py) wrapped around ``DPEnv``, driven by the loop body of
/root/reference/src/trpo.py:47-80 including the double reset of trpo.py:78-79."""
import csv

import numpy as np
import pytest

import common  # noqa: F401
from monitor_contract import MonitorContract, drive_like_trpo

pytestmark = pytest.mark.gpu


def test_monitor_around_dpenv_trpo_loop(tmp_path):
    from deepmimic_mujoco_b200.env import DPEnv
    env = DPEnv(motion="walk", seed=0)
    env.seed(0)
    path = str(tmp_path / "dpenv")
    mon = MonitorContract(env, path)
    with pytest.raises(RuntimeError):
        mon.step(env.action_space.sample())            # monitor.py:52-53: must reset first
    ep_rets, ep_lens, obs = drive_like_trpo(mon, 300)
    assert len(ep_lens) >= 3 and sum(ep_lens) <= 300
    # reward is identically 1 in the shipped env (dp_env_v3.py:117): return == length
    assert all(abs(r - l) < 1e-9 for r, l in zip(ep_rets, ep_lens))
    assert mon.get_episode_lengths() == ep_lens and mon.get_total_steps() == 300
    # a near-random policy from the standing pose +-0.01 falls within a few tens of steps (progress.csv:2-4)
    assert 8 < np.mean(ep_lens) < 80, ep_lens
    if not mon.needs_reset:
        with pytest.raises(RuntimeError):
            mon.reset()                                # monitor.py:44-45: no early reset
    mon.close()
    with open(path + ".monitor.csv") as f:
        assert f.readline().startswith("#")
        rows = list(csv.DictReader(f))
    assert [int(r["l"]) for r in rows] == ep_lens and all(float(r["r"]) == int(r["l"]) for r in rows)
    # the observation after the double reset is the default pose +- 0.01 (reset_model_init), not the mocap frame
    first_after_reset = [o for o in obs[1:] if np.abs(o[:28]).max() <= 0.0101]
    assert len(first_after_reset) >= len(ep_lens) - 1
    env.close()


def test_sim_data_views_write_through():
    """``env.sim.data.qpos[:] = x`` (mujoco-py idiom, dp_env_v3.py:192-197) must change the simulation state."""
    from deepmimic_mujoco_b200.env import DPEnv
    env = DPEnv(motion="walk", seed=0)
    env.reset()
    q = env.sim.data.qpos.copy()
    q2 = q.copy(); q2[2] += 0.25; q2[7:] *= 0.5
    env.sim.data.qpos[:] = q2
    assert np.abs(env.sim.data.qpos - np.float32(q2)).max() < 1e-7
    env.sim.data.qvel[3] = 1.5
    assert abs(env.sim.data.qvel[3] - 1.5) < 1e-7
    env.sim.forward()
    ob, _, _, _ = env.step(np.zeros(28, np.float32))
    assert np.isfinite(ob).all()
    env.set_state(env.mocap.data_config[3], env.mocap.data_vel[3])       # dp_env_v3.py:196 idiom
    assert np.abs(env.sim.data.qpos - np.float32(env.mocap.data_config[3])).max() < 1e-6
    env.close()


def test_bad_clip_ids_rejected_and_launch_counter():
    """Caller-supplied clip ids are validated up front (a bad id would index the clip tables out of bounds), and the
    library counts the kernels it launches (bench.py reports that count)."""
    import torch
    from deepmimic_mujoco_b200.sim import BatchedSim
    with pytest.raises(ValueError):
        BatchedSim(8, motions=("walk", "run"), clip_ids=torch.tensor([0, 1, 2, 0, 1, 0, 1, 0], dtype=torch.int32))
    with pytest.raises(ValueError):
        BatchedSim(8, motions=("walk",), clip_ids=torch.zeros(7, dtype=torch.int32))
    sim = BatchedSim(64, motions=("walk",), seed=0)
    sim.reset()
    n0 = sim.kernel_launches()
    act = torch.zeros(64, 28, device=sim.device)
    for _ in range(16):
        sim.step(act)
    n = sim.kernel_launches() - n0
    assert 16 + 2 <= n <= 32          # 16 step kernels + the scheduler sort on every 8th step (every step at most)
    sim.close()
