"""The CUDA path against the one MuJoCo-PRODUCED artefact of the reference: the episode monitor of its own training
run (tests/golden/ref_episode_lengths.json, made by tests/golden/make_ref_episode_golden.py) and the policy that run
trained in MuJoCo (tests/golden/ref_trained_policy.npz, made by tests/golden/make_ref_policy_golden.py).  Same protocol as the
rows were recorded with (trpo.py:27-80): standing pose +- 0.01 (reset_model_init, dp_env_v3.py:158-164), N(0,1)
actions of the freshly initialised Gaussian policy clamped to the ctrlrange, reward 1.0 per step, done when the CoM
height leaves [0.7, 2.0].  tests/test_oracle_physics.py holds the same checks for the float64 oracle.
(The file name sorts last on purpose: the trained-policy tests were written after round 2's GPU minutes were spent,
so the driver's round-end run is their first on hardware; under `pytest -x` they must not mask the other files.)"""
import numpy as np
import pytest
import torch

import common

pytestmark = pytest.mark.gpu

DEVICE, N_ENVS = "cuda", 2048      # tools/dry_run_gpu_reference_tests.py overrides these (CPU stand-ins, fewer envs)


def test_fall_time_distribution_matches_reference_monitor_log():
    from scipy import stats
    from deepmimic_mujoco_b200.env import DPVecEnv
    ref = common.ref_fall_lengths(100)
    n = N_ENVS
    env = DPVecEnv(n, motions=("walk",), seed=21, reward_mode=0, reset_mode=1, auto_reset=True)
    env.reset()
    g = torch.Generator(device=DEVICE); g.manual_seed(3)
    first = torch.zeros(n, dtype=torch.int32, device=DEVICE)          # length of every env's FIRST episode (no
    for t in range(400):                                              # window bias towards short episodes)
        obs, rew, done, info = env.step(torch.randn(n, 28, device=DEVICE, generator=g))
        new = (done != 0) & (first == 0)
        first = torch.where(new, info["episode_length"].to(torch.int32), first)
        if t > 20 and bool((first > 0).all()):
            break
    lens = first.cpu().numpy().astype(np.float64)
    env.close()
    assert (lens > 0).all(), "some env never fell under N(0,1) torques"
    # oracle, same protocol, 2000 episodes: mean - reference mean = -0.8 (reference sample error 0.8), KS D = 0.08
    assert abs(lens.mean() - ref.mean()) < 3.0, (lens.mean(), ref.mean())
    assert 0.75 < lens.std() / ref.std() < 1.25, (lens.std(), ref.std())
    ks = stats.ks_2samp(lens, ref)
    assert ks.pvalue > 0.01, ks
    assert np.abs(np.percentile(lens, [25, 50, 75]) - np.percentile(ref, [25, 50, 75])).max() <= 3.0


def test_trained_policy_survival_matches_reference_monitor_log():
    """The policy the reference's TRPO run trained in MuJoCo 2.0 (its shipped TensorFlow checkpoint) keeps the MuJoCo
    humanoid up for ~290 steps (monitor rows around the save; a random policy: ~35).  The same weights, same protocol
    (reset_model_init, stochastic actions mean + exp(logstd) N(0,1), done on CoM height), on the CUDA path: first
    episode of 2048 envs.  Acceptance rule and the oracle's numbers (mean 279 +- 3, median 238): common.py,
    tests/test_oracle_physics.py::test_trained_policy_survival_matches_reference_monitor_log."""
    from deepmimic_mujoco_b200.env import DPVecEnv
    pol = common.RefTrainedPolicy()
    ref = pol.monitor_window(50)
    n, horizon = N_ENVS, 3000
    env = DPVecEnv(n, motions=("walk",), seed=22, reward_mode=0, reset_mode=1, auto_reset=True)
    obs = env.reset()
    g = torch.Generator(device=DEVICE); g.manual_seed(4)
    sd = torch.as_tensor(pol.act_std, dtype=torch.float32, device=DEVICE)
    first = torch.zeros(n, dtype=torch.int32, device=DEVICE)
    for t in range(horizon):
        act = pol.torch_mean_action(obs) + sd * torch.randn(n, 28, device=DEVICE, generator=g)
        obs, rew, done, info = env.step(act.contiguous())
        new = (done != 0) & (first == 0)
        first = torch.where(new, info["episode_length"].to(torch.int32), first)
        if t % 100 == 99 and bool((first > 0).all()):
            break
    still_up = int((first == 0).sum())
    first = torch.where(first == 0, torch.full_like(first, horizon), first)   # censored at the horizon (oracle: < 0.1 %)
    lens = first.cpu().numpy().astype(np.float64)
    env.close()
    assert still_up <= n // 100, f"{still_up} envs still up after {horizon} steps"
    assert lens.mean() > 5 * common.ref_fall_lengths(100).mean()
    assert common.trained_policy_verdict(lens, ref), (lens.mean(), np.percentile(lens, [25, 50, 75]), ref.mean(),
                                                      np.percentile(ref, [25, 50, 75]))


def test_reference_checkpoint_through_the_fused_policy_kernel_and_evaluate():
    """The same pin through the product's own rollout pieces: the checkpoint's tensors loaded into MlpPolicy
    (tf_checkpoint.policy_arrays -> load_arrays), actions sampled by the fused policy kernel (Philox noise), the
    batched evaluate task (rollout.evaluate = trpo.py:356-436) with the stochastic policy."""
    from deepmimic_mujoco_b200.env import DPVecEnv
    from deepmimic_mujoco_b200.policy import MlpPolicy
    from deepmimic_mujoco_b200.rollout import evaluate
    from deepmimic_mujoco_b200.tf_checkpoint import policy_arrays
    import os
    g = np.load(os.path.join(common.GOLDEN, "ref_trained_policy.npz"))
    ref = common.RefTrainedPolicy()
    n = N_ENVS
    env = DPVecEnv(n, motions=("walk",), seed=23, reward_mode=0, reset_mode=1, auto_reset=True)
    pi = MlpPolicy(seed=5)
    pi.load_arrays(policy_arrays({k: g[k] for k in g.files if k.startswith("pi/")}, "pi"))
    x = torch.randn(64, 56, device=DEVICE) * 0.5                           # the loaded network is the checkpoint's
    mean = torch.empty(64, 28, device=DEVICE)
    pi.act(False, x, out_mean=mean)
    assert np.abs(mean.cpu().numpy() - ref.mean_action(x.cpu().numpy())).max() < 1e-4
    out = evaluate(pi, env, horizon=3000, stochastic=True)
    lens = out["ep_len"].cpu().numpy().astype(np.float64)
    cut = int((~out["finished"]).sum())
    env.close()
    assert cut <= n // 100, f"{cut} envs still up after 3000 steps"
    assert np.array_equal(out["ep_ret"].cpu().numpy(), lens.astype(np.float32))     # reward 1.0 per step
    assert common.trained_policy_verdict(lens, ref.monitor_window(50)), (lens.mean(), np.percentile(lens, [25, 50, 75]))
