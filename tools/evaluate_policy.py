#!/usr/bin/env python3
"""(GPU box) The reference's evaluate task on the CUDA path:
    python trpo.py --task evaluate --load_model_path <ckpt> [--stochastic_policy]      (/root/reference/src/trpo.py:477-485)
becomes
    python tools/evaluate_policy.py --load_model_path <ckpt> [--stochastic_policy] [--number_trajs 4096]
The TensorFlow checkpoint is read without TensorFlow (deepmimic_mujoco_b200/tf_checkpoint.py) into the fused policy
kernel; every env plays one trajectory from reset_model_init (rollout.evaluate = the batched runner /
traj_1_generator, horizon 1024 as in trpo.py:484) and the two numbers the reference prints are printed.  Without
--load_model_path the policy the reference ships (its 1.0 M-step walk run, tests/golden/ref_trained_policy.npz) is
used; the reference's monitor recorded ~290 steps per episode for it in MuJoCo with the stochastic policy.
NOT YET RUN ON HARDWARE (written after this round's GPU minutes were spent; dry-run on the CPU with the stand-ins of
tools/dry_run_gpu_reference_tests.py: stochastic 285 steps, deterministic cut at horizon + 1); its pieces are covered by
tests/test_tf_checkpoint.py, tests/test_policy_rollout.py (CPU) and the GPU tests of MlpPolicy.act / DPVecEnv.step."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv  # noqa: E402
from deepmimic_mujoco_b200.policy import MlpPolicy  # noqa: E402
from deepmimic_mujoco_b200.rollout import evaluate  # noqa: E402
from deepmimic_mujoco_b200.tf_checkpoint import policy_arrays  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--load_model_path", default=None, help="TensorFlow checkpoint prefix saved by the reference")
    ap.add_argument("--stochastic_policy", action="store_true")
    ap.add_argument("--number_trajs", type=int, default=4096, help="envs = trajectories (the reference plays 100, one by one)")
    ap.add_argument("--horizon", type=int, default=1024)
    ap.add_argument("--motion", default="walk")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    env = DPVecEnv(a.number_trajs, motions=(a.motion,), seed=a.seed, reward_mode=0, reset_mode=1, auto_reset=True)
    pi = MlpPolicy(obs_dim=env.sim.obs_dim, act_dim=env.sim.nu, hid=100, device=env.sim.device, seed=a.seed)
    if a.load_model_path:
        pi.load_tf_checkpoint(a.load_model_path, "pi")
    else:
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_trained_policy.npz"))
        pi.load_arrays(policy_arrays({k: g[k] for k in g.files if k.startswith("pi/")}, "pi"))
    torch.cuda.synchronize()
    t0 = time.time()
    out = evaluate(pi, env, horizon=a.horizon, stochastic=a.stochastic_policy)
    torch.cuda.synchronize()
    dt = time.time() - t0
    lens = out["ep_len"].cpu().numpy()
    print("stochastic policy:" if a.stochastic_policy else "deterministic policy:")
    print("Average length:", float(out["avg_len"]))
    print("Average return:", float(out["avg_ret"]))
    print(f"[{a.number_trajs} trajectories, {int(lens.sum())} env steps counted, {dt:.2f} s; quartiles "
          f"{np.percentile(lens, [25, 50, 75])}, cut by the horizon: {int((~out['finished']).sum())}]")
    env.close()


if __name__ == "__main__":
    main()
