#!/usr/bin/env python3
"""Generate tests/golden/ref_episode_lengths.json from the one artefact of the reference that was PRODUCED BY MuJoCo:
the episode monitor of its own training run (/root/reference/src/log_tmp/DeepMimic/trpo-walk-0/
monitor.json.monitor.csv, rows (r, l, t) written by bench/monitor.py around dp_env_v3.DPEnv; reward is 1.0 per step,
so r == l) and the per-iteration EpLenMean of progress.csv.  Protocol that produced the rows (trpo.py:27-80): the
first episode starts from reset() (mocap RSI); every later one from reset_model_init() (standing pose, U(-0.01, 0.01)
noise on qpos and qvel, dp_env_v3.py:158-164); actions are the freshly initialised Gaussian policy's (mean ~ 0 from
the normc(0.01) output layer, logstd = 0, mlp_policy_trpo.py:14-74) -- i.e. N(0, 1) per actuator, clamped to the
ctrlrange by MuJoCo; an episode ends when the CoM height leaves [0.7, 2.0] (dp_env_v3.py:134-139).  TRPO (max_kl 0.01
per batch) moves the policy slowly: the first 100 episodes (~3500 steps) are within sampling noise of the initial
policy (mean 35.9 / 34.9 / 35.1 over the first 15 / 30 / 100); later ones get longer as the policy learns to stay up
(35.9 over 200, 37.6 over 400).
Run in the build container only."""
import csv
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LOG = "/root/reference/src/log_tmp/DeepMimic/trpo-walk-0"

with open(os.path.join(LOG, "monitor.json.monitor.csv")) as f:
    header = f.readline().strip()
    rows = list(csv.DictReader(f))
lens = [int(r["l"]) for r in rows]
assert all(float(r["r"]) == float(r["l"]) for r in rows[:400])      # reward 1.0 per step (dp_env_v3.py:117)
with open(os.path.join(LOG, "progress.csv")) as f:
    prog = list(csv.DictReader(f))
out = {
    "source": "src/log_tmp/DeepMimic/trpo-walk-0/monitor.json.monitor.csv (+ progress.csv)",
    "header": header,
    "first_episode_from_rsi": lens[0],
    "lengths_from_standing_pose": lens[1:401],
    "progress_eplenmean": [float(p["EpLenMean"]) for p in prog[:8]],
    "progress_timesteps": [int(float(p["TimestepsSoFar"])) for p in prog[:8]],
}
with open(os.path.join(HERE, "ref_episode_lengths.json"), "w") as f:
    json.dump(out, f)
print("episodes", len(lens), "first 100 mean", sum(lens[1:101]) / 100.0, out["progress_eplenmean"][:3])
