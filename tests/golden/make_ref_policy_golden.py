#!/usr/bin/env python3
"""Generate tests/golden/ref_trained_policy.npz from the TRAINED policy the reference ships:
/root/reference/src/checkpoint_tmp/DeepMimic/trpo-walk-0/DeepMimic/trpo-walk-0.{index,data-00000-of-00001}, the
TensorFlow-1 checkpoint its TRPO run wrote after 1.0 M MuJoCo steps (trpo.py:220-224), together with the tail of the
episode monitor of that same run (src/log_tmp/DeepMimic/trpo-walk-0/monitor.json.monitor.csv).

The checkpoint holds a policy that MuJoCo 2.0 itself shaped: under the reference's protocol (trpo.py:27-80) it keeps
the humanoid up for ~270 steps where a random policy falls after ~35.  How long the SAME weights keep the humanoid up in
another implementation of the dynamics is therefore a pin of that implementation against MuJoCo
(tests/test_oracle_physics.py::test_trained_policy_survival_matches_reference_monitor_log, tests/test_z_gpu_reference_log.py).

TensorFlow is not installed here, so the V2 checkpoint ("tensor bundle") is read directly by
deepmimic_mujoco_b200/tf_checkpoint.py (LevelDB-format index table of BundleEntryProto values + raw tensor file).
Network (mlp_policy_trpo.py:24-60): obz = clip((ob - mean) / std, -5, 5) with the RunningMeanStd of
utils/misc_util.py (sum, sumsq, count; std = sqrt(max(sumsq/count - mean^2, 1e-2))), two tanh layers of 100, a linear
head of 28, and a state-independent logstd; the action is mean + exp(logstd) * N(0, 1).
Run in the build container only."""
import csv
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepmimic_mujoco_b200.tf_checkpoint import read_checkpoint  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = "/root/reference/src/checkpoint_tmp/DeepMimic/trpo-walk-0/DeepMimic/trpo-walk-0"
LOG = "/root/reference/src/log_tmp/DeepMimic/trpo-walk-0"


def main():
    t = read_checkpoint(CKPT)
    total = sum(a.nbytes for a in t.values())
    assert total == os.path.getsize(CKPT + ".data-00000-of-00001") == 278264
    pi = {k[3:]: v for k, v in t.items() if k.startswith("pi/")}
    assert pi["polfc1/w"].shape == (56, 100) and pi["polfc2/w"].shape == (100, 100) and pi["polfinal/w"].shape == (100, 28)
    assert pi["logstd"].shape == (1, 28) and pi["obfilter/runningsum"].shape == (56,)
    # "oldpi/*" is the policy one TRPO update earlier (assign_old_eq_new, trpo.py:248): close to, not equal to, pi
    assert max(float(np.abs(v - t["oldpi/" + k]).max()) for k, v in pi.items() if not k.startswith("obfilter")) < 0.05
    with open(os.path.join(LOG, "monitor.json.monitor.csv")) as f:
        f.readline()
        rows = list(csv.DictReader(f))
    lens = np.asarray([int(r["l"]) for r in rows], dtype=np.int32)
    t_mon = np.asarray([float(r["t"]) for r in rows])
    with open(os.path.join(LOG, "progress.csv")) as f:
        prog = list(csv.DictReader(f))
    # WHEN was the checkpoint written?  save_per_iter = 100 (trpo.py:514) and 1942 iterations ran, so the file on disk
    # is the save at the top of iteration 1900 (trpo.py:220-224), i.e. at TimeElapsed of iteration 1899; the monitor's
    # clock started <1 s before the learner's.  Episodes before that index were played by slightly older policies,
    # episodes after it by slightly newer ones (max_kl 0.01 per update; 2-3 episodes per update).
    assert len(prog) == 1942
    k_ckpt = int(np.searchsorted(t_mon, float(prog[1899]["TimeElapsed"])))
    out = {"pi/" + k: v for k, v in pi.items()}                  # keys = the TensorFlow variable names
    out["monitor_last_lengths"] = lens[-400:]                    # the episodes of the last ~110 k steps of training
    out["checkpoint_index"] = np.int32(k_ckpt - (len(lens) - 400))  # position of the save inside monitor_last_lengths
    out["progress_last_eplenmean"] = np.asarray([float(p["EpLenMean"]) for p in prog[-50:]])
    out["progress_last_entropy"] = np.asarray([float(p["entropy"]) for p in prog[-50:]])
    np.savez_compressed(os.path.join(HERE, "ref_trained_policy.npz"), **out)
    ent = float(np.sum(pi["logstd"] + 0.5 * np.log(2 * np.pi * np.e)))
    # the policy's entropy is logged after every update: the checkpoint's logstd gives EXACTLY the value logged by
    # update 1899 (35.71932) -- the weights are the policy that played the episodes from k_ckpt on
    assert abs(ent - float(prog[1899]["entropy"])) < 2e-5, (ent, prog[1899]["entropy"])
    print("tensors", sorted(t), "\nentropy from logstd", ent, "logged by update 1899", prog[1899]["entropy"],
          "\nobfilter count", float(pi["obfilter/count"]), "checkpoint at monitor episode", k_ckpt, "of", len(lens),
          "; mean length +-50 / +-100 episodes around it", lens[k_ckpt - 50: k_ckpt + 50].mean(),
          lens[k_ckpt - 100: k_ckpt + 100].mean())


if __name__ == "__main__":
    main()
