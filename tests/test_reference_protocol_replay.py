"""The trained-policy pin through the reference's OWN code: trpo.py's traj_segment_generator
(source lines executed unchanged), dp_env_v3.DPEnv and bench.Monitor (imported), over the oracle behind a
mujoco-py-shaped adapter -- tools/reference_protocol_replay.py, run in a subprocess because it installs import shims
and changes the working directory.  Build container only (needs /root/reference)."""
import json
import os
import re
import subprocess
import sys

import pytest

import common

REF = "/root/reference/src/trpo.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
def test_reference_loop_env_and_monitor_over_the_oracle_reproduce_the_reference_log():
    tool = os.path.join(common.ROOT, "tools", "reference_protocol_replay.py")
    r = subprocess.run([sys.executable, tool, "150", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout
    m = re.search(r"oracle : mean\s+([0-9.]+)", out)
    assert m and 230.0 < float(m.group(1)) < 350.0, out                   # MuJoCo: 290 / 300; random policy: 35
    verdicts = re.findall(r"rule: (pass|FAIL)", out)
    assert verdicts == ["pass", "pass"], out                              # +-50 and +-100 monitor rows around the save


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
def test_reference_loop_with_a_fresh_policy_reproduces_the_first_episodes_of_the_log():
    """Section 2 (ix) the same way: a freshly initialised policy (normc init, logstd 0) through the reference's loop, env
    class and monitor over the oracle, against the first 100 monitor rows of the reference's run."""
    tool = os.path.join(common.ROOT, "tools", "reference_protocol_replay.py")
    r = subprocess.run([sys.executable, tool, "400", "3", "initial"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert re.findall(r"rule: (pass|FAIL)", r.stdout) == ["pass"], r.stdout


def _replay_updates(mode, iters, seed):
    tool = os.path.join(common.ROOT, "tools", "reference_training_replay.py")
    r = subprocess.run([sys.executable, tool, str(iters), str(seed), mode, "2"], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, DMB_REPLAY_UPDATES="1", OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stderr[-2000:]
    rows = re.findall(r"update expected ([0-9.]+) actual ([\-0-9.]+) meankl ([0-9.]+) stepsize ([0-9.]+)", r.stdout)
    assert len(rows) == 3 * iters, r.stdout
    return [tuple(float(x) for x in row) for row in rows]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
def test_trpo_learner_reproduces_the_statistics_the_reference_learner_logged():
    """deepmimic_mujoco_b200/trpo.py against numbers the reference's TensorFlow learner printed into
    its own log (tests/golden/ref_learner_log.json): fed by the reference's rollout loop and GAE over the oracle with
    the reference's two MPI workers emulated (tools/reference_training_replay.py), the expected improvement of the
    natural-gradient step, the surrogate gain of the accepted step, the KL of the accepted step and the number of
    step halvings must be the reference's -- for the freshly initialised policy (iterations 0-9 of the log: expected
    0.19-0.22, actual 0.19, KL 0.0070, no halving) and for the shipped checkpoint (iterations 1890-1941: expected
    0.34-0.45, exactly one halving in 156 of 156 updates, actual 0.18, KL 0.0049).  With one worker the expected
    improvement is sqrt(2) larger: that is how the two workers were found."""
    import numpy as np
    with open(os.path.join(common.GOLDEN, "ref_learner_log.json")) as f:
        ref = json.load(f)["iterations"]
    first = [u for k in range(10) for u in ref[str(k)]["updates"]]
    late = [u for k in range(1890, 1942) for u in ref[str(k)]["updates"]]
    assert all(u["halvings"] == 0 for u in first) and all(u["halvings"] == 1 for u in late)
    kl_first = np.mean([ref[str(k)]["meankl"] for k in range(10)])
    kl_late = np.mean([ref[str(k)]["meankl"] for k in range(1890, 1942)])

    mine = _replay_updates("fresh", 4, 0)
    e, a, kl, ss = (np.array(c) for c in zip(*mine))
    assert (ss == 1.0).all()
    assert abs(e.mean() / np.mean([u["expected"] for u in first]) - 1) < 0.10, e
    assert abs(a.mean() / np.mean([u["actual"] for u in first]) - 1) < 0.12, a
    assert abs(kl.mean() / kl_first - 1) < 0.10, kl

    mine = _replay_updates("pretrained", 4, 0)
    e, a, kl, ss = (np.array(c) for c in zip(*mine))
    assert (ss == 0.5).sum() >= 11                                        # full step violates 1.5 max_kl, half step passes
    assert abs(np.median(e) / np.median([u["expected"] for u in late]) - 1) < 0.12, e
    assert abs(np.median(a) / np.median([u["actual"] for u in late]) - 1) < 0.20, a
    assert abs(np.median(kl) / kl_late - 1) < 0.10, kl
