"""The trained-policy pin through the reference's OWN code: trpo.py's traj_segment_generator
(source lines executed unchanged), dp_env_v3.DPEnv and bench.Monitor (imported), over the oracle behind a
mujoco-py-shaped adapter -- tools/reference_protocol_replay.py, run in a subprocess because it installs import shims
and changes the working directory.  Build container only (needs /root/reference)."""
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
