"""The CPU-side analysis tools under tools/ keep running (tiny arguments, subprocesses): their full outputs are the
profiles DESIGN.md cites (profiles/r2_box_contact_stats.txt, r2_fall_time_sensitivity.txt)."""
import os
import subprocess
import sys

import common


def _run(tool, *args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(common.ROOT, "tools", tool), *args], capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_box_contact_stats_tool():
    out = _run("box_contact_stats.py", "2", "60")
    assert out.count("env steps with an own-algorithm contact active") == 4          # three clips + the trained policy
    assert "box - plane" in out


def test_dry_run_of_the_gpu_reference_log_tests():
    """tests/test_z_gpu_reference_log.py against oracle-backed stand-ins (the first, cheap test only here; the tool's
    default runs all three)."""
    src = ("import sys; sys.argv = ['x', '96']; sys.path.insert(0, %r); import dry_run_gpu_reference_tests as d; "
           "d.envmod.DPVecEnv, d.polmod.MlpPolicy = d.OracleVecEnv, d.TorchPolicy; "
           "import test_z_gpu_reference_log as t; t.DEVICE, t.N_ENVS = 'cpu', 96; "
           "t.test_fall_time_distribution_matches_reference_monitor_log(); print('ok')") % os.path.join(common.ROOT, "tools")
    r = subprocess.run([sys.executable, "-c", src], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
