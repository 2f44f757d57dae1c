"""The episode-monitor contract the drop-in env has to live under:
/root/reference/src/bench/monitor.py:12-92 wrapped around the env, driven by the loop body of
/root/reference/src/trpo.py:47-80.  tests/monitor_contract.py restates the Monitor (the reference does not travel
to the GPU box); here it is pinned against rows written by the reference's own class
(tests/golden/monitor_golden.json) and, when /root/reference is present, against that class directly."""
import csv
import json
import os
import sys

import pytest

import common
from monitor_contract import MonitorContract, ScriptedEnv, drive_like_trpo, install_gym_shim

GOLD = json.load(open(os.path.join(common.GOLDEN, "monitor_golden.json")))


def run_contract(seed, tmp_path):
    env = ScriptedEnv(seed)
    path = str(tmp_path / f"m{seed}")
    mon = MonitorContract(env, path)
    with pytest.raises(RuntimeError) as e1:
        mon.step(env.action_space.sample())
    ep_rets, ep_lens, _ = drive_like_trpo(mon, 380)
    err2 = None
    if not mon.needs_reset:
        with pytest.raises(RuntimeError) as e2:
            mon.reset()
        err2 = str(e2.value)
    mon.close()
    with open(path + ".monitor.csv") as f:
        header = f.readline()
        rows = [{"r": float(r["r"]), "l": int(r["l"])} for r in csv.DictReader(f)]
    return dict(rows=rows, header_keys=sorted(json.loads(header[1:]).keys()), total_steps=mon.get_total_steps(),
                episode_lengths=mon.get_episode_lengths(), nreset=env.nreset, err_step_before_reset=str(e1.value),
                err_early_reset=err2, loop_ep_lens=ep_lens, loop_ep_rets=ep_rets)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_restated_monitor_matches_reference_golden(seed, tmp_path):
    got, want = run_contract(seed, tmp_path), GOLD[str(seed)]
    assert got == want


@pytest.mark.skipif(not os.path.exists("/root/reference/src/bench/monitor.py"), reason="reference not on this box")
def test_restated_monitor_matches_reference_class(tmp_path):
    install_gym_shim()
    # load the reference file under its own module name (this repo's bench.py may already be imported as `bench`)
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_bench_monitor", "/root/reference/src/bench/monitor.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Monitor = mod.Monitor
    for seed in (5, 6):
        a, b = ScriptedEnv(seed), ScriptedEnv(seed)
        ref, mine = Monitor(a, str(tmp_path / f"ref{seed}")), MonitorContract(b, str(tmp_path / f"mine{seed}"))
        ra, rb = drive_like_trpo(ref, 390), drive_like_trpo(mine, 390)
        assert ra[0] == rb[0] and ra[1] == rb[1]
        assert ref.get_episode_lengths() == mine.get_episode_lengths() and ref.get_total_steps() == mine.get_total_steps()
        assert ref.get_episode_rewards() == mine.get_episode_rewards() and ref.needs_reset == mine.needs_reset
        ref.close(); mine.close()
        la = open(str(tmp_path / f"ref{seed}.monitor.csv")).read().splitlines()[1:]
        lb = open(str(tmp_path / f"mine{seed}.monitor.csv")).read().splitlines()[1:]
        strip_t = lambda ls: [",".join(x.split(",")[:2]) for x in ls]
        assert strip_t(la) == strip_t(lb)
