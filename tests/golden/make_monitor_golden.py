#!/usr/bin/env python3
"""Generate tests/golden/monitor_golden.json by running the REFERENCE episode monitor
(/root/reference/src/bench/monitor.py Monitor + ResultsWriter) here, around the scripted env of
tests/monitor_contract.py, driven by the loop body of trpo.py:47-80.  gym is absent in this container; the
minimal shim in tests/monitor_contract.py provides the two names monitor.py imports (gym, gym.core.Wrapper).
Everything else executed is the reference's own code.  Run in the build container only."""
import csv
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from monitor_contract import ScriptedEnv, drive_like_trpo, install_gym_shim  # noqa: E402

install_gym_shim()
sys.path.insert(0, "/root/reference/src")
from bench.monitor import Monitor  # noqa: E402  (the reference's class)

out = {}
for seed in (0, 1, 2):
    env = ScriptedEnv(seed)
    path = os.path.join(tempfile.mkdtemp(), "golden")
    mon = Monitor(env, path)
    err_step_before_reset = None
    try:
        mon.step(env.action_space.sample())
    except RuntimeError as e:
        err_step_before_reset = str(e)
    ep_rets, ep_lens, _ = drive_like_trpo(mon, 380)
    err_early_reset = None
    if not mon.needs_reset:
        try:
            mon.reset()
        except RuntimeError as e:
            err_early_reset = str(e)
    mon.close()
    with open(path + ".monitor.csv") as f:
        header = f.readline()
        rows = [{"r": float(r["r"]), "l": int(r["l"])} for r in csv.DictReader(f)]
    out[str(seed)] = dict(rows=rows, header_keys=sorted(json.loads(header[1:]).keys()), total_steps=mon.get_total_steps(),
                          episode_lengths=mon.get_episode_lengths(), nreset=env.nreset,
                          err_step_before_reset=err_step_before_reset, err_early_reset=err_early_reset,
                          loop_ep_lens=ep_lens, loop_ep_rets=ep_rets)
with open(os.path.join(HERE, "monitor_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print({k: (len(v["rows"]), v["total_steps"], v["nreset"]) for k, v in out.items()})
