#!/usr/bin/env python3
"""Generate tests/golden/ref_learner_log.json from the reference's own training log
(/root/reference/src/log_tmp/DeepMimic/trpo-walk-0/log.txt): for every TRPO update of iterations 0-9 (the freshly
initialised policy) and 1890-1941 (around the shipped checkpoint, written at the top of iteration 1900) the "Expected"
improvement g . fullstep and the "Actual" surrogate gain of the accepted step, the number of step halvings of the line
search (trpo.py:262-281 messages), and the iteration's logged meankl.  These are numbers the reference's TensorFlow
learner produced; tests/test_reference_protocol_replay.py holds this repo's learner (deepmimic_mujoco_b200/trpo.py)
to them.  Run in the build container only."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
log = open("/root/reference/src/log_tmp/DeepMimic/trpo-walk-0/log.txt").read()
parts = re.split(r"\*+ Iteration (\d+) \*+", log)
out = {}
for i in range(1, len(parts), 2):
    n, body = int(parts[i]), parts[i + 1]
    if not (n < 10 or 1890 <= n <= 1941):
        continue
    ups, cur = [], []
    for line in body.splitlines():
        m = re.match(r"Expected: ([\-0-9.]+) Actual: ([\-0-9.]+)", line)
        if m:
            cur.append((float(m.group(1)), float(m.group(2))))
        elif "Stepsize OK" in line or "couldn't compute" in line:
            ups.append({"expected": cur[0][0], "actual": cur[-1][1], "halvings": len(cur) - 1})
            cur = []
    mk = re.search(r"meankl\s+\|\s+([0-9.e\-]+)", body)
    out[str(n)] = {"updates": ups, "meankl": float(mk.group(1))}
with open(os.path.join(HERE, "ref_learner_log.json"), "w") as f:
    json.dump({"source": "src/log_tmp/DeepMimic/trpo-walk-0/log.txt (line-search messages of trpo.py:262-281 and the meankl "
                         "row), iterations 0-9 and 1890-1941", "iterations": out}, f)
print(len(out), "iterations;", out["0"], out["1900"])
