#!/usr/bin/env python3
"""(build container, CPU) The trained-policy pin of DESIGN.md section 2 (x) WITHOUT restating the protocol: the
reference's own rollout loop (``traj_segment_generator``, /root/reference/src/trpo.py:27-80, its source lines cut out
with ``ast`` and executed unchanged -- trpo.py itself imports TensorFlow), its own env class (dp_env_v3.DPEnv, imported)
wrapped in its own episode monitor (bench/monitor.py Monitor, imported), driven by the policy its TRPO run trained in
MuJoCo (tests/golden/ref_trained_policy.npz, evaluated with numpy).  Only the simulator underneath is this repo's:
the mujoco-py-shaped adapter of tests/golden/make_env_logic_golden.py forwards ``sim.step()`` / ``sim.forward()`` /
``sim.reset()`` to the float64 oracle.  The monitor file this writes is the same artefact as the reference's
src/log_tmp/DeepMimic/trpo-walk-0/monitor.json.monitor.csv; its episode lengths are compared with the reference's rows
around the checkpoint by the rule of tests/common.py::trained_policy_verdict, and with the restated protocol of
tests/test_oracle_physics.py (which skips the reset() before reset_model_init(), the constructor's probe step and the
float32 actions -- none of which should matter, and this run shows whether they do).
With ``initial`` as third argument the policy is a freshly initialised one instead (normc(1.0) hidden layers, normc(0.01)
head, logstd 0, empty observation filter: mlp_policy_trpo.py:24-60, utils/tf_util.py normc_initializer) and the
comparison is with the FIRST 100 monitor rows of the reference's run -- the time-to-fall pin of section 2 (ix).
usage: python tools/reference_protocol_replay.py [episodes=300] [seed=0] [trained|initial]"""
import ast
import csv
import os
import random
import sys
import tempfile

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import common  # noqa: E402
from make_env_logic_golden import install_shims  # noqa: E402


class InitialPolicy:
    """The reference's policy network at initialisation (weights drawn here with numpy: the same distribution as, not
    the same numbers as, TensorFlow's)."""

    def __init__(self, seed):
        rng = np.random.default_rng(1000 + seed)

        def normc(shape, std):
            w = rng.normal(size=shape)
            return w * std / np.sqrt(np.square(w).sum(axis=0, keepdims=True))
        self.layers = [(normc((56, 100), 1.0), np.zeros(100)), (normc((100, 100), 1.0), np.zeros(100)),
                       (normc((100, 28), 0.01), np.zeros(28))]
        self.act_std = np.ones(28)

    def mean_action(self, ob):
        h = np.clip(np.asarray(ob, dtype=np.float64), -5.0, 5.0)          # empty RunningMeanStd: mean 0, std 1
        for i, (w, b) in enumerate(self.layers):
            h = h @ w + b
            if i < 2:
                h = np.tanh(h)
        return h


class NumpyPi:
    """pi.act(stochastic, ob) -> (ac float32 [28], vpred) as mlp_policy_trpo.MlpPolicy.act returns them."""

    def __init__(self, seed, which="trained"):
        self.p = common.RefTrainedPolicy() if which == "trained" else InitialPolicy(seed)
        self.rng = np.random.default_rng(seed)

    def act(self, stochastic, ob):
        mean = self.p.mean_action(np.asarray(ob)[None])[0]
        ac = mean + self.p.act_std * self.rng.normal(size=mean.size) if stochastic else mean
        return ac.astype(np.float32), 0.0


def main():
    episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    which = sys.argv[3] if len(sys.argv) > 3 else "trained"
    import warnings
    warnings.simplefilter("ignore")
    install_shims()
    os.chdir(REF_SRC)
    sys.path.insert(0, REF_SRC)
    from config import Config
    Config.mocap_path = "%s%s/humanoid3d_walk.txt" % (Config.curr_path, Config.motion_folder)     # the run is "trpo-walk-0"
    import dp_env_v3
    from bench.monitor import Monitor
    src = open(os.path.join(REF_SRC, "trpo.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "traj_segment_generator")
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "trpo.py", "exec"), ns)
    random.seed(seed)
    path = os.path.join(tempfile.mkdtemp(), "monitor.json")
    env = Monitor(dp_env_v3.DPEnv(), path)                                  # trpo.py:459-460
    env.seed(seed)                                                          # trpo.py:461
    gen = ns["traj_segment_generator"](NumpyPi(seed, which), env, None, 256, stochastic=True)   # trpo.py:193, 350
    while len(env.get_episode_lengths()) < episodes + 1:
        next(gen)
    env.close()
    with open(path + ".monitor.csv") as f:
        f.readline()
        lens = np.asarray([int(r["l"]) for r in csv.DictReader(f)], dtype=np.float64)
    first, lens = lens[0], lens[1: episodes + 1]      # episode 0 starts from the RSI pose of reset() (trpo.py:32)
    if which != "trained":
        ref = common.ref_fall_lengths(100)
        ks = stats.ks_2samp(lens, ref)
        dq = np.abs(np.percentile(lens, [25, 50, 75]) - np.percentile(ref, [25, 50, 75])).max()
        ok = abs(lens.mean() - ref.mean()) < 3.0 and 0.75 < lens.std() / ref.std() < 1.25 and ks.pvalue > 0.01 and dq <= 3.0
        print(f"{len(lens)} episodes of a freshly initialised policy through the reference's own loop + env class + monitor "
              f"over the oracle (seed {seed})")
        print(f"  oracle : mean {lens.mean():6.2f}  sd {lens.std():5.2f}  quartiles {np.percentile(lens, [25, 50, 75])}")
        print(f"  MuJoCo, first 100 monitor rows: mean {ref.mean():6.2f}  sd {ref.std():5.2f}  quartiles "
              f"{np.percentile(ref, [25, 50, 75])}  KS D {ks.statistic:.3f} p {ks.pvalue:.2f}  rule: {'pass' if ok else 'FAIL'}")
        return
    pol = common.RefTrainedPolicy()
    print(f"{len(lens)} episodes through the reference's own loop + env class + monitor over the oracle (seed {seed}); "
          f"first episode (from the mocap RSI pose): {int(first)} steps")
    print(f"  oracle : mean {lens.mean():6.1f}  sd {lens.std():6.1f}  quartiles {np.percentile(lens, [25, 50, 75])}")
    for w in (50, 100):
        ref = pol.monitor_window(w)
        ks = stats.ks_2samp(lens, ref)
        print(f"  MuJoCo, {2 * w} monitor rows around the checkpoint: mean {ref.mean():6.1f}  sd {ref.std():6.1f}  quartiles "
              f"{np.percentile(ref, [25, 50, 75])}  KS D {ks.statistic:.3f} p {ks.pvalue:.2f}  "
              f"rule: {'pass' if common.trained_policy_verdict(lens, ref) else 'FAIL'}")


if __name__ == "__main__":
    main()
