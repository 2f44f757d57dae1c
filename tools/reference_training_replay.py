#!/usr/bin/env python3
"""(build container, CPU) The reference's TRAINING RUN replayed: does the learner of this repo (trpo.py: the SURVEY 8(f)
rank-2 row), fed by the reference's own rollout code over this repo's physics, learn as fast as the reference's
TensorFlow learner did over MuJoCo?  The learning curve the reference logged (src/log_tmp/DeepMimic/trpo-walk-0/
progress.csv: EpLenMean 37 -> ~270 and policy entropy 39.7 -> 35.7 over 1942 iterations = 1.49 M env steps) is the
target.

What is the reference's, unchanged: ``traj_segment_generator`` and ``add_vtarg_and_adv`` (trpo.py:27-94, their source lines
executed as they are), the env class dp_env_v3.DPEnv and its episode monitor bench.Monitor (imported), the protocol and
hyper-parameters of ``train`` (trpo.py:338-354: 256 steps per segment, g_step 3, max_kl 0.01, 10 CG iterations,
damping 0.1, gamma 0.995, lambda 0.97, 3 value epochs of minibatch 128 at step size 1e-3), the reporting (mean of the
last 40 episodes).  What is this repo's: the simulator under the env (the float64 oracle behind the mujoco-py-shaped
adapter of tests/golden/make_env_logic_golden.py), the policy / value networks and the TRPO update
(deepmimic_mujoco_b200/trpo.py on CPU tensors; weights initialised with the reference's normc scheme).
The reference's run used TWO MPI workers, which its own logs show although no command line is recorded: progress.csv
counts 8918 episodes where rank 0's monitor file has 13 274 (= 2 workers x the last of 3 segments against 1 worker x 3
segments: ratio 2/3), TimestepsSoFar grows by ~512 = 2 x 256 per iteration while the monitor sees 768 = 3 x 256 env
steps, and the logged "Expected" improvement of the first updates (0.19) is 1/sqrt(2) of what one worker's 256-sample
gradient gives (0.25-0.27).  The replay therefore runs two envs (seeds s and s + 10000, trpo.py:341) whose segments
are concatenated for every update: equal-sized workers make the averaged gradient / Fisher-vector product / losses of
trpo.py:174-179 the full-batch ones, per-worker advantage standardisation is kept, and the value-function minibatches
become 256 (128 per worker, gradients averaged).  ``workers=1`` reproduces the single-process reading of the code.
One run is one seed of a noisy process, as the reference's log is; the printed table puts both next to each other.
With ``pretrained`` as third argument the run starts from the reference's shipped checkpoint (policy, value network and
observation filter after update 1899) instead of a fresh initialisation: the learner then has to keep, and keep improving,
a policy that MuJoCo shaped (the reference's own last 42 iterations took EpLenMean from 260 to 269).
usage: python tools/reference_training_replay.py [iterations=1942] [seed=0] [fresh|pretrained] [workers=2]"""
import ast
import csv
import os
import random
import sys
import tempfile
import time
from collections import deque

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_env_logic_golden import install_shims  # noqa: E402
from deepmimic_mujoco_b200.policy import RunningMeanStd  # noqa: E402
from deepmimic_mujoco_b200.trpo import TRPO  # noqa: E402


class CpuPolicy:
    """mlp_policy_trpo.MlpPolicy on CPU tensors: ``params`` / ``ob_rms`` for the learner, ``act`` for the rollout."""

    def __init__(self, seed):
        g = torch.Generator().manual_seed(seed)

        def normc(i, o, std):                                               # utils/tf_util.py normc_initializer
            w = torch.randn(i, o, generator=g)
            return w * std / torch.sqrt((w * w).sum(dim=0, keepdim=True))
        z = torch.zeros
        self.params = dict(vw1=normc(56, 100, 1.0), vb1=z(100), vw2=normc(100, 100, 1.0), vb2=z(100), vw3=normc(100, 1, 1.0),
                           vb3=z(1), pw1=normc(56, 100, 1.0), pb1=z(100), pw2=normc(100, 100, 1.0), pb2=z(100),
                           pw3=normc(100, 28, 0.01), pb3=z(28), logstd=z(28))
        self.ob_rms = RunningMeanStd((56,), "cpu")
        self.rng = np.random.default_rng(seed)
        self.refresh()

    def refresh(self):
        """numpy copies for the per-step act() (the learner owns the torch tensors)."""
        self.n = {k: v.detach().numpy().astype(np.float64) for k, v in self.params.items()}
        self.mean, self.std = self.ob_rms.mean.numpy().astype(np.float64), self.ob_rms.std.numpy().astype(np.float64)

    def act(self, stochastic, ob):
        n = self.n
        x = np.clip((ob - self.mean) / self.std, -5.0, 5.0)
        v = np.tanh(np.tanh(x @ n["vw1"] + n["vb1"]) @ n["vw2"] + n["vb2"]) @ n["vw3"] + n["vb3"]
        m = np.tanh(np.tanh(x @ n["pw1"] + n["pb1"]) @ n["pw2"] + n["pb2"]) @ n["pw3"] + n["pb3"]
        ac = m + np.exp(n["logstd"]) * self.rng.normal(size=28) if stochastic else m
        return ac.astype(np.float32), float(v[0])


def cut(path, name, ns):
    node = next(n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name == name)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def obfilter_report(pi, iters):
    """The observation filter is a by-product that MuJoCo shaped too: the reference's checkpoint holds the sum / sum of
    squares of every observation its two workers fed the filter during 1900 iterations -- per-joint mean angle, angle
    spread and joint-velocity spread of the whole training run.  Print this run's next to it."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_trained_policy.npz"))
    cnt = g["pi/obfilter/count"]
    mr = g["pi/obfilter/runningsum"] / cnt
    sr = np.sqrt(np.maximum(g["pi/obfilter/runningsumsq"] / cnt - mr * mr, 0))
    c = pi.ob_rms.count.numpy()
    m = pi.ob_rms.sum.numpy() / c
    sd = np.sqrt(np.maximum(pi.ob_rms.sumsq.numpy() / c - m * m, 0))
    import common
    names = common.tables().joint_names[1:]
    print(f"observation filter after {iters} iterations (count {float(c):.0f}) | the reference's checkpoint, 1900 iterations "
          f"(count {float(cnt):.0f})")
    print(f"{'joint':18s} {'mean angle':>17s} {'sd angle':>15s} {'sd velocity':>15s}   (this run | reference)")
    for j in range(28):
        print(f"{names[j]:18s} {m[j]:+8.3f} |{mr[j]:+7.3f} {sd[j]:7.3f} |{sr[j]:6.3f} {sd[28 + j]:7.3f} |{sr[28 + j]:6.3f}")
    ra, rv = sd[:28] / sr[:28], sd[28:] / sr[28:]
    print(f"ratio this run / reference over the 28 joints: sd angle median {np.median(ra):.2f} ({ra.min():.2f} .. {ra.max():.2f}), "
          f"sd velocity median {np.median(rv):.2f} ({rv.min():.2f} .. {rv.max():.2f})")


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1942
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    import warnings
    warnings.simplefilter("ignore")
    install_shims()
    os.chdir(REF_SRC)
    sys.path.insert(0, REF_SRC)
    from config import Config
    Config.mocap_path = "%s%s/humanoid3d_walk.txt" % (Config.curr_path, Config.motion_folder)
    import dp_env_v3
    from bench.monitor import Monitor
    ns = {"np": np}
    seg_gen_fn = cut(os.path.join(REF_SRC, "trpo.py"), "traj_segment_generator", ns)
    gae = cut(os.path.join(REF_SRC, "trpo.py"), "add_vtarg_and_adv", ns)
    with open(os.path.join(REF_SRC, "log_tmp/DeepMimic/trpo-walk-0/progress.csv")) as f:
        ref = list(csv.DictReader(f))
    workers = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    random.seed(seed); torch.manual_seed(seed)
    envs = []
    for w in range(workers):
        e = Monitor(dp_env_v3.DPEnv(), os.path.join(tempfile.mkdtemp(), "monitor.json"))
        e.seed(seed + 10000 * w)                                            # workerseed, trpo.py:341
        envs.append(e)
    env = envs[0]                                                           # rank 0: the one whose monitor file is kept
    pi = CpuPolicy(seed)
    start_iter = 0
    if len(sys.argv) > 3 and sys.argv[3] == "pretrained":
        from deepmimic_mujoco_b200.tf_checkpoint import policy_arrays
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_trained_policy.npz"))
        a = policy_arrays({k: g[k] for k in g.files if k.startswith("pi/")}, "pi")
        for k, t in pi.params.items():
            t.copy_(torch.as_tensor(a[k]))
        pi.ob_rms.load(a["ob_sum"], a["ob_sumsq"], a["ob_count"])
        pi.refresh()
        start_iter = 1900
    learner = TRPO(pi, max_kl=0.01, cg_iters=10, cg_damping=0.1, gamma=0.995, lam=0.97, vf_iters=3, vf_stepsize=1e-3,
                   vf_batch=128 * workers)
    gens = [seg_gen_fn(pi, e, None, 256, stochastic=True) for e in envs]
    lenbuf, t0, steps = deque(maxlen=40), time.time(), 0
    print(f"seed {seed}, {workers} worker(s); columns: this run | the reference's log (same iteration)")
    print(f"{'iter':>5s} {'env steps':>10s} | {'EpLenMean':>9s} {'entropy':>8s} {'meankl':>8s} {'ev_tdlam':>8s} | "
          f"{'EpLenMean':>9s} {'entropy':>8s} {'meankl':>8s} {'ev_tdlam':>8s}", flush=True)
    for it in range(iters):
        for _ in range(3):                                                  # g_step (trpo.py:232)
            segs = [next(g) for g in gens]
            for seg in segs:
                gae(seg, 0.995, 0.97)
                seg["adv"] = (seg["adv"] - seg["adv"].mean()) / seg["adv"].std()     # per worker, trpo.py:240
            steps += len(segs[0]["rew"])                                    # rank 0's env steps, as its monitor counts
            t = lambda k: torch.as_tensor(np.concatenate([np.asarray(sg[k]) for sg in segs]), dtype=torch.float32).unsqueeze(1)
            st = learner.update({k: t(k) for k in ("ob", "ac", "adv", "tdlamret", "vpred")})
            pi.refresh()
            if os.environ.get("DMB_REPLAY_UPDATES"):                        # one line per TRPO update (tests)
                print(f"update expected {st.get('expectedimprove', float('nan')):.4f} actual {st['surrgain']:.4f} "
                      f"meankl {st['meankl']:.6f} stepsize {st['stepsize']:.3f}", flush=True)
        for seg in segs:
            lenbuf.extend(seg["ep_lens"])                                   # trpo.py:300-304: every worker's last segment
        if it % (10 if start_iter else 50) == 0 or it == iters - 1:
            r = ref[min(it + start_iter, len(ref) - 1)]
            print(f"{it:5d} {steps:10d} | {np.mean(lenbuf) if lenbuf else float('nan'):9.1f} {st['entropy']:8.3f} "
                  f"{st['meankl']:8.5f} {st.get('ev_tdlam_before', float('nan')):8.3f} | {float(r['EpLenMean']):9.1f} "
                  f"{float(r['entropy']):8.3f} {float(r['meankl']):8.5f} {float(r['ev_tdlam_before']):8.3f}", flush=True)
    if os.environ.get("DMB_REPLAY_SAVE"):                                   # the trained policy, in the reference's format
        from deepmimic_mujoco_b200.tf_checkpoint import policy_tensors, write_checkpoint
        arrays = {k: v.detach().numpy() for k, v in pi.params.items()}
        arrays.update(ob_sum=pi.ob_rms.sum.numpy(), ob_sumsq=pi.ob_rms.sumsq.numpy(), ob_count=pi.ob_rms.count.numpy())
        write_checkpoint(os.environ["DMB_REPLAY_SAVE"], {**policy_tensors(arrays, "pi"), **policy_tensors(arrays, "oldpi")})
    if not start_iter:
        obfilter_report(pi, iters)
    lens = np.asarray(env.get_episode_lengths(), dtype=np.float64)
    env.close()
    print(f"{len(lens)} episodes, {int(lens.sum())} env steps in {time.time() - t0:.0f} s; mean length of the last 200 episodes "
          f"{lens[-200:].mean():.1f} (reference run: 13274 episodes, 1491336 steps, last 200: 296.2)")


if __name__ == "__main__":
    main()
