"""Shared helpers for the test-suite: model/mocap loading, oracle wrappers, state generators."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from deepmimic_mujoco_b200.mjcf import load_tables  # noqa: E402
from deepmimic_mujoco_b200.mocap import concat_clips, load_clip  # noqa: E402
from deepmimic_mujoco_b200.model_blob import default_config, pack_model  # noqa: E402

ASSETS = os.path.join(ROOT, "deepmimic_mujoco_b200", "assets")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_XML = "/root/reference/src/mujoco/humanoid_deepmimic/envs/asset/dp_env_v3.xml"

_tables = None


def tables():
    global _tables
    if _tables is None:
        _tables = load_tables(os.path.join(ASSETS, "dp_env_v3.model.npz"))
    return _tables


def model(max_con=16, max_efc=40):
    return pack_model(tables(), max_con=max_con, max_efc=max_efc)


def clip(name):
    return load_clip(os.path.join(ASSETS, "motions", name + ".npz"), name=name)


def f32(x):
    """Round to fp32-representable float64 (so oracle and GPU start from identical inputs)."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def random_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def airborne_states(rng, n, z=3.0, frac=0.45, vel=2.0):
    """Random poses inside the joint ranges, high above the floor."""
    mt = tables()
    lo, hi = mt.jnt_range[1:, 0], mt.jnt_range[1:, 1]
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
    qpos = np.tile(mt.qpos0, (n, 1))
    qvel = rng.normal(size=(n, mt.nv)) * vel
    for i in range(n):
        qpos[i, 7:] = mid + half * rng.uniform(-frac, frac, size=mt.nq - 7)
        qpos[i, 3:7] = random_quat(rng)
        qpos[i, 2] = z
    return f32(qpos), f32(qvel)


def standing_states(rng, n, noise=0.05, vel=0.3, drop=0.03):
    """Near the default standing pose, feet at / slightly into the floor."""
    mt = tables()
    qpos = np.tile(mt.qpos0, (n, 1))
    qpos[:, 7:] += rng.uniform(-noise, noise, size=(n, mt.nq - 7))
    qpos[:, 2] -= rng.uniform(0.0, drop, size=n)  # root height 0.9 puts the soles ~2 cm above the floor
    for i in range(n):
        q = np.array([1.0, 0, 0, 0]) + rng.normal(size=4) * noise * 0.5
        qpos[i, 3:7] = q / np.linalg.norm(q)
    qvel = rng.normal(size=(n, mt.nv)) * vel
    return f32(qpos), f32(qvel)


def mocap_states(name, idx):
    c = clip(name)
    return f32(c.data_config[idx]), f32(np.nan_to_num(c.data_vel[idx]))


def rollout_states(rng, n, steps_range=(5, 40)):
    """States reached by the ORACLE under random torques from the standing pose (contact-rich, falling)."""
    import oracle.pyoracle as po
    mt = tables()
    o = po.Oracle(model())
    qs, vs, ws = [], [], []
    for i in range(n):
        qpos = mt.qpos0 + rng.uniform(-0.01, 0.01, size=mt.nq)
        qvel = rng.uniform(-0.01, 0.01, size=mt.nv)
        o.set_state(qpos, qvel)
        for _ in range(int(rng.integers(*steps_range))):
            o.d.arr("ctrl")[: mt.nu] = rng.uniform(-0.5, 0.5, size=mt.nu)
            o.step()
        qs.append(o.qpos.copy()); vs.append(o.qvel.copy()); ws.append(o.d.arr("qacc_warmstart")[: mt.nv].copy())
    return f32(np.array(qs)), f32(np.array(vs)), f32(np.array(ws))


# ---- measured-error bookkeeping: the parity tests record the worst error they saw per quantity; conftest.py writes
# the table to gpurun_out/parity_measured.json at the end of a GPU session (evidence for DESIGN.md section 2)
MEASURED = {}


def record(label, key, value):
    d = MEASURED.setdefault(label, {})
    d[key] = max(float(value), d.get(key, 0.0))


def ref_fall_lengths(k=100):
    """Episode lengths of the reference's own MuJoCo run under its freshly initialised policy
    (tests/golden/make_ref_episode_golden.py): the first k episodes that start from the standing pose."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_episode_lengths.json")) as f:
        g = json.load(f)
    return np.asarray(g["lengths_from_standing_pose"][:k], dtype=np.float64)


class RefTrainedPolicy:
    """The policy the reference's own TRPO run trained IN MuJoCo 2.0, read from the TensorFlow checkpoint the reference
    ships (tests/golden/make_ref_policy_golden.py; network of mlp_policy_trpo.py:24-60): obz = clip((ob - mean) / std,
    -5, 5), two tanh layers of 100, a linear head of 28, action = mean + exp(logstd) * N(0, 1).  ``monitor_window(w)``
    = the episode lengths the reference's monitor recorded for the w episodes before and the w after the moment the
    checkpoint was written (older / newer policies within TRPO's max_kl 0.01 per update of the saved one)."""

    def __init__(self):
        g = np.load(os.path.join(GOLDEN, "ref_trained_policy.npz"))
        cnt = g["pi/obfilter/count"]
        self.ob_mean = (g["pi/obfilter/runningsum"] / cnt).astype(np.float32).astype(np.float64)
        var = (g["pi/obfilter/runningsumsq"] / cnt).astype(np.float32).astype(np.float64) - self.ob_mean ** 2
        self.ob_std = np.sqrt(np.maximum(var, 1e-2))                      # utils/misc_util.py:53-54
        self.layers = [(g[f"pi/{n}/w"].astype(np.float64), g[f"pi/{n}/b"].astype(np.float64))
                       for n in ("polfc1", "polfc2", "polfinal")]
        self.act_std = np.exp(g["pi/logstd"][0].astype(np.float64))
        self._lens = g["monitor_last_lengths"].astype(np.float64)
        self._k = int(g["checkpoint_index"])

    def mean_action(self, ob):
        """ob [n, 56] -> mean action [n, 28] (float64 numpy)."""
        h = np.clip((np.asarray(ob, dtype=np.float64) - self.ob_mean) / self.ob_std, -5.0, 5.0)
        for i, (w, b) in enumerate(self.layers):
            h = h @ w + b
            if i < 2:
                h = np.tanh(h)
        return h

    def torch_mean_action(self, ob):
        """The same network on a torch tensor ob [n, 56] (any device, float32) -> mean action [n, 28]."""
        import torch
        c = lambda a: torch.as_tensor(a, dtype=torch.float32, device=ob.device)
        h = torch.clamp((ob - c(self.ob_mean)) / c(self.ob_std), -5.0, 5.0)
        for i, (w, b) in enumerate(self.layers):
            h = h @ c(w) + c(b)
            if i < 2:
                h = torch.tanh(h)
        return h

    def monitor_window(self, w=50):
        return self._lens[max(0, self._k - w): self._k + w].copy()


def trained_policy_verdict(lens, ref):
    """The acceptance rule of the trained-policy pin, shared by the oracle test, the GPU test
    and tools/trained_policy_sensitivity.py: survival time of the reference's MuJoCo-trained policy, this
    implementation (``lens``) vs the reference's own monitor rows (``ref``).  The distribution is heavy-tailed
    (roughly geometric beyond ~100 steps), so the rule uses the mean, the median, the lower quartile and a two-sample
    Kolmogorov-Smirnov test; 100 reference episodes put ~6 % standard error on the reference mean and the policy
    itself drifts by a few % across the window."""
    from scipy import stats
    lens, ref = np.asarray(lens, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return bool(abs(lens.mean() / ref.mean() - 1.0) < 0.2
                and abs(np.median(lens) / np.median(ref) - 1.0) < 0.2
                and abs(np.percentile(lens, 25) / np.percentile(ref, 25) - 1.0) < 0.2
                and stats.ks_2samp(lens, ref).pvalue > 0.01)
