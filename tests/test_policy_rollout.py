"""batched policy inference + rollout/GAE.  GAE is checked on CPU against a literal
restatement of trpo.py:83-94; the fused CUDA kernel against a float64 numpy MlpPolicy (gpu marker)."""
import ctypes as C

import numpy as np
import pytest
import torch

import common


def gae_reference(rew, vpred, new, nextvpred, gamma, lam):
    """trpo.py:83-94 for one env (numpy loops)."""
    new = np.append(new, 0); vpred = np.append(vpred, nextvpred)
    T = len(rew); adv = np.empty(T, "float32"); last = 0
    for t in reversed(range(T)):
        nonterminal = 1 - new[t + 1]
        delta = rew[t] + gamma * vpred[t + 1] * nonterminal - vpred[t]
        adv[t] = last = delta + gamma * lam * nonterminal * last
    return adv, adv + vpred[:-1]


def test_gae_matches_reference_formula():
    from deepmimic_mujoco_b200.rollout import add_vtarg_and_adv
    rng = np.random.default_rng(0)
    T, N = 37, 5
    rew = rng.normal(size=(T, N)).astype(np.float32); vp = rng.normal(size=(T, N)).astype(np.float32)
    new = (rng.uniform(size=(T, N)) < 0.15).astype(np.float32); nxt = rng.normal(size=N).astype(np.float32)
    seg = dict(rew=torch.tensor(rew), vpred=torch.tensor(vp), new=torch.tensor(new), nextvpred=torch.tensor(nxt))
    add_vtarg_and_adv(seg, 0.995, 0.97)
    for i in range(N):
        adv, ret = gae_reference(rew[:, i], vp[:, i], new[:, i], nxt[i], 0.995, 0.97)
        assert np.abs(seg["adv"][:, i].numpy() - adv).max() < 1e-5
        assert np.abs(seg["tdlamret"][:, i].numpy() - ret).max() < 1e-5


def test_gae_matches_the_reference_function_run_here():
    """Golden advantages / TD(lambda) returns produced by the reference's own add_vtarg_and_adv (trpo.py:83-94, its
    source lines executed unchanged: tests/golden/make_learner_golden.py), one env per column."""
    import os
    from deepmimic_mujoco_b200.rollout import add_vtarg_and_adv
    g = np.load(os.path.join(common.GOLDEN, "learner_golden.npz"))
    seg = dict(rew=torch.tensor(g["gae_rew"]), vpred=torch.tensor(g["gae_vpred"]),
               new=torch.tensor(g["gae_new"].astype(np.float32)), nextvpred=torch.tensor(g["gae_nextvpred"]))
    add_vtarg_and_adv(seg, float(g["gae_gamma"]), float(g["gae_lam"]))
    assert np.abs(seg["adv"].numpy() - g["gae_adv"]).max() < 2e-6 * np.abs(g["gae_adv"]).max()
    assert np.abs(seg["tdlamret"].numpy() - g["gae_tdlamret"]).max() < 2e-6 * np.abs(g["gae_tdlamret"]).max()


def test_running_mean_std_semantics():
    pytest.importorskip("torch")
    if not torch.cuda.is_available():
        dev = "cpu"
    else:
        dev = "cuda"
    from deepmimic_mujoco_b200.policy import RunningMeanStd
    r = RunningMeanStd((3,), dev)
    assert torch.allclose(r.mean.cpu(), torch.zeros(3)) and torch.allclose(r.std.cpu(), torch.ones(3))   # eps/eps = 1
    x = torch.tensor(np.random.default_rng(0).normal(2.0, 3.0, size=(1000, 3)), dtype=torch.float32, device=dev)
    r.update(x)
    xs = x.double().cpu().numpy()
    mean = xs.sum(0) / (1000 + 1e-2)
    std = np.sqrt(np.maximum((np.square(xs).sum(0) + 1e-2) / (1000 + 1e-2) - mean ** 2, 1e-2))
    assert np.abs(r.mean.cpu().numpy() - mean).max() < 1e-5 and np.abs(r.std.cpu().numpy() - std).max() < 1e-5


@pytest.mark.gpu
def test_policy_kernel_matches_numpy():
    import oracle.pyoracle as po
    from deepmimic_mujoco_b200.policy import MlpPolicy
    pi = MlpPolicy(seed=3)
    rng = np.random.default_rng(1)
    pi.params["logstd"].copy_(torch.tensor(rng.uniform(-1, 0.3, 28), dtype=torch.float32))
    for k in ("pb1", "pb2", "pb3", "vb1", "vb2", "vb3"):
        pi.params[k].copy_(torch.tensor(rng.normal(size=pi.params[k].shape) * 0.1, dtype=torch.float32))
    n = 333
    ob = torch.tensor(rng.normal(size=(n, 56)) * 3, dtype=torch.float32, device="cuda")
    pi.ob_rms.update(ob)
    mean = torch.empty(n, 28, device="cuda")
    ac, vp = pi.act(True, ob, out_mean=mean, first_row=7)
    P = {k: v.double().cpu().numpy() for k, v in pi.params.items()}
    x = np.clip((ob.double().cpu().numpy() - pi.ob_rms.mean.double().cpu().numpy()) / pi.ob_rms.std.double().cpu().numpy(), -5, 5)
    hv = np.tanh(np.tanh(x @ P["vw1"] + P["vb1"]) @ P["vw2"] + P["vb2"])
    vref = (hv @ P["vw3"] + P["vb3"])[:, 0]
    hp = np.tanh(np.tanh(x @ P["pw1"] + P["pb1"]) @ P["pw2"] + P["pb2"])
    mref = hp @ P["pw3"] + P["pb3"]
    assert np.abs(vp.cpu().numpy() - vref).max() < 2e-5 * max(1, np.abs(vref).max())
    assert np.abs(mean.cpu().numpy() - mref).max() < 2e-5
    # noise: Philox(seed, row + first_row, step, unit/2) + Box-Muller, reproduced on the host
    L = po.lib(); out = (C.c_uint32 * 4)()
    z = np.zeros((n, 28))
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    def philox(k0, k1, c):
        c = list(c)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xffffffff]
            k0 = (k0 + W0) & 0xffffffff; k1 = (k1 + W1) & 0xffffffff
        return c
    for r_ in range(0, n, 37):
        for u in range(28):
            c = philox(3, 0, [r_ + 7, 0, u >> 1, 0x504f4c])
            u1 = ((c[(u & 1) * 2] >> 8) + 0.5) / 16777216.0; u2 = (c[(u & 1) * 2 + 1] >> 8) / 16777216.0
            z[r_, u] = np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2)
        ref = mref[r_] + np.exp(P["logstd"]) * z[r_]
        assert np.abs(ac[r_].cpu().numpy() - ref).max() < 1e-4
    # deterministic mode returns the mean; statistics of the noise are standard normal
    ac2, _ = pi.act(False, ob)
    assert torch.equal(ac2, mean)
    zz = ((ac - mean) / torch.exp(pi.params["logstd"])).cpu().numpy()
    assert abs(zz.mean()) < 0.05 and abs(zz.std() - 1) < 0.05


@pytest.mark.gpu
def test_gae_kernel_matches_reference_formula():
    """dmb_gae (one thread per env) vs the numpy restatement of trpo.py:83-94, and vs the CPU torch loop."""
    from deepmimic_mujoco_b200.rollout import add_vtarg_and_adv
    rng = np.random.default_rng(1)
    T, N = 33, 301
    rew = rng.normal(size=(T, N)).astype(np.float32); vp = rng.normal(size=(T, N)).astype(np.float32)
    new = (rng.uniform(size=(T, N)) < 0.15).astype(np.float32); nxt = rng.normal(size=N).astype(np.float32)
    seg = {k: torch.tensor(v).cuda() for k, v in dict(rew=rew, vpred=vp, new=new, nextvpred=nxt).items()}
    add_vtarg_and_adv(seg, 0.995, 0.97)
    cpu = dict(rew=torch.tensor(rew), vpred=torch.tensor(vp), new=torch.tensor(new), nextvpred=torch.tensor(nxt))
    add_vtarg_and_adv(cpu, 0.995, 0.97)
    assert (seg["adv"].cpu() - cpu["adv"]).abs().max() < 1e-5 and (seg["tdlamret"].cpu() - cpu["tdlamret"]).abs().max() < 1e-5
    for i in (0, 7, 300):
        adv, ret = gae_reference(rew[:, i], vp[:, i], new[:, i], nxt[i], 0.995, 0.97)
        assert np.abs(seg["adv"][:, i].cpu().numpy() - adv).max() < 1e-5
        assert np.abs(seg["tdlamret"][:, i].cpu().numpy() - ret).max() < 1e-5


@pytest.mark.gpu
def test_segment_generator_shapes_and_bookkeeping():
    from deepmimic_mujoco_b200.env import DPVecEnv
    from deepmimic_mujoco_b200.policy import MlpPolicy
    from deepmimic_mujoco_b200.rollout import SegmentGenerator, add_vtarg_and_adv
    env = DPVecEnv(256, motions=("walk",), seed=0, reward_mode=4)
    pi = MlpPolicy(seed=0)
    gen = SegmentGenerator(pi, env, horizon=32)
    seg = next(gen)
    assert seg["ob"].shape == (32, 256, 56) and seg["ac"].shape == (32, 256, 28) and seg["new"][0].all()
    assert torch.isfinite(seg["ob"]).all() and torch.isfinite(seg["vpred"]).all()
    add_vtarg_and_adv(seg, 0.995, 0.97)
    assert seg["adv"].shape == (32, 256) and torch.isfinite(seg["adv"]).all()
    assert len(seg["ep_rets"]) == int(seg["new"][1:].sum() + (gen.cur_new).sum())   # one record per finished episode
    pi.ob_rms.update(seg["ob"])
    seg2 = next(gen)
    assert not seg2["new"][0].all()
    env.close()


def test_evaluate_counts_first_episodes_like_traj_1_generator():
    """rollout.evaluate (the batched runner / traj_1_generator of trpo.py:356-436) on stub env + policy objects (CPU):
    every env's FIRST episode is counted, later auto-reset episodes are ignored, the horizon cut happens after
    horizon + 1 steps as in the reference (`if new or t >= horizon: break` follows the step)."""
    from deepmimic_mujoco_b200.rollout import evaluate

    class Env:
        def __init__(self, ends):
            self.ends, self.t = torch.tensor(ends), 0          # env i is done at steps ends[i], 2 ends[i], ...
        def reset(self):
            self.t = 0
            return torch.zeros(len(self.ends), 4)
        def step(self, ac):
            self.t += 1
            done = ((self.t % self.ends) == 0).to(torch.uint8)
            return torch.full((len(self.ends), 4), float(self.t)), torch.full((len(self.ends),), 0.5), done, {}

    class Pi:
        calls = 0
        def act(self, stochastic, ob):
            Pi.calls += 1
            return torch.zeros(ob.shape[0], 2), torch.zeros(ob.shape[0])

    out = evaluate(Pi(), Env([3, 10, 64, 500]), horizon=99, stochastic=True)
    assert out["ep_len"].tolist() == [3, 10, 64, 100]              # the last env is cut after horizon + 1 steps
    assert out["finished"].tolist() == [True, True, True, False]
    assert torch.allclose(out["ep_ret"], torch.tensor([1.5, 5.0, 32.0, 50.0]))
    assert abs(float(out["avg_len"]) - 177 / 4) < 1e-12 and Pi.calls == 100
    Pi.calls = 0
    out = evaluate(Pi(), Env([3, 10, 20]), horizon=1000)
    assert out["ep_len"].tolist() == [3, 10, 20] and Pi.calls == 64  # stops at the first completion check
