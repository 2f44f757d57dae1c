"""TRPO update restated in PyTorch (CPU tests incl. gloo world_size 2; one GPU test
ties the differentiable policy to the fused inference kernel)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common  # noqa: F401
from deepmimic_mujoco_b200 import trpo as T
from deepmimic_mujoco_b200.policy import RunningMeanStd


class TinyPolicy:
    def __init__(self, od=5, ad=2, hid=7, seed=0, device="cpu"):
        g = torch.Generator().manual_seed(seed)
        r = lambda *s: (torch.randn(*s, generator=g) * 0.3).to(device)
        self.params = dict(pw1=r(od, hid), pb1=r(hid), pw2=r(hid, hid), pb2=r(hid), pw3=r(hid, ad) * 0.1, pb3=r(ad) * 0.1,
                           logstd=r(ad) * 0.1, vw1=r(od, hid), vb1=r(hid), vw2=r(hid, hid), vb2=r(hid), vw3=r(hid, 1), vb3=r(1))
        self.ob_rms = RunningMeanStd((od,), device)


def test_cg_solves_spd_system():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(8, 8)); A = A @ A.T + 8 * np.eye(8); b = rng.normal(size=8)
    At, bt = torch.tensor(A), torch.tensor(b)
    x = T.cg(lambda p: At @ p, bt, cg_iters=8)
    assert np.abs(x.numpy() - np.linalg.solve(A, b)).max() < 1e-6   # stops at |r|^2 < 1e-10 like cg.py


def test_cg_and_explained_variance_match_the_reference_routines_run_here():
    """Golden vectors produced by the reference's own cg (cg.py:2-34, imported) and explained_variance
    (utils/math_util.py:25-38, its source lines executed unchanged): tests/golden/make_learner_golden.py."""
    g = np.load(os.path.join(common.GOLDEN, "learner_golden.npz"))
    A, b = torch.tensor(g["cg_A"]), torch.tensor(g["cg_b"])
    x10 = T.cg(lambda p: A @ p, b, cg_iters=10)                            # the reference's 10 iterations, unconverged
    assert np.abs(x10.numpy() - g["cg_x10"]).max() < 1e-12 * max(1.0, np.abs(g["cg_x10"]).max())
    x100 = T.cg(lambda p: A @ p, b, cg_iters=100)                          # early exit at residual_tol, as a mask here
    assert np.abs(x100.numpy() - g["cg_x100"]).max() < 1e-9
    ev = T.explained_variance(torch.tensor(g["ev_ypred"]), torch.tensor(g["ev_y"]))
    assert abs(float(ev) - float(g["ev"])) < 1e-12
    assert math.isnan(float(T.explained_variance(torch.tensor(g["ev_ypred"]), torch.ones(200, dtype=torch.float64))))
    assert math.isnan(float(g["ev_const"]))


def test_adam_matches_the_reference_mpi_adam_update_run_here():
    """Golden parameter trajectory produced by the reference's own MpiAdam.update (mpi_adam.py:21-35, source lines
    executed on a stub with a one-process communicator): tests/golden/make_learner_golden.py."""
    g = np.load(os.path.join(common.GOLDEN, "learner_golden.npz"))
    params = {"w": torch.tensor(g["adam_theta0"][:25].copy()), "b": torch.tensor(g["adam_theta0"][25:].copy())}
    opt = T.Adam(params, ["w", "b"])
    for k, gk in enumerate(g["adam_grads"]):
        opt.update(torch.tensor(gk), 1e-3)
        th = torch.cat([params["w"], params["b"]]).numpy()
        assert np.abs(th - g["adam_traj"][k]).max() < 2e-7, k


def test_diag_gaussian_formulas():
    rng = np.random.default_rng(1)
    m0, m1 = torch.tensor(rng.normal(size=(4, 3))), torch.tensor(rng.normal(size=(4, 3)))
    l0, l1 = torch.tensor(rng.normal(size=3) * 0.3), torch.tensor(rng.normal(size=3) * 0.3)
    x = torch.tensor(rng.normal(size=(4, 3)))
    s0, s1 = np.exp(l0.numpy()), np.exp(l1.numpy())
    ref_logp = -0.5 * (((x.numpy() - m0.numpy()) / s0) ** 2).sum(1) - 0.5 * math.log(2 * math.pi) * 3 - np.log(s0).sum()
    assert np.abs(T.gauss_logp(m0, l0, x).numpy() - ref_logp).max() < 1e-12
    ref_kl = (np.log(s1 / s0) + (s0 ** 2 + (m0.numpy() - m1.numpy()) ** 2) / (2 * s1 ** 2) - 0.5).sum(1)
    assert np.abs(T.gauss_kl(m0, l0, m1, l1).numpy() - ref_kl).max() < 1e-12
    assert np.allclose(T.gauss_kl(m0, l0, m0, l0).numpy(), 0)
    assert np.allclose(T.gauss_entropy(l0, (4,)).numpy(), (np.log(s0) + 0.5 * math.log(2 * math.pi * math.e)).sum())


def test_fisher_vector_product_matches_explicit_hessian():
    pi = TinyPolicy()
    for k in T.POL_KEYS:
        pi.params[k] = pi.params[k].double().requires_grad_(True)
    for k in T.VF_KEYS:
        pi.params[k] = pi.params[k].double()
    pi.ob_rms.mean = pi.ob_rms.mean.double(); pi.ob_rms.std = pi.ob_rms.std.double()
    ob = torch.randn(40, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    pol = [pi.params[k] for k in T.POL_KEYS]
    with torch.no_grad():
        m_old, ls_old, _ = T.policy_forward(pi.params, pi.ob_rms.mean, pi.ob_rms.std, ob)
        m_old, ls_old = m_old.clone(), ls_old.clone()
    def klfun():
        m, ls, _ = T.policy_forward(pi.params, pi.ob_rms.mean, pi.ob_rms.std, ob)
        return T.gauss_kl(m_old, ls_old, m, ls).mean()
    n = sum(p.numel() for p in pol)
    g = T.flat_grad(klfun(), pol, create_graph=True)
    H = torch.stack([T.flat_grad(g[i], pol, create_graph=True).detach() for i in range(n)])
    v = torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    hv = T.flat_grad((T.flat_grad(klfun(), pol, create_graph=True) * v).sum(), pol)
    assert torch.allclose(hv, H @ v, atol=1e-10)
    assert torch.linalg.eigvalsh(0.5 * (H + H.T)).min() > -1e-10     # Fisher matrix is PSD at the expansion point


def _toy_segment(pi, Tn=64, N=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    ob = torch.randn(Tn, N, 5, generator=g)
    with torch.no_grad():
        mean, logstd, vp = T.policy_forward(pi.params, pi.ob_rms.mean, pi.ob_rms.std, ob.reshape(-1, 5))
    ac = (mean + torch.exp(logstd.detach()) * torch.randn(mean.shape, generator=g)).reshape(Tn, N, 2)
    adv = (ac[..., 0] * ob[..., 0] - ac[..., 1] * ob[..., 1])          # reward correlated with obs -> learnable
    return dict(ob=ob, ac=ac, adv=adv, vpred=vp.reshape(Tn, N), tdlamret=adv + 0.1)


def test_trpo_update_respects_trust_region_and_improves():
    pi = TinyPolicy(seed=1)
    learner = T.TRPO(pi, vf_batch=64)
    seg = _toy_segment(pi)
    before = T.flat_params(pi.params, T.POL_KEYS).detach().clone()
    st = learner.update(seg)
    assert st["stepsize"] > 0 and st["meankl"] <= 0.01 * 1.5 + 1e-9
    assert st["surrgain"] >= -1e-9                                        # surrogate did not get worse
    assert (T.flat_params(pi.params, T.POL_KEYS).detach() - before).abs().max() > 0
    fixed = _toy_segment(pi, seed=7)
    v0 = learner.update(fixed)["vferr"]
    for _ in range(6):
        st = learner.update(fixed)
    assert st["vferr"] < v0                                               # value function fits fixed targets


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pi = TinyPolicy(seed=10 + rank)                # different init per rank: TRPO() broadcasts rank 0's (trpo.py:184)
    learner = T.TRPO(pi, vf_batch=64)
    st = learner.update(_toy_segment(pi, seed=100 + rank))   # different data per rank, averaged gradients
    flat = torch.cat([T.flat_params(pi.params, T.POL_KEYS), T.flat_params(pi.params, T.VF_KEYS)]).detach()
    q.put((rank, flat.numpy(), st["stepsize"]))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_update_keeps_replicas_in_sync():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs: p.join(timeout=60)
    assert np.allclose(res[0][1], res[1][1], atol=1e-6) and res[0][2] == res[1][2]   # trpo.py:285-287 desync check


@pytest.mark.gpu
def test_differentiable_policy_matches_fused_kernel_and_trains():
    from deepmimic_mujoco_b200.env import DPVecEnv
    from deepmimic_mujoco_b200.policy import MlpPolicy
    from deepmimic_mujoco_b200.rollout import SegmentGenerator, add_vtarg_and_adv
    env = DPVecEnv(512, motions=("walk",), seed=0, reward_mode=4)
    pi = MlpPolicy(seed=0)
    gen = SegmentGenerator(pi, env, horizon=16)
    learner = T.TRPO(pi, vf_batch=1024)
    seg = next(gen)
    ob = seg["ob"][3].contiguous()
    mean_k = torch.empty(512, 28, device="cuda")
    _, vp_k = pi.act(False, ob, out_mean=mean_k)
    with torch.no_grad():
        mean_t, _, vp_t = T.policy_forward(pi.params, pi.ob_rms.mean, pi.ob_rms.std, ob)
    assert (mean_k - mean_t).abs().max() < 2e-5 and (vp_k - vp_t).abs().max() < 2e-4 * max(1.0, float(vp_t.abs().max()))
    for it in range(2):
        add_vtarg_and_adv(seg, learner.gamma, learner.lam)
        st = learner.update(seg)
        assert all(math.isfinite(v) for v in st.values()) and st["meankl"] <= 0.015 + 1e-6
        seg = next(gen)
    env.close()
