"""TRPO update + value-function fit on device-resident rollouts, gradients averaged over ranks.

SURVEY.md section 8(f) rank 2.  Restates the learner side of /root/reference/src/trpo.py:97-319 in
PyTorch (autograd instead of the TF1 graph, ``torch.distributed`` all-reduce over NCCL/NVLink -- gloo in
the CPU tests -- instead of the mpi4py ``Allreduce`` calls at trpo.py:178, mpi_adam.py:26):

  * surrogate / KL / entropy of the diagonal Gaussian (distributions.py:220-245),
  * conjugate gradient (cg.py:2-34) on Fisher-vector products (trpo.py:150-163, 228-229; every 5th sample),
  * step scaling by sqrt(shs / max_kl) and the 10-step backtracking line search (trpo.py:255-284),
  * value function: vf_iters epochs of minibatch-128 Adam on (vpred - tdlamret)^2 (trpo.py:288-295,
    mpi_adam.py:6-50).
Hyper-parameters default to trpo.py:350-353.  The policy used for gradients (:func:`policy_forward`) is the
same arithmetic as the fused inference kernel (``dmb_policy_act``); a GPU test checks they agree.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence

import torch
import torch.distributed as dist

POL_KEYS = ("pw1", "pb1", "pw2", "pb2", "pw3", "pb3", "logstd")
VF_KEYS = ("vw1", "vb1", "vw2", "vb2", "vw3", "vb3")


# ---------------------------------------------------------------------------------------------
def allmean(x: torch.Tensor, group=None) -> torch.Tensor:
    """Average over ranks (trpo.py:174-179); identity when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        x = x.clone()
        dist.all_reduce(x, group=group)
        x /= dist.get_world_size(group)
    return x


def policy_forward(params: Dict[str, torch.Tensor], ob_mean: torch.Tensor, ob_std: torch.Tensor, ob: torch.Tensor):
    """(mean [B,A], logstd [A], vpred [B]) -- mlp_policy_trpo.py:33-47 in torch ops (differentiable)."""
    obz = torch.clamp((ob - ob_mean) / ob_std, -5.0, 5.0)
    hv = torch.tanh(torch.tanh(obz @ params["vw1"] + params["vb1"]) @ params["vw2"] + params["vb2"])
    vpred = (hv @ params["vw3"] + params["vb3"])[:, 0]
    hp = torch.tanh(torch.tanh(obz @ params["pw1"] + params["pb1"]) @ params["pw2"] + params["pb2"])
    mean = hp @ params["pw3"] + params["pb3"]
    return mean, params["logstd"], vpred


def gauss_logp(mean, logstd, x):
    """-neglogp of DiagGaussianPd (distributions.py:231-234)."""
    std = torch.exp(logstd)
    return -(0.5 * (((x - mean) / std) ** 2).sum(-1) + 0.5 * math.log(2.0 * math.pi) * x.shape[-1] + logstd.sum(-1))


def gauss_kl(mean0, logstd0, mean1, logstd1):
    """KL(p0 || p1) of diagonal Gaussians (distributions.py:235-238)."""
    std0, std1 = torch.exp(logstd0), torch.exp(logstd1)
    return (logstd1 - logstd0 + (std0 ** 2 + (mean0 - mean1) ** 2) / (2.0 * std1 ** 2) - 0.5).sum(-1)


def gauss_entropy(logstd, batch_shape):
    """distributions.py:239-240."""
    return (logstd + 0.5 * math.log(2.0 * math.pi * math.e)).sum(-1).expand(batch_shape)


def flat_params(params: Dict[str, torch.Tensor], keys: Sequence[str]) -> torch.Tensor:
    return torch.cat([params[k].reshape(-1) for k in keys])


def set_flat_params(params: Dict[str, torch.Tensor], keys: Sequence[str], flat: torch.Tensor) -> None:
    off = 0
    with torch.no_grad():
        for k in keys:
            n = params[k].numel()
            params[k].copy_(flat[off:off + n].view_as(params[k]))
            off += n


def flat_grad(y: torch.Tensor, xs: List[torch.Tensor], create_graph: bool = False) -> torch.Tensor:
    gs = torch.autograd.grad(y, xs, create_graph=create_graph, retain_graph=create_graph, allow_unused=True)
    return torch.cat([(g if g is not None else torch.zeros_like(x)).reshape(-1) for g, x in zip(gs, xs)])


def cg(f_Ax: Callable[[torch.Tensor], torch.Tensor], b: torch.Tensor, cg_iters: int = 10,
       residual_tol: float = 1e-10) -> torch.Tensor:
    """Conjugate gradient, cg.py:2-34 (Demmel p. 312).  The early exit of cg.py:31-32 (``rdotr < residual_tol``)
    is applied as a device-side mask instead of a Python ``break``: once the residual is below the tolerance the
    iterate stops changing, and no iteration waits for a device-to-host copy of ``rdotr``."""
    p, r, x = b.clone(), b.clone(), torch.zeros_like(b)
    rdotr = r.dot(r)
    live = (rdotr >= residual_tol).to(b.dtype)
    for _ in range(cg_iters):
        z = f_Ax(p)
        v = live * rdotr / torch.where(live > 0, p.dot(z), torch.ones_like(rdotr))
        x += v * p
        r -= v * z
        newrdotr = r.dot(r)
        p = torch.where(live > 0, r + (newrdotr / torch.where(live > 0, rdotr, torch.ones_like(rdotr))) * p, p)
        rdotr = newrdotr
        live = live * (rdotr >= residual_tol).to(b.dtype)
    return x


def explained_variance(ypred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """utils/math_util.py:25-38: 1 - Var[y - ypred] / Var[y] (population variance), NaN when Var[y] == 0."""
    vary = y.var(unbiased=False)
    return torch.where(vary == 0, torch.full_like(vary, float("nan")), 1.0 - (y - ypred).var(unbiased=False) / vary)


class Adam:
    """mpi_adam.py:6-50 on a flat parameter view: the gradient is averaged over ranks, then plain Adam."""

    def __init__(self, params: Dict[str, torch.Tensor], keys: Sequence[str], beta1=0.9, beta2=0.999, epsilon=1e-8, group=None):
        self.params, self.keys, self.b1, self.b2, self.eps, self.group = params, list(keys), beta1, beta2, epsilon, group
        n = sum(params[k].numel() for k in keys)
        dev = params[keys[0]].device
        self.m, self.v, self.t = torch.zeros(n, device=dev), torch.zeros(n, device=dev), 0

    def update(self, localg: torch.Tensor, stepsize: float) -> None:
        g = allmean(localg, self.group)
        self.t += 1
        a = stepsize * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * (g * g)
        step = -a * self.m / (torch.sqrt(self.v) + self.eps)
        set_flat_params(self.params, self.keys, flat_params(self.params, self.keys).detach() + step)


# ---------------------------------------------------------------------------------------------
class TRPO:
    """One learner per rank; ``pi`` is a :class:`deepmimic_mujoco_b200.policy.MlpPolicy` (or any object with
    ``params`` and ``ob_rms``)."""

    def __init__(self, pi, max_kl=0.01, cg_iters=10, cg_damping=0.1, gamma=0.995, lam=0.97, vf_iters=3,
                 vf_stepsize=1e-3, entcoeff=0.0, vf_batch=128, group=None):
        self.pi, self.group = pi, group
        self.max_kl, self.cg_iters, self.cg_damping = max_kl, cg_iters, cg_damping
        self.gamma, self.lam, self.vf_iters, self.vf_stepsize, self.entcoeff, self.vf_batch = gamma, lam, vf_iters, vf_stepsize, entcoeff, vf_batch
        for k in POL_KEYS + VF_KEYS:
            pi.params[k].requires_grad_(True)
        self.vfadam = Adam(pi.params, VF_KEYS, group=group)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:   # trpo.py:184-186
            for k in POL_KEYS + VF_KEYS:
                dist.broadcast(pi.params[k].data, src=0, group=group)

    # losses of trpo.py:120-134 for the current parameters against detached old distribution parameters
    def _losses(self, ob, ac, atarg, old_mean, old_logstd):
        mean, logstd, _ = policy_forward(self.pi.params, self.pi.ob_rms.mean, self.pi.ob_rms.std, ob)
        kl = gauss_kl(old_mean, old_logstd, mean, logstd).mean()
        ent = gauss_entropy(logstd, mean.shape[:1]).mean()
        ratio = torch.exp(gauss_logp(mean, logstd, ac) - gauss_logp(old_mean, old_logstd, ac))
        surr = (ratio * atarg).mean()
        entbonus = self.entcoeff * ent
        return surr + entbonus, kl, entbonus, surr, ent

    def update(self, seg: Dict[str, torch.Tensor]) -> Dict[str, float]:
        """seg: output of rollout.SegmentGenerator after rollout.add_vtarg_and_adv ([T, N, ...] tensors)."""
        P = self.pi.params
        ob = seg["ob"].detach().reshape(-1, seg["ob"].shape[-1])
        ac = seg["ac"].detach().reshape(-1, seg["ac"].shape[-1])
        atarg = seg["adv"].detach().reshape(-1)
        tdlamret = seg["tdlamret"].detach().reshape(-1)
        atarg = (atarg - atarg.mean()) / atarg.std(unbiased=False)        # trpo.py:240 (numpy std: population)
        self.pi.ob_rms.update(ob, group=self.group)                       # trpo.py:242
        pol = [P[k] for k in POL_KEYS]
        with torch.no_grad():                                             # assign_old_eq_new (trpo.py:247)
            old_mean, old_logstd, _ = policy_forward(P, self.pi.ob_rms.mean, self.pi.ob_rms.std, ob)
            old_mean, old_logstd = old_mean.clone(), old_logstd.clone()
        losses = self._losses(ob, ac, atarg, old_mean, old_logstd)
        g = allmean(flat_grad(losses[0], pol), self.group)
        lossbefore = allmean(torch.stack([l.detach() for l in losses]), self.group)
        stats = {"optimgain": float(lossbefore[0]), "meankl": 0.0, "surrgain": float(lossbefore[3]), "entropy": float(lossbefore[4]),
                 "stepsize": 0.0}
        if not torch.allclose(g, torch.zeros_like(g)):
            fob, fold_mean = ob[::5], old_mean[::5]                       # fvpargs = every 5th sample (trpo.py:245)

            def fisher_vector_product(p: torch.Tensor) -> torch.Tensor:
                mean, logstd, _ = policy_forward(P, self.pi.ob_rms.mean, self.pi.ob_rms.std, fob)
                kl = gauss_kl(fold_mean, old_logstd, mean, logstd).mean()
                klg = flat_grad(kl, pol, create_graph=True)
                return allmean(flat_grad((klg * p).sum(), pol), self.group) + self.cg_damping * p

            stepdir = cg(fisher_vector_product, g, cg_iters=self.cg_iters)
            assert torch.isfinite(stepdir).all()
            shs = 0.5 * stepdir.dot(fisher_vector_product(stepdir))
            lm = torch.sqrt(shs / self.max_kl)
            fullstep = stepdir / lm
            expectedimprove = float(g.dot(fullstep))
            surrbefore = float(lossbefore[0])
            thbefore = flat_params(P, POL_KEYS).detach().clone()
            stepsize, ok = 1.0, False
            for _ in range(10):
                set_flat_params(P, POL_KEYS, thbefore + fullstep * stepsize)
                with torch.no_grad():
                    ml = allmean(torch.stack(list(self._losses(ob, ac, atarg, old_mean, old_logstd))), self.group)
                surr, kl = float(ml[0]), float(ml[1])
                if torch.isfinite(ml).all() and kl <= self.max_kl * 1.5 and surr - surrbefore >= 0:
                    ok = True
                    break
                stepsize *= 0.5
            if not ok:
                set_flat_params(P, POL_KEYS, thbefore)
                stepsize = 0.0
            stats.update(optimgain=float(ml[0]), meankl=float(ml[1]), surrgain=float(ml[3]), entropy=float(ml[4]),
                         stepsize=stepsize, expectedimprove=expectedimprove)
        # value function (trpo.py:288-295)
        vf = [P[k] for k in VF_KEYS]
        n = ob.shape[0]
        for _ in range(self.vf_iters):
            perm = torch.randperm(n, device=ob.device)
            for s in range(0, n - self.vf_batch + 1, self.vf_batch):      # include_final_partial_batch=False
                idx = perm[s:s + self.vf_batch]
                self.pi.ob_rms.update(ob[idx], group=self.group)
                _, _, vpred = policy_forward(P, self.pi.ob_rms.mean, self.pi.ob_rms.std, ob[idx])
                vferr = ((vpred - tdlamret[idx]) ** 2).mean()
                self.vfadam.update(flat_grad(vferr, vf), self.vf_stepsize)
        with torch.no_grad():
            _, _, vp = policy_forward(P, self.pi.ob_rms.mean, self.pi.ob_rms.std, ob)
            stats["vferr"] = float(((vp - tdlamret) ** 2).mean())
            if "vpred" in seg:                                            # trpo.py:298 (vpred before the update)
                stats["ev_tdlam_before"] = float(explained_variance(seg["vpred"].detach().reshape(-1), tdlamret))
        return stats
