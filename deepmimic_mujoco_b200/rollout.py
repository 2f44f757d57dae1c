"""Device-resident rollout segments + GAE (SURVEY.md section 8(f) rank 1/2 entry points).

Batched restatement of /root/reference/src/trpo.py:27-80 ``traj_segment_generator`` (a horizon of T steps
per env instead of one env, history kept in CUDA tensors instead of numpy arrays, auto-reset inside the
env step instead of the reset()/reset_model_init() pair) and of trpo.py:83-94 ``add_vtarg_and_adv``.
"""
from __future__ import annotations

from typing import Dict

import torch

from .env import DPVecEnv
from .policy import MlpPolicy


def add_vtarg_and_adv(seg: Dict[str, torch.Tensor], gamma: float, lam: float) -> None:
    """GAE(lambda) over [T, N] tensors; seg['new'][t] = 1 if step t starts a new episode (trpo.py:83-94).
    CUDA segments go through one kernel (``dmb_gae``, one thread per env); CPU tensors (tests) use the loop."""
    rew, vpred, new = seg["rew"], seg["vpred"], seg["new"]
    T = rew.shape[0]
    if rew.is_cuda:
        import ctypes as C
        from . import lib as _lib
        L = _lib.load()
        rew, vpred, new = rew.contiguous().float(), vpred.contiguous().float(), new.contiguous().float()
        nextv = seg["nextvpred"].contiguous().float()
        adv, ret = torch.empty_like(rew), torch.empty_like(rew)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(rew.device):
            rc = L.dmb_gae(p(rew), p(vpred), p(new), p(nextv), T, rew.shape[1], float(gamma), float(lam), p(adv), p(ret),
                           C.c_void_p(torch.cuda.current_stream(rew.device).cuda_stream))
        if rc != 0:
            raise _lib.DmbError(f"dmb_gae failed ({rc})")
        seg["adv"], seg["tdlamret"] = adv, ret
        return
    nextv = torch.cat([vpred[1:], seg["nextvpred"][None]], dim=0)
    nextnew = torch.cat([new[1:], torch.zeros_like(new[:1])], dim=0)
    adv = torch.empty_like(rew)
    last = torch.zeros_like(rew[0])
    for t in reversed(range(T)):
        nonterminal = 1.0 - nextnew[t]
        delta = rew[t] + gamma * nextv[t] * nonterminal - vpred[t]
        last = delta + gamma * lam * nonterminal * last
        adv[t] = last
    seg["adv"] = adv
    seg["tdlamret"] = adv + vpred


class SegmentGenerator:
    """Yields {"ob","ac","rew","vpred","new","nextvpred","ep_rets","ep_lens"} with leading dims [T, N]."""

    def __init__(self, pi: MlpPolicy, env: DPVecEnv, horizon: int, stochastic: bool = True):
        self.pi, self.env, self.T, self.stochastic = pi, env, horizon, stochastic
        N, d = env.num_envs, env.sim.device
        self.ob = torch.zeros(horizon, N, env.sim.obs_dim, device=d)
        self.ac = torch.zeros(horizon, N, env.sim.nu, device=d)
        self.rew = torch.zeros(horizon, N, device=d)
        self.vpred = torch.zeros(horizon, N, device=d)
        self.new = torch.zeros(horizon, N, device=d)
        self.cur_ob = env.reset().clone()
        self.cur_new = torch.ones(N, device=d)      # trpo.py:31 `new = True`

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, torch.Tensor]:
        # finished-episode records are kept per step on the device and compacted once per segment: a boolean-mask
        # gather inside the loop would force a host sync every env step
        N, d = self.env.num_envs, self.ob.device
        done_hist = torch.empty(self.T, N, dtype=torch.bool, device=d)
        ret_hist = torch.empty(self.T, N, device=d)
        len_hist = torch.empty(self.T, N, dtype=torch.int32, device=d)
        for t in range(self.T):
            self.ob[t].copy_(self.cur_ob)
            self.new[t].copy_(self.cur_new)
            self.pi.act(self.stochastic, self.cur_ob, out_ac=self.ac[t], out_vpred=self.vpred[t])
            obs, rew, done, info = self.env.step(self.ac[t])
            self.rew[t].copy_(rew)
            self.cur_ob.copy_(obs)
            self.cur_new.copy_(done)
            done_hist[t].copy_(done)
            ret_hist[t].copy_(info["episode_return"])
            len_hist[t].copy_(info["episode_length"])
        nextvpred = torch.empty(N, device=d)
        tmp_ac = torch.empty(N, self.env.sim.nu, device=d)
        self.pi.act(self.stochastic, self.cur_ob, out_ac=tmp_ac, out_vpred=nextvpred)
        nextvpred = nextvpred * (1.0 - self.cur_new)   # trpo.py:56
        return {"ob": self.ob, "ac": self.ac, "rew": self.rew, "vpred": self.vpred, "new": self.new,
                "nextvpred": nextvpred, "ep_rets": ret_hist[done_hist], "ep_lens": len_hist[done_hist]}


def evaluate(pi, env, horizon: int = 1024, stochastic: bool = False) -> Dict[str, torch.Tensor]:
    """The reference's evaluate task (/root/reference/src/trpo.py:356-393 ``runner`` over trpo.py:397-436
    ``traj_1_generator``) for all N envs at once: every env plays ONE trajectory from a fresh reset with
    ``pi.act(stochastic, ob)`` until it is done or has made ``horizon + 1`` steps (the reference's ``t >= horizon``
    break comes after the step).  ``env`` is a DPVecEnv built with ``reset_mode=1`` (the reference calls
    ``reset_model_init`` before each trajectory) and ``auto_reset=True``; an env that has finished keeps stepping
    (its later episodes are ignored), so the loop stays one launch per step with no host round trip except a
    completion check every 64 steps.  Returns ``ep_len`` [N] int32, ``ep_ret`` [N] float32, ``finished`` [N] bool
    (False: cut by the horizon) and the reference's two printed numbers ``avg_len`` / ``avg_ret`` (0-d tensors)."""
    ob = env.reset()
    n = ob.shape[0]
    ep_len = torch.zeros(n, dtype=torch.int32, device=ob.device)
    ep_ret = torch.zeros(n, dtype=torch.float32, device=ob.device)
    running = torch.ones(n, dtype=torch.bool, device=ob.device)
    finished = torch.zeros(n, dtype=torch.bool, device=ob.device)
    for t in range(int(horizon) + 1):
        ac, _ = pi.act(stochastic, ob)
        ob, rew, done, _ = env.step(ac)
        ep_len += running.to(torch.int32)
        ep_ret += torch.where(running, rew, torch.zeros_like(rew))
        ended = running & (done != 0)
        finished |= ended
        running = running & ~ended
        if t % 64 == 63 and not bool(running.any()):
            break
    return {"ep_len": ep_len, "ep_ret": ep_ret, "finished": finished,
            "avg_len": ep_len.double().mean(), "avg_ret": ep_ret.double().mean()}
