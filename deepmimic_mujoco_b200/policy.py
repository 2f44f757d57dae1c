"""Batched policy / value inference and the observation filter, device resident.

SURVEY.md section 8(f) rank 1 -- the immediate caller of the env step in the reference:
  /root/reference/src/mlp_policy_trpo.py:14-74   MlpPolicy (2 x tanh(100) policy and value nets, diag Gaussian)
  /root/reference/src/utils/misc_util.py:32-70   RunningMeanStd (sum / sumsq / count, eps 1e-2, var floor 1e-2)
  /root/reference/src/trpo.py:49                 ac, vpred = pi.act(stochastic, ob)   (batch 1, TensorFlow)
Here one fused CUDA kernel (csrc/dmb_policy.cuh, include/dmb_policy.h) evaluates both networks for all
envs per step; parameters are plain torch tensors (so a PyTorch learner can own and update them).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import lib as _lib


class DmbPolicy(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("hid", C.c_int32), ("pad", C.c_int32)] + [
        (n, C.c_void_p) for n in ("ob_mean", "ob_std", "pw1", "pb1", "pw2", "pb2", "pw3", "pb3", "logstd",
                                   "vw1", "vb1", "vw2", "vb2", "vw3", "vb3")]


class RunningMeanStd:
    """misc_util.py:32-70 on the device (float64 accumulators, float32 mean / std views)."""

    def __init__(self, shape, device, epsilon: float = 1e-2):
        self.sum = torch.zeros(shape, dtype=torch.float64, device=device)
        self.sumsq = torch.full(shape, epsilon, dtype=torch.float64, device=device)
        self.count = torch.tensor(epsilon, dtype=torch.float64, device=device)
        self.mean = torch.zeros(shape, dtype=torch.float32, device=device)
        self.std = torch.ones(shape, dtype=torch.float32, device=device)
        self._refresh()

    def _refresh(self):
        mean = (self.sum / self.count).float()
        self.mean.copy_(mean)
        self.std.copy_(torch.sqrt(torch.clamp((self.sumsq / self.count).float() - mean * mean, min=1e-2)))

    def load(self, total, sumsq, count):
        """Overwrite the accumulators (e.g. with ``obfilter/runningsum|runningsumsq|count`` of a reference checkpoint)."""
        as64 = lambda a: torch.as_tensor(np.array(a, dtype=np.float64), device=self.sum.device)
        self.sum.copy_(as64(total)); self.sumsq.copy_(as64(sumsq)); self.count.copy_(as64(count))
        self._refresh()

    def update(self, x: torch.Tensor, group=None):
        x = x.reshape(-1, x.shape[-1]).double()
        add = torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([x.shape[0]], dtype=torch.float64, device=x.device)])
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            torch.distributed.all_reduce(add, group=group)   # MPI Allreduce SUM in the reference
        n = self.sum.numel()
        self.sum += add[:n]; self.sumsq += add[n:2 * n]; self.count += add[2 * n]
        self._refresh()


def normc_(w: torch.Tensor, std: float, gen: torch.Generator) -> torch.Tensor:
    """utils/tf_util.py normc_initializer: N(0,1) columns rescaled to norm `std`."""
    w.copy_(torch.randn(w.shape, generator=gen, device=w.device))
    w.mul_(std / torch.sqrt((w * w).sum(dim=0, keepdim=True)))
    return w


class MlpPolicy:
    """Parameters + fused act() for N envs.  hid_size=100, num_hid_layers=2 as in trpo.py:346."""

    def __init__(self, obs_dim: int = 56, act_dim: int = 28, hid: int = 100, device=None, seed: int = 0):
        if not torch.cuda.is_available():
            raise _lib.DmbError("MlpPolicy.act runs on the GPU only")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if obs_dim > 64 or hid > 128 or act_dim > 32:
            raise ValueError(f"MlpPolicy: the fused inference kernel (include/dmb_policy.h) keeps both networks' weights in "
                             f"shared memory and supports obs_dim <= 64, hid <= 128, act_dim <= 32; got obs_dim={obs_dim}, "
                             f"hid={hid}, act_dim={act_dim} (the 197-d DeepMimic state of obs_mode=1 does not fit)")
        self.obs_dim, self.act_dim, self.hid = obs_dim, act_dim, hid
        g = torch.Generator(device=self.device); g.manual_seed(seed)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.params: Dict[str, torch.Tensor] = dict(
            vw1=normc_(z(obs_dim, hid), 1.0, g), vb1=z(hid), vw2=normc_(z(hid, hid), 1.0, g), vb2=z(hid),
            vw3=normc_(z(hid, 1), 1.0, g), vb3=z(1),
            pw1=normc_(z(obs_dim, hid), 1.0, g), pb1=z(hid), pw2=normc_(z(hid, hid), 1.0, g), pb2=z(hid),
            pw3=normc_(z(hid, act_dim), 0.01, g), pb3=z(act_dim), logstd=z(act_dim))
        self.ob_rms = RunningMeanStd((obs_dim,), self.device)
        self.L = _lib.load()
        self.L.dmb_policy_act.argtypes = [C.POINTER(DmbPolicy), C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32,
                                          C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.seed = seed
        self.step_count = 0

    def load_arrays(self, arrays) -> None:
        """Overwrite parameters and the observation filter from host arrays named as in ``self.params`` plus
        ``ob_sum`` / ``ob_sumsq`` / ``ob_count`` (tf_checkpoint.policy_arrays); shapes must match this policy."""
        for k, t in self.params.items():
            a = torch.as_tensor(np.array(arrays[k], dtype=np.float32))      # a copy: checkpoint arrays are read-only views
            if tuple(a.shape) != tuple(t.shape):
                raise ValueError(f"{k}: checkpoint shape {tuple(a.shape)} != policy shape {tuple(t.shape)}")
            t.copy_(a)                                           # in place: pointers held by a learner stay valid
        self.ob_rms.load(arrays["ob_sum"], arrays["ob_sumsq"], arrays["ob_count"])

    def load_tf_checkpoint(self, prefix: str, scope: str = "pi") -> None:
        """Restore a policy saved by the reference (``U.load_state(load_model_path)``, trpo.py:207-208 / 367):
        ``prefix`` is the TensorFlow checkpoint path the reference's ``--load_model_path`` takes, ``scope`` the
        variable scope of the policy ("pi"; the checkpoint also holds TRPO's "oldpi").  No TensorFlow needed."""
        from .tf_checkpoint import policy_arrays, read_checkpoint
        self.load_arrays(policy_arrays(read_checkpoint(prefix), scope))

    def host_arrays(self):
        """Parameters and observation-filter accumulators as host numpy arrays (the dict ``load_arrays`` takes)."""
        out = {k: t.detach().cpu().numpy() for k, t in self.params.items()}
        out.update(ob_sum=self.ob_rms.sum.cpu().numpy(), ob_sumsq=self.ob_rms.sumsq.cpu().numpy(),
                   ob_count=self.ob_rms.count.cpu().numpy())
        return out

    def save_tf_checkpoint(self, prefix: str, scopes=("pi", "oldpi")) -> None:
        """Save in the reference's on-disk format (``U.save_state`` = ``tf.train.Saver().save``, trpo.py:220-224): a
        TensorFlow V2 checkpoint holding this policy under every scope in ``scopes`` -- the reference's graph has the
        policy twice, "pi" and TRPO's "oldpi" -- so that its ``--load_model_path`` / ``--pretrained_weight_path``
        restore it."""
        from .tf_checkpoint import policy_tensors, write_checkpoint, write_checkpoint_state
        arrays, tensors = self.host_arrays(), {}
        for sc in scopes:
            tensors.update(policy_tensors(arrays, sc))
        write_checkpoint(prefix, tensors)
        write_checkpoint_state(prefix)

    def _struct(self) -> DmbPolicy:
        p = self.params
        s = DmbPolicy(self.obs_dim, self.act_dim, self.hid, 0, self.ob_rms.mean.data_ptr(), self.ob_rms.std.data_ptr())
        for k in ("pw1", "pb1", "pw2", "pb2", "pw3", "pb3", "logstd", "vw1", "vb1", "vw2", "vb2", "vw3", "vb3"):
            t = p[k]
            assert t.is_contiguous() and t.dtype == torch.float32 and t.device == self.device
            setattr(s, k, t.data_ptr())
        return s

    def act(self, stochastic: bool, ob: torch.Tensor, out_ac: Optional[torch.Tensor] = None,
            out_vpred: Optional[torch.Tensor] = None, out_mean: Optional[torch.Tensor] = None, first_row: int = 0):
        """ac [N,act_dim], vpred [N] for ob [N,obs_dim] (CUDA float32, contiguous)."""
        n = ob.shape[0]
        assert ob.is_contiguous() and ob.dtype == torch.float32 and ob.shape[1] == self.obs_dim
        ac = out_ac if out_ac is not None else torch.empty(n, self.act_dim, dtype=torch.float32, device=self.device)
        vp = out_vpred if out_vpred is not None else torch.empty(n, dtype=torch.float32, device=self.device)
        s = self._struct()
        with torch.cuda.device(self.device):
            rc = self.L.dmb_policy_act(C.byref(s), C.c_void_p(ob.data_ptr()), n, int(bool(stochastic)), C.c_uint64(self.seed),
                                       C.c_uint32(self.step_count), C.c_uint32(first_row), C.c_void_p(ac.data_ptr()),
                                       C.c_void_p(vp.data_ptr()), C.c_void_p(out_mean.data_ptr()) if out_mean is not None else None,
                                       C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise _lib.DmbError(f"dmb_policy_act failed ({rc})")
        self.step_count += 1
        return ac, vp
