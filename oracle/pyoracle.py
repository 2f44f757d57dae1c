"""ctypes binding of the CPU float64 oracle (oracle/dm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/dm_oracle.h.  PARITY UNPINNED.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from deepmimic_mujoco_b200.model_blob import (MAX_BODY, MAX_DOF, MAX_GEOM, MAX_JNT, MAX_M, MAX_Q, MAX_U, DmbConfig,
                                               DmbMocap, DmbModel)

_HERE = os.path.dirname(os.path.abspath(__file__))
MAXCON, MAXEFC = 64, 192
i32, f64 = C.c_int32, C.c_double


class DmoContact(C.Structure):
    _fields_ = [("dist", f64), ("pos", f64 * 3), ("frame", f64 * 9), ("mu", f64),
                ("geom1", i32), ("geom2", i32), ("dim", i32), ("efc_address", i32)]


class DmoData(C.Structure):
    _fields_ = [
        ("qpos", f64 * MAX_Q), ("qvel", f64 * MAX_DOF), ("ctrl", f64 * MAX_U), ("qacc_warmstart", f64 * MAX_DOF),
        ("xpos", (f64 * 3) * MAX_BODY), ("xquat", (f64 * 4) * MAX_BODY), ("xmat", (f64 * 9) * MAX_BODY),
        ("xipos", (f64 * 3) * MAX_BODY), ("xaxis", (f64 * 3) * MAX_JNT),
        ("geom_xpos", (f64 * 3) * MAX_GEOM), ("geom_xmat", (f64 * 9) * MAX_GEOM),
        ("com", f64 * 3),
        ("cinert", (f64 * 10) * MAX_BODY), ("crb", (f64 * 10) * MAX_BODY), ("cdof", (f64 * 6) * MAX_DOF),
        ("qM", f64 * MAX_M), ("qLD", f64 * MAX_M), ("qLDiagInv", f64 * MAX_DOF),
        ("cvel", (f64 * 6) * MAX_BODY), ("cdof_dot", (f64 * 6) * MAX_DOF),
        ("qfrc_bias", f64 * MAX_DOF), ("qfrc_passive", f64 * MAX_DOF), ("qfrc_actuator", f64 * MAX_DOF),
        ("qfrc_smooth", f64 * MAX_DOF), ("qacc_smooth", f64 * MAX_DOF),
        ("ncon", i32), ("nefc", i32), ("solver_iter", i32), ("flags", i32),
        ("contact", DmoContact * MAXCON),
        ("efc_type", i32 * MAXEFC), ("efc_id", i32 * MAXEFC),
        ("efc_J", (f64 * MAX_DOF) * MAXEFC),
        ("efc_pos", f64 * MAXEFC), ("efc_margin", f64 * MAXEFC), ("efc_diagApprox", f64 * MAXEFC),
        ("efc_R", f64 * MAXEFC), ("efc_D", f64 * MAXEFC), ("efc_KBIP", (f64 * 4) * MAXEFC),
        ("efc_vel", f64 * MAXEFC), ("efc_aref", f64 * MAXEFC), ("efc_b", f64 * MAXEFC), ("efc_force", f64 * MAXEFC),
        ("efc_AR", (f64 * MAXEFC) * MAXEFC),
        ("qfrc_constraint", f64 * MAX_DOF), ("qacc", f64 * MAX_DOF),
    ]

    def arr(self, name, *shape):
        a = np.ctypeslib.as_array(getattr(self, name))
        if shape:
            a = a[tuple(slice(0, s) for s in shape)]
        return a


class DmoEnv(C.Structure):
    _fields_ = [("d", DmoData), ("clip", i32), ("idx_init", i32), ("idx_curr", i32), ("ep_len", i32),
                ("seed", C.c_uint64), ("env_id", C.c_uint32), ("reset_count", C.c_uint32),
                ("ep_ret", f64), ("reward_terms", f64 * 5), ("zcom_last", f64)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdm_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("dm_oracle.c", "dm_oracle.h")] + [
        os.path.join(_HERE, "..", "include", "dmb_model.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdm_oracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.dmo_sizeof_data.restype = C.c_ulong
        L.dmo_sizeof_env.restype = C.c_ulong
        L.dmo_sizeof_model.restype = C.c_ulong
        assert L.dmo_sizeof_data() == C.sizeof(DmoData), (L.dmo_sizeof_data(), C.sizeof(DmoData))
        assert L.dmo_sizeof_model() == C.sizeof(DmbModel), (L.dmo_sizeof_model(), C.sizeof(DmbModel))
        assert L.dmo_sizeof_env() == C.sizeof(DmoEnv), (L.dmo_sizeof_env(), C.sizeof(DmoEnv))
        mp, dp = C.POINTER(DmbModel), C.POINTER(DmoData)
        for fn in ("dmo_fwd_position", "dmo_fwd_velocity", "dmo_fwd_actuation", "dmo_fwd_acceleration",
                   "dmo_fwd_constraint", "dmo_forward", "dmo_step", "dmo_kinematics"):
            getattr(L, fn).argtypes = [mp, dp]
            getattr(L, fn).restype = None
        L.dmo_solve_M.argtypes = [mp, dp, C.POINTER(f64)]
        L.dmo_full_M.argtypes = [mp, dp, C.POINTER(f64)]
        ep, cp, mcp = C.POINTER(DmoEnv), C.POINTER(DmbConfig), C.POINTER(DmbMocap)
        L.dmo_env_init.argtypes = [mp, cp, mcp, ep, C.c_uint64, C.c_uint32, i32]
        L.dmo_env_reset.argtypes = [mp, cp, mcp, ep, C.c_int]
        L.dmo_env_set_state.argtypes = [mp, ep, C.POINTER(f64), C.POINTER(f64)]
        L.dmo_env_step.argtypes = [mp, cp, mcp, ep, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64)]
        L.dmo_env_step.restype = C.c_int
        L.dmo_env_obs.argtypes = [mp, ep, C.POINTER(f64)]
        L.dmo_env_obs_dm.argtypes = [mp, cp, mcp, ep, C.POINTER(f64)]
        L.dmo_env_obs_dm.restype = C.c_int
        L.dmo_mocap_sample.argtypes = [mp, mcp, C.c_int, C.c_double, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64)]
        L.dmo_mocap_sample.restype = None
        L.dmo_ref_aux.argtypes = [mp, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64)]
        L.dmo_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.dmo_rollout.argtypes = [mp, cp, mcp, ep, C.c_long, C.c_uint64]
        L.dmo_rollout.restype = C.c_long
        ip, fp = C.POINTER(i32), C.POINTER(f64)
        L.dmo_batch_step.argtypes = [mp, cp, mcp, C.c_int, C.c_uint64, C.c_uint32, fp, fp, fp, ip, ip, ip, ip, fp, ip,
                                     fp, fp, C.c_int, fp, ip, fp, ip, ip, ip, fp]
        L.dmo_batch_step.restype = None
        _lib = L
    return _lib


def dptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(f64))


class Oracle:
    """Single-env float64 physics oracle with direct access to every stage output."""

    def __init__(self, model: DmbModel):
        self.m = model
        self.d = DmoData()
        self.L = lib()
        self.nq, self.nv, self.nu = model.nq, model.nv, model.nu

    def set_state(self, qpos, qvel, ctrl=None, warm=None):
        self.d.arr("qpos")[: self.nq] = qpos
        self.d.arr("qvel")[: self.nv] = qvel
        self.d.arr("ctrl")[: self.nu] = 0.0 if ctrl is None else ctrl
        self.d.arr("qacc_warmstart")[: self.nv] = 0.0 if warm is None else warm
        self.d.flags = 0   # overflow / non-finite flags accumulate until the next set_state

    def forward(self):
        self.L.dmo_forward(C.byref(self.m), C.byref(self.d))

    def step(self):
        self.L.dmo_step(C.byref(self.m), C.byref(self.d))

    def full_M(self):
        M = np.zeros((self.nv, self.nv))
        self.L.dmo_full_M(C.byref(self.m), C.byref(self.d), dptr(M))
        return M

    def solve_M(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        buf = np.zeros(MAX_DOF)
        buf[: self.nv] = x
        self.L.dmo_solve_M(C.byref(self.m), C.byref(self.d), dptr(buf))
        return buf[: self.nv].copy()

    @property
    def qpos(self):
        return self.d.arr("qpos")[: self.nq]

    @property
    def qvel(self):
        return self.d.arr("qvel")[: self.nv]


def batch_step(model, cfg, mcs, seed, first_env_id, state: dict, action, obs_dim: int) -> dict:
    """One oracle env step for a batch of envs from explicit inputs (dmo_batch_step).  ``state`` holds numpy arrays
    qpos [n,nq], qvel [n,nv], warm [n,nv] (float64) and clip, idx_init, idx_curr, ep_len, reset_count (int32), ep_ret
    (float64); it is NOT modified.  Returns the post-step arrays plus obs, reward, done, last_ret, last_len, flags,
    nefc (of the last RK4 stage)."""
    L = lib()
    n = state["qpos"].shape[0]
    f = lambda k: np.ascontiguousarray(state[k], dtype=np.float64).copy()
    i = lambda k: np.ascontiguousarray(state[k], dtype=np.int32).copy()
    out = dict(qpos=f("qpos"), qvel=f("qvel"), warm=f("warm"), clip=i("clip"), idx_init=i("idx_init"),
               idx_curr=i("idx_curr"), ep_len=i("ep_len"), ep_ret=f("ep_ret"), reset_count=i("reset_count"))
    act = np.ascontiguousarray(action, dtype=np.float64)
    out.update(obs=np.zeros((n, obs_dim)), reward=np.zeros(n), done=np.zeros(n, np.int32), last_ret=np.zeros(n),
               last_len=np.zeros(n, np.int32), flags=np.zeros(n, np.int32), nefc=np.zeros(n, np.int32), zcom=np.zeros(n))
    ip = lambda a: a.ctypes.data_as(C.POINTER(i32))
    L.dmo_batch_step(C.byref(model), C.byref(cfg), C.byref(mcs), n, C.c_uint64(seed), C.c_uint32(first_env_id),
                     dptr(out["qpos"]), dptr(out["qvel"]), dptr(out["warm"]), ip(out["clip"]), ip(out["idx_init"]),
                     ip(out["idx_curr"]), ip(out["ep_len"]), dptr(out["ep_ret"]), ip(out["reset_count"]), dptr(act),
                     dptr(out["obs"]), obs_dim, dptr(out["reward"]), ip(out["done"]), dptr(out["last_ret"]),
                     ip(out["last_len"]), ip(out["flags"]), ip(out["nefc"]), dptr(out["zcom"]))
    return out
