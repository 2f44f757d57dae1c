"""BatchedSim: env-major batched state in PyTorch CUDA tensors + the libdmb200 C-ABI.

This is the host side of the hot path: it owns the state tensors (qpos/qvel/warmstart/phase/...),
passes raw device pointers through ctypes and enqueues every call on torch's current CUDA stream.
There is no CPU fallback: constructing a BatchedSim without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import lib as _lib
from .mjcf import ModelTables, load_tables
from .mocap import MocapTables, concat_clips, load_clip
from .model_blob import REF_AUX, DmbConfig, DmbMocap, DmbModel, default_config, pack_model

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def default_model_tables() -> ModelTables:
    """Compiled dp_env_v3.xml tables shipped with the repo (tools/build_assets.py)."""
    return load_tables(os.path.join(ASSETS, "dp_env_v3.model.npz"))


def motion_path(name: str) -> str:
    return os.path.join(ASSETS, "motions", name + ".npz")


def load_motions(names: Sequence[str]) -> MocapTables:
    clips = []
    for n in names:
        clips.append(load_clip(n) if os.path.exists(n) else load_clip(motion_path(n), name=n))
    return concat_clips(clips)


def make_mocap_struct(mc: MocapTables, ref_aux: Optional[np.ndarray] = None):
    """DmbMocap + the numpy arrays that back its pointers (keep them alive)."""
    cfg = np.ascontiguousarray(mc.data_config, dtype=np.float64)
    vel = np.ascontiguousarray(np.nan_to_num(mc.data_vel, nan=0.0, posinf=0.0, neginf=0.0), dtype=np.float64)
    aux = np.zeros((cfg.shape[0], REF_AUX)) if ref_aux is None else np.ascontiguousarray(ref_aux, dtype=np.float64)
    s = DmbMocap()
    s.nclip, s.nframe_total = len(mc.names), cfg.shape[0]
    for k in range(len(mc.names)):
        s.clip_start[k], s.clip_len[k], s.clip_dt[k] = int(mc.clip_start[k]), int(mc.clip_len[k]), float(mc.clip_dt[k])
    dp = C.POINTER(C.c_double)
    s.data_config, s.data_vel, s.ref_aux = cfg.ctypes.data_as(dp), vel.ctypes.data_as(dp), aux.ctypes.data_as(dp)
    return s, (cfg, vel, aux)


class HostArray:
    """Pinned, device-mapped float32 host array (``dmb_host_alloc``): ``array`` / ``tensor`` are numpy / torch views of
    the same bytes, ``dev_ptr`` is the address the step kernel uses to read or write them over PCIe."""

    def __init__(self, L, device_index: int, shape):
        self.shape = tuple(int(x) for x in shape)
        n = int(np.prod(self.shape))
        hp, dp = C.c_void_p(), C.c_void_p()
        rc = L.dmb_host_alloc(int(device_index), C.c_uint64(4 * n), C.byref(hp), C.byref(dp))
        if rc != 0:
            raise _lib.DmbError(f"dmb_host_alloc({4 * n} bytes) failed ({rc})")
        self._L, self.host_ptr, self.dev_ptr = L, hp.value, dp.value
        self._buf = (C.c_float * n).from_address(self.host_ptr)
        self.array = np.frombuffer(self._buf, dtype=np.float32).reshape(self.shape)
        self.tensor = torch.from_numpy(self.array)

    def free(self):
        """Release the pinned memory; the views must not be used afterwards."""
        if self.host_ptr:
            self.array = self.tensor = self._buf = None
            self._L.dmb_host_free(C.c_void_p(self.host_ptr))
            self.host_ptr = self.dev_ptr = 0


class BatchedSim:
    """N envs on one GPU.  All tensors are CUDA, env-major, contiguous."""

    def __init__(self, num_envs: int, motions: Sequence[str] = ("walk",), device: Optional[torch.device] = None,
                 seed: int = 0, first_env_id: int = 0, config: Optional[DmbConfig] = None,
                 model_tables: Optional[ModelTables] = None, clip_ids: Optional[torch.Tensor] = None,
                 max_con: int = 16, max_efc: int = 40, ref_aux: Optional[np.ndarray] = None, rec_depth: int = 4):
        if not torch.cuda.is_available():
            raise _lib.DmbError("BatchedSim needs a CUDA device: the hot path has no CPU fallback")
        self.L = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.N = int(num_envs)
        self.tables = model_tables if model_tables is not None else default_model_tables()
        self.model: DmbModel = pack_model(self.tables, max_con=max_con, max_efc=max_efc)
        self.config: DmbConfig = config if config is not None else default_config()
        self.mocap: MocapTables = load_motions(list(motions))
        self._mc_struct, self._mc_keep = make_mocap_struct(self.mocap, ref_aux)
        self.nq, self.nv, self.nu = self.tables.nq, self.tables.nv, self.tables.nu
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.dmb_create(C.byref(self.model), C.byref(self.config), C.byref(self._mc_struct), self.N,
                                         self.device.index, C.c_uint64(seed), C.c_uint32(first_env_id),
                                         C.byref(self.handle)), None, "dmb_create")
        self.obs_dim = int(self.L.dmb_obs_dim(self.handle))   # 56, or 197 for the DeepMimic state (obs_mode 1)
        d, N = self.device, self.N
        f32, i32 = torch.float32, torch.int32
        self.qpos = torch.zeros(N, _lib.QSTRIDE, dtype=f32, device=d)
        self.qvel = torch.zeros(N, _lib.VSTRIDE, dtype=f32, device=d)
        self.warm = torch.zeros(N, _lib.VSTRIDE, dtype=f32, device=d)
        if clip_ids is not None:
            ci = torch.as_tensor(clip_ids)
            if ci.numel() != N or int(ci.min()) < 0 or int(ci.max()) >= len(self.mocap.names):
                raise ValueError(f"clip_ids must hold {N} ids in [0, {len(self.mocap.names)}) (one per env)")
        self.clip = torch.zeros(N, dtype=i32, device=d) if clip_ids is None else clip_ids.to(d, i32).contiguous()
        self.idx_init = torch.zeros(N, dtype=i32, device=d)
        self.idx_curr = torch.zeros(N, dtype=i32, device=d)
        self.reset_count = torch.zeros(N, dtype=i32, device=d)  # bit pattern of uint32
        self.ep_len = torch.zeros(N, dtype=i32, device=d)
        self.ep_ret = torch.zeros(N, dtype=f32, device=d)
        self.flags = torch.zeros(N, dtype=i32, device=d)
        self.obs = torch.zeros(N, self.obs_dim, dtype=f32, device=d)
        self.reward = torch.zeros(N, dtype=f32, device=d)
        self.done = torch.zeros(N, dtype=torch.uint8, device=d)
        # the packed (obs, reward, done) record is multi-buffered: step t writes rec_buffers[t % rec_depth], so that
        # the all-gather of step t (dist.RecordGather, on its own stream) can overlap the following steps
        self.rec_buffers = [torch.zeros(N, self.obs_dim + 2, dtype=f32, device=d) for _ in range(max(2, int(rec_depth)))]
        self.rec_index = 0
        self.rec = self.rec_buffers[0]    # the record of the latest step
        self.last_ret = torch.zeros(N, dtype=f32, device=d)
        self.last_len = torch.zeros(N, dtype=i32, device=d)
        self._st = _lib.DmbState(*[t.data_ptr() for t in (self.qpos, self.qvel, self.warm, self.clip, self.idx_init,
                                                           self.idx_curr, self.reset_count, self.ep_len, self.ep_ret,
                                                           self.flags)])
        self._out = _lib.DmbStepOut(self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                                    self.rec.data_ptr(), self.last_ret.data_ptr(), self.last_len.data_ptr())
        self._st_ref, self._out_ref = C.byref(self._st), C.byref(self._out)
        qpos0 = torch.tensor(self.tables.qpos0, dtype=f32, device=d)
        self.qpos[:, : self.nq] = qpos0
        self.host_act: Optional[HostArray] = None    # step_host() defaults, allocated on first use
        self.host_rec: Optional[HostArray] = None
        self._host_arrays = []

    # ------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            if torch.cuda.is_available():
                torch.cuda.synchronize(self.device)     # no kernel may still be writing a host-mapped record
            for a in getattr(self, "_host_arrays", []):
                a.free()
            self._host_arrays = []
            self.host_act = self.host_rec = None
            self.L.dmb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_info(self) -> Dict[str, int]:
        g, b, s, w = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self.L.dmb_launch_info(self.handle, C.byref(g), C.byref(b), C.byref(s), C.byref(w))
        return dict(grid=g.value, block=b.value, smem_bytes=s.value, envs_per_cta=w.value)

    def kernel_launches(self) -> int:
        """Kernels launched by ``step`` so far (the scheduler sort + the fused step kernel)."""
        return int(self.L.dmb_kernel_launches(self.handle))

    # ------------------------------------------------------------------------------------
    def reset(self, mask: Optional[torch.Tensor] = None, mode: int = -1) -> torch.Tensor:
        """(Re)initialise envs (all, or where mask != 0); returns the observation tensor [N,56]."""
        mp = None
        if mask is not None:
            mask = mask.to(self.device, torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(self.L.dmb_reset(self.handle, C.byref(self._st), mp, mode, C.c_void_p(self.obs.data_ptr()),
                                        self._stream()), self.handle, "dmb_reset")
        return self.obs

    def step(self, action, rec_host: Optional["HostArray"] = None):
        """One env step for all envs.  action: CUDA float32 [N, nu] (or a HostArray, see ``step_host``).  Returns
        (obs, reward, done) tensors (views of internal buffers, overwritten by the next call).  With ``rec_host`` the
        kernel stores the packed record rows into that pinned host array instead of the device record ``self.rec``."""
        if isinstance(action, HostArray):
            if action.shape != (self.N, self.nu) or not action.dev_ptr:
                raise ValueError("action HostArray must be live and of shape [N, nu]")
            act_ptr = action.dev_ptr
        else:
            if action.device != self.device or action.dtype != torch.float32 or not action.is_contiguous() \
                    or tuple(action.shape) != (self.N, self.nu):
                raise ValueError("action must be a contiguous CUDA float32 tensor of shape [N, nu] on the sim's device")
            act_ptr = action.data_ptr()
        if rec_host is not None:
            if rec_host.shape != (self.N, self.obs_dim + 2) or not rec_host.dev_ptr:
                raise ValueError("rec_host must be a live HostArray of shape [N, obs_dim + 2]")
            self._out.rec = rec_host.dev_ptr
        else:
            self.rec_index = (self.rec_index + 1) % len(self.rec_buffers)
            self.rec = self.rec_buffers[self.rec_index]
            self._out.rec = self.rec.data_ptr()
        if torch.cuda.current_device() == self.device.index:     # the usual case: skip the device-guard round trip
            rc = self.L.dmb_step(self.handle, self._st_ref, act_ptr, self._out_ref, self._stream())
        else:
            with torch.cuda.device(self.device):
                rc = self.L.dmb_step(self.handle, self._st_ref, act_ptr, self._out_ref, self._stream())
        if rc != 0:
            _lib.check(rc, self.handle, "dmb_step")
        return self.obs, self.reward, self.done

    # --- host-mapped I/O (include/dmb.h "Host-mapped I/O"): for a policy that lives on the host ---------------------
    def alloc_host(self, shape) -> HostArray:
        """A pinned, device-mapped float32 host array owned by this sim (freed by ``close``)."""
        a = HostArray(self.L, self.device.index, shape)
        self._host_arrays.append(a)
        return a

    def enable_host_io(self):
        """Allocate the default host action [N, nu] and record [N, obs_dim + 2] arrays of ``step_host``."""
        if self.host_act is None:
            self.host_act = self.alloc_host((self.N, self.nu))
            self.host_rec = self.alloc_host((self.N, self.obs_dim + 2))
        return self.host_act, self.host_rec

    def step_host(self, act: Optional[HostArray] = None, rec: Optional[HostArray] = None) -> HostArray:
        """One env step in ONE launch with host-resident inputs and outputs: the kernel loads every env's action row
        from the pinned host array ``act`` and stores its (obs, reward, done) record row into the pinned host array
        ``rec`` over PCIe -- no cudaMemcpyAsync on either side.  Asynchronous like ``step``: synchronise the stream
        before reading ``rec.array`` (and before overwriting ``act``).  obs / reward / done / last_ret / last_len are
        still written to the device tensors; the device record ``self.rec`` is NOT written by this call.  The two
        sides are independent: ``step(cuda_action, rec_host=rec)`` keeps the action on the device (copied by the DMA
        engine, which moves a large batch faster than SM loads over PCIe do) and only writes the record to the host."""
        if act is None or rec is None:
            self.enable_host_io()
            act = self.host_act if act is None else act
            rec = self.host_rec if rec is None else rec
        if not isinstance(act, HostArray) or not isinstance(rec, HostArray):
            raise ValueError("step_host: act must be a HostArray [N, nu] and rec a HostArray [N, obs_dim + 2]")
        self.step(act, rec_host=rec)
        return rec

    def get_obs(self) -> torch.Tensor:
        with torch.cuda.device(self.device):
            _lib.check(self.L.dmb_get_obs(self.handle, C.byref(self._st), C.c_void_p(self.obs.data_ptr()),
                                          self._stream()), self.handle, "dmb_get_obs")
        return self.obs

    def mocap_sample(self, t, clip_ids: Optional[torch.Tensor] = None):
        """Interpolated reference poses at mocap times ``t`` [n] seconds (phase_mode 1 arithmetic: lerp / slerp
        between frames, loop wrap with root-offset accumulation).  Returns (qpos [n,nq], qvel [n,nv], phase [n])
        CUDA float32 tensors; ``clip_ids`` [n] int (default: clip 0)."""
        t = torch.as_tensor(t, dtype=torch.float64).reshape(-1)
        n = t.numel()
        clip = (torch.zeros(n, dtype=torch.int64) if clip_ids is None else torch.as_tensor(clip_ids).reshape(-1).cpu().long())
        dt = torch.as_tensor(np.asarray(self.mocap.clip_dt), dtype=torch.float64)[clip]
        u = (t.cpu() / dt).to(self.device).contiguous()                 # frame coordinate t / clip_dt, float64
        clip_d = clip.to(self.device, torch.int32).contiguous()
        qpos = torch.empty(n, _lib.QSTRIDE, dtype=torch.float32, device=self.device)
        qvel = torch.empty(n, _lib.VSTRIDE, dtype=torch.float32, device=self.device)
        phase = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.dmb_mocap_sample(self.handle, C.c_void_p(clip_d.data_ptr()), C.c_void_p(u.data_ptr()), n,
                                               C.c_void_p(qpos.data_ptr()), C.c_void_p(qvel.data_ptr()),
                                               C.c_void_p(phase.data_ptr()), self._stream()), self.handle,
                       "dmb_mocap_sample")
        return qpos[:, : self.nq], qvel[:, : self.nv], phase

    def set_state(self, qpos, qvel, warm=None, idx_curr=None):
        """Overwrite state from arrays [N,nq] / [N,nv] (numpy or torch, any float dtype)."""
        q = torch.as_tensor(np.asarray(qpos) if not torch.is_tensor(qpos) else qpos).to(self.device, torch.float32)
        v = torch.as_tensor(np.asarray(qvel) if not torch.is_tensor(qvel) else qvel).to(self.device, torch.float32)
        self.qpos.zero_(); self.qvel.zero_()
        self.qpos[:, : self.nq] = q
        self.qvel[:, : self.nv] = v
        self.warm.zero_()
        if warm is not None:
            w = torch.as_tensor(np.asarray(warm) if not torch.is_tensor(warm) else warm).to(self.device, torch.float32)
            self.warm[:, : self.nv] = w
        if idx_curr is not None:
            self.idx_curr.copy_(torch.as_tensor(idx_curr).to(self.device, torch.int32))

    def get_state(self):
        return (self.qpos[:, : self.nq].double().cpu().numpy(), self.qvel[:, : self.nv].double().cpu().numpy(),
                self.warm[:, : self.nv].double().cpu().numpy())

    # ------------------------------------------------------------------------------------
    def forward_debug(self, ctrl: torch.Tensor) -> Dict[str, np.ndarray]:
        """Run one mj_forward at the current state and return the stage outputs as numpy arrays."""
        stride = self.L.dmb_debug_stride()
        dbg = torch.zeros(self.N, stride, dtype=torch.float32, device=self.device)
        ctrl = ctrl.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.L.dmb_forward_debug(self.handle, C.byref(self._st), C.c_void_p(ctrl.data_ptr()),
                                                C.c_void_p(dbg.data_ptr()), self._stream()), self.handle,
                       "dmb_forward_debug")
        h = dbg.cpu().numpy().astype(np.float64)
        off = lambda n: self.L.dmb_debug_offset(n.encode())
        nb, nv, nM = self.tables.nbody, self.nv, self.tables.nM
        ne = off("efc_R") - off("efc_pos")
        out = dict(
            xpos=h[:, off("xpos"): off("xpos") + nb * 3].reshape(-1, nb, 3),
            xquat=h[:, off("xquat"): off("xquat") + nb * 4].reshape(-1, nb, 4),
            xipos=h[:, off("xipos"): off("xipos") + nb * 3].reshape(-1, nb, 3),
            com=h[:, off("com"): off("com") + 3],
            qM=h[:, off("qM"): off("qM") + nM], qLD=h[:, off("qLD"): off("qLD") + nM],
            qfrc_bias=h[:, off("qfrc_bias"): off("qfrc_bias") + nv],
            qfrc_smooth=h[:, off("qfrc_smooth"): off("qfrc_smooth") + nv],
            qacc_smooth=h[:, off("qacc_smooth"): off("qacc_smooth") + nv],
            ncon=h[:, off("ncon")].astype(int), nefc=h[:, off("nefc")].astype(int), iter=h[:, off("iter")].astype(int),
            z_com=h[:, off("z_com")],
            contact=h[:, off("contact"): off("efc_pos")].reshape(h.shape[0], -1, 16),
            efc_pos=h[:, off("efc_pos"): off("efc_pos") + ne], efc_R=h[:, off("efc_R"): off("efc_R") + ne],
            efc_aref=h[:, off("efc_aref"): off("efc_aref") + ne], efc_b=h[:, off("efc_b"): off("efc_b") + ne],
            efc_force=h[:, off("efc_force"): off("efc_force") + ne],
            efc_AR_diag=h[:, off("efc_AR_diag"): off("efc_AR_diag") + ne],
            qacc=h[:, off("qacc"): off("qacc") + nv],
            cvel=h[:, off("cvel"): off("cvel") + nb * 6].reshape(-1, nb, 6),
        )
        return out
