"""ctypes binding of libdmb200.so (include/dmb.h).  No fallback: if the CUDA library is
missing or no GPU is visible, everything here raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

from .model_blob import DmbConfig, DmbMocap, DmbModel

_PKG = os.path.dirname(os.path.abspath(__file__))
# DMB_LIB selects another in-tree build of the same sources (A/B kernel variants for measurements)
LIB_PATH = os.environ.get("DMB_LIB") or os.path.join(_PKG, "libdmb200.so")
CSRC = os.path.join(_PKG, "csrc")
QSTRIDE = 36
VSTRIDE = 36

# -prec-div=false / -prec-sqrt=false: 2-ulp MUFU-based division and square root instead of the IEEE sequences (15
# instructions + a slow path per division): 7 % less code, +2.3 % env-steps/s, no measurable change of the
# kernel-vs-oracle errors (profiles/r2_c19_*); the kernel is fp32 against a float64 reference either way
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-prec-div=false", "-prec-sqrt=false"]


class DmbState(C.Structure):
    _fields_ = [("qpos", C.c_void_p), ("qvel", C.c_void_p), ("warm", C.c_void_p), ("clip", C.c_void_p),
                ("idx_init", C.c_void_p), ("idx_curr", C.c_void_p), ("reset_count", C.c_void_p),
                ("ep_len", C.c_void_p), ("ep_ret", C.c_void_p), ("flags", C.c_void_p)]


class DmbStepOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p), ("rec", C.c_void_p),
                ("last_ret", C.c_void_p), ("last_len", C.c_void_p)]


class DmbError(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/dmb.cu for sm_100a into the in-tree libdmb200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(_PKG, "..", "include", f) for f in ("dmb.h", "dmb_model.h", "dmb_policy.h")]
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s)):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "dmb.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib: Optional[C.CDLL] = None

EXPORTS = ("dmb_version", "dmb_sizeof_model", "dmb_sizeof_config", "dmb_sizeof_mocap", "dmb_sizeof_tile", "dmb_create", "dmb_destroy", "dmb_reset", "dmb_step", "dmb_get_obs", "dmb_forward_debug",
           "dmb_debug_stride", "dmb_debug_offset", "dmb_launch_info", "dmb_kernel_launches", "dmb_peer_alloc", "dmb_peer_open",
           "dmb_peer_close", "dmb_peer_free", "dmb_set_peer_gather", "dmb_peer_wait", "dmb_set_peer_wait", "dmb_host_alloc", "dmb_host_free", "dmb_host_device_pointer", "dmb_last_error", "dmb_obs_dim", "dmb_mocap_sample", "dmb_get_trace")
POLICY_EXPORTS = ("dmb_policy_act", "dmb_gae")   # include/dmb_policy.h


def load() -> C.CDLL:
    """dlopen the CUDA library; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DmbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    hp = C.c_void_p
    L.dmb_version.restype = C.c_int
    for fn, st in (("dmb_sizeof_model", DmbModel), ("dmb_sizeof_config", DmbConfig), ("dmb_sizeof_mocap", DmbMocap)):
        getattr(L, fn).restype = C.c_int32
        if getattr(L, fn)() != C.sizeof(st):
            raise DmbError(f"ABI mismatch: {fn}() = {getattr(L, fn)()} but ctypes mirror is {C.sizeof(st)} bytes")
    L.dmb_create.argtypes = [C.POINTER(DmbModel), C.POINTER(DmbConfig), C.POINTER(DmbMocap), C.c_int32, C.c_int32,
                             C.c_uint64, C.c_uint32, C.POINTER(hp)]
    L.dmb_destroy.argtypes = [hp]
    L.dmb_reset.argtypes = [hp, C.POINTER(DmbState), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.dmb_step.argtypes = [hp, C.POINTER(DmbState), C.c_void_p, C.POINTER(DmbStepOut), C.c_void_p]
    L.dmb_get_obs.argtypes = [hp, C.POINTER(DmbState), C.c_void_p, C.c_void_p]
    L.dmb_forward_debug.argtypes = [hp, C.POINTER(DmbState), C.c_void_p, C.c_void_p, C.c_void_p]
    L.dmb_get_trace.argtypes = [hp, C.c_void_p, C.c_int32]
    L.dmb_get_trace.restype = C.c_int32
    L.dmb_obs_dim.argtypes = [hp]
    L.dmb_kernel_launches.argtypes = [hp]
    L.dmb_peer_alloc.argtypes = [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]
    L.dmb_peer_open.argtypes = [C.c_int32, C.c_char_p, C.POINTER(C.c_void_p)]
    L.dmb_peer_close.argtypes = [C.c_void_p]
    L.dmb_peer_free.argtypes = [C.c_void_p]
    L.dmb_set_peer_gather.argtypes = [hp, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32]
    L.dmb_peer_wait.argtypes = [hp, C.c_void_p, C.c_int32, C.c_void_p]
    L.dmb_set_peer_wait.argtypes = [hp, C.c_void_p, C.c_int32]
    L.dmb_host_alloc.argtypes = [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.dmb_host_free.argtypes = [C.c_void_p]
    L.dmb_host_device_pointer.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
    L.dmb_kernel_launches.restype = C.c_int64
    L.dmb_obs_dim.restype = C.c_int32
    L.dmb_mocap_sample.argtypes = [hp, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dmb_debug_stride.restype = C.c_int32
    L.dmb_debug_offset.argtypes = [C.c_char_p]
    L.dmb_debug_offset.restype = C.c_int32
    L.dmb_launch_info.argtypes = [hp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32)]
    L.dmb_gae.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dmb_last_error.argtypes = [hp]
    L.dmb_last_error.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int, handle=None, what: str = "") -> None:
    if rc != 0:
        msg = load().dmb_last_error(handle)
        raise DmbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
