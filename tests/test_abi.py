"""C-ABI checks that need no GPU: the library loads, exports every symbol of include/dmb.h,
struct sizes agree with the ctypes mirrors, and the product path fails loudly without CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

import common
from deepmimic_mujoco_b200 import lib
from deepmimic_mujoco_b200.model_blob import default_config
from deepmimic_mujoco_b200.sim import load_motions, make_mocap_struct


def test_library_exports_header_symbols():
    L = lib.load()
    hdr = open(os.path.join(common.ROOT, "include", "dmb.h")).read()
    declared = set(re.findall(r"\b(dmb_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"dmb_status"}
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.dmb_version() == 2
    phdr = open(os.path.join(common.ROOT, "include", "dmb_policy.h")).read()
    assert set(re.findall(r"\b(dmb_(?:policy_[a-z_0-9]+|gae))\s*\(", phdr)) == set(lib.POLICY_EXPORTS)
    for name in lib.POLICY_EXPORTS:
        assert hasattr(L, name), name
    assert L.dmb_debug_stride() > 0 and L.dmb_debug_offset(b"qacc") > 0 and L.dmb_debug_offset(b"nope") == -1


def test_struct_sizes_match_oracle_and_cuda_lib():
    import oracle.pyoracle as po
    po.lib()      # asserts sizeof(dmo_data_t / dmb_model_t / dmo_env_t) against the ctypes mirrors
    lib.load()    # raises on dmb_sizeof_* mismatch


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    L = lib.load()
    m, c = common.model(), default_config()
    mc, keep = make_mocap_struct(load_motions(["walk"]))
    h = C.c_void_p()
    rc = L.dmb_create(C.byref(m), C.byref(c), C.byref(mc), 4, 0, 0, 0, C.byref(h))
    assert rc == -4 and b"no CPU fallback" in L.dmb_last_error(None)
    from deepmimic_mujoco_b200.sim import BatchedSim
    with pytest.raises(lib.DmbError):
        BatchedSim(4)


def test_bad_arguments_are_rejected():
    L = lib.load()
    h = C.c_void_p()
    assert L.dmb_create(None, None, None, 4, 0, 0, 0, C.byref(h)) == -1
    assert L.dmb_destroy(None) == -1
    assert L.dmb_step(None, None, None, None, None) == -1
