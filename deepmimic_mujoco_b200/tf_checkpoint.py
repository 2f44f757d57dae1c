"""Read the reference's saved policies without TensorFlow.

The reference saves and restores its learner with ``tf.train.Saver`` (/root/reference/src/utils/tf_util.py:314-329
``load_state`` / ``save_state``; called from trpo.py:207-208 ``--pretrained_weight_path``, trpo.py:220-224 the periodic
save, trpo.py:367 ``--load_model_path`` of the evaluate task).  What lands on disk is a TensorFlow "V2" checkpoint
(tensor bundle): ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``.  A user who moves the rollout onto the CUDA
path still has those files, so this module reads them directly:

* ``<prefix>.index`` is a LevelDB-format sorted string table, written uncompressed by the bundle writer: data blocks
  of prefix-compressed entries ``(shared, non_shared, value_len as varints, key suffix, value)`` followed by a restart
  array, each block trailed by a 1-byte compression tag + 4-byte CRC; an index block maps to the data blocks; the last
  48 bytes are the footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).
* every value is a ``BundleEntryProto``: dtype (1), shape (2: TensorShapeProto, dims (2) with size (1)), shard_id (3),
  offset (4), size (5), crc32c (6).  The empty key holds the ``BundleHeaderProto`` and is skipped.
* ``<prefix>.data-00000-of-00001`` holds the tensors back to back, little endian, at ``offset`` / ``size``.

:func:`policy_arrays` maps the variables of one ``MlpPolicy`` scope (mlp_policy_trpo.py:24-60: ``polfc1/w`` ...
``vffinal/b``, ``logstd``, ``obfilter/runningsum|runningsumsq|count``) onto the parameter names of
:class:`deepmimic_mujoco_b200.policy.MlpPolicy`; ``MlpPolicy.load_tf_checkpoint`` puts them on the device.
Host-side, load-time code: nothing here is on the hot path.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}     # DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64


class CheckpointError(ValueError):
    pass


def _varint(b: bytes, p: int):
    x = s = 0
    while True:
        if p >= len(b):
            raise CheckpointError("truncated varint")
        c = b[p]; p += 1
        x |= (c & 0x7F) << s
        if c < 0x80:
            return x, p
        s += 7


def _block(b: bytes, off: int, size: int):
    """Entries of one table block as (key, value) pairs."""
    if off + size + 1 > len(b) or size < 4:
        raise CheckpointError("table block outside the index file")
    if b[off + size] != 0:
        raise CheckpointError("compressed table block: only uncompressed checkpoints (the TensorFlow default) are read")
    blk = b[off: off + size]
    nrestart = struct.unpack("<I", blk[-4:])[0]
    end = size - 4 - 4 * nrestart
    if end < 0:
        raise CheckpointError("bad restart array")
    p, key, out = 0, b"", []
    while p < end:
        shared, p = _varint(blk, p)
        non_shared, p = _varint(blk, p)
        vlen, p = _varint(blk, p)
        key = key[:shared] + blk[p: p + non_shared]; p += non_shared
        out.append((key, blk[p: p + vlen])); p += vlen
    return out


def _proto(b: bytes) -> Dict[int, list]:
    """Flat protobuf decode: {field number: [values]}; varints as int, length-delimited fields as bytes."""
    p, out = 0, {}
    while p < len(b):
        tag, p = _varint(b, p)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, p = _varint(b, p)
        elif wire == 2:
            n, p = _varint(b, p); v = b[p: p + n]; p += n
        elif wire == 5:
            v = struct.unpack("<I", b[p: p + 4])[0]; p += 4
        elif wire == 1:
            v = struct.unpack("<Q", b[p: p + 8])[0]; p += 8
        else:
            raise CheckpointError(f"unsupported protobuf wire type {wire}")
        out.setdefault(field, []).append(v)
    return out


def read_checkpoint(prefix: str) -> Dict[str, np.ndarray]:
    """All variables of the checkpoint ``prefix`` (the path given to ``--load_model_path``) as numpy arrays."""
    with open(prefix + ".index", "rb") as f:
        idx = f.read()
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        data = f.read()
    if len(idx) < 48 or struct.unpack("<Q", idx[-8:])[0] != TABLE_MAGIC:
        raise CheckpointError(f"{prefix}.index is not a TensorFlow V2 checkpoint index")
    foot = idx[-48:]
    _, p = _varint(foot, 0); _, p = _varint(foot, p)                      # metaindex handle
    ioff, p = _varint(foot, p); isz, p = _varint(foot, p)                 # index handle
    tensors: Dict[str, np.ndarray] = {}
    for _, handle in _block(idx, ioff, isz):
        off, q = _varint(handle, 0); sz, q = _varint(handle, q)
        for key, val in _block(idx, off, sz):
            if not key:
                continue                                                  # BundleHeaderProto
            e = _proto(val)
            dtype = e.get(1, [0])[0]
            if dtype not in _DTYPES:
                raise CheckpointError(f"variable {key.decode()}: unsupported dtype enum {dtype}")
            if e.get(3, [0])[0] != 0:
                raise CheckpointError("sharded checkpoint (more than one data file) is not supported")
            shape = [_proto(d).get(1, [0])[0] for d in _proto(e[2][0]).get(2, [])] if 2 in e else []
            o, n = e.get(4, [0])[0], e.get(5, [0])[0]
            want = int(np.prod(shape, dtype=np.int64)) * np.dtype(_DTYPES[dtype]).itemsize
            if n != want or o + n > len(data):
                raise CheckpointError(f"variable {key.decode()}: size {n} at offset {o} does not match shape {shape} "
                                      f"/ the data file ({len(data)} bytes)")
            a = np.frombuffer(data[o: o + n], dtype=_DTYPES[dtype])
            tensors[key.decode()] = a.reshape(shape)
    return tensors


# MlpPolicy parameter name <- reference variable name inside the policy scope (mlp_policy_trpo.py:35-46)
_POLICY_VARS = dict(pw1="polfc1/w", pb1="polfc1/b", pw2="polfc2/w", pb2="polfc2/b", pw3="polfinal/w", pb3="polfinal/b",
                    vw1="vffc1/w", vb1="vffc1/b", vw2="vffc2/w", vb2="vffc2/b", vw3="vffinal/w", vb3="vffinal/b",
                    logstd="logstd")


def policy_arrays(tensors: Dict[str, np.ndarray], scope: str = "pi") -> Dict[str, np.ndarray]:
    """The variables of one reference ``MlpPolicy`` scope under the names of policy.MlpPolicy: float32 ``pw1`` ...
    ``vb3``, ``logstd`` [act_dim], and the float64 observation-filter accumulators ``ob_sum``, ``ob_sumsq`` [obs_dim],
    ``ob_count`` [] (utils/misc_util.py:36-50).  Dense kernels are [in][out] in both."""
    out = {}
    for mine, theirs in _POLICY_VARS.items():
        k = f"{scope}/{theirs}"
        if k not in tensors:
            raise CheckpointError(f"checkpoint has no variable {k!r} (scopes: {sorted({n.split('/')[0] for n in tensors})})")
        out[mine] = np.ascontiguousarray(tensors[k], dtype=np.float32)
    out["logstd"] = out["logstd"].reshape(-1)
    for mine, theirs in (("ob_sum", "runningsum"), ("ob_sumsq", "runningsumsq"), ("ob_count", "count")):
        out[mine] = np.asarray(tensors[f"{scope}/obfilter/{theirs}"], dtype=np.float64)
    obs_dim, hid = out["pw1"].shape
    act_dim = out["pw3"].shape[1]
    expect = dict(pb1=(hid,), pw2=(hid, hid), pb2=(hid,), pw3=(hid, act_dim), pb3=(act_dim,), logstd=(act_dim,),
                  vw1=(obs_dim, hid), vb1=(hid,), vw2=(hid, hid), vb2=(hid,), vw3=(hid, 1), vb3=(1,),
                  ob_sum=(obs_dim,), ob_sumsq=(obs_dim,), ob_count=())
    for k, shp in expect.items():
        if out[k].shape != shp:
            raise CheckpointError(f"{scope}/{k}: shape {out[k].shape}, expected {shp}")
    return out
