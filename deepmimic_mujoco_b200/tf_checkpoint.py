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

:func:`write_checkpoint` is the other direction (a policy trained on the CUDA path handed back to the reference's
``--task evaluate`` / ``--pretrained_weight_path``): it lays the files out the way TensorFlow's ``BundleWriter`` over the
LevelDB table builder does -- keys sorted, restart point every 16 entries, one data block, an empty metaindex block, an
index block keyed by the shortest successor of the last key, masked CRC32C in every block trailer and every entry --
and is checked by regenerating the reference's own shipped ``.index`` and ``.data`` files byte for byte from the
tensors read out of them (tests/test_tf_checkpoint.py).
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


def _crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _crc_table()


def crc32c(b: bytes, crc: int = 0) -> int:
    """CRC-32C (Castagnoli), the checksum of LevelDB tables and tensor-bundle entries."""
    crc ^= 0xFFFFFFFF
    tab = _CRC_TABLE
    for x in b:
        crc = tab[(crc ^ x) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(b: bytes) -> int:
    c = crc32c(b)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _vi(x: int) -> bytes:
    out = bytearray()
    while True:
        c = x & 0x7F
        x >>= 7
        out.append(c | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _field(num: int, wire: int, payload: bytes) -> bytes:
    return _vi((num << 3) | wire) + (_vi(len(payload)) + payload if wire == 2 else payload)


def _table_block(entries, restart_interval: int = 16) -> bytes:
    blk, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(blk))
        else:
            while shared < min(len(k), len(prev)) and k[shared] == prev[shared]:
                shared += 1
        blk += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        prev = k
    for r in restarts or [0]:
        blk += struct.pack("<I", r)
    return bytes(blk + struct.pack("<I", len(restarts or [0])))


def _short_successor(key: bytes) -> bytes:
    """LevelDB BytewiseComparator::FindShortSuccessor: cut after the first byte that can be incremented."""
    for i, c in enumerate(key):
        if c != 0xFF:
            return key[:i] + bytes([c + 1])
    return key


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write ``tensors`` (name -> float32 / float64 / int32 / int64 array) as a TensorFlow V2 checkpoint
    ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` that ``tf.train.Saver().restore`` (the reference's
    ``U.load_state``) reads; see the module docstring for the layout."""
    enum = {np.dtype(v): k for k, v in _DTYPES.items()}
    data = bytearray()
    header = _field(1, 0, _vi(1)) + _field(3, 2, _field(1, 0, _vi(1)))     # num_shards 1, version {producer 1}
    entries = [(b"", header)]
    for name in sorted(tensors, key=lambda n: n.encode()):
        a = np.asarray(tensors[name], order="C")
        if a.dtype not in enum:
            raise CheckpointError(f"variable {name}: dtype {a.dtype} is not supported")
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        e = _field(1, 0, _vi(enum[a.dtype])) + _field(2, 2, b"".join(_field(2, 2, _field(1, 0, _vi(d))) for d in a.shape))
        if len(data):
            e += _field(4, 0, _vi(len(data)))
        e += _field(5, 0, _vi(len(raw))) + _field(6, 5, struct.pack("<I", masked_crc32c(raw)))
        entries.append((name.encode(), e))
        data += raw
    out = bytearray()

    def emit(block: bytes):
        off = len(out)
        out.extend(block + b"\0" + struct.pack("<I", masked_crc32c(block + b"\0")))   # type 0 = no compression
        return _vi(off) + _vi(len(block))

    d_handle = emit(_table_block(entries))
    m_handle = emit(_table_block([]))
    i_handle = emit(_table_block([(_short_successor(entries[-1][0]), d_handle)]))
    foot = m_handle + i_handle
    out.extend(foot + b"\0" * (40 - len(foot)) + struct.pack("<Q", TABLE_MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


# MlpPolicy parameter name <- reference variable name inside the policy scope (mlp_policy_trpo.py:35-46)
_POLICY_VARS = dict(pw1="polfc1/w", pb1="polfc1/b", pw2="polfc2/w", pb2="polfc2/b", pw3="polfinal/w", pb3="polfinal/b",
                    vw1="vffc1/w", vb1="vffc1/b", vw2="vffc2/w", vb2="vffc2/b", vw3="vffinal/w", vb3="vffinal/b",
                    logstd="logstd")


def policy_tensors(arrays: Dict[str, np.ndarray], scope: str = "pi") -> Dict[str, np.ndarray]:
    """Inverse of :func:`policy_arrays`: MlpPolicy-named host arrays -> the reference's variables of one scope."""
    out = {f"{scope}/{theirs}": np.asarray(arrays[mine], dtype=np.float32) for mine, theirs in _POLICY_VARS.items()}
    out[f"{scope}/logstd"] = out[f"{scope}/logstd"].reshape(1, -1)          # mlp_policy_trpo.py:46: shape [1, act_dim]
    for mine, theirs in (("ob_sum", "runningsum"), ("ob_sumsq", "runningsumsq"), ("ob_count", "count")):
        out[f"{scope}/obfilter/{theirs}"] = np.asarray(arrays[mine], dtype=np.float64)
    return out


def write_checkpoint_state(prefix: str) -> None:
    """The ``checkpoint`` file tf.train.Saver writes next to the data (for ``tf.train.latest_checkpoint``)."""
    import os
    name = os.path.basename(prefix)
    with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as f:
        f.write(f'model_checkpoint_path: "{name}"\nall_model_checkpoint_paths: "{name}"\n')


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
