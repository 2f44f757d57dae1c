"""deepmimic_mujoco_b200/tf_checkpoint.py: reading the reference's saved policies (tf.train.Saver V2 checkpoints,
utils/tf_util.py:314-329, trpo.py:207-208 / 220-224 / 367) without TensorFlow.

* a bundle written by this test (LevelDB-format table with prefix compression and restart points, BundleEntryProto
  values) round-trips through the reader;
* the checkpoint the reference ships (checkpoint_tmp/DeepMimic/trpo-walk-0) reads to exactly the tensors committed
  in tests/golden/ref_trained_policy.npz (build container only: /root/reference is not on the GPU box);
* the variable-name mapping onto policy.MlpPolicy and its shape checks; malformed files raise CheckpointError."""
import os
import struct

import numpy as np
import pytest
import torch

import common
from deepmimic_mujoco_b200 import tf_checkpoint as tfc

REF_CKPT = "/root/reference/src/checkpoint_tmp/DeepMimic/trpo-walk-0/DeepMimic/trpo-walk-0"


def _vi(x):
    out = bytearray()
    while True:
        c = x & 0x7F
        x >>= 7
        out.append(c | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _pb(field, wire, payload):
    return _vi((field << 3) | wire) + (payload if wire != 2 else _vi(len(payload)) + payload)


def _table_block(entries, restart_every=4):
    blk, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_every == 0:
            restarts.append(len(blk))
        else:
            while shared < min(len(k), len(prev)) and k[shared] == prev[shared]:
                shared += 1
        blk += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        blk += struct.pack("<I", r)
    return bytes(blk + struct.pack("<I", len(restarts)))


def write_bundle(prefix, tensors, compressed_tag=0):
    """Minimal tensor-bundle writer (one data block, one shard), the inverse of tf_checkpoint.read_checkpoint."""
    enum = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}
    data, entries = bytearray(), [(b"", _pb(1, 0, _vi(1)))]                # "" -> BundleHeaderProto{num_shards: 1}
    for name in sorted(tensors):
        a = np.asarray(tensors[name], order="C")              # (ascontiguousarray would turn 0-d into 1-d)
        shape = b"".join(_pb(2, 2, _pb(1, 0, _vi(d))) for d in a.shape)
        e = _pb(1, 0, _vi(enum[a.dtype])) + _pb(2, 2, shape)
        if len(data):
            e += _pb(4, 0, _vi(len(data)))
        e += _pb(5, 0, _vi(a.nbytes)) + _pb(6, 5, struct.pack("<I", 0xDEADBEEF))
        entries.append((name.encode(), e))
        data += a.tobytes()
    blocks = bytearray()
    d = _table_block(entries)
    d_handle = _vi(0) + _vi(len(d)); blocks += d + bytes([compressed_tag]) + b"\0\0\0\0"
    m = _table_block([]); m_handle = _vi(len(blocks)) + _vi(len(m)); blocks += m + b"\0" + b"\0\0\0\0"
    i = _table_block([(b"~", d_handle)]); i_handle = _vi(len(blocks)) + _vi(len(i)); blocks += i + b"\0" + b"\0\0\0\0"
    foot = m_handle + i_handle
    foot += b"\0" * (40 - len(foot)) + struct.pack("<Q", tfc.TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(blocks) + foot)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


def small_policy(rng, scope="pi", obs=5, hid=7, act=3):
    t = {}
    for net, last in (("pol", act), ("vf", 1)):
        t[f"{scope}/{net}fc1/w"] = rng.normal(size=(obs, hid)).astype(np.float32)
        t[f"{scope}/{net}fc1/b"] = rng.normal(size=hid).astype(np.float32)
        t[f"{scope}/{net}fc2/w"] = rng.normal(size=(hid, hid)).astype(np.float32)
        t[f"{scope}/{net}fc2/b"] = rng.normal(size=hid).astype(np.float32)
        t[f"{scope}/{net}final/w"] = rng.normal(size=(hid, last)).astype(np.float32)
        t[f"{scope}/{net}final/b"] = rng.normal(size=last).astype(np.float32)
    t[f"{scope}/logstd"] = rng.normal(size=(1, act)).astype(np.float32)
    t[f"{scope}/obfilter/runningsum"] = rng.normal(size=obs) * 100
    t[f"{scope}/obfilter/runningsumsq"] = rng.uniform(50, 100, size=obs) * 100
    t[f"{scope}/obfilter/count"] = np.float64(100.01)
    return t


def test_bundle_round_trip_and_policy_mapping(tmp_path):
    rng = np.random.default_rng(0)
    t = {**small_policy(rng, "pi"), **small_policy(rng, "oldpi"), "global_step": np.asarray(17, dtype=np.int64),
         "iters": np.arange(6, dtype=np.int32).reshape(2, 3)}
    prefix = str(tmp_path / "model")
    write_bundle(prefix, t)
    r = tfc.read_checkpoint(prefix)
    assert sorted(r) == sorted(t)
    for k in t:
        assert r[k].dtype == np.asarray(t[k]).dtype and r[k].shape == np.shape(t[k]) and np.array_equal(r[k], t[k]), k
    a = tfc.policy_arrays(r, "oldpi")
    assert np.array_equal(a["pw3"], t["oldpi/polfinal/w"]) and np.array_equal(a["vw1"], t["oldpi/vffc1/w"])
    assert a["logstd"].shape == (3,) and a["ob_count"] == 100.01 and a["ob_sum"].dtype == np.float64
    with pytest.raises(tfc.CheckpointError, match="no variable"):
        tfc.policy_arrays(r, "pi2")
    bad = dict(r); bad["pi/polfc2/w"] = np.zeros((7, 6), np.float32)
    with pytest.raises(tfc.CheckpointError, match="shape"):
        tfc.policy_arrays(bad, "pi")


def test_malformed_checkpoints_raise(tmp_path):
    rng = np.random.default_rng(1)
    prefix = str(tmp_path / "m")
    write_bundle(prefix, small_policy(rng), compressed_tag=1)              # snappy-compressed block tag
    with pytest.raises(tfc.CheckpointError, match="compressed"):
        tfc.read_checkpoint(prefix)
    write_bundle(prefix, small_policy(rng))
    raw = open(prefix + ".index", "rb").read()
    with open(prefix + ".index", "wb") as f:
        f.write(raw[:-8] + b"\0" * 8)                                      # wrong magic
    with pytest.raises(tfc.CheckpointError, match="not a TensorFlow"):
        tfc.read_checkpoint(prefix)
    write_bundle(prefix, small_policy(rng))
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.truncate(40)                                                     # data file shorter than the index says
    with pytest.raises(tfc.CheckpointError, match="size"):
        tfc.read_checkpoint(prefix)


@pytest.mark.skipif(not os.path.exists(REF_CKPT + ".index"), reason="reference checkout not present (GPU box)")
def test_reference_checkpoint_reads_to_the_committed_golden():
    r = tfc.read_checkpoint(REF_CKPT)
    assert len(r) == 32 and sum(a.nbytes for a in r.values()) == os.path.getsize(REF_CKPT + ".data-00000-of-00001")
    g = np.load(os.path.join(common.GOLDEN, "ref_trained_policy.npz"))
    for k, v in r.items():
        if k.startswith("pi/"):
            assert np.array_equal(g[k], v), k
    a = tfc.policy_arrays(r, "pi")
    assert a["pw1"].shape == (56, 100) and a["pw3"].shape == (100, 28) and a["vw3"].shape == (100, 1)
    pol = common.RefTrainedPolicy()                                        # the same numbers through the test helper
    assert np.allclose(np.exp(a["logstd"]), pol.act_std)
    # entropy of the diagonal Gaussian = the value the reference logged for the update that preceded the save
    assert abs(float(np.sum(a["logstd"] + 0.5 * np.log(2 * np.pi * np.e))) - 35.71932) < 2e-5


def test_policy_load_arrays_on_host_tensors():
    """MlpPolicy.load_arrays / RunningMeanStd.load (the device copy is plain tensor.copy_): exercised on CPU tensors by
    building the object without its CUDA-only constructor; the network evaluated from the loaded tensors must be the
    one the checkpoint describes."""
    from deepmimic_mujoco_b200.policy import MlpPolicy, RunningMeanStd
    g = np.load(os.path.join(common.GOLDEN, "ref_trained_policy.npz"))
    a = tfc.policy_arrays({k: g[k] for k in g.files if k.startswith("pi/")}, "pi")
    pol = MlpPolicy.__new__(MlpPolicy)
    pol.params = {k: torch.zeros(v.shape) for k, v in a.items() if not k.startswith("ob_")}
    pol.ob_rms = RunningMeanStd((56,), "cpu")
    ptr = pol.params["pw1"].data_ptr()
    pol.load_arrays(a)
    assert pol.params["pw1"].data_ptr() == ptr                             # in place
    ref = common.RefTrainedPolicy()
    assert np.allclose(pol.ob_rms.mean.numpy(), ref.ob_mean, atol=1e-7) and np.allclose(pol.ob_rms.std.numpy(), ref.ob_std, atol=1e-6)
    x = torch.as_tensor(np.random.default_rng(2).normal(size=(16, 56)) * 0.5, dtype=torch.float32)
    h = torch.clamp((x - pol.ob_rms.mean) / pol.ob_rms.std, -5, 5)
    h = torch.tanh(h @ pol.params["pw1"] + pol.params["pb1"]); h = torch.tanh(h @ pol.params["pw2"] + pol.params["pb2"])
    assert np.abs((h @ pol.params["pw3"] + pol.params["pb3"]).numpy() - ref.mean_action(x.numpy())).max() < 1e-4
    bad = dict(a); bad["pw2"] = np.zeros((100, 99), np.float32)
    with pytest.raises(ValueError, match="pw2"):
        pol.load_arrays(bad)
