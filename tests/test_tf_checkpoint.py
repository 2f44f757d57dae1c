"""deepmimic_mujoco_b200/tf_checkpoint.py: reading the reference's saved policies (tf.train.Saver V2 checkpoints,
utils/tf_util.py:314-329, trpo.py:207-208 / 220-224 / 367) without TensorFlow.

* write_checkpoint regenerates the reference's shipped .index / .data / checkpoint files BYTE FOR BYTE from the tensors
  read out of them (table layout, restart points, shortest-successor index key, masked CRC32C of blocks and entries);
* a bundle written by write_checkpoint round-trips through the reader;
* the checkpoint the reference ships (checkpoint_tmp/DeepMimic/trpo-walk-0) reads to exactly the tensors committed
  in tests/golden/ref_trained_policy.npz (build container only: /root/reference is not on the GPU box);
* the variable-name mapping onto policy.MlpPolicy and its shape checks; malformed files raise CheckpointError."""
import os

import numpy as np
import pytest
import torch

import common
from deepmimic_mujoco_b200 import tf_checkpoint as tfc

REF_CKPT = "/root/reference/src/checkpoint_tmp/DeepMimic/trpo-walk-0/DeepMimic/trpo-walk-0"


def write_bundle(prefix, tensors, compressed_tag=0):
    """The package's writer; ``compressed_tag`` overwrites the data block's compression byte to exercise the reader."""
    tfc.write_checkpoint(prefix, tensors)
    if compressed_tag:
        raw = bytearray(open(prefix + ".index", "rb").read())
        foot = raw[-48:]
        _, p = tfc._varint(foot, 0); _, p = tfc._varint(foot, p)          # metaindex handle
        ioff, p = tfc._varint(foot, p); isz, p = tfc._varint(foot, p)      # index handle
        (_, handle), = tfc._block(bytes(raw), ioff, isz)                   # one data block: (offset, size)
        off, q = tfc._varint(handle, 0); sz, q = tfc._varint(handle, q)
        raw[off + sz] = compressed_tag                                      # the block's type byte
        open(prefix + ".index", "wb").write(bytes(raw))


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


@pytest.mark.skipif(not os.path.exists(REF_CKPT + ".index"), reason="reference checkout not present (GPU box)")
def test_writer_regenerates_the_reference_checkpoint_byte_for_byte(tmp_path):
    r = tfc.read_checkpoint(REF_CKPT)
    q = str(tmp_path / "trpo-walk-0")
    tfc.write_checkpoint(q, r)
    tfc.write_checkpoint_state(q)
    for ext in (".index", ".data-00000-of-00001"):
        assert open(q + ext, "rb").read() == open(REF_CKPT + ext, "rb").read(), ext
    assert open(str(tmp_path / "checkpoint")).read() == open(os.path.join(os.path.dirname(REF_CKPT), "checkpoint")).read()
    # and through the policy-level mapping: arrays -> scopes "pi" (the golden's) + "oldpi" -> same names and shapes
    a = tfc.policy_arrays(r, "pi")
    t = {**tfc.policy_tensors(a, "pi"), **tfc.policy_tensors(a, "oldpi")}
    assert sorted(t) == sorted(r)
    for k in r:
        assert t[k].shape == r[k].shape and t[k].dtype == r[k].dtype, k
        if k.startswith("pi/"):
            assert np.array_equal(t[k], r[k]), k


def test_crc32c_known_answers():
    assert tfc.crc32c(b"123456789") == 0xE3069283                         # the CRC-32C check value
    assert tfc.crc32c(b"") == 0 and tfc.crc32c(bytes(32)) == 0x8A9136AA   # RFC 3720 B.4: 32 bytes of zeros


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
    h = pol.host_arrays()                                                  # and back out: save -> read -> same arrays
    import tempfile
    q = os.path.join(tempfile.mkdtemp(), "saved")
    pol.save_tf_checkpoint(q)
    back = tfc.read_checkpoint(q)
    assert sorted(back) == sorted({**tfc.policy_tensors(h, "pi"), **tfc.policy_tensors(h, "oldpi")})
    for sc in ("pi", "oldpi"):
        b = tfc.policy_arrays(back, sc)
        for k in a:
            assert np.array_equal(np.asarray(b[k]), np.asarray(a[k])), (sc, k)
