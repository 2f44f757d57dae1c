"""Mocap compiler vs golden vectors produced by the REFERENCE loader (tests/golden/make_mocap_golden.py)
and the coarse App. C pins of SURVEY.md -- CPU only."""
import os

import numpy as np
import pytest

import common
from deepmimic_mujoco_b200 import mocap

CLIPS = ("walk", "spinkick", "dance_b", "run", "backflip")


@pytest.mark.parametrize("name", CLIPS)
def test_against_reference_loader(name):
    g = np.load(os.path.join(common.GOLDEN, f"mocap_{name}.npz"))
    c = common.clip(name)
    assert c.dt == float(g["dt"])
    assert c.data_config.shape == g["data_config"].shape and c.data_vel.shape == g["data_vel"].shape
    assert np.abs(c.data - g["data"]).max() < 1e-12
    assert np.abs(c.data_config - g["data_config"]).max() < 1e-12
    assert np.abs(c.data_vel - g["data_vel"]).max() < 1e-10
    assert np.all(c.data_vel[0] == 0.0)  # frame 0 has zero velocity (mocap_v2.py:100,110)


OTHER_CLIPS = ("cartwheel", "crawl", "dance_a", "getup_facedown", "getup_faceup", "jump", "kick", "punch", "roll", "spin")


@pytest.mark.parametrize("name", OTHER_CLIPS)
def test_remaining_clips_against_reference_loader_digest(name):
    """The other ten clips the reference ships (src/mujoco/motions/): every 5th frame and the last frame of the
    reference loader's data_config / data_vel, and float64 column sums over all frames (mocap_digest.npz)."""
    g = np.load(os.path.join(common.GOLDEN, "mocap_digest.npz"))
    c = common.clip(name)
    assert c.dt == float(g[name + "/dt"]) and len(c) == int(g[name + "/nframes"])
    rows = g[name + "/rows"]
    assert np.abs(c.data_config[rows] - g[name + "/cfg_rows"]).max() < 1e-12
    assert np.abs(c.data_vel[rows] - g[name + "/vel_rows"]).max() < 1e-10
    assert np.abs(c.data_config.sum(0) - g[name + "/cfg_colsum"]).max() < 1e-10
    assert np.abs(np.abs(c.data_config).sum(0) - g[name + "/cfg_abssum"]).max() < 1e-10
    assert np.abs(c.data_vel.sum(0) - g[name + "/vel_colsum"]).max() < 1e-8


def test_survey_appendix_c_pins():
    walk0 = [0, -0, 0.847532, 0.998678, 0.014104, 0.049423, -0.000698, 0.019375, 0.008037255, -0.09523903, -0, 0, -0,
             -0.1555353, 0.2391943, 0.2073966, 0.170571, 0.3529632, -0.2610683, -0.2456053, 0.581348, 0.02035205,
             -0.5175742, -0.1137634, -0.249116, 0.02055624, -0.0195345, 0.06552698, -0.0560635, 0.1520958, 0.1827421,
             -0.391532, 0.1931168, -0.2978919, -0.08305715]
    assert np.abs(common.clip("walk").data_config[0] - walk0).max() < 2e-6
    sk0 = [0, -0, 0.825094, -0.996905, 0.045699, -0.063904, -0.002929, -0.039919, 0.062062, -0.092066, 0.092995,
           0.423033, 0.115536, -1.442533, 0.470893, -0.454991, 1.122763, 1.307129, -0.762477, -0.247515, 1.445205,
           -0.105896, -0.480345, 0.046555, -0.811415, 0.088775, -0.345502, -0.173389, 0.203981, -0.155389, 0.384096,
           -0.626824, -0.587719, -0.557765, -0.112269]
    assert np.abs(common.clip("spinkick").data_config[0] - sk0).max() < 2e-6
    lens = dict(backflip=29, cartwheel=164, crawl=177, dance_a=98, dance_b=153, getup_facedown=183, getup_faceup=227,
                jump=107, kick=47, punch=65, roll=121, run=25, spin=107, spinkick=78, walk=39)
    for k, n in lens.items():
        assert len(common.clip(k)) == n


def test_transformations_known_answers():
    k = np.load(os.path.join(common.GOLDEN, "transformations_kat.npz"))
    # transformations.py:1092-1093 doctest
    e = mocap.euler_rxyz_from_quat(np.array([0.99810947, 0.06146124, 0, 0]))
    assert np.allclose(e, [0.123, 0, 0], atol=1e-8)
    assert np.abs(mocap.euler_rxyz_from_quat(k["rand_quat_wxyz"]) - k["rand_euler_rxyz"]).max() < 1e-12
    # quaternion_multiply doctest (xyzw): [1,-2,3,4]*[-5,6,7,8] = [-44,-14,48,28]
    a = np.array([4.0, 1, -2, 3]); b = np.array([8.0, -5, 6, 7])
    r = mocap.qmul(a, b)
    assert np.allclose([r[1], r[2], r[3], r[0]], k["quaternion_multiply"])


def test_euler_roundtrip_matches_hinge_stacking():
    """The Euler triple must reproduce the joint rotation as R = Rx Ry Rz (dp_env_v3.xml hinge order)."""
    rng = np.random.default_rng(3)
    q = rng.normal(size=(32, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    e = mocap.euler_rxyz_from_quat(q)
    def rot(axis, a):
        c, s = np.cos(a), np.sin(a)
        m = np.eye(3); i, j = [(1, 2), (0, 2), (0, 1)][axis]
        m[i, i] = c; m[j, j] = c; m[i, j] = -s if axis != 1 else s; m[j, i] = s if axis != 1 else -s
        return m
    R = mocap.quat_to_matrix_xyzw_style(q)
    for k in range(32):
        assert np.allclose(rot(0, e[k, 0]) @ rot(1, e[k, 1]) @ rot(2, e[k, 2]), R[k], atol=1e-12)


def test_pyquaternion_semantics():
    # angle is wrapped to (-pi, pi]; a negative-w quaternion gives the shortest arc through the wrap
    q = np.array([np.cos(2.0), np.sin(2.0), 0, 0])  # rotation of 4 rad about x
    assert np.isclose(mocap.quat_angle(q), 4.0 - 2 * np.pi)
    assert np.allclose(mocap.quat_axis(np.array([1.0, 0, 0, 0])), 0.0)  # null rotation -> zero axis, no NaN
    assert np.allclose(mocap.align_position(np.array([1.0, 2.0, 3.0])), [1, -3, 2])
    # align_rotation maps a rotation about y-up onto the same rotation about z-up
    a = 0.3
    qy = np.array([np.cos(a / 2), 0, np.sin(a / 2), 0])
    assert np.allclose(mocap.align_rotation(qy), [np.cos(a / 2), 0, 0, np.sin(a / 2)])


def test_concat_and_ragged_inputs():
    mc = mocap.concat_clips([common.clip("walk"), common.clip("dance_b"), common.clip("spinkick")])
    assert mc.clip_start.tolist() == [0, 39, 192] and mc.clip_len.tolist() == [39, 153, 78]
    assert mc.data_config.shape == (270, 35) and mc.data_vel.shape == (270, 34)
    one = mocap.compile_frames(np.load(os.path.join(common.ASSETS, "motions", "walk.npz"))["frames"][:1])
    assert one.data_config.shape == (1, 35) and np.all(one.data_vel == 0)
    with pytest.raises(ValueError):
        mocap.compile_frames(np.zeros((0, 44)))
    with pytest.raises(ValueError):
        mocap.compile_frames(np.zeros((3, 43)))
