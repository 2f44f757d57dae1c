"""MJCF compiler pins -- CPU only."""
import os

import numpy as np
import pytest

import common
from deepmimic_mujoco_b200 import mjcf


def test_sizes_and_mass():
    mt = common.tables()
    assert (mt.nq, mt.nv, mt.nu, mt.nbody, mt.njnt, mt.ngeom) == (35, 34, 28, 14, 29, 16)
    assert mt.nM == 310 and mt.npair == 104
    assert abs(mt.total_mass() - 45.0) < 1e-12
    assert mt.timestep == 0.0166 and mt.iterations == 50 and mt.margin == 0.001
    assert (mt.geom_type[mt.pair_geom1] == mjcf.GEOM_PLANE).sum() == 15


def test_tree_and_actuators():
    mt = common.tables()
    assert mt.body_names == ["world", "root", "chest", "neck", "right_shoulder", "right_elbow", "left_shoulder",
                             "left_elbow", "right_hip", "right_knee", "right_ankle", "left_hip", "left_knee", "left_ankle"]
    assert mt.body_parent.tolist() == [-1, 0, 1, 2, 2, 4, 2, 6, 1, 8, 9, 1, 11, 12]
    # actuator i drives qpos[7+i] (control_test.py:53-85)
    assert mt.act_dofadr.tolist() == list(range(6, 34))
    gears = dict(chest=200, neck=50, right_shoulder=100, right_elbow=60, right_hip=200, right_knee=150, right_ankle=90)
    for u in range(mt.nu):
        body = mt.body_names[mt.dof_bodyid[mt.act_dofadr[u]]]
        assert mt.act_gear[u] == gears[body.replace("left", "right")]
    assert np.all(mt.act_ctrlrange == np.array([-0.5, 0.5]))
    # elbows / knees rotate about -y
    for name in ("right_elbow", "left_elbow", "right_knee", "left_knee"):
        assert mt.jnt_axis[mt.joint_names.index(name)].tolist() == [0, -1, 0]
    assert mt.jnt_range[mt.joint_names.index("right_knee")].tolist() == [-2.7, 0.0]


def test_inertials():
    mt = common.tables()
    # sphere: 2/5 m r^2 ; box: m/3 (b^2 + c^2)
    assert np.allclose(mt.body_inertia[1, :3], 0.4 * 6.0 * 0.09 ** 2)
    a, b, c = 0.0885, 0.045, 0.0275
    assert np.allclose(mt.body_inertia[10, :3], [(b * b + c * c) / 3, (a * a + c * c) / 3, (a * a + b * b) / 3])
    # elbow body = capsule (1.0 kg @ z=-0.12) + wrist sphere (0.5 kg @ z=-0.258947)
    assert np.allclose(mt.body_ipos[5], [0, 0, (1.0 * -0.12 + 0.5 * -0.258947) / 1.5])
    assert abs(mt.body_mass[5] - 1.5) < 1e-12
    # capsule inertia against numerical integration of a uniform capsule
    r, h, m = 0.045, 0.18, 1.5
    rng = np.random.default_rng(0)
    pts = rng.uniform([-r, -r, -h / 2 - r], [r, r, h / 2 + r], size=(400000, 3))
    zc = np.clip(pts[:, 2], -h / 2, h / 2)
    inside = pts[:, 0] ** 2 + pts[:, 1] ** 2 + (pts[:, 2] - zc) ** 2 <= r * r
    p = pts[inside]
    ixx = m * np.mean(p[:, 1] ** 2 + p[:, 2] ** 2)
    izz = m * np.mean(p[:, 0] ** 2 + p[:, 1] ** 2)
    assert abs(mt.body_inertia[4, 0] - ixx) / ixx < 0.01 and abs(mt.body_inertia[4, 2] - izz) / izz < 0.01


def test_mass_matrix_properties():
    mt = common.tables()
    rng = np.random.default_rng(1)
    q, _ = common.airborne_states(rng, 1)
    M = mjcf.np_mass_matrix(mt, q[0])
    assert np.allclose(M, M.T) and np.all(np.linalg.eigvalsh(M) > 0)
    # translational block = total mass
    assert np.allclose(M[:3, :3], 45.0 * np.eye(3))
    assert abs(mt.meaninertia - np.trace(mjcf.np_mass_matrix(mt, mt.qpos0)) / mt.nv) < 1e-12


@pytest.mark.skipif(not os.path.exists(common.REFERENCE_XML), reason="reference tree not present (GPU box)")
def test_shipped_asset_matches_reference_xml():
    ref = mjcf.compile_mjcf(common.REFERENCE_XML)
    mt = common.tables()
    for f in ("body_pos", "body_inertia", "body_mass", "jnt_range", "geom_size", "pair_geom1", "pair_geom2",
              "body_invweight0", "dof_invweight0", "act_gear", "dof_Madr"):
        assert np.array_equal(getattr(ref, f), getattr(mt, f)), f
