"""ctypes mirrors of the POD structs in include/dmb_model.h + packing from ModelTables.

The structs cross the C-ABI by pointer (``dmb_create`` in include/dmb.h, ``dmo_*`` in
oracle/dm_oracle.h); field order and capacities here must match the header exactly --
``tests/test_abi.py`` checks ``sizeof`` against both shared libraries.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .mjcf import ModelTables

MAX_BODY, MAX_JNT, MAX_DOF, MAX_Q, MAX_GEOM, MAX_PAIR, MAX_U, MAX_M, MAX_CLIP, MAX_EE, MAX_PART = (
    16, 32, 40, 40, 16, 128, 32, 320, 16, 4, 16)
REF_AUX = 24

i32, f64 = C.c_int32, C.c_double

# DeepMimic PD gains and joint weights (reference: src/mujoco/mocap_util.py:22-29)
PARAMS_KP_KD = {"chest": (1000, 100), "neck": (100, 10), "right_shoulder": (400, 40), "right_elbow": (300, 30),
                "left_shoulder": (400, 40), "left_elbow": (300, 30), "right_hip": (500, 50), "right_knee": (500, 50),
                "right_ankle": (400, 40), "left_hip": (500, 50), "left_knee": (500, 50), "left_ankle": (400, 40)}
JOINT_WEIGHT = {"root": 1, "chest": 0.5, "neck": 0.3, "right_hip": 0.5, "right_knee": 0.3, "right_ankle": 0.2,
                "right_shoulder": 0.3, "right_elbow": 0.2, "right_wrist": 0.0, "left_hip": 0.5, "left_knee": 0.3,
                "left_ankle": 0.2, "left_shoulder": 0.3, "left_elbow": 0.2, "left_wrist": 0.0}
# DeepMimic end effectors (src/data/characters/humanoid3d.txt IsEndEffector): ankles = ankle
# joint origins, wrists = wrist geom centres on the elbow bodies (dp_env_v3.xml:51,65)
END_EFFECTORS = (("right_ankle", (0.0, 0.0, 0.0)), ("right_elbow", (0.0, 0.0, -0.258947)),
                 ("left_ankle", (0.0, 0.0, 0.0)), ("left_elbow", (0.0, 0.0, -0.258947)))


# DeepMimic body parts in the order of src/data/characters/humanoid3d.txt: (owning body, index of the
# part's geom among that body's geoms) -- the wrists are the second geoms of the elbow bodies
DM_PARTS = (("root", 0), ("chest", 0), ("neck", 0), ("right_hip", 0), ("right_knee", 0), ("right_ankle", 0),
            ("right_shoulder", 0), ("right_elbow", 0), ("right_elbow", 1), ("left_hip", 0), ("left_knee", 0),
            ("left_ankle", 0), ("left_shoulder", 0), ("left_elbow", 0), ("left_elbow", 1))


class DmbModel(C.Structure):
    _fields_ = [
        ("nq", i32), ("nv", i32), ("nu", i32), ("nbody", i32), ("njnt", i32), ("ngeom", i32), ("npair", i32), ("nM", i32),
        ("iterations", i32), ("max_con", i32), ("max_efc", i32), ("pad0", i32),
        ("timestep", f64), ("tolerance", f64), ("meaninertia", f64), ("margin", f64),
        ("gravity", f64 * 3), ("solref", f64 * 2), ("solimp", f64 * 5),
        ("body_parent", i32 * MAX_BODY), ("body_depth", i32 * MAX_BODY),
        ("body_jntadr", i32 * MAX_BODY), ("body_jntnum", i32 * MAX_BODY),
        ("body_dofadr", i32 * MAX_BODY), ("body_dofnum", i32 * MAX_BODY),
        ("body_pos", (f64 * 3) * MAX_BODY), ("body_quat", (f64 * 4) * MAX_BODY),
        ("body_ipos", (f64 * 3) * MAX_BODY), ("body_inertia", (f64 * 6) * MAX_BODY),
        ("body_mass", f64 * MAX_BODY), ("body_invweight0", (f64 * 2) * MAX_BODY),
        ("jnt_type", i32 * MAX_JNT), ("jnt_bodyid", i32 * MAX_JNT),
        ("jnt_qposadr", i32 * MAX_JNT), ("jnt_dofadr", i32 * MAX_JNT), ("jnt_limited", i32 * MAX_JNT),
        ("jnt_axis", (f64 * 3) * MAX_JNT), ("jnt_range", (f64 * 2) * MAX_JNT),
        ("dof_bodyid", i32 * MAX_DOF), ("dof_jntid", i32 * MAX_DOF),
        ("dof_parentid", i32 * MAX_DOF), ("dof_Madr", i32 * MAX_DOF),
        ("dof_armature", f64 * MAX_DOF), ("dof_damping", f64 * MAX_DOF), ("dof_invweight0", f64 * MAX_DOF),
        ("geom_type", i32 * MAX_GEOM), ("geom_bodyid", i32 * MAX_GEOM), ("geom_condim", i32 * MAX_GEOM),
        ("geom_size", (f64 * 3) * MAX_GEOM), ("geom_pos", (f64 * 3) * MAX_GEOM), ("geom_quat", (f64 * 4) * MAX_GEOM),
        ("geom_rbound", f64 * MAX_GEOM), ("geom_friction", (f64 * 3) * MAX_GEOM),
        ("pair_geom1", i32 * MAX_PAIR), ("pair_geom2", i32 * MAX_PAIR),
        ("act_dofadr", i32 * MAX_U), ("act_gear", f64 * MAX_U), ("act_ctrlrange", (f64 * 2) * MAX_U),
        ("act_kp", f64 * MAX_U), ("act_kd", f64 * MAX_U),
        ("qpos0", f64 * MAX_Q),
        ("dof_weight", f64 * MAX_DOF),
        ("ee_body", i32 * MAX_EE), ("nee", i32), ("pad1", i32 * 3),
        ("ee_pos", (f64 * 3) * MAX_EE),
        ("npart", i32), ("pad2", i32), ("part_geom", i32 * MAX_PART),
    ]


class DmbConfig(C.Structure):
    _fields_ = [
        ("ctrl_mode", i32), ("reward_mode", i32), ("reset_mode", i32), ("auto_reset", i32),
        ("term_mode", i32), ("fall_body_mask", C.c_uint32), ("phase_mode", i32), ("obs_mode", i32),
        ("z_min", f64), ("z_max", f64), ("reset_noise", f64), ("joint_weight_sum", f64),
        ("w_pose", f64), ("w_vel", f64), ("w_end_eff", f64), ("w_root", f64), ("w_com", f64),
        ("s_pose", f64), ("s_vel", f64), ("s_end_eff", f64), ("s_root", f64), ("s_com", f64), ("s_err", f64),
    ]


class DmbMocap(C.Structure):
    _fields_ = [
        ("nclip", i32), ("nframe_total", i32),
        ("clip_start", i32 * MAX_CLIP), ("clip_len", i32 * MAX_CLIP), ("clip_dt", f64 * MAX_CLIP),
        ("data_config", C.POINTER(f64)), ("data_vel", C.POINTER(f64)), ("ref_aux", C.POINTER(f64)),
    ]


def _fill(dst, src):
    """Copy a numpy array into a (possibly nested) ctypes array prefix."""
    a = np.asarray(src)
    if a.ndim == 1:
        for i, v in enumerate(a):
            dst[i] = v.item()
    else:
        for i in range(a.shape[0]):
            _fill(dst[i], a[i])


def pack_model(mt: ModelTables, max_con: int = 16, max_efc: int = 40) -> DmbModel:
    """ModelTables -> dmb_model_t (adds PD gains, reward weights and end-effector points)."""
    for cap, n, what in ((MAX_BODY, mt.nbody, "bodies"), (MAX_JNT, mt.njnt, "joints"), (MAX_DOF, mt.nv, "dofs"),
                         (MAX_Q, mt.nq, "qpos"), (MAX_GEOM, mt.ngeom, "geoms"), (MAX_PAIR, mt.npair, "pairs"),
                         (MAX_U, mt.nu, "actuators"), (MAX_M, mt.nM, "inertia entries")):
        if n > cap:
            raise ValueError(f"model has {n} {what}, capacity is {cap}")
    m = DmbModel()
    for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "npair", "nM", "iterations"):
        setattr(m, k, int(getattr(mt, k)))
    m.max_con, m.max_efc = int(max_con), int(max_efc)
    m.timestep, m.tolerance, m.meaninertia, m.margin = mt.timestep, mt.tolerance, mt.meaninertia, mt.margin
    for k in ("gravity", "solref", "solimp", "body_parent", "body_depth", "body_jntadr", "body_jntnum", "body_dofadr",
              "body_dofnum", "body_pos", "body_quat", "body_ipos", "body_inertia", "body_mass", "body_invweight0",
              "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "jnt_axis", "jnt_range",
              "dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr", "dof_armature", "dof_damping", "dof_invweight0",
              "geom_type", "geom_bodyid", "geom_condim", "geom_size", "geom_pos", "geom_quat", "geom_rbound",
              "geom_friction", "pair_geom1", "pair_geom2", "act_dofadr", "act_gear", "act_ctrlrange", "qpos0"):
        _fill(getattr(m, k), getattr(mt, k))
    # PD gains per actuator and normalised joint weights per dof, keyed by body name
    wsum = float(sum(JOINT_WEIGHT.values()))
    for u in range(mt.nu):
        body = mt.body_names[mt.dof_bodyid[mt.act_dofadr[u]]]
        kp, kd = PARAMS_KP_KD.get(body, (0.0, 0.0))
        m.act_kp[u], m.act_kd[u] = float(kp), float(kd)
    for d in range(mt.nv):
        body = mt.body_names[mt.dof_bodyid[d]]
        m.dof_weight[d] = JOINT_WEIGHT.get(body, 0.0) / wsum
    nee = 0
    for name, off in END_EFFECTORS:
        if name in mt.body_names:
            m.ee_body[nee] = mt.body_names.index(name)
            for k in range(3):
                m.ee_pos[nee][k] = off[k]
            nee += 1
    m.nee = nee
    # DeepMimic body parts (humanoid3d.txt order) -> geom ids: the k-th geom of the named body
    npart = 0
    for body, k in DM_PARTS:
        if body in mt.body_names:
            b = mt.body_names.index(body)
            gs = [g for g in range(mt.ngeom) if mt.geom_bodyid[g] == b]
            if k < len(gs):
                m.part_geom[npart] = gs[k]
                npart += 1
    m.npart = npart
    return m


def default_config(**kw) -> DmbConfig:
    """Defaults = the live reference env (dp_env_v3.py:42-53,106-139)."""
    c = DmbConfig()
    c.ctrl_mode, c.reward_mode, c.reset_mode, c.auto_reset = 0, 0, 0, 0
    c.term_mode = 0
    c.phase_mode, c.obs_mode = 0, 0
    # DeepMimic walk args: every body except the two ankles (bodies 10 and 13 of dp_env_v3.xml) is a fall contact
    c.fall_body_mask = sum(1 << b for b in range(1, 14) if b not in (10, 13))
    c.z_min, c.z_max, c.reset_noise = 0.7, 2.0, 0.01
    c.joint_weight_sum = float(sum(JOINT_WEIGHT.values()))
    c.w_pose, c.w_vel, c.w_end_eff, c.w_root, c.w_com = 0.5, 0.05, 0.15, 0.2, 0.1
    c.s_pose, c.s_vel, c.s_end_eff, c.s_root, c.s_com, c.s_err = 2.0, 0.1, 40.0, 5.0, 10.0, 1.0
    for k, v in kw.items():
        if not hasattr(c, k):
            raise TypeError(f"unknown config field {k}")
        setattr(c, k, v)
    return c
