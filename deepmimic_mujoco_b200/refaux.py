"""Per-mocap-frame reference features for the 5-term DeepMimic reward (host, load time, numpy).

The original DeepMimic evaluates the kinematic reference character every step
(``cSceneImitate::CalcRewardImitate`` quoted in /root/reference/code.md:979-1146: end-effector
positions relative to the root in the heading frame, CoM velocity).  The reference pose only
depends on the frame index, so these are tabulated once per clip and read by the step kernel:
``aux[f] = [4 end-effector points (12), v_com (3), root quat (4), pad]`` (DMB_REF_AUX = 24).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .mjcf import ModelTables, np_body_jacobian, np_kinematics
from .model_blob import END_EFFECTORS, REF_AUX


def frame_features(mt: ModelTables, qpos: np.ndarray, qvel: np.ndarray) -> np.ndarray:
    xpos, xquat, xmat, xipos, xaxis = np_kinematics(mt, qpos)
    aux = np.zeros(REF_AUX)
    R = xmat[1]
    heading = np.arctan2(R[1, 0], R[0, 0])
    ch, sh = np.cos(heading), np.sin(heading)
    k = 0
    for name, off in END_EFFECTORS:
        b = mt.body_names.index(name)
        w = xpos[b] + xmat[b] @ np.asarray(off)
        rel = np.array([w[0] - xpos[1][0], w[1] - xpos[1][1], w[2]])
        aux[3 * k: 3 * k + 3] = [ch * rel[0] + sh * rel[1], -sh * rel[0] + ch * rel[1], rel[2]]
        k += 1
    p = np.zeros(3)
    for b in range(1, mt.nbody):
        J = np_body_jacobian(mt, xpos, xmat, xaxis, b, xipos[b])
        p += mt.body_mass[b] * (J[0:3] @ qvel)
    aux[12:15] = p / mt.body_mass.sum()
    aux[15:19] = qpos[3:7]
    return aux


def compute_ref_aux(motions: Sequence[str], mt: ModelTables | None = None) -> np.ndarray:
    from .sim import default_model_tables, load_motions
    mt = mt if mt is not None else default_model_tables()
    mc = load_motions(list(motions))
    vel = np.nan_to_num(mc.data_vel, nan=0.0, posinf=0.0, neginf=0.0)
    return np.stack([frame_features(mt, mc.data_config[f], vel[f]) for f in range(mc.data_config.shape[0])])
