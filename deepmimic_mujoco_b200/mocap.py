"""Mocap compiler: DeepMimic motion JSON -> per-frame qpos / qvel tables (float64, host, load time).

Restates the reference loader
  /root/reference/src/mujoco/mocap_v2.py:20-149   MocapDM.load_mocap / read_raw_data / convert_raw_data
  /root/reference/src/mujoco/mocap_util.py:5-48    joint orders, DOF table, align_rotation / align_position
  /root/reference/src/transformations.py:1031-1097,1174-1193  euler_from_quaternion(q_xyzw, 'rxyz')
and the pyquaternion semantics it relies on (un-vendored third party, unpinned in README.md:33;
SURVEY.md App. D): Hamilton product, ``.angle`` wrapped to (-pi, pi], ``.axis`` = 0 for a null
rotation.  Quirks are reproduced on purpose:
  * rotational velocities use the time-reversed difference  q_k^-1 * q_{k-1};
  * 3-DoF joint velocities are rotation-vector rates written into Euler-rate slots;
  * frame 0 has zero velocity; ``dura`` is the *previous* frame's duration.
Output layout = MuJoCo qpos/qvel order of dp_env_v3.xml.
"""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Dict, List, Sequence

import numpy as np

BODY_JOINTS = ["chest", "neck", "right_shoulder", "right_elbow", "left_shoulder", "left_elbow", "right_hip",
               "right_knee", "right_ankle", "left_hip", "left_knee", "left_ankle"]          # MuJoCo order
BODY_JOINTS_IN_DP_ORDER = ["chest", "neck", "right_hip", "right_knee", "right_ankle", "right_shoulder", "right_elbow",
                           "left_hip", "left_knee", "left_ankle", "left_shoulder", "left_elbow"]  # file order
DOF_DEF = {"root": 3, "chest": 3, "neck": 3, "right_shoulder": 3, "right_elbow": 1, "right_wrist": 0,
           "left_shoulder": 3, "left_elbow": 1, "left_wrist": 0, "right_hip": 3, "right_knee": 1, "right_ankle": 3,
           "left_hip": 3, "left_knee": 1, "left_ankle": 3}

_S = np.sqrt(0.5)
_Q_ALIGN_LEFT = np.array([_S, _S, 0.0, 0.0])    # Quaternion(matrix=Rx(+90 deg))   (mocap_util.py:36-38)
_Q_ALIGN_RIGHT = np.array([_S, -_S, 0.0, 0.0])  # Quaternion(matrix=Rx(-90 deg))   (mocap_util.py:33-35)


# ---------------------------------------------------------------------------------------------
# quaternion helpers (w, x, y, z), batched over leading axes
# ---------------------------------------------------------------------------------------------
def qmul(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], axis=-1)


def qconj(q):
    q = np.asarray(q, dtype=np.float64)
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def qnormalise(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q, axis=-1, keepdims=True)
    return np.where(n > 0, q / np.where(n > 0, n, 1.0), q)


def quat_angle(q):
    """pyquaternion ``Quaternion.angle``: 2*atan2(|v|, w) wrapped to (-pi, pi]."""
    q = qnormalise(q)
    th = 2.0 * np.arctan2(np.linalg.norm(q[..., 1:], axis=-1), q[..., 0])
    res = np.mod(th + np.pi, 2 * np.pi) - np.pi
    return np.where(res == -np.pi, np.pi, res)


def quat_axis(q):
    """pyquaternion ``Quaternion.axis``: unit vector part, or zeros when |v| < 1e-17."""
    q = qnormalise(q)
    v = q[..., 1:]
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return np.where(n < 1e-17, 0.0, v / np.where(n < 1e-17, 1.0, n))


def align_rotation(q):
    """y-up -> z-up change of basis  q' = qL * q * qR  (mocap_util.py:31-40)."""
    return qmul(qmul(_Q_ALIGN_LEFT, q), _Q_ALIGN_RIGHT)


def align_position(p):
    """(x, y, z) -> (x, -z, y)  (mocap_util.py:42-48)."""
    p = np.asarray(p, dtype=np.float64)
    return np.stack([p[..., 0], -p[..., 2], p[..., 1]], axis=-1)


def quat_to_matrix_xyzw_style(q_wxyz):
    """Rotation matrix of a (w,x,y,z) quaternion, normalised the way transformations.quaternion_matrix
    does (scale by sqrt(2/|q|^2); identity for a null quaternion)."""
    q = np.asarray(q_wxyz, dtype=np.float64)
    n = np.sum(q * q, axis=-1, keepdims=True)
    ok = n[..., 0] >= np.finfo(float).eps * 4.0
    s = np.sqrt(2.0 / np.where(n > 0, n, 1.0))
    w, x, y, z = np.moveaxis(q * s, -1, 0)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1.0 - y * y - z * z; R[..., 0, 1] = x * y - z * w; R[..., 0, 2] = x * z + y * w
    R[..., 1, 0] = x * y + z * w; R[..., 1, 1] = 1.0 - x * x - z * z; R[..., 1, 2] = y * z - x * w
    R[..., 2, 0] = x * z - y * w; R[..., 2, 1] = y * z + x * w; R[..., 2, 2] = 1.0 - x * x - y * y
    R[~ok] = np.eye(3)
    return R


def euler_rxyz_from_quat(q_wxyz):
    """Intrinsic x-y-z Euler angles (R = Rx(a) Ry(b) Rz(c)); equals
    transformations.euler_from_quaternion([x,y,z,w], axes='rxyz') as called at mocap_v2.py:136-139."""
    R = quat_to_matrix_xyzw_style(q_wxyz)
    cy = np.sqrt(R[..., 2, 2] ** 2 + R[..., 1, 2] ** 2)
    eps = np.finfo(float).eps * 4.0
    a = np.where(cy > eps, np.arctan2(-R[..., 1, 2], R[..., 2, 2]), 0.0)
    b = np.arctan2(R[..., 0, 2], cy)
    c = np.where(cy > eps, np.arctan2(-R[..., 0, 1], R[..., 0, 0]), np.arctan2(R[..., 1, 0], R[..., 1, 1]))
    return np.stack([a, b, c], axis=-1)


def calc_rot_vel(seg0, seg1, dura):
    """MocapDM.calc_rot_vel (mocap_v2.py:64-76): angle(q0^-1 q1)/dura * axis(q0^-1 q1)."""
    qd = qmul(qconj(seg0), seg1)
    return (quat_angle(qd) / dura)[..., None] * quat_axis(qd)


def quat_from_euler_rxyz(e):
    """Quaternion (w,x,y,z) of the hinge triple R = Rx(a) Ry(b) Rz(c); inverse of euler_rxyz_from_quat and
    equal to transformations.quaternion_from_euler(a, b, c, 'rxyz') (transformations.py:1100-1154)."""
    e = np.asarray(e, dtype=np.float64)
    h = 0.5 * e
    z, o = np.zeros_like(h[..., 0]), np.ones_like(h[..., 0])
    qx = np.stack([np.cos(h[..., 0]), np.sin(h[..., 0]), z, z], axis=-1)
    qy = np.stack([np.cos(h[..., 1]), z, np.sin(h[..., 1]), z], axis=-1)
    qz = np.stack([np.cos(h[..., 2]), z, z, np.sin(h[..., 2])], axis=-1)
    del o
    return qmul(qmul(qx, qy), qz)


_EPS = np.finfo(float).eps * 4.0


def quat_slerp(q0, q1, fraction):
    """transformations.quaternion_slerp(q0, q1, fraction, spin=0, shortestpath=True)
    (transformations.py:1270-1308), restated for one pair of quaternions (any component order)."""
    q0 = np.array(q0, dtype=np.float64); q1 = np.array(q1, dtype=np.float64)
    q0 /= np.linalg.norm(q0); q1 /= np.linalg.norm(q1)
    if fraction == 0.0:
        return q0
    if fraction == 1.0:
        return q1
    d = float(np.dot(q0, q1))
    if abs(abs(d) - 1.0) < _EPS:
        return q0
    if d < 0.0:
        d, q1 = -d, -q1
    angle = np.arccos(d)
    if abs(angle) < _EPS:
        return q0
    isin = 1.0 / np.sin(angle)
    return q0 * (np.sin((1.0 - fraction) * angle) * isin) + q1 * (np.sin(fraction * angle) * isin)


def sample_tables(data_config, data_vel, u, body_dofs=None):
    """Time-based mocap lookup with interpolation (BASELINE north_star
    "mocap-frame interpolation"): reference pose at the fractional frame coordinate ``u = t / clip_dt``.

    One cycle of an F-frame clip spans F-1 frame intervals (the last frame closes the loop); past the
    end the clip wraps and the root x/y are shifted by the last frame's root x/y once per completed
    cycle -- the root-offset accumulation of MocapDM.play (mocap_v2.py:168-182).  Inside an interval:
    root position, 1-DoF joints and all velocities are interpolated linearly; the root quaternion and
    the 3-DoF joints (hinge triples <-> quaternions) by quaternion_slerp, back to 'rxyz' Euler angles.
    Returns (qpos[35], qvel[34], cycle, k, alpha)."""
    cfg, vel = np.asarray(data_config, dtype=np.float64), np.asarray(data_vel, dtype=np.float64)
    F = cfg.shape[0]
    if F < 2:
        return cfg[0].copy(), vel[0].copy(), 0, 0, 0.0
    cycle = int(np.floor(u / (F - 1)))
    uu = u - cycle * (F - 1)
    k = min(int(uu), F - 2)
    a = uu - k
    c0, c1 = cfg[k], cfg[k + 1]
    q = c0 + a * (c1 - c0)
    q[0:2] += cycle * cfg[F - 1, 0:2]
    q[3:7] = quat_slerp(c0[3:7], c1[3:7], a)
    off = 7
    for jn in BODY_JOINTS:
        if DOF_DEF[jn] == 3:
            qs = quat_slerp(quat_from_euler_rxyz(c0[off:off + 3]), quat_from_euler_rxyz(c1[off:off + 3]), a)
            q[off:off + 3] = euler_rxyz_from_quat(qs)
        off += DOF_DEF[jn]
    v = vel[k] + a * (vel[k + 1] - vel[k])
    return q, v, cycle, k, a


# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Clip:
    name: str
    dt: float                 # MocapDM.dt = first frame's duration (mocap_v2.py:37)
    loop: str
    durations: np.ndarray     # [F]
    data: np.ndarray          # [F,44]  MocapDM.data (dura, root pos, root quat, joints in MuJoCo order)
    data_config: np.ndarray   # [F,35]  qpos per frame
    data_vel: np.ndarray      # [F,34]  qvel per frame

    def __len__(self):
        return self.data_config.shape[0]

    def sample(self, t: float):
        """Interpolated reference (qpos, qvel) at mocap time ``t`` seconds (see sample_tables)."""
        q, v, _, _, _ = sample_tables(self.data_config, np.nan_to_num(self.data_vel), t / self.dt)
        return q, v


def compile_frames(frames: np.ndarray, name: str = "clip", loop: str = "wrap") -> Clip:
    """frames: [F,44] raw DeepMimic frames (file order, y-up, quats wxyz)."""
    fr = np.asarray(frames, dtype=np.float64)
    if fr.ndim != 2 or fr.shape[1] != 44 or fr.shape[0] < 1:
        raise ValueError("expected Frames of shape [F,44], got %r" % (fr.shape,))
    F = fr.shape[0]
    durations = fr[:, 0].copy()
    root_pos = align_position(fr[:, 1:4])
    root_rot = align_rotation(fr[:, 4:8])
    state: Dict[str, np.ndarray] = {}
    off = 8
    for jn in BODY_JOINTS_IN_DP_ORDER:
        if DOF_DEF[jn] == 1:
            state[jn] = fr[:, off:off + 1].copy()
            off += 1
        else:
            state[jn] = align_rotation(fr[:, off:off + 4])
            off += 4
    assert off == 44
    dura = np.concatenate([durations[:1], durations[:-1]])  # frame k uses durations[k-1]
    data = np.empty((F, 44))
    data[:, 0] = dura
    data[:, 1:4] = root_pos
    data[:, 4:8] = root_rot
    cfg = [root_pos, root_rot]
    vel = [np.zeros((F, 3)), np.zeros((F, 3))]
    with np.errstate(divide="ignore", invalid="ignore"):
        if F > 1:
            vel[0][1:] = (root_pos[1:] - root_pos[:-1]) / dura[1:, None]
            vel[1][1:] = calc_rot_vel(root_rot[1:], root_rot[:-1], dura[1:])
        off = 8
        for jn in BODY_JOINTS:
            v = state[jn]
            if DOF_DEF[jn] == 1:
                data[:, off:off + 1] = v
                off += 1
                jv = np.zeros((F, 1))
                if F > 1:
                    jv[1:] = (v[1:] - v[:-1]) / dura[1:, None]
                cfg.append(v)
            else:
                data[:, off:off + 4] = v
                off += 4
                jv = np.zeros((F, 3))
                if F > 1:
                    jv[1:] = calc_rot_vel(v[1:], v[:-1], dura[1:])
                cfg.append(euler_rxyz_from_quat(v))
            vel.append(jv)
    return Clip(name=name, dt=float(durations[0]), loop=loop, durations=durations, data=data,
                data_config=np.concatenate(cfg, axis=1), data_vel=np.concatenate(vel, axis=1))


def load_clip(path: str, name: str | None = None) -> Clip:
    """Load a DeepMimic motion file: ``*.txt`` JSON ({"Loop", "Frames"}) or a ``*.npz`` with the same
    raw frames (the format the repo ships, see assets/README)."""
    if name is None:
        name = os.path.splitext(os.path.basename(path))[0].replace("humanoid3d_", "")
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        return compile_frames(z["frames"], name=name, loop=str(z["loop"]))
    with open(path, "r") as f:
        d = json.load(f)
    return compile_frames(np.array(d["Frames"], dtype=np.float64), name=name, loop=str(d.get("Loop", "wrap")))


@dataclasses.dataclass
class MocapTables:
    """Concatenated clips for mixed batches (per-clip start/len)."""
    names: List[str]
    clip_start: np.ndarray
    clip_len: np.ndarray
    clip_dt: np.ndarray
    data_config: np.ndarray  # [Ftot,35]
    data_vel: np.ndarray     # [Ftot,34]


def concat_clips(clips: Sequence[Clip]) -> MocapTables:
    starts, n = [], 0
    for c in clips:
        starts.append(n)
        n += len(c)
    return MocapTables(names=[c.name for c in clips], clip_start=np.array(starts, dtype=np.int32),
                       clip_len=np.array([len(c) for c in clips], dtype=np.int32),
                       clip_dt=np.array([c.dt for c in clips]),
                       data_config=np.ascontiguousarray(np.concatenate([c.data_config for c in clips], axis=0)),
                       data_vel=np.ascontiguousarray(np.concatenate([c.data_vel for c in clips], axis=0)))
