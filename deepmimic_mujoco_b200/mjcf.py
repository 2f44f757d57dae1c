"""MJCF-subset model compiler: dp_env_v3.xml -> flat constant tables (float64).

Host-side, load-time only.  Replaces what ``mujoco_py.load_model_from_path`` +
MuJoCo's model compiler do for the reference env
(/root/reference/src/dp_env_v3.py:59 via gym ``MujocoEnv.__init__``) for exactly the
MJCF subset the humanoid uses
(/root/reference/src/mujoco/humanoid_deepmimic/envs/asset/dp_env_v3.xml:1-156):
``compiler angle=radian inertiafromgeom=true``, ``<default>`` for joint/geom/motor,
``<option>``, nested bodies with sphere / capsule(fromto) / box / plane geoms, free and
hinge joints, ``<contact><exclude>``, ``<actuator><motor>``.

The output :class:`ModelTables` is what both the float64 CPU oracle
(``oracle/dm_oracle.c``) and the CUDA library (``csrc/dmb.cu``) consume through the
``dmb_model_t`` struct of ``include/dmb_model.h``.

Derived quantities restated from MuJoCo's documented compiler behaviour (SURVEY.md
App. A / B.11):
  * geom inertias: sphere 2/5 m r^2; capsule = cylinder + two hemispheres at uniform
    density scaled to the given mass; box m/3 (b^2 + c^2) with half-sizes;
  * body inertial frame from its geoms (mass-weighted COM + parallel axis).  We keep the
    full symmetric inertia tensor in the body frame instead of (iquat, diag inertia) --
    physically identical, no eigen-decomposition needed;
  * ``body_invweight0`` / ``dof_invweight0`` / ``stat.meaninertia`` from the dense joint
    space inertia at ``qpos0`` (computed here with an independent numpy CRB, which the
    test-suite also uses to cross-check the oracle's sparse CRB);
  * collision candidate pairs: all geom pairs with ``contype & conaffinity``, minus
    same-body, minus parent-child bodies unless the parent is the world, minus
    ``<exclude>``; ordered the way MuJoCo's body-pair broad phase emits them (body pair
    major, geom minor).
"""
from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET
from typing import List, Optional

import numpy as np

GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 0, 2, 3, 6  # MuJoCo mjtGeom ids
JNT_FREE, JNT_HINGE = 0, 3  # MuJoCo mjtJoint ids

_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "box": GEOM_BOX}


def _floats(s: Optional[str], n: Optional[int] = None, default=None) -> np.ndarray:
    if s is None:
        if default is None:
            raise ValueError("missing attribute")
        return np.asarray(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} floats, got {v.size}: {s!r}")
    return v


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def _z_to_quat(vec):
    """Quaternion rotating +z onto ``vec`` (MuJoCo's mjuu_z2quat)."""
    v = np.asarray(vec, dtype=np.float64)
    v = v / np.linalg.norm(v)
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(z, v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        # parallel (identity) or anti-parallel (half-turn about x)
        return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1.0, 0, 0])
    axis /= s
    ang = np.arctan2(s, v[2])
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])


@dataclasses.dataclass
class ModelTables:
    """Flat float64/int32 constant tables for one articulated model."""

    # sizes
    nq: int
    nv: int
    nu: int
    nbody: int
    njnt: int
    ngeom: int
    npair: int
    nM: int
    # options
    timestep: float
    gravity: np.ndarray
    iterations: int
    tolerance: float
    solref: np.ndarray
    solimp: np.ndarray
    margin: float
    meaninertia: float
    # names (host only)
    body_names: List[str]
    joint_names: List[str]
    geom_names: List[str]
    actuator_names: List[str]
    # bodies
    body_parent: np.ndarray
    body_depth: np.ndarray
    body_jntadr: np.ndarray
    body_jntnum: np.ndarray
    body_dofadr: np.ndarray
    body_dofnum: np.ndarray
    body_pos: np.ndarray      # [nbody,3] in parent frame
    body_quat: np.ndarray     # [nbody,4]
    body_ipos: np.ndarray     # [nbody,3] COM in body frame
    body_inertia: np.ndarray  # [nbody,6] xx,yy,zz,xy,xz,yz about COM, body frame
    body_mass: np.ndarray
    body_invweight0: np.ndarray  # [nbody,2]
    # joints
    jnt_type: np.ndarray
    jnt_bodyid: np.ndarray
    jnt_qposadr: np.ndarray
    jnt_dofadr: np.ndarray
    jnt_axis: np.ndarray
    jnt_limited: np.ndarray
    jnt_range: np.ndarray
    # dofs
    dof_bodyid: np.ndarray
    dof_jntid: np.ndarray
    dof_parentid: np.ndarray
    dof_Madr: np.ndarray
    dof_armature: np.ndarray
    dof_damping: np.ndarray
    dof_invweight0: np.ndarray
    # geoms
    geom_type: np.ndarray
    geom_bodyid: np.ndarray
    geom_condim: np.ndarray
    geom_size: np.ndarray
    geom_pos: np.ndarray
    geom_quat: np.ndarray
    geom_rbound: np.ndarray
    geom_friction: np.ndarray
    geom_mass: np.ndarray
    # candidate collision pairs (type-ordered: pair_geom1 has the smaller geom type)
    pair_geom1: np.ndarray
    pair_geom2: np.ndarray
    # actuators
    act_dofadr: np.ndarray
    act_gear: np.ndarray
    act_ctrlrange: np.ndarray
    # reference configuration
    qpos0: np.ndarray

    def total_mass(self) -> float:
        return float(self.body_mass.sum())


# --------------------------------------------------------------------------------------
# geom inertia helpers
# --------------------------------------------------------------------------------------

def _geom_inertia_diag(gtype: int, size: np.ndarray, mass: float) -> np.ndarray:
    if gtype == GEOM_SPHERE:
        r = size[0]
        return np.full(3, 0.4 * mass * r * r)
    if gtype == GEOM_CAPSULE:
        r, h = size[0], 2.0 * size[1]
        sphere_mass = mass * 4 * r / (4 * r + 3 * h)
        cyl_mass = mass - sphere_mass
        ixx = cyl_mass * (3 * r * r + h * h) / 12.0
        izz = cyl_mass * r * r / 2.0
        sph_i = 0.4 * sphere_mass * r * r
        ixx += sph_i + sphere_mass * h * (3 * r + 2 * h) / 8.0
        izz += sph_i
        return np.array([ixx, ixx, izz])
    if gtype == GEOM_BOX:
        a, b, c = size
        return mass / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    raise ValueError("inertia of geom type %d not supported" % gtype)


def _rbound(gtype: int, size: np.ndarray) -> float:
    if gtype == GEOM_SPHERE:
        return float(size[0])
    if gtype == GEOM_CAPSULE:
        return float(size[0] + size[1])
    if gtype == GEOM_BOX:
        return float(np.linalg.norm(size))
    return 0.0  # plane


# --------------------------------------------------------------------------------------
# parser
# --------------------------------------------------------------------------------------

def compile_mjcf(path: str) -> ModelTables:
    root = ET.parse(path).getroot()
    comp = root.find("compiler")
    if comp is None or comp.get("angle") != "radian" or comp.get("inertiafromgeom") != "true":
        raise ValueError("only <compiler angle=radian inertiafromgeom=true> models are supported")

    dflt = root.find("default")
    dj = dict(dflt.find("joint").attrib) if dflt is not None and dflt.find("joint") is not None else {}
    dg = dict(dflt.find("geom").attrib) if dflt is not None and dflt.find("geom") is not None else {}
    dm = dict(dflt.find("motor").attrib) if dflt is not None and dflt.find("motor") is not None else {}

    opt = root.find("option")
    if opt.get("integrator") != "RK4" or opt.get("solver") != "PGS":
        raise ValueError("only integrator=RK4 solver=PGS is implemented (dp_env_v3.xml:9)")
    timestep = float(opt.get("timestep"))
    iterations = int(opt.get("iterations", 100))

    bodies: List[dict] = [dict(name="world", parent=-1, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                               joints=[], geoms=[])]
    joints: List[dict] = []
    geoms: List[dict] = []

    def gattr(el, key, default=None):
        if key in el.attrib:
            return el.attrib[key]
        return dg.get(key, default)

    def add_geom(el, bid):
        gtype = _GEOM_TYPES[el.get("type", "sphere")]
        size = np.zeros(3)
        pos = _floats(el.get("pos"), 3, default=[0, 0, 0])
        quat = np.array([1.0, 0, 0, 0])
        sz = _floats(el.get("size"))
        if gtype == GEOM_CAPSULE:
            ft = _floats(el.get("fromto"), 6)
            vec = ft[0:3] - ft[3:6]  # MuJoCo: from - to
            size[0] = sz[0]
            size[1] = np.linalg.norm(vec) / 2.0
            pos = 0.5 * (ft[0:3] + ft[3:6])
            quat = _z_to_quat(vec)
        elif gtype == GEOM_PLANE:
            size[:] = sz[:3]
        else:
            size[: sz.size] = sz
        fr = _floats(gattr(el, "friction"), None, default=[1.0, 0.005, 0.0001])
        friction = np.array([1.0, 0.005, 0.0001])
        friction[: fr.size] = fr
        g = dict(name=el.get("name", "geom%d" % len(geoms)), type=gtype, body=bid, size=size, pos=pos, quat=quat,
                 contype=int(gattr(el, "contype", 1)), conaffinity=int(gattr(el, "conaffinity", 1)),
                 condim=int(gattr(el, "condim", 3)), margin=float(gattr(el, "margin", 0.0)),
                 friction=friction, mass=float(el.get("mass", 0.0)) if gtype != GEOM_PLANE else 0.0)
        bodies[bid]["geoms"].append(len(geoms))
        geoms.append(g)

    def add_joint(el, bid):
        jtype = {"free": JNT_FREE, "hinge": JNT_HINGE}[el.get("type", dj.get("type", "hinge"))]
        def ja(key, default):
            return el.attrib.get(key, dj.get(key, default))
        if np.abs(_floats(el.get("pos"), 3, default=[0, 0, 0])).max() != 0.0:
            raise ValueError("joint pos != 0 is not supported by the hot path (all dp_env_v3 joints are at the body origin)")
        j = dict(name=el.get("name"), type=jtype, body=bid,
                 axis=_floats(el.get("axis"), 3, default=[0, 0, 1]),
                 limited=(ja("limited", "false") == "true") and jtype == JNT_HINGE,
                 range=_floats(el.get("range"), 2, default=[0, 0]),
                 armature=float(ja("armature", 0.0)), damping=float(ja("damping", 0.0)))
        if float(ja("stiffness", 0.0)) != 0.0:
            raise ValueError("joint stiffness not supported")
        if jtype == JNT_HINGE:
            j["axis"] = j["axis"] / np.linalg.norm(j["axis"])
        bodies[bid]["joints"].append(len(joints))
        joints.append(j)

    def walk(el, parent):
        bid = len(bodies)
        bodies.append(dict(name=el.get("name"), parent=parent, pos=_floats(el.get("pos"), 3, default=[0, 0, 0]),
                           quat=_floats(el.get("quat"), 4, default=[1, 0, 0, 0]), joints=[], geoms=[]))
        # MuJoCo numbers joints/geoms in document order of the depth-first traversal
        for ch in el:
            if ch.tag == "joint":
                add_joint(ch, bid)
            elif ch.tag == "geom":
                add_geom(ch, bid)
        for ch in el:
            if ch.tag == "body":
                walk(ch, bid)

    wb = root.find("worldbody")
    for ch in wb:
        if ch.tag == "geom":
            add_geom(ch, 0)
    for ch in wb:
        if ch.tag == "body":
            walk(ch, 0)

    nbody, njnt, ngeom = len(bodies), len(joints), len(geoms)

    # ---- joints / dofs ----------------------------------------------------------------
    jnt_qposadr, jnt_dofadr = [], []
    dof_bodyid, dof_jntid, dof_arm, dof_damp = [], [], [], []
    nq = nv = 0
    for ji, j in enumerate(joints):
        jnt_qposadr.append(nq)
        jnt_dofadr.append(nv)
        nd = 6 if j["type"] == JNT_FREE else 1
        nq += 7 if j["type"] == JNT_FREE else 1
        nv += nd
        dof_bodyid += [j["body"]] * nd
        dof_jntid += [ji] * nd
        dof_arm += [j["armature"]] * nd
        dof_damp += [j["damping"]] * nd
    body_jntadr = np.array([b["joints"][0] if b["joints"] else -1 for b in bodies], dtype=np.int32)
    body_jntnum = np.array([len(b["joints"]) for b in bodies], dtype=np.int32)
    body_dofadr = np.full(nbody, -1, dtype=np.int32)
    body_dofnum = np.zeros(nbody, dtype=np.int32)
    for bi, b in enumerate(bodies):
        if b["joints"]:
            body_dofadr[bi] = jnt_dofadr[b["joints"][0]]
            body_dofnum[bi] = sum(6 if joints[j]["type"] == JNT_FREE else 1 for j in b["joints"])
    # dof parent chain: previous dof in same body, else last dof of nearest ancestor with dofs
    dof_parentid = np.full(nv, -1, dtype=np.int32)
    for d in range(nv):
        b = dof_bodyid[d]
        if d > body_dofadr[b]:
            dof_parentid[d] = d - 1
        else:
            p = bodies[b]["parent"]
            while p > 0 and body_dofnum[p] == 0:
                p = bodies[p]["parent"]
            if p > 0:
                dof_parentid[d] = body_dofadr[p] + body_dofnum[p] - 1
    dof_Madr = np.zeros(nv, dtype=np.int32)
    nM = 0
    for d in range(nv):
        dof_Madr[d] = nM
        k = d
        while k >= 0:
            nM += 1
            k = dof_parentid[k]
    body_depth = np.zeros(nbody, dtype=np.int32)
    for bi in range(1, nbody):
        body_depth[bi] = body_depth[bodies[bi]["parent"]] + 1

    # ---- body inertials from geoms ---------------------------------------------------
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_I = np.zeros((nbody, 3, 3))
    for bi, b in enumerate(bodies):
        gs = [geoms[g] for g in b["geoms"] if geoms[g]["type"] != GEOM_PLANE]
        m = sum(g["mass"] for g in gs)
        if m <= 0:
            continue
        com = sum(g["mass"] * g["pos"] for g in gs) / m
        I = np.zeros((3, 3))
        for g in gs:
            R = quat_to_mat(g["quat"])
            Ig = R @ np.diag(_geom_inertia_diag(g["type"], g["size"], g["mass"])) @ R.T
            d = g["pos"] - com
            I += Ig + g["mass"] * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        body_mass[bi], body_ipos[bi], body_I[bi] = m, com, I
    body_inertia = np.stack([body_I[:, 0, 0], body_I[:, 1, 1], body_I[:, 2, 2],
                             body_I[:, 0, 1], body_I[:, 0, 2], body_I[:, 1, 2]], axis=1)

    # ---- collision pairs -------------------------------------------------------------
    excl = set()
    con = root.find("contact")
    names = [b["name"] for b in bodies]
    if con is not None:
        for e in con.findall("exclude"):
            a, b = names.index(e.get("body1")), names.index(e.get("body2"))
            excl.add((min(a, b), max(a, b)))
    pair1, pair2 = [], []
    for b1 in range(nbody):
        for b2 in range(b1 + 1, nbody):
            if (b1, b2) in excl:
                continue
            if b1 != 0 and (bodies[b2]["parent"] == b1 or bodies[b1]["parent"] == b2):
                continue  # parent-child filter (kept when the parent is the world)
            for g1 in bodies[b1]["geoms"]:
                for g2 in bodies[b2]["geoms"]:
                    ga, gb = geoms[g1], geoms[g2]
                    if not ((ga["contype"] & gb["conaffinity"]) or (gb["contype"] & ga["conaffinity"])):
                        continue
                    if ga["type"] > gb["type"]:  # MuJoCo orders a pair by geom type
                        g1o, g2o = g2, g1
                    else:
                        g1o, g2o = g1, g2
                    pair1.append(g1o)
                    pair2.append(g2o)

    # ---- actuators -------------------------------------------------------------------
    act = root.find("actuator")
    jnames = [j["name"] for j in joints]
    act_dofadr, act_gear, act_range, act_names = [], [], [], []
    for mtr in act.findall("motor"):
        ji = jnames.index(mtr.get("joint"))
        act_dofadr.append(jnt_dofadr[ji])
        act_gear.append(float(mtr.get("gear", dm.get("gear", "1")).split()[0]))
        limited = mtr.get("ctrllimited", dm.get("ctrllimited", "false")) == "true"
        rng = _floats(mtr.get("ctrlrange", dm.get("ctrlrange")), 2, default=[0, 0])
        act_range.append(rng if limited else np.array([-np.inf, np.inf]))
        act_names.append(mtr.get("name"))

    qpos0 = np.zeros(nq)
    for ji, j in enumerate(joints):
        if j["type"] == JNT_FREE:
            b = bodies[j["body"]]
            qpos0[jnt_qposadr[ji]: jnt_qposadr[ji] + 3] = b["pos"]
            qpos0[jnt_qposadr[ji] + 3: jnt_qposadr[ji] + 7] = b["quat"]

    margins = {g["margin"] for g in geoms}
    if len(margins) != 1:
        raise ValueError("per-geom margins differ; the hot path assumes one global margin")

    mt = ModelTables(
        nq=nq, nv=nv, nu=len(act_dofadr), nbody=nbody, njnt=njnt, ngeom=ngeom, npair=len(pair1), nM=nM,
        timestep=timestep, gravity=np.array([0.0, 0.0, -9.81]), iterations=iterations, tolerance=1e-8,
        solref=np.array([0.02, 1.0]), solimp=np.array([0.9, 0.95, 0.001, 0.5, 2.0]), margin=margins.pop(),
        meaninertia=1.0,
        body_names=names, joint_names=jnames, geom_names=[g["name"] for g in geoms], actuator_names=act_names,
        body_parent=np.array([b["parent"] for b in bodies], dtype=np.int32), body_depth=body_depth,
        body_jntadr=body_jntadr, body_jntnum=body_jntnum, body_dofadr=body_dofadr, body_dofnum=body_dofnum,
        body_pos=np.array([b["pos"] for b in bodies]), body_quat=np.array([b["quat"] for b in bodies]),
        body_ipos=body_ipos, body_inertia=body_inertia, body_mass=body_mass,
        body_invweight0=np.zeros((nbody, 2)),
        jnt_type=np.array([j["type"] for j in joints], dtype=np.int32),
        jnt_bodyid=np.array([j["body"] for j in joints], dtype=np.int32),
        jnt_qposadr=np.array(jnt_qposadr, dtype=np.int32), jnt_dofadr=np.array(jnt_dofadr, dtype=np.int32),
        jnt_axis=np.array([j["axis"] for j in joints]),
        jnt_limited=np.array([int(j["limited"]) for j in joints], dtype=np.int32),
        jnt_range=np.array([j["range"] for j in joints]),
        dof_bodyid=np.array(dof_bodyid, dtype=np.int32), dof_jntid=np.array(dof_jntid, dtype=np.int32),
        dof_parentid=dof_parentid, dof_Madr=dof_Madr,
        dof_armature=np.array(dof_arm), dof_damping=np.array(dof_damp), dof_invweight0=np.zeros(nv),
        geom_type=np.array([g["type"] for g in geoms], dtype=np.int32),
        geom_bodyid=np.array([g["body"] for g in geoms], dtype=np.int32),
        geom_condim=np.array([g["condim"] for g in geoms], dtype=np.int32),
        geom_size=np.array([g["size"] for g in geoms]), geom_pos=np.array([g["pos"] for g in geoms]),
        geom_quat=np.array([g["quat"] for g in geoms]),
        geom_rbound=np.array([_rbound(g["type"], g["size"]) for g in geoms]),
        geom_friction=np.array([g["friction"] for g in geoms]), geom_mass=np.array([g["mass"] for g in geoms]),
        pair_geom1=np.array(pair1, dtype=np.int32), pair_geom2=np.array(pair2, dtype=np.int32),
        act_dofadr=np.array(act_dofadr, dtype=np.int32), act_gear=np.array(act_gear),
        act_ctrlrange=np.array(act_range), qpos0=qpos0,
    )
    _set_const(mt)
    return mt


# --------------------------------------------------------------------------------------
# independent numpy kinematics / CRB (used for invweight0 and as a test cross-check)
# --------------------------------------------------------------------------------------

def np_kinematics(mt: ModelTables, qpos: np.ndarray):
    """World poses of every body for configuration ``qpos`` (numpy, float64)."""
    xpos = np.zeros((mt.nbody, 3))
    xquat = np.zeros((mt.nbody, 4))
    xquat[0] = [1, 0, 0, 0]
    xaxis = np.zeros((mt.njnt, 3))
    for b in range(1, mt.nbody):
        p = mt.body_parent[b]
        Rp = quat_to_mat(xquat[p])
        pos = xpos[p] + Rp @ mt.body_pos[b]
        quat = quat_mul(xquat[p], mt.body_quat[b])
        for j in range(mt.body_jntadr[b], mt.body_jntadr[b] + mt.body_jntnum[b]):
            qa = mt.jnt_qposadr[j]
            if mt.jnt_type[j] == JNT_FREE:
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
            else:
                xaxis[j] = quat_to_mat(quat) @ mt.jnt_axis[j]
                ang = qpos[qa] - mt.qpos0[qa]
                ql = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * mt.jnt_axis[j]])
                quat = quat_mul(quat, ql)
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    xmat = np.stack([quat_to_mat(q) for q in xquat])
    xipos = xpos + np.einsum("bij,bj->bi", xmat, mt.body_ipos)
    return xpos, xquat, xmat, xipos, xaxis


def np_body_jacobian(mt: ModelTables, xpos, xmat, xaxis, body: int, point: np.ndarray):
    """6 x nv world-frame Jacobian (linear; angular) of ``point`` rigidly attached to ``body``."""
    J = np.zeros((6, mt.nv))
    b = body
    while b > 0:
        for j in range(mt.body_jntadr[b], mt.body_jntadr[b] + mt.body_jntnum[b]):
            da = mt.jnt_dofadr[j]
            if mt.jnt_type[j] == JNT_FREE:
                J[0:3, da:da + 3] = np.eye(3)
                for k in range(3):  # rotational dofs are about the body-frame axes
                    ax = xmat[b][:, k]
                    J[3:6, da + 3 + k] = ax
                    J[0:3, da + 3 + k] = np.cross(ax, point - xpos[b])
            else:
                J[3:6, da] = xaxis[j]
                J[0:3, da] = np.cross(xaxis[j], point - xpos[b])  # joint anchor == body origin
        b = mt.body_parent[b]
    return J


def np_mass_matrix(mt: ModelTables, qpos: np.ndarray) -> np.ndarray:
    """Dense joint-space inertia M(q) (incl. armature) from body Jacobians: sum_b J^T I_b J."""
    xpos, xquat, xmat, xipos, xaxis = np_kinematics(mt, qpos)
    M = np.zeros((mt.nv, mt.nv))
    for b in range(1, mt.nbody):
        if mt.body_mass[b] == 0:
            continue
        J = np_body_jacobian(mt, xpos, xmat, xaxis, b, xipos[b])
        ii = mt.body_inertia[b]
        Ib = np.array([[ii[0], ii[3], ii[4]], [ii[3], ii[1], ii[5]], [ii[4], ii[5], ii[2]]])
        Iw = xmat[b] @ Ib @ xmat[b].T
        M += mt.body_mass[b] * J[0:3].T @ J[0:3] + J[3:6].T @ Iw @ J[3:6]
    M += np.diag(mt.dof_armature)
    return M


def _set_const(mt: ModelTables) -> None:
    """body_invweight0 / dof_invweight0 / meaninertia at qpos0 (MuJoCo's mj_setConst)."""
    M = np_mass_matrix(mt, mt.qpos0)
    Minv = np.linalg.inv(M)
    xpos, xquat, xmat, xipos, xaxis = np_kinematics(mt, mt.qpos0)
    mt.meaninertia = float(np.trace(M) / mt.nv)
    for b in range(1, mt.nbody):
        J = np_body_jacobian(mt, xpos, xmat, xaxis, b, xipos[b])
        A = J @ Minv @ J.T
        mt.body_invweight0[b, 0] = (A[0, 0] + A[1, 1] + A[2, 2]) / 3.0
        mt.body_invweight0[b, 1] = (A[3, 3] + A[4, 4] + A[5, 5]) / 3.0
    for j in range(mt.njnt):
        da = mt.jnt_dofadr[j]
        if mt.jnt_type[j] == JNT_FREE:
            d = np.diag(Minv)[da:da + 6]
            mt.dof_invweight0[da:da + 3] = d[0:3].mean()
            mt.dof_invweight0[da + 3:da + 6] = d[3:6].mean()
        else:
            mt.dof_invweight0[da] = Minv[da, da]


# --------------------------------------------------------------------------------------
# (de)serialisation of compiled tables (shipped asset; /root/reference is absent on the GPU box)
# --------------------------------------------------------------------------------------

_LIST_FIELDS = ("body_names", "joint_names", "geom_names", "actuator_names")


def save_tables(mt: ModelTables, path: str) -> None:
    d = {}
    for f in dataclasses.fields(mt):
        v = getattr(mt, f.name)
        d[f.name] = np.array(v) if f.name in _LIST_FIELDS else np.asarray(v)
    np.savez(path, **d)


def load_tables(path: str) -> ModelTables:
    z = np.load(path, allow_pickle=False)
    kw = {}
    for f in dataclasses.fields(ModelTables):
        v = z[f.name]
        if f.name in _LIST_FIELDS:
            kw[f.name] = [str(s) for s in v]
        elif v.ndim == 0:
            kw[f.name] = v.item()
        else:
            kw[f.name] = v
    return ModelTables(**kw)
