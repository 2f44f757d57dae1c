"""The algebraic identities the CUDA kernels rely on, checked in float64 numpy on the real
humanoid tables -- CPU only.  Each test restates the kernel's formulation (csrc/dmb_device.cuh, function named in
the docstring) next to the direct computation it replaces."""
import numpy as np

import common
import oracle.pyoracle as po
from deepmimic_mujoco_b200 import mjcf


def _chains(mt):
    """anc[d] = ancestors of dof d, nearest first; depth[d] = len(anc[d]); ndesc[d] = number of descendants."""
    anc = []
    for d in range(mt.nv):
        a, c = [], mt.dof_parentid[d]
        while c >= 0:
            a.append(int(c)); c = mt.dof_parentid[c]
        anc.append(a)
    ndesc = [sum(1 for e in range(mt.nv) if d in anc[e]) for d in range(mt.nv)]
    return anc, ndesc


def _factor(mt, q, v):
    """qLD (MuJoCo sparse layout, L entries scaled, D on the diagonal slot) and dense M from the oracle."""
    o = po.Oracle(common.model())
    o.set_state(q, v); o.forward()
    return o, o.d.arr("qLD")[: mt.nM].copy(), o.full_M()


def _dense_L_D(mt, qLD, anc):
    """M = L' D L with L unit lower triangular: L[i][a] stored at qLD[Madr[i] + 1 + rank] (rank: nearest first)."""
    L, D = np.eye(mt.nv), np.zeros(mt.nv)
    for i in range(mt.nv):
        D[i] = qLD[mt.dof_Madr[i]]
        for k, a in enumerate(anc[i]):
            L[i, a] = qLD[mt.dof_Madr[i] + 1 + k]
    return L, D


def test_depth_first_numbering_range_test_and_table_free_addressing():
    """reg_solve_LT / half_solve_rows: 'i is a descendant of a' <=> a < i <= a + ndesc[a], and L[i][a] sits at
    qLD[Lend[i] - depth(a)] with Lend[i] = Madr[i] + depth(i)."""
    mt = common.tables()
    anc, ndesc = _chains(mt)
    rng = np.random.default_rng(0)
    q, v = common.airborne_states(rng, 1)
    o, qLD, M = _factor(mt, q[0], v[0])
    L, D = _dense_L_D(mt, qLD, anc)
    assert np.abs(L.T @ np.diag(D) @ L - M).max() < 1e-10 * np.abs(M).max()
    depth = [len(a) for a in anc]
    for a in range(mt.nv):
        for i in range(mt.nv):
            assert (a in anc[i]) == (a < i <= a + ndesc[a])
            if a in anc[i]:
                assert qLD[mt.dof_Madr[i] + depth[i] - depth[a]] == L[i, a]
    # x <- L^-T x in the kernel's push form (largest id first), against the dense solve
    x = rng.normal(size=mt.nv); y = x.copy()
    for i in range(mt.nv - 1, 0, -1):
        for a in range(mt.nv):
            if a < i <= a + ndesc[a]:
                y[a] -= qLD[mt.dof_Madr[i] + depth[i] - depth[a]] * y[i]
    assert np.abs(y - np.linalg.solve(L.T, x)).max() < 1e-12 * max(1.0, np.abs(x).max())


def test_level_parallel_L_solve():
    """reg_solve_L: in round r every dof with more than r ancestors pulls from its ancestor at depth r (final since
    round r-1); 12 rounds instead of nv-1 steps, same result as the serial forward substitution."""
    mt = common.tables()
    anc, _ = _chains(mt)
    rng = np.random.default_rng(1)
    q, v = common.airborne_states(rng, 1)
    _, qLD, _ = _factor(mt, q[0], v[0])
    L, _ = _dense_L_D(mt, qLD, anc)
    x = rng.normal(size=mt.nv); y = x.copy()
    maxdepth = max(len(a) for a in anc)
    assert maxdepth == 12
    for r in range(maxdepth):
        snap = y.copy()                                  # all lanes shuffle before anyone writes
        for d in range(mt.nv):
            if r < len(anc[d]):
                a = anc[d][len(anc[d]) - 1 - r]          # ancestor at depth r (root side first)
                y[d] -= L[d, a] * snap[a]
    assert np.abs(y - np.linalg.solve(L, x)).max() < 1e-12 * max(1.0, np.abs(x).max())


def test_half_solved_jacobian_identities():
    """forward_eval: with Y_r = D^-1/2 L^-T J_r',  J M^-1 J' = Y Y',  J M^-1 f = Y (D^-1/2 L^-T f)  and
    M^-1 (f + J' lam) = L^-1 D^-1/2 (y_s + Y' lam)."""
    mt = common.tables()
    anc, _ = _chains(mt)
    rng = np.random.default_rng(2)
    q, v = common.standing_states(rng, 1)
    o, qLD, M = _factor(mt, q[0], v[0])
    assert o.d.nefc > 0
    J = np.array([o.d.arr("efc_J")[r][: mt.nv] for r in range(o.d.nefc)])
    L, D = _dense_L_D(mt, qLD, anc)
    Y = (np.diag(D ** -0.5) @ np.linalg.solve(L.T, J.T)).T
    Minv = np.linalg.inv(M)
    assert np.abs(Y @ Y.T - J @ Minv @ J.T).max() < 1e-9 * np.abs(J @ Minv @ J.T).max()
    f = rng.normal(size=mt.nv)
    ys = D ** -0.5 * np.linalg.solve(L.T, f)
    assert np.abs(Y @ ys - J @ Minv @ f).max() < 1e-9 * max(1.0, np.abs(J @ Minv @ f).max())
    lam = rng.uniform(0, 1, size=o.d.nefc)
    qacc = np.linalg.solve(L, D ** -0.5 * (ys + Y.T @ lam))
    assert np.abs(qacc - Minv @ (f + J.T @ lam)).max() < 1e-9 * max(1.0, np.abs(qacc).max())


def test_pgs_sweep_cost_decrease_identity_and_scaled_rows():
    """pgs_sweeps_reg: (1) the cost decrease of a whole Gauss-Seidel sweep, -1/2 (f_e - f_s)' (res_e + res_s) with
    res = A f + b, equals the sum of MuJoCo's per-row decreases -(1/2 d^2 A_ii + d res_i); (2) keeping rows scaled by
    -1/A_ii (increment = max(-f, sres)) gives the same iterates as f <- max(0, f - res/A_ii)."""
    rng = np.random.default_rng(3)
    for n in (3, 9, 24):
        G = rng.normal(size=(n, n + 4)); A = G @ G.T + np.diag(rng.uniform(0.01, 0.1, n))
        b = rng.normal(size=n) * 3
        f = np.maximum(0, rng.normal(size=n)); f2 = f.copy()
        res = A @ f + b
        sres = -res / np.diag(A)
        for sweep in range(4):
            fs, rs = f.copy(), res.copy()
            per_row = 0.0
            for i in range(n):
                fnew = max(0.0, f[i] - res[i] / A[i, i]); d = fnew - f[i]
                per_row += -(0.5 * d * d * A[i, i] + d * res[i])
                f[i] = fnew; res += A[:, i] * d
                mine = max(-f2[i], sres[i])                      # scaled-row form
                f2[i] += mine; sres += (-A[:, i] / np.diag(A)) * mine
            whole = -0.5 * (f - fs) @ (res + rs)
            cost = lambda x: 0.5 * x @ A @ x + x @ b
            assert abs(whole - per_row) < 1e-10 * max(1.0, abs(per_row))
            assert abs(whole - (cost(fs) - cost(f))) < 1e-9 * max(1.0, abs(cost(fs)))
            assert np.abs(f - f2).max() < 1e-12 * max(1.0, np.abs(f).max())
            assert np.abs(sres + res / np.diag(A)).max() < 1e-10 * max(1.0, np.abs(sres).max())


def _qmul(a, b):
    return mjcf.quat_mul(a, b)


def _qrot(q, v):
    u = q[1:]
    t = 2 * np.cross(u, v)
    return v + q[0] * t + np.cross(u, t)


def test_kinematics_as_prefix_composition():
    """kinematics(): world poses = prefix composition of the local transforms (body_pos, body_quat * hinges) along
    the body chains by pointer jumping (2 rounds for depth 4); world hinge axes by peeling the hinges off the final
    body quaternion.  Compared with the sequential mj_kinematics restatement (mjcf.np_kinematics)."""
    mt = common.tables()
    rng = np.random.default_rng(4)
    q, _ = common.airborne_states(rng, 3, frac=0.8)
    for qpos in q:
        xpos, xquat, xmat, xipos, xaxis = mjcf.np_kinematics(mt, qpos)
        nb = mt.nbody
        pos = np.zeros((nb, 3)); quat = np.tile([1.0, 0, 0, 0], (nb, 1)); hinges = [[] for _ in range(nb)]
        for b in range(1, nb):
            pos[b], quat[b] = mt.body_pos[b], mt.body_quat[b]
            for j in range(mt.body_jntadr[b], mt.body_jntadr[b] + mt.body_jntnum[b]):
                qa = mt.jnt_qposadr[j]
                if mt.jnt_type[j] == mjcf.JNT_FREE:
                    pos[b] = qpos[qa:qa + 3]; quat[b] = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
                else:
                    ang = qpos[qa] - mt.qpos0[qa]
                    ql = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * mt.jnt_axis[j]])
                    quat[b] = _qmul(quat[b], ql); hinges.append(None); hinges[b].append((j, ql))
        jump = np.full((3, nb), -1)
        for b in range(1, nb):
            for s in range(3):
                a = b
                for _ in range(1 << s):
                    a = mt.body_parent[a] if a > 0 else 0
                jump[s, b] = a if a > 0 else -1
        depth = int(max(mt.body_depth))
        for s in range(3):
            if (1 << s) >= depth:
                break
            p0, q0 = pos.copy(), quat.copy()
            for b in range(1, nb):
                src = jump[s, b]
                if src >= 0:
                    pos[b] = p0[src] + _qrot(q0[src], p0[b]); quat[b] = _qmul(q0[src], q0[b])
        quat /= np.linalg.norm(quat, axis=1, keepdims=True)
        assert np.abs(pos - xpos).max() < 1e-12 and np.abs(quat - xquat).max() < 1e-12
        for b in range(1, nb):
            qq = quat[b].copy()
            for j, ql in reversed(hinges[b]):
                assert np.abs(_qrot(qq, mt.jnt_axis[j]) - xaxis[j]).max() < 1e-12
                qq = _qmul(qq, ql * np.array([1, -1, -1, -1]))


def test_chain_scan_equals_ancestor_sums():
    """smooth_forces / chain_scan6: inclusive sums along the dof ancestor chains by 4 rounds of pointer jumping."""
    mt = common.tables()
    anc, _ = _chains(mt)
    rng = np.random.default_rng(6)
    w = rng.normal(size=(mt.nv, 6))
    x = w.copy()
    for s in range(4):
        snap = x.copy()
        for d in range(mt.nv):
            if (1 << s) <= len(anc[d]):
                x[d] += snap[anc[d][(1 << s) - 1]]
    for d in range(mt.nv):
        assert np.abs(x[d] - (w[d] + sum(w[a] for a in anc[d]))).max() < 1e-12


def test_slerp_chord_angle_and_nlerp_fallback():
    """quat_slerp (csrc/dmb.cu): the angle from the chord |q1 - q0| = 2 sin(theta/2) equals acos(dot), and below
    0.01 rad normalised lerp deviates from slerp by less than 2e-8 rad."""
    from deepmimic_mujoco_b200.mocap import quat_slerp
    rng = np.random.default_rng(7)
    for th in (1e-4, 1e-3, 9e-3, 0.3, 1.2):
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        q0 = rng.normal(size=4); q0 /= np.linalg.norm(q0)
        dq = np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * ax])
        q1 = mjcf.quat_mul(q0, dq)
        ang = 2 * np.arcsin(min(1.0, 0.5 * np.linalg.norm(q1 - q0)))
        assert abs(ang - np.arccos(np.clip(q0 @ q1, -1, 1))) < 1e-7 and abs(ang - th / 2) < 1e-9
        for f in (0.25, 0.5, 0.9):
            s = quat_slerp(q0, q1, f)
            if th < 0.01:
                n = (1 - f) * q0 + f * q1; n /= np.linalg.norm(n)
                if n @ s < 0:
                    n = -n
                assert 4 * np.arcsin(0.5 * np.linalg.norm(n - s)) < 2e-8     # rotation angle between the two results
