#!/usr/bin/env python3
"""Generate tests/golden/learner_golden.npz by RUNNING the reference's own learner-side numpy routines here (the
SURVEY 8(f) rows rebuilt in rollout.py / trpo.py are checked against these instead of against restatements):

* ``cg`` -- /root/reference/src/cg.py:2-34, imported as a module (pure numpy);
* ``add_vtarg_and_adv`` -- /root/reference/src/trpo.py:83-94.  trpo.py imports TensorFlow / mpi4py at module level
  and cannot be imported, so the function's own source lines are cut out of the file with ``ast`` and executed
  unchanged in a namespace that holds only numpy;
* ``explained_variance`` -- /root/reference/src/utils/math_util.py:25-38, cut out the same way (its module imports scipy
  only, but lives in a package whose __init__ pulls TensorFlow);
* ``MpiAdam.update`` -- /root/reference/src/mpi_adam.py:21-35, the method's source lines bound to a stub object (one-process
  communicator, flat get / set closures): seven steps on seeded gradients.
Run in the build container only."""
import ast
import os
import sys

import numpy as np

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut_function(path, name):
    """The reference's own function object, from its unmodified source text."""
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main():
    sys.path.insert(0, REF_SRC)
    from cg import cg                                                       # the reference module itself
    gae = cut_function(os.path.join(REF_SRC, "trpo.py"), "add_vtarg_and_adv")
    ev = cut_function(os.path.join(REF_SRC, "utils", "math_util.py"), "explained_variance")
    rng = np.random.default_rng(0)
    out = {}
    # GAE: 6 single-env segments of 41 steps with episode starts sprinkled in (trpo.py keeps one env)
    T, N = 41, 6
    rew = rng.normal(size=(T, N)).astype(np.float32)
    vpred = rng.normal(size=(T, N)).astype(np.float32)
    new = (rng.uniform(size=(T, N)) < 0.12).astype(np.int32)
    new[0] = 1
    nextvpred = rng.normal(size=N).astype(np.float32)
    adv, ret = np.zeros((T, N), np.float32), np.zeros((T, N), np.float32)
    for i in range(N):
        seg = {"new": new[:, i].copy(), "vpred": vpred[:, i].copy(), "nextvpred": nextvpred[i], "rew": rew[:, i].copy()}
        gae(seg, 0.995, 0.97)
        adv[:, i], ret[:, i] = seg["adv"], seg["tdlamret"]
    out.update(gae_rew=rew, gae_vpred=vpred, gae_new=new, gae_nextvpred=nextvpred, gae_adv=adv, gae_tdlamret=ret,
               gae_gamma=np.float64(0.995), gae_lam=np.float64(0.97))
    # CG: SPD system, the reference's 10 iterations (not converged) and a run that hits residual_tol
    n = 24
    A = rng.normal(size=(n, n)); A = A @ A.T + 0.5 * np.eye(n)
    b = rng.normal(size=n)
    out.update(cg_A=A, cg_b=b, cg_x10=cg(lambda p: A @ p, b, cg_iters=10), cg_x100=cg(lambda p: A @ p, b, cg_iters=100))
    # explained variance
    y = rng.normal(size=200); yp = 0.7 * y + 0.3 * rng.normal(size=200)
    out.update(ev_y=y, ev_ypred=yp, ev=np.float64(ev(yp, y)), ev_const=np.float64(ev(yp, np.ones(200))))
    # MpiAdam.update (mpi_adam.py:21-35): the method's own source lines, bound to a stub that supplies the three things it
    # touches outside numpy -- a one-process communicator and the flat get / set closures of the variable list
    src = open(os.path.join(REF_SRC, "mpi_adam.py")).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "MpiAdam")
    upd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "update")
    import types
    ns = {"np": np, "MPI": types.SimpleNamespace(SUM="sum")}             # the one mpi4py name the method mentions
    exec(compile(ast.Module(body=[upd], type_ignores=[]), "mpi_adam.py", "exec"), ns)

    class Comm:
        def Allreduce(self, a, b, op=None):
            b[:] = a

        def Get_size(self):
            return 1

    class Stub:
        beta1, beta2, epsilon, scale_grad_by_procs, t = 0.9, 0.999, 1e-08, True, 0

        def __init__(self, theta):
            self.theta = theta.astype("float32")
            self.m = np.zeros(theta.size, "float32"); self.v = np.zeros(theta.size, "float32"); self.comm = Comm()
            self.getflat = lambda: self.theta
            self.setfromflat = lambda x: setattr(self, "theta", np.asarray(x, "float32"))

        def check_synced(self):
            pass
    theta0 = rng.normal(size=40).astype("float32")
    grads = rng.normal(size=(7, 40)).astype("float32") * np.array([1, 1e-3, 10, 1, 0, 1, 1], "float32")[:, None]
    st, traj = Stub(theta0), []
    for gk in grads:
        ns["update"](st, gk, 1e-3)
        traj.append(st.theta.copy())
    out.update(adam_theta0=theta0, adam_grads=grads, adam_traj=np.asarray(traj))
    np.savez_compressed(os.path.join(HERE, "learner_golden.npz"), **out)
    print("gae adv[0]", adv[0], "\ncg residual 10 / 100 iters", np.linalg.norm(A @ out["cg_x10"] - b),
          np.linalg.norm(A @ out["cg_x100"] - b), "\nev", out["ev"], out["ev_const"])


if __name__ == "__main__":
    main()
