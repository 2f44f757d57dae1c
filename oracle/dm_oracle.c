/* dm_oracle.c -- CPU float64 ORACLE (test infrastructure only; see dm_oracle.h).
 *
 * PARITY UNPINNED (no MuJoCo binary, no reference golden vectors).
 *
 * Restates, for the MJCF subset of dp_env_v3.xml, the pipeline that the reference
 * reaches through gym MujocoEnv.do_simulation -> mujoco_py MjSim.step -> mj_step
 * (/root/reference/src/dp_env_v3.py:106-132), following MuJoCo's documented
 * computation stages:
 *   mj_step = mj_forward + mj_RungeKutta(4)
 *   mj_forward = fwdPosition(kinematics, comPos, crb, factorM, collision, makeConstraint)
 *              + fwdVelocity(comVel, passive, rne) + fwdActuation + fwdAcceleration
 *              + fwdConstraint(PGS)
 * Each function cites the MuJoCo routine it restates.  Single-threaded, one env.
 */
#include "dm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999

/* ------------------------------------------------------------------ small math */
static inline double dot3(const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
/* mju_normalize3: returns the norm; tiny vectors become (1,0,0) */
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
static inline void normalize4(double* q) {
  double n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else { q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n; }
}
static inline void mul_quat(double* r, const double* a, const double* b) {
  double w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
  double x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
  double y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
  double z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void quat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w*w + x*x - y*y - z*z; m[1] = 2*(x*y - w*z); m[2] = 2*(x*z + w*y);
  m[3] = 2*(x*y + w*z); m[4] = w*w - x*x + y*y - z*z; m[5] = 2*(y*z - w*x);
  m[6] = 2*(x*z - w*y); m[7] = 2*(y*z + w*x); m[8] = w*w - x*x - y*y + z*z;
}
static inline void axis_angle2quat(double* q, const double* axis, double angle) {
  double s = sin(0.5*angle);
  q[0] = cos(0.5*angle); q[1] = axis[0]*s; q[2] = axis[1]*s; q[3] = axis[2]*s;
}
static inline void mat_vec(double* r, const double* m, const double* v) { /* r = M v */
  double x = m[0]*v[0] + m[1]*v[1] + m[2]*v[2];
  double y = m[3]*v[0] + m[4]*v[1] + m[5]*v[2];
  double z = m[6]*v[0] + m[7]*v[1] + m[8]*v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void matT_vec(double* r, const double* m, const double* v) { /* r = M' v */
  double x = m[0]*v[0] + m[3]*v[1] + m[6]*v[2];
  double y = m[1]*v[0] + m[4]*v[1] + m[7]*v[2];
  double z = m[2]*v[0] + m[5]*v[1] + m[8]*v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
/* spatial algebra, vectors are [angular(3); linear(3)] (MuJoCo convention) */
static inline void cross_motion(double* r, const double* v, const double* s) {
  double a[3], b[3], c[3];
  cross3(a, v, s); cross3(b, v, s + 3); cross3(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static inline void cross_force(double* r, const double* v, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
/* mju_mulInertVec: 10-number c-frame inertia times spatial motion vector */
static inline void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0]*v[0] + i[3]*v[1] + i[4]*v[2] - i[8]*v[4] + i[7]*v[5];
  r[1] = i[3]*v[0] + i[1]*v[1] + i[5]*v[2] + i[8]*v[3] - i[6]*v[5];
  r[2] = i[4]*v[0] + i[5]*v[1] + i[2]*v[2] - i[7]*v[3] + i[6]*v[4];
  r[3] = i[8]*v[1] - i[7]*v[2] + i[9]*v[3];
  r[4] = i[6]*v[2] - i[8]*v[0] + i[9]*v[4];
  r[5] = i[7]*v[0] - i[6]*v[1] + i[9]*v[5];
}

int dmo_version(void) { return 1; }
unsigned long dmo_sizeof_data(void) { return sizeof(dmo_data_t); }
unsigned long dmo_sizeof_env(void) { return sizeof(dmo_env_t); }
unsigned long dmo_sizeof_model(void) { return sizeof(dmb_model_t); }

/* ------------------------------------------------------------------ mj_kinematics */
void dmo_kinematics(const dmb_model_t* m, dmo_data_t* d) {
  memset(d->xpos[0], 0, sizeof(d->xpos[0]));
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2mat(d->xmat[0], d->xquat[0]);
  memset(d->xipos[0], 0, sizeof(d->xipos[0]));
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b];
    double pos[3], quat[4], tmp[3];
    mat_vec(tmp, d->xmat[p], m->body_pos[b]);
    for (int k = 0; k < 3; k++) pos[k] = d->xpos[p][k] + tmp[k];
    mul_quat(quat, d->xquat[p], m->body_quat[b]);
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int qa = m->jnt_qposadr[j];
      if (m->jnt_type[j] == DMB_JNT_FREE) {
        for (int k = 0; k < 3; k++) pos[k] = d->qpos[qa + k];
        for (int k = 0; k < 4; k++) quat[k] = d->qpos[qa + 3 + k];
        normalize4(quat);
        d->xaxis[j][0] = 0; d->xaxis[j][1] = 0; d->xaxis[j][2] = 1;
      } else { /* hinge, anchored at the body origin (jnt_pos == 0) */
        double mat[9], ql[4], qn[4];
        quat2mat(mat, quat);
        mat_vec(d->xaxis[j], mat, m->jnt_axis[j]);
        axis_angle2quat(ql, m->jnt_axis[j], d->qpos[qa] - m->qpos0[qa]);
        mul_quat(qn, quat, ql);
        memcpy(quat, qn, sizeof(qn));
      }
    }
    normalize4(quat);
    memcpy(d->xpos[b], pos, sizeof(pos));
    memcpy(d->xquat[b], quat, sizeof(quat));
    quat2mat(d->xmat[b], quat);
    mat_vec(tmp, d->xmat[b], m->body_ipos[b]);
    for (int k = 0; k < 3; k++) d->xipos[b][k] = pos[k] + tmp[k];
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double tmp[3], q[4];
    mat_vec(tmp, d->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) d->geom_xpos[g][k] = d->xpos[b][k] + tmp[k];
    mul_quat(q, d->xquat[b], m->geom_quat[g]);
    quat2mat(d->geom_xmat[g], q);
  }
}

/* ------------------------------------------------------------------ mj_comPos */
static void com_pos(const dmb_model_t* m, dmo_data_t* d) {
  /* single kinematic tree rooted at body 1: c-frame origin = whole-model CoM */
  double mass = 0, c[3] = {0, 0, 0};
  for (int b = 1; b < m->nbody; b++) {
    mass += m->body_mass[b];
    for (int k = 0; k < 3; k++) c[k] += m->body_mass[b] * d->xipos[b][k];
  }
  for (int k = 0; k < 3; k++) d->com[k] = c[k] / (mass > MINVAL ? mass : MINVAL);
  memset(d->cinert[0], 0, sizeof(d->cinert[0]));
  for (int b = 1; b < m->nbody; b++) {
    /* mju_inertCom with a full body-frame tensor: I_world = R I R', then parallel axis */
    const double* I = m->body_inertia[b];
    const double* R = d->xmat[b];
    double Ib[9] = {I[0], I[3], I[4], I[3], I[1], I[5], I[4], I[5], I[2]};
    double RI[9], W[9];
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) {
      RI[3*r + cc] = R[3*r]*Ib[cc] + R[3*r + 1]*Ib[3 + cc] + R[3*r + 2]*Ib[6 + cc];
    }
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) {
      W[3*r + cc] = RI[3*r]*R[3*cc] + RI[3*r + 1]*R[3*cc + 1] + RI[3*r + 2]*R[3*cc + 2];
    }
    double dif[3], ms = m->body_mass[b];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[b][k] - d->com[k];
    double* ci = d->cinert[b];
    ci[0] = W[0] + ms*(dif[1]*dif[1] + dif[2]*dif[2]);
    ci[1] = W[4] + ms*(dif[0]*dif[0] + dif[2]*dif[2]);
    ci[2] = W[8] + ms*(dif[0]*dif[0] + dif[1]*dif[1]);
    ci[3] = W[1] - ms*dif[0]*dif[1];
    ci[4] = W[2] - ms*dif[0]*dif[2];
    ci[5] = W[5] - ms*dif[1]*dif[2];
    ci[6] = ms*dif[0]; ci[7] = ms*dif[1]; ci[8] = ms*dif[2]; ci[9] = ms;
  }
  /* cdof: motion axis of each dof expressed at the c-frame origin */
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->com[k] - d->xpos[b][k]; /* anchor == body origin */
    if (m->jnt_type[j] == DMB_JNT_FREE) {
      for (int i = 0; i < 3; i++) {
        memset(d->cdof[da + i], 0, sizeof(d->cdof[0]));
        d->cdof[da + i][3 + i] = 1;
      }
      for (int i = 0; i < 3; i++) {
        double ax[3] = {d->xmat[b][i], d->xmat[b][3 + i], d->xmat[b][6 + i]};
        memcpy(d->cdof[da + 3 + i], ax, sizeof(ax));
        cross3(d->cdof[da + 3 + i] + 3, ax, off);
      }
    } else {
      memcpy(d->cdof[da], d->xaxis[j], 3*sizeof(double));
      cross3(d->cdof[da] + 3, d->xaxis[j], off);
    }
  }
}

/* ------------------------------------------------------------------ mj_crb + mj_factorM */
static void crb_factor(const dmb_model_t* m, dmo_data_t* d) {
  memcpy(d->crb, d->cinert, sizeof(d->crb));
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
  }
  memset(d->qM, 0, sizeof(d->qM));
  for (int i = 0; i < m->nv; i++) {
    double buf[6];
    int adr = m->dof_Madr[i];
    d->qM[adr] = m->dof_armature[i];
    mul_inert_vec(buf, d->crb[m->dof_bodyid[i]], d->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[j][k] * buf[k];
      d->qM[adr++] += s;
    }
  }
  /* L'DL in the sparse layout (mj_factorM) */
  memcpy(d->qLD, d->qM, sizeof(d->qM));
  for (int k = m->nv - 1; k >= 0; k--) {
    int Madr_kk = m->dof_Madr[k];
    int Madr_ki = Madr_kk + 1;
    int i = m->dof_parentid[k];
    while (i >= 0) {
      double tmp = d->qLD[Madr_ki] / d->qLD[Madr_kk];
      /* M(i, j) -= M(k,i)/M(k,k) * M(k, j) for j = i and its ancestors */
      int Madr_ij = m->dof_Madr[i], cnt = 0;
      for (int j = i; j >= 0; j = m->dof_parentid[j]) {
        d->qLD[Madr_ij + cnt] -= tmp * d->qLD[Madr_ki + cnt];
        cnt++;
      }
      d->qLD[Madr_ki] = tmp;
      i = m->dof_parentid[i];
      Madr_ki++;
    }
  }
  for (int i = 0; i < m->nv; i++) d->qLDiagInv[i] = 1.0 / d->qLD[m->dof_Madr[i]];
}

/* mj_solveLD: x <- inv(L'*D*L) x */
void dmo_solve_M(const dmb_model_t* m, const dmo_data_t* d, double* x) {
  for (int i = m->nv - 1; i >= 0; i--) {        /* x <- inv(L') x */
    if (x[i] != 0.0) {
      int adr = m->dof_Madr[i] + 1;
      for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= d->qLD[adr++] * x[i];
    }
  }
  for (int i = 0; i < m->nv; i++) x[i] *= d->qLDiagInv[i];
  for (int i = 0; i < m->nv; i++) {             /* x <- inv(L) x */
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= d->qLD[adr++] * x[j];
  }
}

void dmo_full_M(const dmb_model_t* m, const dmo_data_t* d, double* dense) {
  int nv = m->nv;
  memset(dense, 0, sizeof(double)*nv*nv);
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      dense[i*nv + j] = dense[j*nv + i] = d->qM[adr++];
    }
  }
}

/* ------------------------------------------------------------------ collision */
/* mju_makeFrame */
static void make_frame(double* f) {
  normalize3(f);
  if (norm3(f + 3) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double t = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= t*f[k];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

typedef struct { double dist, pos[3], normal[3], yaxis[3]; } rawcon_t;

static int raw_plane_sphere(rawcon_t* c, double margin, const double* ppos, const double* pnormal,
                            const double* spos, double r) {
  double dif[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  double cdist = dot3(dif, pnormal);
  if (cdist > margin + r) return 0;
  c->dist = cdist - r;
  for (int k = 0; k < 3; k++) {
    c->pos[k] = spos[k] - pnormal[k]*(r + 0.5*c->dist);
    c->normal[k] = pnormal[k];
    c->yaxis[k] = 0;
  }
  return 1;
}

static int raw_sphere_sphere(rawcon_t* c, double margin, const double* p1, double r1, const double* p2, double r2) {
  double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double dist = norm3(dif);
  if (dist > margin + r1 + r2) return 0;
  c->dist = dist - r1 - r2;
  if (dist < MINVAL) { c->normal[0] = 1; c->normal[1] = 0; c->normal[2] = 0; }
  else for (int k = 0; k < 3; k++) c->normal[k] = dif[k] / dist;
  for (int k = 0; k < 3; k++) {
    c->pos[k] = p1[k] + c->normal[k]*(r1 + 0.5*c->dist);
    c->yaxis[k] = 0;
  }
  return 1;
}

/* sphere (possibly radius 0 = point) against a box; normal points sphere -> box */
static int raw_sphere_box(rawcon_t* c, double margin, const double* spos, double r,
                          const double* bpos, const double* bmat, const double* bsize) {
  double tmp[3] = {spos[0] - bpos[0], spos[1] - bpos[1], spos[2] - bpos[2]};
  double center[3], clamped[3], deepest[3];
  matT_vec(center, bmat, tmp);
  for (int k = 0; k < 3; k++) {
    clamped[k] = center[k] > bsize[k] ? bsize[k] : (center[k] < -bsize[k] ? -bsize[k] : center[k]);
    deepest[k] = center[k] - clamped[k];
  }
  double dist = norm3(deepest);
  if (dist - r > margin) return 0;
  double nloc[3], ploc[3];
  if (dist <= MINVAL) { /* centre inside the box: push out through the nearest face */
    double closest = 1e30; int kk = 0;
    for (int k = 0; k < 3; k++) {
      double gap = bsize[k] - fabs(center[k]);
      if (gap < closest) { closest = gap; kk = k; }
    }
    double nf[3] = {0, 0, 0};
    nf[kk] = center[kk] > 0 ? 1.0 : -1.0;
    c->dist = -closest - r;
    for (int k = 0; k < 3; k++) { nloc[k] = -nf[k]; ploc[k] = center[k] + nf[k]*0.5*(closest - r); }
  } else {
    c->dist = dist - r;
    for (int k = 0; k < 3; k++) {
      double nout = deepest[k] / dist;
      nloc[k] = -nout;
      ploc[k] = clamped[k] + nout*0.5*c->dist;
    }
  }
  mat_vec(c->normal, bmat, nloc);
  mat_vec(c->pos, bmat, ploc);
  for (int k = 0; k < 3; k++) { c->pos[k] += bpos[k]; c->yaxis[k] = 0; }
  return 1;
}

/* d/dt of squared distance from the point c + t*u to the box (monotone in t) */
static double capbox_dfdt(const double* c, const double* u, const double* s, double t) {
  double g = 0;
  for (int k = 0; k < 3; k++) {
    double p = c[k] + t*u[k];
    if (p > s[k]) g += (p - s[k])*u[k];
    else if (p < -s[k]) g += (p + s[k])*u[k];
  }
  return g;
}

/* capsule against box (own restatement: exact closest point of the segment to the box by
 * root-finding on the piecewise-linear derivative, then sphere-box; optional 2nd contact at
 * the far segment end).  MuJoCo's mjc_CapsuleBox is not reproduced -- see DESIGN.md. */
static int raw_capsule_box(rawcon_t* c, double margin, const double* cpos, const double* cmat, double r, double h,
                           const double* bpos, const double* bmat, const double* bsize) {
  double axis[3] = {cmat[2], cmat[5], cmat[8]};
  double tmp[3] = {cpos[0] - bpos[0], cpos[1] - bpos[1], cpos[2] - bpos[2]};
  double cl[3], u[3];
  matT_vec(cl, bmat, tmp);
  matT_vec(u, bmat, axis);
  /* knots: segment ends and the <=6 parameters where a coordinate crosses a face plane */
  double knot[8]; int nk = 0;
  knot[nk++] = -h;
  for (int k = 0; k < 3; k++) {
    if (fabs(u[k]) > 1e-12) {
      double t1 = (bsize[k] - cl[k]) / u[k], t2 = (-bsize[k] - cl[k]) / u[k];
      if (t1 > -h && t1 < h) knot[nk++] = t1;
      if (t2 > -h && t2 < h) knot[nk++] = t2;
    }
  }
  knot[nk++] = h;
  for (int i = 1; i < nk; i++) { /* insertion sort */
    double v = knot[i]; int j = i - 1;
    while (j >= 0 && knot[j] > v) { knot[j + 1] = knot[j]; j--; }
    knot[j + 1] = v;
  }
  double tlo, thi;
  /* Does the axis run through the box?  Clip the segment against the three slabs (Liang-Barsky): inside the box
   * the distance is zero on the whole stretch [tin, tout], and looking for roots of its derivative there means
   * testing the sign of rounding noise (found by the full-size parity test: fp32 and fp64 disagreed). */
  double tin = -h, tout = h;
  for (int k = 0; k < 3; k++) {
    if (fabs(u[k]) > 1e-12) {
      double ta = (-bsize[k] - cl[k]) / u[k], tb = (bsize[k] - cl[k]) / u[k];
      if (ta > tb) { double t = ta; ta = tb; tb = t; }
      if (ta > tin) tin = ta;
      if (tb < tout) tout = tb;
    } else if (fabs(cl[k]) > bsize[k]) { tin = h; tout = -h; }
  }
  if (tout - tin > 1e-6*h) { tlo = tin; thi = tout; }
  else {
  { /* leftmost root */
    double gprev = capbox_dfdt(cl, u, bsize, knot[0]);
    if (gprev >= 0) tlo = knot[0];
    else {
      tlo = knot[nk - 1];
      for (int i = 1; i < nk; i++) {
        double g = capbox_dfdt(cl, u, bsize, knot[i]);
        if (g >= 0) { tlo = knot[i - 1] - gprev*(knot[i] - knot[i - 1])/(g - gprev); break; }
        gprev = g;
      }
    }
  }
  { /* rightmost root */
    double gnext = capbox_dfdt(cl, u, bsize, knot[nk - 1]);
    if (gnext <= 0) thi = knot[nk - 1];
    else {
      thi = knot[0];
      for (int i = nk - 2; i >= 0; i--) {
        double g = capbox_dfdt(cl, u, bsize, knot[i]);
        if (g <= 0) { thi = knot[i + 1] - gnext*(knot[i] - knot[i + 1])/(g - gnext); break; }
        gnext = g;
      }
    }
  }
  }
  /* a flat stretch of the derivative (the axis runs through the box, or slides parallel to a face): the closest
   * point is not unique.  Take the end nearer the capsule centre (the middle would sit equidistant from the two
   * faces a capsule pierces), moved into the stretch by 1e-3 of its length: the end itself lies exactly ON the box
   * surface, where the sphere-box normal is a 0/0 limit that fp32 and fp64 resolve differently (found by the
   * full-size parity test on spinkick: knee capsule through the other foot's box). */
  double ts = 0.5*(tlo + thi);
  if (thi - tlo > 1e-6*h) ts = fabs(tlo) <= fabs(thi) ? tlo + 1e-3*(thi - tlo) : thi - 1e-3*(thi - tlo);
  int n = 0;
  double sp[3];
  for (int k = 0; k < 3; k++) sp[k] = cpos[k] + ts*axis[k];
  n += raw_sphere_box(c + n, margin, sp, r, bpos, bmat, bsize);
  /* second contact: the segment end farther from ts */
  double te = (h - ts >= ts + h) ? h : -h;
  if (fabs(te - ts) > 0.01*h) {
    for (int k = 0; k < 3; k++) sp[k] = cpos[k] + te*axis[k];
    n += raw_sphere_box(c + n, margin, sp, r, bpos, bmat, bsize);
  }
  return n;
}

/* mjc_CapsuleCapsule: closest points of two segments (+ up to 2 contacts when parallel) */
static int raw_capsule_capsule(rawcon_t* c, double margin, const double* pos1, const double* mat1, double r1, double h1,
                               const double* pos2, const double* mat2, double r2, double h2) {
  double axis1[3] = {mat1[2], mat1[5], mat1[8]}, axis2[3] = {mat2[2], mat2[5], mat2[8]};
  double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  double ma = dot3(axis1, axis1), mb = -dot3(axis1, axis2), mc = dot3(axis2, axis2);
  double u = -dot3(axis1, dif), v = dot3(axis2, dif);
  double det = ma*mc - mb*mb;
  double vec1[3], vec2[3];
  if (fabs(det) >= 1e-10) { /* MuJoCo: mjMINVAL; loosened so fp32 takes the same branch */
    double x1 = (mc*u - mb*v)/det, x2 = (ma*v - mb*u)/det;
    if (x1 > h1) { x1 = h1; x2 = (v - mb*h1)/mc; }
    else if (x1 < -h1) { x1 = -h1; x2 = (v + mb*h1)/mc; }
    if (x2 > h2) { x2 = h2; x1 = (u - mb*h2)/ma; }
    else if (x2 < -h2) { x2 = -h2; x1 = (u + mb*h2)/ma; }
    if (x1 > h1) x1 = h1; else if (x1 < -h1) x1 = -h1;
    for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k]*x1; vec2[k] = pos2[k] + axis2[k]*x2; }
    return raw_sphere_sphere(c, margin, vec1, r1, vec2, r2);
  }
  int n = 0;
  double x1, x2;
  /* x1 = +h1 */
  x2 = (v - mb*h1)/mc; if (x2 > h2) x2 = h2; else if (x2 < -h2) x2 = -h2;
  for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k]*h1; vec2[k] = pos2[k] + axis2[k]*x2; }
  n += raw_sphere_sphere(c + n, margin, vec1, r1, vec2, r2);
  /* x1 = -h1 */
  x2 = (v + mb*h1)/mc; if (x2 > h2) x2 = h2; else if (x2 < -h2) x2 = -h2;
  for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] - axis1[k]*h1; vec2[k] = pos2[k] + axis2[k]*x2; }
  n += raw_sphere_sphere(c + n, margin, vec1, r1, vec2, r2);
  if (n == 2) return n;
  /* x2 = +h2 */
  x1 = (u - mb*h2)/ma; if (x1 > h1) x1 = h1; else if (x1 < -h1) x1 = -h1;
  for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k]*x1; vec2[k] = pos2[k] + axis2[k]*h2; }
  n += raw_sphere_sphere(c + n, margin, vec1, r1, vec2, r2);
  if (n == 2) return n;
  /* x2 = -h2 */
  x1 = (u + mb*h2)/ma; if (x1 > h1) x1 = h1; else if (x1 < -h1) x1 = -h1;
  for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k]*x1; vec2[k] = pos2[k] - axis2[k]*h2; }
  n += raw_sphere_sphere(c + n, margin, vec1, r1, vec2, r2);
  return n;
}

/* box-box (own restatement: vertex-in-box manifold, first 4 in fixed order, no edge-edge).
 * MuJoCo's mjc_BoxBox is not reproduced -- see DESIGN.md. */
static int raw_box_box(rawcon_t* c, double margin, const double* pos1, const double* mat1, const double* size1,
                       const double* pos2, const double* mat2, const double* size2) {
  int n = 0;
  /* exact early-out: if the boxes are separated by more than the margin along a face axis of either
   * box, no vertex of one can be within the margin of the other */
  for (int pass = 0; pass < 2; pass++) {
    const double *pa = pass == 0 ? pos1 : pos2, *ma = pass == 0 ? mat1 : mat2, *sa = pass == 0 ? size1 : size2;
    const double *pb = pass == 0 ? pos2 : pos1, *mb = pass == 0 ? mat2 : mat1, *sb = pass == 0 ? size2 : size1;
    double dc[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    for (int k = 0; k < 3; k++) {
      double ax[3] = {ma[k], ma[3 + k], ma[6 + k]};
      double r = 0;
      for (int j = 0; j < 3; j++) r += sb[j]*fabs(ax[0]*mb[j] + ax[1]*mb[3 + j] + ax[2]*mb[6 + j]);
      if (fabs(dot3(dc, ax)) - sa[k] - r > margin) return 0;
    }
  }
  for (int pass = 0; pass < 2 && n < 4; pass++) {
    const double* vp = pass == 0 ? pos2 : pos1;   /* owner of the vertices */
    const double* vm = pass == 0 ? mat2 : mat1;
    const double* vs = pass == 0 ? size2 : size1;
    const double* bp = pass == 0 ? pos1 : pos2;   /* box being tested against */
    const double* bm = pass == 0 ? mat1 : mat2;
    const double* bs = pass == 0 ? size1 : size2;
    for (int i = 0; i < 8 && n < 4; i++) {
      double loc[3] = {(i & 1) ? vs[0] : -vs[0], (i & 2) ? vs[1] : -vs[1], (i & 4) ? vs[2] : -vs[2]};
      double w[3];
      mat_vec(w, vm, loc);
      for (int k = 0; k < 3; k++) w[k] += vp[k];
      if (raw_sphere_box(c + n, margin, w, 0.0, bp, bm, bs)) {
        /* raw normal is vertex -> box; contact normal must be geom1 -> geom2 */
        if (pass == 0) for (int k = 0; k < 3; k++) c[n].normal[k] = -c[n].normal[k];
        n++;
      }
    }
  }
  return n;
}

static void collision(const dmb_model_t* m, dmo_data_t* d) {
  d->ncon = 0;
  double margin = m->margin;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    const double *pos1 = d->geom_xpos[g1], *pos2 = d->geom_xpos[g2];
    const double *mat1 = d->geom_xmat[g1], *mat2 = d->geom_xmat[g2];
    const double *s1 = m->geom_size[g1], *s2 = m->geom_size[g2];
    rawcon_t rc[4];
    int n = 0;
    if (t1 == DMB_GEOM_PLANE) {
      double nrm[3] = {mat1[2], mat1[5], mat1[8]};
      double dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
      if (dot3(dif, nrm) > m->geom_rbound[g2] + margin) continue; /* bounding-sphere filter */
      if (t2 == DMB_GEOM_SPHERE) {
        n = raw_plane_sphere(rc, margin, pos1, nrm, pos2, s2[0]);
      } else if (t2 == DMB_GEOM_CAPSULE) { /* mjc_PlaneCapsule: two end spheres, frame y-axis along the capsule */
        double axis[3] = {mat2[2], mat2[5], mat2[8]}, e[3];
        for (int k = 0; k < 3; k++) e[k] = pos2[k] + axis[k]*s2[1];
        n += raw_plane_sphere(rc + n, margin, pos1, nrm, e, s2[0]);
        for (int k = 0; k < 3; k++) e[k] = pos2[k] - axis[k]*s2[1];
        n += raw_plane_sphere(rc + n, margin, pos1, nrm, e, s2[0]);
        for (int i = 0; i < n; i++) memcpy(rc[i].yaxis, axis, sizeof(axis));
      } else if (t2 == DMB_GEOM_BOX) { /* mjc_PlaneBox: corners in the lower half-space, first 4 */
        double dist = dot3(dif, nrm);
        for (int i = 0; i < 8 && n < 4; i++) {
          double vec[3] = {(i & 1) ? s2[0] : -s2[0], (i & 2) ? s2[1] : -s2[1], (i & 4) ? s2[2] : -s2[2]};
          double corner[3];
          mat_vec(corner, mat2, vec);
          double ldist = dot3(nrm, corner);
          if (dist + ldist > margin || ldist > 0) continue;
          rc[n].dist = dist + ldist;
          for (int k = 0; k < 3; k++) {
            rc[n].pos[k] = corner[k] + pos2[k] - nrm[k]*0.5*rc[n].dist;
            rc[n].normal[k] = nrm[k];
            rc[n].yaxis[k] = 0;
          }
          n++;
        }
      }
    } else {
      double dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
      double bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      if (dot3(dif, dif) > bound*bound) continue;
      if (t1 == DMB_GEOM_SPHERE && t2 == DMB_GEOM_SPHERE) {
        n = raw_sphere_sphere(rc, margin, pos1, s1[0], pos2, s2[0]);
      } else if (t1 == DMB_GEOM_SPHERE && t2 == DMB_GEOM_CAPSULE) { /* mjc_SphereCapsule */
        double axis[3] = {mat2[2], mat2[5], mat2[8]};
        double vec[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
        double x = dot3(axis, vec);
        x = x > s2[1] ? s2[1] : (x < -s2[1] ? -s2[1] : x);
        double cp[3] = {pos2[0] + axis[0]*x, pos2[1] + axis[1]*x, pos2[2] + axis[2]*x};
        n = raw_sphere_sphere(rc, margin, pos1, s1[0], cp, s2[0]);
      } else if (t1 == DMB_GEOM_SPHERE && t2 == DMB_GEOM_BOX) {
        n = raw_sphere_box(rc, margin, pos1, s1[0], pos2, mat2, s2);
      } else if (t1 == DMB_GEOM_CAPSULE && t2 == DMB_GEOM_CAPSULE) {
        n = raw_capsule_capsule(rc, margin, pos1, mat1, s1[0], s1[1], pos2, mat2, s2[0], s2[1]);
      } else if (t1 == DMB_GEOM_CAPSULE && t2 == DMB_GEOM_BOX) {
        n = raw_capsule_box(rc, margin, pos1, mat1, s1[0], s1[1], pos2, mat2, s2);
      } else if (t1 == DMB_GEOM_BOX && t2 == DMB_GEOM_BOX) {
        n = raw_box_box(rc, margin, pos1, mat1, s1, pos2, mat2, s2);
      }
    }
    for (int i = 0; i < n; i++) {
      if (d->ncon >= m->max_con || d->ncon >= DMO_MAXCON) { d->flags |= 1; return; }
      dmo_contact_t* c = &d->contact[d->ncon++];
      c->dist = rc[i].dist;
      memcpy(c->pos, rc[i].pos, sizeof(c->pos));
      memcpy(c->frame, rc[i].normal, 3*sizeof(double));
      memcpy(c->frame + 3, rc[i].yaxis, 3*sizeof(double));
      make_frame(c->frame);
      c->geom1 = g1; c->geom2 = g2;
      int cd1 = m->geom_condim[g1], cd2 = m->geom_condim[g2];
      c->dim = cd1 > cd2 ? cd1 : cd2;                         /* max condim (equal priorities) */
      double f1 = m->geom_friction[g1][0], f2 = m->geom_friction[g2][0];
      c->mu = f1 > f2 ? f1 : f2;                              /* max friction */
      c->efc_address = -1;
    }
  }
}

/* ------------------------------------------------------------------ mj_makeConstraint */
/* translational Jacobian difference (body2 - body1) at a world point, projected on `dir` */
static void jac_point_dir(const dmb_model_t* m, const dmo_data_t* d, int body, const double* point,
                          const double* dir, double sign, double* row) {
  double off[3] = {point[0] - d->com[0], point[1] - d->com[1], point[2] - d->com[2]};
  for (int b = body; b > 0; b = m->body_parent[b]) {
    for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++) {
      double tmp[3];
      cross3(tmp, d->cdof[i], off);
      for (int k = 0; k < 3; k++) tmp[k] += d->cdof[i][3 + k];
      row[i] += sign * dot3(dir, tmp);
    }
  }
}

static void get_impedance(const double* solimp, double pos, double margin, double* imp) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = dmin < MINIMP ? MINIMP : (dmin > MAXIMP ? MAXIMP : dmin);
  dmax = dmax < MINIMP ? MINIMP : (dmax > MAXIMP ? MAXIMP : dmax);
  mid = mid < MINIMP ? MINIMP : (mid > MAXIMP ? MAXIMP : mid);
  if (power < 1) power = 1;
  if (dmin == dmax || width <= MINVAL) { *imp = 0.5*(dmin + dmax); return; }
  double x = (pos - margin)/width;
  if (x < 0) x = -x;
  if (x >= 1) { *imp = dmax; return; }
  if (x == 0) { *imp = dmin; return; }
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  *imp = dmin + y*(dmax - dmin);
}

static void make_constraint(const dmb_model_t* m, dmo_data_t* d) {
  int nv = m->nv;
  d->nefc = 0;
  int maxefc = m->max_efc < DMO_MAXEFC ? m->max_efc : DMO_MAXEFC;
  /* joint limits (mj_instantiateLimit): lower side then upper side, joint order */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] != DMB_JNT_HINGE) continue;
    double value = d->qpos[m->jnt_qposadr[j]];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[j][side < 0 ? 0 : 1] - value);
      if (dist < 0) { /* jnt_margin = 0 */
        if (d->nefc >= maxefc) { d->flags |= 2; continue; }
        int r = d->nefc++;
        memset(d->efc_J[r], 0, sizeof(d->efc_J[r]));
        d->efc_J[r][m->jnt_dofadr[j]] = -side;
        d->efc_pos[r] = dist; d->efc_margin[r] = 0;
        d->efc_type[r] = 0; d->efc_id[r] = j;
        d->efc_diagApprox[r] = m->dof_invweight0[m->jnt_dofadr[j]];
      }
    }
  }
  /* contacts (mj_instantiateContact): frictionless (dim 1) or pyramidal (dim 3 -> 4 rows) */
  for (int ci = 0; ci < d->ncon; ci++) {
    dmo_contact_t* c = &d->contact[ci];
    int nrow = c->dim == 1 ? 1 : 2*(c->dim - 1);
    if (d->nefc + nrow > maxefc) { d->flags |= 2; d->ncon = ci; break; }
    int b1 = m->geom_bodyid[c->geom1], b2 = m->geom_bodyid[c->geom2];
    double jn[3][DMB_MAX_DOF];
    memset(jn, 0, sizeof(jn));
    int nax = c->dim == 1 ? 1 : 3;
    for (int a = 0; a < nax; a++) {
      jac_point_dir(m, d, b2, c->pos, c->frame + 3*a, 1.0, jn[a]);
      jac_point_dir(m, d, b1, c->pos, c->frame + 3*a, -1.0, jn[a]);
    }
    double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    c->efc_address = d->nefc;
    if (c->dim == 1) {
      int r = d->nefc++;
      memcpy(d->efc_J[r], jn[0], sizeof(d->efc_J[r]));
      d->efc_pos[r] = c->dist; d->efc_margin[r] = m->margin;
      d->efc_type[r] = 1; d->efc_id[r] = ci;
      d->efc_diagApprox[r] = tran;
    } else {
      for (int k = 1; k < c->dim; k++) {
        for (int s = 0; s < 2; s++) {
          int r = d->nefc++;
          double sg = s == 0 ? c->mu : -c->mu;
          for (int i = 0; i < nv; i++) d->efc_J[r][i] = jn[0][i] + sg*jn[k][i];
          for (int i = nv; i < DMB_MAX_DOF; i++) d->efc_J[r][i] = 0;
          d->efc_pos[r] = c->dist; d->efc_margin[r] = m->margin;
          d->efc_type[r] = 2; d->efc_id[r] = ci;
          d->efc_diagApprox[r] = tran + c->mu*c->mu*tran;
        }
      }
    }
  }
  /* mj_makeImpedance */
  double tc = m->solref[0], dr = m->solref[1];
  if (tc < 2*m->timestep) tc = 2*m->timestep; /* refsafe */
  double dmax = m->solimp[1];
  dmax = dmax < MINIMP ? MINIMP : (dmax > MAXIMP ? MAXIMP : dmax);
  double kk = dmax*dmax*tc*tc*dr*dr, bb = dmax*tc;
  kk = 1.0 / (kk > MINVAL ? kk : MINVAL);
  bb = 2.0 / (bb > MINVAL ? bb : MINVAL);
  for (int r = 0; r < d->nefc; r++) {
    double imp;
    get_impedance(m->solimp, d->efc_pos[r], d->efc_margin[r], &imp);
    double R = (1 - imp)*d->efc_diagApprox[r]/imp;
    d->efc_R[r] = R > MINVAL ? R : MINVAL;
    d->efc_KBIP[r][0] = kk; d->efc_KBIP[r][1] = bb; d->efc_KBIP[r][2] = imp; d->efc_KBIP[r][3] = 0;
  }
  /* pyramidal contacts: common R = 2 mu^2 R[first row]  (impratio = 1) */
  for (int ci = 0; ci < d->ncon; ci++) {
    dmo_contact_t* c = &d->contact[ci];
    if (c->dim > 1 && c->efc_address >= 0) {
      int r0 = c->efc_address;
      double Rpy = 2*c->mu*c->mu*d->efc_R[r0];
      for (int k = 0; k < 2*(c->dim - 1); k++) d->efc_R[r0 + k] = Rpy;
    }
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1.0 / d->efc_R[r];
  /* mj_projectConstraint: AR = J M^-1 J' + diag(R) */
  static __thread double B[DMO_MAXEFC][DMB_MAX_DOF];
  for (int r = 0; r < d->nefc; r++) {
    memcpy(B[r], d->efc_J[r], sizeof(B[r]));
    dmo_solve_M(m, d, B[r]);
  }
  for (int r = 0; r < d->nefc; r++) {
    for (int s = 0; s <= r; s++) {
      double a = 0;
      for (int i = 0; i < nv; i++) a += d->efc_J[r][i]*B[s][i];
      d->efc_AR[r][s] = d->efc_AR[s][r] = a;
    }
    d->efc_AR[r][r] += d->efc_R[r];
  }
}

void dmo_fwd_position(const dmb_model_t* m, dmo_data_t* d) {
  dmo_kinematics(m, d);
  com_pos(m, d);
  crb_factor(m, d);
  collision(m, d);
  make_constraint(m, d);
}

/* ------------------------------------------------------------------ mj_fwdVelocity */
/* mj_comVel */
static void com_vel(const dmb_model_t* m, dmo_data_t* d) {
  memset(d->cvel[0], 0, sizeof(d->cvel[0]));
  for (int b = 1; b < m->nbody; b++) {
    double cvel[6];
    memcpy(cvel, d->cvel[m->body_parent[b]], sizeof(cvel));
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == DMB_JNT_FREE) {
        for (int i = 0; i < 3; i++) memset(d->cdof_dot[da + i], 0, sizeof(d->cdof_dot[0]));
        for (int i = 0; i < 3; i++) for (int k = 0; k < 6; k++) cvel[k] += d->cdof[da + i][k]*d->qvel[da + i];
        for (int i = 3; i < 6; i++) cross_motion(d->cdof_dot[da + i], cvel, d->cdof[da + i]);
        for (int i = 3; i < 6; i++) for (int k = 0; k < 6; k++) cvel[k] += d->cdof[da + i][k]*d->qvel[da + i];
      } else {
        cross_motion(d->cdof_dot[da], cvel, d->cdof[da]);
        for (int k = 0; k < 6; k++) cvel[k] += d->cdof[da][k]*d->qvel[da];
      }
    }
    memcpy(d->cvel[b], cvel, sizeof(cvel));
  }
}

void dmo_fwd_velocity(const dmb_model_t* m, dmo_data_t* d) {
  int nv = m->nv;
  com_vel(m, d);
  /* mj_passive: joint damping only */
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i]*d->qvel[i];
  /* mj_rne(flg_acc = 0) */
  double cacc[DMB_MAX_BODY][6], cfrc[DMB_MAX_BODY][6];
  memset(cacc, 0, sizeof(cacc)); memset(cfrc, 0, sizeof(cfrc));
  for (int k = 0; k < 3; k++) cacc[0][3 + k] = -m->gravity[k];
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b];
    memcpy(cacc[b], cacc[p], sizeof(cacc[b]));
    for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++)
      for (int k = 0; k < 6; k++) cacc[b][k] += d->cdof_dot[i][k]*d->qvel[i];
    double t1[6], t2[6];
    mul_inert_vec(cfrc[b], d->cinert[b], cacc[b]);
    mul_inert_vec(t1, d->cinert[b], d->cvel[b]);
    cross_force(t2, d->cvel[b], t1);
    for (int k = 0; k < 6; k++) cfrc[b][k] += t2[k];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    for (int k = 0; k < 6; k++) cfrc[p][k] += cfrc[b][k];
  }
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += d->cdof[i][k]*cfrc[m->dof_bodyid[i]][k];
    d->qfrc_bias[i] = s;
  }
  /* mj_referenceConstraint: efc_vel = J qvel, aref = -b*vel - k*imp*(pos - margin) */
  for (int r = 0; r < d->nefc; r++) {
    double v = 0;
    for (int i = 0; i < nv; i++) v += d->efc_J[r][i]*d->qvel[i];
    d->efc_vel[r] = v;
    d->efc_aref[r] = -d->efc_KBIP[r][1]*v - d->efc_KBIP[r][0]*d->efc_KBIP[r][2]*(d->efc_pos[r] - d->efc_margin[r]);
  }
}

/* ------------------------------------------------------------------ mj_fwdActuation */
void dmo_fwd_actuation(const dmb_model_t* m, dmo_data_t* d) {
  memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
  for (int u = 0; u < m->nu; u++) {
    double c = d->ctrl[u];
    if (c < m->act_ctrlrange[u][0]) c = m->act_ctrlrange[u][0];
    if (c > m->act_ctrlrange[u][1]) c = m->act_ctrlrange[u][1];
    d->qfrc_actuator[m->act_dofadr[u]] += m->act_gear[u]*c;
  }
}

/* ------------------------------------------------------------------ mj_fwdAcceleration */
void dmo_fwd_acceleration(const dmb_model_t* m, dmo_data_t* d) {
  for (int i = 0; i < m->nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  dmo_solve_M(m, d, d->qacc_smooth);
}

/* ------------------------------------------------------------------ mj_fwdConstraint (PGS) */
void dmo_fwd_constraint(const dmb_model_t* m, dmo_data_t* d) {
  int nv = m->nv, nefc = d->nefc;
  d->solver_iter = 0;
  if (nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(d->qacc));
    memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(d->qacc));
    memset(d->qfrc_constraint, 0, sizeof(d->qfrc_constraint));
    return;
  }
  /* efc_b = J qacc_smooth - aref */
  for (int r = 0; r < nefc; r++) {
    double s = 0;
    for (int i = 0; i < nv; i++) s += d->efc_J[r][i]*d->qacc_smooth[i];
    d->efc_b[r] = s - d->efc_aref[r];
  }
  /* warmstart: forces from qacc_warmstart; keep them only if the dual cost is negative */
  {
    double cost = 0;
    for (int r = 0; r < nefc; r++) {
      double jar = -d->efc_aref[r];
      for (int i = 0; i < nv; i++) jar += d->efc_J[r][i]*d->qacc_warmstart[i];
      d->efc_force[r] = jar < 0 ? -d->efc_D[r]*jar : 0.0; /* all rows are unilateral */
    }
    for (int r = 0; r < nefc; r++) {
      double s = 0;
      for (int c = 0; c < nefc; c++) s += d->efc_AR[r][c]*d->efc_force[c];
      cost += d->efc_force[r]*(0.5*s + d->efc_b[r]);
    }
    if (cost > 0) memset(d->efc_force, 0, sizeof(double)*nefc);
  }
  /* mj_solPGS: all rows are scalar with force >= 0 */
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  int iter = 0;
  while (iter < m->iterations) {
    double improvement = 0;
    for (int r = 0; r < nefc; r++) {
      double res = d->efc_b[r];
      for (int c = 0; c < nefc; c++) res += d->efc_AR[r][c]*d->efc_force[c];
      double old = d->efc_force[r];
      double f = old - res / d->efc_AR[r][r];
      if (f < 0) f = 0;
      d->efc_force[r] = f;
      double delta = f - old;
      improvement -= 0.5*delta*delta*d->efc_AR[r][r] + delta*res;
    }
    iter++;
    if (improvement*scale < m->tolerance) break;
  }
  d->solver_iter = iter;
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int r = 0; r < nefc; r++) s += d->efc_J[r][i]*d->efc_force[r];
    d->qfrc_constraint[i] = s;
    d->qacc[i] = s;
  }
  dmo_solve_M(m, d, d->qacc);
  for (int i = 0; i < nv; i++) d->qacc[i] += d->qacc_smooth[i];
  memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc));
}

void dmo_forward(const dmb_model_t* m, dmo_data_t* d) {
  dmo_fwd_position(m, d);
  dmo_fwd_velocity(m, d);
  dmo_fwd_actuation(m, d);
  dmo_fwd_acceleration(m, d);
  dmo_fwd_constraint(m, d);
}

/* ------------------------------------------------------------------ integration */
/* mj_integratePos: free joint = translation + body-frame quaternion update; hinge = scalar */
static void integrate_pos(const dmb_model_t* m, double* qpos, const double* qvel, double h) {
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == DMB_JNT_FREE) {
      for (int k = 0; k < 3; k++) qpos[qa + k] += h*qvel[da + k];
      double ax[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]};
      double ang = h*normalize3(ax), qr[4], qn[4];
      axis_angle2quat(qr, ax, ang);
      normalize4(qpos + qa + 3);
      mul_quat(qn, qpos + qa + 3, qr);
      normalize4(qn);
      memcpy(qpos + qa + 3, qn, sizeof(qn));
    } else {
      qpos[qa] += h*qvel[da];
    }
  }
}

static int state_bad(const dmb_model_t* m, const dmo_data_t* d) {
  for (int i = 0; i < m->nq; i++) if (!(fabs(d->qpos[i]) < 1e10)) return 1;
  for (int i = 0; i < m->nv; i++) if (!(fabs(d->qvel[i]) < 1e10)) return 1;
  return 0;
}

/* mj_step with integrator = RK4 (mj_RungeKutta, N = 4) */
void dmo_step(const dmb_model_t* m, dmo_data_t* d) {
  static const double A[9] = {0.5, 0, 0, 0, 0.5, 0, 0, 0, 1};
  static const double Bw[4] = {1.0/6, 1.0/3, 1.0/3, 1.0/6};
  int nq = m->nq, nv = m->nv;
  double h = m->timestep;
  double X[4][DMB_MAX_Q + DMB_MAX_DOF], F[4][DMB_MAX_DOF], dX[2*DMB_MAX_DOF];
  d->flags = 0;
  if (state_bad(m, d)) { d->flags |= 4; return; }
  dmo_forward(m, d);
  memcpy(X[0], d->qpos, sizeof(double)*nq);
  memcpy(X[0] + nq, d->qvel, sizeof(double)*nv);
  memcpy(F[0], d->qacc, sizeof(double)*nv);
  for (int i = 1; i < 4; i++) {
    memset(dX, 0, sizeof(dX));
    for (int j = 0; j < i; j++) {
      double a = A[(i - 1)*3 + j];
      for (int k = 0; k < nv; k++) { dX[k] += a*X[j][nq + k]; dX[nv + k] += a*F[j][k]; }
    }
    memcpy(X[i], X[0], sizeof(double)*(nq + nv));
    integrate_pos(m, X[i], dX, h);
    for (int k = 0; k < nv; k++) X[i][nq + k] += h*dX[nv + k];
    memcpy(d->qpos, X[i], sizeof(double)*nq);
    memcpy(d->qvel, X[i] + nq, sizeof(double)*nv);
    dmo_forward(m, d);
    memcpy(F[i], d->qacc, sizeof(double)*nv);
  }
  memset(dX, 0, sizeof(dX));
  for (int j = 0; j < 4; j++)
    for (int k = 0; k < nv; k++) { dX[k] += Bw[j]*X[j][nq + k]; dX[nv + k] += Bw[j]*F[j][k]; }
  memcpy(d->qpos, X[0], sizeof(double)*nq);
  for (int k = 0; k < nv; k++) d->qvel[k] = X[0][nq + k] + h*dX[nv + k];
  integrate_pos(m, d->qpos, dX, h);
  if (state_bad(m, d)) d->flags |= 4;
}

/* ====================================================================================
 * Env logic: restates /root/reference/src/dp_env_v3.py DPEnv (step 106-132, _get_obs 62-65,
 * is_done 134-139, calc_config_reward 89-104, reset_model 148-156, reset_model_init 158-164)
 * plus the PD controller intent of mujoco_interface.py:97-107 and the 5-term DeepMimic
 * reward quoted in code.md:979-1146 (adapted to the hinge model).
 * ==================================================================================== */

/* Philox4x32-10; same integer arithmetic as csrc/dmb_math.cuh */
void dmo_philox(uint64_t seed, uint32_t env_id, uint32_t reset_count, uint32_t block, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
  uint32_t c0 = env_id, c1 = reset_count, c2 = block, c3 = 0;
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ key0, n1 = lo1, n2 = hi0 ^ c3 ^ key1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    key0 += W0; key1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline float u01f(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

void dmo_env_obs(const dmb_model_t* m, const dmo_env_t* e, double* obs) {
  int np = m->nq - 7;
  for (int i = 0; i < np; i++) obs[i] = e->d.qpos[7 + i];
  for (int i = 0; i < m->nv - 6; i++) obs[np + i] = e->d.qvel[6 + i];
}

void dmo_env_set_state(const dmb_model_t* m, dmo_env_t* e, const double* qpos, const double* qvel) {
  memcpy(e->d.qpos, qpos, sizeof(double)*m->nq);
  memcpy(e->d.qvel, qvel, sizeof(double)*m->nv);
  memset(e->d.qacc_warmstart, 0, sizeof(e->d.qacc_warmstart));
}

void dmo_env_reset(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e, int mode) {
  if (mode < 0) mode = cfg->reset_mode;
  int len = mc->clip_len[e->clip], start = mc->clip_start[e->clip];
  uint32_t r[4];
  dmo_philox(e->seed, e->env_id, e->reset_count, 0, r);
  int idx = (int)(u01f(r[0]) * (float)len);
  if (idx >= len) idx = len - 1;
  if (mode == 0) {
    /* tables are consumed in fp32 by the CUDA path; round the same way */
    for (int i = 0; i < m->nq; i++) e->d.qpos[i] = (double)(float)mc->data_config[(size_t)(start + idx)*m->nq + i];
    for (int i = 0; i < m->nv; i++) e->d.qvel[i] = (double)(float)mc->data_vel[(size_t)(start + idx)*m->nv + i];
  } else {
    for (int i = 0; i < m->nq + m->nv; i++) {
      dmo_philox(e->seed, e->env_id, e->reset_count, 1u + (uint32_t)(i >> 2), r);
      float nz = (float)cfg->reset_noise * (2.f*u01f(r[i & 3]) - 1.f);
      if (i < m->nq) e->d.qpos[i] = (double)((float)m->qpos0[i] + nz);
      else e->d.qvel[i - m->nq] = (double)nz;
    }
  }
  memset(e->d.qacc_warmstart, 0, sizeof(e->d.qacc_warmstart));
  e->idx_init = idx; e->idx_curr = idx; e->reset_count++;
  e->ep_len = 0; e->ep_ret = 0;
}

void dmo_env_init(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e,
                  uint64_t seed, uint32_t env_id, int32_t clip) {
  memset(e, 0, sizeof(*e));
  e->seed = seed; e->env_id = env_id; e->clip = clip;
  memcpy(e->d.qpos, m->qpos0, sizeof(double)*m->nq);
}

/* quaternion of a hinge triple: R = Rx(a) Ry(b) Rz(c)  (dp_env_v3.xml hinge stacking) */
static void quat_from_xyz(double* q, const double* a) {
  double qx[4] = {cos(0.5*a[0]), sin(0.5*a[0]), 0, 0}, qy[4] = {cos(0.5*a[1]), 0, sin(0.5*a[1]), 0};
  double qz[4] = {cos(0.5*a[2]), 0, 0, sin(0.5*a[2])}, t[4];
  mul_quat(t, qx, qy); mul_quat(q, t, qz);
}
/* geodesic angle in [0, pi] between two unit quaternions (cMathUtil::QuatDiffTheta), computed
 * from the difference quaternion with atan2 so that small angles keep their accuracy in fp32 */
static double quat_diff_theta(const double* a, const double* b) {
  double ac[4] = {a[0], -a[1], -a[2], -a[3]}, qd[4];
  mul_quat(qd, ac, b);
  return 2*atan2(sqrt(qd[1]*qd[1] + qd[2]*qd[2] + qd[3]*qd[3]), fabs(qd[0]));
}
/* heading-frame end-effector points + CoM velocity of the pose currently in d (after
 * dmo_kinematics + com_pos + com_vel) */
static void pose_features(const dmb_model_t* m, const dmo_data_t* d, double* ee /* nee*3 */, double* vcom) {
  const double* R = d->xmat[1]; /* root body */
  double heading = atan2(R[3], R[0]); /* world image of the root x axis */
  double ch = cos(heading), sh = sin(heading);
  for (int k = 0; k < m->nee; k++) {
    int b = m->ee_body[k];
    double w[3], rel[3];
    mat_vec(w, d->xmat[b], m->ee_pos[k]);
    for (int i = 0; i < 3; i++) w[i] += d->xpos[b][i];
    rel[0] = w[0] - d->xpos[1][0]; rel[1] = w[1] - d->xpos[1][1]; rel[2] = w[2]; /* height above ground */
    ee[3*k] = ch*rel[0] + sh*rel[1];
    ee[3*k + 1] = -sh*rel[0] + ch*rel[1];
    ee[3*k + 2] = rel[2];
  }
  double mass = 0, p[3] = {0, 0, 0};
  for (int b = 1; b < m->nbody; b++) {
    double r[3] = {d->xipos[b][0] - d->com[0], d->xipos[b][1] - d->com[1], d->xipos[b][2] - d->com[2]}, t[3];
    cross3(t, d->cvel[b], r);
    for (int i = 0; i < 3; i++) p[i] += m->body_mass[b]*(d->cvel[b][3 + i] + t[i]);
    mass += m->body_mass[b];
  }
  for (int i = 0; i < 3; i++) vcom[i] = p[i]/mass;
}

void dmo_ref_aux(const dmb_model_t* m, const double* qpos, const double* qvel, double* aux) {
  static __thread dmo_data_t d;
  memcpy(d.qpos, qpos, sizeof(double)*m->nq);
  memcpy(d.qvel, qvel, sizeof(double)*m->nv);
  dmo_kinematics(m, &d); com_pos(m, &d); com_vel(m, &d);
  memset(aux, 0, sizeof(double)*DMB_REF_AUX);
  pose_features(m, &d, aux, aux + 12);
  for (int i = 0; i < 4; i++) aux[15 + i] = qpos[3 + i];
}

/* ------------------------------------------------------------------------------------
 * Time-based mocap phase with interpolation (spec pinned against the
 * reference's own routines in tests/golden/make_interp_golden.py).
 * ------------------------------------------------------------------------------------ */
/* transformations.quaternion_slerp(q0, q1, fraction, spin=0, shortestpath=True)
 * (/root/reference/src/transformations.py:1270-1308) */
static void quat_slerp(double* out, const double* qa, const double* qb, double f) {
  const double EPS = 2.220446049250313e-16*4.0;
  double q0[4], q1[4];
  memcpy(q0, qa, sizeof(q0)); memcpy(q1, qb, sizeof(q1));
  normalize4(q0); normalize4(q1);
  if (f == 0.0) { memcpy(out, q0, sizeof(q0)); return; }
  if (f == 1.0) { memcpy(out, q1, sizeof(q1)); return; }
  double d = q0[0]*q1[0] + q0[1]*q1[1] + q0[2]*q1[2] + q0[3]*q1[3];
  if (fabs(fabs(d) - 1.0) < EPS) { memcpy(out, q0, sizeof(q0)); return; }
  if (d < 0.0) { d = -d; for (int i = 0; i < 4; i++) q1[i] = -q1[i]; }
  double angle = acos(d);
  if (fabs(angle) < EPS) { memcpy(out, q0, sizeof(q0)); return; }
  double isin = 1.0/sin(angle);
  double w0 = sin((1.0 - f)*angle)*isin, w1 = sin(f*angle)*isin;
  for (int i = 0; i < 4; i++) out[i] = q0[i]*w0 + q1[i]*w1;
}
/* transformations.euler_from_quaternion([x,y,z,w], 'rxyz') (transformations.py:1031-1097,1174-1193)
 * for a (w,x,y,z) quaternion: R = Rx(a) Ry(b) Rz(c) */
static void euler_rxyz_from_quat(double* e, const double* q) {
  const double EPS = 2.220446049250313e-16*4.0;
  double n = q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3];
  if (n < EPS) { e[0] = e[1] = e[2] = 0; return; }
  double s = sqrt(2.0/n);
  double w = q[0]*s, x = q[1]*s, y = q[2]*s, z = q[3]*s;
  double R00 = 1.0 - y*y - z*z, R01 = x*y - z*w, R02 = x*z + y*w;
  double R10 = x*y + z*w, R11 = 1.0 - x*x - z*z, R12 = y*z - x*w, R22 = 1.0 - x*x - y*y;
  double cy = sqrt(R22*R22 + R12*R12);
  if (cy > EPS) { e[0] = atan2(-R12, R22); e[1] = atan2(R02, cy); e[2] = atan2(-R01, R00); }
  else { e[0] = 0; e[1] = atan2(R02, cy); e[2] = atan2(R10, R11); }
}
/* frame coordinate u -> (cycle, k, alpha, phase) for an F-frame clip: one cycle = F-1 intervals */
static void phase_split(int F, double u, int* cycle, int* k, double* alpha, double* phase) {
  if (F < 2) { *cycle = 0; *k = 0; *alpha = 0; *phase = 0; return; }
  double c = floor(u/(double)(F - 1));
  double uu = u - c*(double)(F - 1);
  int kk = (int)uu;
  if (kk > F - 2) kk = F - 2;
  if (kk < 0) kk = 0;
  *cycle = (int)c; *k = kk; *alpha = uu - (double)kk; *phase = uu/(double)(F - 1);
}

/* interpolated reference pose at frame coordinate u = t / clip_dt (mocap.py sample_tables); table
 * entries are rounded to fp32 first, as the CUDA path stores them */
void dmo_mocap_sample(const dmb_model_t* m, const dmb_mocap_t* mc, int clip, double u, double* qpos, double* qvel,
                      double* phase) {
  int F = mc->clip_len[clip], start = mc->clip_start[clip], cycle, k;
  double a, ph;
  phase_split(F, u, &cycle, &k, &a, &ph);
  int k1 = F < 2 ? k : k + 1;
  const double* c0 = mc->data_config + (size_t)(start + k)*m->nq;
  const double* c1 = mc->data_config + (size_t)(start + k1)*m->nq;
  const double* cl = mc->data_config + (size_t)(start + F - 1)*m->nq;
  const double* v0 = mc->data_vel + (size_t)(start + k)*m->nv;
  const double* v1 = mc->data_vel + (size_t)(start + k1)*m->nv;
  double r0[DMB_MAX_Q], r1[DMB_MAX_Q];
  for (int i = 0; i < m->nq; i++) { r0[i] = (double)(float)c0[i]; r1[i] = (double)(float)c1[i]; }
  for (int i = 0; i < m->nq; i++) qpos[i] = r0[i] + a*(r1[i] - r0[i]);
  qpos[0] += (double)cycle*(double)(float)cl[0];
  qpos[1] += (double)cycle*(double)(float)cl[1];
  quat_slerp(qpos + 3, r0 + 3, r1 + 3, a);
  for (int b = 2; b < m->nbody; b++) {
    if (m->body_dofnum[b] != 3) continue;
    int qa = m->body_dofadr[b] + 1;
    double q0[4], q1[4], qs[4];
    quat_from_xyz(q0, r0 + qa); quat_from_xyz(q1, r1 + qa);
    quat_slerp(qs, q0, q1, a);
    euler_rxyz_from_quat(qpos + qa, qs);
  }
  for (int i = 0; i < m->nv; i++) {
    double x0 = (double)(float)v0[i], x1 = (double)(float)v1[i];
    qvel[i] = x0 + a*(x1 - x0);
  }
  if (phase) *phase = ph;
}

/* frame coordinate of env e `extra` env steps after its current step count (phase_mode 1) */
static double env_frame_coord(const dmb_model_t* m, const dmb_mocap_t* mc, const dmo_env_t* e, int extra) {
  double rate = m->timestep/mc->clip_dt[e->clip];
  return (double)e->idx_init + (double)(e->ep_len + extra)*rate;
}

/* reference pose for the reward: phase_mode 0 = table row `frame` of the clip (tables rounded to fp32
 * as the CUDA path stores them); phase_mode 1 = interpolated at frame coordinate u, with the
 * end-effector / CoM-velocity features recomputed by forward kinematics of the interpolated pose */
static void get_reference(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, const dmo_env_t* e,
                          int frame, double u, double* rq, double* rv, double* aux) {
  if (cfg->phase_mode == 1) {
    dmo_mocap_sample(m, mc, e->clip, u, rq, rv, NULL);
    if (aux) dmo_ref_aux(m, rq, rv, aux);
    return;
  }
  size_t f = (size_t)(mc->clip_start[e->clip] + frame);
  for (int i = 0; i < m->nq; i++) rq[i] = (double)(float)mc->data_config[f*m->nq + i];
  for (int i = 0; i < m->nv; i++) rv[i] = (double)(float)mc->data_vel[f*m->nv + i];
  if (aux) for (int i = 0; i < DMB_REF_AUX; i++) aux[i] = (double)(float)mc->ref_aux[f*DMB_REF_AUX + i];
  if (cfg->reward_mode == 4) {
    /* the integer frame counter wraps (dp_env_v3.py:101-102); the 5-term reward scores the absolute root
     * position, so the reference root keeps moving: every completed pass over the clip adds the last frame's
     * root xy (the root-offset accumulation of MocapDM.play, mocap_v2.py:168-182) */
    const int len = mc->clip_len[e->clip];
    const int cycle = (e->idx_init + e->ep_len)/len;
    const size_t fl = (size_t)(mc->clip_start[e->clip] + len - 1);
    rq[0] += (double)cycle*(double)(float)mc->data_config[fl*m->nq + 0];
    rq[1] += (double)cycle*(double)(float)mc->data_config[fl*m->nq + 1];
  }
}

static double reward_imitate(const dmb_model_t* m, const dmb_config_t* cfg, dmo_env_t* e, const double* rq,
                             const double* rv, const double* aux) {
  dmo_data_t* d = &e->d;
  /* fresh kinematics at the post-step state */
  dmo_kinematics(m, d); com_pos(m, d); com_vel(m, d);
  double ee[3*DMB_MAX_EE], vcom[3];
  pose_features(m, d, ee, vcom);
  double pose_err = 0, vel_err = 0;
  { /* root */
    double q0[4] = {d->qpos[3], d->qpos[4], d->qpos[5], d->qpos[6]}, q1[4] = {rq[3], rq[4], rq[5], rq[6]};
    normalize4(q0); normalize4(q1);
    double th = quat_diff_theta(q0, q1);
    pose_err += m->dof_weight[3]*th*th;
  }
  for (int b = 2; b < m->nbody; b++) {
    int da = m->body_dofadr[b], nd = m->body_dofnum[b];
    if (nd == 3) {
      double q0[4], q1[4];
      quat_from_xyz(q0, d->qpos + da + 1); quat_from_xyz(q1, rq + da + 1);
      double th = quat_diff_theta(q0, q1);
      pose_err += m->dof_weight[da]*th*th;
    } else if (nd == 1) {
      double dq = d->qpos[da + 1] - rq[da + 1];
      pose_err += m->dof_weight[da]*dq*dq;
    }
  }
  for (int i = 3; i < m->nv; i++) { double dv = d->qvel[i] - rv[i]; vel_err += m->dof_weight[i]*dv*dv; }
  double ee_err = 0;
  for (int k = 0; k < 3*m->nee; k++) { double t = ee[k] - aux[k]; ee_err += t*t; }
  if (m->nee > 0) ee_err /= m->nee;
  double root_err = 0;
  {
    double q0[4] = {d->qpos[3], d->qpos[4], d->qpos[5], d->qpos[6]}, q1[4] = {rq[3], rq[4], rq[5], rq[6]};
    normalize4(q0); normalize4(q1);
    double th = quat_diff_theta(q0, q1), p = 0, v = 0, w = 0;
    for (int i = 0; i < 3; i++) {
      p += (d->qpos[i] - rq[i])*(d->qpos[i] - rq[i]);
      v += (d->qvel[i] - rv[i])*(d->qvel[i] - rv[i]);
      w += (d->qvel[3 + i] - rv[3 + i])*(d->qvel[3 + i] - rv[3 + i]);
    }
    root_err = p + 0.1*th*th + 0.01*v + 0.001*w;
  }
  double com_err = 0;
  for (int i = 0; i < 3; i++) com_err += (aux[12 + i] - vcom[i])*(aux[12 + i] - vcom[i]);
  com_err *= 0.1;
  e->reward_terms[0] = exp(-cfg->s_err*cfg->s_pose*pose_err);
  e->reward_terms[1] = exp(-cfg->s_err*cfg->s_vel*vel_err);
  e->reward_terms[2] = exp(-cfg->s_err*cfg->s_end_eff*ee_err);
  e->reward_terms[3] = exp(-cfg->s_err*cfg->s_root*root_err);
  e->reward_terms[4] = exp(-cfg->s_err*cfg->s_com*com_err);
  return cfg->w_pose*e->reward_terms[0] + cfg->w_vel*e->reward_terms[1] + cfg->w_end_eff*e->reward_terms[2] +
         cfg->w_root*e->reward_terms[3] + cfg->w_com*e->reward_terms[4];
}

/* DeepMimic state (obs_mode 1): cCtController::BuildStatePose / BuildStateVel quoted in
 * /root/reference/code.md:287-504 and the MuJoCo prototype /root/reference/src/mujoco/mujoco_env.py:91-124
 * (record_state), on the hinge model: [phase, root height, npart x (pos 3, quat 4), npart x (lin vel 3,
 * ang vel 3)], positions relative to the root and everything rotated into the root heading frame
 * (BuildOriginTrans); quaternions (w,x,y,z) with w >= 0; z-up component order.  Recomputes the
 * kinematics of the CURRENT state (post-reset for auto-reset envs). */
int dmo_env_obs_dm(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e, double* obs) {
  dmo_data_t* d = &e->d;
  dmo_kinematics(m, d); com_pos(m, d); com_vel(m, d);
  const int np = m->npart, len = mc->clip_len[e->clip];
  double phase;
  if (cfg->phase_mode == 1) { int cyc, k; double al; phase_split(len, env_frame_coord(m, mc, e, 0), &cyc, &k, &al, &phase); }
  else phase = (double)e->idx_curr/(double)len;
  const double* R = d->xmat[1];
  double heading = atan2(R[3], R[0]), ch = cos(heading), sh = sin(heading);
  double qhi[4] = {cos(0.5*heading), 0, 0, -sin(0.5*heading)}; /* inverse heading rotation */
  obs[0] = phase;
  obs[1] = d->xpos[1][2];
  for (int p = 0; p < np; p++) {
    int g = m->part_geom[p], b = m->geom_bodyid[g];
    double rel[3] = {d->geom_xpos[g][0] - d->xpos[1][0], d->geom_xpos[g][1] - d->xpos[1][1], d->geom_xpos[g][2] - d->xpos[1][2]};
    double* o = obs + 2 + 7*p;
    o[0] = ch*rel[0] + sh*rel[1]; o[1] = -sh*rel[0] + ch*rel[1]; o[2] = rel[2];
    double q[4];
    mul_quat(q, qhi, d->xquat[b]);
    double sg = q[0] < 0 ? -1.0 : 1.0;
    for (int i = 0; i < 4; i++) o[3 + i] = sg*q[i];
    /* velocity of the part centre: cvel = [omega; v at the c-frame origin (subtree CoM)] */
    double r[3] = {d->geom_xpos[g][0] - d->com[0], d->geom_xpos[g][1] - d->com[1], d->geom_xpos[g][2] - d->com[2]}, t[3];
    cross3(t, d->cvel[b], r);
    double v[3] = {d->cvel[b][3] + t[0], d->cvel[b][4] + t[1], d->cvel[b][5] + t[2]};
    double* ov = obs + 2 + 7*np + 6*p;
    ov[0] = ch*v[0] + sh*v[1]; ov[1] = -sh*v[0] + ch*v[1]; ov[2] = v[2];
    ov[3] = ch*d->cvel[b][0] + sh*d->cvel[b][1]; ov[4] = -sh*d->cvel[b][0] + ch*d->cvel[b][1]; ov[5] = d->cvel[b][2];
  }
  return 2 + 13*np;
}

int dmo_env_step(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e,
                 const double* action, double* obs, double* reward) {
  dmo_data_t* d = &e->d;
  /* action -> ctrl */
  for (int u = 0; u < m->nu; u++) {
    double a = action[u];
    int dof = m->act_dofadr[u];
    if (cfg->ctrl_mode != 0) {
      double perr = a - d->qpos[dof + 1], tau;
      if (cfg->ctrl_mode == 1) tau = m->act_kp[u]*perr + m->act_kd[u]*(perr/m->timestep - d->qvel[dof]);
      else tau = m->act_kp[u]*perr - m->act_kd[u]*d->qvel[dof];
      a = tau / m->act_gear[u];
    }
    if (a != a) a = 0;
    d->ctrl[u] = a;
  }
  dmo_step(m, d);
  int bad = (d->flags & 4) != 0;
  double zc = d->com[2]; /* stale CoM of the last RK4 stage, as mjData.xipos is after mj_step */
  e->zcom_last = zc;
  double rew = 1.0;
  const int len = mc->clip_len[e->clip];
  /* phase_mode 1: the reference is sampled at the post-step time for every reward mode */
  const double u_ref = cfg->phase_mode == 1 ? env_frame_coord(m, mc, e, 1) : 0.0;
  double rq[DMB_MAX_Q], rv[DMB_MAX_DOF], aux[DMB_REF_AUX];
  if (cfg->reward_mode == 1) {
    get_reference(m, cfg, mc, e, e->idx_curr, u_ref, rq, rv, NULL);
    double err = 0;
    for (int j = 0; j < m->nq - 7; j++) err += fabs(d->qpos[7 + j] - rq[7 + j]);
    rew = exp(-err);
    e->idx_curr = (e->idx_curr + 1) % len;
  } else if (cfg->reward_mode == 4) {
    get_reference(m, cfg, mc, e, e->idx_curr, u_ref, rq, rv, aux);
    rew = reward_imitate(m, cfg, e, rq, rv, aux);
    e->idx_curr = (e->idx_curr + 1) % len;
  } else if (cfg->reward_mode == 2 || cfg->reward_mode == 3) {
    /* v2 / v1 rewards: the frame counter advances before the reward is evaluated
     * (dp_env_v2.py:174-183, dp_env_v1.py:143-152); control cost on the raw action */
    e->idx_curr = (e->idx_curr + 1) % len;
    get_reference(m, cfg, mc, e, e->idx_curr, u_ref, rq, rv, NULL);
    double acs = 0;
    for (int u = 0; u < m->nu; u++) acs += action[u]*action[u];
    if (cfg->reward_mode == 2) { /* exp(-scale_err*scale_pose*|qpos[3:] - ref[3:]|_1) */
      double err = 0;
      for (int i = 3; i < m->nq; i++) err += fabs(d->qpos[i] - rq[i]);
      rew = exp(-cfg->s_err*cfg->s_pose*err) - 0.1*acs;
    } else { /* v1: weighted joint angles (root included), L1 velocity, L1 root position */
      double pose = 0, vel = 0, root = 0;
      double q0[4] = {d->qpos[3], d->qpos[4], d->qpos[5], d->qpos[6]}, q1[4] = {rq[3], rq[4], rq[5], rq[6]};
      normalize4(q0); normalize4(q1);
      pose += m->dof_weight[3]*cfg->joint_weight_sum*quat_diff_theta(q0, q1);
      for (int b = 2; b < m->nbody; b++) {
        int da = m->body_dofadr[b], nd = m->body_dofnum[b];
        double w = m->dof_weight[da]*cfg->joint_weight_sum;
        if (nd == 3) { double a[4], c[4]; quat_from_xyz(a, d->qpos + da + 1); quat_from_xyz(c, rq + da + 1); pose += w*quat_diff_theta(a, c); }
        else if (nd == 1) pose += w*fabs(d->qpos[da + 1] - rq[da + 1]);
      }
      for (int i = 3; i < m->nv; i++) vel += fabs(d->qvel[i] - rv[i]);
      for (int i = 0; i < 3; i++) root += fabs(d->qpos[i] - rq[i]);
      rew = cfg->w_pose*exp(-cfg->s_err*cfg->s_pose*pose) + cfg->w_vel*exp(-cfg->s_err*cfg->s_vel*vel) +
            cfg->w_root*exp(-cfg->s_err*cfg->s_root*root) - 0.1*acs;
    }
  }
  if (cfg->phase_mode == 1) { /* idx_curr reports the frame interval of the reference time */
    int cyc, k; double al, ph;
    phase_split(len, u_ref, &cyc, &k, &al, &ph);
    e->idx_curr = k;
  }
  if (bad) rew = 0;
  int done = bad || zc < cfg->z_min || zc > cfg->z_max;
  if (cfg->term_mode == 1) { /* fall contact: a listed body touches the floor (contacts of the last RK4 stage) */
    for (int ci = 0; ci < d->ncon; ci++) {
      const dmo_contact_t* c = &d->contact[ci];
      if (m->geom_type[c->geom1] == DMB_GEOM_PLANE && ((cfg->fall_body_mask >> m->geom_bodyid[c->geom2]) & 1u)) done = 1;
    }
  }
  e->ep_len++; e->ep_ret += rew;
  *reward = rew;
  if (done && cfg->auto_reset) dmo_env_reset(m, cfg, mc, e, cfg->reset_mode);
  else if (bad) { memcpy(d->qpos, m->qpos0, sizeof(double)*m->nq); memset(d->qvel, 0, sizeof(d->qvel)); memset(d->qacc_warmstart, 0, sizeof(d->qacc_warmstart)); }
  if (obs) { if (cfg->obs_mode == 1) dmo_env_obs_dm(m, cfg, mc, e, obs); else dmo_env_obs(m, e, obs); }
  return done;
}

/* CPU baseline loop: random actions a ~ U(-0.5, 0.5)^nu (action_space.sample()), reset on done */
long dmo_rollout(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e,
                 long nsteps, uint64_t action_seed) {
  double action[DMB_MAX_U], obs[2 + 13*DMB_MAX_PART + 2*DMB_MAX_DOF], rew;
  uint32_t r[4];
  dmb_config_t c = *cfg;
  c.auto_reset = 1;
  for (long t = 0; t < nsteps; t++) {
    for (int u = 0; u < m->nu; u += 4) {
      dmo_philox(action_seed, e->env_id, (uint32_t)t, (uint32_t)(u >> 2), r);
      for (int k = 0; k < 4 && u + k < m->nu; k++) action[u + k] = (double)(u01f(r[k]) - 0.5f);
    }
    dmo_env_step(m, &c, mc, e, action, obs, &rew);
  }
  return nsteps;
}

/* One env step for a batch of independent envs, from explicit per-env inputs (the tests feed it the fp32 state
 * of the CUDA path and compare the outputs): env i has global id first_env_id + i (Philox stream of its resets).
 * Arrays are row-major [n][.] doubles / int32; qpos/qvel/warm and the bookkeeping arrays are updated in place.
 * last_ret / last_len = return and length of the episode including this step (the Monitor record,
 * bench/monitor.py:58-76, meaningful where done is set).  Test infrastructure only. */
void dmo_batch_step(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, int n, uint64_t seed,
                    uint32_t first_env_id, double* qpos, double* qvel, double* warm, const int32_t* clip,
                    int32_t* idx_init, int32_t* idx_curr, int32_t* ep_len, double* ep_ret, int32_t* reset_count,
                    const double* action, double* obs, int obs_stride, double* reward, int32_t* done,
                    double* last_ret, int32_t* last_len, int32_t* flags, int32_t* nefc_last, double* zcom) {
  dmo_env_t* e = (dmo_env_t*)malloc(sizeof(dmo_env_t));
  double ob[2 + 13*DMB_MAX_PART + 2*DMB_MAX_DOF];
  for (int i = 0; i < n; i++) {
    dmo_env_init(m, cfg, mc, e, seed, first_env_id + (uint32_t)i, clip[i]);
    memcpy(e->d.qpos, qpos + (size_t)i*m->nq, sizeof(double)*m->nq);
    memcpy(e->d.qvel, qvel + (size_t)i*m->nv, sizeof(double)*m->nv);
    memcpy(e->d.qacc_warmstart, warm + (size_t)i*m->nv, sizeof(double)*m->nv);
    e->idx_init = idx_init[i]; e->idx_curr = idx_curr[i]; e->ep_len = ep_len[i]; e->ep_ret = ep_ret[i];
    e->reset_count = (uint32_t)reset_count[i];
    double rew = 0;
    const double ret0 = e->ep_ret;
    const int len0 = e->ep_len;
    e->d.com[2] = 0;
    done[i] = dmo_env_step(m, cfg, mc, e, action + (size_t)i*m->nu, ob, &rew);
    reward[i] = rew;
    last_ret[i] = ret0 + rew; last_len[i] = len0 + 1;
    if (flags) flags[i] = e->d.flags;
    if (nefc_last) nefc_last[i] = e->d.nefc;
    if (zcom) zcom[i] = e->zcom_last;
    memcpy(obs + (size_t)i*obs_stride, ob, sizeof(double)*obs_stride);
    memcpy(qpos + (size_t)i*m->nq, e->d.qpos, sizeof(double)*m->nq);
    memcpy(qvel + (size_t)i*m->nv, e->d.qvel, sizeof(double)*m->nv);
    memcpy(warm + (size_t)i*m->nv, e->d.qacc_warmstart, sizeof(double)*m->nv);
    idx_init[i] = e->idx_init; idx_curr[i] = e->idx_curr; ep_len[i] = e->ep_len; ep_ret[i] = e->ep_ret;
    reset_count[i] = (int32_t)e->reset_count;
  }
  free(e);
}
