// dmb_device.cuh -- one-warp-per-env forward dynamics (MuJoCo pipeline restated for sm_100a).
//
// Every function below is executed by the 32 lanes of the warp that owns the env tile `S`.
// Lane mappings: bodies/geoms/contacts/joints -> lane, dofs -> lane and
// lane+32, constraint rows -> lane and lane+32.  The reference path being replaced is
// mj_forward inside mj_step (dp_env_v3.py:112 -> mujoco_py MjSim.step); stage names in the
// comments are MuJoCo's.
#pragma once
#include "dmb_math.cuh"
#include "dmb_types.cuh"

namespace dmb {

// The big per-stage phases are folded into the single forward_eval call site (one function body: the compiler
// keeps addresses and model constants in registers across phases and vectorises shared loads; +10 % on B200).
// DMB_INLINE_PHASES=0 keeps them out of line; DMB_INLINE_KIN=1 also inlines kinematics / comPos (two call sites).
#ifndef DMB_INLINE_PHASES
#define DMB_INLINE_PHASES 1
#endif
#ifndef DMB_INLINE_KIN
#define DMB_INLINE_KIN 0
#endif
// DMB_SHARE=1: CTA-wide sharing of the half-solve / Gram work among the warps of a CTA, see forward_eval (+2.4 %)
#ifndef DMB_SHARE
#define DMB_SHARE 1
#endif
#if DMB_INLINE_KIN
#define DMB_KIN_FN __forceinline__
#else
#define DMB_KIN_FN __noinline__
#endif
#if DMB_INLINE_PHASES
#define DMB_PHASE_FN __forceinline__
#else
#define DMB_PHASE_FN __noinline__
#endif

// DMB_PHASE_TIMERS (diagnostic builds only): lane 0 of every warp adds the clock64 deltas between tick points
// into g_phase_cycles[i] (see tools/gpu_phase_timers.py)
#ifndef DMB_PHASE_TIMERS
#define DMB_PHASE_TIMERS 0
#endif
#if DMB_PHASE_TIMERS
__device__ unsigned long long g_phase_cycles[32];
__device__ __forceinline__ void dmb_tick(int i, int lane) {
  __shared__ long long s_tick[32];
  if (lane == 0) {
    const long long t = clock64();
    const int w = threadIdx.y;
    // i < 0: start the clock (kernel entry); i == 0 (start of a forward evaluation) is booked on bucket 25: the step
    // prologue (state load, ctrl) before the first stage, the RK4 update + integratePos between stages
    if (i >= 0) atomicAdd(&g_phase_cycles[i == 0 ? 25 : i], (unsigned long long)(t - s_tick[w]));
    s_tick[w] = t;
  }
}
#define DMB_TICK(i) dmb_tick(i, lane)
#else
#define DMB_TICK(i) do { } while (0)
#endif


// ------------------------------------------------------------------------------------------
// mj_kinematics: lane = body, one tree level per round.  Also writes the world-frame hinge
// axes into cdof[.][0:3] (the angular part of cdof) and the geom poses (lane = geom).
// ------------------------------------------------------------------------------------------
// rotate v by the unit quaternion q:  v + 2 w (u x v) + 2 u x (u x v)
__device__ __forceinline__ V3 qrotv(Q4 q, V3 v) {
  const V3 u = v3(q.x, q.y, q.z);
  const V3 t = 2.f * cross(u, v);
  return v + q.w * t + cross(u, t);
}
// Pose composition is associative, so the world poses are a prefix "sum" of the local transforms along the body
// chains: local (body_pos, body_quat * hinges), then log2(depth) rounds of pointer jumping, lane = body.  The
// world-frame hinge axes follow from the final body quaternion by peeling the hinges off again (an axis is
// invariant under its own hinge rotation).
__device__ DMB_KIN_FN void kinematics(const ModelS& M, EnvS& S, int lane) {
  const int b = lane;
  const bool act = b >= 1 && b < M.nbody;
  const int bb = act ? b : 0;
  float sn[JPB], cs[JPB];
  V3 ax[JPB];
  const int jadr = M.body_jntadr[bb], jnum = act ? M.body_jntnum[bb] : 0;
  V3 pos = ld3(M.body_pos[bb]);
  Q4 quat; quat.w = M.body_quat[bb][0]; quat.x = M.body_quat[bb][1]; quat.y = M.body_quat[bb][2]; quat.z = M.body_quat[bb][3];
#pragma unroll
  for (int k = 0; k < JPB; k++) {
    sn[k] = 0.f; cs[k] = 1.f; ax[k] = v3(1.f, 0.f, 0.f);
    if (k < jnum) {
      const int j = jadr + k, qa = M.jnt_qposadr[j];
      if (M.jnt_type[j] == DMB_JNT_FREE) {
        pos = v3(S.qpos[qa], S.qpos[qa + 1], S.qpos[qa + 2]);
        quat.w = S.qpos[qa + 3]; quat.x = S.qpos[qa + 4]; quat.y = S.qpos[qa + 5]; quat.z = S.qpos[qa + 6];
        quat = qnormalize(quat);
      } else {
        ax[k] = ld3(M.jnt_axis[j]);
        sincosf(0.5f * (S.qpos[qa] - M.jnt_qpos0[j]), &sn[k], &cs[k]);
        Q4 ql; ql.w = cs[k]; ql.x = sn[k] * ax[k].x; ql.y = sn[k] * ax[k].y; ql.z = sn[k] * ax[k].z;
        quat = qmul(quat, ql);
      }
    }
  }
  // prefix composition along the chain: (p, q) <- (p_src + R(q_src) p, q_src q); a free joint makes the pose absolute
  // (free joints only exist on top-level bodies, so nothing is ever composed onto an absolute pose)
  const int j0 = act ? M.body_jump[0][bb] : -1, j1 = act ? M.body_jump[1][bb] : -1, j2 = act ? M.body_jump[2][bb] : -1;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    if ((1 << s) >= M.maxdepth) break;
    const int src = s == 0 ? j0 : (s == 1 ? j1 : j2);
    const float px = __shfl_sync(DMB_FULL, pos.x, src & 31), py = __shfl_sync(DMB_FULL, pos.y, src & 31), pz = __shfl_sync(DMB_FULL, pos.z, src & 31);
    Q4 qs;
    qs.w = __shfl_sync(DMB_FULL, quat.w, src & 31); qs.x = __shfl_sync(DMB_FULL, quat.x, src & 31);
    qs.y = __shfl_sync(DMB_FULL, quat.y, src & 31); qs.z = __shfl_sync(DMB_FULL, quat.z, src & 31);
    if (src >= 0) {
      pos = v3(px, py, pz) + qrotv(qs, pos);
      quat = qmul(qs, quat);
    }
  }
  if (lane == 0) {
    S.o.k.xpos[0] = S.o.k.xpos[1] = S.o.k.xpos[2] = 0.f;
    S.o.k.xquat[0] = 1.f; S.o.k.xquat[1] = S.o.k.xquat[2] = S.o.k.xquat[3] = 0.f;
    S.o.k.xmat[0] = 1.f; S.o.k.xmat[1] = 0.f; S.o.k.xmat[2] = 0.f; S.o.k.xmat[3] = 0.f; S.o.k.xmat[4] = 1.f; S.o.k.xmat[5] = 0.f;
    S.o.k.xmat[6] = 0.f; S.o.k.xmat[7] = 0.f; S.o.k.xmat[8] = 1.f;
    S.o.k.xipos[0] = S.o.k.xipos[1] = S.o.k.xipos[2] = 0.f;
  }
  if (act) {
    quat = qnormalize(quat);
    st3(&S.o.k.xpos[3 * b], pos);
    S.o.k.xquat[4 * b] = quat.w; S.o.k.xquat[4 * b + 1] = quat.x; S.o.k.xquat[4 * b + 2] = quat.y; S.o.k.xquat[4 * b + 3] = quat.z;
    float m[9];
    quat2mat(m, quat);
#pragma unroll
    for (int k = 0; k < 9; k++) S.o.k.xmat[9 * b + k] = m[k];
    st3(&S.o.k.xipos[3 * b], pos + mat_vec(m, ld3(M.body_ipos[b])));
    // world hinge axes, last joint first: axis_k = R(q after hinge k) a_k, then q <- q * conj(hinge k)
    Q4 q = quat;
#pragma unroll
    for (int k = JPB - 1; k >= 0; k--) {
      if (k < jnum && M.jnt_type[jadr + k] == DMB_JNT_HINGE) {
        st3(&S.o.k.cdof[6 * M.jnt_dofadr[jadr + k]], qrotv(q, ax[k]));
        if (k > 0) {
          Q4 qc; qc.w = cs[k]; qc.x = -sn[k] * ax[k].x; qc.y = -sn[k] * ax[k].y; qc.z = -sn[k] * ax[k].z;
          q = qmul(q, qc);
        }
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// mj_comPos: whole-model CoM (warp-shuffle reduction over bodies), cinert (lane = body),
// cdof (lane = dof).
// ------------------------------------------------------------------------------------------
__device__ DMB_KIN_FN void com_pos(const ModelS& M, EnvS& S, int lane) {
  const int b = lane;
  const bool act = b >= 1 && b < M.nbody;
  float ms = act ? M.body_mass[b] : 0.f;
  V3 xi = act ? ld3(&S.o.k.xipos[3 * b]) : v3(0.f, 0.f, 0.f);
  float cx = warp_sum(ms * xi.x) * M.inv_total_mass;
  float cy = warp_sum(ms * xi.y) * M.inv_total_mass;
  float cz = warp_sum(ms * xi.z) * M.inv_total_mass;
  if (lane == 0) { S.com[0] = cx; S.com[1] = cy; S.com[2] = cz; }
  const V3 com = v3(cx, cy, cz);
  if (lane < M.nbody) {
    float* ci = &S.o.k.cinert[10 * b];
    if (!act) {
#pragma unroll
      for (int k = 0; k < 10; k++) ci[k] = 0.f;
    } else {
      const float* I = M.body_inertia[b];
      const float* R = &S.o.k.xmat[9 * b];
      // W = R * Ib * R'
      float Ib[9] = {I[0], I[3], I[4], I[3], I[1], I[5], I[4], I[5], I[2]};
      float RI[9], W[9];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) RI[3 * r + c] = R[3 * r] * Ib[c] + R[3 * r + 1] * Ib[3 + c] + R[3 * r + 2] * Ib[6 + c];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) W[3 * r + c] = RI[3 * r] * R[3 * c] + RI[3 * r + 1] * R[3 * c + 1] + RI[3 * r + 2] * R[3 * c + 2];
      V3 d = xi - com;
      ci[0] = W[0] + ms * (d.y * d.y + d.z * d.z);
      ci[1] = W[4] + ms * (d.x * d.x + d.z * d.z);
      ci[2] = W[8] + ms * (d.x * d.x + d.y * d.y);
      ci[3] = W[1] - ms * d.x * d.y;
      ci[4] = W[2] - ms * d.x * d.z;
      ci[5] = W[5] - ms * d.y * d.z;
      ci[6] = ms * d.x; ci[7] = ms * d.y; ci[8] = ms * d.z; ci[9] = ms;
    }
  }
  for (int d = lane; d < M.nv; d += 32) {
    const int db = M.dof_bodyid[d];
    const V3 off = com - ld3(&S.o.k.xpos[3 * db]);
    float* cd = &S.o.k.cdof[6 * d];
    const int kind = M.dof_kind[d], k = M.dof_axisk[d];
    if (kind == DOF_FREE_TRANS) {
      cd[0] = cd[1] = cd[2] = 0.f;
      cd[3] = k == 0 ? 1.f : 0.f; cd[4] = k == 1 ? 1.f : 0.f; cd[5] = k == 2 ? 1.f : 0.f;
    } else {
      V3 ax;
      if (kind == DOF_FREE_ROT) { ax = v3(S.o.k.xmat[9 * db + k], S.o.k.xmat[9 * db + 3 + k], S.o.k.xmat[9 * db + 6 + k]); st3(cd, ax); }
      else ax = ld3(cd);  // written by kinematics()
      st3(cd + 3, cross(ax, off));
    }
  }
  __syncwarp();
}

#ifndef DMB_FAST_RCP
#define DMB_FAST_RCP 1
#endif
// reciprocal for pivots / scalings: one MUFU (1 ulp) instead of the correctly rounded sequence
__device__ __forceinline__ float rcp(float x) {
#if DMB_FAST_RCP
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return __frcp_rn(x);
#endif
}

// ------------------------------------------------------------------------------------------
// mj_crb + mj_factorM: composite inertias (parents gather children, level by level), the
// nM sparse inertia entries (lane = entry) and the in-place sparse L'DL factorisation.
// ------------------------------------------------------------------------------------------
// All shared-memory reads of a step are issued before its writes (values staged in registers): the
// compiler cannot prove that the tile's arrays do not alias, and a load queued behind an earlier store
// would serialise the ~30-cycle shared-memory round trips of independent lanes' work.
__device__ DMB_PHASE_FN void crb_factor(const ModelS& M, EnvS& S, int lane, float* dbg_qM) {
  const int b = lane;
  if (lane < M.nbody) {
#pragma unroll
    for (int k = 0; k < 10; k++) S.o.k.crb[10 * b + k] = S.o.k.cinert[10 * b + k];
  }
  __syncwarp();
  for (int lev = M.maxdepth - 1; lev >= 1; lev--) {
    if (b >= 1 && b < M.nbody && M.body_depth[b] == lev) {
      const int nc = M.body_nchild[b];
      if (nc > 0) {
        float acc[10];
#pragma unroll
        for (int k = 0; k < 10; k++) acc[k] = S.o.k.crb[10 * b + k];
        for (int c = 0; c < nc; c++) {
          const int ch = M.body_child[b][c];
#pragma unroll
          for (int k = 0; k < 10; k++) acc[k] += S.o.k.crb[10 * ch + k];
        }
#pragma unroll
        for (int k = 0; k < 10; k++) S.o.k.crb[10 * b + k] = acc[k];
      }
    }
    __syncwarp();
  }
  DMB_TICK(11);
  for (int d = lane; d < M.nv; d += 32) mul_inert_vec(&S.o.k.buf6[6 * d], &S.o.k.crb[10 * M.dof_bodyid[d]], &S.o.k.cdof[6 * d]);
  __syncwarp();
  constexpr int NR = NMX / 32;  // rounds of 32 inertia entries
  {
    float val[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int e = lane + 32 * r;
      val[r] = 0.f;
      if (e < M.nM) {
        const int i = M.M_i[e], j = M.M_j[e];
        const float* a = &S.o.k.cdof[6 * j];
        const float* bf = &S.o.k.buf6[6 * i];
        float s = a[0] * bf[0] + a[1] * bf[1] + a[2] * bf[2] + a[3] * bf[3] + a[4] * bf[4] + a[5] * bf[5];
        if (i == j) s += M.dof_armature[i];
        val[r] = s;
      }
    }
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int e = lane + 32 * r;
      if (e < M.nM) { S.qLD[e] = val[r]; if (dbg_qM) dbg_qM[e] = val[r]; }
    }
  }
  __syncwarp();
  // sparse L'DL (mj_factorM): for k = nv-1..0 eliminate dof k from all its ancestors.  Each lane
  // keeps the (p, q) decode of its <= 3 update slots in registers; anc_rowbase[k][p] is the packed
  // address of the row of k's p-th ancestor; rows are left unscaled during the elimination and divided
  // by their pivot in one parallel pass at the end.
  DMB_TICK(12);
  // Per-lane constants of the (p, q) slots; per-step constants come packed in one word (ldl_meta).  Lanes whose
  // slot lies outside the step's pair range still load (their p, q address valid words of the tile) and skip the store.
  int tp[3], tq[3];
#pragma unroll
  for (int u = 0; u < 3; u++) { const int t = lane + 32 * u; tp[u] = t < 78 ? M.tri_p[t] : 0; tq[u] = t < 78 ? M.tri_q[t] : 0; }
  const int qmp0 = tq[0] - tp[0], qmp1 = tq[1] - tp[1], qmp2 = tq[2] - tp[2];
  for (int k = M.nv - 1; k >= 1; k--) {
    const unsigned meta = M.ldl_meta[k];
    const int npair = (int)((meta >> 8) & 0xffu);
    if (npair == 0) continue;
    const float* rowk = &S.qLD[(meta >> 16) + 1];
    const int16_t* rb = M.anc_rowbase[k];
    const float piv = rowk[-1];
    // (slots past the step's pair range read the pivot word instead of a stray address, so that no lane ever reads a
    // word another lane writes in the same step)
    const int pva = (int)(meta >> 16);
    if (npair <= 32) {
      const int d0 = lane < npair ? rb[tp[0]] + qmp0 : pva;
      const float p0 = rowk[tp[0]], q0 = rowk[tq[0]], v0 = S.qLD[d0];
      const float inv = rcp(piv);
      if (lane < npair) S.qLD[d0] = fmaf(-(p0 * inv), q0, v0);
    } else if (npair <= 64) {
      const int d0 = rb[tp[0]] + qmp0, d1 = lane + 32 < npair ? rb[tp[1]] + qmp1 : pva;
      const float p0 = rowk[tp[0]], q0 = rowk[tq[0]], v0 = S.qLD[d0];
      const float p1 = rowk[tp[1]], q1 = rowk[tq[1]], v1 = S.qLD[d1];
      const float inv = rcp(piv);
      S.qLD[d0] = fmaf(-(p0 * inv), q0, v0);
      if (lane + 32 < npair) S.qLD[d1] = fmaf(-(p1 * inv), q1, v1);
    } else {
      const int d0 = rb[tp[0]] + qmp0, d1 = rb[tp[1]] + qmp1, d2 = lane + 64 < npair ? rb[tp[2]] + qmp2 : pva;
      const float p0 = rowk[tp[0]], q0 = rowk[tq[0]], v0 = S.qLD[d0];
      const float p1 = rowk[tp[1]], q1 = rowk[tq[1]], v1 = S.qLD[d1];
      const float p2 = rowk[tp[2]], q2 = rowk[tq[2]], v2 = S.qLD[d2];
      const float inv = rcp(piv);
      S.qLD[d0] = fmaf(-(p0 * inv), q0, v0);
      S.qLD[d1] = fmaf(-(p1 * inv), q1, v1);
      if (lane + 64 < npair) S.qLD[d2] = fmaf(-(p2 * inv), q2, v2);
    }
    __syncwarp();
  }
  DMB_TICK(13);
  // D^-1/2 for the half solves; L entries = row / pivot (1/D = (D^-1/2)^2)
  for (int d = lane; d < M.nv; d += 32) S.dsq[d] = rsqrtf(S.qLD[M.dof_Madr[d]]);
  __syncwarp();
  {
    float val[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int e = lane + 32 * r;
      val[r] = 0.f;
      if (e < M.nM) {
        const int i = M.M_i[e];
        const float sq = S.dsq[i], x = S.qLD[e];
        val[r] = M.M_j[e] != i ? x * (sq * sq) : x;
      }
    }
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int e = lane + 32 * r;
      if (e < M.nM) S.qLD[e] = val[r];
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// mj_comVel + mj_rne(flg_acc=0) + passive + actuation: leaves the smooth generalised force
// qfrc_smooth = passive - bias + actuator in S.x_dv.
// ------------------------------------------------------------------------------------------
// Inclusive sums of a spatial 6-vector along the dof ancestor chains, dof d on lane d & 31 (x: d < 32,
// xh: d >= 32), by pointer jumping: 4 rounds cover chains of up to 16 dofs.
__device__ __forceinline__ void chain_scan6(const ModelS& M, int lane, int nv, const int (&jmp)[4], float* x, float* xh) {
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int src = jmp[s];
#pragma unroll
    for (int k = 0; k < 6; k++) { const float t = __shfl_sync(DMB_FULL, x[k], src & 31); if (src >= 0) x[k] += t; }
  }
  for (int d = 32; d < nv; d++) {  // the few dofs past lane 31 hang off finished sums
    const int p = M.dof_jump[0][d];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const float tl = __shfl_sync(DMB_FULL, x[k], p & 31), th = __shfl_sync(DMB_FULL, xh[k], p & 31);
      if (lane == d - 32 && p >= 0) xh[k] += p < 32 ? tl : th;
    }
  }
}
__device__ DMB_PHASE_FN void smooth_forces(const ModelS& M, EnvS& S, int lane, float* dbgrow) {
  const int nv = M.nv;
  const bool has_lo = lane < nv, has_hi = lane + 32 < nv;
  const int dl = has_lo ? lane : 0, dh = has_hi ? lane + 32 : 0;
  const int jmp[4] = {has_lo ? M.dof_jump[0][dl] : -1, has_lo ? M.dof_jump[1][dl] : -1, has_lo ? M.dof_jump[2][dl] : -1,
                      has_lo ? M.dof_jump[3][dl] : -1};
  // body velocities: V = chain sum of cdof * qvel (mj_comVel)
  const float qvl = has_lo ? S.qvel[dl] : 0.f, qvh = has_hi ? S.qvel[dh] : 0.f;
  float cl[6], ch[6], V[6], Vh[6];
#pragma unroll
  for (int k = 0; k < 6; k++) { cl[k] = S.o.k.cdof[6 * dl + k]; ch[k] = S.o.k.cdof[6 * dh + k]; V[k] = cl[k] * qvl; Vh[k] = ch[k] * qvh; }
  chain_scan6(M, lane, nv, jmp, V, Vh);
  // cdof_dot[d] = crossMotion(velocity seen by dof d, cdof[d]); A = chain sum of cdof_dot * qvel
  const int sl = has_lo ? M.dof_vsrc[dl] : -1, sh = has_hi ? M.dof_vsrc[dh] : -1;
  float E[6], Eh[6], A[6], Ah[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const float t = __shfl_sync(DMB_FULL, V[k], sl & 31);
    const float tl = __shfl_sync(DMB_FULL, V[k], sh & 31), th = __shfl_sync(DMB_FULL, Vh[k], sh & 31);
    E[k] = sl >= 0 ? t : 0.f;
    Eh[k] = sh >= 0 ? (sh < 32 ? tl : th) : 0.f;
  }
  cross_motion(A, E, cl);
  cross_motion(Ah, Eh, ch);
#pragma unroll
  for (int k = 0; k < 6; k++) { A[k] = sl >= 0 ? A[k] * qvl : 0.f; Ah[k] = sh >= 0 ? Ah[k] * qvh : 0.f; }
  chain_scan6(M, lane, nv, jmp, A, Ah);
  // hand the sums of each body's last dof to the body lanes (velocity -> S.o.k.cvel, acceleration -> scratch)
  float* acc = S.o.k.acc;
  if (lane < 6) S.o.k.cvel[lane] = 0.f;
  const int bl = has_lo ? M.dof_lastof[dl] : -1, bh = has_hi ? M.dof_lastof[dh] : -1;
  if (bl >= 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) { S.o.k.cvel[6 * bl + k] = V[k]; acc[6 * bl + k] = A[k]; }
  }
  if (bh >= 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) { S.o.k.cvel[6 * bh + k] = Vh[k]; acc[6 * bh + k] = Ah[k]; }
  }
  __syncwarp();
  // body forces  f = I a + v x* (I v)  with a = -gravity + chain acceleration
  if (lane < M.nbody) {
    const int b = lane;
    float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (b > 0) {
      float v[6], a[6], t1[6], t2[6];
#pragma unroll
      for (int k = 0; k < 6; k++) { v[k] = S.o.k.cvel[6 * b + k]; a[k] = acc[6 * b + k]; }
      a[3] -= M.gravity[0]; a[4] -= M.gravity[1]; a[5] -= M.gravity[2];
      mul_inert_vec(f, &S.o.k.cinert[10 * b], a);
      mul_inert_vec(t1, &S.o.k.cinert[10 * b], v);
      cross_force(t2, v, t1);
#pragma unroll
      for (int k = 0; k < 6; k++) f[k] += t2[k];
    }
#pragma unroll
    for (int k = 0; k < 6; k++) S.o.k.cfrc[6 * b + k] = f[k];
  }
  __syncwarp();
  for (int lev = M.maxdepth - 1; lev >= 1; lev--) {
    const int b = lane;
    if (b >= 1 && b < M.nbody && M.body_depth[b] == lev) {
      const int nc = M.body_nchild[b];
      if (nc > 0) {
        float acc[6];
#pragma unroll
        for (int k = 0; k < 6; k++) acc[k] = S.o.k.cfrc[6 * b + k];
        for (int c = 0; c < nc; c++) {
          const int ch = M.body_child[b][c];
#pragma unroll
          for (int k = 0; k < 6; k++) acc[k] += S.o.k.cfrc[6 * ch + k];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) S.o.k.cfrc[6 * b + k] = acc[k];
      }
    }
    __syncwarp();
  }
  for (int d = lane; d < M.nv; d += 32) {
    const float* cd = &S.o.k.cdof[6 * d];
    const float* cf = &S.o.k.cfrc[6 * M.dof_bodyid[d]];
    const float bias = cd[0] * cf[0] + cd[1] * cf[1] + cd[2] * cf[2] + cd[3] * cf[3] + cd[4] * cf[4] + cd[5] * cf[5];
    const float fs = -M.dof_damping[d] * S.qvel[d] - bias + S.ctrlf[d];
    if (dbgrow) { dbgrow[dbg::qfrc_bias + d] = bias; dbgrow[dbg::qfrc_smooth + d] = fs; }
    S.x_dv[d] = fs;   // qfrc_smooth (x_dv is scratch inside a forward evaluation)
  }
  __syncwarp();
}

// geom poses (lane = geom); runs after the smooth stage, writing over dead phase-A arrays
__device__ __forceinline__ void geom_poses(const ModelS& M, EnvS& S, int lane) {
  if (lane < M.ngeom) {
    const int g = lane, gb = M.geom_bodyid[g];
    st3(&S.o.c.gpos[3 * g], ld3(&S.o.c.xpos[3 * gb]) + mat_vec(&S.o.c.xmat[9 * gb], ld3(M.geom_pos[g])));
    if (M.geom_identq[g]) {
#pragma unroll
      for (int k = 0; k < 9; k++) S.o.c.gmat[9 * g + k] = S.o.c.xmat[9 * gb + k];
    } else {
      Q4 qb; qb.w = S.o.c.xquat[4 * gb]; qb.x = S.o.c.xquat[4 * gb + 1]; qb.y = S.o.c.xquat[4 * gb + 2]; qb.z = S.o.c.xquat[4 * gb + 3];
      Q4 qg; qg.w = M.geom_quat[g][0]; qg.x = M.geom_quat[g][1]; qg.y = M.geom_quat[g][2]; qg.z = M.geom_quat[g][3];
      float m[9];
      quat2mat(m, qmul(qb, qg));
#pragma unroll
      for (int k = 0; k < 9; k++) S.o.c.gmat[9 * g + k] = m[k];
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// narrow phase (restates oracle/dm_oracle.c raw_* functions in fp32)
// ------------------------------------------------------------------------------------------
struct RawCon { float dist; V3 pos, n, y; };

__device__ __forceinline__ int raw_plane_sphere(RawCon& c, float margin, V3 ppos, V3 pn, V3 spos, float r) {
  const float cdist = dot(spos - ppos, pn);
  if (cdist > margin + r) return 0;
  c.dist = cdist - r;
  c.pos = spos - (r + 0.5f * c.dist) * pn;
  c.n = pn;
  c.y = v3(0.f, 0.f, 0.f);
  return 1;
}
__device__ __forceinline__ int raw_sphere_sphere(RawCon& c, float margin, V3 p1, float r1, V3 p2, float r2) {
  V3 dif = p2 - p1;
  const float dist = norm(dif);
  if (dist > margin + r1 + r2) return 0;
  c.dist = dist - r1 - r2;
  c.n = dist < DMB_MINVAL ? v3(1.f, 0.f, 0.f) : (1.0f / dist) * dif;
  c.pos = p1 + (r1 + 0.5f * c.dist) * c.n;
  c.y = v3(0.f, 0.f, 0.f);
  return 1;
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x > hi ? hi : (x < lo ? lo : x); }

__device__ int raw_sphere_box(RawCon& c, float margin, V3 spos, float r, V3 bpos, const float* bmat, V3 bs) {
  const V3 center = matT_vec(bmat, spos - bpos);
  const V3 cl = v3(clampf(center.x, -bs.x, bs.x), clampf(center.y, -bs.y, bs.y), clampf(center.z, -bs.z, bs.z));
  const V3 deep = center - cl;
  const float dist = norm(deep);
  if (dist - r > margin) return 0;
  V3 nloc, ploc;
  if (dist <= DMB_MINVAL) {
    float gx = bs.x - fabsf(center.x), gy = bs.y - fabsf(center.y), gz = bs.z - fabsf(center.z);
    float closest = gx; int kk = 0;
    if (gy < closest) { closest = gy; kk = 1; }
    if (gz < closest) { closest = gz; kk = 2; }
    V3 nf = v3(0.f, 0.f, 0.f);
    if (kk == 0) nf.x = center.x > 0.f ? 1.f : -1.f;
    else if (kk == 1) nf.y = center.y > 0.f ? 1.f : -1.f;
    else nf.z = center.z > 0.f ? 1.f : -1.f;
    c.dist = -closest - r;
    nloc = -1.f * nf;
    ploc = center + (0.5f * (closest - r)) * nf;
  } else {
    c.dist = dist - r;
    const V3 nout = (1.0f / dist) * deep;
    nloc = -1.f * nout;
    ploc = cl + (0.5f * c.dist) * nout;
  }
  c.n = mat_vec(bmat, nloc);
  c.pos = mat_vec(bmat, ploc) + bpos;
  c.y = v3(0.f, 0.f, 0.f);
  return 1;
}

__device__ __forceinline__ float capbox_dfdt(V3 c, V3 u, V3 s, float t) {
  float g = 0.f, p;
  p = c.x + t * u.x; if (p > s.x) g += (p - s.x) * u.x; else if (p < -s.x) g += (p + s.x) * u.x;
  p = c.y + t * u.y; if (p > s.y) g += (p - s.y) * u.y; else if (p < -s.y) g += (p + s.y) * u.y;
  p = c.z + t * u.z; if (p > s.z) g += (p - s.z) * u.z; else if (p < -s.z) g += (p + s.z) * u.z;
  return g;
}

__device__ int raw_capsule_box(RawCon* c, float margin, V3 cpos, const float* cmat, float r, float h, V3 bpos,
                               const float* bmat, V3 bs) {
  const V3 axis = v3(cmat[2], cmat[5], cmat[8]);
  const V3 cl = matT_vec(bmat, cpos - bpos);
  const V3 u = matT_vec(bmat, axis);
  float knot[8];
  int nk = 0;
  knot[nk++] = -h;
  const float cc[3] = {cl.x, cl.y, cl.z}, uu[3] = {u.x, u.y, u.z}, ss[3] = {bs.x, bs.y, bs.z};
  for (int k = 0; k < 3; k++) {
    if (fabsf(uu[k]) > 1e-12f) {
      const float t1 = (ss[k] - cc[k]) / uu[k], t2 = (-ss[k] - cc[k]) / uu[k];
      if (t1 > -h && t1 < h) knot[nk++] = t1;
      if (t2 > -h && t2 < h) knot[nk++] = t2;
    }
  }
  knot[nk++] = h;
  for (int i = 1; i < nk; i++) {
    float v = knot[i];
    int j = i - 1;
    while (j >= 0 && knot[j] > v) { knot[j + 1] = knot[j]; j--; }
    knot[j + 1] = v;
  }
  float tlo, thi;
  // the axis runs through the box: zero distance on the clipped stretch [tin, tout] (see oracle raw_capsule_box)
  float tin = -h, tout = h;
  for (int k = 0; k < 3; k++) {
    if (fabsf(uu[k]) > 1e-12f) {
      float ta = (-ss[k] - cc[k]) / uu[k], tb = (ss[k] - cc[k]) / uu[k];
      if (ta > tb) { const float t = ta; ta = tb; tb = t; }
      tin = fmaxf(tin, ta); tout = fminf(tout, tb);
    } else if (fabsf(cc[k]) > ss[k]) { tin = h; tout = -h; }
  }
  if (tout - tin > 1e-6f * h) { tlo = tin; thi = tout; }
  else {
  {
    float gprev = capbox_dfdt(cl, u, bs, knot[0]);
    if (gprev >= 0.f) tlo = knot[0];
    else {
      tlo = knot[nk - 1];
      for (int i = 1; i < nk; i++) {
        const float g = capbox_dfdt(cl, u, bs, knot[i]);
        if (g >= 0.f) { tlo = knot[i - 1] - gprev * (knot[i] - knot[i - 1]) / (g - gprev); break; }
        gprev = g;
      }
    }
  }
  {
    float gnext = capbox_dfdt(cl, u, bs, knot[nk - 1]);
    if (gnext <= 0.f) thi = knot[nk - 1];
    else {
      thi = knot[0];
      for (int i = nk - 2; i >= 0; i--) {
        const float g = capbox_dfdt(cl, u, bs, knot[i]);
        if (g <= 0.f) { thi = knot[i + 1] - gnext * (knot[i] - knot[i + 1]) / (g - gnext); break; }
        gnext = g;
      }
    }
  }
  }
  // flat stretch (see oracle raw_capsule_box): the end nearer the capsule centre, moved 1e-3 of the stretch inwards
  float ts = 0.5f * (tlo + thi);
  if (thi - tlo > 1e-6f * h) ts = fabsf(tlo) <= fabsf(thi) ? tlo + 1e-3f * (thi - tlo) : thi - 1e-3f * (thi - tlo);
  int n = 0;
  n += raw_sphere_box(c[n], margin, cpos + ts * axis, r, bpos, bmat, bs);
  const float te = (h - ts >= ts + h) ? h : -h;
  if (fabsf(te - ts) > 0.01f * h) n += raw_sphere_box(c[n], margin, cpos + te * axis, r, bpos, bmat, bs);
  return n;
}

__device__ int raw_capsule_capsule(RawCon* c, float margin, V3 pos1, const float* mat1, float r1, float h1, V3 pos2,
                                   const float* mat2, float r2, float h2) {
  const V3 a1 = v3(mat1[2], mat1[5], mat1[8]), a2 = v3(mat2[2], mat2[5], mat2[8]);
  const V3 dif = pos1 - pos2;
  const float ma = dot(a1, a1), mb = -dot(a1, a2), mc = dot(a2, a2);
  const float u = -dot(a1, dif), v = dot(a2, dif);
  const float det = ma * mc - mb * mb;
  if (fabsf(det) >= 1e-10f) {
    float x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > h1) { x1 = h1; x2 = (v - mb * h1) / mc; }
    else if (x1 < -h1) { x1 = -h1; x2 = (v + mb * h1) / mc; }
    if (x2 > h2) { x2 = h2; x1 = (u - mb * h2) / ma; }
    else if (x2 < -h2) { x2 = -h2; x1 = (u + mb * h2) / ma; }
    x1 = clampf(x1, -h1, h1);
    return raw_sphere_sphere(c[0], margin, pos1 + x1 * a1, r1, pos2 + x2 * a2, r2);
  }
  int n = 0;
  float x1, x2;
  x2 = clampf((v - mb * h1) / mc, -h2, h2);
  n += raw_sphere_sphere(c[n], margin, pos1 + h1 * a1, r1, pos2 + x2 * a2, r2);
  x2 = clampf((v + mb * h1) / mc, -h2, h2);
  n += raw_sphere_sphere(c[n], margin, pos1 - h1 * a1, r1, pos2 + x2 * a2, r2);
  if (n == 2) return n;
  x1 = clampf((u - mb * h2) / ma, -h1, h1);
  n += raw_sphere_sphere(c[n], margin, pos1 + x1 * a1, r1, pos2 + h2 * a2, r2);
  if (n == 2) return n;
  x1 = clampf((u + mb * h2) / ma, -h1, h1);
  n += raw_sphere_sphere(c[n], margin, pos1 + x1 * a1, r1, pos2 - h2 * a2, r2);
  return n;
}

__device__ int raw_box_box(RawCon* c, float margin, V3 pos1, const float* mat1, V3 s1, V3 pos2, const float* mat2,
                           V3 s2) {
  int n = 0;
  // exact early-out on the 6 face axes (feet are almost always inside each other's bounding sphere)
  for (int pass = 0; pass < 2; pass++) {
    const V3 pa = pass == 0 ? pos1 : pos2, pb = pass == 0 ? pos2 : pos1;
    const float* ma = pass == 0 ? mat1 : mat2;
    const float* mb = pass == 0 ? mat2 : mat1;
    const V3 sa = pass == 0 ? s1 : s2, sb = pass == 0 ? s2 : s1;
    const V3 dc = pb - pa;
    const float sak[3] = {sa.x, sa.y, sa.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const V3 ax = v3(ma[k], ma[3 + k], ma[6 + k]);
      const float r = sb.x * fabsf(ax.x * mb[0] + ax.y * mb[3] + ax.z * mb[6]) +
                      sb.y * fabsf(ax.x * mb[1] + ax.y * mb[4] + ax.z * mb[7]) +
                      sb.z * fabsf(ax.x * mb[2] + ax.y * mb[5] + ax.z * mb[8]);
      if (fabsf(dot(dc, ax)) - sak[k] - r > margin) return 0;
    }
  }
  for (int pass = 0; pass < 2 && n < 4; pass++) {
    const V3 vp = pass == 0 ? pos2 : pos1;
    const float* vm = pass == 0 ? mat2 : mat1;
    const V3 vs = pass == 0 ? s2 : s1;
    const V3 bp = pass == 0 ? pos1 : pos2;
    const float* bm = pass == 0 ? mat1 : mat2;
    const V3 bs = pass == 0 ? s1 : s2;
    for (int i = 0; i < 8 && n < 4; i++) {
      const V3 loc = v3((i & 1) ? vs.x : -vs.x, (i & 2) ? vs.y : -vs.y, (i & 4) ? vs.z : -vs.z);
      const V3 w = mat_vec(vm, loc) + vp;
      if (raw_sphere_box(c[n], margin, w, 0.f, bp, bm, bs)) {
        if (pass == 0) c[n].n = -1.f * c[n].n;
        n++;
      }
    }
  }
  return n;
}

// mju_makeFrame
__device__ __forceinline__ void make_frame(float* f, V3 n, V3 y) {
  normalize(n);
  if (norm(y) < 0.5f) {
    y = v3(0.f, 0.f, 0.f);
    if (n.y < 0.5f && n.y > -0.5f) y.y = 1.f; else y.z = 1.f;
  }
  y = y - dot(n, y) * n;
  normalize(y);
  st3(f, n); st3(f + 3, y);   // the second tangent is cross(n, y), recomputed where needed
}

// ------------------------------------------------------------------------------------------
// mj_collision: broad phase (lane = candidate pair, 4 rounds) -> compacted survivor list ->
// narrow phase (lane = survivor) -> contacts compacted in pair order by warp prefix sums.
// ------------------------------------------------------------------------------------------
__device__ DMB_PHASE_FN void collision(const ModelS& M, EnvS& S, int lane) {
  const float margin = M.margin;
  int nsurv = 0;
  for (int base = 0; base < M.npair; base += 32) {
    const int p = base + lane;
    bool keep = false;
    if (p < M.npair) {
      const int g1 = M.pair_g1[p], g2 = M.pair_g2[p];
      const V3 dif = ld3(&S.o.c.gpos[3 * g2]) - ld3(&S.o.c.gpos[3 * g1]);
      if (M.geom_type[g1] == DMB_GEOM_PLANE) {
        const V3 nrm = v3(S.o.c.gmat[9 * g1 + 2], S.o.c.gmat[9 * g1 + 5], S.o.c.gmat[9 * g1 + 8]);
        keep = !(dot(dif, nrm) > M.geom_rbound[g2] + margin);
      } else {
        const float bound = M.geom_rbound[g1] + M.geom_rbound[g2] + margin;
        keep = !(dot(dif, dif) > bound * bound);
      }
    }
    const unsigned bal = __ballot_sync(DMB_FULL, keep);
    if (keep) S.o.c.surv[nsurv + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)p;
    nsurv += __popc(bal);
  }
  __syncwarp();
  int ncon = 0;
  for (int base = 0; base < nsurv; base += 32) {
    RawCon rc[4];
    int n = 0, g1 = 0, g2 = 0;
    if (base + lane < nsurv) {
      const int p = S.o.c.surv[base + lane];
      g1 = M.pair_g1[p]; g2 = M.pair_g2[p];
      const int t1 = M.geom_type[g1], t2 = M.geom_type[g2];
      const V3 pos1 = ld3(&S.o.c.gpos[3 * g1]), pos2 = ld3(&S.o.c.gpos[3 * g2]);
      const float* mat1 = &S.o.c.gmat[9 * g1];
      const float* mat2 = &S.o.c.gmat[9 * g2];
      const V3 s1 = ld3(M.geom_size[g1]), s2 = ld3(M.geom_size[g2]);
      if (t1 == DMB_GEOM_PLANE) {
        const V3 nrm = v3(mat1[2], mat1[5], mat1[8]);
        if (t2 == DMB_GEOM_SPHERE) {
          n = raw_plane_sphere(rc[0], margin, pos1, nrm, pos2, s2.x);
        } else if (t2 == DMB_GEOM_CAPSULE) {
          const V3 axis = v3(mat2[2], mat2[5], mat2[8]);
          n += raw_plane_sphere(rc[n], margin, pos1, nrm, pos2 + s2.y * axis, s2.x);
          n += raw_plane_sphere(rc[n], margin, pos1, nrm, pos2 - s2.y * axis, s2.x);
          for (int i = 0; i < n; i++) rc[i].y = axis;
        } else if (t2 == DMB_GEOM_BOX) {
          const float dist = dot(pos2 - pos1, nrm);
          for (int i = 0; i < 8 && n < 4; i++) {
            const V3 vec = v3((i & 1) ? s2.x : -s2.x, (i & 2) ? s2.y : -s2.y, (i & 4) ? s2.z : -s2.z);
            const V3 corner = mat_vec(mat2, vec);
            const float ldist = dot(nrm, corner);
            if (dist + ldist > margin || ldist > 0.f) continue;
            rc[n].dist = dist + ldist;
            rc[n].pos = corner + pos2 - (0.5f * rc[n].dist) * nrm;
            rc[n].n = nrm;
            rc[n].y = v3(0.f, 0.f, 0.f);
            n++;
          }
        }
      } else if (t1 == DMB_GEOM_SPHERE && t2 == DMB_GEOM_SPHERE) {
        n = raw_sphere_sphere(rc[0], margin, pos1, s1.x, pos2, s2.x);
      } else if (t1 == DMB_GEOM_SPHERE && t2 == DMB_GEOM_CAPSULE) {
        const V3 axis = v3(mat2[2], mat2[5], mat2[8]);
        const float x = clampf(dot(axis, pos1 - pos2), -s2.y, s2.y);
        n = raw_sphere_sphere(rc[0], margin, pos1, s1.x, pos2 + x * axis, s2.x);
      } else if (t1 == DMB_GEOM_SPHERE && t2 == DMB_GEOM_BOX) {
        n = raw_sphere_box(rc[0], margin, pos1, s1.x, pos2, mat2, s2);
      } else if (t1 == DMB_GEOM_CAPSULE && t2 == DMB_GEOM_CAPSULE) {
        n = raw_capsule_capsule(rc, margin, pos1, mat1, s1.x, s1.y, pos2, mat2, s2.x, s2.y);
      } else if (t1 == DMB_GEOM_CAPSULE && t2 == DMB_GEOM_BOX) {
        n = raw_capsule_box(rc, margin, pos1, mat1, s1.x, s1.y, pos2, mat2, s2);
      } else if (t1 == DMB_GEOM_BOX && t2 == DMB_GEOM_BOX) {
        n = raw_box_box(rc, margin, pos1, mat1, s1, pos2, mat2, s2);
      }
    }
    const int incl = warp_incl_scan(n, lane);
    const int total = __shfl_sync(DMB_FULL, incl, 31);
    const int off = ncon + incl - n;
    for (int i = 0; i < n; i++) {
      const int ci = off + i;
      if (ci < M.max_con) {
        S.o.c.c_dist[ci] = rc[i].dist;
        st3(&S.o.c.c_pos[3 * ci], rc[i].pos);
        make_frame(&S.o.c.c_frame[6 * ci], rc[i].n, rc[i].y);
        const int cd1 = M.geom_condim[g1], cd2 = M.geom_condim[g2];
        S.c_meta[ci] = g1 | (g2 << 8) | ((cd1 > cd2 ? cd1 : cd2) << 16);   // first row (bits 24..31) set by count_rows
      }
    }
    ncon += total;
  }
  if (ncon > M.max_con) { ncon = M.max_con; if (lane == 0) S.flags |= 1; }
  if (lane == 0) S.ncon = ncon;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// mj_makeConstraint + mj_makeImpedance + mj_referenceConstraint, producing
//   Y rows 0..nefc-1 = J rows (lane = dof), per-row R / aref / dof support (lane = row).
// Limit rows come first (joint order, lower then upper), then contacts in contact order:
// 1 frictionless row (condim 1) or 4 pyramid edges (condim 3).
// Row storage: the tile's PhaseR for up to RF rows; else (OVF) Y in the warp's global scratch slot G and the
// row scalars / AR in the tile's PhaseRbig.
// ------------------------------------------------------------------------------------------
template <bool OVF> struct RowView {
  float *Y, *e_R, *e_aref, *e_b, *e_f, *AR;
  unsigned long long* rowmask;
  signed char* e_src;
  __device__ __forceinline__ RowView(EnvS& S, float* G) {
    if (OVF) {
      Y = G + gs::Y; e_src = reinterpret_cast<signed char*>(G + gs::e_src);
      e_R = S.o.rb.e_R; e_aref = S.o.rb.e_aref; e_b = S.o.rb.e_b; e_f = S.o.rb.e_f; AR = S.o.rb.AR;
      rowmask = S.o.rb.rowmask;
    } else {
      Y = S.o.r.Y; e_R = S.o.r.e_R; e_aref = S.o.r.e_aref; e_b = S.o.r.e_b; e_f = S.o.r.e_f; AR = S.o.r.AR;
      rowmask = S.o.r.rowmask; e_src = S.e_src;
    }
  }
};
// limit row source: -(2 * (1 + joint) + side), side 0 = lower (J = +1), 1 = upper (J = -1)
__device__ __forceinline__ int lim_joint(int src) { return ((-src) >> 1) - 1; }
__device__ __forceinline__ int lim_side(int src) { return (-src) & 1; }

// Row bookkeeping: active limits (lane = joint), contact row addresses (lane = contact), capacity check.
// Writes e_src and the first-row field of c_meta; returns nefc.
// warp that holds the CTA's one shared slot for a stage with more than RF rows (-1: free); reset by k_step
__shared__ int s_slot_owner;

__device__ DMB_PHASE_FN int count_rows(const ModelS& M, EnvS& S, int lane, float* Gglobal, float* slot, int warp) {
  int cnt = 0;
  bool lo = false, hi = false;
  if (lane < M.njnt && M.jnt_limited[lane] && M.jnt_type[lane] == DMB_JNT_HINGE) {
    const float q = S.qpos[M.jnt_qposadr[lane]];
    lo = q - M.jnt_range[lane][0] < 0.f;
    hi = M.jnt_range[lane][1] - q < 0.f;
    cnt = (int)lo + (int)hi;
  }
  const int incl = warp_incl_scan(cnt, lane);
  const int nlimit = __shfl_sync(DMB_FULL, incl, 31);
  int ncon = S.ncon;
  int nrow = 0;
  if (lane < ncon) nrow = cm_dim(S.c_meta[lane]) == 1 ? 1 : 4;
  const int cincl = warp_incl_scan(nrow, lane);
  const int adr = nlimit + cincl - nrow;
  const bool fits = lane < ncon && adr + nrow <= M.max_efc;
  const unsigned fitbal = __ballot_sync(DMB_FULL, fits);
  // contacts are dropped from the first one that does not fit (oracle: make_constraint)
  const unsigned wantbal = ncon >= 32 ? 0xffffffffu : ((1u << ncon) - 1u);
  if (fitbal != wantbal) {
    const int firstbad = __ffs(~fitbal & wantbal) - 1;
    ncon = firstbad;
    if (lane == 0) { S.flags |= 2; S.ncon = ncon; }
  }
  int nefc = nlimit;
  if (ncon > 0) nefc = __shfl_sync(DMB_FULL, adr + nrow, ncon - 1);
  // A stage with more than RF rows keeps Y (and e_src) outside the tile: in the CTA's one shared slot if it is free
  // (claimed until this stage's solve is done), else in the warp's global scratch slot (slow, very rare).
  int big = 0;
  if (nefc > RF) {
    if (lane == 0) big = (slot && atomicCAS(&s_slot_owner, -1, warp) == -1) ? 1 : 2;
    big = __shfl_sync(DMB_FULL, big, 0);
  }
  float* const G = big == 1 ? slot : Gglobal;
  signed char* e_src = big ? reinterpret_cast<signed char*>(G + gs::e_src) : S.e_src;
  {
    int r = incl - cnt;
    if (lo) e_src[r++] = (signed char)(-(2 * (1 + lane)));
    if (hi) e_src[r] = (signed char)(-(2 * (1 + lane) + 1));
  }
  if (lane < ncon) {
    S.c_meta[lane] = (S.c_meta[lane] & 0xffffff) | (adr << 24);
    for (int k = 0; k < nrow; k++) e_src[adr + k] = (signed char)(lane * 4 + k);
  }
  if (lane == 0) { S.nefc = nefc; S.nlimit = nlimit; S.big = big; }
  __syncwarp();
  return nefc;
}

template <bool OVF>
__device__ DMB_PHASE_FN void build_rows(const ModelS& M, EnvS& S, int lane, float* G, float* dbgrow) {
  const RowView<OVF> V(S, G);
  const int nefc = S.nefc, nlimit = S.nlimit, ncon = S.ncon;
  // ---- J rows into Y: lane = dof.  Y lies in front of cdof / the contact geometry, which are read here.
  const V3 com = ld3(S.com);
  for (int d = lane; d < M.nv; d += 32) {
    for (int r = 0; r < nlimit; r++) {
      const int src = V.e_src[r];
      V.Y[r * YS + d] = (M.jnt_dofadr[lim_joint(src)] == d) ? (lim_side(src) ? -1.f : 1.f) : 0.f;
    }
    const V3 ca = ld3(&S.o.c.cdof[6 * d]), cl = ld3(&S.o.c.cdof[6 * d + 3]);
    for (int c = 0; c < ncon; c++) {
      const int cm = S.c_meta[c];
      const int b1 = M.geom_bodyid[cm_g1(cm)], b2 = M.geom_bodyid[cm_g2(cm)];
      const int in2 = (int)((M.body_dofmask[b2] >> d) & 1ull), in1 = (int)((M.body_dofmask[b1] >> d) & 1ull);
      const float sg = (float)(in2 - in1);
      const int a = cm_adr(cm);
      const float* fr = &S.o.c.c_frame[6 * c];
      if (cm_dim(cm) == 1) {
        float jn = 0.f;
        if (sg != 0.f) { const V3 p = cl + cross(ca, ld3(&S.o.c.c_pos[3 * c]) - com); jn = sg * dot(ld3(fr), p); }
        V.Y[a * YS + d] = jn;
      } else {
        float jn = 0.f, j1 = 0.f, j2 = 0.f;
        if (sg != 0.f) {
          const V3 p = cl + cross(ca, ld3(&S.o.c.c_pos[3 * c]) - com);
          const V3 n = ld3(fr), t1 = ld3(fr + 3);
          jn = sg * dot(n, p); j1 = sg * dot(t1, p); j2 = sg * dot(cross(n, t1), p);
        }
        // base rows of the contact frame; half_solve_rows() solves these three together and then
        // expands them to the four pyramid edges  n +- mu t1, n +- mu t2
        V.Y[a * YS + d] = jn;
        V.Y[(a + 1) * YS + d] = j1;
        V.Y[(a + 2) * YS + d] = j2;
      }
    }
  }
  __syncwarp();   // cdof is dead from here on: the row scalars below overwrite it
  // ---- per-row impedance, R, aref: lane = row.  Values are computed for all rows of a round first, then stored
  // (the stores land on cdof, which no lane reads any more, but cvel / the contact geometry are still read)
  for (int r = lane; r < nefc; r += 32) {
    const int src = V.e_src[r];
    float dA, vel, mu = 0.f, pos, mg;
    unsigned long long mask;
    bool pyramid = false;
    if (src < 0) {
      const int j = lim_joint(src), dof = M.jnt_dofadr[j];
      const float q = S.qpos[M.jnt_qposadr[j]];
      pos = lim_side(src) ? M.jnt_range[j][1] - q : q - M.jnt_range[j][0];
      mg = 0.f;
      dA = M.dof_invw[dof];
      vel = lim_side(src) ? -S.qvel[dof] : S.qvel[dof];
      mask = M.dof_ancmask[dof] | (1ull << dof);
    } else {
      const int c = src >> 2, k = src & 3;
      const int cm = S.c_meta[c];
      const int b1 = M.geom_bodyid[cm_g1(cm)], b2 = M.geom_bodyid[cm_g2(cm)];
      mask = M.body_dofmask[b1] | M.body_dofmask[b2];
      pos = S.o.c.c_dist[c]; mg = M.margin;
      const float tran = M.body_invw[b1] + M.body_invw[b2];
      const V3 off = ld3(&S.o.c.c_pos[3 * c]) - com;
      const V3 v2 = ld3(&S.o.c.cvel[6 * b2 + 3]) + cross(ld3(&S.o.c.cvel[6 * b2]), off);
      const V3 v1 = ld3(&S.o.c.cvel[6 * b1 + 3]) + cross(ld3(&S.o.c.cvel[6 * b1]), off);
      const V3 vr = v2 - v1;
      const float* fr = &S.o.c.c_frame[6 * c];
      const V3 n = ld3(fr), t1 = ld3(fr + 3);
      const float vn = dot(n, vr);
      if (cm_dim(cm) == 1) { dA = tran; vel = vn; }
      else {
        pyramid = true;
        mu = fmaxf(M.geom_mu[cm_g1(cm)], M.geom_mu[cm_g2(cm)]);
        const float vt = dot((k >> 1) ? cross(n, t1) : t1, vr);
        vel = vn + ((k & 1) ? -mu : mu) * vt;
        dA = tran + mu * mu * tran;
      }
    }
    // getimpedance (solimp sigmoid)
    float dmin = clampf(M.solimp[0], 0.0001f, 0.9999f), dmax = clampf(M.solimp[1], 0.0001f, 0.9999f);
    const float width = M.solimp[2], mid = clampf(M.solimp[3], 0.0001f, 0.9999f), power = fmaxf(M.solimp[4], 1.f);
    float imp;
    if (dmin == dmax || width <= DMB_MINVAL) imp = 0.5f * (dmin + dmax);
    else {
      const float x = fabsf((pos - mg) / width);
      if (x >= 1.f) imp = dmax;
      else if (x == 0.f) imp = dmin;
      else {
        float y;
        if (power == 1.f) y = x;
        else if (power == 2.f) y = x <= mid ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
        else y = x <= mid ? powf(x, power) / powf(mid, power - 1.f) : 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
        imp = dmin + y * (dmax - dmin);
      }
    }
    float R = fmaxf((1.f - imp) * dA / imp, DMB_MINVAL);
    if (pyramid) R = 2.f * mu * mu * R;
    V.e_R[r] = R;
    V.e_aref[r] = -M.imp_b * vel - M.imp_k * imp * (pos - mg);
    V.rowmask[r] = mask;
    if (dbgrow) dbgrow[dbg::efc_pos + r] = pos;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Single-vector operations on the sparse factor M = L' D L with the vector held in REGISTERS
// (dof d on lane d & 31, register `lo` for d < 32 and `hi` for d >= 32).  One broadcast shuffle
// and one FFMA per dof: ~30 cycles of latency per dof instead of a shared-memory round trip.
// ------------------------------------------------------------------------------------------
// Dofs are numbered depth first, so "i is a descendant of a" is the range test a < i <= a + ndesc[a], and
// L[i][a] sits at qLD[Lend[i] - depth(a)]: no per-(i, a) table lookups.
// x <- L^-T x   (leaves -> root: dof i, largest id first, pushes its value to its ancestors)
__device__ __forceinline__ void reg_solve_LT(const ModelS& M, const EnvS& S, int lane, float& lo, float& hi) {
  const int nv = M.nv;
  const bool has_lo = lane < nv, has_hi = lane + 32 < nv;
  const unsigned nd_lo = has_lo ? (unsigned)M.dof_ndesc[lane] : 0u, nd_hi = has_hi ? (unsigned)M.dof_ndesc[lane + 32] : 0u;
  const float* b_lo = S.qLD - (has_lo ? M.dof_nanc[lane] : 0);
  const float* b_hi = S.qLD - (has_hi ? M.dof_nanc[lane + 32] : 0);
  for (int i = nv - 1; i >= 32; i--) {
    const int e = M.dof_Lend[i];
    const float xi = __shfl_sync(DMB_FULL, hi, i - 32);
    if ((unsigned)(i - lane - 1) < nd_lo) lo = fmaf(-b_lo[e], xi, lo);
    if ((unsigned)(i - lane - 33) < nd_hi) hi = fmaf(-b_hi[e], xi, hi);
  }
#pragma unroll 4
  for (int i = (nv < 32 ? nv : 32) - 1; i >= 1; i--) {
    const int e = M.dof_Lend[i];
    const float xi = __shfl_sync(DMB_FULL, lo, i);
    if ((unsigned)(i - lane - 1) < nd_lo) lo = fmaf(-b_lo[e], xi, lo);
  }
}
// x <- L^-1 x   level-parallel: in round r every dof with more than r ancestors pulls from its ancestor at
// depth r, whose value has been final since round r-1 (12 rounds instead of nv-1 steps; same summation
// order as the serial loop).  Dofs >= 32 (two on the humanoid) are finished by warp reductions.
__device__ __forceinline__ void reg_solve_L(const ModelS& M, const EnvS& S, int lane, float& lo, float& hi) {
  const int nv = M.nv;
  const int dl = lane < nv ? lane : 0;
  const int nd = lane < nv ? M.dof_nanc[dl] : 0;
  const float* Lp = &S.qLD[M.dof_Lend[dl]];
  const uint32_t* ap = reinterpret_cast<const uint32_t*>(M.dof_ancr[dl]);
  static_assert(MAXANC == 12, "three packed words of ancestor ids");
  const uint32_t aw[3] = {ap[0], ap[1], ap[2]};
#pragma unroll
  for (int r = 0; r < MAXANC; r++) {
    const int src = (int)((aw[r >> 2] >> (8 * (r & 3))) & 0xffu);
    const float xs = __shfl_sync(DMB_FULL, lo, src);
    if (r < nd) lo = fmaf(-Lp[-r], xs, lo);
  }
  for (int d = 32; d < nv; d++) {
    const int c = M.dof_nanc[d];
    const float* Ld = &S.qLD[M.dof_Lend[d]];
    const int a = lane < c ? M.dof_ancr[d][lane] : 0;
    const float xl = __shfl_sync(DMB_FULL, lo, a & 31), xh = __shfl_sync(DMB_FULL, hi, a & 31);
    const float term = lane < c ? Ld[-lane] * (a >= 32 ? xh : xl) : 0.f;
    const float sum = warp_sum(term);
    if (lane == d - 32) hi -= sum;
  }
}
// z <- D^1/2 L x  (image of an acceleration in the half-solved space), smem in / smem out
__device__ void mul_L_sqrtD(const ModelS& M, EnvS& S, int lane, const float* x, float* z) {
  for (int i = lane; i < M.nv; i += 32) {
    const int c = M.dof_nanc[i], adr = M.dof_Madr[i] + 1;
    float s = x[i];
    for (int k = 0; k < c; k++) s += S.qLD[adr + k] * x[M.dof_anc[i][k]];
    z[i] = s / S.dsq[i];
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Half solve  Y_r <- D^-1/2 L^-T J_r'  (lane = dof, rows in registers).  S.rowmask[r] holds the
// ancestor-closed dof support of row r (chains of the bodies in contact / of the limited joint);
// only those dofs are visited, deepest first: one broadcast shuffle + one FFMA per dof and row.
// A pyramidal contact arrives as its three frame rows (n, t1, t2), which share one support: they
// are eliminated together (factor entries and index work shared) and then expanded to the four
// edge rows  n +- mu t1,  n +- mu t2  (the elimination is linear).
// ------------------------------------------------------------------------------------------
template <bool OVF>
__device__ DMB_PHASE_FN void half_solve_rows(const ModelS& M, EnvS& S, int lane, int r_begin, int r_end, float* G) {
  const RowView<OVF> V(S, G);
  const bool has_lo = lane < M.nv, has_hi = lane + 32 < M.nv;
  const float dlo = has_lo ? S.dsq[lane] : 0.f, dhi = has_hi ? S.dsq[lane + 32] : 0.f;
  const unsigned nd_lo = has_lo ? (unsigned)M.dof_ndesc[lane] : 0u, nd_hi = has_hi ? (unsigned)M.dof_ndesc[lane + 32] : 0u;
  const float* Lb_lo = S.qLD - (has_lo ? M.dof_nanc[lane] : 0);
  const float* Lb_hi = S.qLD - (has_hi ? M.dof_nanc[lane + 32] : 0);
  const int16_t* Lend = M.dof_Lend;
  int r = r_begin;   // r_begin must be the first row of a group (a limit row, a frictionless row or a pyramid)
  while (r < r_end) {
    const int src = V.e_src[r];
    const bool pyr = src >= 0 && cm_dim(S.c_meta[src >> 2]) == 3;
    float* y = &V.Y[r * YS];
    float a_lo = has_lo ? y[lane] : 0.f, a_hi = has_hi ? y[lane + 32] : 0.f;
    float b_lo = 0.f, b_hi = 0.f, c_lo = 0.f, c_hi = 0.f;
    if (pyr) {
      if (has_lo) { b_lo = y[YS + lane]; c_lo = y[2 * YS + lane]; }
      if (has_hi) { b_hi = y[YS + lane + 32]; c_hi = y[2 * YS + lane + 32]; }
    }
    // support dofs, deepest first: ids >= 32 (they can feed both halves), then ids < 32 (only the low half)
    const unsigned long long sup64 = V.rowmask[r];
    unsigned sup_hi = (unsigned)(sup64 >> 32), sup = (unsigned)sup64 & ~1u;   // dof 0 has no ancestors
    while (sup_hi) {
      const int ih = 31 - __clz(sup_hi);
      sup_hi &= ~(1u << ih);
      const int i = ih + 32, e = M.dof_Lend[i];
      const float Lv = (unsigned)(i - lane - 1) < nd_lo ? Lb_lo[e] : 0.f;
      const float Lh = (unsigned)(i - lane - 33) < nd_hi ? Lb_hi[e] : 0.f;
      const float xa = __shfl_sync(DMB_FULL, a_hi, ih);
      a_lo = fmaf(-Lv, xa, a_lo); a_hi = fmaf(-Lh, xa, a_hi);
      if (pyr) {
        const float xb = __shfl_sync(DMB_FULL, b_hi, ih), xc = __shfl_sync(DMB_FULL, c_hi, ih);
        b_lo = fmaf(-Lv, xb, b_lo); c_lo = fmaf(-Lv, xc, c_lo);
        b_hi = fmaf(-Lh, xb, b_hi); c_hi = fmaf(-Lh, xc, c_hi);
      }
    }
    // (the factor entry is loaded unconditionally -- a non-descendant lane reads some other word of its tile --
    // and masked afterwards: no divergent branch around the load; bfind gives the top support bit directly)
    if (pyr) {
      while (sup) {
        int i;
        asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(sup));
        sup ^= 1u << i;
        const float Lraw = Lb_lo[Lend[i]];
        const float Lv = (unsigned)(i - lane - 1) < nd_lo ? Lraw : 0.f;
        const float xa = __shfl_sync(DMB_FULL, a_lo, i), xb = __shfl_sync(DMB_FULL, b_lo, i), xc = __shfl_sync(DMB_FULL, c_lo, i);
        a_lo = fmaf(-Lv, xa, a_lo); b_lo = fmaf(-Lv, xb, b_lo); c_lo = fmaf(-Lv, xc, c_lo);
      }
    } else {
      while (sup) {
        int i;
        asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(sup));
        sup ^= 1u << i;
        const float Lraw = Lb_lo[Lend[i]];
        const float Lv = (unsigned)(i - lane - 1) < nd_lo ? Lraw : 0.f;
        const float xa = __shfl_sync(DMB_FULL, a_lo, i);
        a_lo = fmaf(-Lv, xa, a_lo);
      }
    }
    a_lo *= dlo; a_hi *= dhi;
    if (pyr) {
      const int cmm = S.c_meta[src >> 2];
      const float mu = fmaxf(M.geom_mu[cm_g1(cmm)], M.geom_mu[cm_g2(cmm)]);
      b_lo *= mu * dlo; b_hi *= mu * dhi; c_lo *= mu * dlo; c_hi *= mu * dhi;
      if (has_lo) { y[lane] = a_lo + b_lo; y[YS + lane] = a_lo - b_lo; y[2 * YS + lane] = a_lo + c_lo; y[3 * YS + lane] = a_lo - c_lo; }
      if (has_hi) {
        y[lane + 32] = a_hi + b_hi; y[YS + lane + 32] = a_hi - b_hi;
        y[2 * YS + lane + 32] = a_hi + c_hi; y[3 * YS + lane + 32] = a_hi - c_hi;
      }
      r += 4;
    } else {
      if (has_lo) y[lane] = a_lo;
      if (has_hi) y[lane + 32] = a_hi;
      r += 1;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }

// Gram matrix AR = Y Y' + diag(R) (packed lower triangle) and b = Y y_s - aref.
// lane = matrix entry: the nefc(nefc+1)/2 pairs (+ nefc entries for b) are dealt out 32 at a time;
// each lane runs the sparse dot product over the intersection of the two row supports.
template <bool OVF>
__device__ DMB_PHASE_FN void gram(const ModelS& M, EnvS& S, int lane, int nefc, int t_begin, int t_end, float* G) {
  const RowView<OVF> V(S, G);
  const int npair = tri(nefc), ntask = min(npair + nefc, t_end);
  for (int t = t_begin + lane; t < ntask; t += 32) {
    // one loop for both kinds of task (a lane with a matrix entry and a lane with an entry of b would otherwise
    // run two loops one after the other): dot product of row r with row c, or with y_s
    int r, c = 0;
    const float* z = S.ys;
    unsigned long long mk;
    if (t < npair) {
      r = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
      if (tri(r + 1) <= t) r++;
      if (tri(r) > t) r--;
      c = t - tri(r);
      z = &V.Y[c * YS];
      mk = V.rowmask[r] & V.rowmask[c];
    } else {
      r = t - npair;
      mk = V.rowmask[r];
    }
    const float* yr = &V.Y[r * YS];
    float acc = 0.f;
    unsigned lo = (unsigned)mk, hi = (unsigned)(mk >> 32);
    while (hi) {  // dofs >= 32
      const int k = 31 - __clz(hi);
      hi ^= 1u << k;
      acc = fmaf(yr[32 + k], z[32 + k], acc);
    }
    while (lo) {
      const int k = 31 - __clz(lo);
      lo ^= 1u << k;
      acc = fmaf(yr[k], z[k], acc);
    }
    if (t < npair) V.AR[t] = (c == r) ? acc + V.e_R[r] : acc;
    else V.e_b[r] = acc - V.e_aref[r];
  }
  __syncwarp();
}

// PGS sweeps over rows [0, nefc) with AR in memory (only used by stages with more than RF rows, whose rows live in
// the global scratch slot): residuals res = AR f + b live in registers (lane = row, and row + 32 when HI); a row
// update broadcasts its force increment and every lane applies one column of AR.  All rows are scalar with
// force >= 0 (limits, frictionless normals, pyramid edges).
// (inlined into its two call sites in solve_constraints<true>: as an out-of-line function its four reference
// parameters lived in local memory and every row update paid several L1/L2 round trips -- 125 cycles per row)
template <bool HI>
__device__ __forceinline__ int pgs_sweeps(const ModelS& M, const float* AR, int lane, int nefc, float& f0, float& f1,
                                          float& res0, float& res1) {
  const int r0 = lane, r1 = lane + 32;
  const bool a0 = r0 < nefc, a1 = HI && r1 < nefc;
  const float d0 = a0 ? AR[tri(r0) + r0] : 1.f, d1 = a1 ? AR[tri(r1) + r1] : 1.f;
  const float inv0 = 1.0f / d0, inv1 = 1.0f / d1;
  const int t0 = tri(r0), t1 = tri(r1);
  const int nlo = HI ? 32 : nefc;
  int iter = 0;
  while (iter < M.iterations) {
    // the owner of a row remembers its increment and the residual it was computed from; the
    // cost decrease  -(0.5 delta^2 AR_ii + delta res_i)  is summed once per sweep
    float dm0 = 0.f, rm0 = 0.f, dm1 = 0.f, rm1 = 0.f;
    // Column i of the packed triangle for lane-row r is AR[tri(r) + i] (i <= r) or AR[tri(i) + r] (i > r): the
    // index moves by 1 while i < r and by i + 1 afterwards.  Loads are unpredicated (indices past the last row
    // are clamped into the triangle; lanes without a row never own an update).
    int idx0 = t0, idx1 = t1;
    float acol0 = AR[idx0 < NTRI ? idx0 : 0];
    float acol1 = HI ? AR[a1 ? idx1 : 0] : 0.f;
#pragma unroll 2
    for (int i = 0; i < nlo; i++) {
      idx0 += (i < r0) ? 1 : i + 1;
      const float an0 = AR[idx0 < NTRI ? idx0 : 0];
      float an1 = 0.f;
      if (HI) { idx1 += (i < r1) ? 1 : i + 1; an1 = AR[(a1 && idx1 < NTRI) ? idx1 : 0]; }
      const float fnew = fmaxf(0.f, f0 - res0 * inv0);
      const float mine = fnew - f0;
      const float delta = __shfl_sync(DMB_FULL, mine, i);
      if (lane == i) { dm0 = mine; rm0 = res0; f0 = fnew; }
      res0 = fmaf(acol0, delta, res0);
      if (HI) res1 = fmaf(acol1, delta, res1);
      acol0 = an0; acol1 = an1;
    }
    if (HI) {
      for (int i = 32; i < nefc; i++) {
        idx0 += i + 1;                                          // i >= 32 > r0
        idx1 += (i < r1) ? 1 : i + 1;
        const float an0 = AR[idx0 < NTRI ? idx0 : 0];
        const float an1 = AR[(a1 && idx1 < NTRI) ? idx1 : 0];
        const float fnew = fmaxf(0.f, f1 - res1 * inv1);
        const float mine = fnew - f1;
        const float delta = __shfl_sync(DMB_FULL, mine, i - 32);
        if (lane == i - 32) { dm1 = mine; rm1 = res1; f1 = fnew; }
        res0 = fmaf(acol0, delta, res0);
        res1 = fmaf(acol1, delta, res1);
        acol0 = an0; acol1 = an1;
      }
    }
    iter++;
    float imp = -(dm0 * (0.5f * dm0 * d0 + rm0));
    if (HI) imp -= dm1 * (0.5f * dm1 * d1 + rm1);
    imp = warp_sum(imp) * M.pgs_scale;
    if (imp < M.tolerance) break;
  }
  return iter;
}

// nefc <= NR (NR = RF = 24 in the tile path, 32 for big stages): this lane's column of AR (= its row, AR is
// symmetric) is held in RF registers, loaded once per solve; a sweep is then a fully unrolled chain of
// row updates with no shared-memory access and no index arithmetic.  Rows are kept scaled by -1/AR_ii
// (sres = -res / AR_ii, scaled column acol * -1/AR_ii), so the increment is max(-f, sres) and the serial chain per
// row is FMNMX -> SHFL -> FFMA.  The column is fetched in chunks of 8 rows behind a uniform branch (most envs have
// fewer than 8 rows); lanes without a row keep a zero column.
template <int NR>
__device__ __forceinline__ int pgs_sweeps_reg(const ModelS& M, const float* AR, int lane, int nefc, float& f0, float& res0) {
  static_assert(NR % 8 == 0 && NR <= 32, "column chunks of 8 rows");
  const int r0 = lane, t0 = tri(r0);
  const bool a0 = r0 < nefc;
  const float d0 = a0 ? AR[t0 + r0] : 1.f;
  const float ninv0 = -rcp(d0);
  float acol[NR];
#pragma unroll
  for (int i = 0; i < NR; i++) acol[i] = 0.f;
#pragma unroll
  for (int c = 0; c < NR / 8; c++) {
    if (8 * c < nefc) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int i = 8 * c + j;
        if (a0 && i < nefc) acol[i] = AR[i <= r0 ? t0 + i : tri(i) + r0] * ninv0;
      }
    }
  }
  float sres = res0 * ninv0;
  int iter = 0;
  while (iter < M.iterations) {
    // The cost decrease of a whole sweep is  -0.5 (f_end - f_start)' (res_end + res_start)  (res = AR f + b,
    // AR symmetric): exact like the per-row sum MuJoCo accumulates, without any per-row bookkeeping.
    const float fs = f0, ss = sres;
#pragma unroll
    for (int i = 0; i < NR; i += 2) {
      if (i >= nefc) break;   // rows are taken in pairs; a row past nefc has no owner and a zero column
      {
        const float mine = fmaxf(-f0, sres);
        const float delta = __shfl_sync(DMB_FULL, mine, i);
        if (lane == i) f0 += mine;
        sres = fmaf(acol[i], delta, sres);
      }
      {
        const float mine = fmaxf(-f0, sres);
        const float delta = __shfl_sync(DMB_FULL, mine, i + 1);
        if (lane == i + 1) f0 += mine;
        sres = fmaf(acol[i + 1], delta, sres);
      }
    }
    iter++;
    const float imp = warp_sum(0.5f * d0 * (f0 - fs) * (sres + ss)) * M.pgs_scale;
    if (imp < M.tolerance) break;
  }
  res0 = -sres * d0;
  return iter;
}

// ------------------------------------------------------------------------------------------
// mj_fwdConstraint: warmstart + PGS on the dual, then qacc = L^-1 D^-1/2 (y_s + Y' f).
// S.qacc holds qacc_warmstart on entry and the new acceleration on exit.
// ------------------------------------------------------------------------------------------
template <bool OVF>
__device__ DMB_PHASE_FN void solve_constraints(const ModelS& M, EnvS& S, int lane, int nefc, float* G) {
  const RowView<OVF> V(S, G);
  int iter = 0;
  if (nefc > 0) {
    const int r0 = lane, r1 = lane + 32;
    const bool a0 = r0 < nefc, a1 = OVF && r1 < nefc;
    // warmstart forces from qacc_warmstart: jar = J qacc_w - aref = Y (D^1/2 L qacc_w) - aref
    float* w = S.x_dv;   // scratch (x_dv is dead inside a forward evaluation)
    mul_L_sqrtD(M, S, lane, S.qacc, w);
    float jar0 = 0.f, jar1 = 0.f;
    if (a0) { const float* y = &V.Y[r0 * YS]; for (int k = 0; k < M.nv; k++) jar0 += y[k] * w[k]; jar0 -= V.e_aref[r0]; }
    if (a1) { const float* y = &V.Y[r1 * YS]; for (int k = 0; k < M.nv; k++) jar1 += y[k] * w[k]; jar1 -= V.e_aref[r1]; }
    float f0 = (a0 && jar0 < 0.f) ? -jar0 / V.e_R[r0] : 0.f;
    float f1 = (a1 && jar1 < 0.f) ? -jar1 / V.e_R[r1] : 0.f;
    if (a0) V.e_f[r0] = f0;
    if (a1) V.e_f[r1] = f1;
    __syncwarp();
    // res = AR f + b ; dual cost = sum f (0.5 (res - b) + b); a positive cost falls back to f = 0
    const float b0 = a0 ? V.e_b[r0] : 0.f, b1 = a1 ? V.e_b[r1] : 0.f;
    float res0 = b0, res1 = b1;
    for (int s = 0; s < nefc; s++) {
      const float fs = V.e_f[s];
      if (fs != 0.f) {
        if (a0) res0 += V.AR[r0 >= s ? tri(r0) + s : tri(s) + r0] * fs;
        if (a1) res1 += V.AR[r1 >= s ? tri(r1) + s : tri(s) + r1] * fs;
      }
    }
    float cost = f0 * 0.5f * (res0 + b0) + f1 * 0.5f * (res1 + b1);
    cost = warp_sum(cost);
    if (cost > 0.f) { f0 = 0.f; f1 = 0.f; res0 = b0; res1 = b1; }
    DMB_TICK(14);
    // up to 32 rows: one row per lane, its column of AR in registers; more: a second row on lanes 0..7, AR in the tile.
    // (Tried: two consecutive rows per lane with a local pair update and AR entries read per step -- one shuffle round
    // trip per two rows.  The 4 x index arithmetic + loads per pair step made it 45 instructions per pair against 16,
    // and a lone warp is issue-limited too: 7.63 M vs 8.47 M env-steps/s when used from 17 rows up.)
    if (OVF) iter = nefc > 32 ? pgs_sweeps<true>(M, V.AR, lane, nefc, f0, f1, res0, res1)
                              : pgs_sweeps_reg<32>(M, V.AR, lane, nefc, f0, res0);
    else iter = pgs_sweeps_reg<RF>(M, V.AR, lane, nefc, f0, res0);
    DMB_TICK(15);
    __syncwarp();
    if (a0) V.e_f[r0] = f0;
    if (a1) V.e_f[r1] = f1;
    __syncwarp();
  }
  if (lane == 0) {
    S.iter = iter;
    S.diag = ((S.diag & 0xffff) + iter) | (max(S.diag >> 16, nefc) << 16);
    // scheduler key (see k_order): modes 0-2 accumulate over the RK stages, 3-4 keep the last stage only
    if (M.cost_mode == 3) S.cost = 4 * nefc * iter;
    else if (M.cost_mode == 4) S.cost = 4 * nefc * (8 + iter);
    else S.cost += M.cost_mode == 1 ? nefc * 16 : (M.cost_mode == 2 ? nefc * (16 + iter) : nefc * iter);
  }
  // t = y_s + sum_r Y_r f_r  (lane = dof), then qacc = L^-1 D^-1/2 t in registers
  float tlo = lane < M.nv ? S.ys[lane] : 0.f, thi = lane + 32 < M.nv ? S.ys[lane + 32] : 0.f;
  for (int r = 0; r < nefc; r++) {
    const float fr = V.e_f[r];
    if (fr != 0.f) {
      if (lane < M.nv) tlo += V.Y[r * YS + lane] * fr;
      if (lane + 32 < M.nv) thi += V.Y[r * YS + lane + 32] * fr;
    }
  }
  if (lane < M.nv) tlo *= S.dsq[lane];
  if (lane + 32 < M.nv) thi *= S.dsq[lane + 32];
  reg_solve_L(M, S, lane, tlo, thi);
  if (lane < M.nv) S.qacc[lane] = tlo;
  if (lane + 32 < M.nv) S.qacc[lane + 32] = thi;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// One full forward evaluation (mj_forward) at (S.qpos, S.qvel, S.ctrlf; S.qacc = qacc_warmstart) -> S.qacc.
// Returns the whole-body CoM height of this evaluation (what DPEnv.is_done reads from the
// stale mjData.xipos after mj_step, dp_env_v3.py:134-139).
// ------------------------------------------------------------------------------------------
// LOCKSTEP: all warps of the CTA walk the phases together (one __syncthreads per phase), so the
// ~100 KB instruction stream of a stage is fetched once per CTA instead of once per warp --
// the v2 profile showed 60% of warp stalls were instruction-fetch (stall_no_inst).  Warps with
// no env (`active` false) only take part in the barriers.
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <bool LOCKSTEP>
__device__ DMB_PHASE_FN float forward_eval(const ModelS& M, EnvS& S, int lane, float* dbgrow, bool active, int bar_id,
                                         int bar_n, EnvS* tiles, int* share_cnt, float* gcta, int warp,
                                         bool use_slot) {
  // the CTA's shared slot lies behind its last tile (k_step); only the single-group lockstep kernel has one
  float* const slot = use_slot ? reinterpret_cast<float*>(tiles + (bar_n >> 5)) : nullptr;
#define DMB_PHASE_SYNC(bit) do { if (LOCKSTEP && (M.sync_mask & (bit))) group_barrier(bar_id, bar_n); } while (0)
  float* G = gcta + (size_t)warp * gs::stride;   // where Y lives in a stage with > RF rows (set by count_rows)
  DMB_TICK(0);
  DMB_PHASE_SYNC(1);
  DMB_TICK(1);
  if (active) {
    kinematics(M, S, lane);
    DMB_TICK(16);
    com_pos(M, S, lane);
    if (dbgrow) {
      for (int i = lane; i < M.nbody * 3; i += 32) { dbgrow[dbg::xpos + i] = S.o.k.xpos[i]; dbgrow[dbg::xipos + i] = S.o.k.xipos[i]; }
      for (int i = lane; i < M.nbody * 4; i += 32) dbgrow[dbg::xquat + i] = S.o.k.xquat[i];
    }
  }
  DMB_TICK(2);
  DMB_PHASE_SYNC(2);
  if (active) crb_factor(M, S, lane, dbgrow ? dbgrow + dbg::qM : nullptr);
  DMB_TICK(3);
  DMB_PHASE_SYNC(4);
  if (active) {
    smooth_forces(M, S, lane, dbgrow);
    DMB_TICK(17);
    // y_s = D^-1/2 L^-T qfrc_smooth (registers)
    float lo = lane < M.nv ? S.x_dv[lane] : 0.f, hi = lane + 32 < M.nv ? S.x_dv[lane + 32] : 0.f;
    reg_solve_LT(M, S, lane, lo, hi);
    if (lane < M.nv) S.ys[lane] = lo * S.dsq[lane];
    if (lane + 32 < M.nv) S.ys[lane + 32] = hi * S.dsq[lane + 32];
  }
  DMB_TICK(4);
  DMB_PHASE_SYNC(8);
  if (active) {
    geom_poses(M, S, lane);
    collision(M, S, lane);
  }
  DMB_TICK(5);
  DMB_PHASE_SYNC(16);
  int nefc = 0;
  if (active) {
    nefc = count_rows(M, S, lane, G, slot, warp);
    if (S.big == 1) G = slot;
    if (dbgrow) {   // contact geometry and cvel are overwritten by the row data below
      for (int i = lane; i < M.nbody * 6; i += 32) dbgrow[dbg::cvel + i] = S.o.c.cvel[i];
      for (int c = lane; c < S.ncon; c += 32) {
        float* cr = dbgrow + dbg::contact + 16 * c;
        const float* fr = &S.o.c.c_frame[6 * c];
        const V3 t2 = cross(ld3(fr), ld3(fr + 3));
        cr[0] = S.o.c.c_dist[c];
        for (int k = 0; k < 3; k++) cr[1 + k] = S.o.c.c_pos[3 * c + k];
        for (int k = 0; k < 6; k++) cr[4 + k] = fr[k];
        cr[10] = t2.x; cr[11] = t2.y; cr[12] = t2.z;
        cr[13] = (float)cm_g1(S.c_meta[c]); cr[14] = (float)cm_g2(S.c_meta[c]); cr[15] = (float)cm_dim(S.c_meta[c]);
      }
      __syncwarp();
    }
    if (nefc > RF) build_rows<true>(M, S, lane, G, dbgrow); else build_rows<false>(M, S, lane, G, dbgrow);
  }
  DMB_TICK(6);
#if DMB_SHARE
  if (LOCKSTEP && share_cnt) {
    // CTA-wide work sharing of the two phases whose cost scales with the row count: every tile of the
    // CTA is in shared memory, so any warp can half-solve a row group or compute a chunk of Gram entries of any env.
    // Tasks are numbered tile by tile and handed out by an atomic counter; warps without an env help too.
    const int W = bar_n >> 5;
    if (!active && lane == 0) { S.nefc = 0; S.nlimit = 0; S.ncon = 0; }
    group_barrier(bar_id, bar_n);                                   // A: the rows of every tile are assembled
    DMB_TICK(18);
    {
      int ng = 0, nl = 0;
      if (lane < W) { nl = tiles[lane].nlimit; ng = tiles[lane].nefc > 0 ? nl + tiles[lane].ncon : 0; }
      const int incl = warp_incl_scan(ng, lane);
      const int total = __shfl_sync(DMB_FULL, incl, 31);
      for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&share_cnt[0], 1);
        t = __shfl_sync(DMB_FULL, t, 0);
        if (t >= total) break;
        const int w = __popc(__ballot_sync(DMB_FULL, incl <= t));    // tile that owns task t
        const int g = t - (w > 0 ? __shfl_sync(DMB_FULL, incl, w - 1) : 0);
        const int nlw = __shfl_sync(DMB_FULL, nl, w);
        EnvS& T = tiles[w];
        const int row = g < nlw ? g : cm_adr(T.c_meta[g - nlw]);
        if (T.nefc > RF) half_solve_rows<true>(M, T, lane, row, row + 1, T.big == 1 ? slot : gcta + (size_t)w * gs::stride);
        else half_solve_rows<false>(M, T, lane, row, row + 1, nullptr);
      }
    }
    DMB_TICK(19);
    group_barrier(bar_id, bar_n);                                   // B: every Y row is half-solved
    DMB_TICK(20);
    if ((warp | lane) == 0) share_cnt[0] = 0;
    {
      int nt = 0, ne = 0;
      if (lane < W) { ne = tiles[lane].nefc; nt = (tri(ne) + ne + 31) >> 5; }
      const int incl = warp_incl_scan(nt, lane);
      const int total = __shfl_sync(DMB_FULL, incl, 31);
      for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&share_cnt[1], 1);
        t = __shfl_sync(DMB_FULL, t, 0);
        if (t >= total) break;
        const int w = __popc(__ballot_sync(DMB_FULL, incl <= t));
        const int c = t - (w > 0 ? __shfl_sync(DMB_FULL, incl, w - 1) : 0);
        const int new_ = __shfl_sync(DMB_FULL, ne, w);
        if (new_ > RF) gram<true>(M, tiles[w], lane, new_, 32 * c, 32 * c + 32, tiles[w].big == 1 ? slot : gcta + (size_t)w * gs::stride);
        else gram<false>(M, tiles[w], lane, new_, 32 * c, 32 * c + 32, nullptr);
      }
    }
    DMB_TICK(21);
    group_barrier(bar_id, bar_n);                                   // C: every AR / b entry is in place
    if ((warp | lane) == 0) share_cnt[1] = 0;
  } else
#endif
  {
    DMB_PHASE_SYNC(32);
    if (active && nefc > 0) {
      if (nefc > RF) {
        half_solve_rows<true>(M, S, lane, 0, nefc, G);
        gram<true>(M, S, lane, nefc, 0, 1 << 30, G);
      } else {
        half_solve_rows<false>(M, S, lane, 0, nefc, nullptr);
        DMB_TICK(7);
        gram<false>(M, S, lane, nefc, 0, 1 << 30, nullptr);
      }
    }
  }
  DMB_TICK(8);
  if (active && dbgrow) {
    // qacc_smooth = L^-1 D^-1/2 y_s for the dump (not needed by the solver)
    float lo = lane < M.nv ? S.ys[lane] * S.dsq[lane] : 0.f, hi = lane + 32 < M.nv ? S.ys[lane + 32] * S.dsq[lane + 32] : 0.f;
    reg_solve_L(M, S, lane, lo, hi);
    if (lane < M.nv) dbgrow[dbg::qacc_smooth + lane] = lo;
    if (lane + 32 < M.nv) dbgrow[dbg::qacc_smooth + lane + 32] = hi;
    for (int e = lane; e < M.nM; e += 32) dbgrow[dbg::qLD + e] = S.qLD[e];
    const RowView<true> Vg(S, G);
    const RowView<false> Vt(S, G);
    for (int r = lane; r < nefc; r += 32) {
      const bool ov = nefc > RF;
      dbgrow[dbg::efc_R + r] = ov ? Vg.e_R[r] : Vt.e_R[r];
      dbgrow[dbg::efc_aref + r] = ov ? Vg.e_aref[r] : Vt.e_aref[r]; dbgrow[dbg::efc_b + r] = ov ? Vg.e_b[r] : Vt.e_b[r];
      dbgrow[dbg::efc_AR_diag + r] = ov ? Vg.AR[tri(r) + r] : Vt.AR[tri(r) + r];
    }
    __syncwarp();
  }
#if DMB_SHARE
  if (!(LOCKSTEP && share_cnt))
#endif
  DMB_PHASE_SYNC(64);
  DMB_TICK(9);
  if (active) {
    if (nefc > RF) solve_constraints<true>(M, S, lane, nefc, G); else solve_constraints<false>(M, S, lane, nefc, G);
    if (nefc > RF && S.big == 1 && !dbgrow) {   // hand the CTA's shared slot back
      __syncwarp();
      if (lane == 0) atomicExch(&s_slot_owner, -1);
    }
    if (dbgrow) {
      const RowView<true> Vg(S, G);
      const RowView<false> Vt(S, G);
      for (int r = lane; r < nefc; r += 32) dbgrow[dbg::efc_force + r] = nefc > RF ? Vg.e_f[r] : Vt.e_f[r];
    }
  }
  DMB_TICK(10);
#undef DMB_PHASE_SYNC
  return active ? S.com[2] : 0.f;
}

}  // namespace dmb
