// dmb.cu -- kernels + C-ABI of libdmb200.so (see include/dmb.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC -prec-div=false -prec-sqrt=false
//
// Kernel design (DESIGN.md has the full account):
//   * one warp per env, W = 28 envs per CTA (896 threads, 72 registers), persistent grid (<= #SMs CTAs): 148 x 28
//     warps hold a 4096-env batch at once.  With more envs than resident warps CTAs pull W envs at a time from a
//     queue sorted by last step's constraint work (k_order); with one round the sorted list is dealt round robin.
//     All warps of a CTA walk the RK stages in lockstep (one instruction stream per SM);
//   * the fp32 model tables are staged once per CTA in shared memory (one TMA bulk copy); each warp owns a 7.8 KB
//     EnvS tile whose overlay is time-shared by the phases of a forward evaluation (kinematics / inertia / RNE
//     temporaries, then geom poses + contact geometry, then the half-solved constraint Jacobian Y and the packed
//     Delassus matrix AR) -- per-env state crosses HBM exactly once per step in each direction (one 128-bit
//     load / store per lane for the three state rows);
//   * the whole RK4 step (4 forward evaluations incl. collision and PGS), mocap lookup /
//     interpolation, reward, termination, auto-reset and the observation (56-d or the 197-d
//     DeepMimic state) are fused in this one kernel.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "dmb_device.cuh"

namespace dmb {

struct DevPtrs {
  int* counter;            // dynamic env scheduler (zeroed by k_order before every step)
  int* cost;               // [N] constraint-work estimate of each env's last step (scheduler key)
  const int* order;        // [N] env ids sorted by decreasing cost
  const ModelS* model;
  const float* mocap_cfg;  // [F][nq]
  const float* mocap_vel;  // [F][nv]
  const float* ref_aux;    // [F][DMB_REF_AUX]
  long long* trace;        // DMB_TRACE=1: [grid][8] globaltimer at kernel start and after each scheduler round
  float* gscratch;         // [grid * W][gs::stride] row storage of stages with more than RF constraint rows
  // fused all-gather (dmb_set_peer_gather): every env's record row is also stored into the gathered [N_global][od+2]
  // buffer of each of the n_peer ranks (peer memory over NVLink, self included) at row row0 + env; the last CTA to
  // finish bumps every peer's arrival flag
  float* peer_rec[DMB_MAX_PEER];
  int* peer_flag[DMB_MAX_PEER];
  int n_peer, row0;
  int* ticket;             // CTAs done in this launch (last one signals the peers and rewinds it)
  const int* wait_flag;    // dmb_set_peer_wait: this launch does not complete before *wait_flag >= wait_target
  int wait_target;
};

// ---------------------------------------------------------------------------------------
// state I/O helpers (lane-coalesced env-major rows)
// ---------------------------------------------------------------------------------------
// The three state rows of an env (qpos / qvel / qacc_warmstart, DMB_QSTRIDE = DMB_VSTRIDE = 36 floats = nine
// 16-byte words each) move with ONE 128-bit load / store per lane: lanes 0-8 qpos, 9-17 qvel, 18-26 warmstart.
static_assert(DMB_QSTRIDE == NQC && DMB_VSTRIDE == NVC && NQC % 4 == 0 && NVC % 4 == 0, "state rows are whole float4s");
static_assert(offsetof(EnvS, qvel) % 16 == 0 && offsetof(EnvS, qacc) % 16 == 0 && sizeof(EnvS) % 16 == 0, "float4 alignment");
__device__ __forceinline__ void load_state(const ModelS& M, EnvS& S, const dmb_state_t& st, int env, int lane) {
  constexpr int NW = NQC / 4;   // float4 words per row
  if (lane < 3 * NW) {
    const int a = lane / NW, w = lane - a * NW;
    const float* src = a == 0 ? st.qpos : (a == 1 ? st.qvel : st.warm);
    float* dst = a == 0 ? S.qpos : (a == 1 ? S.qvel : S.qacc);
    reinterpret_cast<float4*>(dst)[w] = __ldg(reinterpret_cast<const float4*>(src + (size_t)env * NQC) + w);
  }
  if (lane == 0) { S.flags = 0; S.cost = 0; S.diag = 0; S.big = 0; }
  __syncwarp();
}
__device__ __forceinline__ void store_state(const ModelS& M, EnvS& S, const dmb_state_t& st, int env, int lane) {
  constexpr int NW = NQC / 4;
  if (lane < 3 * NW) {
    const int a = lane / NW, w = lane - a * NW;
    float* dst = a == 0 ? st.qpos : (a == 1 ? st.qvel : st.warm);
    const float* src = a == 0 ? S.qpos : (a == 1 ? S.qvel : S.qacc);
    const int n = a == 0 ? M.nq : M.nv;   // the padding words of a row are kept at zero
    float4 v = reinterpret_cast<const float4*>(src)[w];
    if (4 * w + 0 >= n) v.x = 0.f;
    if (4 * w + 1 >= n) v.y = 0.f;
    if (4 * w + 2 >= n) v.z = 0.f;
    if (4 * w + 3 >= n) v.w = 0.f;
    reinterpret_cast<float4*>(dst + (size_t)env * NQC)[w] = v;
  }
}
// obs = qpos[7:] || qvel[6:]   (dp_env_v3.py:62-65)
__device__ __forceinline__ void write_obs(const ModelS& M, const EnvS& S, float* obs, float* rec, int env, int lane) {
  const int np = M.nq - 7, nvv = M.nv - 6, od = np + nvv;
  for (int o = lane; o < od; o += 32) {
    const float v = o < np ? S.qpos[7 + o] : S.qvel[6 + o - np];
    if (obs) obs[(size_t)env * od + o] = v;
    if (rec) rec[(size_t)env * (od + 2) + o] = v;
  }
}

// ---------------------------------------------------------------------------------------
// kinematic features of the pose in (S.qpos, S.qvel), shared by the imitation reward and the
// DeepMimic state
// ---------------------------------------------------------------------------------------
// fresh kinematics + CoM frame + per-dof spatial velocity terms cdof[d] * qvel[d] (in buf6)
__device__ __noinline__ void kin_vel_prep(const ModelS& M, EnvS& S, int lane) {
  kinematics(M, S, lane);
  com_pos(M, S, lane);
  for (int d = lane; d < M.nv; d += 32) {
    const float qv = S.qvel[d];
#pragma unroll
    for (int k = 0; k < 6; k++) S.o.k.buf6[6 * d + k] = S.o.k.cdof[6 * d + k] * qv;
  }
  __syncwarp();
}
// spatial velocity [omega; v at the c-frame origin] of body b: sum over its dof chain
__device__ __forceinline__ void body_cvel(const ModelS& M, const EnvS& S, int b, float* v) {
#pragma unroll
  for (int k = 0; k < 6; k++) v[k] = 0.f;
  unsigned long long mk = M.body_dofmask[b];
  while (mk) {
    const int d = __ffsll((long long)mk) - 1;
    mk &= mk - 1;
#pragma unroll
    for (int k = 0; k < 6; k++) v[k] += S.o.k.buf6[6 * d + k];
  }
}

// ---------------------------------------------------------------------------------------
// time-based mocap phase with interpolation (phase_mode 1); fp32 restatement of
// oracle/dm_oracle.c quat_slerp / euler_rxyz_from_quat / dmo_mocap_sample
// ---------------------------------------------------------------------------------------
// transformations.quaternion_slerp (shortest path).  The angle comes from the chord |q1 - q0| =
// 2 sin(theta/2), which keeps its relative accuracy in fp32 where acos(dot) does not; below
// 0.01 rad the sine weights equal the linear ones to 2e-8 rad.
__device__ __forceinline__ Q4 quat_slerp(Q4 q0, Q4 q1, float f) {
  q0 = qnormalize(q0); q1 = qnormalize(q1);
  const float d = q0.w * q1.w + q0.x * q1.x + q0.y * q1.y + q0.z * q1.z;
  if (d < 0.f) { q1.w = -q1.w; q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z; }
  const float dw = q1.w - q0.w, dx = q1.x - q0.x, dy = q1.y - q0.y, dz = q1.z - q0.z;
  const float c2 = dw * dw + dx * dx + dy * dy + dz * dz;
  float w0 = 1.f - f, w1 = f;
  if (c2 >= 1e-4f) {
    const float th = 2.f * asinf(fminf(1.f, 0.5f * sqrtf(c2)));
    const float isin = 1.f / sinf(th);
    w0 = sinf((1.f - f) * th) * isin;
    w1 = sinf(f * th) * isin;
  }
  Q4 r;
  r.w = w0 * q0.w + w1 * q1.w; r.x = w0 * q0.x + w1 * q1.x; r.y = w0 * q0.y + w1 * q1.y; r.z = w0 * q0.z + w1 * q1.z;
  return qnormalize(r);
}
// transformations.euler_from_quaternion(q, 'rxyz'): angles of R = Rx(a) Ry(b) Rz(c)
__device__ __forceinline__ void euler_rxyz_from_quat(Q4 q, float& a, float& b, float& c) {
  q = qnormalize(q);
  const float s = 1.41421356237f;
  const float w = q.w * s, x = q.x * s, y = q.y * s, z = q.z * s;
  const float R00 = 1.f - y * y - z * z, R01 = x * y - z * w, R02 = x * z + y * w;
  const float R10 = x * y + z * w, R11 = 1.f - x * x - z * z, R12 = y * z - x * w, R22 = 1.f - x * x - y * y;
  const float cy = sqrtf(R22 * R22 + R12 * R12);
  if (cy > 4.8e-7f) { a = atan2f(-R12, R22); b = atan2f(R02, cy); c = atan2f(-R01, R00); }
  else { a = 0.f; b = atan2f(R02, cy); c = atan2f(R10, R11); }
}
// quaternion of a hinge triple, R = Rx(a) Ry(b) Rz(c)
__device__ __forceinline__ Q4 quat_from_xyz(float a, float b, float c) {
  float sa, ca, sb, cb, sc, cc;
  sincosf(0.5f * a, &sa, &ca); sincosf(0.5f * b, &sb, &cb); sincosf(0.5f * c, &sc, &cc);
  Q4 qx; qx.w = ca; qx.x = sa; qx.y = 0.f; qx.z = 0.f;
  Q4 qy; qy.w = cb; qy.x = 0.f; qy.y = sb; qy.z = 0.f;
  Q4 qz; qz.w = cc; qz.x = 0.f; qz.y = 0.f; qz.z = sc;
  return qmul(qmul(qx, qy), qz);
}
// frame coordinate u -> cycle, frame interval k, fraction alpha, phase in [0,1); same double
// arithmetic as the oracle's phase_split (an F-frame clip has F-1 intervals per cycle)
__device__ __forceinline__ void phase_split(int F, double u, int& cycle, int& k, float& alpha, float& phase) {
  cycle = 0; k = 0; alpha = 0.f; phase = 0.f;
  if (F < 2) return;
  const double c = floor(u / (double)(F - 1));
  const double uu = u - c * (double)(F - 1);
  int kk = (int)uu;
  kk = kk > F - 2 ? F - 2 : (kk < 0 ? 0 : kk);
  cycle = (int)c; k = kk; alpha = (float)(uu - (double)kk); phase = (float)(uu / (double)(F - 1));
}
__device__ __forceinline__ double frame_coord(const ModelS& M, int clip, int idx_init, int steps) {
  return __dadd_rn((double)idx_init, __dmul_rn((double)steps, M.clip_rate[clip]));
}
// interpolated reference pose at frame coordinate u into rq[nq] / rv[nv] (shared memory):
// linear for the root position (+ cycle * last frame's root xy: root-offset accumulation of
// MocapDM.play, mocap_v2.py:168-182), 1-DoF joints and velocities (lane = coordinate); slerp for the
// root quaternion and the hinge triples (lane = body).  Returns the frame interval k.
__device__ __noinline__ int mocap_sample(const ModelS& M, const DevPtrs& P, int clip, double u, int lane, float* rq,
                                         float* rv) {
  const int F = M.clip_len[clip], start = M.clip_start[clip];
  int cycle, k;
  float a, ph;
  phase_split(F, u, cycle, k, a, ph);
  const int k1 = F < 2 ? k : k + 1;
  const float* c0 = P.mocap_cfg + (size_t)(start + k) * M.nq;
  const float* c1 = P.mocap_cfg + (size_t)(start + k1) * M.nq;
  const float* cl = P.mocap_cfg + (size_t)(start + F - 1) * M.nq;
  const float* v0 = P.mocap_vel + (size_t)(start + k) * M.nv;
  const float* v1 = P.mocap_vel + (size_t)(start + k1) * M.nv;
  for (int i = lane; i < M.nq; i += 32) { const float x0 = c0[i]; rq[i] = fmaf(a, c1[i] - x0, x0); }
  for (int i = lane; i < M.nv; i += 32) { const float x0 = v0[i]; rv[i] = fmaf(a, v1[i] - x0, x0); }
  __syncwarp();
  if (lane == 1) {
    rq[0] += (float)cycle * cl[0];
    rq[1] += (float)cycle * cl[1];
    Q4 q0; q0.w = c0[3]; q0.x = c0[4]; q0.y = c0[5]; q0.z = c0[6];
    Q4 q1; q1.w = c1[3]; q1.x = c1[4]; q1.y = c1[5]; q1.z = c1[6];
    const Q4 q = quat_slerp(q0, q1, a);
    rq[3] = q.w; rq[4] = q.x; rq[5] = q.y; rq[6] = q.z;
  } else if (lane >= 2 && lane < M.nbody && M.body_dofnum[lane] == 3) {
    const int qa = M.body_dofadr[lane] + 1;
    const Q4 q = quat_slerp(quat_from_xyz(c0[qa], c0[qa + 1], c0[qa + 2]), quat_from_xyz(c1[qa], c1[qa + 1], c1[qa + 2]), a);
    euler_rxyz_from_quat(q, rq[qa], rq[qa + 1], rq[qa + 2]);
  }
  __syncwarp();
  return k;
}
// phase in [0, 1) reported by the DeepMimic state
__device__ __forceinline__ float phase_of(const ModelS& M, int clip, int idx_init, int idx_curr, int ep_len) {
  const int F = M.clip_len[clip];
  if (M.phase_mode == 1) {
    int cycle, k;
    float a, ph;
    phase_split(F, frame_coord(M, clip, idx_init, ep_len), cycle, k, a, ph);
    return ph;
  }
  return (float)idx_curr / (float)F;
}

// DeepMimic state (obs_mode 1; cCtController::BuildStatePose / BuildStateVel, code.md:287-504, and
// record_state, mujoco_env.py:91-124): [phase, root height, npart x (pos 3, quat 4), npart x
// (lin vel 3, ang vel 3)] in the root heading frame, lane = body part.  The row is staged in the
// (dead) composite-inertia arrays of the tile and written out coalesced.
__device__ __noinline__ void write_obs_dm(const ModelS& M, EnvS& S, float phase, float* obs, float* rec, int env, int lane) {
  kin_vel_prep(M, S, lane);
  float* buf = S.o.k.cinert;   // cinert + crb (contiguous, 320 floats) are dead after com_pos
  static_assert(2 + 13 * DMB_MAX_PART <= 2 * NB * 10 && offsetof(PhaseK, crb) == offsetof(PhaseK, cinert) + sizeof(float) * NB * 10, "obs staging");
  const float* R = &S.o.k.xmat[9];
  const float heading = atan2f(R[3], R[0]);
  float sh, ch, shh, chh;
  sincosf(heading, &sh, &ch);
  sincosf(0.5f * heading, &shh, &chh);
  if (lane == 0) { buf[0] = phase; buf[1] = S.o.k.xpos[5]; }
  if (lane < M.npart) {
    const int g = M.part_geom[lane], b = M.geom_bodyid[g];
    const V3 gp = ld3(&S.o.k.xpos[3 * b]) + mat_vec(&S.o.k.xmat[9 * b], ld3(M.geom_pos[g]));
    const V3 rel = gp - ld3(&S.o.k.xpos[3]);
    float* o = buf + 2 + 7 * lane;
    o[0] = ch * rel.x + sh * rel.y; o[1] = -sh * rel.x + ch * rel.y; o[2] = rel.z;
    Q4 qh; qh.w = chh; qh.x = 0.f; qh.y = 0.f; qh.z = -shh;
    Q4 qb; qb.w = S.o.k.xquat[4 * b]; qb.x = S.o.k.xquat[4 * b + 1]; qb.y = S.o.k.xquat[4 * b + 2]; qb.z = S.o.k.xquat[4 * b + 3];
    Q4 q = qmul(qh, qb);
    const float sg = q.w < 0.f ? -1.f : 1.f;
    o[3] = sg * q.w; o[4] = sg * q.x; o[5] = sg * q.y; o[6] = sg * q.z;
    float v[6];
    body_cvel(M, S, b, v);
    const V3 vl = v3(v[3], v[4], v[5]) + cross(v3(v[0], v[1], v[2]), gp - ld3(S.com));
    float* ov = buf + 2 + 7 * M.npart + 6 * lane;
    ov[0] = ch * vl.x + sh * vl.y; ov[1] = -sh * vl.x + ch * vl.y; ov[2] = vl.z;
    ov[3] = ch * v[0] + sh * v[1]; ov[4] = -sh * v[0] + ch * v[1]; ov[5] = v[2];
  }
  __syncwarp();
  const int od = M.obs_dim;
  for (int o = lane; o < od; o += 32) {
    const float v = buf[o];
    if (obs) obs[(size_t)env * od + o] = v;
    if (rec) rec[(size_t)env * (od + 2) + o] = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void emit_obs(const ModelS& M, EnvS& S, int clip, int idx_init, int idx_curr, int ep_len,
                                         float* obs, float* rec, int env, int lane) {
  if (M.obs_mode == 1) write_obs_dm(M, S, phase_of(M, clip, idx_init, idx_curr, ep_len), obs, rec, env, lane);
  else write_obs(M, S, obs, rec, env, lane);
}

// action -> per-dof actuator force (mj_fwdActuation with the ctrl clamp; optional PD)
__device__ __forceinline__ void set_ctrl(const ModelS& M, EnvS& S, const float* action, int env, int lane) {
  for (int d = lane; d < M.nv; d += 32) S.ctrlf[d] = 0.f;
  __syncwarp();
  if (lane < M.nu) {
    const int d = M.act_dofadr[lane];
    float a = action[(size_t)env * M.nu + lane];
    if (M.ctrl_mode != 0) {
      // PD on the joint angle: action = target angle.  mode 1 restates the intent of
      // MujocoInterface.action2torque (mujoco_interface.py:97-107): v_target = p_err/dt;
      // mode 2 is the plain PD  kp (a - q) - kd qvel.  Torque is mapped back to ctrl by the gear.
      const int j = lane;  // actuator u drives hinge joint with qpos index 7+u in this model family
      (void)j;
      const float q = S.qpos[d + 1];  // hinge dof d <-> qpos d+1 (free joint: 7 qpos, 6 dofs)
      const float perr = a - q;
      float tau;
      if (M.ctrl_mode == 1) tau = M.dof_kp[d] * perr + M.dof_kd[d] * (perr / M.pd_dt - S.qvel[d]);
      else tau = M.dof_kp[d] * perr - M.dof_kd[d] * S.qvel[d];
      a = tau / M.dof_gear[d];
    }
    if (!(a == a)) a = 0.f;
    a = fminf(fmaxf(a, M.dof_ctrl_lo[d]), M.dof_ctrl_hi[d]);
    S.ctrlf[d] = M.dof_gear[d] * a;
  }
  __syncwarp();
}

// mj_integratePos from X0 with velocity S.x_dv and step h into S.qpos (lane = joint)
__device__ __forceinline__ void integrate_pos(const ModelS& M, EnvS& S, int lane, float h) {
  if (lane < M.njnt) {
    const int qa = M.jnt_qposadr[lane], da = M.jnt_dofadr[lane];
    if (M.jnt_type[lane] == DMB_JNT_FREE) {
      S.qpos[qa] = S.x_q0[qa] + h * S.x_dv[da];
      S.qpos[qa + 1] = S.x_q0[qa + 1] + h * S.x_dv[da + 1];
      S.qpos[qa + 2] = S.x_q0[qa + 2] + h * S.x_dv[da + 2];
      V3 ax = v3(S.x_dv[da + 3], S.x_dv[da + 4], S.x_dv[da + 5]);
      const float ang = h * normalize(ax);
      float sn, cs;
      sincosf(0.5f * ang, &sn, &cs);
      Q4 qr; qr.w = cs; qr.x = sn * ax.x; qr.y = sn * ax.y; qr.z = sn * ax.z;
      Q4 q0; q0.w = S.x_q0[qa + 3]; q0.x = S.x_q0[qa + 4]; q0.y = S.x_q0[qa + 5]; q0.z = S.x_q0[qa + 6];
      Q4 qn = qnormalize(qmul(qnormalize(q0), qr));
      S.qpos[qa + 3] = qn.w; S.qpos[qa + 4] = qn.x; S.qpos[qa + 5] = qn.y; S.qpos[qa + 6] = qn.z;
    } else {
      S.qpos[qa] = S.x_q0[qa] + h * S.x_dv[da];
    }
  }
}

// mj_step with RK4 (mj_RungeKutta N=4).  X0 velocities and the weighted stage sums live in
// registers (dof d on lane d & 31).  forward_eval has a single call site (code size matters:
// the kernel is instruction-fetch bound).  Returns the CoM height of the last stage evaluation.
template <bool LOCKSTEP>
__device__ __forceinline__ float rk4_step(const ModelS& M, EnvS& S, int lane, bool active, int bar_id, int bar_n,
                                          EnvS* tiles, int* share_cnt, float* gcta, int warp, bool use_slot) {
  const float h = M.timestep;
  const int d0 = lane, d1 = lane + 32;
  const bool a0 = active && d0 < M.nv, a1 = active && d1 < M.nv;
  if (active) for (int i = lane; i < M.nq; i += 32) S.x_q0[i] = S.qpos[i];
  const float v00 = a0 ? S.qvel[d0] : 0.f, v01 = a1 ? S.qvel[d1] : 0.f;
  __syncwarp();
  float zc = 0.f, sv0 = 0.f, sv1 = 0.f, sa0 = 0.f, sa1 = 0.f;
#pragma unroll 1
  for (int st = 0; st < 4; st++) {
    if (st > 0) {
      // X[st] = X0 (+) h * a * (V[st-1], F[st-1])
      const float a = st == 3 ? 1.f : 0.5f;
      if (a0) { S.x_dv[d0] = a * S.qvel[d0]; S.qvel[d0] = v00 + h * (a * S.qacc[d0]); }
      if (a1) { S.x_dv[d1] = a * S.qvel[d1]; S.qvel[d1] = v01 + h * (a * S.qacc[d1]); }
      __syncwarp();
      if (active) integrate_pos(M, S, lane, h);
      __syncwarp();
    }
    zc = forward_eval<LOCKSTEP>(M, S, lane, nullptr, active, bar_id, bar_n, tiles, share_cnt, gcta, warp, use_slot);
    const float bw = (st == 0 || st == 3) ? (1.f / 6.f) : (1.f / 3.f);
    if (a0) { sv0 += bw * S.qvel[d0]; sa0 += bw * S.qacc[d0]; }
    if (a1) { sv1 += bw * S.qvel[d1]; sa1 += bw * S.qacc[d1]; }
  }
  if (a0) { S.x_dv[d0] = sv0; S.qvel[d0] = v00 + h * sa0; }
  if (a1) { S.x_dv[d1] = sv1; S.qvel[d1] = v01 + h * sa1; }
  __syncwarp();
  if (active) integrate_pos(M, S, lane, h);
  __syncwarp();
  return zc;
}

// clip ids come from a caller-owned tensor: out-of-range ids are clamped (BatchedSim rejects them up front)
__device__ __forceinline__ int clip_of(const ModelS& M, const dmb_state_t& st, int env) {
  const int c = st.clip[env];
  return c < 0 ? 0 : (c >= M.nclip ? M.nclip - 1 : c);
}
// reference-state initialisation (dp_env_v3.py:67-71,148-164); Philox keyed by (seed, env id)
__device__ int reset_env(const ModelS& M, EnvS& S, const DevPtrs& P, const dmb_state_t& st, int env, int lane,
                         unsigned long long seed, unsigned first_env_id, int mode) {
  const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
  const unsigned eid = first_env_id + (unsigned)env;
  const unsigned rc = st.reset_count[env];
  const int clip = clip_of(M, st, env);
  const int len = M.clip_len[clip], start = M.clip_start[clip];
  unsigned r[4];
  philox4x32(k0, k1, eid, rc, 0u, 0u, r);
  int idx = (int)(u01(r[0]) * (float)len);
  if (idx >= len) idx = len - 1;
  if (mode == 0) {
    for (int i = lane; i < M.nq; i += 32) S.qpos[i] = P.mocap_cfg[(size_t)(start + idx) * M.nq + i];
    for (int i = lane; i < M.nv; i += 32) S.qvel[i] = P.mocap_vel[(size_t)(start + idx) * M.nv + i];
  } else {
    for (int i = lane; i < M.nq + M.nv; i += 32) {
      philox4x32(k0, k1, eid, rc, 1u + (unsigned)(i >> 2), 0u, r);
      const float nz = M.reset_noise * (2.f * u01(r[i & 3]) - 1.f);
      if (i < M.nq) S.qpos[i] = M.qpos0[i] + nz; else S.qvel[i - M.nq] = nz;
    }
  }
  for (int i = lane; i < M.nv; i += 32) S.qacc[i] = 0.f;   // qacc_warmstart
  if (lane == 0) {
    st.idx_init[env] = idx; st.idx_curr[env] = idx; st.reset_count[env] = rc + 1u;
    st.ep_len[env] = 0; st.ep_ret[env] = 0.f;
  }
  __syncwarp();
  return idx;
}

__device__ __forceinline__ float quat_diff_theta(Q4 a, Q4 b) {
  Q4 ac; ac.w = a.w; ac.x = -a.x; ac.y = -a.y; ac.z = -a.z;
  const Q4 qd = qmul(ac, b);
  return 2.f * atan2f(sqrtf(qd.x * qd.x + qd.y * qd.y + qd.z * qd.z), fabsf(qd.w));
}

// end-effector points in the root heading frame (lane = end effector) and whole-body CoM velocity
// (all lanes) of the pose in (S.qpos, S.qvel); runs fresh kinematics
struct PoseFeat { float ex, ey, ez, vx, vy, vz; };
__device__ __noinline__ PoseFeat pose_features(const ModelS& M, EnvS& S, int lane) {
  kin_vel_prep(M, S, lane);
  PoseFeat out;
  // CoM velocity: sum_b m_b (v_lin + w x (xipos_b - com)) / M
  float px = 0.f, py = 0.f, pz = 0.f;
  if (lane >= 1 && lane < M.nbody) {
    const int b = lane;
    float v[6];
    body_cvel(M, S, b, v);
    const V3 r = ld3(&S.o.k.xipos[3 * b]) - ld3(S.com);
    const V3 vb = v3(v[3], v[4], v[5]) + cross(v3(v[0], v[1], v[2]), r);
    const float ms = M.body_mass[b];
    px = ms * vb.x; py = ms * vb.y; pz = ms * vb.z;
  }
  out.vx = warp_sum(px) * M.inv_total_mass; out.vy = warp_sum(py) * M.inv_total_mass; out.vz = warp_sum(pz) * M.inv_total_mass;
  out.ex = out.ey = out.ez = 0.f;
  if (lane < M.nee) {
    const float* R = &S.o.k.xmat[9];
    const float heading = atan2f(R[3], R[0]);
    float sh, ch;
    sincosf(heading, &sh, &ch);
    const int b = M.ee_body[lane];
    const V3 w = ld3(&S.o.k.xpos[3 * b]) + mat_vec(&S.o.k.xmat[9 * b], ld3(M.ee_pos[lane]));
    const float rx = w.x - S.o.k.xpos[3], ry = w.y - S.o.k.xpos[4];
    out.ex = ch * rx + sh * ry; out.ey = -sh * rx + ch * ry; out.ez = w.z;
  }
  return out;
}

// 5-term DeepMimic imitation reward (code.md:979-1146 adapted to the hinge model; weights and
// scales from dp_env_v3.py:42-53) against the reference pose rq / rv with features `ref`.
// (offx, offy): root offset of the reference accumulated over completed passes of the clip (phase_mode 0)
__device__ float reward_imitate(const ModelS& M, EnvS& S, int lane, const float* rq, const float* rv, PoseFeat ref,
                                float offx, float offy) {
  const PoseFeat cur = pose_features(M, S, lane);
  float ee = 0.f;
  if (lane < M.nee) {
    const float e0 = cur.ex - ref.ex, e1 = cur.ey - ref.ey, e2 = cur.ez - ref.ez;
    ee = e0 * e0 + e1 * e1 + e2 * e2;
  }
  ee = warp_sum(ee);
  if (M.nee > 0) ee /= (float)M.nee;
  // pose error: lane = body (root quaternion on lane 1, hinge triples / single hinges on lanes >= 2)
  float pe = 0.f, th_root = 0.f;
  if (lane == 1) {
    Q4 q0; q0.w = S.qpos[3]; q0.x = S.qpos[4]; q0.y = S.qpos[5]; q0.z = S.qpos[6];
    Q4 q1; q1.w = rq[3]; q1.x = rq[4]; q1.y = rq[5]; q1.z = rq[6];
    th_root = quat_diff_theta(qnormalize(q0), qnormalize(q1));
    pe = M.dof_weight[3] * th_root * th_root;
  } else if (lane >= 2 && lane < M.nbody) {
    const int da = M.body_dofadr[lane], nd = M.body_dofnum[lane];
    if (nd == 3) {
      const float th = quat_diff_theta(quat_from_xyz(S.qpos[da + 1], S.qpos[da + 2], S.qpos[da + 3]),
                                       quat_from_xyz(rq[da + 1], rq[da + 2], rq[da + 3]));
      pe = M.dof_weight[da] * th * th;
    } else if (nd == 1) {
      const float dq = S.qpos[da + 1] - rq[da + 1];
      pe = M.dof_weight[da] * dq * dq;
    }
  }
  th_root = __shfl_sync(DMB_FULL, th_root, 1);
  const float pose_err = warp_sum(pe);
  float ve = 0.f;
  for (int d = lane; d < M.nv; d += 32) if (d >= 3) { const float dv = S.qvel[d] - rv[d]; ve += M.dof_weight[d] * dv * dv; }
  const float vel_err = warp_sum(ve);
  float rp = 0.f, rvv = 0.f, rw = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float a = S.qpos[i] - (rq[i] + (i == 0 ? offx : (i == 1 ? offy : 0.f))), b = S.qvel[i] - rv[i], c = S.qvel[3 + i] - rv[3 + i];
    rp += a * a; rvv += b * b; rw += c * c;
  }
  const float root_err = rp + 0.1f * th_root * th_root + 0.01f * rvv + 0.001f * rw;
  const float cx = ref.vx - cur.vx, cy = ref.vy - cur.vy, cz = ref.vz - cur.vz;
  const float com_err = 0.1f * (cx * cx + cy * cy + cz * cz);
  return M.w_pose * expf(-M.s_err * M.s_pose * pose_err) + M.w_vel * expf(-M.s_err * M.s_vel * vel_err) +
         M.w_ee * expf(-M.s_err * M.s_ee * ee) + M.w_root * expf(-M.s_err * M.s_root * root_err) +
         M.w_com * expf(-M.s_err * M.s_com * com_err);
}

__device__ __forceinline__ bool state_bad(const ModelS& M, const EnvS& S, int lane) {
  bool bad = false;
  for (int i = lane; i < M.nq; i += 32) bad |= !(fabsf(S.qpos[i]) < 1e10f);
  for (int i = lane; i < M.nv; i += 32) bad |= !(fabsf(S.qvel[i]) < 1e10f);
  return __any_sync(DMB_FULL, bad);
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// Model tables global -> shared: one TMA bulk copy (cp.async.bulk, completion on an mbarrier) issued by thread 0
// instead of a word-by-word copy loop on every thread of the CTA.
static_assert(sizeof(ModelS) % 16 == 0, "cp.async.bulk moves multiples of 16 bytes");
__device__ __forceinline__ void stage_model(ModelS* dst, const ModelS* src) {
  __shared__ __align__(8) unsigned long long s_mbar;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar);
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  constexpr unsigned bytes = (unsigned)sizeof(ModelS);
  if ((threadIdx.x | threadIdx.y) == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(bar) : "memory");
  }
  __syncthreads();   // the barrier is initialised and armed before anybody polls it
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar) : "memory");
  }
}

#ifndef DMB_MAXTHREADS
#define DMB_MAXTHREADS 896   // 28 warps: 72 registers per thread
#endif
static_assert(YS % 2 == 1, "odd row stride");
constexpr size_t MODEL_BYTES = (sizeof(ModelS) + 15) & ~(size_t)15;

template <bool LOCKSTEP>
__global__ void __launch_bounds__(DMB_MAXTHREADS, 1) k_step(DevPtrs P, dmb_state_t st, const float* __restrict__ action, dmb_step_out_t out, int N,
                       unsigned long long seed, unsigned first_env_id) {
  extern __shared__ __align__(16) unsigned char smem[];
  ModelS& M = *reinterpret_cast<ModelS*>(smem);
  EnvS* tiles = reinterpret_cast<EnvS*>(smem + MODEL_BYTES);
  stage_model(&M, P.model);
  // (pinning warp / lane in registers with an opaque asm removes the S2R + shift + multiply-add re-derivations of the
  // tile address -- 13 % of the instructions under the 72-register cap -- but costs 110 bytes of extra spills:
  // measured 7.03 M vs 7.21 M env-steps/s without the pin)
  // tile kernels run 2-D blocks (32, W): lane and warp come straight from %tid.x / %tid.y (under the 72-register cap
  // the compiler re-derives them and the tile address all over the kernel; this saves the shift and the mask, +1.2 %)
  const int warp = threadIdx.y, lane = threadIdx.x, W = blockDim.y;
  EnvS& S = tiles[warp];
  const int od = M.obs_dim;
  __shared__ int s_base[8];
  __shared__ int s_diag[4];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 4) s_diag[tid] = 0;
  if (tid == 0) s_slot_owner = -1;
  // (one shared slot per CTA lies behind the tiles: Y of a stage with more than RF rows, see count_rows)
#if DMB_SHARE
  __shared__ int s_share[16];   // task counters of the work sharing (two per lockstep group)
  if (tid < 16) s_share[tid] = 0;
#else
  int* const s_share = nullptr;
#endif
  // lockstep groups: the CTA's warps are split into M.ngroups groups, each with its own named
  // barrier and its own pull from the scheduler
  const int gsz = LOCKSTEP ? W / M.ngroups : 1, grp = LOCKSTEP ? warp / gsz : 0, gw = LOCKSTEP ? warp % gsz : 0;
  const int bar_id = 1 + grp, bar_n = gsz * 32;
  float* const gcta = P.gscratch + (size_t)blockIdx.x * W * gs::stride;
  // Single round (every env resident at once, e.g. 4096 envs on 148 x 28 warps): the cost-sorted list is dealt out
  // round robin, so that every CTA holds the same mix of heavy and light envs and no SM carries more instructions
  // than the others.  Otherwise CTAs pull W consecutive entries at a time, heaviest first.
  const bool spread = LOCKSTEP && M.spread && N <= (int)gridDim.x * W;
  int round = 0;
  if (P.trace && tid == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[blockIdx.x * 8] = t;
  }
  if (P.wait_flag && blockIdx.x == 0 && tid == 0) {
    // consumer side of the fused all-gather folded into this launch (no extra kernel on the stream): one thread
    // polls the arrival flag of an earlier step -- normally long satisfied -- so that the completion of this
    // kernel implies that every rank's rows of that step have landed (acquire, system scope)
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(P.wait_flag) : "memory");
      if (v < P.wait_target) __nanosleep(200);
    } while (v < P.wait_target);
    __threadfence_system();
  }
  for (;;) {
    // Scheduling: envs are handed out in order of decreasing constraint work (k_order); a group
    // takes gsz consecutive entries at a time so that its warps see similar work between the
    // lockstep barriers; without lockstep every warp pulls for itself.
    int env = N;
    if (spread) {
      if (round > 0) break;
      if (tid == 0) s_base[0] = 0;
      __syncthreads();
      // (spread 2 deals the rows of the sorted list alternately left-to-right and right-to-left: the CTA that gets
      // the heaviest env of a row gets the lightest of the next one)
      const int col = (M.spread == 2 && (warp & 1)) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
      const int idx = warp * (int)gridDim.x + col;
      if (idx < N) env = P.order[idx];
    } else if (LOCKSTEP) {
      group_barrier(bar_id, bar_n);
      if (gw == 0 && lane == 0) s_base[grp] = atomicAdd(P.counter, gsz);
      group_barrier(bar_id, bar_n);
      const int idx = s_base[grp] + gw;
      if (idx < N) env = P.order[idx];
    } else {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(P.counter, 1);
      idx = __shfl_sync(DMB_FULL, idx, 0);
      if (idx < N) env = P.order[idx];
    }
    const bool have = env < N;
    if (LOCKSTEP) { if (s_base[grp] >= N) break; }
    else if (!have) break;
    DMB_TICK(-1);
    bool bad = false;
    if (have) {
      load_state(M, S, st, env, lane);
      bad = state_bad(M, S, lane);
      if (!bad) set_ctrl(M, S, action, env, lane);
    }
    // work sharing happens inside a lockstep group: its tiles, its scratch slots, its two task counters
    const float zc = rk4_step<LOCKSTEP>(M, S, lane, have && !bad, bar_id, bar_n, tiles + grp * gsz,
                                        LOCKSTEP ? s_share + 2 * grp : nullptr, gcta + (size_t)grp * gsz * gs::stride, gw,
                                        LOCKSTEP && M.ngroups == 1);
    if (!have) { round++; continue; }
    DMB_TICK(22);
    if (!bad) bad = state_bad(M, S, lane);
    // reward (dp_env_v3.py:117 / 89-104).  Reference pose: phase_mode 0 = table row of the integer frame
    // counter; phase_mode 1 = interpolated at the post-step mocap time into S.x_q0 / S.x_dv (dead here)
    float rew = 1.0f;
    int idx_curr = st.idx_curr[env], idx_init = st.idx_init[env];
    const int clip = clip_of(M, st, env);
    const int len = M.clip_len[clip], start = M.clip_start[clip];
    int ep_len = st.ep_len[env] + 1;
    const int rmode = M.reward_mode;
    if (rmode != 0 && !bad) {
      const float* rq;
      const float* rv;
      int frame = idx_curr;                                   // modes 1, 4: reward first, then advance
      if (rmode == 2 || rmode == 3) frame = (idx_curr + 1) % len;  // v2 / v1: the frame advances first
      int k_ref = 0;
      if (M.phase_mode == 1) {
        k_ref = mocap_sample(M, P, clip, frame_coord(M, clip, idx_init, ep_len), lane, S.x_q0, S.x_dv);
        rq = S.x_q0; rv = S.x_dv;
      } else {
        rq = P.mocap_cfg + (size_t)(start + frame) * M.nq;
        rv = P.mocap_vel + (size_t)(start + frame) * M.nv;
      }
      if (rmode == 1) {
        float e = 0.f;
        for (int j = lane; j < M.nq - 7; j += 32) e += fabsf(S.qpos[7 + j] - rq[7 + j]);
        rew = expf(-warp_sum(e));
      } else if (rmode == 4) {
        PoseFeat ref;
        float offx = 0.f, offy = 0.f;
        if (M.phase_mode == 1) {
          // features of the interpolated reference pose: swap it into the tile, run the kinematics, swap back
          const int i1 = lane + 32;
          const float q0 = S.qpos[lane], q1 = i1 < M.nq ? S.qpos[i1] : 0.f;
          const float v0 = S.qvel[lane], v1 = i1 < M.nv ? S.qvel[i1] : 0.f;
          __syncwarp();
          S.qpos[lane] = rq[lane]; if (i1 < M.nq) S.qpos[i1] = rq[i1];
          S.qvel[lane] = rv[lane]; if (i1 < M.nv) S.qvel[i1] = rv[i1];
          __syncwarp();
          ref = pose_features(M, S, lane);
          __syncwarp();
          S.qpos[lane] = q0; if (i1 < M.nq) S.qpos[i1] = q1;
          S.qvel[lane] = v0; if (i1 < M.nv) S.qvel[i1] = v1;
          __syncwarp();
        } else {
          const float* aux = P.ref_aux + (size_t)(start + frame) * DMB_REF_AUX;
          ref.ex = ref.ey = ref.ez = 0.f;
          if (lane < M.nee) { ref.ex = aux[3 * lane]; ref.ey = aux[3 * lane + 1]; ref.ez = aux[3 * lane + 2]; }
          ref.vx = aux[12]; ref.vy = aux[13]; ref.vz = aux[14];
          // the frame counter wraps, the reference root keeps moving: every completed pass over the clip adds
          // the last frame's root xy (MocapDM.play, mocap_v2.py:168-182); ep_len already counts this step
          const float* cl = P.mocap_cfg + (size_t)(start + len - 1) * M.nq;
          const float cyc = (float)((idx_init + ep_len - 1) / len);
          offx = cyc * cl[0]; offy = cyc * cl[1];
        }
        rew = reward_imitate(M, S, lane, rq, rv, ref, offx, offy);
      } else {
        // v2 / v1 rewards (dp_env_v2.py:116-183, dp_env_v1.py:82-152); the control cost is on the raw action
        float acs = 0.f;
        if (lane < M.nu) { const float a = action[(size_t)env * M.nu + lane]; acs = a * a; }
        acs = warp_sum(acs);
        if (rmode == 2) {
          float e = 0.f;
          for (int i = 3 + lane; i < M.nq; i += 32) e += fabsf(S.qpos[i] - rq[i]);
          rew = expf(-M.s_err * M.s_pose * warp_sum(e)) - 0.1f * acs;
        } else {
          float pe = 0.f;
          if (lane == 1) {
            Q4 q0; q0.w = S.qpos[3]; q0.x = S.qpos[4]; q0.y = S.qpos[5]; q0.z = S.qpos[6];
            Q4 q1; q1.w = rq[3]; q1.x = rq[4]; q1.y = rq[5]; q1.z = rq[6];
            pe = M.dof_weight[3] * quat_diff_theta(qnormalize(q0), qnormalize(q1));
          } else if (lane >= 2 && lane < M.nbody) {
            const int da = M.body_dofadr[lane], nd = M.body_dofnum[lane];
            if (nd == 3) pe = M.dof_weight[da] * quat_diff_theta(quat_from_xyz(S.qpos[da + 1], S.qpos[da + 2], S.qpos[da + 3]),
                                                                 quat_from_xyz(rq[da + 1], rq[da + 2], rq[da + 3]));
            else if (nd == 1) pe = M.dof_weight[da] * fabsf(S.qpos[da + 1] - rq[da + 1]);
          }
          const float pose = warp_sum(pe) * M.joint_weight_sum;
          float ve = 0.f;
          for (int d = lane; d < M.nv; d += 32) if (d >= 3) ve += fabsf(S.qvel[d] - rv[d]);
          const float vel = warp_sum(ve);
          const float root = fabsf(S.qpos[0] - rq[0]) + fabsf(S.qpos[1] - rq[1]) + fabsf(S.qpos[2] - rq[2]);
          rew = M.w_pose * expf(-M.s_err * M.s_pose * pose) + M.w_vel * expf(-M.s_err * M.s_vel * vel) +
                M.w_root * expf(-M.s_err * M.s_root * root) - 0.1f * acs;
        }
      }
      (void)k_ref;
    }
    DMB_TICK(23);
    // frame counter: phase_mode 0 advances by one per step whenever a mocap reward is configured (also for a
    // non-finite state); phase_mode 1 reports the frame interval of the post-step reference time
    if (M.phase_mode == 1) {
      int cyc; float al, ph;
      phase_split(len, frame_coord(M, clip, idx_init, ep_len), cyc, idx_curr, al, ph);
    } else if (rmode != 0) idx_curr = (idx_curr + 1) % len;
    if (bad) rew = 0.f;
    // termination (dp_env_v3.py:134-139) on the CoM height of the last stage evaluation
    bool done = bad || zc < M.z_min || zc > M.z_max;
    if (M.term_mode == 1 && !bad) {  // DeepMimic fall-contact rule on the contacts of the last RK4 stage
      bool fall = false;
      if (lane < S.ncon) fall = M.geom_type[cm_g1(S.c_meta[lane])] == DMB_GEOM_PLANE && ((M.fall_body_mask >> M.geom_bodyid[cm_g2(S.c_meta[lane])]) & 1u);
      done = done || __any_sync(DMB_FULL, fall);
    }
    int flags = S.flags | (bad ? 4 : 0);
    if (P.trace) {   // diagnostics in the upper bits: rows << 8 (6 bits) | PGS sweeps << 14 (8) | warp time in 2-us units << 22
      long long tnow;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
      const long long dt = (tnow - P.trace[blockIdx.x * 8]) / 2000;
      flags |= (min(S.diag >> 16, 63) << 8) | (min(S.diag & 0xffff, 255) << 14) | ((int)min(dt, 1023LL) << 22);
    }
    float ep_ret = st.ep_ret[env] + rew;
    if (lane == 0) {
      out.reward[env] = rew;
      out.done[env] = done ? 1 : 0;
      if (out.rec) { out.rec[(size_t)env * (od + 2) + od] = rew; out.rec[(size_t)env * (od + 2) + od + 1] = done ? 1.f : 0.f; }
      if (out.last_ret) out.last_ret[env] = ep_ret;
      if (out.last_len) out.last_len[env] = ep_len;
      st.flags[env] = flags;
      P.cost[env] = min(255, S.cost >> 4);
      st.idx_curr[env] = idx_curr;
      st.ep_len[env] = ep_len;
      st.ep_ret[env] = ep_ret;
    }
    __syncwarp();
    if (done && M.auto_reset) {
      idx_init = idx_curr = reset_env(M, S, P, st, env, lane, seed, first_env_id, M.reset_mode);
      ep_len = 0;
    } else if (bad) {  // keep a finite state in HBM
      for (int i = lane; i < M.nq; i += 32) S.qpos[i] = M.qpos0[i];
      for (int i = lane; i < M.nv; i += 32) { S.qvel[i] = 0.f; S.qacc[i] = 0.f; }
      __syncwarp();
    }
    emit_obs(M, S, clip, idx_init, idx_curr, ep_len, out.obs, out.rec, env, lane);
    if (P.n_peer > 0) {
      // fused all-gather: the record row (obs, reward, done) goes straight into every rank's gathered buffer --
      // posted stores over NVLink that overlap the rest of the kernel; no collective kernel competes for the SMs
      const size_t rowoff = (size_t)(P.row0 + env) * (od + 2);
      const int np = M.nq - 7;
      for (int o = lane; o < od + 2; o += 32) {
        float v;
        if (o >= od) v = o == od ? rew : (done ? 1.f : 0.f);
        else if (M.obs_mode == 1) v = S.o.k.cinert[o];                 // staged by write_obs_dm
        else v = o < np ? S.qpos[7 + o] : S.qvel[6 + o - np];          // dp_env_v3.py:62-65
        for (int p = 0; p < P.n_peer; p++) P.peer_rec[p][rowoff + o] = v;
      }
    }
    store_state(M, S, st, env, lane);
    __syncwarp();
    DMB_TICK(24);
    ++round;
    if (P.trace && (warp | lane) == 0 && round < 6) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[blockIdx.x * 8 + round] = t;
    }
    if (P.trace && lane == 0) {   // per-CTA diagnostics of the step: most PGS sweeps / rows of one env, scratch-path envs
      atomicMax(&s_diag[0], S.diag & 0xffff);
      atomicMax(&s_diag[1], S.diag >> 16);
      if ((S.diag >> 16) > RF) atomicAdd(&s_diag[2], 1);
    }
  }
  if (P.n_peer > 0) {
    // release pattern of a put-with-signal: every thread fences its peer stores at system scope, the CTA counts itself
    // done, and the last CTA of the launch (which has observed every other CTA's count) signals all peers
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      const int done_ctas = atomicAdd(P.ticket, 1);
      if (done_ctas == (int)gridDim.x - 1) {
        __threadfence_system();
        for (int p = 0; p < P.n_peer; p++) atomicAdd_system(P.peer_flag[p], 1);
        *P.ticket = 0;
      }
    }
  }
  if (P.trace) {   // slot 6: the CTA's last warp is done; slot 7: diagnostics (sweeps | rows << 16 | scratch envs << 24)
    __syncthreads();
    if (tid == 0) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[blockIdx.x * 8 + 6] = t;
      P.trace[blockIdx.x * 8 + 7] = (long long)s_diag[0] | ((long long)s_diag[1] << 16) | ((long long)s_diag[2] << 24);
    }
  }
}

// Counting sort of the envs by the constraint work of their previous step (256 buckets,
// heaviest first) + reset of the scheduler counter.  One CTA; order inside a bucket is arbitrary
// (envs are independent, so results do not depend on the schedule).
__global__ void k_order(const int* __restrict__ cost, int* __restrict__ order, int* counter, int N, int identity) {
  __shared__ int hist[256], cursor[256];
  if (identity) {  // DMB_NO_SORT=1: schedule in index order (experiments)
    for (int i = threadIdx.x; i < N; i += blockDim.x) order[i] = i;
    if (threadIdx.x == 0) *counter = 0;
    return;
  }
  if (threadIdx.x < 256) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(&hist[min(255, max(0, cost[i]))], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 255; b >= 0; b--) { cursor[b] = acc; acc += hist[b]; }
    *counter = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) order[atomicAdd(&cursor[min(255, max(0, cost[i]))], 1)] = i;
}

// consumer side of the fused all-gather: one thread polls this rank's arrival flag (acquire, system scope)
__global__ void k_peer_wait(const int* flag, int target) {
  int v;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v < target) __nanosleep(200);
  } while (v < target);
  __threadfence_system();
}

__global__ void __launch_bounds__(DMB_MAXTHREADS, 1) k_reset(DevPtrs P, dmb_state_t st, const unsigned char* __restrict__ mask, int mode, float* obs, int N,
                        unsigned long long seed, unsigned first_env_id) {
  extern __shared__ __align__(16) unsigned char smem[];
  ModelS& M = *reinterpret_cast<ModelS*>(smem);
  EnvS* tiles = reinterpret_cast<EnvS*>(smem + MODEL_BYTES);
  stage_model(&M, P.model);
  const int warp = threadIdx.y, lane = threadIdx.x, W = blockDim.y;
  EnvS& S = tiles[warp];
  for (int env = blockIdx.x * W + warp; env < N; env += gridDim.x * W) {
    if (mask && !mask[env]) continue;
    const int idx = reset_env(M, S, P, st, env, lane, seed, first_env_id, mode < 0 ? M.reset_mode : mode);
    if (lane == 0) st.flags[env] = 0;
    if (obs) emit_obs(M, S, clip_of(M, st, env), idx, idx, 0, obs, nullptr, env, lane);
    store_state(M, S, st, env, lane);
    __syncwarp();
  }
}

__global__ void k_obs(DevPtrs P, dmb_state_t st, float* obs, int N) {
  const int nq = P.model->nq, nv = P.model->nv;
  const int np = nq - 7, od = np + nv - 6;
  const size_t total = (size_t)N * od;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t env = i / od;
    const int o = (int)(i % od);
    obs[i] = o < np ? st.qpos[env * DMB_QSTRIDE + 7 + o] : st.qvel[env * DMB_VSTRIDE + 6 + o - np];
  }
}

// DeepMimic state of the stored state (obs_mode 1): one warp per env, needs the tile for the kinematics
__global__ void __launch_bounds__(DMB_MAXTHREADS, 1) k_obs_dm(DevPtrs P, dmb_state_t st, float* obs, int N) {
  extern __shared__ __align__(16) unsigned char smem[];
  ModelS& M = *reinterpret_cast<ModelS*>(smem);
  EnvS* tiles = reinterpret_cast<EnvS*>(smem + MODEL_BYTES);
  stage_model(&M, P.model);
  const int warp = threadIdx.y, lane = threadIdx.x, W = blockDim.y;
  EnvS& S = tiles[warp];
  for (int env = blockIdx.x * W + warp; env < N; env += gridDim.x * W) {
    load_state(M, S, st, env, lane);
    emit_obs(M, S, clip_of(M, st, env), st.idx_init[env], st.idx_curr[env], st.ep_len[env], obs, nullptr, env, lane);
    __syncwarp();
  }
}

// Interpolated reference poses for arbitrary (clip, frame coordinate) pairs: one warp per sample
// (dmb_mocap_sample; the kinematic playback of MocapDM.play, mocap_v2.py:151-182, without a simulator)
constexpr int SAMPLE_WARPS = 8;
__global__ void k_mocap_sample(DevPtrs P, const int* __restrict__ clip, const double* __restrict__ u, int n, float* qpos,
                               float* qvel, float* phase) {
  extern __shared__ __align__(16) unsigned char smem[];
  ModelS& M = *reinterpret_cast<ModelS*>(smem);
  float* scratch = reinterpret_cast<float*>(smem + MODEL_BYTES);
  stage_model(&M, P.model);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* rq = scratch + warp * 2 * NQC;
  float* rv = rq + NQC;
  for (int i = blockIdx.x * SAMPLE_WARPS + warp; i < n; i += gridDim.x * SAMPLE_WARPS) {
    int c = clip ? clip[i] : 0;
    c = c < 0 ? 0 : (c >= M.nclip ? M.nclip - 1 : c);
    const double ui = u[i];
    mocap_sample(M, P, c, ui, lane, rq, rv);
    for (int k = lane; k < DMB_QSTRIDE; k += 32) qpos[(size_t)i * DMB_QSTRIDE + k] = k < M.nq ? rq[k] : 0.f;
    for (int k = lane; k < DMB_VSTRIDE; k += 32) qvel[(size_t)i * DMB_VSTRIDE + k] = k < M.nv ? rv[k] : 0.f;
    if (phase && lane == 0) {
      int cyc, kk; float al, ph;
      phase_split(M.clip_len[c], ui, cyc, kk, al, ph);
      phase[i] = ph;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(DMB_MAXTHREADS, 1) k_forward_debug(DevPtrs P, dmb_state_t st, const float* __restrict__ ctrl, float* dbgout, int N) {
  extern __shared__ __align__(16) unsigned char smem[];
  ModelS& M = *reinterpret_cast<ModelS*>(smem);
  EnvS* tiles = reinterpret_cast<EnvS*>(smem + MODEL_BYTES);
  stage_model(&M, P.model);
  const int warp = threadIdx.y, lane = threadIdx.x, W = blockDim.y;
  EnvS& S = tiles[warp];
  for (int env = blockIdx.x * W + warp; env < N; env += gridDim.x * W) {
    load_state(M, S, st, env, lane);
    // ctrl is given directly (no PD) in the debug path
    for (int d = lane; d < M.nv; d += 32) S.ctrlf[d] = 0.f;
    __syncwarp();
    if (lane < M.nu) {
      const int d = M.act_dofadr[lane];
      float a = ctrl[(size_t)env * M.nu + lane];
      a = fminf(fmaxf(a, M.dof_ctrl_lo[d]), M.dof_ctrl_hi[d]);
      S.ctrlf[d] = M.dof_gear[d] * a;
    }
    __syncwarp();
    float* row = dbgout + (size_t)env * dbg::stride;
    for (int i = lane; i < dbg::stride; i += 32) row[i] = 0.f;
    __syncwarp();
    const float zc = forward_eval<false>(M, S, lane, row, true, 0, 0, nullptr, nullptr,
                                         P.gscratch + (size_t)blockIdx.x * W * gs::stride, warp, false);
    for (int i = lane; i < M.nv; i += 32) row[dbg::qacc + i] = S.qacc[i];
    if (lane == 0) {
      row[dbg::com] = S.com[0]; row[dbg::com + 1] = S.com[1]; row[dbg::com + 2] = S.com[2];
      row[dbg::ncon] = (float)S.ncon; row[dbg::nefc] = (float)S.nefc; row[dbg::iter] = (float)S.iter;
      row[dbg::z_com] = zc;
      st.flags[env] = S.flags;
    }
    // mj_forward leaves qacc_warmstart = qacc
    for (int i = lane; i < DMB_VSTRIDE; i += 32) st.warm[(size_t)env * DMB_VSTRIDE + i] = i < M.nv ? S.qacc[i] : 0.f;
    __syncwarp();
  }
}

}  // namespace dmb

// =========================================================================================
// host side
// =========================================================================================
using namespace dmb;

struct dmb_handle_s {
  int device = 0;
  int num_envs = 0;
  unsigned long long seed = 0;
  unsigned first_env_id = 0;
  ModelS hmodel;
  ModelS* dmodel = nullptr;
  float *d_cfg = nullptr, *d_vel = nullptr, *d_aux = nullptr, *d_scratch = nullptr;
  int *d_counter = nullptr, *d_cost = nullptr, *d_order = nullptr;
  long long* d_trace = nullptr;
  int grid = 0, block = 0, smem = 0, envs_per_cta = 0;
  int nu = 0, obs_dim = 0;
  int lockstep = 1;
  int no_sort = 0;
  int sort_period = 1;      // single-round schedule: re-sort the env list every sort_period steps
  long long nstep = 0;
  long long nlaunch = 0;    // kernels launched through this handle so far
  int n_peer = 0, row0 = 0;  // fused all-gather targets (dmb_set_peer_gather)
  float* peer_rec[DMB_MAX_PEER] = {nullptr};
  int* peer_flag[DMB_MAX_PEER] = {nullptr};
  int* d_ticket = nullptr;
  const int* wait_flag = nullptr;   // one-shot: consumed by the next dmb_step
  int wait_target = 0;
  std::string err;
};

static std::string g_err;

static int fail(dmb_handle_t h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_err = msg;
  return code;
}
#define CUDA_TRY(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(h, DMB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

static int build_model(const dmb_model_t* m, const dmb_config_t* c, const dmb_mocap_t* mc, ModelS& S, std::string& why) {
  memset(&S, 0, sizeof(S));
  if (m->nv > YS || m->nv > NVT || m->nq > NQC - 1 || m->nM > NMT || m->nbody > NB || m->njnt > NJ || m->ngeom > NG || m->npair > NP || m->nu > NU ||
      m->nM > NMX || m->nv > 64) { why = "model exceeds kernel capacities"; return DMB_ERR_MODEL; }
  if (m->max_efc > MAXROW || m->max_con > MAXC || m->max_con > 32 || m->max_efc < m->njnt) {
    why = "max_efc must be <= 40 and >= njnt, max_con <= 16 (kernel capacities)"; return DMB_ERR_MODEL;
  }
  if (m->ngeom > 255 || m->max_efc > 255 || m->npair > 256) {
    why = "geom ids / row addresses are packed in bytes"; return DMB_ERR_MODEL;
  }
  S.nq = m->nq; S.nv = m->nv; S.nu = m->nu; S.nbody = m->nbody; S.njnt = m->njnt; S.ngeom = m->ngeom;
  S.npair = m->npair; S.nM = m->nM; S.iterations = m->iterations; S.max_con = m->max_con; S.max_efc = m->max_efc;
  S.timestep = (float)m->timestep; S.tolerance = (float)m->tolerance; S.margin = (float)m->margin;
  S.pgs_scale = (float)(1.0 / (m->meaninertia * (m->nv > 1 ? m->nv : 1)));
  for (int k = 0; k < 3; k++) S.gravity[k] = (float)m->gravity[k];
  for (int k = 0; k < 5; k++) S.solimp[k] = (float)m->solimp[k];
  {  // mj_makeImpedance spring constants (refsafe)
    double tc = m->solref[0], dr = m->solref[1];
    if (tc < 2 * m->timestep) tc = 2 * m->timestep;
    double dmax = m->solimp[1]; dmax = dmax < 1e-4 ? 1e-4 : (dmax > 0.9999 ? 0.9999 : dmax);
    double kk = dmax * dmax * tc * tc * dr * dr, bb = dmax * tc;
    S.imp_k = (float)(1.0 / (kk > 1e-15 ? kk : 1e-15));
    S.imp_b = (float)(2.0 / (bb > 1e-15 ? bb : 1e-15));
  }
  double mass = 0;
  int maxdepth = 0;
  for (int b = 0; b < m->nbody; b++) {
    S.body_parent[b] = (int8_t)m->body_parent[b]; S.body_depth[b] = (int8_t)m->body_depth[b];
    S.body_jntadr[b] = (int8_t)m->body_jntadr[b]; S.body_jntnum[b] = (int8_t)m->body_jntnum[b];
    S.body_dofadr[b] = (int8_t)m->body_dofadr[b]; S.body_dofnum[b] = (int8_t)m->body_dofnum[b];
    if (m->body_jntnum[b] > JPB) { why = "more than 3 joints on a body"; return DMB_ERR_MODEL; }
    for (int k = 0; k < 3; k++) { S.body_pos[b][k] = (float)m->body_pos[b][k]; S.body_ipos[b][k] = (float)m->body_ipos[b][k]; }
    for (int k = 0; k < 4; k++) S.body_quat[b][k] = (float)m->body_quat[b][k];
    for (int k = 0; k < 6; k++) S.body_inertia[b][k] = (float)m->body_inertia[b][k];
    S.body_mass[b] = (float)m->body_mass[b];
    S.body_invw[b] = (float)m->body_invweight0[b][0];
    mass += m->body_mass[b];
    if (m->body_depth[b] > maxdepth) maxdepth = m->body_depth[b];
    if (b > 0) {
      int p = m->body_parent[b];
      if (p > 0) {
        if (S.body_nchild[p] >= 4) { why = "more than 4 children on a body"; return DMB_ERR_MODEL; }
        S.body_child[p][S.body_nchild[p]++] = (int8_t)b;
      }
    }
  }
  S.maxdepth = maxdepth;
  S.inv_total_mass = (float)(1.0 / mass);
  for (int j = 0; j < m->njnt; j++) {
    S.jnt_type[j] = (int8_t)m->jnt_type[j]; S.jnt_qposadr[j] = (int8_t)m->jnt_qposadr[j];
    S.jnt_dofadr[j] = (int8_t)m->jnt_dofadr[j]; S.jnt_limited[j] = (int8_t)m->jnt_limited[j];
    S.jnt_bodyid[j] = (int8_t)m->jnt_bodyid[j];
    for (int k = 0; k < 3; k++) S.jnt_axis[j][k] = (float)m->jnt_axis[j][k];
    S.jnt_range[j][0] = (float)m->jnt_range[j][0]; S.jnt_range[j][1] = (float)m->jnt_range[j][1];
    S.jnt_qpos0[j] = m->jnt_type[j] == DMB_JNT_HINGE ? (float)m->qpos0[m->jnt_qposadr[j]] : 0.f;
    if (m->jnt_type[j] == DMB_JNT_FREE && m->body_parent[m->jnt_bodyid[j]] != 0) {
      why = "free joints are only supported on top-level bodies"; return DMB_ERR_MODEL;
    }
    if (m->jnt_type[j] == DMB_JNT_HINGE && m->jnt_qposadr[j] != m->jnt_dofadr[j] + 1) {
      why = "hinge qpos/dof addressing must be qposadr == dofadr + 1 (single leading free joint)"; return DMB_ERR_MODEL;
    }
  }
  for (int d = 0; d < m->nv; d++) {
    S.dof_bodyid[d] = (int8_t)m->dof_bodyid[d];
    const int j = m->dof_jntid[d];
    if (m->jnt_type[j] == DMB_JNT_FREE) {
      const int k = d - m->jnt_dofadr[j];
      S.dof_kind[d] = k < 3 ? DOF_FREE_TRANS : DOF_FREE_ROT;
      S.dof_axisk[d] = (int8_t)(k % 3);
    } else { S.dof_kind[d] = DOF_HINGE; S.dof_axisk[d] = 0; }
    int c = 0;
    for (int a = m->dof_parentid[d]; a >= 0; a = m->dof_parentid[a]) {
      if (c >= MAXANC) { why = "dof chain longer than 12"; return DMB_ERR_MODEL; }
      S.dof_anc[d][c++] = (int8_t)a;
    }
    S.dof_nanc[d] = (int8_t)c;
    for (int k = 0; k < c; k++) S.anc_rowbase[d][k] = (int16_t)m->dof_Madr[S.dof_anc[d][k]];
    S.dof_Madr[d] = (int16_t)m->dof_Madr[d];
    S.dof_Lend[d] = (int16_t)(m->dof_Madr[d] + c);
    S.ldl_meta[d] = (uint32_t)c | ((uint32_t)(c * (c + 1) / 2) << 8) | ((uint32_t)m->dof_Madr[d] << 16);
    for (int k = 0; k < c; k++) S.dof_ancr[d][k] = (uint8_t)S.dof_anc[d][c - 1 - k];
    if (c > S.maxanc) S.maxanc = c;
    {  // descendants must be the contiguous id range d+1 .. d+ndesc (depth-first numbering, as MuJoCo compiles it)
      int nd = 0;
      for (int e = d + 1; e < m->nv; e++) {
        bool desc = false;
        for (int a = m->dof_parentid[e]; a >= 0; a = m->dof_parentid[a]) if (a == d) { desc = true; break; }
        if (desc) { if (e != d + 1 + nd) { why = "dofs are not numbered depth first"; return DMB_ERR_MODEL; } nd++; }
      }
      S.dof_ndesc[d] = (int8_t)nd;
    }
    S.dof_armature[d] = (float)m->dof_armature[d]; S.dof_damping[d] = (float)m->dof_damping[d];
    S.dof_invw[d] = (float)m->dof_invweight0[d];
    S.dof_act[d] = -1; S.dof_gear[d] = 1.f; S.dof_ctrl_lo[d] = 0.f; S.dof_ctrl_hi[d] = 0.f;
    S.dof_weight[d] = (float)m->dof_weight[d];
    { unsigned long long am = 0; for (int a = m->dof_parentid[d]; a >= 0; a = m->dof_parentid[a]) am |= 1ull << a; S.dof_ancmask[d] = am; }
    int e = m->dof_Madr[d];
    for (int a = d; a >= 0; a = m->dof_parentid[a]) { S.M_i[e] = (uint8_t)d; S.M_j[e] = (uint8_t)a; e++; }
  }
  for (int d = 0; d < m->nv; d++) {
    for (int sft = 0; sft < 4; sft++) {  // 1st, 2nd, 4th, 8th ancestor
      int a = d;
      for (int k = 0; k < (1 << sft) && a >= 0; k++) a = m->dof_parentid[a];
      S.dof_jump[sft][d] = (int8_t)a;
    }
    const int j = m->dof_jntid[d];
    if (S.dof_kind[d] == DOF_HINGE) S.dof_vsrc[d] = (int8_t)m->dof_parentid[d];
    else if (S.dof_kind[d] == DOF_FREE_ROT) S.dof_vsrc[d] = (int8_t)(m->jnt_dofadr[j] + 2);
    else S.dof_vsrc[d] = -1;
    const int b = m->dof_bodyid[d];
    S.dof_lastof[d] = (int8_t)(d == m->body_dofadr[b] + m->body_dofnum[b] - 1 ? b : -1);
  }
  if (maxdepth > 8) { why = "body tree deeper than 8"; return DMB_ERR_MODEL; }
  for (int b = 0; b < m->nbody; b++)
    for (int sft = 0; sft < 3; sft++) {
      int a = b;
      for (int k = 0; k < (1 << sft) && a > 0; k++) a = m->body_parent[a];
      S.body_jump[sft][b] = (int8_t)(b > 0 && a > 0 ? a : -1);
    }
  for (int b = 1; b < m->nbody; b++)
    if (m->body_dofnum[b] < 1) { why = "every moving body needs at least one dof"; return DMB_ERR_MODEL; }
  for (int b = 0; b < m->nbody; b++) {
    unsigned long long mk = 0;
    for (int bb = b; bb > 0; bb = m->body_parent[bb])
      for (int d = m->body_dofadr[bb]; d < m->body_dofadr[bb] + m->body_dofnum[bb]; d++) mk |= 1ull << d;
    S.body_dofmask[b] = mk;
  }
  { int t = 0; for (int q = 0; q < MAXANC; q++) for (int p = 0; p <= q; p++) { S.tri_p[t] = (uint8_t)p; S.tri_q[t] = (uint8_t)q; t++; } }
  for (int g = 0; g < m->ngeom; g++) {
    S.geom_type[g] = (int8_t)m->geom_type[g]; S.geom_bodyid[g] = (int8_t)m->geom_bodyid[g];
    S.geom_condim[g] = (int8_t)m->geom_condim[g];
    if (m->geom_condim[g] != 1 && m->geom_condim[g] != 3) { why = "condim must be 1 or 3"; return DMB_ERR_MODEL; }
    for (int k = 0; k < 3; k++) { S.geom_size[g][k] = (float)m->geom_size[g][k]; S.geom_pos[g][k] = (float)m->geom_pos[g][k]; }
    for (int k = 0; k < 4; k++) S.geom_quat[g][k] = (float)m->geom_quat[g][k];
    S.geom_identq[g] = (m->geom_quat[g][0] == 1.0 && m->geom_quat[g][1] == 0.0 && m->geom_quat[g][2] == 0.0 && m->geom_quat[g][3] == 0.0);
    S.geom_rbound[g] = (float)m->geom_rbound[g]; S.geom_mu[g] = (float)m->geom_friction[g][0];
  }
  for (int p = 0; p < m->npair; p++) { S.pair_g1[p] = (uint8_t)m->pair_geom1[p]; S.pair_g2[p] = (uint8_t)m->pair_geom2[p]; }
  for (int u = 0; u < m->nu; u++) {
    const int d = m->act_dofadr[u];
    S.act_dofadr[u] = (int8_t)d; S.dof_act[d] = (int8_t)u; S.dof_gear[d] = (float)m->act_gear[u];
    S.dof_ctrl_lo[d] = (float)m->act_ctrlrange[u][0]; S.dof_ctrl_hi[d] = (float)m->act_ctrlrange[u][1];
    S.dof_kp[d] = (float)m->act_kp[u]; S.dof_kd[d] = (float)m->act_kd[u];
  }
  for (int i = 0; i < m->nq; i++) S.qpos0[i] = (float)m->qpos0[i];
  S.nee = m->nee;
  for (int e = 0; e < m->nee; e++) { S.ee_body[e] = m->ee_body[e]; for (int k = 0; k < 3; k++) S.ee_pos[e][k] = (float)m->ee_pos[e][k]; }
  S.ctrl_mode = c->ctrl_mode; S.reward_mode = c->reward_mode; S.reset_mode = c->reset_mode; S.auto_reset = c->auto_reset;
  S.joint_weight_sum = (float)c->joint_weight_sum;
  S.term_mode = c->term_mode; S.fall_body_mask = c->fall_body_mask;
  S.phase_mode = c->phase_mode; S.obs_mode = c->obs_mode;
  if (c->phase_mode < 0 || c->phase_mode > 1) { why = "phase_mode must be 0 or 1"; return DMB_ERR_ARG; }
  if (c->obs_mode < 0 || c->obs_mode > 1) { why = "obs_mode must be 0 or 1"; return DMB_ERR_ARG; }
  if (m->npart < 0 || m->npart > DMB_MAX_PART) { why = "npart out of range"; return DMB_ERR_MODEL; }
  S.npart = m->npart;
  for (int p = 0; p < m->npart; p++) {
    if (m->part_geom[p] < 0 || m->part_geom[p] >= m->ngeom) { why = "part_geom out of range"; return DMB_ERR_MODEL; }
    S.part_geom[p] = (int8_t)m->part_geom[p];
  }
  if (c->obs_mode == 1 && m->npart < 1) { why = "obs_mode 1 needs the body-part table (npart >= 1)"; return DMB_ERR_MODEL; }
  S.obs_dim = c->obs_mode == 1 ? 2 + 13 * m->npart : (m->nq - 7) + (m->nv - 6);
  S.z_min = (float)c->z_min; S.z_max = (float)c->z_max; S.reset_noise = (float)c->reset_noise; S.pd_dt = (float)m->timestep;
  S.w_pose = (float)c->w_pose; S.w_vel = (float)c->w_vel; S.w_ee = (float)c->w_end_eff; S.w_root = (float)c->w_root; S.w_com = (float)c->w_com;
  S.s_pose = (float)c->s_pose; S.s_vel = (float)c->s_vel; S.s_ee = (float)c->s_end_eff; S.s_root = (float)c->s_root; S.s_com = (float)c->s_com;
  S.s_err = (float)c->s_err;
  if (c->reward_mode < 0 || c->reward_mode > 4) { why = "reward_mode must be 0..4"; return DMB_ERR_ARG; }
  if (c->ctrl_mode < 0 || c->ctrl_mode > 2) { why = "ctrl_mode must be 0, 1 or 2"; return DMB_ERR_ARG; }
  S.nclip = mc->nclip; S.nframe_total = mc->nframe_total;
  S.ngroups = 1;
  S.cost_mode = 0;
  S.spread = 1;
  if (const char* sp = getenv("DMB_SPREAD")) S.spread = atoi(sp);
  if (const char* cm = getenv("DMB_COST_MODE")) S.cost_mode = atoi(cm);
  S.sync_mask = 0x41;  // barriers at the start of every RK stage and before the constraint solve (sweep on B200)
  if (const char* sm = getenv("DMB_SYNC_MASK")) S.sync_mask = (int)strtol(sm, nullptr, 0);
  if (mc->nclip < 1 || mc->nclip > DMB_MAX_CLIP) { why = "need 1..16 motion clips"; return DMB_ERR_ARG; }
  for (int k = 0; k < mc->nclip; k++) {
    S.clip_start[k] = mc->clip_start[k]; S.clip_len[k] = mc->clip_len[k];
    if (mc->clip_len[k] < 1 || mc->clip_start[k] < 0 || mc->clip_start[k] + mc->clip_len[k] > mc->nframe_total) {
      why = "clip_start / clip_len outside the mocap tables"; return DMB_ERR_ARG;
    }
    S.clip_rate[k] = mc->clip_dt[k] > 0 ? m->timestep / mc->clip_dt[k] : 0.0;
    if (c->phase_mode == 1 && !(mc->clip_dt[k] > 0)) { why = "phase_mode 1 needs clip_dt > 0"; return DMB_ERR_ARG; }
  }
  return DMB_OK;
}

extern "C" {

int dmb_version(void) { return DMB_VERSION; }
int32_t dmb_sizeof_tile(void) { return (int32_t)sizeof(EnvS); }
int32_t dmb_sizeof_model(void) { return (int32_t)sizeof(dmb_model_t); }
int32_t dmb_sizeof_config(void) { return (int32_t)sizeof(dmb_config_t); }
int32_t dmb_sizeof_mocap(void) { return (int32_t)sizeof(dmb_mocap_t); }

const char* dmb_last_error(dmb_handle_t h) { return h ? h->err.c_str() : g_err.c_str(); }

int32_t dmb_debug_stride(void) { return dbg::stride; }

int32_t dmb_debug_offset(const char* name) {
  struct { const char* n; int o; } tab[] = {
      {"xpos", dbg::xpos}, {"xquat", dbg::xquat}, {"xipos", dbg::xipos}, {"com", dbg::com}, {"qM", dbg::qM},
      {"qLD", dbg::qLD}, {"qfrc_bias", dbg::qfrc_bias}, {"qfrc_smooth", dbg::qfrc_smooth},
      {"qacc_smooth", dbg::qacc_smooth}, {"ncon", dbg::ncon}, {"nefc", dbg::nefc}, {"iter", dbg::iter},
      {"z_com", dbg::z_com}, {"contact", dbg::contact}, {"efc_pos", dbg::efc_pos}, {"efc_R", dbg::efc_R},
      {"efc_aref", dbg::efc_aref}, {"efc_b", dbg::efc_b}, {"efc_force", dbg::efc_force},
      {"efc_AR_diag", dbg::efc_AR_diag}, {"qacc", dbg::qacc}, {"cvel", dbg::cvel}};
  if (!name) return -1;
  for (auto& t : tab) if (!strcmp(t.n, name)) return t.o;
  return -1;
}

int dmb_create(const dmb_model_t* model, const dmb_config_t* config, const dmb_mocap_t* mocap, int32_t num_envs,
               int32_t cuda_device, uint64_t seed, uint32_t first_env_id, dmb_handle_t* out) {
  if (!model || !config || !mocap || !out || num_envs <= 0) return fail(nullptr, DMB_ERR_ARG, "dmb_create: bad argument");
  if (!mocap->data_config || !mocap->data_vel || mocap->nframe_total <= 0) return fail(nullptr, DMB_ERR_ARG, "dmb_create: empty mocap tables");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, DMB_ERR_NO_DEVICE, "no CUDA device visible: libdmb200 has no CPU fallback");
  if (cuda_device < 0 || cuda_device >= ndev) return fail(nullptr, DMB_ERR_ARG, "dmb_create: bad device index");
  dmb_handle_t h = new (std::nothrow) dmb_handle_s();
  if (!h) return fail(nullptr, DMB_ERR_ARG, "out of host memory");
  std::string why;
  int rc = build_model(model, config, mocap, h->hmodel, why);
  if (rc != DMB_OK) { delete h; return fail(nullptr, rc, "dmb_create: " + why); }
  h->device = cuda_device; h->num_envs = num_envs; h->seed = seed; h->first_env_id = first_env_id;
  h->nu = model->nu; h->obs_dim = h->hmodel.obs_dim;
  cudaError_t e = cudaSetDevice(cuda_device);
  if (e != cudaSuccess) { delete h; return fail(nullptr, DMB_ERR_CUDA, cudaGetErrorString(e)); }
  const size_t F = (size_t)mocap->nframe_total;
  const size_t ncfg = F * model->nq, nvel = F * model->nv, naux = F * DMB_REF_AUX;
  float* tmp = new float[ncfg > naux ? (ncfg > nvel ? ncfg : nvel) : (naux > nvel ? naux : nvel)];
  auto upload = [&](const double* src, size_t n, float** dst) -> cudaError_t {
    cudaError_t ee = cudaMalloc((void**)dst, n * sizeof(float));
    if (ee != cudaSuccess) return ee;
    for (size_t i = 0; i < n; i++) tmp[i] = src ? (float)src[i] : 0.f;
    return cudaMemcpy(*dst, tmp, n * sizeof(float), cudaMemcpyHostToDevice);
  };
  e = cudaMalloc((void**)&h->dmodel, sizeof(ModelS));
  if (e == cudaSuccess) e = cudaMemcpy(h->dmodel, &h->hmodel, sizeof(ModelS), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_counter, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_ticket, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(h->d_ticket, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_cost, sizeof(int) * (size_t)num_envs);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_order, sizeof(int) * (size_t)num_envs);
  if (e == cudaSuccess) e = cudaMemset(h->d_cost, 0, sizeof(int) * (size_t)num_envs);
  if (e == cudaSuccess) e = upload(mocap->data_config, ncfg, &h->d_cfg);
  if (e == cudaSuccess) e = upload(mocap->data_vel, nvel, &h->d_vel);
  if (e == cudaSuccess) e = upload(mocap->ref_aux, naux, &h->d_aux);
  delete[] tmp;
  if (e != cudaSuccess) { std::string msg = cudaGetErrorString(e); dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, "dmb_create: " + msg); }
  // launch geometry: as many env tiles per CTA as opt-in shared memory allows (28 on B200), one CTA per SM
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, cuda_device);
  if (e != cudaSuccess) { dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, cudaGetErrorString(e)); }
  const void* tile_kernels[] = {(const void*)k_step<false>, (const void*)k_step<true>, (const void*)k_reset,
                                (const void*)k_forward_debug, (const void*)k_obs_dm};
  size_t static_smem = 0;
  int max_regs = 0;
  for (auto fn : tile_kernels) {   // every kernel launched with the tile geometry bounds it
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) { dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, cudaGetErrorString(e)); }
    if (fa.sharedSizeBytes > static_smem) static_smem = fa.sharedSizeBytes;
    if (fa.numRegs > max_regs) max_regs = fa.numRegs;
  }
  const size_t maxsmem = prop.sharedMemPerBlockOptin;
  constexpr size_t SLOT_BYTES = sizeof(float) * gs::stride;   // the CTA's shared big-stage slot
  int W = maxsmem > MODEL_BYTES + static_smem + SLOT_BYTES ? (int)((maxsmem - MODEL_BYTES - static_smem - SLOT_BYTES) / sizeof(EnvS)) : 0;
  if (W > DMB_MAXTHREADS / 32) W = DMB_MAXTHREADS / 32;   // __launch_bounds__ of the tile kernels
  {  // the register file bounds the resident warps as well
    const int regs = ((max_regs + 7) / 8) * 8;
    const int wreg = prop.regsPerMultiprocessor / (32 * (regs > 0 ? regs : 1));
    if (W > wreg) W = wreg;
  }
  if (const char* lenv = getenv("DMB_LOCKSTEP")) h->lockstep = atoi(lenv) != 0;
  if (const char* ns = getenv("DMB_NO_SORT")) h->no_sort = atoi(ns) != 0;
  h->sort_period = 8;
  if (const char* sp = getenv("DMB_SORT_PERIOD")) { int p = atoi(sp); if (p >= 1) h->sort_period = p; }
  if (const char* wenv = getenv("DMB_ENVS_PER_CTA")) { int w = atoi(wenv); if (w >= 1 && w < W) W = w; }
  if (W < 1) { dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, "not enough shared memory for one env tile"); }
  {  // lockstep groups: DMB_GROUPS (default 1); must divide the warps of a CTA
    int G = 1;
    if (const char* genv = getenv("DMB_GROUPS")) G = atoi(genv);
    if (G < 1 || G > 7 || W % G != 0) G = 1;
    h->hmodel.ngroups = G;
    e = cudaMemcpy(h->dmodel, &h->hmodel, sizeof(ModelS), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, cudaGetErrorString(e)); }
  }
  h->envs_per_cta = W; h->block = 32 * W;
  h->smem = (int)(MODEL_BYTES + (size_t)W * sizeof(EnvS) + SLOT_BYTES);
  int need = (num_envs + W - 1) / W;
  int nsm = prop.multiProcessorCount;
  // DMB_RESERVE_SMS=k leaves k SMs without a persistent CTA: a concurrent kernel of the caller (the NCCL all-gather of
  // the previous step's record, on its own stream) cannot share an SM with a CTA that owns the SM's whole shared memory
  if (const char* rs = getenv("DMB_RESERVE_SMS")) { const int k = atoi(rs); if (k > 0 && k < nsm) nsm -= k; }
  h->grid = need < nsm ? need : nsm;
  for (auto fn : tile_kernels) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem);
    if (e != cudaSuccess) { dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); }
  }
  e = cudaMalloc((void**)&h->d_scratch, sizeof(float) * (size_t)gs::stride * (size_t)h->grid * (size_t)W);
  if (e != cudaSuccess) { dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, cudaGetErrorString(e)); }
  if (const char* tr = getenv("DMB_TRACE")) {
    if (atoi(tr) != 0) {
      e = cudaMalloc((void**)&h->d_trace, sizeof(long long) * 8 * (size_t)h->grid);
      if (e == cudaSuccess) e = cudaMemset(h->d_trace, 0, sizeof(long long) * 8 * (size_t)h->grid);
      if (e != cudaSuccess) { dmb_destroy(h); return fail(nullptr, DMB_ERR_CUDA, cudaGetErrorString(e)); }
    }
  }
  *out = h;
  return DMB_OK;
}

/* DMB_TRACE=1 diagnostics: globaltimer (ns) of every CTA of the last dmb_step at kernel start and after each of
 * its scheduler rounds, [grid][8] int64 on the host; returns the grid size (0 when tracing is off). */
int32_t dmb_get_trace(dmb_handle_t h, int64_t* host_out, int32_t max_ctas) {
  if (!h || !host_out) return DMB_ERR_ARG;
  if (!h->d_trace) return 0;
  const int n = h->grid < max_ctas ? h->grid : max_ctas;
  if (cudaMemcpy(host_out, h->d_trace, sizeof(long long) * 8 * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return DMB_ERR_CUDA;
  cudaMemset(h->d_trace, 0, sizeof(long long) * 8 * (size_t)h->grid);
  return n;
}

int dmb_destroy(dmb_handle_t h) {
  if (!h) return DMB_ERR_ARG;
  cudaSetDevice(h->device);
  cudaFree(h->d_counter); cudaFree(h->d_cost); cudaFree(h->d_order); cudaFree(h->dmodel); cudaFree(h->d_cfg); cudaFree(h->d_vel); cudaFree(h->d_aux); cudaFree(h->d_trace); cudaFree(h->d_scratch); cudaFree(h->d_ticket);
  delete h;
  return DMB_OK;
}

static bool state_ok(const dmb_state_t* st) {
  return st && st->qpos && st->qvel && st->warm && st->clip && st->idx_init && st->idx_curr && st->reset_count &&
         st->ep_len && st->ep_ret && st->flags;
}
static DevPtrs devptrs(dmb_handle_t h) { DevPtrs P; P.counter = h->d_counter; P.cost = h->d_cost; P.order = h->d_order; P.model = h->dmodel; P.mocap_cfg = h->d_cfg; P.mocap_vel = h->d_vel; P.ref_aux = h->d_aux; P.trace = h->d_trace; P.gscratch = h->d_scratch; P.n_peer = h->n_peer; P.row0 = h->row0; P.ticket = h->d_ticket; P.wait_flag = nullptr; P.wait_target = 0; for (int p = 0; p < DMB_MAX_PEER; p++) { P.peer_rec[p] = h->peer_rec[p]; P.peer_flag[p] = h->peer_flag[p]; } return P; }

int dmb_reset(dmb_handle_t h, const dmb_state_t* st, const uint8_t* mask, int32_t mode, float* obs, void* stream) {
  if (!h) return DMB_ERR_ARG;
  if (!state_ok(st) || mode < -1 || mode > 1) return fail(h, DMB_ERR_ARG, "dmb_reset: bad argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  k_reset<<<h->grid, dim3(32, h->envs_per_cta), h->smem, (cudaStream_t)stream>>>(devptrs(h), *st, mask, mode, obs, h->num_envs, h->seed, h->first_env_id);
  CUDA_TRY(h, cudaGetLastError());
  return DMB_OK;
}

int dmb_step(dmb_handle_t h, const dmb_state_t* st, const float* action, const dmb_step_out_t* out, void* stream) {
  if (!h) return DMB_ERR_ARG;
  if (!state_ok(st) || !action || !out || !out->obs || !out->reward || !out->done) return fail(h, DMB_ERR_ARG, "dmb_step: bad argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  // Scheduler sort.  With dynamic pulling (more envs than resident warps) it runs every step: it also rewinds the
  // pull counter.  When one round holds every env the sorted list only balances the CTAs' mix of heavy and light
  // envs, which changes slowly: it is refreshed every sort_period steps (the sort costs 10 us + a launch gap).
  const bool single_round = h->lockstep && h->hmodel.spread && h->num_envs <= h->grid * h->envs_per_cta;
  if (!single_round || h->nstep % h->sort_period == 0) {
    k_order<<<1, 1024, 0, (cudaStream_t)stream>>>(h->d_cost, h->d_order, h->d_counter, h->num_envs, h->no_sort);
    h->nlaunch++;
  }
  h->nstep++;
  h->nlaunch++;
  DevPtrs P = devptrs(h);
  P.wait_flag = h->wait_flag; P.wait_target = h->wait_target;   // one-shot (dmb_set_peer_wait)
  h->wait_flag = nullptr; h->wait_target = 0;
  if (h->lockstep) k_step<true><<<h->grid, dim3(32, h->envs_per_cta), h->smem, (cudaStream_t)stream>>>(P, *st, action, *out, h->num_envs, h->seed, h->first_env_id);
  else k_step<false><<<h->grid, dim3(32, h->envs_per_cta), h->smem, (cudaStream_t)stream>>>(P, *st, action, *out, h->num_envs, h->seed, h->first_env_id);
  CUDA_TRY(h, cudaGetLastError());
  return DMB_OK;
}

int dmb_get_obs(dmb_handle_t h, const dmb_state_t* st, float* obs, void* stream) {
  if (!h) return DMB_ERR_ARG;
  if (!state_ok(st) || !obs) return fail(h, DMB_ERR_ARG, "dmb_get_obs: bad argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (h->hmodel.obs_mode == 1) {
    k_obs_dm<<<h->grid, dim3(32, h->envs_per_cta), h->smem, (cudaStream_t)stream>>>(devptrs(h), *st, obs, h->num_envs);
  } else {
    const size_t total = (size_t)h->num_envs * h->obs_dim;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    k_obs<<<blocks, 256, 0, (cudaStream_t)stream>>>(devptrs(h), *st, obs, h->num_envs);
  }
  CUDA_TRY(h, cudaGetLastError());
  return DMB_OK;
}

int dmb_forward_debug(dmb_handle_t h, const dmb_state_t* st, const float* ctrl, float* dbgout, void* stream) {
  if (!h) return DMB_ERR_ARG;
  if (!state_ok(st) || !ctrl || !dbgout) return fail(h, DMB_ERR_ARG, "dmb_forward_debug: bad argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  k_forward_debug<<<h->grid, dim3(32, h->envs_per_cta), h->smem, (cudaStream_t)stream>>>(devptrs(h), *st, ctrl, dbgout, h->num_envs);
  CUDA_TRY(h, cudaGetLastError());
  return DMB_OK;
}

int dmb_mocap_sample(dmb_handle_t h, const int32_t* clip, const double* frame_coord, int32_t n, float* qpos, float* qvel,
                     float* phase, void* stream) {
  if (!h) return DMB_ERR_ARG;
  if (!frame_coord || !qpos || !qvel || n < 0) return fail(h, DMB_ERR_ARG, "dmb_mocap_sample: bad argument");
  if (n == 0) return DMB_OK;
  CUDA_TRY(h, cudaSetDevice(h->device));
  int blocks = (n + SAMPLE_WARPS - 1) / SAMPLE_WARPS;
  if (blocks > 1184) blocks = 1184;
  const size_t smem = MODEL_BYTES + (size_t)SAMPLE_WARPS * 2 * NQC * sizeof(float);
  k_mocap_sample<<<blocks, SAMPLE_WARPS * 32, smem, (cudaStream_t)stream>>>(devptrs(h), clip, frame_coord, n, qpos, qvel, phase);
  CUDA_TRY(h, cudaGetLastError());
  return DMB_OK;
}

int32_t dmb_obs_dim(dmb_handle_t h) { return h ? h->obs_dim : DMB_ERR_ARG; }

/* ---- fused all-gather over NVLink peer memory (see include/dmb.h) ---- */
int dmb_peer_alloc(int32_t cuda_device, uint64_t bytes, void** ptr, uint8_t* handle64) {
  if (!ptr || !handle64 || bytes == 0) return DMB_ERR_ARG;
  if (cudaSetDevice(cuda_device) != cudaSuccess) return DMB_ERR_CUDA;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return DMB_ERR_CUDA;
  cudaIpcMemHandle_t hnd;
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaIpcGetMemHandle(&hnd, p) != cudaSuccess) { cudaFree(p); return DMB_ERR_CUDA; }
  memcpy(handle64, &hnd, 64);
  *ptr = p;
  return DMB_OK;
}
int dmb_peer_open(int32_t cuda_device, const uint8_t* handle64, void** ptr) {
  if (!ptr || !handle64) return DMB_ERR_ARG;
  if (cudaSetDevice(cuda_device) != cudaSuccess) return DMB_ERR_CUDA;
  cudaIpcMemHandle_t hnd;
  memcpy(&hnd, handle64, 64);
  // mapped into THIS device's address space; peer access to the owning GPU is enabled on first use
  if (cudaIpcOpenMemHandle(ptr, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return DMB_ERR_CUDA; }
  return DMB_OK;
}
int dmb_peer_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? DMB_OK : DMB_ERR_CUDA; }
int dmb_peer_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? DMB_OK : DMB_ERR_CUDA; }

/* ---- host-mapped I/O: pinned host memory the step kernel reads (action) / writes (record) directly over PCIe ---- */
int dmb_host_alloc(int32_t cuda_device, uint64_t bytes, void** host_ptr, void** dev_ptr) {
  if (!host_ptr || !dev_ptr || bytes == 0) return DMB_ERR_ARG;
  if (cudaSetDevice(cuda_device) != cudaSuccess) return DMB_ERR_CUDA;
  void *hp = nullptr, *dp = nullptr;
  if (cudaHostAlloc(&hp, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return DMB_ERR_CUDA; }
  if (cudaHostGetDevicePointer(&dp, hp, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(hp); return DMB_ERR_CUDA; }
  memset(hp, 0, bytes);
  *host_ptr = hp; *dev_ptr = dp;
  return DMB_OK;
}
int dmb_host_free(void* host_ptr) { return cudaFreeHost(host_ptr) == cudaSuccess ? DMB_OK : DMB_ERR_CUDA; }
int dmb_host_device_pointer(int32_t cuda_device, void* host_ptr, void** dev_ptr) {
  if (!host_ptr || !dev_ptr) return DMB_ERR_ARG;
  if (cudaSetDevice(cuda_device) != cudaSuccess) return DMB_ERR_CUDA;
  if (cudaHostGetDevicePointer(dev_ptr, host_ptr, 0) != cudaSuccess) { cudaGetLastError(); return DMB_ERR_ARG; }
  return DMB_OK;
}

int dmb_set_peer_gather(dmb_handle_t h, int32_t n_peer, float* const* rec_peer, int32_t* const* flag_peer, int32_t row0) {
  if (!h) return DMB_ERR_ARG;
  if (n_peer < 0 || n_peer > DMB_MAX_PEER || (n_peer > 0 && (!rec_peer || !flag_peer)) || row0 < 0)
    return fail(h, DMB_ERR_ARG, "dmb_set_peer_gather: bad argument");
  h->n_peer = n_peer; h->row0 = row0;
  for (int p = 0; p < DMB_MAX_PEER; p++) {
    h->peer_rec[p] = p < n_peer ? rec_peer[p] : nullptr;
    h->peer_flag[p] = p < n_peer ? (int*)flag_peer[p] : nullptr;
  }
  return DMB_OK;
}
int dmb_set_peer_wait(dmb_handle_t h, const int32_t* flag, int32_t target) {
  if (!h) return DMB_ERR_ARG;
  h->wait_flag = (const int*)flag; h->wait_target = target;
  return DMB_OK;
}
int dmb_peer_wait(dmb_handle_t h, const int32_t* flag, int32_t target, void* stream) {
  if (!h || !flag) return DMB_ERR_ARG;
  CUDA_TRY(h, cudaSetDevice(h->device));
  k_peer_wait<<<1, 1, 0, (cudaStream_t)stream>>>((const int*)flag, target);
  CUDA_TRY(h, cudaGetLastError());
  return DMB_OK;
}

int64_t dmb_kernel_launches(dmb_handle_t h) { return h ? h->nlaunch : DMB_ERR_ARG; }

#if DMB_PHASE_TIMERS
/* diagnostic builds only (-DDMB_PHASE_TIMERS=1): read and clear the per-phase cycle counters */
int32_t dmb_phase_cycles(uint64_t* host_out16) {
  unsigned long long z[32] = {0};
  if (cudaMemcpyFromSymbol(host_out16, dmb::g_phase_cycles, sizeof(z)) != cudaSuccess) return DMB_ERR_CUDA;
  cudaMemcpyToSymbol(dmb::g_phase_cycles, z, sizeof(z));
  return 32;
}
#endif

int dmb_launch_info(dmb_handle_t h, int32_t* grid, int32_t* block, int32_t* smem_bytes, int32_t* envs_per_cta) {
  if (!h) return DMB_ERR_ARG;
  if (grid) *grid = h->grid;
  if (block) *block = h->block;
  if (smem_bytes) *smem_bytes = h->smem;
  if (envs_per_cta) *envs_per_cta = h->envs_per_cta;
  return DMB_OK;
}

}  // extern "C"

#include "dmb_policy.cuh"
