// dmb_types.cuh -- device-side (fp32) model tables and per-env shared-memory tile.
//
// One warp owns one env.  Everything an env needs during one RK4 step lives in its EnvS
// tile in shared memory; the fp32 model (ModelS) is copied once per CTA into shared memory
// because most of its tables are indexed per lane (body / dof / geom ids), which the
// constant cache would serialise.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/dmb.h"

namespace dmb {

constexpr int NB = DMB_MAX_BODY;  // 16
constexpr int NJ = DMB_MAX_JNT;   // 32
constexpr int NVC = 36;           // dof capacity of the kernel (nv <= 36; humanoid: 34)
constexpr int NQC = 36;           // qpos capacity (nq <= 36; humanoid: 35)
constexpr int NG = DMB_MAX_GEOM;  // 16
constexpr int NP = DMB_MAX_PAIR;  // 128
constexpr int NU = DMB_MAX_U;     // 32
constexpr int NMX = DMB_MAX_M;    // 320
constexpr int NMT = 312;          // sparse-inertia capacity of the tile (nM <= 312; humanoid: 310)
constexpr int NVT = 34;           // dof capacity of the tile arrays that need no 16-byte alignment
constexpr int MAXANC = 12;        // longest dof ancestor chain (humanoid: 12)
constexpr int MAXROW = 40;        // constraint-row capacity of the kernel (max_efc <= 40)
constexpr int MAXC = 16;          // contact capacity per env (max_con <= 16)
constexpr int YS = 35;            // row stride of Y (odd -> conflict-free lane=row access; >= nv)
constexpr int NTRI = MAXROW * (MAXROW + 1) / 2;  // packed lower triangle of AR
constexpr int JPB = 3;            // joints per body capacity

enum DofKind : int8_t { DOF_FREE_TRANS = 0, DOF_FREE_ROT = 1, DOF_HINGE = 2 };

struct ModelS {
  // sizes / options
  int nq, nv, nu, nbody, njnt, ngeom, npair, nM;
  int iterations, max_con, max_efc, maxdepth;
  int nclip, nframe_total, nee, sync_mask;  // sync_mask: which lockstep phase barriers are active
  int ngroups, cost_mode, pad_k0, pad_k1;   // ngroups: lockstep groups per CTA (each with its own named barrier)
  float timestep, tolerance, pgs_scale, margin;
  float gravity[3], inv_total_mass;
  float imp_k, imp_b;        // reference spring constants after refsafe (mj_makeImpedance)
  float solimp[5], pad_f;
  // bodies
  int8_t body_parent[NB], body_depth[NB], body_jntadr[NB], body_jntnum[NB], body_dofadr[NB], body_dofnum[NB];
  int8_t body_nchild[NB], body_child[NB][4];
  float body_pos[NB][3], body_quat[NB][4], body_ipos[NB][3], body_inertia[NB][6], body_mass[NB], body_invw[NB];
  unsigned long long body_dofmask[NB];  // dofs on the chain world..b (inclusive)
  // joints
  int8_t jnt_type[NJ], jnt_qposadr[NJ], jnt_dofadr[NJ], jnt_limited[NJ], jnt_bodyid[NJ];
  float jnt_axis[NJ][3], jnt_range[NJ][2], jnt_qpos0[NJ];
  // dofs
  int8_t dof_bodyid[NVC], dof_kind[NVC], dof_axisk[NVC], dof_nanc[NVC], dof_anc[NVC][MAXANC], dof_act[NVC];
  int16_t dof_Madr[NVC];
  int16_t anc_rowbase[NVC][MAXANC];     // dof_Madr of the p-th ancestor of each dof
  unsigned long long dof_ancmask[NVC];  // strict ancestors of each dof
  // depth-first dof numbering: the descendants of dof d are the ids d+1 .. d+dof_ndesc[d]; row d of the factor
  // stores L[d][a] at qLD[dof_Lend[d] - depth(a)] with dof_Lend[d] = dof_Madr[d] + dof_nanc[d]
  uint8_t dof_ancr[NVC][MAXANC];        // ancestor of d at depth r (root side first)
  int8_t dof_ndesc[NVC];
  int16_t dof_Lend[NVC];
  int maxanc, spread;                   // spread: deal the sorted env list round robin over the CTAs when one round suffices
  uint32_t ldl_meta[NVC];               // L'DL step k: chain length | pair count << 8 | row address << 16
  // chain prefix sums by pointer jumping: dof_jump[s][d] = the 2^s-th ancestor of dof d (-1: none);
  // dof_vsrc[d] = dof whose inclusive chain sum is the velocity seen by cdof_dot[d] (mj_comVel; -1: zero);
  // dof_lastof[d] = body whose last dof is d (-1: d is not the last dof of its body)
  int8_t dof_jump[4][NVC], dof_vsrc[NVC], dof_lastof[NVC];
  int8_t body_jump[3][NB];              // 1st, 2nd, 4th ancestor body of each body (-1: none / world)
  float dof_armature[NVC], dof_damping[NVC], dof_invw[NVC], dof_gear[NVC], dof_ctrl_lo[NVC], dof_ctrl_hi[NVC];
  float dof_kp[NVC], dof_kd[NVC], dof_weight[NVC];
  // inertia entries
  uint8_t M_i[NMX], M_j[NMX];
  uint8_t tri_p[80], tri_q[80];  // pair decode for the sparse L'DL update (q-major)
  // geoms
  int8_t geom_type[NG], geom_bodyid[NG], geom_condim[NG], geom_identq[NG];
  float geom_size[NG][3], geom_pos[NG][3], geom_quat[NG][4], geom_rbound[NG], geom_mu[NG];
  uint8_t pair_g1[NP], pair_g2[NP];
  // actuators
  int8_t act_dofadr[NU];
  // env config
  int ctrl_mode, reward_mode, reset_mode, auto_reset;
  int term_mode; unsigned fall_body_mask; int phase_mode, obs_mode;
  int npart, obs_dim, pad_t0, pad_t1;
  int8_t part_geom[DMB_MAX_PART];
  double clip_rate[DMB_MAX_CLIP];  // env steps -> mocap frames: timestep / clip_dt (phase_mode 1)
  float z_min, z_max, reset_noise, pd_dt;
  float joint_weight_sum, pad_w0, pad_w1, pad_w2;
  float w_pose, w_vel, w_ee, w_root, w_com, s_pose, s_vel, s_ee, s_root, s_com, s_err, pad_g;
  // reference pose
  float qpos0[NQC];
  // end effectors + clips
  int ee_body[DMB_MAX_EE];
  float ee_pos[DMB_MAX_EE][3];
  int clip_start[DMB_MAX_CLIP], clip_len[DMB_MAX_CLIP];
};

// Per-env tile (fp32), 7.9 KB: 28 tiles + the model tables fill the 227 KB of one SM, so that all 4096 envs of
// the benchmark batch are resident at once (148 SMs x 28 warps).  `Fixed` lives for the whole env step; the overlay
// `o` is time-shared by the phases of one forward evaluation ("shared-memory tile"):
//   PhaseK  kinematics / inertia / RNE temporaries, cdof, cvel
//   PhaseC  geom poses + broad-phase survivors + contact geometry (xpos/xquat/xmat/cdof/cvel stay where PhaseK put them)
//   PhaseR  the half-solved constraint Jacobian Y, per-row scalars and the packed Delassus matrix AR for up to
//           RF rows.  Y is written while cdof / contact geometry are still read (they lie behind it); the row
//           scalars and AR are written after those are dead.
// Envs with more than RF rows in a stage (about 0.1 % of the benchmark's stage evaluations) keep Y in a per-warp
// scratch slot in global memory and the row scalars / AR in PhaseRbig (same code, template parameter OVF).
#ifndef DMB_RF
#define DMB_RF 24
#endif
constexpr int RF = DMB_RF;        // rows held in the tile
constexpr int NTRI_F = RF * (RF + 1) / 2;

struct PhaseK {
  float xpos[NB * 3], xquat[NB * 4], xmat[NB * 9], xipos[NB * 3];   // [0, 304)
  float cinert[NB * 10], crb[NB * 10];                              // [304, 624)
  union {                                                            // [624, 840)
    float buf6[NVC * 6];                                             //   crb * cdof (M entries); cdof * qvel (features)
    struct { float acc[NB * 6], cfrc[NB * 6]; };                     //   RNE body accelerations / forces
  };
  float cdof[NVC * 6], cvel[NB * 6];                                 // [840, 1152)
};
struct PhaseC {
  float xpos[NB * 3], xquat[NB * 4], xmat[NB * 9];                   // as PhaseK
  float gpos[NG * 3], gmat[NG * 9];                                  // [256, 448)
  unsigned char surv[NP];                                            // [448, 480) broad-phase survivors (pair ids)
  float pad[360];
  float cdof[NVC * 6], cvel[NB * 6];                                 // as PhaseK
  float c_dist[MAXC], c_pos[MAXC * 3], c_frame[MAXC * 6];            // [1152, 1312) contact frame: normal, tangent 1
};
struct PhaseR {
  float Y[RF * YS];                                                  // [0, 840)
  float e_R[RF], e_aref[RF], e_b[RF], e_f[RF];                       // [840, 936)
  unsigned long long rowmask[RF];                                    // [936, 984) dof support of each half-solved row
  float AR[NTRI_F];                                                  // [984, 1284) packed lower triangle of J M^-1 J' + R
};
// Stage with more than RF rows: the half-solved Jacobian Y (MAXROW x YS floats) does not fit the tile and lives in
// the warp's global scratch slot; the latency-critical row scalars and the Delassus matrix (PGS reads a column per
// row update) stay in shared memory, in the part of the overlay that Y leaves free.
struct PhaseRbig {
  float e_R[MAXROW], e_aref[MAXROW], e_b[MAXROW], e_f[MAXROW];      // [0, 160)
  unsigned long long rowmask[MAXROW];                                // [160, 240)
  float AR[NTRI];                                                    // [240, 1060): over cdof, written after J is built
};
union Overlay {
  PhaseK k;
  PhaseC c;
  PhaseR r;
  PhaseRbig rb;
};
static_assert(offsetof(PhaseRbig, AR) <= offsetof(PhaseK, cdof),
              "the big-row scalars are written by the row stage, which still reads cvel and the contact geometry (AR is "
              "written later, by the Gram stage, and may lie over them)");
static_assert(offsetof(PhaseRbig, rowmask) % 8 == 0, "alignment");
static_assert(offsetof(PhaseC, cdof) == offsetof(PhaseK, cdof), "cdof must not move between phases");
static_assert(offsetof(PhaseC, xmat) == offsetof(PhaseK, xmat), "xmat must not move between phases");
static_assert(offsetof(PhaseR, e_R) >= offsetof(PhaseK, cdof), "row scalars may only overwrite cdof / cvel");
static_assert(sizeof(float) * RF * YS <= offsetof(PhaseK, cdof), "Y is written while cdof is read");
static_assert(offsetof(PhaseR, AR) % 4 == 0 && offsetof(PhaseR, rowmask) % 8 == 0, "alignment");

struct EnvS {
  float qpos[NQC], qvel[NVC], ctrlf[NVC];
  float qacc[NVC];                     // also qacc_warmstart: mj_forward leaves warmstart = qacc
  float x_q0[NQC - 1], x_dv[NVC];      // RK4: X0 positions (nq <= 35), stage velocity; x_dv doubles as scratch inside forward_eval
  float qLD[NMT], dsq[NVT], ys[NVT];   // sparse factor, D^-1/2, y_s = D^-1/2 L^-T qfrc_smooth
  float com[3];
  int diag;                            // diagnostics: PGS sweeps summed over the RK stages | largest row count << 16
  int ncon, nefc, nlimit, flags, iter, cost;
  int big;                             // stage with > RF rows: 1 = Y in the CTA's shared slot, 2 = in global scratch
  int c_meta[MAXC];                    // geom1 | geom2 << 8 | condim << 16 | first row << 24
  signed char e_src[RF];               // row source: >=0 contact*4+edge, <0 joint limit (see count_rows)
  Overlay o;
};
static_assert(sizeof(EnvS) % 16 == 0, "tiles start on 16-byte boundaries (float4 state rows)");
__device__ __forceinline__ int cm_g1(int m) { return m & 0xff; }
__device__ __forceinline__ int cm_g2(int m) { return (m >> 8) & 0xff; }
__device__ __forceinline__ int cm_dim(int m) { return (m >> 16) & 0xff; }
__device__ __forceinline__ int cm_adr(int m) { return (m >> 24) & 0xff; }

// Global scratch slot of one warp for a stage with more than RF rows (floats)
namespace gs {
constexpr int Y = 0;                          // MAXROW * YS
constexpr int e_src = Y + MAXROW * YS;        // MAXROW bytes (10 words)
constexpr int stride = ((e_src + MAXROW / 4 + 3) / 4) * 4;
}  // namespace gs

// Debug row layout (floats) for dmb_forward_debug
namespace dbg {
constexpr int xpos = 0;                     // NB*3
constexpr int xquat = xpos + NB * 3;        // NB*4
constexpr int xipos = xquat + NB * 4;       // NB*3
constexpr int com = xipos + NB * 3;         // 4
constexpr int qM = com + 4;                 // NMX
constexpr int qLD = qM + NMX;               // NMX
constexpr int qfrc_bias = qLD + NMX;        // NQC
constexpr int qfrc_smooth = qfrc_bias + NQC;
constexpr int qacc_smooth = qfrc_smooth + NQC;
constexpr int ncon = qacc_smooth + NQC;     // 1
constexpr int nefc = ncon + 1;
constexpr int iter = nefc + 1;
constexpr int z_com = iter + 1;
constexpr int contact = z_com + 1;          // MAXC * 16: dist, pos3, frame9, g1, g2, dim
constexpr int efc_pos = contact + MAXC * 16;
constexpr int efc_R = efc_pos + MAXROW;
constexpr int efc_aref = efc_R + MAXROW;
constexpr int efc_b = efc_aref + MAXROW;
constexpr int efc_force = efc_b + MAXROW;
constexpr int efc_AR_diag = efc_force + MAXROW;
constexpr int qacc = efc_AR_diag + MAXROW;  // NQC
constexpr int cvel = qacc + NQC;            // NB*6
constexpr int stride = cvel + NB * 6;
}  // namespace dbg

}  // namespace dmb
