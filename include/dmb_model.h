/* dmb_model.h -- compiled-model and mocap table structs shared by the CUDA library
 * (libdmb200.so, include/dmb.h) and the CPU oracle (oracle/dm_oracle.c).
 *
 * These are plain-old-data, fixed-capacity, float64 tables produced by the host-side
 * model compiler (deepmimic_mujoco_b200/mjcf.py) from the reference's MJCF
 * (/root/reference/src/mujoco/humanoid_deepmimic/envs/asset/dp_env_v3.xml) -- they
 * replace the mjModel that mujoco_py.load_model_from_path builds for
 * /root/reference/src/dp_env_v3.py:59.  The CUDA side converts to fp32 at create time.
 */
#ifndef DMB_MODEL_H_
#define DMB_MODEL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMB_MAX_BODY 16
#define DMB_MAX_JNT 32
#define DMB_MAX_DOF 40
#define DMB_MAX_Q 40
#define DMB_MAX_GEOM 16
#define DMB_MAX_PAIR 128
#define DMB_MAX_U 32
#define DMB_MAX_M 320
#define DMB_MAX_CLIP 16
#define DMB_MAX_EE 4
#define DMB_MAX_PART 16

/* geom / joint type ids (MuJoCo's numbering) */
#define DMB_GEOM_PLANE 0
#define DMB_GEOM_SPHERE 2
#define DMB_GEOM_CAPSULE 3
#define DMB_GEOM_BOX 6
#define DMB_JNT_FREE 0
#define DMB_JNT_HINGE 3

typedef struct dmb_model {
  /* sizes */
  int32_t nq, nv, nu, nbody, njnt, ngeom, npair, nM;
  /* solver / integrator options (dp_env_v3.xml:9 + MuJoCo defaults) */
  int32_t iterations;   /* PGS sweeps (50) */
  int32_t max_con;      /* contact buffer capacity per env (overflow -> flag) */
  int32_t max_efc;      /* constraint-row capacity per env */
  int32_t pad0;
  double timestep, tolerance, meaninertia, margin;
  double gravity[3];
  double solref[2];
  double solimp[5];
  /* bodies */
  int32_t body_parent[DMB_MAX_BODY], body_depth[DMB_MAX_BODY];
  int32_t body_jntadr[DMB_MAX_BODY], body_jntnum[DMB_MAX_BODY];
  int32_t body_dofadr[DMB_MAX_BODY], body_dofnum[DMB_MAX_BODY];
  double body_pos[DMB_MAX_BODY][3], body_quat[DMB_MAX_BODY][4];
  double body_ipos[DMB_MAX_BODY][3];
  double body_inertia[DMB_MAX_BODY][6]; /* xx yy zz xy xz yz about COM, body frame */
  double body_mass[DMB_MAX_BODY];
  double body_invweight0[DMB_MAX_BODY][2];
  /* joints */
  int32_t jnt_type[DMB_MAX_JNT], jnt_bodyid[DMB_MAX_JNT];
  int32_t jnt_qposadr[DMB_MAX_JNT], jnt_dofadr[DMB_MAX_JNT], jnt_limited[DMB_MAX_JNT];
  double jnt_axis[DMB_MAX_JNT][3], jnt_range[DMB_MAX_JNT][2];
  /* dofs */
  int32_t dof_bodyid[DMB_MAX_DOF], dof_jntid[DMB_MAX_DOF];
  int32_t dof_parentid[DMB_MAX_DOF], dof_Madr[DMB_MAX_DOF];
  double dof_armature[DMB_MAX_DOF], dof_damping[DMB_MAX_DOF], dof_invweight0[DMB_MAX_DOF];
  /* geoms */
  int32_t geom_type[DMB_MAX_GEOM], geom_bodyid[DMB_MAX_GEOM], geom_condim[DMB_MAX_GEOM];
  double geom_size[DMB_MAX_GEOM][3], geom_pos[DMB_MAX_GEOM][3], geom_quat[DMB_MAX_GEOM][4];
  double geom_rbound[DMB_MAX_GEOM], geom_friction[DMB_MAX_GEOM][3];
  /* collision candidates, in MuJoCo's emission order */
  int32_t pair_geom1[DMB_MAX_PAIR], pair_geom2[DMB_MAX_PAIR];
  /* actuators (motors on hinge dofs) */
  int32_t act_dofadr[DMB_MAX_U];
  double act_gear[DMB_MAX_U], act_ctrlrange[DMB_MAX_U][2];
  /* PD gains per actuator (mocap_util.py:22-24 PARAMS_KP_KD) */
  double act_kp[DMB_MAX_U], act_kd[DMB_MAX_U];
  /* reference configuration */
  double qpos0[DMB_MAX_Q];
  /* imitation-reward constants */
  double dof_weight[DMB_MAX_DOF]; /* normalised DeepMimic joint weight per dof (mocap_util.py:26-29) */
  int32_t ee_body[DMB_MAX_EE];    /* end-effector points: body id + local offset */
  int32_t nee, pad1[3];
  double ee_pos[DMB_MAX_EE][3];
  /* DeepMimic body parts for the 197-d state (obs_mode 1): geom id of each part, in the part order of
   * src/data/characters/humanoid3d.txt (root, chest, neck, r_hip, r_knee, r_ankle, r_shoulder, r_elbow,
   * r_wrist, l_hip, ..., l_wrist); a part's position is its geom centre, its rotation the owning body's */
  int32_t npart, pad2;
  int32_t part_geom[DMB_MAX_PART];
} dmb_model_t;

/* Environment-level configuration (reward / control / termination / reset modes). */
typedef struct dmb_config {
  int32_t ctrl_mode;   /* 0: action is motor ctrl (dp_env_v3.py:112); 1: PD "reference intent"
                          (mujoco_interface.py:97-107); 2: plain PD  kp*(a-q) - kd*qvel */
  int32_t reward_mode; /* 0: 1.0 (dp_env_v3.py:117); 1: exp(-L1) (dp_env_v3.py:89-104);
                          2: v2 pose reward - 0.1|a|^2 (dp_env_v2.py:116-183); 3: v1 3-term reward
                          - 0.1|a|^2 (dp_env_v1.py:82-152, mujoco_interface.py:169-210);
                          4: 5-term DeepMimic (code.md:979-1146) */
  int32_t reset_mode;  /* 0: mocap RSI (dp_env_v3.py:148-156); 1: init pose + U(-.01,.01) (158-164) */
  int32_t auto_reset;  /* 1: done envs are re-initialised inside step (vec_env semantics) */
  int32_t term_mode;   /* 0: CoM height only (dp_env_v3.py:134-139); 1: + DeepMimic fall-contact rule
                          (--fall_contact_bodies, src/args/train_humanoid3d_walk_args.txt:20) */
  uint32_t fall_body_mask; /* bit b set: a floor contact of body b ends the episode (all but the ankles) */
  int32_t phase_mode;  /* 0: integer frame stepping, one frame per env step (dp_env_v3.py:101-102);
                          1: time-based phase, reference pose interpolated between frames (lerp / slerp,
                          transformations.py:1270-1308) with loop wrap + root-offset accumulation
                          (mocap_v2.py:168-182); reference time = idx_init*clip_dt + steps*timestep */
  int32_t obs_mode;    /* 0: qpos[7:] || qvel[6:] (dp_env_v3.py:62-65);
                          1: DeepMimic state (code.md:287-504, mujoco_env.py:91-124): phase, root height,
                          npart x (pos 3 + quat 4), npart x (lin vel 3 + ang vel 3) in the root heading frame */
  double z_min, z_max; /* CoM-height termination band (dp_env_v3.py:134-139): 0.7, 2.0 */
  double reset_noise;  /* 0.01 */
  double joint_weight_sum; /* sum of the raw DeepMimic joint weights (mocap_util.py:26-29): 4.8 */
  double w_pose, w_vel, w_end_eff, w_root, w_com;          /* dp_env_v3.py:42-46 */
  double s_pose, s_vel, s_end_eff, s_root, s_com, s_err;   /* dp_env_v3.py:48-53 */
} dmb_config_t;

/* Compiled motion clips (mocap_v2.py:78-149 output, concatenated). Arrays are
 * row-major float64 owned by the caller and copied at create time. */
#define DMB_REF_AUX 24 /* per-frame reference extras for the 5-term reward:
                          [0:12] 4 end-effector points in the root heading frame,
                          [12:15] CoM velocity, [15:19] root quat, [19:24] pad */
typedef struct dmb_mocap {
  int32_t nclip, nframe_total;
  int32_t clip_start[DMB_MAX_CLIP], clip_len[DMB_MAX_CLIP];
  double clip_dt[DMB_MAX_CLIP];
  const double* data_config; /* [nframe_total][nq] */
  const double* data_vel;    /* [nframe_total][nv] */
  const double* ref_aux;     /* [nframe_total][DMB_REF_AUX] */
} dmb_mocap_t;

#ifdef __cplusplus
}
#endif
#endif /* DMB_MODEL_H_ */
