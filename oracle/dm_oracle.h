/* dm_oracle.h -- CPU float64 ORACLE for the batched DeepMimic/MuJoCo hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (deepmimic_mujoco_b200/) may
 * import, link or execute this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in the closed-source
 * MuJoCo 2.0 binary reached through mujoco-py (/root/reference/src/dp_env_v3.py:10-13,112)
 * which is neither vendored in /root/reference nor installable here, and the reference
 * ships no golden vectors for qpos/qvel/contacts/reward.  This file
 * restates MuJoCo's published computation pipeline (mj_step with RK4 + PGS, as
 * documented for the open-sourced >=2.1 engine) for the MJCF subset of dp_env_v3.xml,
 * plus the env logic of dp_env_v3.py.  Deviations are listed in DESIGN.md.
 * The only MuJoCo-PRODUCED data in the reference are the episode monitor of its training run
 * (src/log_tmp/DeepMimic/trpo-walk-0/monitor.json.monitor.csv) and the policy that run trained
 * (src/checkpoint_tmp/DeepMimic/trpo-walk-0, a TensorFlow checkpoint).  The oracle reproduces
 * (a) the time-to-fall distribution of the initial random policy (tests/golden/ref_episode_lengths.json)
 * and (b) the ~290-step survival of the trained policy replayed from the checkpoint
 * (tests/golden/ref_trained_policy.npz; oracle / MuJoCo = 0.94 +- 0.03), also when driven by the
 * reference's own rollout loop, env class and monitor (tools/reference_protocol_replay.py) --
 * tests/test_oracle_physics.py, tests/test_reference_protocol_replay.py.  These are statistical
 * pins of the whole pipeline, not state-level ones, so the status above stands.  The ENV LOGIC
 * (dmo_env_*) is pinned state for state against the reference class dp_env_v3.DPEnv run over an
 * adapter (tests/test_env_logic_golden.py).
 */
#ifndef DM_ORACLE_H_
#define DM_ORACLE_H_

#include "../include/dmb_model.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DMO_MAXCON 64
#define DMO_MAXEFC 192

typedef struct dmo_contact {
  double dist;
  double pos[3];
  double frame[9]; /* row 0 = normal (geom1 -> geom2), rows 1,2 tangents */
  double mu;
  int32_t geom1, geom2, dim, efc_address;
} dmo_contact_t;

/* Everything one env owns; all stage outputs are kept for stage-level parity diffs. */
typedef struct dmo_data {
  /* state */
  double qpos[DMB_MAX_Q], qvel[DMB_MAX_DOF], ctrl[DMB_MAX_U], qacc_warmstart[DMB_MAX_DOF];
  /* position stage */
  double xpos[DMB_MAX_BODY][3], xquat[DMB_MAX_BODY][4], xmat[DMB_MAX_BODY][9], xipos[DMB_MAX_BODY][3];
  double xaxis[DMB_MAX_JNT][3];
  double geom_xpos[DMB_MAX_GEOM][3], geom_xmat[DMB_MAX_GEOM][9];
  double com[3];
  double cinert[DMB_MAX_BODY][10], crb[DMB_MAX_BODY][10], cdof[DMB_MAX_DOF][6];
  double qM[DMB_MAX_M], qLD[DMB_MAX_M], qLDiagInv[DMB_MAX_DOF];
  /* velocity stage */
  double cvel[DMB_MAX_BODY][6], cdof_dot[DMB_MAX_DOF][6];
  double qfrc_bias[DMB_MAX_DOF], qfrc_passive[DMB_MAX_DOF], qfrc_actuator[DMB_MAX_DOF];
  double qfrc_smooth[DMB_MAX_DOF], qacc_smooth[DMB_MAX_DOF];
  /* collision + constraints */
  int32_t ncon, nefc, solver_iter, flags; /* flags: 1 contact overflow, 2 row overflow, 4 non-finite */
  dmo_contact_t contact[DMO_MAXCON];
  int32_t efc_type[DMO_MAXEFC], efc_id[DMO_MAXEFC]; /* type: 0 limit, 1 frictionless, 2 pyramidal */
  double efc_J[DMO_MAXEFC][DMB_MAX_DOF];
  double efc_pos[DMO_MAXEFC], efc_margin[DMO_MAXEFC], efc_diagApprox[DMO_MAXEFC];
  double efc_R[DMO_MAXEFC], efc_D[DMO_MAXEFC], efc_KBIP[DMO_MAXEFC][4];
  double efc_vel[DMO_MAXEFC], efc_aref[DMO_MAXEFC], efc_b[DMO_MAXEFC], efc_force[DMO_MAXEFC];
  double efc_AR[DMO_MAXEFC][DMO_MAXEFC];
  double qfrc_constraint[DMB_MAX_DOF], qacc[DMB_MAX_DOF];
} dmo_data_t;

/* Per-env bookkeeping of the imitation env (dp_env_v3.py DPEnv attributes). */
typedef struct dmo_env {
  dmo_data_t d;
  int32_t clip, idx_init, idx_curr, ep_len;
  uint64_t seed;       /* Philox key */
  uint32_t env_id, reset_count;
  double ep_ret;
  double reward_terms[5];
  double zcom_last;    /* CoM height the last dmo_env_step tested for termination */
} dmo_env_t;

int dmo_version(void);
unsigned long dmo_sizeof_data(void);
unsigned long dmo_sizeof_env(void);
unsigned long dmo_sizeof_model(void);

/* physics (App. B of SURVEY.md): */
void dmo_fwd_position(const dmb_model_t* m, dmo_data_t* d);   /* kinematics, com, crb, factor, collision, constraints */
void dmo_fwd_velocity(const dmb_model_t* m, dmo_data_t* d);   /* comVel, passive, rne bias, efc_vel/aref */
void dmo_fwd_actuation(const dmb_model_t* m, dmo_data_t* d);
void dmo_fwd_acceleration(const dmb_model_t* m, dmo_data_t* d);
void dmo_fwd_constraint(const dmb_model_t* m, dmo_data_t* d); /* PGS */
void dmo_forward(const dmb_model_t* m, dmo_data_t* d);        /* mj_forward */
void dmo_step(const dmb_model_t* m, dmo_data_t* d);           /* mj_step (RK4) */
void dmo_kinematics(const dmb_model_t* m, dmo_data_t* d);
void dmo_solve_M(const dmb_model_t* m, const dmo_data_t* d, double* x); /* x <- M^-1 x */
void dmo_full_M(const dmb_model_t* m, const dmo_data_t* d, double* dense /* nv*nv */);

/* env logic (dp_env_v3.py): */
void dmo_env_init(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e,
                  uint64_t seed, uint32_t env_id, int32_t clip);
void dmo_env_reset(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e, int mode);
void dmo_env_set_state(const dmb_model_t* m, dmo_env_t* e, const double* qpos, const double* qvel);
/* one env step: action[nu] -> obs[nq-7+nv-6], reward, done.  Returns done. */
int dmo_env_step(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e,
                 const double* action, double* obs, double* reward);
void dmo_env_obs(const dmb_model_t* m, const dmo_env_t* e, double* obs);
/* DeepMimic 197-d state of the current state (obs_mode 1); returns its length 2 + 13*npart */
int dmo_env_obs_dm(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e, double* obs);
/* interpolated reference pose at frame coordinate u = t / clip_dt (phase_mode 1) */
void dmo_mocap_sample(const dmb_model_t* m, const dmb_mocap_t* mc, int clip, double u, double* qpos, double* qvel,
                      double* phase);
/* reference-pose extras for one mocap frame (fills DMB_REF_AUX doubles) */
void dmo_ref_aux(const dmb_model_t* m, const double* qpos, const double* qvel, double* aux);
/* counter-based RNG shared bit-for-bit with the CUDA kernels */
void dmo_philox(uint64_t seed, uint32_t env_id, uint32_t reset_count, uint32_t block, uint32_t out[4]);
/* timed rollout used as the CPU baseline: nsteps random-action env steps with auto reset;
 * returns number of env steps done. */
long dmo_rollout(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, dmo_env_t* e,
                 long nsteps, uint64_t action_seed);

/* one env step of n independent envs from explicit inputs (full-size parity tests); see dm_oracle.c */
void dmo_batch_step(const dmb_model_t* m, const dmb_config_t* cfg, const dmb_mocap_t* mc, int n, uint64_t seed,
                    uint32_t first_env_id, double* qpos, double* qvel, double* warm, const int32_t* clip,
                    int32_t* idx_init, int32_t* idx_curr, int32_t* ep_len, double* ep_ret, int32_t* reset_count,
                    const double* action, double* obs, int obs_stride, double* reward, int32_t* done,
                    double* last_ret, int32_t* last_len, int32_t* flags, int32_t* nefc_last, double* zcom);

#ifdef __cplusplus
}
#endif
#endif
