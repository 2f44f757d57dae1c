/* dmb.h -- C-ABI of libdmb200.so: the B200-native batched DeepMimic/MuJoCo env hot path.
 *
 * Drop-in boundary for the per-step path of the reference env
 *   /root/reference/src/dp_env_v3.py:106-132  DPEnv.step
 *     -> gym MujocoEnv.do_simulation -> mujoco_py MjSim.step -> mj_step   (dp_env_v3.py:112)
 *     -> DPEnv._get_obs (62-65), DPEnv.is_done (134-139), calc_config_reward (89-104)
 *   /root/reference/src/dp_env_v3.py:148-164  reset_model / reset_model_init
 *     -> MujocoEnv.set_state + sim.forward()
 * i.e. what a maintainer would bind instead of mujoco_py's Cython `cymj` (see INTEGRATION.md
 * for the ctypes stub).  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *  - every function returns 0 on success or a negative dmb_status; the text of the last
 *    error is available from dmb_last_error() (per handle, or global for create failures);
 *  - all array arguments are DEVICE pointers into caller-owned buffers (PyTorch tensors)
 *    unless the name ends in _host; the library owns only the opaque handle (constant
 *    tables, launch configuration);
 *  - env-major rows, fp32: qpos[N][DMB_QSTRIDE], qvel[N][DMB_VSTRIDE], warm[N][DMB_VSTRIDE]
 *    (strides padded to 16 B multiples), action[N][nu], obs[N][obs_dim] (obs_dim = nq-7 + nv-6 = 56, or
 *    2 + 13*npart = 197 for the DeepMimic state, config.obs_mode 1);
 *  - all work is enqueued on the caller's CUDA stream (pass torch.cuda.current_stream()
 *    .cuda_stream, or NULL for the legacy default stream); calls are asynchronous w.r.t.
 *    the host except create/destroy;
 *  - a handle is bound to one CUDA device and is not thread-safe.
 */
#ifndef DMB_H_
#define DMB_H_

#include <stddef.h>
#include <stdint.h>

#include "dmb_model.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DMB_VERSION 2
#define DMB_QSTRIDE 36 /* floats per qpos row (nq = 35 padded) */
#define DMB_VSTRIDE 36 /* floats per qvel / warmstart row (nv = 34 padded) */
#define DMB_MAX_PEER 8 /* ranks of one NVLink domain served by the fused all-gather */

typedef enum dmb_status {
  DMB_OK = 0,
  DMB_ERR_ARG = -1,      /* bad argument (null pointer, size, mode) */
  DMB_ERR_CUDA = -2,     /* a CUDA runtime call failed */
  DMB_ERR_MODEL = -3,    /* model exceeds compiled capacities / unsupported feature */
  DMB_ERR_NO_DEVICE = -4 /* no CUDA device: there is NO CPU fallback */
} dmb_status;

typedef struct dmb_handle_s* dmb_handle_t;

/* Per-env state, caller-owned device buffers (all env-major). Replaces the mjData /
 * DPEnv attributes the reference keeps per env (sim.data.qpos/qvel/qacc_warmstart,
 * idx_init, idx_curr: dp_env_v3.py:55-57,67-71). */
typedef struct dmb_state {
  float* qpos;           /* [N][DMB_QSTRIDE] */
  float* qvel;           /* [N][DMB_VSTRIDE] */
  float* warm;           /* [N][DMB_VSTRIDE]  qacc_warmstart */
  int32_t* clip;         /* [N] motion clip id of each env */
  int32_t* idx_init;     /* [N] RSI frame (dp_env_v3.py:68) */
  int32_t* idx_curr;     /* [N] current mocap frame = phase (dp_env_v3.py:70,101-102) */
  uint32_t* reset_count; /* [N] Philox counter: number of resets so far */
  int32_t* ep_len;       /* [N] steps in the current episode */
  float* ep_ret;         /* [N] return of the current episode */
  int32_t* flags;        /* [N] bit0 contact overflow, bit1 row overflow, bit2 non-finite state (bits 8.. carry
                          * per-env diagnostics when DMB_TRACE=1, see dmb_get_trace) */
} dmb_state_t;

/* Outputs of one step (caller-owned device buffers). `rec` is optional: the packed
 * [N][obs_dim + 2] record (obs, reward, done-as-float) that is all-gathered across ranks. */
typedef struct dmb_step_out {
  float* obs;       /* [N][obs_dim] post-step (post-reset for auto-reset envs) observation */
  float* reward;    /* [N] */
  uint8_t* done;    /* [N] */
  float* rec;       /* [N][obs_dim + 2] or NULL */
  float* last_ret;  /* [N] return of the episode that just finished (valid where done) or NULL */
  int32_t* last_len;/* [N] length of the episode that just finished or NULL */
} dmb_step_out_t;

int dmb_version(void);
/* ABI guards for foreign-language bindings (ctypes mirrors in model_blob.py) */
int32_t dmb_sizeof_model(void);
int32_t dmb_sizeof_config(void);
int32_t dmb_sizeof_mocap(void);
int32_t dmb_sizeof_tile(void); /* bytes of shared memory per env (one warp) */

/* Build a handle: copies the model / config / mocap tables to `cuda_device` (fp32) and
 * sizes the launch.  `seed` keys the per-env Philox streams (env e uses (seed, first_env_id + e)). */
int dmb_create(const dmb_model_t* model, const dmb_config_t* config, const dmb_mocap_t* mocap,
               int32_t num_envs, int32_t cuda_device, uint64_t seed, uint32_t first_env_id,
               dmb_handle_t* out);
int dmb_destroy(dmb_handle_t h);

/* (Re)initialise envs.  mask: device uint8[N] or NULL (= all).  mode: 0 mocap RSI
 * (dp_env_v3.py:148-156), 1 init pose + U(-noise, noise) (dp_env_v3.py:158-164), -1 = config default.
 * Writes the post-reset observation to obs (may be NULL). */
int dmb_reset(dmb_handle_t h, const dmb_state_t* st, const uint8_t* mask, int32_t mode, float* obs, void* stream);

/* One env step for all N envs: action -> ctrl (PD optional) -> RK4 mj_step -> obs, reward,
 * done (+ in-kernel auto reset when config.auto_reset).  dp_env_v3.py:106-132. */
int dmb_step(dmb_handle_t h, const dmb_state_t* st, const float* action, const dmb_step_out_t* out, void* stream);

/* Observation of the current state: obs_mode 0 = qpos[7:] || qvel[6:] (dp_env_v3.py:62-65);
 * obs_mode 1 = DeepMimic state (code.md:287-504, mujoco_env.py:91-124).  obs is [N][dmb_obs_dim()]. */
int dmb_get_obs(dmb_handle_t h, const dmb_state_t* st, float* obs, void* stream);
int32_t dmb_obs_dim(dmb_handle_t h);

/* Interpolated mocap reference poses (phase_mode 1 arithmetic; replaces the kinematic playback of
 * MocapDM.play, mocap_v2.py:151-182, and transformations.quaternion_slerp, transformations.py:1270-1308):
 * for i < n, the pose of clip[i] (NULL = clip 0) at frame coordinate frame_coord[i] = t / clip_dt (device
 * float64) -> qpos[n][DMB_QSTRIDE], qvel[n][DMB_VSTRIDE], phase[n] in [0,1) (may be NULL). */
int dmb_mocap_sample(dmb_handle_t h, const int32_t* clip, const double* frame_coord, int32_t n, float* qpos, float* qvel,
                     float* phase, void* stream);

/* Stage-level debug: run ONE forward evaluation (mj_forward) at the current state with the
 * given ctrl[N][nu] and dump stage outputs into dbg[N][dmb_debug_stride()] (layout in
 * dmb_debug_layout).  Also updates st->warm like MuJoCo's mj_forward.  Used by the parity tests. */
int dmb_forward_debug(dmb_handle_t h, const dmb_state_t* st, const float* ctrl, float* dbg, void* stream);
int32_t dmb_debug_stride(void);
/* offsets (in floats) of the named sections inside one debug row; returns -1 for unknown names.
 * names: xpos xquat xipos com qM qLD qfrc_bias qfrc_smooth qacc_smooth ncon nefc iter contact
 *        efc_pos efc_R efc_aref efc_b efc_force efc_AR_diag qacc z_com cvel */
int32_t dmb_debug_offset(const char* name);

/* Fused all-gather of the step record over NVLink peer memory ("a single all-gather of (obs, reward,
 * done) per step"; replaces the ncclAllGather that would follow dmb_step).  Every rank allocates its gathered
 * [N_global][obs_dim + 2] buffers and int32 arrival flags with dmb_peer_alloc (cudaMalloc + CUDA IPC handle, zeroed),
 * exchanges the 64-byte handles (any out-of-band channel, e.g. torch.distributed.all_gather_object), maps the other
 * ranks' buffers with dmb_peer_open, and before a step names the n_peer (<= DMB_MAX_PEER, self included) target
 * buffers and flags: dmb_step then stores every env's record row into ALL of them at row row0 + env (posted peer
 * stores from the step kernel's epilogue, overlapping the rest of the kernel) and the last CTA adds 1 to every peer
 * flag (release at system scope).  A consumer enqueues dmb_peer_wait(flag, target) on its stream: one polling thread
 * that returns once `target` ranks' steps have arrived (acquire).  No collective kernel, no SMs taken from the step
 * kernel, no rank waits for another unless it consumes.  n_peer = 0 turns it off. */
int dmb_peer_alloc(int32_t cuda_device, uint64_t bytes, void** ptr, uint8_t* handle64);
int dmb_peer_open(int32_t cuda_device, const uint8_t* handle64, void** ptr);
int dmb_peer_close(void* ptr);
int dmb_peer_free(void* ptr);
int dmb_set_peer_gather(dmb_handle_t h, int32_t n_peer, float* const* rec_peer, int32_t* const* flag_peer, int32_t row0);
int dmb_peer_wait(dmb_handle_t h, const int32_t* flag, int32_t target, void* stream);
/* the same wait folded into the NEXT dmb_step (one thread of its first CTA polls at kernel start, no extra launch): the
 * completion of that step then implies the arrival; flag = NULL cancels */
int dmb_set_peer_wait(dmb_handle_t h, const int32_t* flag, int32_t target);

/* Host-mapped I/O for callers whose policy lives on the host (the reference's own use: trpo.py:49 computes the action
 * on the CPU and reads ob/rew/new back every step, dp_env_v3.py:106-132).  `action` and `out->rec` of dmb_step may point
 * to PINNED, device-mapped host memory: the step kernel then loads each env's action row and stores its record row
 * directly over PCIe (one 112-byte read and 232 bytes of posted writes per env, overlapping the kernel), so a step is
 * ONE launch instead of cudaMemcpyAsync + kernel + cudaMemcpyAsync.  The host may read the record after it has
 * synchronised with the stream.  dmb_host_alloc returns such a buffer (cudaHostAlloc mapped + portable, zeroed) with
 * its host and device addresses; dmb_host_device_pointer resolves the device address of a buffer the caller pinned
 * itself (cudaHostAlloc / cudaHostRegister, e.g. torch's pin_memory()).  All other arrays stay device memory. */
int dmb_host_alloc(int32_t cuda_device, uint64_t bytes, void** host_ptr, void** dev_ptr);
int dmb_host_free(void* host_ptr);
int dmb_host_device_pointer(int32_t cuda_device, void* host_ptr, void** dev_ptr);

/* kernels launched by dmb_step through this handle so far (bench.py reports the count of its timed region) */
int64_t dmb_kernel_launches(dmb_handle_t h);

/* launch geometry chosen at create time (for DESIGN.md / bench reporting) */
int dmb_launch_info(dmb_handle_t h, int32_t* grid, int32_t* block, int32_t* smem_bytes, int32_t* envs_per_cta);

/* Diagnostics (DMB_TRACE=1 in the environment at create time): per-CTA timeline of the last dmb_step,
 * host int64 [grid][8]: [0] globaltimer ns at kernel start, [1..5] when warp 0 finished its 1st..5th scheduler
 * round, [6] when the CTA's last warp was done, [7] most PGS sweeps of one env (summed over the RK stages) |
 * largest row count << 16 | envs that used the global scratch rows << 24; returns the number of CTAs written
 * (0 when tracing is off).  Synchronises the device. */
int32_t dmb_get_trace(dmb_handle_t h, int64_t* host_out, int32_t max_ctas);

const char* dmb_last_error(dmb_handle_t h);

#ifdef __cplusplus
}
#endif
#endif /* DMB_H_ */
