/* dmb_policy.h -- C-ABI of the batched policy-inference entry point of libdmb200.so.
 *
 * SURVEY.md section 8(f) rank 1 ("next" row): replaces the batch-1 TensorFlow `pi.act` call in the
 * rollout loop (/root/reference/src/trpo.py:49) with one fused kernel over all envs:
 *   obz = clip((ob - mean) / std, -5, 5)                  (mlp_policy_trpo.py:33, misc_util.py:50-51)
 *   vpred = vffinal(tanh(vffc2(tanh(vffc1(obz)))))        (mlp_policy_trpo.py:35-37)
 *   mean  = polfinal(tanh(polfc2(tanh(polfc1(obz)))))     (mlp_policy_trpo.py:39-44)
 *   ac    = mean + exp(logstd) * N(0,1)   if stochastic else mean   (distributions.py DiagGaussianPd)
 * All pointers are device pointers (fp32); dense kernels are row-major [in][out] like tf dense.
 */
#ifndef DMB_POLICY_H_
#define DMB_POLICY_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmb_policy {
  int32_t obs_dim, act_dim, hid, pad;     /* 56, 28, 100 */
  const float* ob_mean; const float* ob_std;          /* [obs_dim] running mean / std of the obs filter */
  const float* pw1; const float* pb1;                 /* [obs_dim][hid], [hid] */
  const float* pw2; const float* pb2;                 /* [hid][hid], [hid] */
  const float* pw3; const float* pb3;                 /* [hid][act_dim], [act_dim] */
  const float* logstd;                                /* [act_dim] */
  const float* vw1; const float* vb1;
  const float* vw2; const float* vb2;
  const float* vw3; const float* vb3;                 /* [hid][1], [1] */
} dmb_policy_t;

/* ac[n][act_dim], vpred[n], mean_out[n][act_dim] (may be NULL).  Gaussian noise comes from
 * Philox4x32-10 keyed by (seed), counter (first_row + row, step, unit/4) and Box-Muller, so it is
 * reproducible and independent of the batch partitioning.  Returns 0 or a negative dmb_status. */
int dmb_policy_act(const dmb_policy_t* p, const float* obs, int32_t n, int32_t stochastic, uint64_t seed,
                   uint32_t step, uint32_t first_row, float* ac, float* vpred, float* mean_out, void* stream);

/* GAE(lambda) over a [T][n] rollout segment on the device (replaces the numpy loop of trpo.py:83-94
 * add_vtarg_and_adv): rew, vpred, isnew are [T][n] fp32 (isnew[t] = 1 when step t starts a new episode), nextvpred [n]
 * is the bootstrap value already masked by (1 - new after the last step); writes adv and tdlamret = adv + vpred. */
int dmb_gae(const float* rew, const float* vpred, const float* isnew, const float* nextvpred, int32_t T, int32_t n,
            float gamma, float lam, float* adv, float* tdlamret, void* stream);

#ifdef __cplusplus
}
#endif
#endif
