// dmb_policy.cuh -- fused batched MLP policy + value inference (include/dmb_policy.h).
//
// A warp computes POL_R rows at a time; both networks' weights (2 x ~17 k floats) are staged once per CTA in shared
// memory (139 KB).  A layer is computed with lane = output unit (units lane, lane+32, ...) and a register tile of
// POL_R rows x 4 units per lane: every weight read from shared memory feeds POL_R FFMAs and the POL_R inputs of a
// unit come in one 128-bit broadcast read (activations are kept transposed, [unit][row]).  Round 1 did one row per
// warp with two shared-memory reads per FFMA (94 us per 4096 rows).  Accumulation order per output is unchanged
// (bias, then inputs in index order), so results are bit-identical to the one-row kernel.  0.3 GFLOP per 4096 rows:
// after the tiling the kernel is launch / staging bound, tensor cores would not show (and tf32 would cost the
// fp32 parity of the policy mean).
#pragma once
#include <cuda_runtime.h>

#include "../../include/dmb_policy.h"
#include "dmb_math.cuh"

namespace dmb {

constexpr int POL_MAXH = 128, POL_MAXIN = 64, POL_MAXOUT = 32;
constexpr int POL_R = 4;                       // rows per warp (one float4 of transposed activations)
constexpr int POL_UNITS = POL_MAXH / 32;       // output units per lane

struct PolNet { int w1, b1, w2, b2, w3, b3; };  // offsets (floats) into the shared weight pool

// y[j][r] = act(b[j] + sum_i x[i][r] W[i][j]) for the POL_R rows of the warp; x, y transposed [unit][POL_R]
__device__ __forceinline__ void pol_layer(const float* W, const float* b, const float* xt, float* yt, int nin, int nout,
                                          bool do_tanh, int lane) {
  float acc[POL_UNITS][POL_R];
#pragma unroll
  for (int u = 0; u < POL_UNITS; u++) {
    const int j = lane + 32 * u;
    const float bj = j < nout ? b[j] : 0.f;
#pragma unroll
    for (int r = 0; r < POL_R; r++) acc[u][r] = bj;
  }
  for (int i = 0; i < nin; i++) {
    const float4 xv = *reinterpret_cast<const float4*>(xt + POL_R * i);   // the same word for every lane: broadcast
    const float xr[POL_R] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int u = 0; u < POL_UNITS; u++) {
      const int j = lane + 32 * u;
      const float w = j < nout ? W[i * nout + j] : 0.f;
#pragma unroll
      for (int r = 0; r < POL_R; r++) acc[u][r] = fmaf(xr[r], w, acc[u][r]);
    }
  }
#pragma unroll
  for (int u = 0; u < POL_UNITS; u++) {
    const int j = lane + 32 * u;
    if (j < nout) {
      float4 o;
      o.x = do_tanh ? tanhf(acc[u][0]) : acc[u][0]; o.y = do_tanh ? tanhf(acc[u][1]) : acc[u][1];
      o.z = do_tanh ? tanhf(acc[u][2]) : acc[u][2]; o.w = do_tanh ? tanhf(acc[u][3]) : acc[u][3];
      *reinterpret_cast<float4*>(yt + POL_R * j) = o;
    }
  }
  __syncwarp();
}
static_assert(POL_R == 4, "activations are moved as float4");

__global__ void __launch_bounds__(256) k_policy_act(dmb_policy_t P, const float* __restrict__ obs, int n, int stochastic,
                                                    unsigned long long seed, unsigned step, unsigned first_row, float* ac,
                                                    float* vpred, float* mean_out) {
  extern __shared__ __align__(16) float pool[];
  const int od = P.obs_dim, ad = P.act_dim, H = P.hid;
  PolNet pn, vn;
  int off = 0;
  auto take = [&](int cnt) { int o = off; off += cnt; return o; };
  pn.w1 = take(od * H); pn.b1 = take(H); pn.w2 = take(H * H); pn.b2 = take(H); pn.w3 = take(H * ad); pn.b3 = take(ad);
  vn.w1 = take(od * H); vn.b1 = take(H); vn.w2 = take(H * H); vn.b2 = take(H); vn.w3 = take(H); vn.b3 = take(1);
  const int o_mean = take(od), o_std = take(od), o_logstd = take(ad);
  const int wbase = (off + 3) & ~3;   // 16-byte aligned activations
  auto stage = [&](int dst, const float* src, int cnt) { for (int i = threadIdx.x; i < cnt; i += blockDim.x) pool[dst + i] = src[i]; };
  stage(pn.w1, P.pw1, od * H); stage(pn.b1, P.pb1, H); stage(pn.w2, P.pw2, H * H); stage(pn.b2, P.pb2, H);
  stage(pn.w3, P.pw3, H * ad); stage(pn.b3, P.pb3, ad);
  stage(vn.w1, P.vw1, od * H); stage(vn.b1, P.vb1, H); stage(vn.w2, P.vw2, H * H); stage(vn.b2, P.vb2, H);
  stage(vn.w3, P.vw3, H); stage(vn.b3, P.vb3, 1);
  stage(o_mean, P.ob_mean, od); stage(o_std, P.ob_std, od); stage(o_logstd, P.logstd, ad);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  float* xt = pool + wbase + warp * POL_R * (POL_MAXIN + 2 * POL_MAXH);   // [unit][POL_R]
  float* h1 = xt + POL_R * POL_MAXIN;
  float* h2 = h1 + POL_R * POL_MAXH;
  for (int row0 = (blockIdx.x * W + warp) * POL_R; row0 < n; row0 += gridDim.x * W * POL_R) {
    // obs filter (misc_util.py:50-51); rows past the end repeat the last row and are not stored
    for (int t = lane; t < od * POL_R; t += 32) {
      const int i = t / POL_R, r = t - i * POL_R;
      const int row = min(row0 + r, n - 1);
      const float z = (obs[(size_t)row * od + i] - pool[o_mean + i]) / pool[o_std + i];
      xt[t] = fminf(fmaxf(z, -5.f), 5.f);
    }
    __syncwarp();
    // value net
    pol_layer(pool + vn.w1, pool + vn.b1, xt, h1, od, H, true, lane);
    pol_layer(pool + vn.w2, pool + vn.b2, h1, h2, H, H, true, lane);
#pragma unroll
    for (int r = 0; r < POL_R; r++) {
      float acc = 0.f;
      for (int i = lane; i < H; i += 32) acc = fmaf(h2[POL_R * i + r], pool[vn.w3 + i], acc);
      acc = warp_sum(acc) + pool[vn.b3];
      if (lane == 0 && row0 + r < n) vpred[row0 + r] = acc;
    }
    __syncwarp();
    // policy net
    pol_layer(pool + pn.w1, pool + pn.b1, xt, h1, od, H, true, lane);
    pol_layer(pool + pn.w2, pool + pn.b2, h1, h2, H, H, true, lane);
    if (lane < ad) {
      float m[POL_R];
#pragma unroll
      for (int r = 0; r < POL_R; r++) m[r] = pool[pn.b3 + lane];
      for (int i = 0; i < H; i++) {
        const float4 hv = *reinterpret_cast<const float4*>(h2 + POL_R * i);
        const float w = pool[pn.w3 + i * ad + lane];
        m[0] = fmaf(hv.x, w, m[0]); m[1] = fmaf(hv.y, w, m[1]); m[2] = fmaf(hv.z, w, m[2]); m[3] = fmaf(hv.w, w, m[3]);
      }
#pragma unroll
      for (int r = 0; r < POL_R; r++) {
        const int row = row0 + r;
        if (row >= n) break;
        if (mean_out) mean_out[(size_t)row * ad + lane] = m[r];
        float a = m[r];
        if (stochastic) {
          unsigned rr[4];
          philox4x32((unsigned)seed, (unsigned)(seed >> 32), first_row + (unsigned)row, step, (unsigned)(lane >> 1), 0x504f4cu, rr);
          const float u1 = ((float)(rr[(lane & 1) * 2] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
          const float u2 = u01(rr[(lane & 1) * 2 + 1]);
          const float z = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
          a = fmaf(expf(pool[o_logstd + lane]), z, m[r]);
        }
        ac[(size_t)row * ad + lane] = a;
      }
    }
    __syncwarp();
  }
}

}  // namespace dmb

extern "C" int dmb_policy_act(const dmb_policy_t* p, const float* obs, int32_t n, int32_t stochastic, uint64_t seed,
                              uint32_t step, uint32_t first_row, float* ac, float* vpred, float* mean_out, void* stream) {
  using namespace dmb;
  if (!p || !obs || !ac || !vpred || n <= 0) return DMB_ERR_ARG;
  if (p->obs_dim > POL_MAXIN || p->hid > POL_MAXH || p->act_dim > POL_MAXOUT || p->obs_dim < 1 || p->hid < 1 || p->act_dim < 1)
    return DMB_ERR_MODEL;
  const int od = p->obs_dim, ad = p->act_dim, H = p->hid;
  const int weights = 2 * (od * H + H + H * H + H) + H * ad + ad + H + 1 + 2 * od + ad;
  const int Wp = 8;
  const size_t smem = sizeof(float) * ((size_t)((weights + 3) & ~3) + (size_t)Wp * POL_R * (POL_MAXIN + 2 * POL_MAXH));
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DMB_ERR_NO_DEVICE;
  if (cudaFuncSetAttribute((const void*)k_policy_act, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return DMB_ERR_CUDA;
  int grid = (n + Wp * POL_R - 1) / (Wp * POL_R);
  if (grid > nsm) grid = nsm;
  k_policy_act<<<grid, Wp * 32, smem, (cudaStream_t)stream>>>(*p, obs, n, stochastic, seed, step, first_row, ac, vpred, mean_out);
  return cudaGetLastError() == cudaSuccess ? DMB_OK : DMB_ERR_CUDA;
}

// GAE(lambda) over a [T][n] segment, one thread per env walking its T steps backwards (trpo.py:83-94
// add_vtarg_and_adv): nonterminal = 1 - new[t+1] (0 after the last step is handled by nextvpred being pre-masked),
// delta = rew[t] + gamma * vpred[t+1] * nonterminal - vpred[t], adv[t] = delta + gamma * lam * nonterminal * adv[t+1].
namespace dmb {
__global__ void k_gae(const float* __restrict__ rew, const float* __restrict__ vpred, const float* __restrict__ isnew,
                      const float* __restrict__ nextvpred, int T, int n, float gamma, float lam, float* adv, float* tdlamret) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float last = 0.f, nextv = nextvpred[e], nextnew = 0.f;
  for (int t = T - 1; t >= 0; t--) {
    const size_t k = (size_t)t * n + e;
    const float nonterminal = 1.f - nextnew, v = vpred[k];
    const float delta = rew[k] + gamma * nextv * nonterminal - v;
    last = delta + gamma * lam * nonterminal * last;
    adv[k] = last;
    tdlamret[k] = last + v;
    nextv = v;
    nextnew = isnew[k];
  }
}
}  // namespace dmb

extern "C" int dmb_gae(const float* rew, const float* vpred, const float* isnew, const float* nextvpred, int32_t T, int32_t n,
                       float gamma, float lam, float* adv, float* tdlamret, void* stream) {
  if (!rew || !vpred || !isnew || !nextvpred || !adv || !tdlamret || T <= 0 || n <= 0) return DMB_ERR_ARG;
  dmb::k_gae<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rew, vpred, isnew, nextvpred, T, n, gamma, lam, adv, tdlamret);
  return cudaGetLastError() == cudaSuccess ? DMB_OK : DMB_ERR_CUDA;
}

