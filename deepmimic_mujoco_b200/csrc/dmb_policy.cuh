// dmb_policy.cuh -- fused batched MLP policy + value inference (include/dmb_policy.h).
//
// One warp per row; both networks' weights (2 x ~17 k floats) are staged once per CTA in shared
// memory (139 KB); a layer is computed with lane = output unit (units lane, lane+32, ...), the
// input vector is a broadcast read and the weight row a conflict-free read.  0.15 GFLOP for 4096
// rows: no tensor cores needed -- the point is to keep the rollout on the device.
#pragma once
#include <cuda_runtime.h>

#include "../../include/dmb_policy.h"
#include "dmb_math.cuh"

namespace dmb {

constexpr int POL_MAXH = 128, POL_MAXIN = 64, POL_MAXOUT = 32;

struct PolNet { int w1, b1, w2, b2, w3, b3; };  // offsets (floats) into the shared weight pool

__device__ __forceinline__ void pol_layer(const float* W, const float* b, const float* x, float* y, int nin, int nout,
                                          bool do_tanh, int lane) {
  for (int j = lane; j < nout; j += 32) {
    float acc = b[j];
    for (int i = 0; i < nin; i++) acc = fmaf(x[i], W[i * nout + j], acc);
    y[j] = do_tanh ? tanhf(acc) : acc;
  }
  __syncwarp();
}

__global__ void k_policy_act(dmb_policy_t P, const float* __restrict__ obs, int n, int stochastic, unsigned long long seed,
                             unsigned step, unsigned first_row, float* ac, float* vpred, float* mean_out) {
  extern __shared__ __align__(16) float pool[];
  const int od = P.obs_dim, ad = P.act_dim, H = P.hid;
  PolNet pn, vn;
  int off = 0;
  auto take = [&](int cnt) { int o = off; off += cnt; return o; };
  pn.w1 = take(od * H); pn.b1 = take(H); pn.w2 = take(H * H); pn.b2 = take(H); pn.w3 = take(H * ad); pn.b3 = take(ad);
  vn.w1 = take(od * H); vn.b1 = take(H); vn.w2 = take(H * H); vn.b2 = take(H); vn.w3 = take(H); vn.b3 = take(1);
  const int o_mean = take(od), o_std = take(od), o_logstd = take(ad);
  const int wbase = off;
  auto stage = [&](int dst, const float* src, int cnt) { for (int i = threadIdx.x; i < cnt; i += blockDim.x) pool[dst + i] = src[i]; };
  stage(pn.w1, P.pw1, od * H); stage(pn.b1, P.pb1, H); stage(pn.w2, P.pw2, H * H); stage(pn.b2, P.pb2, H);
  stage(pn.w3, P.pw3, H * ad); stage(pn.b3, P.pb3, ad);
  stage(vn.w1, P.vw1, od * H); stage(vn.b1, P.vb1, H); stage(vn.w2, P.vw2, H * H); stage(vn.b2, P.vb2, H);
  stage(vn.w3, P.vw3, H); stage(vn.b3, P.vb3, 1);
  stage(o_mean, P.ob_mean, od); stage(o_std, P.ob_std, od); stage(o_logstd, P.logstd, ad);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  float* x = pool + wbase + warp * (POL_MAXIN + 2 * POL_MAXH);
  float* h1 = x + POL_MAXIN;
  float* h2 = h1 + POL_MAXH;
  for (int row = blockIdx.x * W + warp; row < n; row += gridDim.x * W) {
    for (int i = lane; i < od; i += 32) {
      const float z = (obs[(size_t)row * od + i] - pool[o_mean + i]) / pool[o_std + i];
      x[i] = fminf(fmaxf(z, -5.f), 5.f);
    }
    __syncwarp();
    // value net
    pol_layer(pool + vn.w1, pool + vn.b1, x, h1, od, H, true, lane);
    pol_layer(pool + vn.w2, pool + vn.b2, h1, h2, H, H, true, lane);
    {
      float acc = 0.f;
      for (int i = lane; i < H; i += 32) acc = fmaf(h2[i], pool[vn.w3 + i], acc);
      acc = warp_sum(acc) + pool[vn.b3];
      if (lane == 0) vpred[row] = acc;
    }
    __syncwarp();
    // policy net
    pol_layer(pool + pn.w1, pool + pn.b1, x, h1, od, H, true, lane);
    pol_layer(pool + pn.w2, pool + pn.b2, h1, h2, H, H, true, lane);
    if (lane < ad) {
      float m = pool[pn.b3 + lane];
      for (int i = 0; i < H; i++) m = fmaf(h2[i], pool[pn.w3 + i * ad + lane], m);
      if (mean_out) mean_out[(size_t)row * ad + lane] = m;
      float a = m;
      if (stochastic) {
        unsigned r[4];
        philox4x32((unsigned)seed, (unsigned)(seed >> 32), first_row + (unsigned)row, step, (unsigned)(lane >> 1), 0x504f4cu, r);
        const float u1 = ((float)(r[(lane & 1) * 2] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
        const float u2 = u01(r[(lane & 1) * 2 + 1]);
        const float z = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
        a = fmaf(expf(pool[o_logstd + lane]), z, m);
      }
      ac[(size_t)row * ad + lane] = a;
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
  const size_t smem = sizeof(float) * ((size_t)weights + (size_t)Wp * (POL_MAXIN + 2 * POL_MAXH));
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return DMB_ERR_NO_DEVICE;
  if (cudaFuncSetAttribute((const void*)k_policy_act, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return DMB_ERR_CUDA;
  int grid = (n + Wp - 1) / Wp;
  if (grid > nsm) grid = nsm;
  k_policy_act<<<grid, Wp * 32, smem, (cudaStream_t)stream>>>(*p, obs, n, stochastic, seed, step, first_row, ac, vpred, mean_out);
  return cudaGetLastError() == cudaSuccess ? DMB_OK : DMB_ERR_CUDA;
}
