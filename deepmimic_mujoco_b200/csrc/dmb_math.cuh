// dmb_math.cuh -- small fp32 device math (quaternions, spatial vectors, warp helpers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace dmb {

#define DMB_FULL 0xffffffffu
#define DMB_MINVAL 1e-15f

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }
// mju_normalize3: tiny vectors become (1,0,0); returns the norm
__device__ __forceinline__ float normalize(V3& a) {
  const float d2 = dot(a, a);
  if (d2 < DMB_MINVAL * DMB_MINVAL) { a = v3(1.f, 0.f, 0.f); return sqrtf(d2); }
  const float s = rsqrtf(d2);   // 2 ulp; one MUFU instead of sqrt + IEEE divide
  a = s * a;
  return d2 * s;
}
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  const float d2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  if (d2 < DMB_MINVAL * DMB_MINVAL) { q.w = 1.f; q.x = q.y = q.z = 0.f; }
  else { const float s = rsqrtf(d2); q.w *= s; q.x *= s; q.y *= s; q.z *= s; }
  return q;
}
__device__ __forceinline__ void quat2mat(float* m, Q4 q) {
  float w = q.w, x = q.x, y = q.y, z = q.z;
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
  m[3] = 2.f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.f * (y * z - w * x);
  m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
__device__ __forceinline__ V3 mat_vec(const float* m, V3 v) {
  return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z,
            m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
__device__ __forceinline__ V3 matT_vec(const float* m, V3 v) {
  return v3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
            m[2] * v.x + m[5] * v.y + m[8] * v.z);
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  float m[9];
  quat2mat(m, q);
  return mat_vec(m, v);
}

// spatial vectors [angular; linear]
__device__ __forceinline__ void cross_motion(float* r, const float* v, const float* s) {
  V3 va = ld3(v), vl = ld3(v + 3), sa = ld3(s), sl = ld3(s + 3);
  st3(r, cross(va, sa));
  st3(r + 3, cross(va, sl) + cross(vl, sa));
}
__device__ __forceinline__ void cross_force(float* r, const float* v, const float* f) {
  V3 va = ld3(v), vl = ld3(v + 3), fa = ld3(f), fl = ld3(f + 3);
  st3(r, cross(va, fa) + cross(vl, fl));
  st3(r + 3, cross(va, fl));
}
__device__ __forceinline__ void mul_inert_vec(float* r, const float* i, const float* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DMB_FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(DMB_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Philox4x32-10 (Salmon et al. 2011), identical integer arithmetic in oracle/dm_oracle.c
__device__ __host__ __forceinline__ void philox4x32(uint32_t key0, uint32_t key1, uint32_t c0, uint32_t c1,
                                                    uint32_t c2, uint32_t c3, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ key0, n1 = lo1, n2 = hi0 ^ c3 ^ key1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    key0 += W0; key1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in [0,1) with 24 random bits (exactly representable in fp32 and fp64)
__device__ __host__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

}  // namespace dmb
