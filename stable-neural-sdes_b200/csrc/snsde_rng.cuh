// In-register Brownian increments: Philox4x32-10 + Box-Muller.
//
// Replaces torchsde.BrownianInterval (constructed implicitly by sdeint because the reference
// never passes bm=, neuralsde.py:78-82).  Bit parity with BrownianInterval is impossible by
// construction (numpy SeedSequence tree + torch generator streams), so the engine defines its
// own counter-based stream and exposes it through snsde_philox_fill for the oracle:
//
//   key = (seed_lo, seed_hi);  counter = (feature j, global_row >> 2, step, 'SNSD')
//   -> 4 normals for global rows 4p..4p+3 of feature j at that step.
//
// Every kernel that needs dW[s][b][j] calls the SAME inline function below, so the increments
// are identical in the FMA kernel, the tcgen05 kernel and snsde_philox_fill, for any sharding.
#pragma once
#include <stdint.h>

namespace snsde {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kStreamTag = 0x534E5344u;      // 'SNSD': Brownian increments W
constexpr uint32_t kStreamTagU = 0x534E5355u;     // 'SNSU': space-time Levy areas of the SRK method

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += kPhiloxW0;
    k.y += kPhiloxW1;
  }
  return c;
}

// u in (0,1]:  (x + 0.5) * 2^-32
__device__ __forceinline__ float u01(uint32_t x) {
  return __fmaf_rn(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  const float r = __fsqrt_rn(__fmul_rn(-2.0f, __logf(u01(xa))));
  float sn, cs;
  __sincosf(__fmul_rn(6.2831853071795865f, u01(xb)), &sn, &cs);
  n0 = __fmul_rn(r, cs);
  n1 = __fmul_rn(r, sn);
}

// Standard normals for global rows 4p..4p+3, feature j, step s.
__device__ __forceinline__ void philox_normals4(unsigned long long seed, uint32_t j, uint32_t p, uint32_t s,
                                                float n[4]) {
  const uint4 r = philox4x32_10(make_uint4(j, p, s, kStreamTag),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  box_muller(r.x, r.y, n[0], n[1]);
  box_muller(r.z, r.w, n[2], n[3]);
}

// The one normal of global row `gb` (= element gb & 3 of philox_normals4(seed, j, gb >> 2, s), bit for bit) when a
// thread serves a single row: one Box-Muller pair instead of two.
__device__ __forceinline__ float philox_normal1(unsigned long long seed, uint32_t j, unsigned long long gb, uint32_t s) {
  const uint4 r = philox4x32_10(make_uint4(j, (uint32_t)(gb >> 2), s, kStreamTag),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const bool hi = (gb & 2ull) != 0ull;
  float n0, n1;
  box_muller(hi ? r.z : r.x, hi ? r.w : r.y, n0, n1);
  return (gb & 1ull) ? n1 : n0;
}

// Second, independent stream for the SRK method: standard normals behind the space-time Levy area.
__device__ __forceinline__ void philox_normals4_u(unsigned long long seed, uint32_t j, uint32_t p, uint32_t s,
                                                  float n[4]) {
  const uint4 r = philox4x32_10(make_uint4(j, p, s, kStreamTagU),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  box_muller(r.x, r.y, n[0], n[1]);
  box_muller(r.z, r.w, n[2], n[3]);
}

// torchsde `bm(t0, t1, return_U=True)` for the space-time Levy area approximation:
//   U = h (W/2 + Hst),  Hst ~ N(0, h/12) independent of W   (brownian_interval.py _H_to_U)
__device__ __forceinline__ float levy_U(float W, float n2, float h, float sqrt_h) {
  const float hst = __fmul_rn(n2, __fmul_rn(sqrt_h, 0.28867513459481287f));
  return __fmul_rn(h, __fadd_rn(__fmul_rn(0.5f, W), hst));
}

// n[lane] without a dynamically indexed (local-memory) array
__device__ __forceinline__ float pick4(const float (&n)[4], int lane) {
  const float lo = (lane & 1) ? n[1] : n[0];
  const float hi = (lane & 1) ? n[3] : n[2];
  return (lane & 2) ? hi : lo;
}

}  // namespace snsde
