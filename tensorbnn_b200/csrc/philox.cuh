// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11) for the momentum
// draw and the Metropolis uniform of the on-device HMC step.  The reference draws from
// TF's stateful generator under tf.random.set_seed(50) (network.py:562), which cannot be
// reproduced outside TF; parity tests inject momentum/uniforms instead.
//
// Stream definition (restated in tests/philox_ref.py):
//   key     = (seed_lo, seed_hi ^ stream_tag)
//   counter = (block index j, chain, call_lo, call_hi)
//   normals for flat elements 4j..4j+3 = Box-Muller on the four outputs (fp32), or on
//   two 53-bit uniforms built from them (fp64: elements 2j, 2j+1 ... see draw_normals).
#pragma once
#include <stdint.h>

namespace tbnn {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t STREAM_MAIN = 0x0u, STREAM_HYPER = 0x48595045u;  // 'HYPE'
constexpr uint32_t UNIFORM_BLOCK = 0xFFFFFFFFu;

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ inline U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c.x, p1 = (uint64_t)PHILOX_M1 * c.z;
    U4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    n.w = (uint32_t)p0;
    c = n;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  return c;
}

__host__ __device__ inline float u01f(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__host__ __device__ inline double u01d(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;  // 53 bits
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// Four fp32 standard normals from one Philox block.
__device__ inline void normals4(U4 r, float (&n)[4]) {
  const float u0 = u01f(r.x), u1 = u01f(r.y), u2 = u01f(r.z), u3 = u01f(r.w);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  n[0] = r0 * c0; n[1] = r0 * s0; n[2] = r1 * c1; n[3] = r1 * s1;
}
// Two fp64 standard normals from one Philox block.
__device__ inline void normals2(U4 r, double (&n)[2]) {
  const double u0 = u01d(r.x, r.y), u1 = u01d(r.z, r.w);
  const double rr = sqrt(-2.0 * log(u0));
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  n[0] = rr * c; n[1] = rr * s;
}

// Standard normal for flat element `idx` of chain `chain`.
template <typename T> __device__ inline T draw_normal(uint64_t seed, uint32_t tag, uint64_t call,
                                                      uint32_t chain, uint32_t idx);
template <> __device__ inline float draw_normal<float>(uint64_t seed, uint32_t tag, uint64_t call,
                                                       uint32_t chain, uint32_t idx) {
  U4 c = {idx >> 2, chain, (uint32_t)call, (uint32_t)(call >> 32)};
  float n[4];
  normals4(philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ tag), n);
  return n[idx & 3];
}
template <> __device__ inline double draw_normal<double>(uint64_t seed, uint32_t tag, uint64_t call,
                                                         uint32_t chain, uint32_t idx) {
  U4 c = {idx >> 1, chain, (uint32_t)call, (uint32_t)(call >> 32)};
  double n[2];
  normals2(philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ tag), n);
  return n[idx & 1];
}
// Uniform in (0,1) for the Metropolis test of chain `chain`.
__device__ inline double draw_uniform(uint64_t seed, uint32_t tag, uint64_t call, uint32_t chain) {
  U4 c = {UNIFORM_BLOCK, chain, (uint32_t)call, (uint32_t)(call >> 32)};
  const U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ tag);
  return u01d(r.x, r.y);
}

}  // namespace tbnn
