// device_math.cuh — Philox4x32-10, the RNG contract, tanh_fast, env dynamics.
// Everything here is per-thread scalar code; the tiled MLP code is in mlp_tile.cuh.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------- Philox4x32-10
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// RNG contract (identical in oracle/ppo_oracle.c): key = seed; ctr = (global env id | epoch
// word, counter_lo, counter_hi, stream).
__host__ __device__ inline void philox_draw(uint64_t seed, uint32_t a, uint64_t counter, uint32_t stream,
                                            uint32_t out[4]) {
  philox4x32_10(a, (uint32_t)counter, (uint32_t)(counter >> 32), stream, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

// the Float64 rand() of StatsBase.sample(Weights) (ppo.jl:26): 53 random bits
__device__ inline double rng_action_uniform(uint64_t seed, uint32_t env, uint64_t step) {
  uint32_t r[4];
  philox_draw(seed, env, step, CRL_STREAM_ACTION, r);
  const uint64_t x = ((uint64_t)r[0] << 32) | r[1];
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
// Box-Muller standard normals for the Gaussian head, one per action dim (A <= 2)
__device__ inline void rng_action_normals(uint64_t seed, uint32_t env, uint64_t step, float z[2]) {
  uint32_t r[4];
  philox_draw(seed, env, step, CRL_STREAM_ACTION, r);
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const float u1 = (float)((r[2 * a] >> 8) + 1u) * (1.0f / 16777216.0f);
    const float u2 = (float)(r[2 * a + 1] >> 8) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    z[a] = rad * cosf(6.283185307179586f * u2);
  }
}
// rand(rng, Float32, 4) of reset! [RLEnvs 0.6.12]: 24-bit uniforms in [0,1)
__device__ inline void rng_reset_uniforms(uint64_t seed, uint32_t env, uint64_t k, float u[4]) {
  uint32_t r[4];
  philox_draw(seed, env, k, CRL_STREAM_RESET, r);
#pragma unroll
  for (int i = 0; i < 4; i++) u[i] = (float)(r[i] >> 8) * (1.0f / 16777216.0f);
}

// Philox-keyed Feistel bijection on [0,B) with cycle walking: the device replacement for
// shuffle(b_inds), ppo.jl:194. Integer-only, so it is bit-identical to the oracle.
__host__ __device__ inline uint32_t feistel_round(uint32_t x, uint32_t k, uint32_t mask) {
  x ^= k;
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  return x & mask;
}
__host__ __device__ inline void perm_keys(uint64_t seed, uint64_t update_index, uint32_t epoch, uint32_t rank,
                                          uint32_t keys[8]) {
  philox_draw(seed, epoch | (rank << 16), update_index, CRL_STREAM_PERM, keys);
  philox_draw(seed, epoch | (rank << 16) | 0x80000000u, update_index, CRL_STREAM_PERM, keys + 4);
}
__host__ __device__ inline uint32_t perm_index(uint32_t i, uint32_t B, int half_bits, const uint32_t* keys) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t x = i;
  do {
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const uint32_t nl = r;
      r = l ^ feistel_round(r, keys[k], mask);
      l = nl;
    }
    x = (l << half_bits) | r;
  } while (x >= B);
  return x;
}
__host__ __device__ inline int perm_half_bits(uint32_t B) {
  int bits = 1;
  while ((1u << bits) < B) bits++;
  return (bits + 1) / 2;
}

// ---------------------------------------------------------------- activations
// NNlib.tanh_fast(::Float32) [NNlib 0.8.21; networks.jl:6]: x*n(x^2)/d(x^2), Horner with fma.
// The division uses MUFU.RCP + one Newton step (<= 1 ulp from IEEE division).
__device__ __forceinline__ float tanh_fast(float x) {
  const float x2 = x * x;
  const float n = __fmaf_rn(x2, __fmaf_rn(x2, __fmaf_rn(x2, __fmaf_rn(x2, 1.587199e-8f, 2.2332108e-5f), 0.0035974074f), 0.1346604f), 1.0f);
  const float d = __fmaf_rn(x2, __fmaf_rn(x2, __fmaf_rn(x2, __fmaf_rn(x2, 8.7767893e-7f, 0.0003453992f), 0.026262015f), 0.4679937f), 1.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = __fmaf_rn(__fmaf_rn(-d, r, 1.0f), r, r);  // Newton refinement
  const float y = x * (n * r);
  return x2 < 66.0f ? y : copysignf(1.0f, x);
}

// packed FP32 (fma.rn.f32x2 / FFMA2: two FMAs per lane per issued instruction)
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
// tanh_fast on two values at once: same operations, same roundings per lane (bit-identical to tanh_fast). The numerator
// is evaluated with negated coefficients and the reciprocal taken of -d, so both Newton FMAs and the final products
// are plain packed instructions (the two sign flips cancel exactly).
__device__ __forceinline__ float2 tanh_fast2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  float2 n = __ffma2_rn(x2, f2s(-1.587199e-8f), f2s(-2.2332108e-5f));
  n = __ffma2_rn(x2, n, f2s(-0.0035974074f));
  n = __ffma2_rn(x2, n, f2s(-0.1346604f));
  n = __ffma2_rn(x2, n, f2s(-1.0f));
  float2 d = __ffma2_rn(x2, f2s(8.7767893e-7f), f2s(0.0003453992f));
  d = __ffma2_rn(x2, d, f2s(0.026262015f));
  d = __ffma2_rn(x2, d, f2s(0.4679937f));
  d = __ffma2_rn(x2, d, f2s(1.0f));
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(-d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(-d.y));
  r = __ffma2_rn(__ffma2_rn(d, r, f2s(1.0f)), r, r);  // Newton step on -1/d
  float2 y = __fmul2_rn(x, __fmul2_rn(n, r));
  y.x = x2.x < 66.0f ? y.x : copysignf(1.0f, x.x);
  y.y = x2.y < 66.0f ? y.y : copysignf(1.0f, x.y);
  return y;
}

// Float64 division by a compile-time constant, correctly rounded (same bits as __ddiv_rn) in six dependent FMA-class
// operations instead of the ~400-cycle division sequence: q0 = RN(a y) with y = RN(1/c), then two residual corrections
// q <- q + (a - q c) y with exact residuals (Markstein; 4e8 random operands checked against IEEE division on the CPU).
__device__ __forceinline__ double ddiv_const(double a, double c, double y) {
  double q = __dmul_rn(a, y);
  q = __fma_rn(__fma_rn(-q, c, a), y, q);
  q = __fma_rn(__fma_rn(-q, c, a), y, q);
  return q;
}

// ---------------------------------------------------------------- environments
// CartPoleEnv(T=Float32) step [RLEnvs 0.6.12, called at multi_thread_env.jl:91]. Float32 params,
// but the Float64 literal 4/3 promotes thetaacc, xacc and the velocity updates to Float64
// (SURVEY §8a row 5u). Explicit _rn intrinsics keep the compiler from contracting mul+add
// so the arithmetic is the same sequence of roundings as the oracle's.
// cartpole_dynamics: the state update alone (no counters), so that the rollout kernel can evaluate it for both actions
// ahead of the sampling; cartpole_outcome: termination and reward from the new state and the already advanced counter.
__device__ inline void cartpole_dynamics(float s[4], int action) {
  const float gravity = 9.8f, masspole = 0.1f, halflength = 0.5f, forcemag = 10.0f, dt = 0.02f;
  const float totalmass = 1.0f + 0.1f;
  const float polemasslength = 0.1f * 0.5f;
  const float force = action == 1 ? forcemag : -forcemag;
  const float xdot = s[1], theta = s[2], thetadot = s[3];
  float sintheta, costheta;
  sincosf(theta, &sintheta, &costheta);
  const float tmp = __fdiv_rn(__fadd_rn(force, __fmul_rn(__fmul_rn(polemasslength, __fmul_rn(thetadot, thetadot)), sintheta)), totalmass);
  const float num = __fsub_rn(__fmul_rn(gravity, sintheta), __fmul_rn(costheta, tmp));
  const float mc = __fdiv_rn(__fmul_rn(masspole, __fmul_rn(costheta, costheta)), totalmass);
  const double den = __dmul_rn((double)halflength, __dsub_rn(4.0 / 3.0, (double)mc));
  const double thetaacc = __ddiv_rn((double)num, den);
  // division by the constant totalmass: correctly rounded without the division sequence (ddiv_const)
  constexpr double tm_d = (double)(1.0f + 0.1f), tm_inv = 1.0 / tm_d;
  const double xacc = __dsub_rn((double)tmp, ddiv_const(__dmul_rn(__dmul_rn((double)polemasslength, thetaacc), (double)costheta), tm_d, tm_inv));
  s[0] = __fadd_rn(s[0], __fmul_rn(dt, xdot));
  s[1] = (float)__dadd_rn((double)s[1], __dmul_rn((double)dt, xacc));
  s[2] = __fadd_rn(s[2], __fmul_rn(dt, thetadot));
  s[3] = (float)__dadd_rn((double)s[3], __dmul_rn((double)dt, thetaacc));
}
__device__ inline void cartpole_outcome(const float s[4], int t, int max_steps, float& reward, bool& done) {
  const float thetathreshold = (float)(12.0 * 2.0 * 3.141592653589793 / 360.0);
  const float xthreshold = 2.4f;
  done = fabsf(s[0]) > xthreshold || fabsf(s[2]) > thetathreshold || t > max_steps;
  reward = done ? 0.0f : 1.0f;
}
__device__ inline void cartpole_step(float s[4], int& t, int action, int max_steps, float& reward, bool& done) {
  t += 1;
  cartpole_dynamics(s, action);
  cartpole_outcome(s, t, max_steps, reward, done);
}
__device__ inline void cartpole_reset(float s[4], int& t, const float u[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) s[i] = __fsub_rn(__fmul_rn(0.1f, u[i]), 0.05f);
  t = 0;
}
// PendulumEnv(T=Float32) _step! [RLEnvs 0.6.12]; state = (theta, thetadot)
__device__ inline void pendulum_step(float s[2], int& t, float a, int max_steps, float& reward, bool& done) {
  const float max_speed = 8.0f, max_torque = 2.0f, g = 10.0f, m = 1.0f, l = 1.0f, dt = 0.05f;
  const double PI = 3.141592653589793;
  t += 1;
  float th = s[0];
  const float thdot = s[1];
  a = a < -max_torque ? -max_torque : (a > max_torque ? max_torque : a);
  const float xp = __fadd_rn(th, (float)PI);
  double an = fmod((double)xp, 2.0 * PI);
  if (an < 0.0) an = __dadd_rn(an, 2.0 * PI);
  an = __dsub_rn(an, PI);
  const double costs = __dadd_rn(__dadd_rn(__dmul_rn(an, an), __dmul_rn(0.1, (double)__fmul_rn(thdot, thdot))),
                                 __dmul_rn(0.001, (double)__fmul_rn(a, a)));
  const float c1 = __fdiv_rn(__fmul_rn(-3.0f, g), __fmul_rn(2.0f, l));
  const float term = __fadd_rn(__fmul_rn(c1, sinf(xp)), __fdiv_rn(__fmul_rn(3.0f, a), __fmul_rn(m, __fmul_rn(l, l))));
  float newthdot = __fadd_rn(thdot, __fmul_rn(term, dt));
  th = __fadd_rn(th, __fmul_rn(newthdot, dt));
  newthdot = newthdot < -max_speed ? -max_speed : (newthdot > max_speed ? max_speed : newthdot);
  s[0] = th;
  s[1] = newthdot;
  done = t >= max_steps;
  reward = (float)(-costs);
}
__device__ inline void pendulum_reset(float s[2], int& t, const float u[4]) {
  s[0] = (float)__dmul_rn(2.0 * 3.141592653589793, (double)__fsub_rn(u[0], 1.0f));
  s[1] = __fmul_rn(2.0f, __fsub_rn(u[1], 1.0f));
  t = 0;
}

template <int ENV> __device__ inline void env_obs(const float* s, float* obs) {
  if (ENV == CRL_ENV_CARTPOLE) {
#pragma unroll
    for (int i = 0; i < 4; i++) obs[i] = s[i];
  } else {
    obs[0] = cosf(s[0]);
    obs[1] = sinf(s[0]);
    obs[2] = s[1];
  }
}
template <int ENV> __device__ inline void env_reset(float* s, int& t, const float u[4]) {
  if (ENV == CRL_ENV_CARTPOLE) cartpole_reset(s, t, u);
  else pendulum_reset(s, t, u);
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
