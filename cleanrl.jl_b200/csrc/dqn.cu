// dqn.cu — vectorised DQN (src/algorithms/dqn.jl) behind the crl_dqn_* entry points of
// include/cleanrl_cuda.h: N CartPole envs stepping in lockstep with an epsilon-greedy Q-network, an HBM-resident
// ring replay buffer, and one learning step (sample without replacement -> TD target from the target net -> MSE on
// the taken action -> backward -> Adam -> optional target copy) every train_freq iterations.
//
//   dqn_act_kernel    32 envs per CTA of 256 threads: lane = env, warp = neuron group. The 10,934 parameters sit in
//                     shared memory; activations are [neuron][env lane] tiles in shared memory (conflict-free; the
//                     weights of four consecutive neurons are one broadcast 128-bit load). Warp 0 then draws epsilon /
//                     the random action from Philox, steps CartPole, appends the transition at
//                     (ptr + step * N + env) % capacity and resets a finished env on the spot. The parameters only
//                     change at a learning step, so ONE launch runs every iteration up to the next learning step
//                     (train_freq of them, at most ACT_MAX_STEPS) with the env state in registers.
//   dqn_learn_kernel  ONE CTA of 512 threads for the batch (<= 128 samples x 4 neuron groups): both parameter sets
//                     and the batch activations live in shared memory (~197 KB, activation rows padded to 129 floats
//                     so that the batch reductions, whose lanes walk down the neuron axis, are conflict-free);
//                     forward target net and q net, gradient reductions over the batch as 4x4 register tiles, every
//                     sum in ascending sample order (deterministic, same order as the oracle), Adam and the target
//                     copy in the same launch.
// Every decision that only depends on counters (epsilon, "learn on this iteration?", "copy the target?") is taken on
// the host, so a run is a plain sequence of launches on one stream without any device-to-host read.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "device_math.cuh"
#include "kernels.h"

namespace {

constexpr int DQ_D = 4, DQ_H1 = 120, DQ_H2 = 84, DQ_A = 2;
constexpr int DQ_W1 = 0, DQ_B1 = DQ_W1 + DQ_H1 * DQ_D, DQ_W2 = DQ_B1 + DQ_H1, DQ_B2 = DQ_W2 + DQ_H2 * DQ_H1,
              DQ_W3 = DQ_B2 + DQ_H2, DQ_B3 = DQ_W3 + DQ_A * DQ_H2, DQ_P = DQ_B3 + DQ_A;
static_assert(DQ_P == CRL_DQN_PARAMS, "parameter count");
constexpr uint32_t STREAM_DQN_ACT = 3u, STREAM_DQN_BATCH = 4u;
constexpr int ACT_E = 32;     // envs per CTA in dqn_act_kernel (one warp-width of samples)
constexpr int ACT_G = DQ_H2 / 4;   // 21 warps: warp g owns neurons 4g..4g+3 of the second layer (one pass) and g, g+21, .. of the first
constexpr int ACT_T = ACT_E * ACT_G;
constexpr int ACT_MAX_STEPS = 16;   // iterations per dqn_act_kernel launch (bounded by the next learning step)
constexpr int LEARN_B = 128;  // max batch size = samples per tile in dqn_learn_kernel
constexpr int LEARN_SP = LEARN_B + 1;   // activation row stride in dqn_learn_kernel: odd, so rows fall on distinct banks
// the learning step spread over several SMs (default): forward/backward for 16 samples per CTA, then one thread per
// 4x4 tile of dW2 (or per single small-array element) which reduces over the batch and applies Adam on the spot
constexpr int LF_S = 16, LF_G = 32, LF_T = LF_S * LF_G, LF_SP = LF_S + 1;
constexpr int LF_MAX_BLOCKS = LEARN_B / LF_S;
constexpr int LU_T = 64;
constexpr int LU_TILES = (DQ_H2 / 4) * (DQ_H1 / 4);
constexpr int LU_TILE_BLOCKS = (LU_TILES + LU_T - 1) / LU_T;
constexpr int LU_SINGLES = DQ_H1 * DQ_D + DQ_H1 + DQ_H2 + DQ_A * DQ_H2 + DQ_A;   // dW1, db1, db2, dW3, db3
constexpr int LU_SINGLE_BLOCKS = (LU_SINGLES + LU_T - 1) / LU_T;
constexpr int LEARN_G = 4;    // neuron groups
constexpr int LEARN_T = LEARN_B * LEARN_G;

struct DqnDev {               // device-resident scalars
  double bp1, bp2;            // beta1^t, beta2^t of Adam
  double last_loss;
  double sum_return, sum_length;
  unsigned long long episodes;
};

struct ActArgs {
  const float* q;
  float* env_state; int* env_t; double* ep_ret; int* ep_len; uint32_t* resets;
  float *b_state, *b_next, *b_reward; int* b_action; uint8_t* b_term;
  DqnDev* dev;
  unsigned long long seed, it0;   // it0 = iteration number (1-based) of step 0 of this launch
  double eps[ACT_MAX_STEPS];      // epsilon of each step (dqn.jl:52, evaluated on the host)
  int n_steps;
  int N, C, ptr, max_steps;
};

// Dense -> relu -> Dense -> relu -> Dense for a tile of samples, neurons spread over G thread groups (CTA-wide
// barriers between the layers: every thread of the CTA must call this). Activations are [neuron][SP] tiles in shared
// memory (row stride SP floats; conflict-free across the lanes of a warp, which are consecutive samples); weight reads
// are warp broadcasts, and in the 120 -> 84 layer a thread owns four consecutive neurons so that their weights are
// one 128-bit load per k (p must be 16-byte aligned). Each neuron is one fmaf chain over ascending k, so the result
// does not depend on the tile shape or on G. qo = [DQ_A][SP].
static_assert(DQ_H2 % 4 == 0 && DQ_W2 % 4 == 0, "128-bit weight loads of the second layer");
template <int SP, int G>
__device__ __forceinline__ void q_forward(const float* __restrict__ p, const float x[DQ_D], float* h1, float* h2, float* qo,
                                          int l, int g) {
  for (int j = g; j < DQ_H1; j += G) {
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < DQ_D; k++) acc = fmaf(p[DQ_W1 + j + DQ_H1 * k], x[k], acc);
    acc += p[DQ_B1 + j];
    h1[j * SP + l] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  for (int j = 4 * g; j < DQ_H2; j += 4 * G) {
    float2 a01 = make_float2(0.0f, 0.0f), a23 = make_float2(0.0f, 0.0f);   // packed FP32: two chains per instruction
#pragma unroll 8
    for (int k = 0; k < DQ_H1; k++) {
      const float4 w = *reinterpret_cast<const float4*>(p + DQ_W2 + j + DQ_H2 * k);
      const float h = h1[k * SP + l];
      a01 = __ffma2_rn(make_float2(w.x, w.y), make_float2(h, h), a01);
      a23 = __ffma2_rn(make_float2(w.z, w.w), make_float2(h, h), a23);
    }
    h2[(j + 0) * SP + l] = fmaxf(a01.x + p[DQ_B2 + j + 0], 0.0f);
    h2[(j + 1) * SP + l] = fmaxf(a01.y + p[DQ_B2 + j + 1], 0.0f);
    h2[(j + 2) * SP + l] = fmaxf(a23.x + p[DQ_B2 + j + 2], 0.0f);
    h2[(j + 3) * SP + l] = fmaxf(a23.y + p[DQ_B2 + j + 3], 0.0f);
  }
  __syncthreads();
  for (int o = g; o < DQ_A; o += G) {
    float acc = 0.0f;
#pragma unroll 4
    for (int k = 0; k < DQ_H2; k++) acc = fmaf(p[DQ_W3 + o + DQ_A * k], h2[k * SP + l], acc);
    qo[o * SP + l] = acc + p[DQ_B3 + o];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(ACT_T) dqn_act_kernel(ActArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* p = smem;                       // [DQ_P]
  float* h1 = p + ((DQ_P + 3) & ~3);     // [120][32]
  float* h2 = h1 + DQ_H1 * ACT_E;        // [84][32]
  float* qo = h2 + DQ_H2 * ACT_E;        // [2][32]
  float* xs = qo + DQ_A * ACT_E;         // [4][32] current observation of the CTA's envs
  for (int i = threadIdx.x; i < DQ_P; i += ACT_T) p[i] = a.q[i];
  const int e = threadIdx.x & (ACT_E - 1), g = threadIdx.x / ACT_E;
  const int n = blockIdx.x * ACT_E + e;
  const bool owner = g == 0, valid = n < a.N;   // warp 0: one lane per env, state in registers for the whole launch
  float st[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  int t = 0, len = 0;
  double ret = 0.0;
  uint32_t rc = 0;
  if (owner) {
    if (valid) {
#pragma unroll
      for (int k = 0; k < 4; k++) st[k] = a.env_state[4 * n + k];
      t = a.env_t[n];
      ret = a.ep_ret[n];
      len = a.ep_len[n];
      rc = a.resets[n];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) xs[k * ACT_E + e] = st[k];
  }
  __syncthreads();
  for (int s = 0; s < a.n_steps; s++) {
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = xs[k * ACT_E + e];
    q_forward<ACT_E, ACT_G>(p, x, h1, h2, qo, e, g);   // q_net(obs), dqn.jl:56 (used only where the epsilon test fails)
    if (owner && valid) {
      const float4 obs = make_float4(st[0], st[1], st[2], st[3]);   // deepcopy(state(env)), dqn.jl:50
      uint32_t r[4];
      philox_draw(a.seed, (uint32_t)n, a.it0 + (unsigned long long)s, STREAM_DQN_ACT, r);
      const double u = (double)((((uint64_t)r[0] << 32) | r[1]) >> 11) * (1.0 / 9007199254740992.0);
      int action;
      if (u < a.eps[s]) action = (int)(r[2] & 1u);                  // rand(action_space(env)), dqn.jl:54
      else action = qo[1 * ACT_E + e] > qo[0 * ACT_E + e] ? 1 : 0;  // argmax: first maximum, dqn.jl:56-57
      float rew;
      bool done;
      cartpole_step(st, t, action, a.max_steps, rew, done);
      const int slot = (int)(((long long)a.ptr + (long long)s * a.N + n) % a.C);   // add!, replay_buffer.jl:23-37, envs in order
      reinterpret_cast<float4*>(a.b_state)[slot] = obs;
      reinterpret_cast<float4*>(a.b_next)[slot] = make_float4(st[0], st[1], st[2], st[3]);
      a.b_action[slot] = action;
      a.b_reward[slot] = rew;
      a.b_term[slot] = done ? 1 : 0;
      ret += (double)rew;
      len += 1;
      if (done) {                          // dqn.jl:80-86
        atomicAdd(&a.dev->episodes, 1ull);
        atomicAdd(&a.dev->sum_return, ret);
        atomicAdd(&a.dev->sum_length, (double)len);
        ret = 0.0;
        len = 0;
        float u4[4];
        rng_reset_uniforms(a.seed, (uint32_t)n, rc, u4);
        rc += 1;
        cartpole_reset(st, t, u4);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) xs[k * ACT_E + e] = st[k];
    }
    __syncthreads();
  }
  if (owner && valid) {
    a.ep_ret[n] = ret;
    a.ep_len[n] = len;
    a.env_t[n] = t;
    a.resets[n] = rc;
#pragma unroll
    for (int k = 0; k < 4; k++) a.env_state[4 * n + k] = st[k];
  }
}

struct LearnArgs {
  float* q; float* tgt; float* m; float* v; float* g;
  const float *b_state, *b_next, *b_reward; const int* b_action; const uint8_t* b_term;
  DqnDev* dev;
  unsigned long long seed, learn_step;
  double gamma, lr;
  int B, size, copy_target;
  // multi-SM path: sample-major scratch [128][rows] between the two kernels, Adam's beta powers from the host
  float *h1T, *h2T, *z2T, *z1T, *xT, *dqT;
  double* loss_part;
  double bp1, bp2;
};

__global__ void __launch_bounds__(LEARN_T) dqn_learn_kernel(LearnArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PP = (DQ_P + 3) & ~3, S = LEARN_B, SP = LEARN_SP;
  float* pq = smem;                       // q_net parameters
  float* pt = pq + PP;                    // target_net parameters
  float* h1 = pt + PP;                    // [120][SP]  (later dz1 in place)
  float* h2 = h1 + DQ_H1 * SP;            // [84][SP]   (later dz2 in place)
  float* xs = h2 + DQ_H2 * SP;            // [4][SP] states of the batch
  float* dq = xs + DQ_D * SP;             // [2][SP]
  float* qo = dq + DQ_A * SP;             // [2][SP] network outputs
  __shared__ uint32_t keys[8];
  __shared__ double red[S / 32];
  const int tid = threadIdx.x, B = a.B;
  const int i = tid & (S - 1), g = tid / S;   // sample, neuron group
  for (int k = tid; k < DQ_P; k += LEARN_T) { pq[k] = a.q[k]; pt[k] = a.tgt[k]; }
  if (tid == 0) {
    philox_draw(a.seed, 0u, a.learn_step, STREAM_DQN_BATCH, keys);
    philox_draw(a.seed, 0x80000000u, a.learn_step, STREAM_DQN_BATCH, keys + 4);
  }
  __syncthreads();
  float s[4] = {0.f, 0.f, 0.f, 0.f}, nx[4] = {0.f, 0.f, 0.f, 0.f};
  int act = 0, term = 0;
  float rew = 0.0f;
  if (i < B) {
    // sample(1:size, B, replace=false), replay_buffer.jl:43: the first B entries of a keyed permutation of [0,size)
    const uint32_t idx = perm_index((uint32_t)i, (uint32_t)a.size, perm_half_bits((uint32_t)a.size), keys);
    const float4 s4 = reinterpret_cast<const float4*>(a.b_state)[idx];
    const float4 n4 = reinterpret_cast<const float4*>(a.b_next)[idx];
    s[0] = s4.x; s[1] = s4.y; s[2] = s4.z; s[3] = s4.w;
    nx[0] = n4.x; nx[1] = n4.y; nx[2] = n4.z; nx[3] = n4.w;
    act = a.b_action[idx];
    rew = a.b_reward[idx];
    term = a.b_term[idx];
  }
  q_forward<SP, LEARN_G>(pt, nx, h1, h2, qo, i, g);                  // target_net(next_state), dqn.jl:99
  const float next_q = fmaxf(qo[0 * SP + i], qo[1 * SP + i]);
  const double td = (double)rew + a.gamma * (double)next_q * (1.0 - (double)term);   // dqn.jl:100
  __syncthreads();
  q_forward<SP, LEARN_G>(pq, s, h1, h2, qo, i, g);                   // q_net(state), dqn.jl:105
  double sq = 0.0;
  if (g == 0) {
    dq[0 * SP + i] = 0.0f;
    dq[1 * SP + i] = 0.0f;
#pragma unroll
    for (int k = 0; k < DQ_D; k++) xs[k * SP + i] = s[k];
    if (i < B) {
      const double diff = td - (double)qo[act * SP + i];
      sq = diff * diff;                                              // Flux.mse, dqn.jl:107
      dq[act * SP + i] = (float)(-2.0 * diff / (double)B);
    }
    sq = warp_sum(sq);
    if ((tid & 31) == 0) red[tid >> 5] = sq;
  }
  __syncthreads();
  if (tid == 0) a.dev->last_loss = ((red[0] + red[1]) + (red[2] + red[3])) / (double)B;
  // ---- backward; every reduction over the batch runs in ascending sample order in one thread. The lanes of a warp
  //      walk down the NEURON axis of the [neuron][SP] tiles here: SP is odd, so they hit distinct banks.
  float* gr = a.g;
  for (int w = tid; w < DQ_A * DQ_H2 + DQ_A; w += LEARN_T) {         // dW3(o,k), db3(o)
    if (w < DQ_A * DQ_H2) {
      const int o = w % DQ_A, k = w / DQ_A;
      float acc = 0.0f;
      for (int b = 0; b < B; b++) acc = fmaf(dq[o * SP + b], h2[k * SP + b], acc);
      gr[DQ_W3 + w] = acc;
    } else {
      const int o = w - DQ_A * DQ_H2;
      float acc = 0.0f;
      for (int b = 0; b < B; b++) acc += dq[o * SP + b];
      gr[DQ_B3 + o] = acc;
    }
  }
  __syncthreads();
  {                                                                  // dz2 = (W3^T dq) .* (h2 > 0), in place
    const float d0 = dq[0 * SP + i], d1 = dq[1 * SP + i];
    for (int k = g; k < DQ_H2; k += LEARN_G) {
      const float dh = fmaf(pq[DQ_W3 + 1 + DQ_A * k], d1, pq[DQ_W3 + 0 + DQ_A * k] * d0);
      h2[k * SP + i] = h2[k * SP + i] > 0.0f ? dh : 0.0f;
    }
  }
  __syncthreads();
  {                                                                  // dW2(j,k) as 4x4 register tiles, db2(j)
    constexpr int JQ = DQ_H2 / 4, KQ = DQ_H1 / 4;                    // tile = neurons j, j+21, j+42, j+63 x inputs k0..k0+3
    static_assert(DQ_H2 % 4 == 0 && DQ_H1 % 4 == 0, "4x4 weight-gradient tiles");
    for (int w = tid; w < JQ * KQ; w += LEARN_T) {
      const int j = w % JQ, k0 = (w / JQ) * 4;
      float acc[4][4];
#pragma unroll
      for (int t = 0; t < 4; t++)
#pragma unroll
        for (int u = 0; u < 4; u++) acc[t][u] = 0.0f;
#pragma unroll 2
      for (int b = 0; b < B; b++) {
        float z[4], hh[4];
#pragma unroll
        for (int t = 0; t < 4; t++) z[t] = h2[(j + JQ * t) * SP + b];
#pragma unroll
        for (int u = 0; u < 4; u++) hh[u] = h1[(k0 + u) * SP + b];
#pragma unroll
        for (int t = 0; t < 4; t++)
#pragma unroll
          for (int u = 0; u < 4; u++) acc[t][u] = fmaf(z[t], hh[u], acc[t][u]);
      }
#pragma unroll
      for (int t = 0; t < 4; t++)
#pragma unroll
        for (int u = 0; u < 4; u++) gr[DQ_W2 + (j + JQ * t) + DQ_H2 * (k0 + u)] = acc[t][u];
    }
    for (int j = tid; j < DQ_H2; j += LEARN_T) {
      float acc = 0.0f;
      for (int b = 0; b < B; b++) acc += h2[j * SP + b];
      gr[DQ_B2 + j] = acc;
    }
  }
  __syncthreads();
  // dz1 = (W2^T dz2) .* (h1 > 0), in place; three inputs k per pass share the dz2 loads, four consecutive j are one
  // 128-bit weight load (ascending j within each chain)
  static_assert(DQ_H1 % (3 * LEARN_G) == 0, "dz1 passes");
  for (int k0 = 3 * g; k0 < DQ_H1; k0 += 3 * LEARN_G) {
    float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
#pragma unroll 3
    for (int j = 0; j < DQ_H2; j += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(pq + DQ_W2 + j + DQ_H2 * (k0 + 0));
      const float4 w1 = *reinterpret_cast<const float4*>(pq + DQ_W2 + j + DQ_H2 * (k0 + 1));
      const float4 w2 = *reinterpret_cast<const float4*>(pq + DQ_W2 + j + DQ_H2 * (k0 + 2));
      const float z0 = h2[(j + 0) * SP + i], z1 = h2[(j + 1) * SP + i], z2 = h2[(j + 2) * SP + i], z3 = h2[(j + 3) * SP + i];
      d0 = fmaf(w0.x, z0, d0); d0 = fmaf(w0.y, z1, d0); d0 = fmaf(w0.z, z2, d0); d0 = fmaf(w0.w, z3, d0);
      d1 = fmaf(w1.x, z0, d1); d1 = fmaf(w1.y, z1, d1); d1 = fmaf(w1.z, z2, d1); d1 = fmaf(w1.w, z3, d1);
      d2 = fmaf(w2.x, z0, d2); d2 = fmaf(w2.y, z1, d2); d2 = fmaf(w2.z, z2, d2); d2 = fmaf(w2.w, z3, d2);
    }
    h1[(k0 + 0) * SP + i] = h1[(k0 + 0) * SP + i] > 0.0f ? d0 : 0.0f;
    h1[(k0 + 1) * SP + i] = h1[(k0 + 1) * SP + i] > 0.0f ? d1 : 0.0f;
    h1[(k0 + 2) * SP + i] = h1[(k0 + 2) * SP + i] > 0.0f ? d2 : 0.0f;
  }
  __syncthreads();
  for (int w = tid; w < DQ_H1 * DQ_D + DQ_H1; w += LEARN_T) {        // dW1(j,k), db1(j)
    if (w < DQ_H1 * DQ_D) {
      const int j = w % DQ_H1, k = w / DQ_H1;
      float acc = 0.0f;
      for (int b = 0; b < B; b++) acc = fmaf(h1[j * SP + b], xs[k * SP + b], acc);
      gr[DQ_W1 + w] = acc;
    } else {
      const int j = w - DQ_H1 * DQ_D;
      float acc = 0.0f;
      for (int b = 0; b < B; b++) acc += h1[j * SP + b];
      gr[DQ_B1 + j] = acc;
    }
  }
  __syncthreads();
  // ---- Flux.Adam(lr) (dqn.jl:41,109), Float64 scalars as in clip_adam_kernel; then the target copy (dqn.jl:111-113)
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  const double bp1 = a.dev->bp1, bp2 = a.dev->bp2;
  __syncthreads();
  for (int k = tid; k < DQ_P; k += LEARN_T) {
    const float d = gr[k];
    const float mt = (float)__dadd_rn(__dmul_rn(b1, (double)a.m[k]), __dmul_rn(1.0 - b1, (double)d));
    const float vt = (float)__dadd_rn(__dmul_rn(b2, (double)a.v[k]), __dmul_rn(__dmul_rn(1.0 - b2, (double)d), (double)d));
    a.m[k] = mt;
    a.v[k] = vt;
    const double den = __dadd_rn(sqrt(__ddiv_rn((double)vt, 1.0 - bp2)), eps);
    const float step = (float)__dmul_rn(__ddiv_rn(__ddiv_rn((double)mt, 1.0 - bp1), den), a.lr);
    const float pn = __fsub_rn(pq[k], step);
    a.q[k] = pn;
    if (a.copy_target) a.tgt[k] = pn;
  }
  if (tid == 0) { a.dev->bp1 = bp1 * b1; a.dev->bp2 = bp2 * b2; }
}


// ---- the learning step on several SMs -------------------------------------------------------------------------------
// dqn_learn_fwd_kernel: CTA c owns samples 16c..16c+15 of the batch (thread = sample x neuron group): gather, target
// and q forward, TD target, loss, dz2, dz1; everything the weight gradients need is written sample-major to a global
// scratch (L2-resident, 220 KB). dqn_learn_upd_kernel: one thread per 4x4 tile of dW2 (or per element of the small
// arrays) sums over the batch in ascending sample order - the same fmaf chains as dqn_learn_kernel - and applies Adam
// to the parameters it has just differentiated, so gradients never leave registers.
__global__ void __launch_bounds__(LF_T) dqn_learn_fwd_kernel(LearnArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PP = (DQ_P + 3) & ~3, SP = LF_SP;
  float* pq = smem;                       // q_net parameters
  float* pt = pq + PP;                    // target_net parameters
  float* h1 = pt + PP;                    // [120][SP]
  float* h2 = h1 + DQ_H1 * SP;            // [84][SP]   (later dz2 in place)
  float* qo = h2 + DQ_H2 * SP;            // [2][SP]
  float* dq = qo + DQ_A * SP;             // [2][SP]
  float* xin = dq + DQ_A * SP;            // [8][SP]: rows 0-3 state, rows 4-7 next_state
  __shared__ uint32_t keys[8];
  __shared__ float s_rew[LF_S];
  __shared__ int s_act[LF_S], s_term[LF_S];
  const int tid = threadIdx.x, B = a.B;
  const int i = tid & (LF_S - 1), g = tid / LF_S;   // sample within the CTA, neuron group
  const int b = blockIdx.x * LF_S + i;              // sample of the batch
  for (int k = tid; k < DQ_P; k += LF_T) { pq[k] = a.q[k]; pt[k] = a.tgt[k]; }
  if (tid == 0) {
    philox_draw(a.seed, 0u, a.learn_step, STREAM_DQN_BATCH, keys);
    philox_draw(a.seed, 0x80000000u, a.learn_step, STREAM_DQN_BATCH, keys + 4);
  }
  __syncthreads();
  if (g == 0) {
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), n4 = s4;
    int act = 0, term = 0;
    float rew = 0.0f;
    if (b < B) {
      // sample(1:size, B, replace=false), replay_buffer.jl:43: the first B entries of a keyed permutation of [0,size)
      const uint32_t idx = perm_index((uint32_t)b, (uint32_t)a.size, perm_half_bits((uint32_t)a.size), keys);
      s4 = reinterpret_cast<const float4*>(a.b_state)[idx];
      n4 = reinterpret_cast<const float4*>(a.b_next)[idx];
      act = a.b_action[idx];
      rew = a.b_reward[idx];
      term = a.b_term[idx];
    }
    xin[0 * SP + i] = s4.x; xin[1 * SP + i] = s4.y; xin[2 * SP + i] = s4.z; xin[3 * SP + i] = s4.w;
    xin[4 * SP + i] = n4.x; xin[5 * SP + i] = n4.y; xin[6 * SP + i] = n4.z; xin[7 * SP + i] = n4.w;
    s_rew[i] = rew; s_act[i] = act; s_term[i] = term;
  }
  __syncthreads();
  float s[4], nx[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { s[k] = xin[k * SP + i]; nx[k] = xin[(4 + k) * SP + i]; }
  q_forward<SP, LF_G>(pt, nx, h1, h2, qo, i, g);                     // target_net(next_state), dqn.jl:99
  const float next_q = fmaxf(qo[0 * SP + i], qo[1 * SP + i]);
  const double td = (double)s_rew[i] + a.gamma * (double)next_q * (1.0 - (double)s_term[i]);   // dqn.jl:100
  __syncthreads();
  q_forward<SP, LF_G>(pq, s, h1, h2, qo, i, g);                      // q_net(state), dqn.jl:105
  if (g == 0) {                                                      // lanes 0-15 of warp 0
    float d0 = 0.0f, d1 = 0.0f;
    double sq = 0.0;
    if (b < B) {
      const int act = s_act[i];
      const double diff = td - (double)qo[act * SP + i];
      sq = diff * diff;                                              // Flux.mse, dqn.jl:107
      const float dd = (float)(-2.0 * diff / (double)B);
      if (act == 0) d0 = dd; else d1 = dd;
      a.dqT[b * DQ_A + 0] = d0;
      a.dqT[b * DQ_A + 1] = d1;
#pragma unroll
      for (int k = 0; k < DQ_D; k++) a.xT[b * DQ_D + k] = s[k];
    }
    dq[0 * SP + i] = d0;
    dq[1 * SP + i] = d1;
#pragma unroll
    for (int o = LF_S / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0x0000ffffu, sq, o);
    if (i == 0) a.loss_part[blockIdx.x] = sq;
  }
  __syncthreads();
  {                                                                  // dz2 = (W3^T dq) .* (h2 > 0), in place
    const float d0 = dq[0 * SP + i], d1 = dq[1 * SP + i];
    for (int k = g; k < DQ_H2; k += LF_G) {
      const float hv = h2[k * SP + i];
      const float dh = fmaf(pq[DQ_W3 + 1 + DQ_A * k], d1, pq[DQ_W3 + 0 + DQ_A * k] * d0);
      const float z = hv > 0.0f ? dh : 0.0f;
      h2[k * SP + i] = z;
      if (b < B) { a.h2T[b * DQ_H2 + k] = hv; a.z2T[b * DQ_H2 + k] = z; }
    }
    if (b < B)
      for (int k = g; k < DQ_H1; k += LF_G) a.h1T[b * DQ_H1 + k] = h1[k * SP + i];
  }
  __syncthreads();
  {                                                                  // dz1 = (W2^T dz2) .* (h1 > 0) for k = g, g+32, g+64, g+96
    static_assert(4 * LF_G >= DQ_H1, "one dz1 pass covers the first hidden layer");
    float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int kk[4];
#pragma unroll
    for (int t = 0; t < 4; t++) kk[t] = min(g + LF_G * t, DQ_H1 - 1);   // clamped for the loads; stores are guarded
#pragma unroll 3
    for (int j = 0; j < DQ_H2; j += 4) {
      const float z0 = h2[(j + 0) * SP + i], z1 = h2[(j + 1) * SP + i], z2 = h2[(j + 2) * SP + i], z3 = h2[(j + 3) * SP + i];
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const float4 w = *reinterpret_cast<const float4*>(pq + DQ_W2 + j + DQ_H2 * kk[t]);
        d[t] = fmaf(w.x, z0, d[t]); d[t] = fmaf(w.y, z1, d[t]); d[t] = fmaf(w.z, z2, d[t]); d[t] = fmaf(w.w, z3, d[t]);
      }
    }
    if (b < B) {
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int k = g + LF_G * t;
        if (k < DQ_H1) a.z1T[b * DQ_H1 + k] = h1[k * SP + i] > 0.0f ? d[t] : 0.0f;
      }
    }
  }
}

// Flux.Adam(lr) (dqn.jl:41,109) for one parameter, Float64 scalars as in clip_adam_kernel; then the target copy
// (dqn.jl:111-113)
__device__ __forceinline__ void dqn_adam_one(const LearnArgs& a, int k, float d) {
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  const float mt = (float)__dadd_rn(__dmul_rn(b1, (double)a.m[k]), __dmul_rn(1.0 - b1, (double)d));
  const float vt = (float)__dadd_rn(__dmul_rn(b2, (double)a.v[k]), __dmul_rn(__dmul_rn(1.0 - b2, (double)d), (double)d));
  a.m[k] = mt;
  a.v[k] = vt;
  const double den = __dadd_rn(sqrt(__ddiv_rn((double)vt, 1.0 - a.bp2)), eps);
  const float step = (float)__dmul_rn(__ddiv_rn(__ddiv_rn((double)mt, 1.0 - a.bp1), den), a.lr);
  const float pn = __fsub_rn(a.q[k], step);
  a.q[k] = pn;
  if (a.copy_target) a.tgt[k] = pn;
}

__global__ void __launch_bounds__(LU_T) dqn_learn_upd_kernel(LearnArgs a, int n_fwd_blocks) {
  const int B = a.B;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double sq = 0.0;
    for (int c = 0; c < n_fwd_blocks; c++) sq += a.loss_part[c];
    a.dev->last_loss = sq / (double)B;
  }
  if (blockIdx.x < LU_TILE_BLOCKS) {                                 // dW2(j,k): neurons j, j+21, j+42, j+63 x inputs k0..k0+3
    const int w = blockIdx.x * LU_T + threadIdx.x;
    if (w >= LU_TILES) return;
    constexpr int JQ = DQ_H2 / 4;
    const int j = w % JQ, k0 = (w / JQ) * 4;
    float2 acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; t++) acc[t][0] = acc[t][1] = make_float2(0.0f, 0.0f);
#pragma unroll 4
    for (int b = 0; b < B; b++) {
      const float4 hh = *reinterpret_cast<const float4*>(a.h1T + b * DQ_H1 + k0);
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const float z = a.z2T[b * DQ_H2 + j + JQ * t];
        acc[t][0] = __ffma2_rn(make_float2(z, z), make_float2(hh.x, hh.y), acc[t][0]);
        acc[t][1] = __ffma2_rn(make_float2(z, z), make_float2(hh.z, hh.w), acc[t][1]);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int base = DQ_W2 + (j + JQ * t) + DQ_H2 * k0;
      dqn_adam_one(a, base + DQ_H2 * 0, acc[t][0].x);
      dqn_adam_one(a, base + DQ_H2 * 1, acc[t][0].y);
      dqn_adam_one(a, base + DQ_H2 * 2, acc[t][1].x);
      dqn_adam_one(a, base + DQ_H2 * 3, acc[t][1].y);
    }
    return;
  }
  const int w = (blockIdx.x - LU_TILE_BLOCKS) * LU_T + threadIdx.x;
  if (w >= LU_SINGLES) return;
  constexpr int N_W1 = DQ_H1 * DQ_D, N_B1 = N_W1 + DQ_H1, N_B2 = N_B1 + DQ_H2, N_W3 = N_B2 + DQ_A * DQ_H2;
  float acc = 0.0f;
  int pidx;
  if (w < N_W1) {                                                    // dW1(j,k)
    const int j = w % DQ_H1, k = w / DQ_H1;
    for (int b = 0; b < B; b++) acc = fmaf(a.z1T[b * DQ_H1 + j], a.xT[b * DQ_D + k], acc);
    pidx = DQ_W1 + w;
  } else if (w < N_B1) {                                             // db1(j)
    const int j = w - N_W1;
    for (int b = 0; b < B; b++) acc += a.z1T[b * DQ_H1 + j];
    pidx = DQ_B1 + j;
  } else if (w < N_B2) {                                             // db2(j)
    const int j = w - N_B1;
    for (int b = 0; b < B; b++) acc += a.z2T[b * DQ_H2 + j];
    pidx = DQ_B2 + j;
  } else if (w < N_W3) {                                             // dW3(o,k)
    const int ww = w - N_B2, o = ww % DQ_A, k = ww / DQ_A;
    for (int b = 0; b < B; b++) acc = fmaf(a.dqT[b * DQ_A + o], a.h2T[b * DQ_H2 + k], acc);
    pidx = DQ_W3 + ww;
  } else {                                                           // db3(o)
    const int o = w - N_W3;
    for (int b = 0; b < B; b++) acc += a.dqT[b * DQ_A + o];
    pidx = DQ_B3 + o;
  }
  dqn_adam_one(a, pidx, acc);
}

constexpr size_t ACT_SMEM = (((DQ_P + 3) & ~3) + (DQ_H1 + DQ_H2 + DQ_A + DQ_D) * ACT_E) * sizeof(float);
constexpr size_t LEARN_SMEM = (2 * ((DQ_P + 3) & ~3) + (DQ_H1 + DQ_H2 + DQ_D + 2 * DQ_A) * LEARN_SP) * sizeof(float);
static_assert(LEARN_SMEM <= 227 * 1024, "dqn_learn shared memory");
constexpr size_t LF_SMEM = (2 * ((DQ_P + 3) & ~3) + (DQ_H1 + DQ_H2 + 2 * DQ_A + 2 * DQ_D) * LF_SP) * sizeof(float);

int dfail(int code, const std::string& msg) { return crl_internal_fail(code, msg.c_str()); }
#define DCK(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess)                                                                                \
      return dfail(CRL_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (dqn.cu)"); \
  } while (0)

template <typename T> cudaError_t dzalloc(T** p, size_t n) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemset(*p, 0, (n ? n : 1) * sizeof(T));
}

}  // namespace

struct crl_dqn_ctx {
  crl_dqn_config cfg;
  cudaStream_t stream;
  float *q, *tgt, *m, *v, *g;
  float* env_state; int* env_t; double* ep_ret; int* ep_len; uint32_t* resets;
  float *b_state, *b_next, *b_reward; int* b_action; uint8_t* b_term;
  DqnDev* dev;
  float *h1T, *h2T, *z2T, *z1T, *xT, *dqT;   // scratch between dqn_learn_fwd_kernel and dqn_learn_upd_kernel
  double* loss_part;
  double bp1, bp2;                            // beta1^t, beta2^t of Adam (multi-SM learning step: kept on the host)
  int size, ptr;
  long long it, learn_steps, launches;
  bool params_set, reset_done, one_cta_learn;
};

__global__ void dqn_reset_kernel(int N, unsigned long long seed, float* env_state, int* env_t, double* ep_ret, int* ep_len,
                                 uint32_t* resets) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float u4[4], st[4];
  int t;
  rng_reset_uniforms(seed, (uint32_t)n, 0, u4);
  cartpole_reset(st, t, u4);
  for (int k = 0; k < 4; k++) env_state[4 * n + k] = st[k];
  env_t[n] = t; ep_ret[n] = 0.0; ep_len[n] = 0; resets[n] = 1;
}

static double linear_schedule(double start_e, double end_e, double duration, double t) {  // dqn.jl:28-31
  const double slope = (end_e - start_e) / duration;
  const double e = slope * t + start_e;
  return e > end_e ? e : end_e;
}

extern "C" CRL_API int crl_dqn_create(const crl_dqn_config* cfg, crl_dqn_ctx** out) {
  if (!cfg || !out) return dfail(CRL_ERR_INVALID, "NULL argument");
  if (cfg->struct_size != (int32_t)sizeof(crl_dqn_config)) return dfail(CRL_ERR_INVALID, "crl_dqn_config.struct_size mismatch");
  if (cfg->num_envs < 1 || cfg->buffer_size < cfg->num_envs) return dfail(CRL_ERR_INVALID, "need 1 <= num_envs <= buffer_size");
  if (cfg->batch_size < 1 || cfg->batch_size > LEARN_B) return dfail(CRL_ERR_INVALID, "batch_size must be in [1, 128]");
  if (cfg->train_freq < 1 || cfg->target_net_freq < 1 || cfg->max_episode_steps < 1 || !(cfg->epsilon_duration > 0.0))
    return dfail(CRL_ERR_INVALID, "train_freq, target_net_freq, max_episode_steps, epsilon_duration must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return dfail(CRL_ERR_CUDA, "no CUDA device: libcleanrl_cuda has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return dfail(CRL_ERR_INVALID, "device ordinal out of range");
  DCK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  DCK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return dfail(CRL_ERR_CUDA, "libcleanrl_cuda is built for sm_100a only");
  crl_dqn_ctx* c = new crl_dqn_ctx();
  memset(c, 0, sizeof(*c));
  c->cfg = *cfg;
  const size_t N = cfg->num_envs, C = cfg->buffer_size;
  DCK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  DCK(dzalloc(&c->q, DQ_P)); DCK(dzalloc(&c->tgt, DQ_P)); DCK(dzalloc(&c->m, DQ_P)); DCK(dzalloc(&c->v, DQ_P)); DCK(dzalloc(&c->g, DQ_P));
  DCK(dzalloc(&c->env_state, 4 * N)); DCK(dzalloc(&c->env_t, N)); DCK(dzalloc(&c->ep_ret, N)); DCK(dzalloc(&c->ep_len, N));
  DCK(dzalloc(&c->resets, N));
  DCK(dzalloc(&c->b_state, 4 * C)); DCK(dzalloc(&c->b_next, 4 * C)); DCK(dzalloc(&c->b_reward, C)); DCK(dzalloc(&c->b_action, C));
  DCK(dzalloc(&c->b_term, C)); DCK(dzalloc(&c->dev, 1));
  DCK(dzalloc(&c->h1T, (size_t)LEARN_B * DQ_H1)); DCK(dzalloc(&c->h2T, (size_t)LEARN_B * DQ_H2)); DCK(dzalloc(&c->z2T, (size_t)LEARN_B * DQ_H2));
  DCK(dzalloc(&c->z1T, (size_t)LEARN_B * DQ_H1)); DCK(dzalloc(&c->xT, (size_t)LEARN_B * DQ_D)); DCK(dzalloc(&c->dqT, (size_t)LEARN_B * DQ_A));
  DCK(dzalloc(&c->loss_part, LF_MAX_BLOCKS));
  DCK(cudaFuncSetAttribute(dqn_learn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LF_SMEM));
  {
    const char* e = getenv("CRL_DQN_ONE_CTA");   // A/B switch: the whole learning step in one CTA (dqn_learn_kernel)
    c->one_cta_learn = e && atoi(e) != 0;
  }
  DCK(cudaFuncSetAttribute(dqn_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACT_SMEM));
  DCK(cudaFuncSetAttribute(dqn_learn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEARN_SMEM));
  *out = c;
  return CRL_OK;
}

extern "C" CRL_API int crl_dqn_destroy(crl_dqn_ctx* c) {
  if (!c) return dfail(CRL_ERR_INVALID, "ctx is NULL");
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  void* ptrs[] = {c->q, c->tgt, c->m, c->v, c->g, c->env_state, c->env_t, c->ep_ret, c->ep_len, c->resets, c->b_state, c->b_next,
                  c->b_reward, c->b_action, c->b_term, c->dev, c->h1T, c->h2T, c->z2T, c->z1T, c->xT, c->dqT, c->loss_part};
  for (void* p : ptrs) if (p) cudaFree(p);
  cudaStreamDestroy(c->stream);
  delete c;
  return CRL_OK;
}

extern "C" CRL_API int crl_dqn_set_params(crl_dqn_ctx* c, const float* params, int32_t n) {
  if (!c || !params) return dfail(CRL_ERR_INVALID, "NULL argument");
  if (n != DQ_P) return dfail(CRL_ERR_INVALID, "DQN parameter vector must have 10934 floats");
  DCK(cudaSetDevice(c->cfg.device));
  DCK(cudaMemcpyAsync(c->q, params, sizeof(float) * DQ_P, cudaMemcpyHostToDevice, c->stream));
  DCK(cudaMemcpyAsync(c->tgt, c->q, sizeof(float) * DQ_P, cudaMemcpyDeviceToDevice, c->stream));   // deepcopy, dqn.jl:40
  DCK(cudaMemsetAsync(c->m, 0, sizeof(float) * DQ_P, c->stream));
  DCK(cudaMemsetAsync(c->v, 0, sizeof(float) * DQ_P, c->stream));
  DqnDev d;
  memset(&d, 0, sizeof(d));
  d.bp1 = 0.9; d.bp2 = 0.999;
  DCK(cudaMemcpyAsync(c->dev, &d, sizeof(d), cudaMemcpyHostToDevice, c->stream));
  DCK(cudaStreamSynchronize(c->stream));
  c->bp1 = 0.9; c->bp2 = 0.999;
  c->params_set = true;
  return CRL_OK;
}

extern "C" CRL_API int crl_dqn_get_params(crl_dqn_ctx* c, float* q_params, float* target_params, int32_t n) {
  if (!c || !q_params) return dfail(CRL_ERR_INVALID, "NULL argument");
  if (n != DQ_P) return dfail(CRL_ERR_INVALID, "DQN parameter vector must have 10934 floats");
  DCK(cudaSetDevice(c->cfg.device));
  DCK(cudaMemcpyAsync(q_params, c->q, sizeof(float) * DQ_P, cudaMemcpyDeviceToHost, c->stream));
  if (target_params) DCK(cudaMemcpyAsync(target_params, c->tgt, sizeof(float) * DQ_P, cudaMemcpyDeviceToHost, c->stream));
  DCK(cudaStreamSynchronize(c->stream));
  return CRL_OK;
}

extern "C" CRL_API int crl_dqn_reset(crl_dqn_ctx* c) {
  if (!c) return dfail(CRL_ERR_INVALID, "ctx is NULL");
  DCK(cudaSetDevice(c->cfg.device));
  const int N = c->cfg.num_envs;
  dqn_reset_kernel<<<(N + 127) / 128, 128, 0, c->stream>>>(N, c->cfg.seed, c->env_state, c->env_t, c->ep_ret, c->ep_len, c->resets);
  DCK(cudaGetLastError());
  c->size = 0; c->ptr = 0; c->it = 0; c->learn_steps = 0;   /* launches keeps counting: it is a lifetime counter */
  c->reset_done = true;
  return CRL_OK;
}

extern "C" CRL_API int crl_dqn_run(crl_dqn_ctx* c, int64_t iterations, crl_dqn_stats* stats) {
  if (!c) return dfail(CRL_ERR_INVALID, "ctx is NULL");
  if (iterations < 0) return dfail(CRL_ERR_INVALID, "iterations must be >= 0");
  if (!c->params_set) return dfail(CRL_ERR_STATE, "crl_dqn_run called before crl_dqn_set_params");
  if (!c->reset_done) return dfail(CRL_ERR_STATE, "crl_dqn_run called before crl_dqn_reset");
  DCK(cudaSetDevice(c->cfg.device));
  const int N = c->cfg.num_envs, C = c->cfg.buffer_size;
  // per-call episode aggregate: clear the three accumulators, keep Adam's powers and the last loss
  DCK(cudaMemsetAsync(&c->dev->sum_return, 0, sizeof(double) * 2 + sizeof(unsigned long long), c->stream));
  double eps = 0.0;
  int64_t k = 0;
  while (k < iterations) {
    // One launch acts for every iteration up to (and including) the next learning step: the parameters are constant in
    // between. Bounded by ACT_MAX_STEPS and by the ring capacity (steps of one launch must not share a slot).
    ActArgs a;
    a.q = c->q; a.env_state = c->env_state; a.env_t = c->env_t; a.ep_ret = c->ep_ret; a.ep_len = c->ep_len; a.resets = c->resets;
    a.b_state = c->b_state; a.b_next = c->b_next; a.b_reward = c->b_reward; a.b_action = c->b_action; a.b_term = c->b_term;
    a.dev = c->dev; a.seed = c->cfg.seed; a.it0 = (unsigned long long)(c->it + 1); a.N = N; a.C = C; a.ptr = c->ptr;
    a.max_steps = c->cfg.max_episode_steps;
    int ns = 0;
    bool learn = false;
    while (k < iterations && ns < ACT_MAX_STEPS && (long long)(ns + 1) * N <= (long long)C && !learn) {
      c->it += 1;
      k += 1;
      const double gs = (double)c->it * (double)N;
      eps = linear_schedule(c->cfg.epsilon_start, c->cfg.epsilon_end, c->cfg.epsilon_duration, gs);   // dqn.jl:52
      a.eps[ns++] = eps;
      c->ptr = (c->ptr + N) % C;
      c->size = c->size + N > C ? C : c->size + N;
      learn = gs > (double)c->cfg.min_buff_size && c->it % c->cfg.train_freq == 0 && c->size >= c->cfg.batch_size;   // dqn.jl:94
    }
    for (int i = ns; i < ACT_MAX_STEPS; i++) a.eps[i] = 0.0;
    a.n_steps = ns;
    dqn_act_kernel<<<(N + ACT_E - 1) / ACT_E, ACT_T, ACT_SMEM, c->stream>>>(a);
    DCK(cudaGetLastError());
    c->launches += 1;
    if (learn) {
      LearnArgs l;
      memset(&l, 0, sizeof(l));
      l.q = c->q; l.tgt = c->tgt; l.m = c->m; l.v = c->v; l.g = c->g;
      l.b_state = c->b_state; l.b_next = c->b_next; l.b_reward = c->b_reward; l.b_action = c->b_action; l.b_term = c->b_term;
      l.dev = c->dev; l.seed = c->cfg.seed; l.learn_step = (unsigned long long)c->learn_steps; l.gamma = c->cfg.gamma;
      l.lr = c->cfg.lr; l.B = c->cfg.batch_size; l.size = c->size;
      l.copy_target = (c->it % c->cfg.target_net_freq == 0) ? 1 : 0;                                             // dqn.jl:111
      if (c->one_cta_learn) {
        dqn_learn_kernel<<<1, LEARN_T, LEARN_SMEM, c->stream>>>(l);
        DCK(cudaGetLastError());
        c->launches += 1;
      } else {
        l.h1T = c->h1T; l.h2T = c->h2T; l.z2T = c->z2T; l.z1T = c->z1T; l.xT = c->xT; l.dqT = c->dqT;
        l.loss_part = c->loss_part; l.bp1 = c->bp1; l.bp2 = c->bp2;
        const int fwd_blocks = (l.B + LF_S - 1) / LF_S;
        dqn_learn_fwd_kernel<<<fwd_blocks, LF_T, LF_SMEM, c->stream>>>(l);
        DCK(cudaGetLastError());
        dqn_learn_upd_kernel<<<LU_TILE_BLOCKS + LU_SINGLE_BLOCKS, LU_T, 0, c->stream>>>(l, fwd_blocks);
        DCK(cudaGetLastError());
        c->bp1 *= 0.9; c->bp2 *= 0.999;   // same Float64 products the one-CTA kernel keeps on the device
        c->launches += 2;
      }
      c->learn_steps += 1;
    }
  }
  if (stats) {
    DqnDev d;
    DCK(cudaMemcpyAsync(&d, c->dev, sizeof(d), cudaMemcpyDeviceToHost, c->stream));
    DCK(cudaStreamSynchronize(c->stream));
    stats->last_loss = d.last_loss; stats->sum_return = d.sum_return; stats->sum_length = d.sum_length; stats->epsilon = eps;
    stats->episodes = (int64_t)d.episodes; stats->learn_steps = c->learn_steps; stats->iterations = c->it;
    stats->kernel_launches = c->launches;
  }
  return CRL_OK;
}

extern "C" CRL_API int crl_dqn_read_buffer(crl_dqn_ctx* c, float* state, int32_t* action, float* reward, float* next_state,
                                           uint8_t* terminal, int32_t* size, int32_t* ptr) {
  if (!c) return dfail(CRL_ERR_INVALID, "ctx is NULL");
  DCK(cudaSetDevice(c->cfg.device));
  const size_t C = c->cfg.buffer_size;
  if (state) DCK(cudaMemcpyAsync(state, c->b_state, C * 16, cudaMemcpyDeviceToHost, c->stream));
  if (next_state) DCK(cudaMemcpyAsync(next_state, c->b_next, C * 16, cudaMemcpyDeviceToHost, c->stream));
  if (action) DCK(cudaMemcpyAsync(action, c->b_action, C * 4, cudaMemcpyDeviceToHost, c->stream));
  if (reward) DCK(cudaMemcpyAsync(reward, c->b_reward, C * 4, cudaMemcpyDeviceToHost, c->stream));
  if (terminal) DCK(cudaMemcpyAsync(terminal, c->b_term, C, cudaMemcpyDeviceToHost, c->stream));
  DCK(cudaStreamSynchronize(c->stream));
  if (size) *size = c->size;
  if (ptr) *ptr = c->ptr;
  return CRL_OK;
}
