/*
 * ppo_oracle.c — CPU restatement of the PPO hot path of sash-a/CleanRL.jl.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product path
 * (libcleanrl_cuda.so) never links, loads or calls anything in oracle/.
 *
 * What it follows (paths relative to the reference repo root):
 *   src/algorithms/ppo.jl:21-32    get_action            -> orc_policy_forward + sample_*
 *   src/algorithms/ppo.jl:34-45    logprob_actions       -> inside orc_loss_grad
 *   src/algorithms/ppo.jl:48-73    gae                   -> orc_gae_raw
 *   src/algorithms/ppo.jl:117-166  rollout loop          -> orc_rollout
 *   src/algorithms/ppo.jl:169-189  bootstrap/GAE/flatten -> orc_gae
 *   src/algorithms/ppo.jl:191-252  epochs/minibatches    -> orc_update_epochs / orc_update_minibatch
 *   src/utils/multi_thread_env.jl:86-133  batched env    -> env arrays inside orc_ctx
 *   src/utils/replay_buffer.jl:15-37      buffer layout  -> [T][N] arrays inside orc_ctx
 *   src/utils/networks.jl:6-13,36-49      two 64-64 MLPs -> mlp_forward / mlp_backward
 *
 * PARITY STATUS. The reference has no tests, no golden vectors, and Julia is not present
 * in this image, so the reference itself cannot be run. What is pinned:
 *   - gae: known-answer vectors derived by hand from ppo.jl:62-72;
 *   - Philox4x32-10: Random123 known-answer vectors;
 *   - loss gradient: cross-checked against a float64 PyTorch autograd restatement of
 *     ppo.jl:213-243 (tests/test_oracle_grad.py).
 * PARITY UNPINNED (restated from memory of un-vendored, version-pinned dependencies;
 * Manifest.toml lines in SURVEY §2.2):
 *   - CartPoleEnv / PendulumEnv dynamics: ReinforcementLearningEnvironments 0.6.12;
 *   - tanh_fast, softmax, logsoftmax: NNlib 0.8.21 (coefficients validated only as
 *     "approximates tanh to < 3e-7 relative");
 *   - Dense/ClipNorm/Adam/update!: Flux 0.13.4; sample(Weights): StatsBase 0.33.21;
 *   - the Gaussian policy (Pendulum) follows CleanRL-Python conventions: the reference
 *     has no continuous-action PPO at all.
 * Random draws are Philox (the reference uses Xoshiro; SURVEY §7.2 "RNG"), so random
 * streams are a contract between this oracle and the CUDA library, not the reference.
 *
 * Float types follow Julia's promotion rules at each cited line; reductions the
 * reference does in Float32 with a BLAS/pairwise order that cannot be pinned are
 * accumulated in double here and rounded where the reference's result type is Float32.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cleanrl_cuda.h"

#define H 64 /* hidden width, networks.jl:36 */
#define MAXD 4
#define MAXA 2
#define MAX_ARRAYS 13

/* ------------------------------------------------------------------ threading */
static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_get_threads(void) { return g_threads; }

typedef void (*range_fn)(int64_t lo, int64_t hi, int tid, void* arg);
typedef struct {
  range_fn fn;
  void* arg;
  int64_t lo, hi;
  int tid;
} job_t;
static void* job_main(void* p) {
  job_t* j = (job_t*)p;
  j->fn(j->lo, j->hi, j->tid, j->arg);
  return NULL;
}
/* one unit of work per env, as multi_thread_env.jl:88-96 spawns one task per env */
static void parallel_for(int64_t n, range_fn fn, void* arg) {
  int nt = g_threads;
  if (nt > n) nt = (int)(n > 0 ? n : 1);
  if (nt <= 1) {
    fn(0, n, 0, arg);
    return;
  }
  pthread_t th[256];
  job_t jobs[256];
  if (nt > 256) nt = 256;
  for (int i = 0; i < nt; i++) {
    jobs[i].fn = fn;
    jobs[i].arg = arg;
    jobs[i].lo = n * i / nt;
    jobs[i].hi = n * (i + 1) / nt;
    jobs[i].tid = i;
    pthread_create(&th[i], NULL, job_main, &jobs[i]);
  }
  for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* RNG contract shared with the CUDA library (csrc/philox.cuh):
 *   key = (seed_lo, seed_hi); ctr = (global env id, counter_lo, counter_hi, stream)
 *   stream 0 = action noise (counter = global policy step), 1 = env reset (counter =
 *   number of resets of that env so far), 2 = minibatch permutation keys. */
#define ORC_STREAM_ACTION 0u
#define ORC_STREAM_RESET 1u
#define ORC_STREAM_PERM 2u
static void philox_draw(uint64_t seed, uint32_t a, uint64_t counter, uint32_t stream, uint32_t out[4]) {
  uint32_t ctr[4] = {a, (uint32_t)counter, (uint32_t)(counter >> 32), stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  orc_philox4x32_10(ctr, key, out);
}
/* the Float64 rand() inside StatsBase.sample(Weights), ppo.jl:26 */
double orc_rng_action_uniform(uint64_t seed, uint32_t env, uint64_t step) {
  uint32_t r[4];
  philox_draw(seed, env, step, ORC_STREAM_ACTION, r);
  uint64_t x = ((uint64_t)r[0] << 32) | r[1];
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
/* Box-Muller in float; one normal per action dim (A <= 2) */
void orc_rng_action_normals(uint64_t seed, uint32_t env, uint64_t step, double z[2]) {
  uint32_t r[4];
  philox_draw(seed, env, step, ORC_STREAM_ACTION, r);
  for (int a = 0; a < 2; a++) {
    float u1 = (float)((r[2 * a] >> 8) + 1u) * (1.0f / 16777216.0f); /* (0,1] */
    float u2 = (float)(r[2 * a + 1] >> 8) * (1.0f / 16777216.0f);    /* [0,1) */
    float rad = sqrtf(-2.0f * logf(u1));
    z[a] = (double)(rad * cosf(6.283185307179586f * u2));
  }
}
/* rand(rng, Float32, 4) of CartPole reset! [RLEnvs 0.6.12] */
void orc_rng_reset_uniforms(uint64_t seed, uint32_t env, uint64_t k, float u[4]) {
  uint32_t r[4];
  philox_draw(seed, env, k, ORC_STREAM_RESET, r);
  for (int i = 0; i < 4; i++) u[i] = (float)(r[i] >> 8) * (1.0f / 16777216.0f);
}

/* ------------------------------------------------------------------ layout */
typedef struct {
  int D, A, S, continuous, P, n_arrays;
  int off[MAX_ARRAYS], size[MAX_ARRAYS];
  int actor, critic, logstd; /* offsets */
} layout_t;

/* Flux.params(actor, critic) order, ppo.jl:196; networks.jl:36-49 */
static int make_layout(int env_kind, layout_t* L) {
  memset(L, 0, sizeof(*L));
  if (env_kind == CRL_ENV_CARTPOLE) { L->D = 4; L->A = 2; L->S = 4; L->continuous = 0; }
  else if (env_kind == CRL_ENV_PENDULUM) { L->D = 3; L->A = 1; L->S = 2; L->continuous = 1; }
  else return -1;
  int o = 0, i = 0;
  for (int net = 0; net < 2; net++) {
    int O = net == 0 ? L->A : 1;
    if (net == 0) L->actor = o; else L->critic = o;
    int sizes[6] = {H * L->D, H, H * H, H, O * H, O};
    for (int k = 0; k < 6; k++) { L->off[i] = o; L->size[i] = sizes[k]; o += sizes[k]; i++; }
  }
  L->logstd = -1;
  if (L->continuous) { L->logstd = o; L->off[i] = o; L->size[i] = L->A; o += L->A; i++; }
  L->P = o;
  L->n_arrays = i;
  return 0;
}
int orc_dims(int env_kind, int32_t* D, int32_t* A, int32_t* S, int32_t* P, int32_t* n_arrays) {
  layout_t L;
  if (make_layout(env_kind, &L)) return CRL_ERR_INVALID;
  *D = L.D; *A = L.A; *S = L.S; *P = L.P; *n_arrays = L.n_arrays;
  return 0;
}
int orc_param_layout(int env_kind, int32_t* offsets, int32_t* sizes) {
  layout_t L;
  if (make_layout(env_kind, &L)) return CRL_ERR_INVALID;
  for (int i = 0; i < L.n_arrays; i++) { offsets[i] = L.off[i]; sizes[i] = L.size[i]; }
  return 0;
}

/* ------------------------------------------------------------------ activations */
/* NNlib.tanh_fast(x::Float32) [NNlib 0.8.21, called through networks.jl:6]: rational
 * x*n(x^2)/d(x^2), evalpoly = Horner with muladd (fma on any FMA CPU). */
float orc_tanh_fast(float x) {
  float x2 = x * x;
  float n = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 1.587199e-8f, 2.2332108e-5f), 0.0035974074f), 0.1346604f), 1.0f);
  float d = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 8.7767893e-7f, 0.0003453992f), 0.026262015f), 0.4679937f), 1.0f);
  if (x2 < 66.0f) return x * (n / d);
  return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : x);
}

/* Flux Dense chain: tanh_fast.(W*x .+ b) twice, identity head (networks.jl:6-13,42-46).
 * W is (out,in) column-major. */
static void mlp_forward(const float* p, int D, int O, const float* x, float* h1, float* h2, float* out) {
  const float* W1 = p;
  const float* b1 = W1 + H * D;
  const float* W2 = b1 + H;
  const float* b2 = W2 + H * H;
  const float* W3 = b2 + H;
  const float* b3 = W3 + O * H;
  for (int j = 0; j < H; j++) {
    float acc = 0.0f;
    for (int k = 0; k < D; k++) acc += W1[j + H * k] * x[k];
    h1[j] = orc_tanh_fast(acc + b1[j]);
  }
  for (int j = 0; j < H; j++) {
    float acc = 0.0f;
    for (int k = 0; k < H; k++) acc += W2[j + H * k] * h1[k];
    h2[j] = orc_tanh_fast(acc + b2[j]);
  }
  for (int o = 0; o < O; o++) {
    float acc = 0.0f;
    for (int k = 0; k < H; k++) acc += W3[o + O * k] * h2[k];
    out[o] = acc + b3[o];
  }
}

/* reverse pass of the chain for one sample; grads accumulate in double.
 * tanh_fast pullback: dx = dy * (1 - y^2) [NNlib @scalar_rule]. */
static void mlp_backward(const float* p, double* g, int D, int O, const float* x, const float* h1,
                         const float* h2, const float* dout) {
  const float* W2 = p + H * D + H;
  const float* W3 = W2 + H * H + H;
  double* gW1 = g;
  double* gb1 = gW1 + H * D;
  double* gW2 = gb1 + H;
  double* gb2 = gW2 + H * H;
  double* gW3 = gb2 + H;
  double* gb3 = gW3 + O * H;
  float dz2[H], dz1[H];
  for (int o = 0; o < O; o++) gb3[o] += dout[o];
  for (int k = 0; k < H; k++) {
    float acc = 0.0f;
    for (int o = 0; o < O; o++) {
      gW3[o + O * k] += (double)(dout[o] * h2[k]);
      acc += W3[o + O * k] * dout[o];
    }
    dz2[k] = acc * (1.0f - h2[k] * h2[k]);
  }
  for (int j = 0; j < H; j++) gb2[j] += dz2[j];
  for (int k = 0; k < H; k++) {
    float acc = 0.0f;
    for (int j = 0; j < H; j++) {
      gW2[j + H * k] += (double)(dz2[j] * h1[k]);
      acc += W2[j + H * k] * dz2[j];
    }
    dz1[k] = acc * (1.0f - h1[k] * h1[k]);
  }
  for (int j = 0; j < H; j++) gb1[j] += dz1[j];
  for (int k = 0; k < D; k++)
    for (int j = 0; j < H; j++) gW1[j + H * k] += (double)(dz1[j] * x[k]);
}

/* NNlib softmax / logsoftmax over one column (ppo.jl:23-24,36-37) */
static void softmax_logsoftmax(const float* z, int A, float* p, float* lp) {
  float m = z[0];
  for (int a = 1; a < A; a++) m = z[a] > m ? z[a] : m;
  float e[MAXA], s = 0.0f;
  for (int a = 0; a < A; a++) { e[a] = expf(z[a] - m); s += e[a]; }
  float ls = logf(s);
  for (int a = 0; a < A; a++) { p[a] = e[a] / s; lp[a] = z[a] - m - ls; }
}

/* ------------------------------------------------------------------ environments */
/* CartPoleEnv(T=Float32) step [RLEnvs 0.6.12; called at multi_thread_env.jl:91]. Params are
 * Float32; the literal 4/3 is Float64, so thetaacc, xacc and both velocity updates are
 * evaluated in Float64 and rounded on store. action is 0-based here:
 * Julia a==2 (push right) <=> action==1. */
void orc_cartpole_step(float* s, int32_t* t, int action, int max_steps, float* reward, uint8_t* done) {
  const float gravity = 9.8f, masscart = 1.0f, masspole = 0.1f, halflength = 0.5f, forcemag = 10.0f, dt = 0.02f;
  const float totalmass = masscart + masspole;
  const float polemasslength = masspole * halflength;
  const float thetathreshold = (float)(12.0 * 2.0 * 3.141592653589793 / 360.0);
  const float xthreshold = 2.4f;
  *t += 1;
  float force = action == 1 ? forcemag : -forcemag;
  float x = s[0], xdot = s[1], theta = s[2], thetadot = s[3];
  (void)x;
  float costheta = cosf(theta), sintheta = sinf(theta);
  float tmp = (force + polemasslength * (thetadot * thetadot) * sintheta) / totalmass;
  double thetaacc = (double)(gravity * sintheta - costheta * tmp) /
                    ((double)halflength * (4.0 / 3.0 - (double)(masspole * (costheta * costheta) / totalmass)));
  double xacc = (double)tmp - (double)polemasslength * thetaacc * (double)costheta / (double)totalmass;
  s[0] = s[0] + dt * xdot;
  s[1] = (float)((double)s[1] + (double)dt * xacc);
  s[2] = s[2] + dt * thetadot;
  s[3] = (float)((double)s[3] + (double)dt * thetaacc);
  int d = fabsf(s[0]) > xthreshold || fabsf(s[2]) > thetathreshold || *t > max_steps;
  *done = (uint8_t)d;
  *reward = d ? 0.0f : 1.0f;
}
/* reset!: state = 0.1f0*rand(rng,Float32,4) .- 0.05f0; t = 0 */
void orc_cartpole_reset(float* s, int32_t* t, const float u[4]) {
  for (int i = 0; i < 4; i++) s[i] = 0.1f * u[i] - 0.05f;
  *t = 0;
}
/* PendulumEnv(T=Float32) _step! [RLEnvs 0.6.12]. state = (theta, thetadot); costs are
 * Float64 because angle_normalize mixes in 2π::Float64; the dynamics stay Float32. */
void orc_pendulum_step(float* s, int32_t* t, float a, int max_steps, float* reward, uint8_t* done) {
  const float max_speed = 8.0f, max_torque = 2.0f, g = 10.0f, m = 1.0f, l = 1.0f, dt = 0.05f;
  *t += 1;
  float th = s[0], thdot = s[1];
  a = a < -max_torque ? -max_torque : (a > max_torque ? max_torque : a);
  const double PI = 3.141592653589793;
  float xp = th + (float)PI; /* Float32 + π -> Float32 */
  double an = fmod((double)xp, 2.0 * PI);
  if (an < 0.0) an += 2.0 * PI; /* Base.mod: result has the sign of the divisor */
  an -= PI;
  double costs = an * an + 0.1 * (double)(thdot * thdot) + 0.001 * (double)(a * a);
  float newthdot = thdot + (-3.0f * g / (2.0f * l) * sinf(th + (float)PI) + 3.0f * a / (m * (l * l))) * dt;
  th = th + newthdot * dt;
  newthdot = newthdot < -max_speed ? -max_speed : (newthdot > max_speed ? max_speed : newthdot);
  s[0] = th;
  s[1] = newthdot;
  *done = (uint8_t)(*t >= max_steps);
  *reward = (float)(-costs);
}
/* reset!: theta = 2π*(rand(T)-1) (Float64 product, stored Float32); thetadot = 2*(rand(T)-1) */
void orc_pendulum_reset(float* s, int32_t* t, const float u[4]) {
  s[0] = (float)(2.0 * 3.141592653589793 * (double)(u[0] - 1.0f));
  s[1] = 2.0f * (u[1] - 1.0f);
  *t = 0;
}
static void env_obs(int env_kind, const float* s, float* obs) {
  if (env_kind == CRL_ENV_CARTPOLE) { for (int i = 0; i < 4; i++) obs[i] = s[i]; }
  else { obs[0] = cosf(s[0]); obs[1] = sinf(s[0]); obs[2] = s[1]; }
}
/* batch wrappers for tests (mirror crl_env_step_raw) */
int orc_env_step_raw(int32_t env_kind, float* state, int32_t* t, const void* action, float* reward,
                     uint8_t* done, int64_t n, int32_t max_steps) {
  for (int64_t i = 0; i < n; i++) {
    if (env_kind == CRL_ENV_CARTPOLE)
      orc_cartpole_step(state + 4 * i, t + i, ((const int32_t*)action)[i], max_steps, reward + i, done + i);
    else if (env_kind == CRL_ENV_PENDULUM)
      orc_pendulum_step(state + 2 * i, t + i, ((const float*)action)[i], max_steps, reward + i, done + i);
    else return CRL_ERR_INVALID;
  }
  return 0;
}
int orc_env_reset_raw(int32_t env_kind, float* state, int32_t* t, const float* u4, int64_t n) {
  for (int64_t i = 0; i < n; i++) {
    if (env_kind == CRL_ENV_CARTPOLE) orc_cartpole_reset(state + 4 * i, t + i, u4 + 4 * i);
    else orc_pendulum_reset(state + 2 * i, t + i, u4 + 4 * i);
  }
  return 0;
}

/* actor+critic forward for n observations (get_action lines ppo.jl:22-24 and ppo.jl:128) */
int orc_policy_forward_raw(int32_t env_kind, const float* params, const float* obs, float* out_policy,
                           float* logp, float* value, int64_t n) {
  layout_t L;
  if (make_layout(env_kind, &L)) return CRL_ERR_INVALID;
  for (int64_t i = 0; i < n; i++) {
    float h1[H], h2[H], z[MAXA], p[MAXA], lp[MAXA], v;
    mlp_forward(params + L.actor, L.D, L.A, obs + L.D * i, h1, h2, z);
    mlp_forward(params + L.critic, L.D, 1, obs + L.D * i, h1, h2, &v);
    for (int a = 0; a < L.A; a++) out_policy[L.A * i + a] = z[a];
    if (!L.continuous) {
      softmax_logsoftmax(z, L.A, p, lp);
      for (int a = 0; a < L.A; a++) logp[L.A * i + a] = lp[a];
    }
    value[i] = v;
  }
  return 0;
}

/* ------------------------------------------------------------------ GAE */
/* gae(), ppo.jl:48-73, for N independent envs + returns = advantages + values (ppo.jl:181).
 * nonterm = 1.0 .- terminals is Float64 (ppo.jl:63) and gae = 0.0 is Float64 (ppo.jl:65), so
 * the recurrence runs in Float64; γ*λ is a Float32 product first (left-assoc, ppo.jl:68).
 * REF_COMPAT reproduces the range quirk at ppo.jl:66 (Q1): `length(rewards)-1:-1:1` is
 * (T-1):-1:1, so adv[T] is never written (we define it as 0) and the bootstrap is unused. */
typedef struct {
  const float *values, *rewards, *next_value;
  const uint8_t *dones, *next_done;
  float *adv, *ret;
  int T;
  int64_t N;
  float gamma, lambda;
  int mode;
} gae_args;
static void gae_range(int64_t lo, int64_t hi, int tid, void* argp) {
  (void)tid;
  gae_args* a = (gae_args*)argp;
  const int T = a->T;
  const int64_t N = a->N;
  const float gl = a->gamma * a->lambda; /* Float32 product */
  for (int64_t n = lo; n < hi; n++) {
    if (a->mode == CRL_GAE_A2C_RETURNS) {
      /* discounted_future_rewards, a2c.jl:13-24 (Float64 there): future[t] = terminals[t] ? 0 : r[t] + γ future[t+1],
       * seeded with final_value; terminals[t] is is_terminated AFTER step t = the flag stored one row later here.
       * advantage = discounted_rewards - values (a2c.jl:83). */
      double fut = (double)a->next_value[n];
      for (int t = T - 1; t >= 0; t--) {
        const int term = t == T - 1 ? a->next_done[n] : a->dones[(int64_t)(t + 1) * N + n];
        fut = term ? 0.0 : (double)a->rewards[(int64_t)t * N + n] + (double)a->gamma * fut;
        a->ret[(int64_t)t * N + n] = (float)fut;
        a->adv[(int64_t)t * N + n] = (float)(fut - (double)a->values[(int64_t)t * N + n]);
      }
      continue;
    }
    double gae = 0.0;
    int tstart;
    if (a->mode == CRL_GAE_REF_COMPAT) {
      a->adv[(int64_t)(T - 1) * N + n] = 0.0f;
      a->ret[(int64_t)(T - 1) * N + n] = 0.0f + a->values[(int64_t)(T - 1) * N + n];
      tstart = T - 2;
    } else {
      tstart = T - 1;
    }
    for (int t = tstart; t >= 0; t--) {
      double nonterm, vnext;
      if (t == T - 1) { nonterm = 1.0 - (double)a->next_done[n]; vnext = (double)a->next_value[n]; }
      else { nonterm = 1.0 - (double)a->dones[(int64_t)(t + 1) * N + n]; vnext = (double)a->values[(int64_t)(t + 1) * N + n]; }
      double v = (double)a->values[(int64_t)t * N + n];
      double delta = (double)a->rewards[(int64_t)t * N + n] + ((double)a->gamma * nonterm) * vnext - v;
      gae = delta + (((double)gl * nonterm) * gae);
      float af = (float)gae;
      a->adv[(int64_t)t * N + n] = af;
      a->ret[(int64_t)t * N + n] = af + a->values[(int64_t)t * N + n];
    }
  }
}
int orc_gae_raw(const float* values, const float* rewards, const uint8_t* dones, const float* next_value,
                const uint8_t* next_done, float* adv, float* ret, int32_t T, int64_t N, float gamma,
                float lambda, int32_t mode) {
  if (T < 1 || N < 0) return CRL_ERR_INVALID;
  gae_args a = {values, rewards, next_value, dones, next_done, adv, ret, T, N, gamma, lambda, mode};
  parallel_for(N, gae_range, &a);
  return 0;
}

/* ------------------------------------------------------------------ context */
typedef struct orc_ctx {
  crl_config cfg;
  layout_t L;
  int N, T, B;
  float *params, *grads, *adam_m, *adam_v;
  double* beta_pow; /* [n_arrays][2] */
  /* envs */
  float* env_state; /* [N][S] */
  int32_t* env_t;
  double* ep_return;
  int32_t* ep_length;
  uint32_t* reset_count;
  float* next_obs; /* [N][D] */
  uint8_t* next_done;
  float* next_value;
  uint64_t policy_step; /* global policy step counter (Philox action stream) */
  uint64_t update_index;
  /* rollout buffer */
  float* state; /* [T][N][D] */
  void* action; /* int32 [T][N] or float [T][N][A] */
  float *logprob, *reward, *value, *advantage, *ret;
  uint8_t* terminal;
  /* episode records of the last rollout */
  uint8_t* ep_done; /* [T][N] */
  double* ep_rec_return;
  int32_t* ep_rec_length;
  float* vnew; /* [M] */
  int rolled, gae_done;
} orc_ctx;

static int check_cfg(const crl_config* c) {
  if (!c || c->struct_size != (int32_t)sizeof(crl_config)) return -1;
  if (c->num_envs < 1 || c->num_steps < 1 || c->num_minibatches < 1 || c->update_epochs < 0) return -1;
  if (((int64_t)c->num_envs * c->num_steps) % c->num_minibatches != 0) return -1; /* Q10 */
  if (c->env_kind != CRL_ENV_CARTPOLE && c->env_kind != CRL_ENV_PENDULUM) return -1;
  return 0;
}

int orc_create(const crl_config* cfg, orc_ctx** out) {
  if (check_cfg(cfg) || !out) return CRL_ERR_INVALID;
  orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
  c->cfg = *cfg;
  make_layout(cfg->env_kind, &c->L);
  c->N = cfg->num_envs; c->T = cfg->num_steps; c->B = c->N * c->T;
  const layout_t* L = &c->L;
  size_t N = c->N, B = c->B;
  c->params = calloc(L->P, 4); c->grads = calloc(L->P, 4); c->adam_m = calloc(L->P, 4); c->adam_v = calloc(L->P, 4);
  c->beta_pow = calloc(L->n_arrays * 2, 8);
  for (int i = 0; i < L->n_arrays; i++) { c->beta_pow[2 * i] = 0.9; c->beta_pow[2 * i + 1] = 0.999; }
  c->env_state = calloc(N * L->S, 4); c->env_t = calloc(N, 4); c->ep_return = calloc(N, 8);
  c->ep_length = calloc(N, 4); c->reset_count = calloc(N, 4); c->next_obs = calloc(N * L->D, 4);
  c->next_done = calloc(N, 1); c->next_value = calloc(N, 4);
  c->state = calloc(B * L->D, 4); c->action = calloc(B * (L->continuous ? L->A : 1), 4);
  c->logprob = calloc(B, 4); c->reward = calloc(B, 4); c->value = calloc(B, 4);
  c->advantage = calloc(B, 4); c->ret = calloc(B, 4); c->terminal = calloc(B, 1);
  c->ep_done = calloc(B, 1); c->ep_rec_return = calloc(B, 8); c->ep_rec_length = calloc(B, 4);
  c->vnew = calloc(B, 4);
  *out = c;
  return 0;
}
int orc_destroy(orc_ctx* c) {
  if (!c) return 0;
  free(c->params); free(c->grads); free(c->adam_m); free(c->adam_v); free(c->beta_pow);
  free(c->env_state); free(c->env_t); free(c->ep_return); free(c->ep_length); free(c->reset_count);
  free(c->next_obs); free(c->next_done); free(c->next_value); free(c->state); free(c->action);
  free(c->logprob); free(c->reward); free(c->value); free(c->advantage); free(c->ret); free(c->terminal);
  free(c->ep_done); free(c->ep_rec_return); free(c->ep_rec_length); free(c->vnew);
  free(c);
  return 0;
}
int orc_set_params(orc_ctx* c, const float* p, int32_t n) {
  if (n != c->L.P) return CRL_ERR_INVALID;
  memcpy(c->params, p, 4 * (size_t)n);
  return 0;
}
int orc_get_params(orc_ctx* c, float* p, int32_t n) {
  if (n != c->L.P) return CRL_ERR_INVALID;
  memcpy(p, c->params, 4 * (size_t)n);
  return 0;
}
int orc_get_grads(orc_ctx* c, float* p, int32_t n) {
  if (n != c->L.P) return CRL_ERR_INVALID;
  memcpy(p, c->grads, 4 * (size_t)n);
  return 0;
}
int orc_get_adam_state(orc_ctx* c, float* m, float* v, double* bp) {
  memcpy(m, c->adam_m, 4 * (size_t)c->L.P); memcpy(v, c->adam_v, 4 * (size_t)c->L.P);
  memcpy(bp, c->beta_pow, 16 * (size_t)c->L.n_arrays);
  return 0;
}
int orc_set_adam_state(orc_ctx* c, const float* m, const float* v, const double* bp) {
  memcpy(c->adam_m, m, 4 * (size_t)c->L.P); memcpy(c->adam_v, v, 4 * (size_t)c->L.P);
  memcpy(c->beta_pow, bp, 16 * (size_t)c->L.n_arrays);
  return 0;
}

static void reset_one(orc_ctx* c, int n, const float u[4]) {
  if (c->cfg.env_kind == CRL_ENV_CARTPOLE) orc_cartpole_reset(c->env_state + 4 * n, c->env_t + n, u);
  else orc_pendulum_reset(c->env_state + 2 * n, c->env_t + n, u);
}
/* reset!(env) at ppo.jl:112 (every env was just constructed => all reset), then
 * next_obs = state(env), next_done = is_terminated(env) (ppo.jl:114-115) */
int orc_env_reset(orc_ctx* c) {
  for (int n = 0; n < c->N; n++) {
    float u[4];
    orc_rng_reset_uniforms(c->cfg.seed, (uint32_t)(c->cfg.env_id_base + n), c->reset_count[n], u);
    c->reset_count[n] += 1;
    reset_one(c, n, u);
    env_obs(c->cfg.env_kind, c->env_state + c->L.S * n, c->next_obs + c->L.D * n);
    c->next_done[n] = 0; c->ep_return[n] = 0.0; c->ep_length[n] = 0;
  }
  return 0;
}
int orc_env_set_state(orc_ctx* c, const float* state, const int32_t* t) {
  memcpy(c->env_state, state, 4 * (size_t)c->N * c->L.S);
  for (int n = 0; n < c->N; n++) {
    c->env_t[n] = t ? t[n] : 0;
    env_obs(c->cfg.env_kind, c->env_state + c->L.S * n, c->next_obs + c->L.D * n);
    c->next_done[n] = 0;
  }
  return 0;
}

/* ------------------------------------------------------------------ rollout */
typedef struct {
  orc_ctx* c;
  const double* action_noise;
  const float* reset_noise;
} rollout_args;

/* The loop ppo.jl:123-166 for envs [lo,hi). Envs are independent within a rollout (weights
 * are constant), so each env runs all T steps; the stored results equal the reference's
 * step-major order. Quirks reproduced: Q2 (obs copied BEFORE reset!, so the step after a
 * termination sees the stale terminal observation while the action drives the freshly reset
 * env), Q3 (at rollout start obs = state(env) refreshed and all done flags false). */
static void rollout_range(int64_t lo, int64_t hi, int tid, void* argp) {
  (void)tid;
  rollout_args* ra = (rollout_args*)argp;
  orc_ctx* c = ra->c;
  const layout_t* L = &c->L;
  const int N = c->N, T = c->T, D = L->D, A = L->A, S = L->S;
  const int kind = c->cfg.env_kind;
  for (int64_t n = lo; n < hi; n++) {
    float obs[MAXD];
    uint8_t done_flag = 0; /* Q3: ppo.jl:170 is_terminated(env) after reset! => false */
    env_obs(kind, c->env_state + S * n, obs); /* Q3: ppo.jl:169 state(env) refreshed */
    for (int t = 0; t < T; t++) {
      const int64_t b = (int64_t)t * N + n;
      c->ep_length[n] += 1; /* ppo.jl:125 */
      float h1[H], h2[H], z[MAXA], v;
      mlp_forward(c->params + L->actor, D, A, obs, h1, h2, z);
      mlp_forward(c->params + L->critic, D, 1, obs, h1, h2, &v); /* ppo.jl:128 */
      float lp_action;
      int act_i = 0;
      float act_f[MAXA] = {0, 0};
      if (!L->continuous) {
        /* get_action ppo.jl:23-29 + StatsBase.sample(Weights(p)): t = rand()*sum(p);
         * walk cw += p[i] while cw < t. rand() is Float64, p and cw Float32. */
        float p[MAXA], lp[MAXA];
        softmax_logsoftmax(z, A, p, lp);
        double u = ra->action_noise ? ra->action_noise[b]
                                    : orc_rng_action_uniform(c->cfg.seed, (uint32_t)(c->cfg.env_id_base + n), c->policy_step + (uint64_t)t);
        float sum = 0.0f;
        for (int a = 0; a < A; a++) sum += p[a];
        double tt = u * (double)sum;
        int i = 0;
        float cw = p[0];
        while ((double)cw < tt && i < A - 1) { i++; cw += p[i]; }
        act_i = i;
        lp_action = lp[i];
        ((int32_t*)c->action)[b] = act_i;
      } else {
        /* Gaussian head (CleanRL-Python convention): a = mean + exp(logstd)*z,
         * logprob = sum_a [-(a-mean)^2/(2 var) - logstd - log(sqrt(2π))] */
        double zn[2];
        if (ra->action_noise) { for (int a = 0; a < A; a++) zn[a] = ra->action_noise[b * A + a]; }
        else orc_rng_action_normals(c->cfg.seed, (uint32_t)(c->cfg.env_id_base + n), c->policy_step + (uint64_t)t, zn);
        float lps = 0.0f;
        for (int a = 0; a < A; a++) {
          float logstd = c->params[L->logstd + a];
          float sd = expf(logstd);
          float zz = (float)zn[a];
          act_f[a] = z[a] + sd * zz;
          float diff = act_f[a] - z[a];
          lps += -(diff * diff) / (2.0f * sd * sd) - logstd - 0.9189385332046727f;
          ((float*)c->action)[b * A + a] = act_f[a];
        }
        lp_action = lps;
      }
      /* Buffer.add! ppo.jl:133-140 (state = next_obs, terminal = next_done of the previous step) */
      for (int d = 0; d < D; d++) c->state[b * D + d] = obs[d];
      c->terminal[b] = done_flag;
      c->logprob[b] = lp_action;
      c->value[b] = v;
      /* env(action) ppo.jl:130 */
      float r;
      uint8_t dn;
      if (kind == CRL_ENV_CARTPOLE) orc_cartpole_step(c->env_state + S * n, c->env_t + n, act_i, c->cfg.max_episode_steps, &r, &dn);
      else orc_pendulum_step(c->env_state + S * n, c->env_t + n, act_f[0], c->cfg.max_episode_steps, &r, &dn);
      c->reward[b] = r;                            /* ppo.jl:132 */
      env_obs(kind, c->env_state + S * n, obs);    /* ppo.jl:143 (before reset!) */
      done_flag = dn;                              /* ppo.jl:144 */
      c->ep_return[n] += (double)r;                /* ppo.jl:145 */
      c->ep_done[b] = dn;
      if (dn) { /* ppo.jl:147-165 */
        c->ep_rec_return[b] = c->ep_return[n];
        c->ep_rec_length[b] = c->ep_length[n];
        c->ep_return[n] = 0.0;
        c->ep_length[n] = 0;
        float u[4];
        if (ra->reset_noise) memcpy(u, ra->reset_noise + 4 * b, 16);
        else orc_rng_reset_uniforms(c->cfg.seed, (uint32_t)(c->cfg.env_id_base + n), c->reset_count[n], u);
        c->reset_count[n] += 1;
        reset_one(c, (int)n, u); /* reset!(env) resets only terminated envs, multi_thread_env.jl:105-111 */
        /* A2C: obs = deepcopy(state(env)) is read after reset!(env) (a2c.jl:108 then :52), so the next transition
         * starts from the reset state; PPO keeps the stale terminal observation (Q2) */
        if (c->cfg.flags & CRL_FLAG_A2C) env_obs(kind, c->env_state + S * n, obs);
      }
    }
    for (int d = 0; d < D; d++) c->next_obs[D * n + d] = obs[d];
    c->next_done[n] = done_flag;
  }
}
int orc_rollout(orc_ctx* c, const double* action_noise, const float* reset_noise) {
  rollout_args ra = {c, action_noise, reset_noise};
  parallel_for(c->N, rollout_range, &ra);
  c->policy_step += (uint64_t)c->T;
  c->rolled = 1;
  c->gae_done = 0;
  return 0;
}

/* ppo.jl:169-181: next_obs = state(env) (refreshed, post-reset), next_values = critic(next_obs),
 * GAE per env, returns. FIXED mode uses the true done flags of the last step. */
int orc_gae(orc_ctx* c) {
  if (!c->rolled) return CRL_ERR_STATE;
  const layout_t* L = &c->L;
  for (int n = 0; n < c->N; n++) {
    float obs[MAXD], h1[H], h2[H];
    env_obs(c->cfg.env_kind, c->env_state + L->S * n, obs);
    mlp_forward(c->params + L->critic, L->D, 1, obs, h1, h2, c->next_value + n);
  }
  int rc = orc_gae_raw(c->value, c->reward, c->terminal, c->next_value, c->next_done, c->advantage, c->ret,
                       c->T, c->N, c->cfg.gamma, c->cfg.gae_lambda, c->cfg.gae_mode);
  c->gae_done = 1;
  return rc;
}

/* ------------------------------------------------------------------ loss + gradient */
typedef struct {
  int env_kind;
  const float* params;
  const int32_t* idx;
  int M;
  const float *states, *logprobs, *advantages, *returns, *values;
  const void* actions;
  float clip_coef, ent_coeff, v_coef;
  /* phase outputs */
  float* vnew;      /* [M] */
  double adv_mean_f, adv_std_f; /* Float32 results held in double */
  float s_unclipped;
  double inv_cnt_term; /* cnt / M */
  double Mg;           /* normalising minibatch size (= M, or world*M when sharded) */
  int phase;
  /* per-thread partials */
  double* part; /* [threads][P + 8] */
  int P;
} loss_args;

/* clip_value_loss = false (ppo.jl:239-241, CRL_FLAG_NO_VCLIP): set around a loss evaluation; not thread-safe across
 * contexts, which the tests do not need */
static int g_no_vclip = 0;
void orc_set_no_vclip(int on) { g_no_vclip = on ? 1 : 0; }

/* per-thread slot layout after the P gradient doubles */
#define SL_SUM_ADV 0
#define SL_SUM_ADV2 1
#define SL_SUM_S 2
#define SL_CNT 3
#define SL_PG 4
#define SL_VMAX 5
#define SL_ENT 6
#define SL_N 8

static void loss_range(int64_t lo, int64_t hi, int tid, void* argp) {
  loss_args* a = (loss_args*)argp;
  layout_t L;
  make_layout(a->env_kind, &L);
  double* g = a->part + (size_t)tid * (a->P + SL_N);
  double* sl = g + a->P;
  const int D = L.D, A = L.A;
  const float c = a->clip_coef;
  for (int64_t i = lo; i < hi; i++) {
    const int b = a->idx[i];
    const float* x = a->states + (int64_t)b * D;
    if (a->phase == 0) {
      /* critic forward + sums for mean/std (ppo.jl:221) and s = mean(newvalue .- R.^2) (ppo.jl:232, Q5) */
      float h1[H], h2[H], v;
      mlp_forward(a->params + L.critic, D, 1, x, h1, h2, &v);
      a->vnew[i] = v;
      double ad = (double)a->advantages[b];
      sl[SL_SUM_ADV] += ad;
      sl[SL_SUM_ADV2] += ad * ad;
      float R = a->returns[b];
      sl[SL_SUM_S] += (double)(v - R * R);
    } else if (a->phase == 1) {
      /* count of elements where the scalar s wins the max (ppo.jl:236) */
      float R = a->returns[b], V = a->values[b], v = a->vnew[i];
      float dv = v - V;
      float cl = dv < -c ? -c : (dv > c ? c : dv);
      float vc = V + cl;
      float vlc = (vc - R) * (vc - R);
      if (a->s_unclipped > vlc) sl[SL_CNT] += 1.0;
    } else {
      float h1a[H], h2a[H], h1c[H], h2c[H], z[MAXA], v;
      mlp_forward(a->params + L.actor, D, A, x, h1a, h2a, z);
      mlp_forward(a->params + L.critic, D, 1, x, h1c, h2c, &v);
      /* (adv .- mean) ./ (std .+ 1e-8): Float32 numerator, Float64 quotient (Q6) */
      float num = a->advantages[b] - (float)a->adv_mean_f;
      double adv_n = (double)num / ((double)(float)a->adv_std_f + 1e-8);
      float newlp;
      float dz[MAXA];
      double dlogstd[MAXA] = {0, 0};
      double ent_sum = 0.0; /* sum over a of the A×M entropy matrix entries for this sample */
      double g_lp;          /* dL/d newlogprob */
      float p[MAXA], lp[MAXA];
      if (!L.continuous) {
        softmax_logsoftmax(z, A, p, lp);
        int act = ((const int32_t*)a->actions)[b];
        newlp = lp[act];
        for (int k = 0; k < A; k++) ent_sum += (double)(-(p[k] * lp[k])); /* ppo.jl:42 (Q4: matrix, not per-sample sums) */
      } else {
        float s = 0.0f;
        for (int k = 0; k < A; k++) {
          float logstd = a->params[L.logstd + k];
          float sd = expf(logstd);
          float diff = ((const float*)a->actions)[(int64_t)b * A + k] - z[k];
          s += -(diff * diff) / (2.0f * sd * sd) - logstd - 0.9189385332046727f;
          ent_sum += (double)(0.5f + 0.9189385332046727f + logstd);
        }
        newlp = s;
      }
      float logratio = newlp - a->logprobs[b];
      float ratio = expf(logratio);
      float lo_c = 1.0f - c, hi_c = 1.0f + c;
      float rc = ratio < lo_c ? lo_c : (ratio > hi_c ? hi_c : ratio);
      double pg1 = -adv_n * (double)ratio; /* ppo.jl:226 */
      double pg2 = -adv_n * (double)rc;    /* ppo.jl:227 */
      double pgm;
      double dratio; /* d max / d ratio */
      if (pg1 > pg2) { pgm = pg1; dratio = -adv_n; }
      else { pgm = pg2; dratio = (ratio >= lo_c && ratio <= hi_c) ? -adv_n : 0.0; }
      sl[SL_PG] += pgm;
      g_lp = dratio * (double)ratio / a->Mg;
      /* value loss, Q5 */
      float R = a->returns[b], V = a->values[b];
      float dvv = v - V;
      float cl = dvv < -c ? -c : (dvv > c ? c : dvv);
      float vc = V + cl;
      float vlc = (vc - R) * (vc - R);
      float vmax = a->s_unclipped > vlc ? a->s_unclipped : vlc;
      double dv_d = a->inv_cnt_term; /* (1/M) * cnt: every v_new_j feeds s */
      if (!(a->s_unclipped > vlc) && dvv >= -c && dvv <= c) dv_d += 2.0 * (double)(vc - R);
      if (g_no_vclip) {              /* clip_value_loss = false: 0.5 * mean((newvalue - R).^2), ppo.jl:239-241 */
        float d = v - R;
        vmax = d * d;
        dv_d = 2.0 * (double)d;
      }
      sl[SL_VMAX] += (double)vmax;
      float dv = (float)((double)a->v_coef * 0.5 / a->Mg * dv_d);
      sl[SL_ENT] += ent_sum;
      /* back through the heads */
      const double ent_scale = (double)a->ent_coeff / ((double)A * a->Mg);
      if (!L.continuous) {
        int act = ((const int32_t*)a->actions)[b];
        double Hs = ent_sum;
        for (int k = 0; k < A; k++) {
          double d = g_lp * ((k == act ? 1.0 : 0.0) - (double)p[k]);
          d += ent_scale * (double)p[k] * ((double)lp[k] + Hs); /* -ent_coeff * d(mean entropy)/dz */
          dz[k] = (float)d;
        }
      } else {
        for (int k = 0; k < A; k++) {
          float logstd = a->params[L.logstd + k];
          float sd = expf(logstd);
          double diff = (double)(((const float*)a->actions)[(int64_t)b * A + k] - z[k]);
          double var = (double)sd * (double)sd;
          dz[k] = (float)(g_lp * diff / var);
          dlogstd[k] = g_lp * (diff * diff / var - 1.0) - ent_scale;
        }
      }
      mlp_backward(a->params + L.actor, g + L.actor, D, A, x, h1a, h2a, dz);
      mlp_backward(a->params + L.critic, g + L.critic, D, 1, x, h1c, h2c, &dv);
      if (L.continuous) for (int k = 0; k < A; k++) g[L.logstd + k] += dlogstd[k];
    }
  }
}

static void reduce_parts(loss_args* a, int nt, double* out) {
  int W = a->P + SL_N;
  for (int k = 0; k < W; k++) {
    double s = 0.0;
    for (int t = 0; t < nt; t++) s += a->part[(size_t)t * W + k];
    out[k] = s;
  }
}

/* The closure at ppo.jl:202-244 and its gradient for one minibatch.
 * stats_out = {loss, pg_loss, v_loss, entropy_loss}; grads_out = un-clipped gradient.
 * vnew_out may be NULL. */
int orc_ppo_loss_raw(int32_t env_kind, const float* params, const int32_t* idx, int32_t M, const float* states,
                     const void* actions, const float* logprobs, const float* advantages, const float* returns,
                     const float* values, float clip_coef, float ent_coeff, float v_coef, float* grads_out,
                     double* stats_out, float* vnew_out) {
  layout_t L;
  if (make_layout(env_kind, &L) || M < 2) return CRL_ERR_INVALID;
  int nt = g_threads > 256 ? 256 : g_threads;
  if (nt > M) nt = M;
  int W = L.P + SL_N;
  loss_args a;
  memset(&a, 0, sizeof(a));
  a.env_kind = env_kind; a.params = params; a.idx = idx; a.M = M; a.states = states; a.actions = actions;
  a.logprobs = logprobs; a.advantages = advantages; a.returns = returns; a.values = values;
  a.clip_coef = clip_coef; a.ent_coeff = ent_coeff; a.v_coef = v_coef; a.P = L.P; a.Mg = (double)M;
  a.vnew = vnew_out ? vnew_out : (float*)malloc(4 * (size_t)M);
  a.part = (double*)calloc((size_t)nt * W, 8);
  double* tot = (double*)calloc(W, 8);
  int save = g_threads;
  g_threads = nt;
  a.phase = 0;
  parallel_for(M, loss_range, &a);
  reduce_parts(&a, nt, tot);
  double mean = tot[L.P + SL_SUM_ADV] / M;
  double var = (tot[L.P + SL_SUM_ADV2] - M * mean * mean) / (M - 1); /* corrected std, ppo.jl:221 */
  if (var < 0) var = 0;
  a.adv_mean_f = (double)(float)mean;
  a.adv_std_f = (double)(float)sqrt(var);
  a.s_unclipped = (float)(tot[L.P + SL_SUM_S] / M);
  memset(a.part, 0, (size_t)nt * W * 8);
  a.phase = 1;
  parallel_for(M, loss_range, &a);
  reduce_parts(&a, nt, tot);
  a.inv_cnt_term = tot[L.P + SL_CNT] / M;
  memset(a.part, 0, (size_t)nt * W * 8);
  a.phase = 2;
  parallel_for(M, loss_range, &a);
  reduce_parts(&a, nt, tot);
  g_threads = save;
  for (int k = 0; k < L.P; k++) grads_out[k] = (float)tot[k];
  double pg_loss = tot[L.P + SL_PG] / M;                           /* ppo.jl:228 */
  double v_loss = 0.5 * (double)(float)(tot[L.P + SL_VMAX] / M);   /* ppo.jl:237 */
  double ent_loss = (double)(float)(tot[L.P + SL_ENT] / ((double)L.A * M)); /* ppo.jl:242 */
  stats_out[1] = pg_loss;
  stats_out[2] = v_loss;
  stats_out[3] = ent_loss;
  stats_out[0] = pg_loss - (double)(ent_coeff * (float)ent_loss) + (double)v_coef * v_loss; /* ppo.jl:243 */
  if (!vnew_out) free(a.vnew);
  free(a.part);
  free(tot);
  return 0;
}

/* One phase of the same closure on ONE SHARD of a minibatch that is split over several ranks
 * (SURVEY §8e): the three minibatch-global scalars are exchanged between phases by the caller
 * (tests/test_dist_gloo.py does it with gloo; the CUDA library with NCCL).
 *   phase 0: out io[0..2] = local Σadv, Σadv², Σ(v_new - R²); fills vnew[M]
 *   phase 1: in io[3] = s (global); out io[4] = local #{s > (clip-R)²}
 *   phase 2: in io[3] = s, io[4] = global count, io[5] = adv mean, io[6] = adv std, io[7] = global M;
 *            out grads[P] (local sums, double), io[0..2] = local Σpg, Σvmax, Σentropy */
int orc_ppo_loss_phase(int32_t env_kind, const float* params, const int32_t* idx, int32_t M, const float* states,
                       const void* actions, const float* logprobs, const float* advantages, const float* returns,
                       const float* values, float clip_coef, float ent_coeff, float v_coef, int32_t phase, double* io,
                       float* vnew, double* grads) {
  layout_t L;
  if (make_layout(env_kind, &L) || M < 1 || phase < 0 || phase > 2) return CRL_ERR_INVALID;
  int W = L.P + SL_N;
  loss_args a;
  memset(&a, 0, sizeof(a));
  a.env_kind = env_kind; a.params = params; a.idx = idx; a.M = M; a.states = states; a.actions = actions;
  a.logprobs = logprobs; a.advantages = advantages; a.returns = returns; a.values = values;
  a.clip_coef = clip_coef; a.ent_coeff = ent_coeff; a.v_coef = v_coef; a.P = L.P; a.vnew = vnew; a.phase = phase;
  a.part = (double*)calloc((size_t)W, 8);
  a.s_unclipped = (float)io[3];
  a.adv_mean_f = io[5]; a.adv_std_f = io[6]; a.Mg = io[7];
  a.inv_cnt_term = phase == 2 ? io[4] / io[7] : 0.0;
  int save = g_threads;
  g_threads = 1;
  parallel_for(M, loss_range, &a);
  g_threads = save;
  if (phase == 0) { io[0] = a.part[L.P + SL_SUM_ADV]; io[1] = a.part[L.P + SL_SUM_ADV2]; io[2] = a.part[L.P + SL_SUM_S]; }
  else if (phase == 1) io[4] = a.part[L.P + SL_CNT];
  else {
    for (int k = 0; k < L.P; k++) grads[k] = a.part[k];
    io[0] = a.part[L.P + SL_PG]; io[1] = a.part[L.P + SL_VMAX]; io[2] = a.part[L.P + SL_ENT];
  }
  free(a.part);
  return 0;
}

/* A2C losses and their gradient for one batch (a2c.jl:78-97): critic_loss = mean((R - v)^2) with the gradient flowing
 * through v; actor_loss = -mean(logp(a) .* advantage) with advantage = R - v held constant. The reference applies two
 * update! calls (critic, then actor) with one optimiser; the parameter sets are disjoint and ClipNorm/Adam are per array,
 * so one combined step is identical. stats_out = {actor+critic, actor_loss, critic_loss, 0}. */
int orc_a2c_loss_raw(int32_t env_kind, const float* params, const int32_t* idx, int32_t M, const float* states,
                     const void* actions, const float* returns, float* grads_out, double* stats_out) {
  layout_t L;
  if (make_layout(env_kind, &L) || M < 1) return CRL_ERR_INVALID;
  double* g = (double*)calloc(L.P, 8);
  double actor = 0.0, critic = 0.0;
  for (int i = 0; i < M; i++) {
    const int b = idx[i];
    const float* x = states + (int64_t)b * L.D;
    float h1a[H], h2a[H], h1c[H], h2c[H], z[MAXA], v, p[MAXA], lp[MAXA], dz[MAXA];
    mlp_forward(params + L.actor, L.D, L.A, x, h1a, h2a, z);
    mlp_forward(params + L.critic, L.D, 1, x, h1c, h2c, &v);
    const double adv = (double)returns[b] - (double)v; /* a2c.jl:83 */
    critic += adv * adv;                               /* a2c.jl:84 */
    const float dv = (float)(-2.0 * adv / (double)M);
    const double g_lp = -adv / (double)M;              /* d(-mean(logp .* adv))/dlogp, a2c.jl:95 */
    double newlp;
    if (!L.continuous) {
      softmax_logsoftmax(z, L.A, p, lp);
      const int act = ((const int32_t*)actions)[b];
      newlp = (double)lp[act];
      for (int k = 0; k < L.A; k++) dz[k] = (float)(g_lp * ((k == act ? 1.0 : 0.0) - (double)p[k]));
    } else {
      float s = 0.0f;
      for (int k = 0; k < L.A; k++) {
        const float logstd = params[L.logstd + k], sd = expf(logstd);
        const float diff = ((const float*)actions)[(int64_t)b * L.A + k] - z[k];
        s += -(diff * diff) / (2.0f * sd * sd) - logstd - 0.9189385332046727f;
        const double var = (double)sd * (double)sd;
        dz[k] = (float)(g_lp * (double)diff / var);
        g[L.logstd + k] += g_lp * ((double)diff * (double)diff / var - 1.0);
      }
      newlp = (double)s;
    }
    actor += -newlp * adv;
    mlp_backward(params + L.actor, g + L.actor, L.D, L.A, x, h1a, h2a, dz);
    mlp_backward(params + L.critic, g + L.critic, L.D, 1, x, h1c, h2c, &dv);
  }
  for (int k = 0; k < L.P; k++) grads_out[k] = (float)g[k];
  stats_out[1] = actor / M;
  stats_out[2] = critic / M;
  stats_out[3] = 0.0;
  stats_out[0] = stats_out[1] + stats_out[2];
  free(g);
  return 0;
}

/* Flux.Optimise.update!(Optimiser(ClipNorm(thresh), Adam(η)), params, gs), ppo.jl:93,250
 * [Flux 0.13.4]: per ARRAY: if norm(Δ) > thresh: Δ *= thresh/norm(Δ); then Adam with
 * Float64 scalars (β=(0.9,0.999), ε=1e-8) on Float32 state, per-array β powers. */
int orc_clip_adam_raw(int32_t env_kind, float* params, const float* grads, float* m, float* v, double* beta_pow,
                      double lr, float clip_norm) {
  layout_t L;
  if (make_layout(env_kind, &L)) return CRL_ERR_INVALID;
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  for (int i = 0; i < L.n_arrays; i++) {
    const int o = L.off[i], n = L.size[i];
    double ss = 0.0;
    for (int k = 0; k < n; k++) ss += (double)grads[o + k] * (double)grads[o + k];
    float nrm = (float)sqrt(ss); /* norm(Δ::Array{Float32}) is Float32 */
    double scale = 1.0;
    int clip = (double)nrm > (double)clip_norm;
    if (clip) scale = (double)clip_norm / (double)nrm;
    double bp1 = beta_pow[2 * i], bp2 = beta_pow[2 * i + 1];
    for (int k = 0; k < n; k++) {
      float d = grads[o + k];
      if (clip) d = (float)((double)d * scale); /* rmul!(Δ, thresh/nrm) */
      float mt = (float)(b1 * (double)m[o + k] + (1.0 - b1) * (double)d);
      float vt = (float)(b2 * (double)v[o + k] + (1.0 - b2) * (double)d * (double)d);
      m[o + k] = mt;
      v[o + k] = vt;
      float step = (float)((double)mt / (1.0 - bp1) / (sqrt((double)vt / (1.0 - bp2)) + eps) * lr);
      params[o + k] = params[o + k] - step;
    }
    beta_pow[2 * i] = bp1 * b1;
    beta_pow[2 * i + 1] = bp2 * b2;
  }
  return 0;
}

int orc_update_minibatch(orc_ctx* c, const int32_t* idx, int32_t M, double lr, crl_loss_stats* stats) {
  if (!c->gae_done) return CRL_ERR_STATE;
  double st[4];
  int rc;
  if (c->cfg.flags & CRL_FLAG_A2C)
    rc = orc_a2c_loss_raw(c->cfg.env_kind, c->params, idx, M, c->state, c->action, c->ret, c->grads, st);
  else {
    g_no_vclip = (c->cfg.flags & CRL_FLAG_NO_VCLIP) ? 1 : 0;
    rc = orc_ppo_loss_raw(c->cfg.env_kind, c->params, idx, M, c->state, c->action, c->logprob, c->advantage,
                          c->ret, c->value, c->cfg.clip_coef, c->cfg.ent_coeff, c->cfg.v_coef, c->grads, st, c->vnew);
    g_no_vclip = 0;
  }
  if (rc) return rc;
  if (stats) { stats->loss = st[0]; stats->pg_loss = st[1]; stats->v_loss = st[2]; stats->entropy_loss = st[3]; }
  return orc_clip_adam_raw(c->cfg.env_kind, c->params, c->grads, c->adam_m, c->adam_v, c->beta_pow, lr, c->cfg.clip_norm);
}

/* ------------------------------------------------------------------ device permutation */
/* Philox-keyed Feistel bijection on [0,B) with cycle walking (SURVEY §7.2 "Permutation at
 * scale"); replaces shuffle(b_inds), ppo.jl:194, when the host passes no permutation. */
static inline uint32_t feistel_round(uint32_t x, uint32_t k, uint32_t mask) {
  x ^= k;
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  return x & mask;
}
void orc_perm_keys(uint64_t seed, uint64_t update_index, uint32_t epoch, uint32_t rank, uint32_t keys[8]) {
  philox_draw(seed, epoch | (rank << 16), update_index, ORC_STREAM_PERM, keys);
  philox_draw(seed, epoch | (rank << 16) | 0x80000000u, update_index, ORC_STREAM_PERM, keys + 4);
}
uint32_t orc_perm_index(uint32_t i, uint32_t B, const uint32_t keys[8]) {
  int bits = 1;
  while ((1u << bits) < B) bits++;
  int hb = (bits + 1) / 2;
  uint32_t mask = (1u << hb) - 1u;
  uint32_t x = i;
  do {
    uint32_t l = x >> hb, r = x & mask;
    for (int k = 0; k < 6; k++) {
      uint32_t nl = r;
      r = l ^ feistel_round(r, keys[k], mask);
      l = nl;
    }
    x = (l << hb) | r;
  } while (x >= B);
  return x;
}
int orc_device_permutation(orc_ctx* c, int64_t update_index, int32_t epoch, int32_t* out) {
  uint32_t keys[8];
  orc_perm_keys(c->cfg.seed, (uint64_t)update_index, (uint32_t)epoch, (uint32_t)c->cfg.rank, keys);
  for (int i = 0; i < c->B; i++) out[i] = (int32_t)orc_perm_index((uint32_t)i, (uint32_t)c->B, keys);
  return 0;
}

/* epochs × minibatches, ppo.jl:193-251. perms = [epochs][B] or NULL (device permutation). */
int orc_update_epochs(orc_ctx* c, const int32_t* perms, double lr, crl_loss_stats* stats) {
  if (!c->gae_done) return CRL_ERR_STATE;
  const int B = c->B, M = B / c->cfg.num_minibatches;
  int32_t* perm = (int32_t*)malloc(4 * (size_t)B);
  int k = 0;
  for (int e = 0; e < c->cfg.update_epochs; e++) {
    if (perms) memcpy(perm, perms + (size_t)e * B, 4 * (size_t)B);
    else orc_device_permutation(c, (int64_t)c->update_index, e, perm);
    for (int start = 0; start < B; start += M) { /* ppo.jl:197,203-204 */
      int rc = orc_update_minibatch(c, perm + start, M, lr, stats ? stats + k : NULL);
      if (rc) { free(perm); return rc; }
      k++;
    }
  }
  free(perm);
  c->update_index += 1;
  return 0;
}

/* one whole update with Philox draws (mirror of crl_train_update) */
int orc_train_update(orc_ctx* c, double lr, crl_loss_stats* stats) {
  int rc = orc_rollout(c, NULL, NULL);
  if (rc) return rc;
  rc = orc_gae(c);
  if (rc) return rc;
  return orc_update_epochs(c, NULL, lr, stats);
}

/* ------------------------------------------------------------------ data access */
static int field_info(orc_ctx* c, int f, void** p, size_t* bytes) {
  const layout_t* L = &c->L;
  size_t N = c->N, B = c->B;
  switch (f) {
    case CRL_F_STATE: *p = c->state; *bytes = B * L->D * 4; break;
    case CRL_F_ACTION: *p = c->action; *bytes = B * (L->continuous ? L->A : 1) * 4; break;
    case CRL_F_LOGPROB: *p = c->logprob; *bytes = B * 4; break;
    case CRL_F_REWARD: *p = c->reward; *bytes = B * 4; break;
    case CRL_F_TERMINAL: *p = c->terminal; *bytes = B; break;
    case CRL_F_VALUE: *p = c->value; *bytes = B * 4; break;
    case CRL_F_ADVANTAGE: *p = c->advantage; *bytes = B * 4; break;
    case CRL_F_RETURN: *p = c->ret; *bytes = B * 4; break;
    case CRL_F_NEXT_OBS: *p = c->next_obs; *bytes = N * L->D * 4; break;
    case CRL_F_NEXT_DONE: *p = c->next_done; *bytes = N; break;
    case CRL_F_NEXT_VALUE: *p = c->next_value; *bytes = N * 4; break;
    case CRL_F_ENV_STATE: *p = c->env_state; *bytes = N * L->S * 4; break;
    case CRL_F_ENV_T: *p = c->env_t; *bytes = N * 4; break;
    case CRL_F_EP_RETURN: *p = c->ep_return; *bytes = N * 8; break;
    case CRL_F_EP_LENGTH: *p = c->ep_length; *bytes = N * 4; break;
    case CRL_F_RESET_COUNT: *p = c->reset_count; *bytes = N * 4; break;
    case CRL_F_VNEW: *p = c->vnew; *bytes = (B / c->cfg.num_minibatches) * 4; break;
    default: return -1;
  }
  return 0;
}
int orc_read_field(orc_ctx* c, int32_t field, void* host, size_t bytes) {
  void* p; size_t nb;
  if (field_info(c, field, &p, &nb) || bytes != nb) return CRL_ERR_INVALID;
  memcpy(host, p, nb);
  return 0;
}
int orc_write_field(orc_ctx* c, int32_t field, const void* host, size_t bytes) {
  void* p; size_t nb;
  if (field_info(c, field, &p, &nb) || bytes != nb) return CRL_ERR_INVALID;
  memcpy(p, host, nb);
  if (field <= CRL_F_VALUE) c->rolled = 1;
  if (field == CRL_F_ADVANTAGE || field == CRL_F_RETURN) c->gae_done = 1;
  return 0;
}
/* episode records of the last rollout in the reference's logging order (step, then env; ppo.jl:149) */
int orc_pop_episodes(orc_ctx* c, crl_episode* out, int32_t max_records, int32_t* n_out, crl_episode_agg* agg) {
  int k = 0;
  crl_episode_agg g = {0, 0.0, 0.0, -INFINITY, 0};
  for (int t = 0; t < c->T; t++)
    for (int n = 0; n < c->N; n++) {
      int64_t b = (int64_t)t * c->N + n;
      if (!c->ep_done[b]) continue;
      g.count++; g.sum_return += c->ep_rec_return[b]; g.sum_length += c->ep_rec_length[b];
      if (c->ep_rec_return[b] > g.max_return) g.max_return = c->ep_rec_return[b];
      if (out && k < max_records) {
        out[k].step = t; out[k].env = n; out[k].length = c->ep_rec_length[b]; out[k]._pad = 0;
        out[k].episode_return = c->ep_rec_return[b];
        k++;
      } else if (out) g.dropped++;
    }
  if (n_out) *n_out = k;
  if (agg) *agg = g;
  return 0;
}
uint64_t orc_policy_step(orc_ctx* c) { return c->policy_step; }

/* ================================================================== DQN (dqn.jl) =====
 * CPU restatement of the vectorised DQN contract in include/cleanrl_cuda.h. TEST INFRASTRUCTURE like the rest of
 * this file. Parity unpinned: Flux Dense/relu/mse/Adam and RLEnvs CartPole are restated from the pinned versions,
 * the reference has no tests or golden vectors. What is pinned: linear_schedule (dqn.jl:28-31) by hand, the
 * backward pass against a float64 NumPy restatement (tests/test_dqn.py). */
#define DQ_H1 120
#define DQ_H2 84
#define DQ_A 2
#define DQ_D 4
#define DQ_W1 0
#define DQ_B1 (DQ_W1 + DQ_H1 * DQ_D)
#define DQ_W2 (DQ_B1 + DQ_H1)
#define DQ_B2 (DQ_W2 + DQ_H2 * DQ_H1)
#define DQ_W3 (DQ_B2 + DQ_H2)
#define DQ_B3 (DQ_W3 + DQ_A * DQ_H2)
#define DQ_P (DQ_B3 + DQ_A)
#define ORC_STREAM_DQN_ACT 3u
#define ORC_STREAM_DQN_BATCH 4u

typedef struct orc_dqn_ctx {
  crl_dqn_config cfg;
  float q[DQ_P], tgt[DQ_P], m[DQ_P], v[DQ_P], g[DQ_P];
  double bp1, bp2;
  float* env_state; int32_t* env_t; double* ep_ret; int32_t* ep_len; uint32_t* resets;
  float *b_state, *b_next, *b_reward; int32_t* b_action; uint8_t* b_term;
  int32_t size, ptr;
  int64_t it, learn_steps;
  double last_loss;
  /* data-parallel shard (orc_dqn_set_shard): world ranks, this one owns the envs env_id_base .. env_id_base + num_envs - 1 */
  int32_t world, rank, env_id_base;
} orc_dqn_ctx;

double orc_dqn_linear_schedule(double start_e, double end_e, double duration, double t) {
  double slope = (end_e - start_e) / duration;   /* dqn.jl:29 */
  double e = slope * t + start_e;
  return e > end_e ? e : end_e;                  /* dqn.jl:30 */
}
/* Chain(Dense(4,120,relu), Dense(120,84,relu), Dense(84,2)), dqn.jl:25; W (out,in) column-major */
static void dqn_forward(const float* p, const float* x, float* h1, float* h2, float* q) {
  for (int j = 0; j < DQ_H1; j++) {
    float acc = 0.0f;
    for (int k = 0; k < DQ_D; k++) acc += p[DQ_W1 + j + DQ_H1 * k] * x[k];
    acc += p[DQ_B1 + j];
    h1[j] = acc > 0.0f ? acc : 0.0f;
  }
  for (int j = 0; j < DQ_H2; j++) {
    float acc = 0.0f;
    for (int k = 0; k < DQ_H1; k++) acc += p[DQ_W2 + j + DQ_H2 * k] * h1[k];
    acc += p[DQ_B2 + j];
    h2[j] = acc > 0.0f ? acc : 0.0f;
  }
  for (int o = 0; o < DQ_A; o++) {
    float acc = 0.0f;
    for (int k = 0; k < DQ_H2; k++) acc += p[DQ_W3 + o + DQ_A * k] * h2[k];
    q[o] = acc + p[DQ_B3 + o];
  }
}
int orc_dqn_forward_raw(const float* params, const float* obs, float* q_out, int64_t n) {
  float h1[DQ_H1], h2[DQ_H2];
  for (int64_t i = 0; i < n; i++) dqn_forward(params, obs + i * DQ_D, h1, h2, q_out + i * DQ_A);
  return 0;
}
/* loss and gradient of one batch (dqn.jl:96-108); everything passed in, nothing sampled */
/* B_scale = size of the GLOBAL batch the mean is taken over (= B unless the batch is one shard of a data-parallel
 * step); sq_sum_out receives the plain sum of squared TD errors of these B samples */
static int dqn_loss_grad(const float* q_params, const float* tgt_params, int32_t B, int32_t B_scale, const float* state,
                         const int32_t* action, const float* reward, const float* next_state, const uint8_t* terminal,
                         double gamma, float* grads, double* sq_sum_out) {
  float* h1 = (float*)malloc(sizeof(float) * (size_t)B * DQ_H1);
  float* h2 = (float*)malloc(sizeof(float) * (size_t)B * DQ_H2);
  float* dq = (float*)calloc((size_t)B * DQ_A, sizeof(float));
  float* dz2 = (float*)malloc(sizeof(float) * (size_t)B * DQ_H2);
  float* dz1 = (float*)malloc(sizeof(float) * (size_t)B * DQ_H1);
  double loss = 0.0;
  for (int i = 0; i < B; i++) {
    float t1[DQ_H1], t2[DQ_H2], qn[DQ_A], qv[DQ_A];
    dqn_forward(tgt_params, next_state + i * DQ_D, t1, t2, qn);
    float next_q = qn[0] > qn[1] ? qn[0] : qn[1];                                       /* maximum, dqn.jl:99 */
    double td = (double)reward[i] + gamma * (double)next_q * (1.0 - (double)terminal[i]); /* dqn.jl:100 */
    dqn_forward(q_params, state + i * DQ_D, h1 + i * DQ_H1, h2 + i * DQ_H2, qv);
    double diff = td - (double)qv[action[i]];
    loss += diff * diff;                                                                 /* Flux.mse, dqn.jl:107 */
    dq[i * DQ_A + action[i]] = (float)(-2.0 * diff / (double)B_scale);
  }
  *sq_sum_out = loss;
  memset(grads, 0, sizeof(float) * DQ_P);
  /* backward, reductions over the batch in ascending sample order */
  for (int o = 0; o < DQ_A; o++) {
    float bs = 0.0f;
    for (int i = 0; i < B; i++) bs += dq[i * DQ_A + o];
    grads[DQ_B3 + o] = bs;
    for (int k = 0; k < DQ_H2; k++) {
      float acc = 0.0f;
      for (int i = 0; i < B; i++) acc += dq[i * DQ_A + o] * h2[i * DQ_H2 + k];
      grads[DQ_W3 + o + DQ_A * k] = acc;
    }
  }
  for (int i = 0; i < B; i++)
    for (int k = 0; k < DQ_H2; k++) {
      float dh = 0.0f;
      for (int o = 0; o < DQ_A; o++) dh += q_params[DQ_W3 + o + DQ_A * k] * dq[i * DQ_A + o];
      dz2[i * DQ_H2 + k] = h2[i * DQ_H2 + k] > 0.0f ? dh : 0.0f;
    }
  for (int j = 0; j < DQ_H2; j++) {
    float bs = 0.0f;
    for (int i = 0; i < B; i++) bs += dz2[i * DQ_H2 + j];
    grads[DQ_B2 + j] = bs;
    for (int k = 0; k < DQ_H1; k++) {
      float acc = 0.0f;
      for (int i = 0; i < B; i++) acc += dz2[i * DQ_H2 + j] * h1[i * DQ_H1 + k];
      grads[DQ_W2 + j + DQ_H2 * k] = acc;
    }
  }
  for (int i = 0; i < B; i++)
    for (int k = 0; k < DQ_H1; k++) {
      float dh = 0.0f;
      for (int j = 0; j < DQ_H2; j++) dh += q_params[DQ_W2 + j + DQ_H2 * k] * dz2[i * DQ_H2 + j];
      dz1[i * DQ_H1 + k] = h1[i * DQ_H1 + k] > 0.0f ? dh : 0.0f;
    }
  for (int j = 0; j < DQ_H1; j++) {
    float bs = 0.0f;
    for (int i = 0; i < B; i++) bs += dz1[i * DQ_H1 + j];
    grads[DQ_B1 + j] = bs;
    for (int k = 0; k < DQ_D; k++) {
      float acc = 0.0f;
      for (int i = 0; i < B; i++) acc += dz1[i * DQ_H1 + j] * state[i * DQ_D + k];
      grads[DQ_W1 + j + DQ_H1 * k] = acc;
    }
  }
  free(h1); free(h2); free(dq); free(dz2); free(dz1);
  return 0;
}
int orc_dqn_loss_raw(const float* q_params, const float* tgt_params, int32_t B, const float* state, const int32_t* action,
                     const float* reward, const float* next_state, const uint8_t* terminal, double gamma, float* grads,
                     double* loss_out) {
  double sq = 0.0;
  int rc = dqn_loss_grad(q_params, tgt_params, B, B, state, action, reward, next_state, terminal, gamma, grads, &sq);
  *loss_out = sq / (double)B;
  return rc;
}
/* Flux.Adam(eta) [Flux 0.13.4], one shared (beta1^t, beta2^t) pair: every array is updated at every step */
static void dqn_adam(orc_dqn_ctx* c) {
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8, lr = c->cfg.lr;
  for (int k = 0; k < DQ_P; k++) {
    float d = c->g[k];
    float mt = (float)(b1 * (double)c->m[k] + (1.0 - b1) * (double)d);
    float vt = (float)(b2 * (double)c->v[k] + (1.0 - b2) * (double)d * (double)d);
    c->m[k] = mt; c->v[k] = vt;
    double den = sqrt((double)vt / (1.0 - c->bp2)) + eps;
    float step = (float)((double)mt / (1.0 - c->bp1) / den * lr);
    c->q[k] = c->q[k] - step;
  }
  c->bp1 *= b1; c->bp2 *= b2;
}
int orc_dqn_create(const crl_dqn_config* cfg, orc_dqn_ctx** out) {
  if (!cfg || !out || cfg->struct_size != (int32_t)sizeof(crl_dqn_config)) return -1;
  if (cfg->num_envs < 1 || cfg->buffer_size < cfg->num_envs || cfg->batch_size < 1 || cfg->batch_size > 128 ||
      cfg->train_freq < 1 || cfg->target_net_freq < 1) return -1;
  orc_dqn_ctx* c = (orc_dqn_ctx*)calloc(1, sizeof(orc_dqn_ctx));
  c->cfg = *cfg;
  int N = cfg->num_envs, C = cfg->buffer_size;
  c->env_state = (float*)calloc((size_t)N * 4, sizeof(float)); c->env_t = (int32_t*)calloc(N, sizeof(int32_t));
  c->ep_ret = (double*)calloc(N, sizeof(double)); c->ep_len = (int32_t*)calloc(N, sizeof(int32_t));
  c->resets = (uint32_t*)calloc(N, sizeof(uint32_t));
  c->b_state = (float*)calloc((size_t)C * 4, sizeof(float)); c->b_next = (float*)calloc((size_t)C * 4, sizeof(float));
  c->b_reward = (float*)calloc(C, sizeof(float)); c->b_action = (int32_t*)calloc(C, sizeof(int32_t));
  c->b_term = (uint8_t*)calloc(C, 1);
  c->bp1 = 0.9; c->bp2 = 0.999;
  c->world = 1; c->rank = 0; c->env_id_base = 0;
  *out = c;
  return 0;
}
int orc_dqn_destroy(orc_dqn_ctx* c) {
  if (!c) return -1;
  free(c->env_state); free(c->env_t); free(c->ep_ret); free(c->ep_len); free(c->resets);
  free(c->b_state); free(c->b_next); free(c->b_reward); free(c->b_action); free(c->b_term); free(c);
  return 0;
}
int orc_dqn_set_params(orc_dqn_ctx* c, const float* p, int32_t n) {
  if (!c || n != DQ_P) return -1;
  memcpy(c->q, p, sizeof(float) * DQ_P); memcpy(c->tgt, p, sizeof(float) * DQ_P);   /* deepcopy, dqn.jl:40 */
  memset(c->m, 0, sizeof(c->m)); memset(c->v, 0, sizeof(c->v));
  c->bp1 = 0.9; c->bp2 = 0.999;
  return 0;
}
int orc_dqn_get_params(orc_dqn_ctx* c, float* q, float* tgt, int32_t n) {
  if (!c || n != DQ_P) return -1;
  if (q) memcpy(q, c->q, sizeof(float) * DQ_P);
  if (tgt) memcpy(tgt, c->tgt, sizeof(float) * DQ_P);
  return 0;
}
int orc_dqn_reset(orc_dqn_ctx* c) {
  if (!c) return -1;
  for (int n = 0; n < c->cfg.num_envs; n++) {
    float u[4];
    c->resets[n] = 0;
    orc_rng_reset_uniforms(c->cfg.seed, (uint32_t)(c->env_id_base + n), c->resets[n], u);
    c->resets[n] += 1;
    orc_cartpole_reset(c->env_state + 4 * n, &c->env_t[n], u);
    c->ep_ret[n] = 0.0; c->ep_len[n] = 0;
  }
  c->size = 0; c->ptr = 0; c->it = 0; c->learn_steps = 0; c->last_loss = 0.0;
  return 0;
}
/* this shard's part of a learning step: gradient of its batch_size samples scaled by 1 / (world * batch_size) into
 * c->g, and the sum of its squared TD errors */
static double dqn_learn_grad(orc_dqn_ctx* c) {
  const int B = c->cfg.batch_size;
  uint32_t keys[8];
  philox_draw(c->cfg.seed, (uint32_t)c->rank, (uint64_t)c->learn_steps, ORC_STREAM_DQN_BATCH, keys);
  philox_draw(c->cfg.seed, 0x80000000u | (uint32_t)c->rank, (uint64_t)c->learn_steps, ORC_STREAM_DQN_BATCH, keys + 4);
  float st[128 * 4], nx[128 * 4], rw[128];
  int32_t ac[128];
  uint8_t tm[128];
  for (int i = 0; i < B; i++) {
    uint32_t idx = orc_perm_index((uint32_t)i, (uint32_t)c->size, keys);   /* sample(1:size, B, replace=false) */
    memcpy(st + 4 * i, c->b_state + 4 * idx, 16); memcpy(nx + 4 * i, c->b_next + 4 * idx, 16);
    rw[i] = c->b_reward[idx]; ac[i] = c->b_action[idx]; tm[i] = c->b_term[idx];
  }
  double sq = 0.0;
  dqn_loss_grad(c->q, c->tgt, B, B * c->world, st, ac, rw, nx, tm, c->cfg.gamma, c->g, &sq);
  return sq;
}
/* one vector step of this shard's envs (dqn.jl:49-92); gs = global step after it */
static void dqn_act_iteration(orc_dqn_ctx* c, double eps, double* sum_ret, double* sum_len, int64_t* episodes) {
  const int N = c->cfg.num_envs, C = c->cfg.buffer_size;
  for (int n = 0; n < N; n++) {
    const uint32_t gid = (uint32_t)(c->env_id_base + n);
    float obs[4], h1[DQ_H1], h2[DQ_H2], q[DQ_A];
    memcpy(obs, c->env_state + 4 * n, 16);                                    /* deepcopy(state(env)), dqn.jl:50 */
    uint32_t r[4];
    philox_draw(c->cfg.seed, gid, (uint64_t)c->it, ORC_STREAM_DQN_ACT, r);
    const double u = (double)((((uint64_t)r[0] << 32) | r[1]) >> 11) * (1.0 / 9007199254740992.0);
    int action;
    if (u < eps) action = (int)(r[2] & 1u);                                   /* rand(action_space), dqn.jl:54 */
    else { dqn_forward(c->q, obs, h1, h2, q); action = q[1] > q[0] ? 1 : 0; } /* argmax: first maximum */
    float rew; uint8_t done;
    orc_cartpole_step(c->env_state + 4 * n, &c->env_t[n], action, c->cfg.max_episode_steps, &rew, &done);
    const int p = (c->ptr + n) % C;                                           /* add!, replay_buffer.jl:23-37 */
    memcpy(c->b_state + 4 * p, obs, 16); memcpy(c->b_next + 4 * p, c->env_state + 4 * n, 16);
    c->b_action[p] = action; c->b_reward[p] = rew; c->b_term[p] = done;
    c->ep_ret[n] += (double)rew; c->ep_len[n] += 1;
    if (done) {                                                               /* dqn.jl:80-86 */
      *episodes += 1; *sum_ret += c->ep_ret[n]; *sum_len += (double)c->ep_len[n];
      c->ep_ret[n] = 0.0; c->ep_len[n] = 0;
      float u4[4];
      orc_rng_reset_uniforms(c->cfg.seed, gid, c->resets[n], u4);
      c->resets[n] += 1;
      orc_cartpole_reset(c->env_state + 4 * n, &c->env_t[n], u4);
    }
  }
  c->ptr = (c->ptr + N) % C;
  c->size = c->size + N > C ? C : c->size + N;
}
/* Data-parallel DQN over `world` shards in lockstep (an extension: the reference has one env and one buffer). Every
 * shard acts with the shared parameters on its own envs (Philox keyed by the GLOBAL env id, so the union of the
 * shards' transitions equals a single run over all envs as long as the parameters agree), keeps its own ring, and
 * contributes batch_size samples of it to a global batch of world * batch_size: the partial gradients are summed in
 * rank order (a library allreduce may use another order: compare with a tolerance), Adam is applied identically
 * everywhere. world = 1 is orc_dqn_run. stats: one per shard, or NULL. */
int orc_dqn_group_run(orc_dqn_ctx** cs, int32_t world, int64_t iterations, crl_dqn_stats* stats) {
  if (!cs || world < 1 || iterations < 0) return -1;
  for (int r = 0; r < world; r++)
    if (!cs[r] || cs[r]->world != world || cs[r]->rank != r || cs[r]->it != cs[0]->it ||
        cs[r]->cfg.num_envs != cs[0]->cfg.num_envs || cs[r]->cfg.batch_size != cs[0]->cfg.batch_size) return -1;
  double sum_ret[64] = {0}, sum_len[64] = {0}, eps = 0.0;
  int64_t episodes[64] = {0};
  if (world > 64) return -1;
  const double n_global = (double)cs[0]->cfg.num_envs * (double)world;
  for (int64_t k = 0; k < iterations; k++) {
    int learn = 1;
    for (int r = 0; r < world; r++) {
      orc_dqn_ctx* c = cs[r];
      c->it += 1;
      const double gs = (double)c->it * n_global;
      eps = orc_dqn_linear_schedule(c->cfg.epsilon_start, c->cfg.epsilon_end, c->cfg.epsilon_duration, gs);
      dqn_act_iteration(c, eps, &sum_ret[r], &sum_len[r], &episodes[r]);
      learn = learn && gs > (double)c->cfg.min_buff_size && c->it % c->cfg.train_freq == 0 && c->size >= c->cfg.batch_size;  /* dqn.jl:94 */
    }
    if (!learn) continue;
    double sq = 0.0;
    float gsum[DQ_P];
    for (int r = 0; r < world; r++) {
      sq += dqn_learn_grad(cs[r]);
      for (int i = 0; i < DQ_P; i++) gsum[i] = r == 0 ? cs[r]->g[i] : gsum[i] + cs[r]->g[i];
    }
    for (int r = 0; r < world; r++) {
      orc_dqn_ctx* c = cs[r];
      memcpy(c->g, gsum, sizeof(gsum));
      c->last_loss = sq / ((double)c->cfg.batch_size * (double)world);
      dqn_adam(c);
      c->learn_steps += 1;
      if (c->it % c->cfg.target_net_freq == 0) memcpy(c->tgt, c->q, sizeof(float) * DQ_P);   /* dqn.jl:111-113 */
    }
  }
  if (stats)
    for (int r = 0; r < world; r++) {
      orc_dqn_ctx* c = cs[r];
      stats[r].last_loss = c->last_loss; stats[r].sum_return = sum_ret[r]; stats[r].sum_length = sum_len[r]; stats[r].epsilon = eps;
      stats[r].episodes = episodes[r]; stats[r].learn_steps = c->learn_steps; stats[r].iterations = c->it;
      stats[r].kernel_launches = 0;   /* the oracle launches nothing */
    }
  return 0;
}
int orc_dqn_set_shard(orc_dqn_ctx* c, int32_t world, int32_t rank, int32_t env_id_base) {
  if (!c || world < 1 || rank < 0 || rank >= world || env_id_base < 0) return -1;
  c->world = world; c->rank = rank; c->env_id_base = env_id_base;
  return 0;
}
int orc_dqn_run(orc_dqn_ctx* c, int64_t iterations, crl_dqn_stats* stats) {
  if (!c || c->world != 1) return -1;   /* a shard of a larger world only steps through orc_dqn_group_run */
  return orc_dqn_group_run(&c, 1, iterations, stats);
}
int orc_dqn_read_buffer(orc_dqn_ctx* c, float* state, int32_t* action, float* reward, float* next_state, uint8_t* terminal,
                        int32_t* size, int32_t* ptr) {
  if (!c) return -1;
  const size_t C = (size_t)c->cfg.buffer_size;
  if (state) memcpy(state, c->b_state, C * 16);
  if (next_state) memcpy(next_state, c->b_next, C * 16);
  if (action) memcpy(action, c->b_action, C * 4);
  if (reward) memcpy(reward, c->b_reward, C * 4);
  if (terminal) memcpy(terminal, c->b_term, C);
  if (size) *size = c->size;
  if (ptr) *ptr = c->ptr;
  return 0;
}
