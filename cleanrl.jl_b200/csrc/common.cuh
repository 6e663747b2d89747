// common.cuh — shared definitions for libcleanrl_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cleanrl_cuda.h"

#define CRL_H 64  // hidden width of both MLPs, networks.jl:36
#define CRL_MAX_ARRAYS 13
#define CRL_MAXD 4
#define CRL_MAXA 2

// Parameter layout = Flux.params(actor, critic) order (ppo.jl:196; networks.jl:36-49).
// Each W is (out,in) column-major as Flux stores it, i.e. element (j,k) at j + out*k, which
// read in C order is the k-major [in][out] matrix the forward GEMM wants.
struct Layout {
  int D, A, S, continuous, P, n_arrays;
  int off[CRL_MAX_ARRAYS], size[CRL_MAX_ARRAYS];
  int actor, critic, logstd;
};

__host__ __device__ inline bool make_layout(int env_kind, Layout* L) {
  if (env_kind == CRL_ENV_CARTPOLE) { L->D = 4; L->A = 2; L->S = 4; L->continuous = 0; }
  else if (env_kind == CRL_ENV_PENDULUM) { L->D = 3; L->A = 1; L->S = 2; L->continuous = 1; }
  else return false;
  int o = 0, i = 0;
  for (int net = 0; net < 2; net++) {
    const int O = net == 0 ? L->A : 1;
    if (net == 0) L->actor = o; else L->critic = o;
    const int sizes[6] = {CRL_H * L->D, CRL_H, CRL_H * CRL_H, CRL_H, O * CRL_H, O};
    for (int k = 0; k < 6; k++) { L->off[i] = o; L->size[i] = sizes[k]; o += sizes[k]; i++; }
  }
  L->logstd = -1;
  if (L->continuous) { L->logstd = o; L->off[i] = o; L->size[i] = L->A; o += L->A; i++; }
  for (int k = i; k < CRL_MAX_ARRAYS; k++) { L->off[k] = o; L->size[k] = 0; }
  L->P = o;
  L->n_arrays = i;
  return true;
}

// compile-time traits per environment
template <int ENV> struct EnvTraits;
template <> struct EnvTraits<CRL_ENV_CARTPOLE> {
  static constexpr int D = 4, A = 2, S = 4, CONT = 0;
  static constexpr int NET_A = CRL_H * 4 + CRL_H + CRL_H * CRL_H + CRL_H + 2 * CRL_H + 2;  // 4610
  static constexpr int NET_C = CRL_H * 4 + CRL_H + CRL_H * CRL_H + CRL_H + CRL_H + 1;      // 4545
  static constexpr int P = NET_A + NET_C;
};
template <> struct EnvTraits<CRL_ENV_PENDULUM> {
  static constexpr int D = 3, A = 1, S = 2, CONT = 1;
  static constexpr int NET_A = CRL_H * 3 + CRL_H + CRL_H * CRL_H + CRL_H + CRL_H + 1;  // 4481
  static constexpr int NET_C = NET_A;
  static constexpr int P = NET_A + NET_C + 1;
};
// offsets inside one net (O = output width)
template <int D, int O> struct NetOff {
  static constexpr int W1 = 0, B1 = CRL_H * D, W2 = B1 + CRL_H, B2 = W2 + CRL_H * CRL_H, W3 = B2 + CRL_H,
                       B3 = W3 + O * CRL_H, SIZE = B3 + O;
};

// device-resident counters the kernels read (keeps CUDA-graph launches parameter-free)
struct DevState {
  unsigned long long policy_step;   // global policy step (Philox action stream counter)
  unsigned long long update_index;  // PPO update counter (Philox permutation stream)
  double lr;                        // opt.eta for the current update (ppo.jl:120)
  int spec_failed;                  // multi-GPU: a speculative minibatch failed verification during this update
  int _pad;
};

// per-rollout episode bookkeeping (device)
struct EpisodeBuf {
  unsigned int count;  // records appended (may exceed capacity)
  unsigned int _pad;
  unsigned long long n_episodes;
  double sum_return;
  double sum_length;
  double max_return;
};

// minibatch statistics shared between mb_stats and loss_grad
struct MbScalars {
  double sum_adv, sum_adv2, sum_s;
  float min_vlc;
  float _pad;
};

#define CRL_STREAM_ACTION 0u
#define CRL_STREAM_RESET 1u
#define CRL_STREAM_PERM 2u

// ---- flag-in-data ("LL") packets for the peer-memory exchanges (NVLink / NVSwitch) ----------------------------------
// A Float64 travels as one 16-byte store {lo, flag, hi, flag}: each 8-byte half is written atomically, so a reader that
// sees `flag` in both halves has the whole value -- no fence, no separate flag store, no counter. flag = the exchange's
// sequence number (never 0, never the value two exchanges ago, which is what the slot held before).
#ifdef __CUDACC__
__device__ __forceinline__ void ll_store(uint4* p, double v, uint32_t flag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)b), "r"(flag), "r"((uint32_t)(b >> 32)), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
#endif
