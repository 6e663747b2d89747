// update.cu — the minibatch update (ppo.jl:191-252) around the forward/loss/backward kernel:
//
//   adv_stats   advantage sums of every minibatch of an update in one launch (speculative chain).
//   mb_stats    critic forward over the minibatch -> v_new[M] + per-CTA sums for the three minibatch-global scalars
//               the reference's loss needs before any gradient exists: mean/std of the advantages (ppo.jl:221) and
//               s = mean(newvalue .- R.^2) (ppo.jl:232, quirk Q5), plus min_i (clip_i - R_i)^2 (exact chain only).
//   mb_count    finalises those scalars and counts #{i : s > (clip_i - R_i)^2} (ppo.jl:236) (exact chain only).
//   loss_grad   the FP32 FFMA variant of the fused forward (actor+critic) + loss + full backward for tiles of 128
//               samples: every 64x64 contraction is a register-tiled FFMA GEMM out of shared memory. The default
//               variant is the tcgen05 kernel in update_tc.cu; this one serves the exact replay, the raw entry point,
//               A2C through crl_update_minibatch and CRL_NO_TC=1.
//   grad_reduce sums the per-CTA partial gradients in a fixed order (deterministic) into the double-precision vector
//               that is exchanged between GPUs (pushed into every peer's memory), or reduced with NCCL.
//   clip_adam   Flux.Optimiser(ClipNorm(0.5), Adam) per parameter array (ppo.jl:93,250) on thread-block clusters.
#include "kernels.h"
#include "mlp_tile.cuh"
#include "update_common.cuh"

namespace {

using namespace crl_upd;

// ------------------------------------------------------------------ mb_stats
template <int ENV> struct StatsSmem {
  using G = TileGeom<8, 1>;
  using E = EnvTraits<ENV>;
  static constexpr int SP = G::S_PAD;
  static constexpr int PARAMS = 0;
  static constexpr int X = (E::NET_C + 3) & ~3;
  static constexpr int H1 = X + CRL_MAXD * SP;
  static constexpr int H2 = H1 + CRL_H * SP;
  static constexpr int KEYS = H2 + CRL_H * SP;
  static constexpr int RED = KEYS + 8;  // 16 doubles
  static constexpr int FLOATS = RED + 32;
  static constexpr size_t BYTES = FLOATS * sizeof(float);
};

template <int ENV>
__global__ void __launch_bounds__(CRL_THREADS, 1) mb_stats_kernel(UpdateArgs a) {
  using G = TileGeom<8, 1>;
  using E = EnvTraits<ENV>;
  using SM = StatsSmem<ENV>;
  using NO = NetOff<E::D, 1>;
  constexpr int SP = G::S_PAD, S = G::S, D = E::D;
  extern __shared__ __align__(16) float smem[];
  float* cp = smem + SM::PARAMS;
  float* xs = smem + SM::X;
  float* h1 = smem + SM::H1;
  float* h2 = smem + SM::H2;
  uint32_t* keys = reinterpret_cast<uint32_t*>(smem + SM::KEYS);
  double* red = reinterpret_cast<double*>(smem + SM::RED);
  const ThreadCoord<G> tc;
  for (int i = threadIdx.x; i < E::NET_C; i += blockDim.x) cp[i] = a.params[E::NET_A + i];
  if (threadIdx.x == 0 && !a.idx.arr) perm_keys(a.idx.seed, a.idx.ds->update_index, a.idx.epoch, a.idx.rank, keys);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.fin->cnt = 0ull;
  __syncthreads();

  double s_adv = 0.0, s_adv2 = 0.0, s_s = 0.0;
  float mn = INFINITY;
  const int tid = threadIdx.x;
  for (int m0 = blockIdx.x * S; m0 < a.M; m0 += gridDim.x * S) {
    const int m = m0 + tid;
    const bool valid = m < a.M;
    float adv = 0.0f, R = 0.0f, V = 0.0f;
    float x[D];
#pragma unroll
    for (int k = 0; k < D; k++) x[k] = 0.0f;
    if (valid) {
      const int b = sample_index(a.idx, keys, m);
      if (D == 4) {
        const float4 v4 = reinterpret_cast<const float4*>(a.states)[b];
        x[0] = v4.x; x[1] = v4.y; x[2] = v4.z; x[D - 1] = v4.w;
      } else {
#pragma unroll
        for (int k = 0; k < D; k++) x[k] = a.states[(long long)b * D + k];
      }
      adv = a.advantages[b];
      R = a.returns[b];
      V = a.values[b];
    }
#pragma unroll
    for (int k = 0; k < D; k++) xs[k * SP + tid] = x[k];
    __syncthreads();
    tile_layer<G, D, EPI_BIAS_TANH>(tc, cp + NO::W1, cp + NO::B1, xs, h1);
    __syncthreads();
    tile_layer<G, CRL_H, EPI_BIAS_TANH>(tc, cp + NO::W2, cp + NO::B2, h1, h2);
    __syncthreads();
    float v = 0.0f;
#pragma unroll 8
    for (int k = 0; k < CRL_H; k++) v = fmaf(cp[NO::W3 + k], h2[k * SP + tid], v);
    v += cp[NO::B3];
    if (valid) {
      a.vnew[m] = v;
      const double ad = (double)adv;
      s_adv += ad;
      s_adv2 += ad * ad;
      s_s += (double)__fsub_rn(v, __fmul_rn(R, R));  // newvalue .- mb_returns .^ 2, ppo.jl:232
      float d, vlc;
      bool inside;
      value_clip(v, V, R, a.clip_coef, d, vlc, inside);
      mn = fminf(mn, vlc);
    }
  }
  const double t_adv = block_sum<8>(s_adv, red);
  const double t_adv2 = block_sum<8>(s_adv2, red);
  const double t_s = block_sum<8>(s_s, red);
  mn = warp_min(mn);
  __shared__ float mred[8];
  __syncthreads();
  if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m8 = mred[0];
#pragma unroll
    for (int w = 1; w < 8; w++) m8 = fminf(m8, mred[w]);
    MbScalars o;
    o.sum_adv = t_adv; o.sum_adv2 = t_adv2; o.sum_s = t_s; o.min_vlc = m8; o._pad = 0.0f;
    a.parts[blockIdx.x] = o;
  }
}

// ------------------------------------------------------------------ mb_count
template <int DUMMY>
__global__ void __launch_bounds__(CRL_THREADS) mb_count_kernel(UpdateArgs a) {
  __shared__ double red[8];
  __shared__ float mred[8];
  __shared__ uint32_t keys[8];
  double s_adv = 0.0, s_adv2 = 0.0, s_s = 0.0;
  float mn = INFINITY;
  for (int i = threadIdx.x; i < a.n_parts_in; i += blockDim.x) {
    const MbScalars p = a.parts_in[i];
    s_adv += p.sum_adv; s_adv2 += p.sum_adv2; s_s += p.sum_s; mn = fminf(mn, p.min_vlc);
  }
  // fixed-order (deterministic) reduction: lanes -> warps -> block
  const double t_adv = block_sum<8>(s_adv, red);
  const double t_adv2 = block_sum<8>(s_adv2, red);
  const double t_s = block_sum<8>(s_s, red);
  mn = warp_min(mn);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = mn;
  if (threadIdx.x == 0 && !a.idx.arr) perm_keys(a.idx.seed, a.idx.ds->update_index, a.idx.epoch, a.idx.rank, keys);
  __syncthreads();
  float m8 = mred[0];
#pragma unroll
  for (int w = 1; w < 8; w++) m8 = fminf(m8, mred[w]);
  const double Mg = (double)a.M * (double)a.world;
  const double mean = t_adv / Mg;
  double var = (t_adv2 - Mg * mean * mean) / (Mg - 1.0);  // corrected std, ppo.jl:221
  if (var < 0.0) var = 0.0;
  const float mean_f = (float)mean, std_f = (float)sqrt(var), s_f = (float)(t_s / Mg);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.fin->adv_mean = mean_f; a.fin->adv_std = std_f; a.fin->s_unclipped = s_f; a.fin->min_vlc = m8;
    a.fin->M_global = Mg;
    a.fin->need_fixup = 0;
  }
  if (!(s_f > m8)) return;  // common case: the scalar never wins the max, count is 0
  unsigned int c = 0;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += gridDim.x * blockDim.x) {
    const int b = sample_index(a.idx, keys, m);
    float d, vlc;
    bool inside;
    value_clip(a.vnew[m], a.values[b], a.returns[b], a.clip_coef, d, vlc, inside);
    c += (s_f > vlc) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&a.fin->cnt, (unsigned long long)c);
}

// ------------------------------------------------------------------ loss_grad
template <int ENV> struct LossSmem {
  using G = TileGeom<8, 2>;
  using E = EnvTraits<ENV>;
  static constexpr int SP = G::S_PAD, S = G::S;
  static constexpr int PARAMS = 0;
  static constexpr int W2T = PARAMS + SmemParams<ENV>::SIZE;  // [2][64][64], [net][j][k]
  static constexpr int X = W2T + 2 * CRL_H * CRL_H;
  static constexpr int H1 = X + CRL_MAXD * SP;
  static constexpr int H2 = H1 + G::ROWS * SP;
  static constexpr int ZO = H2 + G::ROWS * SP;  // z[A][S] then v[S]
  static constexpr int DOUT = ZO + 3 * S;       // dz[A][S] then dv[S]
  static constexpr int KEYS = DOUT + 3 * S;
  static constexpr int RED = KEYS + 8;
  static constexpr int XIN = RED + 32;           // [2][S][4] raw states of the current / next tile (cp.async target)
  static constexpr int SCAL = XIN + 2 * S * 4;   // [2][6][S] adv, old logprob, return, value, action (A <= 2 words)
  static constexpr int FLOATS = SCAL + 2 * 6 * S;
  static constexpr size_t BYTES = FLOATS * sizeof(float);
  static_assert(BYTES <= 227 * 1024, "loss_grad shared memory exceeds 227 KB");
};

template <int ENV> __global__ void param_image_kernel(const float* params, float* image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < EnvTraits<ENV>::P) image_scatter<ENV>(image, i, params[i]);
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Asynchronous gather of one tile (128 samples by permuted index) into shared memory: the loads of tile i+1
// are in flight while tile i runs its backward pass, so the random-access latency never stalls a barrier.
template <int ENV>
__device__ __forceinline__ void prefetch_tile(const UpdateArgs& a, const uint32_t* keys, int m0, float* xin, float* scal) {
  using E = EnvTraits<ENV>;
  constexpr int S = 128, D = E::D, A = E::A;
  const int tid = threadIdx.x;
  if (tid < S) {
    const int m = m0 + tid;
    float* x = xin + tid * 4;
    if (m < a.M) {
      const int b = sample_index(a.idx, keys, m);
      if (D == 4) {
        cp_async16(x, a.states + (long long)b * 4);
      } else {
#pragma unroll
        for (int k = 0; k < D; k++) cp_async4(x + k, a.states + (long long)b * D + k);
        x[3] = 0.0f;
      }
      cp_async4(scal + 0 * S + tid, a.advantages + b);
      cp_async4(scal + 1 * S + tid, a.logprobs + b);
      cp_async4(scal + 2 * S + tid, a.returns + b);
      cp_async4(scal + 3 * S + tid, a.values + b);
#pragma unroll
      for (int k = 0; k < A; k++)
        cp_async4(scal + (4 + k) * S + tid, reinterpret_cast<const float*>(a.actions) + (E::CONT ? (long long)b * A + k : (long long)b));
    } else {
      *reinterpret_cast<float4*>(x) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 6; k++) scal[k * S + tid] = 0.0f;
    }
  }
  cp_async_commit();
}

// barrier over the 4 warps that own one net (actor: warps 0-3, critic: warps 4-7); the two nets only meet at the
// CTA-wide barriers around the per-sample loss
__device__ __forceinline__ void net_sync(int net) { asm volatile("bar.sync %0, %1;" ::"r"(net + 1), "r"(128) : "memory"); }

template <int ENV>
__global__ void __launch_bounds__(CRL_THREADS, 1) loss_grad_kernel(UpdateArgs a) {
  using G = TileGeom<8, 2>;
  using E = EnvTraits<ENV>;
  using SM = LossSmem<ENV>;
  using NO = NetOff<E::D, 1>;
  using NA = NetOff<E::D, E::A>;
  constexpr int SP = G::S_PAD, S = G::S, D = E::D, A = E::A;
  extern __shared__ __align__(16) float smem[];
  float* sp = smem + SM::PARAMS;
  float* w2t = smem + SM::W2T;
  float* xs = smem + SM::X;
  float* h1 = smem + SM::H1;
  float* h2 = smem + SM::H2;
  float* zo = smem + SM::ZO;
  float* dout = smem + SM::DOUT;
  uint32_t* keys = reinterpret_cast<uint32_t*>(smem + SM::KEYS);
  double* red = reinterpret_cast<double*>(smem + SM::RED);
  const ThreadCoord<G> tc;
  const int tid = threadIdx.x;
  // a new peer exchange begins with this minibatch: grad_reduce and the finishing kernel only READ the counter
  if (a.p2p_seq && blockIdx.x == 0 && tid == 0) *a.p2p_seq += 1ull;
  float* xin = smem + SM::XIN;
  float* scal = smem + SM::SCAL;

  if (tid == 0 && !a.idx.arr) perm_keys(a.idx.seed, a.idx.ds->update_index, a.idx.epoch, a.idx.rank, keys);
  __syncthreads();
  prefetch_tile<ENV>(a, keys, blockIdx.x * S, xin, scal);  // overlaps the parameter staging below
  if (a.image) {
    // one elected thread stages parameters + W2^T with bulk asynchronous copies (TMA engine, no registers);
    // completion is signalled on an mbarrier by transaction bytes
    constexpr uint32_t IMG_BYTES = (SmemParams<ENV>::SIZE + 2 * CRL_H * CRL_H) * 4;
    constexpr uint32_t CHUNK = 16384;
    __shared__ __align__(8) unsigned long long pbar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&pbar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(IMG_BYTES) : "memory");
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sp);
      for (uint32_t off = 0; off < IMG_BYTES; off += CHUNK) {
        const uint32_t n = IMG_BYTES - off < CHUNK ? IMG_BYTES - off : CHUNK;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + off), "l"(reinterpret_cast<const char*>(a.image) + off), "r"(n), "r"(bar) : "memory");
      }
    }
    __syncthreads();  // the barrier object is initialised before anyone polls it
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(bar), "r"(0) : "memory");
  } else {
    load_params<ENV>(a.params, sp);
    __syncthreads();
    // W2T[net][j][k] = W2(j,k): the k-major operand of the dh1 = W2^T dz2 contraction
    for (int i = tid; i < 2 * CRL_H * CRL_H; i += blockDim.x) {
      const int net = i / (CRL_H * CRL_H), r = i % (CRL_H * CRL_H), j = r / CRL_H, k = r % CRL_H;
      w2t[i] = sp[net_base<ENV>(net) + NO::W2 + k * CRL_H + j];
    }
  }
  const bool spec = a.mode == LG_SPEC;
  float mean_f, std_f, s_f;
  double Mg, cnt_over_M;
  if (spec) {
    // advantage mean / corrected std of this minibatch from the per-update pre-pass (ppo.jl:221);
    // s is unknown yet: assume it never wins the max (checked by grad_reduce afterwards)
    double sa = 0.0, sa2 = 0.0;
#pragma unroll
    for (int i = 0; i < ADV_CHUNKS; i++) { sa += a.advparts[2 * i]; sa2 += a.advparts[2 * i + 1]; }
    Mg = (double)a.M * (double)a.world;
    const double mean = sa / Mg;
    double var = (sa2 - Mg * mean * mean) / (Mg - 1.0);
    if (var < 0.0) var = 0.0;
    mean_f = (float)mean; std_f = (float)sqrt(var); s_f = 0.0f; cnt_over_M = 0.0;
  } else {
    mean_f = a.fin->adv_mean; std_f = a.fin->adv_std; s_f = a.fin->s_unclipped;
    Mg = a.fin->M_global;
    cnt_over_M = (double)a.fin->cnt / Mg;
  }
  double st_s = 0.0;
  float st_min = INFINITY;
  const float c = a.clip_coef;
  const float lo_c = 1.0f - c, hi_c = 1.0f + c;
  __syncthreads();

  // persistent per-thread gradient accumulators (live for the whole kernel)
  float gW2[4][8];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) gW2[i][j] = 0.0f;
  float gW3 = 0.0f, gb3 = 0.0f;             // head phase owner (tid < 64*(A+1))
  float gW1[D], gb12 = 0.0f;                // row-sum phase: tid<128: dW1 + db1; tid>=128: db2
#pragma unroll
  for (int k = 0; k < D; k++) gW1[k] = 0.0f;
  double st_pg = 0.0, st_vmax = 0.0, st_ent = 0.0, g_logstd[A];
#pragma unroll
  for (int k = 0; k < A; k++) g_logstd[k] = 0.0;

  // phase-specific coordinates
  const int w_net = tid >> 7;          // dW2 phase: warps 0-3 actor, 4-7 critic
  const int w_jq = tid & 7;            // j rows jq + 8*i2
  const int w_kq = (tid & 127) >> 3;   // k rows kq + 16*i
  // head-backward phase: threads of a net own (output o, row k) pairs of that net's head
  const int p5_tl = tid & 127, p5_nout = w_net == 0 ? A : 1;
  const bool p5_active = p5_tl < CRL_H * p5_nout;
  const int p5_k = p5_tl % CRL_H;
  const int p5_o = w_net == 0 ? p5_tl / CRL_H : A;   // index into dout: actor outputs 0..A-1, critic = A
  // row-sum phase: each net's threads own that net's rows
  const int p9_kind = (tid & 127) >> 6;              // 0: dz1 row (db1, dW1), 1: dz2 row (db2)
  const int p9_row = w_net * CRL_H + (tid & 63);

  int tile_no = 0;
  for (int m0 = blockIdx.x * S; m0 < a.M; m0 += gridDim.x * S) {
    // ---- P0: this tile's inputs were gathered asynchronously (prefetch_tile); unpack them
    const int buf = tile_no & 1;
    cp_async_wait_all();
    __syncthreads();
    float s_adv = 0.0f, s_oldlp = 0.0f, s_R = 0.0f, s_V = 0.0f, s_actf[A];
    int s_act = 0;
    bool valid = false;
#pragma unroll
    for (int k = 0; k < A; k++) s_actf[k] = 0.0f;
    if (tid < S) {
      valid = m0 + tid < a.M;
      const float4 x4 = *reinterpret_cast<const float4*>(xin + buf * S * 4 + tid * 4);
      const float xk[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int k = 0; k < D; k++) xs[k * SP + tid] = xk[k];
      const float* sc = scal + buf * 6 * S;
      s_adv = sc[0 * S + tid];
      s_oldlp = sc[1 * S + tid];
      s_R = sc[2 * S + tid];
      s_V = sc[3 * S + tid];
      if (E::CONT) {
#pragma unroll
        for (int k = 0; k < A; k++) s_actf[k] = sc[(4 + k) * S + tid];
      } else {
        s_act = __float_as_int(sc[4 * S + tid]);
      }
    }
    __syncthreads();
    // next tile's gather goes out now and lands while this tile computes
    if (m0 + gridDim.x * S < a.M) prefetch_tile<ENV>(a, keys, m0 + gridDim.x * S, xin + (buf ^ 1) * S * 4, scal + (buf ^ 1) * 6 * S);
    tile_no++;
    // ---- P1/P2 forward, both nets (logprob_actions ppo.jl:35 and critic ppo.jl:214)
    {
      const float* np = sp + net_base<ENV>(tc.net);
      tile_layer<G, D, EPI_BIAS_TANH, false, true>(tc, np + NO::W1, np + NO::B1, xs, h1 + tc.net * CRL_H * SP);
      net_sync(tc.net);
      tile_layer<G, CRL_H, EPI_BIAS_TANH, true, true>(tc, np + NO::W2, np + NO::B2, h1 + tc.net * CRL_H * SP, h2 + tc.net * CRL_H * SP);
      net_sync(tc.net);
    }
    // ---- P3 heads
    {
      const int s = tid & (S - 1);
      if (tid < S) {
        const float* ap = sp + SmemParams<ENV>::ACTOR;
        float acc[A];
#pragma unroll
        for (int o = 0; o < A; o++) acc[o] = 0.0f;
#pragma unroll 8
        for (int k = 0; k < CRL_H; k++) {
          const float h = h2[k * SP + (s ^ act_swz(k))];
#pragma unroll
          for (int o = 0; o < A; o++) acc[o] = fmaf(ap[NA::W3 + k * A + o], h, acc[o]);
        }
#pragma unroll
        for (int o = 0; o < A; o++) zo[o * S + s] = acc[o] + ap[NA::B3 + o];
      } else {
        const float* cp = sp + SmemParams<ENV>::CRITIC;
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < CRL_H; k++) acc = fmaf(cp[NO::W3 + k], h2[(CRL_H + k) * SP + (s ^ act_swz(k))], acc);
        zo[A * S + s] = acc + cp[NO::B3];
      }
    }
    __syncthreads();
    // ---- P4 per-sample loss and its gradient w.r.t. the head outputs (ppo.jl:213-243)
    if (tid < S) {
      const int s = tid;
      float dz[A], dv = 0.0f;
#pragma unroll
      for (int k = 0; k < A; k++) dz[k] = 0.0f;
      if (valid) {
        float z[A];
#pragma unroll
        for (int k = 0; k < A; k++) z[k] = zo[k * S + s];
        const float v = zo[A * S + s];
        // (adv .- mean) ./ (std .+ 1e-8): Float32 numerator, Float64 quotient (Q6)
        const double adv_n = (double)__fsub_rn(s_adv, mean_f) / ((double)std_f + 1e-8);
        float newlp, p[A], lp[A];
        double ent_sum = 0.0;
        if (!E::CONT) {
          float m = z[0];
#pragma unroll
          for (int k = 1; k < A; k++) m = fmaxf(m, z[k]);
          float ex[A], sum = 0.0f;
#pragma unroll
          for (int k = 0; k < A; k++) { ex[k] = expf(__fsub_rn(z[k], m)); sum = __fadd_rn(sum, ex[k]); }
          const float ls = logf(sum);
          newlp = 0.0f;
#pragma unroll
          for (int k = 0; k < A; k++) {
            p[k] = __fdiv_rn(ex[k], sum);
            lp[k] = __fsub_rn(__fsub_rn(z[k], m), ls);
            ent_sum += (double)(-__fmul_rn(p[k], lp[k]));  // ppo.jl:42 (Q4: A x M matrix)
            if (k == s_act) newlp = lp[k];
          }
        } else {
          float acc = 0.0f;
#pragma unroll
          for (int k = 0; k < A; k++) {
            const float logstd = sp[SmemParams<ENV>::LOGSTD + k];
            const float sd = expf(logstd);
            const float diff = __fsub_rn(s_actf[k], z[k]);
            const float q = __fdiv_rn(-__fmul_rn(diff, diff), __fmul_rn(__fmul_rn(2.0f, sd), sd));
            acc = __fadd_rn(acc, __fsub_rn(__fsub_rn(q, logstd), 0.9189385332046727f));
            ent_sum += (double)__fadd_rn(__fadd_rn(0.5f, 0.9189385332046727f), logstd);
            p[k] = 0.0f; lp[k] = 0.0f;
          }
          newlp = acc;
        }
        if (a.algo == 1) {
          // A2C (a2c.jl:78-97): critic_loss = mean((R - v)^2), gradient through v; actor_loss = -mean(logp .* (R - v))
          // with the advantage held constant. One combined step == the reference's two update! calls (disjoint
          // parameter sets, per-array ClipNorm/Adam).
          const double advd = (double)s_R - (double)v;
          st_vmax += advd * advd;
          st_pg += -(double)newlp * advd;
          dv = (float)(-2.0 * advd / Mg);
          const double g_a = -advd / Mg;
          if (!E::CONT) {
#pragma unroll
            for (int k = 0; k < A; k++) dz[k] = (float)(g_a * ((k == s_act ? 1.0 : 0.0) - (double)p[k]));
          } else {
#pragma unroll
            for (int k = 0; k < A; k++) {
              const float sd = expf(sp[SmemParams<ENV>::LOGSTD + k]);
              const double diff = (double)__fsub_rn(s_actf[k], z[k]);
              const double var = (double)sd * (double)sd;
              dz[k] = (float)(g_a * diff / var);
              g_logstd[k] += g_a * (diff * diff / var - 1.0);
            }
          }
        } else {
        const float logratio = __fsub_rn(newlp, s_oldlp);  // ppo.jl:224
        const float ratio = expf(logratio);                // ppo.jl:225
        const float rc = ratio < lo_c ? lo_c : (ratio > hi_c ? hi_c : ratio);
        const double pg1 = -adv_n * (double)ratio;  // ppo.jl:226
        const double pg2 = -adv_n * (double)rc;     // ppo.jl:227
        double pgm, dratio;
        if (pg1 > pg2) { pgm = pg1; dratio = -adv_n; }
        else { pgm = pg2; dratio = (ratio >= lo_c && ratio <= hi_c) ? -adv_n : 0.0; }
        st_pg += pgm;
        const double g_lp = dratio * (double)ratio / Mg;
        // value loss (Q5): 0.5*mean(max.(s, (clip - R)^2)), s a minibatch scalar
        float d_vcR, vlc;
        bool inside;
        value_clip(v, s_V, s_R, c, d_vcR, vlc, inside);
        const bool s_wins = !spec && s_f > vlc && !a.no_vclip;
        double dv_d = cnt_over_M;
        if (!s_wins && inside) dv_d += 2.0 * (double)d_vcR;
        if (a.no_vclip) {
          // clip_value_loss = false: 0.5 * mean((newvalue - R).^2), ppo.jl:239-241 (no minibatch scalar, nothing to verify)
          const float d = __fsub_rn(v, s_R);
          vlc = __fmul_rn(d, d);
          dv_d = 2.0 * (double)d;
          if (spec) a.vnew[m0 + s] = v;
        } else if (spec) {
          st_s += (double)__fsub_rn(v, __fmul_rn(s_R, s_R));  // newvalue .- mb_returns .^ 2, ppo.jl:232
          st_min = fminf(st_min, vlc);
          a.vnew[m0 + s] = v;
        }
        st_vmax += (double)(s_wins ? s_f : vlc);
        dv = (float)((double)a.v_coef * 0.5 / Mg * dv_d);
        st_ent += ent_sum;
        const double ent_scale = (double)a.ent_coeff / ((double)A * Mg);
        if (!E::CONT) {
#pragma unroll
          for (int k = 0; k < A; k++) {
            double d = g_lp * ((k == s_act ? 1.0 : 0.0) - (double)p[k]);
            d += ent_scale * (double)p[k] * ((double)lp[k] + ent_sum);
            dz[k] = (float)d;
          }
        } else {
#pragma unroll
          for (int k = 0; k < A; k++) {
            const float sd = expf(sp[SmemParams<ENV>::LOGSTD + k]);
            const double diff = (double)__fsub_rn(s_actf[k], z[k]);
            const double var = (double)sd * (double)sd;
            dz[k] = (float)(g_lp * diff / var);
            g_logstd[k] += g_lp * (diff * diff / var - 1.0) - ent_scale;
          }
        }
        }  // PPO
      }
#pragma unroll
      for (int k = 0; k < A; k++) dout[k * S + s] = dz[k];
      dout[A * S + s] = dv;
    }
    __syncthreads();
    // ---- P5 head backward: dW3 += h2 * dout^T, db3 += sum dout
    if (p5_active) {
      const int k = p5_k;
      const float* hrow = h2 + (w_net * CRL_H + k) * SP;
      const float* drow = dout + p5_o * S;
      const int hsw = act_swz(k);
      float acc = 0.0f, bacc = 0.0f;
#pragma unroll 4
      for (int s = 0; s < S; s += 4) {
        const float4 h = *reinterpret_cast<const float4*>(hrow + (s ^ hsw));
        const float4 d = *reinterpret_cast<const float4*>(drow + s);
        acc = fmaf(h.x, d.x, acc); acc = fmaf(h.y, d.y, acc); acc = fmaf(h.z, d.z, acc); acc = fmaf(h.w, d.w, acc);
        bacc += (d.x + d.y) + (d.z + d.w);
      }
      gW3 += acc;
      gb3 += bacc;
    }
    net_sync(w_net);
    // ---- P6 dz2 = (W3^T dout) .* (1 - h2^2), in place over this thread's own h2 tile
    {
      float* hb = h2 + tc.net * CRL_H * SP;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int row = tc.jb[j / 4] + (j % 4);
        float w3[A];
        if (tc.net == 0) {
#pragma unroll
          for (int o = 0; o < A; o++) w3[o] = sp[SmemParams<ENV>::ACTOR + NA::W3 + row * A + o];
        } else {
          w3[0] = sp[SmemParams<ENV>::CRITIC + NO::W3 + row];
        }
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
          float4* hp = reinterpret_cast<float4*>(hb + row * SP + (tc.sb[cc] ^ act_swz(row)));
          const float4 h = *hp;
          float4 dh;
          if (tc.net == 0) {
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < A; o++) {
              const float4 d = *reinterpret_cast<const float4*>(dout + o * S + tc.sb[cc]);
              acc4.x = fmaf(w3[o], d.x, acc4.x); acc4.y = fmaf(w3[o], d.y, acc4.y);
              acc4.z = fmaf(w3[o], d.z, acc4.z); acc4.w = fmaf(w3[o], d.w, acc4.w);
            }
            dh = acc4;
          } else {
            const float4 d = *reinterpret_cast<const float4*>(dout + A * S + tc.sb[cc]);
            dh = make_float4(w3[0] * d.x, w3[0] * d.y, w3[0] * d.z, w3[0] * d.w);
          }
          float4 o4;
          o4.x = dh.x * (1.0f - h.x * h.x); o4.y = dh.y * (1.0f - h.y * h.y);
          o4.z = dh.z * (1.0f - h.z * h.z); o4.w = dh.w * (1.0f - h.w * h.w);
          *hp = o4;
        }
      }
    }
    net_sync(w_net);
    // ---- P7 dW2 += dz2 * h1^T (per net): 4 k-rows x 8 j-rows per thread, reduce over samples
    {
      const float* h1b = h1 + w_net * CRL_H * SP;
      const float* d2b = h2 + w_net * CRL_H * SP;
      // packed FP32 (FFMA2): each accumulator pair holds the partial sums over the even / odd samples
      float2 acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = make_float2(0.0f, 0.0f);
#pragma unroll 2
      for (int s = 0; s < S; s += 4) {
        float4 hk[4], dj[8];
#pragma unroll
        for (int i = 0; i < 4; i++) hk[i] = *reinterpret_cast<const float4*>(h1b + (w_kq + 16 * i) * SP + (s ^ act_swz(w_kq + 16 * i)));
#pragma unroll
        for (int j = 0; j < 8; j++) dj[j] = *reinterpret_cast<const float4*>(d2b + (w_jq + 8 * j) * SP + (s ^ act_swz(w_jq + 8 * j)));
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) {
            acc[i][j] = __ffma2_rn(make_float2(hk[i].x, hk[i].y), make_float2(dj[j].x, dj[j].y), acc[i][j]);
            acc[i][j] = __ffma2_rn(make_float2(hk[i].z, hk[i].w), make_float2(dj[j].z, dj[j].w), acc[i][j]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) gW2[i][j] += acc[i][j].x + acc[i][j].y;
    }
    net_sync(w_net);
    // ---- P8 dz1 = (W2^T dz2) .* (1 - h1^2), in place over h1
    tile_layer<G, CRL_H, EPI_DTANH, true, true>(tc, w2t + tc.net * CRL_H * CRL_H, nullptr, h2 + tc.net * CRL_H * SP,
                                                h1 + tc.net * CRL_H * SP);
    net_sync(w_net);
    // ---- P9 row sums of this net's rows: first 64 threads of the net: db1[row], dW1[k][row]; other 64: db2[row]
    {
      const int row = p9_row;
      const float* src = (p9_kind == 0 ? h1 : h2) + row * SP;
      const int rsw = act_swz(row);
      float bacc = 0.0f, wacc[D];
#pragma unroll
      for (int k = 0; k < D; k++) wacc[k] = 0.0f;
#pragma unroll 4
      for (int s = 0; s < S; s += 4) {
        const float4 d = *reinterpret_cast<const float4*>(src + (s ^ rsw));
        bacc += (d.x + d.y) + (d.z + d.w);
        if (p9_kind == 0) {
#pragma unroll
          for (int k = 0; k < D; k++) {
            const float4 x4 = *reinterpret_cast<const float4*>(xs + k * SP + s);
            wacc[k] = fmaf(x4.x, d.x, wacc[k]); wacc[k] = fmaf(x4.y, d.y, wacc[k]);
            wacc[k] = fmaf(x4.z, d.z, wacc[k]); wacc[k] = fmaf(x4.w, d.w, wacc[k]);
          }
        }
      }
      gb12 += bacc;
#pragma unroll
      for (int k = 0; k < D; k++) gW1[k] += wacc[k];
    }
    // no barrier here: the next tile starts with a CTA-wide barrier (or the kernel epilogue follows)
  }

  // ---- write this CTA's partial gradient (every element of [0,P) exactly once)
  float* gp = a.gpart + (long long)blockIdx.x * E::P;
  {
    const int nb = w_net == 0 ? 0 : E::NET_A;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) gp[nb + NO::W2 + (w_jq + 8 * j) + CRL_H * (w_kq + 16 * i)] = gW2[i][j];
  }
  if (p5_active) {
    const int k = p5_k;
    if (w_net == 0) {
      gp[NA::W3 + p5_o + A * k] = gW3;
      if (k == 0) gp[NA::B3 + p5_o] = gb3;
    } else {
      gp[E::NET_A + NO::W3 + k] = gW3;
      if (k == 0) gp[E::NET_A + NO::B3] = gb3;
    }
  }
  {
    const int j = tid & 63;
    const int nb = w_net == 0 ? 0 : E::NET_A;
    if (p9_kind == 0) {
      gp[nb + NO::B1 + j] = gb12;
#pragma unroll
      for (int k = 0; k < D; k++) gp[nb + NO::W1 + j + CRL_H * k] = gW1[k];
    } else {
      gp[nb + NO::B2 + j] = gb12;
    }
  }
  const double t_pg = block_sum<8>(st_pg, red);
  const double t_vm = block_sum<8>(st_vmax, red);
  const double t_en = block_sum<8>(st_ent, red);
  const double t_ss = block_sum<8>(st_s, red);
  __shared__ float mred_s[8];
  st_min = warp_min(st_min);
  __syncthreads();
  if ((tid & 31) == 0) mred_s[tid >> 5] = st_min;
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int w = 1; w < 8; w++) mred_s[0] = fminf(mred_s[0], mred_s[w]);
  }
  double t_ls[A];
#pragma unroll
  for (int k = 0; k < A; k++) t_ls[k] = block_sum<8>(g_logstd[k], red);
  if (tid == 0) {
    double* spp = a.spart + (long long)blockIdx.x * 4;
    spp[0] = t_pg; spp[1] = t_vm; spp[2] = t_en; spp[3] = t_ss;
    if (spec) a.mpart[blockIdx.x] = mred_s[0];
    if (E::CONT) {
#pragma unroll
      for (int k = 0; k < A; k++) gp[E::NET_A + E::NET_C + k] = (float)t_ls[k];
    }
  }
}

// ------------------------------------------------------------------ grad_reduce
// Sums the per-CTA partials in a fixed order (deterministic): 64 elements per block, the partials
// split over 4 thread groups so ~37 independent loads are in flight per thread. The extra last
// block verifies the speculation of LG_SPEC: s = mean(v_new - R^2) must not exceed
// min_i (clip_i - R_i)^2, otherwise the exact kernels that follow redo the minibatch.
__global__ void __launch_bounds__(256) grad_reduce_kernel(UpdateArgs a, int P) {
  __shared__ double sh[4][64];
  const int grid = a.grid_loss;
  // with the peer-memory allreduce the sums are pushed straight into every rank's exchange buffer
  const bool push = a.p2p_data != nullptr;
  const unsigned long long seq = push ? *a.p2p_seq : 0ull;
  const size_t push_off = push ? ((size_t)(seq & 1ull) * a.world + a.rank) * a.p2p_stride : 0;
  double* gsum = a.gsum;
  auto put = [&](int e, double v) {
    if (push) {
      for (int r = 0; r < a.world; r++) reinterpret_cast<double*>(a.p2p_peers[r])[push_off + e] = v;
    } else {
      gsum[e] = v;
    }
  };
  // every block that has pushed its part counts itself in; the last one raises this rank's flag on all peers
  auto arrive = [&]() {
    if (!push) return;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();  // cumulative: orders the whole block's pushes (CTA barrier above) before the count
      const unsigned int old = atomicAdd(a.p2p_count, 1u);
      if ((old + 1u) % gridDim.x == 0u) {
        __threadfence_system();
        for (int r = 0; r < a.world; r++)
          *(reinterpret_cast<volatile unsigned long long*>(a.p2p_peers[r] + a.p2p_flags_off) + a.rank) = seq;
      }
    }
  };
  if (blockIdx.x == gridDim.x - 1) {
    if (a.mode != LG_SPEC) { arrive(); return; }   // exact chain: mb_count has already produced the statistics
    // speculative chain: sum s and this rank's min (clip_i - R_i)^2 ride behind the gradient (every other rank adds 0
    // in this rank's min slot), so ONE sum-exchange delivers gradient, loss sums, sum s and every rank's min; the
    // finishing kernel checks s <= min on the exchanged values
    __shared__ float mred[8];
    float mn = INFINITY;
    for (int c = threadIdx.x; c < grid; c += blockDim.x) mn = fminf(mn, a.mpart[c]);
    mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m8 = mred[0];
      for (int w = 1; w < 8; w++) m8 = fminf(m8, mred[w]);
      for (int r = 0; r < a.world; r++) put(P + 4 + r, (r == a.rank) ? (double)m8 : 0.0);
    }
    arrive();
    return;
  }
  const int el = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + el;
  double s = 0.0;
  if (e < P) {
    int c_lo = 0, c_hi = grid;
    if (a.tc_actor_ctas) {  // one net per CTA: only that net's CTAs hold element e
      const bool critic = e >= a.tc_net_a && e < a.tc_net_a + a.tc_net_c;
      if (critic) c_lo = a.tc_actor_ctas; else c_hi = a.tc_actor_ctas;
    }
#pragma unroll 8
    for (int c = c_lo + g; c < c_hi; c += 4) s += (double)a.gpart[(long long)c * P + e];
  } else if (e < P + 4) {
    for (int c = g; c < grid; c += 4) s += a.spart[(long long)c * 4 + (e - P)];
  }
  sh[g][el] = s;
  __syncthreads();
  if (g == 0 && e < P + 4) put(e, (sh[0][el] + sh[1][el]) + (sh[2][el] + sh[3][el]));
  arrive();
}

// advantage sums of every minibatch of an update in one launch: grid (ADV_CHUNKS, n_sets)
__global__ void __launch_bounds__(256) adv_stats_kernel(AdvStatsArgs a) {
  __shared__ double red[8];
  __shared__ uint32_t keys[8];
  const int set = blockIdx.y, epoch = set / a.nmb, start = (set % a.nmb) * a.M;
  IdxSrc ix = a.idx;
  ix.epoch = (uint32_t)epoch;
  ix.start = (uint32_t)start;
  ix.arr = a.arr_base ? a.arr_base + (long long)epoch * a.B + start : nullptr;
  if (threadIdx.x == 0 && !ix.arr) perm_keys(ix.seed, ix.ds->update_index, ix.epoch, ix.rank, keys);
  __syncthreads();
  const int per = (a.M + ADV_CHUNKS - 1) / ADV_CHUNKS;
  const int lo = blockIdx.x * per, hi = min(a.M, lo + per);
  double sa = 0.0, sa2 = 0.0;
#pragma unroll 4
  for (int m = lo + threadIdx.x; m < hi; m += blockDim.x) {
    const double ad = (double)a.advantages[sample_index(ix, keys, m)];
    sa += ad;
    sa2 += ad * ad;
  }
  sa = block_sum<8>(sa, red);
  sa2 = block_sum<8>(sa2, red);
  if (threadIdx.x < 2) {
    const int pos = (set * ADV_CHUNKS + (int)blockIdx.x) * 2 + (int)threadIdx.x;
    double v = threadIdx.x ? sa2 : sa;
    if (a.x_local) {
      // global sums over the ranks: push this block's value to every peer, collect the peers' values of the same block
      // from this rank's own memory, add in rank order (identical on every rank). No other block is waited for.
      const unsigned long long u = a.idx.ds->update_index;
      const uint32_t flag = (uint32_t)(u + 1ull);
      const size_t row0 = (size_t)(u & 1ull) * a.x_world * a.x_stride;
      for (int r = 0; r < a.x_world; r++)
        if (r != a.x_rank)
          ll_store(reinterpret_cast<uint4*>(a.x_peers[r] + a.x_off) + row0 + (size_t)a.x_rank * a.x_stride + pos, v, flag);
      const uint4* src = reinterpret_cast<const uint4*>(a.x_local + a.x_off) + row0 + pos;
      const long long t0 = clock64();
      double tot = 0.0;
      for (int r = 0; r < a.x_world; r++) {
        double x = v;
        if (r != a.x_rank) {
          uint4 pk = ll_load(src + (size_t)r * a.x_stride);
          while (pk.y != flag || pk.w != flag) {
            if (clock64() - t0 > a.x_timeout) { atomicExch(a.x_err, 1); break; }   // no update is applied after this
            pk = ll_load(src + (size_t)r * a.x_stride);
          }
          x = __longlong_as_double((long long)(((unsigned long long)pk.z << 32) | pk.x));
        }
        tot += x;
      }
      v = tot;
    }
    a.advparts[pos] = v;
  }
}

// raw path: Float32 gradient + loss scalars out of the double sums
__global__ void loss_finalize_kernel(const double* gsum, int P, float* grads_out, double Mg, int A, float ent_coeff,
                                     float v_coef, double* stats_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < P) grads_out[e] = (float)gsum[e];
  if (e == 0 && stats_out) finalize_stats(gsum + P, Mg, A, ent_coeff, v_coef, stats_out);
}

// reduce the per-CTA partial sums of mb_stats to one record (multi-GPU all-gather payload)
__global__ void stats_pack_kernel(const MbScalars* parts, int n, MbScalars* out) {
  __shared__ double red[8];
  __shared__ float mred[8];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  float mn = INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const MbScalars p = parts[i];
    s0 += p.sum_adv; s1 += p.sum_adv2; s2 += p.sum_s; mn = fminf(mn, p.min_vlc);
  }
  s0 = block_sum<8>(s0, red); s1 = block_sum<8>(s1, red); s2 = block_sum<8>(s2, red);
  mn = warp_min(mn);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m8 = mred[0];
    for (int w = 1; w < 8; w++) m8 = fminf(m8, mred[w]);
    MbScalars o;
    o.sum_adv = s0; o.sum_adv2 = s1; o.sum_s = s2; o.min_vlc = m8; o._pad = 0.0f;
    *out = o;
  }
}

// ------------------------------------------------------------------ clip + Adam
// Flux.Optimiser(ClipNorm(thresh), Adam(η)) [Flux 0.13.4], one CTA per parameter array.
// Finishing kernel of a minibatch, one CTA per parameter array (ppo.jl:250):
//   [multi-GPU] one-shot exchange of the reduced sums over NVLink / NVSwitch peer memory: grad_reduce has PUSHED this
//   rank's vector into every rank's exchange buffer (posted remote stores) and its last block raised this rank's
//   sequence flag on every peer; here each block spins on its LOCAL flags and adds the world vectors out of LOCAL
//   memory in rank order, so all ranks hold bit-identical sums and no remote load sits on the critical path. Two
//   slots alternate; a slot is rewritten only after a complete exchange in between, which implies every peer
//   finished reading the older contents.
//   [speculative path] block 0 verifies s = mean(v_new - R^2) <= min_i (clip_i - R_i)^2 with the exchanged values.
//   Then Flux.Optimiser(ClipNorm, Adam): Float32 norm per array, clip, Adam with Float64 scalars, parameter image.
__device__ __forceinline__ double finish_load(const AdamArgs& a, int slot, int e) {
  if (a.p2p_local) {  // every rank's sums were pushed into this rank's buffer: add them in rank order
    double s = 0.0;
    for (int r = 0; r < a.world; r++)
      s += __ldcv(a.p2p_local + ((size_t)slot * a.world + r) * a.p2p_stride + e);
    return s;
  }
  return a.gsum[e];
}

constexpr int ADAM_CL = 8;     // CTAs per parameter array = one thread-block cluster
constexpr int ADAM_NT = 256;   // threads per CTA; 8 x 256 x 2 elements covers the largest array (4096)

__global__ void __cluster_dims__(ADAM_CL, 1, 1) __launch_bounds__(ADAM_NT) clip_adam_kernel(AdamArgs a) {
  // The Float64 Adam arithmetic of a 4096-element array would keep ONE SM's FP64 pipe busy for ~8 us, so each array is
  // spread over a cluster of 8 CTAs; the per-array L2 norm (ClipNorm is per array, ppo.jl:93) is combined through
  // distributed shared memory: every CTA stores its partial sum of squares into all 8 CTAs' shared memory, one
  // cluster barrier, then everybody adds the 8 partials in the same order.
  __shared__ double red[8];
  __shared__ double part[ADAM_CL];
  __shared__ int badpart[ADAM_CL];
  Layout L;
  make_layout(a.env_kind, &L);
  const int i = blockIdx.x / ADAM_CL;         // parameter array
  unsigned int crank;                          // rank of this CTA inside its cluster
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  // Distributed shared memory of a peer CTA may only be written once that CTA is known to be running: arrive on the
  // cluster barrier now, wait for it just before the remote stores (compute-sanitizer racecheck flags it otherwise).
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  const int o = L.off[i], n = L.size[i];
  int slot = 0;
  int bad = 0;   // a peer never delivered (now or in an earlier minibatch): nobody applies an update built on partial sums
  if (a.p2p_local) {
    const unsigned long long q = *a.p2p_seq;
    slot = (int)(q & 1ull);
    if (*reinterpret_cast<volatile int*>(a.p2p_err) != 0) bad = 1;   // sticky: the host reports CRL_ERR_NCCL at the next fetch
    if (!bad && threadIdx.x < a.world) {   // wait until every rank's grad_reduce has pushed this minibatch's sums here
      const volatile unsigned long long* mine =
          reinterpret_cast<const volatile unsigned long long*>(reinterpret_cast<const unsigned char*>(a.p2p_local) + a.p2p_flags_off) + threadIdx.x;
      const long long t0 = clock64();
      while (*mine < q) {
        if (clock64() - t0 > a.timeout_cycles) { atomicExch(a.p2p_err, 1); bad = 1; break; }  // fail instead of hanging
      }
      __threadfence_system();
    }
    bad = __syncthreads_or(bad);
  }
  // this CTA's slice of the array: elements crank*512 + tid + c*256, c < 2
  float g[2];
  double ss = 0.0;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const int k = (int)crank * (2 * ADAM_NT) + threadIdx.x + c * ADAM_NT;
    g[c] = 0.0f;
    if (k < n) {
      g[c] = a.gf ? a.gf[o + k] : (float)(finish_load(a, slot, o + k) * a.grad_scale);  // the Float32 gradient Zygote returns
      ss += (double)g[c] * (double)g[c];
    }
  }
  ss = block_sum<8>(ss, red);
  const double bp1 = a.beta_pow[2 * i], bp2 = a.beta_pow[2 * i + 1];  // read BEFORE the cluster barrier (rank 0 rewrites them after)
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA of the cluster has started
  if (threadIdx.x < ADAM_CL) {
    // distributed shared memory: part[crank] (and the error flag) of CTA `threadIdx.x` of this cluster
    unsigned int local = (unsigned int)__cvta_generic_to_shared(&part[crank]), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(threadIdx.x));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(ss) : "memory");
    local = (unsigned int)__cvta_generic_to_shared(&badpart[crank]);
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(threadIdx.x));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(bad) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  double tot = 0.0;
#pragma unroll
  for (int r = 0; r < ADAM_CL; r++) { tot += part[r]; bad |= badpart[r]; }
  if (bad) return;   // the whole cluster (= the whole parameter array) agrees: no Adam step, no beta powers, no statistics
  const float nrm = (float)sqrt(tot);  // norm(Δ::Array{Float32})::Float32
  const bool clip = (double)nrm > (double)a.clip_norm;
  const double scale = clip ? (double)a.clip_norm / (double)nrm : 1.0;
  const double lr = a.lr_host >= 0.0 ? a.lr_host : a.ds->lr;
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const int k = (int)crank * (2 * ADAM_NT) + threadIdx.x + c * ADAM_NT;
    if (k >= n) continue;
    if (a.grads_out) a.grads_out[o + k] = g[c];
    float d = g[c];
    if (clip) d = (float)__dmul_rn((double)d, scale);  // rmul!(Δ, thresh/nrm)
    const float mt = (float)__dadd_rn(__dmul_rn(b1, (double)a.m[o + k]), __dmul_rn(1.0 - b1, (double)d));
    const float vt = (float)__dadd_rn(__dmul_rn(b2, (double)a.v[o + k]), __dmul_rn(__dmul_rn(1.0 - b2, (double)d), (double)d));
    a.m[o + k] = mt;
    a.v[o + k] = vt;
    const double den = __dadd_rn(sqrt(__ddiv_rn((double)vt, 1.0 - bp2)), eps);
    const float step = (float)__dmul_rn(__ddiv_rn(__ddiv_rn((double)mt, 1.0 - bp1), den), lr);
    const float pnew = __fsub_rn(a.params[o + k], step);
    a.params[o + k] = pnew;
    if (a.image) {
      if (a.env_kind == CRL_ENV_CARTPOLE) image_scatter<CRL_ENV_CARTPOLE>(a.image, o + k, pnew);
      else image_scatter<CRL_ENV_PENDULUM>(a.image, o + k, pnew);
    }
  }
  if (threadIdx.x == 0 && crank == 0) {
    a.beta_pow[2 * i] = bp1 * b1;
    a.beta_pow[2 * i + 1] = bp2 * b2;
    if (i == 0 && !a.gf) {
      double tail[4 + CRL_MAX_WORLD];
      const int nt = 4 + (a.verify ? a.world : 0);
      for (int k = 0; k < nt; k++) tail[k] = finish_load(a, slot, L.P + k);
      if (a.stats_out) finalize_stats(tail, a.M_global * a.stat_ranks, a.A, a.ent_coeff, a.v_coef, a.stats_out, a.algo);
      if (a.verify) {
        float m = INFINITY;
        for (int r = 0; r < a.world; r++) m = fminf(m, (float)tail[4 + r]);
        const float s_f = (float)(tail[3] / a.M_global);
        a.fin->s_unclipped = s_f; a.fin->min_vlc = m; a.fin->M_global = a.M_global; a.fin->cnt = 0ull;
        a.fin->need_fixup = (s_f > m) ? 1 : 0;
        if (s_f > m) a.ds_rw->spec_failed = 1;  // the host replays this update exactly
      }
    }
  }
}

__global__ void advance_kernel(DevState* ds, unsigned long long dstep, unsigned long long dupd) {
  ds->policy_step += dstep;
  ds->update_index += dupd;
}

// all epoch permutations of one update, update_index read from the device (graph-capturable): out[epoch][B]
__global__ void fill_perms_dev_kernel(int32_t* out, uint32_t B, int half_bits, unsigned long long seed,
                                      const DevState* ds, uint32_t rank) {
  __shared__ uint32_t keys[8];
  const uint32_t epoch = blockIdx.y;
  if (threadIdx.x == 0) perm_keys(seed, ds->update_index, epoch, rank, keys);
  __syncthreads();
  int32_t* o = out + (size_t)epoch * B;
  // perm_index() cycle-walks (B is rarely a power of four), and a warp would wait for its unluckiest lane: ~6 Feistel
  // passes instead of the average 2. Here every lane keeps its own walk and moves on to its next index as soon as one
  // lands, so the lanes stay busy; the values are those of perm_index().
  const uint32_t mask = (1u << half_bits) - 1u, stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, x = i;
  while (i < B) {
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const uint32_t nl = r;
      r = l ^ feistel_round(r, keys[k], mask);
      l = nl;
    }
    x = (l << half_bits) | r;
    if (x < B) {
      o[i] = (int32_t)x;
      i += stride;
      x = i;
    }
  }
}

__global__ void fill_perm_kernel(int32_t* out, uint32_t B, int half_bits, unsigned long long seed,
                                 unsigned long long update_index, uint32_t epoch, uint32_t rank) {
  __shared__ uint32_t keys[8];
  if (threadIdx.x == 0) perm_keys(seed, update_index, epoch, rank, keys);
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x)
    out[i] = (int32_t)perm_index(i, B, half_bits, keys);
}

template <typename K> cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

int mb_stats_grid(int M, int sm_count) {
  const int tiles = (M + 255) / 256;
  return tiles < sm_count ? (tiles < 1 ? 1 : tiles) : sm_count;
}
int loss_grad_grid(int M, int sm_count) {
  const int tiles = (M + 127) / 128;
  return tiles < sm_count ? (tiles < 1 ? 1 : tiles) : sm_count;
}

cudaError_t kernels_init_update() {
  cudaError_t e = set_smem(mb_stats_kernel<CRL_ENV_CARTPOLE>, StatsSmem<CRL_ENV_CARTPOLE>::BYTES);
  if (e != cudaSuccess) return e;
  e = set_smem(mb_stats_kernel<CRL_ENV_PENDULUM>, StatsSmem<CRL_ENV_PENDULUM>::BYTES);
  if (e != cudaSuccess) return e;
  e = set_smem(loss_grad_kernel<CRL_ENV_CARTPOLE>, LossSmem<CRL_ENV_CARTPOLE>::BYTES);
  if (e != cudaSuccess) return e;
  e = set_smem(loss_grad_kernel<CRL_ENV_PENDULUM>, LossSmem<CRL_ENV_PENDULUM>::BYTES);
  if (e != cudaSuccess) return e;
  return kernels_init_update_tc();
}

cudaError_t launch_mb_stats(const UpdateArgs& a, int grid, cudaStream_t s) {
  if (a.env_kind == CRL_ENV_CARTPOLE)
    mb_stats_kernel<CRL_ENV_CARTPOLE><<<grid, CRL_THREADS, StatsSmem<CRL_ENV_CARTPOLE>::BYTES, s>>>(a);
  else
    mb_stats_kernel<CRL_ENV_PENDULUM><<<grid, CRL_THREADS, StatsSmem<CRL_ENV_PENDULUM>::BYTES, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_mb_count(const UpdateArgs& a, cudaStream_t s) {
  int grid = (a.M + 1023) / 1024;
  if (grid < 1) grid = 1;
  if (grid > 148) grid = 148;
  mb_count_kernel<0><<<grid, CRL_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_loss_grad(const UpdateArgs& a, cudaStream_t s) {
  if (a.tc_actor_ctas) return launch_loss_grad_tc(a, s);
  if (a.env_kind == CRL_ENV_CARTPOLE)
    loss_grad_kernel<CRL_ENV_CARTPOLE><<<a.grid_loss, CRL_THREADS, LossSmem<CRL_ENV_CARTPOLE>::BYTES, s>>>(a);
  else
    loss_grad_kernel<CRL_ENV_PENDULUM><<<a.grid_loss, CRL_THREADS, LossSmem<CRL_ENV_PENDULUM>::BYTES, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_grad_reduce(const UpdateArgs& a, int P, cudaStream_t s) {
  const int grid = (P + 4 + 63) / 64 + 1;  // + the verification block
  grad_reduce_kernel<<<grid, 256, 0, s>>>(a, P);
  return cudaGetLastError();
}

int param_image_floats(int env_kind) {
  return env_kind == CRL_ENV_CARTPOLE ? TcImage<CRL_ENV_CARTPOLE>::FLOATS : TcImage<CRL_ENV_PENDULUM>::FLOATS;
}
cudaError_t launch_param_image(int env_kind, const float* params, float* image, cudaStream_t s) {
  if (env_kind == CRL_ENV_CARTPOLE)
    param_image_kernel<CRL_ENV_CARTPOLE><<<(EnvTraits<CRL_ENV_CARTPOLE>::P + 255) / 256, 256, 0, s>>>(params, image);
  else
    param_image_kernel<CRL_ENV_PENDULUM><<<(EnvTraits<CRL_ENV_PENDULUM>::P + 255) / 256, 256, 0, s>>>(params, image);
  return cudaGetLastError();
}

cudaError_t launch_adv_stats(const AdvStatsArgs& a, cudaStream_t s) {
  if (a.n_sets < 1) return cudaSuccess;
  adv_stats_kernel<<<dim3(ADV_CHUNKS, a.n_sets), 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_clip_adam(const AdamArgs& a, cudaStream_t s) {
  Layout L;
  if (!make_layout(a.env_kind, &L)) return cudaErrorInvalidValue;
  clip_adam_kernel<<<L.n_arrays * ADAM_CL, ADAM_NT, 0, s>>>(a);  // __cluster_dims__(8): one cluster per parameter array
  return cudaGetLastError();
}

cudaError_t launch_loss_finalize(const double* gsum, int P, float* grads_out, double Mg, int A, float ent_coeff,
                                 float v_coef, double* stats_out, cudaStream_t s) {
  loss_finalize_kernel<<<(P + 255) / 256, 256, 0, s>>>(gsum, P, grads_out, Mg, A, ent_coeff, v_coef, stats_out);
  return cudaGetLastError();
}

cudaError_t launch_stats_pack(const MbScalars* parts, int n, MbScalars* out, cudaStream_t s) {
  stats_pack_kernel<<<1, CRL_THREADS, 0, s>>>(parts, n, out);
  return cudaGetLastError();
}

cudaError_t launch_advance(DevState* ds, unsigned long long d_policy_step, unsigned long long d_update, cudaStream_t s) {
  advance_kernel<<<1, 1, 0, s>>>(ds, d_policy_step, d_update);
  return cudaGetLastError();
}

cudaError_t launch_fill_perms_dev(int32_t* out, uint32_t B, unsigned long long seed, const DevState* ds, int n_epochs,
                                  uint32_t rank, cudaStream_t s) {
  unsigned bx = (B + 4095) / 4096;   // >= 16 indices per thread keeps the per-lane walks balanced
  if (bx > 148 * 2) bx = 148 * 2;
  if (bx < 1) bx = 1;
  fill_perms_dev_kernel<<<dim3(bx, (unsigned)n_epochs), 256, 0, s>>>(out, B, perm_half_bits(B), seed, ds, rank);
  return cudaGetLastError();
}

cudaError_t launch_fill_perm(int32_t* out, uint32_t B, unsigned long long seed, unsigned long long update_index,
                             uint32_t epoch, uint32_t rank, cudaStream_t s) {
  fill_perm_kernel<<<64, 256, 0, s>>>(out, B, perm_half_bits(B), seed, update_index, epoch, rank);
  return cudaGetLastError();
}
