// mlp_tile.cuh — register-tiled FFMA layers over a tile of samples held in shared memory.
//
// A CTA of 256 threads owns a tile of S samples. Activations live in shared memory
// feature-major: act[row][s], row = net*64 + neuron, row stride S_PAD = S + 4 floats (the +4
// makes row-strided float4 accesses hit distinct banks). Each thread owns a TS x TS register
// tile (TS neurons x TS samples). Lane mapping inside a warp: q_lo = lane % 8 picks the neuron
// chunk, p_lo = lane / 8 picks the sample chunk, so a warp-wide float4 weight load touches one
// contiguous 128 B line (8 distinct q_lo, broadcast over p_lo) and a float4 activation load
// touches 64 B (4 distinct p_lo, broadcast over q_lo): one shared-memory wavefront each.
//
// Weight matrices are used exactly as Flux stores them: W (out,in) column-major = k-major
// [in][out] (see common.cuh), so the forward layer needs no transposition.
#pragma once
#include "common.cuh"
#include "device_math.cuh"

#define CRL_THREADS 256

// Geometry of one tile configuration.
//   TS   thread tile edge (4 or 8)
//   NETS 2 = actor+critic side by side (rows 0..127), 1 = a single net (rows 0..63)
template <int TS_, int NETS_> struct TileGeom {
  static constexpr int TS = TS_, NETS = NETS_;
  static constexpr int QN = CRL_H / TS;        // neuron groups per net
  static constexpr int QG = NETS * QN;         // neuron groups total
  static constexpr int PG = CRL_THREADS / QG;  // sample groups
  static constexpr int S = PG * TS;            // samples per tile
  static constexpr int S_PAD = S + 4;
  static constexpr int NC = TS / 4;            // float4 chunks per thread edge
  static constexpr int ROWS = NETS * CRL_H;
  static constexpr int PH = PG / 4;            // warps along the sample axis
  static_assert(QG % 8 == 0 && PG % 4 == 0, "warp = 4 sample groups x 8 neuron groups");
  static_assert((QG / 8) * PH == CRL_THREADS / 32, "warp grid must cover the CTA");
};

template <class G> struct ThreadCoord {
  int p, q, net;
  int jb[G::NC];  // neuron base of each float4 chunk (within the net)
  int sb[G::NC];  // sample base of each float4 chunk
  __device__ ThreadCoord() {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q_lo = lane & 7, p_lo = lane >> 3;
    const int q_hi = w / G::PH, p_hi = w % G::PH;
    p = p_hi * 4 + p_lo;
    q = q_hi * 8 + q_lo;
    net = q / G::QN;
    const int qin = q % G::QN;
    if (G::TS == 8) {
#pragma unroll
      for (int c = 0; c < G::NC; c++) { jb[c] = 4 * qin + 32 * c; sb[c] = 4 * p + (G::S / 2) * c; }
    } else {
      jb[0] = 4 * qin;
      sb[0] = 4 * p;
    }
  }
};

__device__ __forceinline__ float f4_get(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// Column swizzle of the activation tiles: physical column = s ^ act_swz(row). With the 132-float row stride
// the 8 lanes of a quarter-warp that store rows 4*q_lo + j land on only two bank groups (4-way conflict on every
// float4 store); XOR-ing the column with 4*((row >> 3) & 3) spreads them over all eight. The XOR is a multiple of
// 4 below 16, so float4 accesses stay aligned and inside their 16-float group.
__device__ __forceinline__ int act_swz(int row) { return ((row >> 3) & 3) << 2; }

#define EPI_BIAS_TANH 0   // out = tanh_fast(acc + bias)                         (forward)
#define EPI_DTANH 1       // out = acc * (1 - out_old^2), in place over h        (backward)

// One dense layer for this thread's tile:  acc[j][s] = sum_k Wt[k][j] * in[k][s].
//   Wt   shared, [K][64] for this thread's net (k-major)
//   in   shared, row 0 of this net's input block (row stride S_PAD)
//   out  shared, row 0 of this net's output block
template <class G, int K, int EPI, bool SWZ_IN = false, bool SWZ_OUT = false>
__device__ __forceinline__ void tile_layer(const ThreadCoord<G>& tc, const float* __restrict__ Wt,
                                           const float* __restrict__ bias, const float* __restrict__ in,
                                           float* __restrict__ out) {
  constexpr int TS = G::TS, NC = G::NC, SP = G::S_PAD;
  // accumulators are packed pairs along the sample axis so the inner product runs on Blackwell's packed
  // FP32 pipe (fma.rn.f32x2 / FFMA2: two FMAs per lane per issued instruction)
  float2 acc2[TS][TS / 2];
#pragma unroll
  for (int j = 0; j < TS; j++)
#pragma unroll
    for (int s = 0; s < TS / 2; s++) acc2[j][s] = make_float2(0.0f, 0.0f);
#pragma unroll 4
  for (int k = 0; k < K; k++) {
    float4 w[NC], a[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      w[c] = *reinterpret_cast<const float4*>(Wt + k * CRL_H + tc.jb[c]);
      a[c] = *reinterpret_cast<const float4*>(in + k * SP + (SWZ_IN ? (tc.sb[c] ^ act_swz(k)) : tc.sb[c]));
    }
#pragma unroll
    for (int j = 0; j < TS; j++) {
      const float wj = f4_get(w[j / 4], j % 4);
      const float2 ww = make_float2(wj, wj);
#pragma unroll
      for (int c = 0; c < NC; c++) {
        acc2[j][2 * c + 0] = __ffma2_rn(ww, make_float2(a[c].x, a[c].y), acc2[j][2 * c + 0]);
        acc2[j][2 * c + 1] = __ffma2_rn(ww, make_float2(a[c].z, a[c].w), acc2[j][2 * c + 1]);
      }
    }
  }
  float acc[TS][TS];
#pragma unroll
  for (int j = 0; j < TS; j++)
#pragma unroll
    for (int s = 0; s < TS / 2; s++) { acc[j][2 * s] = acc2[j][s].x; acc[j][2 * s + 1] = acc2[j][s].y; }
#pragma unroll
  for (int j = 0; j < TS; j++) {
    const int row = tc.jb[j / 4] + (j % 4);
    float b = 0.0f;
    if (EPI == EPI_BIAS_TANH) b = bias[row];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      float4* dst = reinterpret_cast<float4*>(out + row * SP + (SWZ_OUT ? (tc.sb[c] ^ act_swz(row)) : tc.sb[c]));
      float4 o;
      if (EPI == EPI_BIAS_TANH) {
        // packed tanh_fast2: bit-identical to the scalar tanh_fast, half the issue slots
        const float2 t0 = tanh_fast2(__fadd2_rn(acc2[j][2 * c + 0], make_float2(b, b)));
        const float2 t1 = tanh_fast2(__fadd2_rn(acc2[j][2 * c + 1], make_float2(b, b)));
        o.x = t0.x; o.y = t0.y; o.z = t1.x; o.w = t1.y;
      } else {
        const float4 h = *dst;
        o.x = acc[j][4 * c + 0] * (1.0f - h.x * h.x);
        o.y = acc[j][4 * c + 1] * (1.0f - h.y * h.y);
        o.z = acc[j][4 * c + 2] * (1.0f - h.z * h.z);
        o.w = acc[j][4 * c + 3] * (1.0f - h.w * h.w);
      }
      *dst = o;
    }
  }
}

// The same layer for the rollout's 4 x 4 tiles (EPI_BIAS_TANH), scheduled by hand: the k loop is fully unrolled with the
// operands of the NEXT four k fetched while the current four are multiplied (the rolled loop stalled on its own loads
// at the top of every iteration: two warps per scheduler cannot cover them), and the epilogue reads its biases before
// and stores its outputs after ALL the tanh evaluations (a store between two of them stops the compiler from moving
// the next bias load up, which serialised the four rows). Every accumulator is still one fma chain over ascending k.
template <class G, int K>
__device__ __forceinline__ void tile_layer_fwd4(const ThreadCoord<G>& tc, const float* __restrict__ Wt,
                                                const float* __restrict__ bias, const float* __restrict__ in,
                                                float* __restrict__ out) {
  static_assert(G::TS == 4, "4 x 4 thread tiles");
  constexpr int SP = G::S_PAD, PF = 4;
  const float* wp = Wt + tc.jb[0];
  const float* ap = in + tc.sb[0];
  const float4 bv = *reinterpret_cast<const float4*>(bias + tc.jb[0]);
  float2 acc2[4][2];
#pragma unroll
  for (int j = 0; j < 4; j++) acc2[j][0] = acc2[j][1] = make_float2(0.0f, 0.0f);
  float4 w[PF], a[PF];
#pragma unroll
  for (int i = 0; i < PF; i++) {
    if (i < K) {
      w[i] = *reinterpret_cast<const float4*>(wp + i * CRL_H);
      a[i] = *reinterpret_cast<const float4*>(ap + i * SP);
    }
  }
#pragma unroll
  for (int k0 = 0; k0 < K; k0 += PF) {
    float4 wn[PF], an[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) {
      if (k0 + PF + i < K) {
        wn[i] = *reinterpret_cast<const float4*>(wp + (k0 + PF + i) * CRL_H);
        an[i] = *reinterpret_cast<const float4*>(ap + (k0 + PF + i) * SP);
      }
    }
#pragma unroll
    for (int i = 0; i < PF; i++) {
      if (k0 + i < K) {
        const float2 a01 = make_float2(a[i].x, a[i].y), a23 = make_float2(a[i].z, a[i].w);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float wj = f4_get(w[i], j);
          const float2 ww = make_float2(wj, wj);
          acc2[j][0] = __ffma2_rn(ww, a01, acc2[j][0]);
          acc2[j][1] = __ffma2_rn(ww, a23, acc2[j][1]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < PF; i++) { w[i] = wn[i]; a[i] = an[i]; }
  }
  float4 o[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float b = f4_get(bv, j);
    const float2 t0 = tanh_fast2(__fadd2_rn(acc2[j][0], make_float2(b, b)));
    const float2 t1 = tanh_fast2(__fadd2_rn(acc2[j][1], make_float2(b, b)));
    o[j] = make_float4(t0.x, t0.y, t1.x, t1.y);
  }
#pragma unroll
  for (int j = 0; j < 4; j++) *reinterpret_cast<float4*>(out + (tc.jb[0] + j) * SP + tc.sb[0]) = o[j];
}

// Shared-memory image of the parameters: each net copied to a 16-byte aligned base so the
// float4 weight loads are legal (the flat vector puts the critic at an odd offset).
template <int ENV> struct SmemParams {
  using E = EnvTraits<ENV>;
  static constexpr int ACTOR = 0;
  static constexpr int CRITIC = (E::NET_A + 3) & ~3;
  static constexpr int LOGSTD = CRITIC + ((E::NET_C + 3) & ~3);
  static constexpr int SIZE = LOGSTD + 4;
};

template <int ENV>
__device__ __forceinline__ void load_params(const float* __restrict__ g, float* __restrict__ sp) {
  using E = EnvTraits<ENV>;
  using SP = SmemParams<ENV>;
  for (int i = threadIdx.x; i < E::NET_A; i += blockDim.x) sp[SP::ACTOR + i] = g[i];
  for (int i = threadIdx.x; i < E::NET_C; i += blockDim.x) sp[SP::CRITIC + i] = g[E::NET_A + i];
  if (E::CONT && threadIdx.x < E::A) sp[SP::LOGSTD + threadIdx.x] = g[E::NET_A + E::NET_C + threadIdx.x];
}

// base of net `net` inside the smem parameter image, and its head width
template <int ENV> __device__ __forceinline__ int net_base(int net) {
  return net == 0 ? SmemParams<ENV>::ACTOR : SmemParams<ENV>::CRITIC;
}
