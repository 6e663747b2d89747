// update_common.cuh — helpers shared by the FFMA update kernels (update.cu) and the tcgen05 update kernel
// (update_tc.cu): minibatch index source, the value-loss pieces, block reductions and the geometry of the
// tensor-core weight images that ride behind the FFMA parameter image.
#pragma once
#include "kernels.h"
#include "mlp_tile.cuh"

namespace crl_upd {

__device__ __forceinline__ int sample_index(const IdxSrc& ix, const uint32_t* keys, int m) {
  if (ix.arr) return ix.arr[m];
  return (int)perm_index(ix.start + (uint32_t)m, ix.B, ix.half_bits, keys);
}

// value-loss pieces of ppo.jl:234-235, evaluated identically in every kernel that needs them
__device__ __forceinline__ void value_clip(float v, float V, float R, float c, float& vc_minus_R, float& vlc,
                                           bool& inside) {
  const float dv = __fsub_rn(v, V);
  const float cl = dv < -c ? -c : (dv > c ? c : dv);
  const float vc = __fadd_rn(V, cl);
  vc_minus_R = __fsub_rn(vc, R);
  vlc = __fmul_rn(vc_minus_R, vc_minus_R);
  inside = dv >= -c && dv <= c;
}

template <int NW> __device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NW; w++) s += red[w];
  return s;
}

// ---- tensor-core operand images (kind::tf32, K-major, no swizzle) ------------------------------------------
// A 64 x 64 weight operand with rows n and contraction index k is stored as 8 x 4 core matrices (8 rows of
// 16 bytes): float offset = (n%8)*4 + (n/8)*512 + (k/4)*32 + (k%4), i.e. LBO = 128 B between the two K chunks
// of one MMA (K = 8), SBO = 2048 B between 8-row groups (validated by tools/tc_probe*.cu).
constexpr int TC_W_FLOATS = CRL_H * CRL_H;
constexpr int TC_W_LBO = 128, TC_W_SBO = 2048;
__host__ __device__ constexpr int tc_w_off(int n, int k) { return (n % 8) * 4 + (n / 8) * 512 + (k / 4) * 32 + (k % 4); }

// 3xTF32 split: hi = x rounded to TF32 (10 explicit mantissa bits), lo = x - hi exactly (the tensor core reads
// the upper 19 bits of lo). hi*hi + lo*hi + hi*lo recovers the fp32 product to ~2^-21 relative.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// Per net, behind the FFMA image: [0] W2 as (rows j, k contiguous) hi, [1] same lo  — B operand of z2 = h1 W2^T
//                                 [2] W2 as (rows k, j contiguous) hi, [3] same lo  — B operand of dh1 = dz2 W2
template <int ENV> struct TcImage {
  static constexpr int BASE = SmemParams<ENV>::SIZE + 2 * CRL_H * CRL_H;
  static constexpr int NET_FLOATS = 4 * TC_W_FLOATS;
  static constexpr int FLOATS = BASE + 2 * NET_FLOATS;
};
// j = output neuron, k = input neuron of W2 (Flux stores it at j + 64 k)
template <int ENV> __device__ __forceinline__ void tc_image_scatter(float* image, int net, int j, int k, float v) {
  float* b = image + TcImage<ENV>::BASE + net * TcImage<ENV>::NET_FLOATS;
  const float hi = tf32_hi(v), lo = v - hi;
  b[0 * TC_W_FLOATS + tc_w_off(j, k)] = hi;
  b[1 * TC_W_FLOATS + tc_w_off(j, k)] = lo;
  b[2 * TC_W_FLOATS + tc_w_off(k, j)] = hi;
  b[3 * TC_W_FLOATS + tc_w_off(k, j)] = lo;
}

// flat parameter index -> position(s) in the parameter image
template <int ENV> __device__ __forceinline__ void image_scatter(float* image, int flat, float v) {
  using E = EnvTraits<ENV>;
  using SPm = SmemParams<ENV>;
  using NO = NetOff<E::D, 1>;
  int net, within;
  if (flat < E::NET_A) { net = 0; within = flat; image[SPm::ACTOR + within] = v; }
  else if (flat < E::NET_A + E::NET_C) { net = 1; within = flat - E::NET_A; image[SPm::CRITIC + within] = v; }
  else { image[SPm::LOGSTD + (flat - E::NET_A - E::NET_C)] = v; return; }
  if (within >= NO::W2 && within < NO::W2 + CRL_H * CRL_H) {
    const int e = within - NO::W2, j = e % CRL_H, k = e / CRL_H;  // Flux W2 is (out=j, in=k) at j + 64 k
    image[SPm::SIZE + net * CRL_H * CRL_H + j * CRL_H + k] = v;
    tc_image_scatter<ENV>(image, net, j, k, v);  // hi/lo tensor-core operand images (update_tc.cu)
  }
}

// loss scalars from the reduced sums (ppo.jl:228,237,242,243)
__device__ __forceinline__ void finalize_stats(const double* sums, double Mg, int A, float ent_coeff, float v_coef,
                                               double* out, int algo = 0) {
  if (algo == 1) {  // A2C: @info "Training Statistics" actor_loss critic_loss (a2c.jl:100)
    out[1] = sums[0] / Mg;
    out[2] = sums[1] / Mg;
    out[3] = 0.0;
    out[0] = out[1] + out[2];
    return;
  }
  const double pg = sums[0] / Mg;                                   // ppo.jl:228
  const double vl = 0.5 * (double)(float)(sums[1] / Mg);            // ppo.jl:237
  const double en = (double)(float)(sums[2] / ((double)A * Mg));    // ppo.jl:242
  out[0] = pg - (double)__fmul_rn(ent_coeff, (float)en) + (double)v_coef * vl;  // ppo.jl:243
  out[1] = pg;
  out[2] = vl;
  out[3] = en;
}

}  // namespace crl_upd
