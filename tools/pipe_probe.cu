// pipe_probe.cu — development aid: per-SMSP issue cost (cycles per warp instruction) of the instruction kinds the
// rollout and update kernels are made of, measured with clock64 on one SM: packed and scalar FMA, MUFU.RCP, tanh_fast2,
// and the shared-memory load patterns of the register-tiled layers. Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/pipe_probe tools/pipe_probe.cu -Icleanrl.jl_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "device_math.cuh"

constexpr int IT = 2048;

template <int MODE> __global__ void probe(float* out, long long* cyc, int sel) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0f + 1e-3f * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = make_float2(1.0f + lane * 1e-3f + i, 0.5f + i);
  float2 w = make_float2(1.0001f, 0.9999f);
  float acc_s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc_s[i] = 1.0f + i + lane * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {          // 8 independent packed FMA chains
    for (int it = 0; it < IT; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], w, w);
    }
  } else if (MODE == 1) {   // 16 independent scalar FMA chains
    for (int it = 0; it < IT; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc_s[i] = fmaf(acc_s[i], w.x, w.y);
    }
  } else if (MODE == 2) {   // MUFU.RCP, 8 independent
    for (int it = 0; it < IT; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i].x));
    }
  } else if (MODE == 3) {   // tanh_fast2 on 8 pairs
    for (int it = 0; it < IT / 8; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = tanh_fast2(__fadd2_rn(a[i], w));
    }
  } else if (MODE == 4) {   // LDS.128, quarter-warp contiguous (8 x 16 B), other quarters broadcast the same
    const float4* p = reinterpret_cast<const float4*>(sm) + (lane & 7);
    for (int it = 0; it < IT; it++) {
      const float4 v = p[(it & 63) * 8];
      a[it & 7].x += v.x + v.w;
    }
  } else if (MODE == 5) {   // LDS.128, 32 distinct contiguous addresses (512 B)
    const float4* p = reinterpret_cast<const float4*>(sm) + lane;
    for (int it = 0; it < IT; it++) {
      const float4 v = p[(it & 15) * 32];
      a[it & 7].x += v.x + v.w;
    }
  } else if (MODE == 6) {   // LDS.128, one address for the whole warp
    const float4* p = reinterpret_cast<const float4*>(sm);
    for (int it = 0; it < IT; it++) {
      const float4 v = p[(it & 255)];
      a[it & 7].x += v.x + v.w;
    }
  } else if (MODE == 7) {   // the 4x4 tile inner step: 2 LDS.128 + 8 FFMA2 per k
    const float4* pw = reinterpret_cast<const float4*>(sm) + (lane & 7);
    const float4* pa = reinterpret_cast<const float4*>(sm) + 512 + (lane >> 3);
    for (int it = 0; it < IT; it++) {
      const float4 wv = pw[(it & 31) * 16], av = pa[(it & 31) * 9];
      const float2 a01 = make_float2(av.x, av.y), a23 = make_float2(av.z, av.w);
      a[0] = __ffma2_rn(make_float2(wv.x, wv.x), a01, a[0]); a[1] = __ffma2_rn(make_float2(wv.x, wv.x), a23, a[1]);
      a[2] = __ffma2_rn(make_float2(wv.y, wv.y), a01, a[2]); a[3] = __ffma2_rn(make_float2(wv.y, wv.y), a23, a[3]);
      a[4] = __ffma2_rn(make_float2(wv.z, wv.z), a01, a[4]); a[5] = __ffma2_rn(make_float2(wv.z, wv.z), a23, a[5]);
      a[6] = __ffma2_rn(make_float2(wv.w, wv.w), a01, a[6]); a[7] = __ffma2_rn(make_float2(wv.w, wv.w), a23, a[7]);
    }
  } else if (MODE == 8) {   // the same inner step with scalar FMAs (16 per k)
    const float4* pw = reinterpret_cast<const float4*>(sm) + (lane & 7);
    const float4* pa = reinterpret_cast<const float4*>(sm) + 512 + (lane >> 3);
    for (int it = 0; it < IT; it++) {
      const float4 wv = pw[(it & 31) * 16], av = pa[(it & 31) * 9];
      const float ww[4] = {wv.x, wv.y, wv.z, wv.w}, aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int s = 0; s < 4; s++) acc_s[j * 4 + s] = fmaf(ww[j], aa[s], acc_s[j * 4 + s]);
    }
  }
  double dacc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) dacc[i] = 1.0 + i + lane * 1e-3;
  if (MODE == 9) {          // 8 independent DFMA chains (throughput)
    for (int it = 0; it < IT; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dacc[i] = fma(dacc[i], 1.0000001, 1e-9);
    }
  } else if (MODE == 10) {  // one dependent DFMA chain (latency)
    for (int it = 0; it < IT; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dacc[0] = fma(dacc[0], 1.0000001, 1e-9);
    }
  } else if (MODE == 11) {  // dependent float -> double -> float round trips with a DADD in between
    float f = acc_s[0];
    for (int it = 0; it < IT; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) f = (float)((double)f + 1e-9);
    }
    acc_s[0] = f;
  } else if (MODE == 12) {  // Float64 division, dependent
    for (int it = 0; it < IT / 8; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dacc[0] = 1.0000001 / (dacc[0] + 1e-9);
    }
  }
  const long long t1 = clock64();
  float r = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; i++) r += (float)dacc[i];
#pragma unroll
  for (int i = 0; i < 8; i++) r += a[i].x + a[i].y;
#pragma unroll
  for (int i = 0; i < 16; i++) r += acc_s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[sel] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(float));
  cudaMallocManaged(&cyc, 64 * sizeof(long long));
  const char* names[13] = {"8 FFMA2 chains", "16 FFMA chains", "8 MUFU.RCP", "8 tanh_fast2 (per pair-call)", "LDS.128 8x16B + broadcast quarters",
                          "LDS.128 32 distinct", "LDS.128 one address", "4x4 tile k-step (2 LDS.128 + 8 FFMA2)", "4x4 tile k-step scalar (2 LDS.128 + 16 FFMA)",
                          "8 independent DFMA chains", "1 dependent DFMA chain", "f32->f64, DADD, f64->f32 (dependent)", "Float64 division (dependent)"};
  const int per_it[13] = {8, 16, 8, 1, 1, 1, 1, 1, 1, 8, 8, 8, 1};
  const int its[13] = {IT, IT, IT, IT, IT, IT, IT, IT, IT, IT, IT, IT, IT};
  for (int threads : {32, 128, 256, 512}) {
    printf("threads per SM = %d (%d warp(s) per scheduler)\n", threads, threads / 128 ? threads / 128 : 1);
    for (int m = 0; m < 13; m++) {
      switch (m) {
        case 0: probe<0><<<1, threads>>>(out, cyc, m); break;
        case 1: probe<1><<<1, threads>>>(out, cyc, m); break;
        case 2: probe<2><<<1, threads>>>(out, cyc, m); break;
        case 3: probe<3><<<1, threads>>>(out, cyc, m); break;
        case 4: probe<4><<<1, threads>>>(out, cyc, m); break;
        case 5: probe<5><<<1, threads>>>(out, cyc, m); break;
        case 6: probe<6><<<1, threads>>>(out, cyc, m); break;
        case 7: probe<7><<<1, threads>>>(out, cyc, m); break;
        case 8: probe<8><<<1, threads>>>(out, cyc, m); break;
        case 9: probe<9><<<1, threads>>>(out, cyc, m); break;
        case 10: probe<10><<<1, threads>>>(out, cyc, m); break;
        case 11: probe<11><<<1, threads>>>(out, cyc, m); break;
        case 12: probe<12><<<1, threads>>>(out, cyc, m); break;
      }
      cudaDeviceSynchronize();
      printf("  %-46s %8.2f cycles per warp-level op (warp 0's clock)\n", names[m], (double)cyc[m] / ((double)its[m] * per_it[m]));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
