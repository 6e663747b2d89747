// tf32_peak.cu — measured issue-bound peak of tcgen05.mma kind::tf32 on this GPU: the denominator of the roofline of
// loss_grad_tc_kernel (csrc/update_tc.cu). MEASURED_PEAKS.json carries a bf16 figure only.
//
// One CTA per SM. Operands sit in shared memory in the canonical K-major no-swizzle layout the update kernel uses
// (8-row x 16-byte core matrices, LBO = 128 B between K chunks, SBO = 2048 B between 8-row groups); one thread issues
// M128 x N{256,128,64} x K8 MMAs back to back into two alternating TMEM accumulators and commits once at the end, so
// nothing but the tensor pipe's own rate limits the loop. FLOP = 2 M N K per MMA.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tf32_peak tools/tf32_peak.cu
// Run:   tools/tf32_peak [seconds_sustained]      prints one JSON object
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int LBO = 128, SBO = 2048;          // bytes
constexpr int A_BYTES = 16 * SBO, B_BYTES = 32 * SBO;   // 128 rows, 256 rows; K = 64

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ inline uint32_t make_idesc(int M, int N) {  // kind::tf32, fp32 accumulate, both operands K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ inline void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = tid; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) f[i] = 1e-3f * (float)((i * 37) & 63);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + A_BYTES);
    for (int it = 0; it < iters; it++) {
      const uint32_t d = tmem + (uint32_t)((it & 1) * 256);
#pragma unroll
      for (int ks = 0; ks < 8; ks++) mma_ss(d, make_desc(a0 + ks * 2 * LBO), make_desc(b0 + ks * 2 * LBO), idesc, (it > 1 || ks > 0) ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0 && sink) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tmem));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    sink[blockIdx.x] = __uint_as_float(r);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N> static void measure(int sms, double sustain_s, double* burst_tf, double* sustained_tf) {
  const int smem = A_BYTES + B_BYTES;
  CK(cudaFuncSetAttribute(peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4096;  // x 8 MMAs each
  const double flop = 2.0 * 128 * N * 8 * 8.0 * iters * sms;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) peak_kernel<N><<<sms, 128, smem>>>(iters, nullptr);
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int rep = 0; rep < 10; rep++) {
    CK(cudaEventRecord(e0));
    peak_kernel<N><<<sms, 128, smem>>>(iters, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, (double)ms);
  }
  *burst_tf = flop / (best * 1e-3) / 1e12;
  // sustained: back-to-back launches for sustain_s seconds
  const int n_launch = std::max(8, (int)(sustain_s * 1e3 / best));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < n_launch; i++) peak_kernel<N><<<sms, 128, smem>>>(iters, nullptr);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  *sustained_tf = flop * n_launch / (ms * 1e-3) / 1e12;
  CK(cudaGetLastError());
}

int main(int argc, char** argv) {
  const double sustain_s = argc > 1 ? atof(argv[1]) : 2.0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double b256, s256, b128, s128, b64, s64;
  measure<256>(sms, sustain_s, &b256, &s256);
  measure<128>(sms, sustain_s / 4, &b128, &s128);
  measure<64>(sms, sustain_s / 4, &b64, &s64);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"kind\": \"tcgen05.mma.cta_group::1.kind::tf32 M128 K8, SS operands (K-major, no swizzle)\", "
         "\"tf32_tflops\": %.1f, \"tf32_tflops_sustained\": %.1f, \"n256\": {\"burst\": %.1f, \"sustained\": %.1f}, "
         "\"n128\": {\"burst\": %.1f, \"sustained\": %.1f}, \"n64\": {\"burst\": %.1f, \"sustained\": %.1f}, "
         "\"how\": \"one CTA per SM, one thread issues 32768 MMAs back to back into two alternating TMEM accumulators; burst = best of "
         "10 launches, sustained = back-to-back launches for %.1f s; CUDA events\"}\n",
         prop.name, sms, b256, s256, b256, s256, b128, s128, b64, s64, sustain_s);
  return 0;
}
