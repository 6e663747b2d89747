// tc_probe2.cu — layout discovery for MN-major no-swizzle tcgen05 operands: A is a known-good K-major one-hot
// matrix, B's shared memory is filled with its own 16-byte-chunk index (and, in a second run, the element index
// inside the chunk), so D[k][n] reveals which address the tensor core reads for logical B(n, k).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int XF = 128, YS = 32 * 128;
__host__ __device__ inline int act_off(int s, int f) { return ((f % 4) * 4 + (s % 8) * 16 + (f / 4) * XF + (s / 8) * YS) / 4; }
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void probe(const float* act_g, const float* b_g, float* out, uint32_t lbo, uint32_t sbo, int b_mn, int a_probe) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* act = reinterpret_cast<float*>(smem_raw);   // 64 KB one-hot A (K-major) or probe target
  float* b = act + 128 * 128;                        // 64 KB
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 128; i += blockDim.x) { act[i] = act_g[i]; b[i] = b_g[i]; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    uint64_t ad, bd; uint32_t idesc;
    if (!a_probe) {   // A = one-hot K-major (known good), B = probed
      idesc = make_idesc(128, 64, 0, b_mn);
      ad = make_desc(smem_u32(act), XF, YS);
      bd = make_desc(smem_u32(b), lbo, sbo);
    } else {          // A = probed MN-major (M=128), B = one-hot K-major: D[m][n] = sum_k A(m,k) B(n,k), B(n,k) = [k==n] (n<8)
      idesc = make_idesc(128, 64, 1, 0);
      ad = make_desc(smem_u32(b), lbo, sbo);
      bd = make_desc(smem_u32(act), XF, YS);
    }
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int c = 0; c < 8; c++) out[tid * 64 + c0 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}
int main() {
  const int NE = 128 * 128;
  std::vector<float> onehot(NE, 0.f), chunk(NE), within(NE), out(128 * 64);
  // K-major one-hot: row s has a 1 at k = s (s < 8): X(s,k) = [k == s]
  for (int s = 0; s < 8; s++) onehot[act_off(s, s)] = 1.0f;
  for (int i = 0; i < NE; i++) { chunk[i] = (float)(i / 4); within[i] = (float)(i % 4); }
  float *d1, *dc, *dw, *dout;
  CK(cudaMalloc(&d1, NE * 4)); CK(cudaMalloc(&dc, NE * 4)); CK(cudaMalloc(&dw, NE * 4)); CK(cudaMalloc(&dout, 128 * 64 * 4));
  CK(cudaMemcpy(d1, onehot.data(), NE * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dc, chunk.data(), NE * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, within.data(), NE * 4, cudaMemcpyHostToDevice));
  const int smem = 2 * NE * 4;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Cfg { uint32_t lbo, sbo; int b_mn, a_probe; const char* name; } cfgs[] = {
      {128, 2048, 0, 0, "B K-major control lbo=128 sbo=2048"}, {2048, 128, 1, 0, "B MN-major lbo=2048 sbo=128"}, {128, 2048, 1, 0, "B MN-major lbo=128 sbo=2048"},
      {512, 1024, 1, 0, "B MN-major lbo=512 sbo=1024"}, {2048, 128, 1, 1, "A MN-major lbo=2048 sbo=128"},
      {128, 2048, 1, 1, "A MN-major lbo=128 sbo=2048"}, {512, 1024, 1, 1, "A MN-major lbo=512 sbo=1024"}};
  for (auto& c : cfgs) {
    std::vector<float> oc(128 * 64), ow(128 * 64);
    for (int pass = 0; pass < 2; pass++) {
      CK(cudaMemset(dout, 0, 128 * 64 * 4));
      probe<<<1, 128, smem>>>(d1, pass == 0 ? dc : dw, dout, c.lbo, c.sbo, c.b_mn, c.a_probe);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy((pass == 0 ? oc : ow).data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    }
    printf("== %s : byte address read for logical (mn, k) ==  raw oc[0..3]=%g %g %g %g\n", c.name, oc[0], oc[1], oc[2], oc[3]);
    if (!c.a_probe) {  // D[k][n] = B(n,k)
      for (int k = 0; k < 8; k++) { printf(" k=%d:", k); for (int n : {0, 1, 2, 3, 4, 5, 8, 12, 16, 32, 63}) printf(" n%d@%d", n, (int)(oc[k * 64 + n] * 16 + ow[k * 64 + n] * 4)); printf("\n"); }
    } else {           // D[m][n] = A(m, k=n)
      for (int k = 0; k < 8; k++) { printf(" k=%d:", k); for (int m : {0, 1, 2, 3, 4, 5, 8, 12, 16, 32, 64, 127}) printf(" m%d@%d", m, (int)(oc[m * 64 + k] * 16 + ow[m * 64 + k] * 4)); printf("\n"); }
    }
  }
  return 0;
}
