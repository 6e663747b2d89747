// tc_probe4.cu — layout discovery for MN-major SWIZZLE_128B_BASE32B tcgen05 operands (the only MN-major form kind::tf32 has): A is a known-good K-major one-hot
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
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
  uint64_t d = (uint64_t)layout << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void probe(const float* act_g, const float* b_g, float* out, uint32_t lbo, uint32_t sbo, int b_mn, int a_probe, uint32_t layout) {
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
      bd = make_desc(smem_u32(b), lbo, sbo, layout);
    } else {          // A = probed MN-major (M=128), B = one-hot K-major: D[m][n] = sum_k A(m,k) B(n,k), B(n,k) = [k==n] (n<8)
      idesc = make_idesc(128, 64, 1, 0);
      ad = make_desc(smem_u32(b), lbo, sbo, layout);
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
// hypothesis (CUTLASS Layout_MN_SW128_32B_Atom): atoms of 32 MN x 4 K fp32 (4 rows of 128 B), the 32-byte chunk index
// XOR-ed with the row; LBO = stride between MN atoms, SBO = stride between K atoms
static int hyp(int mn, int k, int lbo, int sbo) {
  const int r = k % 4, c = (mn % 32) / 8;
  return (mn / 32) * lbo + (k / 4) * sbo + r * 128 + ((c ^ r) * 32) + (mn % 8) * 4;
}
int main() {
  const int NE = 128 * 128;
  std::vector<float> onehot(NE, 0.f), f0(NE), f1(NE), f2(NE);
  for (int s = 0; s < 8; s++) onehot[act_off(s, s)] = 1.0f;
  for (int i = 0; i < NE; i++) { f0[i] = (float)(i % 4); f1[i] = (float)((i / 4) % 64); f2[i] = (float)(i / 256); }   // byte = 4 f0 + 16 f1 + 1024 f2
  float *d1, *dv[3], *dout;
  CK(cudaMalloc(&d1, NE * 4)); CK(cudaMalloc(&dout, 128 * 64 * 4));
  CK(cudaMemcpy(d1, onehot.data(), NE * 4, cudaMemcpyHostToDevice));
  const std::vector<float>* src[3] = {&f0, &f1, &f2};
  for (int p = 0; p < 3; p++) { CK(cudaMalloc(&dv[p], NE * 4)); CK(cudaMemcpy(dv[p], src[p]->data(), NE * 4, cudaMemcpyHostToDevice)); }
  const int smem = 2 * NE * 4;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Cfg { uint32_t lbo, sbo; int a_probe; uint32_t layout; } cfgs[] = {
      {1024, 512, 0, 1}, {512, 1024, 0, 1}, {16384, 512, 0, 1}, {1024, 512, 1, 1}, {4096, 512, 1, 1}, {16384, 512, 1, 1}, {512, 2048, 1, 1},
      {8192, 640, 0, 1}, {8192, 768, 0, 1}, {8192, 1536, 0, 1}};   // (an SBO of 528 faults: misaligned address)
  for (auto& c : cfgs) {
    std::vector<float> o[3];
    for (int pass = 0; pass < 3; pass++) {
      o[pass].resize(128 * 64);
      CK(cudaMemset(dout, 0, 128 * 64 * 4));
      probe<<<1, 128, smem>>>(d1, dv[pass], dout, c.lbo, c.sbo, 1, c.a_probe, c.layout);
      CK(cudaGetLastError());
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("lbo=%u sbo=%u: %s\n", c.lbo, c.sbo, cudaGetErrorString(e)); return 1; }   // sticky: stop
      CK(cudaMemcpy(o[pass].data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    }
    const int MN = c.a_probe ? 128 : 64;
    int bad = 0;
    printf("== %s MN-major layout_type=%u lbo=%u sbo=%u\n", c.a_probe ? "A" : "B", c.layout, c.lbo, c.sbo);
    for (int k = 0; k < 8; k++) {
      printf(" k=%d:", k);
      for (int mn = 0; mn < MN; mn++) {
        const int idx = c.a_probe ? mn * 64 + k : k * 64 + mn;   // D[m][n]: A probe: (m = mn, n = k); B probe: (m = k, n = mn)
        const int addr = (int)(o[0][idx] * 4 + o[1][idx] * 16 + o[2][idx] * 1024);
        if (addr != hyp(mn, k, c.lbo, c.sbo)) bad++;
        if (mn == 0 || mn == 8 || mn == 16 || mn == 24 || mn == 32 || mn == 63 || mn == 64 || mn == 96 || mn == 127) printf(" %d@%d", mn, addr);
      }
      printf("\n");
    }
    printf(" mismatches against the hypothesis: %d of %d\n", bad, 8 * MN);
  }
  return 0;
}
