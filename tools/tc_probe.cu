// tc_probe.cu — development probe for the tcgen05 path: validates the shared-memory descriptor
// conventions (no-swizzle canonical layouts, K-major and MN-major), the kind::tf32 instruction
// descriptor, TMEM allocation, commit/mbarrier and tcgen05.ld against a CPU reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// activation layout: element (s, f) of an [S=128][F=128] matrix, core matrix = 8 samples x 4 features (16 B rows)
//   byte address = (f%4)*4 + (s%8)*16 + (f/4)*XF + (s/8)*YS,   XF = 128, YS = (F/4)*128
// weight layout:   element (j, k) of W[64][64]: core matrix = 8 k x 4 j
//   byte address = (j%4)*4 + (k%8)*16 + (j/4)*XJ + (k/8)*YK,   XJ = 128, YK = 16*128
constexpr int S = 128, F = 128, XF = 128, YS = (F / 4) * 128;
constexpr int XJ = 128, YK = 16 * 128;
__host__ __device__ inline int act_off(int s, int f) { return ((f % 4) * 4 + (s % 8) * 16 + (f / 4) * XF + (s / 8) * YS) / 4; }
__host__ __device__ inline int w_off(int j, int k) { return ((j % 4) * 4 + (k % 8) * 16 + (j / 4) * XJ + (k / 8) * YK) / 4; }

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}
__device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 2u << 7;                    // a_format = TF32
  d |= 2u << 10;                   // b_format = TF32
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ inline void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0: D[s][j] = sum_k A(s,k) W(j,k)      A K-major (features 0..63), B = W as MN-major (N=j)
// mode 1: D[s][k] = sum_j A(s,j) W(j,k)      A K-major, B = W as K-major (N=k, K=j)
// mode 2: D[f][k] = sum_s A2(s,f) A(s,k)     A MN-major (M=f over all 128 features of act2), B MN-major (N=k<64 of act)
__global__ void probe_kernel(const float* act_g, const float* act2_g, const float* w_g, float* out, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* act = reinterpret_cast<float*>(smem_raw);             // 64 KB
  float* act2 = act + S * F;                                   // 64 KB
  float* w = act2 + S * F;                                     // 16 KB
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < S * F; i += blockDim.x) { act[i] = act_g[i]; act2[i] = act2_g[i]; }
  for (int i = tid; i < 64 * 64; i += blockDim.x) w[i] = w_g[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(act), a2 = smem_u32(act2), w0 = smem_u32(w);
    if (mode == 0) {
      const uint32_t idesc = make_idesc(128, 64, 0, 1);
      for (int ks = 0; ks < 8; ks++) {  // K = 64 in steps of 8
        const uint64_t ad = make_desc(a0 + ks * 2 * XF, XF, YS);   // K-major: LBO = stride between the two 4-element K chunks, SBO = 8-row group stride
        const uint64_t bd = make_desc(w0 + ks * YK, YK, XJ);       // MN-major: SBO = stride between 4-element N chunks, LBO = stride between K groups of 8
        mma_tf32(tmem, ad, bd, idesc, ks > 0);
      }
    } else if (mode == 1) {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      for (int js = 0; js < 8; js++) {
        const uint64_t ad = make_desc(a0 + js * 2 * XF, XF, YS);
        const uint64_t bd = make_desc(w0 + js * 2 * XJ, XJ, YK);   // K-major B: rows = k (8-row groups at YK), K chunks (j/4) at XJ
        mma_tf32(tmem, ad, bd, idesc, js > 0);
      }
    } else {
      const uint32_t idesc = make_idesc(128, 64, 1, 1);
      for (int ss = 0; ss < 16; ss++) {  // K = 128 samples in steps of 8
        const uint64_t ad = make_desc(a2 + ss * YS, YS, XF);       // MN-major A: M chunks (f/4) at XF, K groups (s/8) at YS
        const uint64_t bd = make_desc(a0 + ss * YS, YS, XF);       // MN-major B: N chunks (k/4) at XF
        mma_tf32(tmem, ad, bd, idesc, ss > 0);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
        "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]),
        "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr + 32));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int c = 0; c < 64; c++) out[tid * 64 + c] = __uint_as_float(r[c]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  std::vector<float> A(S * F), A2(S * F), W(64 * 64), act(S * F), act2(S * F), w(64 * 64);
  srand(1);
  auto rnd = []() { return tf32_trunc((float)rand() / RAND_MAX * 2.0f - 1.0f); };
  for (auto& x : A) x = rnd();
  for (auto& x : A2) x = rnd();
  for (auto& x : W) x = rnd();
  for (int s = 0; s < S; s++) for (int f = 0; f < F; f++) { act[act_off(s, f)] = A[s * F + f]; act2[act_off(s, f)] = A2[s * F + f]; }
  for (int j = 0; j < 64; j++) for (int k = 0; k < 64; k++) w[w_off(j, k)] = W[j * 64 + k];
  float *dact, *dact2, *dw, *dout;
  CK(cudaMalloc(&dact, S * F * 4)); CK(cudaMalloc(&dact2, S * F * 4)); CK(cudaMalloc(&dw, 64 * 64 * 4)); CK(cudaMalloc(&dout, 128 * 64 * 4));
  CK(cudaMemcpy(dact, act.data(), S * F * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dact2, act2.data(), S * F * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w.data(), 64 * 64 * 4, cudaMemcpyHostToDevice));
  const int smem = (2 * S * F + 64 * 64) * 4;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> out(128 * 64);
  int bad = 0;
  for (int mode = 0; mode < 3; mode++) {
    CK(cudaMemset(dout, 0, 128 * 64 * 4));
    probe_kernel<<<1, 128, smem>>>(dact, dact2, dw, dout, mode);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < 128; r++)
      for (int c = 0; c < 64; c++) {
        double ref = 0;
        if (mode == 0) for (int k = 0; k < 64; k++) ref += (double)A[r * F + k] * W[c * 64 + k];
        if (mode == 1) for (int j = 0; j < 64; j++) ref += (double)A[r * F + j] * W[j * 64 + c];
        if (mode == 2) for (int s = 0; s < S; s++) ref += (double)A2[s * F + r] * A[s * F + c];
        maxerr = fmax(maxerr, fabs(ref - out[r * 64 + c]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("mode %d: max |err| = %.3e (max |ref| = %.3f) %s\n", mode, maxerr, maxref, maxerr < 1e-4 ? "OK" : "MISMATCH");
    if (!(maxerr < 1e-4)) bad++;
  }
  printf(bad ? "tc_probe: FAILED\n" : "tc_probe: all modes OK\n");
  return bad;
}
