// tc_probe3.cu — probes for the tcgen05 update kernel (csrc/update_tc.cu):
//   T1  A operand from TMEM (written with tcgen05.st by the lane that owns the sample), B = K-major weight image,
//       3xTF32 with masked / unmasked hi parts and plain 1xTF32 for comparison
//   T2  contraction over samples: A = [X_hi ; X_lo] stacked along M (K-major, no swizzle, LBO = 144 B), B = Y_hi
//       then Y_lo (N = 64): rows r and r+64 of D add up to the full 4-term product
//   T3  A with SBO = 0: the sixteen 8-row groups alias the same 8 rows
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe3 tools/tc_probe3.cu ; run on a B200.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int LBO_W = 128, SBO_W = 16 * 128;   // weight image: rows n, K = 64: (n%8)*16 + (n/8)*SBO_W + (k/4)*LBO_W + (k%4)*4
constexpr int LBO_F = 144, SBO_F = 32 * 144;   // feature-major operand: rows r, K = 128 samples, padded chunk stride
__host__ __device__ inline int w_off(int n, int k) { return ((n % 8) * 16 + (n / 8) * SBO_W + (k / 4) * LBO_W + (k % 4) * 4) / 4; }
__host__ __device__ inline int f_off(int r, int s) { return ((r % 8) * 16 + (r / 8) * SBO_F + (s / 4) * LBO_F + (s % 4) * 4) / 4; }
constexpr int W_FLOATS = 64 * 64, F64_FLOATS = 8 * SBO_F / 4, F128_FLOATS = 16 * SBO_F / 4;

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ inline void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ inline void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ inline void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
               "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
               "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
__device__ inline void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ inline float hi_of(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// mode 0: T1 3xTF32 masked hi   1: T1 3xTF32 raw value as hi (lo computed against the masked hi)   2: T1 1xTF32
// mode 3: T2                    4: T3
__global__ void probe(const float* a_g, const float* whi_g, const float* wlo_g, const float* x_g, const float* yhi_g,
                      const float* ylo_g, const float* alias_g, float* out, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* whi = reinterpret_cast<float*>(smem_raw);
  float* wlo = whi + W_FLOATS;
  float* xs = wlo + W_FLOATS;        // stacked [X_hi ; X_lo], 128 rows
  float* yhi = xs + F128_FLOATS;     // 64 rows
  float* ylo = yhi + F64_FLOATS;
  float* al = ylo + F64_FLOATS;      // 8 rows
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < W_FLOATS; i += blockDim.x) { whi[i] = whi_g[i]; wlo[i] = wlo_g[i]; }
  for (int i = tid; i < F128_FLOATS; i += blockDim.x) xs[i] = x_g[i];
  for (int i = tid; i < F64_FLOATS; i += blockDim.x) { yhi[i] = yhi_g[i]; ylo[i] = ylo_g[i]; }
  for (int i = tid; i < SBO_F / 4; i += blockDim.x) al[i] = alias_g[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  if (mode <= 2) {
    // this thread's sample row -> TMEM columns 64.. (hi) and 128.. (lo)
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float h[16], l[16];
      for (int i = 0; i < 16; i++) {
        const float a = a_g[tid * 64 + c0 + i];
        const float m = hi_of(a);
        h[i] = mode == 0 ? m : a;
        l[i] = a - m;
      }
      tmem_st16(lane_addr + 64 + c0, h);
      tmem_st16(lane_addr + 128 + c0, l);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 64);
    if (mode <= 2) {
      int first = 1;
      const int groups = mode == 2 ? 1 : 3;
      for (int g = 0; g < groups; g++) {
        const uint32_t acol = tmem + (g == 1 ? 128 : 64);       // g=1: A_lo
        const float* wsel = g == 2 ? wlo : whi;                   // g=2: W_lo
        for (int ks = 0; ks < 8; ks++) {
          mma_ts(tmem, acol + ks * 8, make_desc(smem_u32(wsel) + ks * 2 * LBO_W, LBO_W, SBO_W), idesc, first ? 0 : 1);
          first = 0;
        }
      }
    } else if (mode == 3) {
      int first = 1;
      for (int g = 0; g < 2; g++)
        for (int ks = 0; ks < 16; ks++) {
          mma_ss(tmem, make_desc(smem_u32(xs) + ks * 2 * LBO_F, LBO_F, SBO_F),
                 make_desc(smem_u32(g ? ylo : yhi) + ks * 2 * LBO_F, LBO_F, SBO_F), idesc, first ? 0 : 1);
          first = 0;
        }
    } else {
      for (int ks = 0; ks < 16; ks++)
        mma_ss(tmem, make_desc(smem_u32(al) + ks * 2 * LBO_F, LBO_F, 0), make_desc(smem_u32(yhi) + ks * 2 * LBO_F, LBO_F, SBO_F), idesc, ks > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tmem_ld16(lane_addr + c0, v);
    for (int i = 0; i < 16; i++) out[tid * 64 + c0 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static float hmask(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float frand() { return (float)rand() / RAND_MAX * 2.0f - 1.0f; }

int main() {
  srand(7);
  std::vector<float> A(128 * 64), W(64 * 64), X(128 * 64), Y(128 * 64), AL(8 * 128);   // X, Y: [s][feature]
  for (auto& v : A) v = tanhf(2.0f * frand());
  for (auto& v : W) v = 0.3f * frand();
  for (auto& v : X) v = 0.01f * frand() * frand();
  for (auto& v : Y) v = tanhf(2.0f * frand());
  for (int r = 0; r < 8; r++) for (int s = 0; s < 128; s++) AL[r * 128 + s] = r == 0 ? 1.0f : (r < 5 ? hmask(frand()) : 0.0f);
  std::vector<float> whi(W_FLOATS), wlo(W_FLOATS), xs(F128_FLOATS, 0.f), yhi(F64_FLOATS, 0.f), ylo(F64_FLOATS, 0.f), al(SBO_F / 4, 0.f);
  for (int n = 0; n < 64; n++) for (int k = 0; k < 64; k++) { const float w = W[n * 64 + k], h = hmask(w); whi[w_off(n, k)] = h; wlo[w_off(n, k)] = w - h; }
  for (int s = 0; s < 128; s++) for (int j = 0; j < 64; j++) {
    const float x = X[s * 64 + j], xh = hmask(x); xs[f_off(j, s)] = xh; xs[f_off(64 + j, s)] = x - xh;
    const float y = Y[s * 64 + j], yh = hmask(y); yhi[f_off(j, s)] = yh; ylo[f_off(j, s)] = y - yh;
  }
  for (int r = 0; r < 8; r++) for (int s = 0; s < 128; s++) al[f_off(r, s)] = AL[r * 128 + s];
  auto up = [](const std::vector<float>& h) { float* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); return d; };
  float *dA = up(A), *dwhi = up(whi), *dwlo = up(wlo), *dxs = up(xs), *dyhi = up(yhi), *dylo = up(ylo), *dal = up(al), *dout;
  CK(cudaMalloc(&dout, 128 * 64 * 4));
  const int smem = (2 * W_FLOATS + F128_FLOATS + 2 * F64_FLOATS + SBO_F / 4) * 4;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> out(128 * 64);
  const char* names[] = {"T1 A(TMEM) 3xTF32, masked hi", "T1 A(TMEM) 3xTF32, raw value as hi", "T1 A(TMEM) 1xTF32", "T2 stacked hi/lo over samples (LBO=144)", "T3 aliased 8-row A (SBO=0)"};
  int bad = 0;
  for (int mode = 0; mode < 5; mode++) {
    CK(cudaMemset(dout, 0, 128 * 64 * 4));
    probe<<<1, 128, smem>>>(dA, dwhi, dwlo, dxs, dyhi, dylo, dal, dout, mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0, maxerr32 = 0;
    const int rows = mode == 3 ? 64 : 128;
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < 64; c++) {
        double ref = 0; float ref32 = 0; float got;
        if (mode <= 2) { for (int k = 0; k < 64; k++) { ref += (double)A[r * 64 + k] * W[c * 64 + k]; ref32 = fmaf(A[r * 64 + k], W[c * 64 + k], ref32); } got = out[r * 64 + c]; }
        else if (mode == 3) { for (int s = 0; s < 128; s++) { ref += (double)X[s * 64 + r] * Y[s * 64 + c]; ref32 = fmaf(X[s * 64 + r], Y[s * 64 + c], ref32); } got = out[r * 64 + c] + out[(r + 64) * 64 + c]; }
        else { for (int s = 0; s < 128; s++) { ref += (double)AL[(r % 8) * 128 + s] * hmask(Y[s * 64 + c]); } ref32 = (float)ref; got = out[r * 64 + c]; }
        maxerr = fmax(maxerr, fabs(ref - got));
        maxerr32 = fmax(maxerr32, fabs(ref - ref32));
        maxref = fmax(maxref, fabs(ref));
      }
    const double rel = maxerr / maxref;
    const bool ok = mode == 2 ? rel < 5e-3 : rel < 2e-6;
    printf("%-44s max|err| %.3e  max|ref| %.3e  rel %.3e   (fp32 fmaf chain: %.3e)  %s\n", names[mode], maxerr, maxref, rel, maxerr32 / maxref, ok ? "OK" : "MISMATCH");
    if (!ok && mode != 1) bad++;
  }
  printf(bad ? "tc_probe3: FAILED\n" : "tc_probe3: OK\n");
  return bad;
}
