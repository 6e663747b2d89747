/* c_abi_smoke.c -- a pure-C consumer of include/cleanrl_cuda.h (no Python, no ctypes): the library is dlopen'ed and
 * driven through the header alone, the way a Julia `ccall` (or any other FFI) would drive it:
 *   create -> set_params -> env_reset -> train_update x 3 (lag-1 fetch) -> fetch -> get_params -> destroy
 * Build: gcc -std=c99 -Wall -Iinclude tests/c_abi_smoke.c -o c_abi_smoke -ldl -lm
 * Run:   ./c_abi_smoke path/to/libcleanrl_cuda.so        (needs a B200; exit code 0 = ok) */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cleanrl_cuda.h"

#define LOAD(name)                                                        \
  name##_t p_##name = (name##_t)dlsym(lib, #name);                        \
  if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }
#define CHECK(call)                                                                               \
  do {                                                                                            \
    int rc__ = (call);                                                                            \
    if (rc__ != CRL_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, p_crl_last_error()); return 1; } \
  } while (0)

typedef int (*crl_version_t)(void);
typedef const char* (*crl_last_error_t)(void);
typedef int (*crl_create_t)(const crl_config*, crl_ctx**);
typedef int (*crl_destroy_t)(crl_ctx*);
typedef int (*crl_dims_t)(const crl_ctx*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*);
typedef int (*crl_set_params_t)(crl_ctx*, const float*, int32_t);
typedef int (*crl_get_params_t)(crl_ctx*, float*, int32_t);
typedef int (*crl_env_reset_t)(crl_ctx*);
typedef int (*crl_train_update_t)(crl_ctx*, double);
typedef int (*crl_fetch_update_at_t)(crl_ctx*, int32_t, crl_loss_stats*, crl_episode_agg*);
typedef int (*crl_read_field_t)(crl_ctx*, int32_t, void*, size_t);
typedef int (*crl_kernel_launches_t)(const crl_ctx*, uint64_t*);
typedef int (*crl_sync_t)(crl_ctx*);

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s libcleanrl_cuda.so\n", argv[0]); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  LOAD(crl_version) LOAD(crl_last_error) LOAD(crl_create) LOAD(crl_destroy) LOAD(crl_dims) LOAD(crl_set_params)
  LOAD(crl_get_params) LOAD(crl_env_reset) LOAD(crl_train_update) LOAD(crl_fetch_update_at) LOAD(crl_read_field)
  LOAD(crl_kernel_launches) LOAD(crl_sync)
  if (p_crl_version() != CRL_VERSION) { fprintf(stderr, "version mismatch\n"); return 1; }

  crl_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.struct_size = (int32_t)sizeof(cfg);
  cfg.env_kind = CRL_ENV_CARTPOLE;
  cfg.num_envs = 256; cfg.num_steps = 32; cfg.num_minibatches = 4; cfg.update_epochs = 4;   /* PPOConfig shape, more envs */
  cfg.max_episode_steps = 500; cfg.gae_mode = CRL_GAE_REF_COMPAT; cfg.world_size = 1;
  cfg.gamma = 0.99f; cfg.gae_lambda = 0.95f; cfg.clip_coef = 0.2f; cfg.ent_coeff = 0.01f; cfg.v_coef = 0.5f; cfg.clip_norm = 0.5f;
  cfg.seed = 1;

  /* argument validation comes before any device work and reports through crl_last_error */
  crl_ctx* h = NULL;
  crl_config bad = cfg;
  bad.num_minibatches = 7;   /* 256*32 is not divisible by 7: ppo.jl:197,203 would index out of bounds (Q10) */
  if (p_crl_create(&bad, &h) != CRL_ERR_INVALID || h != NULL || strlen(p_crl_last_error()) == 0) {
    fprintf(stderr, "bad config was not rejected\n");
    return 1;
  }

  CHECK(p_crl_create(&cfg, &h));
  int32_t D, A, S, P, n_arrays;
  CHECK(p_crl_dims(h, &D, &A, &S, &P, &n_arrays));
  if (D != 4 || A != 2 || P != 9155 || n_arrays != 12) { fprintf(stderr, "unexpected dims %d %d %d %d\n", D, A, P, n_arrays); return 1; }
  float* p0 = (float*)malloc(sizeof(float) * P);
  float* p1 = (float*)malloc(sizeof(float) * P);
  unsigned int s = 12345u;
  for (int i = 0; i < P; i++) { s = s * 1664525u + 1013904223u; p0[i] = ((float)(s >> 8) / 16777216.0f - 0.5f) * 0.25f; }
  CHECK(p_crl_set_params(h, p0, P));
  CHECK(p_crl_env_reset(h));

  const int nmb = cfg.num_minibatches * cfg.update_epochs;
  crl_loss_stats* st = (crl_loss_stats*)calloc(nmb, sizeof(crl_loss_stats));
  crl_episode_agg agg;
  long long episodes = 0;
  for (int u = 0; u < 3; u++) {
    CHECK(p_crl_train_update(h, 2.5e-4));
    if (u >= 1) { CHECK(p_crl_fetch_update_at(h, 1, st, &agg)); episodes += agg.count; }   /* log update u-1 while u runs */
  }
  CHECK(p_crl_fetch_update_at(h, 0, st, &agg));
  episodes += agg.count;
  for (int k = 0; k < nmb; k++)
    if (!isfinite(st[k].loss) || !isfinite(st[k].pg_loss) || !(st[k].v_loss >= 0.0) || !(st[k].entropy_loss > 0.0)) {
      fprintf(stderr, "minibatch %d: bad statistics %g %g %g %g\n", k, st[k].loss, st[k].pg_loss, st[k].v_loss, st[k].entropy_loss);
      return 1;
    }
  CHECK(p_crl_get_params(h, p1, P));
  double moved = 0.0;
  for (int i = 0; i < P; i++) {
    if (!isfinite(p1[i])) { fprintf(stderr, "parameter %d is not finite\n", i); return 1; }
    moved += fabs((double)p1[i] - (double)p0[i]);
  }
  if (!(moved > 0.0)) { fprintf(stderr, "the parameters did not move\n"); return 1; }
  unsigned char* term = (unsigned char*)malloc((size_t)cfg.num_envs * cfg.num_steps);
  CHECK(p_crl_read_field(h, CRL_F_TERMINAL, term, (size_t)cfg.num_envs * cfg.num_steps));
  for (int n = 0; n < cfg.num_envs; n++)
    if (term[n] != 0) { fprintf(stderr, "Q3 violated: terminal flag set at step 0\n"); return 1; }
  uint64_t launches = 0;
  CHECK(p_crl_kernel_launches(h, &launches));
  if (launches < 3 * (4 + (uint64_t)nmb)) { fprintf(stderr, "only %llu kernel launches\n", (unsigned long long)launches); return 1; }
  CHECK(p_crl_sync(h));
  CHECK(p_crl_destroy(h));
  printf("c_abi_smoke ok: %llu kernels, %lld episodes, last loss %.6f, sum |dp| %.4f\n", (unsigned long long)launches, episodes,
         st[nmb - 1].loss, moved);
  free(p0); free(p1); free(st); free(term);
  dlclose(lib);
  return 0;
}
