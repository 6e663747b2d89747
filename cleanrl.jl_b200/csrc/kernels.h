// kernels.h — host-visible launchers of the sm_100a kernels (internal; the public ABI is
// include/cleanrl_cuda.h).
#pragma once
#include "common.cuh"

// sets the message crl_last_error() returns (api.cu) and hands the code back
int crl_internal_fail(int code, const char* msg);
// NCCL (resolved with dlopen in api.cu) for the other translation units: opaque communicator, in-place sum
int crl_internal_nccl_comm_init(void** comm, int world, int rank, const void* id128);
int crl_internal_nccl_allreduce_sum(void* comm, void* buf, size_t count, int is_double, cudaStream_t s);
int crl_internal_nccl_allgather(void* comm, const void* send, void* recv, size_t bytes, cudaStream_t s);   // device buffers
void crl_internal_nccl_comm_destroy(void* comm);

struct RolloutArgs {
  const float* params;
  const DevState* ds;
  unsigned long long seed;
  int env_id_base;
  int N, T, max_steps;
  // env state (in/out)
  float* env_state;
  int* env_t;
  double* ep_return;
  int* ep_length;
  uint32_t* reset_count;
  float* next_obs;
  uint8_t* next_done;
  float* next_value;
  // rollout buffer (out)
  float* state;
  void* action;
  float* logprob;
  float* reward;
  float* value;
  uint8_t* terminal;
  // optional injected noise (device copies)
  const double* action_noise;
  const float* reset_noise;
  // episode records
  EpisodeBuf* eb;
  crl_episode* records;
  int ep_capacity;
  // A2C (a2c.jl:52,108): obs = deepcopy(state(env)) is taken AFTER reset!(env), so the step after a termination
  // observes the reset state; PPO keeps the stale terminal observation (Q2, ppo.jl:143 before :164)
  int fresh_obs_after_reset;
};

// where the minibatch sample indices come from: an explicit device array, or the
// Philox-keyed Feistel permutation evaluated in registers (no index array in memory).
struct IdxSrc {
  const int32_t* arr;  // device int32[M] or nullptr
  uint32_t start;      // offset of this minibatch inside the epoch permutation
  uint32_t B;          // batch size (permutation domain)
  int half_bits;
  uint32_t epoch, rank;
  unsigned long long seed;
  const DevState* ds;  // update_index lives on the device
};

struct MbFinal {
  float adv_mean, adv_std, s_unclipped, min_vlc;
  double M_global;
  unsigned long long cnt;  // #{i : s > (clip_i - R_i)^2}, ppo.jl:236 (Q5)
  int need_fixup;          // speculative pass found s > min (clip_i - R_i)^2: the host replays the update exactly
  int _pad;
};

#define LG_EXACT 0  // minibatch scalars (mean, std, s, count) are known: fin is valid
#define LG_SPEC 1   // speculate that the scalar s never wins the max (ppo.jl:236); verify afterwards
#define ADV_CHUNKS 32

// per-minibatch advantage sums for a whole update (they do not depend on the parameters)
struct AdvStatsArgs {
  IdxSrc idx;               // seed / ds / rank / B / half_bits (epoch and start are derived per set)
  const int32_t* arr_base;  // device permutations [epochs][B] (or one index list), nullptr = device permutation
  int B, M, nmb, n_sets;
  const float* advantages;
  double* advparts;         // [n_sets][ADV_CHUNKS][2] = (sum adv, sum adv^2)
  // multi-GPU: the sums of all ranks are exchanged by this kernel itself over peer memory (flag-in-data packets, rows
  // [2 slots][world][x_stride] at byte offset x_off of the exchange buffers; slot/flag from the update index) and added
  // in rank order, so advparts holds the GLOBAL sums on every rank when the kernel ends. nullptr = local sums only.
  unsigned char* const* x_peers;
  const unsigned char* x_local;
  size_t x_off;
  int x_world, x_rank, x_stride;
  int* x_err;
  long long x_timeout;
};

#define CRL_MAX_WORLD 16

struct AdamArgs {
  int env_kind;
  float* params;
  const double* gsum;  // [P+4] reduced sums (double), or nullptr to read gf
  const float* gf;     // [P] Float32 gradient (raw entry point)
  double grad_scale;   // 1, or 1/world with CRL_FLAG_LOCAL_STATS
  double stat_ranks;   // ranks whose loss sums were added into gsum[P..] when M_global is local
  float* image;        // parameter image kept in sync with params (may be nullptr)
  float* grads_out;    // [P] un-clipped Float32 gradient (may be nullptr)
  float* m;
  float* v;
  double* beta_pow;    // [n_arrays][2]
  const DevState* ds;  // lr read from device when lr_host < 0
  double lr_host;
  float clip_norm;
  float ent_coeff, v_coef;
  double M_global;
  int A;
  double* stats_out;   // 4 doubles: loss, pg_loss, v_loss, entropy_loss (may be nullptr)
  // ---- speculative throughput path: the finishing kernel also exchanges the reduced sums with the peers
  //      (NVLink peer memory), verifies the speculation and records a failure for the host
  const double* p2p_local;              // this rank's exchange buffer (every rank pushes into it), or nullptr
  const unsigned long long* p2p_seq;    // exchange sequence number (advanced by loss_grad)
  int* p2p_err;
  int p2p_stride;
  size_t p2p_flags_off;
  int world, rank;
  int algo;                             // 0 = PPO loss scalars, 1 = A2C (actor_loss, critic_loss)
  int verify;                           // 1: check s <= min (clip-R)^2 and set ds_rw->spec_failed otherwise
  int M, P;
  DevState* ds_rw;
  MbFinal* fin;
  // ---- fused tail (loss_grad_tc_kernel finishes the minibatch itself, see UpdateArgs::fuse_tail)
  unsigned long long* grid_bar;         // monotonic arrival counter of the grid barriers
  double* normpart;                     // [n_arrays][grid] per-CTA partial sums of squares of the Float32 gradient
  unsigned char* const* ll_peers;       // device array [world]: every rank's exchange buffer (this rank's included)
  const unsigned char* ll_local;        // this rank's exchange buffer, or nullptr (single GPU)
  size_t ll_off;                        // byte offset of the flag-in-data ("LL") region inside an exchange buffer
  int ll_stride;                        // 16-byte packets per (slot, rank) row of that region
  long long timeout_cycles;             // give up waiting for a peer after this many clock64 ticks (error, no update)
};

struct UpdateArgs {
  int env_kind;
  int algo;            // 0 = PPO clipped surrogate (ppo.jl:213-243), 1 = A2C losses (a2c.jl:78-97)
  int no_vclip;        // PPO with clip_value_loss = false: v_loss = 0.5 mean((newvalue - R)^2), ppo.jl:239-241
  const float* params;
  const float* image;  // shared-memory image of the parameters (see param_image_floats) or nullptr
  IdxSrc idx;
  int M;  // local minibatch size
  // flattened rollout data (ppo.jl:184-189)
  const float* states;
  const void* actions;
  const float* logprobs;
  const float* advantages;
  const float* returns;
  const float* values;
  float clip_coef, ent_coeff, v_coef;
  // scratch
  float* vnew;           // [M]
  MbScalars* parts;      // per-CTA partial sums written by mb_stats
  int n_parts_cap;
  const MbScalars* parts_in;  // what mb_count reduces (local parts, or the all-gathered ranks)
  int n_parts_in;
  MbFinal* fin;
  int world;
  float* gpart;          // [grid][P] per-CTA partial gradients
  double* spart;         // [grid][4] per-CTA partial loss sums
  int grid_loss;
  double* gsum;          // [P + 4] reduced gradient (+ loss sums) in double: the allreduce buffer
  int mode;              // LG_EXACT | LG_SPEC
  const double* advparts;  // [ADV_CHUNKS][2] of this minibatch (LG_SPEC)
  float* mpart;          // [grid] per-CTA min (clip_i - R_i)^2 (LG_SPEC)
  int rank;
  // peer-memory allreduce (NVLink/NVSwitch): grad_reduce PUSHES its sums into every rank's exchange buffer
  // (layout [2 slots][world][p2p_stride] doubles, then [world] arrival flags); the last block to finish raises this
  // rank's flag on every peer, and clip_adam only ever reads its own memory
  double* p2p_data;                     // own exchange buffer (non-null = peer path active)
  int p2p_stride;
  unsigned long long* p2p_seq;          // exchange sequence number: loss_grad advances it, slot = seq & 1
  unsigned char* const* p2p_peers;      // device array [world] of all ranks' exchange buffers (this rank's included)
  size_t p2p_flags_off;                 // byte offset of the arrival flags inside an exchange buffer
  unsigned int* p2p_count;              // local count of grad_reduce blocks that have pushed their part
  // tcgen05 kernel (update_tc.cu): CTAs [0, tc_actor_ctas) own the actor, the rest the critic, and each CTA writes
  // only its own net's slice of gpart; 0 = the FFMA kernel (every CTA writes all P elements)
  int tc_actor_ctas;
  int tc_net_a, tc_net_c;               // sizes of the actor / critic slices of the flat parameter vector
  int values_fresh;                     // `values` holds the critic's output under the CURRENT parameters (rollout ->
                                        // one update, crl_train_update): lets the A2C actor CTAs take R - v from it
  // Fused tail (tcgen05 kernel, speculative chain): after writing its partial gradient every CTA crosses a grid barrier,
  // reduces ITS slab of the gradient over all CTAs' partials in the fixed order of grad_reduce, exchanges the slab with
  // the peers (flag-in-data packets pushed over NVLink, no fence, no counter), and after a second grid barrier applies
  // per-array clip + Adam to the slab: ONE cooperative launch per minibatch instead of three kernels.
  int fuse_tail;
  AdamArgs adam;
};

// per-device opt-in to large dynamic shared memory; call once per device outside stream capture
cudaError_t kernels_init_rollout();
cudaError_t kernels_init_update();
cudaError_t launch_episode_buf_init(EpisodeBuf* eb, cudaStream_t s);
cudaError_t launch_rollout(int env_kind, const RolloutArgs& a, cudaStream_t s);
cudaError_t launch_env_step_raw(int env_kind, float* state, int* t, const void* action, float* reward,
                                uint8_t* done, long long n, int max_steps, cudaStream_t s);
cudaError_t launch_policy_forward_raw(int env_kind, const float* params, const float* obs, float* out_policy,
                                      float* logp, float* value, long long n, cudaStream_t s);
cudaError_t launch_gae(const float* values, const float* rewards, const uint8_t* dones, const float* next_value,
                       const uint8_t* next_done, float* adv, float* ret, int T, long long N, float gamma,
                       float lambda, int mode, cudaStream_t s);

int mb_stats_grid(int M, int sm_count);
int loss_grad_grid(int M, int sm_count);
// The parameter image is the exact shared-memory layout loss_grad wants (each net at a 16-byte aligned base,
// then both W2 matrices transposed), kept in global memory so a CTA can stage it with one bulk async copy.
int param_image_floats(int env_kind);
cudaError_t launch_param_image(int env_kind, const float* params, float* image, cudaStream_t s);
cudaError_t launch_adv_stats(const AdvStatsArgs& a, cudaStream_t s);
cudaError_t launch_mb_stats(const UpdateArgs& a, int grid, cudaStream_t s);
cudaError_t launch_mb_count(const UpdateArgs& a, cudaStream_t s);
cudaError_t launch_loss_grad(const UpdateArgs& a, cudaStream_t s);
// tensor-core variant: loss_grad_tc_plan decides whether it applies and sets grid_loss / tc_actor_ctas
cudaError_t kernels_init_update_tc();
// full_grid: one CTA per SM even when the minibatch has fewer tiles (the fused tail's reduce + Adam work is spread over the
// grid: with the two CTAs a 32-sample minibatch needs, the tail alone took 100 us)
int loss_grad_tc_plan(UpdateArgs* a, int sm_count, bool full_grid = false);
cudaError_t launch_loss_grad_tc(const UpdateArgs& a, cudaStream_t s);
cudaError_t launch_grad_reduce(const UpdateArgs& a, int P, cudaStream_t s);
cudaError_t launch_clip_adam(const AdamArgs& a, cudaStream_t s);

cudaError_t launch_loss_finalize(const double* gsum, int P, float* grads_out, double Mg, int A, float ent_coeff,
                                 float v_coef, double* stats_out, cudaStream_t s);
cudaError_t launch_stats_pack(const MbScalars* parts, int n, MbScalars* out, cudaStream_t s);
cudaError_t launch_env_reset(int env_kind, int N, unsigned long long seed, int env_id_base, float* env_state,
                             int* env_t, double* ep_return, int* ep_length, uint32_t* reset_count, float* next_obs,
                             uint8_t* next_done, cudaStream_t s);
cudaError_t launch_env_refresh(int env_kind, int N, const float* env_state, float* next_obs, uint8_t* next_done,
                               cudaStream_t s);
cudaError_t launch_advance(DevState* ds, unsigned long long d_policy_step, unsigned long long d_update, cudaStream_t s);
// all epoch permutations of the current update (update_index from DevState): out[n_epochs][B]
cudaError_t launch_fill_perms_dev(int32_t* out, uint32_t B, unsigned long long seed, const DevState* ds, int n_epochs,
                                  uint32_t rank, cudaStream_t s);
cudaError_t launch_fill_perm(int32_t* out, uint32_t B, unsigned long long seed, unsigned long long update_index,
                             uint32_t epoch, uint32_t rank, cudaStream_t s);
