# CleanRLCuda.jl — `ccall` binding of libcleanrl_cuda.so (include/cleanrl_cuda.h) and the drop-in
# `ppo(config)` that replaces src/algorithms/ppo.jl:75-254 of sash-a/CleanRL.jl.
#
# STATUS: this file is the reference-side binding a maintainer would add. Julia is not installed
# in the build image (nor on the GPU boxes), so it has NOT been executed; the Python package next
# to it (cleanrl.jl_b200/*.py) is the executed, tested host side and is a line-for-line mirror of
# the control flow below. Keep the two in sync.
#
# Usage inside the reference repo:
#   include("utils/CleanRLCuda.jl")          # after logger.jl / networks.jl in src/CleanRL.jl
#   CleanRL.CleanRLCuda.ppo()                                         # PPOConfig() defaults, as ppo.jl:75
#   CleanRL.CleanRLCuda.ppo(PPOConfig(num_envs=4096, num_steps=128))
#   julia --project run_ppo.jl --num_envs 4096 --num_steps 128         # argparse_struct(PPOConfig()) |> ppo
module CleanRLCuda

import LinearAlgebra
import Random
using Random: shuffle

const LIB = get(ENV, "CLEANRL_CUDA_LIB", "libcleanrl_cuda.so")

const CRL_ENV_CARTPOLE = Int32(0)
const CRL_ENV_PENDULUM = Int32(1)
const CRL_GAE_REF_COMPAT = Int32(0)
const CRL_GAE_FIXED = Int32(1)
const CRL_GAE_A2C_RETURNS = Int32(2)
const CRL_FLAG_LOCAL_STATS = UInt32(1)
const CRL_FLAG_A2C = UInt32(2)
const CRL_FLAG_NO_VCLIP = UInt32(4)   # PPOConfig.clip_value_loss = false (ppo.jl:16)

# field ids for crl_read_field / crl_write_field
const F_STATE, F_ACTION, F_LOGPROB, F_REWARD, F_TERMINAL, F_VALUE, F_ADVANTAGE, F_RETURN = Int32.(0:7)

# mirror of `struct crl_config` (88 bytes, seed at offset 80; same field order; checked by struct_size)
struct CrlConfig
  struct_size::Int32
  env_kind::Int32
  num_envs::Int32
  num_steps::Int32
  num_minibatches::Int32
  update_epochs::Int32
  max_episode_steps::Int32
  gae_mode::Int32
  device::Int32
  world_size::Int32
  rank::Int32
  env_id_base::Int32
  episode_capacity::Int32
  flags::UInt32
  gamma::Float32
  gae_lambda::Float32
  clip_coef::Float32
  ent_coeff::Float32
  v_coef::Float32
  clip_norm::Float32
  seed::UInt64
end

struct LossStats            # crl_loss_stats, the "Training Statistics" record (ppo.jl:247)
  loss::Float64
  pg_loss::Float64
  v_loss::Float64
  entropy_loss::Float64
end

struct Episode              # crl_episode, one "Episode Statistics" record (ppo.jl:152-157)
  step::Int32
  env::Int32
  length::Int32
  _pad::Int32
  episode_return::Float64
end

struct EpisodeAgg
  count::Int64
  sum_return::Float64
  sum_length::Float64
  max_return::Float64
  dropped::Int64
end

last_error() = unsafe_string(ccall((:crl_last_error, LIB), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : error("libcleanrl_cuda error $rc: $(last_error())")

mutable struct Handle
  ptr::Ptr{Cvoid}
  cfg::CrlConfig
  function Handle(cfg::CrlConfig)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:crl_create, LIB), Cint, (Ref{CrlConfig}, Ref{Ptr{Cvoid}}), cfg, out))
    h = new(out[], cfg)
    finalizer(h -> (h.ptr != C_NULL && ccall((:crl_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr); h.ptr = C_NULL), h)
    h
  end
end

getfield_or(c, f::Symbol, default) = hasproperty(c, f) ? getproperty(c, f) : default

function make_config(config; env_kind=CRL_ENV_CARTPOLE, num_envs=config.num_envs, device=0, world_size=1, rank=0,
                     env_id_base=0, max_steps=500, gae_mode=CRL_GAE_REF_COMPAT, seed=UInt64(1), flags=UInt32(0),
                     num_minibatches=getfield_or(config, :num_minibatches, 1), update_epochs=getfield_or(config, :update_epochs, 1))
  CrlConfig(Int32(sizeof(CrlConfig)), env_kind, num_envs, config.num_steps, num_minibatches,
            update_epochs, max_steps, gae_mode, device, world_size, rank, env_id_base, 0, flags,
            config.gamma, getfield_or(config, :gae_lambda, 1.0f0), getfield_or(config, :clip_coef, 0.2f0),
            getfield_or(config, :ent_coeff, 0.0f0), getfield_or(config, :v_coef, 0.5f0), 0.5f0, seed)
end

# Flux.params(actor, critic) order (ppo.jl:196); each W is already (out,in) column-major in Flux,
# which is exactly the flat layout the library expects: no transposition.
flat_params(actor, critic) = reduce(vcat, [vec(Float32.(p)) for p in Iterators.flatten((Flux_params(actor), Flux_params(critic)))])
Flux_params(chain) = reduce(vcat, [[l.weight, l.bias] for l in chain.layers])

function set_params!(h::Handle, p::Vector{Float32})
  GC.@preserve p check(ccall((:crl_set_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int32), h.ptr, p, length(p)))
end
function get_params(h::Handle, n::Integer)
  p = Vector{Float32}(undef, n)
  GC.@preserve p check(ccall((:crl_get_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int32), h.ptr, p, n))
  p
end

env_reset!(h::Handle) = check(ccall((:crl_env_reset, LIB), Cint, (Ptr{Cvoid},), h.ptr))
rollout!(h::Handle) = check(ccall((:crl_rollout, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float32}), h.ptr, C_NULL, C_NULL))
gae!(h::Handle) = check(ccall((:crl_gae, LIB), Cint, (Ptr{Cvoid},), h.ptr))
train_update!(h::Handle, lr::Float64) = check(ccall((:crl_train_update, LIB), Cint, (Ptr{Cvoid}, Float64), h.ptr, lr))

# perms: Matrix{Int32} (B, update_epochs) of 1-BASED indices as shuffle(1:B) produces (ppo.jl:191-194);
# the C ABI is 0-based, so subtract 1 here.
function update_epochs!(h::Handle, perms::Union{Nothing,Matrix{Int32}}, lr::Float64)
  n = h.cfg.update_epochs * h.cfg.num_minibatches
  stats = Vector{LossStats}(undef, n)
  if perms === nothing
    GC.@preserve stats check(ccall((:crl_update_epochs, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Float64, Ptr{LossStats}),
                                   h.ptr, C_NULL, lr, stats))
  else
    p0 = perms .- Int32(1)
    GC.@preserve p0 stats check(ccall((:crl_update_epochs, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Float64, Ptr{LossStats}),
                                      h.ptr, p0, lr, stats))
  end
  stats
end

function fetch_update(h::Handle)
  n = h.cfg.update_epochs * h.cfg.num_minibatches
  stats = Vector{LossStats}(undef, n)
  agg = Ref(EpisodeAgg(0, 0.0, 0.0, 0.0, 0))
  GC.@preserve stats check(ccall((:crl_fetch_update, LIB), Cint, (Ptr{Cvoid}, Ptr{LossStats}, Ref{EpisodeAgg}), h.ptr, stats, agg))
  stats, agg[]
end

# results of the update enqueued `lag` calls before the latest (0 or 1): with lag = 1 the host logs update u-1 while
# update u is running (crl_fetch_update_at; the Python host's pipelined loop in ppo_algo.py does exactly this)
function fetch_update(h::Handle, lag::Integer)
  n = h.cfg.update_epochs * h.cfg.num_minibatches
  stats = Vector{LossStats}(undef, n)
  agg = Ref(EpisodeAgg(0, 0.0, 0.0, 0.0, 0))
  GC.@preserve stats check(ccall((:crl_fetch_update_at, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{LossStats}, Ref{EpisodeAgg}),
                                 h.ptr, Int32(lag), stats, agg))
  stats, agg[]
end

# speculative updates that failed their on-device check and were replayed exactly (performance counter)
function spec_replays(h::Handle)
  n = Ref{UInt64}(0)
  check(ccall((:crl_spec_replays, LIB), Cint, (Ptr{Cvoid}, Ref{UInt64}), h.ptr, n))
  n[]
end

function pop_episodes(h::Handle, max_records::Integer=1 << 20)
  recs = Vector{Episode}(undef, max_records)
  n = Ref{Int32}(0)
  agg = Ref(EpisodeAgg(0, 0.0, 0.0, 0.0, 0))
  GC.@preserve recs check(ccall((:crl_pop_episodes, LIB), Cint, (Ptr{Cvoid}, Ptr{Episode}, Int32, Ref{Int32}, Ref{EpisodeAgg}),
                                h.ptr, recs, max_records, n, agg))
  resize!(recs, n[]), agg[]
end

# rb.data.*-shaped views: Julia column-major (D,N,T) == the library's [T][N][D], so a plain copy
function read_field(h::Handle, field::Int32, ::Type{T}, dims...) where {T}
  a = Array{T}(undef, dims...)
  GC.@preserve a check(ccall((:crl_read_field, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t), h.ptr, field, a, sizeof(a)))
  a
end

# gae(values, rewards, terminals, γ, λ) with the reference's signature (ppo.jl:48): one env, values [0,k],
# rewards [1,k], terminals [0,k]. Runs crl_gae_raw on device copies made with CUDA.jl.
# (Needs `using CUDA`; left as the call shape only because CUDA.jl is a declared but unused dependency of
# the reference, Project.toml:9.)
#   crl_gae_raw(values[1:k], rewards, terminals[1:k], values[k+1:k+1], terminals[k+1:k+1], adv, ret, k, 1, γ, λ, mode, C_NULL)

# orthogonal initialisation of networks.jl:36-49 without Flux, so that `ppo()` needs no keyword argument: two separate
# 64-64 tanh MLPs, gains sqrt(2) on the hidden layers, 0.01 on the actor head, 1.0 on the critic head, zero biases
# (Flux.orthogonal = Q of a QR of a Gaussian matrix, signs fixed by diag(R)). Flat Flux.params(actor, critic) order.
function default_params(D::Integer, A::Integer; continuous::Bool=false, seed::Integer=1)
  rng = Random.Xoshiro(seed)
  function orth(rows, cols, gain)
    flat = randn(rng, Float64, max(rows, cols), min(rows, cols))
    Q, R = LinearAlgebra.qr(flat)
    Qm = Matrix(Q) * LinearAlgebra.Diagonal(sign.(LinearAlgebra.diag(R)))
    W = rows >= cols ? Qm : Qm'
    Float32.(gain .* W[1:rows, 1:cols])
  end
  function net(out, head_gain)
    [vec(orth(64, D, sqrt(2.0))); zeros(Float32, 64); vec(orth(64, 64, sqrt(2.0))); zeros(Float32, 64);
     vec(orth(out, 64, head_gain)); zeros(Float32, out)]
  end
  p = [net(A, 0.01); net(1, 1.0)]
  continuous ? [p; zeros(Float32, A)] : p     # Gaussian head: state-independent log-std, initialised to 0
end

"""
    ppo(config::PPOConfig = PPOConfig(); actor = nothing, critic = nothing, env_id = "CartPole", seed = 1,
        log_episodes = false)

Drop-in for `CleanRL.ppo` (ppo.jl:75): `ppo()` with no arguments trains CartPole with the reference's defaults. Same
config struct, same `@info` record names and keys. `actor`/`critic` may be the Flux chains of
`Networks.make_actor_critic` (networks.jl:36-49), used only to initialise the device parameters; without them the same
orthogonal initialisation is done here. The loop is the one of `cleanrl.jl_b200/ppo_algo.py` (the executed and tested
host): one asynchronous `crl_train_update` per update, and the records of update u-1 are fetched (lag = 1) and logged
while update u runs, so the stream never waits for the logger.
"""
function ppo(config=nothing; actor=nothing, critic=nothing, env_id::String="CartPole", seed::Integer=1,
             log_episodes::Bool=false)
  config === nothing && (config = Main.CleanRL.PPOConfig())       # ppo.jl:75 default argument
  config.normalize_advantages || error("normalize_advantages=false is not supported by the reference (ppo.jl:219-222 throws)")
  nt = config.num_envs                                            # ppo.jl:76
  continuous = env_id == "Pendulum"
  h = Handle(make_config(config; env_kind=continuous ? CRL_ENV_PENDULUM : CRL_ENV_CARTPOLE,
                         max_steps=continuous ? 200 : 500, seed=UInt64(seed),
                         flags=config.clip_value_loss ? UInt32(0) : CRL_FLAG_NO_VCLIP))
  p = (actor === nothing || critic === nothing) ?
      default_params(continuous ? 3 : 4, continuous ? 1 : 2; continuous=continuous, seed=seed) :
      flat_params(actor, critic)                                  # ppo.jl:85-87,196
  set_params!(h, p)
  batch_size = config.num_steps * nt                              # ppo.jl:89
  num_updates = config.total_timesteps ÷ batch_size               # ppo.jl:91
  global_step = 0                                                 # ppo.jl:106
  last_log_step = 0
  start_time = time()                                             # ppo.jl:111
  env_reset!(h)                                                   # ppo.jl:112-115

  function log_update(stats, agg, gs)                             # records of one finished update
    if agg !== nothing && agg.count > 0
      steps_per_sec = trunc(gs / (time() - start_time))           # ppo.jl:148
      log_step_inc = last_log_step == 0 ? 0 : gs - last_log_step  # ppo.jl:156
      @info "Episode Statistics" episode_return = agg.sum_return / agg.count episode_length = agg.sum_length / agg.count global_step = gs steps_per_sec log_step_increment = log_step_inc
      last_log_step = gs
    end
    for s in stats                                                # ppo.jl:246-248, one record per minibatch
      log_step_inc = last_log_step == 0 ? 0 : gs - last_log_step
      @info "Training Statistics" loss = s.loss pg_loss = s.pg_loss v_loss = s.v_loss entropy_loss = s.entropy_loss log_step_increment = log_step_inc
      last_log_step = gs
    end
  end

  pending = 0                                                     # global_step of the update not logged yet (0 = none)
  for update in 1:num_updates                                     # ppo.jl:117
    lr_now = Float64(config.lr)
    if config.anneal_lr
      frac = 1.0 - (update - 1.0) / num_updates                   # ppo.jl:119
      lr_now = frac * config.lr                                   # ppo.jl:120
    end
    step_base = global_step
    if log_episodes
      # stage-by-stage path: same call sequence as the reference's loop body, one record per episode
      rollout!(h)                                                 # ppo.jl:123-166 in one launch
      recs, _ = pop_episodes(h)
      for r in recs                                               # (step, env) order == ppo.jl:149
        gs = step_base + (r.step + 1) * nt
        steps_per_sec = trunc(gs / (time() - start_time))         # ppo.jl:148
        log_step_inc = last_log_step == 0 ? 0 : gs - last_log_step
        @info "Episode Statistics" episode_return = r.episode_return episode_length = Float64(r.length) global_step = gs steps_per_sec log_step_increment = log_step_inc
        last_log_step = gs
      end
      global_step += batch_size
      gae!(h)                                                     # ppo.jl:169-181
      log_update(update_epochs!(h, nothing, lr_now), nothing, global_step)   # ppo.jl:191-252 (device permutation)
    else
      train_update!(h, lr_now)                                    # whole update, one CUDA-graph launch, asynchronous
      global_step += batch_size
      if pending != 0
        stats, agg = fetch_update(h, 1)                           # update u-1: does not wait for update u
        log_update(stats, agg, pending)
      end
      pending = global_step
    end
  end
  if pending != 0
    stats, agg = fetch_update(h, 0)
    log_update(stats, agg, pending)
  end
  get_params(h, length(p))
end

"""
    a2c(config; actor, critic)

Vectorised counterpart of `CleanRL.a2c` (a2c.jl:27-110) on the same handle: `CRL_FLAG_A2C` selects the A2C losses
(a2c.jl:78-97), `CRL_GAE_A2C_RETURNS` the discounted-return scan (a2c.jl:13-24); one `crl_train_update` is one
rollout of `num_envs x num_steps` transitions followed by the critic and actor steps. `config` needs the fields
`num_envs, num_steps, total_timesteps, lr, gamma, seed` (a2c.jl:1-11 has no env vector; see a2c_algo.py).
"""
function a2c(config; actor, critic)
  cfg = make_config(config; gae_mode=CRL_GAE_A2C_RETURNS, flags=CRL_FLAG_A2C, num_minibatches=1, update_epochs=1)
  h = Handle(cfg)
  p = flat_params(actor, critic)
  set_params!(h, p)
  env_reset!(h)
  batch = config.num_envs * config.num_steps
  global_step = 0
  start_time = time()
  for update in 1:(config.total_timesteps ÷ batch)
    train_update!(h, Float64(config.lr))
    global_step += batch
    stats, agg = fetch_update(h)
    @info "Training Statistics" actor_loss = stats[1].pg_loss critic_loss = stats[1].v_loss            # a2c.jl:100
    if agg.count > 0
      steps_per_sec = trunc(global_step / (time() - start_time))
      @info "Episode Statistics" episode_return = agg.sum_return / agg.count episode_length = agg.sum_length / agg.count global_step steps_per_sec   # a2c.jl:106
    end
  end
  get_params(h, length(p))
end

# ---- DQN (dqn.jl) over the crl_dqn_* entry points ------------------------------------------------------------
struct CrlDqnConfig          # crl_dqn_config
  struct_size::Int32
  num_envs::Int32
  buffer_size::Int32
  min_buff_size::Int32
  batch_size::Int32
  train_freq::Int32
  target_net_freq::Int32
  max_episode_steps::Int32
  device::Int32
  _pad::Int32
  lr::Float64
  gamma::Float64
  epsilon_start::Float64
  epsilon_end::Float64
  epsilon_duration::Float64
  seed::UInt64
end
struct CrlDqnStats           # crl_dqn_stats
  last_loss::Float64
  sum_return::Float64
  sum_length::Float64
  epsilon::Float64
  episodes::Int64
  learn_steps::Int64
  iterations::Int64
  kernel_launches::Int64
end

"""
    comm_unique_id() -> Vector{UInt8}

The 128 bytes of `crl_comm_unique_id` (rank 0 calls it and ships them to the other ranks: MPI.jl, a file, sockets).
"""
function comm_unique_id()
  id = zeros(UInt8, 128)
  GC.@preserve id check(ccall((:crl_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
  id
end

"""
    dqn_comm_init(h, unique_id::Vector{UInt8}, world_size, rank, env_id_base)

Data-parallel DQN over the GPUs of one box (`crl_dqn_comm_init`): one handle per GPU and Julia process, called on all
ranks after `crl_dqn_set_params` and before `crl_dqn_reset` with the 128 bytes rank 0 got from `comm_unique_id()`.
`num_envs`, `buffer_size` and `batch_size` of the handle are then per rank.
"""
function dqn_comm_init(h::Ptr{Cvoid}, unique_id::Vector{UInt8}, world_size::Integer, rank::Integer, env_id_base::Integer)
  length(unique_id) == 128 || error("unique id must have 128 bytes")
  GC.@preserve unique_id check(ccall((:crl_dqn_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Int32),
                                     h, unique_id, world_size, rank, env_id_base))
end

"""
    dqn(config::DQNConfig; q_net, num_envs = 1, seed = 1)

Drop-in for `CleanRL.dqn` (dqn.jl:34): same `DQNConfig`, same two `@info` records. `q_net` is the Flux chain of
`make_nn` (dqn.jl:22-26), used only to initialise the device parameters. `num_envs = 1` reproduces the reference's
schedule step for step; larger values step that many CartPole envs in lockstep on the GPU.
"""
function dqn(config; q_net, num_envs::Integer=1, seed::Integer=1)
  cfg = CrlDqnConfig(Int32(sizeof(CrlDqnConfig)), num_envs, config.buffer_size, config.min_buff_size, config.batch_size,
                     config.train_freq, config.target_net_freq, 200, 0, 0, config.lr, config.gamma, config.epsilon_start,
                     config.epsilon_end, config.epsilon_duration, UInt64(seed))
  out = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:crl_dqn_create, LIB), Cint, (Ref{CrlDqnConfig}, Ref{Ptr{Cvoid}}), cfg, out))
  h = out[]
  p = reduce(vcat, [vec(Float32.(x)) for x in Flux_params(q_net)])                    # Flux.params(q_net), dqn.jl:103
  GC.@preserve p check(ccall((:crl_dqn_set_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int32), h, p, length(p)))
  check(ccall((:crl_dqn_reset, LIB), Cint, (Ptr{Cvoid},), h))                          # reset!(env), dqn.jl:48
  chunk = max(1, config.log_frequencey ÷ num_envs)
  n_iter = config.total_timesteps ÷ num_envs
  done = 0
  start_time = time()
  while done < n_iter
    k = min(chunk, n_iter - done)
    st = Ref(CrlDqnStats(0, 0, 0, 0, 0, 0, 0, 0))
    check(ccall((:crl_dqn_run, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{CrlDqnStats}), h, k, st))   # dqn.jl:49-118
    done += k
    global_step = done * num_envs
    s = st[]
    if s.episodes > 0
      steps_per_sec = trunc(global_step / (time() - start_time))
      @info "Episode Statistics" episode_return = s.sum_return / s.episodes episode_length = s.sum_length / s.episodes global_step ϵ = s.epsilon steps_per_sec   # dqn.jl:82
    end
    s.learn_steps > 0 && @info "Training Statistics" loss = s.last_loss                          # dqn.jl:116
  end
  q = Vector{Float32}(undef, length(p))
  GC.@preserve q check(ccall((:crl_dqn_get_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32), h, q, C_NULL, length(q)))
  check(ccall((:crl_dqn_destroy, LIB), Cint, (Ptr{Cvoid},), h))
  q
end

end # module
