"""ppo(config) — the drop-in entry point, mirror of src/algorithms/ppo.jl:75-254.

Same structure as the reference's ppo(): logger, env vector, actor/critic, batch arithmetic,
optimiser, then `for update in 1:num_updates` with lr annealing, rollout, GAE and the
epoch/minibatch loop — but every stage is one call into libcleanrl_cuda.so, and in the default
(throughput) mode a whole update is a single asynchronous call replayed as a CUDA graph.

Julia cannot run in this image, so this Python module plays the role of the Julia host code;
cleanrl.jl_b200/julia/CleanRLCuda.jl is the equivalent `ccall` shim.
"""
import time

import numpy as np

from . import _abi
from . import logger as Logger
from . import networks as Networks
from . import parallel
from .config import PPOConfig
from .handle import PPOHandle, comm_unique_id

ENV_KINDS = {"CartPole": _abi.CRL_ENV_CARTPOLE, "Pendulum": _abi.CRL_ENV_PENDULUM}
GAE_MODES = {"ref_compat": _abi.CRL_GAE_REF_COMPAT, "fixed": _abi.CRL_GAE_FIXED}


def make_crl_config(config, num_envs_local=None, device=0, world_size=1, rank=0, env_id_base=0):
    """PPOConfig -> crl_config (hyper-parameters are rounded to Float32 as in the Julia struct)."""
    if config.env_id not in ENV_KINDS:
        raise ValueError("env_id must be one of %s" % sorted(ENV_KINDS))
    if not config.normalize_advantages:
        # ppo.jl:219-222: the `if` has no else branch, mb_advantages becomes `nothing` and the
        # next line throws. Keep the error behaviour.
        raise ValueError("normalize_advantages=false is not supported by the reference (ppo.jl:219-222 throws)")
    kind = ENV_KINDS[config.env_id]
    max_steps = config.max_steps or (500 if kind == _abi.CRL_ENV_CARTPOLE else 200)
    return _abi.make_config(
        env_kind=kind, num_envs=num_envs_local if num_envs_local is not None else config.num_envs,
        num_steps=config.num_steps, num_minibatches=config.num_minibatches, update_epochs=config.update_epochs,
        max_episode_steps=max_steps, gae_mode=GAE_MODES[config.gae_mode], device=device, world_size=world_size,
        rank=rank, env_id_base=env_id_base, flags=(_abi.CRL_FLAG_LOCAL_STATS if config.local_stats else 0) | (0 if config.clip_value_loss else _abi.CRL_FLAG_NO_VCLIP),
        gamma=config.gamma, gae_lambda=config.gae_lambda, clip_coef=config.clip_coef, ent_coeff=config.ent_coeff,
        v_coef=config.v_coef, clip_norm=config.clip_norm, seed=config.seed)


def annealed_lr(config, update, num_updates):
    """ppo.jl:118-121: frac = 1.0 - (update - 1.0) / num_updates; eta = frac * lr (Float64 * Float32).
    `update` is 1-based as in the reference."""
    if not config.anneal_lr:
        return float(np.float32(config.lr))
    frac = 1.0 - (update - 1.0) / num_updates
    return frac * float(np.float32(config.lr))


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def ppo(config=PPOConfig(), *, logger=None, initial_params=None, on_update=None, max_updates=None, device=None):
    """Train PPO. Returns a summary dict (the reference returns nothing).

    Multi-GPU: when torch.distributed is initialised with world_size k, every process owns
    num_envs/k envs, its slice of the rollout buffer and of GAE; gradients are summed with one
    NCCL allreduce per minibatch. Only rank 0 logs.
    """
    dist = _dist()
    rank, local_rank, world = (dist.get_rank(), parallel.dist_info()[1], dist.get_world_size()) if dist else (0, 0, 1)
    nt = config.num_envs  # ppo.jl:76 (global env count)
    if logger is None and rank == 0:
        logger = Logger.make_logger(config.run_name, to_terminal=False)  # ppo.jl:77
    env_base, n_local = parallel.shard_envs(nt, world, rank)
    cfg = make_crl_config(config, n_local, device if device is not None else local_rank, world, rank, env_base)
    h = PPOHandle(cfg)  # env vector, buffer, optimiser state: ppo.jl:80-104
    try:
        if world > 1:
            h.comm_init(parallel.exchange_unique_id(comm_unique_id))
        # actor, critic = Networks.make_actor_critic(...) .|> Flux.f32   (ppo.jl:85-87)
        if initial_params is None:
            initial_params = Networks.init_params(h.continuous, h.d["D"], h.d["A"], seed=config.seed)
        h.set_params(initial_params)

        batch_size = config.num_steps * nt                       # ppo.jl:89
        num_updates = config.total_timesteps // batch_size       # ppo.jl:91
        if max_updates is not None:
            num_updates_run = min(num_updates, max_updates)
        else:
            num_updates_run = num_updates

        global_step = 0       # ppo.jl:106
        last_log_step = 0     # ppo.jl:107
        h2d_bytes = 4 * len(initial_params)
        d2h_bytes = 0
        start_time = time.time()  # ppo.jl:111
        h.env_reset()             # ppo.jl:112-115
        rng = np.random.default_rng(config.seed + 7919 * rank)

        state = {"last_log_step": last_log_step, "episodes": 0, "ret_sum": 0.0, "d2h": 0, "last_stats": None}
        host = {"enqueue_s": 0.0, "fetch_s": 0.0, "log_s": 0.0}  # where the host thread spends its time

        def log_update(stats, agg, gs):
            """records of one finished update (rank 0): aggregated episodes + one record per minibatch"""
            state["d2h"] += stats.nbytes + 40
            state["last_stats"] = stats
            if rank != 0:
                return
            t_log = time.perf_counter()
            if agg is not None and agg.count > 0:
                # throughput mode: one aggregated "Episode Statistics" record per rollout
                steps_per_sec = np.trunc(gs / max(time.time() - start_time, 1e-9))   # ppo.jl:148
                inc = 0 if state["last_log_step"] == 0 else gs - state["last_log_step"]  # ppo.jl:156
                logger.info("Episode Statistics", episode_return=agg.sum_return / agg.count,
                            episode_length=agg.sum_length / agg.count, global_step=gs,
                            steps_per_sec=steps_per_sec, log_step_increment=inc)
                state["last_log_step"] = gs
                state["episodes"] += agg.count
                state["ret_sum"] += agg.sum_return
            for row in stats:  # ppo.jl:246-248, one record per minibatch
                inc = 0 if state["last_log_step"] == 0 else gs - state["last_log_step"]
                logger.info("Training Statistics", loss=row[0], pg_loss=row[1], v_loss=row[2],
                            entropy_loss=row[3], log_step_increment=inc)
                state["last_log_step"] = gs
            host["log_s"] += time.perf_counter() - t_log

        pending = None  # (update, global_step) of the update whose results have not been logged yet
        for update in range(1, num_updates_run + 1):  # ppo.jl:117
            lr_now = annealed_lr(config, update, num_updates)
            step_base = global_step
            if config.log_episodes or config.host_shuffle:
                # stage-by-stage path: same call sequence as the reference's loop body
                h.rollout()                                  # ppo.jl:123-166
                if config.log_episodes and rank == 0:
                    recs, _ = h.pop_episodes()
                    state["d2h"] += 24 * len(recs)
                    for (step, env, length, ep_ret) in recs:  # (step, env) order = ppo.jl:149
                        gs = step_base + (step + 1) * nt      # ppo.jl:124
                        steps_per_sec = np.trunc(gs / max(time.time() - start_time, 1e-9))  # ppo.jl:148
                        inc = 0 if state["last_log_step"] == 0 else gs - state["last_log_step"]  # ppo.jl:156
                        logger.info("Episode Statistics", episode_return=ep_ret, episode_length=float(length),
                                    global_step=gs, steps_per_sec=steps_per_sec, log_step_increment=inc)  # ppo.jl:157
                        state["last_log_step"] = gs  # ppo.jl:161
                        state["episodes"] += 1
                        state["ret_sum"] += ep_ret
                global_step += config.num_steps * nt
                h.gae()                                      # ppo.jl:169-181
                perms = None
                if config.host_shuffle:                      # shuffle(b_inds), ppo.jl:194
                    perms = np.stack([rng.permutation(h.B) for _ in range(config.update_epochs)]).astype(np.int32)
                    h2d_bytes += perms.nbytes
                stats = h.update_epochs(perms, lr_now)       # ppo.jl:191-252
                log_update(stats, None, global_step)
                if on_update is not None:
                    on_update(update, h, stats, None)
            else:
                # throughput path: the whole update is one asynchronous CUDA-graph launch; the host logs
                # update u-1 (double-buffered results) while the GPU runs update u
                t0 = time.perf_counter()
                h.train_update(lr_now)
                t1 = time.perf_counter()
                host["enqueue_s"] += t1 - t0
                global_step += config.num_steps * nt
                h2d_bytes += 8
                if pending is not None:
                    stats, agg = h.fetch_update(lag=1)
                    host["fetch_s"] += time.perf_counter() - t1
                    log_update(stats, agg, pending[1])
                    if on_update is not None:
                        on_update(pending[0], h, stats, agg)
                pending = (update, global_step)
        if pending is not None:
            stats, agg = h.fetch_update(lag=0)
            log_update(stats, agg, pending[1])
            if on_update is not None:
                on_update(pending[0], h, stats, agg)
        last_stats, episodes, ret_sum, d2h_bytes = state["last_stats"], state["episodes"], state["ret_sum"], state["d2h"]
        h.sync()
        elapsed = time.time() - start_time
        return {
            "global_step": global_step, "num_updates": num_updates_run, "elapsed_s": elapsed,
            "steps_per_sec": global_step / max(elapsed, 1e-9), "last_stats": last_stats,
            "episodes": int(episodes), "mean_episode_return": (ret_sum / episodes) if episodes else float("nan"),
            "params": h.get_params(), "kernel_launches": h.kernel_launches(),
            "h2d_bytes": h2d_bytes, "d2h_bytes": d2h_bytes, "host_s": host,
        }
    finally:
        h.close()
