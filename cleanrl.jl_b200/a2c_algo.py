"""a2c(config) — vectorised A2C on the same CUDA kernels (BASELINE config 3), mirroring
src/algorithms/a2c.jl.

What is kept from the reference: `A2CConfig` (a2c.jl:1-10), `discounted_future_rewards` (a2c.jl:13-24, as the
CRL_GAE_A2C_RETURNS scan), the two loss expressions (critic `mean(advantage .^ 2)`, actor
`-mean(log_probs .* advantage)`, a2c.jl:78-97), `Flux.Optimiser(ClipNorm(0.5), Adam(lr))` (a2c.jl:36) and the record
names/keys (a2c.jl:100,106). What differs, because the reference is single-env and episodic: the rollout is a fixed
`num_steps` over `num_envs` vectorised envs (bootstrapped with the critic at the cut), and the two `update!` calls
are one combined step (identical, the parameter sets are disjoint). Parity is therefore claimed only for the return
scan and the two loss expressions, against the oracle. As in a2c.jl:108 then :52, the observation after a termination is
the RESET state (`CRL_FLAG_A2C` makes the rollout kernel refresh it; the PPO path keeps ppo.jl's stale terminal
observation), so no sample pairs a terminal observation with a new episode's return."""
import time
from dataclasses import dataclass

import numpy as np

from . import _abi
from . import logger as Logger
from . import networks as Networks
from .handle import PPOHandle


@dataclass(frozen=True)
class A2CConfig:
    # --- the reference's fields, a2c.jl:2-9
    run_name: str = "a2c"            # the reference formats now(); a2c.jl:2
    lr: float = 0.0001
    total_timesteps: int = 1_000_000
    min_replay_size: int = 512
    gamma: float = 0.99
    # --- vectorised extension (BASELINE config 3: 16,384 envs)
    num_envs: int = 16
    num_steps: int = 32               # steps per env between updates (the n of the n-step return)
    env_id: str = "CartPole"
    seed: int = 1
    clip_norm: float = 0.5            # ClipNorm(0.5), a2c.jl:36


def make_crl_config(config, device=0):
    kind = {"CartPole": _abi.CRL_ENV_CARTPOLE, "Pendulum": _abi.CRL_ENV_PENDULUM}[config.env_id]
    return _abi.make_config(env_kind=kind, num_envs=config.num_envs, num_steps=config.num_steps, num_minibatches=1,
                            update_epochs=1, gae_mode=_abi.CRL_GAE_A2C_RETURNS, flags=_abi.CRL_FLAG_A2C, device=device,
                            gamma=config.gamma, gae_lambda=1.0, clip_norm=config.clip_norm, seed=config.seed)


def a2c(config=A2CConfig(), *, logger=None, initial_params=None, max_updates=None, device=0):
    if logger is None:
        logger = Logger.make_logger("a2c|%s" % config.run_name)  # a2c.jl:31
    h = PPOHandle(make_crl_config(config, device))
    try:
        if initial_params is None:
            initial_params = Networks.init_params(h.continuous, h.d["D"], h.d["A"], seed=config.seed)  # a2c.jl:35
        h.set_params(initial_params)
        batch = config.num_envs * config.num_steps
        num_updates = config.total_timesteps // batch
        if max_updates is not None:
            num_updates = min(num_updates, max_updates)
        start_time = time.time()  # a2c.jl:49
        h.env_reset()             # a2c.jl:50
        global_step, episodes, ret_sum, last = 0, 0, 0.0, None
        pending = None
        for update in range(1, num_updates + 1):
            h.train_update(config.lr)
            global_step += batch
            if pending is not None:
                stats, agg = h.fetch_update(lag=1)
                last = _log(logger, stats, agg, pending, start_time)
                episodes += agg.count
                ret_sum += agg.sum_return
            pending = global_step
        if pending is not None:
            stats, agg = h.fetch_update(lag=0)
            last = _log(logger, stats, agg, pending, start_time)
            episodes += agg.count
            ret_sum += agg.sum_return
        elapsed = time.time() - start_time
        return {"global_step": global_step, "num_updates": num_updates, "elapsed_s": elapsed,
                "steps_per_sec": global_step / max(elapsed, 1e-9), "actor_loss": None if last is None else float(last[1]),
                "critic_loss": None if last is None else float(last[2]), "episodes": int(episodes),
                "mean_episode_return": (ret_sum / episodes) if episodes else float("nan"), "params": h.get_params()}
    finally:
        h.close()


def _log(logger, stats, agg, global_step, start_time):
    row = stats[-1]
    logger.info("Training Statistics", actor_loss=row[1], critic_loss=row[2])  # a2c.jl:100
    if agg.count > 0:
        steps_per_sec = np.trunc(global_step / max(time.time() - start_time, 1e-9))  # a2c.jl:105
        logger.info("Episode Statistics", episode_return=agg.sum_return / agg.count,
                    episode_length=agg.sum_length / agg.count, global_step=global_step, steps_per_sec=steps_per_sec)  # a2c.jl:106
    return row
