"""dqn(config) — host mirror of CleanRL.dqn (src/algorithms/dqn.jl:34-121) over the crl_dqn_* entry points.

The reference steps ONE CartPole env; here `num_envs` envs step in lockstep on the GPU (num_envs = 1 reproduces the
reference's schedule exactly: see include/cleanrl_cuda.h, "DQN"). `DQNConfig` keeps the reference's field names,
types and defaults (dqn.jl:1-20, including the `log_frequencey` spelling) and adds `num_envs` and `seed`.

Arithmetic: dqn.jl:37 builds `CartPoleEnv()` with its default T = Float64, so the reference's env state (and the replay
buffer rows typed after it) are Float64. The kernels step the Float32 env of the PPO path (ppo.jl:82), whose accelerations
are already evaluated in Float64 (the `4/3` literal promotes them), so only the stored state is rounded: per step
|s32 - s64| <= 4e-7 + 3e-7 |s64| per component, every discrete quantity (action, reward, termination away from a threshold,
counters) identical; `tests/test_dqn.py::test_float32_env_substitute_for_the_float64_cartpole_of_dqn_and_a2c` asserts it.
The same holds for a2c.jl:33."""
import ctypes as C
import dataclasses
import datetime
import time

import numpy as np

from . import _abi
from . import _lib as L
from . import logger as Logger
from . import parallel
from .handle import comm_unique_id


def _dist():
    """torch.distributed when it is initialised with more than one rank, else None (torch is only plumbing)"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except ImportError:
        pass
    return None


@dataclasses.dataclass
class DQNConfig:
    run_name: str = dataclasses.field(default_factory=lambda: datetime.datetime.now().strftime("%y-%m-%d|%H:%M:%S"))
    log_frequencey: int = 1000          # dqn.jl:4 (sic)
    total_timesteps: int = 500_000      # dqn.jl:6
    buffer_size: int = 10_000           # dqn.jl:8
    min_buff_size: int = 200            # dqn.jl:9
    lr: float = 0.0001                  # dqn.jl:11
    train_freq: int = 10                # dqn.jl:12
    target_net_freq: int = 100          # dqn.jl:13
    batch_size: int = 120               # dqn.jl:14
    gamma: float = 0.99                 # dqn.jl:15
    epsilon_start: float = 1.0          # dqn.jl:17
    epsilon_end: float = 0.05           # dqn.jl:18
    epsilon_duration: float = 10_000.0  # dqn.jl:19 (Float64)
    # ---- not in the reference
    num_envs: int = 1                   # vectorised envs (the reference: one env, dqn.jl:38)
    seed: int = 1
    max_steps: int = 200                # CartPoleEnv() default


def linear_schedule(start_e, end_e, duration, t):
    """dqn.jl:28-31"""
    slope = (end_e - start_e) / duration
    return max(slope * t + start_e, end_e)


def init_q_params(seed=0):
    """make_nn (dqn.jl:22-26): Chain(Dense(4,120,relu), Dense(120,84,relu), Dense(84,2)) with Flux's default
    glorot_uniform weights and zero biases, flattened in Flux.params order, each W (out,in) column-major."""
    rng = np.random.default_rng(seed)
    out = []
    for fan_in, fan_out in ((4, 120), (120, 84), (84, 2)):
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        w = rng.uniform(-lim, lim, size=(fan_out, fan_in)).astype(np.float32)
        out.append(np.asfortranarray(w).ravel(order="F"))
        out.append(np.zeros(fan_out, np.float32))
    p = np.concatenate(out)
    assert p.size == _abi.CRL_DQN_PARAMS
    return p


class DQNHandle:
    """object wrapper over crl_dqn_ctx"""

    def __init__(self, cfg):
        self.lib = L.load()
        self.cfg = cfg
        h = C.c_void_p()
        L.check(self.lib.crl_dqn_create(C.byref(cfg), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            L.check(self.lib.crl_dqn_destroy(self.h))
            self.h = None

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float32)
        L.check(self.lib.crl_dqn_set_params(self.h, L.ptr(p), p.size))

    def get_params(self):
        q = np.zeros(_abi.CRL_DQN_PARAMS, np.float32)
        t = np.zeros(_abi.CRL_DQN_PARAMS, np.float32)
        L.check(self.lib.crl_dqn_get_params(self.h, L.ptr(q), L.ptr(t), q.size))
        return q, t

    def reset(self):
        L.check(self.lib.crl_dqn_reset(self.h))

    def comm_init(self, unique_id: bytes, world_size, rank, env_id_base):
        """make this handle rank `rank` of a data-parallel world (collective; after set_params, before reset)"""
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        L.check(self.lib.crl_dqn_comm_init(self.h, buf, int(world_size), int(rank), int(env_id_base)))

    def run(self, iterations):
        st = _abi.crl_dqn_stats()
        L.check(self.lib.crl_dqn_run(self.h, int(iterations), C.byref(st)))
        return st

    def read_buffer(self):
        cap = self.cfg.buffer_size
        out = {"state": np.zeros((cap, 4), np.float32), "action": np.zeros(cap, np.int32), "reward": np.zeros(cap, np.float32),
               "next_state": np.zeros((cap, 4), np.float32), "terminal": np.zeros(cap, np.uint8)}
        size, ptr = C.c_int32(), C.c_int32()
        L.check(self.lib.crl_dqn_read_buffer(self.h, L.ptr(out["state"]), L.ptr(out["action"]), L.ptr(out["reward"]),
                                             L.ptr(out["next_state"]), L.ptr(out["terminal"]), C.byref(size), C.byref(ptr)))
        out["size"], out["ptr"] = size.value, ptr.value
        return out


def make_crl_dqn_config(config, device=0):
    return _abi.make_dqn_config(num_envs=config.num_envs, buffer_size=config.buffer_size, min_buff_size=config.min_buff_size,
                                batch_size=config.batch_size, train_freq=config.train_freq,
                                target_net_freq=config.target_net_freq, max_episode_steps=config.max_steps, device=device,
                                lr=config.lr, gamma=config.gamma, epsilon_start=config.epsilon_start,
                                epsilon_end=config.epsilon_end, epsilon_duration=config.epsilon_duration, seed=config.seed)


def dqn(config=None, logger=None, params=None, device=0, distributed=None):
    """Runs `total_timesteps` env steps (dqn.jl:49) and returns a summary dict. Logs the reference's two records:
    "Episode Statistics" (episode_return, episode_length, global_step, ϵ, steps_per_sec: dqn.jl:82) aggregated over the
    episodes that ended since the last log, and "Training Statistics" (loss: dqn.jl:116) every `log_frequencey` steps."""
    config = config or DQNConfig()
    # Data-parallel over the GPUs of one box when torch.distributed is initialised (one process per GPU): every rank owns
    # num_envs / k envs, its own ring of buffer_size / k transitions and batch_size / k samples of each learning step;
    # one gradient allreduce per learning step. Only rank 0 logs (its own shard's episodes).
    # `distributed=False` keeps this call on one GPU even inside a torchrun job (independent replicas).
    dist = _dist() if distributed in (None, True) else None
    if distributed and dist is None:
        raise ValueError("distributed=True needs an initialised torch.distributed process group with more than one rank")
    rank, local_rank, world = (dist.get_rank(), parallel.dist_info()[1], dist.get_world_size()) if dist else (0, device, 1)
    own_logger = logger is None and rank == 0
    if logger is None and rank == 0:
        logger = Logger.make_logger("dqn|%s" % config.run_name, to_terminal=False)   # dqn.jl:35
    local = config
    if world > 1:
        env_base, n_local = parallel.shard_envs(config.num_envs, world, rank)
        if config.batch_size % world or config.buffer_size % world:
            raise ValueError("batch_size and buffer_size must be divisible by the number of GPUs")
        local = dataclasses.replace(config, num_envs=n_local, batch_size=config.batch_size // world,
                                    buffer_size=config.buffer_size // world)
    h = DQNHandle(make_crl_dqn_config(local, local_rank))
    h.set_params(init_q_params(config.seed) if params is None else params)
    if world > 1:
        h.comm_init(parallel.exchange_unique_id(comm_unique_id), world, rank, env_base)
    h.reset()
    n_iter = config.total_timesteps // config.num_envs
    chunk = max(1, config.log_frequencey // config.num_envs)
    start = time.time()
    done_iter, episodes, last = 0, 0, None
    while done_iter < n_iter:
        k = min(chunk, n_iter - done_iter)
        last = h.run(k)
        done_iter += k
        global_step = done_iter * config.num_envs
        if logger is None:
            continue
        if last.episodes > 0:
            episodes += last.episodes
            logger.info("Episode Statistics", episode_return=last.sum_return / last.episodes,
                        episode_length=last.sum_length / last.episodes, global_step=global_step, ϵ=last.epsilon,
                        steps_per_sec=int(global_step / max(time.time() - start, 1e-9)))
        if last.learn_steps > 0:
            logger.info("Training Statistics", loss=last.last_loss)
    wall = time.time() - start
    q, tgt = h.get_params()
    h.close()
    if own_logger:
        logger.close()
    return {"global_step": n_iter * config.num_envs, "episodes": episodes, "learn_steps": int(last.learn_steps) if last else 0,
            "last_loss": float(last.last_loss) if last else float("nan"), "steps_per_sec": n_iter * config.num_envs / max(wall, 1e-9),
            "params": q, "target_params": tgt,
            "last_mean_return": (last.sum_return / last.episodes) if last and last.episodes else float("nan")}
