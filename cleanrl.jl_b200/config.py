"""PPOConfig and ConfigParser.argparse_struct — mirrors of ppo.jl:1-19 and config_parser.jl:18-40.

Field names, types and defaults of the reference's PPOConfig are kept verbatim (including the
spelling `ent_coeff` vs `v_coef`). The extra fields at the bottom configure what the reference
hard-codes; their defaults reproduce the reference's behaviour.
"""
import argparse
import dataclasses
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class PPOConfig:
    # --- the reference's 14 fields, ppo.jl:2-18
    total_timesteps: int = 500_000
    num_steps: int = 32
    num_envs: int = 4
    num_minibatches: int = 4
    update_epochs: int = 4

    lr: float = float(np.float32(2.5e-4))
    gamma: float = float(np.float32(0.99))
    gae_lambda: float = float(np.float32(0.95))

    clip_coef: float = float(np.float32(0.2))
    ent_coeff: float = float(np.float32(0.01))
    v_coef: float = float(np.float32(0.5))

    normalize_advantages: bool = True
    clip_value_loss: bool = True
    anneal_lr: bool = True

    # --- new fields (defaults = what the reference hard-codes)
    env_id: str = "CartPole"      # ppo.jl:79-83 hard-codes CartPoleEnv; "Pendulum" selects the Gaussian policy
    max_steps: int = 0            # 0 = env default (500 CartPole, ppo.jl:82; 200 Pendulum)
    seed: int = 1                 # Philox key (the reference seeds Xoshiro from hash(threadid), ppo.jl:81)
    gae_mode: str = "ref_compat"  # "ref_compat" = ppo.jl:66 as written (Q1); "fixed" = bootstrap from the last step
    clip_norm: float = 0.5        # ClipNorm(0.5), ppo.jl:93
    run_name: str = "ppo-2-test"  # ppo.jl:77
    log_episodes: bool = False    # True: one "Episode Statistics" record per episode in (step, env) order (Q11)
    host_shuffle: bool = False    # True: host permutation per epoch (ppo.jl:194); False: device permutation
    local_stats: bool = False     # multi-GPU only: per-shard minibatch statistics


def _parse_bool(s):
    if isinstance(s, bool):
        return s
    if s.lower() in ("true", "1", "yes"):
        return True
    if s.lower() in ("false", "0", "no"):
        return False
    raise argparse.ArgumentTypeError("expected true/false, got %r" % s)


def argparse_struct(s, argv=None):
    """config_parser.jl:18-40: one `--<field>` option per struct field, typed and defaulted from
    the instance; returns a new struct of the same type with the CLI overrides applied."""
    parser = argparse.ArgumentParser()
    for f in dataclasses.fields(s):
        value = getattr(s, f.name)
        typ = _parse_bool if isinstance(value, bool) else type(value)
        parser.add_argument("--%s" % f.name, default=value, type=typ)
    args = parser.parse_args(argv)
    return type(s)(**vars(args))
