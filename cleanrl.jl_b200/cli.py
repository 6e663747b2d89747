"""Runner: `python -m cleanrl_jl_b200 {ppo,a2c,dqn} [--field value ...]`.

The reference leaves this wiring as a TODO ("Make individual file runners e.g experiments/run_ppo.(jl/sh)", README.md:24):
`ConfigParser.argparse_struct` exists (config_parser.jl:18-40) but nothing calls it. Here every field of the algorithm's
config struct is a command-line option, typed and defaulted from the struct (argparse_struct), and the parsed struct goes
to the algorithm entry point:   argparse_struct(PPOConfig()) |> ppo.

Multi-GPU: launch under torchrun (`python -m torch.distributed.run --nproc-per-node N -m cleanrl_jl_b200 ppo ...`); the
env vector is sharded over the ranks and rank 0 logs."""
import json
import os
import sys


def _configs():
    from .a2c_algo import A2CConfig
    from .config import PPOConfig
    from .dqn_algo import DQNConfig
    return {"ppo": PPOConfig, "a2c": A2CConfig, "dqn": DQNConfig}


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    algos = _configs()
    if not argv or argv[0] in ("-h", "--help") or argv[0] not in algos:
        sys.stderr.write("usage: python -m cleanrl_jl_b200 {%s} [--<config field> <value> ...]\n"
                         "       (every field of the algorithm's config struct is an option; --help after the algorithm lists them)\n"
                         % ",".join(sorted(algos)))
        return 2
    from .config import argparse_struct
    algo = argv[0]
    config = argparse_struct(algos[algo](), argv[1:])
    world = int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        local_rank = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        from . import a2c, dqn, ppo
        res = {"ppo": ppo, "a2c": a2c, "dqn": dqn}[algo](config)
    finally:
        if dist is not None:
            dist.destroy_process_group()
    if int(os.environ.get("RANK", 0)) == 0:
        keep = {k: v for k, v in res.items() if isinstance(v, (int, float, str)) or v is None}
        print(json.dumps({"algorithm": algo, **keep}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
