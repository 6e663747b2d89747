#!/bin/sh
# PPO with command-line hyper-parameters (the reference's README.md:24 TODO): every PPOConfig field is an option, e.g.
#   experiments/run_ppo.sh --num_envs 4096 --num_steps 128 --total_timesteps 100000000
#   GPUS=8 experiments/run_ppo.sh --env_id Pendulum --num_envs 65536 --num_steps 128
cd "$(dirname "$0")/.." || exit 1
if [ "${GPUS:-1}" -gt 1 ]; then
  exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port "${PORT:-29500}" \
    cleanrl_jl_b200.py ppo "$@"
fi
exec python -m cleanrl_jl_b200 ppo "$@"
