#!/bin/sh
# vectorised A2C (a2c.jl) with command-line hyper-parameters: experiments/run_a2c.sh --num_envs 16384 --num_steps 32
cd "$(dirname "$0")/.." || exit 1
exec python -m cleanrl_jl_b200 a2c "$@"
