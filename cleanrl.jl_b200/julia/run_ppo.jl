# run_ppo.jl -- the runner the reference lists as a TODO ("Make individual file runners e.g experiments/run_ppo.(jl/sh)",
# README.md:24): every PPOConfig field is a command-line option (ConfigParser.argparse_struct, config_parser.jl:18-40)
# and the parsed struct goes to the CUDA-backed ppo():
#
#   CLEANRL_CUDA_LIB=/path/to/libcleanrl_cuda.so julia --project run_ppo.jl --num_envs 4096 --num_steps 128
#
# NOT EXECUTED in this repository's CI (no Julia in the image); the executed equivalent is
# `python -m cleanrl_jl_b200 ppo --num_envs 4096 --num_steps 128` (cleanrl.jl_b200/cli.py).
using CleanRL
using CleanRL: PPOConfig
using CleanRL.ConfigParser: argparse_struct
include(joinpath(@__DIR__, "CleanRLCuda.jl"))

CleanRL.Logger.make_logger("ppo-cuda")                 # logger.jl:7: terminal + TensorBoard + JSON sinks
argparse_struct(PPOConfig()) |> CleanRLCuda.ppo
