"""Runner / CLI wiring (the reference's README.md:24 TODO): `python -m cleanrl_jl_b200 <algo> --field value`
builds the algorithm's config struct with ConfigParser.argparse_struct (config_parser.jl:18-40) and calls the entry point."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, timeout=300):
    return subprocess.run([sys.executable, "-m", "cleanrl_jl_b200", *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_cli_usage_and_field_options():
    p = run()
    assert p.returncode == 2 and "{a2c,dqn,ppo}" in p.stderr
    p = run("ppo", "--help")
    assert p.returncode == 0
    for field in ("--total_timesteps", "--num_steps", "--num_envs", "--num_minibatches", "--update_epochs", "--lr", "--gamma",
                  "--gae_lambda", "--clip_coef", "--ent_coeff", "--v_coef", "--normalize_advantages", "--clip_value_loss", "--anneal_lr"):
        assert field in p.stdout, field                      # the reference's 14 PPOConfig fields, ppo.jl:2-18
    p = run("dqn", "--help")
    for field in ("--buffer_size", "--min_buff_size", "--train_freq", "--target_net_freq", "--batch_size", "--epsilon_duration"):
        assert field in p.stdout, field                      # dqn.jl:1-20
    p = run("ppo", "--num_envs", "three")
    assert p.returncode == 2 and "invalid int value" in p.stderr


def test_cli_parses_into_the_config_struct():
    from cleanrl_jl_b200 import PPOConfig, argparse_struct
    c = argparse_struct(PPOConfig(), ["--num_envs", "4096", "--num_steps", "128", "--lr", "1e-3", "--anneal_lr", "false", "--env_id", "Pendulum"])
    assert (c.num_envs, c.num_steps, c.lr, c.anneal_lr, c.env_id) == (4096, 128, 1e-3, False, "Pendulum")
    assert c.total_timesteps == PPOConfig().total_timesteps and isinstance(c, PPOConfig)


@pytest.mark.gpu
@pytest.mark.parametrize("algo,args", [
    ("ppo", ["--num_envs", "64", "--num_steps", "16", "--total_timesteps", "20480"]),
    ("ppo", ["--env_id", "Pendulum", "--num_envs", "64", "--num_steps", "16", "--total_timesteps", "10240"]),
    ("a2c", ["--num_envs", "64", "--num_steps", "16", "--total_timesteps", "20480"]),
    ("dqn", ["--num_envs", "16", "--total_timesteps", "8000", "--min_buff_size", "500", "--buffer_size", "4096"]),
])
def test_cli_runs_the_algorithms(torch_cuda, tmp_path, algo, args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "cleanrl_jl_b200.py"), algo, *args], cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["algorithm"] == algo and res["global_step"] > 0 and res["steps_per_sec"] > 0
