"""Learning smoke test (SURVEY §4 tier 4): the public ppo(config) call on a B200 must actually learn CartPole, with
the reference's semantics (quirks Q1-Q7 included) and its default hyper-parameters except the env count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ppo_cartpole_mean_return_rises(tmp_path, torch_cuda):
    from cleanrl_jl_b200 import PPOConfig, ppo
    from cleanrl_jl_b200 import logger as Logger
    cfg = PPOConfig(num_envs=256, num_steps=64, total_timesteps=256 * 64 * 120, seed=3)
    lg = Logger.make_logger("learn", to_terminal=False, to_tensorboard=False, to_json=True, log_dir=str(tmp_path))
    hist = []
    res = ppo(cfg, logger=lg, on_update=lambda u, h, stats, agg: hist.append((u, agg.sum_return / max(agg.count, 1), agg.count)))
    lg.close()
    assert res["num_updates"] == 120 and len(hist) == 120
    first = np.mean([r for _, r, _ in hist[:5]])
    last = np.mean([r for _, r, _ in hist[-10:]])
    assert np.isfinite(res["last_stats"]).all()
    assert first < 40            # random policy: ~20 steps per episode
    assert last > 3 * first and last > 100, (first, last)
    # the JSON log holds the reference's record names and keys
    import json
    recs = [json.loads(l) for l in open(tmp_path / "learn.json")]
    assert {r["msg"] for r in recs} == {"Episode Statistics", "Training Statistics"}
    assert sum(r["msg"] == "Training Statistics" for r in recs) == 120 * 16


def test_ppo_pendulum_runs_and_a2c_runs(tmp_path, torch_cuda):
    from cleanrl_jl_b200 import A2CConfig, PPOConfig, a2c, ppo
    from cleanrl_jl_b200 import logger as Logger
    lg = Logger.make_logger("pend", to_terminal=False, to_tensorboard=False, to_json=False)
    res = ppo(PPOConfig(env_id="Pendulum", num_envs=128, num_steps=64, total_timesteps=128 * 64 * 10), logger=lg)
    assert np.isfinite(res["last_stats"]).all() and np.isfinite(res["params"]).all()
    res = a2c(A2CConfig(num_envs=256, num_steps=32, total_timesteps=256 * 32 * 60, lr=1e-3), logger=lg)
    assert np.isfinite(res["params"]).all() and res["episodes"] > 0 and np.isfinite(res["critic_loss"])
