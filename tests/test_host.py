"""Host-side mirror of the reference's Julia interface (config, CLI parser, networks, logger,
sharding, lr schedule). CPU only."""
import dataclasses
import json
import os

import numpy as np
import pytest

from cleanrl_jl_b200 import config as ConfigParser
from cleanrl_jl_b200 import logger as Logger
from cleanrl_jl_b200 import networks as Networks
from cleanrl_jl_b200 import parallel
from cleanrl_jl_b200.config import PPOConfig

F = np.float32


def test_ppoconfig_defaults_match_ppo_jl():
    c = PPOConfig()
    ref = dict(total_timesteps=500_000, num_steps=32, num_envs=4, num_minibatches=4, update_epochs=4,
               lr=F(2.5e-4), gamma=F(0.99), gae_lambda=F(0.95), clip_coef=F(0.2), ent_coeff=F(0.01), v_coef=F(0.5),
               normalize_advantages=True, clip_value_loss=True, anneal_lr=True)  # ppo.jl:2-18
    names = [f.name for f in dataclasses.fields(c)]
    assert len(ref) == 14 and names[:14] == list(ref)  # same order, same spelling (ent_coeff vs v_coef)
    for k, v in ref.items():
        assert getattr(c, k) == (float(v) if isinstance(v, np.floating) else v), k
    with pytest.raises(dataclasses.FrozenInstanceError):
        c.lr = 1.0  # Julia structs are immutable
    assert (c.num_steps * c.num_envs, c.num_steps * c.num_envs // c.num_minibatches,
            c.total_timesteps // (c.num_steps * c.num_envs)) == (128, 32, 3906)  # ppo.jl:89-91


def test_argparse_struct_round_trip():
    c = ConfigParser.argparse_struct(PPOConfig(), [])
    assert c == PPOConfig()
    c = ConfigParser.argparse_struct(PPOConfig(), ["--num_envs", "4096", "--lr", "0.001", "--anneal_lr", "false",
                                                   "--env_id", "Pendulum"])
    assert (c.num_envs, c.lr, c.anneal_lr, c.env_id) == (4096, 0.001, False, "Pendulum")
    assert type(c) is PPOConfig and c.num_steps == 32
    with pytest.raises(SystemExit):
        ConfigParser.argparse_struct(PPOConfig(), ["--no_such_field", "1"])
    with pytest.raises(SystemExit):
        ConfigParser.argparse_struct(PPOConfig(), ["--num_envs", "four"])


def test_annealed_lr_matches_ppo_jl_118_121():
    from cleanrl_jl_b200.ppo_algo import annealed_lr
    c = PPOConfig()
    n = 3906
    assert annealed_lr(c, 1, n) == float(F(2.5e-4))
    assert annealed_lr(c, n, n) == (1.0 - (n - 1.0) / n) * float(F(2.5e-4))
    assert annealed_lr(dataclasses.replace(c, anneal_lr=False), 77, n) == float(F(2.5e-4))


def test_make_crl_config_and_error_behaviour(abi):
    from cleanrl_jl_b200.ppo_algo import make_crl_config
    cfg = make_crl_config(PPOConfig())
    assert (cfg.env_kind, cfg.num_envs, cfg.num_steps, cfg.max_episode_steps, cfg.gae_mode) == (0, 4, 32, 500, 0)
    assert cfg.clip_norm == 0.5 and F(cfg.gamma) == F(0.99)
    cfg = make_crl_config(PPOConfig(env_id="Pendulum", gae_mode="fixed"), 8, 3, 4, 2, 16)
    assert (cfg.env_kind, cfg.num_envs, cfg.device, cfg.world_size, cfg.rank, cfg.env_id_base, cfg.max_episode_steps,
            cfg.gae_mode) == (1, 8, 3, 4, 2, 16, 200, 1)
    with pytest.raises(ValueError):  # ppo.jl:219-222 throws for normalize_advantages=false (Q6)
        make_crl_config(PPOConfig(normalize_advantages=False))
    with pytest.raises(ValueError):
        make_crl_config(PPOConfig(env_id="HalfCheetah"))


def test_orthogonal_init_like_flux():
    rng = np.random.default_rng(0)
    for rows, cols, gain in [(64, 4, np.sqrt(2)), (64, 64, np.sqrt(2)), (2, 64, 0.01), (1, 64, 1.0)]:
        W = Networks.orthogonal(rng, rows, cols, gain)
        assert W.shape == (rows, cols) and W.dtype == np.float32
        G = W @ W.T if rows <= cols else W.T @ W
        np.testing.assert_allclose(G, gain ** 2 * np.eye(min(rows, cols)), atol=1e-5)


def test_make_actor_critic_and_flat_layout(olib, abi):
    actor, critic = Networks.make_actor_critic(2, 4, seed=3)
    assert [w.shape for w, _ in actor] == [(64, 4), (64, 64), (2, 64)]
    assert [w.shape for w, _ in critic] == [(64, 4), (64, 64), (1, 64)]
    assert all(np.all(b == 0) for _, b in actor + critic)
    flat = Networks.flatten_params(actor, critic)
    off, size = olib.param_layout(abi.CRL_ENV_CARTPOLE)
    assert flat.size == 9155 == off[-1] + size[-1]
    # W (out,in) column-major: element (j,k) at j + out*k
    assert flat[off[0] + 5 + 64 * 3] == actor[0][0][5, 3]
    assert flat[off[4] + 1 + 2 * 10] == actor[2][0][1, 10]
    assert flat[off[8] + 7 + 64 * 9] == critic[1][0][7, 9]
    a2, c2, ls = Networks.unflatten_params(flat, 4, 2)
    assert ls is None and all(np.array_equal(x[0], y[0]) for x, y in zip(actor + critic, a2 + c2))
    # the oracle's forward agrees with a plain NumPy evaluation of the same layers
    obs = np.random.default_rng(1).standard_normal((7, 4)).astype(F)
    pol, logp, val = olib.policy_forward_raw(abi.CRL_ENV_CARTPOLE, flat, obs)
    h = obs.T
    for W, b in actor[:2]:
        h = np.tanh(W @ h + b[:, None])
    np.testing.assert_allclose(pol, (actor[2][0] @ h + actor[2][1][:, None]).T, rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError):
        Networks.make_actor_critic(2, 4, hidden_sizes=(32, 32))
    p = Networks.init_params(True, 3, 1, seed=0)
    assert p.size == 4481 * 2 + 1 and p[-1] == 0


def test_logger_records_and_step_increment(tmp_path):
    lg = Logger.make_logger("run", to_terminal=False, to_tensorboard=True, to_json=True, log_dir=str(tmp_path))
    assert Logger.global_logger() is lg
    lg.info("Episode Statistics", episode_return=12.0, episode_length=12.0, global_step=128, steps_per_sec=1e6,
            log_step_increment=0)
    lg.info("Training Statistics", loss=0.5, pg_loss=0.1, v_loss=0.2, entropy_loss=0.3, log_step_increment=128)
    lg.close()
    recs = [json.loads(l) for l in open(tmp_path / "run.json")]
    assert [r["msg"] for r in recs] == ["Episode Statistics", "Training Statistics"]
    assert set(recs[0]["kwargs"]) == {"episode_return", "episode_length", "global_step", "steps_per_sec", "log_step_increment"}
    assert set(recs[1]["kwargs"]) == {"loss", "pg_loss", "v_loss", "entropy_loss", "log_step_increment"}
    assert lg.step == 128 and lg.records == 2
    scal = tmp_path / "run" / "scalars.jsonl"
    if scal.exists():  # tensorboard not installed: same (tag, step, value) triples as JSON lines
        rows = [json.loads(l) for l in open(scal)]
        assert {"tag": "Training Statistics/loss", "step": 128, "value": 0.5} in rows
    null = Logger.make_logger("x", to_terminal=False, to_tensorboard=False, to_json=False)
    null.info("anything", a=1)


def test_logger_record_without_step_increment_advances_by_one(tmp_path):
    """TensorBoardLogger.jl advances its step by 1 per record unless `log_step_increment` says otherwise: the DQN and A2C
    records (dqn.jl:82,116; a2c.jl:100,106) carry no such key, so each of them gets its own step"""
    lg = Logger.make_logger("dq", to_terminal=False, to_tensorboard=True, to_json=False, log_dir=str(tmp_path))
    lg.info("Training Statistics", actor_loss=0.1, critic_loss=0.2)
    lg.info("Episode Statistics", episode_return=9.0, episode_length=9, global_step=9, steps_per_sec=10.0)
    lg.info("Training Statistics", loss=0.5, log_step_increment=0)   # PPO-style explicit zero still means zero
    lg.close()
    assert lg.step == 2 and lg.records == 3
    scal = tmp_path / "dq" / "scalars.jsonl"
    if scal.exists():
        rows = [json.loads(l) for l in open(scal)]
        assert {"tag": "Training Statistics/actor_loss", "step": 1, "value": 0.1} in rows
        assert {"tag": "Episode Statistics/episode_return", "step": 2, "value": 9.0} in rows
        assert {"tag": "Training Statistics/loss", "step": 2, "value": 0.5} in rows


def _crc32c(data):
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    c = 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked(data):
    c = _crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_native_tensorboard_writer_produces_valid_event_files(tmp_path):
    """crl_tb_* (csrc/tblog.cu): TFRecord framing with masked CRC-32C, one Event per record, "<message>/<key>" tags at
    the logger's step (TensorBoardLogger.jl semantics, logger.jl:7-29)"""
    import glob
    import struct
    lg = Logger.make_logger("tb", to_terminal=False, to_tensorboard=True, to_json=False, log_dir=str(tmp_path))
    lg.info("Episode Statistics", episode_return=12.5, episode_length=12.0, global_step=128, steps_per_sec=1e6, log_step_increment=0)
    lg.info("Training Statistics", loss=0.5, pg_loss=-0.25, v_loss=0.125, entropy_loss=0.3, log_step_increment=128)
    lg.close()
    files = glob.glob(str(tmp_path / "tb" / "events.out.tfevents.*"))
    assert len(files) == 1
    raw = open(files[0], "rb").read()
    recs, off = [], 0
    while off < len(raw):
        (n,) = struct.unpack_from("<Q", raw, off)
        assert struct.unpack_from("<I", raw, off + 8)[0] == _masked(raw[off:off + 8])
        data = raw[off + 12:off + 12 + n]
        assert struct.unpack_from("<I", raw, off + 12 + n)[0] == _masked(data)
        recs.append(data)
        off += 16 + n
    assert len(recs) == 3 and b"brain.Event:2" in recs[0]
    assert b"Episode Statistics/episode_return" in recs[1] and b"Training Statistics/entropy_loss" in recs[2]
    assert struct.pack("<f", 12.5) in recs[1] and struct.pack("<f", -0.25) in recs[2]
    try:
        from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    except Exception:
        return
    ea = EventAccumulator(str(tmp_path / "tb"))
    ea.Reload()
    assert set(ea.Tags()["scalars"]) == {"Episode Statistics/" + k for k in ("episode_return", "episode_length", "global_step", "steps_per_sec")} | \
        {"Training Statistics/" + k for k in ("loss", "pg_loss", "v_loss", "entropy_loss")}
    ev = ea.Scalars("Training Statistics/pg_loss")
    assert len(ev) == 1 and ev[0].step == 128 and ev[0].value == -0.25
    assert ea.Scalars("Episode Statistics/episode_return")[0].step == 0


def test_shard_envs():
    assert parallel.shard_envs(65536, 8, 3) == (3 * 8192, 8192)
    assert parallel.shard_envs(4, 1, 0) == (0, 4)
    with pytest.raises(ValueError):
        parallel.shard_envs(10, 4, 0)
    assert parallel.dist_info() == (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
                                    int(os.environ.get("WORLD_SIZE", 1)))
