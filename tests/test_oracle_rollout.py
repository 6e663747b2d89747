"""Rollout bookkeeping of the oracle against the reference's cited semantics (SURVEY §4 table:
done flags, episode counters, the stale-obs-after-done behaviour, buffer layout). CPU only."""
import numpy as np
import pytest

from conftest import rand_params

F = np.float32


def make(olib, abi, kind=0, N=8, T=12, **kw):
    cfg = abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=2, update_epochs=1, **kw)
    o = olib.create(cfg)
    p = rand_params(olib, kind, seed=1)
    o.set_params(p)
    return o, p


def test_injected_noise_equals_philox_draws(olib, abi):
    """the two ways of feeding randomness are the same computation"""
    N, T, seed = 8, 12, 42
    o1, p = make(olib, abi, N=N, T=T, seed=seed)
    o2, _ = make(olib, abi, N=N, T=T, seed=seed)
    o1.env_reset(); o2.env_reset()
    an = np.array([[olib.action_uniform(seed, n, t) for n in range(N)] for t in range(T)])
    # reset k=0 was consumed by env_reset; an env's first in-rollout reset uses counter 1 (a second
    # termination cannot happen within 12 steps of the first)
    rn = np.array([[olib.reset_uniforms(seed, n, 1) for n in range(N)] for t in range(T)], F)
    o1.rollout()
    o2.rollout(an, rn)
    for f in (abi.CRL_F_STATE, abi.CRL_F_ACTION, abi.CRL_F_LOGPROB, abi.CRL_F_VALUE, abi.CRL_F_REWARD, abi.CRL_F_TERMINAL,
              abi.CRL_F_ENV_STATE, abi.CRL_F_RESET_COUNT):
        np.testing.assert_array_equal(o1.read_field(f), o2.read_field(f))
    assert o1.read_field(abi.CRL_F_RESET_COUNT).max() <= 2


def test_stale_observation_after_done_and_fresh_obs_at_rollout_start(olib, abi):
    """Q2: after env i terminates at step s, step s+1 stores (and acts on) the terminal observation
    with terminal=true while the env itself was reset. Q3: the next rollout starts from the
    refreshed state with terminal=false."""
    N, T = 4, 10
    o, p = make(olib, abi, N=N, T=T, seed=3)
    st = np.zeros((N, 4), F)
    st[1] = [0.0, 0.0, 0.205, 2.0]   # falls over on the first step
    st[2] = [2.39, 3.0, 0.0, 0.0]    # leaves the track on the first step
    o.env_set_state(st, np.array([0, 0, 0, 499], np.int32))  # env 3 hits max_steps at its 2nd step
    rn = np.random.default_rng(0).random((T, N, 4)).astype(F)
    an = np.full((T, N), 0.5)
    o.rollout(an, rn)
    term = o.read_field(abi.CRL_F_TERMINAL)
    states = o.read_field(abi.CRL_F_STATE)
    rew = o.read_field(abi.CRL_F_REWARD)
    assert list(term[0]) == [0, 0, 0, 0]            # Q3
    assert list(term[1]) == [0, 1, 1, 0]
    assert term[2, 3] == 1 and term[1, 3] == 0      # t=500 ok, t=501 > max_steps
    np.testing.assert_array_equal(states[0], st)    # first stored obs = current state
    assert abs(states[1, 1, 2]) > 0.2094395 and abs(states[1, 2, 0]) > 2.4  # stale terminal obs (Q2)
    assert rew[0, 1] == 0.0 and rew[0, 2] == 0.0 and rew[0, 0] == 1.0       # reward 0 on the terminal step
    # the env was reset from the injected noise of the step that terminated it, then stepped once
    s_reset, t_reset = olib.env_reset_raw(abi.CRL_ENV_CARTPOLE, rn[0, 1:3])
    acts = o.read_field(abi.CRL_F_ACTION)
    s_next, _, _, _ = olib.env_step_raw(abi.CRL_ENV_CARTPOLE, s_reset, t_reset, acts[1, 1:3], 500)
    np.testing.assert_array_equal(states[2, 1:3], s_next)   # the reset state itself is never observed
    # episode records in (step, env) order with the reference's counters (ppo.jl:124-125,145-162)
    recs, agg = o.pop_episodes()
    assert [(r[0], r[1]) for r in recs][:3] == [(0, 1), (0, 2), (1, 3)]
    assert recs[0][2] == 1 and recs[0][3] == 0.0    # length 1, return 0 (terminal step pays 0)
    assert recs[2][2] == 2 and recs[2][3] == 1.0
    assert agg.count == len(recs)
    # reset counters
    rc = o.read_field(abi.CRL_F_RESET_COUNT)
    assert rc[1] >= 1 and rc[2] >= 1 and rc[3] >= 1
    # episode lengths: steps since the env's last termination (ppo.jl:125,159)
    el = o.read_field(abi.CRL_F_EP_LENGTH)
    last_done = [max([r[0] for r in recs if r[1] == n], default=-1) for n in range(N)]
    assert list(el) == [T - 1 - ld for ld in last_done]
    # second rollout: Q3 again
    env_state_before = o.read_field(abi.CRL_F_ENV_STATE)
    o.rollout(an, rn)
    term2 = o.read_field(abi.CRL_F_TERMINAL)
    assert term2[0].sum() == 0
    # the first observation of a rollout is the refreshed env state (ppo.jl:169), not the stale obs
    np.testing.assert_array_equal(o.read_field(abi.CRL_F_STATE)[0], env_state_before)


def test_buffer_layout_and_flat_index(olib, abi):
    """element (d,n,t) of `state` at d + D*n + D*N*t; scalars at n + N*t (replay_buffer.jl:16,28; ppo.jl:184-189)"""
    N, T = 5, 6
    o, p = make(olib, abi, N=N, T=T, seed=9)
    o.env_reset()
    o.rollout()
    s = o.read_field(abi.CRL_F_STATE)
    assert s.shape == (T, N, 4)
    flat = s.ravel()
    assert flat[2 + 4 * 3 + 4 * N * 4] == s[4, 3, 2]
    v = o.read_field(abi.CRL_F_VALUE)
    assert v.ravel()[3 + N * 4] == v[4, 3]
    # minibatch slices are contiguous slices of one permutation per epoch (ppo.jl:191-204)
    o.gae()
    perm = np.random.default_rng(0).permutation(N * T).astype(np.int32)[None]
    stats = o.update_epochs(perm, 1e-3)
    assert stats.shape == (2, 4)


@pytest.mark.parametrize("kind", [0, 1])
def test_train_update_runs_and_learning_signal(olib, abi, kind):
    """a few oracle updates: losses finite, parameters move, every array gets a gradient"""
    o, p = make(olib, abi, kind=kind, N=16, T=32, seed=5)
    o.env_reset()
    for u in range(3):
        st = o.train_update(2.5e-4)
        assert np.all(np.isfinite(st))
    p2 = o.get_params()
    assert np.abs(p2 - p).max() > 1e-4
    off, size = olib.param_layout(kind)
    g = o.get_grads()
    assert all(np.abs(g[a:a + b]).max() > 0 for a, b in zip(off, size))
    # per-array β powers advanced once per minibatch (Flux keeps them per parameter array)
    _, _, bp = o.get_adam_state()
    np.testing.assert_allclose(bp[:, 0], 0.9 ** (1 + 3 * 2), rtol=1e-12)


def test_oracle_rejects_bad_configs(olib, abi):
    with pytest.raises(ValueError):
        olib.create(abi.make_config(num_envs=3, num_steps=5, num_minibatches=2))  # Q10
    o = olib.create(abi.make_config())
    with pytest.raises(RuntimeError):
        o.gae()  # before any rollout
