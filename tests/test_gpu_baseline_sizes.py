"""GPU-vs-oracle parity AT BASELINE.json's sizes (the small-size comparisons live in test_gpu_parity.py):

  configs[1]  PPO CartPole, 4096 envs x 128 steps, 4 epochs x 4 minibatches of 131,072
  configs[3]  PPO Pendulum (Gaussian head), one 8,192-env shard x 128 steps
  configs[2]  A2C CartPole, 16,384 envs x 32 steps, one fused update per rollout
  configs[4]  DQN CartPole, 4096 envs, 1,048,576-transition HBM replay, >= 200 iterations

Every stage of the CUDA path is checked against the CPU oracle on the GPU's own inputs ("teacher forced": a rollout of
half a million chaotic env steps cannot be compared trajectory against trajectory, every STEP of it can): per-step
policy/value outputs, sampled actions against the Philox uniforms, env transitions, bit-exact GAE / return scan, the
raw gradients of the tcgen05 (3xTF32) update kernel per parameter array, post-step parameters and a whole update.
Tolerances are north_star's: integer/flag quantities bit-exact, fp32 quantities rtol 1e-5 per step (atol stated where a
sum cancels); raw 131,072-term gradient sums rtol 1e-4 with atol 3e-6 of the largest element of the array."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import rand_params

pytestmark = pytest.mark.gpu
F = np.float32
RTOL = 1e-5
ALL_FIELDS = ("STATE", "ACTION", "LOGPROB", "REWARD", "TERMINAL", "VALUE", "ADVANTAGE", "RETURN")


def oracle_forward(olib, kind, p, obs, threads=None):
    """olib.policy_forward_raw over row chunks on a thread pool (ctypes releases the GIL)"""
    threads = threads or min(32, os.cpu_count() or 1)
    chunks = np.array_split(np.arange(obs.shape[0]), threads * 4)
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda c: olib.policy_forward_raw(kind, p, np.ascontiguousarray(obs[c])), chunks))
    return [np.concatenate([o[i] for o in outs]) for i in range(3)]


def per_array(olib, kind, g_gpu, g_ref, rtol, atol_of_max, what):
    offs, sizes = olib.param_layout(kind)
    for i, (off, size) in enumerate(zip(offs, sizes)):
        a, b = g_gpu[off:off + size], g_ref[off:off + size]
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_of_max * np.abs(b).max(), err_msg="%s, parameter array %d" % (what, i))


@pytest.mark.parametrize("kind,N", [(0, 4096), (1, 8192)])
def test_ppo_update_at_baseline_size_matches_oracle(crl, olib, abi, torch_cuda, kind, N):
    from cleanrl_jl_b200.handle import PPOHandle
    T, mb, epochs, seed = 128, 4, 4, 11
    olib.set_threads(os.cpu_count() or 1)
    cfg = abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=mb, update_epochs=epochs, seed=seed)
    h, o = PPOHandle(cfg), olib.create(cfg)
    d = olib.dims(kind)
    p = rand_params(olib, kind, seed=4)
    if kind == 1:
        p[-1] = -0.5
    for x in (h, o):
        x.set_params(p)
        x.env_reset()
    np.testing.assert_allclose(h.read_field(abi.CRL_F_ENV_STATE), o.read_field(abi.CRL_F_ENV_STATE), rtol=1e-6, atol=1e-7)

    # ---- rollout: every stored step re-derived by the oracle from the GPU's own previous state
    h.rollout()
    h.gae()
    f = {name: h.read_field(getattr(abi, "CRL_F_" + name)) for name in ALL_FIELDS}
    states, actions, term, rew = f["STATE"], f["ACTION"], f["TERMINAL"], f["REWARD"]
    pol, logp, val = oracle_forward(olib, kind, p, states.reshape(-1, d["D"]))
    np.testing.assert_allclose(f["VALUE"].ravel(), val, rtol=RTOL, atol=2e-6)
    if kind == 0:
        lp_taken = logp[np.arange(N * T), actions.ravel()]
        np.testing.assert_allclose(f["LOGPROB"].ravel(), lp_taken, rtol=RTOL, atol=2e-6)
        # the action is the inverse-CDF sample of the Philox uniform (ppo.jl:26): checked wherever the uniform is not
        # within rounding distance of the threshold
        probs0 = np.exp(logp.astype(np.float64)).reshape(T, N, 2)[..., 0]
        checked = 0
        for t in range(0, T, 5):
            for n in range(t % 7, N, 61):
                u = olib.action_uniform(seed, n, t)
                if abs(u - probs0[t, n]) > 1e-5:
                    assert actions[t, n] == (1 if probs0[t, n] < u else 0), (t, n)
                    checked += 1
        assert checked > 1500
        # transitions (all 127 x 4096 of them): state[t+1] = step(state[t], action[t]) unless an episode boundary intervenes
        for t in range(T - 1):
            ok = (term[t] == 0) & (term[t + 1] == 0)
            s1, _, r1, d1 = olib.env_step_raw(kind, states[t], np.zeros(N, np.int32), actions[t], 500)
            np.testing.assert_allclose(states[t + 1][ok], s1[ok], rtol=RTOL, atol=1e-6)
            assert np.all(rew[t][ok] == 1.0) and np.all(d1[ok] == 0)
            assert np.all(rew[t][(term[t] == 0) & (term[t + 1] == 1)] == 0.0)
        assert term[0].sum() == 0 and term.sum() > N // 8      # Q3, and plenty of episode ends inside the rollout
    else:
        mean = pol.reshape(T, N)
        sd = np.exp(p[-1])
        lp = -(actions[..., 0] - mean) ** 2 / (2 * sd * sd) - p[-1] - 0.9189385332046727
        np.testing.assert_allclose(f["LOGPROB"], lp, rtol=1e-4, atol=1e-5)
        z = (actions[..., 0] - mean) / sd
        assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
        # pendulum reward and next observation from (obs, action): theta from (cos, sin), RLEnvs' cost expression
        th = np.arctan2(states[..., 1].astype(np.float64), states[..., 0].astype(np.float64))
        a = np.clip(actions[..., 0].astype(np.float64), -2, 2)
        cost = th ** 2 + 0.1 * states[..., 2].astype(np.float64) ** 2 + 0.001 * a ** 2
        np.testing.assert_allclose(rew, -cost, rtol=2e-5, atol=2e-5)
    # episode bookkeeping is bit-exact against the terminal flags
    recs, agg = h.pop_episodes()
    nd = h.read_field(abi.CRL_F_NEXT_DONE)
    assert agg.count == term[1:].sum() + nd.sum() and agg.dropped == 0

    # ---- GAE: bit-exact on the GPU's own buffers
    adv_o, ret_o = olib.gae_raw(f["VALUE"], f["REWARD"], f["TERMINAL"], h.read_field(abi.CRL_F_NEXT_VALUE), nd,
                                F(0.99), F(0.95), abi.CRL_GAE_REF_COMPAT)
    np.testing.assert_array_equal(f["ADVANTAGE"], adv_o)
    np.testing.assert_array_equal(f["RETURN"], ret_o)

    # ---- one minibatch of 131,072 (262,144) samples through the tcgen05 kernel: raw gradients per array
    for name in ALL_FIELDS:
        o.write_field(getattr(abi, "CRL_F_" + name), f[name])
    rng = np.random.default_rng(2)
    B, M = N * T, N * T // mb
    perms = np.stack([rng.permutation(B) for _ in range(epochs)]).astype(np.int32)
    lr = float(F(2.5e-4))
    sh = h.update_minibatch(perms[0, :M], lr)
    so = o.update_minibatch(perms[0, :M], lr)
    sh, so = np.array([sh.loss, sh.pg_loss, sh.v_loss, sh.entropy_loss]), np.array([so.loss, so.pg_loss, so.v_loss, so.entropy_loss])
    # pg_loss (and with it loss) is a mean of O(1) terms that cancels to ~1e-3: atol is 1e-6 of a summand
    np.testing.assert_allclose(sh, so, rtol=RTOL, atol=1e-6)
    per_array(olib, kind, h.get_grads(), o.get_grads(), 1e-4, 3e-6, "tcgen05 raw gradient vs oracle")
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=RTOL, atol=1e-6)
    mh, vh, bph = h.get_adam_state()
    mo, vo, bpo = o.get_adam_state()
    np.testing.assert_array_equal(bph, bpo)
    per_array(olib, kind, mh, mo, 1e-4, 3e-6, "Adam m")

    # ---- the whole update (16 more minibatches with Adam steps in between) from there
    sh = h.update_epochs(perms, lr)
    so = o.update_epochs(perms, lr)
    np.testing.assert_allclose(sh, so, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=2e-5, atol=2e-6)
    assert h.spec_replays() == 0
    h.close()
    o.close()


def test_a2c_update_at_baseline_size_matches_oracle(crl, olib, abi, torch_cuda):
    """configs[2]: 16,384 envs x 32 steps through crl_train_update (CUDA graph, tcgen05 kernel, A2C losses)"""
    from cleanrl_jl_b200.handle import PPOHandle
    kind, N, T, seed = 0, 16384, 32, 13
    olib.set_threads(os.cpu_count() or 1)
    cfg = abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=1, update_epochs=1, seed=seed,
                          gae_mode=abi.CRL_GAE_A2C_RETURNS, flags=abi.CRL_FLAG_A2C, gae_lambda=1.0)
    h, o = PPOHandle(cfg), olib.create(cfg)
    p0 = rand_params(olib, kind, seed=2)
    h.set_params(p0)
    h.env_reset()
    # start close to falling so that episodes end (and envs reset) inside the 32 steps
    st = h.read_field(abi.CRL_F_ENV_STATE)
    st[: N // 4, 2] = 0.17
    st[: N // 4, 3] = 1.2
    h.env_set_state(st, np.zeros(N, np.int32))
    lr = 1e-4
    h.train_update(lr)
    sh, agg = h.fetch_update()
    f = {name: h.read_field(getattr(abi, "CRL_F_" + name)) for name in ALL_FIELDS}
    term = f["TERMINAL"]
    assert term[1:].sum() > N // 8 and agg.count >= term[1:].sum()
    # per-step forward under the pre-update parameters
    pol, logp, val = oracle_forward(olib, kind, p0, f["STATE"].reshape(-1, 4))
    np.testing.assert_allclose(f["VALUE"].ravel(), val, rtol=RTOL, atol=2e-6)
    np.testing.assert_allclose(f["LOGPROB"].ravel(), logp[np.arange(N * T), f["ACTION"].ravel()], rtol=RTOL, atol=2e-6)
    # a2c.jl:108 then :52: the observation after a termination is the reset state
    tt, nn = np.nonzero(term[1:])
    assert np.all(np.abs(f["STATE"][tt + 1, nn]) <= 0.05)
    # discounted_future_rewards (a2c.jl:13-24): bit-exact on the GPU's buffers
    adv_o, ret_o = olib.gae_raw(f["VALUE"], f["REWARD"], term, h.read_field(abi.CRL_F_NEXT_VALUE),
                                h.read_field(abi.CRL_F_NEXT_DONE), F(0.99), F(1.0), abi.CRL_GAE_A2C_RETURNS)
    np.testing.assert_array_equal(f["RETURN"], ret_o)
    np.testing.assert_array_equal(f["ADVANTAGE"], adv_o)
    # the fused update on identical inputs
    o.set_params(p0)
    for name in ALL_FIELDS:
        o.write_field(getattr(abi, "CRL_F_" + name), f[name])
    so = o.update_minibatch(np.arange(N * T, dtype=np.int32), lr)
    so = np.array([so.loss, so.pg_loss, so.v_loss, so.entropy_loss])
    np.testing.assert_allclose(sh[-1], so, rtol=2e-5, atol=1e-6)
    per_array(olib, kind, h.get_grads(), o.get_grads(), 1e-4, 3e-6, "A2C raw gradient vs oracle")
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=RTOL, atol=1e-6)
    assert h.spec_replays() == 0
    h.close()
    o.close()


def test_dqn_at_baseline_size_matches_oracle(crl, olib, abi, torch_cuda):
    """configs[4]: 4096 envs, 1M-transition ring, the reference's schedule (learn every 10 iterations after 10,000 steps,
    target copy every 100, batch 120) for 240 iterations = 983,040 transitions, 24 learning steps"""
    from cleanrl_jl_b200.dqn_algo import DQNHandle, init_q_params
    from oracle.oracle import OracleDQN
    N, iters = 4096, 240
    cfg = abi.make_dqn_config(num_envs=N, buffer_size=1 << 20, min_buff_size=10_000, batch_size=120, train_freq=10,
                              target_net_freq=100, epsilon_duration=float(N * 150), seed=9)
    h, o = DQNHandle(cfg), OracleDQN(olib, cfg)
    p = init_q_params(3)
    for x in (h, o):
        x.set_params(p)
        x.reset()
    for chunk in (80, 80, 80):
        sh, so = h.run(chunk), o.run(chunk)
        assert (sh.iterations, sh.learn_steps, sh.episodes) == (so.iterations, so.learn_steps, so.episodes)
        assert sh.epsilon == so.epsilon
        assert sh.sum_return == so.sum_return and sh.sum_length == so.sum_length
        bh, bo = h.read_buffer(), o.read_buffer()
        assert (bh["size"], bh["ptr"]) == (bo["size"], bo["ptr"])
        n = bh["size"]
        np.testing.assert_array_equal(bh["action"][:n], bo["action"][:n])      # epsilon draws, argmax, ring order
        np.testing.assert_array_equal(bh["terminal"][:n], bo["terminal"][:n])
        np.testing.assert_array_equal(bh["reward"][:n], bo["reward"][:n])
        # trajectory level: episodes run up to 200 chained steps without teacher forcing, and CartPole amplifies a rounding
        # difference by ~e^0.08 per step (23 of 2.6 M elements exceed 1e-5 relative, the largest by 2e-5 absolute)
        np.testing.assert_allclose(bh["state"][:n], bo["state"][:n], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(bh["next_state"][:n], bo["next_state"][:n], rtol=1e-3, atol=1e-4)
        # step level (north_star's rtol 1e-5 per step): every stored transition re-derived from the GPU's own state
        s1, _, r1, _ = olib.env_step_raw(0, bh["state"][:n], np.zeros(n, np.int32), bh["action"][:n], 10 ** 6)
        np.testing.assert_allclose(bh["next_state"][:n], s1, rtol=RTOL, atol=1e-6)
        np.testing.assert_array_equal(bh["reward"][:n], r1)
        qh, th = h.get_params()
        qo, to = o.get_params()
        np.testing.assert_allclose(qh, qo, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(th, to, rtol=1e-4, atol=2e-6)
        assert abs(sh.last_loss - so.last_loss) <= 1e-4 * abs(so.last_loss) + 1e-7
    assert so.learn_steps == 24 and so.episodes > 10_000 and bo["size"] == iters * N
    h.close()
    o.close()
