"""A2C widening: the return scan `discounted_future_rewards` (a2c.jl:13-24) and the two loss expressions
(a2c.jl:78-97). CPU: oracle vs a hand-derived known answer, an independent NumPy restatement and float64 autograd.
GPU: the CUDA path through the C ABI vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import rand_params
from test_oracle_grad import TanhFast, make_batch

F = np.float32


def discounted_future_rewards(rewards, terminals, final_value, gamma):
    """line-by-line NumPy restatement of a2c.jl:13-24 (1-based indices turned 0-based)"""
    n = len(rewards)
    fut = np.zeros(n)
    fut[0] = 0.0 if terminals[-1] else rewards[-1] + gamma * final_value
    rr, tt = rewards[:-1][::-1], terminals[:-1][::-1]
    for i, (r, t) in enumerate(zip(rr, tt)):
        fut[i + 1] += 0.0 if t else r + gamma * fut[i]
    return fut[::-1]


def test_returns_known_answer_and_restatement(olib, abi):
    # hand-derived: rewards [1,1,0,1], terminals after each step [F,F,T,F], final_value 0.5, γ = 0.99
    #   future = [1 + 0.99*1, 1 + 0.99*0, 0, 1 + 0.99*0.5] = [1.99, 1.0, 0.0, 1.495]
    v = np.array([[0.1], [0.2], [0.3], [0.4]], F)
    r = np.array([[1], [1], [0], [1]], F)
    enter = np.array([[0], [0], [0], [1]], np.uint8)  # flag entering step t = terminal after step t-1
    adv, ret = olib.gae_raw(v, r, enter, np.array([0.5], F), np.array([0], np.uint8), F(0.99), F(1.0), abi.CRL_GAE_A2C_RETURNS)
    np.testing.assert_allclose(ret[:, 0], [1.99, 1.0, 0.0, 1.495], rtol=1e-6)
    np.testing.assert_allclose(adv[:, 0], ret[:, 0] - v[:, 0], rtol=1e-6)
    rng = np.random.default_rng(0)
    T, N = 40, 9
    v = rng.standard_normal((T, N)).astype(F)
    r = rng.standard_normal((T, N)).astype(F)
    after = rng.random((T, N)) < 0.15   # is_terminated after step t
    enter = np.zeros((T, N), np.uint8)
    enter[1:] = after[:-1]
    fv = rng.standard_normal(N).astype(F)
    adv, ret = olib.gae_raw(v, r, enter, fv, after[-1].astype(np.uint8), F(0.99), F(1.0), abi.CRL_GAE_A2C_RETURNS)
    for n in range(N):
        exp = discounted_future_rewards(r[:, n].astype(np.float64), after[:, n], float(fv[n]), float(F(0.99)))
        np.testing.assert_allclose(ret[:, n], exp, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("kind", [0, 1])
def test_a2c_loss_gradient_matches_float64_autograd(olib, kind):
    d = olib.dims(kind)
    off, size = olib.param_layout(kind)
    B, M = 200, 64
    p = rand_params(olib, kind, seed=7)
    if kind == 1:
        p[-1] = -0.2
    states, actions, _, _, ret, _ = make_batch(olib, kind, B, 3)
    idx = np.random.default_rng(1).permutation(B)[:M].astype(np.int32)
    g, st = olib.a2c_loss_raw(kind, p, idx, states, actions, ret)
    pt = torch.tensor(p.astype(np.float64), requires_grad=True)
    D, A = d["D"], d["A"]
    x = torch.tensor(states.astype(np.float64))[idx].T

    def net(base, out):
        W1 = pt[off[base]:off[base] + size[base]].reshape(D, 64).T
        b1 = pt[off[base + 1]:off[base + 1] + 64]
        W2 = pt[off[base + 2]:off[base + 2] + 4096].reshape(64, 64).T
        b2 = pt[off[base + 3]:off[base + 3] + 64]
        W3 = pt[off[base + 4]:off[base + 4] + 64 * out].reshape(64, out).T
        b3 = pt[off[base + 5]:off[base + 5] + out]
        return W3 @ TanhFast.apply(W2 @ TanhFast.apply(W1 @ x + b1[:, None]) + b2[:, None]) + b3[:, None]

    z, values = net(0, A), net(6, 1)[0]
    R = torch.tensor(ret.astype(np.float64))[idx]
    advantage = R - values                       # a2c.jl:83
    critic_loss = (advantage ** 2).mean()        # a2c.jl:84
    if kind == 0:
        lp = torch.log_softmax(z, 0)[torch.tensor(actions[idx], dtype=torch.int64), torch.arange(M)]
    else:
        logstd = pt[off[12]:off[12] + A]
        a = torch.tensor(actions.astype(np.float64)).reshape(-1, A)[idx].T
        lp = (-(a - z) ** 2 / (2 * torch.exp(logstd)[:, None] ** 2) - logstd[:, None] - 0.9189385332046727).sum(0)
    actor_loss = -(lp * advantage.detach()).mean()  # a2c.jl:95 (advantage is a captured constant there)
    (actor_loss + critic_loss).backward()
    np.testing.assert_allclose(st[1:3], [actor_loss.item(), critic_loss.item()], rtol=2e-5, atol=1e-6)
    gt = pt.grad.numpy()
    np.testing.assert_allclose(g, gt, rtol=2e-3, atol=2e-5 * np.abs(gt).max())


@pytest.mark.parametrize("a2c_flag", [True, False])
def test_a2c_rollout_observes_the_reset_state_after_a_termination(olib, abi, a2c_flag):
    """a2c.jl:108 resets the env and a2c.jl:52 then copies state(env): the transition after a termination starts from
    the RESET state. PPO (ppo.jl:143 before :164, quirk Q2) keeps the stale terminal observation instead."""
    N, T = 48, 24
    kw = dict(num_minibatches=1, update_epochs=1, gae_mode=abi.CRL_GAE_A2C_RETURNS, flags=abi.CRL_FLAG_A2C, gae_lambda=1.0) \
        if a2c_flag else dict(num_minibatches=4, update_epochs=1)
    o = olib.create(abi.make_config(env_kind=0, num_envs=N, num_steps=T, seed=3, **kw))
    from conftest import rand_params
    o.set_params(rand_params(olib, 0, seed=2))
    rng = np.random.default_rng(5)
    st = (rng.random((N, 4)) * 0.1 - 0.05).astype(F)
    st[:, 2] = rng.uniform(0.15, 0.2, N)   # about to fall
    st[:, 3] = 1.5
    o.env_set_state(st, np.zeros(N, np.int32))
    o.rollout(rng.random((T, N)), rng.random((T, N, 4)).astype(F))
    term, states = o.read_field(abi.CRL_F_TERMINAL), o.read_field(abi.CRL_F_STATE)
    tt, nn = np.nonzero(term[1:])
    assert len(tt) > 20
    after = states[tt + 1, nn]
    fresh = np.all(np.abs(after) <= 0.05, axis=1)   # reset!: 0.1 rand - 0.05 per component
    fallen = (np.abs(after[:, 0]) > 2.4) | (np.abs(after[:, 2]) > 0.2094395)
    if a2c_flag:
        assert fresh.all() and not fallen.any()
    else:
        assert fallen.all() and not fresh.any()


def test_a2c_flag_is_rejected_with_several_minibatches_or_epochs(olib, abi):
    """the recorded values are the current critic's output for the first minibatch only (a2c.jl:79-85)"""
    from cleanrl_jl_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    for bad in (dict(num_minibatches=2, update_epochs=1, gae_mode=abi.CRL_GAE_A2C_RETURNS),
                dict(num_minibatches=1, update_epochs=2, gae_mode=abi.CRL_GAE_A2C_RETURNS),
                dict(num_minibatches=1, update_epochs=1, gae_mode=abi.CRL_GAE_FIXED)):
        cfg = abi.make_config(env_kind=0, num_envs=8, num_steps=8, flags=abi.CRL_FLAG_A2C, **bad)
        h = C.c_void_p()
        assert lib.crl_create(C.byref(cfg), C.byref(h)) == abi.CRL_ERR_INVALID
        assert b"CRL_FLAG_A2C" in lib.crl_last_error()


def test_a2c_config_defaults_match_a2c_jl():
    from cleanrl_jl_b200.a2c_algo import A2CConfig, make_crl_config
    c = A2CConfig()
    assert (c.lr, c.total_timesteps, c.min_replay_size, c.gamma) == (0.0001, 1_000_000, 512, 0.99)  # a2c.jl:4-9
    cfg = make_crl_config(c)
    assert (cfg.gae_mode, cfg.flags, cfg.num_minibatches, cfg.update_epochs, cfg.clip_norm) == (2, 2, 1, 1, 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1])
def test_a2c_cuda_path_matches_oracle(crl, olib, abi, torch_cuda, kind):
    from cleanrl_jl_b200.handle import PPOHandle
    N, T = 96, 20
    cfg = abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=1, update_epochs=1, seed=13,
                          gae_mode=abi.CRL_GAE_A2C_RETURNS, flags=abi.CRL_FLAG_A2C, gae_lambda=1.0)
    h, o = PPOHandle(cfg), olib.create(cfg)
    p = rand_params(olib, kind, seed=2)
    if kind == 1:
        p[-1] = -0.3
    for x in (h, o):
        x.set_params(p)
        x.env_reset()
    # return scan: bit-exact on identical inputs (raw entry)
    torch = torch_cuda
    rng = np.random.default_rng(5)
    v = rng.standard_normal((T, N)).astype(F); r = rng.standard_normal((T, N)).astype(F)
    d = (rng.random((T, N)) < 0.1).astype(np.uint8); nv = rng.standard_normal(N).astype(F); nd = (rng.random(N) < 0.1).astype(np.uint8)
    adv_o, ret_o = olib.gae_raw(v, r, d, nv, nd, F(0.99), F(1.0), abi.CRL_GAE_A2C_RETURNS)
    t = [torch.from_numpy(a).cuda() for a in (v, r, d, nv, nd)]
    adv, ret = torch.zeros((T, N), device="cuda"), torch.zeros((T, N), device="cuda")
    crl.check(crl.load().crl_gae_raw(*[crl.ptr(a) for a in t], crl.ptr(adv), crl.ptr(ret), T, N, 0.99, 1.0, abi.CRL_GAE_A2C_RETURNS, None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ret.cpu().numpy(), ret_o)
    np.testing.assert_array_equal(adv.cpu().numpy(), adv_o)
    # three full updates (rollout + returns + combined step) with the shared Philox streams
    for u in range(3):
        h.train_update(1e-4)
        sh, _ = h.fetch_update()
        so = o.train_update(1e-4)
        np.testing.assert_allclose(sh, so, rtol=5e-4, atol=1e-5, err_msg="update %d" % u)
        np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=1e-4, atol=2e-6)
        np.testing.assert_array_equal(h.read_field(abi.CRL_F_TERMINAL), o.read_field(abi.CRL_F_TERMINAL))
    assert h.spec_replays() == 0
    # gradient on identical inputs
    for name in ("STATE", "ACTION", "RETURN", "ADVANTAGE", "VALUE", "LOGPROB", "REWARD", "TERMINAL"):
        o.write_field(getattr(abi, "CRL_F_" + name), h.read_field(getattr(abi, "CRL_F_" + name)))
    o.set_params(h.get_params())
    m, vv, bp = h.get_adam_state()
    o.set_adam_state(m, vv, bp)
    idx = np.arange(N * T, dtype=np.int32)
    h.update_minibatch(idx, 1e-4); o.update_minibatch(idx, 1e-4)
    gh, go = h.get_grads(), o.get_grads()
    np.testing.assert_allclose(gh, go, rtol=1e-3, atol=2e-5 * np.abs(go).max())
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=1e-5, atol=1e-6)
    h.close()
