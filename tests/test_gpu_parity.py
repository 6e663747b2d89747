"""GPU parity tests: every CUDA kernel, called through the C ABI (libcleanrl_cuda.so), against
the CPU oracle on identical seeded inputs. Tolerances (north_star): done flags, episode
counters, actions and permutations bit-exact; GAE bit-exact (Float64 recurrence, no FMA
contraction); everything else within fp32 tolerance rtol 1e-5 per step (atol stated per test).
Run with `pytest -m gpu` on a B200."""
import ctypes as C

import numpy as np
import pytest

from conftest import rand_params

pytestmark = pytest.mark.gpu
F = np.float32
RTOL = 1e-5


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------ raw kernels
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("T,N", [(1, 5), (2, 33), (32, 4), (128, 4096), (37, 1001), (16, 148 * 1024 * 4)])
def test_gae_raw_bit_exact(crl, olib, torch_cuda, mode, T, N):
    torch = torch_cuda
    rng = np.random.default_rng(1234 + T + N)
    v = rng.standard_normal((T, N)).astype(F)
    r = rng.standard_normal((T, N)).astype(F)
    d = (rng.random((T, N)) < 0.05).astype(np.uint8)
    nv = rng.standard_normal(N).astype(F)
    nd = (rng.random(N) < 0.05).astype(np.uint8)
    adv_o, ret_o = olib.gae_raw(v, r, d, nv, nd, F(0.99), F(0.95), mode)
    tv, tr, td, tnv, tnd = [dev(torch, a) for a in (v, r, d, nv, nd)]
    adv = torch.full((T, N), float("nan"), dtype=torch.float32, device="cuda")
    ret = torch.full((T, N), float("nan"), dtype=torch.float32, device="cuda")
    crl.check(crl.load().crl_gae_raw(crl.ptr(tv), crl.ptr(tr), crl.ptr(td), crl.ptr(tnv), crl.ptr(tnd), crl.ptr(adv),
                                     crl.ptr(ret), T, N, 0.99, 0.95, mode, None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(adv.cpu().numpy(), adv_o)
    np.testing.assert_array_equal(ret.cpu().numpy(), ret_o)


def test_gae_raw_empty_and_errors(crl, torch_cuda):
    lib = crl.load()
    t = torch_cuda.zeros(4, device="cuda")
    assert lib.crl_gae_raw(crl.ptr(t), crl.ptr(t), crl.ptr(t), None, None, crl.ptr(t), crl.ptr(t), 4, 0, 0.99, 0.95, 0, None) == 0
    assert lib.crl_gae_raw(crl.ptr(t), crl.ptr(t), crl.ptr(t), None, None, crl.ptr(t), crl.ptr(t), 0, 1, 0.99, 0.95, 0, None) == -1
    assert lib.crl_gae_raw(crl.ptr(t), crl.ptr(t), crl.ptr(t), None, None, crl.ptr(t), crl.ptr(t), 4, 1, 0.99, 0.95, 1, None) == -1
    assert b"bootstrap" in lib.crl_last_error()
    assert lib.crl_gae_raw(None, crl.ptr(t), crl.ptr(t), None, None, crl.ptr(t), crl.ptr(t), 4, 1, 0.99, 0.95, 0, None) == -1


@pytest.mark.parametrize("kind", [0, 1])
def test_env_step_raw(crl, olib, torch_cuda, kind):
    torch = torch_cuda
    rng = np.random.default_rng(kind)
    n = 100_003
    S = olib.dims(kind)["S"]
    if kind == 0:
        s0 = (rng.random((n, 4)) * np.array([5.0, 4.0, 0.5, 4.0]) - np.array([2.5, 2.0, 0.25, 2.0])).astype(F)
        a = rng.integers(0, 2, n).astype(np.int32)
        t0 = rng.integers(0, 505, n).astype(np.int32)
        ms = 500
    else:
        s0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-8, 8, n)], 1).astype(F)
        a = rng.uniform(-3, 3, n).astype(F)
        t0 = rng.integers(0, 203, n).astype(np.int32)
        ms = 200
    so, to, ro, do = olib.env_step_raw(kind, s0, t0, a, ms)
    ts, tt, ta = dev(torch, s0), dev(torch, t0), dev(torch, a)
    tr = torch.zeros(n, dtype=torch.float32, device="cuda")
    td = torch.zeros(n, dtype=torch.uint8, device="cuda")
    crl.check(crl.load().crl_env_step_raw(kind, crl.ptr(ts), crl.ptr(tt), crl.ptr(ta), crl.ptr(tr), crl.ptr(td), n, ms, None))
    torch.cuda.synchronize()
    sg, dg = ts.cpu().numpy(), td.cpu().numpy()
    np.testing.assert_array_equal(tt.cpu().numpy(), to)
    np.testing.assert_allclose(sg, so, rtol=RTOL, atol=1e-6)
    # done flags are bit-exact except where a state lands within rounding distance of a threshold
    if kind == 0:
        near = (np.abs(np.abs(so[:, 0]) - 2.4) < 1e-5) | (np.abs(np.abs(so[:, 2]) - 0.20943952) < 1e-6)
    else:
        near = np.zeros(n, bool)
    assert near.sum() < 20
    np.testing.assert_array_equal(dg[~near], do[~near])
    np.testing.assert_allclose(tr.cpu().numpy()[~near], ro[~near], rtol=RTOL, atol=1e-6)
    assert 0.005 < do.mean() < 0.98  # both outcomes exercised


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 5000])
def test_policy_forward_raw(crl, olib, torch_cuda, kind, n):
    torch = torch_cuda
    d = olib.dims(kind)
    p = rand_params(olib, kind, seed=5)
    obs = (np.random.default_rng(n).standard_normal((n, d["D"])) * np.array([1.0, 1.5, 0.1, 1.5])[:d["D"]]).astype(F)
    pol_o, logp_o, val_o = olib.policy_forward_raw(kind, p, obs)
    tp, tobs = dev(torch, p), dev(torch, obs)
    pol = torch.zeros((n, d["A"]), dtype=torch.float32, device="cuda")
    logp = torch.zeros((n, d["A"]), dtype=torch.float32, device="cuda")
    val = torch.zeros(n, dtype=torch.float32, device="cuda")
    crl.check(crl.load().crl_policy_forward_raw(kind, crl.ptr(tp), crl.ptr(tobs), crl.ptr(pol), crl.ptr(logp), crl.ptr(val), n, None))
    torch.cuda.synchronize()
    # the test parameters scale the actor head x30 (logits up to +-20), so the absolute error of a
    # 64-term fp32 dot product is ~1e-6 * 20; atol is set relative to that magnitude
    np.testing.assert_allclose(pol.cpu().numpy(), pol_o, rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(val.cpu().numpy(), val_o, rtol=RTOL, atol=2e-6)
    if kind == 0:
        np.testing.assert_allclose(logp.cpu().numpy(), logp_o, rtol=RTOL, atol=1e-5)
        assert np.abs(pol_o[:, 0] - pol_o[:, 1]).max() > 0.05  # non-degenerate policy


def _loss_inputs(olib, kind, B, seed, small_returns=False):
    from test_oracle_grad import make_batch
    return make_batch(olib, kind, B, seed, small_returns)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("M,small_returns", [(2, False), (32, False), (127, False), (128, True), (129, False),
                                             (4096, False), (20000, True)])
def test_ppo_loss_raw(crl, olib, torch_cuda, kind, M, small_returns):
    torch = torch_cuda
    d = olib.dims(kind)
    B = max(2 * M, 64)
    p = rand_params(olib, kind, seed=9)
    if kind == 1:
        p[-1] = -0.3
    if small_returns:
        p[olib.param_layout(kind)[0][11]] = 1.5
    states, actions, logprobs, adv, ret, val = _loss_inputs(olib, kind, B, 21 + M, small_returns)
    idx = np.random.default_rng(M).permutation(B)[:M].astype(np.int32)
    c, ent_c, v_c = float(F(0.2)), float(F(0.01)), float(F(0.5))
    olib.set_threads(8)
    g_o, st_o, vnew_o = olib.ppo_loss_raw(kind, p, idx, states, actions, logprobs, adv, ret, val, c, ent_c, v_c)
    olib.set_threads(1)
    t = [dev(torch, a) for a in (p, idx, states, actions, logprobs, adv, ret, val)]
    g = torch.zeros(d["P"], dtype=torch.float32, device="cuda")
    st = torch.zeros(4, dtype=torch.float64, device="cuda")
    crl.check(crl.load().crl_ppo_loss_raw(kind, crl.ptr(t[0]), crl.ptr(t[1]), M, *[crl.ptr(x) for x in t[2:]],
                                          c, ent_c, v_c, crl.ptr(g), crl.ptr(st), None))
    torch.cuda.synchronize()
    np.testing.assert_allclose(st.cpu().numpy(), st_o, rtol=RTOL, atol=1e-7)
    gg = g.cpu().numpy()
    off, size = olib.param_layout(kind)
    for i in range(d["n_arrays"]):
        sl = slice(off[i], off[i] + size[i])
        scale = np.abs(g_o[sl]).max()
        np.testing.assert_allclose(gg[sl], g_o[sl], rtol=1e-4, atol=2e-6 * scale + 1e-12, err_msg="array %d" % i)
    if small_returns:
        s = np.mean(vnew_o - ret[idx] ** 2)
        dv = np.clip(vnew_o - val[idx], -c, c)
        assert np.sum(s > (val[idx] + dv - ret[idx]) ** 2) > 0  # Q5 count path exercised


@pytest.mark.parametrize("kind", [0, 1])
def test_clip_adam_raw(crl, olib, torch_cuda, kind):
    torch = torch_cuda
    d = olib.dims(kind)
    rng = np.random.default_rng(4)
    p = rng.standard_normal(d["P"]).astype(F)
    m = np.zeros(d["P"], F)
    v = np.zeros(d["P"], F)
    bp = np.tile(np.array([0.9, 0.999]), (d["n_arrays"], 1))
    tp, tm, tv, tbp = dev(torch, p), dev(torch, m), dev(torch, v), dev(torch, bp)
    lr = float(F(2.5e-4)) * 0.731
    for it in range(3):
        g = (rng.standard_normal(d["P"]) * (0.002 if it == 1 else 0.05)).astype(F)
        p, m, v, bp = olib.clip_adam_raw(kind, p, g, m, v, bp, lr, 0.5)
        tg = dev(torch, g)
        crl.check(crl.load().crl_clip_adam_raw(kind, crl.ptr(tp), crl.ptr(tg), crl.ptr(tm), crl.ptr(tv), crl.ptr(tbp), lr, 0.5, None))
        torch.cuda.synchronize()
        # Float64 scalar arithmetic with _rn intrinsics: expected bit-exact; allow 1 ulp for sqrt/div
        np.testing.assert_allclose(tm.cpu().numpy(), m, rtol=1e-7, atol=0)
        np.testing.assert_allclose(tv.cpu().numpy(), v, rtol=1e-7, atol=0)
        np.testing.assert_allclose(tp.cpu().numpy(), p, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(tbp.cpu().numpy(), bp, rtol=1e-15)


# ------------------------------------------------------------------ handle API
def make_pair(crl, olib, abi, kind=0, N=64, T=16, mb=4, epochs=2, seed=3, gae_mode=0, params_seed=1, **kw):
    from cleanrl_jl_b200.handle import PPOHandle
    cfg = abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=mb, update_epochs=epochs, seed=seed,
                          gae_mode=gae_mode, **kw)
    h = PPOHandle(cfg)
    o = olib.create(cfg)
    p = rand_params(olib, kind, seed=params_seed)
    if kind == 1:
        p[-1] = -0.5
    h.set_params(p)
    o.set_params(p)
    return h, o


BUF_FIELDS = ["STATE", "ACTION", "LOGPROB", "REWARD", "TERMINAL", "VALUE"]
EXACT = {"ACTION", "TERMINAL", "REWARD"}


def compare_buffers(h, o, abi, kind, fields=BUF_FIELDS, rtol=RTOL, atol=2e-6):
    for name in fields:
        f = getattr(abi, "CRL_F_" + name)
        a, b = h.read_field(f), o.read_field(f)
        if name in EXACT and not (kind == 1 and name in ("ACTION", "REWARD")):
            np.testing.assert_array_equal(a, b, err_msg=name)
        else:
            np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=name)


@pytest.mark.parametrize("kind", [0, 1])
def test_env_reset_matches_oracle(crl, olib, abi, torch_cuda, kind):
    h, o = make_pair(crl, olib, abi, kind, N=100, T=4, mb=1)
    h.env_reset(); o.env_reset()
    for f in (abi.CRL_F_ENV_STATE, abi.CRL_F_NEXT_OBS):
        np.testing.assert_allclose(h.read_field(f), o.read_field(f), rtol=1e-6, atol=1e-7)
    for f in (abi.CRL_F_ENV_T, abi.CRL_F_RESET_COUNT, abi.CRL_F_NEXT_DONE, abi.CRL_F_EP_LENGTH):
        np.testing.assert_array_equal(h.read_field(f), o.read_field(f))
    if kind == 0:
        s = h.read_field(abi.CRL_F_ENV_STATE)
        assert np.all(s >= -0.05) and np.all(s < 0.05) and s.std() > 0.02


@pytest.mark.parametrize("kind", [0, 1])
def test_rollout_injected_noise_matches_oracle(crl, olib, abi, torch_cuda, kind):
    """identical injected uniforms / normals / reset noise; short horizon so fp32 differences stay small"""
    N, T = 96, 24
    h, o = make_pair(crl, olib, abi, kind, N=N, T=T, mb=4)
    rng = np.random.default_rng(8)
    d = olib.dims(kind)
    if kind == 0:
        st = (rng.random((N, 4)) * 0.1 - 0.05).astype(F)
        st[:N // 3, 2] = rng.uniform(0.15, 0.2, N // 3)  # about to fall: terminations + resets inside the rollout (Q2)
        st[:N // 3, 3] = 1.5
        t0 = np.zeros(N, np.int32)
        t0[-8:] = 495  # hits max_steps=500 mid-rollout
        an = rng.random((T, N))
    else:
        st = np.stack([rng.uniform(-3, 3, N), rng.uniform(-1, 1, N)], 1).astype(F)
        t0 = np.zeros(N, np.int32)
        t0[:N // 2] = rng.integers(180, 199, N // 2)  # done at t >= 200
        an = rng.standard_normal((T, N, d["A"]))
    rn = rng.random((T, N, 4)).astype(F)
    for x in (h, o):
        x.env_set_state(st, t0)
        x.rollout(an, rn)
    # 24 chained steps: fp32 differences compound through the dynamics, so this trajectory-level
    # comparison uses rtol 1e-4 / atol 2e-5; per-step parity at rtol 1e-5 is the teacher-forced test
    compare_buffers(h, o, abi, kind, rtol=1e-4, atol=2e-5)
    term = o.read_field(abi.CRL_F_TERMINAL)
    assert term[0].sum() == 0 and term[1:].sum() > 10  # Q3 and at least some terminations
    for f in (abi.CRL_F_ENV_T, abi.CRL_F_RESET_COUNT, abi.CRL_F_NEXT_DONE, abi.CRL_F_EP_LENGTH):
        np.testing.assert_array_equal(h.read_field(f), o.read_field(f))
    for f in (abi.CRL_F_ENV_STATE, abi.CRL_F_NEXT_OBS, abi.CRL_F_EP_RETURN):
        np.testing.assert_allclose(h.read_field(f), o.read_field(f), rtol=1e-4, atol=2e-5)
    # episode records: same episodes in the reference's (step, env) logging order (Q11)
    rh, ah = h.pop_episodes()
    ro, ao = o.pop_episodes()
    assert [(r[0], r[1], r[2]) for r in rh] == [(r[0], r[1], r[2]) for r in ro]
    np.testing.assert_allclose([r[3] for r in rh], [r[3] for r in ro], rtol=1e-4)
    assert ah.count == ao.count == len(ro) and ah.dropped == 0
    np.testing.assert_allclose([ah.sum_return, ah.sum_length, ah.max_return], [ao.sum_return, ao.sum_length, ao.max_return], rtol=1e-4)
    # Q2: the observation stored after a termination is the stale terminal observation
    if kind == 0:
        states = h.read_field(abi.CRL_F_STATE)
        tt, nn = np.nonzero(term[1:])
        assert np.all((np.abs(states[tt + 1, nn, 0]) > 2.4) | (np.abs(states[tt + 1, nn, 2]) > 0.2094395) | (t0[nn] > 400))
    # GAE on top (both modes): bit-exact given the oracle's own buffers
    for x in (h, o):
        x.gae()
    np.testing.assert_allclose(h.read_field(abi.CRL_F_NEXT_VALUE), o.read_field(abi.CRL_F_NEXT_VALUE), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(h.read_field(abi.CRL_F_ADVANTAGE), o.read_field(abi.CRL_F_ADVANTAGE), rtol=1e-4, atol=5e-5)
    h.close()


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("N", [100, 148 * 32 + 20])
def test_rollout_ragged_env_counts(crl, olib, abi, torch_cuda, kind, N):
    """env counts that are not a multiple of the 32 envs a CTA owns: one grid smaller than the GPU (100 envs, last CTA 4
    of 32 lanes) and one larger (4756 envs: the two-CTAs-per-SM variant of the rollout kernel, last CTA 20 lanes).
    Device RNG on both sides, short horizon, high pole velocity so that episodes end and reset inside the rollout."""
    T = 8
    h, o = make_pair(crl, olib, abi, kind, N=N, T=T, mb=4, seed=29)
    rng = np.random.default_rng(5)
    if kind == 0:
        st = (rng.random((N, 4)) * 0.1 - 0.05).astype(F)
        st[::3, 2] = 0.19
        st[::3, 3] = 2.0
        t0 = np.zeros(N, np.int32)
    else:
        st = np.stack([rng.uniform(-3, 3, N), rng.uniform(-1, 1, N)], 1).astype(F)
        t0 = np.zeros(N, np.int32)
        t0[::2] = 196  # done at t >= 200
    for x in (h, o):
        x.env_set_state(st, t0)
        x.rollout()
    compare_buffers(h, o, abi, kind, rtol=1e-4, atol=2e-5)
    term = o.read_field(abi.CRL_F_TERMINAL)
    assert term[1:].sum() > N // 8
    for f in (abi.CRL_F_ENV_T, abi.CRL_F_RESET_COUNT, abi.CRL_F_NEXT_DONE, abi.CRL_F_EP_LENGTH):
        np.testing.assert_array_equal(h.read_field(f), o.read_field(f))
    for f in (abi.CRL_F_ENV_STATE, abi.CRL_F_NEXT_OBS, abi.CRL_F_EP_RETURN):
        np.testing.assert_allclose(h.read_field(f), o.read_field(f), rtol=1e-4, atol=2e-5)
    rh, ah = h.pop_episodes()
    ro, ao = o.pop_episodes()
    assert [(r[0], r[1], r[2]) for r in rh] == [(r[0], r[1], r[2]) for r in ro]
    for x in (h, o):
        x.gae()
    np.testing.assert_allclose(h.read_field(abi.CRL_F_NEXT_VALUE), o.read_field(abi.CRL_F_NEXT_VALUE), rtol=1e-4, atol=2e-5)
    h.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_rollout_philox_teacher_forced(crl, olib, abi, torch_cuda, kind):
    """device RNG, long horizon: every stored step is re-derived by the oracle from the GPU's own
    previous state (per-step parity, immune to chaotic divergence)"""
    N, T = 256, 128
    h, o = make_pair(crl, olib, abi, kind, N=N, T=T, mb=4, seed=11)
    h.env_reset(); o.env_reset()
    h.rollout()
    p = h.get_params()
    d = olib.dims(kind)
    states = h.read_field(abi.CRL_F_STATE)
    actions = h.read_field(abi.CRL_F_ACTION)
    term = h.read_field(abi.CRL_F_TERMINAL)
    rew = h.read_field(abi.CRL_F_REWARD)
    pol, logp, val = olib.policy_forward_raw(kind, p, states.reshape(-1, d["D"]))
    np.testing.assert_allclose(h.read_field(abi.CRL_F_VALUE).ravel(), val, rtol=RTOL, atol=2e-6)
    if kind == 0:
        lp_taken = logp[np.arange(N * T), actions.ravel()]
        np.testing.assert_allclose(h.read_field(abi.CRL_F_LOGPROB).ravel(), lp_taken, rtol=RTOL, atol=2e-6)
        # the action is the inverse-CDF sample of the Philox uniform (ppo.jl:26)
        probs = np.exp(logp.astype(np.float64)).reshape(T, N, 2)
        for t in range(0, T, 17):
            for n in range(0, N, 13):
                u = olib.action_uniform(11, n, t)
                if abs(u - probs[t, n, 0]) > 1e-5:
                    assert actions[t, n] == (1 if probs[t, n, 0] < u else 0)
        # transitions: state[t+1] = step(state[t], action[t]) whenever neither step is flagged terminal
        for t in range(T - 1):
            ok = (term[t] == 0) & (term[t + 1] == 0)
            s1, _, r1, d1 = olib.env_step_raw(kind, states[t], np.zeros(N, np.int32), actions[t], 500)
            np.testing.assert_allclose(states[t + 1][ok], s1[ok], rtol=RTOL, atol=1e-6)
            assert np.all(rew[t][ok] == 1.0) and np.all(d1[ok] == 0)
            flagged = (term[t] == 0) & (term[t + 1] == 1)
            assert np.all(rew[t][flagged] == 0.0)  # reward 0 on the terminating step
        assert term[0].sum() == 0 and term.sum() > 0
    else:
        mean = pol.reshape(T, N)
        sd = np.exp(p[-1])
        z = (actions[..., 0] - mean) / sd
        lp = -(actions[..., 0] - mean) ** 2 / (2 * sd * sd) - p[-1] - 0.9189385332046727
        np.testing.assert_allclose(h.read_field(abi.CRL_F_LOGPROB), lp, rtol=1e-4, atol=1e-5)
        assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05
        zo = np.array([olib.action_normals(11, n, 5)[0] for n in range(N)])
        np.testing.assert_allclose(z[5], zo, rtol=1e-3, atol=2e-4)
    # episode bookkeeping is bit-exact against the terminal flags
    recs, agg = h.pop_episodes()
    nd = h.read_field(abi.CRL_F_NEXT_DONE)
    assert agg.count == term[1:].sum() + nd.sum() == len(recs)
    assert recs == sorted(recs, key=lambda r: (r[0], r[1]))
    h.close()


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("gae_mode", [0, 1])
def test_full_update_matches_oracle(crl, olib, abi, torch_cuda, kind, gae_mode):
    """rollout (injected noise) + GAE + 2 epochs x 4 minibatches with host permutations:
    losses, gradients, Adam state and post-step parameters vs the oracle."""
    N, T = 64, 16
    h, o = make_pair(crl, olib, abi, kind, N=N, T=T, mb=4, epochs=2, gae_mode=gae_mode)
    rng = np.random.default_rng(17)
    an = rng.random((T, N)) if kind == 0 else rng.standard_normal((T, N, 1))
    rn = rng.random((T, N, 4)).astype(F)
    perms = np.stack([rng.permutation(N * T) for _ in range(2)]).astype(np.int32)
    lr = float(F(2.5e-4)) * 0.9
    for x in (h, o):
        x.env_reset()
        x.rollout(an, rn)
        x.gae()
    np.testing.assert_allclose(h.read_field(abi.CRL_F_ADVANTAGE), o.read_field(abi.CRL_F_ADVANTAGE), rtol=1e-4, atol=1e-5)
    # feed the GPU's own buffers to the oracle so the update is compared on identical inputs
    for name in BUF_FIELDS + ["ADVANTAGE", "RETURN"]:
        f = getattr(abi, "CRL_F_" + name)
        o.write_field(f, h.read_field(f))
    sh = h.update_epochs(perms, lr)
    so = o.update_epochs(perms, lr)
    np.testing.assert_allclose(sh, so, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=RTOL, atol=1e-6)
    mh, vh, bph = h.get_adam_state()
    mo, vo, bpo = o.get_adam_state()
    np.testing.assert_allclose(bph, bpo, rtol=1e-14)
    np.testing.assert_allclose(mh, mo, rtol=1e-3, atol=1e-6 * np.abs(mo).max())
    gh, go = h.get_grads(), o.get_grads()
    np.testing.assert_allclose(gh, go, rtol=1e-2, atol=1e-4 * np.abs(go).max())
    h.close()


def test_single_minibatch_first_step_parameters(crl, olib, abi, torch_cuda):
    """the reference's default shape: N=4, T=32, minibatch 32 (ppo.jl:2-6), one optimiser step"""
    h, o = make_pair(crl, olib, abi, 0, N=4, T=32, mb=4, epochs=4, seed=2)
    for x in (h, o):
        x.env_reset()
        x.rollout()
        x.gae()
    compare_buffers(h, o, abi, 0)
    for name in BUF_FIELDS + ["ADVANTAGE", "RETURN"]:
        f = getattr(abi, "CRL_F_" + name)
        o.write_field(f, h.read_field(f))
    idx = np.random.default_rng(0).permutation(128)[:32].astype(np.int32)
    lr = float(F(2.5e-4))
    s1, s2 = h.update_minibatch(idx, lr), o.update_minibatch(idx, lr)
    # pg_loss is a mean of 32 terms -adv_n * ratio of magnitude ~1 that cancel to ~1e-7 on the first step; ratio =
    # exp(~1e-8) is quantised to 1 or 1 + 2^-23, so the sum carries ~1e-7 of rounding noise: atol 5e-7 = 5e-7 of a summand
    np.testing.assert_allclose([s1.loss, s1.pg_loss, s1.v_loss, s1.entropy_loss],
                               [s2.loss, s2.pg_loss, s2.v_loss, s2.entropy_loss], rtol=RTOL, atol=5e-7)
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(h.read_field(abi.CRL_F_VNEW)[:32], o.read_field(abi.CRL_F_VNEW)[:32], rtol=RTOL, atol=2e-6)
    h.close()


@pytest.mark.parametrize("B_shape", [(4, 32), (64, 16), (100, 7), (4096, 128)])
def test_device_permutation_bit_exact(crl, olib, abi, torch_cuda, B_shape):
    N, T = B_shape
    h, o = make_pair(crl, olib, abi, 0, N=N, T=T, mb=1, epochs=1, seed=99)
    for (u, e) in [(0, 0), (0, 3), (5, 1), (2 ** 33 + 1, 2)]:
        ph, po = h.device_permutation(u, e), o.device_permutation(u, e)
        np.testing.assert_array_equal(ph, po)
        assert np.array_equal(np.sort(ph), np.arange(N * T))
    h.close()


@pytest.mark.parametrize("kind", [0, 1])
def test_train_update_graph_matches_oracle(crl, olib, abi, torch_cuda, kind):
    """crl_train_update (CUDA graph, device RNG and device permutation) for 3 updates vs the oracle
    running the same Philox streams; short horizon keeps fp32 divergence below tolerance."""
    N, T = 64, 16
    h, o = make_pair(crl, olib, abi, kind, N=N, T=T, mb=4, epochs=2, seed=23)
    h.env_reset(); o.env_reset()
    for u in range(3):
        lr = float(F(2.5e-4)) * (1 - u / 10)
        h.train_update(lr)
        sh, agg = h.fetch_update()
        so = o.train_update(lr)
        np.testing.assert_allclose(sh, so, rtol=5e-4, atol=1e-5, err_msg="update %d" % u)
        np.testing.assert_array_equal(h.read_field(abi.CRL_F_TERMINAL), o.read_field(abi.CRL_F_TERMINAL))
        if kind == 0:
            np.testing.assert_array_equal(h.read_field(abi.CRL_F_ACTION), o.read_field(abi.CRL_F_ACTION))
        np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=1e-4, atol=2e-6)
        _, ao = o.pop_episodes()
        assert agg.count == ao.count
    assert h.kernel_launches() >= 3 * (4 + 8 * 1)  # per update: init, rollout, gae, adv_stats + one fused kernel per minibatch (3 unfused)
    assert h.spec_replays() == 0
    h.close()


def test_default_config_shape_matches_oracle(crl, olib, abi, torch_cuda):
    """BASELINE.json configs[0]: PPOConfig() defaults (ppo.jl:2-6): 4 envs x 32 steps, 4 epochs x 4 minibatches of 32
    samples. One 128-sample tile holds a whole minibatch, so 146 of the update kernel's 148 CTAs have no tile and only
    take part in the fused tail (reduce, clip, Adam on their slab)."""
    N, T = 4, 32
    h, o = make_pair(crl, olib, abi, 0, N=N, T=T, mb=4, epochs=4, seed=1)
    h.env_reset(); o.env_reset()
    for u in range(4):
        lr = float(F(2.5e-4)) * (1 - u / 10)
        h.train_update(lr)
        sh, agg = h.fetch_update()
        so = o.train_update(lr)
        np.testing.assert_allclose(sh, so, rtol=5e-4, atol=1e-5, err_msg="update %d" % u)
        np.testing.assert_array_equal(h.read_field(abi.CRL_F_TERMINAL), o.read_field(abi.CRL_F_TERMINAL))
        np.testing.assert_array_equal(h.read_field(abi.CRL_F_ACTION), o.read_field(abi.CRL_F_ACTION))
        np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=1e-4, atol=2e-6)
    assert h.kernel_launches() >= 4 * (4 + 16)   # one fused launch per minibatch
    assert h.spec_replays() == 0
    h.close()


def test_failed_speculation_is_replayed_exactly(crl, olib, abi, torch_cuda):
    """gamma = lambda = 0 with a critic bias of 1.5: returns are in {0, 1, v}, so the scalar s = mean(v - R^2) of the
    value loss (Q5) exceeds min (clip - R)^2 = 0 and the speculative update fails its on-device verification. The
    library must notice, restore its snapshot and replay the update with the exact kernels: results still match the
    oracle, and the replay counter shows the slow path really ran."""
    N, T = 64, 16
    h, o = make_pair(crl, olib, abi, 0, N=N, T=T, mb=4, epochs=2, seed=29, gamma=0.0, gae_lambda=0.0)
    p = h.get_params()
    p[olib.param_layout(0)[0][11]] = 1.5
    h.set_params(p); o.set_params(p)
    h.env_reset(); o.env_reset()
    lr = float(F(2.5e-4))
    outs = []
    for u in range(3):
        h.train_update(lr)
        if u >= 1:
            outs.append(h.fetch_update(lag=1)[0])   # pipelined fetch: validates update u-1 while u is in flight
    outs.append(h.fetch_update(lag=0)[0])
    for u in range(3):
        so = o.train_update(lr)
        np.testing.assert_allclose(outs[u], so, rtol=5e-4, atol=1e-5, err_msg="update %d" % u)
    assert h.spec_replays() >= 1
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=1e-4, atol=2e-6)
    np.testing.assert_array_equal(h.read_field(abi.CRL_F_TERMINAL), o.read_field(abi.CRL_F_TERMINAL))
    h.close()


def test_call_order_and_argument_errors(crl, olib, abi, torch_cuda):
    from cleanrl_jl_b200.handle import PPOHandle
    from cleanrl_jl_b200._lib import CleanRLCudaError
    h = PPOHandle(abi.make_config(num_envs=4, num_steps=8, num_minibatches=2))
    with pytest.raises(CleanRLCudaError) as e:
        h.gae()
    assert e.value.code == abi.CRL_ERR_STATE
    with pytest.raises(CleanRLCudaError) as e:
        h.update_epochs(None, 1e-3)
    assert e.value.code == abi.CRL_ERR_STATE
    with pytest.raises(CleanRLCudaError) as e:
        h.set_params(np.zeros(3, F))
    assert e.value.code == abi.CRL_ERR_INVALID
    h.close()
    for bad in (dict(num_envs=3, num_steps=5, num_minibatches=2),  # Q10: not divisible
                dict(num_envs=0), dict(env_kind=7), dict(num_envs=1, num_steps=1, num_minibatches=1)):
        with pytest.raises(CleanRLCudaError) as e:
            PPOHandle(abi.make_config(**bad))
        assert e.value.code == abi.CRL_ERR_INVALID
    cfg = abi.make_config()
    cfg.struct_size = 8
    with pytest.raises(CleanRLCudaError):
        PPOHandle(cfg)
    cfg = abi.make_config(device=63)
    with pytest.raises(CleanRLCudaError) as e:
        PPOHandle(cfg)
    assert e.value.code == abi.CRL_ERR_CUDA


@pytest.mark.parametrize("kind", [0, 1])
def test_unclipped_value_loss_update_matches_oracle(crl, olib, abi, torch_cuda, kind):
    """PPOConfig.clip_value_loss = false (ppo.jl:239-241): stage-by-stage update and the graph path against the oracle"""
    h, o = make_pair(crl, olib, abi, kind, N=64, T=16, mb=4, epochs=2, seed=6, flags=abi.CRL_FLAG_NO_VCLIP)
    for x in (h, o):
        x.env_reset()
    for u in range(2):
        sh = h.train_update(2.5e-4) or h.fetch_update()[0]
        so = o.train_update(2.5e-4)
        np.testing.assert_allclose(sh, so, rtol=1e-4, atol=2e-6, err_msg="update %d" % u)
        np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=1e-5, atol=1e-6)
    assert h.spec_replays() == 0
    # and the exact chain (crl_update_epochs) from identical buffers
    for x in (h, o):
        x.rollout(); x.gae()
    for name in BUF_FIELDS + ["ADVANTAGE", "RETURN"]:
        f = getattr(abi, "CRL_F_" + name)
        o.write_field(f, h.read_field(f))
    sh, so = h.update_epochs(None, 1e-4), o.update_epochs(None, 1e-4)
    np.testing.assert_allclose(sh, so, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(h.get_params(), o.get_params(), rtol=1e-5, atol=1e-6)
    h.close()
