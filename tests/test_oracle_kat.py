"""Pins the CPU oracle against known-answer vectors. CPU only."""
import numpy as np
import pytest

F = np.float32


def test_philox_random123_kat(olib):
    # Random123 kat_vectors, philox4x32-10
    assert list(olib.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(olib.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(olib.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _gae_1env(olib, values, rewards, terminals, mode):
    v = np.array(values, F)
    T = len(rewards)
    adv, ret = olib.gae_raw(v[:T, None], np.array(rewards, F)[:, None], np.array(terminals[:T], np.uint8)[:, None],
                            v[T:], np.array(terminals[T:], np.uint8), F(0.99), F(0.95), mode)
    return adv[:, 0], ret[:, 0]


def test_gae_kat1_terminal_before_last_step(olib, abi):
    # KAT-1, hand-derived from ppo.jl:62-72 (Float64 accumulation, γλ multiplied in Float32)
    vals, rews, term = [0.5, 0.6, 0.7, 0.2, 0.9], [1, 1, 0, 1], [0, 0, 0, 1, 0]
    adv, ret = _gae_1env(olib, vals, rews, term, abi.CRL_GAE_REF_COMPAT)
    np.testing.assert_array_equal(adv[:3], np.array([1.5027883052825928, 0.4346499741077423, -0.699999988079071], F))
    assert adv[3] == 0.0  # never written by the reference (Q1); defined as 0 here
    np.testing.assert_array_equal(ret, adv + np.array(vals[:4], F))
    adv_f, _ = _gae_1env(olib, vals, rews, term, abi.CRL_GAE_FIXED)
    np.testing.assert_array_equal(adv_f, np.array([1.5027883052825928, 0.4346499741077423, -0.699999988079071,
                                                   1.690999984741211], F))


def test_gae_kat2_no_terminals(olib, abi):
    vals, rews, term = [0.5, 0.6, 0.7, 0.2, 0.9], [1, 1, 0, 1], [0, 0, 0, 0, 0]
    adv, _ = _gae_1env(olib, vals, rews, term, abi.CRL_GAE_REF_COMPAT)
    np.testing.assert_array_equal(adv[:3], np.array([1.6779272556304932, 0.620868980884552, -0.5019999742507935], F))
    adv_f, _ = _gae_1env(olib, vals, rews, term, abi.CRL_GAE_FIXED)
    np.testing.assert_array_equal(adv_f, np.array([3.0846874713897705, 2.116626501083374, 1.0883855819702148,
                                                   1.690999984741211], F))
    assert F(0.99) * F(0.95) == F(0.940500020980835)


def test_gae_matches_float64_numpy_restatement(olib, abi):
    """independent restatement of ppo.jl:62-72 in NumPy float64, random inputs"""
    rng = np.random.default_rng(3)
    T, N = 37, 19
    v = rng.standard_normal((T + 1, N)).astype(F)
    r = rng.standard_normal((T, N)).astype(F)
    d = (rng.random((T + 1, N)) < 0.1).astype(np.uint8)
    g, lam = F(0.99), F(0.95)
    adv, ret = olib.gae_raw(v[:T], r, d[:T], v[T], d[T], g, lam, abi.CRL_GAE_FIXED)
    exp = np.zeros((T, N), F)
    gl = np.float64(g * lam)
    for n in range(N):
        gae = 0.0
        for t in range(T - 1, -1, -1):
            nonterm = 1.0 - float(d[t + 1, n])
            delta = float(r[t, n]) + float(g) * nonterm * float(v[t + 1, n]) - float(v[t, n])
            gae = delta + gl * nonterm * gae
            exp[t, n] = gae
    np.testing.assert_array_equal(adv, exp)
    np.testing.assert_array_equal(ret, exp + v[:T])
    # REF_COMPAT = the same scan restarted at T-1 with zero carry and no bootstrap
    adv_c, _ = olib.gae_raw(v[:T], r, d[:T], v[T], d[T], g, lam, abi.CRL_GAE_REF_COMPAT)
    adv_s, _ = olib.gae_raw(v[:T - 1], r[:T - 1], d[:T - 1], v[T - 1], d[T - 1], g, lam, abi.CRL_GAE_FIXED)
    np.testing.assert_array_equal(adv_c[:T - 1], adv_s)
    assert np.all(adv_c[T - 1] == 0)


def test_cartpole_kat3(olib, abi):
    s, t, r, d = olib.env_step_raw(abi.CRL_ENV_CARTPOLE, np.array([[0, 0, 0.05, 0]], F), [0], [1], 500)
    np.testing.assert_array_equal(s[0], np.array([0.0, 0.19437053799629211, 0.05000000074505806, -0.27649757266044617], F))
    assert t[0] == 1 and r[0] == 1.0 and d[0] == 0
    assert F(12 * 2 * np.pi / 360) == F(0.20943951606750488)


def test_cartpole_matches_gym_equations_and_termination(olib, abi):
    """cross-check against the classic Barto-Sutton/Gym Euler equations in float64"""
    rng = np.random.default_rng(0)
    n = 512
    s0 = (rng.random((n, 4)) * 0.4 - 0.2).astype(F)
    a = rng.integers(0, 2, n)
    s1, t1, r, d = olib.env_step_raw(abi.CRL_ENV_CARTPOLE, s0, np.zeros(n, np.int32), a, 500)
    x, xd, th, thd = [s0[:, i].astype(np.float64) for i in range(4)]
    force = np.where(a == 1, 10.0, -10.0)
    tmp = (force + 0.05 * thd ** 2 * np.sin(th)) / 1.1
    thacc = (9.8 * np.sin(th) - np.cos(th) * tmp) / (0.5 * (4 / 3 - 0.1 * np.cos(th) ** 2 / 1.1))
    xacc = tmp - 0.05 * thacc * np.cos(th) / 1.1
    exp = np.stack([x + 0.02 * xd, xd + 0.02 * xacc, th + 0.02 * thd, thd + 0.02 * thacc], 1)
    np.testing.assert_allclose(s1, exp, rtol=2e-6, atol=2e-7)
    done = (np.abs(s1[:, 0]) > F(2.4)) | (np.abs(s1[:, 2]) > F(0.20943951606750488))
    np.testing.assert_array_equal(d, done.astype(np.uint8))
    np.testing.assert_array_equal(r, np.where(done, 0.0, 1.0).astype(F))
    # t > max_steps terminates (max_steps=500 at ppo.jl:82): the 501st step
    s, t, r, d = olib.env_step_raw(abi.CRL_ENV_CARTPOLE, np.zeros((2, 4), F), [499, 500], [1, 1], 500)
    assert list(t) == [500, 501] and list(d) == [0, 1] and list(r) == [1.0, 0.0]


def test_pendulum_step_and_reset(olib, abi):
    rng = np.random.default_rng(1)
    n = 256
    s0 = np.stack([rng.uniform(-7, 7, n), rng.uniform(-8, 8, n)], 1).astype(F)
    a = rng.uniform(-3, 3, n).astype(F)
    s1, t1, r, d = olib.env_step_raw(abi.CRL_ENV_PENDULUM, s0, np.zeros(n, np.int32), a, 200)
    th, thd = s0[:, 0].astype(np.float64), s0[:, 1].astype(np.float64)
    ac = np.clip(a.astype(np.float64), -2, 2)
    an = np.mod(th + np.pi, 2 * np.pi) - np.pi
    cost = an ** 2 + 0.1 * thd ** 2 + 0.001 * ac ** 2
    nthd = thd + (-3 * 10.0 / 2 * np.sin(th + np.pi) + 3.0 * ac) * 0.05
    nth = th + nthd * 0.05
    np.testing.assert_allclose(s1[:, 0], nth, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(s1[:, 1], np.clip(nthd, -8, 8), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(r, -cost, rtol=1e-5, atol=1e-5)
    assert not d.any()
    _, t, _, d = olib.env_step_raw(abi.CRL_ENV_PENDULUM, s0[:2], [198, 199], a[:2], 200)
    assert list(d) == [0, 1]
    st, t = olib.env_reset_raw(abi.CRL_ENV_PENDULUM, np.array([[0.25, 0.75, 0, 0]], F))
    np.testing.assert_allclose(st[0], [2 * np.pi * (0.25 - 1), 2 * (0.75 - 1)], rtol=1e-6)
    st, t = olib.env_reset_raw(abi.CRL_ENV_CARTPOLE, np.array([[0.0, 0.5, 1.0, 0.25]], F))
    np.testing.assert_allclose(st[0], [-0.05, 0.0, 0.05, -0.025], atol=1e-8)


def test_tanh_fast_is_a_few_ulp_from_tanh(olib):
    x = np.linspace(-9, 9, 20001).astype(F)
    y = olib.tanh_fast(x)
    ref = np.tanh(x.astype(np.float64))
    assert np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1e-30)) < 4e-7
    assert olib.tanh_fast([100.0])[0] == 1.0 and olib.tanh_fast([-100.0])[0] == -1.0 and olib.tanh_fast([0.0])[0] == 0.0


def test_rng_contract_ranges(olib):
    us = [olib.action_uniform(7, e, s) for e in range(20) for s in range(20)]
    assert min(us) >= 0.0 and max(us) < 1.0 and 0.35 < np.mean(us) < 0.65
    u = olib.reset_uniforms(7, 3, 5)
    assert u.dtype == np.float32 and np.all(u >= 0) and np.all(u < 1)
    assert not np.array_equal(u, olib.reset_uniforms(7, 3, 6))
    z = np.array([olib.action_normals(11, e, 0) for e in range(4000)])
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1.0) < 0.05


@pytest.mark.parametrize("B", [2, 7, 128, 1000, 4096, 524288])
def test_device_permutation_is_a_bijection(olib, abi, B):
    T = 1
    cfg = abi.make_config(num_envs=B, num_steps=T, num_minibatches=1, seed=5)
    c = olib.create(cfg)
    p0 = c.device_permutation(0, 0)
    assert np.array_equal(np.sort(p0), np.arange(B))
    p1 = c.device_permutation(0, 1)
    p2 = c.device_permutation(1, 0)
    if B > 16:
        assert not np.array_equal(p0, p1) and not np.array_equal(p0, p2)
        assert np.mean(p0 == np.arange(B)) < 0.05


def test_clip_adam_matches_independent_float64_restatement(olib, abi):
    """Flux.Optimiser(ClipNorm(0.5), Adam(η)) per array [Flux 0.13.4] restated in NumPy."""
    rng = np.random.default_rng(5)
    kind = abi.CRL_ENV_CARTPOLE
    d = olib.dims(kind)
    off, size = olib.param_layout(kind)
    p = rng.standard_normal(d["P"]).astype(F)
    m = np.zeros(d["P"], F)
    v = np.zeros(d["P"], F)
    bp = np.tile(np.array([0.9, 0.999]), (d["n_arrays"], 1))
    pe, me, ve, bpe = p.copy(), m.copy(), v.copy(), bp.copy()
    lr = float(F(2.5e-4))
    for it in range(3):
        g = (rng.standard_normal(d["P"]) * (0.002 if it == 1 else 0.05)).astype(F)  # it==1: below the clip threshold
        p, m, v, bp = olib.clip_adam_raw(kind, p, g, m, v, bp, lr, 0.5)
        for i in range(d["n_arrays"]):
            sl = slice(off[i], off[i] + size[i])
            gi = g[sl].copy()
            nrm = F(np.sqrt(np.sum(gi.astype(np.float64) ** 2)))
            if nrm > 0.5:
                gi = (gi.astype(np.float64) * (0.5 / float(nrm))).astype(F)
            me[sl] = (0.9 * me[sl].astype(np.float64) + (1 - 0.9) * gi.astype(np.float64)).astype(F)
            ve[sl] = (0.999 * ve[sl].astype(np.float64) + (1 - 0.999) * gi.astype(np.float64) * gi.astype(np.float64)).astype(F)
            step = (me[sl].astype(np.float64) / (1 - bpe[i, 0]) / (np.sqrt(ve[sl].astype(np.float64) / (1 - bpe[i, 1])) + 1e-8) * lr).astype(F)
            pe[sl] = pe[sl] - step
            bpe[i] *= [0.9, 0.999]
        np.testing.assert_array_equal(p, pe)
        np.testing.assert_array_equal(m, me)
        np.testing.assert_array_equal(v, ve)
        np.testing.assert_allclose(bp, bpe, rtol=1e-15)
    # per-ARRAY clipping (Q7): the 64x64 arrays were clipped, the tiny bias arrays were not
    assert d["n_arrays"] == 12 and d["P"] == 9155


def test_layout_matches_flux_params_order(olib, abi):
    off, size = olib.param_layout(abi.CRL_ENV_CARTPOLE)
    assert list(size) == [256, 64, 4096, 64, 128, 2, 256, 64, 4096, 64, 64, 1]
    assert list(off) == list(np.cumsum([0] + list(size[:-1])))
    d = olib.dims(abi.CRL_ENV_PENDULUM)
    off, size = olib.param_layout(abi.CRL_ENV_PENDULUM)
    assert d == {"D": 3, "A": 1, "S": 2, "P": 4481 * 2 + 1, "n_arrays": 13}
    assert list(size) == [192, 64, 4096, 64, 64, 1] * 2 + [1]
