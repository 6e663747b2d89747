"""Size-independent properties of the oracle's GAE scan, minibatch permutation, CartPole step and clip+Adam
(hypothesis-driven, CPU only): what must hold for ANY input, next to the known-answer vectors of test_oracle_kat.py.
The GPU suite checks the same kernels against the oracle; these properties are what makes the oracle itself credible
at shapes no hand-derived vector covers (T = 1, N = 1, all-terminal, prime batch sizes)."""
import numpy as np
from hypothesis import given, settings, strategies as st

F = np.float32
SET = dict(max_examples=40, deadline=None)


def _inputs(rng, T, N, p_done):
    v = rng.standard_normal((T + 1, N)).astype(F)
    r = rng.standard_normal((T, N)).astype(F)
    d = (rng.random((T + 1, N)) < p_done).astype(np.uint8)
    return v, r, d


@settings(**SET)
@given(T=st.integers(1, 40), N=st.integers(1, 17), seed=st.integers(0, 2 ** 31), p=st.sampled_from([0.0, 0.1, 1.0]))
def test_gae_lambda_zero_is_the_td_residual(olib, abi, T, N, seed, p):
    """lambda = 0: adv[t] = r[t] + gamma * (1 - done[t+1]) * v[t+1] - v[t], evaluated in Float64 (ppo.jl:63-69)"""
    v, r, d = _inputs(np.random.default_rng(seed), T, N, p)
    g = F(0.99)
    adv, ret = olib.gae_raw(v[:T], r, d[:T], v[T], d[T], g, F(0.0), abi.CRL_GAE_FIXED)
    delta = r.astype(np.float64) + np.float64(g) * (1.0 - d[1:].astype(np.float64)) * v[1:].astype(np.float64) - v[:T].astype(np.float64)
    np.testing.assert_array_equal(adv, delta.astype(F))
    np.testing.assert_array_equal(ret, adv + v[:T])      # returns = advantages + values, ppo.jl:181


@settings(**SET)
@given(T=st.integers(1, 40), N=st.integers(1, 9), seed=st.integers(0, 2 ** 31))
def test_gae_all_terminal_cuts_every_bootstrap(olib, abi, T, N, seed):
    v, r, d = _inputs(np.random.default_rng(seed), T, N, 1.0)
    adv, _ = olib.gae_raw(v[:T], r, d[:T], v[T], d[T], F(0.99), F(0.95), abi.CRL_GAE_FIXED)
    np.testing.assert_array_equal(adv, (r.astype(np.float64) - v[:T].astype(np.float64)).astype(F))


@settings(**SET)
@given(T=st.integers(2, 30), N=st.integers(2, 12), seed=st.integers(0, 2 ** 31), mode=st.sampled_from([0, 1]))
def test_gae_envs_are_independent_and_time_is_causal_backwards(olib, abi, T, N, seed, mode):
    rng = np.random.default_rng(seed)
    v, r, d = _inputs(rng, T, N, 0.15)
    g, lam = F(0.99), F(0.95)
    adv, _ = olib.gae_raw(v[:T], r, d[:T], v[T], d[T], g, lam, mode)
    # permuting the env columns permutes the result
    perm = rng.permutation(N)
    adv_p, _ = olib.gae_raw(v[:T, perm], r[:, perm], d[:T, perm], v[T, perm], d[T, perm], g, lam, mode)
    np.testing.assert_array_equal(adv_p, adv[:, perm])
    # a change of the reward at time t0 cannot reach advantages at later times (reverse-time scan)
    t0 = int(rng.integers(0, T))
    r2 = r.copy()
    r2[t0] += F(1.0)
    adv2, _ = olib.gae_raw(v[:T], r2, d[:T], v[T], d[T], g, lam, mode)
    np.testing.assert_array_equal(adv2[t0 + 1:], adv[t0 + 1:])
    if mode == abi.CRL_GAE_FIXED or t0 < T - 1:
        assert np.all(adv2[t0] != adv[t0])


@settings(**SET)
@given(N=st.integers(1, 9), seed=st.integers(0, 2 ** 31))
def test_gae_ref_compat_never_writes_the_last_step(olib, abi, N, seed):
    """Q1 (ppo.jl:62,66): the loop starts at T-1; adv[T] is defined as 0 here. T = 1 leaves nothing to scan."""
    for T in (1, 2, 5):
        v, r, d = _inputs(np.random.default_rng(seed + T), T, N, 0.2)
        adv, ret = olib.gae_raw(v[:T], r, d[:T], v[T], d[T], F(0.99), F(0.95), abi.CRL_GAE_REF_COMPAT)
        assert np.all(adv[T - 1] == 0)
        np.testing.assert_array_equal(ret[T - 1], v[T - 1])


@settings(**SET)
@given(B=st.one_of(st.integers(1, 70), st.sampled_from([127, 128, 129, 521, 1024, 4099])), upd=st.integers(0, 1000),
       epoch=st.integers(0, 7))
def test_device_permutation_is_a_bijection_for_any_batch_size(olib, abi, B, upd, epoch):
    """shuffle(1:B) (ppo.jl:194) is replaced by a keyed permutation: it must be a bijection of [0, B) for every B,
    including B = 1, primes, and sizes that are not a power of two (cycle walking)."""
    # B = num_envs * num_steps; one minibatch, one epoch is enough to instantiate the handle
    n_envs = B if B <= 64 or B % 2 else 2
    n_steps = B // n_envs
    cfg = abi.make_config(num_envs=n_envs, num_steps=n_steps, num_minibatches=1, update_epochs=8, seed=17)
    c = olib.create(cfg)
    p = c.device_permutation(upd, epoch)
    assert p.shape == (B,)
    assert np.array_equal(np.sort(p), np.arange(B))
    c.close()


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 33))
def test_cartpole_step_termination_and_reset_ranges(olib, abi, seed, n):
    """done <=> |x| > 2.4 or |theta| > 12 deg or t > max_steps; reward = done ? 0 : 1; reset states in [-0.05, 0.05)"""
    rng = np.random.default_rng(seed)
    s = np.column_stack([rng.uniform(-2.6, 2.6, n), rng.uniform(-2, 2, n), rng.uniform(-0.25, 0.25, n), rng.uniform(-2, 2, n)]).astype(F)
    t = rng.integers(0, 12, n).astype(np.int32)
    a = rng.integers(0, 2, n).astype(np.int32)
    s2, t2, r, d = olib.env_step_raw(abi.CRL_ENV_CARTPOLE, s, t, a, 10)
    thr = F(12.0 * 2.0 * np.pi / 360.0)
    exp = (np.abs(s2[:, 0]) > F(2.4)) | (np.abs(s2[:, 2]) > thr) | (t2 > 10)
    np.testing.assert_array_equal(d.astype(bool), exp)
    np.testing.assert_array_equal(r, np.where(exp, F(0), F(1)))
    np.testing.assert_array_equal(t2, t + 1)
    # positions integrate the OLD velocities (explicit Euler, RLEnvs CartPoleEnv)
    np.testing.assert_array_equal(s2[:, 0], s[:, 0] + F(0.02) * s[:, 1])
    np.testing.assert_array_equal(s2[:, 2], s[:, 2] + F(0.02) * s[:, 3])
    u4 = rng.random((n, 4)).astype(F)
    s0, t0 = olib.env_reset_raw(abi.CRL_ENV_CARTPOLE, u4)
    assert np.all(t0 == 0) and np.all(s0 >= F(-0.05)) and np.all(s0 <= F(0.05))
    np.testing.assert_array_equal(s0, F(0.1) * u4 - F(0.05))


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 2 ** 31), scale=st.floats(2.0, 50.0))
def test_clip_adam_is_scale_invariant_above_the_clip_threshold(olib, abi, seed, scale):
    """ClipNorm(0.5) per array (Q7): once every array's norm exceeds 0.5, multiplying the gradient by any factor > 1
    changes nothing (up to Float32 rounding of the rescale)."""
    kind = abi.CRL_ENV_CARTPOLE
    d = olib.dims(kind)
    rng = np.random.default_rng(seed)
    p = rng.standard_normal(d["P"]).astype(F)
    g = rng.standard_normal(d["P"]).astype(F)          # every array has norm >> 0.5 (the smallest holds 1 element ~ N(0,1)?)
    off, size = olib.param_layout(kind)
    for i in range(d["n_arrays"]):                      # make sure even the 1- and 2-element arrays are above the threshold
        sl = slice(off[i], off[i] + size[i])
        n = np.linalg.norm(g[sl].astype(np.float64))
        if n < 1.0:
            g[sl] = (g[sl] / max(n, 1e-3)).astype(F)
    z = np.zeros(d["P"], F)
    bp = np.tile(np.array([0.9, 0.999]), (d["n_arrays"], 1))
    lr = 1e-3
    p1, m1, v1, _ = olib.clip_adam_raw(kind, p, g, z, z, bp, lr, 0.5)
    p2, m2, v2, _ = olib.clip_adam_raw(kind, p, (g * F(scale)).astype(F), z, z, bp, lr, 0.5)
    np.testing.assert_allclose(m2, m1, rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(p2, p1, rtol=0, atol=2e-7)
