"""Second oracle for the gradient: the loss closure ppo.jl:213-243, with quirks
Q4/Q5/Q6, restated in float64 PyTorch and differentiated with autograd, against the hand-derived
backward pass of oracle/ppo_oracle.c. CPU only."""
import numpy as np
import pytest
import torch

from conftest import rand_params

F = np.float32
NC = [1.0, 0.1346604, 0.0035974074, 2.2332108e-5, 1.587199e-8]
DC = [1.0, 0.4679937, 0.026262015, 0.0003453992, 8.7767893e-7]


class TanhFast(torch.autograd.Function):
    """NNlib.tanh_fast with its scalar rule dx = dy*(1 - y^2)"""

    @staticmethod
    def forward(ctx, x):
        x2 = x * x
        n = sum(c * x2 ** i for i, c in enumerate(NC))
        d = sum(c * x2 ** i for i, c in enumerate(DC))
        y = torch.where(x2 < 66.0, x * n / d, torch.sign(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g * (1 - y * y)


def torch_loss(params, layout, dims, idx, states, actions, logprobs, adv, ret, val, c, ent_c, v_c, continuous, clip_vloss=True):
    off, size = layout
    D, A = dims["D"], dims["A"]
    t = lambda a: torch.tensor(np.asarray(a, np.float64))
    x = t(states)[idx].T  # D x M

    def net(base, out):
        W1 = params[off[base]:off[base] + size[base]].reshape(D, 64).T  # column-major (out,in)
        b1 = params[off[base + 1]:off[base + 1] + 64]
        W2 = params[off[base + 2]:off[base + 2] + 4096].reshape(64, 64).T
        b2 = params[off[base + 3]:off[base + 3] + 64]
        W3 = params[off[base + 4]:off[base + 4] + 64 * out].reshape(64, out).T
        b3 = params[off[base + 5]:off[base + 5] + out]
        h1 = TanhFast.apply(W1 @ x + b1[:, None])
        h2 = TanhFast.apply(W2 @ h1 + b2[:, None])
        return W3 @ h2 + b3[:, None]

    z = net(0, A)            # A x M
    newvalue = net(6, 1)[0]  # M
    mb_adv, mb_lp, mb_val, mb_ret = t(adv)[idx], t(logprobs)[idx], t(val)[idx], t(ret)[idx]
    M = len(idx)
    if not continuous:
        probs = torch.softmax(z, 0)
        lps = torch.log_softmax(z, 0)
        act = torch.tensor(np.asarray(actions)[idx], dtype=torch.int64)
        newlp = lps[act, torch.arange(M)]
        entropy = -(probs * lps)  # ppo.jl:42: the A x M matrix (Q4)
    else:
        logstd = params[off[12]:off[12] + A]
        a = t(actions).reshape(-1, A)[idx].T
        var = torch.exp(logstd)[:, None] ** 2
        newlp = (-(a - z) ** 2 / (2 * var) - logstd[:, None] - 0.9189385332046727).sum(0)
        entropy = (0.5 + 0.9189385332046727 + logstd)[:, None].expand(A, M)
    adv_n = (mb_adv - mb_adv.mean()) / (mb_adv.std(unbiased=True) + 1e-8)  # ppo.jl:221
    ratio = torch.exp(newlp - mb_lp)
    pg_loss = torch.maximum(-adv_n * ratio, -adv_n * torch.clamp(ratio, 1 - c, 1 + c)).mean()  # ppo.jl:226-228
    s = (newvalue - mb_ret ** 2).mean()  # ppo.jl:232 (Q5: a scalar)
    v_clipped = mb_val + torch.clamp(newvalue - mb_val, -c, c)
    v_loss = 0.5 * torch.maximum(s, (v_clipped - mb_ret) ** 2).mean()  # ppo.jl:234-237
    if not clip_vloss:
        v_loss = 0.5 * ((newvalue - mb_ret) ** 2).mean()  # ppo.jl:239-241
    entropy_loss = entropy.mean()  # ppo.jl:242
    loss = pg_loss - ent_c * entropy_loss + v_c * v_loss  # ppo.jl:243
    return loss, pg_loss, v_loss, entropy_loss


def make_batch(olib, kind, B, seed, small_returns=False):
    rng = np.random.default_rng(seed)
    d = olib.dims(kind)
    states = (rng.standard_normal((B, d["D"])) * 0.5).astype(F)
    if kind == 0:
        actions = rng.integers(0, d["A"], B).astype(np.int32)
    else:
        actions = rng.standard_normal((B, d["A"])).astype(F)
    logprobs = (-0.7 + 0.2 * rng.standard_normal(B)).astype(F)
    adv = rng.standard_normal(B).astype(F) * 2 + 0.3
    ret = (rng.standard_normal(B) * (0.1 if small_returns else 3.0)).astype(F)
    val = (rng.standard_normal(B) * 0.5).astype(F)
    return states, actions, logprobs, adv, ret, val


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("small_returns", [False, True])
def test_oracle_backward_matches_float64_autograd(olib, kind, small_returns):
    d = olib.dims(kind)
    layout = olib.param_layout(kind)
    B, M = 300, 96
    p = rand_params(olib, kind, seed=3 + kind)
    if kind == 1:
        p[-1] = -0.3
    if small_returns:
        p[layout[0][11]] = 1.5  # critic head bias: pushes v_new up so that s = mean(v - R^2) wins the max (Q5)
    states, actions, logprobs, adv, ret, val = make_batch(olib, kind, B, 11, small_returns)
    idx = np.random.default_rng(2).permutation(B)[:M].astype(np.int32)
    c, ent_c, v_c = float(F(0.2)), float(F(0.01)), float(F(0.5))
    g, stats, vnew = olib.ppo_loss_raw(kind, p, idx, states, actions, logprobs, adv, ret, val, c, ent_c, v_c)
    pt = torch.tensor(p.astype(np.float64), requires_grad=True)
    loss, pg, vl, en = torch_loss(pt, layout, d, idx, states, actions, logprobs, adv, ret, val, c, ent_c, v_c, kind == 1)
    loss.backward()
    gt = pt.grad.numpy()
    np.testing.assert_allclose(stats, [loss.item(), pg.item(), vl.item(), en.item()], rtol=2e-5, atol=1e-6)
    scale = np.abs(gt).max()
    assert scale > 1e-4
    np.testing.assert_allclose(g, gt, rtol=2e-3, atol=2e-5 * scale)
    # every parameter array receives gradient
    off, size = layout
    for i in range(d["n_arrays"]):
        assert np.abs(g[off[i]:off[i] + size[i]]).max() > 0
    if small_returns:
        # the Q5 count path really was exercised: the scalar wins for some samples
        s = np.mean(vnew - ret[idx] ** 2)
        dv = np.clip(vnew - val[idx], -c, c)
        assert np.sum(s > (val[idx] + dv - ret[idx]) ** 2) > 0


def test_tanh_fast_pullback_choice_is_immaterial():
    """Zygote may differentiate the rational function itself (ForwardDiff duals inside the fused
    broadcast) instead of using 1 - y^2. The two derivatives agree to ~1e-6 absolute, far below
    the parity tolerance, so the choice cannot be observed at rtol 1e-5 on gradients."""
    x = torch.linspace(-8, 8, 40001, dtype=torch.float64, requires_grad=True)
    x2 = x * x
    y = x * sum(c * x2 ** i for i, c in enumerate(NC)) / sum(c * x2 ** i for i, c in enumerate(DC))
    (dy,) = torch.autograd.grad(y.sum(), x)
    assert (dy - (1 - y.detach() ** 2)).abs().max().item() < 3e-6


def test_loss_rejects_tiny_minibatch(olib):
    states, actions, logprobs, adv, ret, val = make_batch(olib, 0, 8, 0)
    with pytest.raises(AssertionError):
        olib.ppo_loss_raw(0, rand_params(olib, 0), np.array([0], np.int32), states, actions, logprobs, adv, ret, val,
                          0.2, 0.01, 0.5)


@pytest.mark.parametrize("kind", [0, 1])
def test_oracle_unclipped_value_loss_matches_float64_autograd(olib, kind):
    """PPOConfig.clip_value_loss = false (ppo.jl:16,239-241; CRL_FLAG_NO_VCLIP)"""
    d = olib.dims(kind)
    layout = olib.param_layout(kind)
    B, M = 300, 96
    p = rand_params(olib, kind, seed=5 + kind)
    if kind == 1:
        p[-1] = -0.3
    states, actions, logprobs, adv, ret, val = make_batch(olib, kind, B, 13)
    idx = np.random.default_rng(4).permutation(B)[:M].astype(np.int32)
    c, ent_c, v_c = float(F(0.2)), float(F(0.01)), float(F(0.5))
    olib.lib.orc_set_no_vclip(1)
    try:
        g, stats, vnew = olib.ppo_loss_raw(kind, p, idx, states, actions, logprobs, adv, ret, val, c, ent_c, v_c)
    finally:
        olib.lib.orc_set_no_vclip(0)
    pt = torch.tensor(p.astype(np.float64), requires_grad=True)
    loss, pg, vl, en = torch_loss(pt, layout, d, idx, states, actions, logprobs, adv, ret, val, c, ent_c, v_c, kind == 1,
                                  clip_vloss=False)
    loss.backward()
    gt = pt.grad.numpy()
    np.testing.assert_allclose(stats, [loss.item(), pg.item(), vl.item(), en.item()], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(g, gt, rtol=2e-3, atol=2e-5 * np.abs(gt).max())
