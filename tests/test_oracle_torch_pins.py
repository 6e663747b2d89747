"""Third-party pins for the oracle (the reference has no tests and its arithmetic lives in un-vendored
Julia packages). PyTorch's own modules implement the same published algorithms (Linear + tanh, softmax/log_softmax,
Adam with bias correction and eps outside the square root) and ARE present in this image, so they serve as an
independent implementation: the oracle must agree with them, not only with restatements written in this repo.
CPU only; nothing here touches the CUDA path."""
import numpy as np
import pytest
import torch

from conftest import rand_params
from test_oracle_grad import torch_loss

F = np.float32


def torch_nets(params, layout, D, A):
    """the two Flux chains of networks.jl:36-49 as torch.nn modules in float64 (true tanh instead of tanh_fast)"""
    off, size = layout

    def chain(base, out):
        dims = [(D, 64), (64, 64), (64, out)]
        layers = []
        for i, (fin, fout) in enumerate(dims):
            lin = torch.nn.Linear(fin, fout).double()
            W = params[off[base + 2 * i]:off[base + 2 * i] + size[base + 2 * i]].astype(np.float64)
            b = params[off[base + 2 * i + 1]:off[base + 2 * i + 1] + size[base + 2 * i + 1]].astype(np.float64)
            with torch.no_grad():
                lin.weight.copy_(torch.tensor(W.reshape(fin, fout).T))  # Flux (out,in) column-major = torch (out,in) row-major^T
                lin.bias.copy_(torch.tensor(b))
            layers.append(lin)
            if i < 2:
                layers.append(torch.nn.Tanh())
        return torch.nn.Sequential(*layers)

    return chain(0, A), chain(6, 1)


@pytest.mark.parametrize("kind", [0, 1])
def test_policy_forward_matches_torch_modules(olib, kind):
    """actor/critic forward + softmax/logsoftmax (ppo.jl:22-24, networks.jl:6-13) against torch.nn.Linear/Tanh and
    torch.log_softmax. tanh_fast is a few ulp from tanh, so logits agree to ~1e-6 absolute."""
    d = olib.dims(kind)
    rng = np.random.default_rng(11)
    params = rand_params(olib, kind, seed=3)
    obs = (rng.standard_normal((257, d["D"])) * 0.7).astype(F)
    pol, logp, val = olib.policy_forward_raw(kind, params, obs)
    actor, critic = torch_nets(params, olib.param_layout(kind), d["D"], d["A"])
    with torch.no_grad():
        x = torch.tensor(obs.astype(np.float64))
        z = actor(x)
        v = critic(x)[:, 0]
    np.testing.assert_allclose(pol, z.numpy(), rtol=1e-5, atol=5e-6)
    np.testing.assert_allclose(val, v.numpy(), rtol=1e-5, atol=5e-6)
    if kind == 0:
        np.testing.assert_allclose(logp, torch.log_softmax(z, 1).numpy(), rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(np.exp(logp.astype(np.float64)).sum(1), 1.0, atol=1e-6)


def test_clip_adam_matches_torch_optim_adam(olib, abi):
    """Flux.Optimiser(ClipNorm(0.5), Adam(eta)) (ppo.jl:93,250) against torch.optim.Adam in float64 with the
    per-array clip applied by hand: both divide m and v by (1 - beta^t) and add eps OUTSIDE the square root, so they
    are the same update. Oracle state is Float32 (Flux keeps Float32 moments): agreement to Float32 rounding."""
    kind = abi.CRL_ENV_CARTPOLE
    d = olib.dims(kind)
    off, size = olib.param_layout(kind)
    rng = np.random.default_rng(17)
    p0 = rng.standard_normal(d["P"]).astype(F)
    tp = [torch.nn.Parameter(torch.tensor(p0[off[i]:off[i] + size[i]].astype(np.float64))) for i in range(d["n_arrays"])]
    lr = float(F(2.5e-4))
    opt = torch.optim.Adam(tp, lr=lr, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.copy(), np.zeros(d["P"], F), np.zeros(d["P"], F)
    bp = np.tile(np.array([0.9, 0.999]), (d["n_arrays"], 1))
    for it in range(6):
        g = (rng.standard_normal(d["P"]) * (0.002 if it % 2 else 0.05)).astype(F)
        p, m, v, bp = olib.clip_adam_raw(kind, p, g, m, v, bp, lr, 0.5)
        for i, q in enumerate(tp):
            gi = torch.tensor(g[off[i]:off[i] + size[i]].astype(np.float64))
            nrm = gi.norm()
            if nrm > 0.5:  # Flux ClipNorm: rescale only above the threshold, per array (Q7)
                gi = gi * (0.5 / nrm)
            q.grad = gi
        opt.step()
        ref = np.concatenate([q.detach().numpy() for q in tp])
        # the step is ~lr per element; Float32 moments perturb it by ~1e-7 relative
        np.testing.assert_allclose(p, ref, rtol=0, atol=2e-7 * (it + 1) + 1.2e-7 * np.abs(ref).max())
        assert np.abs(p - p0).max() > 0.5 * lr  # the update moved the parameters by about lr per step
    st = opt.state[tp[2]]
    np.testing.assert_allclose(m[off[2]:off[2] + size[2]], st["exp_avg"].numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(v[off[2]:off[2] + size[2]], st["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("kind", [0, 1])
def test_minibatch_step_matches_autograd_plus_torch_adam(olib, abi, kind):
    """One whole minibatch step of the oracle's handle (rollout -> GAE -> loss -> backward -> ClipNorm -> Adam,
    ppo.jl:123-250) against float64 autograd of the restated loss followed by torch.optim.Adam. Pins the composition
    (flat index n + N t, per-array clip, parameter order), not only the pieces."""
    cfg = abi.make_config(env_kind=kind, num_envs=16, num_steps=12, num_minibatches=2, update_epochs=1, seed=9)
    o = olib.create(cfg)
    d = olib.dims(kind)
    layout = olib.param_layout(kind)
    off, size = layout
    p0 = rand_params(olib, kind, seed=4, scale=0.5)
    o.set_params(p0)
    o.env_reset()
    o.rollout()
    o.gae()
    B = cfg.num_envs * cfg.num_steps
    flat = lambda f: o.read_field(f).reshape(B, -1).squeeze()
    states = o.read_field(abi.CRL_F_STATE).reshape(B, d["D"])
    actions = flat(abi.CRL_F_ACTION)
    logprobs, adv, ret, val = (flat(f) for f in (abi.CRL_F_LOGPROB, abi.CRL_F_ADVANTAGE, abi.CRL_F_RETURN, abi.CRL_F_VALUE))
    idx = np.random.default_rng(2).permutation(B)[:B // 2].astype(np.int32)
    lr = 1e-3
    st = o.update_minibatch(idx, lr)

    tparams = torch.tensor(p0.astype(np.float64), requires_grad=True)
    loss, pg, vl, ent = torch_loss(tparams, layout, d, idx.astype(np.int64), states, actions, logprobs, adv, ret, val,
                                   cfg.clip_coef, cfg.ent_coeff, cfg.v_coef, kind == 1)
    loss.backward()
    np.testing.assert_allclose([st.loss, st.pg_loss, st.v_loss, st.entropy_loss],
                               [loss.item(), pg.item(), vl.item(), ent.item()], rtol=2e-5, atol=2e-6)
    g = tparams.grad.numpy()
    tp = [torch.nn.Parameter(torch.tensor(p0[off[i]:off[i] + size[i]].astype(np.float64))) for i in range(d["n_arrays"])]
    opt = torch.optim.Adam(tp, lr=lr, betas=(0.9, 0.999), eps=1e-8)
    for i, q in enumerate(tp):
        gi = torch.tensor(g[off[i]:off[i] + size[i]])
        nrm = gi.norm()
        q.grad = gi * (0.5 / nrm) if nrm > 0.5 else gi
    opt.step()
    ref = np.concatenate([q.detach().numpy() for q in tp])
    got = o.get_params()
    # the first Adam step is lr * sign(g) wherever |g| >> eps: compare the STEP, not the parameter
    step_ref, step_got = ref - p0.astype(np.float64), got.astype(np.float64) - p0.astype(np.float64)
    big = np.abs(g) > 1e-5
    assert big.mean() > 0.5
    np.testing.assert_allclose(step_got[big], step_ref[big], rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=lr * 0.05)
    o.close()
