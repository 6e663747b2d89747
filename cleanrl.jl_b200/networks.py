"""Networks — mirror of src/utils/networks.jl (host-side initialisation only; the forward and
backward passes run in the CUDA library).

make_actor_critic builds two separate D-64-64-out tanh MLPs (networks.jl:36-49): hidden layers
orthogonal with gain sqrt(2), actor head gain 0.01, critic head gain 1.0, zero biases. The
flat vector handed to crl_set_params follows Flux.params(actor, critic) order (ppo.jl:196) with
every W stored (out,in) column-major.
"""
import numpy as np


def orthogonal(rng, rows, cols, gain=1.0):
    """Flux.orthogonal(rng, rows, cols; gain) [Flux 0.13.4]: QR of a Gaussian matrix with the sign
    fix, transposed when rows < cols. The random stream is NumPy's, not Julia's."""
    if rows < cols:
        return np.ascontiguousarray(orthogonal(rng, cols, rows, gain).T)
    mat = rng.standard_normal((rows, cols)).astype(np.float32)
    q, r = np.linalg.qr(mat)
    q = q * np.sign(np.diag(r))[None, :]
    return (q * np.float32(gain)).astype(np.float32)


def mlp(layer_sizes, rng, gain=np.sqrt(2.0)):
    """networks.jl:6-13: Dense(in, out, tanh_fast; init=orthogonal(gain)) per consecutive pair."""
    return [(orthogonal(rng, layer_sizes[i + 1], layer_sizes[i], gain), np.zeros(layer_sizes[i + 1], np.float32))
            for i in range(len(layer_sizes) - 1)]


def make_actor_critic(n_actions, obs_dim, hidden_sizes=(64, 64), rng=None, seed=0):
    """networks.jl:36-49. Returns (actor, critic): lists of (W[out,in], b[out])."""
    if tuple(hidden_sizes) != (64, 64):
        raise ValueError("the CUDA kernels are specialised for the reference's 64-64 hidden layers (networks.jl:36)")
    rng = rng or np.random.default_rng(seed)
    sizes = [obs_dim] + list(hidden_sizes)
    actor = mlp(sizes, rng) + [(orthogonal(rng, n_actions, hidden_sizes[-1], 0.01), np.zeros(n_actions, np.float32))]
    critic = mlp(sizes, rng) + [(orthogonal(rng, 1, hidden_sizes[-1], 1.0), np.zeros(1, np.float32))]
    return actor, critic


def flatten_params(actor, critic, logstd=None):
    """Flux.params(actor, critic) order; W column-major (out,in)."""
    parts = []
    for net in (actor, critic):
        for W, b in net:
            parts.append(np.asarray(W, np.float32).flatten(order="F"))
            parts.append(np.asarray(b, np.float32).ravel())
    if logstd is not None:
        parts.append(np.asarray(logstd, np.float32).ravel())
    return np.concatenate(parts).astype(np.float32)


def unflatten_params(flat, obs_dim, n_actions, continuous=False):
    """inverse of flatten_params -> (actor, critic, logstd|None)"""
    flat = np.asarray(flat, np.float32)
    o = 0
    nets = []
    for out in (n_actions, 1):
        layers = []
        for (i, j) in ((obs_dim, 64), (64, 64), (64, out)):
            W = flat[o:o + i * j].reshape((j, i), order="F")
            o += i * j
            b = flat[o:o + j]
            o += j
            layers.append((W.copy(), b.copy()))
        nets.append(layers)
    logstd = flat[o:o + n_actions].copy() if continuous else None
    return nets[0], nets[1], logstd


def init_params(env_kind_continuous, obs_dim, n_actions, seed=0):
    actor, critic = make_actor_critic(n_actions, obs_dim, seed=seed)
    logstd = np.zeros(n_actions, np.float32) if env_kind_continuous else None
    return flatten_params(actor, critic, logstd)
