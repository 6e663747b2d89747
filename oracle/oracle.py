"""ctypes loader for the CPU oracle (oracle/ppo_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Nothing under cleanrl.jl_b200/ imports this.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_here))
from cleanrl_jl_b200 import _abi  # noqa: E402  (struct definitions only)

_FIELD_DTYPES = {
    _abi.CRL_F_STATE: np.float32, _abi.CRL_F_LOGPROB: np.float32, _abi.CRL_F_REWARD: np.float32,
    _abi.CRL_F_TERMINAL: np.uint8, _abi.CRL_F_VALUE: np.float32, _abi.CRL_F_ADVANTAGE: np.float32,
    _abi.CRL_F_RETURN: np.float32, _abi.CRL_F_NEXT_OBS: np.float32, _abi.CRL_F_NEXT_DONE: np.uint8,
    _abi.CRL_F_NEXT_VALUE: np.float32, _abi.CRL_F_ENV_STATE: np.float32, _abi.CRL_F_ENV_T: np.int32,
    _abi.CRL_F_EP_RETURN: np.float64, _abi.CRL_F_EP_LENGTH: np.int32, _abi.CRL_F_RESET_COUNT: np.uint32,
    _abi.CRL_F_VNEW: np.float32,
}


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    so = os.path.join(_here, "libppo_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_here, "ppo_oracle.c")):
        subprocess.check_call(["make", "-B", "-C", _here], stdout=subprocess.DEVNULL)
    return so


def _ptr(a, ctype=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class OracleLib:
    def __init__(self, fast=False):
        build()
        name = "libppo_oracle_fast.so" if fast else "libppo_oracle.so"
        self.lib = C.CDLL(os.path.join(_here, name))
        L = self.lib
        L.orc_rng_action_uniform.restype = C.c_double
        L.orc_rng_action_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.orc_tanh_fast.restype = C.c_float
        L.orc_tanh_fast.argtypes = [C.c_float]
        L.orc_policy_step.restype = C.c_uint64
        L.orc_perm_index.restype = C.c_uint32
        L.orc_gae_raw.argtypes = [C.c_void_p] * 7 + [C.c_int32, C.c_int64, C.c_float, C.c_float, C.c_int32]
        L.orc_env_step_raw.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_int32]
        L.orc_env_reset_raw.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_policy_forward_raw.argtypes = [C.c_int32] + [C.c_void_p] * 5 + [C.c_int64]
        L.orc_ppo_loss_raw.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 6 + \
            [C.c_float] * 3 + [C.c_void_p] * 3
        L.orc_ppo_loss_phase.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 6 + \
            [C.c_float] * 3 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_a2c_loss_raw.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 5
        L.orc_clip_adam_raw.argtypes = [C.c_int32] + [C.c_void_p] * 5 + [C.c_double, C.c_float]
        L.orc_create.argtypes = [C.POINTER(_abi.crl_config), C.POINTER(C.c_void_p)]
        for f in ("orc_destroy", "orc_env_reset", "orc_gae"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_get_params.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_get_grads.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_get_adam_state.argtypes = [C.c_void_p] * 4
        L.orc_set_adam_state.argtypes = [C.c_void_p] * 4
        L.orc_env_set_state.argtypes = [C.c_void_p] * 3
        L.orc_rollout.argtypes = [C.c_void_p] * 3
        L.orc_update_minibatch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
        L.orc_update_epochs.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_train_update.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.orc_device_permutation.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.orc_read_field.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]
        L.orc_write_field.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t]
        L.orc_pop_episodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_policy_step.argtypes = [C.c_void_p]
        L.orc_philox4x32_10.argtypes = [C.c_void_p] * 3
        L.orc_rng_reset_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p]
        L.orc_rng_action_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p]
        L.orc_dims.argtypes = [C.c_int32] + [C.c_void_p] * 5
        L.orc_param_layout.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]

    # ---- stateless helpers -------------------------------------------------------
    def set_threads(self, n):
        self.lib.orc_set_threads(int(n))

    def dims(self, env_kind):
        v = [C.c_int32() for _ in range(5)]
        assert self.lib.orc_dims(env_kind, *[C.byref(x) for x in v]) == 0
        return dict(zip("D A S P n_arrays".split(), [x.value for x in v]))

    def param_layout(self, env_kind):
        n = self.dims(env_kind)["n_arrays"]
        off = np.zeros(n, np.int32)
        size = np.zeros(n, np.int32)
        assert self.lib.orc_param_layout(env_kind, _ptr(off), _ptr(size)) == 0
        return off, size

    def philox(self, ctr, key):
        ctr = np.asarray(ctr, np.uint32)
        key = np.asarray(key, np.uint32)
        out = np.zeros(4, np.uint32)
        self.lib.orc_philox4x32_10(_ptr(ctr), _ptr(key), _ptr(out))
        return out

    def reset_uniforms(self, seed, env, k):
        u = np.zeros(4, np.float32)
        self.lib.orc_rng_reset_uniforms(seed, env, k, _ptr(u))
        return u

    def action_uniform(self, seed, env, step):
        return self.lib.orc_rng_action_uniform(seed, env, step)

    def action_normals(self, seed, env, step):
        z = np.zeros(2, np.float64)
        self.lib.orc_rng_action_normals(seed, env, step, _ptr(z))
        return z

    def tanh_fast(self, x):
        return np.array([self.lib.orc_tanh_fast(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32)

    def gae_raw(self, values, rewards, dones, next_value, next_done, gamma, lam, mode):
        T, N = values.shape
        values = np.ascontiguousarray(values, np.float32)
        rewards = np.ascontiguousarray(rewards, np.float32)
        dones = np.ascontiguousarray(dones, np.uint8)
        next_value = np.ascontiguousarray(next_value, np.float32)
        next_done = np.ascontiguousarray(next_done, np.uint8)
        adv = np.zeros((T, N), np.float32)
        ret = np.zeros((T, N), np.float32)
        rc = self.lib.orc_gae_raw(_ptr(values), _ptr(rewards), _ptr(dones), _ptr(next_value), _ptr(next_done),
                                  _ptr(adv), _ptr(ret), T, N, gamma, lam, mode)
        assert rc == 0, rc
        return adv, ret

    def env_step_raw(self, env_kind, state, t, action, max_steps):
        state = np.array(state, np.float32, copy=True)
        t = np.array(t, np.int32, copy=True)
        n = t.shape[0]
        action = np.ascontiguousarray(action, np.int32 if env_kind == _abi.CRL_ENV_CARTPOLE else np.float32)
        reward = np.zeros(n, np.float32)
        done = np.zeros(n, np.uint8)
        rc = self.lib.orc_env_step_raw(env_kind, _ptr(state), _ptr(t), _ptr(action), _ptr(reward), _ptr(done), n, max_steps)
        assert rc == 0, rc
        return state, t, reward, done

    def env_reset_raw(self, env_kind, u4):
        u4 = np.ascontiguousarray(u4, np.float32)
        n = u4.shape[0]
        S = self.dims(env_kind)["S"]
        state = np.zeros((n, S), np.float32)
        t = np.zeros(n, np.int32)
        self.lib.orc_env_reset_raw(env_kind, _ptr(state), _ptr(t), _ptr(u4), n)
        return state, t

    def policy_forward_raw(self, env_kind, params, obs):
        d = self.dims(env_kind)
        params = np.ascontiguousarray(params, np.float32)
        obs = np.ascontiguousarray(obs, np.float32)
        n = obs.shape[0]
        pol = np.zeros((n, d["A"]), np.float32)
        logp = np.zeros((n, d["A"]), np.float32)
        val = np.zeros(n, np.float32)
        rc = self.lib.orc_policy_forward_raw(env_kind, _ptr(params), _ptr(obs), _ptr(pol), _ptr(logp), _ptr(val), n)
        assert rc == 0, rc
        return pol, logp, val

    def ppo_loss_raw(self, env_kind, params, idx, states, actions, logprobs, advantages, returns, values,
                     clip_coef, ent_coeff, v_coef):
        d = self.dims(env_kind)
        params = np.ascontiguousarray(params, np.float32)
        idx = np.ascontiguousarray(idx, np.int32)
        states = np.ascontiguousarray(states, np.float32)
        actions = np.ascontiguousarray(actions, np.int32 if env_kind == _abi.CRL_ENV_CARTPOLE else np.float32)
        arrs = [np.ascontiguousarray(a, np.float32) for a in (logprobs, advantages, returns, values)]
        grads = np.zeros(d["P"], np.float32)
        stats = np.zeros(4, np.float64)
        vnew = np.zeros(idx.shape[0], np.float32)
        rc = self.lib.orc_ppo_loss_raw(env_kind, _ptr(params), _ptr(idx), idx.shape[0], _ptr(states), _ptr(actions),
                                       _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]),
                                       clip_coef, ent_coeff, v_coef, _ptr(grads), _ptr(stats), _ptr(vnew))
        assert rc == 0, rc
        return grads, stats, vnew

    def a2c_loss_raw(self, env_kind, params, idx, states, actions, returns):
        d = self.dims(env_kind)
        params = np.ascontiguousarray(params, np.float32)
        idx = np.ascontiguousarray(idx, np.int32)
        states = np.ascontiguousarray(states, np.float32)
        actions = np.ascontiguousarray(actions, np.int32 if env_kind == _abi.CRL_ENV_CARTPOLE else np.float32)
        returns = np.ascontiguousarray(returns, np.float32)
        grads = np.zeros(d["P"], np.float32)
        stats = np.zeros(4, np.float64)
        rc = self.lib.orc_a2c_loss_raw(env_kind, _ptr(params), _ptr(idx), idx.shape[0], _ptr(states), _ptr(actions),
                                       _ptr(returns), _ptr(grads), _ptr(stats))
        assert rc == 0, rc
        return grads, stats

    def ppo_loss_phase(self, env_kind, params, idx, states, actions, logprobs, advantages, returns, values,
                       clip_coef, ent_coeff, v_coef, phase, io, vnew):
        """one phase of the loss on one shard; io (8 doubles) and vnew are updated in place"""
        d = self.dims(env_kind)
        params = np.ascontiguousarray(params, np.float32)
        idx = np.ascontiguousarray(idx, np.int32)
        states = np.ascontiguousarray(states, np.float32)
        actions = np.ascontiguousarray(actions, np.int32 if env_kind == _abi.CRL_ENV_CARTPOLE else np.float32)
        arrs = [np.ascontiguousarray(a, np.float32) for a in (logprobs, advantages, returns, values)]
        grads = np.zeros(d["P"], np.float64)
        assert io.dtype == np.float64 and io.size == 8 and vnew.dtype == np.float32 and vnew.size == idx.size
        rc = self.lib.orc_ppo_loss_phase(env_kind, _ptr(params), _ptr(idx), idx.shape[0], _ptr(states), _ptr(actions),
                                         _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]),
                                         clip_coef, ent_coeff, v_coef, phase, _ptr(io), _ptr(vnew), _ptr(grads))
        assert rc == 0, rc
        return grads

    def clip_adam_raw(self, env_kind, params, grads, m, v, beta_pow, lr, clip_norm):
        params = np.array(params, np.float32, copy=True)
        m = np.array(m, np.float32, copy=True)
        v = np.array(v, np.float32, copy=True)
        beta_pow = np.array(beta_pow, np.float64, copy=True)
        grads = np.ascontiguousarray(grads, np.float32)
        rc = self.lib.orc_clip_adam_raw(env_kind, _ptr(params), _ptr(grads), _ptr(m), _ptr(v), _ptr(beta_pow), lr, clip_norm)
        assert rc == 0, rc
        return params, m, v, beta_pow

    def perm_index_keys(self, seed, update_index, epoch, rank):
        keys = np.zeros(8, np.uint32)
        self.lib.orc_perm_keys.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        self.lib.orc_perm_keys(seed, update_index, epoch, rank, _ptr(keys))
        return keys

    def create(self, cfg):
        return OracleCtx(self, cfg)


class OracleCtx:
    """Mirror of the crl_ctx handle API on the CPU oracle."""

    def __init__(self, olib, cfg):
        self.o = olib
        self.L = olib.lib
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.L.orc_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise ValueError("orc_create failed: %d" % rc)
        self.h = h
        self.d = olib.dims(cfg.env_kind)
        self.N, self.T = cfg.num_envs, cfg.num_steps
        self.B = self.N * self.T

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle call failed: %d" % rc)

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float32)
        self._ck(self.L.orc_set_params(self.h, _ptr(p), p.size))

    def get_params(self):
        p = np.zeros(self.d["P"], np.float32)
        self._ck(self.L.orc_get_params(self.h, _ptr(p), p.size))
        return p

    def get_grads(self):
        p = np.zeros(self.d["P"], np.float32)
        self._ck(self.L.orc_get_grads(self.h, _ptr(p), p.size))
        return p

    def get_adam_state(self):
        m = np.zeros(self.d["P"], np.float32)
        v = np.zeros(self.d["P"], np.float32)
        bp = np.zeros((self.d["n_arrays"], 2), np.float64)
        self._ck(self.L.orc_get_adam_state(self.h, _ptr(m), _ptr(v), _ptr(bp)))
        return m, v, bp

    def set_adam_state(self, m, v, bp):
        m = np.ascontiguousarray(m, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        bp = np.ascontiguousarray(bp, np.float64)
        self._ck(self.L.orc_set_adam_state(self.h, _ptr(m), _ptr(v), _ptr(bp)))

    def env_reset(self):
        self._ck(self.L.orc_env_reset(self.h))

    def env_set_state(self, state, t=None):
        state = np.ascontiguousarray(state, np.float32)
        t = None if t is None else np.ascontiguousarray(t, np.int32)
        self._ck(self.L.orc_env_set_state(self.h, _ptr(state), _ptr(t)))

    def rollout(self, action_noise=None, reset_noise=None):
        an = None if action_noise is None else np.ascontiguousarray(action_noise, np.float64)
        rn = None if reset_noise is None else np.ascontiguousarray(reset_noise, np.float32)
        self._ck(self.L.orc_rollout(self.h, _ptr(an), _ptr(rn)))

    def gae(self):
        self._ck(self.L.orc_gae(self.h))

    def update_minibatch(self, idx, lr):
        idx = np.ascontiguousarray(idx, np.int32)
        st = _abi.crl_loss_stats()
        self._ck(self.L.orc_update_minibatch(self.h, _ptr(idx), idx.size, lr, C.byref(st)))
        return st

    def update_epochs(self, perms, lr):
        n = self.cfg.update_epochs * self.cfg.num_minibatches
        st = (_abi.crl_loss_stats * n)()
        perms = None if perms is None else np.ascontiguousarray(perms, np.int32)
        self._ck(self.L.orc_update_epochs(self.h, _ptr(perms), lr, st))
        return np.array([[s.loss, s.pg_loss, s.v_loss, s.entropy_loss] for s in st])

    def train_update(self, lr):
        n = self.cfg.update_epochs * self.cfg.num_minibatches
        st = (_abi.crl_loss_stats * n)()
        self._ck(self.L.orc_train_update(self.h, lr, st))
        return np.array([[s.loss, s.pg_loss, s.v_loss, s.entropy_loss] for s in st])

    def device_permutation(self, update_index, epoch):
        out = np.zeros(self.B, np.int32)
        self._ck(self.L.orc_device_permutation(self.h, update_index, epoch, _ptr(out)))
        return out

    def field_shape(self, field):
        d, N, T = self.d, self.N, self.T
        cont = self.cfg.env_kind == _abi.CRL_ENV_PENDULUM
        return {
            _abi.CRL_F_STATE: (T, N, d["D"]),
            _abi.CRL_F_ACTION: (T, N, d["A"]) if cont else (T, N),
            _abi.CRL_F_NEXT_OBS: (N, d["D"]), _abi.CRL_F_ENV_STATE: (N, d["S"]),
            _abi.CRL_F_VNEW: (self.B // self.cfg.num_minibatches,),
        }.get(field, (T, N) if field <= _abi.CRL_F_RETURN else (N,))

    def field_dtype(self, field):
        if field == _abi.CRL_F_ACTION:
            return np.float32 if self.cfg.env_kind == _abi.CRL_ENV_PENDULUM else np.int32
        return _FIELD_DTYPES[field]

    def read_field(self, field):
        a = np.zeros(self.field_shape(field), self.field_dtype(field))
        self._ck(self.L.orc_read_field(self.h, field, _ptr(a), a.nbytes))
        return a

    def write_field(self, field, a):
        a = np.ascontiguousarray(a, self.field_dtype(field))
        assert a.shape == tuple(self.field_shape(field)), (a.shape, self.field_shape(field))
        self._ck(self.L.orc_write_field(self.h, field, _ptr(a), a.nbytes))

    def pop_episodes(self, max_records=1 << 20):
        recs = (_abi.crl_episode * max_records)()
        n = C.c_int32()
        agg = _abi.crl_episode_agg()
        self._ck(self.L.orc_pop_episodes(self.h, recs, max_records, C.byref(n), C.byref(agg)))
        return [(r.step, r.env, r.length, r.episode_return) for r in recs[:n.value]], agg


class OracleDQN:
    """Mirror of the crl_dqn_* API on the CPU oracle (TEST INFRASTRUCTURE, like everything in this module)."""

    def __init__(self, olib, cfg):
        self.L = olib.lib
        self.cfg = cfg
        self.L.orc_dqn_linear_schedule.restype = C.c_double
        self.L.orc_dqn_linear_schedule.argtypes = [C.c_double] * 4
        self.L.orc_dqn_run.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        self.L.orc_dqn_loss_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        for f in ("orc_dqn_destroy", "orc_dqn_reset"):
            getattr(self.L, f).argtypes = [C.c_void_p]
        self.L.orc_dqn_set_params.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        self.L.orc_dqn_get_params.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        self.L.orc_dqn_read_buffer.argtypes = [C.c_void_p] * 8
        self.L.orc_dqn_forward_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        h = C.c_void_p()
        if self.L.orc_dqn_create(C.byref(cfg), C.byref(h)) != 0:
            raise ValueError("orc_dqn_create failed")
        self.h = h

    def close(self):
        if self.h:
            self.L.orc_dqn_destroy(self.h)
            self.h = None

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float32)
        assert self.L.orc_dqn_set_params(self.h, _ptr(p), p.size) == 0

    def get_params(self):
        q = np.zeros(_abi.CRL_DQN_PARAMS, np.float32)
        t = np.zeros(_abi.CRL_DQN_PARAMS, np.float32)
        assert self.L.orc_dqn_get_params(self.h, _ptr(q), _ptr(t), q.size) == 0
        return q, t

    def reset(self):
        assert self.L.orc_dqn_reset(self.h) == 0

    def run(self, iterations):
        st = _abi.crl_dqn_stats()
        assert self.L.orc_dqn_run(self.h, int(iterations), C.byref(st)) == 0
        return st

    def set_shard(self, world, rank, env_id_base):
        """make this context rank `rank` of a data-parallel world (it then only steps through dqn_group_run)"""
        self.L.orc_dqn_set_shard.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        assert self.L.orc_dqn_set_shard(self.h, int(world), int(rank), int(env_id_base)) == 0

    @staticmethod
    def group_run(shards, iterations):
        """orc_dqn_group_run: all shards of a data-parallel world in lockstep; returns one crl_dqn_stats per shard"""
        L = shards[0].L
        k = len(shards)
        hs = (C.c_void_p * k)(*[s.h for s in shards])
        st = (_abi.crl_dqn_stats * k)()
        L.orc_dqn_group_run.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
        rc = L.orc_dqn_group_run(hs, k, int(iterations), st)
        if rc != 0:
            raise ValueError("orc_dqn_group_run failed: %d" % rc)
        return list(st)

    def read_buffer(self):
        cap = self.cfg.buffer_size
        out = {"state": np.zeros((cap, 4), np.float32), "action": np.zeros(cap, np.int32), "reward": np.zeros(cap, np.float32),
               "next_state": np.zeros((cap, 4), np.float32), "terminal": np.zeros(cap, np.uint8)}
        size, ptr = C.c_int32(), C.c_int32()
        assert self.L.orc_dqn_read_buffer(self.h, _ptr(out["state"]), _ptr(out["action"]), _ptr(out["reward"]),
                                          _ptr(out["next_state"]), _ptr(out["terminal"]), C.byref(size), C.byref(ptr)) == 0
        out["size"], out["ptr"] = size.value, ptr.value
        return out

    def linear_schedule(self, s, e, d, t):
        return self.L.orc_dqn_linear_schedule(s, e, d, t)

    def forward(self, params, obs):
        params = np.ascontiguousarray(params, np.float32)
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 4)
        q = np.zeros((obs.shape[0], 2), np.float32)
        assert self.L.orc_dqn_forward_raw(_ptr(params), _ptr(obs), _ptr(q), obs.shape[0]) == 0
        return q

    def loss_raw(self, q_params, tgt_params, state, action, reward, next_state, terminal, gamma):
        B = len(action)
        arrs = [np.ascontiguousarray(q_params, np.float32), np.ascontiguousarray(tgt_params, np.float32),
                np.ascontiguousarray(state, np.float32), np.ascontiguousarray(action, np.int32),
                np.ascontiguousarray(reward, np.float32), np.ascontiguousarray(next_state, np.float32),
                np.ascontiguousarray(terminal, np.uint8)]
        g = np.zeros(_abi.CRL_DQN_PARAMS, np.float32)
        loss = C.c_double()
        assert self.L.orc_dqn_loss_raw(_ptr(arrs[0]), _ptr(arrs[1]), B, _ptr(arrs[2]), _ptr(arrs[3]), _ptr(arrs[4]), _ptr(arrs[5]),
                                       _ptr(arrs[6]), float(gamma), _ptr(g), C.byref(loss)) == 0
        return g, loss.value
