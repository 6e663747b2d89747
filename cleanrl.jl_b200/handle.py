"""PPOHandle — thin object wrapper over the crl_ctx handle API (one handle per GPU)."""
import ctypes as C

import numpy as np

from . import _abi
from . import _lib as L

_FIELD_DTYPES = {
    _abi.CRL_F_STATE: np.float32, _abi.CRL_F_LOGPROB: np.float32, _abi.CRL_F_REWARD: np.float32,
    _abi.CRL_F_TERMINAL: np.uint8, _abi.CRL_F_VALUE: np.float32, _abi.CRL_F_ADVANTAGE: np.float32,
    _abi.CRL_F_RETURN: np.float32, _abi.CRL_F_NEXT_OBS: np.float32, _abi.CRL_F_NEXT_DONE: np.uint8,
    _abi.CRL_F_NEXT_VALUE: np.float32, _abi.CRL_F_ENV_STATE: np.float32, _abi.CRL_F_ENV_T: np.int32,
    _abi.CRL_F_EP_RETURN: np.float64, _abi.CRL_F_EP_LENGTH: np.int32, _abi.CRL_F_RESET_COUNT: np.uint32,
    _abi.CRL_F_VNEW: np.float32,
}


def _stats_array(st):
    return np.array([[s.loss, s.pg_loss, s.v_loss, s.entropy_loss] for s in st], np.float64).reshape(-1, 4)


class PPOHandle:
    def __init__(self, cfg):
        self.lib = L.load()
        self.cfg = cfg
        h = C.c_void_p()
        L.check(self.lib.crl_create(C.byref(cfg), C.byref(h)))
        self.h = h
        v = [C.c_int32() for _ in range(5)]
        L.check(self.lib.crl_dims(self.h, *[C.byref(x) for x in v]))
        self.d = dict(zip("D A S P n_arrays".split(), [x.value for x in v]))
        self.N, self.T = cfg.num_envs, cfg.num_steps
        self.B = self.N * self.T
        self.M = self.B // cfg.num_minibatches
        self.n_mb = cfg.update_epochs * cfg.num_minibatches
        self.continuous = cfg.env_kind == _abi.CRL_ENV_PENDULUM

    # ---- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.lib.crl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        L.check(self.lib.crl_sync(self.h))

    # ---- parameters / optimiser
    def param_layout(self):
        n = self.d["n_arrays"]
        off = np.zeros(n, np.int32)
        size = np.zeros(n, np.int32)
        L.check(self.lib.crl_param_layout(self.h, L.ptr(off), L.ptr(size), n))
        return off, size

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float32)
        L.check(self.lib.crl_set_params(self.h, L.ptr(p), p.size))

    def get_params(self):
        p = np.zeros(self.d["P"], np.float32)
        L.check(self.lib.crl_get_params(self.h, L.ptr(p), p.size))
        return p

    def get_grads(self):
        p = np.zeros(self.d["P"], np.float32)
        L.check(self.lib.crl_get_grads(self.h, L.ptr(p), p.size))
        return p

    def get_adam_state(self):
        m = np.zeros(self.d["P"], np.float32)
        v = np.zeros(self.d["P"], np.float32)
        bp = np.zeros((self.d["n_arrays"], 2), np.float64)
        L.check(self.lib.crl_get_adam_state(self.h, L.ptr(m), L.ptr(v), L.ptr(bp)))
        return m, v, bp

    def set_adam_state(self, m, v, bp):
        m = np.ascontiguousarray(m, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        bp = np.ascontiguousarray(bp, np.float64)
        L.check(self.lib.crl_set_adam_state(self.h, L.ptr(m), L.ptr(v), L.ptr(bp)))

    # ---- envs / rollout / GAE
    def env_reset(self):
        L.check(self.lib.crl_env_reset(self.h))

    def env_set_state(self, state, t=None):
        state = np.ascontiguousarray(state, np.float32)
        assert state.shape == (self.N, self.d["S"])
        t = None if t is None else np.ascontiguousarray(t, np.int32)
        L.check(self.lib.crl_env_set_state(self.h, L.ptr(state), L.ptr(t)))

    def rollout(self, action_noise=None, reset_noise=None):
        an = None if action_noise is None else np.ascontiguousarray(action_noise, np.float64)
        rn = None if reset_noise is None else np.ascontiguousarray(reset_noise, np.float32)
        if an is not None:
            assert an.size == self.B * (self.d["A"] if self.continuous else 1)
        if rn is not None:
            assert rn.size == self.B * 4
        L.check(self.lib.crl_rollout(self.h, L.ptr(an), L.ptr(rn)))

    def gae(self):
        L.check(self.lib.crl_gae(self.h))

    # ---- update
    def update_minibatch(self, idx, lr):
        idx = np.ascontiguousarray(idx, np.int32)
        st = _abi.crl_loss_stats()
        L.check(self.lib.crl_update_minibatch(self.h, L.ptr(idx), idx.size, float(lr), C.byref(st)))
        return st

    def update_epochs(self, perms, lr):
        st = (_abi.crl_loss_stats * max(self.n_mb, 1))()
        if perms is not None:
            perms = np.ascontiguousarray(perms, np.int32)
            assert perms.shape == (self.cfg.update_epochs, self.B)
        L.check(self.lib.crl_update_epochs(self.h, L.ptr(perms), float(lr), st))
        return _stats_array(st[:self.n_mb])

    def device_permutation(self, update_index, epoch):
        out = np.zeros(self.B, np.int32)
        L.check(self.lib.crl_device_permutation(self.h, int(update_index), int(epoch), L.ptr(out)))
        return out

    def train_update(self, lr):
        """enqueue rollout + GAE + all epochs (asynchronous)"""
        L.check(self.lib.crl_train_update(self.h, float(lr)))

    def fetch_update(self, lag=0):
        """results of the latest update (lag=0) or of the one before it (lag=1, does not wait for the latest)"""
        st = np.empty((max(self.n_mb, 1), 4), np.float64)   # crl_loss_stats = 4 doubles (static_assert in api.cu)
        agg = _abi.crl_episode_agg()
        L.check(self.lib.crl_fetch_update_at(self.h, int(lag), st.ctypes.data_as(C.c_void_p), C.byref(agg)))
        return st[:self.n_mb], agg

    # ---- data
    def field_shape(self, field):
        d, N, T = self.d, self.N, self.T
        special = {
            _abi.CRL_F_STATE: (T, N, d["D"]),
            _abi.CRL_F_ACTION: (T, N, d["A"]) if self.continuous else (T, N),
            _abi.CRL_F_NEXT_OBS: (N, d["D"]), _abi.CRL_F_ENV_STATE: (N, d["S"]), _abi.CRL_F_VNEW: (self.M,),
        }
        if field in special:
            return special[field]
        return (T, N) if field <= _abi.CRL_F_RETURN else (N,)

    def field_dtype(self, field):
        if field == _abi.CRL_F_ACTION:
            return np.float32 if self.continuous else np.int32
        return _FIELD_DTYPES[field]

    def read_field(self, field):
        a = np.zeros(self.field_shape(field), self.field_dtype(field))
        L.check(self.lib.crl_read_field(self.h, field, L.ptr(a), a.nbytes))
        return a

    def write_field(self, field, a):
        a = np.ascontiguousarray(a, self.field_dtype(field))
        assert a.shape == tuple(self.field_shape(field)), (a.shape, self.field_shape(field))
        L.check(self.lib.crl_write_field(self.h, field, L.ptr(a), a.nbytes))

    def pop_episodes(self, max_records=1 << 20):
        recs = (_abi.crl_episode * max_records)()
        n = C.c_int32()
        agg = _abi.crl_episode_agg()
        L.check(self.lib.crl_pop_episodes(self.h, recs, max_records, C.byref(n), C.byref(agg)))
        return [(r.step, r.env, r.length, r.episode_return) for r in recs[:n.value]], agg

    # ---- multi-GPU / instrumentation
    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        L.check(self.lib.crl_comm_init(self.h, buf))

    def kernel_launches(self):
        n = C.c_uint64()
        L.check(self.lib.crl_kernel_launches(self.h, C.byref(n)))
        return n.value

    def spec_replays(self):
        n = C.c_uint64()
        L.check(self.lib.crl_spec_replays(self.h, C.byref(n)))
        return n.value

    def profile(self, enable):
        L.check(self.lib.crl_profile(self.h, 1 if enable else 0))

    def profile_read(self, reset=True):
        kt = _abi.crl_kernel_times()
        L.check(self.lib.crl_profile_read(self.h, C.byref(kt), 1 if reset else 0))
        return {name: {"ms": kt.ms[i], "launches": kt.launches[i]} for i, name in enumerate(_abi.KERNEL_NAMES)}

    def stream(self):
        s = C.c_void_p()
        L.check(self.lib.crl_stream(self.h, C.byref(s)))
        return s.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    L.check(L.load().crl_comm_unique_id(buf))
    return buf.raw
