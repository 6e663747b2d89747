"""Subprocess body of tests/test_gpu_tc.py: one BASELINE-size PPO minibatch + one full update on cuda:0, results to an
.npz. Run once with CRL_NO_TC=1 (FFMA update kernel) and once without (tcgen05 update kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from cleanrl_jl_b200 import _abi, networks  # noqa: E402
from cleanrl_jl_b200.handle import PPOHandle  # noqa: E402

out, kind = sys.argv[1], int(sys.argv[2])
N, T = 4096, 128
cfg = _abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=4, update_epochs=4, seed=11)
h = PPOHandle(cfg)
D, A = (4, 2) if kind == 0 else (3, 1)
h.set_params(networks.init_params(kind == 1, D, A, seed=5))
h.env_reset()
h.rollout()
h.gae()
res = {}
for name in ("STATE", "ACTION", "ADVANTAGE", "RETURN"):
    res[name] = h.read_field(getattr(_abi, "CRL_F_" + name))
idx = np.random.default_rng(0).permutation(N * T)[: N * T // 4].astype(np.int32)
s = h.update_minibatch(idx, 2.5e-4)
res["mb_stats"] = np.array([s.loss, s.pg_loss, s.v_loss, s.entropy_loss])
res["mb_grads"] = h.get_grads()
res["mb_params"] = h.get_params()
h.train_update(2.5e-4)
st, _ = h.fetch_update()
res["upd_stats"] = np.asarray(st)
res["upd_params"] = h.get_params()
res["replays"] = np.array([h.spec_replays()])
h.close()
np.savez(out, **res)
