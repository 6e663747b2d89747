"""Generates tests/golden/golden_v1.npz.

The reference (Julia) cannot run in this image and ships no golden vectors, so the committed fixtures are
(a) the hand-derived known-answer vectors of SURVEY §8c (kat.json, written by hand, NOT by this script) and
(b) outputs of the KAT-pinned CPU oracle on small seeded inputs, frozen here so that both the oracle and the
CUDA path are regression-checked against bytes in the repository.   Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cleanrl_jl_b200 import _abi  # noqa: E402
from oracle.oracle import OracleLib  # noqa: E402
from conftest import rand_params  # noqa: E402

F = np.float32


def main():
    olib = OracleLib()
    out = {}
    rng = np.random.default_rng(20261017)
    # GAE, both modes
    T, N = 16, 12
    v, r = rng.standard_normal((T, N)).astype(F), rng.standard_normal((T, N)).astype(F)
    d = (rng.random((T, N)) < 0.15).astype(np.uint8)
    nv, nd = rng.standard_normal(N).astype(F), (rng.random(N) < 0.15).astype(np.uint8)
    out.update(gae_values=v, gae_rewards=r, gae_dones=d, gae_next_value=nv, gae_next_done=nd)
    for mode in (0, 1):
        adv, ret = olib.gae_raw(v, r, d, nv, nd, F(0.99), F(0.95), mode)
        out["gae_adv_mode%d" % mode], out["gae_ret_mode%d" % mode] = adv, ret
    for kind, name in ((0, "cartpole"), (1, "pendulum")):
        dd = olib.dims(kind)
        p = rand_params(olib, kind, seed=42)
        if kind == 1:
            p[-1] = -0.4
        n = 24
        s0 = (rng.standard_normal((n, dd["S"])) * (0.08 if kind == 0 else 1.5)).astype(F)
        a = rng.integers(0, 2, n).astype(np.int32) if kind == 0 else rng.uniform(-2.5, 2.5, n).astype(F)
        s1, t1, rw, dn = olib.env_step_raw(kind, s0, np.zeros(n, np.int32), a, 500 if kind == 0 else 200)
        obs = (rng.standard_normal((n, dd["D"])) * 0.5).astype(F)
        pol, logp, val = olib.policy_forward_raw(kind, p, obs)
        out.update({name + "_params": p, name + "_s0": s0, name + "_a": a, name + "_s1": s1, name + "_rew": rw,
                    name + "_done": dn, name + "_obs": obs, name + "_pol": pol, name + "_logp": logp, name + "_val": val})
        # one full update on a tiny config with injected noise and a host permutation
        Nn, Tt = 8, 8
        cfg = _abi.make_config(env_kind=kind, num_envs=Nn, num_steps=Tt, num_minibatches=2, update_epochs=2, seed=7)
        o = olib.create(cfg)
        o.set_params(p)
        o.env_reset()
        an = rng.random((Tt, Nn)) if kind == 0 else rng.standard_normal((Tt, Nn, 1))
        rn = rng.random((Tt, Nn, 4)).astype(F)
        perms = np.stack([rng.permutation(Nn * Tt) for _ in range(2)]).astype(np.int32)
        o.rollout(an, rn)
        o.gae()
        stats = o.update_epochs(perms, 2.5e-4)
        out.update({name + "_upd_an": an, name + "_upd_rn": rn, name + "_upd_perms": perms, name + "_upd_stats": stats,
                    name + "_upd_params": o.get_params(), name + "_upd_adv": o.read_field(_abi.CRL_F_ADVANTAGE),
                    name + "_upd_state": o.read_field(_abi.CRL_F_STATE), name + "_upd_action": o.read_field(_abi.CRL_F_ACTION),
                    name + "_upd_terminal": o.read_field(_abi.CRL_F_TERMINAL)})
    # device permutation (integer-only: bit-exact everywhere)
    cfg = _abi.make_config(num_envs=25, num_steps=40, num_minibatches=1, seed=99)
    o = olib.create(cfg)
    out["perm_1000_u3_e2"] = o.device_permutation(3, 2)
    out["philox_reset_u"] = np.stack([olib.reset_uniforms(99, e, k) for e in range(4) for k in range(3)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
    print("wrote golden_v1.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
