"""Committed fixtures: hand-derived KATs (tests/golden/kat.json) and frozen oracle outputs
(tests/golden/golden_v1.npz, written by tests/golden/make_golden.py). The CPU half pins the oracle to the
bytes in the repository; the GPU half checks the CUDA path against the same bytes through the C ABI."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(HERE, "golden", "kat.json")))


def test_oracle_matches_hand_derived_kats(olib, abi, kat):
    g = kat["gae"]
    v = np.array(g["values"], F)
    for name, mode in (("ref_compat", 0), ("fixed", 1)):
        for k in ("kat1", "kat2"):
            term = np.array(g[k + "_terminals"], np.uint8)
            adv, _ = olib.gae_raw(v[:4, None], np.array(g["rewards"], F)[:, None], term[:4, None], v[4:], term[4:],
                                  F(0.99), F(0.95), mode)
            exp = np.array(g["%s_%s" % (k, name)], F)
            np.testing.assert_array_equal(adv[:len(exp), 0], exp)
    c = kat["cartpole_step"]
    s, _, _, _ = olib.env_step_raw(0, np.array([c["state"]], F), [0], [c["action_julia_1_based"] - 1], 500)
    np.testing.assert_array_equal(s[0], np.array(c["next_state"], F))
    for case in kat["philox4x32_10"]["cases"]:
        assert list(olib.philox(case["ctr"], case["key"])) == case["out"]


def test_oracle_reproduces_frozen_outputs(olib, abi, gold):
    for mode in (0, 1):
        adv, ret = olib.gae_raw(gold["gae_values"], gold["gae_rewards"], gold["gae_dones"], gold["gae_next_value"],
                                gold["gae_next_done"], F(0.99), F(0.95), mode)
        np.testing.assert_array_equal(adv, gold["gae_adv_mode%d" % mode])
        np.testing.assert_array_equal(ret, gold["gae_ret_mode%d" % mode])
    for kind, name in ((0, "cartpole"), (1, "pendulum")):
        s1, _, rw, dn = olib.env_step_raw(kind, gold[name + "_s0"], np.zeros(24, np.int32), gold[name + "_a"],
                                          500 if kind == 0 else 200)
        np.testing.assert_allclose(s1, gold[name + "_s1"], rtol=1e-6, atol=1e-7)  # libm sin/cos may differ by an ulp
        np.testing.assert_array_equal(dn, gold[name + "_done"])
        pol, logp, val = olib.policy_forward_raw(kind, gold[name + "_params"], gold[name + "_obs"])
        np.testing.assert_allclose(pol, gold[name + "_pol"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(val, gold[name + "_val"], rtol=1e-6, atol=1e-6)
    cfg = abi.make_config(num_envs=25, num_steps=40, num_minibatches=1, seed=99)
    np.testing.assert_array_equal(olib.create(cfg).device_permutation(3, 2), gold["perm_1000_u3_e2"])
    u = np.stack([olib.reset_uniforms(99, e, k) for e in range(4) for k in range(3)])
    np.testing.assert_array_equal(u, gold["philox_reset_u"])


def test_oracle_dqn_reproduces_frozen_outputs():
    """tests/golden/golden_dqn_v1.npz (make_golden_dqn.py): the oracle's DQN, single process and two-shard group.
    Discrete outputs bit-exact; parameters to 1e-6 (libm sin/cos may differ by an ulp between hosts)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_dqn", os.path.join(HERE, "golden", "make_golden_dqn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now, gold = mod.generate(), np.load(os.path.join(HERE, "golden", "golden_dqn_v1.npz"))
    assert set(now) == set(gold.files)
    for k in gold.files:
        if k.endswith(("_action", "_terminal", "_reward")):
            np.testing.assert_array_equal(now[k], gold[k], err_msg=k)
        else:
            np.testing.assert_allclose(now[k], gold[k], rtol=1e-6, atol=1e-7, err_msg=k)
    assert gold["single_scalars"][5] > 10 and gold["group0_scalars"][3] > 10          # learning steps happened
    np.testing.assert_array_equal(gold["group0_q"], gold["group1_q"])                  # replicas agree


@pytest.mark.gpu
def test_cuda_path_matches_frozen_outputs(crl, abi, gold, torch_cuda):
    from cleanrl_jl_b200.handle import PPOHandle
    torch = torch_cuda
    lib = crl.load()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    T, N = gold["gae_values"].shape
    for mode in (0, 1):
        t = [dev(gold[k]) for k in ("gae_values", "gae_rewards", "gae_dones", "gae_next_value", "gae_next_done")]
        adv = torch.zeros((T, N), device="cuda")
        ret = torch.zeros((T, N), device="cuda")
        crl.check(lib.crl_gae_raw(*[crl.ptr(x) for x in t], crl.ptr(adv), crl.ptr(ret), T, N, 0.99, 0.95, mode, None))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(adv.cpu().numpy(), gold["gae_adv_mode%d" % mode])
        np.testing.assert_array_equal(ret.cpu().numpy(), gold["gae_ret_mode%d" % mode])
    for kind, name in ((0, "cartpole"), (1, "pendulum")):
        cfg = abi.make_config(env_kind=kind, num_envs=8, num_steps=8, num_minibatches=2, update_epochs=2, seed=7)
        h = PPOHandle(cfg)
        h.set_params(gold[name + "_params"])
        h.env_reset()
        h.rollout(gold[name + "_upd_an"], gold[name + "_upd_rn"])
        h.gae()
        np.testing.assert_array_equal(h.read_field(abi.CRL_F_TERMINAL), gold[name + "_upd_terminal"])
        if kind == 0:
            np.testing.assert_array_equal(h.read_field(abi.CRL_F_ACTION), gold[name + "_upd_action"])
        np.testing.assert_allclose(h.read_field(abi.CRL_F_STATE), gold[name + "_upd_state"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(h.read_field(abi.CRL_F_ADVANTAGE), gold[name + "_upd_adv"], rtol=1e-4, atol=2e-5)
        stats = h.update_epochs(gold[name + "_upd_perms"], 2.5e-4)
        np.testing.assert_allclose(stats, gold[name + "_upd_stats"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(h.get_params(), gold[name + "_upd_params"], rtol=1e-5, atol=2e-6)
        h.close()
    h = PPOHandle(abi.make_config(num_envs=25, num_steps=40, num_minibatches=1, seed=99))
    np.testing.assert_array_equal(h.device_permutation(3, 2), gold["perm_1000_u3_e2"])
    h.close()
