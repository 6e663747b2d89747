"""Generates tests/golden/golden_dqn_v1.npz: frozen outputs of the oracle's DQN restatement (single process and a
two-shard data-parallel group) on a small seeded configuration, so that refactors of oracle/ppo_oracle.c are checked
against bytes in the repository (the reference ships no golden vectors).   Run: python tests/golden/make_golden_dqn.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cleanrl_jl_b200 import _abi  # noqa: E402
from cleanrl_jl_b200.dqn_algo import init_q_params  # noqa: E402
from oracle.oracle import OracleDQN, OracleLib  # noqa: E402

KW = dict(buffer_size=256, min_buff_size=32, batch_size=8, train_freq=4, target_net_freq=12, epsilon_duration=600.0, seed=21)
ITERS = 80


def run_single(olib):
    o = OracleDQN(olib, _abi.make_dqn_config(num_envs=6, **KW))
    o.set_params(init_q_params(11))
    o.reset()
    st = o.run(ITERS)
    q, t = o.get_params()
    b = o.read_buffer()
    o.close()
    return dict(single_q=q, single_tgt=t, single_action=b["action"], single_terminal=b["terminal"], single_reward=b["reward"],
                single_state=b["state"], single_scalars=np.array([st.last_loss, st.sum_return, st.sum_length, st.epsilon,
                                                                  st.episodes, st.learn_steps, b["size"], b["ptr"]], np.float64))


def run_group(olib):
    shards = []
    for r in range(2):
        o = OracleDQN(olib, _abi.make_dqn_config(num_envs=3, **KW))
        o.set_shard(2, r, 3 * r)
        o.set_params(init_q_params(11))
        o.reset()
        shards.append(o)
    st = OracleDQN.group_run(shards, ITERS)
    out = {}
    for r, o in enumerate(shards):
        q, _ = o.get_params()
        b = o.read_buffer()
        out["group%d_q" % r] = q
        out["group%d_action" % r] = b["action"]
        out["group%d_terminal" % r] = b["terminal"]
        out["group%d_scalars" % r] = np.array([st[r].last_loss, st[r].sum_return, st[r].episodes, st[r].learn_steps, b["size"]], np.float64)
        o.close()
    return out


def generate():
    olib = OracleLib()
    out = run_single(olib)
    out.update(run_group(olib))
    return out


if __name__ == "__main__":
    arrays = generate()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_dqn_v1.npz"), **arrays)
    print("wrote golden_dqn_v1.npz with", len(arrays), "arrays")
