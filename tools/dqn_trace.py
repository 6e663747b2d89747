"""Runs 30 DQN iterations (3 acting launches, 3 learning steps) so that a -DDQN_TRACE build prints its clock stamps.
Build with `CRL_NVCC_EXTRA=-DDQN_TRACE python cleanrl.jl_b200/build.py --force` first."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleanrl_jl_b200 import _abi  # noqa: E402
from cleanrl_jl_b200.dqn_algo import DQNHandle, init_q_params  # noqa: E402

cfg = _abi.make_dqn_config(num_envs=4096, buffer_size=1 << 20, min_buff_size=10_000, batch_size=120, train_freq=10,
                           target_net_freq=100, epsilon_duration=5e6, seed=1)
h = DQNHandle(cfg)
h.set_params(init_q_params(1))
h.reset()
h.run(30)
h.close()
