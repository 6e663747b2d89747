"""Short, graph-free workload for ncu: 2 PPO updates at BASELINE configs[1] (CartPole, 4096 envs
x 128 steps) and one GAE launch at N=2^20 (2.29 GB); `dqn`: 40 DQN iterations at BASELINE configs[4]. Used by the ncu
recipes in profiles/README.md."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CRL_NO_GRAPH"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cleanrl_jl_b200 import _abi, _lib, networks  # noqa: E402
from cleanrl_jl_b200.handle import PPOHandle  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.cuda.set_device(0)
if what in ("all", "ppo"):
    cfg = _abi.make_config(num_envs=4096, num_steps=128, num_minibatches=4, update_epochs=4, seed=1)
    h = PPOHandle(cfg)
    h.set_params(networks.init_params(False, 4, 2, seed=1))
    h.env_reset()
    for _ in range(2):
        h.train_update(2.5e-4)
    h.sync()
    h.close()
if what in ("all", "gae"):
    T, N = 128, 1 << 20
    lib = _lib.load()
    v = torch.randn((T, N), device="cuda")
    r = torch.randn((T, N), device="cuda")
    d = (torch.rand((T, N), device="cuda") < 0.01).to(torch.uint8)
    adv = torch.empty((T, N), device="cuda")
    ret = torch.empty((T, N), device="cuda")
    for _ in range(2):
        _lib.check(lib.crl_gae_raw(_lib.ptr(v), _lib.ptr(r), _lib.ptr(d), None, None, _lib.ptr(adv), _lib.ptr(ret),
                                   T, N, 0.99, 0.95, 0, None))
    torch.cuda.synchronize()
if what == "dqn":
    from cleanrl_jl_b200.dqn_algo import DQNHandle, init_q_params  # noqa: E402
    cfg = _abi.make_dqn_config(num_envs=4096, buffer_size=1 << 20, min_buff_size=10_000, batch_size=120, train_freq=10,
                               target_net_freq=100, epsilon_duration=5e6, seed=1)
    h = DQNHandle(cfg)
    h.set_params(init_q_params(1))
    h.reset()
    h.run(40)   # 4 acting launches of 10 iterations + 4 learning steps
    h.close()
print("profile target done")
