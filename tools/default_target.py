"""ncu target: 3 updates at PPOConfig() defaults (4 envs x 32 steps, 16 minibatches of 32), graph-free"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CRL_NO_GRAPH"] = "1"
from cleanrl_jl_b200 import networks
from cleanrl_jl_b200.config import PPOConfig
from cleanrl_jl_b200.handle import PPOHandle
from cleanrl_jl_b200.ppo_algo import make_crl_config
d = PPOConfig()
h = PPOHandle(make_crl_config(d, d.num_envs, 0, 1, 0, 0))
h.set_params(networks.init_params(False, 4, 2, seed=1))
h.env_reset()
for _ in range(3):
    h.train_update(2.5e-4)
h.sync()
h.close()
