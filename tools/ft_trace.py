"""development aid: one BASELINE-size update with the -DFT_TRACE build (prints the fused tail's phase clocks)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cleanrl_jl_b200 import _abi, networks
from cleanrl_jl_b200.handle import PPOHandle
cfg = _abi.make_config(env_kind=0, num_envs=4096, num_steps=128, num_minibatches=4, update_epochs=1, seed=1)
h = PPOHandle(cfg)
h.set_params(networks.init_params(False, 4, 2, seed=1))
h.env_reset()
os.environ["CRL_NO_GRAPH"] = "1"
for _ in range(2):
    h.train_update(2.5e-4)
h.sync()
h.close()
