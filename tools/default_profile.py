"""development aid: where the time of ppo(PPOConfig()) goes (4 envs x 32 steps, 16 minibatches of 32): device time per
update (graph launches back to back, one sync), host time per stage of the public loop, cProfile of the loop."""
import cProfile, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleanrl_jl_b200 import _abi, networks, logger as Logger
from cleanrl_jl_b200.config import PPOConfig
from cleanrl_jl_b200.handle import PPOHandle
from cleanrl_jl_b200.ppo_algo import ppo, make_crl_config

d = PPOConfig()
cfg = make_crl_config(d, d.num_envs, 0, 1, 0, 0)
h = PPOHandle(cfg)
h.set_params(networks.init_params(False, 4, 2, seed=1))
h.env_reset()
for _ in range(20):
    h.train_update(2.5e-4)
h.sync()
n = 2000
t0 = time.perf_counter()
for _ in range(n):
    h.train_update(2.5e-4)
t1 = time.perf_counter()
h.sync()
t2 = time.perf_counter()
print("device-bound loop: %.3f ms per update (enqueue alone %.3f ms)" % ((t2 - t0) / n * 1e3, (t1 - t0) / n * 1e3))
h.close()
tmp = tempfile.mkdtemp(prefix="crl_prof_")
lg = Logger.make_logger("prof_default", to_terminal=False, to_tensorboard=True, log_dir=tmp)
pr = cProfile.Profile()
pr.enable()
res = ppo(PPOConfig(total_timesteps=n * d.num_envs * d.num_steps), logger=lg, device=0)
pr.disable()
lg.close()
print("ppo(): %.3f ms per update, host %s" % (res["elapsed_s"] / res["num_updates"] * 1e3, {k: round(v / res["num_updates"] * 1e3, 4) for k, v in res["host_s"].items()}))
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
