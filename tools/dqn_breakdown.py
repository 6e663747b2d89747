"""Per-kernel cost of the DQN path from three schedules on one GPU (no profiler): acting only, the reference
schedule (learn every 10 iterations), and a learning step after every iteration. Prints microseconds per iteration
and the implied cost of one acting step and one learning step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleanrl_jl_b200 import _abi  # noqa: E402
from cleanrl_jl_b200.dqn_algo import DQNHandle, init_q_params  # noqa: E402


def run(train_freq, min_buff, iters=2000, N=4096):
    cfg = _abi.make_dqn_config(num_envs=N, buffer_size=1 << 20, min_buff_size=min_buff, batch_size=120, train_freq=train_freq,
                               target_net_freq=100, epsilon_duration=5e6, seed=1)
    h = DQNHandle(cfg)
    h.set_params(init_q_params(1))
    h.reset()
    h.run(300)
    t0 = time.perf_counter()
    st = h.run(iters)
    dt = time.perf_counter() - t0
    l0 = st.kernel_launches
    h.close()
    return dt / iters * 1e6, st


if __name__ == "__main__":
    act, _ = run(10, 2**31 - 1)
    ref, st = run(10, 10_000)
    every, st1 = run(1, 10_000)
    print("act only           : %.2f us/iteration" % act)
    print("reference schedule : %.2f us/iteration -> learn step ~ %.1f us" % (ref, (ref - act) * 10))
    print("learn every step   : %.2f us/iteration -> single-step act launch + learn ~ %.1f us" % (every, every))
