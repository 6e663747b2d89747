"""bench.py plumbing that needs no GPU: both arms describe the same workload, and the roofline `traffic` figures come from
the newest committed ncu summary that holds the kernel (profiles/)."""
import argparse
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _args(**kw):
    d = dict(gpus=1, steps=20, warmup=5, impl="ours", envs_per_gpu=4096, env="CartPole", no_cpu_baseline=False,
             no_secondary=False, gae_n=1 << 20, local_stats=False, dqn_mode="replicas", algo="ppo")
    d.update(kw)
    return argparse.Namespace(**d)


def test_both_arms_state_the_same_config():
    b = _bench()
    for gpus in (1, 8):
        ours = b.workload_config(_args(gpus=gpus, impl="ours"), gpus)
        ref = b.workload_config(_args(gpus=gpus, impl="reference"), gpus)
        assert ours == ref
        assert "workload" in ours and "model" not in ours
        assert ours["num_envs_global"] == 4096 * gpus


def test_roofline_traffic_comes_from_the_committed_ncu_summaries():
    b = _bench()
    tc, tc_file = b.ncu_traffic("loss_grad_tc_kernel")
    assert tc_file and tc_file.startswith("profiles/") and "ncu" in tc_file
    assert 5e6 < tc < 5e7          # the 16.5 MB rollout buffer read once under ncu (cold L2), per launch
    gae, gae_file = b.ncu_traffic("gae_kernel<0, 4, 4>")
    assert gae_file and 2.0e9 < gae < 2.4e9   # 2.29 GB algorithmic at T=128, N=2^20
    assert b.ncu_traffic("no_such_kernel") == (None, None)
