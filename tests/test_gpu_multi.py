"""Multi-GPU parity (needs >= 2 B200s; skipped on a single-GPU box): 2 ranks under torchrun/NCCL against the
single-process oracle on the union minibatches (SURVEY §4 tier 3)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_multi_gpu_sharded_update_matches_single_process_oracle(tmp_path, torch_cuda):
    ngpu = torch_cuda.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2|4|8)")
    nproc = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)  # one rank per GPU
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
           "--master-port", "29544", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    res = json.load(open(out))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from multi_gpu_checks import verdict
    assert set(res) == {"kind0", "kind1"} and all("normal" in r and "fail" in r for r in res.values()), res
    # union-of-shards rollout == single-process oracle; ranks bit-identical; post-step parameters and loss statistics vs
    # the oracle on the union minibatches; speculative == exact; a forced speculation failure is replayed exactly
    assert verdict(res) == [], res


def test_multi_gpu_sharded_dqn_matches_oracle_group(tmp_path, torch_cuda):
    """data-parallel DQN (crl_dqn_comm_init): every rank's shard against the oracle's group run of all shards"""
    ngpu = torch_cuda.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2|4|8)")
    nproc = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    out = tmp_path / "dqn.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
           "--master-port", "29546", os.path.join(ROOT, "tests", "dqn_multi_worker.py"), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    r = json.load(open(out))
    assert r["learn_steps"] > 10, r
    assert r["exact"], r                    # actions, rewards, terminals, ring order, counters: bit-exact per shard
    assert r["ranks_agree"], r              # replicated parameters stay bit-identical across ranks
    assert r["param_maxerr"] < 2e-5, r      # vs the oracle's rank-ordered gradient sum
    assert r["loss_relerr"] < 1e-4, r
