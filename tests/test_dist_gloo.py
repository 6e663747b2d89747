"""N>1 host path on CPU: world_size-2 gloo processes exercise the sharding plan, the
unique-id / timing plumbing and the stats -> count -> gradient exchange protocol (with the
oracle as the per-shard compute) and compare against the single-process result."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_and_exchange_protocol(tmp_path):
    out = tmp_path / "res.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    res = json.load(open(out))
    assert res["world"] == 2 and res["uid_ok"]
    assert res["max"] == 11.0 and res["sum"] == [2.0, 1.0]
    assert res["shard_rollout_ok"]                      # union of shards == single-process rollout + GAE, bit-exact
    assert res["grad_maxerr_False"] < 1e-6 and res["grad_maxerr_True"] < 1e-6
    assert res["pg_err_False"] < 1e-12 and res["pg_err_True"] < 1e-12
    assert res["cnt_False"] == 0 and res["cnt_True"] > 0  # the count collective carries real data in the Q5 case
