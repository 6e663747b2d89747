"""worker for tests/test_gpu_multi.py::test_multi_gpu_sharded_dqn_matches_oracle_group: one process per GPU (torchrun,
NCCL). Every rank owns n_local envs, its own ring and batch_size samples of each learning step (crl_dqn_comm_init);
the result must match the oracle's data-parallel group run (orc_dqn_group_run) of all shards: actions, rewards,
terminals and ring order bit-exact for this rank's shard, parameters within fp32 tolerance (the allreduce may sum
the shards in another order than the oracle's rank order), replicas bit-identical across ranks."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cleanrl_jl_b200 import _abi, parallel  # noqa: E402
from cleanrl_jl_b200.dqn_algo import DQNHandle, init_q_params  # noqa: E402
from cleanrl_jl_b200.handle import comm_unique_id  # noqa: E402
from oracle.oracle import OracleDQN, OracleLib  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, local_rank, world = parallel.dist_info()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    olib = OracleLib()
    n_local, iters = 16, 90
    kw = dict(num_envs=n_local, buffer_size=1024, min_buff_size=64, batch_size=24, train_freq=4, target_net_freq=12,
              epsilon_duration=float(iters * n_local * world), seed=13)
    p = init_q_params(7)
    h = DQNHandle(_abi.make_dqn_config(device=local_rank, **kw))
    h.set_params(p)
    h.comm_init(parallel.exchange_unique_id(comm_unique_id), world, rank, rank * n_local)
    h.reset()
    shards = []
    for r in range(world):
        o = OracleDQN(olib, _abi.make_dqn_config(**kw))
        o.set_shard(world, r, r * n_local)
        o.set_params(p)
        o.reset()
        shards.append(o)
    res = {"exact": True, "param_maxerr": 0.0, "loss_relerr": 0.0}
    for chunk in (30, 30, 30):
        sh = h.run(chunk)
        so = OracleDQN.group_run(shards, chunk)[rank]
        bh, bo = h.read_buffer(), shards[rank].read_buffer()
        n = bo["size"]
        ok = (sh.iterations, sh.learn_steps, sh.episodes, sh.epsilon, sh.sum_return) == \
             (so.iterations, so.learn_steps, so.episodes, so.epsilon, so.sum_return)
        ok = ok and (bh["size"], bh["ptr"]) == (bo["size"], bo["ptr"])
        for f in ("action", "terminal", "reward"):
            ok = ok and np.array_equal(bh[f][:n], bo[f][:n])
        ok = ok and np.allclose(bh["state"][:n], bo["state"][:n], rtol=1e-5, atol=2e-6)
        res["exact"] = bool(res["exact"] and ok)
        qh, th = h.get_params()
        qo, to = shards[rank].get_params()
        res["param_maxerr"] = max(res["param_maxerr"], float(np.abs(qh - qo).max()), float(np.abs(th - to).max()))
        if so.learn_steps:
            res["loss_relerr"] = max(res["loss_relerr"], abs(sh.last_loss - so.last_loss) / max(abs(so.last_loss), 1e-12))
        res["learn_steps"] = int(so.learn_steps)
    # replicas identical across ranks
    q = torch.tensor(h.get_params()[0], device="cuda")
    lo, hi = q.clone(), q.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["ranks_agree"] = bool(torch.equal(lo, hi))
    flags = torch.tensor([float(res["exact"]), res["param_maxerr"], res["loss_relerr"]], dtype=torch.float64, device="cuda")
    worst = flags.clone()
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    best = flags.clone()
    dist.all_reduce(best, op=dist.ReduceOp.MIN)
    h.close()
    if rank == 0:
        res.update(exact=bool(best[0].item() == 1.0), param_maxerr=float(worst[1].item()), loss_relerr=float(worst[2].item()),
                   world=world)
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
