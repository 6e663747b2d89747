"""worker for tests/test_dist_gloo.py: one process per (CPU) rank, gloo backend, 127.0.0.1."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cleanrl_jl_b200 import _abi, parallel  # noqa: E402
from oracle.oracle import OracleLib  # noqa: E402
from conftest import rand_params  # noqa: E402
from test_oracle_grad import make_batch  # noqa: E402

F = np.float32


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert parallel.dist_info()[0] == rank and parallel.dist_info()[2] == world
    res = {"rank": rank, "world": world}
    olib = OracleLib()

    # 1. unique-id plumbing (the real id comes from crl_comm_unique_id on a GPU box)
    uid = parallel.exchange_unique_id(lambda: bytes(range(128)))
    res["uid_ok"] = uid == bytes(range(128))
    res["max"] = parallel.max_over_ranks(10.0 + rank)
    res["sum"] = parallel.sum_over_ranks([1.0, float(rank)]).tolist()

    # 2. env sharding: rank r owns envs [r*N/k, (r+1)*N/k); per-env Philox keys are global ids, so the
    #    union of the shards' rollouts + GAE is the single-process result
    kind, N, T, seed = 0, 16, 24, 77
    base, n_local = parallel.shard_envs(N, world, rank)
    p = rand_params(olib, kind, seed=2)
    cfg = _abi.make_config(env_kind=kind, num_envs=n_local, num_steps=T, num_minibatches=2, update_epochs=1, seed=seed,
                           world_size=world, rank=rank, env_id_base=base)
    o = olib.create(cfg)
    o.set_params(p)
    o.env_reset()
    o.rollout()
    o.gae()
    shard = {name: o.read_field(getattr(_abi, "CRL_F_" + name)) for name in
             ("STATE", "ACTION", "LOGPROB", "REWARD", "TERMINAL", "VALUE", "ADVANTAGE", "RETURN")}
    gathered = [None] * world
    dist.all_gather_object(gathered, shard)
    if rank == 0:
        cfg_g = _abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=2, update_epochs=1, seed=seed)
        g = olib.create(cfg_g)
        g.set_params(p)
        g.env_reset()
        g.rollout()
        g.gae()
        ok = True
        for name in shard:
            full = g.read_field(getattr(_abi, "CRL_F_" + name))
            cat = np.concatenate([s[name] for s in gathered], axis=1)
            ok = ok and np.array_equal(full, cat)
        res["shard_rollout_ok"] = bool(ok)

    # 3. the three-collective minibatch protocol (stats, count, gradient) over gloo vs the
    #    single-process loss on the union minibatch
    for small in (False, True):
        B, M_local = 256, 40
        pp = rand_params(olib, kind, seed=4)
        if small:
            pp[olib.param_layout(kind)[0][11]] = 1.5
        states, actions, logprobs, adv, ret, val = make_batch(olib, kind, B, 5, small)
        perm = np.random.default_rng(1).permutation(B)[:M_local * world].astype(np.int32)
        idx = perm[rank * M_local:(rank + 1) * M_local]
        c, ec, vc = float(F(0.2)), float(F(0.01)), float(F(0.5))
        io = np.zeros(8)
        vnew = np.zeros(M_local, F)
        args = (kind, pp, idx, states, actions, logprobs, adv, ret, val, c, ec, vc)
        olib.ppo_loss_phase(*args, 0, io, vnew)
        sums = parallel.sum_over_ranks(io[:3])                      # collective A
        Mg = float(M_local * world)
        mean = sums[0] / Mg
        var = max((sums[1] - Mg * mean * mean) / (Mg - 1), 0.0)
        io[3] = float(F(sums[2] / Mg))
        io[5], io[6], io[7] = float(F(mean)), float(F(np.sqrt(var))), Mg
        olib.ppo_loss_phase(*args, 1, io, vnew)
        io[4] = parallel.sum_over_ranks([io[4]])[0]                  # collective B
        g_local = olib.ppo_loss_phase(*args, 2, io, vnew)
        g_sum = parallel.sum_over_ranks(np.concatenate([g_local, io[:3]]))  # collective C
        if rank == 0:
            g_ref, st_ref, _ = olib.ppo_loss_raw(kind, pp, perm, states, actions, logprobs, adv, ret, val, c, ec, vc)
            scale = np.abs(g_ref).max()
            res["grad_maxerr_%s" % small] = float(np.abs(g_sum[:-3].astype(F) - g_ref).max() / scale)
            res["pg_err_%s" % small] = float(abs(g_sum[-3] / Mg - st_ref[1]))
            res["cnt_%s" % small] = float(io[4])
    if rank == 0:
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
