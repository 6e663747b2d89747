"""Multi-GPU parity checks shared by tests/multi_gpu_worker.py (pytest, `gpurun --gpus N`) and bench.py (the
`multi_parity` object of every N > 1 bench line, so the driver's scaling record carries a correctness verdict).
One process per GPU (torchrun, NCCL already initialised by the caller). Each rank owns N/world envs; the sharded
update must reproduce the single-process ORACLE update on the union minibatches, the replicated parameters must stay
bit-identical across ranks, and a failed speculation must be replayed exactly."""
import os

import numpy as np

from cleanrl_jl_b200 import _abi, parallel
from cleanrl_jl_b200.handle import PPOHandle, comm_unique_id

F = np.float32


def rand_params(olib, env_kind, seed=0, scale=1.0):
    """same recipe as tests/conftest.py::rand_params (kept here so that bench.py does not import pytest's conftest)"""
    from cleanrl_jl_b200 import networks
    d = olib.dims(env_kind)
    rng = np.random.default_rng(seed)
    p = networks.init_params(env_kind == 1, d["D"], d["A"], seed=seed)
    p = p + (0.05 * scale * rng.standard_normal(p.shape)).astype(np.float32)
    off, size = olib.param_layout(env_kind)
    p[off[4]:off[4] + size[4]] *= 30.0
    return p.astype(np.float32)


def run_checks(dist, olib, kinds=(0, 1), scenarios=("normal", "fail"), N=64, T=16):
    """returns the result dict on rank 0 (None elsewhere); collective: every rank must call it"""
    rank, local_rank, world = parallel.dist_info()
    res = {}
    for kind in kinds:
        mb, epochs, seed = 4, 2, 31
        base, n_local = parallel.shard_envs(N, world, rank)
        p = rand_params(olib, kind, seed=3)
        if kind == 1:
            p[-1] = -0.4
        cfg = _abi.make_config(env_kind=kind, num_envs=n_local, num_steps=T, num_minibatches=mb, update_epochs=epochs,
                               seed=seed, device=local_rank, world_size=world, rank=rank, env_id_base=base)
        h = PPOHandle(cfg)
        h.comm_init(parallel.exchange_unique_id(comm_unique_id))
        h.set_params(p)
        h.env_reset()
        h.rollout()
        h.gae()
        B_local = n_local * T
        rng = np.random.default_rng(100 + rank)
        perms = np.stack([rng.permutation(B_local) for _ in range(epochs)]).astype(np.int32)
        fields = {name: h.read_field(getattr(_abi, "CRL_F_" + name)) for name in
                  ("STATE", "ACTION", "LOGPROB", "REWARD", "TERMINAL", "VALUE", "ADVANTAGE", "RETURN")}
        lr = 2.0e-4
        stats = h.update_epochs(perms, lr)
        params_after = h.get_params()
        gathered = [None] * world
        dist.all_gather_object(gathered, {"fields": fields, "perms": perms, "params": params_after, "stats": stats})
        if rank == 0:
            # single-process oracle on the union: same rollout (global Philox ids), union minibatches
            cfg_g = _abi.make_config(env_kind=kind, num_envs=N, num_steps=T, num_minibatches=mb, update_epochs=epochs, seed=seed)
            o = olib.create(cfg_g)
            o.set_params(p)
            o.env_reset()
            o.rollout()
            o.gae()
            ok_rollout = True
            for name in fields:
                cat = np.concatenate([g["fields"][name] for g in gathered], axis=1)
                full = o.read_field(getattr(_abi, "CRL_F_" + name))
                if name in ("TERMINAL",) or (name == "ACTION" and kind == 0):
                    ok_rollout = ok_rollout and np.array_equal(cat, full)
                else:
                    ok_rollout = ok_rollout and np.allclose(cat, full, rtol=1e-4, atol=2e-5)
                o.write_field(getattr(_abi, "CRL_F_" + name), cat)  # identical inputs for the update comparison
            M_local = B_local // mb
            st_o = []
            for e in range(epochs):
                for k in range(mb):
                    idx = []
                    for r, g in enumerate(gathered):
                        loc = g["perms"][e, k * M_local:(k + 1) * M_local]
                        n_l, t_l = loc % n_local, loc // n_local
                        idx.append((r * n_local + n_l) + N * t_l)
                    s = o.update_minibatch(np.concatenate(idx).astype(np.int32), lr)
                    st_o.append([s.loss, s.pg_loss, s.v_loss, s.entropy_loss])
            p_o = o.get_params()
            res["kind%d" % kind] = {
                "rollout_ok": bool(ok_rollout),
                "ranks_agree": bool(all(np.array_equal(g["params"], gathered[0]["params"]) for g in gathered)),
                "param_maxerr": float(np.max(np.abs(gathered[0]["params"] - p_o) / (np.abs(p_o) + 1e-2))),
                # pg_loss is a mean of O(1) terms that cancels to ~1e-4 .. 1e-7: the floor is 1e-2 of a summand, not of the sum
                "stats_maxerr": float(np.max(np.abs(np.array(st_o) - gathered[0]["stats"]) / (np.abs(np.array(st_o)) + 1e-2))),
            }
        h.close()
        # throughput path (CUDA graph, speculative loss_grad + ONE allreduce per minibatch) against the exact
        # statistics-exchange path, from identical starting states. scenario "fail": gamma = lambda = 0 and a critic
        # bias of 1.5 make s = mean(v - R^2) win the max (Q5), so the on-device verification fails and the host
        # replays the update exactly: the results must then be BIT-identical to the exact path.
        for scenario in scenarios:
            kw = dict(gamma=0.0, gae_lambda=0.0) if scenario == "fail" else {}
            p2 = p.copy()
            if scenario == "fail":
                p2[olib.param_layout(kind)[0][11]] = 1.5
            outs = {}
            for mode in ("spec", "exact"):
                cfg2 = _abi.make_config(env_kind=kind, num_envs=n_local, num_steps=T, num_minibatches=mb, update_epochs=epochs,
                                        seed=seed, device=local_rank, world_size=world, rank=rank, env_id_base=base, **kw)
                h2 = PPOHandle(cfg2)
                h2.comm_init(parallel.exchange_unique_id(comm_unique_id))
                h2.set_params(p2)
                h2.env_reset()
                if mode == "exact":
                    os.environ["CRL_MULTI_EXACT"] = "1"
                for u in range(3):
                    h2.train_update(lr)
                    if u >= 1:
                        h2.fetch_update(lag=1)
                st2, agg2 = h2.fetch_update()
                os.environ.pop("CRL_MULTI_EXACT", None)
                outs[mode] = (h2.get_params(), st2, h2.spec_replays())
                h2.close()
            g2 = [None] * world
            dist.all_gather_object(g2, outs["spec"][0])
            if rank == 0:
                ps, pe = outs["spec"][0], outs["exact"][0]
                res["kind%d" % kind][scenario] = {
                    "ranks_agree": bool(all(np.array_equal(x, g2[0]) for x in g2)),
                    "finite": bool(np.all(np.isfinite(outs["spec"][1]))),
                    "replays": int(outs["spec"][2]), "replays_exact_mode": int(outs["exact"][2]),
                    "bit_identical": bool(np.array_equal(ps, pe)),
                    "maxerr": float(np.max(np.abs(ps - pe) / (np.abs(pe) + 1e-2))),
                    "stats_maxerr": float(np.max(np.abs(outs["spec"][1] - outs["exact"][1]) / (np.abs(outs["exact"][1]) + 1e-2))),
                }
    return res if rank == 0 else None


def verdict(res):
    """the assertions of tests/test_gpu_multi.py as a list of failures (empty = parity holds)"""
    bad = []

    def need(cond, what):
        if not cond:
            bad.append(what)
    for kind, r in res.items():
        need(r["rollout_ok"], kind + ": union of the shards' rollouts != single-process oracle rollout")
        need(r["ranks_agree"], kind + ": replicated parameters differ across ranks")
        need(r["param_maxerr"] < 2e-5, kind + ": post-step parameters vs oracle on the union minibatches: %g" % r["param_maxerr"])
        need(r["stats_maxerr"] < 1e-4, kind + ": loss statistics vs oracle: %g" % r["stats_maxerr"])
        n = r.get("normal")
        if n is not None:
            need(n["ranks_agree"] and n["finite"] and n["replays"] == 0, kind + ": speculative path: %r" % (n,))
            need(n["maxerr"] < 2e-5 and n["stats_maxerr"] < 1e-4, kind + ": speculative != exact: %r" % (n,))
        f = r.get("fail")
        if f is not None:
            need(f["ranks_agree"] and f["finite"] and f["replays_exact_mode"] == 0, kind + ": forced failure: %r" % (f,))
            if kind == "kind0":
                # CartPole with gamma = 0: R in {0, 1, v} so s = mean(v - R^2) > min (clip - R)^2 = 0: verification fails
                need(f["replays"] >= 1, kind + ": the forced speculation failure was not detected")
            if f["replays"] >= 1:
                need(f["bit_identical"], kind + ": a failed speculation was not replayed exactly")
            else:
                need(f["maxerr"] < 2e-5, kind + ": speculative != exact (fail scenario): %g" % f["maxerr"])
    return bad
