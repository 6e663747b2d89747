"""Sharding of the env vector over GPUs. One process per GPU; torch.distributed is
used only as plumbing (rendezvous, unique-id broadcast, max-over-ranks timing)."""
import os

import numpy as np


def shard_envs(num_envs_global, world_size, rank):
    """GPU `rank` of `world_size` owns envs [base, base+n). Requires divisibility so that every
    shard has the same minibatch size (the loss is normalised by world_size*M)."""
    if num_envs_global % world_size != 0:
        raise ValueError("num_envs = %d is not divisible by the number of GPUs = %d" % (num_envs_global, world_size))
    n = num_envs_global // world_size
    return rank * n, n


def dist_info():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = defaults)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def broadcast_bytes(payload, src=0, group=None):
    """Broadcast a bytes object from rank `src` with torch.distributed (gloo or nccl)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    if dist.get_rank(group) == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
        n = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
    else:
        n = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(n, src, group=group)
    if dist.get_rank(group) != src:
        t = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src, group=group)
    return bytes(t.cpu().numpy().tolist())


def exchange_unique_id(make_id, group=None):
    """rank 0 calls make_id() (crl_comm_unique_id); everyone receives the 128 bytes."""
    import torch.distributed as dist
    payload = make_id() if dist.get_rank(group) == 0 else b""
    return broadcast_bytes(payload, 0, group)


def max_over_ranks(value, group=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(values, group=None):
    import torch
    import torch.distributed as dist
    arr = np.asarray(values, np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arr
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(arr, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
