"""worker for tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL); the checks live in multi_gpu_checks.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cleanrl_jl_b200 import parallel  # noqa: E402
from oracle.oracle import OracleLib  # noqa: E402
from multi_gpu_checks import run_checks  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, local_rank, world = parallel.dist_info()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = run_checks(dist, OracleLib())
    if rank == 0:
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
