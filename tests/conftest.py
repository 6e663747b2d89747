import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def olib():
    """the CPU oracle (test infrastructure)"""
    from oracle.oracle import OracleLib
    return OracleLib()


@pytest.fixture(scope="session")
def abi():
    from cleanrl_jl_b200 import _abi
    return _abi


@pytest.fixture(scope="session")
def crl():
    """the CUDA library binding; GPU tests must go through the C ABI"""
    from cleanrl_jl_b200 import _lib
    _lib.load()
    return _lib


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    torch.cuda.set_device(0)
    return torch


def rand_params(olib, env_kind, seed=0, scale=1.0):
    """random-but-reasonable parameters: orthogonal init like networks.jl, plus small biases"""
    from cleanrl_jl_b200 import networks
    d = olib.dims(env_kind)
    rng = np.random.default_rng(seed)
    p = networks.init_params(env_kind == 1, d["D"], d["A"], seed=seed)
    p = p + (0.05 * scale * rng.standard_normal(p.shape)).astype(np.float32)
    # make the actor head non-trivial so that action probabilities are not all 0.5
    off, size = olib.param_layout(env_kind)
    p[off[4]:off[4] + size[4]] *= 30.0
    return p.astype(np.float32)
