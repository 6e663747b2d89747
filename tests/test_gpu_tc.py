"""The two implementations of the minibatch update -- tcgen05 3xTF32 (csrc/update_tc.cu) and FP32 FFMA
(csrc/update.cu) -- must agree at BASELINE.json's full size (4096 envs x 128 steps, minibatch 131,072): a
size-independent property that complements the small-size comparisons with the oracle in test_gpu_parity.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(tmp_path, kind, no_tc):
    out = str(tmp_path / ("r_%d_%d.npz" % (kind, no_tc)))
    env = dict(os.environ)
    env.pop("CRL_NO_TC", None)
    if no_tc:
        env["CRL_NO_TC"] = "1"
    subprocess.run([sys.executable, os.path.join(HERE, "tc_worker.py"), out, str(kind)], check=True, env=env, timeout=300)
    return np.load(out)


@pytest.mark.parametrize("kind", [0, 1])
def test_tcgen05_update_matches_ffma_update_at_full_size(tmp_path, torch_cuda, abi, olib, kind):
    tc, ff = _run(tmp_path, kind, 0), _run(tmp_path, kind, 1)
    # identical inputs: the rollout and GAE kernels are the same code in both runs
    for name in ("STATE", "ACTION", "ADVANTAGE", "RETURN"):
        np.testing.assert_array_equal(tc[name], ff[name], err_msg=name)
    # one minibatch of 131,072 samples: loss statistics and raw gradients
    # (pg_loss is a cancelling mean of O(1) terms: atol tied to the summands)
    np.testing.assert_allclose(tc["mb_stats"], ff["mb_stats"], rtol=1e-5, atol=5e-7)
    offs, sizes = olib.param_layout(kind)
    g1, g2 = tc["mb_grads"], ff["mb_grads"]
    for off, size in zip(offs, sizes):
        a, b = g1[off:off + size], g2[off:off + size]
        # 3xTF32: each 64-term (or 131,072-term) dot product is within ~1e-6 of the largest one in its array
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=3e-6 * np.abs(b).max(), err_msg="array at %d" % off)
    np.testing.assert_allclose(tc["mb_params"], ff["mb_params"], rtol=1e-5, atol=1e-6)
    # one full update (16 minibatches with Adam steps in between) from there
    np.testing.assert_allclose(tc["upd_stats"], ff["upd_stats"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(tc["upd_params"], ff["upd_params"], rtol=2e-5, atol=2e-6)
    assert tc["replays"][0] == 0 and ff["replays"][0] == 0
