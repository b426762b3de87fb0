"""
Multi-GPU parity (``-m gpu``, needs >= 2 GPUs; skipped otherwise): the conformation-sharded objective over NCCL equals the
single-GPU objective on the same conformations to 1e-12 relative, for every weighting method, for f and f_batch.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world):
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(REPO, "tests", "dist_gpu_worker.py")]
    res = subprocess.run(cmd, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    out = {w: np.load(os.path.join(REPO, "gpurun_out", "dist_values_%s_w%d.npy" % (w, world)))
           for w in ("uniform", "boltzmann", "non_boltzmann")}
    grads = {w: np.load(os.path.join(REPO, "gpurun_out", "dist_grad_%s_w%d.npy" % (w, world))) for w in ("uniform", "boltzmann")}
    extra = {k: np.load(os.path.join(REPO, "gpurun_out", "dist_%s_w%d.npy" % (k, world)))
             for k in ("ebatch_shard", "ebatch_full", "lls_shard", "lls_full")}
    return out, grads, extra


def test_sharded_objective_equals_single_gpu():
    import torch
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs")
    single, single_g, _ = _run(1)
    multi, multi_g, extra = _run(2)
    # energy-only batch through the contraction kernel and the LLS fit: 2 ranks on shards == rank 0 alone on everything
    assert np.all(np.abs(extra["ebatch_shard"] - extra["ebatch_full"]) <= 1e-10 * np.abs(extra["ebatch_full"]))
    assert np.abs(extra["lls_shard"] - extra["lls_full"]).max() <= 1e-8 * np.abs(extra["lls_full"]).max()
    for w in single_g:   # analytic gradient: the shards' term gradients are summed by one allreduce
        assert np.abs(single_g[w] - multi_g[w]).max() <= 1e-10 * np.abs(single_g[w]).max(), w
    for w in single:
        assert np.all(np.abs(single[w] - multi[w]) <= 1e-12 * np.abs(single[w])), (w, single[w], multi[w])
        # f and f_batch agree with each other
        assert np.all(np.abs(single[w][:3] - single[w][3:]) <= 1e-12 * np.abs(single[w][:3]))
