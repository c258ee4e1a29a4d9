"""Multi-rank GPU tests of the SNP-sharded path: 2 ranks, one SNP shard each, exchange over peer memory.
On a box with >= 2 GPUs the ranks use one GPU each and NCCL for the host plumbing; on a single-GPU box both ranks share
cuda:0 with gloo plumbing -- the library code under test (CUDA IPC mapping, grm_peer_finalize_kernel,
peer_allreduce_kernel, TSQR) is the same."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(world, backend=None, timeout=600):
    port = 29600 + (os.getpid() % 2000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        if backend:
            env["EB_BACKEND"] = backend
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=timeout)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-4000:])
        assert "ok" in o


def test_two_ranks_sharded_grm_pca_fpca():
    _run(2)


def test_three_ranks_shared_gpu_gloo():
    """odd world size, ranks sharing one GPU: uneven shards and the same-device IPC path"""
    import torch
    if torch.cuda.device_count() >= 3:
        _run(3)
    else:
        _run(3, backend="gloo")
