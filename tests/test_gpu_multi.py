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


def _run(world, backend=None, timeout=600, extra_env=None):
    port = 29600 + (os.getpid() % 2000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        if backend:
            env["EB_BACKEND"] = backend
        env.update(extra_env or {})
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


def test_two_ranks_sharded_integer_grm():
    """the same worker with every sharded GRM pass on the exact integer tensor-core path (grm_i8.cu): tiles pushed to their owners after
    the integer GEMM, compared with the single-GPU FP64 DMMA result"""
    _run(2, extra_env={"EB_TEST_GRM_METHOD": "2"})


def test_three_ranks_shared_gpu_gloo():
    """odd world size, ranks sharing one GPU: uneven shards and the same-device IPC path"""
    import torch
    if torch.cuda.device_count() >= 3:
        _run(3)
    else:
        _run(3, backend="gloo")


def test_in_process_communicator_threads():
    """eb_local_comm: two host threads of THIS process, one context each (both on cuda:0 when the box has one GPU, else one
    GPU each) -- the C-caller route to several GPUs; same-process ranks use each other's device pointers directly."""
    import threading
    import numpy as np
    import torch
    from eig_b200 import capi, parallel, synth
    world = 2
    ngpu = torch.cuda.device_count()
    nsnp, nind = 2500, 300
    P = synth.packed_genotypes(13, nsnp, nind, missing=0.1, npops=3, delta=0.25)
    single = capi.Context(0)
    single.upload_packed(P, nind); single.set_rows(None)
    ref = single.grm(want_xtx=True)
    ev0, vec0 = single.fpca(3, 6, 2, seed=9)
    lc = capi.LocalComm(world)
    out = [None] * world; err = [None] * world

    def work(rank):
        try:
            ctx = capi.Context(rank if ngpu >= world else 0)
            ctx.set_comm_struct(lc.rank_comm(rank), keep=lc)
            s0, s1 = parallel.shard_snps(nsnp, rank, world)
            ctx.upload_packed(P[s0:s1], nind); ctx.set_rows(None)
            r = ctx.grm(want_xtx=True)
            ev, vec = ctx.fpca(3, 6, 2, seed=9)
            res = ctx.pca_full(numeigs=3, numoutliter=2)
            out[rank] = (r, ev, vec, res, (s0, s1))
            ctx.set_comm(None); ctx.close()
        except Exception as ex:  # noqa: BLE001
            err[rank] = ex

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert all(not t.is_alive() for t in th), "a rank is stuck in a collective"
    assert err == [None] * world, err
    for rank in range(world):
        r, ev, vec, res, (s0, s1) = out[rank]
        assert np.abs(r["XTX"] - ref["XTX"]).max() <= 1e-12 * np.abs(ref["XTX"]).max() and r["nused"] == ref["nused"]
        assert np.array_equal(r["used"], ref["used"][s0:s1])
        assert np.abs(ev - ev0).max() <= 1e-9 * ev0[0]
    assert np.array_equal(out[0][0]["XTX"], out[1][0]["XTX"]), "reduced GRM must be bit-identical on both ranks"
    assert np.array_equal(out[0][3]["lambda_"], out[1][3]["lambda_"])
    lc.close(); single.close()


def test_c_shim_runs(tmp_path):
    """examples/smartpca_shim.c (plain C99 against include/eigb200.h): the single-context entry and the thread-per-GPU entry
    give the spectrum of the Python-driven run"""
    import ctypes
    import shutil
    import subprocess
    import numpy as np
    from eig_b200 import capi, synth
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    capi.lib()
    out = str(tmp_path / "libshim.so")
    cmd = [gcc, "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "smartpca_shim.c"), "-o", out, "-L", os.path.join(ROOT, "eig_b200"), "-leigb200", "-lpthread",
           "-Wl,-rpath," + os.path.join(ROOT, "eig_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = ctypes.CDLL(out)
    nsnp, nind, k = 3000, 200, 3
    P = synth.packed_genotypes(8, nsnp, nind, missing=0.05, npops=3, delta=0.3)
    ctx = capi.Context(0); ctx.upload_packed(P, nind)
    want = ctx.pca_full(numeigs=k, numoutliter=5)
    co_want, _, _ = ctx.evec_coords(want["evecs"])
    ctx.close()
    vp = ctypes.c_void_p
    for entry in ("eb_shim_full_mode", "eb_shim_threads"):
        xi = np.arange(nind, dtype=np.int32); lam = np.zeros(nind); vec = np.zeros((k, nind)); co = np.zeros((k, nind))
        if entry == "eb_shim_full_mode":
            rc = L.eb_shim_full_mode(P.ctypes.data_as(vp), ctypes.c_int64(nsnp), ctypes.c_int64(P.shape[1]), ctypes.c_int(nind), xi.ctypes.data_as(vp),
                                     ctypes.c_int(nind), ctypes.c_int(k), ctypes.c_int(5), lam.ctypes.data_as(vp), vec.ctypes.data_as(vp), co.ctypes.data_as(vp))
        else:
            rc = L.eb_shim_threads(P.ctypes.data_as(vp), ctypes.c_int64(nsnp), ctypes.c_int64(P.shape[1]), ctypes.c_int(nind), xi.ctypes.data_as(vp),
                                   ctypes.c_int(nind), ctypes.c_int(k), lam.ctypes.data_as(vp), vec.ctypes.data_as(vp))
        assert rc == 0
        n = len(want["lambda_"])
        assert np.abs(lam[:n] - want["lambda_"]).max() <= 1e-12 * want["lambda_"][0], entry
        got_vec = vec.reshape(-1)[:k * n].reshape(k, n)          # eb_pca_full packs [numeigs][nrows_final]
        assert np.abs(np.abs(np.einsum("ij,ij->i", got_vec, want["evecs"])) - 1).max() < 1e-9
        if entry == "eb_shim_full_mode" and n == nind:
            assert np.abs(co - co_want).max() <= 1e-9 * np.abs(co_want).max()
