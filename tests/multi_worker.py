"""Worker of the multi-rank GPU tests: one process per rank (spawned by tests/test_gpu_multi.py, or by torchrun on a
multi-GPU box).  Every rank owns one SNP shard; the exchange steps run as the library's kernels over peer memory
(CUDA IPC).  With fewer GPUs than ranks all ranks share cuda:0 (gloo plumbing) -- the peer path is the same code."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eig_b200 import capi, parallel, synth  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    backend = os.environ.get("EB_BACKEND", "nccl" if ngpu >= world else "gloo")
    dev_index = rank if ngpu >= world else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    comm = parallel.TorchComm(device=dev)

    nsnp, nind = 3000, 333
    P = synth.packed_genotypes(11, nsnp, nind, missing=0.1, npops=3, delta=0.25)
    P[5] = 0xFF                                   # an all-missing SNP (dropped by the rule) in rank 0's shard
    s0, s1 = parallel.shard_snps(nsnp, rank, world)

    single = capi.Context(dev_index)              # the whole matrix on one GPU: the expected result
    single.upload_packed(P, nind); single.set_rows(None)
    ref = single.grm(want_xtx=True)

    ctx = capi.Context(dev_index)
    ctx.set_comm(comm)
    if os.environ.get("EB_TEST_GRM_METHOD"):      # 2: the sharded passes take the integer tensor-core GRM (the single-GPU expectation stays FP64 DMMA)
        ctx.set_option("grm_method", int(os.environ["EB_TEST_GRM_METHOD"]))
    # 1. the all-reduce kernel alone
    v = np.arange(1000, dtype=np.float64) * (rank + 1)
    out = ctx.peer_allreduce_test(v)
    assert np.array_equal(out, np.arange(1000, dtype=np.float64) * (world * (world + 1) // 2)), "peer all-reduce"
    # 2. sharded GRM == single-GPU GRM; per-SNP outputs are the shard's slice; nused is the total
    ctx.upload_packed(P[s0:s1], nind); ctx.set_rows(None)
    for rep in range(2):                          # second pass reuses the mapped peer buffers
        r = ctx.grm(want_xtx=True)
        assert abs(r["y"] - ref["y"]) <= 1e-13 * ref["y"], (r["y"], ref["y"])
        assert np.abs(r["XTX"] - ref["XTX"]).max() <= 1e-12 * np.abs(ref["XTX"]).max()
        assert r["nused"] == ref["nused"], (r["nused"], ref["nused"])
        for k in ("c0", "c1", "nmiss", "used", "xmean", "xfancy"):
            assert np.array_equal(r[k], ref[k][s0:s1]), k
    # bit-identical on every rank
    h = torch.from_numpy(r["XTX"].view(np.int64).reshape(-1).copy())
    hs = [torch.empty_like(h) for _ in range(world)]
    if backend == "nccl":
        hd = h.to(dev); hsd = [torch.empty_like(hd) for _ in range(world)]
        dist.all_gather(hsd, hd); hs = [t.cpu() for t in hsd]
    else:
        dist.all_gather(hs, h)
    assert all(torch.equal(hs[0], t) for t in hs), "GRM differs between ranks"
    lam, vec = ctx.eig(4)
    rl, rv = single.eig(4)
    assert np.abs(lam - rl).max() <= 1e-9 * rl[0]
    for i in range(3):
        assert abs(abs(vec[i] @ rv[i]) - 1) < 1e-9
    # 3. rows subset + outlier loop: same decisions and spectrum as the single-GPU run
    pd = np.array([0.05] * 3 + [1.5])             # two planted outliers (tests/test_gpu_eig.py::test_pca_full_outlier_loop)
    g = synth.genotypes(4, nsnp, nind, missing=0.02, npops=4, pop_delta=pd)
    g0 = synth.genotypes(4, nsnp, nind, missing=0.02, npops=4, pop_delta=np.array([0.05] * 4))
    sel = (synth.pop_of(nind, 4) == 3) & (np.arange(nind) < nind - 2)
    g[:, sel] = g0[:, sel]
    Q = synth.pack(g)
    xi = np.arange(5, nind, dtype=np.int32)
    single.upload_packed(Q, nind)
    a = single.pca_full(xindex=xi, numeigs=4, numoutliter=3)
    ctx.upload_packed(Q[s0:s1], nind)
    b = ctx.pca_full(xindex=xi, numeigs=4, numoutliter=3)
    assert np.array_equal(a["removed_index"], b["removed_index"]) and a["niter"] == b["niter"] and a["nused"] == b["nused"]
    assert len(a["removed_index"]) >= 1, "test needs at least one outlier"
    assert np.abs(a["lambda_"] - b["lambda_"]).max() <= 1e-9 * a["lambda_"][0]
    assert np.array_equal(b["used"], a["used"][s0:s1])
    # 3b. .evec coordinates (loadings, projections, lsqproj) and shrinkmode on SNP shards == single GPU
    xi2 = np.arange(3, nind - 4, dtype=np.int32)
    single.upload_packed(P, nind); single.set_rows(xi2); single.grm(want_snp=False)
    ctx.upload_packed(P[s0:s1], nind); ctx.set_rows(xi2); ctx.grm(want_snp=False)
    l0, v0_ = single.eig(4); l1, v1_ = ctx.eig(4)
    assert np.array_equal(v0_, v1_) or np.abs(np.abs(np.einsum("ij,ij->i", v0_, v1_)) - 1).max() < 1e-12
    c0, e0, k0 = single.evec_coords(v0_)
    c1, e1, k1 = ctx.evec_coords(v0_)
    assert np.array_equal(k0, k1) and np.abs(e0 - e1).max() <= 1e-10 * np.abs(e0).max()
    assert np.abs(c0 - c1).max() <= 1e-10 * np.abs(c0).max(), np.abs(c0 - c1).max()
    for new in (False, True):
        a0, la0, ok0 = single.shrink_coords(3, newshrink=new)
        a1, la1, ok1 = ctx.shrink_coords(3, newshrink=new)
        assert ok0.all() and ok1.all() and np.abs(la0 - la1).max() <= 1e-12 * la0[0]
        sg = np.sign((a0 * a1).sum(1))
        assert np.abs(a0 - a1 * sg[:, None]).max() < 1e-8, (new, np.abs(a0 - a1 * sg[:, None]).max())
    # 4. sharded fastmode == single-GPU fastmode
    single.upload_packed(P, nind); single.set_rows(None)
    ctx.upload_packed(P[s0:s1], nind); ctx.set_rows(None)
    for rep in range(2):
        ev0, vec0 = single.fpca(4, 8, 3, seed=77)
        ev1, vec1 = ctx.fpca(4, 8, 3, seed=77)
        assert np.abs(ev1 - ev0).max() <= 1e-9 * ev0[0], (ev0, ev1)
        for k in range(4):
            assert abs(abs(vec0[:, k] @ vec1[:, k]) - 1) < 1e-8, k
    # 5. fastmode at smartpca's defaults (K = 10, L = 20, I = 10) on a spectrum whose ten leading eigenvalues are separated:
    #    SNP shards (all-reduced sketch, TSQR) == single GPU to the north_star bar
    g12 = synth.genotypes(5, 4000, 400, missing=0.02, npops=12, pop_delta=np.linspace(0.15, 0.5, 12))
    P12 = synth.pack(g12)
    t0, t1 = parallel.shard_snps(4000, rank, world)
    single.upload_packed(P12, 400); single.set_rows(None)
    ctx.upload_packed(P12[t0:t1], 400); ctx.set_rows(None)
    ev0, vec0 = single.fpca(10, 20, 10, seed=7)
    ev1, vec1 = ctx.fpca(10, 20, 10, seed=7)
    assert (np.abs(ev1 - ev0) / ev0).max() <= 1e-9, (ev0, ev1)
    assert np.abs(np.abs((vec0 * vec1).sum(0)) - 1).max() <= 1e-9
    # 6. collective eigensolver: n >= 4096 on a sharded GRM -> the subspace iteration splits its block mat-vecs by row tiles over the
    #    ranks (stream-ordered all-reduce of the 64 x n block); same spectrum / vectors as the single-GPU solve, identical on all ranks
    nb, mb = 4200, 3000
    Pb = synth.pack(synth.genotypes(9, mb, nb, missing=0.01, npops=6, pop_delta=np.linspace(0.2, 0.5, 6)))
    u0, u1 = parallel.shard_snps(mb, rank, world)
    single.upload_packed(Pb, nb); single.set_rows(None); single.grm(want_snp=False)
    ctx.upload_packed(Pb[u0:u1], nb); ctx.set_rows(None); ctx.grm(want_snp=False)
    la, va = single.eig(5)
    ctx.set_option("dist_min", 2048)          # default 8192: force the row-distributed band reduction / subspace iteration at this n
    ctx.set_option("eig_vectors", 1)          # leading vectors by the (row-distributed) subspace iteration
    lb, vb = ctx.eig(5)
    tm = ctx.timings()
    assert tm["eig_method"] == 2 and tm["chfsi_matvecs"] > 0
    assert np.abs(la - lb).max() <= 1e-9 * la[0]
    assert np.abs(np.abs(np.einsum("ij,ij->i", va, vb)) - 1).max() <= 1e-9
    ctx.set_option("eig_vectors", 0)          # default: back-transformation through the reflectors of the row-distributed band reduction
    lc, vc = ctx.eig(5)
    tm = ctx.timings()
    assert tm["eig_method"] == 2 and tm["chfsi_matvecs"] == 0
    assert np.abs(la - lc).max() <= 1e-9 * la[0]
    assert np.abs(np.abs(np.einsum("ij,ij->i", va, vc)) - 1).max() <= 1e-9
    vb = vc
    hv = torch.from_numpy(vb.view(np.int64).reshape(-1).copy())
    if backend == "nccl":
        hvd = hv.to(dev); hall = [torch.empty_like(hvd) for _ in range(world)]
        dist.all_gather(hall, hvd); hall = [t.cpu() for t in hall]
    else:
        hall = [torch.empty_like(hv) for _ in range(world)]
        dist.all_gather(hall, hv)
    assert all(torch.equal(hall[0], t) for t in hall), "eigenvectors differ between ranks"
    dist.barrier()
    ctx.close(); single.close()
    dist.destroy_process_group()
    print("rank %d/%d ok (backend %s, device %d)" % (rank, world, backend, dev_index))


if __name__ == "__main__":
    main()
