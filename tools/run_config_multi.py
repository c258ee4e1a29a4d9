"""Multi-GPU probe (torchrun, one rank per GPU): a BASELINE configuration with the SNPs sharded over the ranks (STRONG
scaling: the total SNP count is fixed), full-mode smartpca core through the collective eb_pca_full (GRM passes with the
fused finalize+reduce over peer memory, eigensolver replicated).  Rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/run_config_multi.py C4
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eig_b200 import capi, parallel, synth  # noqa: E402

CONFIGS = {"C2": (5000, 600000, 0.0), "C3": (20000, 1200000, 0.30), "C4": (50000, 600000, 0.0), "C5": (200000, 500000, 0.0)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    N, M, miss = CONFIGS[name]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(local)
    if world > 1:
        ctx.set_comm(parallel.TorchComm(device=dev))
    s0, s1 = parallel.shard_snps(M, rank, world)
    rl = synth.rlen_for(N)
    slab = torch.empty((s1 - s0, rl), dtype=torch.uint8, device=dev)
    ctx.synth_packed_device(slab.data_ptr(), s1 - s0, rl, N, seed=1, s0=s0, missing=miss, npops=4, delta=0.05)
    ctx.adopt_packed_device(slab.data_ptr(), s1 - s0, rl, N)
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if name == "C5":                      # fastmode: kjg_fpca K=10, L=20, I=10 on the shards (all-reduced sketch, TSQR)
        ctx.set_rows(None)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ev, vec = ctx.fpca(10, 20, 10, seed=123)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter() - t0
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            flops = (4.0 * 10 * 20 + 2 * 20 + 2 * 11 * 20) * M * N
            print(json.dumps(dict(config=name, n_gpus=world, N=N, M=M, scaling="strong", fpca_s=float(tt[0]), fpca_tflops=flops / float(tt[0]) / 1e12,
                                  eval_top=ev[:4].tolist(), unit_err=float(np.abs((vec * vec).sum(0) - 1).max()))), flush=True)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    t0 = time.perf_counter()
    res = ctx.pca_full(numeigs=10, numoutliter=5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter() - t0
    tm = ctx.timings()
    extra = {}
    if "coords" in sys.argv[2:]:          # lsqproject: the .evec coordinate sequence smartpca.c:1440-1564 on the shards
        t1 = time.perf_counter(); co, es, ok = ctx.evec_coords(res["evecs"]); torch.cuda.synchronize()
        extra["evec_coords_s"] = time.perf_counter() - t1; extra["coords_ok"] = bool(ok.all())
    if "shrink" in sys.argv[2:]:          # shrinkmode: doshrinkp on the shards
        t1 = time.perf_counter(); sc, sl, sok = ctx.shrink_coords(10, newshrink=False); torch.cuda.synchronize()
        extra["shrink_s"] = time.perf_counter() - t1; extra["shrink_ok"] = bool(sok.all())
        extra["shrink_unit_err"] = float(np.abs((sc * sc).sum(1) - 1).max()); extra["shrink_lam"] = sl[:4].tolist()
    if world > 1:
        dist.barrier()
    extra["total_s"] = time.perf_counter() - t0
    tt = torch.tensor([t, res["secs_grm"], res["secs_eig"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        lam = res["lambda_"]
        print(json.dumps(dict(config=name, n_gpus=world, N=N, M=M, missing=miss, scaling="strong", core_s=float(tt[0]), grm_s=float(tt[1]), eig_s=float(tt[2]),
                              passes=res["niter"], removed=len(res["removed_index"]), nused=int(res["nused"]),
                              snp_indiv2_per_s=float(N) * N * res["nused"] * res["niter"] / float(tt[1]),
                              grm_kernel_ms=tm["grm_ms"], finalize_reduce_ms=tm["finalize_ms"], tridiag_ms=tm["tridiag_ms"], bisect_ms=tm["bisect_ms"],
                              vectors_ms=tm["vectors_ms"], lam_top=lam[:4].tolist(), lam_sum=float(lam.sum()), lam_min=float(lam.min()), **extra)), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
