"""End-to-end full-mode smartpca run through the C-ABI, file to file (process start -> exit, input in the page cache):
PACKEDANCESTRYMAP .geno + .snp + .ind  ->  eb_upload_packed_file -> eb_pca_full -> eb_evec_coords -> .eval + .evec.

  python tools/smartpca_e2e.py gen  DIR NIND NSNP [MISSING]      write a synthetic Hardy-Weinberg dataset (device generator)
  python tools/smartpca_e2e.py run  DIR [numoutevec] [numoutlieriter]   the timed pipeline; prints one JSON line
No torch import in `run`: numpy + ctypes + libeigb200 only."""
import json
import os
import sys
import time

T0 = time.perf_counter()
import numpy as np  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eig_b200 import capi, synth  # noqa: E402


def gen(d, nind, nsnp, missing):
    import torch
    os.makedirs(d, exist_ok=True)
    ctx = capi.Context(0)
    rl = synth.rlen_for(nind)
    ids = ["ind%d" % i for i in range(nind)]; snps = ["rs%d" % i for i in range(nsnp)]
    pops = ["Pop%d" % k for k in synth.pop_of(nind, 4)]
    hdr = ("GENO %7d %7d %x %x" % (nind, nsnp, capi.hash_ids(ids) & 0xffffffff, capi.hash_ids(snps) & 0xffffffff)).encode()
    with open(os.path.join(d, "data.geno"), "wb") as f:
        f.write(hdr + b"\0" * (rl - len(hdr)))
        step = max(1, (1 << 30) // rl)
        for s0 in range(0, nsnp, step):
            n = min(step, nsnp - s0)
            buf = torch.empty((n, rl), dtype=torch.uint8, device="cuda")
            ctx.synth_packed_device(buf.data_ptr(), n, rl, nind, seed=1, s0=s0, missing=missing, npops=4, delta=0.05); ctx.sync()
            f.write(buf.cpu().numpy().tobytes())
    with open(os.path.join(d, "data.snp"), "w") as f:
        per = max(1, (nsnp + 21) // 22)
        f.writelines("%20s %2d %12.6f %12d A C\n" % (snps[k], k // per + 1, ((k % per) * 1000 + 1000) * 1e-8, (k % per) * 1000 + 1000) for k in range(nsnp))
    with open(os.path.join(d, "data.ind"), "w") as f:
        f.writelines("%20s U %s\n" % (ids[k], pops[k]) for k in range(nind))
    print("written", d, nind, nsnp)


def run(d, numoutevec, numoutlieriter):
    t = {"import_s": time.perf_counter() - T0}
    t1 = time.perf_counter()
    ind = [l.split() for l in open(os.path.join(d, "data.ind"))]
    ids = [r[0] for r in ind]; groups = [r[2] for r in ind]
    snps = [l.split(None, 1)[0] for l in open(os.path.join(d, "data.snp"))]
    nind, nsnp = len(ids), len(snps)
    ih, sh = capi.hash_ids(ids), capi.hash_ids(snps)
    t["read_ind_snp_s"] = time.perf_counter() - t1; t1 = time.perf_counter()
    ctx = capi.Context(0)
    t["context_s"] = time.perf_counter() - t1; t1 = time.perf_counter()
    ctx.upload_packed_file(os.path.join(d, "data.geno"), nind, nsnp, ihash=ih, shash=sh)
    t["load_geno_s"] = time.perf_counter() - t1; t1 = time.perf_counter()
    res = ctx.pca_full(numeigs=numoutevec, numoutliter=numoutlieriter)
    t["pca_full_s"] = time.perf_counter() - t1; t["grm_s"] = res["secs_grm"]; t["eig_s"] = res["secs_eig"]; t1 = time.perf_counter()
    coords, es, ok = ctx.evec_coords(res["evecs"])
    t["evec_coords_s"] = time.perf_counter() - t1; t1 = time.perf_counter()
    capi.write_eval(os.path.join(d, "out.eval"), res["lambda_"])
    capi.write_evec(os.path.join(d, "out.evec"), res["lambda_"][:numoutevec], ids, groups, coords)
    t["write_s"] = time.perf_counter() - t1
    t["total_s"] = time.perf_counter() - T0
    print(json.dumps(dict(nind=nind, nsnp=nsnp, numoutevec=numoutevec, numoutlieriter=numoutlieriter, passes=res["niter"], removed=len(res["removed_index"]),
                          lam_top=res["lambda_"][:3].tolist(), geno_GB=os.path.getsize(os.path.join(d, "data.geno")) / 1e9, **t)))


if __name__ == "__main__":
    if sys.argv[1] == "gen":
        gen(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]) if len(sys.argv) > 5 else 0.0)
    else:
        run(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 10, int(sys.argv[4]) if len(sys.argv) > 4 else 5)
