"""Integer tensor-core GRM (grm_i8.cu) against the FP64 DMMA kernel on one GPU: time, 8-bit tensor throughput, agreement.
usage: probe_grm_i8.py [nind] [nsnp] [missing] [slices] [check]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from eig_b200 import capi, synth  # noqa: E402

nind = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
nsnp = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
missing = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
slices = int(sys.argv[4]) if len(sys.argv) > 4 else 0
check = int(sys.argv[5]) if len(sys.argv) > 5 else 1
slabs = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [0]
syncs = [int(x) for x in sys.argv[7].split(",")] if len(sys.argv) > 7 else [0]
pair = int(sys.argv[8]) if len(sys.argv) > 8 else 1
ctx = capi.Context(0)
ctx.set_option("i8_pair", pair)
rl = synth.rlen_for(nind)
slab = torch.empty((nsnp, rl), dtype=torch.uint8, device="cuda")
ctx.synth_packed_device(slab.data_ptr(), nsnp, rl, nind, seed=1, s0=0, missing=missing)
ctx.adopt_packed_device(slab.data_ptr(), nsnp, rl, nind)
ctx.set_rows(None)
ctx.set_option("i8_slices", slices)
ref = None
for method, slab, sync in ([(1, 0, 0)] if check else []) + [(2, s, y) for s in slabs for y in syncs]:
    ctx.set_option("grm_method", method)
    ctx.set_option("i8_slab", slab)
    ctx.set_option("i8_sync", sync)
    if method == 2:
        print("slab cap %d SNPs, sync lag %d, pair %d" % (slab, sync, pair))
    for rep in range(2):
        t0 = time.perf_counter()
        r = ctx.grm(want_snp=False)
        tm = ctx.timings()
        line = "method %d pass %d: %.1f ms wall, grm %.1f ms, %.2f TFLOP/s FP64-equivalent" % (
            method, rep, (time.perf_counter() - t0) * 1e3, tm["grm_ms"], nind * (nind + 1.0) * r["nused"] / tm["grm_ms"] / 1e9)
        if method == 2:
            line += ", %d digits x %d bases, %d flagged blocks, %.1f TOP/s 8-bit" % (
                tm["i8_slices"], tm["i8_segments"], tm["i8_flag_blocks"], tm["i8_tera_ops"] / (tm["grm_ms"] * 1e-3))
        print(line, flush=True)
    ptr, ld, n = ctx.grm_device_ptr() if hasattr(ctx, "grm_device_ptr") else (None, None, None)
    if check and ptr:
        X = torch.empty(0)
        import ctypes
        sz = ld * ld
        buf = torch.empty(sz, dtype=torch.float64, device="cuda")
        ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(sz * 8), 3)
        m = buf.view(ld, ld)[:n, :n]
        if ref is None:
            ref = m.clone()
        else:
            print("max |i8 - dmma| / max |dmma| = %.3e; symmetric: %s" % (
                float((m - ref).abs().max() / ref.abs().max()), bool(torch.equal(m, m.T))), flush=True)
