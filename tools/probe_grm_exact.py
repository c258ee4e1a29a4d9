"""GPU probe: GRM kernel at scale -- run-to-run bitwise repeatability and agreement with an independent FP64 product
(torch.matmul on the decoded matrix: a checker, not a product path)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from eig_b200 import capi, synth
c = capi.Context(0)
for spec in sys.argv[1:]:
    N, M = (int(v) for v in spec.split("x"))
    rl = synth.rlen_for(N)
    buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
    c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=3, missing=0.1, npops=3, delta=0.1)
    c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
    outs = []
    for rep in range(3):
        r = c.grm(want_xtx=True, want_snp=(rep == 0))
        outs.append(r["XTX"] * r["y"])
        if rep == 0:
            xm, xf, used = r["xmean"], r["xfancy"], r["used"]
    same = [bool(np.array_equal(outs[0], o)) for o in outs[1:]]
    # independent product: decode on the GPU with torch, chunked
    P = buf.cpu().numpy()
    ref = torch.zeros((N, N), dtype=torch.float64, device="cuda")
    xm_t = torch.from_numpy(xm).cuda(); xf_t = torch.from_numpy(xf).cuda(); us = torch.from_numpy(used.astype(np.float64)).cuda()
    for a in range(0, M, 20000):
        b = min(M, a + 20000)
        g = torch.from_numpy(synth.unpack(P[a:b], N).astype(np.float64)).cuda()      # [b-a][N], -1 missing
        x = torch.where(g < 0, torch.zeros_like(g), g * xf_t[a:b, None] - xm_t[a:b, None]) * us[a:b, None]
        ref += x.T @ x
    ref = ref.cpu().numpy()
    err = np.abs(outs[0] - ref).max() / np.abs(ref).max()
    nbad = int((np.abs(outs[0] - ref) > 1e-9 * np.abs(ref).max()).sum())
    print("N %d M %d nsplit %d: repeat-bitwise %s  max rel err vs torch %.3e  bad elems %d" % (N, M, c.timings()["nsplit"], same, err, nbad), flush=True)
    del buf
