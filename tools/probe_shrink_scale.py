"""GPU probe: full eigenbasis and shrinkmode at scale (timing + self-consistency)."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
from eig_b200 import capi, synth
c = capi.Context(0)
for spec in sys.argv[1:]:
    parts = spec.split("x")
    N, M, k = (int(v) for v in parts[:3]); variant = parts[3] if len(parts) > 3 else "both"
    rl = synth.rlen_for(N)
    buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
    c.synth_packed_device(buf.data_ptr(), M, rl, N, seed=2, missing=0.1, npops=4, delta=0.08)
    c.adopt_packed_device(buf.data_ptr(), M, rl, N); c.set_rows(None)
    t0 = time.time(); r = c.grm(want_snp=False); tg = time.time() - t0
    rec = dict(N=N, M=M, k=k, grm_s=tg)
    if k > 0:
        t0 = time.time(); co, lam, ok = c.shrink_coords(k, newshrink=(variant == "new")); rec["shrink_%s_s" % ("new" if variant == "new" else "old")] = time.time() - t0
        tm = c.timings(); rec.update(eig_tridiag_ms=tm["tridiag_ms"], eig_bisect_ms=tm["bisect_ms"], eig_vectors_ms=tm["vectors_ms"])
        rec["ok"] = bool(ok.all()); rec["lam"] = lam[:4].tolist(); rec["unit"] = float(np.abs((co * co).sum(1) - 1).max())
        rec["flops_blocks"] = 2.0 * k * N * N * M
        rec["block_tflops_lower_bound"] = rec["flops_blocks"] / (time.time() - t0) / 1e12
        if variant == "both":
            t0 = time.time(); co2, lam2, ok2 = c.shrink_coords(k, newshrink=True); rec["shrink_new_s"] = time.time() - t0
            rec["old_vs_new_maxdiff_per_vec"] = np.abs(co - co2).max(axis=1).tolist()
    else:
        c.set_option("eig_method", 1)
        A = None
        t0 = time.time(); lam, vec = c.eig(N, want_lambda=True); rec["full_basis_s"] = time.time() - t0
        c.set_option("eig_method", 0)
        tm = c.timings(); rec.update(eig_tridiag_ms=tm["tridiag_ms"], eig_bisect_ms=tm["bisect_ms"], eig_vectors_ms=tm["vectors_ms"])
        sub = vec[:: max(1, N // 512)]
        rec["orth_sample"] = float(np.abs(sub @ sub.T - np.eye(len(sub))).max()); rec["sum_lam"] = float(lam.sum())
    print(json.dumps(rec), flush=True)
    del buf
