"""GPU probe (run under gpurun): the BASELINE.json configurations C2..C5 on ONE B200, stage by stage.

usage: python tools/measure_configs.py [C2] [C3] [C4] [C5] [--out gpurun_out/configs.json]
Every stage is timed with the library's CUDA events where it has them (eb_get_timings) and by wall clock around the
synchronous C-ABI call otherwise; rates are ALGORITHMIC work / time:
  stats / gather : N*M/4 packed bytes read per pass                         -> GB/s  (HBM bound)
  grm            : N(N+1) M_used flops                                      -> TFLOP/s (FP64 tensor bound)
  fastmode       : (4 I L + 2 L + 2 (I+1) L) M N flops, (I+2) N M / 4 bytes  -> TFLOP/s and GB/s
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from eig_b200 import capi, synth  # noqa: E402

CONFIGS = {
    "C2": dict(N=5000, M=600000, missing=0.0, mode="full"),
    "C3": dict(N=20000, M=1200000, missing=0.30, mode="full", lsq=True),
    "C4": dict(N=50000, M=600000, missing=0.0, mode="full"),
    "C5": dict(N=200000, M=500000, missing=0.0, mode="fast"),
}


def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return time.perf_counter() - t0, r


def run(name, cfg, ctx, scale=1.0):
    N, M = cfg["N"], int(cfg["M"] * scale)
    rl = synth.rlen_for(N)
    rec = dict(config=name, N=N, M=M, missing=cfg["missing"], packed_GB=M * rl / 1e9)
    buf = torch.empty((M, rl), dtype=torch.uint8, device="cuda")
    t, _ = wall(lambda: (ctx.synth_packed_device(buf.data_ptr(), M, rl, N, seed=1, missing=cfg["missing"], npops=4, delta=0.05), ctx.sync()))
    rec["synth_s"] = t
    ctx.adopt_packed_device(buf.data_ptr(), M, rl, N)
    t, _ = wall(lambda: ctx.set_rows(None))
    tm = ctx.timings()
    rec["gather_ms"] = tm["gather_ms"]; rec["gather_GBs"] = 2 * M * (N / 4) / (tm["gather_ms"] * 1e-3) / 1e9      # read + write
    if cfg["mode"] == "full":
        t, r = wall(lambda: ctx.grm(want_snp=False))
        tm = ctx.timings()
        rec.update(grm_wall_s=t, stats_ms=tm["stats_ms"], grm_ms=tm["grm_ms"], finalize_ms=tm["finalize_ms"], nsplit=tm["nsplit"], nused=r["nused"])
        rec["stats_GBs"] = M * (N / 4) / (tm["stats_ms"] * 1e-3) / 1e9
        rec["grm_tflops"] = N * (N + 1.0) * r["nused"] / (tm["grm_ms"] * 1e-3) / 1e12
        rec["snp_indiv2_per_s"] = float(N) * N * r["nused"] / t
        t, (lam, vec) = wall(lambda: ctx.eig(10))
        tm = ctx.timings()
        rec.update(eig_wall_s=t, eig_method=tm["eig_method"], tridiag_ms=tm["tridiag_ms"], bisect_ms=tm["bisect_ms"], vectors_ms=tm["vectors_ms"],
                   chfsi_iters=tm["chfsi_iters"], chfsi_matvecs=tm["chfsi_matvecs"], lam_top=lam[:3].tolist())
        rec["tridiag_tflops"] = (4.0 / 3.0) * N ** 3 / (tm["tridiag_ms"] * 1e-3) / 1e12 if tm["tridiag_ms"] > 0 else None
        if cfg.get("lsq", True):
            t, (co, es, ok) = wall(lambda: ctx.evec_coords(vec))
            rec["evec_coords_s"] = t
            # loadings (1 sweep XA, 10 cols) + projections (1 sweep XTB) + lsqproj (XTB with 10 + 56 columns)
            rec["evec_coords_tflops"] = 2.0 * N * M * (10 + 10 + 10 + 56) / t / 1e12
        t, _ = wall(lambda: ctx.pop_counts(np.asarray(synth.pop_of(N, 4), np.int32), 4))
        rec["pop_counts_s"] = t
        rec["core_s"] = rec["grm_wall_s"] + rec["eig_wall_s"] + rec.get("evec_coords_s", 0.0)
    else:
        K, L, I = 10, 20, 10
        t, (ev, vec) = wall(lambda: ctx.fpca(K, L, I, seed=123))
        rec.update(fpca_wall_s=t, K=K, L=L, I=I, eval_top=ev[:3].tolist())
        flops = (4.0 * I * L + 2 * L + 2 * (I + 1) * L) * M * N
        rec["fpca_tflops"] = flops / t / 1e12
        rec["fpca_GBs"] = (I + 2 + I) * N * M / 4 / t / 1e9       # 2I+1 product sweeps + B sweep, each reads the packed matrix once
    del buf
    torch.cuda.empty_cache()
    return rec


def main():
    names = [a for a in sys.argv[1:] if a in CONFIGS] or list(CONFIGS)
    out_path = "gpurun_out/configs.json"
    if "--out" in sys.argv:
        out_path = sys.argv[sys.argv.index("--out") + 1]
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    ctx = capi.Context(0)
    out = {"microbench_tflops": dict(zip(("dmma", "dfma"), ctx.microbench_fp64()))}
    for nm in names:
        try:
            out[nm] = run(nm, CONFIGS[nm], ctx, scale)
        except Exception as ex:  # keep the earlier configurations' numbers
            out[nm] = {"error": repr(ex)[:500]}
        print(json.dumps(out[nm]), flush=True)
        json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
