#!/usr/bin/env python
"""bench.py -- smartpca hot path on B200: GRM accumulation throughput (SNP*indiv^2/s) on synthetic Hardy-Weinberg genotypes.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]   (torchrun launches N>1, one rank per GPU)

Workload (BASELINE.json configs[1]): 5,000 individuals x 600,000 SNPs, no missing data, full-mode smartpca,
numoutevec=10, no outlier removal.  A "step" = one pass of the region smartpca.c:1088-1236 over one packed slab:
per-SNP allele counts + normalisation + drop rule, the packed->FP64 symmetric rank-M update, mirror + trace.
  value : inputs resident in HBM when the timed region starts (library entry eb_grm on an adopted device slab)
  e2e   : the same pass through the C-ABI with HOST buffers (eb_upload_packed from pinned memory, eb_set_rows,
          eb_grm, per-SNP outputs copied back) -- H2D/D2H inside the timed region
N>1: SNPs shard across ranks (each rank owns its own 600k-SNP slab: weak scaling); the one exchange step is the
reduction of the partial N x N FP64 GRMs, fused with the split-K finalize/mirror in the library's own kernel over peer
memory (NVLink; CUDA IPC between the torchrun ranks), inside the timed region.  --reduce nccl (or a box without IPC)
uses an NCCL all-reduce on the library's buffer instead and says so in config.parallelism.
The reference arm (--impl reference) times the reference's own CPU code for the same region (oracle/_ref, built from
/root/reference by oracle/Makefile) with all host threads on a bounded SNP sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IND = 5000
N_SNP = 600000
SEED = 1
METRIC = "grm_snp_indiv2_per_s"
UNIT = "SNP*indiv^2/s"
CPU_SAMPLE_SNPS = 20000


def workload_config(extra=None):
    cfg = {"workload": "smartpca full mode 5000 indiv x 600000 SNPs (BASELINE configs[1]), synthetic Hardy-Weinberg p~U(0.05,0.95), "
                       "no missing, fancynorm+altnormstyle YES, no outlier removal",
           "nindiv": N_IND, "nsnp_per_gpu": N_SNP, "seed": SEED,
           "l2": "packed input 750 MB per step > 126 MB L2 (no explicit flush needed)"}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.rows = []; self.stop_flag = False; self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def mark(self):
        """index of the next sample: call at the start / end of the timed region"""
        return len(self.rows)

    def finish(self, first=0, last=None):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        rows = self.rows[first:last] if (last is None or last > first) else self.rows[first:]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline_run(steps, warmup, budget_s=100.0, nthreads=None):
    """Reference CPU implementation of the same region on a bounded SNP sample (sized by a short calibration so that
    warmup+steps passes fit in ~budget_s of wall time)."""
    from eig_b200 import synth
    from oracle import bindings as ob
    nthreads = nthreads or os.cpu_count()
    kind = "reference" if ob.ref() is not None else "port"
    fn = (lambda P: ob.ref_grm(P, N_IND, nthreads=nthreads)) if kind == "reference" else (lambda P: ob.port_grm(P, N_IND))
    cal = 1000 if kind == "reference" else 100
    Pc = synth.packed_genotypes(SEED, cal, N_IND)
    t0 = time.perf_counter(); fn(Pc); rate = cal / (time.perf_counter() - t0)          # SNPs per second
    nsnp = int(min(CPU_SAMPLE_SNPS, max(cal, rate * budget_s / (steps + warmup))))
    P = synth.packed_genotypes(SEED, nsnp, N_IND)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = fn(P)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    used = int(r["used"].sum())
    sec = float(np.mean(times))
    threads_used = min(nthreads, 127) if kind == "reference" else 1
    return {"value": used * float(N_IND) ** 2 / sec, "unit": UNIT, "cores": threads_used, "kind": kind,
            "sample": "first %d of 600000 SNPs x %d individuals per step, %.2f s/step, %d threads, region smartpca.c:1088-1236 "
                      "(getcolxz_binary1/2 + domult_increment_lookup + symit2)" % (nsnp, N_IND, sec, threads_used), "secs_per_step": sec}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    cb = cpu_baseline_run(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": cb["secs_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config({"host_cpus": os.cpu_count()}),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from eig_b200 import capi, parallel, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = capi.Context(local)          # raises if the CUDA library / a B200 is missing: no fallback

    nind, nsnp = N_IND, N_SNP
    rl = synth.rlen_for(nind)
    slab = torch.empty((nsnp, rl), dtype=torch.uint8, device=dev)
    ctx.synth_packed_device(slab.data_ptr(), nsnp, rl, nind, seed=SEED, s0=rank * nsnp)     # each rank: its own SNP shard
    ctx.sync()
    host = torch.empty((nsnp, rl), dtype=torch.uint8, pin_memory=True)
    host.copy_(slab); torch.cuda.synchronize()
    host_np = host.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # exchange step: the library's peer-memory kernel unless asked otherwise / unavailable on this box (all ranks agree)
    reduce_mode = "none"
    if world > 1:
        reduce_mode = args.reduce
        if reduce_mode == "peer":
            ok = 1.0
            try:
                ctx.set_comm(parallel.TorchComm(device=dev))
                chk = ctx.peer_allreduce_test(np.full(64, float(rank + 1)))
                ok = 1.0 if np.all(chk == world * (world + 1) / 2) else 0.0
            except capi.EigB200Error as ex:
                print("rank %d: peer path unavailable: %s" % (rank, ex), file=sys.stderr)
                ok = 0.0
            t = torch.tensor([ok], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if t.item() < 1.0:
                ctx.set_comm(None)
                reduce_mode = "nccl"

    def reduce_partials():
        ptr, ld, n = ctx.grm_device_ptr()
        t = parallel.device_view(ptr, (ld, ld), dev)
        dist.all_reduce(t)             # every rank ends with the full GRM (the eigensolver then runs replicated / row-distributed)
        torch.cuda.synchronize()

    def step_resident():
        if world == 1 or reduce_mode == "peer":
            return ctx.grm(want_snp=False)
        r = ctx.grm(want_snp=False, partial=True)
        reduce_partials()
        r["y"], _ = ctx.grm_finish()
        return r

    def step_e2e():
        ctx.upload_packed(host_np, nind)
        ctx.set_rows(None)
        if world == 1 or reduce_mode == "peer":
            return ctx.grm(want_snp=True)
        r = ctx.grm(want_snp=True, partial=True)
        reduce_partials()
        r["y"], _ = ctx.grm_finish()
        return r

    def timed(fn, steps, warmup, sample_clocks=False):
        # the clock sampler (an nvidia-smi child per rank) starts BEFORE the warm-up: its start-up (NVML enumeration of
        # every GPU of the box) contends with CUDA calls and must not fall into the timed region; only the samples taken
        # between the two marks are reported
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start(); time.sleep(1.0)
        for _ in range(warmup):
            fn()
        barrier()
        m0 = sampler.mark() if sampler else 0
        ctx.reset_launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        kern = []
        e0.record()
        for _ in range(steps):
            r = fn()
            kern.append(ctx.timings())
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count()
        clocks = sampler.finish(m0, sampler.mark()) if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r, kern, launches, clocks

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    # clock ramp (not a step of the workload): a fresh box idles at low clocks and the first second of FP64 tensor work runs
    # up to 20 % slow (seen as 524 vs 444 ms kernels on the first run after boot); spin the DMMA issue-rate probe for ~2 s
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < 2.0:
        ctx.microbench_fp64()
    # resident arm
    ctx.adopt_packed_device(slab.data_ptr(), nsnp, rl, nind)
    ctx.set_rows(None)
    ms, r, kern, launches, clocks = timed(step_resident, steps, warmup, sample_clocks=True)
    own_used = int(ctx.snp_used_count())           # this shard's SNPs that entered XTX
    used = torch.tensor([own_used], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(used)
    if reduce_mode == "peer":
        assert int(used.item()) == int(r["nused"]), "library total of used SNPs disagrees with the per-shard sum"
    units = float(used.item()) * float(nind) ** 2
    ms_per_step = ms / steps
    value = units / (ms_per_step * 1e-3)
    grm_ms = float(np.mean([k["grm_ms"] for k in kern]))
    flops = float(nind) * (nind + 1.0) * own_used
    achieved = flops / (grm_ms * 1e-3) / 1e12

    # e2e arm (host buffers through the C-ABI)
    ems, er, _, _, _ = timed(step_e2e, max(1, min(steps, 5)), 1)
    esteps = max(1, min(steps, 5))
    e2e_value = units / (ems / esteps * 1e-3)
    h2d = int(nsnp * rl + 4 * nind)
    d2h = int(nsnp * (4 * 3 + 1 + 8 * 2) + 16)

    line = None
    if rank == 0:
        # FP64 pipe peak: MEASURED_PEAKS.json carries no FP64 entry, so measure cuBLAS DGEMM here (same box, same run)
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize(); best = 1e30
        for _ in range(5):
            q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
            q0.record(); torch.matmul(a, b); q1.record(); torch.cuda.synchronize(); best = min(best, q0.elapsed_time(q1))
        peak = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        del a, b
        dmma, dfma = ctx.microbench_fp64()
        # full pipeline once (adds the eigensolver): upload -> rows -> GRM -> all eigenvalues + 10 vectors
        if world == 1:
            t0 = time.perf_counter()
            ctx.upload_packed(host_np, nind); ctx.set_rows(None); ctx.grm(want_snp=True); lam, vec = ctx.eig(10)
            pipeline_s = time.perf_counter() - t0
            eig_t = ctx.timings()
        else:
            pipeline_s, eig_t = None, {}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "grm_syrk_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_baseline_run(1, 0, budget_s=20.0)
                cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:      # the checker is optional for the product line
                cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config({"parallelism": ("snp-shard x%d + %s" % (world, "fused finalize+reduce kernel over peer memory (CUDA IPC / NVLink)" if reduce_mode == "peer"
                                                            else "nccl all_reduce of partial GRM")) if world > 1 else "single GPU",
                                           "nsplit": kern[-1]["nsplit"]}),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ems / esteps},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "kernel": "grm_syrk_kernel (FP64 DMMA.8x8x4)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "traffic": traffic, "kernel_ms": grm_ms,
                             "peak_source": "cuBLAS DGEMM fp64 8192^3 best-of-5 measured in this run (MEASURED_PEAKS.json has no FP64 entry); "
                                            "microbench issue rates: DMMA %.1f, DFMA %.1f TFLOP/s" % (dmma, dfma),
                             "algorithmic": "N(N+1)*M_used flops per launch (lower triangle incl. diagonal, FMA=2)"},
                "cpu_baseline": cb,
                "kernel_ms": {k: float(np.mean([q[k] for q in kern])) for k in ("stats_ms", "grm_ms", "finalize_ms")},
                "smartpca_core_s": pipeline_s,
                "eig_ms": {k: eig_t.get(k) for k in ("tridiag_ms", "bisect_ms", "vectors_ms")} if eig_t else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"], help="N>1 exchange step: library peer-memory kernel | NCCL all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
