"""2-GPU probe: where the non-kernel time of a first GRM pass at N = 50,000 goes (allocation, IPC mapping, reduce)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from eig_b200 import capi, parallel, synth
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = capi.Context(local); ctx.set_comm(parallel.TorchComm(device=dev))
N, M = 50000, 20000
rl = synth.rlen_for(N)
slab = torch.empty((M, rl), dtype=torch.uint8, device=dev)
ctx.synth_packed_device(slab.data_ptr(), M, rl, N, seed=1, s0=rank * M)
ctx.adopt_packed_device(slab.data_ptr(), M, rl, N); ctx.sync()
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.set_rows(None); t1 = time.perf_counter()
    r = ctx.grm(want_snp=False); torch.cuda.synchronize(); t2 = time.perf_counter()
    if rank == 0:
        print("pass %d: set_rows %.3f s, grm %.3f s (kernel %.1f ms, reduce %.1f ms)" % (rep, t1 - t0, t2 - t1, ctx.timings()["grm_ms"], ctx.timings()["finalize_ms"]), flush=True)
dist.barrier(); dist.destroy_process_group()
