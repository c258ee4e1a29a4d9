"""SNP-shard data parallelism for the GRM pass: one process per GPU, torch.distributed for plumbing.

Each rank owns a contiguous SNP shard and accumulates a partial N x N FP64 GRM with its own eb_ctx; the one exchange
step is an all-reduce of the partial GRMs (NCCL over NVLink on GPUs; gloo in the CPU tests), performed IN PLACE on the
library's device buffer, after which every rank finishes the pass (trace, y) and can run the eigensolver.
Per-SNP integer outputs need no exchange (each SNP lives on exactly one rank).
"""
import numpy as np


def shard_snps(nsnp, rank, world):
    """Contiguous, near-even SNP range [s0, s1) of this rank."""
    base, rem = divmod(int(nsnp), int(world))
    s0 = rank * base + min(rank, rem)
    return s0, s0 + base + (1 if rank < rem else 0)


class _CudaArray:
    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                         "strides": None}


def device_view(ptr, shape, device):
    """Zero-copy torch view of library-owned device memory."""
    import torch
    return torch.as_tensor(_CudaArray(ptr, shape), device=device)


def allreduce_sum_(t, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_numpy_sum(a, group=None):
    """Host-side variant used by the gloo tests and for small per-rank scalars."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    allreduce_sum_(t, group)
    return t.numpy()


class TorchComm:
    """eb_comm plumbing on top of torch.distributed: an all-gather of small host records and a barrier.  With the nccl
    backend the records travel through a device staging tensor; with gloo they stay on the host.  Nothing else of the
    sharded path goes through torch: the reductions themselves are the library's kernels over peer memory."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.on_device = dist.get_backend(group) == "nccl"

    def allgather_host(self, src_ptr, dst_ptr, nbytes):
        import ctypes
        import torch
        import torch.distributed as dist
        src = torch.frombuffer((ctypes.c_ubyte * nbytes).from_address(src_ptr), dtype=torch.uint8).clone()
        if self.on_device:
            src = src.to(self.device)
        out = torch.empty(self.world * nbytes, dtype=torch.uint8, device=src.device)
        dist.all_gather_into_tensor(out, src, group=self.group)
        out = out.cpu().contiguous()
        ctypes.memmove(dst_ptr, out.data_ptr(), self.world * nbytes)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.on_device:
            t = torch.zeros(1, device=self.device)
            dist.all_reduce(t, group=self.group)
            torch.cuda.synchronize(self.device)
        else:
            dist.barrier(group=self.group)


class ShardedGrm:
    """GRM pass over SNP shards: ctx holds THIS rank's shard (already uploaded / adopted, rows set)."""

    def __init__(self, ctx, device=None, group=None):
        self.ctx, self.device, self.group = ctx, device, group

    def grm(self, want_snp=True, **opts):
        import torch
        r = self.ctx.grm(want_snp=want_snp, partial=True, **opts)
        ptr, ld, n = self.ctx.grm_device_ptr()
        t = device_view(ptr, (ld, ld), self.device)
        allreduce_sum_(t, self.group)
        torch.cuda.synchronize(self.device)
        r["y"], _ = self.ctx.grm_finish()
        return r
