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


def _world(group=None):
    import torch.distributed as dist
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
