"""world_size-2 gloo test of the multi-GPU host logic (SNP sharding + all-reduce of partial GRMs).  The per-rank partial
is produced by the oracle here (no GPU); on the GPU box the same plumbing runs over NCCL on the library's device buffer."""
import os
import subprocess
import sys

import numpy as np

from eig_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["EB_ROOT"])
from eig_b200 import parallel, synth
from oracle import bindings as ob
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["EB_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
nsnp, nind = 900, 64
P = synth.packed_genotypes(5, nsnp, nind, missing=0.1)
s0, s1 = parallel.shard_snps(nsnp, rank, 2)
part = ob.port_grm(P[s0:s1], nind)                      # this rank's shard only
xtx = parallel.allreduce_numpy_sum(part["XTX"].copy())
nused = parallel.allreduce_numpy_sum(np.array([int(part["used"].sum())], dtype=np.int64))
full = ob.port_grm(P, nind)
assert np.abs(xtx - full["XTX"]).max() < 1e-10 * np.abs(full["XTX"]).max()
assert int(nused[0]) == int(full["used"].sum())
assert np.array_equal(part["c0"], full["c0"][s0:s1]) and np.array_equal(part["used"], full["used"][s0:s1])
y = np.trace(xtx) / (nind - 1)
assert abs(y - full["y"]) < 1e-12 * full["y"]
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_shard_snps_partition():
    for nsnp, world in [(10, 3), (600000, 8), (7, 8), (128, 1)]:
        rng = [parallel.shard_snps(nsnp, r, world) for r in range(world)]
        assert rng[0][0] == 0 and rng[-1][1] == nsnp
        assert all(rng[i][1] == rng[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in rng]
        assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_partial_grm_allreduce(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), EB_ROOT=ROOT, EB_PORT=str(port), MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


COMM_WORKER = r'''
import ctypes as C, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["EB_ROOT"])
from eig_b200 import capi, parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["EB_PORT"], rank=int(os.environ["RANK"]), world_size=2)
comm = parallel.TorchComm()
assert (comm.rank, comm.world, comm.on_device) == (dist.get_rank(), 2, False)
# drive the plumbing exactly the way libeigb200 does: through the C callback types of eb_comm
ag = capi.ALLGATHER_CB(lambda user, src, dst, n: (comm.allgather_host(src, dst, n), 0)[1])
bar = capi.BARRIER_CB(lambda user: (comm.barrier(), 0)[1])
st = capi.Comm(comm.rank, comm.world, ag, bar, None)
rec = np.arange(96, dtype=np.uint8) + 100 * comm.rank           # sizeof(PeerRecord) = 96
out = np.zeros(2 * 96, np.uint8)
assert st.allgather_host(None, rec.ctypes.data, out.ctypes.data, 96) == 0
assert np.array_equal(out[:96], np.arange(96, dtype=np.uint8)) and np.array_equal(out[96:], (np.arange(96) + 100).astype(np.uint8))
assert st.barrier(None) == 0
dist.destroy_process_group()
print("rank", comm.rank, "ok")
'''


def test_gloo_world2_comm_callbacks():
    """eb_comm plumbing (all-gather of host records + barrier) over gloo, called through the C callback types"""
    port = 31500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), EB_ROOT=ROOT, EB_PORT=str(port), MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, "-c", COMM_WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
