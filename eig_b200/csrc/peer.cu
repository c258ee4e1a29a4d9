// peer.cu -- the exchange steps of the SNP-sharded path as kernels over peer memory (NVLink 5 / NVSwitch).
//
// One context per GPU owns one SNP shard.  The launcher (torchrun + torch.distributed in bench/tests; MPI, pipes or
// threads elsewhere) supplies only host-side plumbing through eb_comm: an all-gather of small host records and a
// barrier.  Device buffers are exported with CUDA IPC (or used directly when two ranks live in one process), after
// which every exchange is one kernel that reads the peers' buffers with plain loads and writes its reduced slice into
// every peer with plain stores (all-reduce) or lets the peers pull it (GRM):
//
//   grm_peer_reduce_kernel / grm_peer_gather_kernel : the split-K plane sum + symit2 mirror of grm_finalize_kernel FUSED
//       with the cross-GPU reduction of the partial GRMs.  Lower-triangle 32x32 blocks are dealt round-robin to the
//       ranks; the owner sums rank 0's planes, then rank 1's, ... (fixed order => the result is bit-identical on every
//       rank and from run to run) into its own XTX and its own plane 0; after a barrier every rank pulls the other
//       owners' blocks and mirrors them locally.  Only the lower triangle ever crosses NVLink (2 x 7/8 of it per rank,
//       a quarter of what an all-reduce of the square moves) and only the plane buffers are mapped between processes.
//   peer_allreduce_kernel    : in-place one-shot all-reduce of an FP64 buffer (fastmode's N x L sketch and N x (I+1)L
//       projection, kjg_fpca.c:123,79): rank r owns a contiguous slice, pulls it from every rank, pushes the sum back.
//
// Ordering: a host barrier separates "all inputs written" from the kernel, and the kernel's completion (stream sync on
// every rank + barrier) from any consumer, so no device-side flags are needed and a hung peer cannot wedge the GPU.
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include "common.cuh"
#include "tile_order.cuh"

namespace eb {

static int comm_barrier(eb_ctx* c) {
  if (c->comm.barrier(c->comm.user) != 0) { set_error("eb_comm.barrier failed"); return EB_ERR_STATE; }
  return 0;
}

int peer_allgather_host(eb_ctx* c, const void* src, void* dst, int64_t bytes) {
  if (c->comm.allgather_host(c->comm.user, src, dst, bytes) != 0) { set_error("eb_comm.allgather_host failed"); return EB_ERR_STATE; }
  return 0;
}

int peer_bury(eb_ctx* c) {
  bool any = false;
  for (auto& reg : c->peer) any |= !reg.graveyard.empty();
  // every rank calls this at the same point; the barrier guarantees that all peers have already re-mapped
  int rc = comm_barrier(c);
  if (rc) return rc;
  if (any)
    for (auto& reg : c->peer) { for (void* p : reg.graveyard) cudaFree(p); reg.graveyard.clear(); }
  return 0;
}

void peer_release(eb_ctx* c) {
  for (auto& reg : c->peer) {
    for (size_t r = 0; r < reg.mapped.size(); r++)
      if (reg.opened[r] && reg.mapped[r]) cudaIpcCloseMemHandle(reg.mapped[r]);
    reg.mapped.clear(); reg.opened.clear(); reg.rec.clear();
  }
}

// Publish `local` (the base of a cudaMalloc allocation of `bytes`) in slot `slot` and map every peer's buffer of the
// same slot.  Collective.  Handles are re-opened only when a rank's allocation changed since the last call.
int peer_exchange(eb_ctx* c, int slot, void* local, size_t bytes, int aux) {
  const int W = c->comm.world, me = c->comm.rank;
  PeerRegion& reg = c->peer[slot];
  if ((int)reg.rec.size() != W) {
    reg.rec.assign(W, PeerRecord{}); reg.mapped.assign(W, nullptr); reg.opened.assign(W, false);
  }
  PeerRecord mine;
  memset(&mine, 0, sizeof(mine));
  if (reg.rec[me].ptr == (uint64_t)(uintptr_t)local && reg.rec[me].bytes == bytes) mine = reg.rec[me];
  else EB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)mine.handle, local));
  mine.ptr = (uint64_t)(uintptr_t)local; mine.bytes = bytes; mine.pid = (int64_t)getpid(); mine.device = c->device; mine.aux = aux;
  std::vector<PeerRecord> all(W);
  int rc;
  if ((rc = peer_allgather_host(c, &mine, all.data(), sizeof(PeerRecord)))) return rc;
  for (int r = 0; r < W; r++) {
    const PeerRecord& n = all[r];
    const PeerRecord& o = reg.rec[r];
    const bool same = o.ptr == n.ptr && o.bytes == n.bytes && o.pid == n.pid && o.device == n.device &&
                      !memcmp(o.handle, n.handle, sizeof(n.handle)) && reg.mapped[r] != nullptr;
    if (!same) {
      if (reg.opened[r] && reg.mapped[r]) { cudaIpcCloseMemHandle(reg.mapped[r]); }
      reg.opened[r] = false; reg.mapped[r] = nullptr;
      if (r == me) reg.mapped[r] = local;
      else if (n.pid == mine.pid) {                       // same process (threads / several contexts): direct peer access
        if (n.device != c->device) {
          cudaError_t e = cudaDeviceEnablePeerAccess(n.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            set_error("peer access %d -> %d: %s", c->device, n.device, cudaGetErrorString(e)); cudaGetLastError(); return EB_ERR_CUDA;
          }
          cudaGetLastError();
        }
        reg.mapped[r] = (void*)(uintptr_t)n.ptr;
      } else {
        void* p = nullptr;
        cudaIpcMemHandle_t h;
        memcpy(&h, n.handle, sizeof(h));
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          set_error("cudaIpcOpenMemHandle (rank %d, device %d -> %d): %s", r, n.device, c->device, cudaGetErrorString(e));
          cudaGetLastError();
          return EB_ERR_CUDA;
        }
        reg.mapped[r] = p; reg.opened[r] = true;
      }
    }
    reg.rec[r] = n;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ sharded GRM exchange
// grm_syrk_kernel's epilogue has already stored every partial 128 x 128 tile into its owner's receive buffer (GrmPush,
// common.cuh): the reduce-scatter traffic rode under the DMMA work.  What is left, all of it STREAM-ORDERED (no host
// barrier, no host collective in the steady state):
//   signal(0)  my pushes are complete (kernel boundary + fence.sys) -> flag in every peer, with my used-SNP count
//   wait(0)    every rank's pushes have landed in my receive buffer
//   reduce     owned tiles: sum of the world * nsplit slots in a fixed order (bit-identical everywhere and from run to run)
//              -> slot 0 in place (for the peers to pull) and my XTX (+ mirror image: symit2)
//   signal(1) / wait(1), gather: pull every other owner's reduced tiles from its slot 0 into my XTX (+ mirror)
//   signal(2)  done pulling; the NEXT pass waits for everybody's (2) before its first tile store (peer_grm_wait_idle)
// Flags are monotone epochs in a small exported buffer; a wait gives up after GRM_WAIT_TIMEOUT_NS and raises the error
// word instead of wedging the GPU when a peer died.
// One configuration cannot spin on the device: two ranks that are contexts of ONE process on ONE GPU (the single-GPU test
// set-up of eb_local_comm).  There a lagging rank's cudaFree synchronises the whole device and would wait for the leading
// rank's spinning kernel, which waits for the lagging rank: the waits then go through the host (stream sync + eb_comm.barrier).

constexpr unsigned long long GRM_WAIT_TIMEOUT_NS = 60ull * 1000000000ull;
constexpr int FLAG_ERR = 3 * 16 + 2 * 16 * 2;     // word index of the error flag

struct GrmFlagArgs {
  unsigned long long* flags[EB_MAX_WORLD];         // every rank's flag buffer (mine included)
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// one warp: lane w tells rank w "step `word0 / 16` of epoch `epoch` is complete on rank a.rank" (optionally with a payload word).
// Flag rows (16 words each): 0-2 the three phases of the sharded GRM, 8 / 9 the two phases of the stream-ordered all-reduce.
__global__ void __launch_bounds__(32) grm_signal_kernel(const GrmFlagArgs a, int word0, unsigned long long epoch,
                                                        const unsigned long long* __restrict__ payload) {
  const int w = threadIdx.x;
  if (w >= a.world) return;
  __threadfence_system();
  if (payload) {
    a.flags[w][3 * 16 + ((epoch & 1) * 16 + a.rank) * 2] = payload[0];
    __threadfence_system();
  }
  st_release_sys(a.flags[w] + word0 + a.rank, epoch);
}

// one warp: lane w waits until rank w has signalled `phase` of `epoch` into MY flag buffer
__global__ void __launch_bounds__(32) grm_wait_kernel(unsigned long long* mine, int world, int word0, unsigned long long epoch) {
  const int w = threadIdx.x;
  if (w < world) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(mine + word0 + w) < epoch) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > GRM_WAIT_TIMEOUT_NS) { mine[FLAG_ERR] = 1ull + (unsigned long long)w; break; }
      __nanosleep(200);
    }
  }
  __syncwarp();
  __threadfence_system();
}

// 32 x 32 block (bi, bj) of a dense 128 x 128 lower-triangle tile (ti, tj) -> xtx and its mirror image.  `diag`: the tile lies on the
// diagonal, so blocks above its own diagonal are skipped and the diagonal blocks are mirrored inside themselves.
__device__ __forceinline__ void emit_block(double (&tile)[32][33], int ti, int tj, int bi, int bj, bool diag, int npad, double* __restrict__ xtx) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool dblk = diag && bi == bj;
  const size_t r0 = (size_t)ti * TILE + bi * 32, c0 = (size_t)tj * TILE + bj * 32;
  for (int r = ty; r < 32; r += 8) {
    double v = tile[r][tx];
    if (dblk && tx > r) v = tile[tx][r];
    xtx[(r0 + r) * npad + c0 + tx] = v;
    if (!dblk) xtx[(c0 + r) * npad + r0 + tx] = tile[tx][r];
  }
}

struct GrmReduceArgs {
  const double* recv[EB_MAX_WORLD];     // every rank's receive buffer (mine at [rank])
  int world, rank;
};

// CTA b of rank r finalises its b-th owned tile (global tile r + b * world)
__global__ void __launch_bounds__(256) grm_push_reduce_kernel(double* __restrict__ recv, int world, int rank, int nslots, int ntri, int npad,
                                                              double* __restrict__ xtx) {
  __shared__ double tile[32][33];
  const int t = rank + blockIdx.x * world;
  if (t >= ntri) return;
  int ti, tj; tile_decode_banded(t, npad / TILE, ti, tj);      // same walk as grm_syrk_kernel: tile index t -> position
  const bool diag = ti == tj;
  double* slot0 = recv + (size_t)blockIdx.x * nslots * (TILE * TILE);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int bi = 0; bi < 4; bi++)
    for (int bj = 0; bj < 4; bj++) {
      if (diag && bj > bi) continue;
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const size_t idx = (size_t)(bi * 32 + r) * TILE + bj * 32 + tx;
        double v = 0.0;
        for (int s = 0; s < nslots; s++) v += slot0[(size_t)s * (TILE * TILE) + idx];
        slot0[idx] = v;
        tile[r][tx] = v;
      }
      __syncthreads();
      emit_block(tile, ti, tj, bi, bj, diag, npad, xtx);
    }
}

// CTA b pulls the b-th tile NOT owned by this rank from its owner's slot 0
__global__ void __launch_bounds__(256) grm_push_gather_kernel(const GrmReduceArgs a, int nslots, int ntri, int npad, double* __restrict__ xtx) {
  __shared__ double tile[32][33];
  const int grp = blockIdx.x / (a.world - 1), k = blockIdx.x % (a.world - 1);
  const int t = grp * a.world + (k < a.rank ? k : k + 1);
  if (t >= ntri) return;
  const int owner = t % a.world;
  int ti, tj; tile_decode_banded(t, npad / TILE, ti, tj);      // same walk as grm_syrk_kernel: tile index t -> position
  const bool diag = ti == tj;
  const double* src = a.recv[owner] + (size_t)(t / a.world) * nslots * (TILE * TILE);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int bi = 0; bi < 4; bi++)
    for (int bj = 0; bj < 4; bj++) {
      if (diag && bj > bi) continue;
      __syncthreads();
      for (int r = ty; r < 32; r += 8) tile[r][tx] = src[(size_t)(bi * 32 + r) * TILE + bj * 32 + tx];
      __syncthreads();
      emit_block(tile, ti, tj, bi, bj, diag, npad, xtx);
    }
}

static int flag_args(eb_ctx* c, GrmFlagArgs* a) {
  memset(a, 0, sizeof(*a));
  a->world = c->comm.world; a->rank = c->comm.rank;
  for (int r = 0; r < a->world; r++) {
    a->flags[r] = (unsigned long long*)c->peer[PEER_SLOT_FLAGS].mapped[r];
    if (!a->flags[r]) { set_error("sharded GRM: flag buffer of rank %d is not mapped (peer_grm_setup not called)", r); return EB_ERR_STATE; }
  }
  return 0;
}

// Make sure every rank's receive buffer is large enough for the current matrix and mapped everywhere.  The geometry is a function
// of (npad, world, nsplit) alone and every rank calls with the same rows, so ranks agree without talking; the host collectives
// below run only when the matrix outgrew what is mapped (normally once, before the first kernel of the run is launched, so
// cudaIpcOpenMemHandle does not wait behind a running SYRK).
int peer_grm_setup(eb_ctx* c, int nsplit) {
  const int W = c->comm.world;
  const int T = c->npad / TILE, ntri = T * (T + 1) / 2;
  const size_t owned_max = (size_t)(ntri + W - 1) / W;
  const size_t need = owned_max * (size_t)(W * nsplit) * (TILE * TILE);
  c->grm_geom_npad = c->npad; c->grm_geom_nsplit = nsplit; c->grm_recv_need = need;
  const bool mapped = (int)c->peer[PEER_SLOT_PARTIAL].mapped.size() == W && (int)c->peer[PEER_SLOT_FLAGS].mapped.size() == W;
  if (mapped && c->grm_recv.p && c->grm_recv.n >= need) return 0;
  int rc;
  // grow only (outlier passes shrink the matrix and keep the mapping)
  EB_CUDA(cudaStreamSynchronize(c->stream));
  if ((rc = c->grm_recv.ensure(need))) return rc;
  if ((rc = peer_exchange(c, PEER_SLOT_PARTIAL, c->grm_recv.p, c->grm_recv.n * sizeof(double), 0))) return rc;
  if ((rc = peer_flags_setup(c))) return rc;
  for (int r = 0; r < W; r++)
    if (c->peer[PEER_SLOT_PARTIAL].rec[r].bytes < need * sizeof(double)) {
      set_error("sharded GRM: rank %d holds a smaller receive buffer (%llu < %llu bytes): every shard must use the same rows", r,
                (unsigned long long)c->peer[PEER_SLOT_PARTIAL].rec[r].bytes, (unsigned long long)(need * sizeof(double)));
      return EB_ERR_STATE;
    }
  return peer_bury(c);       // barrier: everybody has re-mapped, superseded allocations can go
}

// The small flag / mailbox buffer every rank exports once per communicator (collective; a no-op once it is mapped).
int peer_flags_setup(eb_ctx* c) {
  const int W = c->comm.world;
  if (c->grm_flags.p && (int)c->peer[PEER_SLOT_FLAGS].mapped.size() == W && c->peer[PEER_SLOT_FLAGS].mapped[c->comm.rank]) return 0;
  int rc;
  if (!c->grm_flags.p) {
    if ((rc = c->grm_flags.ensure(GRM_FLAG_WORDS))) return rc;
    EB_CUDA(cudaMemsetAsync(c->grm_flags.p, 0, sizeof(unsigned long long) * GRM_FLAG_WORDS, c->stream));
    EB_CUDA(cudaStreamSynchronize(c->stream));
    c->grm_epoch = 0; c->ar_epoch = 0;
  }
  if ((rc = peer_exchange(c, PEER_SLOT_FLAGS, c->grm_flags.p, c->grm_flags.n * sizeof(unsigned long long), 0))) return rc;
  // every rank sees every record, so all ranks take the same decision
  c->grm_host_sync = false;
  for (int r = 0; r < W; r++)
    for (int q = r + 1; q < W; q++) {
      const PeerRecord &a = c->peer[PEER_SLOT_FLAGS].rec[r], &b = c->peer[PEER_SLOT_FLAGS].rec[q];
      if (a.pid == b.pid && a.device == b.device) c->grm_host_sync = true;
    }
  return 0;
}

// wait until every rank has signalled `phase` of `epoch`: a spinning warp on the stream, or (grm_host_sync) through the host
static int grm_wait(eb_ctx* c, int phase, unsigned long long epoch) {
  if (c->grm_host_sync) {
    EB_CUDA(cudaStreamSynchronize(c->stream));
    return comm_barrier(c);
  }
  grm_wait_kernel<<<1, 32, 0, c->stream>>>(c->grm_flags.p, c->comm.world, phase * 16, epoch);
  EB_CHECK_LAUNCH(c);
  return 0;
}

int peer_grm_push_args(eb_ctx* c, GrmPush* out) {
  memset(out, 0, sizeof(*out));
  out->world = c->comm.world; out->rank = c->comm.rank;
  for (int r = 0; r < out->world; r++) {
    out->recv[r] = (double*)c->peer[PEER_SLOT_PARTIAL].mapped[r];
    if (!out->recv[r]) { set_error("sharded GRM: receive buffer of rank %d is not mapped (peer_grm_setup not called)", r); return EB_ERR_STATE; }
  }
  return 0;
}

int peer_grm_wait_idle(eb_ctx* c) {
  if (c->grm_epoch == 0 || !c->grm_flags.p || c->grm_host_sync) return 0;   // host-sync mode closes every pass with its own barrier
  return grm_wait(c, 2, c->grm_epoch);
}

int peer_grm_finalize(eb_ctx* c, int nsplit) {
  const int W = c->comm.world, me = c->comm.rank;
  int rc;
  GrmFlagArgs fa;
  if ((rc = flag_args(c, &fa))) return rc;
  GrmReduceArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.world = W; ra.rank = me;
  for (int r = 0; r < W; r++) ra.recv[r] = (const double*)c->peer[PEER_SLOT_PARTIAL].mapped[r];
  const unsigned long long ep = ++c->grm_epoch;
  const int T = c->npad / TILE, ntri = T * (T + 1) / 2, nslots = W * nsplit;
  grm_signal_kernel<<<1, 32, 0, c->stream>>>(fa, 0 * 16, ep, reinterpret_cast<const unsigned long long*>(c->nused_d.p));   // payload: my used-SNP count
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[5], c->stream));
  if ((rc = grm_wait(c, 0, ep))) return rc;
  EB_CUDA(cudaEventRecord(c->ev[6], c->stream));
  const int mine = (ntri - me + W - 1) / W;
  if (mine > 0) {
    grm_push_reduce_kernel<<<mine, 256, 0, c->stream>>>(c->grm_recv.p, W, me, nslots, ntri, c->npad, c->xtx.p);
    EB_CHECK_LAUNCH(c);
  }
  grm_signal_kernel<<<1, 32, 0, c->stream>>>(fa, 1 * 16, ep, nullptr);
  EB_CHECK_LAUNCH(c);
  if ((rc = grm_wait(c, 1, ep))) return rc;
  const int groups = (ntri + W - 1) / W;
  if (W > 1 && groups > 0) {
    grm_push_gather_kernel<<<groups * (W - 1), 256, 0, c->stream>>>(ra, nslots, ntri, c->npad, c->xtx.p);
    EB_CHECK_LAUNCH(c);
  }
  grm_signal_kernel<<<1, 32, 0, c->stream>>>(fa, 2 * 16, ep, nullptr);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[4], c->stream));
  if (c->grm_host_sync) return grm_wait(c, 2, ep);      // nothing left outstanding between passes in this mode
  return 0;
}

// after the stream has been synchronised: total of the ranks' used-SNP counts (mailbox of this epoch) and the error word
int peer_grm_collect(eb_ctx* c, long long* nused_total) {
  unsigned long long h[GRM_FLAG_WORDS];
  EB_CUDA(cudaMemcpyAsync(h, c->grm_flags.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  if (h[FLAG_ERR]) {
    set_error("sharded GRM: rank %d never signalled (waited %d s): a peer died or the ranks do not run the same passes", (int)h[FLAG_ERR] - 1,
              (int)(GRM_WAIT_TIMEOUT_NS / 1000000000ull));
    EB_CUDA(cudaMemsetAsync(c->grm_flags.p + FLAG_ERR, 0, sizeof(unsigned long long), c->stream));
    return EB_ERR_STATE;
  }
  long long tot = 0;
  for (int r = 0; r < c->comm.world; r++) tot += (long long)h[3 * 16 + ((c->grm_epoch & 1) * 16 + r) * 2];
  if (nused_total) *nused_total = tot;
  return 0;
}

struct AllreduceArgs {
  double* buf[EB_MAX_WORLD];
  int world, rank;
};

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const AllreduceArgs a, int64_t v0, int64_t v1) {
  for (int64_t i = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += (int64_t)gridDim.x * blockDim.x) {
    double2 s = make_double2(0.0, 0.0);
    for (int w = 0; w < a.world; w++) {
      const double2 t = reinterpret_cast<const double2*>(a.buf[w])[i];
      s.x += t.x; s.y += t.y;
    }
    for (int w = 0; w < a.world; w++) reinterpret_cast<double2*>(a.buf[(w + a.rank) % a.world])[i] = s;
  }
}

// In-place sum over ranks of `count` doubles at the start of an exported allocation (count even).  Collective.
int peer_allreduce(eb_ctx* c, int slot, double* buf, size_t alloc_doubles, int64_t count) {
  const int W = c->comm.world;
  int rc;
  if (count & 1) { set_error("peer_allreduce: odd element count"); return EB_ERR_ARG; }
  if ((rc = peer_exchange(c, slot, buf, alloc_doubles * sizeof(double), (int)(count & 0x7fffffff)))) return rc;
  AllreduceArgs a;
  memset(&a, 0, sizeof(a));
  a.world = W; a.rank = c->comm.rank;
  for (int r = 0; r < W; r++) {
    if (c->peer[slot].rec[r].aux != (int)(count & 0x7fffffff)) { set_error("peer_allreduce: rank %d disagrees on the element count", r); return EB_ERR_STATE; }
    a.buf[r] = (double*)c->peer[slot].mapped[r];
  }
  EB_CUDA(cudaStreamSynchronize(c->stream));
  if ((rc = comm_barrier(c))) return rc;
  const int64_t nvec = count / 2, per = (nvec + W - 1) / W;
  const int64_t v0 = std::min<int64_t>(nvec, per * a.rank), v1 = std::min<int64_t>(nvec, v0 + per);
  if (v1 > v0) {
    const int grid = (int)std::min<int64_t>((v1 - v0 + 255) / 256, (int64_t)c->num_sms * 8);
    peer_allreduce_kernel<<<grid, 256, 0, c->stream>>>(a, v0, v1);
    EB_CHECK_LAUNCH(c);
  }
  EB_CUDA(cudaStreamSynchronize(c->stream));
  if ((rc = comm_barrier(c))) return rc;
  return 0;
}

// The same all-reduce without the host in the loop: "my input is complete" / "my slice is reduced and pushed" travel as device
// flags (rows 8 and 9 of the flag buffer, epoch ar_epoch), so a chain of kernels, all-reduces and more kernels stays on the stream --
// what the collective subspace iteration needs for its hundreds of 64 x n products.  exchange: publish / map `buf` first (host
// collective; pass true whenever the buffer may have been reallocated, all ranks alike).
int peer_allreduce_stream(eb_ctx* c, int slot, double* buf, size_t alloc_doubles, int64_t count, bool exchange, int64_t offset) {
  const int W = c->comm.world;
  int rc;
  if ((count & 1) || (offset & 1)) { set_error("peer_allreduce_stream: odd element count / offset"); return EB_ERR_ARG; }
  if (exchange) {
    if ((rc = peer_exchange(c, slot, buf, alloc_doubles * sizeof(double), (int)(count & 0x7fffffff)))) return rc;
    if ((rc = peer_flags_setup(c))) return rc;
    if ((rc = peer_bury(c))) return rc;
  }
  AllreduceArgs a;
  memset(&a, 0, sizeof(a));
  a.world = W; a.rank = c->comm.rank;
  for (int r = 0; r < W; r++) {
    a.buf[r] = (double*)c->peer[slot].mapped[r];
    if (!a.buf[r]) { set_error("peer_allreduce_stream: buffer of rank %d is not mapped", r); return EB_ERR_STATE; }
    a.buf[r] += offset;                      // `count` doubles starting `offset` doubles into every rank's exported allocation
  }
  GrmFlagArgs fa;
  if ((rc = flag_args(c, &fa))) return rc;
  const unsigned long long ep = ++c->ar_epoch;
  auto sync_step = [&](int row) -> int {
    if (c->grm_host_sync) { EB_CUDA(cudaStreamSynchronize(c->stream)); return comm_barrier(c); }
    grm_signal_kernel<<<1, 32, 0, c->stream>>>(fa, row * 16, ep, nullptr);
    EB_CHECK_LAUNCH(c);
    grm_wait_kernel<<<1, 32, 0, c->stream>>>(c->grm_flags.p, W, row * 16, ep);
    EB_CHECK_LAUNCH(c);
    return 0;
  };
  if ((rc = sync_step(8))) return rc;
  const int64_t nvec = count / 2, per = (nvec + W - 1) / W;
  const int64_t v0 = std::min<int64_t>(nvec, per * a.rank), v1 = std::min<int64_t>(nvec, v0 + per);
  if (v1 > v0) {
    const int grid = (int)std::min<int64_t>((v1 - v0 + 255) / 256, (int64_t)c->num_sms * 8);
    peer_allreduce_kernel<<<grid, 256, 0, c->stream>>>(a, v0, v1);
    EB_CHECK_LAUNCH(c);
  }
  return sync_step(9);
}

// Sum over ranks of an arbitrary device buffer (any address, any owner): staged through the exported scratch allocation.
// For the small per-individual / per-sample sums of the projection, lsqproj and shrinkmode passes.  Collective.
int peer_allreduce_any(eb_ctx* c, double* buf, int64_t count) {
  if (!c->has_comm) return 0;
  const int64_t cnt2 = (count + 1) & ~1ll;
  int rc;
  if ((rc = c->peer_scratch.ensure((size_t)cnt2))) return rc;
  if (cnt2 != count) EB_CUDA(cudaMemsetAsync(c->peer_scratch.p + count, 0, sizeof(double), c->stream));
  EB_CUDA(cudaMemcpyAsync(c->peer_scratch.p, buf, sizeof(double) * count, cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = peer_allreduce(c, PEER_SLOT_T, c->peer_scratch.p, c->peer_scratch.n, cnt2))) return rc;
  EB_CUDA(cudaMemcpyAsync(buf, c->peer_scratch.p, sizeof(double) * count, cudaMemcpyDeviceToDevice, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return peer_bury(c);
}

// host-side sum over ranks of a few doubles (in place)
int peer_sum_host(eb_ctx* c, double* v, int count) {
  if (!c->has_comm) return 0;
  std::vector<double> all((size_t)count * c->comm.world);
  int rc;
  if ((rc = peer_allgather_host(c, v, all.data(), sizeof(double) * count))) return rc;
  for (int i = 0; i < count; i++) {
    double s = 0.0;
    for (int r = 0; r < c->comm.world; r++) s += all[(size_t)r * count + i];
    v[i] = s;
  }
  return 0;
}

}  // namespace eb

extern "C" int eb_set_comm(eb_ctx* c, const eb_comm* comm) {
  if (!c) return EB_ERR_ARG;
  cudaSetDevice(c->device);
  // leaving a communicator: the peers may still be pulling the last pass out of my receive buffer
  if (c->has_comm && c->grm_epoch > 0 && c->grm_flags.p) { eb::peer_grm_wait_idle(c); cudaStreamSynchronize(c->stream); }
  eb::peer_release(c);
  c->grm_recv.defer = c->fpG.defer = c->fpB.defer = c->fpS.defer = c->peer_scratch.defer = c->chfsi_sum.defer = nullptr;
  for (auto& reg : c->peer) { for (void* p : reg.graveyard) cudaFree(p); reg.graveyard.clear(); }
  c->grm_epoch = 0; c->ar_epoch = 0; c->grm_recv.release(); c->grm_flags.release(); c->chfsi_sum.release();
  if (!comm || comm->world <= 1) { c->has_comm = false; memset(&c->comm, 0, sizeof(c->comm)); c->comm.world = 1; return 0; }
  if (comm->world > eb::EB_MAX_WORLD || comm->rank < 0 || comm->rank >= comm->world || !comm->allgather_host || !comm->barrier) {
    eb::set_error("eb_set_comm: need 0 <= rank < world <= %d and both callbacks", eb::EB_MAX_WORLD);
    return EB_ERR_ARG;
  }
  c->comm = *comm; c->has_comm = true;
  c->grm_recv.defer = &c->peer[eb::PEER_SLOT_PARTIAL].graveyard;
  c->fpG.defer = &c->peer[eb::PEER_SLOT_A].graveyard;
  c->fpB.defer = &c->peer[eb::PEER_SLOT_B].graveyard;
  c->fpS.defer = &c->peer[eb::PEER_SLOT_C].graveyard;
  c->peer_scratch.defer = &c->peer[eb::PEER_SLOT_T].graveyard;
  c->chfsi_sum.defer = &c->peer[eb::PEER_SLOT_W].graveyard;
  return 0;
}

// ------------------------------------------------------------------------------------------ in-process communicator
// `world` host threads of ONE process, one context (GPU) each: what a C caller such as smartpca.c needs to use every GPU of
// the box without MPI or torch.  The callbacks are plain C functions over a pthread barrier and a shared staging buffer;
// peer_exchange sees equal pids and uses the peers' device pointers directly (cudaDeviceEnablePeerAccess), no IPC handles.
#include <pthread.h>
struct eb_local_comm {
  int world;
  pthread_barrier_t bar;
  std::vector<unsigned char> stage;
  struct Rank { eb_local_comm* lc; int rank; };
  std::vector<Rank> ranks;
};
static int local_barrier(void* user) {
  auto* r = (eb_local_comm::Rank*)user;
  const int rc = pthread_barrier_wait(&r->lc->bar);
  return (rc == 0 || rc == PTHREAD_BARRIER_SERIAL_THREAD) ? 0 : 1;
}
static int local_allgather(void* user, const void* src, void* dst, int64_t bytes) {
  auto* r = (eb_local_comm::Rank*)user;
  eb_local_comm* lc = r->lc;
  if (local_barrier(user)) return 1;                                   // the previous collective has been read by everyone
  if (r->rank == 0 && lc->stage.size() < (size_t)bytes * lc->world) lc->stage.resize((size_t)bytes * lc->world);
  if (local_barrier(user)) return 1;
  memcpy(lc->stage.data() + (size_t)r->rank * bytes, src, (size_t)bytes);
  if (local_barrier(user)) return 1;
  memcpy(dst, lc->stage.data(), (size_t)bytes * lc->world);
  return local_barrier(user);
}
extern "C" eb_local_comm* eb_local_comm_create(int world) {
  if (world < 1 || world > eb::EB_MAX_WORLD) { eb::set_error("eb_local_comm_create: world must be in 1..%d", eb::EB_MAX_WORLD); return nullptr; }
  eb_local_comm* lc = new eb_local_comm();
  lc->world = world;
  if (pthread_barrier_init(&lc->bar, nullptr, (unsigned)world) != 0) { delete lc; eb::set_error("eb_local_comm_create: pthread_barrier_init failed"); return nullptr; }
  lc->ranks.resize(world);
  for (int r = 0; r < world; r++) lc->ranks[r] = {lc, r};
  return lc;
}
extern "C" int eb_local_comm_get(eb_local_comm* lc, int rank, eb_comm* out) {
  if (!lc || !out || rank < 0 || rank >= lc->world) { eb::set_error("eb_local_comm_get: bad argument"); return EB_ERR_ARG; }
  out->rank = rank; out->world = lc->world; out->allgather_host = local_allgather; out->barrier = local_barrier; out->user = &lc->ranks[rank];
  return 0;
}
extern "C" void eb_local_comm_destroy(eb_local_comm* lc) {
  if (!lc) return;
  pthread_barrier_destroy(&lc->bar);
  delete lc;
}

extern "C" int eb_peer_allreduce_test(eb_ctx* c, double* host_io, int64_t count) {
  // testing aid: all-reduce a host vector through the peer kernel (upload, exchange, download)
  if (!c || !c->has_comm) { eb::set_error("eb_peer_allreduce_test: no communicator set"); return EB_ERR_STATE; }
  EB_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->peer_scratch.ensure((size_t)count))) return rc;
  EB_CUDA(cudaMemcpyAsync(c->peer_scratch.p, host_io, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
  if ((rc = eb::peer_allreduce(c, eb::PEER_SLOT_T, c->peer_scratch.p, c->peer_scratch.n, count))) return rc;
  EB_CUDA(cudaMemcpyAsync(host_io, c->peer_scratch.p, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return eb::peer_bury(c);
}
