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

__device__ __forceinline__ void tri_decode32(int t, int& ti, int& tj) {
  int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= t) r++;
  while (r * (r + 1) / 2 > t) r--;
  ti = r; tj = t - r * (r + 1) / 2;
}

struct GrmPeerArgs {
  const double* part[EB_MAX_WORLD];
  int nsplit[EB_MAX_WORLD];
  int world, rank;
};

// Phase 1 (reduce): CTA b of rank r finalises lower-triangle 32x32 block r + b*world: sum of every rank's planes in a fixed
// order, stored (with its mirror image) into the LOCAL xtx and, for the peers to pull, into the local plane 0 in place.
__global__ void __launch_bounds__(256) grm_peer_reduce_kernel(const GrmPeerArgs a, int npad, int nblocks, double* __restrict__ own_plane0,
                                                              double* __restrict__ xtx) {
  __shared__ double tile[32][33];
  const int gb = a.rank + blockIdx.x * a.world;
  if (gb >= nblocks) return;
  int bi, bj; tri_decode32(gb, bi, bj);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const size_t plane = (size_t)npad * npad;
  for (int r = ty; r < 32; r += 8) {
    const size_t idx = (size_t)(bi * 32 + r) * npad + bj * 32 + tx;
    double v = 0.0;
    for (int w = 0; w < a.world; w++) {
      const double* p = a.part[w] + idx;
      for (int s = 0; s < a.nsplit[w]; s++) v += p[s * plane];
    }
    own_plane0[idx] = v;                                     // complete block (the pad / upper part of diagonal blocks included)
    if (bi == bj && tx > r) v = 0.0;
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    double v = tile[r][tx];
    if (bi == bj && tx > r) v = tile[tx][r];
    xtx[(size_t)(bi * 32 + r) * npad + bj * 32 + tx] = v;
    if (bi != bj) xtx[(size_t)(bj * 32 + r) * npad + bi * 32 + tx] = tile[tx][r];
  }
}

// Phase 2 (gather): every block owned by another rank is pulled from that rank's plane 0 into the local xtx (+ mirror).
// CTA b handles the b-th block that is NOT owned by this rank.
__global__ void __launch_bounds__(256) grm_peer_gather_kernel(const GrmPeerArgs a, int npad, int nblocks, double* __restrict__ xtx) {
  __shared__ double tile[32][33];
  // b-th non-owned block: blocks are owned round-robin, so within each group of `world` consecutive blocks exactly one is ours
  const int grp = blockIdx.x / (a.world - 1), k = blockIdx.x % (a.world - 1);
  const int gb = grp * a.world + (k < a.rank ? k : k + 1);
  if (gb >= nblocks) return;
  const int owner = gb % a.world;
  int bi, bj; tri_decode32(gb, bi, bj);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const double* src = a.part[owner];
  for (int r = ty; r < 32; r += 8) {
    double v = src[(size_t)(bi * 32 + r) * npad + bj * 32 + tx];
    if (bi == bj && tx > r) v = 0.0;
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    double v = tile[r][tx];
    if (bi == bj && tx > r) v = tile[tx][r];
    xtx[(size_t)(bi * 32 + r) * npad + bj * 32 + tx] = v;
    if (bi != bj) xtx[(size_t)(bj * 32 + r) * npad + bi * 32 + tx] = tile[tx][r];
  }
}

// Reduce the per-rank split-K planes (left by grm_syrk_kernel in c->partial) across ranks into every rank's c->xtx.
// Pull only: a rank maps nothing but the peers' plane buffers (one cudaIpcOpenMemHandle per peer; measured 0.1-0.2 s per
// 20 GB buffer the first time, and the open waits for running kernels, so it cannot be hidden behind the SYRK kernel).
// peer_grm_prepare: publish / map the plane buffers (later passes reuse the mappings).
// peer_grm_finalize: barrier, reduce kernel (owned blocks), barrier, gather kernel (everybody else's blocks), barrier.
int peer_grm_prepare(eb_ctx* c) {
  const int W = c->comm.world;
  int rc;
  if (c->npad > (1 << 24)) { set_error("multi-GPU GRM: matrix too large for the exchange record"); return EB_ERR_ARG; }
  if ((rc = peer_exchange(c, PEER_SLOT_PARTIAL, c->partial.p, c->partial.n * sizeof(double), c->nsplit + 64 * c->npad))) return rc;
  const size_t plane = (size_t)c->npad * c->npad;
  for (int r = 0; r < W; r++) {
    const PeerRecord& pr = c->peer[PEER_SLOT_PARTIAL].rec[r];
    const int ns = pr.aux & 63, np = pr.aux >> 6;
    if (np != c->npad || ns < 1 || (size_t)ns * plane * sizeof(double) > pr.bytes) {
      set_error("multi-GPU GRM: rank %d has a different matrix size (npad %d vs %d): every shard must use the same rows", r, np, c->npad);
      return EB_ERR_STATE;
    }
  }
  return 0;
}

int peer_grm_finalize(eb_ctx* c) {
  const int W = c->comm.world;
  int rc;
  GrmPeerArgs a;
  memset(&a, 0, sizeof(a));
  a.world = W; a.rank = c->comm.rank;
  for (int r = 0; r < W; r++) {
    a.part[r] = (const double*)c->peer[PEER_SLOT_PARTIAL].mapped[r];
    a.nsplit[r] = c->peer[PEER_SLOT_PARTIAL].rec[r].aux & 63;
    if (!a.part[r]) { set_error("peer_grm_finalize: planes of rank %d are not mapped (peer_grm_prepare not called)", r); return EB_ERR_STATE; }
  }
  const bool dbg = getenv("EB_DEBUG") != nullptr;
  const auto tnow = [] { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double tq0 = tnow();
  EB_CUDA(cudaStreamSynchronize(c->stream));     // my planes are complete ...
  if ((rc = comm_barrier(c))) return rc;         // ... and so are everybody else's
  const double tq1 = tnow();
  const int T32 = c->npad / 32, nblocks = T32 * (T32 + 1) / 2;
  const int mine = (nblocks - a.rank + W - 1) / W;
  EB_CUDA(cudaEventRecord(c->ev[3], c->stream));
  if (mine > 0) {
    grm_peer_reduce_kernel<<<mine, 256, 0, c->stream>>>(a, c->npad, nblocks, c->partial.p, c->xtx.p);
    EB_CHECK_LAUNCH(c);
  }
  EB_CUDA(cudaStreamSynchronize(c->stream));     // my reduced blocks are in my plane 0 ...
  if ((rc = comm_barrier(c))) return rc;         // ... and everybody else's in theirs
  const int groups = (nblocks + W - 1) / W;
  if (W > 1 && groups > 0) {
    grm_peer_gather_kernel<<<groups * (W - 1), 256, 0, c->stream>>>(a, c->npad, nblocks, c->xtx.p);
    EB_CHECK_LAUNCH(c);
  }
  EB_CUDA(cudaEventRecord(c->ev[4], c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  if (dbg) fprintf(stderr, "[peer_grm_finalize] rank %d: wait for all ranks %.3f s, reduce + gather %.3f s\n", a.rank, tq1 - tq0, tnow() - tq1);
  return peer_bury(c);                            // barrier: nobody overwrites its planes while a peer still pulls from them
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
  eb::peer_release(c);
  c->partial.defer = c->fpG.defer = c->fpB.defer = c->fpS.defer = c->peer_scratch.defer = nullptr;
  for (auto& reg : c->peer) { for (void* p : reg.graveyard) cudaFree(p); reg.graveyard.clear(); }
  if (!comm || comm->world <= 1) { c->has_comm = false; memset(&c->comm, 0, sizeof(c->comm)); c->comm.world = 1; return 0; }
  if (comm->world > eb::EB_MAX_WORLD || comm->rank < 0 || comm->rank >= comm->world || !comm->allgather_host || !comm->barrier) {
    eb::set_error("eb_set_comm: need 0 <= rank < world <= %d and both callbacks", eb::EB_MAX_WORLD);
    return EB_ERR_ARG;
  }
  c->comm = *comm; c->has_comm = true;
  c->partial.defer = &c->peer[eb::PEER_SLOT_PARTIAL].graveyard;
  c->fpG.defer = &c->peer[eb::PEER_SLOT_A].graveyard;
  c->fpB.defer = &c->peer[eb::PEER_SLOT_B].graveyard;
  c->fpS.defer = &c->peer[eb::PEER_SLOT_C].graveyard;
  c->peer_scratch.defer = &c->peer[eb::PEER_SLOT_T].graveyard;
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
