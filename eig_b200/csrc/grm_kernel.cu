// grm_kernel.cu -- K1: symmetric rank-M update XTX = sum_s x_s x_s^T straight from 2-bit packed genotypes.
//
// Replaces domult_increment_lookup + block_increment_binary (smartpca.c:3426-3495, 3361-3423) and symit2
// (smartpca.c:480-508).  x_is = table[s][code(i,s)], table = {cc0,cc1,cc2,0} built by snp_stats_kernel.
//
// Design (sm_100a):
//   * FP64 tensor path = mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); tcgen05 has no f64 kind.
//   * CTA = 128x128 output tile, 8 warps (2x4, 64x32 each, 64 FP64 accumulators/thread); lane 0 of warp 0 doubles
//     as the TMA producer.  Lower-triangle tiles only, walked in bands of 12 tile rows, column by column (tile_order.cuh), so that
//     the 148 concurrent tiles share ~12 row operands and ~12 column operands in L2; optional split over SNP chunks into separate partial
//     buffers (deterministic, no atomics) so that small N still fills 148 SMs for many waves.
//   * Per stage the producer TMA-loads KT=128 SNPs x 32 bytes for the row tile and the column tile
//     (cp.async.bulk.tensor.2d) and the 128x4 FP64 decode table (cp.async.bulk), mbarrier full/empty ring.
//   * The decode of k-step i+1 is issued before the 32 DMMAs of k-step i (explicit software pipeline, LDS only).
//   * Operands never exist in memory as FP64: each thread pulls the 16 (A) / 8 (B) packed bytes that hold
//     its 8 / 4 fragment elements of SNP k, extracts the 2-bit codes with shifts and reads the FP64 value
//     from the per-SNP 4-entry table in shared memory (bank-conflict free: 4 SNPs x 4 entries x 8 B = 128 B).
//   * Algorithmic work: N*(N+1)*M flops ~ N^2 M; bytes are negligible (2 bits / element) => FP64-pipe bound.
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "tile_order.cuh"

namespace eb {

constexpr int STAGES = 4;
constexpr int GRM_THREADS = 256;
constexpr int STAGE_A = KT * 32;          // bytes
constexpr int STAGE_T = KT * 4 * 8;       // bytes
constexpr int STAGE_BYTES = 2 * STAGE_A + STAGE_T;
constexpr int GRM_SMEM = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE_%=;\n"
      "bra LAB_WAIT_%=;\n"
      "LAB_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}

// tile pair index t -> (ti >= tj)
__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
  int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= t) r++;
  while (r * (r + 1) / 2 > t) r--;
  ti = r; tj = t - r * (r + 1) / 2;
}

// WARPS_M x WARPS_N consumer warps, each TB x UB m8n8 blocks (tile = 128 x 128).  There is no dedicated producer warp:
// lane 0 of warp 0 keeps the TMA ring PREFETCH stages ahead, which leaves the full 64K registers to 8 warps (the
// software-pipelined decode needs ~240/thread).  m8 row blocks that lie entirely in the pad region (rows >= nrows)
// are skipped with a warp-uniform predicate.
template <int WARPS_M, int WARPS_N, int TB, int UB, int UNROLL>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, 1)
grm_syrk_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ table, double* __restrict__ partial,
                   int npad, int nrows, int ntiles_tri, int nsplit, int nkblocks, unsigned long long* __restrict__ prof,
                   const __grid_constant__ GrmPush push) {
  static_assert(WARPS_M * TB * 8 == TILE && WARPS_N * UB * 8 == TILE, "tile must be 128 x 128");
  static_assert(TB == 4 || TB == 8, "A segment is 8 or 16 bytes");
  static_assert(UB == 4 || UB == 8, "B segment is 8 or 16 bytes");
  constexpr int NWARPS = WARPS_M * WARPS_N;
  constexpr int PREFETCH = STAGES - 2;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, NWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // self-measurement (4 words per CTA): which SM, wall-clock span (globaltimer, ns) and SM cycles of this CTA, so that the host
  // can report the EFFECTIVE SM clock of the launch and the number of SMs it really ran on (bench.py: roofline.kernel_clock)
  unsigned long long prof_t0 = 0, prof_c0 = 0;
  if (threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_t0));
    prof_c0 = (unsigned long long)clock64();
  }

  const int nitems = ntiles_tri * nsplit;
  // producer cursor (meaningful in thread 0 only)
  int p_item = blockIdx.x, p_kb = 0, p_kb1 = 0, p_ti = 0, p_tj = 0;
  uint32_t p_stage = 0, p_phase = 0;
  bool p_open = false;          // p_item decoded
  int ahead = 0;                // blocks issued but not yet consumed
  auto produce = [&]() {
    // keep the ring PREFETCH blocks ahead of the consumer
    while (ahead < PREFETCH) {
      if (!p_open) {
        if (p_item >= nitems) return;
        const int chunk = p_item / ntiles_tri, t = p_item - chunk * ntiles_tri;
        tile_decode_banded(t, npad / TILE, p_ti, p_tj);
        p_kb = (int)(((long long)nkblocks * chunk) / nsplit);
        p_kb1 = (int)(((long long)nkblocks * (chunk + 1)) / nsplit);
        if (p_kb >= p_kb1) { p_item += gridDim.x; continue; }      // empty chunk (fewer SNP blocks than chunks): nothing to load
        p_open = true;
      }
      mbar_wait(empty + p_stage, p_phase ^ 1);
      uint8_t* sb = smem + p_stage * STAGE_BYTES;
      mbar_expect_tx(full + p_stage, STAGE_BYTES);
      tma_load_2d(sb, &tmap, p_ti * 32, p_kb * KT, full + p_stage);
      tma_load_2d(sb + STAGE_A, &tmap, p_tj * 32, p_kb * KT, full + p_stage);
      bulk_load_1d(sb + 2 * STAGE_A, table + (size_t)p_kb * KT * 4, STAGE_T, full + p_stage);
      if (++p_stage == STAGES) { p_stage = 0; p_phase ^= 1; }
      ahead++;
      if (++p_kb == p_kb1) { p_open = false; p_item += gridDim.x; }
    }
  };

  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, q = lane & 3, h = lane >> 4;
  const uint32_t sh = ((3 - (g & 3)) << 1) + (h << 3);

  double acc[TB][UB][2];
#pragma unroll
  for (int t = 0; t < TB; t++)
#pragma unroll
    for (int u = 0; u < UB; u++) acc[t][u][0] = acc[t][u][1] = 0.0;

  uint32_t stage = 0, phase = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int chunk = item / ntiles_tri, t_ = item - chunk * ntiles_tri;
    int ti, tj; tile_decode_banded(t_, npad / TILE, ti, tj);      // L2-friendly walk of the triangle (tile_order.cuh)
    const int kb0 = (int)(((long long)nkblocks * chunk) / nsplit), kb1 = (int)(((long long)nkblocks * (chunk + 1)) / nsplit);
    // valid m8 row blocks of this warp (rows beyond nrows are padding and decode to zero anyway)
    int tvalid = (nrows - (ti * TILE + wm * TB * 8) + 7) >> 3;
    tvalid = tvalid < 0 ? 0 : (tvalid > TB ? TB : tvalid);
    for (int kb = kb0; kb < kb1; kb++) {
      if (threadIdx.x == 0) produce();
      __syncwarp();
      mbar_wait(full + stage, phase);
      // 32-bit shared-space addresses (explicit LDS; the table row base is 32-byte aligned so OR replaces ADD)
      const uint32_t sb = smem_u32(smem + stage * STAGE_BYTES);
      const uint32_t pa = sb + wm * (TB * 2) + q * 32;
      const uint32_t pb = sb + STAGE_A + wn * (UB * 2) + q * 32;
      const uint32_t pt = sb + 2 * STAGE_A + q * 32;
      // decode one k-step (4 SNPs): packed bytes -> 2-bit codes -> FP64 values from the per-SNP table
      auto decode = [&](int kk, double (&a)[TB], double (&b)[UB]) {
        const uint32_t tk = pt + kk * 32;
        uint32_t wa[TB / 2], wb[UB / 2];
        if (TB == 8) { const uint4 w = lds_v4(pa + kk * 32); wa[0] = w.x; wa[1] = w.y; wa[TB / 2 - 2] = w.z; wa[TB / 2 - 1] = w.w; }
        else { const uint2 w = lds_v2(pa + kk * 32); wa[0] = w.x; wa[1] = w.y; }
        if (UB == 8) { const uint4 w = lds_v4(pb + kk * 32); wb[0] = w.x; wb[1] = w.y; wb[UB / 2 - 2] = w.z; wb[UB / 2 - 1] = w.w; }
        else { const uint2 w = lds_v2(pb + kk * 32); wb[0] = w.x; wb[1] = w.y; }
#pragma unroll
        for (int i = 0; i < TB / 2; i++) {
          const uint32_t v = wa[i] >> sh;
          a[2 * i] = lds_f64(tk | ((v << 3) & 0x18));
          a[2 * i + 1] = lds_f64(tk | ((v >> 13) & 0x18));
        }
#pragma unroll
        for (int i = 0; i < UB / 2; i++) {
          const uint32_t v = wb[i] >> sh;
          b[2 * i] = lds_f64(tk | ((v << 3) & 0x18));
          b[2 * i + 1] = lds_f64(tk | ((v >> 13) & 0x18));
        }
      };
      if (tvalid == TB) {
        if (UNROLL == 2) {
          // software pipeline: the operands of k-step i+1 are decoded before the 32 DMMAs of k-step i are issued, so
          // the two warps of an SMSP never both sit in a decode bubble while the DMMA pipe drains
          double a0[TB], b0[UB], a1[TB], b1[UB];
          decode(0, a0, b0);
#pragma unroll 1
          for (int kk = 0; kk < KT; kk += 8) {
            decode(kk + 4, a1, b1);
#pragma unroll
            for (int t = 0; t < TB; t++)
#pragma unroll
              for (int u = 0; u < UB; u++) dmma884(acc[t][u][0], acc[t][u][1], a0[t], b0[u]);
            if (kk + 8 < KT) decode(kk + 8, a0, b0);
#pragma unroll
            for (int t = 0; t < TB; t++)
#pragma unroll
              for (int u = 0; u < UB; u++) dmma884(acc[t][u][0], acc[t][u][1], a1[t], b1[u]);
          }
        } else {
#pragma unroll 1
          for (int kk = 0; kk < KT; kk += 4) {
            double a[TB], b[UB];
            decode(kk, a, b);
#pragma unroll
            for (int t = 0; t < TB; t++)
#pragma unroll
              for (int u = 0; u < UB; u++) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
          }
        }
      } else if (tvalid > 0) {
#pragma unroll 1
        for (int kk = 0; kk < KT; kk += 4) {
          double a[TB], b[UB];
          decode(kk, a, b);
#pragma unroll
          for (int t = 0; t < TB; t++) {
            if (t < tvalid) {
#pragma unroll
              for (int u = 0; u < UB; u++) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
            }
          }
        }
      }
      // release the stage: every lane fences its own shared-memory reads, the warp converges, lane 0 arrives (the
      // arrive alone does not order the still outstanding operand loads of the last k-step -- see dmma_tile.cuh)
      asm volatile("fence.acq_rel.cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);
      if (threadIdx.x == 0) ahead--;
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // single GPU: the tile goes to its place in split-K plane `chunk`.  SNPs sharded over several GPUs: it goes straight into the
    // receive buffer of the rank that owns tile t_ (plain stores over NVLink when that is a peer), see GrmPush
    double* out;
    size_t ldo;
    if (push.world > 1) {
      const int owner = t_ % push.world;
      out = push.recv[owner] + ((size_t)(t_ / push.world) * (size_t)(push.world * nsplit) + (size_t)(push.rank * nsplit + chunk)) * (TILE * TILE);
      ldo = TILE;
    } else {
      out = partial + (size_t)chunk * npad * npad + (size_t)ti * TILE * npad + (size_t)tj * TILE;
      ldo = (size_t)npad;
    }
#pragma unroll
    for (int t = 0; t < TB; t++) {
      const size_t row = (size_t)(wm * (TB * 8) + t * 8 + g);
#pragma unroll
      for (int u = 0; u < UB; u++) {
        const size_t col = (size_t)(wn * (UB * 8) + u * 8 + q * 2);
        *reinterpret_cast<double2*>(out + row * ldo + col) = make_double2(acc[t][u][0], acc[t][u][1]);
        acc[t][u][0] = acc[t][u][1] = 0.0;
      }
    }
  }
  if (threadIdx.x == 0 && prof) {
    unsigned long long t1; unsigned int smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned long long c1 = (unsigned long long)clock64();
    prof[4 * blockIdx.x + 0] = smid; prof[4 * blockIdx.x + 1] = prof_t0; prof[4 * blockIdx.x + 2] = t1; prof[4 * blockIdx.x + 3] = c1 - prof_c0;
  }
}

// Sum the split partials in a fixed order, mirror the lower triangle (replaces symit2) -> full symmetric xtx.
__global__ void __launch_bounds__(256) grm_finalize_kernel(const double* __restrict__ partial, int nsplit, int npad,
                                                           double* __restrict__ xtx) {
  __shared__ double tile[32][33];
  // blockIdx.x enumerates lower-triangle 32x32 blocks
  int bi, bj; tri_decode(blockIdx.x, bi, bj);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const size_t plane = (size_t)npad * npad;
  for (int r = ty; r < 32; r += 8) {
    const size_t idx = (size_t)(bi * 32 + r) * npad + bj * 32 + tx;
    double v = 0.0;
    for (int c = 0; c < nsplit; c++) v += partial[c * plane + idx];
    if (bi == bj && tx > r) v = 0.0;          // upper part of a diagonal block comes from the mirror below
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

// trace over the first n diagonal entries, fixed-order tree => deterministic
__global__ void __launch_bounds__(1024) trace_kernel(const double* __restrict__ xtx, int n, int ld, double* __restrict__ out) {
  __shared__ double s[1024];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) v += xtx[(size_t)i * ld + i];
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_work_tensormap(eb_ctx* c, CUtensorMap* map) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return EB_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)c->wpitch, (cuuint64_t)c->mpad};
  cuuint64_t strides[1] = {(cuuint64_t)c->wpitch};
  cuuint32_t box[2] = {32, (cuuint32_t)KT};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)c->work.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (pitch %lld rows %lld)", (int)r, (long long)c->wpitch, (long long)c->mpad); return EB_ERR_CUDA; }
  return 0;
}

int grm_trace(eb_ctx* c) {
  trace_kernel<<<1, 1024, 0, c->stream>>>(c->xtx.p, c->nrows, c->npad, c->trace_d.p);
  EB_CHECK_LAUNCH(c);
  double tr = 0.0;
  EB_CUDA(cudaMemcpyAsync(&tr, c->trace_d.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  c->y = tr / (double)(c->nrows - 1);
  return 0;
}

// split-K factor: enough (tile, chunk) items for ~40 waves over the SMs.  Sharded: a function of the matrix size alone, so that every
// rank derives the same receive-buffer geometry without talking (a chunk may then be empty on a rank with few SNP blocks: it
// stores a zero tile).  Single GPU: also bounded by the number of SNP blocks and by free memory.
int grm_nsplit_for(const eb_ctx* c, bool sharded) {
  const int T = c->npad / TILE;
  const int ntri = T * (T + 1) / 2;
  int nsplit = (40 * c->num_sms + ntri - 1) / ntri;
  if (sharded) return grm_use_i8(c) ? 1 : std::max(1, std::min(nsplit, 16));
  const int nkb = (int)(c->mpad / KT);
  nsplit = std::max(1, std::min(nsplit, std::min(nkb, 32)));
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  const size_t plane = (size_t)c->npad * c->npad * sizeof(double);
  const size_t have = c->partial.n * sizeof(double);
  while (nsplit > 1 && (size_t)nsplit * plane > have + freeb / 2) nsplit--;
  return nsplit;
}

int grm_accumulate(eb_ctx* c, bool finalize_local, bool push_mode) {
  if (grm_use_i8(c)) { c->tm.grm_method = 2; return grm_accumulate_i8(c, finalize_local, push_mode); }
  c->tm.grm_method = 1;
  const int T = c->npad / TILE;
  const int ntri = T * (T + 1) / 2;
  const int nkb = (int)(c->mpad / KT);
  const int nsplit = push_mode ? c->grm_geom_nsplit : grm_nsplit_for(c, false);
  c->nsplit = nsplit;
  int rc;
  GrmPush push;
  memset(&push, 0, sizeof(push));
  push.world = 1;
  if (push_mode) { if ((rc = peer_grm_push_args(c, &push))) return rc; }
  else if ((rc = c->partial.ensure((size_t)nsplit * c->npad * c->npad))) return rc;
  if ((rc = c->xtx.ensure((size_t)c->npad * c->npad))) return rc;
  if ((rc = c->trace_d.ensure(1))) return rc;

  CUtensorMap map;
  if ((rc = make_work_tensormap(c, &map))) return rc;
  const char* ev = getenv("EB_GRM_VARIANT");            // 1 = non-pipelined decode (debug / A-B only)
  const int variant = ev ? atoi(ev) : 2;
  // per device (not per process): several contexts on different GPUs may live in one process (eb_local_comm)
  EB_CUDA(cudaFuncSetAttribute(grm_syrk_kernel<2, 4, 8, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRM_SMEM));
  EB_CUDA(cudaFuncSetAttribute(grm_syrk_kernel<2, 4, 8, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRM_SMEM));
  const int nitems = ntri * nsplit;
  const int grid = std::min(nitems, c->num_sms);
  if ((rc = c->grmprof_d.ensure((size_t)4 * c->num_sms))) return rc;
  c->grm_grid = grid;
  EB_CUDA(cudaEventRecord(c->ev[2], c->stream));
  if (variant == 1)
    grm_syrk_kernel<2, 4, 8, 4, 1><<<grid, GRM_THREADS, GRM_SMEM, c->stream>>>(map, c->table_d.p, c->partial.p, c->npad, c->nrows, ntri, nsplit, nkb, c->grmprof_d.p, push);
  else
    grm_syrk_kernel<2, 4, 8, 4, 2><<<grid, GRM_THREADS, GRM_SMEM, c->stream>>>(map, c->table_d.p, c->partial.p, c->npad, c->nrows, ntri, nsplit, nkb, c->grmprof_d.p, push);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[3], c->stream));
  c->tm.grm_launches = 2;
  if (!finalize_local) return 0;       // multi-GPU: peer_grm_finalize() sums the planes of every rank (peer.cu)
  const int T32 = c->npad / 32;
  grm_finalize_kernel<<<T32 * (T32 + 1) / 2, 256, 0, c->stream>>>(c->partial.p, nsplit, c->npad, c->xtx.p);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[4], c->stream));
  return 0;
}

// dense path: lower-tile accumulator (c->partial, one plane) -> full symmetric xtx
int grm_dense_finalize(eb_ctx* c) {
  const int T32 = c->npad / 32;
  grm_finalize_kernel<<<T32 * (T32 + 1) / 2, 256, 0, c->stream>>>(c->partial.p, 1, c->npad, c->xtx.p);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// ---------------------------------------------------------------------------------------------- FP64 microbenchmarks
// DMMA issue-rate probe with the instruction stream of the real kernel minus the decode: 8 warps per SM, 8 x 4 m8n8 blocks per warp
// (64 accumulators per thread, 32 DMMAs per k-step, 8 distinct A and 4 distinct B operands).  (Round 1's probe issued 16 DMMAs with ONE
// A and ONE B register from 32 warps per SM and read 29.7 TFLOP/s on most leases -- below what grm_syrk_kernel sustains -- so it
// measured that artificial stream, not the pipe.)
__global__ void __launch_bounds__(256, 1) dmma_bench_kernel(double* out, int iters) {
  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; t++)
#pragma unroll
    for (int u = 0; u < 4; u++) { acc[t][u][0] = 0.0; acc[t][u][1] = 0.0; }
  double a[8], b[4];
#pragma unroll
  for (int t = 0; t < 8; t++) a[t] = 1.0 + (threadIdx.x + t) * 1e-9;
#pragma unroll
  for (int u = 0; u < 4; u++) b[u] = 1.0 - (threadIdx.x + u) * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
      for (int u = 0; u < 4; u++) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 8; t++)
#pragma unroll
    for (int u = 0; u < 4; u++) s += acc[t][u][0] + acc[t][u][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dfma_bench_kernel(double* out, int iters) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = i;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int microbench_fp64(eb_ctx* c, double* dmma, double* dfma) {
  DevBuf<double> out;
  const int blocks = c->num_sms * 4, threads = 256, iters = 20000;
  const int dblocks = c->num_sms, diters = 40000;      // DMMA probe: one CTA of 8 warps per SM, like grm_syrk_kernel
  int rc;
  if ((rc = out.ensure((size_t)blocks * threads))) return rc;
  cudaEvent_t e0, e1;
  EB_CUDA(cudaEventCreate(&e0)); EB_CUDA(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    EB_CUDA(cudaEventRecord(e0, c->stream));
    dmma_bench_kernel<<<dblocks, threads, 0, c->stream>>>(out.p, diters);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaEventRecord(e1, c->stream));
    EB_CUDA(cudaEventSynchronize(e1));
    EB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  }
  // per warp per iteration: 32 DMMA x (8*8*4 FMA) x 2 flops
  *dmma = (double)dblocks * (threads / 32) * (double)diters * 32.0 * 512.0 / (ms * 1e-3) / 1e12;
  for (int rep = 0; rep < 2; rep++) {
    EB_CUDA(cudaEventRecord(e0, c->stream));
    dfma_bench_kernel<<<blocks, threads, 0, c->stream>>>(out.p, iters);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaEventRecord(e1, c->stream));
    EB_CUDA(cudaEventSynchronize(e1));
    EB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  }
  *dfma = (double)blocks * threads * (double)iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

}  // namespace eb
