// grm_kernel.cu -- K1: symmetric rank-M update XTX = sum_s x_s x_s^T straight from 2-bit packed genotypes.
//
// Replaces domult_increment_lookup + block_increment_binary (smartpca.c:3426-3495, 3361-3423) and symit2
// (smartpca.c:480-508).  x_is = table[s][code(i,s)], table = {cc0,cc1,cc2,0} built by snp_stats_kernel.
//
// Design (sm_100a):
//   * FP64 tensor path = mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); tcgen05 has no f64 kind.
//   * CTA = 128x128 output tile, 8 consumer warps (2x4, 64x32 each, 64 FP64 accumulators/thread) + 1 TMA
//     producer warp.  Lower-triangle tiles only; optional split over SNP chunks into separate partial
//     buffers (deterministic, no atomics) so that small N still fills 148 SMs for many waves.
//   * Per stage the producer TMA-loads KT=128 SNPs x 32 bytes for the row tile and the column tile
//     (cp.async.bulk.tensor.2d) and the 128x4 FP64 decode table (cp.async.bulk), mbarrier full/empty ring.
//   * Operands never exist in memory as FP64: each thread pulls the 16 (A) / 8 (B) packed bytes that hold
//     its 8 / 4 fragment elements of SNP k, extracts the 2-bit codes with shifts and reads the FP64 value
//     from the per-SNP 4-entry table in shared memory (bank-conflict free: 4 SNPs x 4 entries x 8 B = 128 B).
//   * Algorithmic work: N*(N+1)*M flops ~ N^2 M; bytes are negligible (2 bits / element) => FP64-pipe bound.
#include <algorithm>
#include "common.cuh"

namespace eb {

constexpr int STAGES = 4;
constexpr int CONSUMER_WARPS = 8;
constexpr int GRM_THREADS = (CONSUMER_WARPS + 1) * 32;
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

// tile pair index t -> (ti >= tj)
__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
  int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= t) r++;
  while (r * (r + 1) / 2 > t) r--;
  ti = r; tj = t - r * (r + 1) / 2;
}

__global__ void __launch_bounds__(GRM_THREADS, 1)
grm_syrk_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ table, double* __restrict__ partial,
                int npad, int ntiles_tri, int nsplit, int nkblocks) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nitems = ntiles_tri * nsplit;
  uint32_t stage = 0, phase = 0;

  if (warp == CONSUMER_WARPS) {
    // ===== TMA producer warp (one elected lane) =====
    if (lane == 0) {
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chunk = item / ntiles_tri, t = item - chunk * ntiles_tri;   // chunk-major: co-resident CTAs share one SNP range in L2
        int ti, tj; tri_decode(t, ti, tj);
        const int kb0 = (int)(((long long)nkblocks * chunk) / nsplit), kb1 = (int)(((long long)nkblocks * (chunk + 1)) / nsplit);
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sb = smem + stage * STAGE_BYTES;
          mbar_expect_tx(full + stage, STAGE_BYTES);
          tma_load_2d(sb, &tmap, ti * 32, kb * KT, full + stage);
          tma_load_2d(sb + STAGE_A, &tmap, tj * 32, kb * KT, full + stage);
          bulk_load_1d(sb + 2 * STAGE_A, table + (size_t)kb * KT * 4, STAGE_T, full + stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int wm = warp >> 2, wn = warp & 3;     // 2 x 4 warps, 64 x 32 each
  const int g = lane >> 2, q = lane & 3;       // fragment row / k index
  const int h = lane >> 4;                     // which byte of the pair holds this thread's individual
  const uint32_t sh = ((3 - (g & 3)) << 1) + (h << 3);   // bit position of element t=0 inside word 0

  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; t++)
#pragma unroll
    for (int u = 0; u < 4; u++) acc[t][u][0] = acc[t][u][1] = 0.0;

  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int chunk = item / ntiles_tri, t_ = item - chunk * ntiles_tri;
    int ti, tj; tri_decode(t_, ti, tj);
    const int kb0 = (int)(((long long)nkblocks * chunk) / nsplit), kb1 = (int)(((long long)nkblocks * (chunk + 1)) / nsplit);
    for (int kb = kb0; kb < kb1; kb++) {
      mbar_wait(full + stage, phase);
      const uint8_t* sb = smem + stage * STAGE_BYTES;
      const uint8_t* pa = sb + wm * 16 + q * 32;
      const uint8_t* pb = sb + STAGE_A + wn * 8 + q * 32;
      const uint8_t* pt = sb + 2 * STAGE_A + q * 32;
#pragma unroll 1
      for (int kk = 0; kk < KT; kk += 4) {
        const uint4 wa = *reinterpret_cast<const uint4*>(pa + kk * 32);
        const uint2 wb = *reinterpret_cast<const uint2*>(pb + kk * 32);
        const uint8_t* tk = pt + kk * 32;
        double a[8], b[4];
        {
          const uint32_t v0 = wa.x >> sh, v1 = wa.y >> sh, v2 = wa.z >> sh, v3 = wa.w >> sh;
          a[0] = *reinterpret_cast<const double*>(tk + ((v0 << 3) & 0x18));
          a[1] = *reinterpret_cast<const double*>(tk + ((v0 >> 13) & 0x18));
          a[2] = *reinterpret_cast<const double*>(tk + ((v1 << 3) & 0x18));
          a[3] = *reinterpret_cast<const double*>(tk + ((v1 >> 13) & 0x18));
          a[4] = *reinterpret_cast<const double*>(tk + ((v2 << 3) & 0x18));
          a[5] = *reinterpret_cast<const double*>(tk + ((v2 >> 13) & 0x18));
          a[6] = *reinterpret_cast<const double*>(tk + ((v3 << 3) & 0x18));
          a[7] = *reinterpret_cast<const double*>(tk + ((v3 >> 13) & 0x18));
          const uint32_t u0 = wb.x >> sh, u1 = wb.y >> sh;
          b[0] = *reinterpret_cast<const double*>(tk + ((u0 << 3) & 0x18));
          b[1] = *reinterpret_cast<const double*>(tk + ((u0 >> 13) & 0x18));
          b[2] = *reinterpret_cast<const double*>(tk + ((u1 << 3) & 0x18));
          b[3] = *reinterpret_cast<const double*>(tk + ((u1 >> 13) & 0x18));
        }
#pragma unroll
        for (int t = 0; t < 8; t++)
#pragma unroll
          for (int u = 0; u < 4; u++) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // epilogue: accumulators -> partial[chunk], lower tile (ti,tj)
    double* out = partial + (size_t)chunk * npad * npad;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const size_t row = (size_t)ti * TILE + wm * 64 + t * 8 + g;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const size_t col = (size_t)tj * TILE + wn * 32 + u * 8 + q * 2;
        *reinterpret_cast<double2*>(out + row * npad + col) = make_double2(acc[t][u][0], acc[t][u][1]);
        acc[t][u][0] = acc[t][u][1] = 0.0;
      }
    }
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

int grm_accumulate(eb_ctx* c) {
  const int T = c->npad / TILE;
  const int ntri = T * (T + 1) / 2;
  const int nkb = (int)(c->mpad / KT);
  // enough (tile, chunk) items for ~40 waves over the SMs, bounded by the number of SNP blocks and by memory
  int nsplit = (40 * c->num_sms + ntri - 1) / ntri;
  nsplit = std::max(1, std::min(nsplit, std::min(nkb, 32)));
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  const size_t plane = (size_t)c->npad * c->npad * sizeof(double);
  const size_t have = c->partial.n * sizeof(double);
  while (nsplit > 1 && (size_t)nsplit * plane > have + freeb / 2) nsplit--;
  c->nsplit = nsplit;
  int rc;
  if ((rc = c->partial.ensure((size_t)nsplit * c->npad * c->npad))) return rc;
  if ((rc = c->xtx.ensure((size_t)c->npad * c->npad))) return rc;
  if ((rc = c->trace_d.ensure(1))) return rc;

  CUtensorMap map;
  if ((rc = make_work_tensormap(c, &map))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    EB_CUDA(cudaFuncSetAttribute(grm_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRM_SMEM));
    attr_set = true;
  }
  const int nitems = ntri * nsplit;
  const int grid = std::min(nitems, c->num_sms);
  EB_CUDA(cudaEventRecord(c->ev[2], c->stream));
  grm_syrk_kernel<<<grid, GRM_THREADS, GRM_SMEM, c->stream>>>(map, c->table_d.p, c->partial.p, c->npad, ntri, nsplit, nkb);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[3], c->stream));
  const int T32 = c->npad / 32;
  grm_finalize_kernel<<<T32 * (T32 + 1) / 2, 256, 0, c->stream>>>(c->partial.p, nsplit, c->npad, c->xtx.p);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[4], c->stream));
  c->tm.grm_launches = 2;
  return 0;
}

// ---------------------------------------------------------------------------------------------- FP64 microbenchmarks
__global__ void __launch_bounds__(256) dmma_bench_kernel(double* out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
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
  int rc;
  if ((rc = out.ensure((size_t)blocks * threads))) return rc;
  cudaEvent_t e0, e1;
  EB_CUDA(cudaEventCreate(&e0)); EB_CUDA(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    EB_CUDA(cudaEventRecord(e0, c->stream));
    dmma_bench_kernel<<<blocks, threads, 0, c->stream>>>(out.p, iters);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaEventRecord(e1, c->stream));
    EB_CUDA(cudaEventSynchronize(e1));
    EB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  }
  // per warp per iteration: 16 DMMA x (8*8*4 FMA) x 2 flops
  *dmma = (double)blocks * (threads / 32) * (double)iters * 16.0 * 512.0 / (ms * 1e-3) / 1e12;
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
