// dmma_tile.cuh -- FP64 tensor-core (DMMA.8x8x4) tile machinery shared by the eigensolver kernels.
//
// One CTA = 4 consumer warps (2 x 2, each 64 x 32 of a 128 x 64 output tile, 64 FP64 accumulators per thread) plus one
// producer warp whose lane 0 drives a 2-deep TMA ring.  Two CTAs are resident per SM, so one CTA's epilogue (global
// read-modify-write) overlaps the other's DMMA stream.  Operand tiles are fetched with cp.async.bulk.tensor.2d using
// boxes that are 4 doubles WIDER than the tile: the surplus columns act as row padding, which makes the row stride
// == 4 (mod 16) doubles and every DMMA fragment load bank-conflict free without a swizzle.
#pragma once
#include "common.cuh"

namespace eb {

constexpr int DT_M = 128, DT_N = 64, DT_KC = 32;
constexpr int DT_LD_K = DT_KC + 4;     // 36 : tiles stored [row][k]
constexpr int DT_LD_M = DT_M + 4;      // 132: A tiles stored [k][row]
constexpr int DT_LD_N = DT_N + 4;      // 68 : B tiles stored [k][col]
constexpr int DT_A_BYTES = DT_M * DT_LD_K * 8;   // 36864 (>= 32*132*8 = 33792)
constexpr int DT_B_BYTES = DT_N * DT_LD_K * 8;   // 18432 (>= 32*68*8  = 17408)
constexpr int DT_A_KM_BYTES = DT_KC * DT_LD_M * 8;
constexpr int DT_B_KN_BYTES = DT_KC * DT_LD_N * 8;
constexpr int DT_STAGE_BYTES = DT_A_BYTES + DT_B_BYTES;
constexpr int DT_STAGES = 2;
constexpr int DT_SMEM = DT_STAGES * DT_STAGE_BYTES + 2 * DT_STAGES * 8 + 128;
constexpr int DT_CONSUMERS = 128;
constexpr int DT_THREADS = DT_CONSUMERS + 32;

__device__ __forceinline__ uint32_t dt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dt_smem_u32(bar)) : "memory");
}
// Release a pipeline stage after reading it: every lane fences its own shared-memory reads, the warp converges, lane 0
// arrives.  (mbarrier.arrive alone did not keep ptxas / the hardware from performing the arrive while the last operand
// loads of the stage were still queued behind DMMAs.)
__device__ __forceinline__ void dt_release_stage(uint64_t* bar, int lane) {
  asm volatile("fence.acq_rel.cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) dt_mbar_arrive(bar);
}
__device__ __forceinline__ void dt_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DT_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DT_DONE_%=;\n"
      "bra DT_WAIT_%=;\n"
      "DT_DONE_%=:\n"
      "}\n" ::"r"(dt_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void dt_tma_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   dt_smem_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(dt_smem_u32(bar))
               : "memory");
}
// volatile: the DMMAs keep their program order with respect to the operand loads (dt_lds) and to the mbarrier arrive
// that releases the stage.  A DMMA cannot issue before its source registers are back from shared memory, so an arrive
// placed after the last DMMA of a stage is ordered after the completion of every operand load of that stage.  (With
// movable DMMAs ptxas scheduled the arrive ahead of the last k-step's DMMAs; under a deep DMMA backlog the arrive then
// overtook the still queued loads by enough cycles for the producer's next TMA write to land first.)
__device__ __forceinline__ void dt_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Shared-memory operand load as an explicit, volatile LDS: it keeps its program order with respect to the mbarrier waits
// and arrives (also volatile asm).  A plain C++ load through a `const double* __restrict__` was treated by the compiler as
// invariant memory: it turned into generic LD.E instructions scheduled BELOW the warp-level reconvergence point and all
// the way down to the empty-barrier arrive, and persistent CTAs with short items (syr2k: 2-4 k-chunks per tile, 4+ tiles
// per CTA from n ~ 4000) then lost parts of operand tiles to the producer's next TMA write (observed as run-to-run
// different band matrices, eigenvalue errors 1e-5 .. O(1) for n >= 4096).
__device__ __forceinline__ double dt_lds(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// One DT_KC-deep stage of DMMAs.  A_KM: A tile stored [k][row] (ld 132) else [row][k] (ld 36);
// B_KN: B tile stored [k][col] (ld 68) else [col][k] (ld 36).  As / Bs are shared-space byte addresses.
// The fragments of k-step i+1 are loaded before the 32 DMMAs of k-step i are issued (register double buffer).
template <bool A_KM, bool B_KN>
__device__ __forceinline__ void dt_stage_mma(uint32_t As, uint32_t Bs, double (&acc)[8][4][2], int wm, int wn, int g, int q) {
  const uint32_t pa = As + 8u * (uint32_t)(A_KM ? q * DT_LD_M + wm * 64 + g : (wm * 64 + g) * DT_LD_K + q);
  const uint32_t pb = Bs + 8u * (uint32_t)(B_KN ? q * DT_LD_N + wn * 32 + g : (wn * 32 + g) * DT_LD_K + q);
  constexpr uint32_t A_K = 8u * (A_KM ? DT_LD_M : 1), A_T = 8u * (A_KM ? 8 : 8 * DT_LD_K);
  constexpr uint32_t B_K = 8u * (B_KN ? DT_LD_N : 1), B_U = 8u * (B_KN ? 8 : 8 * DT_LD_K);
  double a[2][8], b[2][4];
#pragma unroll
  for (int t = 0; t < 8; t++) a[0][t] = dt_lds(pa + t * A_T);
#pragma unroll
  for (int u = 0; u < 4; u++) b[0][u] = dt_lds(pb + u * B_U);
#pragma unroll
  for (int ks = 0; ks < DT_KC / 4; ks++) {
    const int cur = ks & 1, nxt = cur ^ 1;
    if (ks + 1 < DT_KC / 4) {
#pragma unroll
      for (int t = 0; t < 8; t++) a[nxt][t] = dt_lds(pa + (ks + 1) * 4 * A_K + t * A_T);
#pragma unroll
      for (int u = 0; u < 4; u++) b[nxt][u] = dt_lds(pb + (ks + 1) * 4 * B_K + u * B_U);
    }
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
      for (int u = 0; u < 4; u++) dt_dmma(acc[t][u][0], acc[t][u][1], a[cur][t], b[cur][u]);
  }
}

// All accumulators through one volatile asm with a memory clobber: every DMMA that feeds them has completed (and has
// therefore read its source registers) before anything after this point is issued.
__device__ __forceinline__ void dt_acc_fence(double (&acc)[8][4][2]) {
#pragma unroll
  for (int t = 0; t < 8; t++)
    asm volatile("" : "+d"(acc[t][0][0]), "+d"(acc[t][0][1]), "+d"(acc[t][1][0]), "+d"(acc[t][1][1]), "+d"(acc[t][2][0]), "+d"(acc[t][2][1]),
                      "+d"(acc[t][3][0]), "+d"(acc[t][3][1])
                 :
                 : "memory");
}

typedef CUresult (*PFN_encodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled_t get_tensormap_encoder();
// FP64 row-major matrix [rows][cols] with leading dimension ld (doubles); box = boxc x boxr elements, zero fill out of bounds
int make_f64_tensormap(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int boxc, int boxr);

}  // namespace eb
