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
__device__ __forceinline__ void dt_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One DT_KC-deep stage of DMMAs.  A_KM: A tile stored [k][row] (ld 132) else [row][k] (ld 36);
// B_KN: B tile stored [k][col] (ld 68) else [col][k] (ld 36).
template <bool A_KM, bool B_KN>
__device__ __forceinline__ void dt_stage_mma(const double* __restrict__ As, const double* __restrict__ Bs, double (&acc)[8][4][2],
                                             int wm, int wn, int g, int q) {
  const double* pa = A_KM ? As + q * DT_LD_M + wm * 64 + g : As + (wm * 64 + g) * DT_LD_K + q;
  const double* pb = B_KN ? Bs + q * DT_LD_N + wn * 32 + g : Bs + (wn * 32 + g) * DT_LD_K + q;
#pragma unroll
  for (int kk = 0; kk < DT_KC; kk += 4) {
    double a[8], b[4];
#pragma unroll
    for (int t = 0; t < 8; t++) a[t] = A_KM ? pa[kk * DT_LD_M + t * 8] : pa[t * 8 * DT_LD_K + kk];
#pragma unroll
    for (int u = 0; u < 4; u++) b[u] = B_KN ? pb[kk * DT_LD_N + u * 8] : pb[u * 8 * DT_LD_K + kk];
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
      for (int u = 0; u < 4; u++) dt_dmma(acc[t][u][0], acc[t][u][1], a[t], b[u]);
  }
}

typedef CUresult (*PFN_encodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled_t get_tensormap_encoder();
// FP64 row-major matrix [rows][cols] with leading dimension ld (doubles); box = boxc x boxr elements, zero fill out of bounds
int make_f64_tensormap(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int boxc, int boxr);

}  // namespace eb
