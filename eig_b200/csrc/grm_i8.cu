// grm_i8.cu -- K1b: the same symmetric rank-M update as grm_kernel.cu (domult_increment_lookup + block_increment_binary,
// smartpca.c:3426-3495 / 3361-3423), computed EXACTLY on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// The FP64 pipe of sm_100a ends at ~37 TFLOP/s (DMMA.8x8x4); the integer tensor pipe is ~100x wider.  The GRM admits an exact integer
// formulation because a column of the normalised genotype matrix only takes three values: x_is = v_is (a_s + b_s k_is) with the raw
// genotype k in {0,1,2}, the validity bit v and two per-SNP FP64 numbers a_s = -mean * scale, b_s = scale.  With the two integer bases
// h = k v in {0,1,2} and v in {0,1} one has, for every pair of individuals,
//        x_is x_js = h_is T1_s[k_js] + v_is T2_s[k_js],     T1[k] = ab + b^2 k,   T2[k] = a^2 + ab k   (0 for a missing genotype),
// i.e. two integer-times-table products per SNP (each is unsymmetric, their sum is the symmetric GRM; only its lower triangle is
// computed).  Every table entry is cut into NSL signed 7-bit digits of one fixed-point scale (T = sum_k m_k 2^(E - 7(k+1)), |m_k| <= 127):
// digit k of T[code] is ONE byte of the weighted operand, the other operand is the bare basis, and
// C_k = sum_s h_is m_k(T1_s[k_js]) + v_is m_k(T2_s[k_js]) is a u8 x s8 -> s32 tensor-core product that is EXACT.
// XTX = sum_k 2^(E-7(k+1)) C_k is then accumulated in FP64 by the epilogue.  With NSL = 8 the table entries carry 56 bits below the
// largest one -- more than the 53 of the FP64 products the DMMA path rounds -- and the sum over SNPs has NO rounding error at all, so the
// result is closer to the exact GRM than either FP64 implementation (the tests compare all three).
// SNPs without a missing genotype (class F) need one basis only:  x_i x_j = b^2 k_i k_j + ab (k_i + k_j) + a^2, the last two terms
// being rank one (an FP64 mat-vec over the packed matrix, added by the finalize kernel); 128-SNP blocks that contain only such SNPs
// are skipped in the v segment, so complete data costs NSL integer passes and data with missing genotypes 2 NSL.
//
// Kernels:
//   i8_prep_kernel       per SNP: class, the two 3-entry tables, the largest exponent, block flags, rank-one coefficients
//   i8_rank1_kernel      r_i = sum_{s in F} a_s b_s k_is  (packed matrix read once, FP64, fixed summation order)
//   i8_transform_kernel  packed 2-bit rows -> byte operands [matrix][SNP][individual] (MN-major for the MMA, PRMT as a 4-entry LUT)
//   grm_i8_kernel        persistent, warp-specialised: TMA producer (128B-swizzled 3-D tensor maps) -> 4-stage mbarrier ring ->
//                        one thread issuing tcgen05.mma.kind::i8 (M 128 x N 256 x K 32, both operands MN-major) into a
//                        double-buffered TMEM accumulator -> 4 epilogue warps (tcgen05.ld, s32 -> f64, scaled FP64 add into the tile)
//   grm_i8_pair_kernel   the same on CTA pairs (cluster of 2, tcgen05.mma.cta_group::2, 256 x 256 tiles, 6 stages) with a pass-level
//                        synchronisation of the clusters that keeps their shared operand columns in L2 -- the default
//   grm_i8_finalize_kernel / grm_i8_push_kernel   rank-one terms, mirror (symit2) | tiles to their owners' receive buffers (sharded)
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "tile_order.cuh"
#include "tc5.cuh"

namespace eb {

constexpr int I8_BK = 128;                       // SNPs per pipeline stage
constexpr int I8_TM = 128, I8_TN = 256;          // accumulator tile: M side = GRM columns, N side = GRM rows (stored transposed: coalesced)
constexpr int I8_STAGES = 4;
constexpr int I8_STAGE_A = I8_BK * 128;          // bytes: [128 SNPs][128 individuals]
constexpr int I8_STAGE_B = 2 * I8_STAGE_A;       // two 128-individual atoms
constexpr int I8_STAGE = I8_STAGE_A + I8_STAGE_B;
constexpr int I8_SMEM = I8_STAGES * I8_STAGE + 1024 + 256;
constexpr int I8_THREADS = 192;                  // warp 0 TMA, warp 1 MMA + TMEM allocation, warps 2..5 epilogue
constexpr int I8_MAXSL = 10;

struct I8Args {
  int npad, ntiles, nkb, nsl, nseg, first;
  double scale[I8_MAXSL];                        // 2^(E - 7 (k + 1))
  int sync_lag;                                  // pair kernel: -1 = free-running clusters, else passes a cluster may run ahead of the slowest
  unsigned int* sync_ctr;                        // pair kernel: arrivals at pass boundaries (zeroed before the launch), [2] = time-outs of the whole pass
  const int2* tiles;                             // (nb, mb) in L2-friendly order
  const uint8_t* kbflag;                         // per 128-SNP block of THIS slab: contains a used SNP with a missing genotype
  double* out;                                   // lower-tile accumulator [npad][npad]
};

// instruction descriptor for kind::i8: D = s32, A/B 8-bit (signed flag per operand), both MN-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ constexpr uint32_t i8_idesc(int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(I8_TN >> 3) << 17) |
         ((uint32_t)(I8_TM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------ the GEMM
// Work items: (tile, group).  A tile is 256 GRM rows (N side, block nb) x 128 GRM columns (M side, block mb) of the lower triangle.
// A group is one digit k; its (one or two) segments accumulate into one TMEM buffer over all the
// SNP blocks of the slab, then the epilogue adds scale[k] * C to the FP64 tile.  All three roles walk the same item sequence.
struct I8Walk {
  int gps, ngroups;
  __device__ __forceinline__ I8Walk(const I8Args& a) {
    gps = 1;
    ngroups = a.nsl;
  }
  __device__ __forceinline__ void group(const I8Args& a, int g, int& k, int& seg0, int& seg1, bool& neg) const {
    k = g; seg0 = 0; seg1 = a.nseg; neg = false;
  }
};

__global__ void __launch_bounds__(I8_THREADS, 1)
grm_i8_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ I8Args args) {
  extern __shared__ uint8_t i8_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(i8_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + I8_STAGES * I8_STAGE);
  uint64_t* empty = full + I8_STAGES;
  uint64_t* tfull = empty + I8_STAGES;          // [2] accumulator buffer ready for the epilogue
  uint64_t* tempty = tfull + 2;                 // [2] accumulator buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < I8_STAGES; s++) { i8_mbar_init(full + s, 1); i8_mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; b++) { i8_mbar_init(tfull + b, 1); i8_mbar_init(tempty + b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(i8_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const I8Walk walk(args);
  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
        const int2 tl = args.tiles[t];
        const int xa = tl.y * I8_TM, xb = tl.x * I8_TN;
        for (int g = 0; g < walk.ngroups; g++) {
          int k, seg0, seg1; bool neg;
          walk.group(args, g, k, seg0, seg1, neg);
          for (int seg = seg0; seg < seg1; seg++) {
            for (int kb = 0; kb < args.nkb; kb++) {
              if (seg > 0 && !args.kbflag[kb]) continue;
              i8_mbar_wait(empty + stage, phase ^ 1);
              uint8_t* sb = smem + stage * I8_STAGE;
              i8_mbar_expect_tx(full + stage, I8_STAGE);
              i8_tma_load_3d(sb, &mapA, xa, kb * I8_BK, seg, full + stage);
              i8_tma_load_3d(sb + I8_STAGE_A, &mapB, xb, kb * I8_BK, seg * args.nsl + k, full + stage);
              i8_tma_load_3d(sb + 2 * I8_STAGE_A, &mapB, xb + 128, kb * I8_BK, seg * args.nsl + k, full + stage);
              if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
        for (int g = 0; g < walk.ngroups; g++, it++) {
          int k, seg0, seg1; bool neg;
          walk.group(args, g, k, seg0, seg1, neg);
          const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
          i8_mbar_wait(tempty + buf, bphase ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem_base + buf * I8_TN;
          uint32_t acc = 0;
          for (int seg = seg0; seg < seg1; seg++) {
            const uint32_t idesc = i8_idesc(0, 1);                   // A: unsigned basis, B: signed digits
            for (int kb = 0; kb < args.nkb; kb++) {
              if (seg > 0 && !args.kbflag[kb]) continue;
              i8_mbar_wait(full + stage, phase);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t sa = i8_smem_u32(smem + stage * I8_STAGE);
              const uint32_t sbb = sa + I8_STAGE_A;
#pragma unroll
              for (int j = 0; j < I8_BK / 32; j++) {
                // one instruction = 32 SNPs = four 8-row groups (1024 bytes each); B: two 128-individual atoms 16 KB apart
                const uint64_t ad = i8_smem_desc(sa + j * 4096, I8_STAGE_A, 1024);
                const uint64_t bd = i8_smem_desc(sbb + j * 4096, I8_STAGE_A, 1024);
                i8_mma(tacc, ad, bd, idesc, acc);
                acc = 1;
              }
              i8_commit(empty + stage);          // frees the stage once these MMAs have read it
              if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
            }
          }
          i8_commit(tfull + buf);                // accumulator complete
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue: TMEM -> FP64 accumulate into the tile
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;                 // M index = GRM column inside the tile
    uint32_t it = 0;
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
      const int2 tl = args.tiles[t];
      const int nb = tl.x, mb = tl.y;
      for (int g = 0; g < walk.ngroups; g++, it++) {
        int k, seg0, seg1; bool neg;
        walk.group(args, g, k, seg0, seg1, neg);
        const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
        const double sc = neg ? -args.scale[k] : args.scale[k];
        const bool store_only = args.first && g == 0;
        i8_mbar_wait(tfull + buf, bphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + buf * I8_TN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int cg = 0; cg < I8_TN / 32; cg++) {
          const int row0 = nb * I8_TN + cg * 32;
          if (row0 >= args.npad) break;                         // beyond the matrix (odd number of 128-row tiles)
          if (row0 + 31 < mb * I8_TM) continue;                 // entirely above the diagonal
          uint32_t v[32];
          i8_tmem_ld32(taddr + cg * 32, v);
          double* p = args.out + (size_t)row0 * args.npad + (size_t)mb * I8_TM + m;
          if (store_only) {
#pragma unroll
            for (int c = 0; c < 32; c++) p[(size_t)c * args.npad] = (double)(int)v[c] * sc;
          } else {
            double o[32];
#pragma unroll
            for (int c = 0; c < 32; c++) o[c] = p[(size_t)c * args.npad];
#pragma unroll
            for (int c = 0; c < 32; c++) p[(size_t)c * args.npad] = fma((double)(int)v[c], sc, o[c]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) i8_mbar_arrive(tempty + buf);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------------------ the GEMM, CTA pairs
// Same work, two SMs per tile (cluster of 2, tcgen05.mma.cta_group::2): the pair computes 256 GRM columns (M side, 128 per CTA) x 256
// GRM rows (N side); every CTA loads ITS half of both operands (16 KB + 16 KB per 128-SNP stage instead of 16 + 32 for half the work),
// the leader's thread issues one M 256 x N 256 x K 32 instruction per 32 SNPs, and both CTAs drain their 128 accumulator lanes.
// L2 -> SM traffic per multiply-add drops by a third and the 6 (not 4) stages fit the same shared memory.
constexpr int I8P_STAGES = 6;
constexpr int I8P_STAGE = 2 * I8_STAGE_A;                 // 32 KB per CTA and stage
constexpr int I8P_SMEM = I8P_STAGES * I8P_STAGE + 1024 + 256;

__device__ __forceinline__ uint32_t i8_ld_acquire_u32(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t i8_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t i8_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void i8_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void i8_mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of this CTA's operand half; the transaction bytes are counted on the LEADER's barrier (cluster address)
__device__ __forceinline__ void i8_tma_load_3d_pair(void* dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          i8_smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar_cluster_addr)
      : "memory");
}
__device__ __forceinline__ constexpr uint32_t i8p_idesc(int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void i8p_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void i8p_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(i8_smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
grm_i8_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ I8Args args) {
  extern __shared__ uint8_t i8_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(i8_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + I8P_STAGES * I8P_STAGE);     // used in the leader only
  uint64_t* empty = full + I8P_STAGES;          // per CTA: the pair's MMAs have read this CTA's stage
  uint64_t* tfull = empty + I8P_STAGES;         // [2] per CTA: accumulator buffer complete
  uint64_t* tempty = tfull + 2;                 // [2] leader only: both CTAs have drained the buffer (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = i8_cluster_rank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < I8P_STAGES; s++) { i8_mbar_init(full + s, 1); i8_mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; b++) { i8_mbar_init(tfull + b, 1); i8_mbar_init(tempty + b, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(i8_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  i8_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const I8Walk walk(args);
  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs, each its own halves)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int pass = 0;
      for (int t = cluster_id; t < args.ntiles; t += nclusters) {
        const int2 tl = args.tiles[t];
        const int xa = tl.y * 256 + (int)rank * 128, xb = tl.x * 256 + (int)rank * 128;
        for (int g = 0; g < walk.ngroups; g++, pass++) {
          int k, seg0, seg1; bool neg;
          walk.group(args, g, k, seg0, seg1, neg);
          // Keep the clusters of a wave within sync_lag passes of each other: they stream the same operand columns through L2, and the
          // lines only survive there while everybody reads them at about the same time (free-running clusters drift apart, every one
          // then pulls its own copy from DRAM and the kernel becomes DRAM bound: 35 % L2 hit rate instead of ~85 %).
          if (rank == 0 && args.sync_lag >= 0) {
            atomicAdd(args.sync_ctr, 1u);
            const long long target = ((long long)pass + 1 - args.sync_lag) * (long long)nclusters;
            const long long t0 = clock64();
            while ((long long)i8_ld_acquire_u32(args.sync_ctr) < target) {
              __nanosleep(100);
              if (clock64() - t0 > (1ll << 32)) { atomicAdd(args.sync_ctr + 2, 1u); break; }      // ~2 s: never wedge the GPU
            }
          }
          for (int seg = seg0; seg < seg1; seg++) {
            for (int kb = 0; kb < args.nkb; kb++) {
              if (seg > 0 && !args.kbflag[kb]) continue;
              i8_mbar_wait(empty + stage, phase ^ 1);
              uint8_t* sb = smem + stage * I8P_STAGE;
              if (rank == 0) i8_mbar_expect_tx(full + stage, 2 * I8P_STAGE);
              const uint32_t lead_bar = i8_mapa(i8_smem_u32(full + stage), 0);
              i8_tma_load_3d_pair(sb, &mapA, xa, kb * I8_BK, seg, lead_bar);
              i8_tma_load_3d_pair(sb + I8_STAGE_A, &mapB, xb, kb * I8_BK, seg * args.nsl + k, lead_bar);
              if (++stage == I8P_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      // a cluster without a tile in the last round still owes that round's arrivals
      if (rank == 0 && args.sync_lag >= 0) {
        const int rounds = (args.ntiles + nclusters - 1) / nclusters;
        const int owed = rounds * walk.ngroups - pass;
        if (owed > 0) atomicAdd(args.sync_ctr, (unsigned int)owed);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread of the leader CTA)
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int t = cluster_id; t < args.ntiles; t += nclusters) {
        for (int g = 0; g < walk.ngroups; g++, it++) {
          int k, seg0, seg1; bool neg;
          walk.group(args, g, k, seg0, seg1, neg);
          const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
          i8_mbar_wait(tempty + buf, bphase ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem_base + buf * 256;
          uint32_t acc = 0;
          for (int seg = seg0; seg < seg1; seg++) {
            const uint32_t idesc = i8p_idesc(0, 1);                  // A: unsigned basis, B: signed digits
            for (int kb = 0; kb < args.nkb; kb++) {
              if (seg > 0 && !args.kbflag[kb]) continue;
              i8_mbar_wait(full + stage, phase);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t sa = i8_smem_u32(smem + stage * I8P_STAGE);
              const uint32_t sbb = sa + I8_STAGE_A;
#pragma unroll
              for (int j = 0; j < I8_BK / 32; j++) {
                const uint64_t ad = i8_smem_desc(sa + j * 4096, I8_STAGE_A, 1024);
                const uint64_t bd = i8_smem_desc(sbb + j * 4096, I8_STAGE_A, 1024);
                i8p_mma(tacc, ad, bd, idesc, acc);
                acc = 1;
              }
              i8p_commit(empty + stage);
              if (++stage == I8P_STAGES) { stage = 0; phase ^= 1; }
            }
          }
          i8p_commit(tfull + buf);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (both CTAs: 128 accumulator lanes each)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t it = 0;
    for (int t = cluster_id; t < args.ntiles; t += nclusters) {
      const int2 tl = args.tiles[t];
      const int nb = tl.x;
      const int col0 = tl.y * 256 + (int)rank * 128;         // first GRM column of this CTA's half
      for (int g = 0; g < walk.ngroups; g++, it++) {
        int k, seg0, seg1; bool neg;
        walk.group(args, g, k, seg0, seg1, neg);
        const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
        const double sc = neg ? -args.scale[k] : args.scale[k];
        const bool store_only = args.first && g == 0;
        i8_mbar_wait(tfull + buf, bphase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + buf * 256 + ((uint32_t)(q * 32) << 16);
        if (col0 < args.npad) {
#pragma unroll 1
          for (int cg = 0; cg < 8; cg++) {
            const int row0 = nb * 256 + cg * 32;
            if (row0 >= args.npad) break;
            if (row0 + 31 < col0) continue;
            uint32_t v[32];
            i8_tmem_ld32(taddr + cg * 32, v);
            double* p = args.out + (size_t)row0 * args.npad + (size_t)col0 + m;
            if (store_only) {
#pragma unroll
              for (int c = 0; c < 32; c++) p[(size_t)c * args.npad] = (double)(int)v[c] * sc;
            } else {
              double o[32];
#pragma unroll
              for (int c = 0; c < 32; c++) o[c] = p[(size_t)c * args.npad];
#pragma unroll
              for (int c = 0; c < 32; c++) p[(size_t)c * args.npad] = fma((double)(int)v[c], sc, o[c]);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) i8_mbar_arrive_remote(i8_mapa(i8_smem_u32(tempty + buf), 0));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  i8_cluster_sync();                                 // the peer's shared memory and barriers stay alive until both are done
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------ preparation
// prep[0] = largest exponent of a weight (frexp convention: w < 2^E), prep[1] = blocks with a missing genotype, prep[2] = used SNPs,
// prep[3] = sum of the exponents of w2 over the used SNPs (for the digit count)
// Per-SNP 3-entry tables of the weighted operand (entry = genotype code 0, 1, 2; missing -> 0):
//   T1[k] = ab + b^2 k  pairs with the genotype basis h,   T2[k] = a^2 + ab k  pairs with the validity basis v,
//   x_i x_j = h_i T1[k_j] + v_i T2[k_j]   for every pair of individuals (zero as soon as one of them is missing).
// SNPs without a missing genotype (class F): T1[k] = b^2 k only, the rest is rank one (coef = ab, asq = a^2).
__device__ __forceinline__ void i8_tables(const double* __restrict__ table, const int* __restrict__ nmiss, const uint8_t* __restrict__ used,
                                          int64_t s, int64_t nsnp, bool& use, bool& classF, double& a, double& b, double (&T1)[3], double (&T2)[3]) {
  use = s < nsnp && used[s];
  a = b = 0.0;
  T1[0] = T1[1] = T1[2] = T2[0] = T2[1] = T2[2] = 0.0;
  classF = true;
  if (!use) return;
  const double t0 = table[4 * s], t2 = table[4 * s + 2];
  a = t0; b = 0.5 * (t2 - t0);
  classF = nmiss[s] == 0;
  const double bb = b * b, ab = a * b;
  if (classF) { T1[1] = bb; T1[2] = 2.0 * bb; return; }
  T1[0] = ab; T1[1] = ab + bb; T1[2] = ab + 2.0 * bb;
  T2[0] = a * a; T2[1] = a * a + ab; T2[2] = a * a + 2.0 * ab;
}

__global__ void __launch_bounds__(256) i8_prep_kernel(const double* __restrict__ table, const int* __restrict__ nmiss, const uint8_t* __restrict__ used,
                                                      int64_t nsnp, int64_t mpad, uint8_t* __restrict__ kbflag, double* __restrict__ coef,
                                                      double* __restrict__ asq, long long* __restrict__ prep) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= mpad) return;
  bool use, classF; double a, b, T1[3], T2[3];
  i8_tables(table, nmiss, used, s, nsnp, use, classF, a, b, T1, T2);
  coef[s] = (use && classF) ? a * b : 0.0;
  asq[s] = (use && classF) ? a * a : 0.0;
  if (!use) return;
  const double w2 = b * b;
  const double wmax = fmax(fmax(fabs(T1[0]), fmax(fabs(T1[1]), fabs(T1[2]))), fmax(fabs(T2[0]), fmax(fabs(T2[1]), fabs(T2[2]))));
  int e = 0, e2 = 0;
  frexp(wmax, &e); frexp(w2, &e2);
  if (wmax > 0.0) atomicMax(reinterpret_cast<int*>(prep), e + 4096);     // biased: exponents may be negative
  if (!classF) kbflag[s / I8_BK] = 1;
  atomicAdd(reinterpret_cast<unsigned long long*>(prep + 2), 1ull);
  if (w2 > 0.0) atomicAdd(reinterpret_cast<unsigned long long*>(prep + 3), (unsigned long long)(long long)(e2 + 4096));
}

__global__ void __launch_bounds__(256) i8_count_flags_kernel(const uint8_t* __restrict__ kbflag, int nkb, long long* __restrict__ prep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nkb && kbflag[i]) atomicAdd(reinterpret_cast<unsigned long long*>(prep + 1), 1ull);
}

// r_part[chunk][i] = sum over the SNPs of the chunk of coef_s * k_is (class-F SNPs have no missing genotype among the real rows; the
// pad rows read code 3 and are masked by the finalize kernel).  One thread = one 32-bit word = 16 individuals.
constexpr int I8_R1_CHUNK = 1024;
__global__ void __launch_bounds__(128) i8_rank1_kernel(const uint8_t* __restrict__ work, int64_t wpitch, int64_t mpad, const double* __restrict__ coef,
                                                       double* __restrict__ rpart, int npad, int nchunks) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (int)(wpitch >> 2)) return;
  for (int chunk = blockIdx.y; chunk < nchunks; chunk += gridDim.y) {      // grid.y is capped at 65,535: very long SNP axes wrap around
    const int64_t s0 = (int64_t)chunk * I8_R1_CHUNK, s1 = min(s0 + (int64_t)I8_R1_CHUNK, mpad);
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0.0;
    for (int64_t s = s0; s < s1; s++) {
      const double cf = coef[s];
      if (cf == 0.0) continue;                                           // warp-uniform
      const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(work + s * wpitch) + w);
      const double c2 = cf + cf;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint32_t code = (x >> (((i >> 2) << 3) + ((3 - (i & 3)) << 1))) & 3u;
        acc[i] += (code & 1u) ? cf : 0.0;
        acc[i] += (code & 2u) ? c2 : 0.0;
      }
    }
    double* out = rpart + (size_t)chunk * npad + (size_t)w * 16;
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = acc[i];
  }
}
// r[i] = sum over chunks (fixed order); r[npad] = sum_s asq[s] (fixed-order tree of block 0)
__global__ void __launch_bounds__(256) i8_rank1_reduce_kernel(const double* __restrict__ rpart, int nchunks, int npad, const double* __restrict__ asq,
                                                              int64_t mpad, double* __restrict__ r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) {
    double v = 0.0;
    for (int c = 0; c < nchunks; c++) v += rpart[(size_t)c * npad + i];
    r[i] = v;
  }
  if (blockIdx.x == 0) {
    __shared__ double sh[256];
    double v = 0.0;
    for (int64_t s = threadIdx.x; s < mpad; s += 256) v += asq[s];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) r[npad] = sh[0];
  }
}

// Packed rows of one slab -> byte operands.  A[seg][row][npad]: the bare basis (h | v);
// B[seg * nsl + k][row][npad]: digit k (signed) of the segment's table entry T1 / T2 for the genotype.  One block per SNP; one thread per 32-bit packed word
// (16 individuals): the four 2-bit codes of a byte become the selector of one PRMT over the 4-entry table of the output matrix.
__global__ void __launch_bounds__(256) i8_transform_kernel(const uint8_t* __restrict__ work, int64_t wpitch, const double* __restrict__ table,
                                                           const int* __restrict__ nmiss, const uint8_t* __restrict__ used, const uint8_t* __restrict__ kbflag,
                                                           int64_t nsnp, int64_t s_first, int ks_alloc, int npad, int nsl, int E, int nseg,
                                                           uint8_t* __restrict__ A, uint8_t* __restrict__ B) {
  __shared__ uint32_t lut[2][1 + I8_MAXSL];
  const int row = blockIdx.x;
  const int64_t s = s_first + row;
  const bool flagged = kbflag[row / I8_BK] != 0;
  const int segs = (nseg == 2 && flagged) ? 2 : 1;
  if (threadIdx.x < 2 * (1 + nsl)) {
    const int seg = threadIdx.x / (1 + nsl), j = threadIdx.x - seg * (1 + nsl);
    bool use, classF; double a, b, T1[3], T2[3];
    i8_tables(table, nmiss, used, s, nsnp, use, classF, a, b, T1, T2);
    uint32_t word;
    if (j == 0) {
      word = seg == 0 ? 0x00020100u : 0x00010101u;                   // bare basis for the codes 0, 1, 2, 3: h | v
    } else {
      // digit k (7 bits, most significant first, with the sign of the value) of the three table entries: one byte each
      const int k = j - 1;
      word = 0;
#pragma unroll
      for (int cd = 0; cd < 3; cd++) {
        const double t = seg == 0 ? T1[cd] : T2[cd];
        unsigned long long qv = __double2ull_rn(ldexp(fabs(t), 7 * nsl - E));
        const unsigned long long qmax = (nsl * 7 >= 64) ? ~0ull : ((1ull << (7 * nsl)) - 1ull);
        if (qv > qmax) qv = qmax;
        int mk = (int)((qv >> (7 * (nsl - 1 - k))) & 127ull);
        if (t < 0.0) mk = -mk;
        word |= ((uint32_t)(mk & 0xFF)) << (8 * cd);
      }
    }
    lut[seg][j] = word;
  }
  __syncthreads();
  const int words = (int)(wpitch >> 2);
  const size_t mat = (size_t)ks_alloc * npad;
  for (int w = threadIdx.x; w < words; w += blockDim.x) {
    const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(work + s * wpitch) + w);
    uint32_t sel[4];
#pragma unroll
    for (int bi = 0; bi < 4; bi++) {
      const uint32_t by = (x >> (8 * bi)) & 0xFFu;
      sel[bi] = (by >> 6) | (((by >> 4) & 3u) << 4) | (((by >> 2) & 3u) << 8) | ((by & 3u) << 12);
    }
    const size_t off = (size_t)row * npad + (size_t)w * 16;
    for (int seg = 0; seg < segs; seg++) {
      const uint32_t la = lut[seg][0];
      *reinterpret_cast<uint4*>(A + seg * mat + off) =
          make_uint4(__byte_perm(la, 0, sel[0]), __byte_perm(la, 0, sel[1]), __byte_perm(la, 0, sel[2]), __byte_perm(la, 0, sel[3]));
      for (int k = 0; k < nsl; k++) {
        const uint32_t lb = lut[seg][1 + k];
        *reinterpret_cast<uint4*>(B + (size_t)(seg * nsl + k) * mat + off) =
            make_uint4(__byte_perm(lb, 0, sel[0]), __byte_perm(lb, 0, sel[1]), __byte_perm(lb, 0, sel[2]), __byte_perm(lb, 0, sel[3]));
      }
    }
  }
}

// lower tiles of `acc` (+ the rank-one terms of the class-F SNPs) -> full symmetric xtx (symit2, smartpca.c:480-508)
__global__ void __launch_bounds__(256) grm_i8_finalize_kernel(const double* __restrict__ acc, int npad, int nrows, const double* __restrict__ r,
                                                              double* __restrict__ xtx) {
  __shared__ double tile[32][33];
  int bi, bj;
  {
    const int t = blockIdx.x;
    int rr = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((rr + 1) * (rr + 2) / 2 <= t) rr++;
    while (rr * (rr + 1) / 2 > t) rr--;
    bi = rr; bj = t - rr * (rr + 1) / 2;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const double cc = r[npad];
  for (int rw = ty; rw < 32; rw += 8) {
    const int i = bi * 32 + rw, j = bj * 32 + tx;
    double v = acc[(size_t)i * npad + j];
    if (i < nrows && j < nrows) v += r[i] + r[j] + cc;
    else v = 0.0;
    if (bi == bj && tx > rw) v = 0.0;
    tile[rw][tx] = v;
  }
  __syncthreads();
  for (int rw = ty; rw < 32; rw += 8) {
    double v = tile[rw][tx];
    if (bi == bj && tx > rw) v = tile[tx][rw];
    xtx[(size_t)(bi * 32 + rw) * npad + bj * 32 + tx] = v;
    if (bi != bj) xtx[(size_t)(bj * 32 + rw) * npad + bi * 32 + tx] = tile[tx][rw];
  }
}

// SNPs sharded over several GPUs: every lower-triangle 128 x 128 tile (same numbering and ownership as grm_syrk_kernel's epilogue,
// common.cuh GrmPush) goes to the receive buffer of the rank that owns it, rank-one terms included; plain 16-byte stores over NVLink.
__global__ void __launch_bounds__(256) grm_i8_push_kernel(const double* __restrict__ acc, int npad, int nrows, const double* __restrict__ r, int ntiles,
                                                          int nsplit, const __grid_constant__ GrmPush push) {
  const double cc = r[npad];
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    int ti, tj;
    tile_decode_banded(t, npad / TILE, ti, tj);
    const int owner = t % push.world;
    double* base = push.recv[owner] + ((size_t)(t / push.world) * (size_t)(push.world * nsplit) + (size_t)(push.rank * nsplit)) * (TILE * TILE);
    for (int e = threadIdx.x; e < TILE * TILE / 2; e += blockDim.x) {
      const int rw = e >> 6, c2 = (e & 63) << 1;
      const int i = ti * TILE + rw, j = tj * TILE + c2;
      double2 v = *reinterpret_cast<const double2*>(acc + (size_t)i * npad + j);
      const double ri = (i < nrows) ? r[i] + cc : 0.0;
      v.x = (i < nrows && j < nrows) ? v.x + ri + r[j] : 0.0;
      v.y = (i < nrows && j + 1 < nrows) ? v.y + ri + r[j + 1] : 0.0;
      *reinterpret_cast<double2*>(base + (size_t)rw * TILE + c2) = v;
      for (int ch = 1; ch < nsplit; ch++) *reinterpret_cast<double2*>(base + (size_t)ch * TILE * TILE + (size_t)rw * TILE + c2) = make_double2(0.0, 0.0);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled_i8)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int i8_make_map(CUtensorMap* map, const uint8_t* base, int npad, int rows, int nmats) {
  static PFN_encodeTiled_i8 enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiled_i8>(p);
  }
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return EB_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)npad, (cuuint64_t)rows, (cuuint64_t)nmats};
  cuuint64_t strides[2] = {(cuuint64_t)npad, (cuuint64_t)npad * (cuuint64_t)rows};
  cuuint32_t box[3] = {128, (cuuint32_t)I8_BK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (i8 operands) failed: %d (npad %d rows %d mats %d)", (int)r, npad, rows, nmats); return EB_ERR_CUDA; }
  return 0;
}

// lower-triangle tiles (nb over 256 GRM rows, mb over 128 GRM columns, mb <= 2 nb + 1) in bands of I8_BAND tile rows walked column by
// column, so that the tiles the persistent CTAs work on at any time share their operands in L2 (see tile_order.cuh)
constexpr int I8_BAND = 8;
static void i8_tile_order(int npad, std::vector<int2>& tiles) {
  const int MB = npad / I8_TM, NB = (npad + I8_TN - 1) / I8_TN;
  tiles.clear();
  for (int b0 = 0; b0 < NB; b0 += I8_BAND) {
    const int b1 = std::min(NB, b0 + I8_BAND);
    const int mbmax = std::min(MB - 1, 2 * (b1 - 1) + 1);
    for (int mb = 0; mb <= mbmax; mb++)
      for (int nb = b0; nb < b1; nb++)
        if (mb <= 2 * nb + 1) tiles.push_back(make_int2(nb, mb));
  }
}

// CTA pairs: 256 x 256 tiles (nb, mb2 <= nb), same banding
static void i8_pair_tile_order(int npad, std::vector<int2>& tiles) {
  const int NB = (npad + 255) / 256;
  tiles.clear();
  for (int b0 = 0; b0 < NB; b0 += I8_BAND) {
    const int b1 = std::min(NB, b0 + I8_BAND);
    for (int mb = 0; mb < b1; mb++)
      for (int nb = std::max(b0, mb); nb < b1; nb++) tiles.push_back(make_int2(nb, mb));
  }
}

bool grm_use_i8(const eb_ctx* c) {
  if (c->opt_grm_method == 1) return false;
  if (c->opt_grm_method == 2) return true;
  return c->nrows >= c->opt_i8_min;
}

// work + table -> acc (c->partial, lower tiles, rank-one terms NOT yet added) -> xtx (finalize_local) | owners' receive buffers (push)
int grm_accumulate_i8(eb_ctx* c, bool finalize_local, bool push_mode) {
  int rc;
  const int npad = c->npad;
  const int nkb_all = (int)(c->mpad / I8_BK);
  const size_t plane = (size_t)npad * npad;
  c->nsplit = push_mode ? c->grm_geom_nsplit : 1;
  if ((rc = c->partial.ensure(plane))) return rc;
  if ((rc = c->xtx.ensure(plane))) return rc;
  if ((rc = c->trace_d.ensure(1))) return rc;
  if ((rc = c->i8_flag.ensure((size_t)nkb_all))) return rc;
  if ((rc = c->i8_coef.ensure((size_t)c->mpad * 2))) return rc;
  if ((rc = c->i8_prep.ensure(8))) return rc;
  const int nchunks = (int)((c->mpad + I8_R1_CHUNK - 1) / I8_R1_CHUNK);
  if ((rc = c->i8_r.ensure((size_t)(nchunks + 1) * npad + 8))) return rc;
  double* coef = c->i8_coef.p;
  double* asq = coef + c->mpad;
  double* rpart = c->i8_r.p;
  double* rvec = rpart + (size_t)nchunks * npad;

  EB_CUDA(cudaEventRecord(c->ev[2], c->stream));
  EB_CUDA(cudaMemsetAsync(c->i8_flag.p, 0, (size_t)nkb_all, c->stream));
  EB_CUDA(cudaMemsetAsync(c->i8_prep.p, 0, 8 * sizeof(long long), c->stream));
  i8_prep_kernel<<<(unsigned)((c->mpad + 255) / 256), 256, 0, c->stream>>>(c->table_d.p, c->nmiss_d.p, c->used_d.p, c->nsnp, c->mpad, c->i8_flag.p,
                                                                           coef, asq, c->i8_prep.p);
  EB_CHECK_LAUNCH(c);
  i8_count_flags_kernel<<<(nkb_all + 255) / 256, 256, 0, c->stream>>>(c->i8_flag.p, nkb_all, c->i8_prep.p);
  EB_CHECK_LAUNCH(c);
  long long prep[4] = {0, 0, 0, 0};
  std::vector<uint8_t> flag_h((size_t)nkb_all);
  EB_CUDA(cudaMemcpyAsync(prep, c->i8_prep.p, sizeof(prep), cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaMemcpyAsync(flag_h.data(), c->i8_flag.p, (size_t)nkb_all, cudaMemcpyDeviceToHost, c->stream));
  // the rank-one terms of the SNPs without missing genotypes run while the host decides the geometry
  {
    dim3 grid((unsigned)(((c->wpitch >> 2) + 127) / 128), (unsigned)std::min(nchunks, 65535));
    i8_rank1_kernel<<<grid, 128, 0, c->stream>>>(c->work.p, c->wpitch, c->mpad, coef, rpart, npad, nchunks);
    EB_CHECK_LAUNCH(c);
    i8_rank1_reduce_kernel<<<(npad + 255) / 256, 256, 0, c->stream>>>(rpart, nchunks, npad, asq, c->mpad, rvec);
    EB_CHECK_LAUNCH(c);
  }
  EB_CUDA(cudaStreamSynchronize(c->stream));
  const int Emax = (int)(reinterpret_cast<int*>(prep)[0]) - 4096;
  const long long nflag = prep[1], nuse = prep[2];
  const bool any_weight = reinterpret_cast<int*>(prep)[0] != 0;
  // digits: 52 bits below the typical weight (mean exponent of w2 over the used SNPs), 7 bits each
  int nsl = c->opt_i8_slices;
  if (nsl <= 0) {
    const double eavg = nuse > 0 ? (double)prep[3] / (double)nuse - 4096.0 : (double)Emax;
    nsl = (int)ceil((52.0 + std::max(0.0, (double)Emax - eavg)) / 7.0);
    nsl = std::max(7, std::min(nsl, 9));
  }
  nsl = std::max(1, std::min(nsl, 9));
  const int nseg_all = nflag > 0 ? 2 : 1;
  c->tm.i8_slices = nsl; c->tm.i8_segments = nseg_all; c->tm.i8_flag_blocks = (int)nflag; c->tm.i8_slab_rows = 0;

  // slab: as many SNP blocks as the operand budget allows (A: nseg matrices, B: nseg * nsl matrices of [rows][npad] bytes)
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  const size_t have = c->i8_ops.n;
  const size_t per_row = (size_t)nseg_all * (1 + nsl) * (size_t)npad;
  size_t budget = std::min(have + (size_t)((double)freeb * 0.4), (size_t)48 << 30);    // leave room for the eigensolver's copy of the matrix
  if (c->opt_i8_slab > 0) budget = std::min(budget, (size_t)c->opt_i8_slab * per_row);
  long long rows = (long long)(budget / per_row) / I8_BK * I8_BK;
  rows = std::min<long long>(rows, c->mpad);
  rows = std::min<long long>(rows, 1 << 20);
  if (rows < I8_BK) { set_error("grm (i8): not enough device memory for one 128-SNP operand block (%zu bytes per SNP row)", per_row); return EB_ERR_NOMEM; }
  if ((rc = c->i8_ops.ensure((size_t)rows * per_row))) return rc;
  c->tm.i8_slab_rows = (int)rows;
  uint8_t* Aop = c->i8_ops.p;
  uint8_t* Bop = Aop + (size_t)nseg_all * rows * npad;

  std::vector<int2> tiles;
  const bool pair = c->opt_i8_pair != 0;
  if (pair) i8_pair_tile_order(npad, tiles); else i8_tile_order(npad, tiles);
  if ((rc = c->i8_tiles.ensure(tiles.size() * 2))) return rc;
  EB_CUDA(cudaMemcpyAsync(c->i8_tiles.p, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, c->stream));

  CUtensorMap mapA, mapB;
  if ((rc = i8_make_map(&mapA, Aop, npad, (int)rows, nseg_all))) return rc;
  if ((rc = i8_make_map(&mapB, Bop, npad, (int)rows, nseg_all * nsl))) return rc;
  EB_CUDA(cudaFuncSetAttribute(grm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
  EB_CUDA(cudaFuncSetAttribute(grm_i8_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8P_SMEM));

  I8Args args;
  memset(&args, 0, sizeof(args));
  args.npad = npad; args.ntiles = (int)tiles.size(); args.nsl = nsl;
  for (int k = 0; k < nsl; k++) args.scale[k] = ldexp(1.0, Emax - 7 * (k + 1));
  args.tiles = reinterpret_cast<const int2*>(c->i8_tiles.p);
  args.out = c->partial.p;
  int grid = pair ? 2 * std::min((int)tiles.size(), c->num_sms / 2) : std::min((int)tiles.size(), c->num_sms);
  if (pair) {
    // the pass-level synchronisation spins on a global counter: every cluster of the grid must be resident
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(I8_THREADS); cfg.dynamicSmemBytes = I8P_SMEM;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int ncl = 0;
    EB_CUDA(cudaOccupancyMaxActiveClusters(&ncl, grm_i8_pair_kernel, &cfg));
    if (ncl < 1) { set_error("grm (i8): no CTA pair fits on this device"); return EB_ERR_CUDA; }
    grid = std::min(grid, 2 * ncl);
  }
  const double tile_macs = pair ? 256.0 * 256.0 : (double)I8_TM * I8_TN;
  c->grm_grid = 0;                                       // no per-CTA self-measurement on this path
  bool first = true;
  double ops = 0.0;
  c->i8_nlaunch = 0;
  if (!any_weight) {
    EB_CUDA(cudaMemsetAsync(c->partial.p, 0, plane * sizeof(double), c->stream));     // no used SNP at all
  } else {
    for (long long s0 = 0; s0 < c->mpad; s0 += rows) {
      const int ks = (int)std::min<long long>(rows, c->mpad - s0);
      const int kb0 = (int)(s0 / I8_BK), nkb = ks / I8_BK;
      int nfl = 0;
      for (int kb = 0; kb < nkb; kb++) nfl += flag_h[kb0 + kb] ? 1 : 0;
      const int nseg = nfl > 0 ? 2 : 1;
      i8_transform_kernel<<<ks, 256, 0, c->stream>>>(c->work.p, c->wpitch, c->table_d.p, c->nmiss_d.p, c->used_d.p, c->i8_flag.p + kb0, c->nsnp, s0,
                                                     (int)rows, npad, nsl, Emax, nseg, Aop, Bop);
      EB_CHECK_LAUNCH(c);
      args.nkb = nkb; args.nseg = nseg; args.first = first ? 1 : 0; args.kbflag = c->i8_flag.p + kb0;
      // events around every integer GEMM launch: the kernel-only time of the roofline line (eb_timings.i8_gemm_ms)
      while (c->i8_ev.size() < 2 * (size_t)(c->i8_nlaunch + 1)) {
        cudaEvent_t e;
        EB_CUDA(cudaEventCreate(&e));
        c->i8_ev.push_back(e);
      }
      EB_CUDA(cudaEventRecord(c->i8_ev[2 * c->i8_nlaunch], c->stream));
      if (pair) {
        args.sync_lag = c->opt_i8_sync;
        args.sync_ctr = reinterpret_cast<unsigned int*>(c->i8_prep.p + 4);
        EB_CUDA(cudaMemsetAsync(c->i8_prep.p + 4, 0, sizeof(int), c->stream));
        grm_i8_pair_kernel<<<grid, I8_THREADS, I8P_SMEM, c->stream>>>(mapA, mapB, args);
      }
      else grm_i8_kernel<<<grid, I8_THREADS, I8_SMEM, c->stream>>>(mapA, mapB, args);
      EB_CHECK_LAUNCH(c);
      EB_CUDA(cudaEventRecord(c->i8_ev[2 * c->i8_nlaunch + 1], c->stream));
      c->i8_nlaunch++;
      first = false;
      ops += (double)tiles.size() * tile_macs * 2.0 * (double)I8_BK * nsl * ((double)nkb + (double)nfl * (nseg == 2 ? 1 : 0));
    }
  }
  c->tm.i8_tera_ops = (float)(ops * 1e-12);
  EB_CUDA(cudaMemcpyAsync(c->i8_sync_h, c->i8_prep.p + 4, sizeof(c->i8_sync_h), cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaEventRecord(c->ev[3], c->stream));
  c->tm.grm_launches = 2;
  if (push_mode) {
    GrmPush push;
    memset(&push, 0, sizeof(push));
    if ((rc = peer_grm_push_args(c, &push))) return rc;
    const int T = npad / TILE;
    grm_i8_push_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(c->partial.p, npad, c->nrows, rvec, T * (T + 1) / 2, c->grm_geom_nsplit, push);
    EB_CHECK_LAUNCH(c);
    return 0;
  }
  if (!finalize_local) { set_error("grm (i8): a non-local finalize needs the push exchange"); return EB_ERR_STATE; }
  const int T32 = npad / 32;
  grm_i8_finalize_kernel<<<T32 * (T32 + 1) / 2, 256, 0, c->stream>>>(c->partial.p, npad, c->nrows, rvec, c->xtx.p);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[4], c->stream));
  return 0;
}

}  // namespace eb
