// pg_i8.cu -- the packed x skinny products of fastmode and of the projection passes (kjg_fpca_XTXA/_XA/_XTB, kjg_fpca.c:104-178;
// smartpca.c:1485-1525) on the 5th-generation tensor cores, with the packed matrix decoded INSIDE the kernel.
//
//   MODE_XA  : Out_t[l][s] = sum_i x_si In_t[l][i]      rows = SNPs,        K = individuals
//   MODE_XTB : Out_t[l][i] = sum_s x_si In_t[l][s]      rows = individuals, K = SNPs
// with x_si = table[s][code(s,i)], and every table of the library linear in the genotype: x = v (a_s + b_s k), a = t0, b = (t2 - t0)/2.
// The FP64 DMMA version (packed_gemm_kernel, fpca_kernels.cu) is bound by the FP64 pipe at ~100 GB/s of packed reads.  Here the packed
// operand becomes two byte matrices -- the validity basis v in {0,1} and the genotype basis h = k v in {0,1,2} -- and the skinny FP64
// operand becomes ND = 8 signed 7-bit digit matrices per column (one power-of-two scale per column, 56 bits below the column's largest
// entry), so that every product is an exact u8 x s8 -> s32 tensor-core product:
//   XA : out[s][l] = a_s sum_q 2^(e_l - 7(q+1)) Cv[s][(l,q)] + b_s sum_q ... Ch[s][(l,q)],  Cv = V D^T, Ch = H D^T, D = digits of In
//   XTB: out[i][l] = sum_q 2^(e_l - 7(q+1)) C[i][(l,q)],  C = V^T Da^T + H^T Db^T,  Da / Db = digits of a_s In[l][s] / b_s In[l][s]
// One CTA = one tile of 128 output rows (x one K split); per stage of 128 K elements: TMA brings the 4 KB packed sub-tile (128 SNPs x 32
// bytes) and the digit rows (N x 128 bytes, 128-byte swizzle); four warps turn the packed sub-tile into the two 16 KB byte operands
// directly in the UMMA shared-memory layout (one PRMT per 4 genotypes and basis; the SAME bytes are a K-major operand for XA -- rows =
// SNPs -- and an MN-major operand for XTB -- rows of K = SNPs); one thread issues the tcgen05.mma.kind::i8 instructions into TMEM; the
// four warps then read the accumulators back, apply the digit scales (and a_s, b_s) in FP64 and store the output columns.
// The packed matrix is read once per product (N M / 4 bytes) -- the kernels the north star asks for: bandwidth / decode bound.
#include <algorithm>
#include <cmath>
#include <vector>
#include "common.cuh"
#include "tc5.cuh"

namespace eb {

constexpr int PGI_ND = 8;                 // digits per FP64 value
constexpr int PGI_BK = 128;               // K elements per stage
constexpr int PGI_TILE_A = 128 * 128;     // one byte operand tile
constexpr int PGI_PACKED = 128 * 32;
constexpr int PGI_MAXC = 32;              // output columns per launch (N = 8 x columns <= 256)
constexpr int PGI_DECW = 4;               // decode / epilogue warps
constexpr int PGI_THREADS = 64 + 32 * PGI_DECW;
enum { PGI_XA = 0, PGI_XTB = 1 };

struct PgiArgs {
  int ncp;                  // padded columns of this launch (even), N = ncp * 8
  int ncols;                // real columns
  int nk, nsplit;           // K stages in total, K splits (grid.y)
  int nraw;                 // depth of the TMA ring (digit rows + packed sub-tile)
  int tile0;                // first row tile of this launch (blockIdx.x counts from it)
  int64_t rows;             // valid output rows (global row index < rows)
  int64_t out_row0;         // global row that maps to output row 0
  int64_t ld_l, ld_r;       // output element (row, column l) at out[(row - out_row0) * ld_r + l * ld_l]
  int64_t plane_stride;
  double oscale;
  const double* table;      // [mpad][4]
  const double* colscale;   // [ncp]: 2^(e_l)
  double* out;              // Out_t[l][row] (or plane ks)
};

__device__ __forceinline__ void pgi_tma_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(i8_smem_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(i8_smem_u32(bar))
               : "memory");
}
// K-major operand, 128-byte swizzle: rows of 128 bytes along K, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t pgi_desc_kmajor(uint32_t addr) { return i8_smem_desc(addr, 16, 1024); }
__device__ __forceinline__ uint32_t pgi_idesc(int a_mn_major, int n) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// Two rings: the RAW ring (digit rows + packed sub-tile, filled by TMA, `nraw` stages deep) and the OPERAND ring (the two decoded byte
// tiles, 2 stages).  A raw stage is free again when the decode warps have read its packed sub-tile AND the MMAs have read its digit
// rows (PGI_DECW + 1 arrivals); an operand stage when the MMAs that read it have completed (tcgen05.commit).
template <int MODE>
__global__ void __launch_bounds__(PGI_THREADS, 1)
pg_i8_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapD, const __grid_constant__ PgiArgs args) {
  extern __shared__ uint8_t pgi_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pgi_smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = args.ncp * PGI_ND;
  const int nB = MODE == PGI_XTB ? 2 : 1;                 // digit matrices per stage
  const int d_bytes = nB * N * 128;
  const int raw_bytes = d_bytes + PGI_PACKED;             // digits | packed   (N * 128 is a multiple of 2048: everything stays 1024-aligned)
  const int nraw = args.nraw;
  constexpr int NOP = 2;
  uint8_t* ops = smem;                                    // NOP x (A_v | A_h)
  uint8_t* raw = smem + NOP * 2 * PGI_TILE_A;
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(raw + nraw * raw_bytes);
  uint64_t* raw_empty = raw_full + 8;
  uint64_t* op_full = raw_empty + 8;
  uint64_t* op_empty = op_full + NOP;
  uint64_t* accbar = op_empty + NOP;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nraw; s++) { i8_mbar_init(raw_full + s, 1); i8_mbar_init(raw_empty + s, PGI_DECW + 1); }
    for (int s = 0; s < NOP; s++) { i8_mbar_init(op_full + s, PGI_DECW); i8_mbar_init(op_empty + s, 1); }
    i8_mbar_init(accbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapD) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(i8_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int tile = args.tile0 + blockIdx.x, ks = blockIdx.y, zg = blockIdx.z;      // zg: group of ncp output columns (digit rows zg * N ..)
  const int kb0 = (int)(((long long)args.nk * ks) / args.nsplit), kb1 = (int)(((long long)args.nk * (ks + 1)) / args.nsplit);

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t rs = 0, rph = 0;
      for (int kb = kb0; kb < kb1; kb++) {
        i8_mbar_wait(raw_empty + rs, rph ^ 1);
        uint8_t* sb = raw + rs * raw_bytes;
        i8_mbar_expect_tx(raw_full + rs, (uint32_t)raw_bytes);
        // packed sub-tile: 128 SNP rows x 32 bytes (128 individuals)
        if (MODE == PGI_XA) pgi_tma_2d(sb + d_bytes, &mapW, kb * 32, tile * 128, raw_full + rs);
        else pgi_tma_2d(sb + d_bytes, &mapW, tile * 32, kb * 128, raw_full + rs);
        for (int b = 0; b < nB; b++) pgi_tma_2d(sb + b * N * 128, &mapD, kb * 128, (b * (int)gridDim.z + zg) * N, raw_full + rs);
        if (++rs == (uint32_t)nraw) { rs = 0; rph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      uint32_t rs = 0, rph = 0, os = 0, oph = 0, acc = 0;
      const uint32_t idesc = pgi_idesc(MODE == PGI_XTB ? 1 : 0, N);
      for (int kb = kb0; kb < kb1; kb++) {
        i8_mbar_wait(raw_full + rs, rph);
        i8_mbar_wait(op_full + os, oph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = i8_smem_u32(ops + os * 2 * PGI_TILE_A);
        const uint32_t sd = i8_smem_u32(raw + rs * raw_bytes);
#pragma unroll
        for (int j = 0; j < PGI_BK / 32; j++) {
          if (MODE == PGI_XA) {
            // A K-major (rows = SNPs): 32 K bytes further inside the 128-byte row; the same digit tile for both bases
            const uint64_t bd = pgi_desc_kmajor(sd + j * 32);
            i8_mma(tmem_base, pgi_desc_kmajor(sa + j * 32), bd, idesc, acc);
            i8_mma(tmem_base + 256, pgi_desc_kmajor(sa + PGI_TILE_A + j * 32), bd, idesc, acc);
          } else {
            // A MN-major (K rows = SNPs, 128 individuals along MN): 32 K rows = four 1024-byte groups further
            i8_mma(tmem_base, i8_smem_desc(sa + j * 4096, PGI_TILE_A, 1024), pgi_desc_kmajor(sd + j * 32), idesc, acc);
            i8_mma(tmem_base, i8_smem_desc(sa + PGI_TILE_A + j * 4096, PGI_TILE_A, 1024), pgi_desc_kmajor(sd + N * 128 + j * 32), idesc, 1u);
          }
          acc = 1;
        }
        i8_commit(op_empty + os);
        i8_commit(raw_empty + rs);
        if (++rs == (uint32_t)nraw) { rs = 0; rph ^= 1; }
        if (++os == NOP) { os = 0; oph ^= 1; }
      }
      i8_commit(accbar);
    }
    __syncwarp();
  } else {
    // ===================================================================== decode warps, then epilogue
    const int dt = threadIdx.x - 64;                 // 0 .. 32 PGI_DECW - 1
    uint32_t rs = 0, rph = 0, os = 0, oph = 0;
    for (int kb = kb0; kb < kb1; kb++) {
      i8_mbar_wait(raw_full + rs, rph);
      i8_mbar_wait(op_empty + os, oph ^ 1);
      uint8_t* sb = ops + os * 2 * PGI_TILE_A;
      const uint32_t* pk = reinterpret_cast<const uint32_t*>(raw + rs * raw_bytes + d_bytes);
#pragma unroll
      for (int it = 0; it < 1024 / (32 * PGI_DECW); it++) {
        const int w = it * (32 * PGI_DECW) + dt;    // word index: row r = w >> 3 (SNP), j = w & 7 (16 individuals)
        const int r = w >> 3, j = w & 7;
        const uint32_t x = pk[w];
        uint32_t ov[4], oh[4];
#pragma unroll
        for (int bi = 0; bi < 4; bi++) {
          const uint32_t by = (x >> (8 * bi)) & 0xFFu;
          const uint32_t sel = (by >> 6) | (((by >> 4) & 3u) << 4) | (((by >> 2) & 3u) << 8) | ((by & 3u) << 12);
          ov[bi] = __byte_perm(0x00010101u, 0, sel);
          oh[bi] = __byte_perm(0x00020100u, 0, sel);
        }
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sb + off) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<uint4*>(sb + PGI_TILE_A + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) { i8_mbar_arrive(op_full + os); i8_mbar_arrive(raw_empty + rs); }
      if (++rs == (uint32_t)nraw) { rs = 0; rph ^= 1; }
      if (++os == NOP) { os = 0; oph ^= 1; }
    }
    // ---- epilogue: this thread owns output row (tile * 128 + m)
    const int q = warp & 3, m = q * 32 + lane;
    const int64_t row = (int64_t)tile * 128 + m;
    i8_mbar_wait(accbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    double a_s = 1.0, b_s = 0.0;
    if (MODE == PGI_XA && row < args.rows) {
      const double t0 = args.table[4 * row], t2 = args.table[4 * row + 2];
      a_s = t0; b_s = 0.5 * (t2 - t0);
    }
    double* outp = args.out + (size_t)ks * args.plane_stride + (size_t)(row - args.out_row0) * args.ld_r;
    // two warps share a TMEM lane quarter: they take alternate 32-column chunks
    for (int c0 = 32 * ((warp - 2) >> 2); c0 < N; c0 += 32 * (PGI_DECW / 4)) {      // 32 accumulator columns = 4 output columns x 8 digits
      uint32_t v[32], h[32];
      i8_tmem_ld32(taddr + c0, v);
      if (MODE == PGI_XA) i8_tmem_ld32(taddr + 256 + c0, h);
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        const int l = zg * args.ncp + (c0 >> 3) + cc;
        if (l >= args.ncols) break;
        const double sc = args.colscale[l];
        double sv = 0.0, sh = 0.0, f = 1.0 / 128.0;
#pragma unroll
        for (int d = 0; d < PGI_ND; d++) {
          sv += (double)(int)v[cc * 8 + d] * f;
          if (MODE == PGI_XA) sh += (double)(int)h[cc * 8 + d] * f;
          f *= 1.0 / 128.0;
        }
        const double y = (MODE == PGI_XA ? (a_s * sv + b_s * sh) : sv) * sc * args.oscale;
        if (row < args.rows) outp[(size_t)l * args.ld_l] = y;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------ digit slicing
// colmax[b][l] = max_k |w_b(k) In[l][k]|, w = 1 (XA) or a_k, b_k (XTB; the two share one scale per column: max over both)
template <int MODE>
__global__ void __launch_bounds__(256) pgi_colmax_kernel(const double* __restrict__ In, int64_t ld_in, int64_t klen, const double* __restrict__ table,
                                                         unsigned long long* __restrict__ colmax) {      // klen: VALID K entries
  __shared__ double red[256];
  const int l = blockIdx.y;
  double mx = 0.0;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < klen; k += (int64_t)gridDim.x * 256) {
    const double x = fabs(In[(size_t)l * ld_in + k]);
    if (MODE == PGI_XTB) {
      const double t0 = table[4 * k], t2 = table[4 * k + 2];
      mx = fmax(mx, fmax(fabs(t0), fabs(0.5 * (t2 - t0))) * x);
    } else mx = fmax(mx, x);
  }
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicMax(colmax + l, (unsigned long long)__double_as_longlong(red[0]));     // non-negative doubles order like integers
}
// colscale[l] = 2^(e_l), e_l from frexp of the column maximum (value < 2^e)
__global__ void pgi_colscale_kernel(const unsigned long long* __restrict__ colmax, int ncp, double* __restrict__ colscale) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= ncp) return;
  const double mx = __longlong_as_double((long long)colmax[l]);
  int e = 0;
  if (mx > 0.0) frexp(mx, &e);
  colscale[l] = ldexp(1.0, e);
}
// D[b][(l, q)][k] = digit q of w_b(k) In[l][k] / 2^(e_l): sign x 7 bits, most significant first.  One thread = 16 consecutive k of one column.
template <int MODE>
__global__ void __launch_bounds__(256) pgi_slice_kernel(const double* __restrict__ In, int64_t ld_in, int64_t klen, int64_t kpad, int ncols, int ncp,
                                                        const double* __restrict__ table, const double* __restrict__ colscale, int8_t* __restrict__ D) {
  const int64_t k0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
  const int l = blockIdx.y;
  if (k0 >= kpad) return;
  const int nB = MODE == PGI_XTB ? 2 : 1;
  const double inv = 1.0 / colscale[l];
  for (int b = 0; b < nB; b++) {
    uint32_t dig[PGI_ND][4];
#pragma unroll
    for (int d = 0; d < PGI_ND; d++) dig[d][0] = dig[d][1] = dig[d][2] = dig[d][3] = 0;
    if (l < ncols) {
#pragma unroll
      for (int t = 0; t < 16; t++) {
        const int64_t k = k0 + t;
        double x = k < klen ? In[(size_t)l * ld_in + k] : 0.0;
        if (MODE == PGI_XTB) {
          const double t0 = table[4 * k], t2 = table[4 * k + 2];
          x *= b == 0 ? t0 : 0.5 * (t2 - t0);
        }
        const double ax = fabs(x) * inv;                                 // < 1
        unsigned long long qv = __double2ull_rn(ldexp(ax, 7 * PGI_ND));
        const unsigned long long qmax = (1ull << (7 * PGI_ND)) - 1ull;
        if (qv > qmax) qv = qmax;
        const bool neg = x < 0.0;
#pragma unroll
        for (int d = 0; d < PGI_ND; d++) {
          int mk = (int)((qv >> (7 * (PGI_ND - 1 - d))) & 127ull);
          if (neg) mk = -mk;
          dig[d][t >> 2] |= ((uint32_t)(mk & 0xFF)) << (8 * (t & 3));
        }
      }
    }
#pragma unroll
    for (int d = 0; d < PGI_ND; d++) {
      int8_t* dst = D + ((size_t)b * ncp * PGI_ND + (size_t)l * PGI_ND + d) * kpad + k0;
      *reinterpret_cast<uint4*>(dst) = make_uint4(dig[d][0], dig[d][1], dig[d][2], dig[d][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled_pgi)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int pgi_make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t pitch, int boxc, int boxr, bool swizzle) {
  static PFN_encodeTiled_pgi enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiled_pgi>(p);
  }
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return EB_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)boxc, (cuuint32_t)boxr};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (pg_i8: %lld x %lld, box %d x %d) failed: %d", (long long)rows, (long long)pitch, boxc, boxr, (int)r); return EB_ERR_CUDA; }
  return 0;
}

__global__ void __launch_bounds__(256) pgi_sum_planes_kernel(const double* __restrict__ Part, int64_t plane_stride, int nsplit, int64_t ld, int64_t rows,
                                                             double* __restrict__ Out, int64_t ld_out) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int l = blockIdx.y;
  if (r >= rows) return;
  double v = 0.0;
  for (int ks = 0; ks < nsplit; ks++) v += Part[(size_t)ks * plane_stride + (size_t)l * ld + r];
  Out[(size_t)l * ld_out + r] = v;
}

// same contract as launch_packed_gemm<MODE> of fpca_kernels.cu (mode 0 = XA: rows SNPs, 1 = XTB: rows individuals)
template <int MODE>
static int pg_i8_run(eb_ctx* c, const uint8_t* work, int64_t wpitch, int npad, const double* table, const double* In_t, int64_t ld_in, double* Out_t,
                     int64_t ld_out, int ncols, double oscale) {
  int rc;
  const int64_t rows = MODE == PGI_XTB ? npad : c->mpad;
  const int64_t klen = MODE == PGI_XTB ? c->mpad : npad;        // both multiples of 128
  const int nk = (int)(klen / PGI_BK);
  const unsigned gx = (unsigned)((rows + 127) / 128);
  int nsplit = (int)std::min<int64_t>(std::min<int64_t>(16, std::max(1, nk / 16)), (8LL * c->num_sms + gx - 1) / gx);
  nsplit = std::max(1, nsplit);
  // s32 accumulators: at most 2 x 127 + 127 per K element, so a split must stay below 2^31 / 381 = 5.6 M elements (whole-genome SNP counts)
  nsplit = std::max(nsplit, (int)((klen + (4ll << 20) - 1) / (4ll << 20)));
  CUtensorMap mapW, mapD;
  if ((rc = pgi_make_map(&mapW, work, c->mpad, wpitch, 32, 128, false))) return rc;
  const int nB = MODE == PGI_XTB ? 2 : 1;
  if ((rc = c->pgi_scale.ensure(2 * PGI_MAXC + 8))) return rc;
  unsigned long long* colmax = reinterpret_cast<unsigned long long*>(c->pgi_scale.p);
  double* colscale = c->pgi_scale.p + PGI_MAXC + 4;
  int done = 0;
  while (done < ncols) {
    const int take = std::min(PGI_MAXC, ncols - done);
    const int ncp = (take + 1) & ~1;
    const int N = ncp * PGI_ND;
    const double* in = In_t + (size_t)done * ld_in;
    double* out = Out_t + (size_t)done * ld_out;
    if ((rc = c->pgi_digits.ensure((size_t)nB * N * klen))) return rc;
    EB_CUDA(cudaMemsetAsync(colmax, 0, sizeof(unsigned long long) * PGI_MAXC, c->stream));
    const unsigned gk = (unsigned)std::min<int64_t>((klen + 255) / 256, 4 * c->num_sms);
    pgi_colmax_kernel<MODE><<<dim3(gk, take), 256, 0, c->stream>>>(in, ld_in, klen, table, colmax);
    EB_CHECK_LAUNCH(c);
    pgi_colscale_kernel<<<(ncp + 63) / 64, 64, 0, c->stream>>>(colmax, ncp, colscale);
    EB_CHECK_LAUNCH(c);
    pgi_slice_kernel<MODE><<<dim3((unsigned)((klen / 16 + 255) / 256), ncp), 256, 0, c->stream>>>(in, ld_in, klen, klen, take, ncp, table, colscale,
                                                                                                  reinterpret_cast<int8_t*>(c->pgi_digits.p));
    EB_CHECK_LAUNCH(c);
    if ((rc = pgi_make_map(&mapD, c->pgi_digits.p, (int64_t)nB * N, klen, 128, N, true))) return rc;
    PgiArgs a;
    a.ncp = ncp; a.ncols = take; a.nk = nk; a.nsplit = nsplit; a.rows = rows; a.oscale = oscale; a.table = table;
    a.tile0 = 0; a.out_row0 = 0; a.ld_l = ld_out; a.ld_r = 1;
    a.colscale = colscale;
    a.plane_stride = 0; a.out = out;
    if (nsplit > 1) {
      a.plane_stride = (int64_t)take * ld_out;
      if ((rc = c->pg_part.ensure((size_t)nsplit * a.plane_stride))) return rc;
      a.out = c->pg_part.p;
    }
    const size_t raw_bytes = (size_t)nB * N * 128 + PGI_PACKED;
    const size_t fixed = (size_t)2 * 2 * PGI_TILE_A + 1024 + 512;
    a.nraw = (int)std::max<size_t>(2, std::min<size_t>(6, (227 * 1024 - fixed) / raw_bytes));
    const size_t smem = fixed + (size_t)a.nraw * raw_bytes;
    if (smem > 227 * 1024) { set_error("pg_i8: stage layout does not fit shared memory (N %d)", N); return EB_ERR_STATE; }
    EB_CUDA(cudaFuncSetAttribute(pg_i8_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pg_i8_kernel<MODE><<<dim3(gx, nsplit), PGI_THREADS, smem, c->stream>>>(mapW, mapD, a);
    EB_CHECK_LAUNCH(c);
    if (nsplit > 1) {
      pgi_sum_planes_kernel<<<dim3((unsigned)((rows + 255) / 256), take), 256, 0, c->stream>>>(c->pg_part.p, a.plane_stride, nsplit, ld_out, rows, out, ld_out);
      EB_CHECK_LAUNCH(c);
    }
    done += take;
  }
  return 0;
}

// ---- rows = SNPs against a WIDE dense operand (shrinkmode: F_i = D Enew_i^T, one m x m matrix per eigenvector, smartpca.c:4340-4347):
// the digits of all columns are cut once (pg_i8_slice_wide), then every SNP block is one launch whose grid.z walks the column groups.
// digits: [ncp32 * 8][kpad] bytes, colscale: [ncp32], ncp32 = columns rounded up to a multiple of 32
int pg_i8_slice_wide(eb_ctx* c, const double* In_t, int64_t ld_in, int64_t kvalid, int64_t kpad, int ncols, uint8_t* digits, double* colscale) {
  int rc;
  const int ncp = (ncols + PGI_MAXC - 1) / PGI_MAXC * PGI_MAXC;
  if ((rc = c->pgi_scale.ensure((size_t)ncp + 8))) return rc;
  unsigned long long* colmax = reinterpret_cast<unsigned long long*>(c->pgi_scale.p);
  EB_CUDA(cudaMemsetAsync(colmax, 0, sizeof(unsigned long long) * ncp, c->stream));
  const unsigned gk = (unsigned)std::min<int64_t>((kvalid + 255) / 256, 8);
  for (int l0 = 0; l0 < ncols; l0 += 32768) {
    const int nl = std::min(32768, ncols - l0);
    pgi_colmax_kernel<PGI_XA><<<dim3(gk, nl), 256, 0, c->stream>>>(In_t + (size_t)l0 * ld_in, ld_in, kvalid, nullptr, colmax + l0);
    EB_CHECK_LAUNCH(c);
  }
  pgi_colscale_kernel<<<(ncp + 63) / 64, 64, 0, c->stream>>>(colmax, ncp, colscale);
  EB_CHECK_LAUNCH(c);
  for (int l0 = 0; l0 < ncp; l0 += 32768) {
    const int nl = std::min(32768, ncp - l0);
    pgi_slice_kernel<PGI_XA><<<dim3((unsigned)((kpad / 16 + 255) / 256), nl), 256, 0, c->stream>>>(
        In_t + (size_t)l0 * ld_in, ld_in, kvalid, kpad, std::max(0, std::min(nl, ncols - l0)), nl, nullptr, colscale + l0,
        reinterpret_cast<int8_t*>(digits) + (size_t)l0 * PGI_ND * kpad);
    EB_CHECK_LAUNCH(c);
  }
  return 0;
}
// out[(s - s0) * ld_r + a] = sum_t x_st In[a][t] for the SNP rows [s0, s0 + nb) (multiples of 128) and all `ncols` columns
int pg_i8_rows_wide(eb_ctx* c, const uint8_t* work, int64_t wpitch, int npad, const double* table, const uint8_t* digits, const double* colscale,
                    int ncols, int64_t s0, int nb, double* out, int64_t ld_r) {
  int rc;
  const int ncg = (ncols + PGI_MAXC - 1) / PGI_MAXC;
  const int N = PGI_MAXC * PGI_ND;
  CUtensorMap mapW, mapD;
  if ((rc = pgi_make_map(&mapW, work, c->mpad, wpitch, 32, 128, false))) return rc;
  if ((rc = pgi_make_map(&mapD, digits, (int64_t)ncg * N, npad, 128, N, true))) return rc;
  PgiArgs a;
  a.ncp = PGI_MAXC; a.ncols = ncols; a.nk = npad / PGI_BK; a.nsplit = 1; a.rows = s0 + nb; a.oscale = 1.0; a.table = table;
  a.colscale = colscale; a.tile0 = (int)(s0 / 128); a.out_row0 = s0; a.ld_l = 1; a.ld_r = ld_r; a.plane_stride = 0; a.out = out;
  const size_t raw_bytes = (size_t)N * 128 + PGI_PACKED;
  const size_t fixed = (size_t)2 * 2 * PGI_TILE_A + 1024 + 512;
  a.nraw = (int)std::max<size_t>(2, std::min<size_t>(6, (227 * 1024 - fixed) / raw_bytes));
  const size_t smem = fixed + (size_t)a.nraw * raw_bytes;
  EB_CUDA(cudaFuncSetAttribute(pg_i8_kernel<PGI_XA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pg_i8_kernel<PGI_XA><<<dim3((unsigned)(nb / 128), 1, (unsigned)ncg), PGI_THREADS, smem, c->stream>>>(mapW, mapD, a);
  EB_CHECK_LAUNCH(c);
  return 0;
}

int pg_i8_launch(eb_ctx* c, int mode, const uint8_t* work, int64_t wpitch, int npad, const double* table, const double* In_t, int64_t ld_in,
                 double* Out_t, int64_t ld_out, int ncols, double oscale) {
  if (mode == PGI_XA) return pg_i8_run<PGI_XA>(c, work, wpitch, npad, table, In_t, ld_in, Out_t, ld_out, ncols, oscale);
  return pg_i8_run<PGI_XTB>(c, work, wpitch, npad, table, In_t, ld_in, Out_t, ld_out, ncols, oscale);
}

}  // namespace eb
