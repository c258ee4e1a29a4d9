// fpca_kernels.cu -- fastmode randomised PCA (kjg_fpca, kjg_fpca.c:24-101) and the projection passes
// (smartpca.c:1485-1525) as products of the 2-bit packed matrix with skinny FP64 matrices.
//
// Replaces kjg_fpca_XTXA/_XA/_XTB (kjg_fpca.c:104-178: decode 256 SNP rows to FP64 with
// kjg_geno_get_normalized_rows, gval.c:173-213, then gsl_blas_dgemm) and LAPACKE_dgesvd (kjg_gsl.c:203).
//
// All skinny matrices are kept TRANSPOSED on the device (Mt[col][len], each logical column contiguous):
//   packed_gemm<XA>  : Out_t[l][s] = sum_i x_si In_t[l][i]      rows = SNPs,        K = individuals
//   packed_gemm<XTB> : Out_t[l][i] = sum_s x_si In_t[l][s]      rows = individuals, K = SNPs
// with x_si = table[s][code(s,i)] decoded in registers and fed to FP64 DMMA (m8n8k4); the dense operand is staged
// through shared memory.  Orthonormalisation of the M x (I+1)L sketch uses Householder QR (unconditionally stable;
// the sketch blocks differ in scale by (lambda_1/lambda_k)^I, so Gram-based QR is not an option); the final small
// SVD uses the Gram matrix of B only for the leading K triplets, where it is accurate to rounding.
#include <algorithm>
#include <cmath>
#include <vector>
#include "dmma_tile.cuh"

namespace eb {

constexpr int PG_ROWS = 256;      // output rows per CTA (8 warps x 32)
constexpr int PG_KT = 64;         // K elements per stage
constexpr int PG_LD = PG_KT + 4;  // dense tile row stride in doubles (== 4 mod 16 -> conflict-free B fragments)
// TMA ring depth: a stage of the narrow shapes (fastmode: 24 columns) is only ~1.5 us of DMMA work, so the ring runs four stages ahead
__host__ __device__ constexpr int pg_stages(int nblk) { return nblk <= 4 ? 6 : 4; }
enum { MODE_XA = 0, MODE_XTB = 1 };

__device__ __forceinline__ void dmma884f(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t pg_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pg_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pg_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pg_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pg_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pg_u32(bar)) : "memory");
}
__device__ __forceinline__ void pg_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PG_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PG_DONE_%=;\n"
      "bra PG_WAIT_%=;\n"
      "PG_DONE_%=:\n"
      "}\n" ::"r"(pg_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void pg_tma_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(pg_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(pg_u32(bar))
               : "memory");
}
__device__ __forceinline__ void pg_bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pg_u32(dst)), "l"(src),
               "r"(bytes), "r"(pg_u32(bar))
               : "memory");
}
// explicit, volatile shared loads: they keep their program order with respect to the mbarrier waits / arrives (see dmma_tile.cuh)
__device__ __forceinline__ double pg_lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pg_lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 pg_lds_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// Stage layout (bytes): dense tile In[NC][PG_LD] doubles | packed tile (XTB: [PG_KT snps][64 B = 256 individuals];
// XA: [256 snps][16 B = 64 individuals]) | XTB only: decode table [PG_KT][4] doubles.  XA keeps its table (256 rows, fixed per CTA)
// outside the ring.  Everything arrives by TMA (cp.async.bulk.tensor.2d for the two tiles, cp.async.bulk for the table) into a
// pg_stages()-deep mbarrier full / empty ring driven by thread 0, which stays two stages short of the ring depth ahead of the consumers -- the same
// structure as grm_syrk_kernel.  (Round 1 staged with per-thread cp.async into a 2-stage ring with two __syncthreads per stage:
// DMMA pipe 83 % / 69 % active for X A / X^T B.)
// The dense box is 4 doubles wider than the 64-element K window, which pads the shared-memory rows to 68 doubles (== 4 mod 16):
// conflict-free B fragments without a swizzle (the trick of dmma_tile.cuh).
// grid = (row tiles, k splits): split ks covers K stages [nk ks / nsplit, nk (ks + 1) / nsplit) and writes its own output plane
// (summed in a fixed order by pg_sum_planes_kernel), so that a few hundred row tiles still give tens of waves over the SMs.
template <int MODE, int NBLK>
__global__ void __launch_bounds__(256, 1)
packed_gemm_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapIn, const double* __restrict__ table,
                   int64_t mpad, int npad, double* __restrict__ Out_t, int64_t ld_out, int64_t plane_stride, int ncols, double oscale,
                   int nk, int nsplit) {
  constexpr int NC = NBLK * 8;
  constexpr int PG_STAGES = pg_stages(NBLK);
  constexpr int IN_BYTES = NC * PG_LD * 8;
  constexpr int W_BYTES = MODE == MODE_XTB ? PG_KT * 64 : PG_ROWS * 16;
  constexpr int T_BYTES = MODE == MODE_XTB ? PG_KT * 4 * 8 : 0;
  constexpr int STAGE_BYTES = IN_BYTES + W_BYTES + T_BYTES;
  extern __shared__ __align__(128) uint8_t pg_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pg_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PG_STAGES * STAGE_BYTES);
  uint64_t* empty = full + PG_STAGES;
  double* T_fix = reinterpret_cast<double*>(smem + PG_STAGES * STAGE_BYTES + 2 * PG_STAGES * 8);      // XA: [PG_ROWS][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, h = lane >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * PG_ROWS;          // first output row (SNP or individual)
  const int ks = blockIdx.y;
  const int kb0 = (int)(((long long)nk * ks) / nsplit), kb1 = (int)(((long long)nk * (ks + 1)) / nsplit);

  if (threadIdx.x == 0) {
    for (int s = 0; s < PG_STAGES; s++) { pg_mbar_init(full + s, 1); pg_mbar_init(empty + s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == MODE_XA) {
    for (int idx = threadIdx.x; idx < PG_ROWS * 4; idx += 256) {
      const int64_t s = r0 + (idx >> 2);
      T_fix[idx] = s < mpad ? table[s * 4 + (idx & 3)] : 0.0;
    }
  }
  __syncthreads();

  double acc[4][NBLK][2];
#pragma unroll
  for (int t = 0; t < 4; t++)
#pragma unroll
    for (int u = 0; u < NBLK; u++) acc[t][u][0] = acc[t][u][1] = 0.0;

  // producer cursor (thread 0)
  int p_kb = kb0, ahead = 0;
  uint32_t p_stage = 0, p_phase = 0;
  auto produce = [&]() {
    while (ahead < PG_STAGES - 2 && p_kb < kb1) {
      pg_mbar_wait(empty + p_stage, p_phase ^ 1);
      uint8_t* sb = smem + p_stage * STAGE_BYTES;
      pg_mbar_expect_tx(full + p_stage, STAGE_BYTES);
      const int k0 = p_kb * PG_KT;
      pg_tma_2d(sb, &mapIn, k0, 0, full + p_stage);
      if (MODE == MODE_XTB) {
        pg_tma_2d(sb + IN_BYTES, &mapW, (int)(r0 / 4), k0, full + p_stage);
        pg_bulk_1d(sb + IN_BYTES + W_BYTES, table + (size_t)k0 * 4, T_BYTES, full + p_stage);
      } else {
        pg_tma_2d(sb + IN_BYTES, &mapW, k0 / 4, (int)r0, full + p_stage);
      }
      if (++p_stage == PG_STAGES) { p_stage = 0; p_phase ^= 1; }
      ahead++; p_kb++;
    }
  };

  uint32_t stage = 0, phase = 0;
  const uint32_t sh = ((3 - (g & 3)) << 1) + (h << 3);
  for (int kb = kb0; kb < kb1; kb++) {
    if (threadIdx.x == 0) produce();
    __syncwarp();
    pg_mbar_wait(full + stage, phase);
    const uint32_t sb = pg_u32(smem + stage * STAGE_BYTES);
    const uint32_t in_s = sb, w_s = sb + IN_BYTES, t_s = sb + IN_BYTES + W_BYTES, tfix = pg_u32(T_fix);
#pragma unroll(NBLK <= 4 ? 4 : 2)
    for (int kk = 0; kk < PG_KT; kk += 4) {
      double a[4], b[NBLK];
      if (MODE == MODE_XTB) {
        const uint2 w = pg_lds_v2(w_s + (kk + q) * 64 + warp * 8);
        const uint32_t tk = t_s + (kk + q) * 32;
        const uint32_t v0 = w.x >> sh, v1 = w.y >> sh;
        a[0] = pg_lds_f64(tk + ((v0 & 3) << 3)); a[1] = pg_lds_f64(tk + (((v0 >> 16) & 3) << 3));
        a[2] = pg_lds_f64(tk + ((v1 & 3) << 3)); a[3] = pg_lds_f64(tk + (((v1 >> 16) & 3) << 3));
      } else {
        // individuals k0+kk .. +3 live in byte kk/4 of the 16-byte row segment; this thread's code is q (MSB first)
        const int byte = kk >> 2;
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int row = warp * 32 + t * 8 + g;
          const uint32_t bv = pg_lds_u8(w_s + row * 16 + byte);
          a[t] = pg_lds_f64(tfix + row * 32 + (((bv >> ((3 - q) << 1)) & 3) << 3));
        }
      }
#pragma unroll
      for (int u = 0; u < NBLK; u++) b[u] = pg_lds_f64(in_s + ((u * 8 + g) * PG_LD + kk + q) * 8);
#pragma unroll
      for (int t = 0; t < 4; t++)
#pragma unroll
        for (int u = 0; u < NBLK; u++) dmma884f(acc[t][u][0], acc[t][u][1], a[t], b[u]);
    }
    // release the stage (fence + converge + one arrive per warp, see dt_release_stage)
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) pg_mbar_arrive(empty + stage);
    if (threadIdx.x == 0) ahead--;
    if (++stage == PG_STAGES) { stage = 0; phase ^= 1; }
  }
  const int64_t rlimit = MODE == MODE_XTB ? npad : mpad;
  double* out = Out_t + (size_t)ks * plane_stride;
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int64_t r = r0 + warp * 32 + t * 8 + g;
    if (r >= rlimit) continue;
#pragma unroll
    for (int u = 0; u < NBLK; u++) {
      const int l = u * 8 + q * 2;
      if (l < ncols) out[(size_t)l * ld_out + r] = acc[t][u][0] * oscale;
      if (l + 1 < ncols) out[(size_t)(l + 1) * ld_out + r] = acc[t][u][1] * oscale;
    }
  }
}

// Out[l][r] = sum over planes (fixed order) of Part[ks][l][r]
__global__ void __launch_bounds__(256) pg_sum_planes_kernel(const double* __restrict__ Part, int64_t plane_stride, int nsplit, int64_t ld, int64_t rows,
                                                            double* __restrict__ Out, int64_t ld_out) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int l = blockIdx.y;
  if (r >= rows) return;
  double v = 0.0;
  for (int ks = 0; ks < nsplit; ks++) v += Part[(size_t)ks * plane_stride + (size_t)l * ld + r];
  Out[(size_t)l * ld_out + r] = v;
}

// view of a 2-bit working matrix [mpad][npad/4] (the PCA rows by default; lsqproj builds one over all listed individuals)
struct PackedView { const uint8_t* work; int64_t wpitch; int npad; };

PFN_encodeTiled_t get_tensormap_encoder();
static int make_u8_tensormap(CUtensorMap* map, const uint8_t* base, int64_t rows, int64_t pitch, int boxc, int boxr) {
  PFN_encodeTiled_t enc = get_tensormap_encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return EB_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)boxc, (cuuint32_t)boxr};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(u8 %lld x %lld box %d x %d) failed: %d", (long long)rows, (long long)pitch, boxc, boxr, (int)r); return EB_ERR_CUDA; }
  return 0;
}

template <int MODE>
static int launch_packed_gemm(eb_ctx* c, const double* table, const double* In_t, int64_t ld_in, double* Out_t, int64_t ld_out,
                              int ncols, double oscale, const PackedView* view = nullptr) {
  const PackedView dflt = {c->work.p, c->wpitch, c->npad};
  const PackedView pv = view ? *view : dflt;
  // integer tensor-core version with in-kernel decode (pg_i8.cu): same contract, exact digit products
  {
    const bool fits = (pv.npad % 128) == 0 && (pv.wpitch % 16) == 0 && (c->mpad % 128) == 0;
    const bool big = pv.npad >= c->opt_pg_i8_min && c->mpad >= c->opt_pg_i8_min;
    if (fits && (c->opt_pg_method == 2 || (c->opt_pg_method == 0 && big)))
      return pg_i8_launch(c, MODE == MODE_XA ? 0 : 1, pv.work, pv.wpitch, pv.npad, table, In_t, ld_in, Out_t, ld_out, ncols, oscale);
  }
  const int64_t rows = MODE == MODE_XTB ? pv.npad : c->mpad;
  const int64_t Klen = MODE == MODE_XTB ? c->mpad : pv.npad;
  const int nk = (int)(Klen / PG_KT);
  const unsigned gx = (unsigned)((rows + PG_ROWS - 1) / PG_ROWS);
  // k splits: ~24 waves of CTAs over the SMs, at least 16 stages per split
  int nsplit = (int)std::min<int64_t>(std::min<int64_t>(16, std::max(1, nk / 16)), (24LL * c->num_sms + gx - 1) / gx);
  nsplit = std::max(1, nsplit);
  if ((ld_in & 1)) { set_error("packed_gemm: leading dimension of the dense operand must be even"); return EB_ERR_ARG; }
  CUtensorMap mapW;
  int rc;
  if (MODE == MODE_XTB) { if ((rc = make_u8_tensormap(&mapW, pv.work, c->mpad, pv.wpitch, 64, PG_KT))) return rc; }
  else { if ((rc = make_u8_tensormap(&mapW, pv.work, c->mpad, pv.wpitch, 16, PG_ROWS))) return rc; }
  int done = 0;
  while (done < ncols) {
    const int rem = ncols - done;
    const int nblk = std::min(8, (rem + 7) / 8);
    const int take = std::min(rem, nblk * 8);
    const double* in = In_t + (size_t)done * ld_in;
    double* out = Out_t + (size_t)done * ld_out;
    double* dst = out; int64_t plane = 0;
    if (nsplit > 1) {
      plane = (int64_t)take * ld_out;
      if ((rc = c->pg_part.ensure((size_t)nsplit * plane))) return rc;
      dst = c->pg_part.p;
    }
    CUtensorMap mapIn;
    if ((rc = make_f64_tensormap(&mapIn, in, take, Klen, ld_in, PG_LD, nblk * 8))) return rc;
    dim3 grid(gx, nsplit);
#define PG_CASE(NB_)                                                                                                              \
  case NB_: {                                                                                                                     \
    const size_t smem = (size_t)pg_stages(NB_) * ((NB_ * 8) * PG_LD * 8 + (MODE == MODE_XTB ? PG_KT * 64 + PG_KT * 32 : PG_ROWS * 16)) + \
                        2 * pg_stages(NB_) * 8 + (MODE == MODE_XA ? PG_ROWS * 32 : 0) + 128;                                         \
    EB_CUDA(cudaFuncSetAttribute(packed_gemm_kernel<MODE, NB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    packed_gemm_kernel<MODE, NB_><<<grid, 256, smem, c->stream>>>(mapW, mapIn, table, c->mpad, pv.npad, dst, ld_out, plane, take,  \
                                                                 oscale, nk, nsplit);                                             \
  } break;
    switch (nblk) {
      PG_CASE(1) PG_CASE(2) PG_CASE(3) PG_CASE(4) PG_CASE(5) PG_CASE(6) PG_CASE(7) PG_CASE(8)
    }
#undef PG_CASE
    EB_CHECK_LAUNCH(c);
    if (nsplit > 1) {
      dim3 g2((unsigned)((rows + 255) / 256), take);
      pg_sum_planes_kernel<<<g2, 256, 0, c->stream>>>(c->pg_part.p, plane, nsplit, ld_out, rows, out, ld_out);
      EB_CHECK_LAUNCH(c);
    }
    done += take;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ per-SNP tables
// fastmode table, gval.c:73-79: mean = xmean/xfancy; gtable[k] = ((k - mean) * xfancy) / sqrt(2); missing -> 0
__global__ void fpca_table_kernel(int64_t nsnp, int64_t mpad, const double* __restrict__ xmean, const double* __restrict__ xfancy,
                                  const int* __restrict__ nmiss, double* __restrict__ table) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= mpad) return;
  double t0 = 0, t1 = 0, t2 = 0;
  if (s < nsnp && nmiss[s] >= 0) {
    const double xf = xfancy[s], mean = __ddiv_rn(xmean[s], xf), r2 = __dsqrt_rn(2.0);
    t0 = __ddiv_rn(__dmul_rn(__dadd_rn(0.0, -mean), xf), r2);
    t1 = __ddiv_rn(__dmul_rn(__dadd_rn(1.0, -mean), xf), r2);
    t2 = __ddiv_rn(__dmul_rn(__dadd_rn(2.0, -mean), xf), r2);
  }
  reinterpret_cast<double4*>(table)[s] = make_double4(t0, t1, t2, 0.0);
}

// projection table, fixxrow (qpsubs.c:338-352): g*xfancy - xmean for observed genotypes of used SNPs, else 0
__global__ void fix_table_kernel(int64_t nsnp, int64_t mpad, const double* __restrict__ xmean, const double* __restrict__ xfancy,
                                 const uint8_t* __restrict__ used, double* __restrict__ table) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= mpad) return;
  double t0 = 0, t1 = 0, t2 = 0;
  if (s < nsnp && used[s]) {
    const double xf = xfancy[s], xm = xmean[s];
    t0 = __dadd_rn(0.0, -xm); t1 = __dadd_rn(xf, -xm); t2 = __dadd_rn(__dmul_rn(2.0, xf), -xm);
  }
  reinterpret_cast<double4*>(table)[s] = make_double4(t0, t1, t2, 0.0);
}

// ------------------------------------------------------------------------------------------ Householder QR on transposed storage
__device__ __forceinline__ double bsum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; i++) r += sh[i];
  return r;
}

// reflector j from row j of At (positions j..len-1); v stored in place with v[j] = 1; tau[j], diag[j] = beta
__global__ void __launch_bounds__(1024) qr_reflector_kernel(double* __restrict__ At, int64_t ld, int64_t len, int j, double* __restrict__ tau,
                                                            double* __restrict__ diag) {
  __shared__ double sh[32];
  double* row = At + (size_t)j * ld;
  double s = 0.0;
  for (int64_t i = j + 1 + threadIdx.x; i < len; i += blockDim.x) s += row[i] * row[i];
  s = bsum(s, sh);
  const double alpha = row[j];
  double beta = alpha, tj = 0.0, scal = 0.0;
  if (s != 0.0) {
    const double nrm = sqrt(alpha * alpha + s);
    beta = alpha >= 0.0 ? -nrm : nrm;
    tj = (beta - alpha) / beta;
    scal = 1.0 / (alpha - beta);
  }
  __syncthreads();
  for (int64_t i = j + 1 + threadIdx.x; i < len; i += blockDim.x) row[i] *= scal;
  if (threadIdx.x == 0) { row[j] = 1.0; tau[j] = tj; diag[j] = beta; }
}

// apply H = I - tau v v^T (v = Vt row jv, support [jv, len)) to rows [r_first, r_first + gridDim.x) of Ct
__global__ void __launch_bounds__(1024) qr_apply_kernel(const double* __restrict__ Vt, int64_t ldv, int jv, const double* __restrict__ tau,
                                                        double* __restrict__ Ct, int64_t ldc, int r_first, int64_t len) {
  __shared__ double sh[32];
  const double tj = tau[jv];
  if (tj == 0.0) return;
  const double* v = Vt + (size_t)jv * ldv;
  double* z = Ct + (size_t)(r_first + blockIdx.x) * ldc;
  double s = 0.0;
  for (int64_t i = jv + threadIdx.x; i < len; i += blockDim.x) s += v[i] * z[i];
  s = bsum(s, sh) * tj;
  for (int64_t i = jv + threadIdx.x; i < len; i += blockDim.x) z[i] -= s * v[i];
}

__global__ void set_identity_rows_kernel(double* __restrict__ Ut, int64_t ld, int c) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < c) Ut[(size_t)r * ld + r] = 1.0;
}

// Gram matrix of the rows of Bt: G[a][b] = sum_i Bt[a][i] Bt[b][i]  (one block per pair a >= b, fixed-order reduction)
__global__ void __launch_bounds__(256) gram_rows_kernel(const double* __restrict__ Bt, int64_t ld, int64_t len, int c, double* __restrict__ G) {
  __shared__ double sh[32];
  const int a = blockIdx.y, b = blockIdx.x;
  if (b > a) return;
  const double* ra = Bt + (size_t)a * ld;
  const double* rb = Bt + (size_t)b * ld;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) s += ra[i] * rb[i];
  s = bsum(s, sh);
  if (threadIdx.x == 0) { G[(size_t)a * c + b] = s; G[(size_t)b * c + a] = s; }
}

// U_t[k][i] = (sum_l V[k][l] Bt[l][i]) / sigma_k   ;  sigma_k = sqrt(lam[k])
__global__ void __launch_bounds__(256) left_vectors_kernel(const double* __restrict__ Bt, int64_t ld, int64_t len, int c, const double* __restrict__ V,
                                                           const double* __restrict__ lam, int K, double* __restrict__ Ut) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= len) return;
  double s = 0.0;
  for (int l = 0; l < c; l++) s += V[(size_t)k * c + l] * Bt[(size_t)l * ld + i];
  Ut[(size_t)k * len + i] = s / sqrt(lam[k]);
}

__global__ void col_sumsq_kernel(const double* __restrict__ At, int64_t ld, int64_t len, double* __restrict__ out) {
  __shared__ double sh[32];
  const double* r = At + (size_t)blockIdx.x * ld;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) s += r[i] * r[i];
  s = bsum(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

static int qr_orthonormal_rows(eb_ctx* c, double* At, int64_t ld, int64_t len, int ncol, double* Ut, double* tau_d, double* diag_d) {
  // At (ncol rows of length len) is overwritten by its reflectors; Ut receives an orthonormal basis of span(At rows)
  for (int j = 0; j < ncol && j < len; j++) {
    qr_reflector_kernel<<<1, 1024, 0, c->stream>>>(At, ld, len, j, tau_d, diag_d);
    EB_CHECK_LAUNCH(c);
    if (j + 1 < ncol) {
      qr_apply_kernel<<<ncol - j - 1, 1024, 0, c->stream>>>(At, ld, j, tau_d, At, ld, j + 1, len);
      EB_CHECK_LAUNCH(c);
    }
  }
  EB_CUDA(cudaMemsetAsync(Ut, 0, sizeof(double) * (size_t)ncol * ld, c->stream));
  set_identity_rows_kernel<<<(ncol + 127) / 128, 128, 0, c->stream>>>(Ut, ld, ncol);
  EB_CHECK_LAUNCH(c);
  for (int j = std::min<int64_t>(ncol, len) - 1; j >= 0; j--) {
    qr_apply_kernel<<<ncol - j, 1024, 0, c->stream>>>(At, ld, j, tau_d, Ut, ld, j, len);
    EB_CHECK_LAUNCH(c);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ TSQR pieces (SNP-sharded fastmode)
// St[l][col0 + j] = R[j][l] of the local QR (At after qr_reflector: R above the diagonal in place, diagonal in diag)
__global__ void extract_r_kernel(const double* __restrict__ At, int64_t ld, const double* __restrict__ diag, int ncol,
                                 double* __restrict__ St, int64_t lds, int col0) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y;
  if (j >= ncol) return;
  St[(size_t)l * lds + col0 + j] = j < l ? At[(size_t)l * ld + j] : (j == l ? diag[l] : 0.0);
}

// Out[l][s] = sum_j C[l][j] In[j][s]   (C: ncol x ncol with leading dimension ldc; In/Out: ncol rows of length len)
__global__ void __launch_bounds__(256) small_left_mult_kernel(const double* __restrict__ C, int64_t ldc, const double* __restrict__ In,
                                                              int64_t ldi, double* __restrict__ Out, int64_t ldo, int ncol, int64_t len) {
  extern __shared__ double Cs[];          // [8][ncol]
  const int l0 = blockIdx.y * 8;
  for (int idx = threadIdx.x; idx < 8 * ncol; idx += blockDim.x) {
    const int ll = idx / ncol, j = idx - ll * ncol;
    Cs[idx] = l0 + ll < ncol ? C[(size_t)(l0 + ll) * ldc + j] : 0.0;
  }
  __syncthreads();
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= len) return;
  double acc[8];
#pragma unroll
  for (int ll = 0; ll < 8; ll++) acc[ll] = 0.0;
  for (int j = 0; j < ncol; j++) {
    const double x = In[(size_t)j * ldi + s];
#pragma unroll
    for (int ll = 0; ll < 8; ll++) acc[ll] = fma(Cs[ll * ncol + j], x, acc[ll]);
  }
#pragma unroll
  for (int ll = 0; ll < 8; ll++)
    if (l0 + ll < ncol) Out[(size_t)(l0 + ll) * ldo + s] = acc[ll];
}

// ------------------------------------------------------------------------------------------ fastmode driver
// With a communicator (eb_set_comm) the SNPs of X are sharded over the ranks: Q_i = X G is local (the rows of
// Q live with the shard), G = X^T Q_i / m and B = X^T Q are sums of per-shard partials -> one in-place all-reduce over
// peer memory each, and the orthonormalisation of the row-sharded sketch is a TSQR (local Householder QR, all-reduce of
// the stacked cw x cw R factors, QR of the stack, Q_g <- U_g U2_g).  Everything after B is replicated.
int fpca_run(eb_ctx* c, int fancynorm, int altnormstyle, size_t K, size_t L, size_t I, long seed, double* eval, double* evec) {
  int rc;
  const int64_t m = c->nsnp, mpad = c->mpad;
  const int n = c->nrows, npad = c->npad;
  const int cw = (int)((I + 1) * L);
  const bool shard = c->has_comm;
  const int W = shard ? c->comm.world : 1, me = shard ? c->comm.rank : 0;
  int64_t m_total = m;
  if (shard) {
    std::vector<long long> all(W);
    long long mine[1] = {(long long)m};
    if ((rc = peer_allgather_host(c, mine, all.data(), sizeof(long long)))) return rc;
    m_total = 0;
    for (long long v : all) m_total += v;
  }
  if ((int64_t)cw > m || cw > n) {
    set_error("eb_fpca: (I+1)*L = %d exceeds the matrix dimensions (%lld SNPs%s x %d)", cw, (long long)m, shard ? " in this shard" : "", n);
    return EB_ERR_ARG;
  }
  // per-SNP statistics over the current rows (no drop rule: every uploaded SNP stays a row of X, gval.c:56-86)
  eb_grm_opts o = {fancynorm, altnormstyle, 0, 2147483647, nullptr, nullptr};
  if ((rc = launch_stats(c, &o))) return rc;
  c->grm_valid = false;
  DevBuf<double> ftab, Qt, Ut, tau, diag, gram, uk, U2t;
  const int64_t lds = ((int64_t)W * cw + 1) & ~1ll;
  if ((rc = ftab.ensure((size_t)mpad * 4)) || (rc = c->fpG.ensure((size_t)L * npad)) || (rc = Qt.ensure((size_t)cw * mpad)) ||
      (rc = Ut.ensure((size_t)cw * mpad)) || (rc = c->fpB.ensure((size_t)cw * npad)) || (rc = tau.ensure(cw)) || (rc = diag.ensure(cw)) ||
      (rc = gram.ensure((size_t)cw * cw)) || (rc = uk.ensure((size_t)K * n)))
    return rc;
  if (shard && ((rc = c->fpS.ensure((size_t)cw * lds)) || (rc = U2t.ensure((size_t)cw * lds)))) return rc;
  double* Gt = c->fpG.p;
  double* Bt = c->fpB.p;
  fpca_table_kernel<<<(unsigned)((mpad + 255) / 256), 256, 0, c->stream>>>(m, mpad, c->xmean_d.p, c->xfancy_d.p, c->nmiss_d.p, ftab.p);
  EB_CHECK_LAUNCH(c);
  // G1 <- seeded Gaussians (host RNG, bit-exact with kjg_gsl.c:145-186), uploaded transposed; identical on every rank
  {
    std::vector<double> G((size_t)n * L), T((size_t)L * npad, 0.0);
    eb_gauss_matrix(seed, (size_t)n, L, G.data());
    for (int i = 0; i < n; i++) for (size_t l = 0; l < L; l++) T[l * npad + i] = G[(size_t)i * L + l];
    EB_CUDA(cudaMemcpyAsync(Gt, T.data(), sizeof(double) * T.size(), cudaMemcpyHostToDevice, c->stream));
    EB_CUDA(cudaStreamSynchronize(c->stream));
  }
  EB_CUDA(cudaMemsetAsync(Qt.p, 0, sizeof(double) * (size_t)cw * mpad, c->stream));
  const double inv_m = 1.0 / (double)m_total;
  for (size_t it = 0; it <= I; it++) {
    double* Qi = Qt.p + (size_t)it * L * mpad;
    if ((rc = launch_packed_gemm<MODE_XA>(c, ftab.p, Gt, npad, Qi, mpad, (int)L, 1.0))) return rc;      // Q_i = X G       (kjg_fpca.c:121,148)
    if (it == I) break;
    if ((rc = launch_packed_gemm<MODE_XTB>(c, ftab.p, Qi, mpad, Gt, npad, (int)L, inv_m))) return rc;    // G = X^T Q_i / m (kjg_fpca.c:123,53)
    if (shard && (rc = peer_allreduce(c, PEER_SLOT_A, Gt, c->fpG.n, (int64_t)L * npad))) return rc;
  }
  // Q <- orthonormal basis of its column span (kjg_fpca.c:63-71 keeps U of the SVD; any orthonormal basis of the same
  // span gives the same B B^T and therefore the same leading singular triplets)
  if ((rc = qr_orthonormal_rows(c, Qt.p, mpad, m, cw, Ut.p, tau.p, diag.p))) return rc;
  const double* Qfinal = Ut.p;
  if (shard) {
    double* St = c->fpS.p;
    EB_CUDA(cudaMemsetAsync(St, 0, sizeof(double) * (size_t)cw * lds, c->stream));
    {
      dim3 grid((cw + 127) / 128, cw);
      extract_r_kernel<<<grid, 128, 0, c->stream>>>(Qt.p, mpad, diag.p, cw, St, lds, me * cw);
      EB_CHECK_LAUNCH(c);
    }
    if ((rc = peer_allreduce(c, PEER_SLOT_C, St, c->fpS.n, (int64_t)cw * lds))) return rc;     // every rank: the stacked R factors
    if ((rc = qr_orthonormal_rows(c, St, lds, (int64_t)W * cw, cw, U2t.p, tau.p, diag.p))) return rc;
    {
      dim3 grid((unsigned)((m + 255) / 256), (cw + 7) / 8);
      small_left_mult_kernel<<<grid, 256, sizeof(double) * 8 * cw, c->stream>>>(U2t.p + (size_t)me * cw, lds, Ut.p, mpad, Qt.p, mpad, cw, m);
      EB_CHECK_LAUNCH(c);
    }
    Qfinal = Qt.p;       // the reflectors are no longer needed; pad columns [m, mpad) meet all-zero table rows in X^T Q
  }
  // B = X^T Q   (kjg_fpca.c:79-80)
  if ((rc = launch_packed_gemm<MODE_XTB>(c, ftab.p, Qfinal, mpad, Bt, npad, cw, 1.0))) return rc;
  if (shard && (rc = peer_allreduce(c, PEER_SLOT_B, Bt, c->fpB.n, (int64_t)cw * npad))) return rc;
  // leading K left singular vectors / values of B through the cw x cw Gram matrix
  {
    dim3 grid(cw, cw);
    gram_rows_kernel<<<grid, 256, 0, c->stream>>>(Bt, npad, n, cw, gram.p);
    EB_CHECK_LAUNCH(c);
  }
  std::vector<double> lam(cw), V((size_t)K * cw);
  if ((rc = eig_resident(c, gram.p, cw, cw, 1.0, (int)K, lam.data(), V.data()))) return rc;
  {
    dim3 grid((n + 255) / 256, (unsigned)K);
    left_vectors_kernel<<<grid, 256, 0, c->stream>>>(Bt, npad, n, cw, c->zvec_d.p, c->lambda_d.p, (int)K, uk.p);
    EB_CHECK_LAUNCH(c);
  }
  std::vector<double> U((size_t)K * n);
  EB_CUDA(cudaMemcpyAsync(U.data(), uk.p, sizeof(double) * U.size(), cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t k = 0; k < K; k++) {
    eval[k] = lam[k] * (1.0 / (double)m_total);                // S^2 / m, kjg_fpca.c:91-95
    for (int i = 0; i < n; i++) evec[(size_t)i * K + k] = U[k * n + i];
  }
  if (shard && (rc = peer_bury(c))) return rc;
  return 0;
}

__global__ void mm_table_kernel(int64_t nsnp, int64_t mpad, int nrows, const int* __restrict__ c0, const int* __restrict__ nmiss,
                                const double* __restrict__ xfancy, const uint8_t* __restrict__ used, double* __restrict__ table);

// ------------------------------------------------------------------------------------------ projections (smartpca.c:1485-1525)
int project_run(eb_ctx* c, const double* evecs, int numeigs, double* ffvecs, double* fxvecs, double* fxscal) {
  int rc;
  const int64_t m = c->nsnp, mpad = c->mpad;
  const int n = c->nrows, npad = c->npad;
  DevBuf<double> Ft, FFt, FXt, ftab, ss;
  if ((rc = Ft.ensure((size_t)numeigs * npad)) || (rc = FFt.ensure((size_t)numeigs * mpad)) || (rc = FXt.ensure((size_t)numeigs * npad)) ||
      (rc = ftab.ensure((size_t)mpad * 4)) || (rc = ss.ensure(numeigs)))
    return rc;
  // fvecs = 10 * evecs (setfvecs, smartpca.c:1444)
  std::vector<double> T((size_t)numeigs * npad, 0.0);
  for (int j = 0; j < numeigs; j++) for (int i = 0; i < n; i++) T[(size_t)j * npad + i] = 10.0 * evecs[(size_t)j * n + i];
  EB_CUDA(cudaMemcpyAsync(Ft.p, T.data(), sizeof(double) * T.size(), cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaMemsetAsync(FFt.p, 0, sizeof(double) * (size_t)numeigs * mpad, c->stream));
  // ffvecs[j][s] = sum_k fvecs[j][k] x_ks with the columns of getcolxf (smartpca.c:1487, 3564-3597): (g - ymean) * yfancy WITHOUT the
  // SNP weight (weightname only enters the GRM columns, smartpca.c:1178-1180); dropped / ignored SNPs contribute zero columns
  mm_table_kernel<<<(unsigned)((mpad + 255) / 256), 256, 0, c->stream>>>(m, mpad, n, c->c0_d.p, c->nmiss_d.p, c->xfancy_d.p, c->used_d.p, ftab.p);
  EB_CHECK_LAUNCH(c);
  if ((rc = launch_packed_gemm<MODE_XA>(c, ftab.p, Ft.p, npad, FFt.p, mpad, numeigs, 1.0))) return rc;
  fix_table_kernel<<<(unsigned)((mpad + 255) / 256), 256, 0, c->stream>>>(m, mpad, c->xmean_d.p, c->xfancy_d.p, c->used_d.p, ftab.p);
  EB_CHECK_LAUNCH(c);
  if ((rc = launch_packed_gemm<MODE_XTB>(c, ftab.p, FFt.p, mpad, FXt.p, npad, numeigs, 1.0))) return rc;
  if ((rc = peer_allreduce_any(c, FXt.p, (int64_t)numeigs * npad))) return rc;      // SNP shards: sum of the per-shard projections
  col_sumsq_kernel<<<numeigs, 256, 0, c->stream>>>(FXt.p, npad, n, ss.p);
  EB_CHECK_LAUNCH(c);
  std::vector<double> s(numeigs);
  EB_CUDA(cudaMemcpyAsync(s.data(), ss.p, sizeof(double) * numeigs, cudaMemcpyDeviceToHost, c->stream));
  if (ffvecs) EB_CUDA(cudaMemcpy2DAsync(ffvecs, sizeof(double) * m, FFt.p, sizeof(double) * mpad, sizeof(double) * m, numeigs, cudaMemcpyDeviceToHost, c->stream));
  if (fxvecs) EB_CUDA(cudaMemcpy2DAsync(fxvecs, sizeof(double) * n, FXt.p, sizeof(double) * npad, sizeof(double) * n, numeigs, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  if (fxscal) for (int j = 0; j < numeigs; j++) fxscal[j] = 1.0 / sqrt(s[j]);
  return 0;
}

// ------------------------------------------------------------------------------------------ lsqproj (smartpca.c:4606-4757)
// mask table: 1 for an observed genotype of a used SNP, else 0 (rows of the normal equations, smartpca.c:4695-4706)
__global__ void mask_table_kernel(int64_t nsnp, int64_t mpad, const uint8_t* __restrict__ used, double* __restrict__ table) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= mpad) return;
  const double v = (s < nsnp && used[s]) ? 1.0 : 0.0;
  reinterpret_cast<double4*>(table)[s] = make_double4(v, v, v, 0.0);
}
// E[j][s] = fxscal[j] * ffvecs[j][s] (emat, smartpca.c:4702), zero in the pad
__global__ void lsq_scale_kernel(const double* __restrict__ FFt, int64_t ld, int64_t nsnp, int64_t mpad, const double* __restrict__ fxscal,
                                 double* __restrict__ Et) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (s >= mpad) return;
  Et[(size_t)j * ld + s] = s < nsnp ? fxscal[j] * FFt[(size_t)j * ld + s] : 0.0;
}
// pair rows: Pt[pair(a,b)][s] = E[a][s] E[b][s] for a <= b, plus one all-ones row (-> number of valid SNPs per individual)
__global__ void lsq_pairs_kernel(const double* __restrict__ Et, int64_t ld, int64_t nsnp, int64_t mpad, int k, double* __restrict__ Pt) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= mpad) return;
  int pr = 0;
  for (int a = 0; a < k; a++) {
    const double ea = Et[(size_t)a * ld + s];
    for (int b = a; b < k; b++) Pt[(size_t)(pr++) * ld + s] = ea * Et[(size_t)b * ld + s];
  }
  Pt[(size_t)pr * ld + s] = s < nsnp ? 1.0 : 0.0;
}
// one thread per individual: normal equations -> choldc / cholsl in the reference's operation order (linsubs.c:331-393)
constexpr int LSQ_KMAX = 32;       // numeigs of eb_lsqproj / eb_evec_coords (documented in eigb200.h); the k x k system lives in local memory
__global__ void __launch_bounds__(128) lsq_solve_kernel(const double* __restrict__ Nt, const double* __restrict__ Rt, int64_t ld, int nlist, int k,
                                                        double* __restrict__ At, int* __restrict__ nvalid, uint8_t* __restrict__ ok) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nlist) return;
  double a[LSQ_KMAX * LSQ_KMAX], p[LSQ_KMAX], x[LSQ_KMAX];
  int pr = 0;
  for (int i = 0; i < k; i++)
    for (int j = i; j < k; j++) { const double v = Nt[(size_t)(pr++) * ld + q]; a[i * k + j] = v; a[j * k + i] = v; }
  const int kk = (int)(Nt[(size_t)pr * ld + q] + 0.5);
  nvalid[q] = kk;
  bool good = kk > k;
  if (good) {
    for (int i = 0; i < k && good; i++)
      for (int j = i; j < k; j++) {
        double sum = a[i * k + j];
        for (int m = i - 1; m >= 0; m--) sum -= a[i * k + m] * a[j * k + m];
        if (i == j) { if (sum <= 0.0) { good = false; break; } p[i] = sqrt(sum); }
        else a[j * k + i] = sum / p[i];
      }
  }
  if (good) {
    for (int i = 0; i < k; i++) { double sum = Rt[(size_t)i * ld + q]; for (int m = i - 1; m >= 0; m--) sum -= a[i * k + m] * x[m]; x[i] = sum / p[i]; }
    for (int i = k - 1; i >= 0; i--) { double sum = x[i]; for (int m = i + 1; m < k; m++) sum -= a[m * k + i] * x[m]; x[i] = sum / p[i]; }
  }
  for (int i = 0; i < k; i++) At[(size_t)i * ld + q] = good ? x[i] : 0.0;
  ok[q] = good ? 1 : 0;
}

int lsqproj_run(eb_ctx* c, const int* indiv, int nlist, const double* ffvecs, const double* fxscal, int k, double* acoeffs, double* bcoeffs,
                int* nvalid, uint8_t* ok) {
  if (k < 1 || k > LSQ_KMAX) { set_error("eb_lsqproj: numeigs must be in 1..%d", LSQ_KMAX); return EB_ERR_ARG; }
  int rc;
  const int64_t m = c->nsnp, mpad = c->mpad;
  const int npad2 = (nlist + 255) / 256 * 256;
  const int64_t wp2 = npad2 / 4;
  const int npairs = k * (k + 1) / 2 + 1;
  DevBuf<uint8_t> work2, ok_d;
  DevBuf<int> list_d, nv_d;
  DevBuf<double> FFt, Et, Pt, ftab, mtab, Rt, Nt, At, sc_d;
  if ((rc = work2.ensure((size_t)mpad * wp2)) || (rc = list_d.ensure(nlist)) || (rc = FFt.ensure((size_t)k * mpad)) || (rc = Et.ensure((size_t)k * mpad)) ||
      (rc = Pt.ensure((size_t)npairs * mpad)) || (rc = ftab.ensure((size_t)mpad * 4)) || (rc = mtab.ensure((size_t)mpad * 4)) ||
      (rc = Rt.ensure((size_t)k * npad2)) || (rc = Nt.ensure((size_t)npairs * npad2)) || (rc = At.ensure((size_t)k * npad2)) ||
      (rc = sc_d.ensure(k)) || (rc = nv_d.ensure(npad2)) || (rc = ok_d.ensure(npad2)))
    return rc;
  for (int i = 0; i < nlist; i++)
    if (indiv[i] < 0 || indiv[i] >= c->numindivs) { set_error("eb_lsqproj: individual index %d out of range", indiv[i]); return EB_ERR_ARG; }
  EB_CUDA(cudaMemcpyAsync(list_d.p, indiv, sizeof(int) * nlist, cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaMemsetAsync(FFt.p, 0, sizeof(double) * (size_t)k * mpad, c->stream));
  EB_CUDA(cudaMemcpy2DAsync(FFt.p, sizeof(double) * mpad, ffvecs, sizeof(double) * m, sizeof(double) * m, k, cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaMemcpyAsync(sc_d.p, fxscal, sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
  if ((rc = launch_gather_into(c, list_d.p, nlist, work2.p, wp2))) return rc;
  const unsigned gm = (unsigned)((mpad + 255) / 256);
  fix_table_kernel<<<gm, 256, 0, c->stream>>>(m, mpad, c->xmean_d.p, c->xfancy_d.p, c->used_d.p, ftab.p);
  EB_CHECK_LAUNCH(c);
  mask_table_kernel<<<gm, 256, 0, c->stream>>>(m, mpad, c->used_d.p, mtab.p);
  EB_CHECK_LAUNCH(c);
  lsq_scale_kernel<<<dim3(gm, k), 256, 0, c->stream>>>(FFt.p, mpad, m, mpad, sc_d.p, Et.p);
  EB_CHECK_LAUNCH(c);
  lsq_pairs_kernel<<<gm, 256, 0, c->stream>>>(Et.p, mpad, m, mpad, k, Pt.p);
  EB_CHECK_LAUNCH(c);
  const PackedView pv = {work2.p, wp2, npad2};
  // rr[j][q] = sum_s x_qs e_sj (also bcoeffs, smartpca.c:4743-4746);  co[(a,b)][q] = sum_s m_qs e_sa e_sb
  if ((rc = launch_packed_gemm<MODE_XTB>(c, ftab.p, Et.p, mpad, Rt.p, npad2, k, 1.0, &pv))) return rc;
  if ((rc = launch_packed_gemm<MODE_XTB>(c, mtab.p, Pt.p, mpad, Nt.p, npad2, npairs, 1.0, &pv))) return rc;
  if ((rc = peer_allreduce_any(c, Rt.p, (int64_t)k * npad2)) || (rc = peer_allreduce_any(c, Nt.p, (int64_t)npairs * npad2))) return rc;
  lsq_solve_kernel<<<(nlist + 127) / 128, 128, 0, c->stream>>>(Nt.p, Rt.p, npad2, nlist, k, At.p, nv_d.p, ok_d.p);
  EB_CHECK_LAUNCH(c);
  if (acoeffs) EB_CUDA(cudaMemcpy2DAsync(acoeffs, sizeof(double) * nlist, At.p, sizeof(double) * npad2, sizeof(double) * nlist, k, cudaMemcpyDeviceToHost, c->stream));
  if (bcoeffs) EB_CUDA(cudaMemcpy2DAsync(bcoeffs, sizeof(double) * nlist, Rt.p, sizeof(double) * npad2, sizeof(double) * nlist, k, cudaMemcpyDeviceToHost, c->stream));
  if (nvalid) EB_CUDA(cudaMemcpyAsync(nvalid, nv_d.p, sizeof(int) * nlist, cudaMemcpyDeviceToHost, c->stream));
  if (ok) EB_CUDA(cudaMemcpyAsync(ok, ok_d.p, nlist, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------ shrinkmode
// doshrinkp / doshrinkp2 (smartpca.c:4223-4419, 4022-4220): leave-one-out ("shrunk") coordinates of every PCA sample.
// For eigenvector i and sample a the reference perturbs the normalised GRM X by removing a's row and column
// (dd = -(e_a x_a^T + x_a e_a^T), smartpca.c:4332-4338), applies first-order perturbation theory over the FULL eigenbasis
// (4339-4347), re-centres / re-normalises the perturbed vector with entry a zeroed (4359-4364), recomputes the SNP loadings
// from the other samples (4366-4368) and re-projects sample a by least squares on its observed genotypes (doproj, 3986-4019).
// The reference does this with an m x m scratch matrix per (i, a) and a dense m x n FP64 copy of the data:
// O(k m (m^2 + m n)) scalar flops.  Here, per eigenvector i:
//   S_i[k'][a] = (e_k' . ww_a) / (lam_i - lam_k')   in closed form (ww_a has only the row/column-a structure)
//   Enew_i     = rows { norme0_a(e_i + S_i[:,a]^T E) }              one m x m x m DMMA GEMM + a row kernel
//   F_i[s][a]  = sum_t x_ts Enew_i[a][t]                            DMMA GEMM per SNP block against the decoded block
//   per-sample sums over s (all / observed only) of F_i^2, F_i x_a, F_i F_l or F_i ff_l       reduction kernel
// and a k x k Cholesky per sample.  mmat is never materialised beyond one SNP block.
constexpr int SHR_KMAX = 32;

// (g - ymean) * yfancy of getcolxf (smartpca.c:3564-3597 via fvadjust 2236-2279): no SNP weight; zero rows for unused SNPs
__global__ void mm_table_kernel(int64_t nsnp, int64_t mpad, int nrows, const int* __restrict__ c0, const int* __restrict__ nmiss,
                                const double* __restrict__ xfancy, const uint8_t* __restrict__ used, double* __restrict__ table) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= mpad) return;
  double t0 = 0, t1 = 0, t2 = 0;
  if (s < nsnp && used[s]) {
    const double ym = __ddiv_rn((double)c0[s], (double)(nrows - nmiss[s])), yf = xfancy[s];
    t0 = __dmul_rn(__dadd_rn(0.0, -ym), yf); t1 = __dmul_rn(__dadd_rn(1.0, -ym), yf); t2 = __dmul_rn(__dadd_rn(2.0, -ym), yf);
  }
  reinterpret_cast<double4*>(table)[s] = make_double4(t0, t1, t2, 0.0);
}

// Xn = G * scale restricted to n x n -> X (ld), plus x1[a] = sum_t Xn[a][t] and dg[a] = Xn[a][a]   (one block per row)
__global__ void __launch_bounds__(256) shr_scale_rows_kernel(const double* __restrict__ G, int64_t ldg, int n, double scale, double* __restrict__ X,
                                                             int64_t ldx, double* __restrict__ x1, double* __restrict__ dg) {
  __shared__ double sh[32];
  const int a = blockIdx.x;
  double s = 0.0;
  for (int t = threadIdx.x; t < ldx; t += blockDim.x) {
    const double v = t < n ? G[(size_t)a * ldg + t] * scale : 0.0;
    X[(size_t)a * ldx + t] = v;
    s += v;
  }
  s = bsum(s, sh);
  if (threadIdx.x == 0) { x1[a] = s; dg[a] = G[(size_t)a * ldg + a] * scale; }
}

// norme (smartpca.c:3955-3963) on row blockIdx.x of E: subtract the mean, scale to unit length; mu/sc remember the map
// e = sc * e~ + mu so that X e~ can be written with the eigen-relation of the original vector
__global__ void __launch_bounds__(256) shr_norme_rows_kernel(double* __restrict__ E, int64_t ld, int n, double* __restrict__ mu, double* __restrict__ sc) {
  __shared__ double sh[32];
  double* r = E + (size_t)blockIdx.x * ld;
  double s = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) s += r[t];
  const double mean = bsum(s, sh) / (double)n;
  double q = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) { const double v = r[t] - mean; q += v * v; }
  const double nrm = sqrt(bsum(q, sh));
  for (int t = threadIdx.x; t < n; t += blockDim.x) r[t] = (r[t] - mean) * (1.0 / nrm);
  if (threadIdx.x == 0) { mu[blockIdx.x] = mean; sc[blockIdx.x] = nrm; }
}

// S[k'][a] for eigenvector i (see the derivation above); also delta / ymul per sample (smartpca.c:4339, 4348-4357 / 4143-4152)
__global__ void __launch_bounds__(256) shr_coeff_kernel(const double* __restrict__ E, int64_t ld, int n, int k, int i, const double* __restrict__ lam,
                                                        const double* __restrict__ mu, const double* __restrict__ sc, const double* __restrict__ x1,
                                                        const double* __restrict__ dg, int newshrink, double* __restrict__ S,
                                                        double* __restrict__ ymul) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, kk = blockIdx.y;
  if (a >= ld) return;
  if (a >= n) { S[(size_t)kk * ld + a] = 0.0; return; }
  const double li = lam[i], lk = lam[kk];
  const double ei = E[(size_t)i * ld + a], ek = E[(size_t)kk * ld + a], xaa = dg[a];
  const double Pi = (li * (sc[i] * ei + mu[i]) - mu[i] * x1[a]) / sc[i];                    // (X e~_i)[a]
  const double Pk = kk < k ? (lk * (sc[kk] * ek + mu[kk]) - mu[kk] * x1[a]) / sc[kk] : lk * ek;
  const double eco = -ei * (Pk - xaa * ek) - ek * Pi;
  S[(size_t)kk * ld + a] = kk == i ? 0.0 : eco / (li - lk);
  if (kk == i) {
    const double delta = ei * (xaa * ei - 2.0 * Pi);
    const bool good = newshrink ? (li > -delta) : (li > delta);
    ymul[a] = good ? li / (li + delta) : 1.0;
  }
}

// row a of C holds ediff; -> enew = norme(centre_a(e~_i + ediff)) (smartpca.c:4359-4364), zero in the pad columns
__global__ void __launch_bounds__(256) shr_enew_kernel(double* __restrict__ C, int64_t ld, int n, const double* __restrict__ ei) {
  __shared__ double sh[32];
  const int a = blockIdx.x;
  double* r = C + (size_t)a * ld;
  double s = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) { const double v = t == a ? 0.0 : ei[t] + r[t]; r[t] = v; s += v; }
  const double y = bsum(s, sh) / (double)(n - 1);
  double s2 = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) { const double v = t == a ? 0.0 : r[t] - y; r[t] = v; s2 += v; }
  const double mean = bsum(s2, sh) / (double)n;
  double q = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) { const double v = r[t] - mean; q += v * v; }
  const double inv = 1.0 / sqrt(bsum(q, sh));
  for (int t = threadIdx.x; t < ld; t += blockDim.x) r[t] = t < n ? (r[t] - mean) * inv : 0.0;
}

// D[s - s0][t] = table[s][code(s, t)] for the SNP block [s0, s0 + nb)
__global__ void __launch_bounds__(256) shr_decode_kernel(const uint8_t* __restrict__ work, int64_t wpitch, const double* __restrict__ table, int64_t s0,
                                                         int npad, double* __restrict__ D) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = s0 + blockIdx.y;
  if (t >= npad) return;
  const int code = (work[s * wpitch + (t >> 2)] >> ((3 - (t & 3)) * 2)) & 3;
  D[(size_t)blockIdx.y * npad + t] = table[s * 4 + code];
}

// ff[j][s] *= scale[j]   (the 1/sqrt(mean square) of smartpca.c:4309-4310, sums taken over all SNP shards on the host)
__global__ void __launch_bounds__(256) shr_scale_rows2_kernel(double* __restrict__ FF, int64_t ld, int64_t len, const double* __restrict__ scale) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < len) FF[(size_t)blockIdx.y * ld + t] *= scale[blockIdx.y];
}

// Per (sample a, eigenvector i = blockIdx.y, SNP slice blockIdx.z): sums over the SNPs of this block of
//   slot l < k : NEW  F_i F_l over observed SNPs (l >= i)     OLD  F_i ff_l over observed SNPs (l != i), F_i^2 observed (l == i)
//   slot k     : F_i^2 over all SNPs   (normalisation, smartpca.c:4367)
//   slot k + 1 : F_i x_a               (right-hand side; x_a = 0 where missing)
// Ft: [k][nb][npad] (F_i[s][a]); acc: [nz][k][k + 2][npad], each thread owns its slots (no atomics, fixed order).
template <bool NEWSHRINK>
__global__ void __launch_bounds__(128) shr_reduce_kernel(const double* __restrict__ Ft, int nb, int npad, int n, int k,
                                                         const uint8_t* __restrict__ work, int64_t wpitch, const double* __restrict__ ftab,
                                                         const uint8_t* __restrict__ used, const double* __restrict__ FF, int64_t ldf, int64_t s0,
                                                         double* __restrict__ acc) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, z = blockIdx.z, nz = gridDim.z;
  if (a >= n) return;
  const int b0 = (int)(((long long)nb * z) / nz), b1 = (int)(((long long)nb * (z + 1)) / nz);
  double g[SHR_KMAX], sq = 0.0, rx = 0.0;
#pragma unroll
  for (int l = 0; l < SHR_KMAX; l++) g[l] = 0.0;
  const size_t plane = (size_t)nb * npad;
  const int sh = (3 - (a & 3)) * 2;
  for (int b = b0; b < b1; b++) {
    const int64_t s = s0 + b;
    const int code = (work[s * wpitch + (a >> 2)] >> sh) & 3;
    const bool v = code != 3 && used[s];
    const double fi = Ft[(size_t)i * plane + (size_t)b * npad + a];
    sq = fma(fi, fi, sq);
    rx = fma(fi, ftab[s * 4 + code], rx);
    if (v) {
#pragma unroll
      for (int l = 0; l < SHR_KMAX; l++) {
        if (l < k) {
          if (NEWSHRINK) { if (l >= i) g[l] = fma(fi, Ft[(size_t)l * plane + (size_t)b * npad + a], g[l]); }
          else g[l] = fma(fi, l == i ? fi : FF[(size_t)l * ldf + s], g[l]);
        }
      }
    }
  }
  double* o = acc + ((size_t)(z * k + i) * (k + 2)) * npad + a;
#pragma unroll
  for (int l = 0; l < SHR_KMAX; l++)
    if (l < k) o[(size_t)l * npad] += g[l];
  o[(size_t)k * npad] += sq;
  o[(size_t)(k + 1) * npad] += rx;
}

__device__ __forceinline__ bool shr_cholsolve(double* a, double* p, double* x, const double* rr, int k) {
  // choldc / cholsl in the reference's operation order (nicksrc/linsubs.c:331-393), as lsq_solve_kernel
  for (int i = 0; i < k; i++)
    for (int j = i; j < k; j++) {
      double sum = a[i * k + j];
      for (int m = i - 1; m >= 0; m--) sum -= a[i * k + m] * a[j * k + m];
      if (i == j) { if (!(sum > 0.0)) return false; p[i] = sqrt(sum); }
      else a[j * k + i] = sum / p[i];
    }
  for (int i = 0; i < k; i++) { double sum = rr[i]; for (int m = i - 1; m >= 0; m--) sum -= a[i * k + m] * x[m]; x[i] = sum / p[i]; }
  for (int i = k - 1; i >= 0; i--) { double sum = x[i]; for (int m = i + 1; m < k; m++) sum -= a[m * k + i] * x[m]; x[i] = sum / p[i]; }
  return true;
}

// doshrinkp2: one regression per sample on all k leave-one-out loadings (smartpca.c:4165-4169) -> snew[i][a] = ans[i] * ymul[i][a]
__global__ void __launch_bounds__(128) shr_solve_new_kernel(const double* __restrict__ acc, int nz, int npad, int n, int k, double ncols,
                                                            const double* __restrict__ ymul, double* __restrict__ snew, uint8_t* __restrict__ okf) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  double co[SHR_KMAX * SHR_KMAX], p[SHR_KMAX], x[SHR_KMAX], rr[SHR_KMAX], yn[SHR_KMAX];
  auto A = [&](int i, int q) { double s = 0.0; for (int z = 0; z < nz; z++) s += acc[((size_t)(z * k + i) * (k + 2) + q) * npad + a]; return s; };
  for (int i = 0; i < k; i++) yn[i] = sqrt(A(i, k) / ncols);
  for (int i = 0; i < k; i++) {
    rr[i] = A(i, k + 1) / yn[i];
    for (int l = i; l < k; l++) { const double v = A(i, l) / (yn[i] * yn[l]); co[i * k + l] = v; co[l * k + i] = v; }
  }
  const bool good = shr_cholsolve(co, p, x, rr, k);
  for (int i = 0; i < k; i++) snew[(size_t)i * npad + a] = good ? x[i] * ymul[(size_t)i * npad + a] : 0.0;
  okf[a] = good ? 1 : 0;
}

// doshrinkp: for every (a, i) a regression on the base loadings with row i replaced (smartpca.c:4369-4374)
// Nt: [k(k+1)/2 + 1][npad] base normal-equation pairs (lsq_pairs_kernel order), Rt: [k][npad] base right-hand sides
__global__ void __launch_bounds__(128) shr_solve_old_kernel(const double* __restrict__ acc, int nz, int npad, int n, int k, double ncols,
                                                            const double* __restrict__ Nt, const double* __restrict__ Rt,
                                                            const double* __restrict__ ymul, double* __restrict__ snew, uint8_t* __restrict__ okf) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (a >= n) return;
  double co[SHR_KMAX * SHR_KMAX], p[SHR_KMAX], x[SHR_KMAX], rr[SHR_KMAX];
  auto A = [&](int q) { double s = 0.0; for (int z = 0; z < nz; z++) s += acc[((size_t)(z * k + i) * (k + 2) + q) * npad + a]; return s; };
  int pr = 0;
  for (int j = 0; j < k; j++) {
    rr[j] = Rt[(size_t)j * npad + a];
    for (int l = j; l < k; l++) { const double v = Nt[(size_t)(pr++) * npad + a]; co[j * k + l] = v; co[l * k + j] = v; }
  }
  const double yn = sqrt(A(k) / ncols);
  for (int l = 0; l < k; l++) {
    const double v = l == i ? A(i) / (yn * yn) : A(l) / yn;
    co[i * k + l] = v; co[l * k + i] = v;
  }
  rr[i] = A(k + 1) / yn;
  const bool good = shr_cholsolve(co, p, x, rr, k);
  snew[(size_t)i * npad + a] = good ? x[i] * ymul[(size_t)i * npad + a] : 0.0;
  if (!good) okf[a] = 0;
}

int shrink_run(eb_ctx* c, int k, int newshrink, double* coords, double* lambda_out, uint8_t* ok_out) {
  if (k < 1 || k > SHR_KMAX) { set_error("eb_shrink_coords: numeigs must be in 1..%d", SHR_KMAX); return EB_ERR_ARG; }
  int rc;
  const int64_t m = c->nsnp, mpad = c->mpad;
  const int n = c->nrows, npad = c->npad, N = c->numindivs;
  if (k >= n) { set_error("eb_shrink_coords: numeigs must be smaller than the number of PCA rows"); return EB_ERR_ARG; }
  // ---- full eigenbasis of Xn = XTX / y (smartpca.c:4285-4290; the second trace normalisation is the identity up to rounding)
  // (all n vectors stay on the device in c->zvec_d, row pitch c->zvec_ld)
  std::vector<double> lam(n);
  if ((rc = eig_resident(c, c->xtx.p, c->npad, n, 1.0 / c->y, n, lam.data(), nullptr))) return rc;
  for (int j = 0; j < k; j++) lambda_out[j] = lam[j];
  const int npairs = k * (k + 1) / 2 + 1;
  const int nb_max = (int)std::min<int64_t>(mpad, 2048);
  const int nz = std::max(1, std::min(16, (4 * c->num_sms * 128) / std::max(1, n * k)));
  DevBuf<double> E, X, x1, dg, mu, sc, lam_d, S, En, ymul, mtab, ftab, vtab, FF, Pt, Nt, Rt, D, Ft, acc, snew;
  DevBuf<uint8_t> okf;
  if ((rc = E.ensure((size_t)n * npad)) || (rc = X.ensure((size_t)n * npad)) || (rc = x1.ensure(npad)) || (rc = dg.ensure(npad)) ||
      (rc = mu.ensure(SHR_KMAX)) || (rc = sc.ensure(SHR_KMAX)) || (rc = lam_d.ensure(n)) || (rc = S.ensure((size_t)n * npad)) ||
      (rc = En.ensure((size_t)k * n * npad)) || (rc = ymul.ensure((size_t)k * npad)) || (rc = mtab.ensure((size_t)mpad * 4)) ||
      (rc = ftab.ensure((size_t)mpad * 4)) || (rc = vtab.ensure((size_t)mpad * 4)) || (rc = FF.ensure((size_t)k * mpad)) ||
      (rc = Pt.ensure((size_t)npairs * mpad)) || (rc = Nt.ensure((size_t)npairs * npad)) || (rc = Rt.ensure((size_t)k * npad)) ||
      (rc = D.ensure((size_t)nb_max * npad)) || (rc = Ft.ensure((size_t)k * nb_max * npad)) ||
      (rc = acc.ensure((size_t)nz * k * (k + 2) * npad)) || (rc = snew.ensure((size_t)k * npad)) || (rc = okf.ensure(npad)))
    return rc;
  // the reference divides by ncols = |xsnplist|; the factor is common to every loading vector and cancels in the final
  // unit-length normalisation of printevecs, so the uploaded SNP count serves
  double ncols = (double)m;
  if ((rc = peer_sum_host(c, &ncols, 1))) return rc;
  EB_CUDA(cudaMemsetAsync(E.p, 0, sizeof(double) * (size_t)n * npad, c->stream));
  EB_CUDA(cudaMemcpy2DAsync(E.p, sizeof(double) * npad, c->zvec_d.p, sizeof(double) * c->zvec_ld, sizeof(double) * n, n, cudaMemcpyDeviceToDevice, c->stream));
  EB_CUDA(cudaMemcpyAsync(lam_d.p, lam.data(), sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  shr_scale_rows_kernel<<<n, 256, 0, c->stream>>>(c->xtx.p, c->npad, n, 1.0 / c->y, X.p, npad, x1.p, dg.p);
  EB_CHECK_LAUNCH(c);
  shr_norme_rows_kernel<<<k, 256, 0, c->stream>>>(E.p, npad, n, mu.p, sc.p);                 // smartpca.c:4305-4307
  EB_CHECK_LAUNCH(c);
  const unsigned gm = (unsigned)((mpad + 255) / 256);
  mm_table_kernel<<<gm, 256, 0, c->stream>>>(m, mpad, n, c->c0_d.p, c->nmiss_d.p, c->xfancy_d.p, c->used_d.p, mtab.p);
  EB_CHECK_LAUNCH(c);
  fix_table_kernel<<<gm, 256, 0, c->stream>>>(m, mpad, c->xmean_d.p, c->xfancy_d.p, c->used_d.p, ftab.p);
  EB_CHECK_LAUNCH(c);
  mask_table_kernel<<<gm, 256, 0, c->stream>>>(m, mpad, c->used_d.p, vtab.p);
  EB_CHECK_LAUNCH(c);
  // ---- base loadings ffvecs (smartpca.c:4309-4311) and the old-style projection of EVERY individual (4313-4318)
  EB_CUDA(cudaMemsetAsync(FF.p, 0, sizeof(double) * (size_t)k * mpad, c->stream));
  if ((rc = launch_packed_gemm<MODE_XA>(c, mtab.p, E.p, npad, FF.p, mpad, k, 1.0))) return rc;
  {
    // ff_j /= sqrt(sum_s ff_j[s]^2 / ncols) with the sum taken over every shard's SNPs
    DevBuf<double> q_d;
    std::vector<double> q(k);
    if ((rc = q_d.ensure(k))) return rc;
    col_sumsq_kernel<<<k, 256, 0, c->stream>>>(FF.p, mpad, m, q_d.p);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaMemcpyAsync(q.data(), q_d.p, sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
    EB_CUDA(cudaStreamSynchronize(c->stream));
    if ((rc = peer_sum_host(c, q.data(), k))) return rc;
    for (int j = 0; j < k; j++) q[j] = 1.0 / sqrt(q[j] / ncols);
    EB_CUDA(cudaMemcpyAsync(q_d.p, q.data(), sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
    shr_scale_rows2_kernel<<<dim3((unsigned)((mpad + 255) / 256), k), 256, 0, c->stream>>>(FF.p, mpad, m, q_d.p);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaStreamSynchronize(c->stream));
  }
  std::vector<double> ffh((size_t)k * m), ones(k, 1.0), ss((size_t)k * N);
  std::vector<uint8_t> okall(N);
  std::vector<int> all(N);
  for (int i = 0; i < N; i++) all[i] = i;
  EB_CUDA(cudaMemcpy2DAsync(ffh.data(), sizeof(double) * m, FF.p, sizeof(double) * mpad, sizeof(double) * m, k, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  if ((rc = lsqproj_run(c, all.data(), N, ffh.data(), ones.data(), k, ss.data(), nullptr, nullptr, okall.data()))) return rc;
  // ---- base normal equations of the PCA rows (needed by the one-row-replaced regressions of doshrinkp)
  if (!newshrink) {
    lsq_pairs_kernel<<<gm, 256, 0, c->stream>>>(FF.p, mpad, m, mpad, k, Pt.p);
    EB_CHECK_LAUNCH(c);
    if ((rc = launch_packed_gemm<MODE_XTB>(c, ftab.p, FF.p, mpad, Rt.p, npad, k, 1.0))) return rc;
    if ((rc = launch_packed_gemm<MODE_XTB>(c, vtab.p, Pt.p, mpad, Nt.p, npad, npairs, 1.0))) return rc;
    if ((rc = peer_allreduce_any(c, Rt.p, (int64_t)k * npad)) || (rc = peer_allreduce_any(c, Nt.p, (int64_t)npairs * npad))) return rc;
  }
  // ---- leave-one-out eigenvectors, one m x m block per eigenvector
  for (int i = 0; i < k; i++) {
    shr_coeff_kernel<<<dim3((npad + 255) / 256, n), 256, 0, c->stream>>>(E.p, npad, n, k, i, lam_d.p, mu.p, sc.p, x1.p, dg.p, newshrink, S.p,
                                                                          ymul.p + (size_t)i * npad);
    EB_CHECK_LAUNCH(c);
    double* Ei = En.p + (size_t)i * n * npad;
    if ((rc = launch_gemm(c, true, true, S.p, npad, E.p, npad, Ei, npad, n, n, n))) return rc;       // ediff rows
    shr_enew_kernel<<<n, 256, 0, c->stream>>>(Ei, npad, n, E.p + (size_t)i * npad);
    EB_CHECK_LAUNCH(c);
  }
  // ---- SNP blocks: loadings of every leave-one-out vector and the per-sample sums
  // Integer tensor-core path (pg_i8.cu): the k leave-one-out matrices are cut into 7-bit digit rows ONCE, then every SNP block is one
  // launch per eigenvector with the packed block decoded inside the kernel; FP64 path: decode the block to FP64, DMMA GEMM.
  const int ncp32 = (n + 31) / 32 * 32;
  const size_t dig_bytes = (size_t)ncp32 * 8 * npad;
  bool wide_i8 = (c->opt_pg_method == 2 || (c->opt_pg_method == 0 && n >= c->opt_pg_i8_min && mpad >= c->opt_pg_i8_min)) && (npad % 128) == 0 &&
                 (nb_max % 128) == 0;
  DevBuf<uint8_t> dig_all;
  DevBuf<double> csc_all;
  if (wide_i8) {
    size_t freeb = 0, totalb = 0;
    cudaMemGetInfo(&freeb, &totalb);
    if ((size_t)k * dig_bytes > freeb / 2) wide_i8 = false;        // not enough room for the digit rows of all k matrices: FP64 path
  }
  if (wide_i8) {
    if ((rc = dig_all.ensure((size_t)k * dig_bytes)) || (rc = csc_all.ensure((size_t)k * ncp32))) return rc;
    for (int i = 0; i < k; i++)
      if ((rc = pg_i8_slice_wide(c, En.p + (size_t)i * n * npad, npad, n, npad, n, dig_all.p + (size_t)i * dig_bytes, csc_all.p + (size_t)i * ncp32))) return rc;
  }
  EB_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * (size_t)nz * k * (k + 2) * npad, c->stream));
  for (int64_t s0 = 0; s0 < mpad; s0 += nb_max) {
    const int nb = (int)std::min<int64_t>(nb_max, mpad - s0);
    if (wide_i8) {
      for (int i = 0; i < k; i++)
        if ((rc = pg_i8_rows_wide(c, c->work.p, c->wpitch, npad, mtab.p, dig_all.p + (size_t)i * dig_bytes, csc_all.p + (size_t)i * ncp32, n, s0, nb,
                                  Ft.p + (size_t)i * nb * npad, npad)))
          return rc;
    } else {
      shr_decode_kernel<<<dim3((npad + 255) / 256, nb), 256, 0, c->stream>>>(c->work.p, c->wpitch, mtab.p, s0, npad, D.p);
      EB_CHECK_LAUNCH(c);
      for (int i = 0; i < k; i++)
        if ((rc = launch_gemm(c, false, false, D.p, npad, En.p + (size_t)i * n * npad, npad, Ft.p + (size_t)i * nb * npad, npad, nb, n, n))) return rc;
    }
    const dim3 grid((n + 127) / 128, k, nz);
    if (newshrink)
      shr_reduce_kernel<true><<<grid, 128, 0, c->stream>>>(Ft.p, nb, npad, n, k, c->work.p, c->wpitch, ftab.p, c->used_d.p, FF.p, mpad, s0, acc.p);
    else
      shr_reduce_kernel<false><<<grid, 128, 0, c->stream>>>(Ft.p, nb, npad, n, k, c->work.p, c->wpitch, ftab.p, c->used_d.p, FF.p, mpad, s0, acc.p);
    EB_CHECK_LAUNCH(c);
  }
  if ((rc = peer_allreduce_any(c, acc.p, (int64_t)nz * k * (k + 2) * npad))) return rc;      // per-sample sums over every shard's SNPs
  EB_CUDA(cudaMemsetAsync(okf.p, 1, npad, c->stream));
  if (newshrink) shr_solve_new_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(acc.p, nz, npad, n, k, ncols, ymul.p, snew.p, okf.p);
  else shr_solve_old_kernel<<<dim3((n + 127) / 128, k), 128, 0, c->stream>>>(acc.p, nz, npad, n, k, ncols, Nt.p, Rt.p, ymul.p, snew.p, okf.p);
  EB_CHECK_LAUNCH(c);
  std::vector<double> sn((size_t)k * npad);
  std::vector<uint8_t> okh(npad);
  EB_CUDA(cudaMemcpyAsync(sn.data(), snew.p, sizeof(double) * (size_t)k * npad, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaMemcpyAsync(okh.data(), okf.p, npad, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  // ---- PCA rows take the shrunk value (smartpca.c:4381-4389); printevecs: 10 x, unit length per eigenvector (3849-3866)
  for (int j = 0; j < k; j++)
    for (int t = 0; t < n; t++) ss[(size_t)j * N + c->xindex_h[t]] = sn[(size_t)j * npad + t];
  for (int t = 0; t < n; t++) if (!okh[t]) okall[c->xindex_h[t]] = 0;
  for (int j = 0; j < k; j++) {
    double q = 0.0;
    for (int i = 0; i < N; i++) { const double v = 10.0 * ss[(size_t)j * N + i]; q += v * v; }
    const double inv = 1.0 / sqrt(q);
    for (int i = 0; i < N; i++) coords[(size_t)j * N + i] = 10.0 * ss[(size_t)j * N + i] * inv;
  }
  if (ok_out) memcpy(ok_out, okall.data(), N);
  return 0;
}

}  // namespace eb

// Seeded Gaussian start matrix: kjg_gsl.c:96-113 (GSL mt19937, seed 0 -> 4357) and kjg_gsl.c:145-186
// (Marsaglia polar pairs from -1+2*uniform_pos; an odd last column gets a plain uniform).  Host code, bit-exact.
extern "C" void eb_gauss_matrix(long seed, size_t n, size_t L, double* out) {
  uint32_t mt[624];
  int mti = 624;
  unsigned long s = (unsigned long)seed;
  if (s == 0) s = 4357;
  mt[0] = (uint32_t)s;
  for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  auto next = [&]() -> uint32_t {
    if (mti >= 624) {
      for (int k = 0; k < 624; k++) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  };
  auto upos = [&]() -> double { double x; do { x = next() / 4294967296.0; } while (x == 0); return x; };
  for (size_t i = 0; i < n; i++) {
    double* row = out + i * L;
    size_t j = 0;
    for (; j + 1 < L; j += 2) {
      double x0, x1, r2;
      do { x0 = -1 + 2 * upos(); x1 = -1 + 2 * upos(); r2 = x0 * x0 + x1 * x1; } while (r2 > 1.0 || r2 == 0);
      r2 = sqrt(-2.0 * log(r2) / r2);
      row[j] = x0 * r2; row[j + 1] = x1 * r2;
    }
    if (L % 2) row[L - 1] = upos();
  }
}
