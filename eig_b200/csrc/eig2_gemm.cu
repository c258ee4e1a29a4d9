// eig2_gemm.cu -- the two FP64 tensor-core products the large-N eigensolver is built from (see dmma_tile.cuh).
//
//   sym_skinny_kernel : Wpart[ks][c][i] = sum_{k in range ks} A[i][k] * Bt[c][k]      (A symmetric, n x n; Bt 64 x n)
//       Only the LOWER triangle of A (and complete diagonal 128-tiles) is read: for an output row tile T the k-tiles left
//       of the diagonal come from A[T][J] as [row][k] boxes, the ones right of it from A[J][T] as [k][row] boxes.
//       Used for W = A22 V in the band reduction and for the block mat-vecs of the subspace iteration.
//   syr2k_lower_kernel: A[i][c] -= sum_{k<64} V[k][i] Z[k][c] + Z[k][i] V[k][c]        (VZ = [Vt; Zt], 128 x n)
//       over lower 128 x 64 tiles (both halves of diagonal 128-tiles), in place.
// Tiles are aligned to absolute multiples of 128; the skinny operands are indexed by absolute row and are zero above the
// active sub-block, so no tile needs edge predicates on the operand side (TMA zero-fills beyond n).
#include <algorithm>
#include <cstdlib>
#include "dmma_tile.cuh"

namespace eb {

PFN_encodeTiled_t get_tensormap_encoder() {
  static PFN_encodeTiled_t fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled_t>(p);
  }
  return fn;
}

int make_f64_tensormap(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int boxc, int boxr) {
  PFN_encodeTiled_t enc = get_tensormap_encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return EB_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)boxc, (cuuint32_t)boxr};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f64 %lld x %lld ld %lld box %d x %d) failed: %d", (long long)rows, (long long)cols, (long long)ld, boxc, boxr, (int)r);
    return EB_ERR_CUDA;
  }
  return 0;
}

struct DtSmem {
  uint8_t* base;
  uint64_t* full;
  uint64_t* empty;
  __device__ double* A(int s) const { return reinterpret_cast<double*>(base + s * DT_STAGE_BYTES); }
  __device__ double* B(int s) const { return reinterpret_cast<double*>(base + s * DT_STAGE_BYTES + DT_A_BYTES); }
  __device__ uint32_t a32(int s) const { return dt_smem_u32(base) + s * DT_STAGE_BYTES; }
  __device__ uint32_t b32(int s) const { return dt_smem_u32(base) + s * DT_STAGE_BYTES + DT_A_BYTES; }
};

__device__ __forceinline__ DtSmem dt_setup(uint8_t* raw) {
  DtSmem s;
  s.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~uintptr_t(127));
  s.full = reinterpret_cast<uint64_t*>(s.base + DT_STAGES * DT_STAGE_BYTES);
  s.empty = s.full + DT_STAGES;
  if (threadIdx.x == 0) {
    for (int i = 0; i < DT_STAGES; i++) { dt_mbar_init(s.full + i, 1); dt_mbar_init(s.empty + i, DT_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  return s;
}

// ------------------------------------------------------------------------------------------------ W = A * B^T
// items = ntile * ksplit; item -> (ks = item / ntile, T = t0 + tile_first + (item % ntile) * tile_stride): a rank of a row-tile-sharded
// product owns every tile_stride-th row tile (single GPU: first 0, stride 1).  k chunks of 32 cover [t0*128, n).
__global__ void __launch_bounds__(DT_THREADS, 2)
sym_skinny_kernel(const __grid_constant__ CUtensorMap mapA_mk, const __grid_constant__ CUtensorMap mapA_km,
                  const __grid_constant__ CUtensorMap mapB, double* __restrict__ Wpart, int64_t ldw, int n, int t0, int ntile, int ksplit,
                  int tile_first, int tile_stride, int full_rows) {
  extern __shared__ __align__(128) uint8_t dt_raw[];
  const DtSmem sm = dt_setup(dt_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = ntile * ksplit;
  const int nchunks = (n - t0 * DT_M + DT_KC - 1) / DT_KC;

  if (warp == DT_CONSUMERS / 32) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int ks = item / ntile, T = t0 + tile_first + (item % ntile) * tile_stride;
        const int c0 = (int)(((long long)nchunks * ks) / ksplit), c1 = (int)(((long long)nchunks * (ks + 1)) / ksplit);
        for (int kc = c0; kc < c1; kc++) {
          const int k0 = t0 * DT_M + kc * DT_KC;
          dt_mbar_wait(sm.empty + stage, phase ^ 1);
          const bool mk = full_rows || k0 < (T + 1) * DT_M;      // full_rows: the square is stored, no transposed blocks
          dt_mbar_expect_tx(sm.full + stage, (mk ? DT_A_BYTES : DT_A_KM_BYTES) + DT_B_BYTES);
          if (mk) dt_tma_2d(sm.A(stage), &mapA_mk, k0, T * DT_M, sm.full + stage);
          else dt_tma_2d(sm.A(stage), &mapA_km, T * DT_M, k0, sm.full + stage);
          dt_tma_2d(sm.B(stage), &mapB, k0, 0, sm.full + stage);
          if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; t++)
#pragma unroll
    for (int u = 0; u < 4; u++) acc[t][u][0] = acc[t][u][1] = 0.0;
  uint32_t stage = 0, phase = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int ks = item / ntile, T = t0 + tile_first + (item % ntile) * tile_stride;
    const int c0 = (int)(((long long)nchunks * ks) / ksplit), c1 = (int)(((long long)nchunks * (ks + 1)) / ksplit);
    for (int kc = c0; kc < c1; kc++) {
      const int k0 = t0 * DT_M + kc * DT_KC;
      dt_mbar_wait(sm.full + stage, phase);
      if (full_rows || k0 < (T + 1) * DT_M) dt_stage_mma<false, false>(sm.a32(stage), sm.b32(stage), acc, wm, wn, g, q);
      else dt_stage_mma<true, false>(sm.a32(stage), sm.b32(stage), acc, wm, wn, g, q);
      dt_release_stage(sm.empty + stage, lane);
      if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
    }
    // epilogue: transposed store Wpart[ks][c][row]
    double* out = Wpart + (size_t)ks * DT_N * ldw;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int row = T * DT_M + wm * 64 + t * 8 + g;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int col = wn * 32 + u * 8 + q * 2;
        if (row < n) {
          out[(size_t)col * ldw + row] = acc[t][u][0];
          out[(size_t)(col + 1) * ldw + row] = acc[t][u][1];
        }
        acc[t][u][0] = acc[t][u][1] = 0.0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ A -= V Z^T + Z V^T
// item -> (I, J64): I = 128-row tile in [t0, tn), J64 = 64-col tile in [2*t0, 2*I+1]
__device__ __forceinline__ void syr2k_item(int item, int t0, int& I, int& J64) {
  int r = (int)((sqrtf(4.0f * (float)item + 1.0f) - 1.0f) * 0.5f);
  while ((long long)(r + 1) * (r + 2) <= item) r++;
  while ((long long)r * (r + 1) > item) r--;
  I = t0 + r;
  J64 = 2 * t0 + (item - r * (r + 1));
}

// rect_stride > 0: the row-distributed reduction (every rank stores the full square and owns every rect_stride-th 128-row tile):
// item -> (I = t0 + rect_first + (item / rect_cols) * rect_stride, J64 = 2 * t0 + item % rect_cols), all 64-column tiles of the row.
__device__ __forceinline__ void syr2k_item_any(int item, int t0, int rect_first, int rect_stride, int rect_cols, int& I, int& J64) {
  if (rect_stride > 0) { I = t0 + rect_first + (item / rect_cols) * rect_stride; J64 = 2 * t0 + item % rect_cols; }
  else syr2k_item(item, t0, I, J64);
}

__global__ void __launch_bounds__(DT_THREADS, 2)
syr2k_lower_kernel(const __grid_constant__ CUtensorMap mapVZ_km, const __grid_constant__ CUtensorMap mapVZ_kn,
                   double* __restrict__ A, int64_t lda, int n, int t0, int nitems, int nkc, int boff, int krows, double sgn,
                   int rect_first, int rect_stride, int rect_cols) {
  extern __shared__ __align__(128) uint8_t dt_raw[];
  const DtSmem sm = dt_setup(dt_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == DT_CONSUMERS / 32) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int I, J64; syr2k_item_any(item, t0, rect_first, rect_stride, rect_cols, I, J64);
        for (int kc = 0; kc < nkc; kc++) {
          dt_mbar_wait(sm.empty + stage, phase ^ 1);
          dt_mbar_expect_tx(sm.full + stage, DT_A_KM_BYTES + DT_B_KN_BYTES);
          dt_tma_2d(sm.A(stage), &mapVZ_km, I * DT_M, kc * DT_KC, sm.full + stage);
          int brow = kc * DT_KC + boff;
          if (brow >= krows) brow -= krows;
          dt_tma_2d(sm.B(stage), &mapVZ_kn, J64 * DT_N, brow, sm.full + stage);
          if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; t++)
#pragma unroll
    for (int u = 0; u < 4; u++) acc[t][u][0] = acc[t][u][1] = 0.0;
  uint32_t stage = 0, phase = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    int I, J64; syr2k_item_any(item, t0, rect_first, rect_stride, rect_cols, I, J64);
    // the tile this item will read-modify-write: pull its 512 lines (128 rows x 512 B) into L2 now, under the DMMA main loop, so
    // that the epilogue's dependent loads pay L2 latency instead of HBM latency (the main loop is only 4 k-chunks long)
    {
      const int tid = threadIdx.x;                 // 128 consumer threads: one row each, 4 lines per row
      const int row = I * DT_M + tid;
      if (row < n) {
        const double* p = A + (size_t)row * lda + (size_t)J64 * DT_N;
#pragma unroll
        for (int l = 0; l < 4; l++)
          if (J64 * DT_N + l * 16 < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + l * 16));
      }
    }
    for (int kc = 0; kc < nkc; kc++) {
      dt_mbar_wait(sm.full + stage, phase);
      dt_stage_mma<true, true>(sm.a32(stage), sm.b32(stage), acc, wm, wn, g, q);
      dt_release_stage(sm.empty + stage, lane);
      if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
    }
    // read-modify-write of the 128 x 64 tile in batches of two row blocks: 8 independent 16-byte loads in flight per
    // thread before the first dependent store (a one-at-a-time RMW would serialise 32 global round trips)
    const bool interior = (I + 1) * DT_M <= n && (J64 + 1) * DT_N <= n;
#pragma unroll
    for (int t2 = 0; t2 < 8; t2 += 2) {
      if (interior) {
        double2 v[2][4];
#pragma unroll
        for (int tt = 0; tt < 2; tt++)
#pragma unroll
          for (int u = 0; u < 4; u++)
            v[tt][u] = __ldcg(reinterpret_cast<const double2*>(A + (size_t)(I * DT_M + wm * 64 + (t2 + tt) * 8 + g) * lda + J64 * DT_N + wn * 32 + u * 8 + q * 2));
#pragma unroll
        for (int tt = 0; tt < 2; tt++)
#pragma unroll
          for (int u = 0; u < 4; u++) {
            v[tt][u].x += sgn * acc[t2 + tt][u][0]; v[tt][u].y += sgn * acc[t2 + tt][u][1];
            *reinterpret_cast<double2*>(A + (size_t)(I * DT_M + wm * 64 + (t2 + tt) * 8 + g) * lda + J64 * DT_N + wn * 32 + u * 8 + q * 2) = v[tt][u];
          }
      } else {
#pragma unroll
        for (int tt = 0; tt < 2; tt++) {
          const int row = I * DT_M + wm * 64 + (t2 + tt) * 8 + g;
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int col = J64 * DT_N + wn * 32 + u * 8 + q * 2;
            if (row < n && col < n) {
              double* p = A + (size_t)row * lda + col;
              p[0] += sgn * acc[t2 + tt][u][0];
              if (col + 1 < n) p[1] += sgn * acc[t2 + tt][u][1];
            }
          }
        }
      }
#pragma unroll
      for (int tt = 0; tt < 2; tt++)
#pragma unroll
        for (int u = 0; u < 4; u++) acc[t2 + tt][u][0] = acc[t2 + tt][u][1] = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ C = op(A) op(B)^T (general)
// C[i][j] = sum_k a(i,k) b(j,k) for i < M, j < N.  A_KM: A stored [k][i] (K x M) else [i][k] (M x K);
// B_KN: B stored [k][j] (K x N) else [j][k] (N x K).  128 x 64 tiles, persistent over items (n-tile fastest so that
// the CTAs resident at one time share A row panels through L2); edges rely on the TMA zero fill.
template <bool A_KM, bool B_KN>
__global__ void __launch_bounds__(DT_THREADS, 2)
dt_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, double* __restrict__ C, int64_t ldc,
               int M, int N, int K, int mt, int nt, double alpha, double beta) {
  extern __shared__ __align__(128) uint8_t dt_raw[];
  const DtSmem sm = dt_setup(dt_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = mt * nt;
  const int nchunks = (K + DT_KC - 1) / DT_KC;
  constexpr uint32_t TX = (A_KM ? DT_A_KM_BYTES : DT_A_BYTES) + (B_KN ? DT_B_KN_BYTES : DT_B_BYTES);

  if (warp == DT_CONSUMERS / 32) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int im = item / nt, in = item - im * nt;
        for (int kc = 0; kc < nchunks; kc++) {
          const int k0 = kc * DT_KC;
          dt_mbar_wait(sm.empty + stage, phase ^ 1);
          dt_mbar_expect_tx(sm.full + stage, TX);
          if (A_KM) dt_tma_2d(sm.A(stage), &mapA, im * DT_M, k0, sm.full + stage);
          else dt_tma_2d(sm.A(stage), &mapA, k0, im * DT_M, sm.full + stage);
          if (B_KN) dt_tma_2d(sm.B(stage), &mapB, in * DT_N, k0, sm.full + stage);
          else dt_tma_2d(sm.B(stage), &mapB, k0, in * DT_N, sm.full + stage);
          if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; t++)
#pragma unroll
    for (int u = 0; u < 4; u++) acc[t][u][0] = acc[t][u][1] = 0.0;
  uint32_t stage = 0, phase = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int im = item / nt, in = item - im * nt;
    for (int kc = 0; kc < nchunks; kc++) {
      dt_mbar_wait(sm.full + stage, phase);
      dt_stage_mma<A_KM, B_KN>(sm.a32(stage), sm.b32(stage), acc, wm, wn, g, q);
      dt_release_stage(sm.empty + stage, lane);
      if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int row = im * DT_M + wm * 64 + t * 8 + g;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int col = in * DT_N + wn * 32 + u * 8 + q * 2;
        if (row < M) {
          double* cp = C + (size_t)row * ldc + col;
          if (col + 1 < N) {
            double2 v = make_double2(alpha * acc[t][u][0], alpha * acc[t][u][1]);
            if (beta != 0.0) { const double2 o = __ldcg(reinterpret_cast<const double2*>(cp)); v.x += beta * o.x; v.y += beta * o.y; }
            *reinterpret_cast<double2*>(cp) = v;
          } else if (col < N) {
            cp[0] = alpha * acc[t][u][0] + (beta != 0.0 ? beta * cp[0] : 0.0);
          }
        }
        acc[t][u][0] = acc[t][u][1] = 0.0;
      }
    }
  }
}

template <bool A_KM, bool B_KN>
static int launch_gemm_t(eb_ctx* c, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N, int K,
                         double alpha, double beta) {
  CUtensorMap ma, mb;
  int rc;
  if (A_KM) { if ((rc = make_f64_tensormap(&ma, A, K, M, lda, DT_LD_M, DT_KC))) return rc; }
  else { if ((rc = make_f64_tensormap(&ma, A, M, K, lda, DT_LD_K, DT_M))) return rc; }
  if (B_KN) { if ((rc = make_f64_tensormap(&mb, B, K, N, ldb, DT_LD_N, DT_KC))) return rc; }
  else { if ((rc = make_f64_tensormap(&mb, B, N, K, ldb, DT_LD_K, DT_N))) return rc; }
  // function attributes are per device: set on every launch (microseconds) rather than behind a process-wide flag, so that several
  // contexts on different GPUs of one process (eb_local_comm) all get the shared-memory opt-in
  EB_CUDA(cudaFuncSetAttribute(dt_gemm_kernel<A_KM, B_KN>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
  const int mt = (M + DT_M - 1) / DT_M, nt = (N + DT_N - 1) / DT_N;
  const int grid = std::min(mt * nt, 2 * c->num_sms);
  dt_gemm_kernel<A_KM, B_KN><<<grid, DT_THREADS, DT_SMEM, c->stream>>>(ma, mb, C, ldc, M, N, K, mt, nt, alpha, beta);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// C (M x N, ldc; must be 16-byte aligned rows) = alpha op(A) op(B)^T + beta C on the FP64 tensor cores; a_km / b_kn select the storage
// of the operands (see dt_gemm_kernel).  Leading dimensions must be even (TMA strides are multiples of 16 bytes).
int launch_gemm(eb_ctx* c, bool a_km, bool b_kn, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                int M, int N, int K, double alpha, double beta) {
  if ((lda & 1) || (ldb & 1) || (ldc & 1)) { set_error("launch_gemm: leading dimensions must be even"); return EB_ERR_ARG; }
  if (!a_km && !b_kn) return launch_gemm_t<false, false>(c, A, lda, B, ldb, C, ldc, M, N, K, alpha, beta);
  if (a_km && b_kn) return launch_gemm_t<true, true>(c, A, lda, B, ldb, C, ldc, M, N, K, alpha, beta);
  if (a_km) return launch_gemm_t<true, false>(c, A, lda, B, ldb, C, ldc, M, N, K, alpha, beta);
  return launch_gemm_t<false, true>(c, A, lda, B, ldb, C, ldc, M, N, K, alpha, beta);
}

int dt_resident_ctas(eb_ctx* c) {
  // per context (= per device): the opt-in and the occupancy belong to the device the context lives on
  if (!c->dt_slots) {
    cudaFuncSetAttribute(sym_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM);
    cudaFuncSetAttribute(syr2k_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM);
    int a = 0, b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sym_skinny_kernel, DT_THREADS, DT_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, syr2k_lower_kernel, DT_THREADS, DT_SMEM);
    c->dt_slots = std::max(1, std::min(a, b));
    if (const char* e = getenv("EB_DBG_SLOTS")) c->dt_slots = std::max(1, atoi(e));
  }
  return c->dt_slots * c->num_sms;
}

// Wpart[ks][64][ldw] (ks < ksplit_out) = A[:, t0*128:] * Bt^T restricted to row tiles >= t0.  A: n x n (lda), Bt: 64 x n (ldb).
int launch_sym_skinny(eb_ctx* c, const double* A, int64_t lda, int n, int t0, const double* Bt, int64_t ldb, double* Wpart, int64_t ldw,
                      int max_ksplit, int* ksplit_out, int tile_first, int tile_stride, bool full_rows) {
  CUtensorMap mk, km, mb;
  int rc;
  if ((rc = make_f64_tensormap(&mk, A, n, n, lda, DT_LD_K, DT_M))) return rc;
  if ((rc = make_f64_tensormap(&km, A, n, n, lda, DT_LD_M, DT_KC))) return rc;
  if ((rc = make_f64_tensormap(&mb, Bt, DT_N, n, ldb, DT_LD_K, DT_N))) return rc;
  const int slots = dt_resident_ctas(c);
  const int ntile_all = (n + DT_M - 1) / DT_M - t0;
  const int ntile = ntile_all > tile_first ? (ntile_all - tile_first + tile_stride - 1) / tile_stride : 0;     // row tiles of this rank
  if (ntile <= 0) { *ksplit_out = 0; return 0; }
  const int nchunks = (n - t0 * DT_M + DT_KC - 1) / DT_KC;
  // pick the k split that fills the resident CTA slots best (each split needs >= 8 chunks of work)
  int best = 1; double beste = -1.0;
  for (int ks = 1; ks <= max_ksplit && ks * 8 <= std::max(nchunks, 8); ks++) {
    const int items = ntile * ks;
    const double waves = (double)items / slots;
    const double eff = waves / std::ceil(waves) - 0.004 * ks;
    if (eff > beste + 1e-9) { beste = eff; best = ks; }
  }
  if (const char* e = getenv("EB_DBG_KSPLIT")) best = std::max(1, std::min(atoi(e), max_ksplit));
  *ksplit_out = best;
  const int grid = std::min(ntile * best, slots);
  sym_skinny_kernel<<<grid, DT_THREADS, DT_SMEM, c->stream>>>(mk, km, mb, Wpart, ldw, n, t0, ntile, best, tile_first, tile_stride, full_rows ? 1 : 0);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// A[t0*128:, t0*128:] (lower 128x64 tiles) -= V Z^T + Z V^T with VZ = [Vt(64 rows); Zt(64 rows)] x n (ldv)
int launch_syr2k_lower(eb_ctx* c, double* A, int64_t lda, int n, int t0, const double* VZ, int64_t ldv, int tile_first, int tile_stride) {
  CUtensorMap km, kn;
  int rc;
  if ((rc = make_f64_tensormap(&km, VZ, 128, n, ldv, DT_LD_M, DT_KC))) return rc;
  if ((rc = make_f64_tensormap(&kn, VZ, 128, n, ldv, DT_LD_N, DT_KC))) return rc;
  const int nt = (n + DT_M - 1) / DT_M - t0;
  if (nt <= 0) return 0;
  if (tile_stride > 0) {
    // row-distributed: this rank's row tiles x every 64-column tile of the trailing square (full storage)
    const int rows = nt > tile_first ? (nt - tile_first + tile_stride - 1) / tile_stride : 0;
    const int cols = (n + DT_N - 1) / DT_N - 2 * t0;
    if (rows <= 0 || cols <= 0) return 0;
    const int nitems = rows * cols;
    const int grid = std::min(nitems, dt_resident_ctas(c));
    syr2k_lower_kernel<<<grid, DT_THREADS, DT_SMEM, c->stream>>>(km, kn, A, lda, n, t0, nitems, 4, 64, 128, -1.0, tile_first, tile_stride, cols);
    EB_CHECK_LAUNCH(c);
    return 0;
  }
  const int nitems = nt * (nt + 1);
  const int grid = std::min(nitems, dt_resident_ctas(c));
  syr2k_lower_kernel<<<grid, DT_THREADS, DT_SMEM, c->stream>>>(km, kn, A, lda, n, t0, nitems, 4, 64, 128, -1.0, 0, 0, 0);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// A (lower 128x64 tiles, n x n, lda) += T^T T with T = [krows][n] (ldt), krows a multiple of 32 (zero rows as padding):
// the dense-block accumulation that replaces domult_increment_normal (smartpca.c:3531-3561).
int launch_syrk_lower_add(eb_ctx* c, double* A, int64_t lda, int n, const double* T, int64_t ldt, int krows) {
  CUtensorMap km, kn;
  int rc;
  if ((rc = make_f64_tensormap(&km, T, krows, n, ldt, DT_LD_M, DT_KC))) return rc;
  if ((rc = make_f64_tensormap(&kn, T, krows, n, ldt, DT_LD_N, DT_KC))) return rc;
  const int nt = (n + DT_M - 1) / DT_M;
  const int nitems = nt * (nt + 1);
  const int grid = std::min(nitems, dt_resident_ctas(c));
  syr2k_lower_kernel<<<grid, DT_THREADS, DT_SMEM, c->stream>>>(km, kn, A, lda, n, 0, nitems, krows / DT_KC, 0, krows, 1.0, 0, 0, 0);
  EB_CHECK_LAUNCH(c);
  return 0;
}

}  // namespace eb
