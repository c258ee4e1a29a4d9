// eig2_kernels.cu -- large-N symmetric eigensolver on a device-resident FP64 matrix (replaces eigvecs() -> dspev_,
// eigsubs.c:39-55 / eigx.c:97-117, for the sizes where a one-stage reduction is HBM-bound).
//
//   spectrum (all eigenvalues, what .eval / Tracy-Widom consume, smartpca.c:1312-1431):
//     stage 1  dense -> band (bandwidth 64): per 64-column panel a grid-cooperative Householder QR (one grid barrier per
//              column, panel resident in shared memory), W = A22 V and the rank-128 update A22 -= V Z^T + Z V^T on the FP64
//              tensor cores (eig2_gemm.cu).  Only the lower triangle is maintained.
//     stage 2  band -> tridiagonal by column-wise bulge chasing; one CTA per sweep, sweeps pipelined two blocks apart
//              through release/acquire progress counters, band (51 MB at n = 50k) resident in L2.
//     then the Sturm-count bisection of eig_kernels.cu.
//   leading eigenvectors (what ridoutlier and the .evec consume, smartpca.c:1250,1444): Chebyshev-filtered subspace
//     iteration on the ORIGINAL matrix with a 64-wide block: block mat-vecs on the tensor cores, shifted CholeskyQR,
//     Rayleigh-Ritz through a 64 x 64 parallel Jacobi solver.  No back-transformation through the two stages is needed.
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "dmma_tile.cuh"

namespace cg = cooperative_groups;

namespace eb {

int launch_sym_skinny(eb_ctx* c, const double* A, int64_t lda, int n, int t0, const double* Bt, int64_t ldb, double* Wpart, int64_t ldw,
                      int max_ksplit, int* ksplit_out, int tile_first = 0, int tile_stride = 1, bool full_rows = false);
int launch_syr2k_lower(eb_ctx* c, double* A, int64_t lda, int n, int t0, const double* VZ, int64_t ldv, int tile_first = 0, int tile_stride = 0);

constexpr int BW = 64;             // band width after stage 1 == panel width
constexpr int MAX_KSPLIT = 4;

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier for kernels launched cooperatively (all CTAs co-resident): monotonic arrival counter, zeroed by the
// host before the launch.  A bounded spin turns a lost CTA into an error flag instead of a hung GPU.
__device__ __forceinline__ void grid_barrier(int* bar, int target, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1);
    long long spins = 0;
    while (ld_acquire(bar) < target) {
      if (++spins > (1ll << 26)) { atomicExch(err, 3); break; }
    }
    __threadfence();
  }
  __syncthreads();
}

// =================================================================================================== panel QR
struct PanelParams {
  double* A; int64_t lda; int n; int j;
  double* VZ; int64_t ldv;
  double* T;          // [64][64] row-major, upper triangular
  double* gpart;      // [2][G][64]
  double* rowk;       // [2][64]
  double* G2part;     // [G][64][64]
  double* G2;         // [64][64]
  double* gscratch;   // chunk storage when it does not fit shared memory
  int* bar; int* err; // grid barrier counter (zeroed before launch), error flag
  int rows_per; int use_global;
};

__global__ void __launch_bounds__(256, 1) panel_qr_kernel(PanelParams p) {
  extern __shared__ __align__(16) double psm[];
  const int G = gridDim.x, tid = threadIdx.x, c = tid & 63, part = tid >> 6;
  int bar_target = 0;
  const int r0 = p.j + BW, np = p.n - r0;
  const int row0 = blockIdx.x * p.rows_per;
  const int rows = max(0, min(p.rows_per, np - row0));
  double* P = p.use_global ? p.gscratch + (size_t)blockIdx.x * p.rows_per * 64 : psm;
  double* sh = p.use_global ? psm : psm + (size_t)p.rows_per * 64;
  double (*red)[64] = reinterpret_cast<double (*)[64]>(sh);     // [4][64]
  double* gsum = sh + 256;
  double* rk = sh + 320;
  double* scal_s = sh + 384;
  double* tau_s = sh + 448;
  const bool owner = blockIdx.x == 0;                            // rows_per >= 64: panel rows 0..63 live in CTA 0

  for (int idx = tid; idx < rows * 64; idx += 256) {
    const int i = idx >> 6, cc = idx & 63;
    P[idx] = p.A[(size_t)(r0 + row0 + i) * p.lda + p.j + cc];
  }
  if (blockIdx.x == G - 1) {
    // absolute rows j..j+63 of Vt are above the active block from now on
    for (int idx = tid; idx < 64 * 64; idx += 256) p.VZ[(size_t)(idx >> 6) * p.ldv + p.j + (idx & 63)] = 0.0;
  }
  if (tid < 64) { scal_s[tid] = 0.0; tau_s[tid] = 0.0; }
  __syncthreads();

  const int nr = min(64, np - 1);
  for (int k = 0; k < nr; k++) {
    double acc = 0.0;
    if (c >= k) {
      int i = part;
      if (owner) while (i <= k) i += 4;                        // only rows below the diagonal enter the dots
      for (; i < rows; i += 4) acc += P[i * 64 + k] * P[i * 64 + c];
    }
    red[part][c] = acc;
    __syncthreads();
    if (tid < 64) {
      __stcg(&p.gpart[((size_t)(k & 1) * G + blockIdx.x) * 64 + tid], red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid]);
      if (owner) __stcg(&p.rowk[(k & 1) * 64 + tid], P[k * 64 + tid]);
    }
    grid_barrier(p.bar, bar_target += G, p.err);
    acc = 0.0;
    for (int b = part; b < G; b += 4) acc += ldcg(&p.gpart[((size_t)(k & 1) * G + b) * 64 + c]);
    __syncthreads();                                            // red[] of the first phase has been consumed
    red[part][c] = acc;
    __syncthreads();
    if (tid < 64) { gsum[tid] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid]; rk[tid] = ldcg(&p.rowk[(k & 1) * 64 + tid]); }
    __syncthreads();
    const double alpha = rk[k], sigma = gsum[k];
    double tau = 0.0, scal = 0.0, beta = alpha;
    if (sigma != 0.0) {
      const double nrm = sqrt(alpha * alpha + sigma);
      beta = alpha >= 0.0 ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scal = 1.0 / (alpha - beta);
    }
    if (c > k && tau != 0.0) {
      const double vtp = rk[c] + scal * gsum[c];
      const double f = tau * scal * vtp;
      int i = part;
      if (owner) {
        if ((k & 3) == part) P[k * 64 + c] -= tau * vtp;        // row k itself (v_k = 1)
        while (i <= k) i += 4;
      }
      for (; i < rows; i += 4) P[i * 64 + c] -= P[i * 64 + k] * f;
    }
    if (tid == 0) {
      if (owner) P[k * 64 + k] = beta;
      scal_s[k] = scal; tau_s[k] = tau;
    }
    __syncthreads();
  }

  // R back into the matrix (band part): rows r0..r0+63, columns j..j+63, upper triangular
  if (owner) {
    for (int idx = tid; idx < min(rows, 64) * 64; idx += 256) {
      const int i = idx >> 6, cc = idx & 63;
      p.A[(size_t)(r0 + i) * p.lda + p.j + cc] = (i <= cc) ? P[idx] : 0.0;
    }
  }
  __syncthreads();
  // P -> V (unit lower trapezoidal; columns >= nr and columns with tau == 0 are zero below the diagonal)
  for (int idx = tid; idx < rows * 64; idx += 256) {
    const int i = idx >> 6, cc = idx & 63, gi = row0 + i;
    double v = 0.0;
    if (cc < nr) v = gi > cc ? P[idx] * scal_s[cc] : (gi == cc ? 1.0 : 0.0);
    P[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < rows * 64; idx += 256) {
    const int cc = idx / rows, i = idx - cc * rows;
    p.VZ[(size_t)cc * p.ldv + r0 + row0 + i] = P[i * 64 + cc];
  }
  // partial Gram V^T V (a < c part is what T needs)
  {
    double g2[16];
#pragma unroll
    for (int a = 0; a < 16; a++) g2[a] = 0.0;
    for (int i = 0; i < rows; i++) {
      const double vc = P[i * 64 + c];
#pragma unroll
      for (int a = 0; a < 16; a++) g2[a] += P[i * 64 + part * 16 + a] * vc;
    }
#pragma unroll
    for (int a = 0; a < 16; a++) __stcg(&p.G2part[((size_t)blockIdx.x * 64 + part * 16 + a) * 64 + c], g2[a]);
  }
  grid_barrier(p.bar, bar_target += G, p.err);
  for (int a = blockIdx.x; a < 64; a += G) {
    double acc = 0.0;
    for (int b = part; b < G; b += 4) acc += ldcg(&p.G2part[((size_t)b * 64 + a) * 64 + c]);
    __syncthreads();
    red[part][c] = acc;
    __syncthreads();
    if (tid < 64) __stcg(&p.G2[a * 64 + tid], red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid]);
  }
  grid_barrier(p.bar, bar_target += G, p.err);
  if (blockIdx.x == 0) {
    // T (dlarft, forward columnwise): T[0:k,k] = -tau_k T[0:k,0:k] (V^T v_k)
    double* Ts = P;                 // 64 x 65 (P is at least 64 x 64 ... use stride 64 to stay inside the chunk)
    double* Gs = p.use_global ? p.gscratch + 4096 : psm + 4096;   // second 64x64 block of the chunk area (rows_per >= 128 guaranteed by host)
    for (int idx = tid; idx < 4096; idx += 256) { Ts[idx] = 0.0; Gs[idx] = ldcg(&p.G2[idx]); }
    __syncthreads();
    for (int k = 0; k < 64; k++) {
      const double tk = tau_s[k];
      if (tid < k && tk != 0.0) {
        double s = 0.0;
        for (int m = tid; m < k; m++) s += Ts[tid * 64 + m] * Gs[m * 64 + k];
        Ts[tid * 64 + k] = -tk * s;
      }
      if (tid == k) Ts[k * 64 + k] = tk;
      __syncthreads();
    }
    for (int idx = tid; idx < 4096; idx += 256) p.T[idx] = Ts[idx];
  }
}

// =================================================================================================== skinny helpers
// Out[c][i] = a1 * sum_ks Wpart[ks][c][i] + a2 * Y1[c][i] + a3 * Y0[c][i]   for i in [i_lo, n), zero for [i_zero0, i_lo)
__global__ void __launch_bounds__(256) combine_kernel(const double* __restrict__ Wpart, int ksplit, int64_t ldw, double* __restrict__ Out,
                                                      int64_t ldo, int n, int i_zero0, int i_lo, double a1, const double* __restrict__ Y1,
                                                      double a2, const double* __restrict__ Y0, double a3, int64_t ldy,
                                                      const double* __restrict__ Y2 = nullptr, double a4 = 0.0) {
  const int i = i_zero0 + blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
  if (i >= n) return;
  double v = 0.0;
  if (i >= i_lo) {
    for (int ks = 0; ks < ksplit; ks++) v += Wpart[((size_t)ks * 64 + c) * ldw + i];
    v *= a1;
    if (Y1) v += a2 * Y1[(size_t)c * ldy + i];
    if (Y0) v += a3 * Y0[(size_t)c * ldy + i];
    if (Y2) v += a4 * Y2[(size_t)c * ldy + i];
  }
  Out[(size_t)c * ldo + i] = v;
}

// Gpart[chunk][a][c] = sum_{i in chunk} Xt[a][i] * Yt[c][i], chunks of GR_CHUNK over [i0, n)
__global__ void __launch_bounds__(256) gram64_kernel(const double* __restrict__ Xt, int64_t ldx, const double* __restrict__ Yt, int64_t ldy,
                                                     int i0, int n, int GR_CHUNK, double* __restrict__ Gpart) {
  __shared__ double Xs[64][33], Ys[64][33];
  const int tid = threadIdx.x, ta = tid >> 4, tc = tid & 15;
  const int lo = i0 + blockIdx.x * GR_CHUNK, hi = min(n, lo + GR_CHUNK);
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
  for (int s = lo; s < hi; s += 32) {
    __syncthreads();
    for (int idx = tid; idx < 64 * 32; idx += 256) {
      const int r = idx >> 5, ii = idx & 31, i = s + ii;
      Xs[r][ii] = i < hi ? Xt[(size_t)r * ldx + i] : 0.0;
      Ys[r][ii] = i < hi ? Yt[(size_t)r * ldy + i] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int ii = 0; ii < 32; ii++) {
      double xa[4], yb[4];
#pragma unroll
      for (int a = 0; a < 4; a++) xa[a] = Xs[ta * 4 + a][ii];
#pragma unroll
      for (int b = 0; b < 4; b++) yb[b] = Ys[tc * 4 + b][ii];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] += xa[a] * yb[b];
    }
  }
  double* out = Gpart + (size_t)blockIdx.x * 4096;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) out[(ta * 4 + a) * 64 + tc * 4 + b] = acc[a][b];
}

// Out[c][i] = sum_k C1[k][c] X1[k][i] (+ sum_k C2[k][c] X2[k][i]) for i in [i0, n); C row-major 64 x 64
__global__ void __launch_bounds__(256) left_mult_kernel(const double* __restrict__ C1, const double* __restrict__ X1,
                                                        const double* __restrict__ C2, const double* __restrict__ X2, int64_t ldx,
                                                        double* __restrict__ Out, int64_t ldo, int i0, int n) {
  extern __shared__ __align__(16) double lm[];
  double* C1s = lm;            // [64][64]
  double* X1s = lm + 4096;     // [64][64]
  double* C2s = lm + 8192;
  double* X2s = lm + 12288;
  const int tid = threadIdx.x, il = tid & 63, cgp = tid >> 6;
  const int ibase = i0 + blockIdx.x * 64;
  for (int idx = tid; idx < 4096; idx += 256) {
    const int k = idx >> 6, ii = idx & 63, i = ibase + ii;
    C1s[idx] = C1[idx];
    X1s[idx] = i < n ? X1[(size_t)k * ldx + i] : 0.0;
    if (C2) { C2s[idx] = C2[idx]; X2s[idx] = i < n ? X2[(size_t)k * ldx + i] : 0.0; }
  }
  __syncthreads();
  double acc[16];
#pragma unroll
  for (int cc = 0; cc < 16; cc++) acc[cc] = 0.0;
  for (int k = 0; k < 64; k++) {
    const double x = X1s[k * 64 + il];
#pragma unroll
    for (int cc = 0; cc < 16; cc++) acc[cc] += C1s[k * 64 + cgp * 16 + cc] * x;
  }
  if (C2) {
    for (int k = 0; k < 64; k++) {
      const double x = X2s[k * 64 + il];
#pragma unroll
      for (int cc = 0; cc < 16; cc++) acc[cc] += C2s[k * 64 + cgp * 16 + cc] * x;
    }
  }
  const int i = ibase + il;
  if (i < n) {
#pragma unroll
    for (int cc = 0; cc < 16; cc++) Out[(size_t)(cgp * 16 + cc) * ldo + i] = acc[cc];
  }
}

// one CTA: S = sum_chunks Spart; M = T^T S T; C1 = T, C2 = -M/2
__global__ void __launch_bounds__(256) sbr_small_kernel(const double* __restrict__ Spart, int nchunk, const double* __restrict__ T,
                                                        double* __restrict__ C1, double* __restrict__ C2) {
  extern __shared__ __align__(16) double ss[];
  double* S = ss; double* Ts = ss + 4096; double* U = ss + 8192;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 4096; idx += 256) {
    double v = 0.0;
    for (int ch = 0; ch < nchunk; ch++) v += Spart[(size_t)ch * 4096 + idx];
    S[idx] = v; Ts[idx] = T[idx];
  }
  __syncthreads();
  // U = S T
  for (int idx = tid; idx < 4096; idx += 256) {
    const int r = idx >> 6, cc = idx & 63;
    double v = 0.0;
    for (int k = 0; k <= cc; k++) v += S[r * 64 + k] * Ts[k * 64 + cc];      // T upper triangular
    U[idx] = v;
  }
  __syncthreads();
  // M = T^T U
  for (int idx = tid; idx < 4096; idx += 256) {
    const int r = idx >> 6, cc = idx & 63;
    double v = 0.0;
    for (int k = 0; k <= r; k++) v += Ts[k * 64 + r] * U[k * 64 + cc];
    C2[idx] = -0.5 * v;
    C1[idx] = Ts[idx];
  }
}

// =================================================================================================== band storage
// AB[col][d] = A[col+d][col] for d <= BW (lower triangle of A), zero for BW < d < 2*BW
__global__ void __launch_bounds__(128) band_extract_kernel(const double* __restrict__ A, int64_t lda, int n, double* __restrict__ AB) {
  const int col = blockIdx.x, d = threadIdx.x;
  double v = 0.0;
  if (d <= BW && col + d < n) v = A[(size_t)(col + d) * lda + col];
  AB[(size_t)col * (2 * BW) + d] = v;
}
__global__ void band_de_kernel(const double* __restrict__ AB, int n, double* __restrict__ d, double* __restrict__ e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { d[i] = AB[(size_t)i * (2 * BW)]; e[i] = i < n - 1 ? AB[(size_t)i * (2 * BW) + 1] : 0.0; }
}

// =================================================================================================== bulge chasing

constexpr int BC_THREADS = 256;
constexpr int BC_LD = 65;
constexpr int BC_DONE = 0x3fffffff;
constexpr int BC_SMEM = (2 * 64 * BC_LD + 64 * 8 + 256 * 3 + 8) * 8;

struct HouseSh { double tau, beta; };

// reflector from x[0..len) held in xs (shared): leaves v in vs (v[0] = 1, zero beyond len), tau/beta in hs.  warp 0 only.
__device__ __forceinline__ void bc_make_reflector(const double* xs, int len, double* vs, HouseSh* hs, int lane) {
  const double x0 = xs[lane], x1 = xs[lane + 32];
  double s = (lane >= 1 && lane < len ? x0 * x0 : 0.0) + (lane + 32 < len ? x1 * x1 : 0.0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const double alpha = __shfl_sync(0xffffffffu, x0, 0);
  double tau = 0.0, beta = alpha, scal = 0.0;
  if (s != 0.0) {
    const double nrm = sqrt(alpha * alpha + s);
    beta = alpha >= 0.0 ? -nrm : nrm;
    tau = (beta - alpha) / beta;
    scal = 1.0 / (alpha - beta);
  }
  vs[lane] = lane == 0 ? 1.0 : (lane < len ? x0 * scal : 0.0);
  vs[lane + 32] = lane + 32 < len ? x1 * scal : 0.0;
  if (lane == 0) { hs->tau = tau; hs->beta = beta; }
}

#ifdef EB_BC_PROFILE
__device__ unsigned long long g_bc_prof[8];
#define BC_TICK(slot) do { if (tid == 0 && blockIdx.x == 1) { const long long t_ = clock64(); g_bc_prof[slot] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define BC_TICK(slot) do { } while (0)
#endif

// One CTA per sweep (column s of the band is reduced to tridiagonal form and its bulge chased off the end); sweep s may
// run block k once sweep s-1 has finished block k+1.  Thread (i = tid & 63, part = tid >> 6) owns row i, columns
// part*16 .. part*16+15 of the current 64 x 64 off-diagonal block B and diagonal block D in REGISTERS (one L2 round trip
// per step); shared memory holds the copies the cross-thread reductions need.  Per step: B <- H2 (B H1) fused into one
// update (w = tau1 B v1 gives column 0 of B H1, hence H2, before B is touched), D <- H2 D H2.
// Q2 != nullptr: every reflector is kept for the back-transformation of eigenvectors (q2_apply_kernel): sweep s owns n - 1 - s doubles at
// q2_off(s, n), element t belongs to row s + 1 + t; reflector k of the sweep occupies elements 64 k .. 64 k + len - 1 with tau in place of
// the implicit v[0] = 1.
__host__ __device__ __forceinline__ size_t q2_off(long long s, long long n) { return (size_t)(s * (n - 1) - s * (s - 1) / 2); }

__global__ void __launch_bounds__(BC_THREADS, 2) bulge_chase_kernel(double* __restrict__ AB, int n, int* __restrict__ prog, int* __restrict__ err,
                                                                    double* __restrict__ Q2) {
#ifdef EB_BC_PROFILE
  long long t_prev = clock64();
#endif
  extern __shared__ __align__(16) double bcs[];
  double* Bs = bcs;                       // [64][BC_LD]  Bs[c][i]
  double* Ds = bcs + 64 * BC_LD;          // [64][BC_LD]  full symmetric
  double* va = Ds + 64 * BC_LD;           // reflector ping
  double* vb = va + 64;                   // reflector pong
  double* xs = vb + 64;
  double* wv = xs + 64;                   // tau1 * B v1
  double* us = wv + 64;                   // u[c]
  double* ps = us + 64;                   // p[i]
  double* w2 = ps + 64;                   // two-sided w
  double* spare = w2 + 64;
  double* red = spare + 64;               // [256]
  double* redU = red + 256;               // [256]
  double* redP = redU + 256;              // [256]
  double* red2 = redP + 256;              // [4]
  __shared__ HouseSh hs;
  __shared__ int abort_flag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, i = tid & 63, part = tid >> 6, c0 = part * 16;
  constexpr int LDAB = 2 * BW;
  if (tid == 0) abort_flag = 0;
  __syncthreads();

  auto wait_for = [&](int s_prev, int need) {
    if (s_prev >= 0) {
      if (tid == 0) {
        long long spins = 0;
        while (ld_acquire(prog + s_prev) < need) {
          if (spins > 4) __nanosleep(spins < 64 ? 20 : 400);
          if (++spins > (1ll << 24)) { abort_flag = 1; atomicExch(err, 1); break; }
          if (spins % 4096 == 0 && ld_acquire(err)) { abort_flag = 1; break; }
        }
      }
      __syncthreads();
    }
  };
  auto publish = [&](int s, int val) {
    __syncthreads();                       // every thread's band stores precede the barrier ...
    if (tid == 0) { __threadfence(); st_release(prog + s, val); }   // ... and are published cumulatively by one fence + release
  };
  // two-sided update pieces shared by block 0 and the chase steps (D rows held in dreg; v, tau given)
  double breg[16], dreg[16];

  for (int s = blockIdx.x; s < n - 2; s += gridDim.x) {
    // ---------------------------------------------------------------- block 0: annihilate column s below the sub-diagonal
    wait_for(s - 1, 2);
    if (abort_flag) return;
    int r0 = s + 1, la = min(BW, n - r0);
    double* v1 = va; double* v2 = vb;
    {
      const double x = (tid < 64 && tid < la) ? __ldcg(&AB[(size_t)s * LDAB + 1 + tid]) : 0.0;
#pragma unroll
      for (int it = 0; it < 16; it++) {
        const int cc = c0 + it;
        const bool okd = i >= cc && i < la;
        const double td = __ldcg(&AB[okd ? (size_t)(r0 + cc) * LDAB + (i - cc) : (size_t)r0 * LDAB]);
        dreg[it] = okd ? td : 0.0;
      }
      asm volatile("" ::: "memory");
      if (tid < 64) xs[tid] = x;
#pragma unroll
      for (int it = 0; it < 16; it++) {
        const int cc = c0 + it;
        if (i >= cc) { Ds[cc * BC_LD + i] = dreg[it]; Ds[i * BC_LD + cc] = dreg[it]; }
      }
    }
    __syncthreads();
    if (warp == 0) bc_make_reflector(xs, la, v1, &hs, lane);
    __syncthreads();
    double tau = hs.tau;
    if (tid < la) __stcg(&AB[(size_t)s * LDAB + 1 + tid], tid == 0 ? hs.beta : 0.0);
    double* q2s = Q2 ? Q2 + q2_off(s, n) : nullptr;
    if (q2s && tid < la) __stcs(&q2s[tid], tid == 0 ? tau : v1[tid]);
    {
      double pr = 0.0;
#pragma unroll
      for (int it = 0; it < 16; it++) { dreg[it] = Ds[(c0 + it) * BC_LD + i]; pr += dreg[it] * v1[c0 + it]; }
      redP[part * 64 + i] = pr;
      __syncthreads();
      if (tid < 64) ps[tid] = tau * (redP[tid] + redP[64 + tid] + redP[128 + tid] + redP[192 + tid]);
      __syncthreads();
      if (warp == 0) {
        double sv = ps[lane] * v1[lane] + ps[lane + 32] * v1[lane + 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        const double al = -0.5 * tau * sv;
        w2[lane] = ps[lane] + al * v1[lane];
        w2[lane + 32] = ps[lane + 32] + al * v1[lane + 32];
      }
      __syncthreads();
      const double vi = v1[i], wi = w2[i];
#pragma unroll
      for (int it = 0; it < 16; it++) {
        const int cc = c0 + it;
        if (i >= cc && i < la) __stcg(&AB[(size_t)(r0 + cc) * LDAB + (i - cc)], dreg[it] - (vi * w2[cc] + wi * v1[cc]));
      }
    }
    publish(s, 1);

    // ---------------------------------------------------------------- chase the bulge down the band
    for (int k = 1;; k++) {
      const int r1 = r0 + la;
      const int lb = min(BW, n - r1);
      if (lb <= 0) break;
      BC_TICK(0);
      wait_for(s - 1, k + 2);
      if (abort_flag) return;
      BC_TICK(1);
      // B: rows r1..r1+lb-1, cols r0..r0+la-1 ;  D: rows/cols r1..r1+lb-1 (lower part loaded, mirrored through shared memory)
#pragma unroll
      for (int it = 0; it < 16; it++) {
        const int cc = c0 + it;
        // unconditional loads from a clamped (always valid) address + select: straight-line code, all 32 loads in flight
        const bool okb = cc < la && i < lb, okd = i >= cc && i < lb;
        const double tb = __ldcg(&AB[okb ? (size_t)(r0 + cc) * LDAB + (la + i - cc) : (size_t)r0 * LDAB]);
        const double td = __ldcg(&AB[okd ? (size_t)(r1 + cc) * LDAB + (i - cc) : (size_t)r0 * LDAB]);
        breg[it] = okb ? tb : 0.0;
        dreg[it] = okd ? td : 0.0;
      }
      asm volatile("" ::: "memory");      // all 32 loads are issued before the first use (one L2 round trip, not 32)
      BC_TICK(2);
      {
        double acc = 0.0;
#pragma unroll
        for (int it = 0; it < 16; it++) {
          const int cc = c0 + it;
          Bs[cc * BC_LD + i] = breg[it];
          if (i >= cc) { Ds[cc * BC_LD + i] = dreg[it]; Ds[i * BC_LD + cc] = dreg[it]; }
          acc += breg[it] * v1[cc];
        }
        red[part * 64 + i] = acc;
      }
      BC_TICK(3);
      __syncthreads();                                                       // (1)
      BC_TICK(4);
      if (tid < 64) {
        const double w = tau * (red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]);
        wv[tid] = w;
        xs[tid] = Bs[tid] - w;                                               // column 0 of B H1 (v1[0] = 1)
      }
      __syncthreads();                                                       // (2)
      if (warp == 0) bc_make_reflector(xs, lb, v2, &hs, lane);
      __syncthreads();                                                       // (3)
      const double tau2 = hs.tau;
      if (q2s && tid < lb) __stcs(&q2s[(r1 - s - 1) + tid], tid == 0 ? tau2 : v2[tid]);
      {
        // u_raw[c = i] over rows c0..c0+15, gamma = v2 . wv, p_raw[i] over columns c0..c0+15
        double ur = 0.0, gp = 0.0, pr = 0.0;
#pragma unroll
        for (int it = 0; it < 16; it++) {
          const int r = c0 + it;
          const double v2r = v2[r];
          ur += v2r * Bs[i * BC_LD + r];
          gp += v2r * wv[r];
          dreg[it] = Ds[r * BC_LD + i];
          pr += dreg[it] * v2r;
        }
        redU[part * 64 + i] = ur;
        redP[part * 64 + i] = pr;
        if (i == 0) red2[part] = gp;
      }
      __syncthreads();                                                       // (4)
      if (tid < 64) {
        const double gamma = red2[0] + red2[1] + red2[2] + red2[3];
        us[tid] = tau2 * ((redU[tid] + redU[64 + tid] + redU[128 + tid] + redU[192 + tid]) - gamma * v1[tid]);
      } else if (tid < 128) {
        const int j = tid - 64;
        ps[j] = tau2 * (redP[j] + redP[64 + j] + redP[128 + j] + redP[192 + j]);
      }
      __syncthreads();                                                       // (5)
      if (warp == 0) {
        double sv = ps[lane] * v2[lane] + ps[lane + 32] * v2[lane + 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        const double al = -0.5 * tau2 * sv;
        w2[lane] = ps[lane] + al * v2[lane];
        w2[lane + 32] = ps[lane + 32] + al * v2[lane + 32];
      }
      __syncthreads();                                                       // (6)
      BC_TICK(5);
      {
        const double wvi = wv[i], v2i = v2[i], w2i = w2[i];
        const double beta2 = hs.beta;
#pragma unroll
        for (int it = 0; it < 16; it++) {
          const int cc = c0 + it;
          double bn = breg[it] - (wvi * v1[cc] + v2i * us[cc]);
          if (cc == 0) bn = i == 0 ? beta2 : 0.0;                            // column 0 is (beta, 0, ..., 0)
          if (cc < la && i < lb) __stcg(&AB[(size_t)(r0 + cc) * LDAB + (la + i - cc)], bn);
          if (i >= cc && i < lb) __stcg(&AB[(size_t)(r1 + cc) * LDAB + (i - cc)], dreg[it] - (v2i * w2[cc] + w2i * v2[cc]));
        }
      }
      publish(s, k + 1);                                                     // (7)
      BC_TICK(6);
#ifdef EB_BC_PROFILE
      if (tid == 0 && blockIdx.x == 1) g_bc_prof[7] += 1;
#endif
      r0 = r1; la = lb; tau = tau2;
      double* t = v1; v1 = v2; v2 = t;
    }
    publish(s, BC_DONE);
  }
}

// =================================================================================================== 64 x 64 dense helpers
// Parallel two-sided Jacobi (round-robin ordering).  In: Hpart (nchunk x 64 x 64 partial Grams, summed and symmetrised here).
// Out: theta[64] descending, Z[k][c] (row-major) = eigenvector c.
__global__ void __launch_bounds__(256) jacobi64_kernel(const double* __restrict__ Hpart, int nchunk, double* __restrict__ theta,
                                                       double* __restrict__ Z) {
  extern __shared__ __align__(16) double jsm[];
  double* H = jsm; double* V = jsm + 64 * 65;
  __shared__ double cs[32], sn[32], redw[8];
  __shared__ int pp[32], qq[32], rank[64];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 4096; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    double a = 0.0, b = 0.0;
    for (int ch = 0; ch < nchunk; ch++) { a += Hpart[(size_t)ch * 4096 + r * 64 + c]; b += Hpart[(size_t)ch * 4096 + c * 64 + r]; }
    H[r * 65 + c] = 0.5 * (a + b);
    V[r * 65 + c] = r == c ? 1.0 : 0.0;
  }
  __syncthreads();
  __shared__ int nrot;
  for (int sweep = 0; sweep < 30; sweep++) {
    if (tid == 0) nrot = 0;
    __syncthreads();
    for (int round = 0; round < 63; round++) {
      if (tid < 32) {
        int a, b;
        if (tid == 0) { a = 63; b = round; }
        else { a = (round + tid) % 63; b = (round - tid + 63) % 63; }
        const int p = min(a, b), q = max(a, b);
        pp[tid] = p; qq[tid] = q;
        const double hpq = H[p * 65 + q], hpp = H[p * 65 + p], hqq = H[q * 65 + q];
        double c = 1.0, s = 0.0;
        if (fabs(hpq) > 1.0e-16 * sqrt(fabs(hpp * hqq)) && fabs(hpq) > 1e-300) {
          const double th = (hqq - hpp) / (2.0 * hpq);
          const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(1.0 + th * th));
          c = 1.0 / sqrt(1.0 + t * t); s = t * c;
          atomicAdd(&nrot, 1);
        }
        cs[tid] = c; sn[tid] = s;
      }
      __syncthreads();
      // columns: X[:, p], X[:, q] <- X J for H and V
      for (int idx = tid; idx < 32 * 64; idx += 256) {
        const int pr = idx >> 6, r = idx & 63;
        const double c = cs[pr], s = sn[pr];
        if (s == 0.0) continue;
        const int p = pp[pr], q = qq[pr];
        const double hp = H[r * 65 + p], hq = H[r * 65 + q];
        H[r * 65 + p] = c * hp - s * hq; H[r * 65 + q] = s * hp + c * hq;
        const double vp = V[r * 65 + p], vq = V[r * 65 + q];
        V[r * 65 + p] = c * vp - s * vq; V[r * 65 + q] = s * vp + c * vq;
      }
      __syncthreads();
      // rows: H[p, :], H[q, :] <- J^T H ; the annihilated pair is set to exactly zero
      for (int idx = tid; idx < 32 * 64; idx += 256) {
        const int pr = idx >> 6, cc = idx & 63;
        const double c = cs[pr], s = sn[pr];
        if (s == 0.0) continue;
        const int p = pp[pr], q = qq[pr];
        const double hp = H[p * 65 + cc], hq = H[q * 65 + cc];
        double np_ = c * hp - s * hq, nq_ = s * hp + c * hq;
        if (cc == q) np_ = 0.0;
        if (cc == p) nq_ = 0.0;
        H[p * 65 + cc] = np_; H[q * 65 + cc] = nq_;
      }
      __syncthreads();
    }
    if (nrot == 0) break;
    __syncthreads();
  }
  // sort descending
  if (tid < 64) {
    const double me = H[tid * 65 + tid];
    int rk = 0;
    for (int j = 0; j < 64; j++) { const double o = H[j * 65 + j]; rk += (o > me) || (o == me && j < tid); }
    rank[tid] = rk;
    theta[rk] = me;
  }
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += 256) {
    const int k = idx >> 6, c = idx & 63;
    Z[k * 64 + rank[c]] = V[k * 65 + c];
  }
}

// G = sum Gpart (+ shift * trace on the diagonal) = R^T R ; out Rinv (row-major, upper) so that W Rinv is orthonormal.
__global__ void __launch_bounds__(256) cholinv64_kernel(const double* __restrict__ Gpart, int nchunk, double shift_rel, double* __restrict__ Rinv,
                                                        int* __restrict__ err) {
  extern __shared__ __align__(16) double csm[];
  double* Gs = csm; double* Ri = csm + 64 * 65;
  __shared__ double tr;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 4096; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    double a = 0.0, b = 0.0;
    for (int ch = 0; ch < nchunk; ch++) { a += Gpart[(size_t)ch * 4096 + r * 64 + c]; b += Gpart[(size_t)ch * 4096 + c * 64 + r]; }
    Gs[r * 65 + c] = 0.5 * (a + b);
  }
  __syncthreads();
  if (tid == 0) { double s = 0.0; for (int k = 0; k < 64; k++) s += Gs[k * 65 + k]; tr = s; }
  __syncthreads();
  if (tid < 64) Gs[tid * 65 + tid] += shift_rel * tr;
  __syncthreads();
  // right-looking Cholesky, upper factor R stored in the upper triangle of Gs: G = R^T R
  for (int k = 0; k < 64; k++) {
    if (tid == 0) {
      double d = Gs[k * 65 + k];
      if (!(d > 0.0)) { atomicExch(err, 2); d = 1.0; }
      Gs[k * 65 + k] = sqrt(d);
    }
    __syncthreads();
    const double rkk = Gs[k * 65 + k];
    if (tid > k && tid < 64) Gs[k * 65 + tid] /= rkk;
    __syncthreads();
    for (int idx = tid; idx < 4096; idx += 256) {
      const int r = idx >> 6, c = idx & 63;
      if (r > k && c >= r) Gs[r * 65 + c] -= Gs[k * 65 + r] * Gs[k * 65 + c];
    }
    __syncthreads();
  }
  // Rinv: solve R x = e_c column by column (thread c), x upper triangular
  if (tid < 64) {
    const int c = tid;
    for (int r = 63; r >= 0; r--) {
      double v = 0.0;
      if (r <= c) {
        v = r == c ? 1.0 : 0.0;
        for (int m = r + 1; m <= c; m++) v -= Gs[r * 65 + m] * Ri[m * 65 + c];
        v /= Gs[r * 65 + r];
      }
      Ri[r * 65 + c] = v;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 4096; idx += 256) Rinv[idx] = Ri[(idx >> 6) * 65 + (idx & 63)];
}

// res2[c] = sum_i (AV[c][i] - theta[c] V[c][i])^2 : one CTA per column (fixed-order reduction)
__global__ void __launch_bounds__(256) resnorm_kernel(const double* __restrict__ AVt, const double* __restrict__ Vt, int64_t ld, int n,
                                                      const double* __restrict__ theta, double* __restrict__ res2) {
  __shared__ double red[256];
  const int c = blockIdx.x;
  const double th = theta[c];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { const double r = AVt[(size_t)c * ld + i] - th * Vt[(size_t)c * ld + i]; s += r * r; }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) res2[c] = red[0];
}

__global__ void __launch_bounds__(256) randinit_kernel(double* __restrict__ Vt, int64_t ld, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
  if (i >= n) return;
  uint64_t x = ((uint64_t)(c + 1) << 40) ^ (uint64_t)i * 0x9E3779B97F4A7C15ull;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  Vt[(size_t)c * ld + i] = ((double)(x >> 11) * (1.0 / 9007199254740992.0)) - 0.5;
}

__global__ void __launch_bounds__(256) normalize_rows_kernel(double* __restrict__ Vt, int64_t ld, int n) {
  __shared__ double red[256];
  double* v = Vt + (size_t)blockIdx.x * ld;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += v[i] * v[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  const double inv = 1.0 / sqrt(red[0]);
  for (int i = threadIdx.x; i < n; i += 256) v[i] *= inv;
}


// =================================================================================================== eigenvector back-transformation
// x = Q1 Q2 z for the leading eigenvectors z of the tridiagonal matrix: Q1 = product of the stage-1 block reflectors I - V T V^T (V kept
// below the band in the panel's own dead columns of the working matrix, T in a side buffer), Q2 = product of the bulge-chasing reflectors.
// For the 10-40 vectors smartpca prints this is O(n^2 nvec) work against the ~270 n^2 x 64 block mat-vecs of the subspace iteration.

// A[r0 + gi][j + cc] = V[cc][r0 + gi] for gi > cc (strictly below the unit diagonal: outside the band, never read again by stage 1 / 2)
__global__ void __launch_bounds__(256) save_v_kernel(double* __restrict__ A, int64_t lda, int n, int j, const double* __restrict__ VZ, int64_t ldv) {
  __shared__ double t[64][65];
  const int r0 = j + BW, i0 = blockIdx.x * 64;
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    const int cc = idx >> 6, r = idx & 63, gi = i0 + r;
    t[r][cc] = (r0 + gi < n) ? VZ[(size_t)cc * ldv + r0 + gi] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    const int r = idx >> 6, cc = idx & 63, gi = i0 + r;
    if (r0 + gi < n && gi > cc) A[(size_t)(r0 + gi) * lda + j + cc] = t[r][cc];
  }
}

// Z <- Q2 Z: sweeps in reverse order; the reflectors of one sweep act on disjoint 64-row blocks (one warp each), a grid barrier separates
// the sweeps.  Z: [nvec][ldz].
constexpr int Q2_THREADS = 512;
__global__ void __launch_bounds__(Q2_THREADS, 1) q2_apply_kernel(const double* __restrict__ Q2, int n, int nvec, double* __restrict__ Z, int64_t ldz,
                                                                  int* __restrict__ bar, int* __restrict__ err) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = Q2_THREADS / 32;
  const int G = gridDim.x;
  int bar_target = 0;
  for (int s = n - 3; s >= 0; s--) {
    const int L = n - 1 - s, nref = (L + 63) >> 6;
    const double* q = Q2 + q2_off(s, n);
    for (int k = blockIdx.x * nw + warp; k < nref; k += G * nw) {
      const int len = min(64, L - 64 * k), row0 = s + 1 + 64 * k;
      double v0 = lane < len ? __ldcs(&q[64 * k + lane]) : 0.0;
      const double v1 = lane + 32 < len ? __ldcs(&q[64 * k + lane + 32]) : 0.0;
      const double tau = __shfl_sync(0xffffffffu, v0, 0);
      if (lane == 0) v0 = 1.0;
      if (tau != 0.0) {
        for (int j0 = 0; j0 < nvec; j0 += 8) {
          double z0[8], z1[8], d[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const bool ok = j0 + j < nvec;
            double* zr = Z + (size_t)(j0 + j) * ldz + row0;
            z0[j] = (ok && lane < len) ? __ldcg(&zr[lane]) : 0.0;
            z1[j] = (ok && lane + 32 < len) ? __ldcg(&zr[lane + 32]) : 0.0;
            d[j] = v0 * z0[j] + v1 * z1[j];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < 8; j++) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (j0 + j < nvec) {
              double* zr = Z + (size_t)(j0 + j) * ldz + row0;
              const double f = tau * d[j];
              if (lane < len) __stcg(&zr[lane], z0[j] - f * v0);
              if (lane + 32 < len) __stcg(&zr[lane + 32], z1[j] - f * v1);
            }
          }
        }
      }
    }
    if (G > 1) grid_barrier(bar, bar_target += G, err);
    else __syncthreads();
  }
}

// stage-1 block reflector of the panel at column j: Gpart[chunk][v][c] = sum over the chunk's rows of V[i][c] Z[v][i]
constexpr int Q1_ROWS = 256;       // rows per CTA
constexpr int Q1_MAXV = 40;
__device__ __forceinline__ double q1_v(const double* __restrict__ A, int64_t lda, int r0, int j, int i, int c) {
  const int gi = i - r0;
  return gi > c ? A[(size_t)i * lda + j + c] : (gi == c ? 1.0 : 0.0);
}
__global__ void __launch_bounds__(256) q1_dot_kernel(const double* __restrict__ A, int64_t lda, int n, int j, int nvec, const double* __restrict__ Z,
                                                     int64_t ldz, double* __restrict__ Gpart) {
  __shared__ double red[4][64];
  const int r0 = j + BW, c = threadIdx.x & 63, part = threadIdx.x >> 6;
  const int i0 = r0 + blockIdx.x * Q1_ROWS, i1 = min(n, i0 + Q1_ROWS);
  double acc[Q1_MAXV];
#pragma unroll
  for (int v = 0; v < Q1_MAXV; v++) acc[v] = 0.0;
  for (int i = i0 + part; i < i1; i += 4) {
    const double vv = q1_v(A, lda, r0, j, i, c);
#pragma unroll
    for (int v = 0; v < Q1_MAXV; v++)
      if (v < nvec) acc[v] += vv * Z[(size_t)v * ldz + i];
  }
  for (int v = 0; v < nvec; v++) {
    double a = 0.0;
#pragma unroll
    for (int u = 0; u < Q1_MAXV; u++) if (u == v) a = acc[u];
    red[part][c] = a;
    __syncthreads();
    if (part == 0) Gpart[((size_t)blockIdx.x * nvec + v) * 64 + c] = red[0][c] + red[1][c] + red[2][c] + red[3][c];
    __syncthreads();
  }
}
// Z[v][i] -= sum_c V[i][c] (T G_v)[c],  G_v = sum over chunks (fixed order)
__global__ void __launch_bounds__(256) q1_update_kernel(const double* __restrict__ A, int64_t lda, int n, int j, int nvec, double* __restrict__ Z,
                                                        int64_t ldz, const double* __restrict__ Gpart, int nchunk, const double* __restrict__ T) {
  __shared__ double Gs[Q1_MAXV][64];
  __shared__ double TG[Q1_MAXV][64];
  double (*Vs)[65] = reinterpret_cast<double (*)[65]>(&Gs[0][0]);      // [32][65], reuses Gs once T G is formed
  const int r0 = j + BW;
  for (int idx = threadIdx.x; idx < nvec * 64; idx += 256) {
    double a = 0.0;
    for (int ch = 0; ch < nchunk; ch++) a += Gpart[(size_t)ch * nvec * 64 + idx];
    Gs[idx >> 6][idx & 63] = a;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nvec * 64; idx += 256) {
    const int v = idx >> 6, cp = idx & 63;
    double a = 0.0;
    for (int cc = cp; cc < 64; cc++) a += T[cp * 64 + cc] * Gs[v][cc];         // T upper triangular, row-major
    TG[v][cp] = a;
  }
  __syncthreads();
  const int i0 = r0 + blockIdx.x * Q1_ROWS;
  for (int b = 0; b < Q1_ROWS; b += 32) {
    for (int idx = threadIdx.x; idx < 32 * 64; idx += 256) {
      const int r = idx >> 6, cc = idx & 63, i = i0 + b + r;
      Vs[r][cc] = i < n ? q1_v(A, lda, r0, j, i, cc) : 0.0;
    }
    __syncthreads();
    const int r = threadIdx.x & 31, i = i0 + b + r;
    for (int v = threadIdx.x >> 5; v < nvec; v += 8) {
      double a = 0.0;
#pragma unroll 8
      for (int cc = 0; cc < 64; cc++) a += Vs[r][cc] * TG[v][cc];
      if (i < n) Z[(size_t)v * ldz + i] -= a;
    }
    __syncthreads();
  }
}

// =================================================================================================== host drivers
static int coop_launch(eb_ctx* c, const void* fn, dim3 grid, dim3 block, void** args, size_t smem) {
  EB_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, c->stream));
  c->launches++;
  return 0;
}

// chunk length for gram64: a multiple of 32, at least 256, at most ~2 CTAs per SM
static int gram_chunk(const eb_ctx* c, int len) {
  int ch = (len + 2 * c->num_sms - 1) / (2 * c->num_sms);
  ch = (std::max(ch, 256) + 31) & ~31;
  return ch;
}

struct Eig2Work {
  double *VZ, *Wpart, *Wt, *T, *C1, *C2, *gpart, *rowk, *G2part, *G2, *Spart, *AB, *gscratch;
  int* prog; int* err;
  int64_t ldv;
};

// ---- debugging aids (EB_DBG_REF_W / EB_DBG_REF_SYR2K): plain reference kernels for the two tensor-core products
__global__ void __launch_bounds__(256) dbg_ref_w_kernel(const double* __restrict__ A, int64_t lda, int n, int r0, const double* __restrict__ V,
                                                        int64_t ldv, double* __restrict__ Wt) {
  const int i = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
  if (i >= n) return;
  double s = 0.0;
  if (i >= r0)
    for (int k = r0; k < n; k++) s += (k <= i ? A[(size_t)i * lda + k] : A[(size_t)k * lda + i]) * V[(size_t)c * ldv + k];
  Wt[(size_t)c * ldv + i] = s;
}
__global__ void __launch_bounds__(256) dbg_ref_syr2k_kernel(double* __restrict__ A, int64_t lda, int n, int lo, const double* __restrict__ VZ, int64_t ldv) {
  const int col = lo + blockIdx.x * 256 + threadIdx.x, row = lo + blockIdx.y;
  if (col >= n || row >= n) return;
  if (col > row && (col >> 7) != (row >> 7)) return;          // lower tiles + complete diagonal 128-tiles
  double s = 0.0;
  for (int k = 0; k < 64; k++) s += VZ[(size_t)k * ldv + row] * VZ[(size_t)(64 + k) * ldv + col] + VZ[(size_t)(64 + k) * ldv + row] * VZ[(size_t)k * ldv + col];
  A[(size_t)row * lda + col] -= s;
}

// ---- row-distributed stage 1 (collective): every rank holds the FULL square and keeps its own 128-row tiles (tile % world == rank)
// current; W = A22 V and the rank-128 update run on the owned rows only, the 64 x n panel (before its QR) and the 64 x n product
// block (after it) are summed over the ranks -- the other ranks' rows are zero, so the sum is a gather and every rank ends up with
// bit-identical V, W, Z and band.  The panel QR and the O(n 64^2) glue run replicated.
// stage[cc][i] = A[i][col0 + cc] for owned rows i in [row0, n), zero elsewhere (cc < ncols; rows of the other 64 - ncols stay zero)
__global__ void __launch_bounds__(256) dist_pack_kernel(const double* __restrict__ A, int64_t lda, int n, int row0, int col0, int ncols,
                                                        int rank, int world, double* __restrict__ stage, int64_t lds) {
  __shared__ double t[64][65];
  const int i0 = row0 + blockIdx.x * 64;
  const bool own = ((i0 >> 7) % world) == rank;           // row0 and the 64-row blocks are 64-aligned: a block lies in one 128-row tile
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    const int r = idx >> 6, cc = idx & 63, i = i0 + r;
    t[r][cc] = (own && i < n && cc < ncols) ? A[(size_t)i * lda + col0 + cc] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    const int cc = idx >> 6, r = idx & 63, i = i0 + r;
    if (i < n) stage[(size_t)cc * lds + i] = t[r][cc];
  }
}
__global__ void __launch_bounds__(256) dist_unpack_kernel(double* __restrict__ A, int64_t lda, int n, int row0, int col0, int ncols,
                                                          const double* __restrict__ stage, int64_t lds) {
  __shared__ double t[64][65];
  const int i0 = row0 + blockIdx.x * 64;
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    const int cc = idx >> 6, r = idx & 63, i = i0 + r;
    t[r][cc] = i < n ? stage[(size_t)cc * lds + i] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
    const int r = idx >> 6, cc = idx & 63, i = i0 + r;
    if (i < n && cc < ncols) A[(size_t)i * lda + col0 + cc] = t[r][cc];
  }
}
// Out[c][i] = sum_ks Wpart[ks][c][i] for owned rows i >= i_lo, zero for every other row of [i_zero0, n)
__global__ void __launch_bounds__(256) dist_combine_kernel(const double* __restrict__ Wpart, int ksplit, int64_t ldw, double* __restrict__ Out,
                                                           int64_t ldo, int n, int i_zero0, int i_lo, int rank, int world) {
  const int i = i_zero0 + blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
  if (i >= n) return;
  double v = 0.0;
  if (i >= i_lo && ((i >> 7) % world) == rank)
    for (int ks = 0; ks < ksplit; ks++) v += Wpart[((size_t)ks * 64 + c) * ldw + i];
  Out[(size_t)c * ldo + i] = v;
}

// all ranks: columns [col0, col0 + ncols) of rows [row0, n) become current everywhere (row0 a multiple of 64)
static int dist_gather_cols(eb_ctx* c, double* A, int64_t lda, int n, int row0, int col0, int ncols, double* xbuf, int64_t ldx, bool& first) {
  cudaStream_t st = c->stream;
  int rc;
  double* stage = xbuf + (size_t)64 * ldx;
  const int nb = (n - row0 + 63) / 64;
  if (nb <= 0) return 0;
  if (row0 > 0) EB_CUDA(cudaMemsetAsync(stage, 0, sizeof(double) * 64 * ldx, st));     // rows above row0 must not carry old sums
  dist_pack_kernel<<<nb, 256, 0, st>>>(A, lda, n, row0, col0, ncols, c->comm.rank, c->comm.world, stage, ldx);
  EB_CHECK_LAUNCH(c);
  if ((rc = peer_allreduce_stream(c, PEER_SLOT_W, xbuf, c->chfsi_sum.n, (int64_t)64 * ldx, first, (int64_t)64 * ldx))) return rc;
  first = false;
  dist_unpack_kernel<<<nb, 256, 0, st>>>(A, lda, n, row0, col0, ncols, stage, ldx);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// Stage 1 + stage 2: A (n x n, lda, lower triangle + complete diagonal tiles valid; destroyed) -> d, e (unscaled).
// collective: all ranks of the communicator call with the SAME full symmetric matrix (both triangles valid); see above.
int two_stage_tridiag(eb_ctx* c, double* A, int64_t lda, int n, double* d, double* e, bool collective, bool keep_q) {
  cudaStream_t st = c->stream;
  int rc;
  // keep_q: the stage-1 block reflectors stay in the dead part of A (+ their T factors in c->eigT), the stage-2 reflectors go to c->eigQ2
  // (or into the GRM's dead lower-tile accumulator when that is large enough): two_stage_backtransform() turns eigenvectors of the
  // tridiagonal matrix into eigenvectors of A
  double* q2 = nullptr;
  c->eig_npanels = 0; c->eig_q2 = nullptr;
  if (keep_q) {
    const size_t need = q2_off(n - 2 > 0 ? n - 2 : 0, n) + 64;
    if (c->partial.p && c->partial.n >= need) q2 = c->partial.p;
    else { if ((rc = c->eigQ2.ensure(need))) return rc; q2 = c->eigQ2.p; }
    if ((rc = c->eigT.ensure((size_t)(n / BW + 2) * 4096))) return rc;
  }
  const bool dist = collective && c->has_comm && c->comm.world > 1;
  const int NW = dist ? c->comm.world : 1, me = dist ? c->comm.rank : 0;
  bool first_exchange = true;
  const int64_t ldv = ((int64_t)n + 7) & ~7ll;
  const int G_max = c->num_sms;
  const int nchunk_max = (n + 255) / 256 + 1;
  // workspace (doubles)
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += (cnt + 15) & ~size_t(15); return o; };
  const size_t oVZ = take((size_t)128 * ldv), oWp = take((size_t)MAX_KSPLIT * 64 * ldv), oWt = take((size_t)64 * ldv), oT = take(4096),
               oC1 = take(4096), oC2 = take(4096), ogp = take((size_t)2 * G_max * 64), ork = take(128), oG2p = take((size_t)G_max * 4096),
               oG2 = take(4096), oSp = take((size_t)nchunk_max * 4096), oAB = take((size_t)n * 2 * BW);
  // panel chunk: shared memory holds up to PANEL_SMEM_ROWS rows; beyond that the chunk lives in global scratch
  constexpr int SH_EXTRA = 512 * 8;                    // reduction scratch (doubles * 8)
  const int max_rows_smem = (int)((227 * 1024 - SH_EXTRA - 1024) / 512);
  int rows_per0 = std::max(128, (((n - BW + G_max - 1) / G_max) + 3) & ~3);
  const bool use_global = rows_per0 > max_rows_smem;
  const size_t oGs = take(use_global ? (size_t)G_max * rows_per0 * 64 : 16);
  const size_t oInt = take(((size_t)n + 64) / 2 + 16);
  if ((rc = c->eig2w.ensure(off))) return rc;
  double* W = c->eig2w.p;
  Eig2Work w;
  w.VZ = W + oVZ; w.Wpart = W + oWp; w.Wt = W + oWt; w.T = W + oT; w.C1 = W + oC1; w.C2 = W + oC2; w.gpart = W + ogp; w.rowk = W + ork;
  w.G2part = W + oG2p; w.G2 = W + oG2; w.Spart = W + oSp; w.AB = W + oAB; w.gscratch = W + oGs; w.ldv = ldv;
  w.prog = reinterpret_cast<int*>(W + oInt); w.err = w.prog + n + 8;
  EB_CUDA(cudaMemsetAsync(w.VZ, 0, sizeof(double) * 128 * ldv, st));
  EB_CUDA(cudaMemsetAsync(w.prog, 0, sizeof(int) * ((size_t)n + 16), st));
  double* xbuf = nullptr;                              // exported: [0, 64 ldv) product block, [64 ldv, 128 ldv) panel stage
  if (dist) {
    if ((rc = c->chfsi_sum.ensure((size_t)128 * ldv))) return rc;
    xbuf = c->chfsi_sum.p;
    EB_CUDA(cudaMemsetAsync(xbuf, 0, sizeof(double) * 128 * ldv, st));
  }

  // per device, not per process: set on every call (several contexts on different GPUs may live in one process)
  EB_CUDA(cudaFuncSetAttribute(panel_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  EB_CUDA(cudaFuncSetAttribute(left_mult_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
  EB_CUDA(cudaFuncSetAttribute(sbr_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 8));

  const bool dbg_sync = getenv("EB_DBG_SYNC") != nullptr, dbg_ref_w = getenv("EB_DBG_REF_W") != nullptr,
             dbg_ref_syr2k = getenv("EB_DBG_REF_SYR2K") != nullptr;
  // EB_EIG_PROFILE=1: per-phase device time of stage 1 (events around every phase, one sync per panel) on stderr
  const bool prof = getenv("EB_EIG_PROFILE") != nullptr;
  cudaEvent_t pe[6] = {};
  double pacc[5] = {0, 0, 0, 0, 0};
  if (prof) for (auto& e_ : pe) cudaEventCreate(&e_);
  auto mark = [&](int i) { if (prof) cudaEventRecord(pe[i], st); };
  int j_done = 0;
  for (int j = 0; n - j - BW >= 2; j += BW) {
    const int r0 = j + BW, np = n - r0, t0 = r0 / DT_M;
    // ---- distributed: the panel (diagonal block included) is current only on the owners of its rows: gather it
    mark(0);
    if (dist && (rc = dist_gather_cols(c, A, lda, n, j, j, BW, xbuf, ldv, first_exchange))) return rc;
    mark(1);
    // ---- panel QR
    PanelParams pp;
    pp.A = A; pp.lda = lda; pp.n = n; pp.j = j; pp.VZ = w.VZ; pp.ldv = ldv; pp.T = w.T; pp.gpart = w.gpart; pp.rowk = w.rowk;
    pp.G2part = w.G2part; pp.G2 = w.G2; pp.gscratch = w.gscratch; pp.use_global = use_global ? 1 : 0;
    pp.bar = w.err + 1; pp.err = w.err;
    EB_CUDA(cudaMemsetAsync(pp.bar, 0, sizeof(int), st));
    int rows_per = std::max(128, (((np + G_max - 1) / G_max) + 3) & ~3);
    if (use_global) rows_per = rows_per0;
    const int G = (np + rows_per - 1) / rows_per;
    pp.rows_per = rows_per;
    const size_t smem = use_global ? (size_t)(8192 + 512) * 8 : (size_t)rows_per * 512 + SH_EXTRA;
    void* args[] = {&pp};
    if ((rc = coop_launch(c, (const void*)panel_qr_kernel, dim3(G), dim3(256), args, smem))) return rc;
    if (keep_q) {
      save_v_kernel<<<(np + 63) / 64, 256, 0, st>>>(A, lda, n, j, w.VZ, ldv);
      EB_CHECK_LAUNCH(c);
      EB_CUDA(cudaMemcpyAsync(c->eigT.p + (size_t)c->eig_npanels * 4096, w.T, sizeof(double) * 4096, cudaMemcpyDeviceToDevice, st));
      c->eig_npanels++;
    }
    mark(2);
    // ---- W = A22 V
    int ksplit = 1;
    const int i_z0 = t0 * DT_M;
    if (dbg_ref_w) {
      dbg_ref_w_kernel<<<dim3((n + 255) / 256, 64), 256, 0, st>>>(A, lda, n, r0, w.VZ, ldv, w.Wt);
      EB_CHECK_LAUNCH(c);
    } else if (dist) {
      // owned row tiles only (full rows of the stored square), then the sum over ranks = gather of the 64 x n block
      if ((rc = launch_sym_skinny(c, A, lda, n, t0, w.VZ, ldv, w.Wpart, ldv, MAX_KSPLIT, &ksplit, (me - t0 % NW + NW) % NW, NW, true))) return rc;
      dim3 grid((n + 255) / 256, 64);                 // every row: the block is summed over the ranks as a whole, stale rows must be zero
      dist_combine_kernel<<<grid, 256, 0, st>>>(w.Wpart, ksplit, ldv, xbuf, ldv, n, 0, r0, me, NW);
      EB_CHECK_LAUNCH(c);
      if ((rc = peer_allreduce_stream(c, PEER_SLOT_W, xbuf, c->chfsi_sum.n, (int64_t)64 * ldv, first_exchange, 0))) return rc;
      first_exchange = false;
      EB_CUDA(cudaMemcpyAsync(w.Wt, xbuf, sizeof(double) * 64 * ldv, cudaMemcpyDeviceToDevice, st));
    } else {
      if ((rc = launch_sym_skinny(c, A, lda, n, t0, w.VZ, ldv, w.Wpart, ldv, MAX_KSPLIT, &ksplit))) return rc;
      dim3 grid((n - i_z0 + 255) / 256, 64);
      combine_kernel<<<grid, 256, 0, st>>>(w.Wpart, ksplit, ldv, w.Wt, ldv, n, i_z0, r0, 1.0, nullptr, 0.0, nullptr, 0.0, ldv);
      EB_CHECK_LAUNCH(c);
    }
    mark(3);
    // ---- S = V^T W ; C1 = T ; C2 = -T^T S T / 2 ; Z = W C1 + V C2
    const int gch = gram_chunk(c, n - r0);
    const int nchunk = (n - r0 + gch - 1) / gch;
    gram64_kernel<<<nchunk, 256, 0, st>>>(w.VZ, ldv, w.Wt, ldv, r0, n, gch, w.Spart);
    EB_CHECK_LAUNCH(c);
    sbr_small_kernel<<<1, 256, 12288 * 8, st>>>(w.Spart, nchunk, w.T, w.C1, w.C2);
    EB_CHECK_LAUNCH(c);
    left_mult_kernel<<<(n - i_z0 + 63) / 64, 256, 16384 * 8, st>>>(w.C1, w.Wt, w.C2, w.VZ, ldv, w.VZ + (size_t)64 * ldv, ldv, i_z0, n);
    EB_CHECK_LAUNCH(c);
    mark(4);
    // ---- A22 -= V Z^T + Z V^T
    if (dbg_ref_syr2k) {
      dbg_ref_syr2k_kernel<<<dim3((n - i_z0 + 255) / 256, n - i_z0), 256, 0, st>>>(A, lda, n, i_z0, w.VZ, ldv);
      EB_CHECK_LAUNCH(c);
    } else if (dist) {
      if ((rc = launch_syr2k_lower(c, A, lda, n, t0, w.VZ, ldv, (me - t0 % NW + NW) % NW, NW))) return rc;
    } else if ((rc = launch_syr2k_lower(c, A, lda, n, t0, w.VZ, ldv))) return rc;
    if (dbg_sync) EB_CUDA(cudaStreamSynchronize(st));
    if (prof) {
      mark(5);
      EB_CUDA(cudaStreamSynchronize(st));
      for (int i = 0; i < 5; i++) { float ms = 0; cudaEventElapsedTime(&ms, pe[i], pe[i + 1]); pacc[i] += ms; }
    }
    j_done = j + BW;
  }
  if (prof) {
    fprintf(stderr, "[eig profile] n %d stage 1: panel gather %.1f ms, panel QR %.1f ms, W = A V (+ sum over ranks) %.1f ms, 64x64 glue %.1f ms, rank-128 update %.1f ms\n", n,
            pacc[0], pacc[1], pacc[2], pacc[3], pacc[4]);
    for (auto& e_ : pe) cudaEventDestroy(e_);
  }
  if (dist) {
    // the trailing block that got no panel of its own (2 .. 65 columns): its band entries live with the owners of its rows
    const int row0 = j_done & ~63;
    for (int col0 = row0; col0 < n; col0 += 64)
      if ((rc = dist_gather_cols(c, A, lda, n, row0, col0, std::min(64, n - col0), xbuf, ldv, first_exchange))) return rc;
  }

  // ---- stage 2
  EB_CUDA(cudaEventRecord(c->ev[10], st));            // end of dense -> band (eb_timings.band_ms / chase_ms)
  band_extract_kernel<<<n, 128, 0, st>>>(A, lda, n, w.AB);
  EB_CHECK_LAUNCH(c);
  if (c->dbg_band_h) {   // testing aid: the band matrix before bulge chasing ([n][128], AB[col][d] = A[col+d][col])
    EB_CUDA(cudaMemcpyAsync(c->dbg_band_h, w.AB, sizeof(double) * (size_t)n * 2 * BW, cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaStreamSynchronize(st));
  }
  if (n > 2) {
    int per_sm = 0;
    EB_CUDA(cudaFuncSetAttribute(bulge_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM));
    EB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bulge_chase_kernel, BC_THREADS, BC_SMEM));
    per_sm = std::max(1, std::min(per_sm, 3));
    // sweeps run two blocks apart, so at most ~n/128 are ever active: idle CTAs would only add polling traffic
    const int useful = n / (2 * BW) + 4;
    const int grid = std::max(1, std::min(std::min(per_sm * c->num_sms, n - 2), useful));
    double* ab = w.AB; int nn = n; int* prog = w.prog; int* err = w.err;
    void* args[] = {&ab, &nn, &prog, &err, &q2};
    if ((rc = coop_launch(c, (const void*)bulge_chase_kernel, dim3(grid), dim3(BC_THREADS), args, BC_SMEM))) return rc;
  }
#ifdef EB_BC_PROFILE
  {
    unsigned long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_bc_prof, sizeof(h));
    const double k = h[7] ? (double)h[7] : 1.0;
    fprintf(stderr, "[bc profile] steps %llu cycles/step: other %.0f wait %.0f load-issue %.0f load-done+smem %.0f sync1 %.0f compute %.0f store+publish %.0f\n", h[7],
            h[0] / k, h[1] / k, h[2] / k, h[3] / k, h[4] / k, h[5] / k, h[6] / k);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(g_bc_prof, h, sizeof(h));
  }
#endif
  band_de_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.AB, n, d, e);
  EB_CHECK_LAUNCH(c);
  int herr = 0;
  EB_CUDA(cudaMemcpyAsync(&herr, w.err, sizeof(int), cudaMemcpyDeviceToHost, st));
  EB_CUDA(cudaStreamSynchronize(st));
  if (herr) { set_error("two_stage_tridiag: a device-side wait timed out (code %d: 1 = bulge-chase predecessor, 3 = panel grid barrier)", herr); return EB_ERR_NUMERIC; }
  c->eig_q2 = q2;
  return 0;
}

// Z [nvec][ldz]: eigenvectors of the tridiagonal matrix -> eigenvectors of the matrix two_stage_tridiag(keep_q) reduced (A is its
// working copy with the stage-1 reflectors in place).  Replicated on every rank of a collective solve (O(n^2 nvec) work).
int two_stage_backtransform(eb_ctx* c, const double* A, int64_t lda, int n, int nvec, double* Z, int64_t ldz) {
  cudaStream_t st = c->stream;
  int rc;
  if (!c->eig_q2 || nvec > Q1_MAXV) { set_error("two_stage_backtransform: reflectors were not kept (or more than %d vectors)", Q1_MAXV); return EB_ERR_STATE; }
  if (n > 2) {
    int* bar = reinterpret_cast<int*>(c->eig2w.p);            // the stage-1 workspace is dead: [0] barrier counter, [1] error flag
    EB_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(int), st));
    int per_sm = 0;
    EB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q2_apply_kernel, Q2_THREADS, 0));
    const int want = ((n + 63) / 64 + Q2_THREADS / 32 - 1) / (Q2_THREADS / 32);       // one warp per reflector of the longest sweep
    int grid = std::max(1, std::min(want, std::max(1, per_sm) * c->num_sms));
    const double* q2 = c->eig_q2; int nn = n, nv = nvec; int64_t ld = ldz; int* err = bar + 1;
    void* args[] = {&q2, &nn, &nv, &Z, &ld, &bar, &err};
    if ((rc = coop_launch(c, (const void*)q2_apply_kernel, dim3(grid), dim3(Q2_THREADS), args, 0))) return rc;
  }
  if (getenv("EB_EIG_PROFILE")) {
    static double t_prev = 0;
    cudaStreamSynchronize(st);
    fprintf(stderr, "[eig profile] vectors: Q2 applied at cpu clock %.1f ms (delta to previous print)\n", ((double)clock() / CLOCKS_PER_SEC - t_prev) * 1e3);
    t_prev = (double)clock() / CLOCKS_PER_SEC;
  }
  const int G_max = (n + Q1_ROWS - 1) / Q1_ROWS + 1;
  if ((rc = c->eigW.ensure((size_t)G_max * Q1_MAXV * 64))) return rc;
  for (int p = c->eig_npanels - 1; p >= 0; p--) {
    const int j = p * BW, r0 = j + BW, np = n - r0;
    const int nchunk = (np + Q1_ROWS - 1) / Q1_ROWS;
    q1_dot_kernel<<<nchunk, 256, 0, st>>>(A, lda, n, j, nvec, Z, ldz, c->eigW.p);
    EB_CHECK_LAUNCH(c);
    q1_update_kernel<<<nchunk, 256, 0, st>>>(A, lda, n, j, nvec, Z, ldz, c->eigW.p, nchunk, c->eigT.p + (size_t)p * 4096);
    EB_CHECK_LAUNCH(c);
  }
  int herr = 0;
  if (n > 2) {
    EB_CUDA(cudaMemcpyAsync(&herr, reinterpret_cast<int*>(c->eig2w.p) + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaStreamSynchronize(st));
    if (herr) { set_error("two_stage_backtransform: grid barrier timed out"); return EB_ERR_NUMERIC; }
  }
  return 0;
}

// Cm[j][c] = dvec[j] * sum_chunks Gpart[chunk][j][c]   (coefficients of the deflation term L^T diag(d) (L X^T))
__global__ void __launch_bounds__(256) defl_coef_kernel(const double* __restrict__ Gpart, int nchunk, const double* __restrict__ dvec,
                                                        double* __restrict__ Cm) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  double v = 0.0;
  for (int ch = 0; ch < nchunk; ch++) v += Gpart[(size_t)ch * 4096 + idx];
  Cm[idx] = v * dvec[idx >> 6];
}

// Leading nvec eigenpairs of the symmetric matrix A (n x n, lda; lower triangle + complete diagonal tiles valid, preserved).
// theta_h[nvec] (unscaled), vec_d: device [nvec][n] unit vectors.
// lo0: lower end of the spectrum when the caller knows it (the bisection's smallest eigenvalue), else 0 (a GRM is positive
// semi-definite); the damped interval of the filter starts there, so an indefinite matrix needs it.
// c->tm.chfsi_converged / chfsi_resid tell the caller whether the strict tolerance was reached (see eb_timings).
// collective: every rank of the communicator holds the SAME matrix (the reduced GRM of a sharded pass) and calls this together; the
// block mat-vecs are then split by row tiles of A over the ranks and the 64 x n result block is summed (= gathered: the other
// ranks' rows are zero) by the stream-ordered all-reduce over peer memory, bit-identical on every rank.  Everything else of the
// iteration is O(n 64^2) and runs replicated.
__global__ void __launch_bounds__(256) combine_own_kernel(const double* __restrict__ Wpart, int ksplit, int64_t ldw, double* __restrict__ Out,
                                                          int64_t ldo, int n, int tile_first, int tile_stride) {
  const int i = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
  if (i >= n) return;
  const int T = i >> 7;
  if (T < tile_first || (T - tile_first) % tile_stride != 0) return;        // not my row tile: stays zero
  double v = 0.0;
  for (int ks = 0; ks < ksplit; ks++) v += Wpart[((size_t)ks * 64 + c) * ldw + i];
  Out[(size_t)c * ldo + i] = v;
}

int chfsi_top(eb_ctx* c, const double* A, int64_t lda, int n, int nvec, double* theta_h, double* vec_d, int* iters_out, int* matvecs_out,
              double lo0, bool collective) {
  cudaStream_t st = c->stream;
  int rc;
  const int64_t ld = ((int64_t)n + 7) & ~7ll;
  const int gch = gram_chunk(c, n);
  const int nchunk = (n + gch - 1) / gch;
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += (cnt + 15) & ~size_t(15); return o; };
  size_t oY[5];
  for (int i = 0; i < 5; i++) oY[i] = take((size_t)64 * ld);
  const size_t oL = take((size_t)64 * ld), oCorr = take((size_t)64 * ld), oCm = take(4096), oDv = take(64);
  const size_t oWp = take((size_t)MAX_KSPLIT * 64 * ld), oGp = take((size_t)nchunk * 4096), oZ = take(4096), oTh = take(64), oRes = take(64),
               oErr = take(16);
  if ((rc = c->chfsiw.ensure(off))) return rc;
  double* W = c->chfsiw.p;
  double* Y[5];
  for (int i = 0; i < 5; i++) Y[i] = W + oY[i];
  double *Wpart = W + oWp, *Gpart = W + oGp, *Zm = W + oZ, *theta_d = W + oTh, *res_d = W + oRes;
  int* err_d = reinterpret_cast<int*>(W + oErr);
  double *Lk = W + oL, *Corr = W + oCorr, *Cm = W + oCm, *dvec_d = W + oDv;
  EB_CUDA(cudaMemsetAsync(Lk, 0, sizeof(double) * 64 * ld, st));
  EB_CUDA(cudaMemsetAsync(err_d, 0, sizeof(int), st));
  for (int i = 0; i < 5; i++) EB_CUDA(cudaMemsetAsync(Y[i], 0, sizeof(double) * 64 * ld, st));
  constexpr int SM64x2 = 2 * 64 * 65 * 8;
  EB_CUDA(cudaFuncSetAttribute(left_mult_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
  EB_CUDA(cudaFuncSetAttribute(jacobi64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM64x2));
  EB_CUDA(cudaFuncSetAttribute(cholinv64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM64x2));
  const dim3 cgrid((n + 255) / 256, 64);
  const int lm_grid = (n + 63) / 64;
  int nmat = 0;

  // Locked (converged) leading Ritz vectors live in Lk (rows >= nlock are zero).  With nlock > 0 the filter runs on the
  // deflated operator A' = A - L^T diag(theta_j - c) L, which moves the locked eigenvalues to the centre of the damped
  // interval: outlier eigenvalues (population structure puts them 100-1000x above the bulk) then no longer bound the
  // filter degree through the 1e8 dynamic-range budget, and rounding-level components along them cannot grow.
  int nlock = 0;
  const bool shard = collective && c->has_comm && c->comm.world > 1;
  bool first_exchange = true;
  if (shard && (rc = c->chfsi_sum.ensure((size_t)64 * ld))) return rc;
  auto matvec = [&](const double* X, double* Out, double a1, const double* Y1, double a2, const double* Y0, double a3) -> int {
    int ks = 1, r;
    if (shard) {
      const int W = c->comm.world, me = c->comm.rank;
      if ((r = launch_sym_skinny(c, A, lda, n, 0, X, ld, Wpart, ld, MAX_KSPLIT, &ks, me, W))) return r;
      EB_CUDA(cudaMemsetAsync(c->chfsi_sum.p, 0, sizeof(double) * 64 * ld, st));
      if (ks > 0) {
        combine_own_kernel<<<cgrid, 256, 0, st>>>(Wpart, ks, ld, c->chfsi_sum.p, ld, n, me, W);
        EB_CHECK_LAUNCH(c);
      }
      if ((r = peer_allreduce_stream(c, PEER_SLOT_W, c->chfsi_sum.p, c->chfsi_sum.n, (int64_t)64 * ld, first_exchange))) return r;
      first_exchange = false;
      if (nlock > 0) {
        gram64_kernel<<<nchunk, 256, 0, st>>>(Lk, ld, X, ld, 0, n, gch, Gpart);
        EB_CHECK_LAUNCH(c);
        defl_coef_kernel<<<16, 256, 0, st>>>(Gpart, nchunk, dvec_d, Cm);
        EB_CHECK_LAUNCH(c);
        left_mult_kernel<<<lm_grid, 256, 16384 * 8, st>>>(Cm, Lk, nullptr, nullptr, ld, Corr, ld, 0, n);
        EB_CHECK_LAUNCH(c);
        combine_kernel<<<cgrid, 256, 0, st>>>(c->chfsi_sum.p, 1, ld, Out, ld, n, 0, 0, a1, Y1, a2, Y0, a3, ld, Corr, -a1);
      } else {
        combine_kernel<<<cgrid, 256, 0, st>>>(c->chfsi_sum.p, 1, ld, Out, ld, n, 0, 0, a1, Y1, a2, Y0, a3, ld);
      }
      EB_CHECK_LAUNCH(c);
      nmat++;
      return 0;
    }
    if ((r = launch_sym_skinny(c, A, lda, n, 0, X, ld, Wpart, ld, MAX_KSPLIT, &ks))) return r;
    if (nlock > 0) {
      gram64_kernel<<<nchunk, 256, 0, st>>>(Lk, ld, X, ld, 0, n, gch, Gpart);
      EB_CHECK_LAUNCH(c);
      defl_coef_kernel<<<16, 256, 0, st>>>(Gpart, nchunk, dvec_d, Cm);
      EB_CHECK_LAUNCH(c);
      left_mult_kernel<<<lm_grid, 256, 16384 * 8, st>>>(Cm, Lk, nullptr, nullptr, ld, Corr, ld, 0, n);
      EB_CHECK_LAUNCH(c);
      combine_kernel<<<cgrid, 256, 0, st>>>(Wpart, ks, ld, Out, ld, n, 0, 0, a1, Y1, a2, Y0, a3, ld, Corr, -a1);
    } else {
      combine_kernel<<<cgrid, 256, 0, st>>>(Wpart, ks, ld, Out, ld, n, 0, 0, a1, Y1, a2, Y0, a3, ld);
    }
    EB_CHECK_LAUNCH(c);
    nmat++;
    return 0;
  };
  // X <- orthonormal basis of span(X) (shifted CholeskyQR, then two plain passes); result ends up in *X, *tmp is scratch
  auto chol_qr = [&](double*& X, double*& tmp) -> int {
    for (int pass = 0; pass < 3; pass++) {
      gram64_kernel<<<nchunk, 256, 0, st>>>(X, ld, X, ld, 0, n, gch, Gpart);
      EB_CHECK_LAUNCH(c);
      const double shift = pass == 0 ? 11.0 * (64.0 * n + 64.0 * 65.0) * 1.1e-16 : 0.0;
      cholinv64_kernel<<<1, 256, SM64x2, st>>>(Gpart, nchunk, shift, Zm, err_d);
      EB_CHECK_LAUNCH(c);
      left_mult_kernel<<<lm_grid, 256, 16384 * 8, st>>>(Zm, X, nullptr, nullptr, ld, tmp, ld, 0, n);
      EB_CHECK_LAUNCH(c);
      std::swap(X, tmp);
    }
    return 0;
  };

  double *V = Y[0], *AV = Y[1], *Ya = Y[2], *Yb = Y[3], *Yc = Y[4];
  randinit_kernel<<<cgrid, 256, 0, st>>>(V, ld, n);
  EB_CHECK_LAUNCH(c);
  if ((rc = chol_qr(V, Ya))) return rc;

  double th[64], res2[64];
  const bool debug = getenv("EB_DEBUG") != nullptr;
  const double tol = std::max(2e-14, 6e-16 * sqrt((double)n));
  double prev_worst = 1e300, prev2_worst = 1e300, last_worst = 1e300;
  double lo = std::min(0.0, lo0);
  int outer = 0;
  bool converged = false, strict = false;
  const int maxouter = 80;
  for (outer = 0; outer < maxouter; outer++) {
    // Rayleigh-Ritz (on A itself)
    {
      const int keep = nlock;
      nlock = 0;
      rc = matvec(V, AV, 1.0, nullptr, 0.0, nullptr, 0.0);
      nlock = keep;
      if (rc) return rc;
    }
    gram64_kernel<<<nchunk, 256, 0, st>>>(V, ld, AV, ld, 0, n, gch, Gpart);
    EB_CHECK_LAUNCH(c);
    jacobi64_kernel<<<1, 256, SM64x2, st>>>(Gpart, nchunk, theta_d, Zm);
    EB_CHECK_LAUNCH(c);
    left_mult_kernel<<<lm_grid, 256, 16384 * 8, st>>>(Zm, V, nullptr, nullptr, ld, Ya, ld, 0, n);
    EB_CHECK_LAUNCH(c);
    left_mult_kernel<<<lm_grid, 256, 16384 * 8, st>>>(Zm, AV, nullptr, nullptr, ld, Yb, ld, 0, n);
    EB_CHECK_LAUNCH(c);
    std::swap(V, Ya); std::swap(AV, Yb);
    resnorm_kernel<<<64, 256, 0, st>>>(AV, V, ld, n, theta_d, res_d);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaMemcpyAsync(th, theta_d, sizeof(double) * 64, cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaMemcpyAsync(res2, res_d, sizeof(double) * 64, cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaStreamSynchronize(st));
    const double anorm = std::max(fabs(th[0]), fabs(th[63]));
    // lock the leading Ritz pairs that have reached the final tolerance (monotone: a locked pair stays locked)
    while (nlock < nvec && sqrt(res2[nlock]) <= tol * anorm) nlock++;
    double worst = 0.0;
    for (int i = nlock; i < nvec; i++) worst = std::max(worst, sqrt(res2[i]));
    if (debug) fprintf(stderr, "[chfsi] outer %d matvecs %d theta0 %.6e theta_k %.6e cut %.6e worst_res/anorm %.3e\n", outer, nmat, th[0],
                       th[std::max(nvec - 1, 0)], th[63], worst / anorm);
    // converged: residual at the rounding floor of an n-term FP64 mat-vec, or stagnating just above it
    last_worst = anorm > 0.0 ? worst / anorm : 0.0;
    if (!(anorm > 0.0) || nlock >= nvec || worst <= tol * anorm) { converged = true; strict = true; break; }
    if (worst <= 1e-11 * anorm && worst > 0.5 * prev_worst && prev_worst > 0.5 * prev2_worst) { converged = true; break; }
    prev2_worst = prev_worst; prev_worst = worst;
    // Chebyshev filter damping [lo, cut]
    const double cut = th[63];
    lo = std::min(lo, cut);
    double cc = 0.5 * (cut + lo), ee = 0.5 * (cut - lo);
    if (!(ee > 0.0)) ee = 1e-3 * anorm;
    const double th_ref = th[nlock];             // largest eigenvalue the filter still has to resolve
    const double xi1 = std::max((th_ref - cc) / ee, 1.0 + 1e-12);
    int deg = (int)floor(acosh(1e8) / std::max(acosh(xi1), 1e-6));
    deg = std::max(2, std::min(deg, 40));
    double sigma1 = ee / std::max(th_ref - cc, ee * (1.0 + 1e-12)), sigma = sigma1;
    if (nlock > 0) {
      double dv[64];
      for (int j = 0; j < 64; j++) dv[j] = j < nlock ? th[j] - cc : 0.0;
      EB_CUDA(cudaMemcpyAsync(dvec_d, dv, sizeof(dv), cudaMemcpyHostToDevice, st));
      EB_CUDA(cudaStreamSynchronize(st));
      EB_CUDA(cudaMemcpyAsync(Lk, V, sizeof(double) * (size_t)nlock * ld, cudaMemcpyDeviceToDevice, st));
    }
    // Y1 = (A V - c V) sigma1/e   (A V is already in AV)
    {
      const double a1 = sigma1 / ee;
      combine_kernel<<<cgrid, 256, 0, st>>>(AV, 1, ld, Ya, ld, n, 0, 0, a1, V, -cc * a1, nullptr, 0.0, ld);
      EB_CHECK_LAUNCH(c);
      // under the deflated operator the locked columns map to (c - c) v = 0
      if (nlock > 0) EB_CUDA(cudaMemsetAsync(Ya, 0, sizeof(double) * (size_t)nlock * ld, st));
    }
    double *y0 = V, *y1 = Ya, *y2 = Yb;          // AV and Yc are free scratch now
    double* spare = AV;
    for (int jd = 2; jd <= deg; jd++) {
      const double sn = 1.0 / (2.0 / sigma1 - sigma);
      const double a1 = 2.0 * sn / ee;
      if ((rc = matvec(y1, y2, a1, y1, -cc * a1, y0, -sigma * sn))) return rc;
      double* t = y0; y0 = y1; y1 = y2; y2 = (jd == 2) ? spare : t;
      if (jd == 2) spare = t;                    // V's buffer joins the rotation after its last use
      sigma = sn;
    }
    // the locked columns go back in unchanged, then orthonormalise the block -> V
    if (nlock > 0) EB_CUDA(cudaMemcpyAsync(y1, Lk, sizeof(double) * (size_t)nlock * ld, cudaMemcpyDeviceToDevice, st));
    double* X = y1; double* tmp = Yc;
    if ((rc = chol_qr(X, tmp))) return rc;
    // re-assign buffer roles: V = X, the other four are scratch
    double* all[5] = {Y[0], Y[1], Y[2], Y[3], Y[4]};
    int k = 0; double* others[4];
    for (int i = 0; i < 5; i++) if (all[i] != X) others[k++] = all[i];
    V = X; AV = others[0]; Ya = others[1]; Yb = others[2]; Yc = others[3];
  }
  int herr = 0;
  EB_CUDA(cudaMemcpyAsync(&herr, err_d, sizeof(int), cudaMemcpyDeviceToHost, st));
  EB_CUDA(cudaStreamSynchronize(st));
  if (herr) { set_error("chfsi_top: Cholesky breakdown in the block orthonormalisation"); return EB_ERR_NUMERIC; }
  if (!converged) {
    // not at the rounding floor after maxouter filters: still usable if the residuals are far below what any consumer
    // resolves (the .evec prints 4-6 decimals, ridoutlier compares z-scores with 6.0); otherwise let the caller fall back
    if (!(last_worst <= 1e-9)) { set_error("chfsi_top: no convergence after %d outer iterations (worst residual / |A| = %.2e)", maxouter, last_worst); return EB_ERR_NUMERIC; }
  }
  // the caller is told when the pairs were accepted above the strict tolerance (stagnation just above the rounding floor, or the
  // relaxed 1e-9 bar after maxouter filters): eb_timings.chfsi_converged = 0 and chfsi_resid = worst |A v - theta v| / |A|
  c->tm.chfsi_converged = strict ? 1 : 0;
  c->tm.chfsi_resid = (float)last_worst;
  normalize_rows_kernel<<<nvec, 256, 0, st>>>(V, ld, n);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaMemcpy2DAsync(vec_d, sizeof(double) * n, V, sizeof(double) * ld, sizeof(double) * n, nvec, cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < nvec; i++) theta_h[i] = th[i];
  if (iters_out) *iters_out = outer + 1;
  if (matvecs_out) *matvecs_out = nmat;
  return 0;
}

}  // namespace eb
