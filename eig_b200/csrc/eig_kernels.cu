// eig_kernels.cu -- symmetric eigensolver on a device-resident FP64 matrix.
//
// Replaces eigvecs() -> packsym -> eigxv_ -> LAPACK dspev_ (eigsubs.c:39-55,145-155; eigx.c:97-117):
// all eigenvalues (descending) + the leading nvec eigenvectors, which is everything smartpca.c consumes
// (lambda[] for .eval / Tracy-Widom, smartpca.c:1312-1431; evecs rows 0..numeigs-1, smartpca.c:1250,1444).
//
//   1. blocked Householder tridiagonalisation (dlatrd/dsytrd scheme, panel width NB): per column one
//      HBM-bound GEMV over the trailing matrix + skinny corrections; per panel one FP64 tensor-core (DMMA)
//      rank-2NB update of the full symmetric trailing block.  Reflector j is kept in row j of the matrix.
//   2. all eigenvalues of the tridiagonal by Sturm-count bisection, one thread per eigenvalue.
//   3. leading eigenvectors of the tridiagonal by inverse iteration with cluster re-orthogonalisation
//      (dstein scheme), then back-transformation through the stored reflectors.
#include <algorithm>
#include <algorithm>
#include <cmath>
#include <vector>
#include <time.h>
#include "common.cuh"

namespace eb {

constexpr int NB = 32;            // panel width
constexpr int KA_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* sh) {   // fixed-order => deterministic
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; i++) r += sh[i];
  return r;
}

struct Refl { double beta, tau, scal; };
// dlarfg from alpha = x[0] and xnorm2 = sum x[1:]^2
__device__ __forceinline__ Refl make_reflector(double alpha, double xnorm2) {
  Refl r;
  if (xnorm2 == 0.0) { r.beta = alpha; r.tau = 0.0; r.scal = 0.0; return r; }
  const double nrm = sqrt(alpha * alpha + xnorm2);
  r.beta = alpha >= 0.0 ? -nrm : nrm;
  r.tau = (r.beta - alpha) / r.beta;
  r.scal = 1.0 / (alpha - r.beta);
  return r;
}

// KA: apply the panel's pending rank-2k update to row j (columns j..n-1), record d[j], and emit per-block
// partial sums of squares of the part that will be annihilated (columns j+2..n-1).
__global__ void __launch_bounds__(KA_THREADS) tri_row_update_kernel(double* __restrict__ A, int64_t lda, int n, int j, int k,
                                                                    const double* __restrict__ Vp, const double* __restrict__ Wp,
                                                                    double* __restrict__ d, double* __restrict__ partA,
                                                                    double* __restrict__ alpha_slot) {
  __shared__ double vj[NB], wj[NB], red[KA_THREADS / 32];
  if ((int)threadIdx.x < k) { vj[threadIdx.x] = Vp[(size_t)threadIdx.x * n + j]; wj[threadIdx.x] = Wp[(size_t)threadIdx.x * n + j]; }
  __syncthreads();
  const int c = j + blockIdx.x * blockDim.x + threadIdx.x;
  double sq = 0.0;
  if (c < n) {
    double a = A[(size_t)j * lda + c];
    for (int m = 0; m < k; m++) a -= vj[m] * Wp[(size_t)m * n + c] + wj[m] * Vp[(size_t)m * n + c];
    A[(size_t)j * lda + c] = a;
    if (c == j) d[j] = a;
    if (c == j + 1) alpha_slot[0] = a;   // row j is overwritten with v later; keep alpha = x[0]
    if (c >= j + 2) sq = a * a;
  }
  const double s = block_sum(sq, red);
  if (threadIdx.x == 0) partA[blockIdx.x] = s;
}

// KB: p_raw = A[j+1:, j+1:] * v  (one warp per row) with v formed on the fly from row j; extra blocks compute the
// correction dots s1[m] = Wp[m].v, s2[m] = Vp[m].v (one warp per dot).
__global__ void __launch_bounds__(256) tri_gemv_kernel(const double* __restrict__ A, int64_t lda, int n, int j, int k,
                                                       const double* __restrict__ Vp, const double* __restrict__ Wp,
                                                       const double* __restrict__ partA, int npartA, const double* __restrict__ alpha_slot,
                                                       double* __restrict__ praw, double* __restrict__ s12, double* __restrict__ e, double* __restrict__ tau,
                                                       int gemv_blocks) {
  __shared__ Refl rf;
  if (threadIdx.x == 0) {
    double x2 = 0.0;
    for (int i = 0; i < npartA; i++) x2 += partA[i];
    rf = make_reflector(alpha_slot[0], x2);
    if (blockIdx.x == 0) { e[j] = rf.beta; tau[j] = rf.tau; }
  }
  __syncthreads();
  const double scal = rf.scal;
  const double* x = A + (size_t)j * lda;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = j + 1;
  if ((int)blockIdx.x < gemv_blocks) {
    const int row = c0 + blockIdx.x * 8 + warp;
    if (row >= n) return;
    const double* ar = A + (size_t)row * lda;
    double acc = 0.0;
    for (int c = c0 + lane; c < n; c += 32) {
      const double v = (c == c0) ? 1.0 : x[c] * scal;
      acc += ar[c] * v;
    }
    acc = warp_sum(acc);
    if (lane == 0) praw[row] = acc;
  } else {
    const int id = (blockIdx.x - gemv_blocks) * 8 + warp;     // 0..2k-1
    if (id >= 2 * k) return;
    const double* src = (id < k) ? Wp + (size_t)id * n : Vp + (size_t)(id - k) * n;
    double acc = 0.0;
    for (int c = c0 + lane; c < n; c += 32) {
      const double v = (c == c0) ? 1.0 : x[c] * scal;
      acc += src[c] * v;
    }
    acc = warp_sum(acc);
    if (lane == 0) s12[id] = acc;
  }
}

// KC: p = tau * (p_raw - V s1 - W s2); store v (panel row k and row j of A); per-block partial of p.v
__global__ void __launch_bounds__(KA_THREADS) tri_correct_kernel(double* __restrict__ A, int64_t lda, int n, int j, int k,
                                                                 double* __restrict__ Vp, const double* __restrict__ Wp,
                                                                 const double* __restrict__ partA, int npartA,
                                                                 const double* __restrict__ alpha_slot,
                                                                 const double* __restrict__ praw, const double* __restrict__ s12,
                                                                 double* __restrict__ p, double* __restrict__ partC) {
  __shared__ double s1[NB], s2[NB], red[KA_THREADS / 32];
  __shared__ Refl rf;
  if (threadIdx.x == 0) {
    double x2 = 0.0;
    for (int i = 0; i < npartA; i++) x2 += partA[i];
    rf = make_reflector(alpha_slot[0], x2);
  }
  if ((int)threadIdx.x < k) { s1[threadIdx.x] = s12[threadIdx.x]; s2[threadIdx.x] = s12[k + threadIdx.x]; }
  __syncthreads();
  const int c = j + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  double dot = 0.0;
  if (c < n) {
    const double v = (c == j + 1) ? 1.0 : A[(size_t)j * lda + c] * rf.scal;
    double acc = praw[c];
    for (int m = 0; m < k; m++) acc -= Vp[(size_t)m * n + c] * s1[m] + Wp[(size_t)m * n + c] * s2[m];
    acc *= rf.tau;
    p[c] = acc;
    Vp[(size_t)k * n + c] = v;
    A[(size_t)j * lda + c] = v;
    dot = acc * v;
  }
  __syncthreads();   // all reads of row j happen before rf is reused; (A row j writes are per-thread own element)
  const double s = block_sum(dot, red);
  if (threadIdx.x == 0) partC[blockIdx.x] = s;
}

// KD: w = p - (tau/2)(p.v) v  -> panel row k
__global__ void __launch_bounds__(KA_THREADS) tri_w_kernel(int n, int j, int k, const double* __restrict__ Vp, double* __restrict__ Wp,
                                                           const double* __restrict__ p, const double* __restrict__ partC, int npartC,
                                                           const double* __restrict__ tau) {
  __shared__ double al;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < npartC; i++) s += partC[i];
    al = -0.5 * tau[j] * s;
  }
  __syncthreads();
  const int c = j + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) Wp[(size_t)k * n + c] = p[c] + al * Vp[(size_t)k * n + c];
}

// Trailing update  A[j1:, j1:] -= sum_m V[m][i] W[m][c] + W[m][i] V[m][c]   (full symmetric block, DMMA).
// CTA tile 128x128, 8 warps (64x32 each); K = 2*kp staged in shared memory in chunks of 16.
__device__ __forceinline__ void dmma884e(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int TU_K = 16, TU_LD = 132;   // == 4 mod 16 doubles: conflict-free DMMA fragment reads
__global__ void __launch_bounds__(256, 1) tri_trailing_kernel(double* __restrict__ A, int64_t lda, int n, int j1, int kp,
                                                              const double* __restrict__ Vp, const double* __restrict__ Wp) {
  __shared__ double Ps[TU_K][TU_LD], Qs[TU_K][TU_LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3, g = lane >> 2, q = lane & 3;
  const int i0 = j1 + blockIdx.y * 128, c0 = j1 + blockIdx.x * 128;
  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; t++)
#pragma unroll
    for (int u = 0; u < 4; u++) acc[t][u][0] = acc[t][u][1] = 0.0;
  const int K = 2 * kp;
  for (int k0 = 0; k0 < K; k0 += TU_K) {
    __syncthreads();
    // P[kk][i] = (kk<kp ? V : W)[kk][i0+i] ; Q[kk][c] = (kk<kp ? W : V)[kk][c0+c]
    for (int idx = threadIdx.x; idx < TU_K * 128; idx += 256) {
      const int kk = idx >> 7, ii = idx & 127, kg = k0 + kk;
      double pv = 0.0, qv = 0.0;
      if (kg < K) {
        const bool first = kg < kp;
        const int m = first ? kg : kg - kp;
        const double* Pm = (first ? Vp : Wp) + (size_t)m * n;
        const double* Qm = (first ? Wp : Vp) + (size_t)m * n;
        if (i0 + ii < n) pv = Pm[i0 + ii];
        if (c0 + ii < n) qv = Qm[c0 + ii];
      }
      Ps[kk][ii] = pv; Qs[kk][ii] = qv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TU_K; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int t = 0; t < 8; t++) a[t] = Ps[kk + q][wm * 64 + t * 8 + g];
#pragma unroll
      for (int u = 0; u < 4; u++) b[u] = Qs[kk + q][wn * 32 + u * 8 + g];
#pragma unroll
      for (int t = 0; t < 8; t++)
#pragma unroll
        for (int u = 0; u < 4; u++) dmma884e(acc[t][u][0], acc[t][u][1], a[t], b[u]);
    }
  }
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int row = i0 + wm * 64 + t * 8 + g;
    if (row >= n) continue;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int col = c0 + wn * 32 + u * 8 + q * 2;
      double* ptr = A + (size_t)row * lda + col;
      if (col < n) ptr[0] -= acc[t][u][0];
      if (col + 1 < n) ptr[1] -= acc[t][u][1];
    }
  }
}

// last 2x2 block -> d[n-2], e[n-2], d[n-1]; also applies the eigenvalue scale to d and e
__global__ void tri_tail_scale_kernel(const double* __restrict__ A, int64_t lda, int n, double* __restrict__ d, double* __restrict__ e,
                                      double scale) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (n >= 2) { d[n - 2] = A[(size_t)(n - 2) * lda + n - 2]; e[n - 2] = A[(size_t)(n - 2) * lda + n - 1]; }
    d[n - 1] = A[(size_t)(n - 1) * lda + n - 1];
    e[n - 1] = 0.0;
  }
}
__global__ void scale_de_kernel(int n, double* __restrict__ d, double* __restrict__ e, double scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { d[i] *= scale; e[i] *= scale; }
}

// ------------------------------------------------------------------------------------------ bisection
// Gershgorin bounds + norm of T (single block)
__global__ void __launch_bounds__(1024) tri_bounds_kernel(int n, const double* __restrict__ d, const double* __restrict__ e,
                                                          double* __restrict__ bounds) {
  __shared__ double smin[32], smax[32];
  double lo = 1e300, hi = -1e300;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < n - 1 ? fabs(e[i]) : 0.0);
    lo = fmin(lo, d[i] - r); hi = fmax(hi, d[i] + r);
  }
  for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { lo = fmin(lo, smin[i]); hi = fmax(hi, smax[i]); }
    const double tn = fmax(fabs(lo), fabs(hi));
    bounds[0] = lo - 2.0 * tn * 2.3e-16 * n - 1e-300;
    bounds[1] = hi + 2.0 * tn * 2.3e-16 * n + 1e-300;
    bounds[2] = tn;
  }
}

// Number of eigenvalues of the tridiagonal (d, e) below x.  LAPACK's dstebz runs the ratio recurrence q_i = d_i - x - e_{i-1}^2 / q_{i-1}
// and counts negative q_i; a chain of n dependent FP64 DIVISIONS (~280 cycles each on this part, measured: 0.38 s for n = 50,000).
// The same count comes out of the three-term recurrence of the leading principal minors p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}
// (q_i = p_i / p_{i-1}: q_i < 0 iff p_i and p_{i-1} differ in sign) -- two dependent DFMAs per step -- with the pair rescaled by a power
// of two every 8 steps (exact, so no sign changes) and dstebz's pivot floor kept: |q_i| < pivmin counts as a negative pivot.
__device__ __forceinline__ int sturm_count(int n, const double* __restrict__ d, const double* __restrict__ e2, double x, double pivmin) {
  double pm = 1.0, p = d[0] - x;
  if (fabs(p) < pivmin) p = -pivmin;
  int cnt = p < 0.0;
  for (int i = 1; i < n; i++) {
    double pn = fma(d[i] - x, p, -e2[i - 1] * pm);
    if (fabs(pn) < pivmin * fabs(p)) pn = -pivmin * p;             // q_i = -pivmin
    cnt += (pn < 0.0) != (p < 0.0);
    pm = p; p = pn;
    if ((i & 7) == 0) {
      // bring the larger of the pair to [1, 2): multiply both by 2^-(exponent)
      const int ea = (__double2hiint(p) >> 20) & 0x7ff, eb = (__double2hiint(pm) >> 20) & 0x7ff;
      const int em = ea > eb ? ea : eb;
      if (em > 0 && em < 0x7fe) {
        const double sc = __hiloint2double((2046 - em) << 20, 0);
        p *= sc; pm *= sc;
      }
    }
  }
  return cnt;
}

__global__ void __launch_bounds__(128) e2_kernel(int n, const double* __restrict__ e, double* __restrict__ e2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) e2[i] = e[i] * e[i];
}

// SEC lanes find the k-th smallest eigenvalue together by (SEC + 1)-section: every round lane i counts the eigenvalues below
// lo + (i + 1)(hi - lo) / (SEC + 1) and the bracket shrinks to the sub-interval that holds eigenvalue k -- log2(SEC + 1) bits per Sturm
// sweep instead of one.  A sweep is a chain of n dependent divisions, so the kernel is latency bound with one thread per eigenvalue
// (50,000 threads on 148 SMs); the extra lanes are free until the SMs fill up.  Output descending: lam[n-1-k].  Indices [k0, k1).
template <int SEC>
__global__ void __launch_bounds__(128) tri_bisect_kernel(int n, const double* __restrict__ d, const double* __restrict__ e2,
                                                         const double* __restrict__ bounds, double* __restrict__ lam, int k0, int k1) {
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / SEC, sub = threadIdx.x % SEC;
  const int k = k0 + gid;
  const bool live = k < k1;                         // whole groups are live or not; dead groups still take part in the shuffles
  double lo = bounds[0], hi = bounds[1];
  const double tn = bounds[2];
  const double pivmin = fmax(2.3e-308 * fmax(1.0, tn * tn), 1e-300);
  const double atol = 2.0 * 2.220446049250313e-16 * tn * 0.25 + 2.0 * pivmin;
  const unsigned gmask = SEC == 32 ? 0xffffffffu : (((1u << SEC) - 1u) << ((threadIdx.x & 31) / SEC * SEC));
  for (int it = 0; it < 200; it++) {
    const double w = hi - lo;
    const double mid = 0.5 * (lo + hi);
    if (w <= atol + 2.220446049250313e-16 * fmax(fabs(lo), fabs(hi)) || mid <= lo || mid >= hi) break;     // uniform over the group
    const double x = SEC == 1 ? mid : lo + w * ((double)(sub + 1) / (double)(SEC + 1));
    const bool above = live ? sturm_count(n, d, e2, x, pivmin) > k : true;       // eigenvalue k lies below x
    if (SEC == 1) { if (above) hi = x; else lo = x; continue; }
    // lanes are ordered by x: the first lane whose point lies above eigenvalue k closes the bracket from the right
    const unsigned m = (__ballot_sync(gmask, above) & gmask) >> ((threadIdx.x & 31) / SEC * SEC);
    const int first = m ? __ffs(m) - 1 : SEC;        // points 0 .. first-1 are at or below eigenvalue k
    const double nlo = first > 0 ? lo + w * ((double)first / (double)(SEC + 1)) : lo;
    const double nhi = first < SEC ? lo + w * ((double)(first + 1) / (double)(SEC + 1)) : hi;
    lo = nlo; hi = nhi;
  }
  if (live && sub == 0) lam[n - 1 - k] = 0.5 * (lo + hi);
}

// ------------------------------------------------------------------------------------------ inverse iteration
// Single block. For each of the nvec leading eigenvalues: solve (T - lam I) z = b with partial-pivoting LU
// (thread 0, sequential), re-orthogonalise against earlier vectors of the same cluster, normalise; repeat.
// work: 5*n doubles + n ints ; Z: [nvec][n]
__global__ void __launch_bounds__(1024) tri_invit_kernel(int n, int nvec, const double* __restrict__ d, const double* __restrict__ e,
                                                         const double* __restrict__ lam, const double* __restrict__ bounds,
                                                         double* __restrict__ work, int* __restrict__ ipiv, double* __restrict__ Z) {
  __shared__ double red[32];
  double* dl = work;            // sub-diagonal multipliers
  double* dd = work + n;        // U diagonal
  double* du = work + 2 * n;    // U first super-diagonal
  double* du2 = work + 3 * n;   // U second super-diagonal
  double* b = work + 4 * n;
  const double tn = bounds[2];
  const double eps = 2.220446049250313e-16;
  const double ortol = 1e-3 * tn;
  int cluster_start = 0;
  double prev_shift = 0.0;
  for (int v = 0; v < nvec; v++) {
    double shift = lam[v];
    if (v > 0) {
      if (fabs(lam[v - 1] - lam[v]) > ortol) cluster_start = v;
      // keep shifts of near-coincident eigenvalues separated (dstein's perturbation)
      const double sep = 10.0 * eps * fabs(shift) + 10.0 * eps * tn * 1e-3;
      if (prev_shift - shift < sep) shift = prev_shift - sep;
    }
    prev_shift = shift;
    double* z = Z + (size_t)v * n;
    // deterministic pseudo-random start
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      uint32_t hsh = (uint32_t)(i * 2654435761u) ^ (uint32_t)((v + 1) * 40503u);
      hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
      b[i] = 0.5 + (double)(hsh & 0xFFFF) / 65536.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // LU of tridiagonal (T - shift I) with partial pivoting (dgttrf scheme)
      for (int i = 0; i < n; i++) { dd[i] = d[i] - shift; if (i < n - 1) { dl[i] = e[i]; du[i] = e[i]; } du2[i] = 0.0; }
      const double piv0 = eps * tn;
      for (int i = 0; i < n - 1; i++) {
        if (fabs(dd[i]) >= fabs(dl[i])) {
          if (fabs(dd[i]) < piv0) dd[i] = piv0;
          const double f = dl[i] / dd[i];
          dl[i] = f; dd[i + 1] -= f * du[i]; ipiv[i] = 0;
        } else {
          const double f = dd[i] / dl[i];
          dd[i] = dl[i]; dl[i] = f;
          const double t = du[i];
          du[i] = dd[i + 1]; dd[i + 1] = t - f * du[i];
          if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
          ipiv[i] = 1;
        }
      }
      if (fabs(dd[n - 1]) < piv0) dd[n - 1] = piv0;
    }
    __syncthreads();
    for (int iter = 0; iter < 4; iter++) {
      if (threadIdx.x == 0) {
        // forward: L y = P b
        for (int i = 0; i < n - 1; i++) {
          if (ipiv[i]) { const double t = b[i]; b[i] = b[i + 1]; b[i + 1] = t - dl[i] * b[i]; }
          else b[i + 1] -= dl[i] * b[i];
        }
        // backward: U x = y
        b[n - 1] /= dd[n - 1];
        if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / dd[n - 2];
        for (int i = n - 3; i >= 0; i--) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / dd[i];
      }
      __syncthreads();
      // re-orthogonalise within the cluster (modified Gram-Schmidt)
      for (int u = cluster_start; u < v; u++) {
        const double* zu = Z + (size_t)u * n;
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s += zu[i] * b[i];
        s = block_sum(s, red);
        for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] -= s * zu[i];
        __syncthreads();
      }
      double s = 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) s += b[i] * b[i];
      s = block_sum(s, red);
      const double inv = 1.0 / sqrt(s);
      for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] *= inv;
      __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] = b[i];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ back-transformation
// One block per eigenvector: z <- H_0 H_1 ... H_{n-3} z, reflector j stored in row j of A (columns j+1..n-1, v[j+1]=1).
__global__ void __launch_bounds__(1024) tri_backtransform_kernel(const double* __restrict__ A, int64_t lda, int n,
                                                                 const double* __restrict__ tau, double* __restrict__ Z) {
  __shared__ double red[32];
  double* z = Z + (size_t)blockIdx.x * n;
  for (int j = n - 3; j >= 0; j--) {
    const double tj = tau[j];
    if (tj == 0.0) continue;
    const double* v = A + (size_t)j * lda;
    double s = 0.0;
    for (int c = j + 1 + threadIdx.x; c < n; c += blockDim.x) s += v[c] * z[c];
    s = block_sum(s, red) * tj;
    for (int c = j + 1 + threadIdx.x; c < n; c += blockDim.x) z[c] -= s * v[c];
    __syncthreads();
  }
  // normalise (unit 2-norm, as LAPACK returns)
  double s = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) s += z[c] * z[c];
  s = block_sum(s, red);
  const double inv = 1.0 / sqrt(s);
  for (int c = threadIdx.x; c < n; c += blockDim.x) z[c] *= inv;
}

// ------------------------------------------------------------------------------------------ all eigenvectors (nvec > 64)
// The single-block inverse iteration above is sequential in the vectors and the per-vector back-transformation re-reads the
// whole reflector matrix for every vector; both are fine for the 10-40 leading vectors smartpca prints, not for the full
// basis that eigvecs() promises (eigsubs.c:39-55) and shrinkmode consumes (smartpca.c:4292, 4340-4347).  Full-basis path:
//   tri_invit_batch_kernel : one THREAD per eigenvector; dgttrf-style pivoted LU of T - shift I and the dgtts2-style solves
//                            run sequentially in the thread over interleaved work arrays ([row][vector] => coalesced)
//   tri_invit_cluster_kernel: groups of numerically coincident eigenvalues only: sequential inverse iteration with
//                            re-orthogonalisation inside every iteration (one block per group)
//   blocked back-transformation: 64 reflectors at a time as I - V T V^T (dlarft), applied to all vectors with two FP64
//                            tensor-core GEMMs per panel (launch_gemm), restricted to the columns the reflectors touch.
__global__ void __launch_bounds__(128) tri_invit_batch_kernel(int n, int v0, int nb, const double* __restrict__ d, const double* __restrict__ e,
                                                              const double* __restrict__ shifts, double tn, double* __restrict__ dl,
                                                              double* __restrict__ dd, double* __restrict__ du, double* __restrict__ du2,
                                                              double* __restrict__ b, uint8_t* __restrict__ ipiv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const int v = v0 + t;
#define IX(i) ((size_t)(i) * nb + t)
  const double shift = shifts[v];
  const double piv0 = 2.220446049250313e-16 * tn;
  {
    double dd_i = d[0] - shift, du_i = n > 1 ? e[0] : 0.0;
    for (int i = 0; i < n - 1; i++) {
      double dl_i = e[i];
      double dd_n = d[i + 1] - shift, du_n = i + 1 < n - 1 ? e[i + 1] : 0.0, du2_i = 0.0;
      uint8_t pv = 0;
      if (fabs(dd_i) >= fabs(dl_i)) {
        if (fabs(dd_i) < piv0) dd_i = piv0;
        const double f = dl_i / dd_i;
        dl_i = f; dd_n -= f * du_i;
      } else {
        const double f = dd_i / dl_i;
        dd_i = dl_i; dl_i = f;
        const double tt = du_i;
        du_i = dd_n; dd_n = tt - f * du_i;
        if (i < n - 2) { du2_i = du_n; du_n = -f * du_n; }
        pv = 1;
      }
      dl[IX(i)] = dl_i; dd[IX(i)] = dd_i; du[IX(i)] = du_i; du2[IX(i)] = du2_i; ipiv[IX(i)] = pv;
      dd_i = dd_n; du_i = du_n;
    }
    if (fabs(dd_i) < piv0) dd_i = piv0;
    dd[IX(n - 1)] = dd_i;
  }
  for (int i = 0; i < n; i++) {            // deterministic pseudo-random start (same generator as tri_invit_kernel)
    uint32_t hsh = (uint32_t)(i * 2654435761u) ^ (uint32_t)((v + 1) * 40503u);
    hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
    b[IX(i)] = 0.5 + (double)(hsh & 0xFFFF) / 65536.0;
  }
  double sc = 1.0;
  for (int iter = 0; iter < 4; iter++) {
    double cur = b[IX(0)] * sc;              // forward: L y = P b
    for (int i = 0; i < n - 1; i++) {
      double nxt = b[IX(i + 1)] * sc;
      const double m = dl[IX(i)];
      if (ipiv[IX(i)]) { const double tt = cur; cur = nxt; nxt = tt - m * cur; }
      else nxt -= m * cur;
      b[IX(i)] = cur;
      cur = nxt;
    }
    double x1 = cur / dd[IX(n - 1)];         // backward: U x = y
    b[IX(n - 1)] = x1;
    double ss = x1 * x1, x2 = 0.0;
    if (n > 1) {
      const double x0 = (b[IX(n - 2)] - du[IX(n - 2)] * x1) / dd[IX(n - 2)];
      b[IX(n - 2)] = x0; ss += x0 * x0; x2 = x1; x1 = x0;
    }
    for (int i = n - 3; i >= 0; i--) {
      const double x0 = (b[IX(i)] - du[IX(i)] * x1 - du2[IX(i)] * x2) / dd[IX(i)];
      b[IX(i)] = x0; ss += x0 * x0; x2 = x1; x1 = x0;
    }
    sc = 1.0 / sqrt(ss);
  }
  for (int i = 0; i < n; i++) b[IX(i)] *= sc;
#undef IX
}

// Zt[v0 + t][i] = Bi[i][t]   (interleaved work array -> one contiguous row per eigenvector)
__global__ void __launch_bounds__(256) invit_transpose_kernel(const double* __restrict__ Bi, int n, int nb, int v0, double* __restrict__ Zt, int64_t ldz) {
  __shared__ double tile[32][33];
  const int t0 = blockIdx.x * 32, i0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = (i0 + r < n && t0 + tx < nb) ? Bi[(size_t)(i0 + r) * nb + t0 + tx] : 0.0;
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (t0 + r < nb && i0 + tx < n) Zt[(size_t)(v0 + t0 + r) * ldz + i0 + tx] = tile[tx][r];
}

// One block per group of numerically coincident eigenvalues (rows [start, start + len) of Zt): inverse iteration with
// re-orthogonalisation against the earlier members INSIDE every iteration (dstein), as tri_invit_kernel does.  Independent
// inverse iterations followed by Gram-Schmidt are not enough here: within a coincident group every start vector converges
// towards the same few directions and the orthogonalisation then cancels catastrophically (seen as 1e-5 orthogonality with
// a 230-fold zero eigenvalue).  work: 5n doubles per block, ipiv: n ints per block.
__global__ void __launch_bounds__(1024) tri_invit_cluster_kernel(int n, const double* __restrict__ d, const double* __restrict__ e,
                                                                 const double* __restrict__ shifts, double tn, const int* __restrict__ starts,
                                                                 const int* __restrict__ lens, double* __restrict__ work_all,
                                                                 int* __restrict__ ipiv_all, double* __restrict__ Zt, int64_t ldz) {
  __shared__ double red[32];
  double* work = work_all + (size_t)blockIdx.x * 5 * n;
  int* ipiv = ipiv_all + (size_t)blockIdx.x * n;
  double* dl = work; double* dd = work + n; double* du = work + 2 * n; double* du2 = work + 3 * n; double* b = work + 4 * n;
  const int s0 = starts[blockIdx.x], len = lens[blockIdx.x];
  const double eps = 2.220446049250313e-16;
  for (int v = s0; v < s0 + len; v++) {
    const double shift = shifts[v];
    double* z = Zt + (size_t)v * ldz;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      uint32_t hsh = (uint32_t)(i * 2654435761u) ^ (uint32_t)((v + 1) * 40503u);
      hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
      b[i] = 0.5 + (double)(hsh & 0xFFFF) / 65536.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 0; i < n; i++) { dd[i] = d[i] - shift; if (i < n - 1) { dl[i] = e[i]; du[i] = e[i]; } du2[i] = 0.0; }
      const double piv0 = eps * tn;
      for (int i = 0; i < n - 1; i++) {
        if (fabs(dd[i]) >= fabs(dl[i])) {
          if (fabs(dd[i]) < piv0) dd[i] = piv0;
          const double f = dl[i] / dd[i];
          dl[i] = f; dd[i + 1] -= f * du[i]; ipiv[i] = 0;
        } else {
          const double f = dd[i] / dl[i];
          dd[i] = dl[i]; dl[i] = f;
          const double t = du[i];
          du[i] = dd[i + 1]; dd[i + 1] = t - f * du[i];
          if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
          ipiv[i] = 1;
        }
      }
      if (fabs(dd[n - 1]) < piv0) dd[n - 1] = piv0;
    }
    __syncthreads();
    for (int iter = 0; iter < 4; iter++) {
      if (threadIdx.x == 0) {
        for (int i = 0; i < n - 1; i++) {
          if (ipiv[i]) { const double t = b[i]; b[i] = b[i + 1]; b[i + 1] = t - dl[i] * b[i]; }
          else b[i + 1] -= dl[i] * b[i];
        }
        b[n - 1] /= dd[n - 1];
        if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / dd[n - 2];
        for (int i = n - 3; i >= 0; i--) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / dd[i];
      }
      __syncthreads();
      for (int pass = 0; pass < 2; pass++)
        for (int u = s0; u < v; u++) {
          const double* zu = Zt + (size_t)u * ldz;
          double sacc = 0.0;
          for (int i = threadIdx.x; i < n; i += blockDim.x) sacc += zu[i] * b[i];
          sacc = block_sum(sacc, red);
          for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] -= sacc * zu[i];
          __syncthreads();
        }
      double q = 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) q += b[i] * b[i];
      q = block_sum(q, red);
      const double inv = 1.0 / sqrt(q);
      for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] *= inv;
      __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] = b[i];
    __syncthreads();
  }
}

// Vp[k][c] = reflector j0 + k as a dense row (zero up to column j0 + k, then row j0 + k of A; rows >= kp zero)
__global__ void __launch_bounds__(256) bt_extract_kernel(const double* __restrict__ A, int64_t lda, int n, int j0, int kp, double* __restrict__ Vp,
                                                         int64_t ldv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (c >= ldv) return;
  const int j = j0 + k;
  Vp[(size_t)k * ldv + c] = (k < kp && c > j && c < n) ? A[(size_t)j * lda + c] : 0.0;
}

// partial Gram of the 64 rows of Vp over a column chunk: Gp[chunk][a][b]
__global__ void __launch_bounds__(256) bt_gram_kernel(const double* __restrict__ Vp, int64_t ldv, int c0, int n, int chunk, double* __restrict__ Gp) {
  __shared__ double Xs[64][33];
  const int tid = threadIdx.x, ta = tid >> 4, tc = tid & 15;
  const int lo = c0 + blockIdx.x * chunk, hi = min(n, lo + chunk);
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
  for (int s0 = lo; s0 < hi; s0 += 32) {
    __syncthreads();
    for (int idx = tid; idx < 64 * 32; idx += 256) {
      const int r = idx >> 5, ii = idx & 31, i = s0 + ii;
      Xs[r][ii] = i < hi ? Vp[(size_t)r * ldv + i] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int ii = 0; ii < 32; ii++) {
      double xa[4], xb[4];
#pragma unroll
      for (int a = 0; a < 4; a++) xa[a] = Xs[ta * 4 + a][ii];
#pragma unroll
      for (int b = 0; b < 4; b++) xb[b] = Xs[tc * 4 + b][ii];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] += xa[a] * xb[b];
    }
  }
  double* out = Gp + (size_t)blockIdx.x * 4096;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) out[(ta * 4 + a) * 64 + tc * 4 + b] = acc[a][b];
}

// T of the block reflector H_j0 ... H_{j0+63} = I - V T V^T (dlarft, forward columnwise): T[0:k,k] = -tau_k T[0:k,0:k] (V^T v_k)
__global__ void __launch_bounds__(256) bt_tfactor_kernel(const double* __restrict__ Gp, int nchunk, const double* __restrict__ tau, int kp,
                                                         double* __restrict__ Gs, double* __restrict__ T) {
  __shared__ double Ts[64 * 64];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 4096; idx += 256) {
    double g = 0.0;
    for (int ch = 0; ch < nchunk; ch++) g += Gp[(size_t)ch * 4096 + idx];
    Gs[idx] = g; Ts[idx] = 0.0;
  }
  __threadfence_block();
  __syncthreads();
  for (int k = 0; k < 64; k++) {
    const double tk = k < kp ? tau[k] : 0.0;
    if (tid < k && tk != 0.0) {
      double sacc = 0.0;
      for (int m = tid; m < k; m++) sacc += Ts[tid * 64 + m] * Gs[m * 64 + k];
      Ts[tid * 64 + k] = -tk * sacc;
    }
    if (tid == k) Ts[k * 64 + k] = tk;
    __syncthreads();
  }
  for (int idx = tid; idx < 4096; idx += 256) T[idx] = Ts[idx];
}

// W2[v][a] = sum_b W[v][b] T[a][b]
__global__ void __launch_bounds__(256) bt_wt_kernel(const double* __restrict__ W, const double* __restrict__ T, int nvec, double* __restrict__ W2) {
  __shared__ double Ts[64][65];
  for (int idx = threadIdx.x; idx < 4096; idx += 256) Ts[idx >> 6][idx & 63] = T[idx];
  __syncthreads();
  const int a = threadIdx.x & 63;
  for (int v = blockIdx.x * 4 + (threadIdx.x >> 6); v < nvec; v += gridDim.x * 4) {
    const double* w = W + (size_t)v * 64;
    double sacc = 0.0;
#pragma unroll 8
    for (int b = 0; b < 64; b++) sacc += w[b] * Ts[a][b];
    W2[(size_t)v * 64 + a] = sacc;
  }
}

__global__ void __launch_bounds__(256) normalize_rows_full_kernel(double* __restrict__ Zt, int64_t ldz, int n) {
  __shared__ double red[32];
  double* z = Zt + (size_t)blockIdx.x * ldz;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) q += z[i] * z[i];
  q = block_sum(q, red);
  const double inv = 1.0 / sqrt(q);
  for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] *= inv;
}

// eigenvectors 0..nvec-1 of the tridiagonal (d, e; lam descending on the device and the host) and their back-transformation
// through the reflectors kept in the rows of A.  Result: c->zvec_d, [nvec][ldz] with ldz = n rounded up to even.
// eigenvectors 0..nvec-1 of the tridiagonal (d, e; lam descending, host copy given) -> c->zvec_d [nvec][ldz]: batched inverse iteration
// (one thread per vector), dstein-style sequential treatment of numerically coincident groups
static int tri_vectors(eb_ctx* c, int n, int nvec, const double* d, const double* e, const double* lam_h, double tn, int64_t ldz) {
  cudaStream_t st = c->stream;
  int rc;
  if ((rc = c->zvec_d.ensure((size_t)nvec * ldz))) return rc;
  EB_CUDA(cudaMemsetAsync(c->zvec_d.p, 0, sizeof(double) * (size_t)nvec * ldz, st));
  // shifts: dstein's separation of near-coincident eigenvalues; groups that need re-orthogonalisation
  const double eps = 2.220446049250313e-16;
  std::vector<double> shifts(nvec);
  std::vector<int> starts, lens;
  {
    const double tight = 1e-6 * tn;
    double prev = 0.0;
    int cs = 0;
    for (int v = 0; v < nvec; v++) {
      double sh = lam_h[v];
      if (v > 0) {
        const double sep = 10.0 * eps * fabs(sh) + 10.0 * eps * tn * 1e-3;
        if (prev - sh < sep) sh = prev - sep;
        if (fabs(lam_h[v - 1] - lam_h[v]) > tight) {
          if (v - cs > 1) { starts.push_back(cs); lens.push_back(v - cs); }
          cs = v;
        }
      }
      shifts[v] = sh; prev = sh;
    }
    if (nvec - cs > 1) { starts.push_back(cs); lens.push_back(nvec - cs); }
  }
  DevBuf<double> sh_d, wk;
  DevBuf<uint8_t> piv;
  DevBuf<int> cl;
  const int batch = (int)std::min<int64_t>(nvec, std::max<int64_t>(256, (int64_t)(3ll << 30) / (5 * 8 * (int64_t)n)));   // <= 3 GiB of work arrays
  if ((rc = sh_d.ensure(nvec)) || (rc = wk.ensure((size_t)5 * n * batch)) || (rc = piv.ensure((size_t)n * batch))) return rc;
  EB_CUDA(cudaMemcpyAsync(sh_d.p, shifts.data(), sizeof(double) * nvec, cudaMemcpyHostToDevice, st));
  const size_t nbsz = (size_t)n * batch;
  for (int v0 = 0; v0 < nvec; v0 += batch) {
    const int nb = std::min(batch, nvec - v0);
    tri_invit_batch_kernel<<<(nb + 127) / 128, 128, 0, st>>>(n, v0, nb, d, e, sh_d.p, tn, wk.p, wk.p + nbsz, wk.p + 2 * nbsz, wk.p + 3 * nbsz,
                                                             wk.p + 4 * nbsz, piv.p);
    EB_CHECK_LAUNCH(c);
    invit_transpose_kernel<<<dim3((nb + 31) / 32, (n + 31) / 32), 256, 0, st>>>(wk.p + 4 * nbsz, n, nb, v0, c->zvec_d.p, ldz);
    EB_CHECK_LAUNCH(c);
  }
  EB_CUDA(cudaStreamSynchronize(st));      // shifts (host vector) consumed
  if (!starts.empty()) {
    const int ncl = (int)starts.size();
    if ((rc = cl.ensure((size_t)2 * ncl))) return rc;
    EB_CUDA(cudaMemcpyAsync(cl.p, starts.data(), sizeof(int) * ncl, cudaMemcpyHostToDevice, st));
    EB_CUDA(cudaMemcpyAsync(cl.p + ncl, lens.data(), sizeof(int) * ncl, cudaMemcpyHostToDevice, st));
    DevBuf<double> cw;
    DevBuf<int> cp;
    if ((rc = cw.ensure((size_t)ncl * 5 * n)) || (rc = cp.ensure((size_t)ncl * n))) return rc;
    tri_invit_cluster_kernel<<<ncl, 1024, 0, st>>>(n, d, e, sh_d.p, tn, cl.p, cl.p + ncl, cw.p, cp.p, c->zvec_d.p, ldz);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

// modified Gram-Schmidt over the rows of Z (twice), unit 2-norm: the leading vectors of a gap-free spectrum come out of independent
// inverse iterations orthogonal to ~eps / gap only
__global__ void __launch_bounds__(1024) mgs_rows_kernel(double* __restrict__ Z, int64_t ldz, int n, int nvec) {
  __shared__ double red[32];
  for (int v = 0; v < nvec; v++) {
    double* z = Z + (size_t)v * ldz;
    for (int pass = 0; pass < 2; pass++)
      for (int u = 0; u < v; u++) {
        const double* zu = Z + (size_t)u * ldz;
        double sacc = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) sacc += zu[i] * z[i];
        sacc = block_sum(sacc, red);
        for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] -= sacc * zu[i];
        __syncthreads();
      }
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) q += z[i] * z[i];
    q = block_sum(q, red);
    const double inv = 1.0 / sqrt(q);
    for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] *= inv;
    __syncthreads();
  }
}

static int full_basis(eb_ctx* c, double* A, int64_t lda, int n, int nvec, const double* d, const double* e, const double* tau,
                      const std::vector<double>& lam_h, double tn, int64_t* ldz_out) {
  cudaStream_t st = c->stream;
  int rc;
  const int64_t ldz = ((int64_t)n + 1) & ~1ll;
  *ldz_out = ldz;
  if ((rc = tri_vectors(c, n, nvec, d, e, lam_h.data(), tn, ldz))) return rc;
  DevBuf<double> Vp, Gp, T, W, W2;
  // back-transformation: z <- H_0 H_1 ... H_{n-3} z, reflectors grouped 64 at a time, last group first
  const int nrefl = std::max(0, n - 2);
  if (nrefl > 0) {
    const int chunk = 512;
    const int nchunk_max = (n + chunk - 1) / chunk + 1;
    if ((rc = Vp.ensure((size_t)64 * ldz)) || (rc = Gp.ensure((size_t)(nchunk_max + 1) * 4096)) || (rc = T.ensure(4096)) ||
        (rc = W.ensure((size_t)nvec * 64)) || (rc = W2.ensure((size_t)nvec * 64)))
      return rc;
    const int npanel = (nrefl + 63) / 64;
    for (int pnl = npanel - 1; pnl >= 0; pnl--) {
      const int j0 = pnl * 64, kp = std::min(64, nrefl - j0);
      const int c0 = j0 & ~1;                                   // reflectors of this group are zero left of column j0 + 1
      bt_extract_kernel<<<dim3((unsigned)((ldz + 255) / 256), 64), 256, 0, st>>>(A, lda, n, j0, kp, Vp.p, ldz);
      EB_CHECK_LAUNCH(c);
      const int nchunk = (n - c0 + chunk - 1) / chunk;
      bt_gram_kernel<<<nchunk, 256, 0, st>>>(Vp.p, ldz, c0, n, chunk, Gp.p);
      EB_CHECK_LAUNCH(c);
      bt_tfactor_kernel<<<1, 256, 0, st>>>(Gp.p, nchunk, tau + j0, kp, Gp.p + (size_t)nchunk_max * 4096, T.p);
      EB_CHECK_LAUNCH(c);
      // W = Z V   (nvec x 64);  W2 = W T^T;  Z -= W2 V^T
      if ((rc = launch_gemm(c, false, false, c->zvec_d.p + c0, ldz, Vp.p + c0, ldz, W.p, 64, nvec, 64, n - c0))) return rc;
      bt_wt_kernel<<<std::min((nvec + 3) / 4, 4 * c->num_sms), 256, 0, st>>>(W.p, T.p, nvec, W2.p);
      EB_CHECK_LAUNCH(c);
      if ((rc = launch_gemm(c, false, true, W2.p, 64, Vp.p + c0, ldz, c->zvec_d.p + c0, ldz, nvec, n - c0, 64, -1.0, 1.0))) return rc;
    }
  }
  normalize_rows_full_kernel<<<nvec, 256, 0, st>>>(c->zvec_d.p, ldz, n);
  EB_CHECK_LAUNCH(c);
  return 0;
}

__global__ void copy_matrix_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < n) dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c];
}

// ------------------------------------------------------------------------------------------ host driver
// all eigenvalues of the tridiagonal (d, e; scaled in place by `scale`) -> c->lambda_d descending
// Section width for `cnt` eigenvalues: with the division-free Sturm sweep the kernel is issue bound once ~64k threads are in flight
// and latency bound below, and a (SEC + 1)-section spends SEC / log2(SEC + 1) times the work of a bisection -- so widen only while
// the threads are free.
static int launch_bisect(eb_ctx* c, int n, const double* d, const double* e2, const double* bounds, double* lam, int k0, int k1) {
  const int cnt = k1 - k0;
  if (cnt <= 0) return 0;
  cudaStream_t st = c->stream;
  const int want = 65536 / cnt;
  if (want >= 8) tri_bisect_kernel<8><<<(unsigned)(((int64_t)cnt * 8 + 127) / 128), 128, 0, st>>>(n, d, e2, bounds, lam, k0, k1);
  else if (want >= 4) tri_bisect_kernel<4><<<(unsigned)(((int64_t)cnt * 4 + 127) / 128), 128, 0, st>>>(n, d, e2, bounds, lam, k0, k1);
  else if (want >= 2) tri_bisect_kernel<2><<<(unsigned)(((int64_t)cnt * 2 + 127) / 128), 128, 0, st>>>(n, d, e2, bounds, lam, k0, k1);
  else tri_bisect_kernel<1><<<(cnt + 127) / 128, 128, 0, st>>>(n, d, e2, bounds, lam, k0, k1);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// dist (collective solves, after the row-distributed reduction has mapped the exchange block): every rank bisects n / world of the
// eigenvalues and the vector is summed over the ranks (the other entries are zero) -- bit-identical everywhere
static int tridiag_spectrum(eb_ctx* c, int n, double* d, double* e, double* e2, double* bounds, double scale, bool dist = false) {
  cudaStream_t st = c->stream;
  scale_de_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d, e, scale);
  EB_CHECK_LAUNCH(c);
  e2_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, e, e2);
  EB_CHECK_LAUNCH(c);
  tri_bounds_kernel<<<1, 1024, 0, st>>>(n, d, e, bounds);
  EB_CHECK_LAUNCH(c);
  if (dist && c->has_comm && c->comm.world > 1 && c->chfsi_sum.p && c->chfsi_sum.n >= (size_t)n + 2) {
    const int W = c->comm.world, me = c->comm.rank;
    const int k0 = (int)((long long)n * me / W), k1 = (int)((long long)n * (me + 1) / W);
    const int64_t cnt = ((int64_t)n + 1) & ~1ll;
    int rc;
    EB_CUDA(cudaMemsetAsync(c->chfsi_sum.p, 0, sizeof(double) * cnt, st));
    // a rank bisects n / world eigenvalues: the lanes the other eigenvalues would have used go into the section width
    if ((rc = launch_bisect(c, n, d, e2, bounds, c->chfsi_sum.p, k0, k1))) return rc;
    if ((rc = peer_allreduce_stream(c, PEER_SLOT_W, c->chfsi_sum.p, c->chfsi_sum.n, cnt, false, 0))) return rc;
    EB_CUDA(cudaMemcpyAsync(c->lambda_d.p, c->chfsi_sum.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  return launch_bisect(c, n, d, e2, bounds, c->lambda_d.p, 0, n);
}

// large-n path: spectrum by two-stage tridiagonalisation (only if lambda_h), leading vectors by subspace iteration
// Deterministic sign: every returned eigenvector has its entry of largest magnitude positive (first such entry on ties).  LAPACK's sign
// is whatever dsteqr's rotations leave (the reference fixes one only with `topright:`, smartpca.c:1267-1281); a fixed convention makes
// the one-stage, full-basis and subspace-iteration paths -- and every GPU of a sharded run -- agree, and it is the sign the reference's
// own example outputs carry (POPGEN/example.evec, EIGENSTRAT/example.pca.evec).
__global__ void __launch_bounds__(256) sign_fix_kernel(double* __restrict__ Z, int64_t ldz, int n) {
  __shared__ double sv[256];
  __shared__ int si[256];
  double* z = Z + (size_t)blockIdx.x * ldz;
  double best = -1.0; int bi = 0;
  for (int i = threadIdx.x; i < n; i += 256) { const double a = fabs(z[i]); if (a > best) { best = a; bi = i; } }
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const double a = sv[threadIdx.x + o]; const int ia = si[threadIdx.x + o];
      if (a > sv[threadIdx.x] || (a == sv[threadIdx.x] && ia < si[threadIdx.x])) { sv[threadIdx.x] = a; si[threadIdx.x] = ia; }
    }
    __syncthreads();
  }
  const bool flip = z[si[0]] < 0.0;
  __syncthreads();
  if (flip) for (int i = threadIdx.x; i < n; i += 256) z[i] = -z[i];
}

static int eig_two_stage(eb_ctx* c, const double* A_d, int64_t lda_in, int n, double scale, int nvec, double* lambda_h, double* evecs_h, bool collective) {
  cudaStream_t st = c->stream;
  int rc;
  c->tm.tridiag_ms = c->tm.bisect_ms = c->tm.vectors_ms = c->tm.band_ms = c->tm.chase_ms = 0.f;
  c->tm.chfsi_iters = c->tm.chfsi_matvecs = 0;
  c->tm.chfsi_converged = 1; c->tm.chfsi_resid = 0.f;
  double lo0 = 0.0;                 // smallest eigenvalue (unscaled) when the spectrum was computed: the filter's lower end
  if ((rc = c->lambda_d.ensure(n))) return rc;
  if (lambda_h) {
    const int64_t lda = ((int64_t)n + 15) & ~15ll;
    if ((rc = c->eigA.ensure((size_t)lda * n))) return rc;
    const size_t wn = (size_t)n * 4 + 64;
    if ((rc = c->eigw.ensure(wn))) return rc;
    double *d = c->eigw.p, *e = d + n, *e2 = e + n, *bounds = e2 + n;
    // spectrum AND vectors: keep the reflectors of both stages and back-transform the eigenvectors of the tridiagonal matrix
    const bool backtr = nvec > 0 && nvec <= 40 && c->opt_eig_vectors == 0;
    EB_CUDA(cudaEventRecord(c->ev[5], st));
    dim3 grid((n + 255) / 256, n);
    copy_matrix_kernel<<<grid, 256, 0, st>>>(A_d, lda_in, c->eigA.p, lda, n);
    EB_CHECK_LAUNCH(c);
    if ((rc = two_stage_tridiag(c, c->eigA.p, lda, n, d, e, collective && n >= c->opt_dist_min, backtr))) return rc;
    EB_CUDA(cudaEventRecord(c->ev[6], st));
    if ((rc = tridiag_spectrum(c, n, d, e, e2, bounds, scale, collective && n >= c->opt_dist_min))) return rc;
    EB_CUDA(cudaEventRecord(c->ev[7], st));
    EB_CUDA(cudaMemcpyAsync(lambda_h, c->lambda_d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->tm.tridiag_ms, c->ev[5], c->ev[6]);
    cudaEventElapsedTime(&c->tm.bisect_ms, c->ev[6], c->ev[7]);
    cudaEventElapsedTime(&c->tm.band_ms, c->ev[5], c->ev[10]);
    cudaEventElapsedTime(&c->tm.chase_ms, c->ev[10], c->ev[6]);
    if (scale > 0.0) lo0 = lambda_h[n - 1] / scale;
    if (backtr) {
      double bnd[4];
      EB_CUDA(cudaMemcpyAsync(bnd, bounds, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
      EB_CUDA(cudaStreamSynchronize(st));
      EB_CUDA(cudaEventRecord(c->ev[8], st));
      const int64_t ldz = ((int64_t)n + 1) & ~1ll;
      c->zvec_ld = ldz;
      const bool prof = getenv("EB_EIG_PROFILE") != nullptr;
      auto lap = [&](const char* what, double& t0) {
        if (!prof) return;
        cudaStreamSynchronize(st);
        const double t1 = (double)clock() / CLOCKS_PER_SEC;
        fprintf(stderr, "[eig profile] vectors: %s %.1f ms\n", what, (t1 - t0) * 1e3);
        t0 = t1;
      };
      double tp = (double)clock() / CLOCKS_PER_SEC;
      if ((rc = tri_vectors(c, n, nvec, d, e, lambda_h, bnd[2], ldz))) return rc;
      lap("inverse iteration on T", tp);
      mgs_rows_kernel<<<1, 1024, 0, st>>>(c->zvec_d.p, ldz, n, nvec);
      EB_CHECK_LAUNCH(c);
      lap("Gram-Schmidt", tp);
      if ((rc = two_stage_backtransform(c, c->eigA.p, lda, n, nvec, c->zvec_d.p, ldz))) return rc;
      lap("Q1 Q2 z", tp);
      normalize_rows_full_kernel<<<nvec, 256, 0, st>>>(c->zvec_d.p, ldz, n);
      EB_CHECK_LAUNCH(c);
      sign_fix_kernel<<<nvec, 256, 0, st>>>(c->zvec_d.p, ldz, n);
      EB_CHECK_LAUNCH(c);
      EB_CUDA(cudaEventRecord(c->ev[9], st));
      if (evecs_h)
        EB_CUDA(cudaMemcpy2DAsync(evecs_h, sizeof(double) * n, c->zvec_d.p, sizeof(double) * ldz, sizeof(double) * n, nvec, cudaMemcpyDeviceToHost, st));
      EB_CUDA(cudaStreamSynchronize(st));
      cudaEventElapsedTime(&c->tm.vectors_ms, c->ev[8], c->ev[9]);
      c->ritz.assign(lambda_h, lambda_h + nvec);
      return 0;
    }
  }
  if (nvec > 0) {
    if ((rc = c->zvec_d.ensure((size_t)nvec * n))) return rc;
    c->zvec_ld = n;
    std::vector<double> th(nvec);
    EB_CUDA(cudaEventRecord(c->ev[8], st));
    if ((rc = chfsi_top(c, A_d, lda_in, n, nvec, th.data(), c->zvec_d.p, &c->tm.chfsi_iters, &c->tm.chfsi_matvecs, lo0, collective && n >= c->opt_dist_min))) {
      if (rc != EB_ERR_NUMERIC) return rc;
      // the subspace iteration gave up: take the vectors from the one-stage path (slower, direct)
      const int keep = c->opt_eig_method;
      c->opt_eig_method = 1;
      rc = eig_resident(c, A_d, lda_in, n, scale, nvec, nullptr, evecs_h);
      c->opt_eig_method = keep;
      c->tm.eig_method = 2;
      return rc;
    }
    sign_fix_kernel<<<nvec, 256, 0, st>>>(c->zvec_d.p, n, n);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaEventRecord(c->ev[9], st));
    if (evecs_h) EB_CUDA(cudaMemcpyAsync(evecs_h, c->zvec_d.p, sizeof(double) * (size_t)nvec * n, cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->tm.vectors_ms, c->ev[8], c->ev[9]);
    c->ritz.assign(th.begin(), th.end());
    for (auto& v : c->ritz) v *= scale;
  }
  return 0;
}

bool eig_uses_two_stage(const eb_ctx* c, int n, int nvec) {
  int method = c->opt_eig_method;
  if (method == 0) method = n >= c->opt_two_stage_min ? 2 : 1;
  return method == 2 && nvec <= 40 && n >= 256;      // block width 64 bounds the subspace iteration
}

int eig_resident(eb_ctx* c, const double* A_d, int64_t lda_in, int n, double scale, int nvec, double* lambda_h, double* evecs_h, bool collective) {
  if (n <= 0) return 0;
  nvec = std::max(0, std::min(nvec, n));
  const int method = (eig_uses_two_stage(c, n, nvec) && !(lda_in & 1)) ? 2 : 1;
  c->tm.eig_method = method;
  if (method == 2) return eig_two_stage(c, A_d, lda_in, n, scale, nvec, lambda_h, evecs_h, collective);
  cudaStream_t st = c->stream;
  int rc;
  const int64_t lda = (n + 15) & ~15;
  if ((rc = c->eigA.ensure((size_t)lda * n))) return rc;
  if ((rc = c->eigV.ensure((size_t)NB * n))) return rc;
  if ((rc = c->eigW.ensure((size_t)NB * n))) return rc;
  // eigw layout: d[n] e[n] tau[n] e2[n] p[n] praw[n] partA[1024] partC[1024] s12[2NB] bounds[4] work[5n] | ipiv
  const size_t wn = (size_t)n * 11 + 2048 + 2 * NB + 8;
  if ((rc = c->eigw.ensure(wn + (size_t)n))) return rc;
  if ((rc = c->lambda_d.ensure(n))) return rc;
  if ((rc = c->zvec_d.ensure((size_t)std::max(nvec, 1) * n))) return rc;
  double* A = c->eigA.p;
  double *d = c->eigw.p, *e = d + n, *tau = e + n, *e2 = tau + n, *p = e2 + n, *praw = p + n, *partA = praw + n, *partC = partA + 1024,
         *s12 = partC + 1024, *bounds = s12 + 2 * NB, *work = bounds + 8;
  int* ipiv = reinterpret_cast<int*>(work + (size_t)5 * n);
  double *Vp = c->eigV.p, *Wp = c->eigW.p;

  EB_CUDA(cudaEventRecord(c->ev[5], st));
  {
    dim3 grid((n + 255) / 256, n);
    copy_matrix_kernel<<<grid, 256, 0, st>>>(A_d, lda_in, A, lda, n);
    EB_CHECK_LAUNCH(c);
  }
  EB_CUDA(cudaMemsetAsync(tau, 0, sizeof(double) * n, st));
  EB_CUDA(cudaMemsetAsync(e, 0, sizeof(double) * n, st));

  const int ncols = std::max(0, n - 2);   // columns that get a reflector: j = 0..n-3
  for (int j0 = 0; j0 < ncols; j0 += NB) {
    const int kp = std::min(NB, ncols - j0);
    for (int k = 0; k < kp; k++) {
      const int j = j0 + k;
      const int len = n - j;                 // columns j..n-1
      const int gA = (len + KA_THREADS - 1) / KA_THREADS;
      tri_row_update_kernel<<<gA, KA_THREADS, 0, st>>>(A, lda, n, j, k, Vp, Wp, d, partA, bounds + 4);
      EB_CHECK_LAUNCH(c);
      const int rows = n - j - 1;
      const int gemv_blocks = (rows + 7) / 8, dot_blocks = (2 * k + 7) / 8;
      tri_gemv_kernel<<<gemv_blocks + dot_blocks, 256, 0, st>>>(A, lda, n, j, k, Vp, Wp, partA, gA, bounds + 4, praw, s12, e, tau, gemv_blocks);
      EB_CHECK_LAUNCH(c);
      const int gC = (rows + KA_THREADS - 1) / KA_THREADS;
      tri_correct_kernel<<<gC, KA_THREADS, 0, st>>>(A, lda, n, j, k, Vp, Wp, partA, gA, bounds + 4, praw, s12, p, partC);
      EB_CHECK_LAUNCH(c);
      tri_w_kernel<<<gC, KA_THREADS, 0, st>>>(n, j, k, Vp, Wp, p, partC, gC, tau);
      EB_CHECK_LAUNCH(c);
    }
    const int j1 = j0 + kp;
    const int rem = n - j1;
    if (rem > 0) {
      dim3 grid((rem + 127) / 128, (rem + 127) / 128);
      tri_trailing_kernel<<<grid, 256, 0, st>>>(A, lda, n, j1, kp, Vp, Wp);
      EB_CHECK_LAUNCH(c);
    }
  }
  tri_tail_scale_kernel<<<1, 32, 0, st>>>(A, lda, n, d, e, scale);
  EB_CHECK_LAUNCH(c);
  EB_CUDA(cudaEventRecord(c->ev[6], st));
  if ((rc = tridiag_spectrum(c, n, d, e, e2, bounds, scale))) return rc;
  EB_CUDA(cudaEventRecord(c->ev[7], st));
  int64_t ldz = n;
  c->zvec_ld = n;
  if (nvec > 64) {
    std::vector<double> lam_h(n);
    double bnd[4];
    EB_CUDA(cudaMemcpyAsync(lam_h.data(), c->lambda_d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaMemcpyAsync(bnd, bounds, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
    EB_CUDA(cudaStreamSynchronize(st));
    if ((rc = full_basis(c, A, lda, n, nvec, d, e, tau, lam_h, bnd[2], &ldz))) return rc;
    c->zvec_ld = ldz;
  } else if (nvec > 0) {
    tri_invit_kernel<<<1, 1024, 0, st>>>(n, nvec, d, e, c->lambda_d.p, bounds, work, ipiv, c->zvec_d.p);
    EB_CHECK_LAUNCH(c);
    tri_backtransform_kernel<<<nvec, 1024, 0, st>>>(A, lda, n, tau, c->zvec_d.p);
    EB_CHECK_LAUNCH(c);
  }
  if (nvec > 0) {
    sign_fix_kernel<<<nvec, 256, 0, st>>>(c->zvec_d.p, ldz, n);
    EB_CHECK_LAUNCH(c);
  }
  EB_CUDA(cudaEventRecord(c->ev[1], st));
  if (lambda_h) EB_CUDA(cudaMemcpyAsync(lambda_h, c->lambda_d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (evecs_h && nvec > 0)
    EB_CUDA(cudaMemcpy2DAsync(evecs_h, sizeof(double) * n, c->zvec_d.p, sizeof(double) * ldz, sizeof(double) * n, nvec, cudaMemcpyDeviceToHost, st));
  EB_CUDA(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&c->tm.tridiag_ms, c->ev[5], c->ev[6]);
  cudaEventElapsedTime(&c->tm.bisect_ms, c->ev[6], c->ev[7]);
  cudaEventElapsedTime(&c->tm.vectors_ms, c->ev[7], c->ev[1]);
  return 0;
}

}  // namespace eb
