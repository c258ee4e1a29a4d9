/* oracle/eig_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the smartpca hot path of DReichLab/EIG (EIGENSOFT 8.0.0), used as the
 * checker for the CUDA product in eig_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; libeigb200.so never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle_pins.py against
 *   (1) the reference's golden files (POPGEN/example.{evec,eval}, POPGEN/grmjunk,
 *       CONVERTF/example.packedancestrymapgeno -- copied as fixtures under tests/golden/), and
 *   (2) the unmodified reference compiled into oracle/_ref/libeigref.so (oracle/Makefile, ref_harness.c).
 * Third-party arithmetic the reference calls but does not vendor (src/Makefile:3 "-lgsl -lopenblas",
 * versions unpinned): LAPACK dspev (eigx.c:107), CBLAS dgemm (kjg_fpca.c:121..173), LAPACKE dgesvd
 * (kjg_gsl.c:203), GSL mt19937 (kjg_gsl.c:96-113).  Their published algorithms are restated here
 * (Householder tridiagonalisation + implicit QL; one-sided Jacobi SVD; Matsumoto-Nishimura MT19937 with
 * GSL's seeding) -- results agree with the LAPACK-backed reference to rounding, not bitwise.
 *
 * All file:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- 2-bit packed genotype accessor: admutils.c:718-735 (rbuff) + admutils.c:575-592 (getgtypes) ----
 * individual k lives in byte k>>2, bits (3-(k&3))*2; codes 0,1,2 = allele count, 3 = missing -> -1 */
static inline int gt (const uint8_t * row, long k)
{
  int g = (row[k >> 2] >> ((3 - (k & 3)) << 1)) & 3;
  return g == 3 ? -1 : g;
}

/* ---- per-SNP allele counts over the selected rows: smartpca.c:3261-3276 (getcolxz_binary1),
 * qpsubs.c:240-248 (getrawcol).  nmiss[s] = number of missing among selected rows. */
void orc_snp_counts (const uint8_t * packed, long nsnp, long rlen, const int *xindex, int nrows,
                     int *c0, int *c1, int *nmiss)
{
  for (long s = 0; s < nsnp; s++) {
    const uint8_t *row = packed + s * rlen;
    int a = 0, b = 0, m = 0;
    for (int j = 0; j < nrows; j++) {
      int g = gt (row, xindex[j]);
      if (g < 0) { m++; continue; }
      a += g; b += 2 - g;
    }
    c0[s] = a; c1[s] = b; nmiss[s] = m;
  }
}

/* ---- per-individual valid-genotype counts over kept SNPs: admutils.c:1075-1097 (numvalidgtallind) ---- */
void orc_indiv_valid_counts (const uint8_t * packed, long nsnp, long rlen, int numindivs,
                             const uint8_t * snp_keep, int *nvalid)
{
  memset (nvalid, 0, sizeof (int) * numindivs);
  for (long s = 0; s < nsnp; s++) {
    if (snp_keep && !snp_keep[s]) continue;
    const uint8_t *row = packed + s * rlen;
    for (int j = 0; j < numindivs; j++) if (gt (row, j) >= 0) nvalid[j]++;
  }
}

/* ---- per-SNP normalisation: smartpca.c:2282-2313 (fvadjust_binary) then smartpca.c:3291-3295.
 * returns -999 if all missing, else nmiss.  cc[3] = yfancy*(k - ymean); *pmean = ymean*yfancy
 * (what the caller stores in xmean[col]); *pfancy = yfancy. */
int orc_fvadjust_binary (int c0, int nmiss, int n, int fancynorm, int altnormstyle,
                         double *cc, double *pmean, double *pfancy)
{
  double p, ynum, ysum, y, ymean, yfancy = 1.0;
  if (n == nmiss) { cc[0] = cc[1] = cc[2] = 0; *pmean = 0; *pfancy = 0; return -999; }
  ynum = n - nmiss; ysum = c0; ymean = ysum / ynum;
  cc[0] = -ymean; cc[1] = 1.0 - ymean; cc[2] = 2.0 - ymean;
  if (fancynorm) {
    p = 0.5 * ymean;
    if (!altnormstyle) p = (ysum + 1.0) / (2.0 * ynum + 2.0);
    y = p * (1.0 - p);
    if (y > 0.0) yfancy = 1.0 / sqrt (y);
  }
  cc[0] *= yfancy; cc[1] *= yfancy; cc[2] *= yfancy;
  *pmean = ymean * yfancy; *pfancy = yfancy;
  return nmiss;
}

/* ---- SNP drop rule: smartpca.c:1131-1144 ---- */
int orc_snp_dropped (int n0, int n1, int tt /* nmiss or -1 */ , int minallelecnt, int maxmissing)
{
  int t = n0 < n1 ? n0 : n1;
  return (t < minallelecnt) || (tt > maxmissing) || (tt < 0) || (t == 0);
}

/* ---- one pass of the GRM region smartpca.c:1088-1236.
 * Mathematically XTX = sum_s x_s x_s^T with x_is = cc_s[g_is] (0 if missing); accumulated per SNP in
 * SNP order over the lower triangle like block_increment_normal (smartpca.c:3498-3528), then mirrored
 * (symit2, smartpca.c:480).  The reference's lookup path sums 5-SNP partial sums first
 * (smartpca.c:3449-3483), so the two agree to rounding only.
 * Outputs per input SNP: c0,c1 (-1,-1 if all missing), nmiss (-1 if all missing), used, xmean, xfancy.
 * XTX is nrows*nrows row-major, NOT divided by y; *y_out = trace/(nrows-1) (smartpca.c:1230). */
int orc_grm (const uint8_t * packed, long nsnp, long rlen, const int *xindex, int nrows,
             int fancynorm, int altnormstyle, int minallelecnt, int maxmissing, const double *weights,
             int *c0, int *c1, int *nmiss, uint8_t * used, double *xmean, double *xfancy,
             double *XTX, double *y_out)
{
  double *x = (double *) malloc (sizeof (double) * nrows);
  int *g = (int *) malloc (sizeof (int) * nrows);
  memset (XTX, 0, sizeof (double) * (size_t) nrows * nrows);
  for (long s = 0; s < nsnp; s++) {
    const uint8_t *row = packed + s * rlen;
    int a = 0, b = 0, m = 0, tt; double cc[3];
    for (int j = 0; j < nrows; j++) {
      g[j] = gt (row, xindex[j]);
      if (g[j] < 0) { m++; continue; }
      a += g[j]; b += 2 - g[j];
    }
    tt = orc_fvadjust_binary (a, m, nrows, fancynorm, altnormstyle, cc, xmean + s, xfancy + s);
    if (tt < -99) { a = b = -1; tt = -1; }
    c0[s] = a; c1[s] = b; nmiss[s] = tt; used[s] = 0;
    if (orc_snp_dropped (a, b, tt, minallelecnt, maxmissing)) continue;
    used[s] = 1;
    if (weights) { cc[0] *= weights[s]; cc[1] *= weights[s]; cc[2] *= weights[s]; }
    for (int j = 0; j < nrows; j++) x[j] = g[j] < 0 ? 0.0 : cc[g[j]];
    for (int i = 0; i < nrows; i++) {
      double xi = x[i]; double *r = XTX + (size_t) i * nrows;
      if (xi == 0.0) continue;
      for (int j = 0; j <= i; j++) r[j] += xi * x[j];
    }
  }
  double tr = 0;
  for (int i = 0; i < nrows; i++) {
    tr += XTX[(size_t) i * nrows + i];
    for (int j = 0; j < i; j++) XTX[(size_t) j * nrows + i] = XTX[(size_t) i * nrows + j];
  }
  *y_out = tr / (double) (nrows - 1);
  free (x); free (g);
  return 0;
}

/* ---- symmetric eigensolver with the eigvecs() contract of eigsubs.c:39-55:
 * mat row-major n*n symmetric (left unchanged), evals descending, evecs[i*n+j] = component j of
 * eigenvector i, unit 2-norm, sign arbitrary.  Algorithm: the one LAPACK dspev publishes
 * (Householder reduction to tridiagonal form, then implicit-shift QL with accumulated transforms). */
static double hyp (double a, double b) { return hypot (a, b); }

int orc_eigvecs (const double *mat, double *evals, double *evecs, int n)
{
  if (n <= 0) return 0;
  double *z = (double *) malloc (sizeof (double) * (size_t) n * n);  /* z[i*n+j], columns become vectors */
  double *d = (double *) malloc (sizeof (double) * n), *e = (double *) malloc (sizeof (double) * n);
  memcpy (z, mat, sizeof (double) * (size_t) n * n);
#define Z(i,j) z[(size_t)(i)*n+(j)]
  /* Householder reduction, working upward from the last row (lower triangle is referenced) */
  for (int i = n - 1; i > 0; i--) {
    int l = i - 1; double h = 0, scale = 0;
    if (l > 0) {
      for (int k = 0; k <= l; k++) scale += fabs (Z (i, k));
      if (scale == 0.0) e[i] = Z (i, l);
      else {
        for (int k = 0; k <= l; k++) { Z (i, k) /= scale; h += Z (i, k) * Z (i, k); }
        double f = Z (i, l), g = f >= 0 ? -sqrt (h) : sqrt (h);
        e[i] = scale * g; h -= f * g; Z (i, l) = f - g; f = 0;
        for (int j = 0; j <= l; j++) {
          Z (j, i) = Z (i, j) / h; g = 0;
          for (int k = 0; k <= j; k++) g += Z (j, k) * Z (i, k);
          for (int k = j + 1; k <= l; k++) g += Z (k, j) * Z (i, k);
          e[j] = g / h; f += e[j] * Z (i, j);
        }
        double hh = f / (h + h);
        for (int j = 0; j <= l; j++) {
          f = Z (i, j); e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; k++) Z (j, k) -= f * e[k] + g * Z (i, k);
        }
      }
    } else e[i] = Z (i, l);
    d[i] = h;
  }
  d[0] = 0; e[0] = 0;
  for (int i = 0; i < n; i++) {
    int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; j++) {
        double g = 0;
        for (int k = 0; k <= l; k++) g += Z (i, k) * Z (k, j);
        for (int k = 0; k <= l; k++) Z (k, j) -= g * Z (k, i);
      }
    }
    d[i] = Z (i, i); Z (i, i) = 1.0;
    for (int j = 0; j <= l; j++) Z (j, i) = Z (i, j) = 0.0;
  }
  /* implicit QL on (d,e), accumulating rotations into z */
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0;
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = fabs (d[m]) + fabs (d[m + 1]);
        if (fabs (e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 300) { free (z); free (d); free (e); return -1; }
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]), r = hyp (g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs (r) : -fabs (r)));
        double s = 1, c = 1, p = 0; int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = c * e[i];
          e[i + 1] = r = hyp (f, g);
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0; break; }
          s = f / r; c = g / r; g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b; d[i + 1] = g + (p = s * r); g = c * r - b;
          for (int k = 0; k < n; k++) {
            f = Z (k, i + 1); Z (k, i + 1) = s * Z (k, i) + c * f; Z (k, i) = c * Z (k, i) - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0;
      }
    } while (m != l);
  }
  /* sort descending, emit rows */
  int *ord = (int *) malloc (sizeof (int) * n);
  for (int i = 0; i < n; i++) ord[i] = i;
  for (int i = 1; i < n; i++) { int t = ord[i], j = i - 1; while (j >= 0 && d[ord[j]] < d[t]) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = t; }
  for (int i = 0; i < n; i++) {
    int c = ord[i]; double nn = 0; evals[i] = d[c];
    if (!evecs) continue;
    for (int k = 0; k < n; k++) nn += Z (k, c) * Z (k, c);
    nn = 1.0 / sqrt (nn);
    for (int k = 0; k < n; k++) evecs[(size_t) i * n + k] = Z (k, c) * nn;
  }
#undef Z
  free (ord); free (z); free (d); free (e);
  return 0;
}

/* ---- outlier detection: smartsubs.c:18-93 (ridoutlier), outliermode 0 and 1.
 * vecno[j] = first eigenvector flagging row j (-1 if none), score[j] its z-score. returns nbad. */
int orc_ridoutlier (const double *evecs, int n, int neigs, double thresh, int mode,
                    int *badlist, int *vecno, double *score)
{
  int nbad = 0;
  if (mode > 1 || n < 3) return 0;
  double *ww = (double *) malloc (sizeof (double) * n), *w2 = (double *) malloc (sizeof (double) * n);
  int *vbad = (int *) calloc (n, sizeof (int));
  for (int j = 0; j < n; j++) vecno[j] = -1;
  for (int i = 0; i < neigs; i++) {
    memcpy (ww, evecs + (size_t) i * n, sizeof (double) * n);
    if (mode == 0) {
      double y1 = 0, y2 = 0;
      for (int j = 0; j < n; j++) y1 += ww[j];
      y1 /= (double) n;
      for (int j = 0; j < n; j++) ww[j] += -y1;
      for (int j = 0; j < n; j++) y2 += ww[j] * ww[j];
      y2 = sqrt (y2 / (double) n);
      for (int j = 0; j < n; j++) ww[j] *= 1.0 / y2;
      for (int j = 0; j < n; j++) if (fabs (ww[j]) > thresh) { vbad[j] = 1; if (vecno[j] < 0) { vecno[j] = i; score[j] = ww[j]; } }
    } else {
      for (int j = 0; j < n; j++) {
        double yy = ww[j], y1 = 0, y2 = 0, zz; ww[j] = 0;
        for (int k = 0; k < n; k++) y1 += ww[k];
        y1 /= (double) (n - 1);
        for (int k = 0; k < n; k++) w2[k] = ww[k] - y1;
        w2[j] = 0;
        for (int k = 0; k < n; k++) y2 += w2[k] * w2[k];
        y2 = sqrt (y2 / (double) n); zz = (yy - y1) / y2;
        if (fabs (zz) > thresh) { vbad[j] = 1; if (vecno[j] < 0) { vecno[j] = i; score[j] = zz; } }
        ww[j] = yy;
      }
    }
  }
  for (int j = 0; j < n; j++) if (vbad[j]) badlist[nbad++] = j;
  free (ww); free (w2); free (vbad);
  return nbad;
}

/* ---- MT19937 as GSL seeds and draws it (kjg_gsl.c:96-113 uses gsl_rng_default = mt19937):
 * mt[0]=seed (0 -> 4357), mt[i] = 1812433253*(mt[i-1]^(mt[i-1]>>30))+i; uniform_pos = get()/2^32, 0 rejected. */
typedef struct { uint32_t mt[624]; int mti; } orc_mt;
static void mt_seed (orc_mt * r, unsigned long s)
{
  if (s == 0) s = 4357;
  r->mt[0] = (uint32_t) s;
  for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t) i;
  r->mti = 624;
}
static uint32_t mt_get (orc_mt * r)
{
  uint32_t *mt = r->mt, y;
  if (r->mti >= 624) {
    for (int k = 0; k < 624; k++) {
      y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    r->mti = 0;
  }
  y = mt[r->mti++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static double mt_upos (orc_mt * r) { double x; do { x = mt_get (r) / 4294967296.0; } while (x == 0); return x; }
uint32_t orc_mt_first (unsigned long seed) { orc_mt r; mt_seed (&r, seed); return mt_get (&r); }

/* ---- seeded Gaussian n x L matrix, row-major: kjg_gsl.c:145-186 (Marsaglia polar pairs; an odd last
 * column receives a plain uniform, kjg_gsl.c:183-184) ---- */
void orc_gauss_matrix (long seed, long n, long L, double *out)
{
  orc_mt r; mt_seed (&r, (unsigned long) seed);
  for (long i = 0; i < n; i++) {
    double *data = out + i * L; long j;
    for (j = 0; j + 1 < L; j += 2) {
      double x0, x1, r2;
      do { x0 = -1 + 2 * mt_upos (&r); x1 = -1 + 2 * mt_upos (&r); r2 = x0 * x0 + x1 * x1; } while (r2 > 1.0 || r2 == 0);
      r2 = sqrt (-2.0 * log (r2) / r2);
      data[j] = x0 * r2; data[j + 1] = x1 * r2;
    }
    if (L % 2) data[L - 1] = mt_upos (&r);
  }
}

/* ---- fastmode per-SNP 4-entry tables: gval.c:56-86 (setgval) via getcolxz/fvadjust (smartpca.c:3129,2236):
 * gtable[s][k] = (k - mean)*fancy/sqrt(2), gtable[s][3] = 0.  mono[s]=1 when min(n0,n1)==0 (gval.c:81-83). */
void orc_gtable (const uint8_t * packed, long nsnp, long rlen, const int *xindex, int nrows,
                 int fancynorm, int altnormstyle, double *gtable, uint8_t * mono)
{
  for (long s = 0; s < nsnp; s++) {
    const uint8_t *row = packed + s * rlen;
    double ynum = 0, ysum = 0, ymean, yfancy = 1.0, xm, xf; int a = 0, b = 0;
    for (int j = 0; j < nrows; j++) { int g = gt (row, xindex[j]); if (g < 0) continue; ynum += 1; ysum += g; a += g; b += 2 - g; }
    if (ynum == 0.0) { xm = 0; xf = 0; a = b = -1; }
    else {
      ymean = ysum / ynum;
      if (fancynorm) { double p = 0.5 * ymean, y; if (!altnormstyle) p = (ysum + 1.0) / (2.0 * ynum + 2.0); y = p * (1.0 - p); if (y > 0.0) yfancy = 1.0 / sqrt (y); }
      xm = ymean * yfancy; xf = yfancy;
    }
    double mean = xm / xf;
    for (int k = 0; k < 3; k++) { double y = ((double) k) - mean; y *= xf; gtable[s * 4 + k] = y / sqrt (2.0); }
    gtable[s * 4 + 3] = 0;
    if (mono) mono[s] = ((a < b ? a : b) == 0);
  }
}

/* one-sided Jacobi SVD of row-major A (m x c, m>=c): A := U (left singular vectors, columns ordered by
 * descending singular value), S[c].  Published Hestenes algorithm; stands in for LAPACKE_dgesvd('O','S')
 * at kjg_gsl.c:203 (V is discarded by the caller, kjg_fpca.c:63-71,85). */
static void jacobi_svd_u (double *A, long m, long c, double *S)
{
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (long p = 0; p < c - 1; p++) for (long q = p + 1; q < c; q++) {
      double al = 0, be = 0, ga = 0;
      for (long i = 0; i < m; i++) { double x = A[i * c + p], y = A[i * c + q]; al += x * x; be += y * y; ga += x * y; }
      if (ga == 0.0 || fabs (ga) <= 1e-16 * sqrt (al * be)) continue;
      off = fmax (off, fabs (ga) / sqrt (al * be));
      double zeta = (be - al) / (2.0 * ga), t = (zeta >= 0 ? 1.0 : -1.0) / (fabs (zeta) + sqrt (1.0 + zeta * zeta));
      double cs = 1.0 / sqrt (1.0 + t * t), sn = cs * t;
      for (long i = 0; i < m; i++) { double x = A[i * c + p], y = A[i * c + q]; A[i * c + p] = cs * x - sn * y; A[i * c + q] = sn * x + cs * y; }
    }
    if (off < 1e-15) break;
  }
  long *ord = (long *) malloc (sizeof (long) * c);
  for (long j = 0; j < c; j++) { double s = 0; for (long i = 0; i < m; i++) s += A[i * c + j] * A[i * c + j]; S[j] = sqrt (s); ord[j] = j; }
  for (long i = 1; i < c; i++) { long t = ord[i], j = i - 1; while (j >= 0 && S[ord[j]] < S[t]) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = t; }
  double *row = (double *) malloc (sizeof (double) * c), *S2 = (double *) malloc (sizeof (double) * c);
  for (long j = 0; j < c; j++) S2[j] = S[ord[j]];
  for (long i = 0; i < m; i++) {
    for (long j = 0; j < c; j++) row[j] = S2[j] > 0 ? A[i * c + ord[j]] / S2[j] : 0.0;
    memcpy (A + i * c, row, sizeof (double) * c);
  }
  memcpy (S, S2, sizeof (double) * c);
  free (ord); free (row); free (S2);
}

/* ---- fastmode randomised PCA: kjg_fpca.c:24-101 with products kjg_fpca.c:104-178 and the decode of
 * gval.c:173-213.  X is m x n (SNP-major), x_si = gtable[s][code(s, xindex[i])].
 * eval[K], evec[n*K] row-major (as kjg_fpca leaves it, before smartpca.c:971 transposes). */
int orc_fpca (const uint8_t * packed, long m, long rlen, const int *xindex, long n, int fancynorm,
              int altnormstyle, long K, long L, long I, long seed, double *eval, double *evec)
{
  if (K >= L || I == 0) return -1;
  long c = (I + 1) * L;
  double *gtab = (double *) malloc (sizeof (double) * 4 * m);
  double *G1 = (double *) malloc (sizeof (double) * n * L), *G2 = (double *) malloc (sizeof (double) * n * L);
  double *Q = (double *) calloc ((size_t) m * c, sizeof (double)), *B = (double *) calloc ((size_t) n * c, sizeof (double));
  double *x = (double *) malloc (sizeof (double) * n), *S = (double *) malloc (sizeof (double) * c);
  orc_gtable (packed, m, rlen, xindex, (int) n, fancynorm, altnormstyle, gtab, NULL);
  orc_gauss_matrix (seed, n, L, G1);
#define DECODE(s) do { const uint8_t *row_ = packed + (s) * rlen; for (long i_ = 0; i_ < n; i_++) { int g_ = (row_[xindex[i_] >> 2] >> ((3 - (xindex[i_] & 3)) << 1)) & 3; x[i_] = gtab[(s) * 4 + g_]; } } while (0)
  for (long it = 0; it <= I; it++) {
    /* Q_it = X * G1 */
    for (long s = 0; s < m; s++) {
      DECODE (s);
      for (long l = 0; l < L; l++) { double a = 0; for (long i = 0; i < n; i++) a += x[i] * G1[i * L + l]; Q[s * c + it * L + l] = a; }
    }
    if (it == I) break;
    /* G2 = X^T * Q_it / m */
    memset (G2, 0, sizeof (double) * n * L);
    for (long s = 0; s < m; s++) {
      DECODE (s);
      for (long i = 0; i < n; i++) { double xi = x[i]; if (xi != 0.0) for (long l = 0; l < L; l++) G2[i * L + l] += xi * Q[s * c + it * L + l]; }
    }
    for (long i = 0; i < n * L; i++) G2[i] *= 1.0 / m;
    double *t = G1; G1 = G2; G2 = t;
  }
  jacobi_svd_u (Q, m, c, S);
  for (long s = 0; s < m; s++) {
    DECODE (s);
    for (long i = 0; i < n; i++) { double xi = x[i]; if (xi != 0.0) for (long l = 0; l < c; l++) B[i * c + l] += xi * Q[s * c + l]; }
  }
  jacobi_svd_u (B, n, c, S);
  for (long i = 0; i < n; i++) for (long k = 0; k < K; k++) evec[i * K + k] = B[i * c + k];
  for (long k = 0; k < K; k++) eval[k] = S[k] * S[k] * (1.0 / m);
#undef DECODE
  free (gtab); free (G1); free (G2); free (Q); free (B); free (x); free (S);
  return 0;
}

/* ---- SNP loadings and sample projections: smartpca.c:1444 (setfvecs: fvecs = 10*evecs rows),
 * smartpca.c:1485-1494 (ffvecs[j*ncols+s] = sum_k fvecs[j*nrows+k]*x_ks) and smartpca.c:1504-1525
 * (fxvecs via loadxdataind/fixxrow (qpsubs.c:322-352); fxscal[j] = 1/sqrt(sum_i fxvecs_ji^2)).
 * used[s]==0 SNPs are "ignore" -> zero column (smartpca.c:3576-3579; loadxdataind leaves -1 -> 0). */
void orc_project (const uint8_t * packed, long ncols, long rlen, const int *xindex, int nrows,
                  const uint8_t * used, const double *xmean, const double *xfancy,
                  const double *evecs /* numeigs rows of length nrows */ , int numeigs,
                  double *ffvecs, double *fxvecs, double *fxscal)
{
  memset (ffvecs, 0, sizeof (double) * (size_t) numeigs * ncols);
  for (long s = 0; s < ncols; s++) {
    if (!used[s]) continue;
    const uint8_t *row = packed + s * rlen;
    double mean = xfancy[s] != 0 ? xmean[s] / xfancy[s] : 0;
    for (int j = 0; j < numeigs; j++) {
      double a = 0;
      for (int k = 0; k < nrows; k++) { int g = gt (row, xindex[k]); if (g >= 0) a += 10.0 * evecs[(size_t) j * nrows + k] * ((g - mean) * xfancy[s]); }
      ffvecs[(size_t) j * ncols + s] = a;
    }
  }
  for (int j = 0; j < numeigs; j++) fxscal[j] = 0;
  for (int i = 0; i < nrows; i++) for (int j = 0; j < numeigs; j++) {
    double y = 0;
    for (long s = 0; s < ncols; s++) {
      if (!used[s]) continue;
      int g = gt (packed + s * rlen, xindex[i]);
      if (g >= 0) y += (g * xfancy[s] - xmean[s]) * ffvecs[(size_t) j * ncols + s];
    }
    fxvecs[(size_t) j * nrows + i] = y; fxscal[j] += y * y;
  }
  for (int j = 0; j < numeigs; j++) fxscal[j] = 1.0 / sqrt (fxscal[j]);
}

/* ---------------------------------------------------------------------------------------------------------------
 * lsqproj + seteigscale: the .evec coordinates (restates smartpca.c:4606-4757, regsubs.c:8-77 regressit,
 * nicksrc/linsubs.c:296-393 solvit/choldc/cholsl, smartpca.c:4758-4781 seteigscale, smartpca.c:1553-1564).
 * indiv[nlist]: the non-ignored individuals (the reference walks all of indivmarkers, smartpca.c:4651).
 * For individual q: rows kk over SNPs with used[s] and a valid genotype; emat[kk][j] = fxscal[j]*ffvecs[j][s],
 * rhs[kk] = g*xfancy[s]-xmean[s]; normal equations + Cholesky; bcoeffs[j] = fxscal[j]*sum_s xrow[s]*ffvecs[j][s]
 * (missing -> 0).  An individual with kk <= numeigs (or a non positive definite normal matrix) gets ok = 0 and
 * zero coefficients.  nvalid[q] = kk. */
static int orc_choldc (double *a, int n, double *p)
{
  for (int i = 0; i < n; i++) p[i] = 0;
  for (int i = 0; i < n; i++)
    for (int j = i; j < n; j++) {
      double sum = a[i * n + j];
      for (int k = i - 1; k >= 0; k--) sum -= a[i * n + k] * a[j * n + k];
      if (i == j) { if (sum <= 0.0) return -1; p[i] = sqrt (sum); }
      else a[j * n + i] = sum / p[i];
    }
  return 1;
}
static void orc_cholsl (const double *a, int n, const double *p, const double *b, double *x)
{
  for (int i = 0; i < n; i++) { double sum = b[i]; for (int k = i - 1; k >= 0; k--) sum -= a[i * n + k] * x[k]; x[i] = sum / p[i]; }
  for (int i = n - 1; i >= 0; i--) { double sum = x[i]; for (int k = i + 1; k < n; k++) sum -= a[k * n + i] * x[k]; x[i] = sum / p[i]; }
}

void orc_lsqproj (const uint8_t * packed, long ncols, long rlen, const int *indiv, int nlist,
                  const uint8_t * used, const double *xmean, const double *xfancy,
                  const double *ffvecs /* [numeigs][ncols] */ , const double *fxscal, int numeigs,
                  double *acoeffs /* [numeigs][nlist] */ , double *bcoeffs /* [numeigs][nlist] */ , int *nvalid, uint8_t * ok)
{
  const int n = numeigs;
  double *co = (double *) malloc (sizeof (double) * n * n), *rr = (double *) malloc (sizeof (double) * n),
    *p = (double *) malloc (sizeof (double) * n), *ans = (double *) malloc (sizeof (double) * n), *e = (double *) malloc (sizeof (double) * n);
  for (int q = 0; q < nlist; q++) {
    memset (co, 0, sizeof (double) * n * n); memset (rr, 0, sizeof (double) * n);
    int kk = 0;
    for (int j = 0; j < n; j++) bcoeffs[(size_t) j * nlist + q] = 0;
    for (long s = 0; s < ncols; s++) {
      if (!used[s]) continue;
      const int g = gt (packed + s * rlen, indiv[q]);
      if (g < 0) continue;
      const double x = g * xfancy[s] - xmean[s];
      for (int j = 0; j < n; j++) e[j] = fxscal[j] * ffvecs[(size_t) j * ncols + s];
      for (int j = 0; j < n; j++) {
        rr[j] += e[j] * x;
        for (int k = j; k < n; k++) co[j * n + k] = co[k * n + j] += e[j] * e[k];
      }
      ++kk;
    }
    nvalid[q] = kk; ok[q] = 0;
    for (int j = 0; j < n; j++) acoeffs[(size_t) j * nlist + q] = 0;
    /* bcoeffs: fxscal[j] * vdot(xrow, ffvecs_j) over all SNPs with missing -> 0 == rr up to summation order */
    for (int j = 0; j < n; j++) {
      double y = 0;
      for (long s = 0; s < ncols; s++) {
        if (!used[s]) continue;
        const int g = gt (packed + s * rlen, indiv[q]);
        if (g >= 0) y += (g * xfancy[s] - xmean[s]) * ffvecs[(size_t) j * ncols + s];
      }
      bcoeffs[(size_t) j * nlist + q] = fxscal[j] * y;
    }
    if (kk <= n) continue;
    if (orc_choldc (co, n, p) < 0) continue;
    orc_cholsl (co, n, p, rr, ans);
    for (int j = 0; j < n; j++) acoeffs[(size_t) j * nlist + q] = ans[j];
    ok[q] = 1;
  }
  free (co); free (rr); free (p); free (ans); free (e);
}

/* seteigscale (smartpca.c:4758-4781): eigscale[j] = <a_j, b_j> / <a_j, a_j> over the PCA rows; rowpos[k] = position of
 * PCA row k inside the list the coefficients are stored for. */
void orc_seteigscale (const double *acoeffs, const double *bcoeffs, int nlist, const int *rowpos, int nrows, int numeigs, double *eigscale)
{
  for (int j = 0; j < numeigs; j++) {
    double ab = 0, aa = 0;
    for (int k = 0; k < nrows; k++) {
      const double a = acoeffs[(size_t) j * nlist + rowpos[k]], b = bcoeffs[(size_t) j * nlist + rowpos[k]];
      ab += a * b; aa += a * a;
    }
    eigscale[j] = ab / aa;
  }
}

/* dense path (restates block_increment_normal, smartpca.c:3498-3528, + symit2): XTX = sum_s x_s x_s^T, SNP by SNP in
 * the reference's order; tblock_all[ncols][nrows]. */
void orc_dense_grm (const double *tblock_all, long ncols, int nrows, double *XTX)
{
  memset (XTX, 0, sizeof (double) * (size_t) nrows * nrows);
  for (long s = 0; s < ncols; s++) {
    const double *x = tblock_all + s * nrows;
    for (int i = 0; i < nrows; i++) { const double xi = x[i]; for (int j = 0; j <= i; j++) XTX[(size_t) i * nrows + j] += xi * x[j]; }
  }
  for (int i = 0; i < nrows; i++) for (int j = 0; j < i; j++) XTX[(size_t) j * nrows + i] = XTX[(size_t) i * nrows + j];
}

/* per-population genotype-class counts (the counting loop of fstcolyy, qpsubs.c:1256-1281) and the Fst numerator /
 * denominator of one SNP for every population pair, non-inbreed branch (qpsubs.c:1303-1340).
 * counts[k*3+g]; estn/estd numeg*numeg with the reference's initial values (0 / -1, diagonal of estd 0). */
void orc_pop_counts (const uint8_t * packed, long nsnp, long rlen, const int *xindex, const int *xtypes, int nrows, int npops, int *counts)
{
  memset (counts, 0, sizeof (int) * (size_t) nsnp * npops * 3);
  for (long s = 0; s < nsnp; s++)
    for (int i = 0; i < nrows; i++) {
      const int k = xtypes[i];
      if (k < 0 || k >= npops) continue;
      const int g = gt (packed + s * rlen, xindex[i]);
      if (g >= 0) counts[((size_t) s * npops + k) * 3 + g]++;
    }
}
void orc_fstcol (const int *cnt /* [numeg][3] */ , int numeg, double *estn, double *estd)
{
  for (int i = 0; i < numeg * numeg; i++) { estn[i] = 0.0; estd[i] = -1.0; }
  for (int a = 0; a < numeg; a++) estd[a * numeg + a] = 0.0;
  for (int i = 0; i < numeg; i++)
    for (int j = i + 1; j < numeg; j++) {
      const double ya = cnt[i * 3 + 1] + 2 * cnt[i * 3 + 2], yb = cnt[i * 3 + 1] + 2 * cnt[i * 3 + 0];
      const double yaa = cnt[j * 3 + 1] + 2 * cnt[j * 3 + 2], ybb = cnt[j * 3 + 1] + 2 * cnt[j * 3 + 0];
      const double zz = yaa + ybb, z = ya + yb;
      if ((z < 1.5) || (zz < 1.5)) continue;
      double yt = ya + yb;
      const double p1 = ya / yt, h1 = ya * yb / (yt * (yt - 1.0));
      yt = yaa + ybb;
      const double p2 = yaa / yt, h2 = yaa * ybb / (yt * (yt - 1.0));
      double en = (p1 - p2) * (p1 - p2);
      en -= h1 / z; en -= h2 / zz;
      double ed = en; ed += h1; ed += h2;
      estn[i * numeg + j] = estn[j * numeg + i] = en;
      estd[i * numeg + j] = estd[j * numeg + i] = ed;
    }
}
