/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into libeigb200.so).
 *
 * Thin ctypes-callable wrappers around the UNMODIFIED reference functions on the smartpca hot
 * path.  The reference translation unit is compiled from where it lies under /root/reference
 * (textual #include below, `main` renamed) so that its file-static thread table
 * (smartpca.c:526-534) is reachable; nothing from the reference is copied into this repository.
 * Built by oracle/Makefile into oracle/_ref/libeigref.so (git-ignored, travels with gpurun).
 *
 * Each wrapper only marshals flat arrays into the reference's SNP / Indiv structs and then runs
 * the reference's own calls in the order smartpca.c:main makes them:
 *   refh_grm        -> smartpca.c:1088-1236  (getcolxz_binary1/2, domult_increment_lookup, symit2)
 *   refh_eigvecs    -> eigsubs.c:39 (dspev_ via eigx.c:97)
 *   refh_ridoutlier -> smartsubs.c:18
 *   refh_fpca       -> gval.c:31 setgval + kjg_fpca.c:24 kjg_fpca
 *   refh_gauss      -> kjg_gsl.c:96,166 (seeded Gaussian matrix)
 */
#define main smartpca_reference_main
#include "eigensrc/smartpca.c"
#undef main
#include <time.h>
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_rng.h>
#include "kjg_gsl.h"

static double now_s (void) { struct timespec t; clock_gettime (CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static int dummy_gtypes[1];
static SNP *hsnps; static SNP **hsnpp; static Indiv *hind; static Indiv **hindp;
static void hfree (void) { free (hsnps); free (hsnpp); free (hind); free (hindp); hsnps = NULL; hsnpp = NULL; hind = NULL; hindp = NULL; }
static void hbuild (const unsigned char *packed, long nsnp, long rl, int nind)
{
  long i;
  hfree ();
  hsnps = (SNP *) calloc (nsnp, sizeof (SNP)); hsnpp = (SNP **) calloc (nsnp, sizeof (SNP *));
  hind = (Indiv *) calloc (nind, sizeof (Indiv)); hindp = (Indiv **) calloc (nind, sizeof (Indiv *));
  for (i = 0; i < nsnp; i++) {
    snprintf (hsnps[i].ID, IDSIZE, "s%ld", i);
    hsnps[i].pbuff = (char *) packed + i * rl; hsnps[i].ngtypes = nind; hsnps[i].gtypes = dummy_gtypes;
    hsnps[i].chrom = 1; hsnps[i].weight = 1.0; hsnpp[i] = hsnps + i;
  }
  for (i = 0; i < nind; i++) { snprintf (hind[i].ID, IDSIZE, "i%ld", i); hind[i].affstatus = YES; hind[i].egroup = "P"; hindp[i] = hind + i; }
  packmode = YES;
}

/* One pass of the GRM region. Outputs are per input SNP (length nsnp):
 * c0,c1 (=n0,n1; -1 if all missing), nmiss (-1 if all missing), used (1 if it entered XTX), xmean,xfancy.
 * XTX_out: nrows*nrows row-major after symit2, NOT yet divided by y; *y_out = trace/(nrows-1).
 * secs[0] = wall seconds of the per-SNP loop (stats+pack+lookup), secs[1] = seconds inside domult_increment_lookup. */
static int grm_core (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, int nrows,
              int fancy, int altnorm, int minac, int maxmiss, const double *weights, int nthreads,
              int *c0, int *c1, int *nmiss, unsigned char *used, double *xmean_o, double *xfancy_o,
              double *XTX_out, double *y_out, double *secs, int finish)
{
  long i; int n0, n1, t, tt, xblock = 0, blocksize = 20; uint32_t thread_ct; double y, t0, t1, tl = 0;
  pthread_t threads[MAX_THREADS];
  int *rawcol, *xidx; uintptr_t *bc, *bm; double *tblock, *lut, cc[3];
  hbuild (packed, nsnp, rl, nind);
  fancynorm = fancy; altnormstyle = altnorm;
  thread_ct = nthreads < 1 ? 1 : nthreads; if (thread_ct > MAX_THREADS) thread_ct = MAX_THREADS;
  if (thread_ct > (uint32_t) nrows * 2) { thread_ct = nrows / 2; if (!thread_ct) thread_ct = 1; }
  triangle_fill (g_thread_start, nrows, thread_ct, 0, 1, 0, 1);
  ZALLOC (rawcol, nrows, int); ZALLOC (bc, nrows, uintptr_t); ZALLOC (bm, nrows, uintptr_t);
  ZALLOC (tblock, 3 * blocksize, double); ZALLOC (lut, 131072, double); ZALLOC (xidx, nrows, int);
  memcpy (xidx, xindex_in, sizeof (int) * nrows);
  vzero (XTX_out, ((long) nrows * (nrows + 1)) / 2);
  t0 = now_s ();
  for (i = 0; i < nsnp; i++) {
    tt = getcolxz_binary1 (rawcol, cc, hsnpp[i], xidx, nrows, (int) i, xmean_o, xfancy_o, &n0, &n1);
    c0[i] = n0; c1[i] = n1; nmiss[i] = tt; used[i] = 0;
    t = MIN (n0, n1);
    if ((t < minac) || (tt > maxmiss) || (tt < 0) || (t == 0)) continue;
    getcolxz_binary2 (rawcol, bc, bm, xblock, nrows);
    if (weights) vst (cc, cc, weights[i], 3);
    copyarr (cc, &(tblock[xblock * 3]), 3);
    used[i] = 1; ++xblock;
    if (xblock == blocksize) {
      t1 = now_s ();
      domult_increment_lookup (threads, thread_ct, XTX_out, tblock, bc, bm, xblock, nrows, lut);
      tl += now_s () - t1;
      memset (bc, 0, sizeof (uintptr_t) * nrows); memset (bm, 0, sizeof (uintptr_t) * nrows);
      vzero (tblock, 3 * blocksize); xblock = 0;
    }
  }
  if (xblock > 0) { t1 = now_s (); domult_increment_lookup (threads, thread_ct, XTX_out, tblock, bc, bm, xblock, nrows, lut); tl += now_s () - t1; }
  if (secs) { secs[0] = now_s () - t0; secs[1] = tl; }
  if (finish) {
    symit2 (XTX_out, nrows);
    y = trace (XTX_out, nrows) / (double) (nrows - 1);
    *y_out = y;
  }
  free (rawcol); free (bc); free (bm); free (tblock); free (lut); free (xidx); hfree ();
  return 0;
}

int refh_grm (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, int nrows,
              int fancy, int altnorm, int minac, int maxmiss, const double *weights, int nthreads,
              int *c0, int *c1, int *nmiss, unsigned char *used, double *xmean_o, double *xfancy_o,
              double *XTX_out, double *y_out, double *secs)
{
  return grm_core (packed, nsnp, rl, nind, xindex_in, nrows, fancy, altnorm, minac, maxmiss, weights, nthreads, c0, c1, nmiss, used,
                   xmean_o, xfancy_o, XTX_out, y_out, secs, 1);
}

/* the per-SNP loop alone (smartpca.c:1116-1221) into a caller-provided packed lower triangle of nrows(nrows+1)/2 doubles:
 * the timing arm for matrices whose square (symit2 needs nrows^2 doubles, once per pass) is not worth allocating for a
 * bounded SNP sample -- bench.py's CPU baseline at 50,000 individuals. */
int refh_grm_loop (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, int nrows,
                   int fancy, int altnorm, int minac, int maxmiss, int nthreads,
                   int *c0, int *c1, int *nmiss, unsigned char *used, double *xmean_o, double *xfancy_o, double *XTX_tri, double *secs)
{
  double y;
  return grm_core (packed, nsnp, rl, nind, xindex_in, nrows, fancy, altnorm, minac, maxmiss, NULL, nthreads, c0, c1, nmiss, used,
                   xmean_o, xfancy_o, XTX_tri, &y, secs, 0);
}

void refh_eigvecs (double *mat, double *evals, double *evecs, int n) { eigvecs (mat, evals, evecs, n); }

int refh_ridoutlier (double *evecs, int n, int neigs, double thresh, int mode, int *badlist, int *vecno, double *score)
{
  OUTLINFO **oi; int k, nbad;
  oi = (OUTLINFO **) calloc (n, sizeof (OUTLINFO *));
  for (k = 0; k < n; k++) oi[k] = (OUTLINFO *) calloc (1, sizeof (OUTLINFO));
  setoutliermode (mode);
  nbad = ridoutlier (evecs, n, neigs, thresh, badlist, oi);
  for (k = 0; k < n; k++) { vecno[k] = oi[k]->vecno; score[k] = oi[k]->score; free (oi[k]); }
  free (oi);
  return nbad;
}

/* fastmode: evec is n*K row-major exactly as kjg_fpca leaves it (before smartpca.c:971 transposes). */
int refh_fpca (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, int nrows,
               int fancy, int altnorm, long K, long L, long I, long seed_in, double *eval, double *evec, double *secs)
{
  int *xidx, *xt; double t0;
  hbuild (packed, nsnp, rl, nind);
  fancynorm = fancy; altnormstyle = altnorm; usepopsformissing = NO; seed = seed_in;
  ZALLOC (xidx, nrows, int); ZALLOC (xt, nrows, int); memcpy (xidx, xindex_in, sizeof (int) * nrows);
  setgval (hsnpp, nrows, hindp, nind, xidx, xt, (int) nsnp);
  t0 = now_s ();
  kjg_fpca (K, L, I, eval, evec);
  if (secs) secs[0] = now_s () - t0;
  unsetgval (); free (xidx); free (xt); hfree ();
  return 0;
}

void refh_gauss (long seed_in, long n, long L, double *out)
{
  gsl_matrix_view v = gsl_matrix_view_array (out, n, L); gsl_rng *r;
  seed = seed_in; r = kjg_gsl_rng_init (); kjg_gsl_ran_ugaussian_matrix (r, &v.matrix); gsl_rng_free (r);
}

unsigned long refh_mt_first (unsigned long s) { gsl_rng *r; unsigned long v; gsl_rng_default_seed = s; r = gsl_rng_alloc (gsl_rng_default); v = gsl_rng_get (r); gsl_rng_free (r); return v; }
void refh_openblas_threads (int n) { extern void openblas_set_num_threads (int); openblas_set_num_threads (n); }

/* .evec coordinates: the reference's own post-eigen sequence, smartpca.c:1440-1564 (setfvecs, getcolxf loadings,
 * loadxdataind/fixxrow sample projections, lsqproj, seteigscale, mulmat) on flat inputs.  lsqproj keeps file-static
 * work arrays sized by its first call (smartpca.c:4614,4630-4636), so every invocation runs in a forked child and
 * hands its results back through shared anonymous mappings.
 * acoeffs_o / bcoeffs_o: [k][nind] (all individuals; rows of ignored individuals stay 0), ignored_o[nind]. */
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
int refh_evec_coords (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, int nrows,
                      int fancy, int altnorm, const unsigned char *used, const double *xmean_in, const double *xfancy_in,
                      const double *evecs_in, int k, const unsigned char *indiv_ignore,
                      double *acoeffs_o, double *bcoeffs_o, double *eigscale_o, double *ffvecs_o, double *fxscal_o, unsigned char *ignored_o)
{
  size_t na = sizeof (double) * (size_t) k * nind, nf = sizeof (double) * (size_t) k * nsnp;
  size_t tot = 2 * na + nf + sizeof (double) * 2 * k + nind;
  unsigned char *sh = mmap (NULL, tot, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (sh == MAP_FAILED) return -1;
  pid_t pid = fork ();
  if (pid < 0) return -2;
  if (pid == 0) {
    long i; int j, kk_; double y;
    double *acoeffs = (double *) sh, *bcoeffs = (double *) (sh + na), *ffv = (double *) (sh + 2 * na), *esc = (double *) (sh + 2 * na + nf),
      *fxs = esc + k; unsigned char *ign = sh + 2 * na + nf + sizeof (double) * 2 * k;
    double *fvecs, *fxvecs, *cc, *xrow, *eigscmat; int *xidx, *xt;
    fclose (stdout); stdout = fopen ("/dev/null", "w");
    hbuild (packed, nsnp, rl, nind);
    fancynorm = fancy; altnormstyle = altnorm; usepopsformissing = NO; regmode = YES; plotmode = NO; printcover = NO;
    numeigs = k; numindivs = nind;
    for (i = 0; i < nsnp; i++) hsnps[i].ignore = used[i] ? NO : YES;
    for (i = 0; i < nind; i++) { hind[i].idnum = (int) i; hind[i].ignore = indiv_ignore && indiv_ignore[i] ? YES : NO; }
    ZALLOC (xmean, nsnp, double); ZALLOC (xfancy, nsnp, double);
    memcpy (xmean, xmean_in, sizeof (double) * nsnp); memcpy (xfancy, xfancy_in, sizeof (double) * nsnp);
    ZALLOC (xidx, nrows, int); memcpy (xidx, xindex_in, sizeof (int) * nrows);
    ZALLOC (xt, nind, int);
    ZALLOC (fvecs, (long) nrows * k, double); ZALLOC (fxvecs, (long) nrows * k, double); ZALLOC (cc, nrows > 3 ? nrows : 3, double);
    ZALLOC (xrow, nsnp, double);
    setfvecs (fvecs, (double *) evecs_in, nrows, k);
    for (i = 0; i < nsnp; i++) {
      getcolxf (cc, hsnpp[i], xidx, nrows, (int) i, NULL, NULL);
      for (j = 0; j < k; j++) for (kk_ = 0; kk_ < nrows; kk_++) ffv[j * nsnp + i] += fvecs[j * nrows + kk_] * cc[kk_];
    }
    for (i = 0; i < nrows; i++) {
      loadxdataind (xrow, hsnpp, xidx[i], (int) nsnp);
      fixxrow (xrow, xmean, xfancy, (int) nsnp);
      for (j = 0; j < k; j++) { y = fxvecs[j * nrows + i] = vdot (xrow, ffv + j * nsnp, (int) nsnp); fxs[j] += y * y; }
    }
    for (j = 0; j < k; j++) fxs[j] = 1.0 / sqrt (fxs[j]);
    lsqproj (-99, hsnpp, (int) nsnp, hindp, nind, fxs, ffv, acoeffs, bcoeffs, xt, 1);
    seteigscale (esc, acoeffs, bcoeffs, xidx, nrows, k);
    ZALLOC (eigscmat, k * k, double);
    setdiag (eigscmat, esc, k);
    mulmat (acoeffs, eigscmat, acoeffs, k, k, nind);
    for (i = 0; i < nind; i++) ign[i] = hind[i].ignore ? 1 : 0;
    _exit (0);
  }
  int status = 0;
  waitpid (pid, &status, 0);
  if (!WIFEXITED (status) || WEXITSTATUS (status) != 0) { munmap (sh, tot); return -3; }
  memcpy (acoeffs_o, sh, na); memcpy (bcoeffs_o, sh + na, na); memcpy (ffvecs_o, sh + 2 * na, nf);
  memcpy (eigscale_o, sh + 2 * na + nf, sizeof (double) * k); memcpy (fxscal_o, sh + 2 * na + nf + sizeof (double) * k, sizeof (double) * k);
  memcpy (ignored_o, sh + 2 * na + nf + sizeof (double) * 2 * k, nind);
  munmap (sh, tot);
  return 0;
}

/* shrinkmode: the reference's own doshrinkp / doshrinkp2 (smartpca.c:4223-4419, 4022-4220) on flat inputs, with the state
 * smartpca.c:main has at the call (1642-1662): XTX = normalised GRM of the last pass (nrxtx = nrows), xmean/xfancy of that
 * pass, mmat filled by getcolxf.  Runs in a forked child (doshrinkp frees XTX, printevecs is once-only).
 * coords_o [k][nind]: the values printevecs leaves in xcoeffs (what the .evec file holds in shrinkmode). */
int refh_shrink (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, int nrows,
                 int fancy, int altnorm, const unsigned char *used, const double *xmean_in, const double *xfancy_in,
                 const double *XTXn, int k, int newshrink_in, double *coords_o)
{
  size_t na = sizeof (double) * (size_t) k * nind;
  unsigned char *sh = mmap (NULL, na, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (sh == MAP_FAILED) return -1;
  pid_t pid = fork ();
  if (pid < 0) return -2;
  if (pid == 0) {
    long i; int j; double *cc, *mmat, *xco = (double *) sh; int *xidx;
    fclose (stdout); stdout = fopen ("/dev/null", "w");
    hbuild (packed, nsnp, rl, nind);
    fancynorm = fancy; altnormstyle = altnorm; usepopsformissing = NO; regmode = YES; plotmode = NO; easymode = NO;
    shrinkmode = YES; newshrink = newshrink_in; toprightindex = -1; flip = NULL; verbose = NO;
    numeigs = k; numindivs = nind; indivmarkers = hindp; ofile = fopen ("/dev/null", "w");
    for (i = 0; i < nsnp; i++) hsnps[i].ignore = used[i] ? NO : YES;
    for (i = 0; i < nind; i++) hind[i].idnum = (int) i;
    ZALLOC (xmean, nsnp, double); ZALLOC (xfancy, nsnp, double);
    memcpy (xmean, xmean_in, sizeof (double) * nsnp); memcpy (xfancy, xfancy_in, sizeof (double) * nsnp);
    ZALLOC (xidx, nrows, int); memcpy (xidx, xindex_in, sizeof (int) * nrows);
    ZALLOC (XTX, (long) nrows * nrows, double); memcpy (XTX, XTXn, sizeof (double) * (size_t) nrows * nrows); nrxtx = nrows;
    ZALLOC (cc, nrows > 3 ? nrows : 3, double); ZALLOC (mmat, (long) nrows * nsnp, double);
    for (i = 0; i < nsnp; i++) {                                   /* smartpca.c:1644-1650 */
      getcolxf (cc, hsnpp[i], xidx, nrows, (int) i, NULL, NULL);
      for (j = 0; j < nrows; j++) mmat[(long) j * nsnp + i] = cc[j];
    }
    doshrinkp (mmat, nrows, (int) nsnp, xidx, hsnpp, xco);
    _exit (0);
  }
  int status = 0;
  waitpid (pid, &status, 0);
  if (!WIFEXITED (status) || WEXITSTATUS (status) != 0) { munmap (sh, na); return -3; }
  memcpy (coords_o, sh, na);
  munmap (sh, na);
  return 0;
}

/* dense path: the reference's domult_increment_normal (smartpca.c:3531-3561) over consecutive blocks of `blocksize`
 * columns, then symit2 (smartpca.c:480-508).  XTX_out: nrows*nrows (packed lower triangle while accumulating). */
int refh_dense_grm (const double *tblock_all, long ncols, int nrows, int blocksize, int nthreads, double *XTX_out)
{
  pthread_t threads[MAX_THREADS]; uint32_t thread_ct; long s; double *tb;
  thread_ct = nthreads < 1 ? 1 : nthreads; if (thread_ct > MAX_THREADS) thread_ct = MAX_THREADS;
  if (thread_ct > (uint32_t) nrows * 2) { thread_ct = nrows / 2; if (!thread_ct) thread_ct = 1; }
  triangle_fill (g_thread_start, nrows, thread_ct, 0, 1, 0, 1);
  vzero (XTX_out, ((long) nrows * (nrows + 1)) / 2);
  ZALLOC (tb, (long) blocksize * nrows, double);
  for (s = 0; s < ncols; s += blocksize) {
    int nb = (int) MIN ((long) blocksize, ncols - s);
    memcpy (tb, tblock_all + s * nrows, sizeof (double) * (size_t) nb * nrows);
    domult_increment_normal (threads, thread_ct, XTX_out, tb, nb, nrows);
  }
  symit2 (XTX_out, nrows);
  free (tb);
  return 0;
}

/* usepopsformissing: the reference's getcolxz (smartpca.c:3129-3216) for every SNP with the global switched on: normalised FP64
 * columns cols[nsnp][nrows] (population means filled in for missing genotypes), n0 / n1, the missing count after the fill, xmean, xfancy */
int refh_popfill_cols (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, const int *xtypes_in, int nrows, int npops,
                       int fancy, int altnorm, int *c0, int *c1, int *nmiss, double *xmean_o, double *xfancy_o, double *cols)
{
  long s; int *xidx, *xt, n0, n1, keep_usepops = usepopsformissing, keep_maxpops = maxpops;
  hbuild (packed, nsnp, rl, nind);
  fancynorm = fancy; altnormstyle = altnorm; usepopsformissing = YES; maxpops = npops;
  ZALLOC (xidx, nrows, int); ZALLOC (xt, nrows, int);
  memcpy (xidx, xindex_in, sizeof (int) * nrows); memcpy (xt, xtypes_in, sizeof (int) * nrows);
  for (s = 0; s < nsnp; s++) {
    nmiss[s] = getcolxz (cols + s * nrows, hsnpp[s], xidx, xt, nrows, (int) s, xmean_o, xfancy_o, &n0, &n1);
    c0[s] = n0; c1[s] = n1;
  }
  usepopsformissing = keep_usepops; maxpops = keep_maxpops;
  free (xidx); free (xt); hfree ();
  return 0;
}

/* the reference's fstcolyy (qpsubs.c:1205-1346) for every SNP: estn/estd [nsnp][numeg*numeg] */
int refh_fstcol (const unsigned char *packed, long nsnp, long rl, int nind, const int *xindex_in, const int *xtypes_in, int nrows, int numeg,
                 double *estn, double *estd)
{
  long s; int *xidx, *xt;
  hbuild (packed, nsnp, rl, nind);
  ZALLOC (xidx, nrows, int); ZALLOC (xt, nrows, int);
  memcpy (xidx, xindex_in, sizeof (int) * nrows); memcpy (xt, xtypes_in, sizeof (int) * nrows);
  setinbreed (NO);
  for (s = 0; s < nsnp; s++) fstcolyy (estn + s * numeg * numeg, estd + s * numeg * numeg, hsnpp[s], xidx, xt, nrows, numeg);
  free (xidx); free (xt); hfree ();
  return 0;
}
