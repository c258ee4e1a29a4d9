/* examples/smartpca_shim.c -- the calls a maintainer adds to EIGENSOFT's smartpca.c, as a free-standing C
 * translation unit.  It is compiled (gcc -std=c99 -Wall -Werror) and linked against libeigb200.so by tests/test_capi_cpu.py
 * to prove that include/eigb200.h is a plain-C header and that every symbol used here resolves; it only RUNS on a B200.
 *
 *   eb_shim_full_mode : replaces the outlier-iteration region smartpca.c:1077-1265 (GRM passes + eigvecs + ridoutlier)
 *                       and the .evec coordinate sequence smartpca.c:1440-1564
 *   eb_shim_threads   : the same on every GPU of the box, one host thread per GPU (eb_local_comm)
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "eigb200.h"

#define EB(call) do { if ((call) != 0) { fprintf (stderr, "(eigb200) %s\n", eb_last_error ()); exit (1); } } while (0)

/* packed: [nsnp][rlen] as packgenos lays it out (mcio.c:2825-2863); xindex: rows after loadindx (qpsubs.c:202) */
int eb_shim_full_mode (const uint8_t * packed, int64_t nsnp, int64_t rlen, int numindivs, int *xindex, int nrows,
                       int numeigs, int numoutlieriter, double *lambda, double *evecs, double *coords)
{
  eb_ctx *e = eb_create (-1);
  eb_pca_opts o;
  eb_pca_result res;
  uint8_t *used = (uint8_t *) malloc ((size_t) nsnp);
  double *xmean = (double *) malloc (sizeof (double) * (size_t) nsnp), *xfancy = (double *) malloc (sizeof (double) * (size_t) nsnp);
  int *rm_idx = (int *) malloc (sizeof (int) * (size_t) nrows), *rm_it = (int *) malloc (sizeof (int) * (size_t) nrows);
  int *rm_vec = (int *) malloc (sizeof (int) * (size_t) nrows), i;
  double *rm_score = (double *) malloc (sizeof (double) * (size_t) nrows);
  if (e == NULL) { fprintf (stderr, "(eigb200) %s\n", eb_last_error ()); return 1; }
  memset (&o, 0, sizeof (o));
  o.grm.fancynorm = 1; o.grm.altnormstyle = 1; o.grm.minallelecnt = 1; o.grm.maxmissing = 9999999;   /* smartpca defaults */
  o.numeigs = numeigs; o.numoutliter = numoutlieriter; o.numoutleigs = 10; o.outlthresh = 6.0; o.outliermode = 0;
  EB (eb_upload_packed (e, packed, nsnp, rlen, numindivs));
  EB (eb_pca_full (e, &o, xindex, nrows, lambda, evecs, used, xmean, xfancy, rm_idx, rm_it, rm_vec, rm_score, &res));
  for (i = 0; i < res.nremoved; i++)           /* smartpca.c:1258-1260 */
    printf ("REMOVED outlier individual %d iter %d evec %d sigmage %9.3f\n", rm_idx[i], rm_it[i], rm_vec[i], rm_score[i]);
  if (coords != NULL) EB (eb_evec_coords (e, evecs, numeigs, NULL, numindivs, coords, NULL, NULL));
  free (used); free (xmean); free (xfancy); free (rm_idx); free (rm_it); free (rm_vec); free (rm_score);
  eb_destroy (e);
  return 0;
}

struct shard_job {
  eb_local_comm *lc; int rank, world;
  const uint8_t *packed; int64_t nsnp, rlen; int numindivs, nrows, numeigs; const int *xindex;
  double *lambda, *evecs;
};

static void *shard_thread (void *arg)
{
  struct shard_job *j = (struct shard_job *) arg;
  eb_ctx *e = eb_create (j->rank);
  eb_comm cm;
  eb_pca_opts o;
  eb_pca_result res;
  int64_t base = j->nsnp / j->world, rem = j->nsnp % j->world;
  int64_t s0 = j->rank * base + (j->rank < rem ? j->rank : rem), s1 = s0 + base + (j->rank < rem ? 1 : 0);
  int *xi = (int *) malloc (sizeof (int) * (size_t) j->nrows);
  if (e == NULL) { fprintf (stderr, "(eigb200) %s\n", eb_last_error ()); exit (1); }
  memcpy (xi, j->xindex, sizeof (int) * (size_t) j->nrows);
  memset (&o, 0, sizeof (o));
  o.grm.fancynorm = 1; o.grm.altnormstyle = 1; o.grm.minallelecnt = 1; o.grm.maxmissing = 9999999;
  o.numeigs = j->numeigs; o.numoutliter = 5; o.numoutleigs = 10; o.outlthresh = 6.0;
  EB (eb_local_comm_get (j->lc, j->rank, &cm));
  EB (eb_set_comm (e, &cm));
  EB (eb_upload_packed (e, j->packed + s0 * j->rlen, s1 - s0, j->rlen, j->numindivs));
  /* collective: every thread receives the same lambda / evecs; rank 0's copy is the one the caller keeps */
  EB (eb_pca_full (e, &o, xi, j->nrows, j->rank == 0 ? j->lambda : (double *) malloc (sizeof (double) * (size_t) j->nrows),
                   j->rank == 0 ? j->evecs : NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, &res));
  EB (eb_set_comm (e, NULL));
  eb_destroy (e);
  free (xi);
  return NULL;
}

int eb_shim_threads (const uint8_t * packed, int64_t nsnp, int64_t rlen, int numindivs, const int *xindex, int nrows, int numeigs,
                     double *lambda, double *evecs)
{
  int world = eb_device_count (), r;
  pthread_t th[16];
  struct shard_job job[16];
  eb_local_comm *lc;
  if (world < 1) { fprintf (stderr, "(eigb200) no GPU\n"); return 1; }
  if (world > 16) world = 16;
  lc = eb_local_comm_create (world);
  if (lc == NULL) { fprintf (stderr, "(eigb200) %s\n", eb_last_error ()); return 1; }
  for (r = 0; r < world; r++) {
    struct shard_job j = { lc, r, world, packed, nsnp, rlen, numindivs, nrows, numeigs, xindex, lambda, evecs };
    job[r] = j;
    pthread_create (&th[r], NULL, shard_thread, &job[r]);
  }
  for (r = 0; r < world; r++) pthread_join (th[r], NULL);
  eb_local_comm_destroy (lc);
  return 0;
}
