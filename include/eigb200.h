/* eigb200.h -- C-ABI of libeigb200.so: the B200 (sm_100a) replacement for EIGENSOFT smartpca's hot path.
 *
 * Plain C, no C++/torch types.  All host buffers are caller-allocated and caller-freed; device state is
 * owned by an opaque eb_ctx (one per GPU).  Every function returns 0 on success and a negative code on
 * failure, with text in eb_last_error(); the patched smartpca.c turns non-zero into fatalx() to keep the
 * reference's fail-hard convention (strsubs.c:213).  Functions are called from smartpca's single main
 * thread and are not re-entrant per context.  There is no CPU fallback anywhere behind this header.
 *
 * Each entry point names the reference interface it replaces (file:line under DReichLab/EIG).
 */
#ifndef EIGB200_H
#define EIGB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct eb_ctx eb_ctx;

#define EB_OK 0
#define EB_ERR_CUDA (-1)
#define EB_ERR_ARG (-2)
#define EB_ERR_STATE (-3)
#define EB_ERR_NOMEM (-4)
#define EB_ERR_NUMERIC (-5)

/* -------- lifetime -------- */
const char *eb_last_error (void);
int eb_version (void);
/* device < 0 -> current CUDA device.  Fails (NULL) when no sm_100 GPU is visible. */
eb_ctx *eb_create (int device);
void eb_destroy (eb_ctx *);
int eb_device_count (void);
/* stream the library launches on (cudaStream_t as void*), for callers that time with CUDA events */
void *eb_stream (eb_ctx *);
int eb_sync (eb_ctx *);
/* counters since eb_create / last reset: number of kernel launches this library made */
int64_t eb_launch_count (eb_ctx *);
void eb_reset_launch_count (eb_ctx *);

/* -------- genotype store (replaces SNP.pbuff / packgenos, admutils.h:49, mcio.c:2825-2863) --------
 * Layout is the reference's: SNP-major, `rlen` bytes per SNP, 4 genotypes per byte MSB-first,
 * 0/1/2 = allele count, 3 = missing (admutils.c:718-735). */
int eb_upload_packed (eb_ctx *, const uint8_t * packed /* [nsnp][rlen] contiguous */ ,
                      int64_t nsnp, int64_t rlen, int numindivs);
/* same, from per-SNP row pointers (snp_pbuff[i] = xsnplist[i]->pbuff, qpsubs.c:222-237) */
int eb_upload_packed_rows (eb_ctx *, const uint8_t * const *snp_pbuff, int64_t nsnp, int64_t rlen, int numindivs);
/* adopt a slab that already lives in device memory (caller keeps ownership; pitch in bytes) */
int eb_adopt_packed_device (eb_ctx *, const void *dev_packed, int64_t nsnp, int64_t pitch, int numindivs);
/* fill a device slab with synthetic Hardy-Weinberg genotypes (eig_b200/synth.py documents the generator);
 * used by bench/tests so that 50k x 600k never has to be generated on the host.  SNPs [s0, s0+nsnp). */
int eb_synth_packed_device (eb_ctx *, void *dev_packed, int64_t nsnp, int64_t pitch, int numindivs,
                            uint64_t seed, int64_t s0, double missing, int npops, double delta);

/* PACKEDANCESTRYMAP genotype file -> device slab: inpack(), mcio.c:2769-2879.  The header record is checked with the
 * reference's own rules and messages (individual / SNP counts, and the two ID hashes when check_hash != 0: hasharr,
 * admutils.c:651-682, over the .ind / .snp IDs in file order); the payload streams through pinned staging buffers
 * (disk read of chunk i+1 overlaps the H2D copy of chunk i); male X heterozygotes become missing on the device
 * (checkxval, mcio.c:1606-1618) when snp_is_x[nsnp] / indiv_is_male[numindivs] are given (NULL = no X rule). */
int eb_hash_ids (const char *const *ids, int n);
int eb_packed_file_header (const char *path, int *nind, int *nsnp, int *ihash, int *shash, int64_t * rlen, int64_t * file_bytes);
int eb_upload_packed_file (eb_ctx *, const char *path, int numindivs, int64_t nsnp, int check_hash, int ihash, int shash,
                           const uint8_t * snp_is_x, const uint8_t * indiv_is_male);
int eb_download_packed (eb_ctx *, uint8_t * out /* [nsnp][rlen] */ );

/* output files with the reference's formats: .eval "%12.6f\n" per eigenvalue (smartpca.c:1425-1431); .evec header
 * "%20s " "#eigvals:" + "%9.3f " per eigenvalue, then per individual "%20s " ID, "%10.4f  " (hiprec: "%12.6f  ") per
 * coordinate, "%15s\n" population (smartpca.c:1433-1437, 1574-1591); text GRM "a b nsnp %0.6f" scaled to mean
 * diagonal 1 (dumpgrm, smartpca.c:3770-3805).  coords [numeigs][nout]. */
int eb_write_eval (const char *path, const double *lambda, int n);
int eb_write_evec (const char *path, const double *lambda, int numeigs, const char *const *ids, const char *const *groups,
                   const double *coords, int nout, int hiprec);
int eb_write_grm (const char *path, const double *XTX, int nrows, int numsnps);
/* grmbinary: YES (dumpgrmbin, smartpca.c:3704-3766): <prefix>.N.bin = numsnps as int32 per lower-triangle entry, <prefix>.bin =
 * the entries scaled to mean diagonal 1 as float32 (GCTA layout) */
int eb_write_grm_bin (const char *prefix, const double *XTX, int nrows, int numsnps);

/* rows used = xindex[0..nrows) ascending into 0..numindivs-1 (loadindx, qpsubs.c:202-219).
 * xindex == NULL selects all individuals.  Re-gathers the working matrix on the device. */
int eb_set_rows (eb_ctx *, const int *xindex, int nrows);

/* -------- integer reductions (bit-exact) -------- */
/* per SNP over the current rows: c0 = sum g, c1 = sum (2-g), nmiss (getcolxz_binary1, smartpca.c:3261-3276;
 * numvalidgtx, admutils.c:1136: nvalid = nrows - nmiss) */
int eb_snp_counts (eb_ctx *, int *c0, int *c1, int *nmiss);
/* per individual (all numindivs) non-missing count over SNPs with snp_keep[s] != 0 (NULL = all):
 * numvalidgtallind, admutils.c:1075-1097 */
int eb_indiv_valid_counts (eb_ctx *, const uint8_t * snp_keep, int *nvalid);

/* per population k (xtypes[row] in 0..npops-1, other values skipped; xtypes indexed like the current rows) and SNP s:
 * counts[(s*npops + k)*3 + g] = members with genotype g = 0,1,2.  This is the gather + count loop of fstcolyy
 * (qpsubs.c:1205-1281: ddd[k] = {n1 + 2 n2, n1 + 2 n0}; inbreed mode uses the classes) that dofstnumx
 * (qpsubs.c:2488-2736) runs for every SNP, and of the per-population validity passes (smartpca.c:844,870). */
int eb_pop_counts (eb_ctx *, const int *xtypes, int npops, int *counts);

/* -------- GRM accumulation: the region smartpca.c:1088-1236 -------- */
typedef struct {
  int fancynorm;                /* usenorm:      smartpca.c:2130 ; globals.h:5 default YES */
  int altnormstyle;             /* altnormstyle: smartpca.c:2136 ; default YES */
  int minallelecnt;             /* smartpca.c:2164 ; default 1 */
  int maxmissing;               /* smartpca.c:2167 ; default 9999999 */
  const uint8_t *snp_ignore;    /* [nsnp] or NULL: SNPs already flagged ignore (skipped like loadsnpx does) */
  const double *snp_weight;     /* [nsnp] or NULL: weightname (smartpca.c:1178-1180) */
} eb_grm_opts;

/* Per-SNP outputs (any pointer may be NULL), indexed like the uploaded SNPs:
 *   c0,c1   n0,n1 of getcolxz_binary1 (-1,-1 if all missing);  nmiss (-1 if all missing)
 *   used    1 if the SNP entered XTX, 0 if dropped by smartpca.c:1131-1144 or ignored
 *   xmean   ymean*yfancy, xfancy yfancy (smartpca.c:3291-3295)   -- bit-exact vs the reference
 * y_out     trace(XTX)/(nrows-1) (smartpca.c:1230).  nused_out = number of SNPs used.
 * XTX_host  NULL, or nrows*nrows doubles receiving XTX/y (the matrix the reference holds after line 1236).
 * The GRM stays resident on the device for eb_eig(). */
int eb_grm (eb_ctx *, const eb_grm_opts * opts, int *c0, int *c1, int *nmiss, uint8_t * used,
            double *xmean, double *xfancy, double *y_out, int64_t * nused_out, double *XTX_host);

/* usepopsformissing: YES (smartpca.c:2150): getcolxz (smartpca.c:3129-3216) replaces a missing genotype by the mean of the
 * individual's population at that SNP (when the population has data) before fvadjust (2236-2279), so mean and scale are taken over
 * observed and filled values and the columns are arbitrary FP64 numbers -- the reference's dense path (smartpca.c:995-1014).  All of
 * it runs on the device; the columns never exist on the host.  xtypes[nrows]: population of every current row, 0..npops-1 (other
 * values: no population, stays missing).  Outputs as eb_grm, except nmiss = genotypes still missing AFTER the fill (getcolxz's
 * return value) and that xmean / xfancy agree with the reference to rounding, not bit for bit (the reference adds the filled
 * values row by row, the kernel population by population).  The GRM stays resident for eb_eig; the packed-table passes
 * (eb_project, eb_lsqproj, eb_evec_coords, eb_shrink_coords) refuse to run on it.  Not collective. */
int eb_grm_popfill (eb_ctx *, const eb_grm_opts * opts, const int *xtypes, int npops, int *c0, int *c1, int *nmiss, uint8_t * used,
                    double *xmean, double *xfancy, double *y_out, int64_t * nused_out, double *XTX_host);

/* dense path: getcolxz + domult_increment_normal + block_increment_normal (smartpca.c:3129-3216, 3531-3561, 3498-3528),
 * which the reference takes when usepopsformissing / ldregress make the normalised columns arbitrary FP64 values
 * (smartpca.c:995-1014).  The host keeps producing tblock (nblock rows of nrows doubles, smartpca.c:1198-1218); each call
 * adds sum_s x_s x_s^T to the resident accumulator; _end mirrors (symit2), takes the trace and leaves the GRM resident
 * for eb_eig exactly like eb_grm. */
int eb_grm_dense_begin (eb_ctx *, int nrows);
int eb_grm_dense_add (eb_ctx *, const double *tblock /* [nblock][nrows] */ , int nblock);
int eb_grm_dense_end (eb_ctx *, double *y_out, double *XTX_host /* or NULL */ );

/* multi-GPU (one context per SNP shard): device pointer / leading dimension of the UNNORMALISED partial
 * XTX left by eb_grm_partial, so that the caller's NCCL reduce can run on it in place, and the call that
 * finishes the pass (trace, y) after the reduce. */
int eb_grm_partial (eb_ctx *, const eb_grm_opts * opts, int *c0, int *c1, int *nmiss, uint8_t * used,
                    double *xmean, double *xfancy, int64_t * nused_out);
void *eb_grm_device_ptr (eb_ctx *, int64_t * ld_out, int64_t * n_out);
int eb_grm_finish (eb_ctx *, double *y_out, double *XTX_host);

/* -------- multi-GPU: one context per SNP shard, exchange steps as kernels over peer memory --------
 * The launcher supplies host-side plumbing only: an all-gather of small host records (every rank contributes `bytes`
 * bytes, dst receives world*bytes in rank order) and a barrier; both return 0 on success.  Device buffers are shared
 * through CUDA IPC (directly when ranks are contexts of one process) and reduced by the library's own kernels over
 * NVLink: no device pointer ever crosses this interface.  With a communicator set, eb_grm / eb_pca_full / eb_fpca
 * are COLLECTIVE: every rank calls them with the same rows and options on its own SNP shard; the reduced GRM (hence
 * y, lambda, evecs, outlier decisions) is bit-identical on every rank, per-SNP outputs stay per shard, and nused_out is
 * the total over shards.  This replaces nothing in the reference (smartpca is single-node pthreads,
 * smartpca.c:3331-3358); it is what "numthreads" becomes across the GPUs of one box. */
typedef struct {
  int rank, world;
  int (*allgather_host) (void *user, const void *src, void *dst, int64_t bytes);
  int (*barrier) (void *user);
  void *user;
} eb_comm;
int eb_set_comm (eb_ctx *, const eb_comm * comm /* NULL or world <= 1: single GPU */ );
/* In-process communicator: `world` host threads of one process, one context (GPU) each -- what a C caller such as
 * smartpca.c needs to drive every GPU of the box without MPI or torch.  eb_local_comm_get fills the eb_comm of one rank
 * (plain C callbacks over a pthread barrier); each thread passes its own to eb_set_comm and then makes the collective calls.
 * The handle must outlive the contexts that use it. */
typedef struct eb_local_comm eb_local_comm;
eb_local_comm *eb_local_comm_create (int world);
int eb_local_comm_get (eb_local_comm *, int rank, eb_comm * out);
void eb_local_comm_destroy (eb_local_comm *);
/* testing aid: in-place sum over ranks of a host vector through the peer all-reduce kernel (count even) */
int eb_peer_allreduce_test (eb_ctx *, double *host_io, int64_t count);
/* SNPs of THIS context's shard that entered XTX in the last GRM pass (eb_grm's nused_out is the total over shards) */
int64_t eb_snp_used_count (eb_ctx *);

/* -------- symmetric eigensolver on the resident GRM: eigvecs(), eigsubs.c:39-55 / dspev_, eigx.c:107 --------
 * lambda[nrows] descending (all eigenvalues of XTX/y); evecs[nvec*nrows], row i = unit eigenvector i.
 * lambda may be NULL (leading vectors only: what the outlier iterations need, smartpca.c:1250) and nvec may be 0
 * (spectrum only).  Sign: every returned vector has its entry of largest magnitude positive (LAPACK's sign is arbitrary; this
 * convention is deterministic across solver paths and GPUs and is the one POPGEN/example.evec happens to carry).
 * With a communicator set and the GRM of a sharded eb_grm resident the call is COLLECTIVE: from n = "dist_min" the dense -> band
 * reduction runs row-distributed (owned 128-row tiles of the products, panel / product block summed over peer memory) and the
 * subspace iteration splits its block mat-vecs by row tiles; results are bit-identical on every rank. */
int eb_eig (eb_ctx *, int nvec, double *lambda, double *evecs);
/* standalone drop-in with the reference's contract (mat row-major n*n, preserved): include/eigsubs.h:6-7 */
int eb_eigvecs (eb_ctx *, const double *mat, double *evals, double *evecs, int n, int nvec);

/* drop-in symbols under the reference's own names (include/eigsubs.h:6-7): same contract as eigsubs.c:21,39 (mat
 * preserved, eigenvalues descending, row i of evecs = vector i, fatal on failure).  eigvecs fills all n vectors at any n
 * (callers that only need the leading ones should use eb_eigvecs with nvec, which is much cheaper at large n). */
void eigvecs (double *mat, double *evals, double *evecs, int n);
void eigvals (double *mat, double *evals, int n);

/* eigensolver selection: key "eig_method" = 0 auto | 1 one-stage | 2 two-stage + subspace iteration;
 * "two_stage_min" = n at which auto switches to 2; "dist_min" = n from which a COLLECTIVE eb_eig (communicator set, GRM from a
 * sharded eb_grm) splits the band reduction and the subspace iteration over the ranks (default 8192);
 * "eig_vectors" = 0 (default): a two-stage solve that returns spectrum AND vectors back-transforms the eigenvectors of the
 * tridiagonal matrix through the kept reflectors of both stages | 1: subspace iteration on the original matrix.
 * GRM kernel selection (eb_grm, eb_pca_full; domult_increment_lookup, smartpca.c:3426-3495):
 * "grm_method" = 0 auto (integer path from "i8_min" rows, default 4096) | 1 FP64 DMMA (grm_syrk_kernel) | 2 exact integer tensor
 * cores (grm_i8_pair_kernel: tcgen05.mma kind::i8, s32 accumulators in TMEM, FP64 weights as 7-bit digits, FP64 accumulation);
 * "i8_slices" = 0 auto (>= 52 bits below the typical per-SNP weight: 8 digits on ordinary data) | 1..9 digits;
 * "i8_slab" = cap on the SNP rows per operand slab (0: what memory allows); "i8_pair" = 1 CTA pairs (default) | 0 single CTAs;
 * "i8_sync" = passes a cluster may run ahead of the slowest (default 0, -1 = free running).  All ranks of a communicator must use the same "grm_method".
 * Packed x skinny products (fastmode, loadings / projections, lsqproj, shrinkmode's mat-vecs; kjg_fpca.c:104-178):
 * "pg_method" = 0 auto (integer tensor cores from "pg_i8_min" rows and SNPs, default 2048) | 1 FP64 DMMA (packed_gemm_kernel) |
 * 2 integer tensor cores with in-kernel decode (pg_i8_kernel).
 * Returns EB_ERR_ARG for an unknown key. */
int eb_set_option (eb_ctx *, const char *key, int value);

/* testing aid: two-stage tridiagonalisation alone.  d[n], e[n] of the similar tridiagonal (unscaled); band (may be
 * NULL) receives the intermediate band matrix, [n][128] with band[col][k] = B[col+k][col], k <= 64. */
int eb_debug_tridiag (eb_ctx *, const double *mat, int n, double *d, double *e, double *band);

/* -------- outlier detection: ridoutlier(), smartsubs.c:18-93 (decisions bit-exact) -------- */
int eb_ridoutlier (const double *evecs, int n, int neigs, double thresh, int outliermode,
                   int *badlist, int *vecno, double *score);

/* -------- whole full-mode pass with outlier iterations: smartpca.c:1077-1265 -------- */
typedef struct {
  eb_grm_opts grm;
  int numeigs;                  /* numoutevec, smartpca.c:2118 */
  int numoutliter;              /* numoutlieriter, default 5 */
  int numoutleigs;              /* numoutlierevec, default 10 */
  double outlthresh;            /* outliersigmathresh, default 6.0 */
  int outliermode;              /* default 0 */
} eb_pca_opts;
typedef struct {
  int nrows_final;              /* rows left after outlier removal */
  int niter;                    /* GRM+eig passes executed */
  int nremoved;                 /* outliers removed in total */
  int64_t nused;                /* SNPs used in the last pass */
  double y;                     /* trace/(nrows-1) of the last pass */
  double secs_grm, secs_eig, secs_total;
} eb_pca_result;
/* xindex_io: in = initial rows, out = surviving rows (nrows_final).  removed_*: per removal (capacity nrows):
 * original individual index, iteration (1-based), eigenvector number, z-score.
 * lambda and evecs are written on EVERY pass with the row count of that pass, so they must hold nrows (the INITIAL count) and
 * numeigs * nrows doubles; on return the first nrows_final / numeigs * nrows_final entries are the result, packed
 * [numeigs][nrows_final].  snp_used/xmean/xfancy as in eb_grm for the last pass. */
int eb_pca_full (eb_ctx *, const eb_pca_opts * opts, int *xindex_io, int nrows,
                 double *lambda, double *evecs, uint8_t * snp_used, double *xmean, double *xfancy,
                 int *removed_index, int *removed_iter, int *removed_vecno, double *removed_score,
                 eb_pca_result * res);

/* -------- fastmode: kjg_fpca(), kjg_fpca.c:24 after setgval(), gval.c:31 --------
 * eval[K], evec[n*K] row-major as kjg_fpca leaves it (smartpca.c:971 transposes afterwards).
 * The seeded Gaussian start matrix reproduces kjg_gsl.c:96-186 bit for bit. */
int eb_fpca (eb_ctx *, int fancynorm, int altnormstyle, size_t K, size_t L, size_t I, long seed,
             double *eval, double *evec);
void eb_gauss_matrix (long seed, size_t n, size_t L, double *out);
/* fastmode drop-in under the reference's own name (include/kjg_fpca.h:22; kjg_fpca.c:24): same arguments, same outputs, exit(1)
 * on K >= L or I == 0 (kjg_fpca.c:26-29).  The reference hands the data over through gval.c's file statics (setgval, gval.c:31-87);
 * here the hand-over is eb_setgval_packed with plain pointers (snp_pbuff[i] = xsnplist[i]->pbuff, the PCA rows xindex, the
 * globals fancynorm / altnormstyle / seed of smartpca.c), called by the setgval replacement a maintainer compiles against the
 * reference headers (integration/eb_gval.c).  mono_out[ncols] (may be NULL) = 1 where setgval's side effect sets
 * cupt->ignore (min(n0, n1) == 0, gval.c:80-82).  Both run on the process-wide context of the eigvecs() drop-in. */
int eb_setgval_packed (const uint8_t * const *snp_pbuff, int64_t ncols, int64_t rlen, int numindivs, const int *xindex, int nrows,
                       int fancynorm, int altnormstyle, long seed, uint8_t * mono_out);
void eb_unsetgval (void);
void kjg_fpca (size_t K, size_t L, size_t I, double *eval, double *evec);

/* -------- next rows: SNP loadings / sample projections, smartpca.c:1485-1525 -------- */
int eb_project (eb_ctx *, const double *evecs, int numeigs, double *ffvecs /* [numeigs][nsnp] */ ,
                double *fxvecs /* [numeigs][nrows] */ , double *fxscal /* [numeigs] */ );

/* lsqproj(), smartpca.c:4606-4757 (regressit, regsubs.c:8-77; solvit/choldc/cholsl, nicksrc/linsubs.c:296-393):
 * least-squares coordinates of every listed individual on the SNP loadings, using only its observed genotypes.
 * indiv[nindiv] ascending indices into 0..numindivs-1 (NULL = 0..nindiv-1): the reference walks all non-ignored
 * individuals, PCA rows or not (projected populations).  acoeffs/bcoeffs [numeigs][nindiv] (smartpca.c:4713,4745);
 * nvalid = rows of the individual's regression; ok = 0 where the reference ignores the individual
 * ("insufficient data", nvalid <= numeigs) -- its coefficients are 0.  Uses xmean/xfancy/used of the last eb_grm.
 * Limit: numeigs <= 32 for eb_lsqproj, eb_evec_coords and eb_shrink_coords (EB_ERR_ARG beyond; the reference has no limit). */
int eb_lsqproj (eb_ctx *, const int *indiv, int nindiv, const double *ffvecs /* [numeigs][nsnp] */ , const double *fxscal,
                int numeigs, double *acoeffs, double *bcoeffs, int *nvalid, uint8_t * ok);
/* the .evec values: the whole sequence smartpca.c:1440-1564 (setfvecs, loadings, projections, lsqproj, seteigscale,
 * acoeffs * eigscale).  coords [numeigs][nindiv]; every PCA row must be in the list. */
int eb_evec_coords (eb_ctx *, const double *evecs /* [numeigs][nrows] */ , int numeigs, const int *indiv, int nindiv,
                    double *coords, double *eigscale /* [numeigs] or NULL */ , uint8_t * ok /* [nindiv] or NULL */ );

/* shrinkmode: doshrinkp (smartpca.c:4223-4419) or, with newshrink != 0, doshrinkp2 (4022-4220): leave-one-out
 * ("shrunk") coordinates of every PCA sample, least-squares coordinates (doproj, 3986-4019) of everybody else.
 * Needs the GRM of the last eb_grm resident.  coords [numeigs][numindivs] are the values printevecs writes in
 * shrinkmode (3849-3866: x10, unit length per eigenvector, one row per individual of the store); lambda_out[numeigs]
 * are the eigenvalues of the .evec header; ok[numindivs] (may be NULL) is 0 where a regression was singular.
 * Collective on a sharded context.  Cost: one full eigendecomposition + 2 numeigs nrows^2 nsnp tensor-core flops. */
int eb_shrink_coords (eb_ctx *, int numeigs, int newshrink, double *coords, double *lambda_out, uint8_t * ok);
/* testing aid: C[M][N] = op(A) op(B)^T on the FP64 tensor cores; a_km: A given as [K][M] else [M][K]; b_kn: B as [K][N]
 * else [N][K] (all dense row-major host arrays) */
int eb_debug_gemm (eb_ctx *, int a_km, int b_kn, const double *A, const double *B, double *C, int M, int N, int K);

/* Tracy-Widom statistics of the spectrum: the loop smartpca.c:1336-1366 (twstats.c:58-77) around dotwcalc
 * (statsubs.c:1680-1725) / twnorm (1655-1677).  lambda[m] descending with m = eb_numgtz(lambda, nrows) (statsubs.c:1727).
 * znval > 0: fixed effective number of markers (smartpca passes MAX(nrows, znval)); <= 0: estimated per eigenvalue.
 * tw[i] / zn[i]: the "twstat" / "effect. n" columns (-1 where the reference prints NA: fewer than minm eigenvalues left).
 * The p-value stays with the caller's twtail() table lookup (statsubs.c:1590, POPGEN/twtable).  Host arithmetic, O(m) by
 * suffix sums where the reference re-sums the tail for every eigenvalue (O(m^2)). */
int eb_numgtz (const double *lambda, int n);
int eb_tw_stats (const double *lambda, int m, double znval, int minm, double *tw, double *zn);
/* the "p-value" column: twtail -> gettw (statsubs.c:1590, 1811-1860) on the caller's table (POPGEN/twtable rows: x, right
 * tail, density; n rows, x ascending), cubic interpolation inside the table and the reference's formulas outside it */
double eb_tw_tail (double twstat, const double *tab_x, const double *tab_tail, const double *tab_pdf, int n);

/* -------- measurement helpers -------- */
/* last pass timings measured with CUDA events on the library's stream (milliseconds) */
typedef struct {
  float gather_ms, stats_ms, grm_ms, finalize_ms, tridiag_ms, bisect_ms, vectors_ms;
  int grm_launches; int nsplit;
  int eig_method;               /* 1 = one-stage Householder, 2 = two-stage (band) + subspace iteration */
  int chfsi_iters, chfsi_matvecs;
  /* self-measurement of the last grm_syrk_kernel launch (every CTA records %smid, %globaltimer and clock64):
   * effective SM clock = median over CTAs of cycles / wall time; distinct SMs the CTAs ran on; shortest / longest CTA;
   * first CTA start -> last CTA end.  Explains a slow launch without a profiler: clock, SM count or a late CTA. */
  float grm_sm_mhz, grm_cta_min_ms, grm_cta_max_ms, grm_span_ms;
  int grm_sms;
  int chfsi_converged;          /* 1: every requested pair reached the strict residual tolerance; 0: accepted at the relaxed one */
  float chfsi_resid;            /* largest relative residual |A v - theta v| / |A| among the returned pairs */
  float exchange_wait_ms;       /* multi-GPU: time the stream spent waiting for the slowest rank's tiles before the reduce */
  float band_ms, chase_ms;      /* two-stage tridiagonalisation: dense -> band (DMMA products + panel QR), band -> tridiagonal */
  int grm_method;               /* GRM of the last pass: 1 = FP64 DMMA (grm_syrk_kernel), 2 = exact integer tensor cores (grm_i8_kernel) */
  int i8_slices, i8_segments;   /* integer path: 7-bit weight digits; 1 = complete data (one basis), 3 = missing genotypes present */
  int i8_flag_blocks;           /* integer path: 128-SNP blocks that contain a missing genotype (they take the two extra bases) */
  float i8_tera_ops;            /* integer path: 1e12 8-bit multiply-adds x 2 issued by the last pass (all digits and bases) */
  int i8_slab_rows;             /* integer path: SNP rows per GEMM launch (operand slab) of the last pass */
  float i8_gemm_ms;             /* integer path: summed device time of the grm_i8 GEMM launches of the last pass (CUDA events) */
} eb_timings;
int eb_get_timings (eb_ctx *, eb_timings * t);
/* FP64 DMMA / DFMA issue-rate microbenchmarks (TFLOP/s) used as roofline cross-checks */
int eb_microbench_fp64 (eb_ctx *, double *dmma_tflops, double *dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif
