#!/usr/bin/env python
"""Regenerate integration/smartpca_b200.patch: the edits a maintainer makes to src/eigensrc/smartpca.c (EIGENSOFT 8.0.0,
smartpca v18140) so that its hot path runs in libeigb200.so.  Run where the reference checkout is available:

    python integration/make_patch.py [/root/reference]

The script applies a handful of anchored insertions to a scratch copy of the reference file and writes `diff -U2` of the two;
the repository keeps only the patch (no reference source).  Call sites (reference line numbers):
  936-950   create the context, upload xsnplist[i]->pbuff once          (eb_create, eb_upload_packed_rows)
  1116-1239 per pass: rows, per-SNP counts + drop rule + GRM + eigen     (eb_set_rows, eb_grm, eb_eig); the reference's own
            loop keeps the log lines (" snp ... ignored", logdeletedsnp, "total number of snps killed in pass") in SNP order
  1297      XTX for dumpgrm / dotpops / printxcorr                       (eb_grm_finish)
  1485-1525 SNP loadings, sample projections, fxscal                     (eb_project)
  1553      lsqproj                                                      (eb_lsqproj)
  1642-1658, 2030-2045 shrinkmode                                        (eb_shrink_coords)
fastmode (957-963) needs no source edit: setgval comes from integration/eb_gval.c and kjg_fpca from the library.
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "src", "eigensrc", "smartpca.c")

HELPERS = r'''
/* ---- libeigb200: the GRM / eigen / projection hot path on the GPU (INTEGRATION.md) ---- */
static eb_ctx *ebctx = NULL;
/* the passes that decode the packed matrix with the GRM's own tables (loadings, projections, lsqproj, shrinkmode) run on the GPU
   unless usepopsformissing made the columns population-dependent: those stay with the reference's host code */
#define EB_PACKED (ebctx != NULL && !usepopsformissing)
static SNP **eb_snps0 = NULL;   /* the SNP list as uploaded (the initial xsnplist) */
static int eb_ncols0 = 0;
static double *eb_ffvecs0 = NULL;       /* SNP loadings indexed like the uploaded list (for eb_lsqproj) */
extern long rlen;               /* packit.h: bytes per SNP of the packed store */

static void
eb_fail (const char *what)
{
  fatalx ("(libeigb200) %s: %s\n", what, eb_last_error ());
}

/* position of every SNP of the current xsnplist inside the uploaded list (both are in snpmarkers order) */
static void
eb_colmap (int *map, SNP ** xsnplist, int ncols)
{
  int i, p = 0;
  for (i = 0; i < ncols; ++i) {
    while ((p < eb_ncols0) && (eb_snps0[p] != xsnplist[i]))
      ++p;
    if (p == eb_ncols0)
      fatalx ("(libeigb200) snp %s is not part of the uploaded list\n", xsnplist[i]->ID);
    map[i] = p++;
  }
}

/* lsqproj (smartpca.c:4606) for every non-ignored individual: same outputs, same "insufficient data" handling */
static void
eb_lsqproj_all (Indiv ** indm, int nind, int neigs, double *fxscal, double *acoeffs, double *bcoeffs)
{
  int *list, *nvalid, nl = 0, q, i, j;
  double *a, *b;
  unsigned char *ok;
  ZALLOC (list, nind, int);
  for (i = 0; i < nind; ++i)
    if (!indm[i]->ignore)
      list[nl++] = i;
  ZALLOC (a, neigs * nl + 1, double);
  ZALLOC (b, neigs * nl + 1, double);
  ZALLOC (nvalid, nl + 1, int);
  ZALLOC (ok, nl + 1, unsigned char);
  printf ("lsqproj called!\n");
  if (nl > 0 && eb_lsqproj (ebctx, list, nl, eb_ffvecs0, fxscal, neigs, a, b, nvalid, ok) != 0)
    eb_fail ("eb_lsqproj");
  for (q = 0; q < nl; ++q) {
    i = list[q];
    if (!ok[q]) {
      indm[i]->ignore = YES;
      printf ("%s ignored (insufficient data\n", indm[i]->ID);
      continue;
    }
    for (j = 0; j < neigs; ++j) {
      acoeffs[j * nind + i] = a[j * nl + q];
      bcoeffs[j * nind + i] = b[j * nl + q];
    }
  }
  free (list);
  free (a);
  free (b);
  free (nvalid);
  free (ok);
}

/* doshrinkp / doshrinkp2 (smartpca.c:4223 / 4022) including the .evec they print through printevecs (3837-3871) */
static void
eb_doshrink (double *xcoeffs)
{
  double *co, *lam, y;
  unsigned char *ok;
  int i, j;
  Indiv *indx;
  printf ("doshrink called\n");
  fflush (stdout);
  ZALLOC (co, numeigs * numindivs, double);
  ZALLOC (lam, numeigs, double);
  ZALLOC (ok, numindivs, unsigned char);
  if (eb_shrink_coords (ebctx, numeigs, newshrink, co, lam, ok) != 0)
    eb_fail ("eb_shrink_coords");
  fprintf (ofile, "%20s ", "#eigvals:");
  for (j = 0; j < numeigs; j++)
    fprintf (ofile, "%9.3f ", lam[j]);
  fprintf (ofile, "\n");
  printf ("writing eigenvecotrs for  %d samples\n", numindivs);
  for (i = 0; i < numindivs; i++) {
    indx = indivmarkers[i];
    fprintf (ofile, "%20s ", indx->ID);
    for (j = 0; j < numeigs; j++) {
      y = co[j * numindivs + i];
      if (indx->flag == 7777)
        y *= edgarw[j];
      if (hiprec)
        fprintf (ofile, "%12.6f  ", y);
      else
        fprintf (ofile, "%10.4f  ", y);
      if (xcoeffs != NULL)
        xcoeffs[j * numindivs + i] = y;
    }
    fprintf (ofile, "%15s\n", indx->egroup);
  }
  fflush (ofile);
  free (co);
  free (lam);
  free (ok);
  printf ("doshrink exited\n");
}
'''

UPLOAD = r'''
  if ((ldregress == 0) && (!fastmode || numeigs == 0) && (!fstonly) && packmode && (getenv ("EIGB200_OFF") == NULL)) {
    /* the GRM (lookup path, or the dense path of usepopsformissing, smartpca.c:995-1014) runs on the GPU: upload the packed
       genotypes of the used SNPs once */
    const uint8_t **eb_rows;
    ebctx = eb_create (-1);
    if (ebctx == NULL)
      eb_fail ("eb_create");
    eb_ncols0 = ncols;
    ZALLOC (eb_snps0, ncols + 1, SNP *);
    ZALLOC (eb_rows, ncols + 1, const uint8_t *);
    for (i = 0; i < ncols; ++i) {
      eb_snps0[i] = xsnplist[i];
      eb_rows[i] = (const uint8_t *) xsnplist[i]->pbuff;
    }
    if (eb_upload_packed_rows (ebctx, eb_rows, ncols, rlen, numindivs) != 0)
      eb_fail ("eb_upload_packed_rows");
    free (eb_rows);
    printf ("libeigb200: %d snps x %d individuals resident on the GPU\n", ncols, numindivs);
  }
'''

PASS = r'''
    if (ebctx) {
      /* one pass of smartpca.c:1116-1239 on the GPU; the loop below only replays the per-SNP decisions for the log */
      int *ec0, *ec1, *enm, *emap, ep, env;
      unsigned char *eused, *eign;
      double *exm, *exf, *ewt = NULL;
      int64_t enused;
      eb_grm_opts eo;
      ZALLOC (ec0, eb_ncols0, int);
      ZALLOC (ec1, eb_ncols0, int);
      ZALLOC (enm, eb_ncols0, int);
      ZALLOC (emap, ncols + 1, int);
      ZALLOC (eused, eb_ncols0, unsigned char);
      ZALLOC (eign, eb_ncols0, unsigned char);
      ZALLOC (exm, eb_ncols0, double);
      ZALLOC (exf, eb_ncols0, double);
      eb_colmap (emap, xsnplist, ncols);
      for (i = 0; i < eb_ncols0; ++i)
        eign[i] = 1;
      for (i = 0; i < ncols; ++i)
        eign[emap[i]] = 0;
      if (weightmode) {
        ZALLOC (ewt, eb_ncols0, double);
        for (i = 0; i < eb_ncols0; ++i)
          ewt[i] = eb_snps0[i]->weight;
      }
      eo.fancynorm = fancynorm;
      eo.altnormstyle = altnormstyle;
      eo.minallelecnt = minallelecnt;
      eo.maxmissing = maxmissing;
      eo.snp_ignore = eign;
      eo.snp_weight = ewt;
      if (eb_set_rows (ebctx, xindex, nrows) != 0)
        eb_fail ("eb_set_rows");
      if (usepopsformissing) {
        if (eb_grm_popfill (ebctx, &eo, xtypes, numeg, ec0, ec1, enm, eused, exm, exf, &y, &enused, NULL) != 0)
          eb_fail ("eb_grm_popfill");
      }
      else if (eb_grm (ebctx, &eo, ec0, ec1, enm, eused, exm, exf, &y, &enused, NULL) != 0)
        eb_fail ("eb_grm");
      for (i = 0; i < ncols; i++) {
        cupt = xsnplist[i];
        ep = emap[i];
        n0 = ec0[ep];
        n1 = ec1[ep];
        tt = enm[ep];
        xmean[i] = exm[ep];
        xfancy[i] = exf[ep];
        t = MIN (n0, n1);
        if ((t < minallelecnt) || (tt > maxmissing) || (tt < 0) || (t == 0)) {
          t = MAX (t, 0);
          tt = MAX (tt, 0);
          cupt->ignore = YES;
          logdeletedsnp (cupt->ID, "minallelecnt", deletesnpoutname);
          if (nkill < 10)
            printf (" snp %20s ignored . allelecnt: %5d  missing: %5d\n", cupt->ID, t, tt);
          ++nkill;
          continue;
        }
        ++ynumsnps;
        ++nused;
      }
      if ((int64_t) nused != enused)
        fatalx ("(libeigb200) used-SNP count mismatch: %d vs %ld\n", nused, (long) enused);
      printf ("total number of snps killed in pass: %d  used: %d\n", nkill, nused);
      env = MAX (numeigs, numoutleigs);
      env = MIN (env, nrows);
      if (eb_eig (ebctx, env, lambda, evecs) != 0)
        eb_fail ("eb_eig");
      free (ec0);
      free (ec1);
      free (enm);
      free (emap);
      free (eused);
      free (eign);
      free (exm);
      free (exf);
      if (ewt != NULL)
        free (ewt);
      goto eb_after_eig;
    }
'''

FETCH_XTX = r'''
  if (ebctx) {
    /* host consumers of the normalised matrix (dumpgrm, dotpops, printxcorr) read XTX: fetch it once */
    if (eb_grm_finish (ebctx, NULL, XTX) != 0)
      eb_fail ("eb_grm_finish");
  }
'''

PROJECT = r'''
    if (EB_PACKED) {
      /* SNP loadings, sample projections and fxscal (smartpca.c:1485-1525) in one call */
      int *emap;
      ZALLOC (eb_ffvecs0, numeigs * eb_ncols0 + 1, double);
      ZALLOC (emap, ncols + 1, int);
      if (eb_project (ebctx, evecs, numeigs, eb_ffvecs0, fxvecs, fxscal) != 0)
        eb_fail ("eb_project");
      eb_colmap (emap, xsnplist, ncols);
      for (j = 0; j < numeigs; j++)
        for (i = 0; i < ncols; i++)
          ffvecs[j * ncols + i] = eb_ffvecs0[j * eb_ncols0 + emap[i]];
      free (emap);
    }
    else
'''


def edit(src):
    def once(s, old, new):
        assert s.count(old) == 1, "anchor not unique / missing: %r (%d)" % (old[:60], s.count(old))
        return s.replace(old, new)

    s = src
    s = once(s, '#include "globals.h"\n', '#include "globals.h"\n#include "eigb200.h"\n')
    s = once(s, "int grmbinary = NO;", "int grmbinary = NO;")          # anchor check only
    # helpers go right before main's prototypes end: after the estedgar prototype (file scope, all globals declared above it)
    s = once(s, "void estedgar(double *edgarw, double *lambdav, int lentop, int lenspec, double gamm, double yjfac) ;\n",
             "void estedgar(double *edgarw, double *lambdav, int lentop, int lenspec, double gamm, double yjfac) ;\n" + HELPERS)
    # upload + no dense mmat on the GPU path
    s = once(s, "  if (shrinkmode) {\n    ZALLOC (mmat, nrows * ncols, double);\n    regmode = YES;\n  }\n",
             UPLOAD + "  if (shrinkmode) {\n    if (!EB_PACKED)\n      ZALLOC (mmat, nrows * ncols, double);\n    regmode = YES;\n  }\n")
    # the pass
    s = once(s, "    for (i = 0; i < ncols; i++) {\n      cupt = xsnplist[i];\n      chrom = cupt->chrom;\n",
             PASS + "    for (i = 0; i < ncols; i++) {\n      cupt = xsnplist[i];\n      chrom = cupt->chrom;\n")
    s = once(s, "    eigvecs (XTX, lambda, evecs, nrows);\n", "    eigvecs (XTX, lambda, evecs, nrows);\n  eb_after_eig:;\n")
    s = once(s, "    printf (\"number of samples after outlier removal: %d\\n\", nrows);\n  }\n",
             "    printf (\"number of samples after outlier removal: %d\\n\", nrows);\n  }\n" + FETCH_XTX)
    # loadings / projections
    s = once(s, "    for (i = 0; i < ncols; i++) {\n      cupt = xsnplist[i];\n      getcolxf (cc, cupt, xindex, nrows, i, NULL, NULL);\n\n      for (j = 0; j < numeigs; j++) {\n",
             PROJECT + "    for (i = 0; i < ncols; i++) {\n      cupt = xsnplist[i];\n      getcolxf (cc, cupt, xindex, nrows, i, NULL, NULL);\n\n      for (j = 0; j < numeigs; j++) {\n")
    s = once(s, "      xtypes[i] = k;\n\n      loadxdataind (xrow, xsnplist, xindex[i], ncols);\n      fixxrow (xrow, xmean, xfancy, ncols);\n",
             "      xtypes[i] = k;\n      if (EB_PACKED)\n        continue;\n\n      loadxdataind (xrow, xsnplist, xindex[i], ncols);\n      fixxrow (xrow, xmean, xfancy, ncols);\n")
    s = once(s, "      y = fxscal[j];\n      fxscal[j] = 1.0 / sqrt (y);       // standard\n",
             "      if (EB_PACKED)\n        break;\n      y = fxscal[j];\n      fxscal[j] = 1.0 / sqrt (y);       // standard\n")
    # lsqproj
    s = once(s, "      lsqproj(-99, xsnplist, ncols, indivmarkers, numindivs, fxscal, ffvecs, acoeffs, bcoeffs, xtypes, numeg) ; \n",
             "      if (EB_PACKED) eb_lsqproj_all (indivmarkers, numindivs, numeigs, fxscal, acoeffs, bcoeffs) ;\n      else\n"
             "      lsqproj(-99, xsnplist, ncols, indivmarkers, numindivs, fxscal, ffvecs, acoeffs, bcoeffs, xtypes, numeg) ; \n")
    # shrinkmode (two call sites)
    old_fill = "  for (i = 0; i < ncols; ++i) {\n    cupt = xsnplist[i];\n    getcolxf (cc, cupt, xindex, nrows, i, NULL, NULL);\n    for (j = 0; j < nrows; ++j) {\n      mmat[j * ncols + i] = cc[j];\n"
    assert s.count(old_fill) == 2
    s = s.replace(old_fill, "  if (!EB_PACKED)\n" + old_fill)
    s = once(s, "  doshrinkp (mmat, nrows, ncols, xindex, xsnplist, xcoeffs) ;\n",
             "  if (EB_PACKED) eb_doshrink (xcoeffs) ;\n  else\n  doshrinkp (mmat, nrows, ncols, xindex, xsnplist, xcoeffs) ;\n")
    s = once(s, "  doshrinkp (mmat, nrows, ncols, xindex, xsnplist, xcoeffs);\n",
             "  if (EB_PACKED) eb_doshrink (xcoeffs);\n  else\n  doshrinkp (mmat, nrows, ncols, xindex, xsnplist, xcoeffs);\n")
    return s


def main():
    src = open(SRC).read()
    out = edit(src)
    with tempfile.TemporaryDirectory() as d:
        a = os.path.join(d, "a"); b = os.path.join(d, "b")
        os.makedirs(os.path.join(a, "src", "eigensrc")); os.makedirs(os.path.join(b, "src", "eigensrc"))
        open(os.path.join(a, "src", "eigensrc", "smartpca.c"), "w").write(src)
        open(os.path.join(b, "src", "eigensrc", "smartpca.c"), "w").write(out)
        r = subprocess.run(["diff", "-U2", "a/src/eigensrc/smartpca.c", "b/src/eigensrc/smartpca.c"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 1, r.stderr
        patch = r.stdout
    # stable header (no timestamps)
    lines = patch.split("\n")
    lines[0] = "--- a/src/eigensrc/smartpca.c"
    lines[1] = "+++ b/src/eigensrc/smartpca.c"
    open(os.path.join(HERE, "smartpca_b200.patch"), "w").write("\n".join(lines))
    print("wrote", os.path.join(HERE, "smartpca_b200.patch"), "(%d lines)" % len(lines))


if __name__ == "__main__":
    main()
