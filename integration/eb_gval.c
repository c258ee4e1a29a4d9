/* integration/eb_gval.c -- what replaces src/gval.c when smartpca is linked against libeigb200.so.
 *
 * Compiled against the REFERENCE's headers (include/admutils.h for SNP / Indiv) by integration/Makefile; this is the binding a
 * maintainer adds, not part of the library.  The reference hands the genotypes to kjg_fpca through gval.c's file statics
 * (setgval, gval.c:31-87); here setgval hands plain pointers to the library (eb_setgval_packed, include/eigb200.h) and applies
 * its one side effect (SNPs with min(n0, n1) == 0 get ignore = YES, gval.c:80-82).  kjg_fpca itself (kjg_fpca.c:24) is exported
 * by the library under its own name, so smartpca.c:957-963 compiles and links unchanged.
 * getgval / getggval (gval.c:90-150) are only reachable from printevecs' branch behind `fatalx ("... not yet implemented!")`
 * (smartpca.c:3875-3877); they are kept as hard failures.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#include <nicklib.h>

#include "admutils.h"
#include "mcio.h"
#include "gval.h"
#include "eigb200.h"

extern long rlen;               /* packit.h (mcio.c): bytes per SNP of the packed store */
extern int packmode;
extern int fancynorm, altnormstyle, usepopsformissing;   /* globals.h / smartpca.c */
extern long seed;               /* smartpca.c:167 */

void
setgval (SNP ** xsnps, int nrows, Indiv ** indivmarkers, int numindivs, int *xindex, int *xtypes, int ncols)
{
  const uint8_t **rows;
  uint8_t *mono;
  int i;

  if (!packmode)
    fatalx ("(libeigb200 setgval) genotypes are not in packed mode\n");
  if (usepopsformissing)
    fatalx ("(libeigb200 setgval) usepopsformissing is not supported with fastmode on the GPU\n");
  ZALLOC (rows, ncols, const uint8_t *);
  ZALLOC (mono, ncols, uint8_t);
  for (i = 0; i < ncols; ++i)
    rows[i] = (const uint8_t *) xsnps[i]->pbuff;
  if (eb_setgval_packed (rows, ncols, rlen, numindivs, xindex, nrows, fancynorm, altnormstyle, seed, mono) != 0)
    fatalx ("(libeigb200 setgval) %s\n", eb_last_error ());
  for (i = 0; i < ncols; ++i)
    if (mono[i])
      xsnps[i]->ignore = YES;   /* side-effect, gval.c:80-82 */
  free (rows);
  free (mono);
}

void
unsetgval ()
{
  eb_unsetgval ();
}

int
getgval (int row, int col, double *val)
{
  fatalx ("(libeigb200) getgval is not available: the genotype table lives on the GPU\n");
  return -1;
}

int
getggval (int indindx, int col, double *val)
{
  fatalx ("(libeigb200) getggval is not available: the genotype table lives on the GPU\n");
  return -1;
}
