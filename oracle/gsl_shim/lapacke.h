/* Prototype-only stand-in for <lapacke.h>: the symbols are exported by the OpenBLAS
 * shared object the oracle links (see oracle/Makefile). Used by kjg_gsl.c:118-209. */
#ifndef EIGB200_LAPACKE_SHIM_H
#define EIGB200_LAPACKE_SHIM_H
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#define lapack_int int
double LAPACKE_dlange (int layout, char norm, lapack_int m, lapack_int n, const double *a, lapack_int lda);
lapack_int LAPACKE_dgeqrf (int layout, lapack_int m, lapack_int n, double *a, lapack_int lda, double *tau);
lapack_int LAPACKE_dorgqr (int layout, lapack_int m, lapack_int n, lapack_int k, double *a, lapack_int lda, const double *tau);
lapack_int LAPACKE_dgesvd (int layout, char jobu, char jobvt, lapack_int m, lapack_int n, double *a, lapack_int lda,
                           double *s, double *u, lapack_int ldu, double *vt, lapack_int ldvt, double *superb);
#endif
