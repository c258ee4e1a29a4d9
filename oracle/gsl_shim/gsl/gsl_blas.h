/* gsl_blas subset: dgemm on row-major gsl_matrix, forwarded to CBLAS (OpenBLAS). */
#ifndef EIGB200_GSL_SHIM_BLAS_H
#define EIGB200_GSL_SHIM_BLAS_H
#include "gsl_shim_core.h"
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER_t;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;
int gsl_blas_dgemm (CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                    const gsl_matrix * A, const gsl_matrix * B, double beta, gsl_matrix * C);
#endif
