/* Minimal GSL look-alike used ONLY to compile the reference's fastmode sources
 * (src/ksrc/kjg_fpca.c, src/ksrc/kjg_gsl.c) into oracle/_ref/ -- test infrastructure.
 * GSL itself is not installed in this image and is not vendored by the reference
 * (src/Makefile:3 links -lgsl).  Only the entry points those two files call are provided:
 * kjg_fpca.c:35-100, kjg_gsl.c:96-209.  Semantics follow the public GSL 2.x manual. */
#ifndef EIGB200_GSL_SHIM_CORE_H
#define EIGB200_GSL_SHIM_CORE_H
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct { size_t size; size_t stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; void *block; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef gsl_matrix_view gsl_matrix_const_view;

gsl_matrix *gsl_matrix_alloc (size_t n1, size_t n2);
void gsl_matrix_free (gsl_matrix * m);
gsl_matrix_view gsl_matrix_submatrix (gsl_matrix * m, size_t i, size_t j, size_t n1, size_t n2);
gsl_matrix_const_view gsl_matrix_const_submatrix (const gsl_matrix * m, size_t i, size_t j, size_t n1, size_t n2);
gsl_matrix_view gsl_matrix_view_array (double *base, size_t n1, size_t n2);
int gsl_matrix_scale (gsl_matrix * m, double x);
void gsl_matrix_set_zero (gsl_matrix * m);
int gsl_matrix_memcpy (gsl_matrix * dst, const gsl_matrix * src);
static inline double *gsl_matrix_ptr (gsl_matrix * m, size_t i, size_t j) { return m->data + i * m->tda + j; }
static inline double gsl_matrix_get (const gsl_matrix * m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set (gsl_matrix * m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }

gsl_vector *gsl_vector_alloc (size_t n);
void gsl_vector_free (gsl_vector * v);
gsl_vector_view gsl_vector_subvector (gsl_vector * v, size_t off, size_t n);
gsl_vector_view gsl_vector_view_array (double *base, size_t n);
int gsl_vector_mul (gsl_vector * a, const gsl_vector * b);
int gsl_vector_scale (gsl_vector * a, double x);
int gsl_vector_memcpy (gsl_vector * dst, const gsl_vector * src);
static inline double gsl_vector_get (const gsl_vector * v, size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set (gsl_vector * v, size_t i, double x) { v->data[i * v->stride] = x; }
#endif
