/* Implementation of the GSL subset declared in gsl/ -- test infrastructure for oracle/_ref only.
 * MT19937 follows Matsumoto & Nishimura's published generator as GSL exposes it:
 * seeding mt[i] = 1812433253 * (mt[i-1] ^ (mt[i-1] >> 30)) + i, seed 0 -> 4357,
 * uniform = get()/2^32, uniform_pos rejects 0.  Known answers (GSL manual, "Random number
 * environment variables" example): seed 0 -> first gsl_rng_get = 4293858116,
 * seed 123 -> 2991312382.  Checked in tests/test_oracle_pins.py. */
#include <stdio.h>
#include "gsl/gsl_matrix.h"
#include "gsl/gsl_rng.h"
#include "gsl/gsl_blas.h"

void cblas_dgemm (int order, int ta, int tb, int M, int N, int K, double alpha, const double *A, int lda,
                  const double *B, int ldb, double beta, double *C, int ldc);

gsl_matrix *gsl_matrix_alloc (size_t n1, size_t n2)
{
  gsl_matrix *m = (gsl_matrix *) malloc (sizeof (gsl_matrix));
  m->size1 = n1; m->size2 = n2; m->tda = n2;
  m->data = (double *) malloc (sizeof (double) * (n1 * n2 ? n1 * n2 : 1));
  m->block = m->data; m->owner = 1;
  if (!m->data) { fprintf (stderr, "gsl_shim: out of memory (%zu x %zu)\n", n1, n2); exit (1); }
  return m;
}
void gsl_matrix_free (gsl_matrix * m) { if (!m) return; if (m->owner) free (m->data); free (m); }
gsl_matrix_view gsl_matrix_submatrix (gsl_matrix * m, size_t i, size_t j, size_t n1, size_t n2)
{
  gsl_matrix_view v; v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = m->tda;
  v.matrix.data = m->data + i * m->tda + j; v.matrix.block = m->block; v.matrix.owner = 0; return v;
}
gsl_matrix_const_view gsl_matrix_const_submatrix (const gsl_matrix * m, size_t i, size_t j, size_t n1, size_t n2)
{ return gsl_matrix_submatrix ((gsl_matrix *) m, i, j, n1, n2); }
gsl_matrix_view gsl_matrix_view_array (double *base, size_t n1, size_t n2)
{
  gsl_matrix_view v; v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = n2;
  v.matrix.data = base; v.matrix.block = 0; v.matrix.owner = 0; return v;
}
int gsl_matrix_scale (gsl_matrix * m, double x)
{ for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] *= x; return 0; }
void gsl_matrix_set_zero (gsl_matrix * m)
{ for (size_t i = 0; i < m->size1; i++) memset (m->data + i * m->tda, 0, sizeof (double) * m->size2); }
int gsl_matrix_memcpy (gsl_matrix * d, const gsl_matrix * s)
{ for (size_t i = 0; i < s->size1; i++) memcpy (d->data + i * d->tda, s->data + i * s->tda, sizeof (double) * s->size2); return 0; }

gsl_vector *gsl_vector_alloc (size_t n)
{
  gsl_vector *v = (gsl_vector *) malloc (sizeof (gsl_vector));
  v->size = n; v->stride = 1; v->data = (double *) malloc (sizeof (double) * (n ? n : 1)); v->block = v->data; v->owner = 1; return v;
}
void gsl_vector_free (gsl_vector * v) { if (!v) return; if (v->owner) free (v->data); free (v); }
gsl_vector_view gsl_vector_subvector (gsl_vector * v, size_t off, size_t n)
{ gsl_vector_view w; w.vector.size = n; w.vector.stride = v->stride; w.vector.data = v->data + off * v->stride; w.vector.block = v->block; w.vector.owner = 0; return w; }
gsl_vector_view gsl_vector_view_array (double *base, size_t n)
{ gsl_vector_view w; w.vector.size = n; w.vector.stride = 1; w.vector.data = base; w.vector.block = 0; w.vector.owner = 0; return w; }
int gsl_vector_mul (gsl_vector * a, const gsl_vector * b)
{ for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] *= b->data[i * b->stride]; return 0; }
int gsl_vector_scale (gsl_vector * a, double x) { for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] *= x; return 0; }
int gsl_vector_memcpy (gsl_vector * d, const gsl_vector * s)
{ for (size_t i = 0; i < s->size; i++) d->data[i * d->stride] = s->data[i * s->stride]; return 0; }

int gsl_blas_dgemm (CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, double alpha, const gsl_matrix * A,
                    const gsl_matrix * B, double beta, gsl_matrix * C)
{
  int M = (int) C->size1, N = (int) C->size2;
  int K = (int) (ta == CblasNoTrans ? A->size2 : A->size1);
  cblas_dgemm (CblasRowMajor, ta, tb, M, N, K, alpha, A->data, (int) A->tda, B->data, (int) B->tda, beta, C->data, (int) C->tda);
  return 0;
}

/* ---- MT19937 ---- */
static const gsl_rng_type mt19937_type = { "mt19937" };
const gsl_rng_type *gsl_rng_default = &mt19937_type;
unsigned long int gsl_rng_default_seed = 0;

const gsl_rng_type *gsl_rng_env_setup (void)
{
  const char *s = getenv ("GSL_RNG_SEED");
  gsl_rng_default = &mt19937_type;      /* GSL_RNG_TYPE other than mt19937 is not supported by the shim */
  gsl_rng_default_seed = s ? strtoul (s, 0, 0) : 0;
  return gsl_rng_default;
}
void gsl_rng_set (gsl_rng * r, unsigned long int s)
{
  if (s == 0) s = 4357;
  r->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < 624; i++)
    r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long) i) & 0xffffffffUL;
  r->mti = 624;
}
gsl_rng *gsl_rng_alloc (const gsl_rng_type * T)
{ gsl_rng *r = (gsl_rng *) malloc (sizeof (gsl_rng)); r->type = T; gsl_rng_set (r, gsl_rng_default_seed); return r; }
void gsl_rng_free (gsl_rng * r) { free (r); }
const char *gsl_rng_name (const gsl_rng * r) { return r->type->name; }
unsigned long int gsl_rng_get (const gsl_rng * cr)
{
  gsl_rng *r = (gsl_rng *) cr;
  unsigned long *mt = r->mt, y;
  if (r->mti >= 624) {
    int k;
    for (k = 0; k < 624 - 397; k++) { y = (mt[k] & 0x80000000UL) | (mt[k + 1] & 0x7fffffffUL); mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0); }
    for (; k < 623; k++) { y = (mt[k] & 0x80000000UL) | (mt[k + 1] & 0x7fffffffUL); mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0); }
    y = (mt[623] & 0x80000000UL) | (mt[0] & 0x7fffffffUL); mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
    r->mti = 0;
  }
  y = mt[r->mti++];
  y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680UL; y ^= (y << 15) & 0xefc60000UL; y ^= (y >> 18);
  return y & 0xffffffffUL;
}
double gsl_rng_uniform (const gsl_rng * r) { return gsl_rng_get (r) / 4294967296.0; }
double gsl_rng_uniform_pos (const gsl_rng * r) { double x; do { x = gsl_rng_uniform (r); } while (x == 0); return x; }
