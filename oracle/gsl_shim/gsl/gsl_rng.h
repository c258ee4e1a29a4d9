/* gsl_rng subset: MT19937 only (GSL's default generator), see gsl_shim_core.h for scope. */
#ifndef EIGB200_GSL_SHIM_RNG_H
#define EIGB200_GSL_SHIM_RNG_H
#include "gsl_shim_core.h"
typedef struct { const char *name; } gsl_rng_type;
typedef struct { const gsl_rng_type *type; unsigned long mt[624]; int mti; } gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
extern unsigned long int gsl_rng_default_seed;
const gsl_rng_type *gsl_rng_env_setup (void);
gsl_rng *gsl_rng_alloc (const gsl_rng_type * T);
void gsl_rng_free (gsl_rng * r);
void gsl_rng_set (gsl_rng * r, unsigned long int seed);
unsigned long int gsl_rng_get (const gsl_rng * r);
double gsl_rng_uniform (const gsl_rng * r);
double gsl_rng_uniform_pos (const gsl_rng * r);
const char *gsl_rng_name (const gsl_rng * r);
#endif
