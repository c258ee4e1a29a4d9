#include "gsl_shim_core.h"
