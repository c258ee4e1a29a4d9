// fpca_kernels.cu -- fastmode (kjg_fpca) and projection passes on the packed matrix.  (placeholder: filled in next)
#include <cmath>
#include "common.cuh"

namespace eb {
int fpca_run(eb_ctx*, int, int, size_t, size_t, size_t, long, double*, double*) {
  set_error("eb_fpca: not built yet");
  return EB_ERR_STATE;
}
int project_run(eb_ctx*, const double*, int, double*, double*, double*) {
  set_error("eb_project: not built yet");
  return EB_ERR_STATE;
}
}  // namespace eb

// Seeded Gaussian start matrix: kjg_gsl.c:96-113 (GSL mt19937, seed 0 -> 4357) and kjg_gsl.c:145-186
// (Marsaglia polar pairs from -1+2*uniform_pos; an odd last column gets a plain uniform).  Host code, bit-exact.
extern "C" void eb_gauss_matrix(long seed, size_t n, size_t L, double* out) {
  uint32_t mt[624];
  int mti = 624;
  unsigned long s = (unsigned long)seed;
  if (s == 0) s = 4357;
  mt[0] = (uint32_t)s;
  for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  auto next = [&]() -> uint32_t {
    if (mti >= 624) {
      for (int k = 0; k < 624; k++) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  };
  auto upos = [&]() -> double { double x; do { x = next() / 4294967296.0; } while (x == 0); return x; };
  for (size_t i = 0; i < n; i++) {
    double* row = out + i * L;
    size_t j = 0;
    for (; j + 1 < L; j += 2) {
      double x0, x1, r2;
      do { x0 = -1 + 2 * upos(); x1 = -1 + 2 * upos(); r2 = x0 * x0 + x1 * x1; } while (r2 > 1.0 || r2 == 0);
      r2 = sqrt(-2.0 * log(r2) / r2);
      row[j] = x0 * r2; row[j + 1] = x1 * r2;
    }
    if (L % 2) row[L - 1] = upos();
  }
}
