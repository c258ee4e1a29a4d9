// tile_order.cuh -- order in which the persistent GRM kernel walks the lower-triangle tiles.
//
// The 148 CTAs work on 148 CONSECUTIVE tile indices at any time and stream the packed row operand (tile row ti) and column operand
// (tile column tj) of their tile through L2 in lockstep.  Row-major order makes them share one row operand and 148 different column
// operands (L2 hit rate 44 %, 309 GB of DRAM reads per 60,000-SNP launch at 50,000 rows).  Here the triangle is cut into bands of
// TILE_BAND tile rows and every band is walked column by column, so that 148 consecutive tiles form a ~12 x 12 patch: ~12 row
// operands + ~12 column operands for 296 operand streams.
#pragma once

namespace eb {

constexpr int TILE_BAND = 12;

// t in [0, T (T + 1) / 2) -> (ti >= tj).  Bands b = 0, 1, ...: tile rows [b S, min(b S + S, T)); inside a band: tj ascending, ti ascending.
__host__ __device__ __forceinline__ void tile_decode_banded(int t, int T, int& ti, int& tj) {
  const int S = TILE_BAND;
  // tiles before band b (all earlier bands are S rows high): S^2 b (b - 1) / 2 + b S (S + 1) / 2
  int b = (int)((sqrtf(1.0f + 8.0f * (float)t / (float)(S * S)) - 1.0f) * 0.5f);
  if (b < 0) b = 0;
  while ((long long)S * S * (b + 1) * b / 2 + (long long)(b + 1) * S * (S + 1) / 2 <= t) b++;
  while (b > 0 && (long long)S * S * b * (b - 1) / 2 + (long long)b * S * (S + 1) / 2 > t) b--;
  const int r0 = b * S;
  int r = t - (int)((long long)S * S * b * (b - 1) / 2 + (long long)b * S * (S + 1) / 2);
  const int h = (T - r0) < S ? (T - r0) : S;          // rows of this band
  const int rect = r0 * h;                            // columns left of the band's diagonal block: h tiles each
  if (r < rect) { tj = r / h; ti = r0 + r % h; return; }
  r -= rect;
  int c = 0;
  while (r >= h - c) { r -= h - c; c++; }             // diagonal block, column c has h - c tiles
  tj = r0 + c; ti = r0 + c + r;
}

}  // namespace eb
