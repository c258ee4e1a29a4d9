// pack_kernels.cu -- HBM-bound integer/byte kernels over the 2-bit packed genotype store.
//   gather_rows_kernel   : raw slab [nsnp][rlen] + xindex -> working matrix [mpad][npad/4] (selected rows, pad = 3)
//                          replaces getrawcol/getgtypes gathers (qpsubs.c:240-248, admutils.c:575-592)
//   snp_stats_kernel     : per-SNP c0,c1,nmiss (getcolxz_binary1, smartpca.c:3261-3276) fused with the
//                          normalisation (fvadjust_binary, smartpca.c:2282-2313), the drop rule
//                          (smartpca.c:1131-1144) and the 4-entry decode table the GRM kernel consumes
//   indiv_counts_kernel  : numvalidgtallind (admutils.c:1075-1097)
//   synth_kernel         : synthetic Hardy-Weinberg generator (eig_b200/synth.py is the host twin)
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace eb {

// ---------------------------------------------------------------------------------------------- gather
// One thread builds one 32-bit word of the working row = 16 consecutive selected individuals.
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* __restrict__ raw, int64_t raw_pitch, int64_t nsnp,
                                                          int64_t mpad, const int* __restrict__ xindex, int nrows,
                                                          uint8_t* __restrict__ work, int64_t wpitch) {
  const int wordsPerRow = (int)(wpitch >> 2);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= wordsPerRow) return;
  int src[16];
#pragma unroll
  for (int t = 0; t < 16; t++) {
    int j = w * 16 + t;
    src[t] = j < nrows ? xindex[j] : -1;
  }
  // 16 consecutive individuals starting on a byte boundary: the word is 4 consecutive bytes of the raw row (same MSB-first layout) -- the common case (all individuals,
  // or long runs between a few removed outliers)
  bool run = src[0] >= 0 && (src[0] & 3) == 0;
#pragma unroll
  for (int t = 1; t < 16; t++) run = run && src[t] == src[0] + t;
  const int b0 = src[0] >> 2;
  for (int64_t s = blockIdx.y; s < mpad; s += gridDim.y) {
    uint32_t out = 0xFFFFFFFFu;
    if (s < nsnp) {
      const uint8_t* row = raw + s * raw_pitch;
      if (run) {
        out = (uint32_t)row[b0] | ((uint32_t)row[b0 + 1] << 8) | ((uint32_t)row[b0 + 2] << 16) | ((uint32_t)row[b0 + 3] << 24);
        reinterpret_cast<uint32_t*>(work + s * wpitch)[w] = out;
        continue;
      }
      out = 0;
#pragma unroll
      for (int t = 0; t < 16; t++) {
        uint32_t code = 3;
        if (src[t] >= 0) code = (row[src[t] >> 2] >> ((3 - (src[t] & 3)) << 1)) & 3u;
        // byte t>>2 of the word (little endian), individual t&3 inside the byte MSB-first
        out |= code << (((t >> 2) << 3) + ((3 - (t & 3)) << 1));
      }
    }
    reinterpret_cast<uint32_t*>(work + s * wpitch)[w] = out;
  }
}

// Fast path of the same gather for the working matrix of the PCA rows: the host has classified every 32-bit word of the working
// row once per row selection (wsrc[w] >= 0: the word is the 4 consecutive raw bytes starting at wsrc[w], i.e. 16 consecutive
// individuals that start on a byte boundary -- all individuals, or the long runs between a few removed outliers; -1: mixed word,
// gathered genotype by genotype; -2: pad word).  One thread builds one 16-byte vector (64 individuals) per SNP row.
__global__ void __launch_bounds__(256) gather_rows_fast_kernel(const uint8_t* __restrict__ raw, int64_t raw_pitch, int64_t nsnp, int64_t mpad,
                                                               const int* __restrict__ xindex, int nrows, const int4* __restrict__ wsrc,
                                                               uint8_t* __restrict__ work, int64_t wpitch) {
  const int vecs = (int)(wpitch >> 4);
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= vecs) return;
  const int4 src4 = wsrc[v];
  const int src[4] = {src4.x, src4.y, src4.z, src4.w};
  const bool rows_aligned = ((raw_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(raw) & 3) == 0);
  for (int64_t s = blockIdx.y; s < mpad; s += gridDim.y) {
    uint32_t out[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (s < nsnp) {
      const uint8_t* row = raw + s * raw_pitch;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (src[k] >= 0) {
          const uint8_t* p = row + src[k];
          if (rows_aligned && (src[k] & 3) == 0) out[k] = __ldg(reinterpret_cast<const uint32_t*>(p));
          else out[k] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
        } else if (src[k] == -1) {
          uint32_t o = 0;
          const int j0 = (v * 4 + k) * 16;
#pragma unroll
          for (int t = 0; t < 16; t++) {
            uint32_t code = 3;
            const int j = j0 + t;
            if (j < nrows) { const int q = xindex[j]; code = (__ldg(row + (q >> 2)) >> ((3 - (q & 3)) << 1)) & 3u; }
            o |= code << (((t >> 2) << 3) + ((3 - (t & 3)) << 1));
          }
          out[k] = o;
        }
      }
    }
    reinterpret_cast<uint4*>(work + s * wpitch)[v] = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

int launch_gather(eb_ctx* c) {
  // classify the words of the working row (host, once per row selection)
  const int words = (int)(c->wpitch >> 2);
  std::vector<int> wsrc((size_t)words);
  for (int w = 0; w < words; w++) {
    const int j0 = w * 16;
    if (j0 >= c->nrows) { wsrc[w] = -2; continue; }
    bool run = j0 + 16 <= c->nrows && (c->xindex_h[j0] & 3) == 0;
    for (int t = 1; run && t < 16; t++) run = c->xindex_h[j0 + t] == c->xindex_h[j0] + t;
    wsrc[w] = run ? (c->xindex_h[j0] >> 2) : -1;
  }
  int rc;
  if ((rc = c->wsrc_d.ensure((size_t)words))) return rc;
  EB_CUDA(cudaMemcpyAsync(c->wsrc_d.p, wsrc.data(), sizeof(int) * words, cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));      // wsrc is a local
  const int vecs = (int)(c->wpitch >> 4);
  dim3 block(256), grid((vecs + 255) / 256, (unsigned)std::min<int64_t>(c->mpad, 65535));
  gather_rows_fast_kernel<<<grid, block, 0, c->stream>>>(c->raw, c->raw_pitch, c->nsnp, c->mpad, c->xindex_d.p, c->nrows,
                                                         reinterpret_cast<const int4*>(c->wsrc_d.p), c->work.p, c->wpitch);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// gather an arbitrary individual list (device array) into a caller-provided working matrix [mpad][wpitch]
int launch_gather_into(eb_ctx* c, const int* list_d, int nlist, uint8_t* dst, int64_t wpitch) {
  const int wordsPerRow = (int)(wpitch >> 2);
  dim3 block(256), grid((wordsPerRow + 255) / 256, (unsigned)std::min<int64_t>(c->mpad, 65535));
  gather_rows_kernel<<<grid, block, 0, c->stream>>>(c->raw, c->raw_pitch, c->nsnp, c->mpad, list_d, nlist, dst, wpitch);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// ---------------------------------------------------------------------------------------------- stats
__device__ __forceinline__ void count_word(uint32_t x, int& n1, int& n2, int& n3) {
  const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
  n1 += __popc(lo & ~hi);
  n2 += __popc(hi & ~lo);
  n3 += __popc(hi & lo);
}

// LPR lanes per SNP row of the working matrix (32 / LPR rows per warp: short rows -- 1.25 KB at 5,000 individuals -- would leave a
// whole warp with two or three loads in flight); four independent 16-byte loads per lane and trip; shuffle reduction of the three
// code counts inside the lane group.
template <int LPR>
__global__ void __launch_bounds__(256) snp_stats_kernel(const uint8_t* __restrict__ work, int64_t wpitch, int64_t nsnp, int64_t mpad,
                                                        int nrows, int npad, int fancynorm, int altnormstyle, int minallelecnt,
                                                        int maxmissing, const uint8_t* __restrict__ ignore,
                                                        const double* __restrict__ weight, int* __restrict__ c0o,
                                                        int* __restrict__ c1o, int* __restrict__ nmisso, uint8_t* __restrict__ usedo,
                                                        double* __restrict__ xmeano, double* __restrict__ xfancyo,
                                                        double* __restrict__ table, unsigned long long* __restrict__ nused) {
  constexpr int RPW = 32 / LPR;                      // rows per warp
  const int lane = threadIdx.x & (LPR - 1);         // lane inside the row's group
  const int grp = (threadIdx.x & 31) / LPR;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int vecs = (int)(wpitch >> 4);
  const int padcount = npad - nrows;
  // every lane of a warp runs the same number of trips (mpad is a multiple of 128, hence of RPW)
  for (int64_t s = warp * RPW + grp; s < mpad; s += nwarps * RPW) {
    double t0 = 0, t1 = 0, t2 = 0;
    if (s < nsnp) {
      const uint4* row = reinterpret_cast<const uint4*>(work + s * wpitch);
      int n1 = 0, n2 = 0, n3 = 0;
      int v = lane;
      for (; v + 3 * LPR < vecs; v += 4 * LPR) {
        const uint4 q0 = __ldg(row + v), q1 = __ldg(row + v + LPR), q2 = __ldg(row + v + 2 * LPR), q3 = __ldg(row + v + 3 * LPR);
        count_word(q0.x, n1, n2, n3); count_word(q0.y, n1, n2, n3); count_word(q0.z, n1, n2, n3); count_word(q0.w, n1, n2, n3);
        count_word(q1.x, n1, n2, n3); count_word(q1.y, n1, n2, n3); count_word(q1.z, n1, n2, n3); count_word(q1.w, n1, n2, n3);
        count_word(q2.x, n1, n2, n3); count_word(q2.y, n1, n2, n3); count_word(q2.z, n1, n2, n3); count_word(q2.w, n1, n2, n3);
        count_word(q3.x, n1, n2, n3); count_word(q3.y, n1, n2, n3); count_word(q3.z, n1, n2, n3); count_word(q3.w, n1, n2, n3);
      }
      for (; v < vecs; v += LPR) {
        const uint4 q = __ldg(row + v);
        count_word(q.x, n1, n2, n3); count_word(q.y, n1, n2, n3);
        count_word(q.z, n1, n2, n3); count_word(q.w, n1, n2, n3);
      }
      // groups of one warp may disagree on s < nsnp in the pad tail: synchronise the group's lanes only
      const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (grp * LPR));
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        n1 += __shfl_xor_sync(gmask, n1, o);
        n2 += __shfl_xor_sync(gmask, n2, o);
        n3 += __shfl_xor_sync(gmask, n3, o);
      }
      if (lane == 0) {
        int nmiss = n3 - padcount;
        int a = n1 + 2 * n2;                       // c0 = sum g
        int b = 2 * (nrows - nmiss) - a;           // c1 = sum (2-g)
        double xm = 0.0, xf = 0.0;
        int tt = nmiss;
        if (nmiss == nrows) { a = -1; b = -1; tt = -1; }
        else {
          // fvadjust_binary, smartpca.c:2290-2305, with explicitly rounded operations (no FMA contraction)
          const double ynum = (double)(nrows - nmiss), ysum = (double)a;
          const double ymean = __ddiv_rn(ysum, ynum);
          double yfancy = 1.0;
          if (fancynorm) {
            double p = __dmul_rn(0.5, ymean);
            if (!altnormstyle) p = __ddiv_rn(__dadd_rn(ysum, 1.0), __dadd_rn(__dmul_rn(2.0, ynum), 2.0));
            const double y = __dmul_rn(p, __dadd_rn(1.0, -p));
            if (y > 0.0) yfancy = __ddiv_rn(1.0, __dsqrt_rn(y));
          }
          t0 = __dmul_rn(-ymean, yfancy);
          t1 = __dmul_rn(__dadd_rn(1.0, -ymean), yfancy);
          t2 = __dmul_rn(__dadd_rn(2.0, -ymean), yfancy);
          xm = __dmul_rn(ymean, yfancy); xf = yfancy;
        }
        const int t = a < b ? a : b;
        bool drop = (t < minallelecnt) || (tt > maxmissing) || (tt < 0) || (t == 0);
        if (ignore && ignore[s]) { drop = true; xm = 0.0; xf = 0.0; }
        if (drop) { t0 = t1 = t2 = 0.0; }
        else if (weight) { const double w = weight[s]; t0 = __dmul_rn(t0, w); t1 = __dmul_rn(t1, w); t2 = __dmul_rn(t2, w); }
        c0o[s] = a; c1o[s] = b; nmisso[s] = tt; usedo[s] = drop ? 0 : 1;
        xmeano[s] = xm; xfancyo[s] = xf;
        if (!drop) atomicAdd(nused, 1ull);
      }
    }
    if (lane == 0) {
      double4 row4 = make_double4(t0, t1, t2, 0.0);
      reinterpret_cast<double4*>(table)[s] = row4;
    }
  }
}

// ---------------------------------------------------------------------------------------------- per-population counts
// Working matrix gathered in population-sorted order, every population's segment padded to whole 32-bit words (16
// individuals, pad = code 3).  One warp per SNP; for each population a popcount sweep over its words.
// out[(s * npops + k) * 3 + g] = number of members of population k with genotype g at SNP s  (fstcolyy's ddd[k],
// qpsubs.c:1256-1281: c0 = n1 + 2 n2, c1 = n1 + 2 n0; inbreed mode uses the classes directly).
__global__ void __launch_bounds__(256) pop_counts_kernel(const uint8_t* __restrict__ work, int64_t wpitch, int64_t nsnp, int npops,
                                                         const int* __restrict__ seg_word0, int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t s = warp; s < nsnp; s += nwarps) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(work + s * wpitch);
    for (int k = 0; k < npops; k++) {
      const int w0 = seg_word0[k], w1 = seg_word0[k + 1];
      int n1 = 0, n2 = 0, n3 = 0;
      for (int w = w0 + lane; w < w1; w += 32) count_word(__ldg(row + w), n1, n2, n3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        n3 += __shfl_xor_sync(0xffffffffu, n3, o);
      }
      if (lane == 0) {
        int* o3 = out + ((size_t)s * npops + k) * 3;
        o3[0] = (w1 - w0) * 16 - n1 - n2 - n3; o3[1] = n1; o3[2] = n2;
      }
    }
  }
}

int launch_pop_counts(eb_ctx* c, const uint8_t* work3, int64_t wp3, int npops, const int* seg_word0_d, int* out_d) {
  const int blocks = c->num_sms * 8;
  pop_counts_kernel<<<blocks, 256, 0, c->stream>>>(work3, wp3, c->nsnp, npops, seg_word0_d, out_d);
  EB_CHECK_LAUNCH(c);
  return 0;
}

int launch_stats(eb_ctx* c, const eb_grm_opts* o) {
  const int vecs = (int)(c->wpitch >> 4);
  const int lpr = vecs <= 128 ? 8 : (vecs <= 512 ? 16 : 32);       // lanes per row: ~16 loads per lane for the short rows
  const int rowsPerBlock = 8 * (32 / lpr);
  int64_t blocks = (c->mpad + rowsPerBlock - 1) / rowsPerBlock;
  blocks = std::min<int64_t>(blocks, (int64_t)c->num_sms * 16);
  EB_CUDA(cudaMemsetAsync(c->nused_d.p, 0, sizeof(long long), c->stream));
#define EB_STATS_LAUNCH(LPR_)                                                                                                      \
  snp_stats_kernel<LPR_><<<(unsigned)blocks, 256, 0, c->stream>>>(                                                                  \
      c->work.p, c->wpitch, c->nsnp, c->mpad, c->nrows, c->npad, o->fancynorm, o->altnormstyle, o->minallelecnt, o->maxmissing,      \
      o->snp_ignore ? c->ignore_d.p : nullptr, o->snp_weight ? c->weight_d.p : nullptr, c->c0_d.p, c->c1_d.p, c->nmiss_d.p,           \
      c->used_d.p, c->xmean_d.p, c->xfancy_d.p, c->table_d.p, reinterpret_cast<unsigned long long*>(c->nused_d.p))
  if (lpr == 8) EB_STATS_LAUNCH(8);
  else if (lpr == 16) EB_STATS_LAUNCH(16);
  else EB_STATS_LAUNCH(32);
#undef EB_STATS_LAUNCH
  EB_CHECK_LAUNCH(c);
  return 0;
}

// ---------------------------------------------------------------------------------------------- usepopsformissing
// getcolxz with usepopsformissing (smartpca.c:3129-3216) + fvadjust (2236-2279) for every SNP: a missing genotype of an individual
// whose population has data at the SNP takes that population's mean; mean / scale are then taken over observed AND filled values.
// One warp per SNP of the working matrix; per-population sums live in the warp's slice of shared memory (integer atomics).
// Outputs per SNP: c0, c1 (observed genotypes only), nmiss AFTER the fill (-1: no value at all), used, xmean, xfancy, the 3-entry
// table of the observed genotypes {(g - ymean) yfancy w, .., 0} and the fill values fill[s][k] = (mean_k - ymean) yfancy w (0 when
// population k has no data at s).  The filled sum is accumulated population by population (the reference adds row by row): the
// results agree to rounding, not bit for bit -- the dense path's bar is 1e-11 relative on the GRM.
__global__ void __launch_bounds__(256) popfill_stats_kernel(const uint8_t* __restrict__ work, int64_t wpitch, int64_t nsnp, int64_t mpad, int nrows,
                                                            const int* __restrict__ xt, int npops, int fancynorm, int altnormstyle, int minallelecnt,
                                                            int maxmissing, const uint8_t* __restrict__ ignore, const double* __restrict__ weight,
                                                            int* __restrict__ c0o, int* __restrict__ c1o, int* __restrict__ nmisso, int* __restrict__ nmiss0o,
                                                            uint8_t* __restrict__ usedo, double* __restrict__ xmeano, double* __restrict__ xfancyo,
                                                            double* __restrict__ table, double* __restrict__ fill, unsigned long long* __restrict__ nused) {
  extern __shared__ int pf_sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int* psum = pf_sm + (size_t)wib * 3 * npops;
  int* pnum = psum + npops;
  int* pmis = pnum + npops;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int words = (int)(wpitch >> 2);
  for (int64_t s = warp; s < mpad; s += nwarps) {
    for (int k = lane; k < 3 * npops; k += 32) psum[k] = 0;
    __syncwarp();
    int c0 = 0, nobs = 0, nmiss0 = 0, lost = 0;          // lost: missing genotypes of individuals outside every population
    if (s < nsnp) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(work + s * wpitch);
      for (int w = lane; w < words; w += 32) {
        const uint32_t x = __ldg(row + w);
        const int j0 = w * 16;
#pragma unroll
        for (int t = 0; t < 16; t++) {
          const int j = j0 + t;
          if (j >= nrows) break;
          const int code = (x >> (((t >> 2) << 3) + ((3 - (t & 3)) << 1))) & 3;
          const int k = xt[j];
          if (code < 3) {
            c0 += code; nobs++;
            if (k >= 0 && k < npops) { atomicAdd(psum + k, code); atomicAdd(pnum + k, 1); }
          } else {
            nmiss0++;
            if (k >= 0 && k < npops) atomicAdd(pmis + k, 1); else lost++;
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o); nobs += __shfl_xor_sync(0xffffffffu, nobs, o);
      nmiss0 += __shfl_xor_sync(0xffffffffu, nmiss0, o); lost += __shfl_xor_sync(0xffffffffu, lost, o);
    }
    __syncwarp();
    double fsum = 0.0; int filled = 0, remain = 0;
    for (int k = lane; k < npops; k += 32) {
      if (pnum[k] > 0) { fsum += (double)pmis[k] * __ddiv_rn((double)psum[k], (double)pnum[k]); filled += pmis[k]; }
      else remain += pmis[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      fsum += __shfl_xor_sync(0xffffffffu, fsum, o); filled += __shfl_xor_sync(0xffffffffu, filled, o);
      remain += __shfl_xor_sync(0xffffffffu, remain, o);
    }
    remain += lost;
    // every lane holds the same totals: the normalisation is computed redundantly, lane 0 writes the per-SNP scalars
    int a = c0, b = 2 * nobs - c0, tt = remain;
    double ym = 0.0, yf = 0.0, xm = 0.0;
    const double ynum = (double)(nobs + filled), ysum = (double)c0 + fsum;
    bool drop;
    if (!(ynum > 0.0) || s >= nsnp) { a = -1; b = -1; tt = -1; drop = true; }
    else {
      ym = __ddiv_rn(ysum, ynum);
      yf = 1.0;
      if (fancynorm) {
        double p = __dmul_rn(0.5, ym);
        if (!altnormstyle) p = __ddiv_rn(__dadd_rn(ysum, 1.0), __dadd_rn(__dmul_rn(2.0, ynum), 2.0));
        const double y = __dmul_rn(p, __dadd_rn(1.0, -p));
        if (y > 0.0) yf = __ddiv_rn(1.0, __dsqrt_rn(y));
      }
      xm = __dmul_rn(ym, yf);
      const int t = a < b ? a : b;
      drop = (t < minallelecnt) || (tt > maxmissing) || (t == 0);
    }
    if (s < nsnp && ignore && ignore[s]) { drop = true; xm = 0.0; yf = 0.0; }
    const double w = (!drop && weight) ? weight[s] : 1.0;
    for (int k = lane; k < npops; k += 32)
      fill[(size_t)s * npops + k] = (!drop && pnum[k] > 0) ? __dmul_rn(__dmul_rn(__dadd_rn(__ddiv_rn((double)psum[k], (double)pnum[k]), -ym), yf), w) : 0.0;
    if (lane == 0) {
      double t0 = 0, t1 = 0, t2 = 0;
      if (!drop) {
        t0 = __dmul_rn(__dmul_rn(-ym, yf), w); t1 = __dmul_rn(__dmul_rn(__dadd_rn(1.0, -ym), yf), w); t2 = __dmul_rn(__dmul_rn(__dadd_rn(2.0, -ym), yf), w);
      }
      reinterpret_cast<double4*>(table)[s] = make_double4(t0, t1, t2, 0.0);
      if (s < nsnp) {
        c0o[s] = a; c1o[s] = b; nmisso[s] = tt; nmiss0o[s] = nmiss0; usedo[s] = drop ? 0 : 1;
        xmeano[s] = xm; xfancyo[s] = (tt < 0) ? 0.0 : yf;
        if (!drop) atomicAdd(nused, 1ull);
      }
    }
    __syncwarp();
  }
}

// FP64 columns of SNPs [s0, s0 + nb) -> blk[b][i] (row pitch npad): table value of an observed genotype, the population's fill value
// for a missing one, zero in the pad
__global__ void __launch_bounds__(256) popfill_cols_kernel(const uint8_t* __restrict__ work, int64_t wpitch, int64_t s0, int nb, int nrows, int npad,
                                                           const int* __restrict__ xt, int npops, const double* __restrict__ table,
                                                           const double* __restrict__ fill, double* __restrict__ blk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= npad || b >= nb) return;
  const int64_t s = s0 + b;
  double v = 0.0;
  if (i < nrows) {
    const int code = (work[s * wpitch + (i >> 2)] >> ((3 - (i & 3)) << 1)) & 3;
    if (code < 3) v = table[s * 4 + code];
    else { const int k = xt[i]; if (k >= 0 && k < npops) v = fill[(size_t)s * npops + k]; }
  }
  blk[(size_t)b * npad + i] = v;
}

int launch_popfill_stats(eb_ctx* c, const eb_grm_opts* o, const int* xt_d, int npops, int* nmiss_after_d, double* fill_d) {
  const int warps = 8;
  const size_t smem = sizeof(int) * warps * 3 * (size_t)npops;
  if (smem > 200 * 1024) { set_error("usepopsformissing: too many populations (%d)", npops); return EB_ERR_ARG; }
  EB_CUDA(cudaFuncSetAttribute(popfill_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
  int64_t blocks = std::min<int64_t>((c->mpad + warps - 1) / warps, (int64_t)c->num_sms * 8);
  EB_CUDA(cudaMemsetAsync(c->nused_d.p, 0, sizeof(long long), c->stream));
  popfill_stats_kernel<<<(unsigned)blocks, 256, smem, c->stream>>>(
      c->work.p, c->wpitch, c->nsnp, c->mpad, c->nrows, xt_d, npops, o->fancynorm, o->altnormstyle, o->minallelecnt, o->maxmissing,
      o->snp_ignore ? c->ignore_d.p : nullptr, o->snp_weight ? c->weight_d.p : nullptr, c->c0_d.p, c->c1_d.p, nmiss_after_d, c->nmiss_d.p,
      c->used_d.p, c->xmean_d.p, c->xfancy_d.p, c->table_d.p, fill_d, reinterpret_cast<unsigned long long*>(c->nused_d.p));
  EB_CHECK_LAUNCH(c);
  return 0;
}

int launch_popfill_cols(eb_ctx* c, int64_t s0, int nb, const int* xt_d, int npops, const double* fill_d, double* blk) {
  dim3 grid((c->npad + 255) / 256, nb);
  popfill_cols_kernel<<<grid, 256, 0, c->stream>>>(c->work.p, c->wpitch, s0, nb, c->nrows, c->npad, xt_d, npops, c->table_d.p, fill_d, blk);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// ---------------------------------------------------------------------------------------------- per-individual counts
// thread = one raw byte column (4 individuals); blockIdx.y strides over SNPs; integer atomics (order-independent).
__global__ void __launch_bounds__(256) indiv_counts_kernel(const uint8_t* __restrict__ raw, int64_t raw_pitch, int64_t nsnp,
                                                           int numindivs, const uint8_t* __restrict__ keep, int* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int nbytes = (numindivs + 3) >> 2;
  if (b >= nbytes) return;
  int n[4] = {0, 0, 0, 0};
  for (int64_t s = blockIdx.y; s < nsnp; s += gridDim.y) {
    if (keep && !keep[s]) continue;
    const uint32_t x = raw[s * raw_pitch + b];
#pragma unroll
    for (int t = 0; t < 4; t++) n[t] += (((x >> ((3 - t) << 1)) & 3u) != 3u);
  }
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int j = b * 4 + t;
    if (j < numindivs && n[t]) atomicAdd(out + j, n[t]);
  }
}

int launch_indiv_counts(eb_ctx* c, const uint8_t* keep_d, int* out_d) {
  const int nbytes = (c->numindivs + 3) >> 2;
  EB_CUDA(cudaMemsetAsync(out_d, 0, sizeof(int) * c->numindivs, c->stream));
  dim3 grid((nbytes + 255) / 256, (unsigned)std::min<int64_t>(std::max<int64_t>(c->nsnp / 64, 1), 2048));
  indiv_counts_kernel<<<grid, 256, 0, c->stream>>>(c->raw, c->raw_pitch, c->nsnp, c->numindivs, keep_d, out_d);
  EB_CHECK_LAUNCH(c);
  return 0;
}

// ---------------------------------------------------------------------------------------------- synthetic generator
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

__global__ void __launch_bounds__(256) synth_kernel(uint8_t* __restrict__ dst, int64_t nsnp, int64_t pitch, int numindivs,
                                                    uint64_t seed, int64_t s0, uint32_t Tmiss, int use_missing, int npops,
                                                    double delta) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pitch) return;
  const uint64_t C0 = 0x9E3779B97F4A7C15ull, C1 = 0xBF58476D1CE4E5B9ull, C2 = 0x94D049BB133111EBull;
  const uint64_t CM = 0xD6E8FEB86659FD93ull, CP = 0xA0761D6478BD642Full;
  const double sqrt12 = __dsqrt_rn(12.0);
  for (int64_t r = blockIdx.y; r < nsnp; r += gridDim.y) {
    const uint64_t s = (uint64_t)(s0 + r);
    const uint64_t base = seed * C0 + s * C1;
    const double u0 = (double)(mix64(base + 0xFFFFFFFFull * C2) >> 11) / 9007199254740992.0;
    const double p = __dadd_rn(0.05, __dmul_rn(0.9, u0));
    uint32_t byte = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int i = b * 4 + t;
      uint32_t code = 3;
      if (i < numindivs) {
        const uint64_t key = base + (uint64_t)i * C2;
        const uint64_t h = mix64(key);
        double pi = p;
        if (npops > 1) {
          const uint64_t k = ((uint64_t)i * (uint64_t)npops) / (uint64_t)numindivs;
          const double u = (double)(mix64((base + k * C2) ^ CP) >> 11) / 9007199254740992.0;
          const double t1 = __dmul_rn(__dadd_rn(u, -0.5), sqrt12);
          const double t2 = __dmul_rn(delta, t1);
          const double t3 = __dsqrt_rn(__dmul_rn(p, __dadd_rn(1.0, -p)));
          pi = __dadd_rn(p, __dmul_rn(t2, t3));
          pi = fmin(fmax(pi, 0.01), 0.99);
        }
        const uint64_t T = (uint64_t)__dmul_rn(pi, 4294967296.0);
        code = (uint32_t)((h & 0xFFFFFFFFull) < T) + (uint32_t)((h >> 32) < T);
        if (use_missing && (uint32_t)(mix64(key ^ CM) & 0xFFFFFFFFull) < Tmiss) code = 3;
      }
      byte |= code << ((3 - t) << 1);
    }
    dst[r * pitch + b] = (uint8_t)byte;
  }
}

int launch_synth(eb_ctx* c, uint8_t* dst, int64_t nsnp, int64_t pitch, int numindivs, uint64_t seed, int64_t s0,
                 double missing, int npops, double delta) {
  dim3 grid((unsigned)((pitch + 255) / 256), (unsigned)std::min<int64_t>(nsnp, 65535));
  const uint32_t Tm = (uint32_t)(uint64_t)(missing * 4294967296.0);
  synth_kernel<<<grid, 256, 0, c->stream>>>(dst, nsnp, pitch, numindivs, seed, s0, Tm, missing > 0.0 ? 1 : 0, npops, delta);
  EB_CHECK_LAUNCH(c);
  return 0;
}

}  // namespace eb
