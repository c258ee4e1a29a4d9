// capi.cu -- extern "C" entry points of libeigb200.so (declared in include/eigb200.h) and host orchestration.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include "common.cuh"

namespace eb {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace eb

using namespace eb;

extern "C" {

const char* eb_last_error(void) { return g_err; }
int eb_version(void) { return 100; }

int eb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

eb_ctx* eb_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    set_error("eb_create: no CUDA device visible (libeigb200 has no CPU fallback)");
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= n) { set_error("eb_create: device %d out of range (%d visible)", device, n); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { set_error("eb_create: cudaSetDevice(%d) failed", device); return nullptr; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("eb_create: cudaGetDeviceProperties failed"); return nullptr; }
  if (prop.major != 10) {
    set_error("eb_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return nullptr;
  }
  eb_ctx* c = new eb_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("eb_create: stream"); delete c; return nullptr; }
  for (int i = 0; i < 12; i++) cudaEventCreate(&c->ev[i]);
  return c;
}

void eb_destroy(eb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->has_comm && c->grm_epoch > 0 && c->grm_flags.p) peer_grm_wait_idle(c);   // peers may still pull from my receive buffer
  cudaStreamSynchronize(c->stream);
  peer_release(c);                                           // close the IPC mappings of the peers' buffers
  for (auto& reg : c->peer) { for (void* p : reg.graveyard) cudaFree(p); reg.graveyard.clear(); }
  for (int i = 0; i < 12; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (cudaEvent_t e : c->i8_ev) cudaEventDestroy(e);
  cudaStreamDestroy(c->stream);
  delete c;
}

void* eb_stream(eb_ctx* c) { return c ? (void*)c->stream : nullptr; }
int eb_sync(eb_ctx* c) {
  if (!c) return EB_ERR_ARG;
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int64_t eb_launch_count(eb_ctx* c) { return c ? c->launches : 0; }
void eb_reset_launch_count(eb_ctx* c) { if (c) c->launches = 0; }

static int set_store(eb_ctx* c, int64_t nsnp, int numindivs) {
  if (nsnp <= 0 || numindivs <= 0) { set_error("bad shape: nsnp=%lld numindivs=%d", (long long)nsnp, numindivs); return EB_ERR_ARG; }
  c->nsnp = nsnp; c->numindivs = numindivs;
  c->mpad = (nsnp + SNP_PAD - 1) / SNP_PAD * SNP_PAD;
  c->rows_set = false; c->grm_valid = false;
  int rc;
  if ((rc = c->c0_d.ensure(c->mpad)) || (rc = c->c1_d.ensure(c->mpad)) || (rc = c->nmiss_d.ensure(c->mpad)) ||
      (rc = c->used_d.ensure(c->mpad)) || (rc = c->ignore_d.ensure(c->mpad)) || (rc = c->xmean_d.ensure(c->mpad)) ||
      (rc = c->xfancy_d.ensure(c->mpad)) || (rc = c->weight_d.ensure(c->mpad)) || (rc = c->table_d.ensure(c->mpad * 4)) ||
      (rc = c->nused_d.ensure(1)))
    return rc;
  return 0;
}

int eb_upload_packed(eb_ctx* c, const uint8_t* packed, int64_t nsnp, int64_t rlen, int numindivs) {
  if (!c || !packed) { set_error("eb_upload_packed: null argument"); return EB_ERR_ARG; }
  if (rlen * 4 < numindivs) { set_error("eb_upload_packed: rlen %lld too small for %d individuals", (long long)rlen, numindivs); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->raw_own.ensure((size_t)nsnp * rlen))) return rc;
  EB_CUDA(cudaMemcpyAsync(c->raw_own.p, packed, (size_t)nsnp * rlen, cudaMemcpyHostToDevice, c->stream));
  c->raw = c->raw_own.p; c->raw_pitch = rlen;
  return set_store(c, nsnp, numindivs);
}

int eb_upload_packed_rows(eb_ctx* c, const uint8_t* const* rows, int64_t nsnp, int64_t rlen, int numindivs) {
  if (!c || !rows) { set_error("eb_upload_packed_rows: null argument"); return EB_ERR_ARG; }
  if (rlen * 4 < numindivs) { set_error("eb_upload_packed_rows: rlen too small"); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->raw_own.ensure((size_t)nsnp * rlen))) return rc;
  // coalesce runs of rows that are contiguous in host memory (the usual case: one packgenos slab)
  int64_t i = 0;
  while (i < nsnp) {
    int64_t j = i + 1;
    while (j < nsnp && rows[j] == rows[j - 1] + rlen) j++;
    EB_CUDA(cudaMemcpyAsync(c->raw_own.p + (size_t)i * rlen, rows[i], (size_t)(j - i) * rlen, cudaMemcpyHostToDevice, c->stream));
    i = j;
  }
  c->raw = c->raw_own.p; c->raw_pitch = rlen;
  return set_store(c, nsnp, numindivs);
}

int eb_adopt_packed_device(eb_ctx* c, const void* dev, int64_t nsnp, int64_t pitch, int numindivs) {
  if (!c || !dev) { set_error("eb_adopt_packed_device: null argument"); return EB_ERR_ARG; }
  if (pitch * 4 < numindivs) { set_error("eb_adopt_packed_device: pitch too small"); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  c->raw_own.release();
  c->raw = static_cast<const uint8_t*>(dev); c->raw_pitch = pitch;
  return set_store(c, nsnp, numindivs);
}

int eb_synth_packed_device(eb_ctx* c, void* dev, int64_t nsnp, int64_t pitch, int numindivs, uint64_t seed, int64_t s0,
                           double missing, int npops, double delta) {
  if (!c || !dev) { set_error("eb_synth_packed_device: null argument"); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  return launch_synth(c, static_cast<uint8_t*>(dev), nsnp, pitch, numindivs, seed, s0, missing, npops, delta);
}

int eb_set_rows(eb_ctx* c, const int* xindex, int nrows) {
  if (!c) return EB_ERR_ARG;
  if (!c->raw) { set_error("eb_set_rows: no genotype store uploaded"); return EB_ERR_STATE; }
  if (!xindex) nrows = c->numindivs;
  if (nrows <= 0 || nrows > c->numindivs) { set_error("eb_set_rows: nrows=%d out of range", nrows); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  c->xindex_h.resize(nrows);
  for (int i = 0; i < nrows; i++) {
    const int v = xindex ? xindex[i] : i;
    if (v < 0 || v >= c->numindivs || (i > 0 && v <= c->xindex_h[i - 1])) { set_error("eb_set_rows: xindex must be ascending within [0,numindivs)"); return EB_ERR_ARG; }
    c->xindex_h[i] = v;
  }
  c->nrows = nrows;
  c->npad = (nrows + TILE - 1) / TILE * TILE;
  c->wpitch = c->npad / 4;
  int rc;
  if ((rc = c->xindex_d.ensure(nrows))) return rc;
  if ((rc = c->work.ensure((size_t)c->mpad * c->wpitch))) return rc;
  // the synchronous copy also orders the host vector against later reuse
  EB_CUDA(cudaMemcpyAsync(c->xindex_d.p, c->xindex_h.data(), sizeof(int) * nrows, cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaEventRecord(c->ev[0], c->stream));
  if ((rc = launch_gather(c))) return rc;
  EB_CUDA(cudaEventRecord(c->ev[1], c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->tm.gather_ms, c->ev[0], c->ev[1]);
  c->rows_set = true; c->grm_valid = false;
  return 0;
}

static int need_rows(eb_ctx* c, const char* who) {
  if (!c) return EB_ERR_ARG;
  if (!c->rows_set) { set_error("%s: call eb_upload_packed + eb_set_rows first", who); return EB_ERR_STATE; }
  EB_CUDA(cudaSetDevice(c->device));
  return 0;
}

static int stage_opts(eb_ctx* c, const eb_grm_opts* o) {
  if (o->snp_ignore) EB_CUDA(cudaMemcpyAsync(c->ignore_d.p, o->snp_ignore, (size_t)c->nsnp, cudaMemcpyHostToDevice, c->stream));
  if (o->snp_weight) EB_CUDA(cudaMemcpyAsync(c->weight_d.p, o->snp_weight, sizeof(double) * (size_t)c->nsnp, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

static int fetch_snp_outputs(eb_ctx* c, int* c0, int* c1, int* nmiss, uint8_t* used, double* xmean, double* xfancy, int64_t* nused) {
  const size_t m = (size_t)c->nsnp;
  if (c0) EB_CUDA(cudaMemcpyAsync(c0, c->c0_d.p, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
  if (c1) EB_CUDA(cudaMemcpyAsync(c1, c->c1_d.p, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
  if (nmiss) EB_CUDA(cudaMemcpyAsync(nmiss, c->nmiss_d.p, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
  if (used) EB_CUDA(cudaMemcpyAsync(used, c->used_d.p, m, cudaMemcpyDeviceToHost, c->stream));
  if (xmean) EB_CUDA(cudaMemcpyAsync(xmean, c->xmean_d.p, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
  if (xfancy) EB_CUDA(cudaMemcpyAsync(xfancy, c->xfancy_d.p, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
  long long nu = 0;
  EB_CUDA(cudaMemcpyAsync(&nu, c->nused_d.p, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  // the GRM kernel's self-measurement (4 words per CTA)
  std::vector<unsigned long long> prof;
  if (c->grm_grid > 0 && c->grmprof_d.p) {
    prof.resize((size_t)4 * c->grm_grid);
    EB_CUDA(cudaMemcpyAsync(prof.data(), c->grmprof_d.p, sizeof(unsigned long long) * prof.size(), cudaMemcpyDeviceToHost, c->stream));
  }
  EB_CUDA(cudaStreamSynchronize(c->stream));
  c->nused = nu;
  if (nused) *nused = nu;
  if (!prof.empty()) {
    const int g = c->grm_grid;
    std::vector<double> mhz(g);
    std::vector<unsigned long long> sms(g);
    unsigned long long tmin = ~0ull, tmax = 0;
    double dmin = 1e300, dmax = 0.0;
    for (int b = 0; b < g; b++) {
      const unsigned long long t0 = prof[4 * b + 1], t1 = prof[4 * b + 2];
      const double ns = (double)(t1 - t0);
      sms[b] = prof[4 * b];
      mhz[b] = ns > 0 ? (double)prof[4 * b + 3] / ns * 1e3 : 0.0;
      tmin = std::min(tmin, t0); tmax = std::max(tmax, t1);
      dmin = std::min(dmin, ns); dmax = std::max(dmax, ns);
    }
    std::sort(mhz.begin(), mhz.end()); std::sort(sms.begin(), sms.end());
    c->tm.grm_sm_mhz = (float)mhz[g / 2];
    c->tm.grm_sms = (int)(std::unique(sms.begin(), sms.end()) - sms.begin());
    c->tm.grm_cta_min_ms = (float)(dmin * 1e-6); c->tm.grm_cta_max_ms = (float)(dmax * 1e-6);
    c->tm.grm_span_ms = (float)((double)(tmax - tmin) * 1e-6);
  }
  return 0;
}

int eb_snp_counts(eb_ctx* c, int* c0, int* c1, int* nmiss) {
  int rc;
  if ((rc = need_rows(c, "eb_snp_counts"))) return rc;
  eb_grm_opts o = {1, 1, 0, 2147483647, nullptr, nullptr};
  if ((rc = launch_stats(c, &o))) return rc;
  c->grm_valid = false;
  if ((rc = fetch_snp_outputs(c, c0, c1, nmiss, nullptr, nullptr, nullptr, nullptr))) return rc;
  // plain counts: the all-missing sentinel (-1) of getcolxz_binary1 is reported as c0=c1=0, nmiss=nrows
  for (int64_t s = 0; s < c->nsnp; s++) {
    if (nmiss && nmiss[s] < 0) { nmiss[s] = c->nrows; if (c0) c0[s] = 0; if (c1) c1[s] = 0; }
    else if (!nmiss && c0 && c0[s] < 0) { c0[s] = 0; if (c1) c1[s] = 0; }
  }
  return 0;
}

int eb_indiv_valid_counts(eb_ctx* c, const uint8_t* snp_keep, int* nvalid) {
  if (!c || !nvalid) return EB_ERR_ARG;
  if (!c->raw) { set_error("eb_indiv_valid_counts: no genotype store uploaded"); return EB_ERR_STATE; }
  EB_CUDA(cudaSetDevice(c->device));
  DevBuf<int> out;
  int rc;
  if ((rc = out.ensure(c->numindivs))) return rc;
  if (snp_keep) EB_CUDA(cudaMemcpyAsync(c->ignore_d.p, snp_keep, (size_t)c->nsnp, cudaMemcpyHostToDevice, c->stream));
  if ((rc = launch_indiv_counts(c, snp_keep ? c->ignore_d.p : nullptr, out.p))) return rc;
  EB_CUDA(cudaMemcpyAsync(nvalid, out.p, sizeof(int) * c->numindivs, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// per-population genotype-class counts over the current PCA rows (the getrawcol + ddd[k] loops of fstcolyy,
// qpsubs.c:1205-1281, and of the per-population validity passes)
int eb_pop_counts(eb_ctx* c, const int* xtypes, int npops, int* counts) {
  int rc;
  if ((rc = need_rows(c, "eb_pop_counts"))) return rc;
  if (!xtypes || !counts || npops <= 0) { set_error("eb_pop_counts: bad argument"); return EB_ERR_ARG; }
  // population-sorted individual list, each segment padded to a multiple of 16 with -1 (gathered as missing)
  std::vector<int> cnt(npops, 0), seg(npops + 1, 0);
  for (int i = 0; i < c->nrows; i++) if (xtypes[i] >= 0 && xtypes[i] < npops) cnt[xtypes[i]]++;
  for (int k = 0; k < npops; k++) seg[k + 1] = seg[k] + (cnt[k] + 15) / 16;
  const int nwords = std::max(seg[npops], 1);
  const int64_t wp3 = ((int64_t)nwords * 4 + 15) / 16 * 16;
  std::vector<int> list((size_t)wp3 * 4, -1), fill(npops, 0);
  for (int i = 0; i < c->nrows; i++) {
    const int k = xtypes[i];
    if (k < 0 || k >= npops) continue;
    list[(size_t)seg[k] * 16 + fill[k]++] = c->xindex_h[i];
  }
  DevBuf<int> list_d, seg_d, out_d;
  DevBuf<uint8_t> work3;
  const size_t nout = (size_t)c->nsnp * npops * 3;
  if ((rc = list_d.ensure(list.size())) || (rc = seg_d.ensure(npops + 1)) || (rc = out_d.ensure(nout)) ||
      (rc = work3.ensure((size_t)c->mpad * wp3)))
    return rc;
  EB_CUDA(cudaMemcpyAsync(list_d.p, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaMemcpyAsync(seg_d.p, seg.data(), sizeof(int) * (npops + 1), cudaMemcpyHostToDevice, c->stream));
  if ((rc = launch_gather_into(c, list_d.p, (int)list.size(), work3.p, wp3))) return rc;
  if ((rc = launch_pop_counts(c, work3.p, wp3, npops, seg_d.p, out_d.p))) return rc;
  EB_CUDA(cudaMemcpyAsync(counts, out_d.p, sizeof(int) * nout, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static int grm_pass(eb_ctx* c, const eb_grm_opts* opts, int* c0, int* c1, int* nmiss, uint8_t* used, double* xmean, double* xfancy,
                    int64_t* nused_out, bool peer) {
  int rc;
  if ((rc = need_rows(c, "eb_grm"))) return rc;
  if (!opts) { set_error("eb_grm: opts is NULL"); return EB_ERR_ARG; }
  if (c->nrows < 2) { set_error("eb_grm: need at least 2 rows"); return EB_ERR_ARG; }
  if ((rc = stage_opts(c, opts))) return rc;
  // SNPs sharded over several GPUs: receive buffers sized / mapped for this matrix (host collectives only when it grew: normally
  // once, here, before any kernel of the pass is in flight), then everything below is stream-ordered
  if (peer && (rc = peer_grm_setup(c, grm_nsplit_for(c, true)))) return rc;
  EB_CUDA(cudaEventRecord(c->ev[0], c->stream));
  if ((rc = launch_stats(c, opts))) return rc;
  EB_CUDA(cudaEventRecord(c->ev[1], c->stream));
  if (peer && (rc = peer_grm_wait_idle(c))) return rc;           // nobody still pulls the previous pass out of my receive buffer
  if ((rc = grm_accumulate(c, !peer, peer))) return rc;
  if (peer && (rc = peer_grm_finalize(c, c->nsplit))) return rc; // signal / wait / reduce / signal / wait / gather (peer.cu)
  if ((rc = fetch_snp_outputs(c, c0, c1, nmiss, used, xmean, xfancy, nused_out))) return rc;
  cudaEventElapsedTime(&c->tm.stats_ms, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&c->tm.grm_ms, c->ev[2], c->ev[3]);
  c->tm.i8_gemm_ms = 0.f;
  if (c->tm.grm_method == 2) {
    for (int i = 0; i < c->i8_nlaunch; i++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->i8_ev[2 * i], c->i8_ev[2 * i + 1]);
      c->tm.i8_gemm_ms += ms;
    }
    c->tm.grm_launches = c->i8_nlaunch;
  }
  if (c->tm.grm_method == 2 && c->i8_sync_h[2]) {
    set_error("grm (i8): %u pass synchronisations timed out (clusters not co-resident?)", c->i8_sync_h[2]);
    return EB_ERR_STATE;
  }
  c->tm.exchange_wait_ms = 0.f;
  if (peer) {
    long long tot = 0;
    if ((rc = peer_grm_collect(c, &tot))) return rc;
    c->nused_total = tot;
    if (nused_out) *nused_out = tot;
    cudaEventElapsedTime(&c->tm.exchange_wait_ms, c->ev[5], c->ev[6]);   // spinning for the slowest rank's tiles
  } else {
    c->nused_total = c->nused;
  }
  cudaEventElapsedTime(&c->tm.finalize_ms, c->ev[3], c->ev[4]);
  c->grm_valid = true;
  c->grm_collective = peer;
  c->grm_popfill = false;
  return 0;
}

int eb_grm_partial(eb_ctx* c, const eb_grm_opts* opts, int* c0, int* c1, int* nmiss, uint8_t* used, double* xmean, double* xfancy,
                   int64_t* nused_out) {
  return grm_pass(c, opts, c0, c1, nmiss, used, xmean, xfancy, nused_out, false);
}

void* eb_grm_device_ptr(eb_ctx* c, int64_t* ld, int64_t* n) {
  if (!c || !c->xtx.p) return nullptr;
  if (ld) *ld = c->npad;
  if (n) *n = c->nrows;
  return c->xtx.p;
}

__global__ void scale_copy_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int n, double s) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
  if (col < n) dst[(size_t)row * n + col] = src[(size_t)row * lds + col] * s;
}

int eb_grm_finish(eb_ctx* c, double* y_out, double* XTX_host) {
  if (!c || !c->grm_valid) { set_error("eb_grm_finish: no GRM resident"); return EB_ERR_STATE; }
  EB_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = grm_trace(c))) return rc;
  if (std::isnan(c->y)) { set_error("bad XTX matrix"); return EB_ERR_NUMERIC; }                 // smartpca.c:1231
  if (c->y <= 0.0) { set_error("XTX has zero trace (perhaps no data)"); return EB_ERR_NUMERIC; }  // smartpca.c:1234
  if (y_out) *y_out = c->y;
  if (XTX_host) {
    // XTX / y, computed as XTX * (1/y) like vst(XTX, XTX, 1.0/y, ...) at smartpca.c:1236
    DevBuf<double> tmp;
    const int n = c->nrows;
    if ((rc = tmp.ensure((size_t)n * n))) return rc;
    dim3 grid((n + 255) / 256, n);
    scale_copy_kernel<<<grid, 256, 0, c->stream>>>(c->xtx.p, c->npad, tmp.p, n, 1.0 / c->y);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaMemcpyAsync(XTX_host, tmp.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, c->stream));
    EB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int eb_grm(eb_ctx* c, const eb_grm_opts* opts, int* c0, int* c1, int* nmiss, uint8_t* used, double* xmean, double* xfancy,
           double* y_out, int64_t* nused_out, double* XTX_host) {
  int rc;
  if ((rc = grm_pass(c, opts, c0, c1, nmiss, used, xmean, xfancy, nused_out, c && c->has_comm))) return rc;
  return eb_grm_finish(c, y_out, XTX_host);
}

// ---- usepopsformissing: YES.  getcolxz (smartpca.c:3129-3216) fills a missing genotype with the mean of the individual's population
// before fvadjust, which makes the columns arbitrary FP64 values: the reference then takes its dense path (smartpca.c:995-1014,
// domult_increment_normal).  Here the per-population sums, the fill, the normalisation and the drop rule run on the device
// (popfill_stats_kernel), the FP64 columns are materialised 1024 SNPs at a time in device memory (popfill_cols_kernel) and go
// through the dense tensor-core SYRK -- nothing but the per-SNP outputs crosses PCIe.
int eb_grm_popfill(eb_ctx* c, const eb_grm_opts* opts, const int* xtypes, int npops, int* c0, int* c1, int* nmiss, uint8_t* used,
                   double* xmean, double* xfancy, double* y_out, int64_t* nused_out, double* XTX_host) {
  int rc;
  if ((rc = need_rows(c, "eb_grm_popfill"))) return rc;
  if (!opts || !xtypes || npops < 1) { set_error("eb_grm_popfill: bad argument"); return EB_ERR_ARG; }
  if (c->has_comm) { set_error("eb_grm_popfill: not available on a sharded context (the dense path is not collective)"); return EB_ERR_STATE; }
  if (c->nrows < 2) { set_error("eb_grm_popfill: need at least 2 rows"); return EB_ERR_ARG; }
  if ((rc = stage_opts(c, opts))) return rc;
  DevBuf<int> xt_d, nmiss_after;
  DevBuf<double> fill;
  const size_t plane = (size_t)c->npad * c->npad;
  if ((rc = xt_d.ensure(c->npad)) || (rc = nmiss_after.ensure(c->mpad)) || (rc = fill.ensure((size_t)c->mpad * npops)) ||
      (rc = c->partial.ensure(plane)) || (rc = c->xtx.ensure(plane)) || (rc = c->trace_d.ensure(1)) ||
      (rc = c->dense_blk.ensure((size_t)1024 * c->npad)))
    return rc;
  std::vector<int> xt(c->npad, -1);
  for (int i = 0; i < c->nrows; i++) xt[i] = xtypes[i];
  EB_CUDA(cudaMemcpyAsync(xt_d.p, xt.data(), sizeof(int) * c->npad, cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaMemsetAsync(c->partial.p, 0, sizeof(double) * plane, c->stream));
  EB_CUDA(cudaEventRecord(c->ev[0], c->stream));
  if ((rc = launch_popfill_stats(c, opts, xt_d.p, npops, nmiss_after.p, fill.p))) return rc;
  EB_CUDA(cudaEventRecord(c->ev[1], c->stream));
  EB_CUDA(cudaEventRecord(c->ev[2], c->stream));
  for (int64_t s0 = 0; s0 < c->mpad; s0 += 1024) {
    const int nb = (int)std::min<int64_t>(1024, c->mpad - s0);       // mpad is a multiple of 128: whole k-chunks of 32
    if ((rc = launch_popfill_cols(c, s0, nb, xt_d.p, npops, fill.p, c->dense_blk.p))) return rc;
    if ((rc = launch_syrk_lower_add(c, c->partial.p, c->npad, c->npad, c->dense_blk.p, c->npad, nb))) return rc;
  }
  EB_CUDA(cudaEventRecord(c->ev[3], c->stream));
  c->nsplit = 1; c->grm_grid = 0;
  if ((rc = grm_dense_finalize(c))) return rc;
  EB_CUDA(cudaEventRecord(c->ev[4], c->stream));
  // per-SNP outputs: nmiss is the count AFTER the fill (what getcolxz returns); the device keeps the raw count for later passes
  if (nmiss) EB_CUDA(cudaMemcpyAsync(nmiss, nmiss_after.p, sizeof(int) * (size_t)c->nsnp, cudaMemcpyDeviceToHost, c->stream));
  if ((rc = fetch_snp_outputs(c, c0, c1, nullptr, used, xmean, xfancy, nused_out))) return rc;
  cudaEventElapsedTime(&c->tm.stats_ms, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&c->tm.grm_ms, c->ev[2], c->ev[3]);
  cudaEventElapsedTime(&c->tm.finalize_ms, c->ev[3], c->ev[4]);
  c->nused_total = c->nused;
  c->grm_valid = true; c->grm_collective = false; c->grm_popfill = true;
  return eb_grm_finish(c, y_out, XTX_host);
}

// ---- dense path: getcolxz + domult_increment_normal (smartpca.c:3129-3216, 3531-3561), used by the reference when
// usepopsformissing / ldregress make the columns arbitrary FP64 values.  The host keeps producing the normalised
// columns (tblock rows); the library accumulates XTX += sum_s x_s x_s^T on the FP64 tensor cores.
int eb_grm_dense_begin(eb_ctx* c, int nrows) {
  if (!c || nrows < 2) { set_error("eb_grm_dense_begin: bad argument"); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  c->nrows = nrows;
  c->npad = (nrows + TILE - 1) / TILE * TILE;
  c->rows_set = false;      // no packed working matrix behind this GRM
  c->grm_valid = false;
  int rc;
  const size_t plane = (size_t)c->npad * c->npad;
  if ((rc = c->partial.ensure(plane)) || (rc = c->xtx.ensure(plane)) || (rc = c->trace_d.ensure(1)) ||
      (rc = c->dense_blk.ensure((size_t)1024 * c->npad)))
    return rc;
  EB_CUDA(cudaMemsetAsync(c->partial.p, 0, sizeof(double) * plane, c->stream));
  c->dense_open = true;
  c->nsplit = 1;
  return 0;
}

int eb_grm_dense_add(eb_ctx* c, const double* tblock, int nblock) {
  if (!c || !c->dense_open) { set_error("eb_grm_dense_add: call eb_grm_dense_begin first"); return EB_ERR_STATE; }
  if (!tblock || nblock < 0) { set_error("eb_grm_dense_add: bad argument"); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  int rc;
  // the reference refuses columns that are not centred (domult_increment_normal's ycheck, smartpca.c:3545-3549)
  for (int b = 0; b < nblock; b++) {
    double s = 0.0;
    for (int i = 0; i < c->nrows; i++) s += tblock[(size_t)b * c->nrows + i];
    if (fabs(s) > .00001) { set_error("bad ycheck"); return EB_ERR_NUMERIC; }
  }
  for (int done = 0; done < nblock; done += 1024) {
    const int nb = std::min(1024, nblock - done), kr = (nb + 31) / 32 * 32;
    EB_CUDA(cudaMemsetAsync(c->dense_blk.p, 0, sizeof(double) * (size_t)kr * c->npad, c->stream));
    EB_CUDA(cudaMemcpy2DAsync(c->dense_blk.p, sizeof(double) * c->npad, tblock + (size_t)done * c->nrows, sizeof(double) * c->nrows,
                              sizeof(double) * c->nrows, nb, cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_syrk_lower_add(c, c->partial.p, c->npad, c->npad, c->dense_blk.p, c->npad, kr))) return rc;
    EB_CUDA(cudaStreamSynchronize(c->stream));      // the caller may reuse tblock
  }
  return 0;
}

int eb_grm_dense_end(eb_ctx* c, double* y_out, double* XTX_host) {
  if (!c || !c->dense_open) { set_error("eb_grm_dense_end: call eb_grm_dense_begin first"); return EB_ERR_STATE; }
  EB_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = grm_dense_finalize(c))) return rc;
  c->dense_open = false;
  c->grm_valid = true;
  c->grm_collective = false;
  return eb_grm_finish(c, y_out, XTX_host);
}

int eb_eig(eb_ctx* c, int nvec, double* lambda, double* evecs) {
  if (!c || !c->grm_valid || c->y <= 0.0) { set_error("eb_eig: no normalised GRM resident (call eb_grm first)"); return EB_ERR_STATE; }
  EB_CUDA(cudaSetDevice(c->device));
  return eig_resident(c, c->xtx.p, c->npad, c->nrows, 1.0 / c->y, nvec, lambda, evecs, c->has_comm && c->grm_collective);
}

int eb_eigvecs(eb_ctx* c, const double* mat, double* evals, double* evecs, int n, int nvec) {
  if (!c || !mat || n <= 0) { set_error("eb_eigvecs: bad argument"); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  DevBuf<double> A;
  int rc;
  const int64_t lda = ((int64_t)n + 15) & ~15ll;      // 128-byte rows: legal TMA pitch for any n
  if ((rc = A.ensure((size_t)n * lda))) return rc;
  EB_CUDA(cudaMemcpy2DAsync(A.p, sizeof(double) * lda, mat, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice, c->stream));
  return eig_resident(c, A.p, lda, n, 1.0, nvec, evals, evecs);
}

// testing aid: two-stage tridiagonalisation of a host matrix; d[n], e[n] (unscaled), band[n*128] or NULL
int eb_debug_tridiag(eb_ctx* c, const double* mat, int n, double* d, double* e, double* band) {
  if (!c || !mat || n < 3) { set_error("eb_debug_tridiag: bad argument"); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  DevBuf<double> A, de;
  int rc;
  const int64_t lda = ((int64_t)n + 15) & ~15ll;
  if ((rc = A.ensure((size_t)n * lda)) || (rc = de.ensure((size_t)2 * n))) return rc;
  EB_CUDA(cudaMemcpy2DAsync(A.p, sizeof(double) * lda, mat, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice, c->stream));
  c->dbg_band_h = band;
  rc = two_stage_tridiag(c, A.p, lda, n, de.p, de.p + n);
  c->dbg_band_h = nullptr;
  if (rc) return rc;
  EB_CUDA(cudaMemcpyAsync(d, de.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaMemcpyAsync(e, de.p + n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- drop-in symbols with the reference's own names and contract (include/eigsubs.h:6-7; eigsubs.c:21,39): mat row-major
// n x n, preserved; eigenvalues descending; evecs row i = eigenvector i; failure = message on stderr + exit(1) like
// eigx.c:109-116.  They run on a process-wide context on the current device.
static eb_ctx* g_dropin_ctx = nullptr;
static eb_ctx* dropin_ctx() {
  if (!g_dropin_ctx) g_dropin_ctx = eb_create(-1);
  if (!g_dropin_ctx) { fprintf(stderr, "eigvecs (libeigb200): %s\n", eb_last_error()); exit(1); }
  return g_dropin_ctx;
}
void eigvecs(double* mat, double* evals, double* evecs, int n) {
  eb_ctx* c = dropin_ctx();
  // the reference fills all n vectors (eigsubs.c:39-55): so does this, at any n (full-basis path of the one-stage solver)
  if (eb_eigvecs(c, mat, evals, evecs, n, n) != 0) { fprintf(stderr, "eigvecs (libeigb200): %s\n", eb_last_error()); exit(1); }
}
void eigvals(double* mat, double* evals, int n) {
  eb_ctx* c = dropin_ctx();
  if (eb_eigvecs(c, mat, evals, nullptr, n, 0) != 0) { fprintf(stderr, "eigvals (libeigb200): %s\n", eb_last_error()); exit(1); }
}

// ---- fastmode under the reference's own name: kjg_fpca (include/kjg_fpca.h:22, kjg_fpca.c:24) after the hand-over that
// setgval (gval.c:31-87) does through file statics.  setgval itself takes the reference's SNP / Indiv structs, so it lives in the
// few lines of glue a maintainer compiles against the reference headers (integration/eb_gval.c: setgval -> eb_setgval_packed);
// everything behind it is plain pointers.  Failure = message on stderr + exit(1), as kjg_fpca.c:26-29 does.
static struct { bool set; int fancynorm, altnormstyle; long seed; } g_gval = {false, 1, 1, 0};
int eb_setgval_packed(const uint8_t* const* snp_pbuff, int64_t ncols, int64_t rlen, int numindivs, const int* xindex, int nrows,
                      int fancynorm, int altnormstyle, long seed, uint8_t* mono_out) {
  eb_ctx* c = dropin_ctx();
  int rc;
  for (int i = 1; i < nrows; i++)
    if (xindex[i] < xindex[i - 1]) { fprintf(stderr, "xindex not sorted\n"); exit(1); }     // gval.c:49-54
  if ((rc = eb_upload_packed_rows(c, snp_pbuff, ncols, rlen, numindivs))) return rc;
  if ((rc = eb_set_rows(c, xindex, nrows))) return rc;
  if (mono_out) {
    // side effect of setgval: SNPs with min(n0, n1) == 0 over the PCA rows get ignore = YES (gval.c:80-82); the row stays in the table
    std::vector<int> c0(ncols), c1(ncols);
    if ((rc = eb_snp_counts(c, c0.data(), c1.data(), nullptr))) return rc;
    for (int64_t s = 0; s < ncols; s++) mono_out[s] = std::min(c0[s], c1[s]) == 0 ? 1 : 0;
  }
  g_gval.set = true; g_gval.fancynorm = fancynorm; g_gval.altnormstyle = altnormstyle; g_gval.seed = seed;
  return 0;
}
void eb_unsetgval(void) { g_gval.set = false; }
void kjg_fpca(size_t K, size_t L, size_t I, double* eval, double* evec) {
  if (K >= L) { fprintf(stderr, "kjg_fpca (libeigb200): K >= L\n"); exit(1); }
  if (I == 0) { fprintf(stderr, "kjg_fpca (libeigb200): I == 0\n"); exit(1); }
  if (!g_gval.set) { fprintf(stderr, "kjg_fpca (libeigb200): setgval has not handed over the genotypes\n"); exit(1); }
  if (eb_fpca(dropin_ctx(), g_gval.fancynorm, g_gval.altnormstyle, K, L, I, g_gval.seed, eval, evec) != 0) {
    fprintf(stderr, "kjg_fpca (libeigb200): %s\n", eb_last_error());
    exit(1);
  }
}

int eb_set_option(eb_ctx* c, const char* key, int value) {
  if (!c || !key) return EB_ERR_ARG;
  if (!strcmp(key, "eig_method")) { c->opt_eig_method = value; return 0; }
  if (!strcmp(key, "two_stage_min")) { c->opt_two_stage_min = value; return 0; }
  if (!strcmp(key, "dist_min")) { c->opt_dist_min = value; return 0; }
  if (!strcmp(key, "eig_vectors")) { c->opt_eig_vectors = value; return 0; }
  if (!strcmp(key, "grm_method")) { c->opt_grm_method = value; return 0; }
  if (!strcmp(key, "pg_method")) { c->opt_pg_method = value; return 0; }
  if (!strcmp(key, "pg_i8_min")) { c->opt_pg_i8_min = value; return 0; }
  if (!strcmp(key, "i8_min")) { c->opt_i8_min = value; return 0; }
  if (!strcmp(key, "i8_slices")) { c->opt_i8_slices = value; return 0; }
  if (!strcmp(key, "i8_slab")) { c->opt_i8_slab = value; return 0; }
  if (!strcmp(key, "i8_pair")) { c->opt_i8_pair = value; return 0; }
  if (!strcmp(key, "i8_sync")) { c->opt_i8_sync = value; return 0; }
  set_error("eb_set_option: unknown key '%s'", key);
  return EB_ERR_ARG;
}

// ridoutlier, smartsubs.c:18-93: same operation order as the reference so that |z| > thresh decisions agree bit for bit
int eb_ridoutlier(const double* evecs, int n, int neigs, double thresh, int outliermode, int* badlist, int* vecno, double* score) {
  if (outliermode > 1 || n < 3) return 0;
  std::vector<double> ww(n), w2(n);
  std::vector<int> vbad(n, 0);
  for (int j = 0; j < n; j++) vecno[j] = -1;
  for (int i = 0; i < neigs; i++) {
    for (int j = 0; j < n; j++) ww[j] = evecs[(size_t)i * n + j];
    if (outliermode == 0) {
      double y1 = 0.0, y2 = 0.0;
      for (int j = 0; j < n; j++) y1 += ww[j];
      y1 /= (double)n;
      for (int j = 0; j < n; j++) ww[j] = ww[j] + (-y1);
      for (int j = 0; j < n; j++) y2 += ww[j] * ww[j];
      y2 = y2 / (double)n;
      y2 = sqrt(y2);
      const double r = 1.0 / y2;
      for (int j = 0; j < n; j++) ww[j] = ww[j] * r;
      for (int j = 0; j < n; j++)
        if (fabs(ww[j]) > thresh) { vbad[j] = 1; if (vecno[j] < 0) { vecno[j] = i; score[j] = ww[j]; } }
    } else {
      for (int j = 0; j < n; j++) {
        const double yy = ww[j];
        ww[j] = 0;
        double y1 = 0.0, y2 = 0.0;
        for (int k = 0; k < n; k++) y1 += ww[k];
        y1 /= (double)(n - 1);
        for (int k = 0; k < n; k++) w2[k] = ww[k] + (-y1);
        w2[j] = 0;
        for (int k = 0; k < n; k++) y2 += w2[k] * w2[k];
        y2 = sqrt(y2 / (double)n);
        double zz = yy - y1;
        zz /= y2;
        if (fabs(zz) > thresh) { vbad[j] = 1; if (vecno[j] < 0) { vecno[j] = i; score[j] = zz; } }
        ww[j] = yy;
      }
    }
  }
  int nbad = 0;
  for (int j = 0; j < n; j++) if (vbad[j]) badlist[nbad++] = j;
  return nbad;
}

int eb_pca_full(eb_ctx* c, const eb_pca_opts* o, int* xindex_io, int nrows, double* lambda, double* evecs, uint8_t* snp_used,
                double* xmean, double* xfancy, int* removed_index, int* removed_iter, int* removed_vecno, double* removed_score,
                eb_pca_result* res) {
  if (!c || !o || !xindex_io || !lambda || !res) { set_error("eb_pca_full: null argument"); return EB_ERR_ARG; }
  const double t_start = now_s();
  memset(res, 0, sizeof(*res));
  int rc;
  std::vector<uint8_t> ignore(c->nsnp, 0), used(c->nsnp);
  if (o->grm.snp_ignore) memcpy(ignore.data(), o->grm.snp_ignore, c->nsnp);
  std::vector<int> xi(xindex_io, xindex_io + nrows);
  std::vector<int> bad(nrows), vecno(nrows);
  std::vector<double> score(nrows);
  const int numoutiter = o->numoutliter >= 1 ? o->numoutliter + 1 : 1;      // smartpca.c:1028-1040
  const int outmode = o->numoutliter >= 1 ? o->outliermode : 2;
  int nremoved = 0;
  for (int iter = 1; iter <= numoutiter; iter++) {
    if ((rc = eb_set_rows(c, xi.data(), (int)xi.size()))) return rc;
    eb_grm_opts g = o->grm;
    g.snp_ignore = ignore.data();
    double y = 0; int64_t nused = 0;
    double t0 = now_s();
    if ((rc = eb_grm(c, &g, nullptr, nullptr, nullptr, used.data(), xmean, xfancy, &y, &nused, nullptr))) return rc;
    res->secs_grm += now_s() - t0;
    // SNPs dropped in a pass stay ignored in later passes (cupt->ignore = YES, smartpca.c:1136; loadsnpx 1085)
    for (int64_t s = 0; s < c->nsnp; s++) if (!used[s]) ignore[s] = 1;
    const int n = (int)xi.size();
    const int nv = std::max(std::min(o->numeigs, n), std::min(std::min(o->numoutleigs, n - 1), n));
    std::vector<double> ev((size_t)std::max(nv, 1) * n);
    // Large n: the outlier passes only need the leading vectors (smartpca.c:1250); the full spectrum (.eval, Tracy-Widom)
    // is computed once, on the pass that turns out to be the last.  Small n: one-stage solver, everything each pass.
    const bool two = eig_uses_two_stage(c, n, nv);
    // a pass that cannot be followed by another one (no outlier iterations asked for, or all of them used up) takes spectrum and
    // vectors from ONE two-stage solve (vectors by back-transformation through the kept reflectors)
    const bool last_for_sure = iter == numoutiter;
    t0 = now_s();
    if ((rc = eb_eig(c, nv, (two && !last_for_sure) ? nullptr : lambda, ev.data()))) return rc;
    res->secs_eig += now_s() - t0;
    res->niter = iter; res->y = y; res->nused = nused; res->nrows_final = n;
    const int keep = std::min(o->numeigs, n);
    if (evecs) memcpy(evecs, ev.data(), sizeof(double) * (size_t)keep * n);
    if (snp_used) memcpy(snp_used, used.data(), c->nsnp);
    int nbad = 0;
    if (iter <= o->numoutliter) {                                          // last pass skips outliers, smartpca.c:1246
      const int neigs = std::min(o->numoutleigs, n - 1);                   // smartpca.c:1249
      nbad = eb_ridoutlier(ev.data(), n, neigs, o->outlthresh, outmode, bad.data(), vecno.data(), score.data());
    }
    if (c->has_comm) {
      // every rank holds the same reduced GRM bit for bit, so the decisions must agree; a mismatch would desynchronise
      // the collective passes that follow -- fail loudly instead
      std::vector<long long> all((size_t)2 * c->comm.world);
      long long mine[2] = {nbad, 0};
      for (int b = 0; b < nbad; b++) mine[1] = mine[1] * 1000003ll + bad[b];
      if ((rc = peer_allgather_host(c, mine, all.data(), sizeof(mine)))) return rc;
      for (int r = 0; r < c->comm.world; r++)
        if (all[2 * r] != mine[0] || all[2 * r + 1] != mine[1]) { set_error("eb_pca_full: ranks disagree on the outlier list (pass %d)", iter); return EB_ERR_STATE; }
    }
    if (nbad == 0) {
      if (two && !last_for_sure) {
        t0 = now_s();
        if ((rc = eb_eig(c, 0, lambda, nullptr))) return rc;
        res->secs_eig += now_s() - t0;
      }
      break;
    }
    std::vector<char> kill(n, 0);
    for (int b = 0; b < nbad; b++) {
      const int j = bad[b];
      kill[j] = 1;
      if (removed_index) removed_index[nremoved] = xi[j];
      if (removed_iter) removed_iter[nremoved] = iter;
      if (removed_vecno) removed_vecno[nremoved] = vecno[j];
      if (removed_score) removed_score[nremoved] = score[j];
      nremoved++;
    }
    std::vector<int> nxt;
    for (int j = 0; j < n; j++) if (!kill[j]) nxt.push_back(xi[j]);
    xi.swap(nxt);
  }
  res->nremoved = nremoved;
  for (size_t i = 0; i < xi.size(); i++) xindex_io[i] = xi[i];
  res->secs_total = now_s() - t_start;
  return 0;
}

int eb_fpca(eb_ctx* c, int fancynorm, int altnormstyle, size_t K, size_t L, size_t I, long seed, double* eval, double* evec) {
  int rc;
  if ((rc = need_rows(c, "eb_fpca"))) return rc;
  if (K >= L || I == 0) { set_error("eb_fpca: need K < L and I > 0 (kjg_fpca.c:26-29)"); return EB_ERR_ARG; }
  return fpca_run(c, fancynorm, altnormstyle, K, L, I, seed, eval, evec);
}

int eb_project(eb_ctx* c, const double* evecs, int numeigs, double* ffvecs, double* fxvecs, double* fxscal) {
  int rc;
  if ((rc = need_rows(c, "eb_project"))) return rc;
  if (!c->grm_valid) { set_error("eb_project: run eb_grm first (needs the per-SNP normalisation)"); return EB_ERR_STATE; }
  if (c->grm_popfill) { set_error("eb_project: the resident GRM was built with usepopsformissing (eb_grm_popfill); its projection passes stay on the host"); return EB_ERR_STATE; }
  return project_run(c, evecs, numeigs, ffvecs, fxvecs, fxscal);
}

int eb_lsqproj(eb_ctx* c, const int* indiv, int nindiv, const double* ffvecs, const double* fxscal, int numeigs, double* acoeffs,
               double* bcoeffs, int* nvalid, uint8_t* ok) {
  int rc;
  if ((rc = need_rows(c, "eb_lsqproj"))) return rc;
  if (!c->grm_valid) { set_error("eb_lsqproj: run eb_grm first (needs the per-SNP normalisation)"); return EB_ERR_STATE; }
  if (c->grm_popfill) { set_error("eb_lsqproj: the resident GRM was built with usepopsformissing (eb_grm_popfill); its projection passes stay on the host"); return EB_ERR_STATE; }
  if (!ffvecs || !fxscal || nindiv <= 0) { set_error("eb_lsqproj: bad argument"); return EB_ERR_ARG; }
  std::vector<int> all;
  if (!indiv) { all.resize(nindiv); for (int i = 0; i < nindiv; i++) all[i] = i; indiv = all.data(); }
  return lsqproj_run(c, indiv, nindiv, ffvecs, fxscal, numeigs, acoeffs, bcoeffs, nvalid, ok);
}

// smartpca.c:1440-1564: setfvecs -> SNP loadings -> sample projections -> lsqproj -> seteigscale -> scaled coordinates
int eb_evec_coords(eb_ctx* c, const double* evecs, int numeigs, const int* indiv, int nindiv, double* coords, double* eigscale, uint8_t* ok) {
  int rc;
  if ((rc = need_rows(c, "eb_evec_coords"))) return rc;
  if (!evecs || !coords || nindiv <= 0) { set_error("eb_evec_coords: bad argument"); return EB_ERR_ARG; }
  std::vector<int> all;
  if (!indiv) { all.resize(nindiv); for (int i = 0; i < nindiv; i++) all[i] = i; indiv = all.data(); }
  // position of every PCA row inside the list (sqz, smartpca.c:3664-3684, needs all of them)
  std::vector<int> pos(c->numindivs, -1);
  for (int i = 0; i < nindiv; i++) {
    if (indiv[i] < 0 || indiv[i] >= c->numindivs) { set_error("eb_evec_coords: individual index out of range"); return EB_ERR_ARG; }
    pos[indiv[i]] = i;
  }
  for (int r = 0; r < c->nrows; r++)
    if (pos[c->xindex_h[r]] < 0) { set_error("eb_evec_coords: PCA row %d (individual %d) is not in the output list", r, c->xindex_h[r]); return EB_ERR_ARG; }
  const int64_t m = c->nsnp;
  std::vector<double> ff((size_t)numeigs * m), sc(numeigs), a((size_t)numeigs * nindiv), b((size_t)numeigs * nindiv);
  std::vector<uint8_t> okv(nindiv);
  if ((rc = eb_project(c, evecs, numeigs, ff.data(), nullptr, sc.data()))) return rc;
  if ((rc = eb_lsqproj(c, indiv, nindiv, ff.data(), sc.data(), numeigs, a.data(), b.data(), nullptr, okv.data()))) return rc;
  for (int j = 0; j < numeigs; j++) {
    double ab = 0.0, aa = 0.0;
    for (int r = 0; r < c->nrows; r++) {
      const int q = pos[c->xindex_h[r]];
      ab += a[(size_t)j * nindiv + q] * b[(size_t)j * nindiv + q];
      aa += a[(size_t)j * nindiv + q] * a[(size_t)j * nindiv + q];
    }
    const double es = ab / aa;
    if (eigscale) eigscale[j] = es;
    for (int q = 0; q < nindiv; q++) coords[(size_t)j * nindiv + q] = a[(size_t)j * nindiv + q] * es;
  }
  if (ok) memcpy(ok, okv.data(), nindiv);
  return 0;
}

int eb_shrink_coords(eb_ctx* c, int numeigs, int newshrink, double* coords, double* lambda_out, uint8_t* ok) {
  int rc;
  if ((rc = need_rows(c, "eb_shrink_coords"))) return rc;
  if (!c->grm_valid || c->y <= 0.0) { set_error("eb_shrink_coords: run eb_grm first (needs the resident GRM and the per-SNP normalisation)"); return EB_ERR_STATE; }
  if (!coords || !lambda_out) { set_error("eb_shrink_coords: null argument"); return EB_ERR_ARG; }
  if (c->grm_popfill) { set_error("eb_shrink_coords: not available after eb_grm_popfill"); return EB_ERR_STATE; }
  return shrink_run(c, numeigs, newshrink, coords, lambda_out, ok);
}

// testing aid: C = op(A) op(B)^T through the general FP64 tensor-core GEMM
int eb_debug_gemm(eb_ctx* c, int a_km, int b_kn, const double* A, const double* B, double* C, int M, int N, int K) {
  if (!c || !A || !B || !C) return EB_ERR_ARG;
  EB_CUDA(cudaSetDevice(c->device));
  const int64_t lda = ((a_km ? M : K) + 1) & ~1ll, ldb = ((b_kn ? N : K) + 1) & ~1ll, ldc = (N + 1) & ~1ll;
  const int ra = a_km ? K : M, rb = b_kn ? K : N;
  DevBuf<double> Ad, Bd, Cd;
  int rc;
  if ((rc = Ad.ensure((size_t)ra * lda)) || (rc = Bd.ensure((size_t)rb * ldb)) || (rc = Cd.ensure((size_t)M * ldc))) return rc;
  EB_CUDA(cudaMemsetAsync(Ad.p, 0, sizeof(double) * ra * lda, c->stream));
  EB_CUDA(cudaMemsetAsync(Bd.p, 0, sizeof(double) * rb * ldb, c->stream));
  EB_CUDA(cudaMemcpy2DAsync(Ad.p, sizeof(double) * lda, A, sizeof(double) * (a_km ? M : K), sizeof(double) * (a_km ? M : K), ra, cudaMemcpyHostToDevice, c->stream));
  EB_CUDA(cudaMemcpy2DAsync(Bd.p, sizeof(double) * ldb, B, sizeof(double) * (b_kn ? N : K), sizeof(double) * (b_kn ? N : K), rb, cudaMemcpyHostToDevice, c->stream));
  if ((rc = launch_gemm(c, a_km != 0, b_kn != 0, Ad.p, lda, Bd.p, ldb, Cd.p, ldc, M, N, K))) return rc;
  EB_CUDA(cudaMemcpy2DAsync(C, sizeof(double) * N, Cd.p, sizeof(double) * ldc, sizeof(double) * N, M, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- Tracy-Widom statistics (host arithmetic on the spectrum): the loop smartpca.c:1336-1366 / twstats.c:58-77 around
// dotwcalc (statsubs.c:1680-1725) and twnorm (statsubs.c:1655-1677, Johnstone 2001).  The reference re-sums the tail of the
// spectrum for every eigenvalue (O(m^2): 2.5e9 flops at m = 50,000); suffix sums make it O(m).
int eb_numgtz(const double* lambda, int n) {        // statsubs.c:1727-1742
  int num = 0;
  for (int k = 0; k < n; k++) if (lambda[k] > .000001) num++;
  return num;
}

static double tw_norm(double lam, double p, double n) {
  if (n < 0.0 || p < 0.0) return -10.0;
  if (n < p) return tw_norm(lam, n, p);
  const double y1 = sqrt(n - 1) + sqrt(p);
  const double mu = y1 * y1;
  const double y2 = (1.0 / sqrt(n - 1)) + 1.0 / sqrt(p);
  const double phi = y1 * pow(y2, 1.0 / 3.0);
  return (lam - mu) / phi;
}

int eb_tw_stats(const double* lambda, int m, double znval, int minm, double* tw, double* zn) {
  if (!lambda || m < 0 || !tw || !zn) { set_error("eb_tw_stats: bad argument"); return EB_ERR_ARG; }
  std::vector<double> s1((size_t)m + 1, 0.0), s2((size_t)m + 1, 0.0);
  for (int i = m - 1; i >= 0; i--) { s1[i] = s1[i + 1] + lambda[i]; s2[i] = s2[i + 1] + lambda[i] * lambda[i]; }
  for (int i = 0; i < m; i++) {
    const int mm = m - i;
    const double lsum = s1[i];
    tw[i] = zn[i] = -1.0;
    if (mm < minm || lsum <= 0.0) continue;
    const double tm = (double)mm, y = tm / lsum;
    if (znval > 0.0) {
      zn[i] = znval;
      tw[i] = tw_norm(lambda[i] * y * znval, tm, znval);
    } else {
      const double bot = s2[i] * y * y - tm;       // sum (lambda_j * y)^2 - m
      const double z = (double)mm * (double)(mm + 2) / bot;
      zn[i] = z;
      tw[i] = tw_norm(lambda[i] * y * z, tm, z);
    }
  }
  return 0;
}

// Tracy-Widom right tail from the caller's table (POPGEN/twtable: x, tail, density; what `twstats -t twtable` and the table
// compiled into the reference hold): twtail -> gettw (statsubs.c:1590-1604, 1811-1860), firstgtx / fgtx (1744-1766), cinterp
// (cubic Hermite, statsubs.c), including the reference's behaviour beyond the table (below: tail 1; above: the
// Margetis-Edelman DENSITY formula twdensx is what gettw returns as the tail, statsubs.c:1837-1841).
static int tw_fgtx(const double* tab, int lo, int hi, double val) {
  if (val >= tab[hi]) return hi + 1;
  if (val < tab[lo]) return lo;
  const int k = (lo + hi) / 2;
  if (val <= tab[k]) return tw_fgtx(tab, lo + 1, k, val);
  return tw_fgtx(tab, k, hi - 1, val);
}
double eb_tw_tail(double x, const double* tab_x, const double* tab_tail, const double* tab_pdf, int n) {
  if (!tab_x || !tab_tail || !tab_pdf || n < 2) return -1.0;
  const int k = tw_fgtx(tab_x, 0, n - 1, x);
  if (k <= 0) return 1.0;
  if (k >= n) {
    if (x <= 0.0) return 0.0;
    const double sqrt_pi = 1.0 / 0.5641895835477562869480795;       // SQRT_PI, include/statsubs.h:12,17
    const double lbot = log(sqrt_pi * 4.0);
    return exp(-0.25 * log(x) + -2.0 * pow(x, 1.5) / 3.0 - lbot);
  }
  const double x0 = tab_x[k - 1], x1 = tab_x[k], f0 = tab_tail[k - 1], f0p = -tab_pdf[k - 1], f1 = tab_tail[k], f1p = -tab_pdf[k];
  const double inc = x1 - x0, yval = (x - x0) / inc;
  const double a0 = f0, a1 = f0p * inc, cc0 = f1 - (a0 + a1), cc1 = f1p * inc - a1, a2 = 3 * cc0 - cc1, a3 = cc1 - 2 * cc0;
  double f = a3;
  f *= yval; f += a2; f *= yval; f += a1; f *= yval; f += a0;
  return f;
}

int64_t eb_snp_used_count(eb_ctx* c) { return c ? c->nused : 0; }

int eb_get_timings(eb_ctx* c, eb_timings* t) {
  if (!c || !t) return EB_ERR_ARG;
  *t = c->tm; t->nsplit = c->nsplit;
  return 0;
}

int eb_microbench_fp64(eb_ctx* c, double* dmma, double* dfma) {
  if (!c) return EB_ERR_ARG;
  EB_CUDA(cudaSetDevice(c->device));
  double a = 0, b = 0;
  int rc = microbench_fp64(c, &a, &b);
  if (dmma) *dmma = a;
  if (dfma) *dfma = b;
  return rc;
}

}  // extern "C"
