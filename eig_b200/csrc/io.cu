// io.cu -- the steps either side of the hot path: PACKEDANCESTRYMAP genotype file -> device slab, and
// the .eval / .evec / grm text writers with the reference's exact formats.
//
//   eb_hash_ids              hasharr / hashit, admutils.c:651-682 (order dependent, int arithmetic with wrap-around)
//   eb_packed_file_header    the "GENO %7d %7d %x %x" record, mcio.c:2402, 2812-2826
//   eb_upload_packed_file    inpack, mcio.c:2769-2879: the reference reads the file in 1 GiB read() chunks into packgenos
//                            and then walks every genotype through checkxval (mcio.c:1606-1618, male X hets -> missing).
//                            Here the file streams through two pinned staging buffers straight into HBM (read of chunk
//                            i+1 overlaps the H2D copy of chunk i) and the X-het rule runs as one kernel on the slab.
//   eb_write_eval/evec/grm   smartpca.c:1425-1437, 1570-1591, 3770-3805
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <cmath>
#include "common.cuh"

using namespace eb;

namespace eb {

// male X heterozygotes -> missing (checkxval, mcio.c:1606-1618); one thread per byte of an X-chromosome SNP row
__global__ void __launch_bounds__(256) xhet_mask_kernel(uint8_t* __restrict__ slab, int64_t pitch, int64_t nsnp, int numindivs,
                                                        const uint8_t* __restrict__ snp_is_x, const uint8_t* __restrict__ male_bits) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;          // byte within the row = individuals 4b..4b+3
  if (b * 4 >= numindivs) return;
  const uint8_t mb = male_bits[b];                              // 2-bit field 3 where the individual is male
  if (!mb) return;
  for (int64_t s = blockIdx.y; s < nsnp; s += gridDim.y) {
    if (!snp_is_x[s]) continue;
    uint8_t v = slab[s * pitch + b];
    // het = code 01: low bit set, high bit clear
    const uint8_t het = (uint8_t)((v & 0x55) & ~((v >> 1) & 0x55));
    const uint8_t hit = (uint8_t)((het | (het << 1)) & mb);
    if (hit) slab[s * pitch + b] = v | hit;                     // 01 -> 11 (missing)
  }
}

}  // namespace eb

extern "C" {

int eb_hash_ids(const char* const* ids, int n) {
  uint32_t hash = 0;
  for (int i = 0; i < n; i++) {
    uint32_t th = 0;
    for (const char* p = ids[i]; *p; p++) { th *= 23u; th += (uint32_t)(int)*p; }
    hash *= 17u;
    hash ^= th;
  }
  return (int)hash;
}

int eb_packed_file_header(const char* path, int* nind, int* nsnp, int* ihash, int* shash, int64_t* rlen, int64_t* file_bytes) {
  if (!path) { set_error("eb_packed_file_header: null path"); return EB_ERR_ARG; }
  const int fd = open(path, O_RDONLY);
  if (fd < 0) { set_error("(ispack) bad open %s: %s", path, strerror(errno)); return EB_ERR_ARG; }
  char buf[64] = {0};
  const ssize_t t = read(fd, buf, 48);
  struct stat st;
  fstat(fd, &st);
  close(fd);
  if (t < 48) { set_error("(inpack) bad read %s", path); return EB_ERR_ARG; }
  int xi = 0, xs = 0; unsigned int hi = 0, hs = 0;
  if (strncmp(buf, "GENO", 4) != 0 || sscanf(buf, "GENO %d %d %x %x", &xi, &xs, &hi, &hs) != 4) {
    set_error("%s is not a PACKEDANCESTRYMAP genotype file (no GENO header)", path);
    return EB_ERR_ARG;
  }
  // rlen = max(48, ceil(nind/4)), mcio.c:2788-2790
  const int64_t rl = std::max<int64_t>(48, ((int64_t)xi * 2 + 7) / 8);
  if (nind) *nind = xi;
  if (nsnp) *nsnp = xs;
  if (ihash) *ihash = (int)hi;
  if (shash) *shash = (int)hs;
  if (rlen) *rlen = rl;
  if (file_bytes) *file_bytes = (int64_t)st.st_size;
  return 0;
}

int eb_upload_packed_file(eb_ctx* c, const char* path, int numindivs, int64_t nsnp, int check_hash, int ihash, int shash,
                          const uint8_t* snp_is_x, const uint8_t* indiv_is_male) {
  if (!c || !path) { set_error("eb_upload_packed_file: null argument"); return EB_ERR_ARG; }
  int xi, xs, hi, hs; int64_t rl, fbytes;
  int rc;
  if ((rc = eb_packed_file_header(path, &xi, &xs, &hi, &hs, &rl, &fbytes))) return rc;
  // the reference's own consistency checks and messages (mcio.c:2812-2826)
  if (xi != numindivs) { set_error("OOPS number of individuals %d != %d in input files", numindivs, xi); return EB_ERR_ARG; }
  if (xs != nsnp) { set_error("OOPS number of SNPs %lld != %d in input file: %s", (long long)nsnp, xs, path); return EB_ERR_ARG; }
  if (check_hash) {
    if (hi != ihash) { set_error("OOPS indiv file has changed since genotype file was created"); return EB_ERR_ARG; }
    if (hs != shash) { set_error("OOPS snp file has changed since genotype file was created"); return EB_ERR_ARG; }
  }
  const int64_t packlen = rl * nsnp;
  if (fbytes < rl + packlen) { set_error("(inpack) bad data read (length mismatch) %lld %lld", (long long)(fbytes - rl), (long long)packlen); return EB_ERR_ARG; }
  EB_CUDA(cudaSetDevice(c->device));
  if ((rc = c->raw_own.ensure((size_t)packlen))) return rc;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) { set_error("(ispack) bad open %s: %s", path, strerror(errno)); return EB_ERR_ARG; }
  // two pinned staging buffers: read() of chunk i+1 overlaps the H2D copy of chunk i
  const size_t CH = 64u << 20;
  uint8_t* stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2];
  for (int i = 0; i < 2; i++) {
    if (cudaMallocHost((void**)&stage[i], CH) != cudaSuccess) { close(fd); set_error("cudaMallocHost(staging) failed"); return EB_ERR_NOMEM; }
    cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
  }
  int64_t off = 0; int cur = 0; rc = 0;
  bool used[2] = {false, false};
  while (off < packlen && rc == 0) {
    const size_t want = (size_t)std::min<int64_t>(CH, packlen - off);
    if (used[cur]) cudaEventSynchronize(done[cur]);
    size_t got = 0;
    while (got < want) {
      const ssize_t t = pread(fd, stage[cur] + got, want - got, (off_t)(rl + off + got));
      if (t <= 0) { set_error("(inpack) bad data read at offset %lld", (long long)(off + got)); rc = EB_ERR_ARG; break; }
      got += (size_t)t;
    }
    if (rc) break;
    if (cudaMemcpyAsync(c->raw_own.p + off, stage[cur], want, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { set_error("H2D copy failed"); rc = EB_ERR_CUDA; break; }
    cudaEventRecord(done[cur], c->stream);
    used[cur] = true;
    off += (int64_t)want; cur ^= 1;
  }
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; i++) { cudaFreeHost(stage[i]); cudaEventDestroy(done[i]); }
  close(fd);
  if (rc) return rc;
  c->raw = c->raw_own.p; c->raw_pitch = rl;
  if (snp_is_x && indiv_is_male) {
    // male mask per packed byte (2-bit fields, MSB-first like the genotypes)
    std::vector<uint8_t> mb((size_t)rl, 0);
    for (int k = 0; k < numindivs; k++) if (indiv_is_male[k]) mb[k >> 2] |= (uint8_t)(3u << ((3 - (k & 3)) << 1));
    DevBuf<uint8_t> mb_d, sx_d;
    if ((rc = mb_d.ensure((size_t)rl)) || (rc = sx_d.ensure((size_t)nsnp))) return rc;
    EB_CUDA(cudaMemcpyAsync(mb_d.p, mb.data(), (size_t)rl, cudaMemcpyHostToDevice, c->stream));
    EB_CUDA(cudaMemcpyAsync(sx_d.p, snp_is_x, (size_t)nsnp, cudaMemcpyHostToDevice, c->stream));
    dim3 grid((unsigned)((rl + 255) / 256), (unsigned)std::min<int64_t>(nsnp, 32768));
    xhet_mask_kernel<<<grid, 256, 0, c->stream>>>(c->raw_own.p, rl, nsnp, numindivs, sx_d.p, mb_d.p);
    EB_CHECK_LAUNCH(c);
    EB_CUDA(cudaStreamSynchronize(c->stream));
  }
  c->nsnp = nsnp; c->numindivs = numindivs;
  c->mpad = (nsnp + SNP_PAD - 1) / SNP_PAD * SNP_PAD;
  c->rows_set = false; c->grm_valid = false;
  if ((rc = c->c0_d.ensure(c->mpad)) || (rc = c->c1_d.ensure(c->mpad)) || (rc = c->nmiss_d.ensure(c->mpad)) ||
      (rc = c->used_d.ensure(c->mpad)) || (rc = c->ignore_d.ensure(c->mpad)) || (rc = c->xmean_d.ensure(c->mpad)) ||
      (rc = c->xfancy_d.ensure(c->mpad)) || (rc = c->weight_d.ensure(c->mpad)) || (rc = c->table_d.ensure(c->mpad * 4)) ||
      (rc = c->nused_d.ensure(1)))
    return rc;
  return 0;
}

// copy of the resident raw slab back to the host (tests; also lets the shim keep SNP.pbuff valid after the X-het rule)
int eb_download_packed(eb_ctx* c, uint8_t* out) {
  if (!c || !c->raw || !out) { set_error("eb_download_packed: no genotype store"); return EB_ERR_STATE; }
  EB_CUDA(cudaSetDevice(c->device));
  EB_CUDA(cudaMemcpy2DAsync(out, (size_t)c->raw_pitch, c->raw, (size_t)c->raw_pitch, (size_t)c->raw_pitch, (size_t)c->nsnp, cudaMemcpyDeviceToHost, c->stream));
  EB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------------- writers (host)
int eb_write_eval(const char* path, const double* lambda, int n) {
  FILE* f = fopen(path, "w");
  if (!f) { set_error("can't open file %s of type w", path); return EB_ERR_ARG; }
  for (int j = 0; j < n; j++) fprintf(f, "%12.6f\n", lambda[j]);               // smartpca.c:1428
  fclose(f);
  return 0;
}

int eb_write_evec(const char* path, const double* lambda, int numeigs, const char* const* ids, const char* const* groups,
                  const double* coords, int nout, int hiprec) {
  FILE* f = fopen(path, "w");
  if (!f) { set_error("can't open file %s of type w", path); return EB_ERR_ARG; }
  fprintf(f, "%20s ", "#eigvals:");                                           // smartpca.c:1433-1437
  for (int j = 0; j < numeigs; j++) fprintf(f, "%9.3f ", lambda[j]);
  fprintf(f, "\n");
  for (int i = 0; i < nout; i++) {                                             // smartpca.c:1574-1591
    fprintf(f, "%20s ", ids[i]);
    for (int j = 0; j < numeigs; j++) {
      const double y = coords[(size_t)j * nout + i];
      if (hiprec) fprintf(f, "%12.6f  ", y); else fprintf(f, "%10.4f  ", y);
    }
    fprintf(f, "%15s\n", groups[i]);
  }
  fclose(f);
  return 0;
}

// text GRM, smartpca.c:3770-3805: "a b numsnps value" over the lower triangle, scaled to mean diagonal 1
int eb_write_grm(const char* path, const double* XTX, int nrows, int numsnps) {
  FILE* f = fopen(path, "w");
  if (!f) { set_error("can't open file %s of type w", path); return EB_ERR_ARG; }
  double tr = 0.0;
  for (int a = 0; a < nrows; a++) tr += XTX[(size_t)a * nrows + a];
  const double recip = ((double)nrows) / tr;
  for (int a = 0; a < nrows; a++)
    for (int b = 0; b <= a; b++) fprintf(f, "%d %d %d %0.6f\n", a + 1, b + 1, numsnps, XTX[(size_t)a * nrows + b] * recip);
  fclose(f);
  return 0;
}

// grmbinary: YES -- dumpgrmbin, smartpca.c:3704-3766 (GCTA layout): <prefix>.N.bin holds the SNP count as a 4-byte int for every
// lower-triangle entry, <prefix>.bin the entries XTX[a][b] / (trace / nrows) as 4-byte floats, rows a = 0.., b <= a.
int eb_write_grm_bin(const char* prefix, const double* XTX, int nrows, int numsnps) {
  if (!prefix || !XTX || nrows <= 0) { set_error("eb_write_grm_bin: bad argument"); return EB_ERR_ARG; }
  const size_t numout = (size_t)nrows * ((size_t)nrows + 1) / 2;
  std::string pn = std::string(prefix) + ".N.bin", pb = std::string(prefix) + ".bin";
  FILE* f = fopen(pn.c_str(), "wb");
  if (!f) { set_error("open failed for %s", pn.c_str()); return EB_ERR_ARG; }
  {
    std::vector<int32_t> buf(std::min<size_t>(numout, 1 << 20), (int32_t)numsnps);
    for (size_t done = 0; done < numout; done += buf.size())
      if (fwrite(buf.data(), 4, std::min(buf.size(), numout - done), f) == 0) { fclose(f); set_error("(outpack) bad write"); return EB_ERR_ARG; }
  }
  fclose(f);
  f = fopen(pb.c_str(), "wb");
  if (!f) { set_error("open failed for %s", pb.c_str()); return EB_ERR_ARG; }
  double tr = 0.0;
  for (int a = 0; a < nrows; a++) tr += XTX[(size_t)a * nrows + a];
  const double y_norm = tr / (double)nrows;
  std::vector<float> row((size_t)nrows);
  for (int a = 0; a < nrows; a++) {
    for (int b = 0; b <= a; b++) row[b] = (float)(XTX[(size_t)a * nrows + b] / y_norm);
    if (fwrite(row.data(), 4, (size_t)a + 1, f) != (size_t)a + 1) { fclose(f); set_error("(outpack) bad write"); return EB_ERR_ARG; }
  }
  fclose(f);
  return 0;
}

}  // extern "C"
