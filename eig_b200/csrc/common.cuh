// common.cuh -- shared declarations for libeigb200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/eigb200.h"

namespace eb {

void set_error(const char* fmt, ...);

#define EB_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      eb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
      return EB_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define EB_CHECK_LAUNCH(ctx)                                                                 \
  do {                                                                                       \
    (ctx)->launches++;                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                     \
    if (e_ != cudaSuccess) {                                                                 \
      eb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return EB_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  std::vector<void*>* defer = nullptr;   // exported to peers: the old allocation is freed only after they unmapped it
  int ensure(size_t count) {
    if (count <= n && p) return 0;
    if (p) { if (defer) defer->push_back(p); else cudaFree(p); }
    p = nullptr; n = 0;
    if (count == 0) return 0;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); return EB_ERR_NOMEM; }
    n = count;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

// multi-GPU exchange over peer memory (peer.cu)
constexpr int EB_MAX_WORLD = 16;
enum { PEER_SLOT_PARTIAL = 0, PEER_SLOT_FLAGS = 1, PEER_SLOT_A = 2, PEER_SLOT_B = 3, PEER_SLOT_C = 4, PEER_SLOT_T = 5, PEER_SLOT_W = 6, PEER_SLOTS = 7 };
// Where grm_syrk_kernel stores a finished 128 x 128 tile when the SNPs are sharded over `world` GPUs: lower-triangle tile t
// belongs to rank t % world, and every rank stores its partial tile STRAIGHT INTO THE OWNER'S receive buffer over NVLink
// (slot [t / world][rank * nsplit + chunk], a dense 128 x 128 block), so the reduce-scatter traffic overlaps the DMMA work
// of the tiles that follow.  world <= 1: tiles go to the local split-K planes.
struct GrmPush {
  double* recv[EB_MAX_WORLD];
  int world, rank;
};
constexpr int GRM_FLAG_WORDS = 256;   // per rank: flags[3][16], mailbox[2][16][2], error word (unsigned long long each)
struct PeerRecord {             // what a rank publishes about one exported allocation
  unsigned char handle[64];     // cudaIpcMemHandle_t
  uint64_t ptr, bytes;          // device address in the owner's process, size of the allocation
  int64_t pid;
  int32_t device, aux;          // aux: slot-specific (GRM planes: nsplit; XTX: npad; all-reduce: element count)
};
struct PeerRegion {
  std::vector<PeerRecord> rec;
  std::vector<void*> mapped;    // this process's address of every rank's buffer
  std::vector<bool> opened;     // mapped through cudaIpcOpenMemHandle (must be closed)
  std::vector<void*> graveyard; // my superseded allocations of this slot, freed once every peer has re-mapped
};

constexpr int TILE = 128;       // GRM tile edge in individuals (32 packed bytes)
constexpr int KT = 128;         // SNPs per pipeline stage of the GRM kernel
constexpr int SNP_PAD = KT;     // working matrix SNP count is padded to this (pad rows are all-missing)

}  // namespace eb

struct eb_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;

  // raw slab as uploaded (reference layout, pitch = rlen or caller pitch)
  eb::DevBuf<uint8_t> raw_own;
  const uint8_t* raw = nullptr;
  int64_t nsnp = 0, raw_pitch = 0;
  int numindivs = 0;

  // working matrix: selected rows only, SNP-major, pitch = npad/4 bytes, pad genotypes = 3
  eb::DevBuf<uint8_t> work;
  eb::DevBuf<int> xindex_d;
  eb::DevBuf<int> wsrc_d;         // per 32-bit word of the working row: raw byte offset of a 16-individual run | -1 mixed | -2 pad
  std::vector<int> xindex_h;
  int nrows = 0;        // selected individuals
  int npad = 0;         // nrows rounded up to TILE
  int64_t mpad = 0;     // nsnp rounded up to SNP_PAD
  int64_t wpitch = 0;   // bytes per SNP row of the working matrix
  bool rows_set = false;

  // per-SNP device arrays (length mpad)
  eb::DevBuf<int> c0_d, c1_d, nmiss_d;
  eb::DevBuf<uint8_t> used_d, ignore_d;
  eb::DevBuf<double> xmean_d, xfancy_d, weight_d;
  eb::DevBuf<double> table_d;     // [mpad][4]: cc0,cc1,cc2,0 (zero rows for unused SNPs)
  eb::DevBuf<long long> nused_d;  // 1

  // GRM
  eb::DevBuf<double> partial;     // nsplit * npad * npad
  eb::DevBuf<double> xtx;         // npad * npad, full symmetric, UNNORMALISED
  eb::DevBuf<double> trace_d;     // 1
  eb::DevBuf<int> workctr_d;      // 1
  eb::DevBuf<unsigned long long> grmprof_d;   // [num_sms][4]: smid, globaltimer at CTA start / end, SM cycles (grm_syrk_kernel)
  int grm_grid = 0;               // CTAs of the last grm_syrk_kernel launch
  int dt_slots = 0;               // resident CTAs per SM of the DMMA tile kernels on this context's device (eig2_gemm.cu)
  eb::DevBuf<double> dense_blk;   // dense path staging: [1024][npad]
  bool dense_open = false;
  int nsplit = 1;
  bool grm_valid = false;
  bool grm_popfill = false;       // the resident GRM came from eb_grm_popfill: the packed-table projection passes do not apply
  bool grm_collective = false;    // the resident GRM is the reduced matrix of a sharded pass: identical on every rank of the communicator
  double y = 0.0;                 // trace/(nrows-1)
  int64_t nused = 0;              // SNPs of THIS shard that entered XTX
  int64_t nused_total = 0;        // over all shards (== nused without a communicator)

  // eigensolver workspace
  eb::DevBuf<double> eigA;        // npad*npad working copy (reflectors end up in its lower part)
  eb::DevBuf<double> eigw;        // misc vectors
  eb::DevBuf<double> eigV, eigW;  // panels
  eb::DevBuf<double> lambda_d, zvec_d;
  int64_t zvec_ld = 0;             // row pitch of zvec_d after the last solve (n, or n rounded up to even for the full basis)
  eb::DevBuf<double> eig2w, chfsiw;   // two-stage reduction / subspace-iteration workspaces (eig2_kernels.cu)
  eb::DevBuf<double> eigQ2, eigT;     // kept reflectors for the eigenvector back-transformation: bulge chasing (n^2 / 2), stage-1 T factors
  double* eig_q2 = nullptr;           // where the stage-2 reflectors of the last two_stage_tridiag(keep_q) are (eigQ2 or the dead GRM accumulator)
  int eig_npanels = 0;
  int opt_eig_vectors = 0;            // two-stage path, spectrum + vectors: 0 = back-transformation of the tridiagonal eigenvectors, 1 = subspace iteration
  std::vector<double> ritz;        // scaled Ritz values of the last subspace iteration
  double* dbg_band_h = nullptr;    // eb_debug_tridiag only
  int opt_eig_method = 0;          // 0 auto, 1 one-stage (dsytrd-style), 2 two-stage + subspace iteration
  int opt_two_stage_min = 1536;    // auto: n at which the two-stage path takes over
  int opt_dist_min = 8192;         // collective solves: n from which the band reduction / subspace iteration are split over the ranks

  // multi-GPU (peer.cu)
  eb_comm comm = {0, 1, nullptr, nullptr, nullptr};
  bool has_comm = false;
  eb::PeerRegion peer[eb::PEER_SLOTS];
  eb::DevBuf<double> peer_scratch;
  eb::DevBuf<double> chfsi_sum;             // collective subspace iteration: the 64 x n product block that is summed over the ranks
  unsigned long long ar_epoch = 0;          // stream-ordered all-reduces done (all ranks in lockstep)
  eb::DevBuf<double> grm_recv;              // sharded GRM: [owned tiles][world * nsplit][128 x 128] partial tiles pushed by every rank
  eb::DevBuf<unsigned long long> grm_flags; // sharded GRM: device-side flags / mailbox written by the peers (GRM_FLAG_WORDS)
  unsigned long long grm_epoch = 0;         // one per sharded GRM pass (all ranks in lockstep)
  size_t grm_recv_need = 0;                 // doubles the current geometry needs in grm_recv
  int grm_geom_npad = 0, grm_geom_nsplit = 0;   // geometry the peers have agreed on
  bool grm_host_sync = false;               // two ranks are contexts of ONE process on ONE device: waits go through the host (see peer.cu)
  eb::DevBuf<double> pg_part;         // split-K planes of the packed products (fpca_kernels.cu)
  eb::DevBuf<double> fpG, fpB, fpS;   // fastmode buffers that take part in an exchange (persistent: peers map them)

  // exact integer tensor-core GRM (grm_i8.cu)
  eb::DevBuf<uint8_t> i8_ops;       // byte operands of one SNP slab: A[nseg][rows][npad], B[nseg * nsl][rows][npad]
  eb::DevBuf<uint8_t> i8_flag;      // per 128-SNP block: a used SNP with a missing genotype
  eb::DevBuf<double> i8_coef;       // [2][mpad]: a b and a^2 of the SNPs without missing genotypes (rank-one terms)
  eb::DevBuf<double> i8_r;          // rank-one partial sums per SNP chunk, r[npad], sum a^2
  eb::DevBuf<long long> i8_prep;    // largest weight exponent, flagged blocks, used SNPs, exponent sum
  eb::DevBuf<int> i8_tiles;         // (nb, mb) tile order
  std::vector<cudaEvent_t> i8_ev;   // start / stop of every integer GEMM launch of the last pass
  int i8_nlaunch = 0;
  unsigned int i8_sync_h[4] = {0, 0, 0, 0};   // [2]: pass-synchronisation time-outs of the last pass (0 in a healthy run)
  int opt_grm_method = 0;           // 0 auto (integer path from opt_i8_min rows), 1 FP64 DMMA, 2 integer tensor cores
  int opt_i8_min = 4096;
  int opt_i8_slices = 0;            // 0 auto (52 bits below the typical weight), else the digit count (1..9)
  int opt_i8_slab = 0;              // 0 = as many SNPs per slab as memory allows, else the cap (tests: several slabs)
  int opt_i8_pair = 1;              // 1 = CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles), 0 = one CTA per 128 x 256 tile
  int opt_i8_sync = 0;              // pair kernel: passes a cluster may run ahead of the slowest one (-1: no synchronisation)

  // packed x skinny products on the integer tensor cores (pg_i8.cu)
  eb::DevBuf<uint8_t> pgi_digits;   // 7-bit digit rows of the skinny operand: [1 or 2][columns x 8][K]
  eb::DevBuf<double> pgi_scale;     // per-column maxima / scales
  int opt_pg_method = 0;            // 0 auto (integer path from pg_i8_min rows and K), 1 FP64 DMMA, 2 integer tensor cores
  int opt_pg_i8_min = 2048;

  eb_timings tm = {};
  cudaEvent_t ev[12] = {};
};

namespace eb {
// pack_kernels.cu
int launch_gather(eb_ctx* c);
int launch_gather_into(eb_ctx* c, const int* list_d, int nlist, uint8_t* dst, int64_t wpitch);
int launch_stats(eb_ctx* c, const eb_grm_opts* o);
int launch_pop_counts(eb_ctx* c, const uint8_t* work3, int64_t wp3, int npops, const int* seg_word0_d, int* out_d);
int launch_indiv_counts(eb_ctx* c, const uint8_t* keep_d, int* out_d);
int launch_popfill_stats(eb_ctx* c, const eb_grm_opts* o, const int* xt_d, int npops, int* nmiss_after_d, double* fill_d);
int launch_popfill_cols(eb_ctx* c, int64_t s0, int nb, const int* xt_d, int npops, const double* fill_d, double* blk);
int launch_synth(eb_ctx* c, uint8_t* dst, int64_t nsnp, int64_t pitch, int numindivs, uint64_t seed, int64_t s0,
                 double missing, int npops, double delta);
// grm_kernel.cu
int grm_accumulate(eb_ctx* c, bool finalize_local = true, bool push = false);   // work+table -> split-K planes [-> xtx] | -> owners' receive buffers
int grm_nsplit_for(const eb_ctx* c, bool sharded);
int grm_trace(eb_ctx* c);        // recompute trace_d / y from xtx
// pg_i8.cu
int pg_i8_launch(eb_ctx* c, int mode, const uint8_t* work, int64_t wpitch, int npad, const double* table, const double* In_t, int64_t ld_in,
                 double* Out_t, int64_t ld_out, int ncols, double oscale);
int pg_i8_slice_wide(eb_ctx* c, const double* In_t, int64_t ld_in, int64_t kvalid, int64_t kpad, int ncols, uint8_t* digits, double* colscale);
int pg_i8_rows_wide(eb_ctx* c, const uint8_t* work, int64_t wpitch, int npad, const double* table, const uint8_t* digits, const double* colscale,
                    int ncols, int64_t s0, int nb, double* out, int64_t ld_r);
// grm_i8.cu
bool grm_use_i8(const eb_ctx* c);
int grm_accumulate_i8(eb_ctx* c, bool finalize_local, bool push);
int microbench_fp64(eb_ctx* c, double* dmma, double* dfma);
// eig_kernels.cu
int eig_resident(eb_ctx* c, const double* A_d, int64_t lda, int n, double scale, int nvec, double* lambda_h, double* evecs_h, bool collective = false);
bool eig_uses_two_stage(const eb_ctx* c, int n, int nvec);
// eig2_gemm.cu
int launch_syrk_lower_add(eb_ctx* c, double* A, int64_t lda, int n, const double* T, int64_t ldt, int krows);
int launch_gemm(eb_ctx* c, bool a_km, bool b_kn, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                int M, int N, int K, double alpha = 1.0, double beta = 0.0);
// grm_kernel.cu (dense path)
int grm_dense_finalize(eb_ctx* c);
// eig2_kernels.cu
int two_stage_tridiag(eb_ctx* c, double* A, int64_t lda, int n, double* d, double* e, bool collective = false, bool keep_q = false);
int two_stage_backtransform(eb_ctx* c, const double* A, int64_t lda, int n, int nvec, double* Z, int64_t ldz);
int chfsi_top(eb_ctx* c, const double* A, int64_t lda, int n, int nvec, double* theta_h, double* vec_d, int* iters_out, int* matvecs_out,
              double lo0 = 0.0, bool collective = false);
// fpca_kernels.cu
int fpca_run(eb_ctx* c, int fancynorm, int altnormstyle, size_t K, size_t L, size_t I, long seed, double* eval, double* evec);
int project_run(eb_ctx* c, const double* evecs, int numeigs, double* ffvecs, double* fxvecs, double* fxscal);
int shrink_run(eb_ctx* c, int k, int newshrink, double* coords, double* lambda_out, uint8_t* ok_out);
int lsqproj_run(eb_ctx* c, const int* indiv, int nlist, const double* ffvecs, const double* fxscal, int k, double* acoeffs, double* bcoeffs,
                int* nvalid, uint8_t* ok);
// peer.cu
int peer_allgather_host(eb_ctx* c, const void* src, void* dst, int64_t bytes);
int peer_exchange(eb_ctx* c, int slot, void* local, size_t bytes, int aux);
int peer_grm_setup(eb_ctx* c, int nsplit);      // collective only when the matrix outgrew the mapped buffers
int peer_grm_push_args(eb_ctx* c, GrmPush* out);
int peer_grm_wait_idle(eb_ctx* c);              // stream-ordered: peers finished pulling the previous pass
int peer_grm_finalize(eb_ctx* c, int nsplit);   // stream-ordered: signal, wait, reduce, signal, wait, gather, signal
int peer_grm_collect(eb_ctx* c, long long* nused_total);   // after a stream sync: mailbox + error word
int peer_allreduce(eb_ctx* c, int slot, double* buf, size_t alloc_doubles, int64_t count);
int peer_allreduce_stream(eb_ctx* c, int slot, double* buf, size_t alloc_doubles, int64_t count, bool exchange, int64_t offset = 0);   // stream-ordered (device flags)
int peer_flags_setup(eb_ctx* c);
int peer_allreduce_any(eb_ctx* c, double* buf, int64_t count);
int peer_sum_host(eb_ctx* c, double* v, int count);
void peer_release(eb_ctx* c);
int peer_bury(eb_ctx* c);        // collective: barrier, then free superseded exported allocations

}  // namespace eb
