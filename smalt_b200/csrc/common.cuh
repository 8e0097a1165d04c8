// common.cuh - shared declarations of the sm_100a kernels behind include/smalt_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <vector>
#include "../../include/smalt_b200.h"

namespace smb {

// 8x8 substitution matrix over the 3-bit alphabet + affine gap costs
// (score.c:138-173; gap costs positive as returned by scoreGetProfile, score.c:682-683).
// Passed to kernels by value (lives in the constant bank of the launch parameters).
struct Scoring {
  int match, mismatch, gap_init, gap_ext;
  signed char S[64];  // S[ref*8 + read]
};

// Where a kernel reads sequences from.
struct SeqSrc {
  const uint8_t *arena;     // concatenated code bytes (reads, explicit windows)
  const uint32_t *packed;   // 3-bit packed reference, 10 bases / word (sequence.c:1360-1424)
  uint64_t packed_nbases;
};

// base i of the packed reference: bits 3*(9 - i%10) of word i/10
__device__ __forceinline__ uint32_t packed_base(const uint32_t *__restrict__ w, uint64_t i) {
  const uint64_t wi = i / 10u;
  const uint32_t r = (uint32_t)(i - wi * 10u);
  return (__ldg(w + wi) >> (3u * (9u - r))) & 7u;
}

__device__ __forceinline__ uint32_t ref_base(const SeqSrc &s, bool packed, uint64_t off, uint32_t i) {
  return packed ? packed_base(s.packed, off + i) : (uint32_t)(__ldg(s.arena + off + i) & 7u);
}

// base j of the profiled sequence: the read, or its reverse complement
// (complement of the 2-bit codes, everything else unchanged - seqFastqAppendSegment
// with reverse+codec, sequence.c; codtab_complement sequence.c:305)
__device__ __forceinline__ uint32_t read_base(const uint8_t *__restrict__ arena, uint64_t off,
                                              uint32_t len, bool rc, uint32_t j) {
  uint32_t c = __ldg(arena + off + (rc ? (len - 1u - j) : j)) & 7u;
  if (rc && c < 4u) c = 3u - c;
  return c;
}

// Opt-in to large dynamic shared memory, once per DEVICE (the attribute is per device; contexts of
// several GPUs can live in one process and worker threads race here harmlessly).
template <class F>
inline cudaError_t ensure_dyn_smem(F *func, int bytes, std::atomic<unsigned long long> &done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
};

// ---- K2 ----
constexpr int SW_SLOT_LONG2 = 17, SW_SLOTS = 18;
struct SwPlan {
  std::vector<int> order;   // task indices grouped by columns-per-lane class
  int count[SW_SLOTS], start[SW_SLOTS]; // [1..8] 32-bit classes, [9..16] paired 16-bit classes, [17] long reads paired
  int max_grid;
  uint32_t bstride;         // rows of the boundary strips (multi-block reads), 0 if none
  size_t strip_bytes;
};
void plan_sw(const smb_sw_task *h_tasks, int ntasks, int sm_count, const Scoring &sc, SwPlan &plan);
cudaError_t launch_sw_score(const Scoring &sc, const SeqSrc &src, const smb_sw_task *d_tasks,
                            const SwPlan &plan, int *d_counters, const int *d_order, void *d_strips,
                            int32_t *d_scores, int32_t *d_errs, cudaStream_t st, int *nlaunch);

struct BandOut {        // device buffers of smb_band_align_batch
  smb_ali_result *results;  // [ntasks * max_res]
  uint32_t *nres;           // [ntasks]
  uint8_t *diff;            // per-task arenas, task i at diff_off[i], capacity diff_cap[i]
  int32_t *errs;            // [ntasks]
  unsigned long long *cells;
  uint32_t *dused;          // [ntasks] DiffStr bytes written by the task (may be nullptr)
};

// ---- K3 output compaction (compact.cu) ----
struct CompactTotals {
  unsigned long long nresults, ndiff;
  int capacity_flag, pad;
  unsigned long long ncig;   // CIGAR text bytes (output stage on)
};
struct GatherCigar;   // cigar.cuh
int compact_tiles(int n);
cudaError_t launch_compact_scan(const uint32_t *nres, const uint32_t *dused, const int32_t *errs, int n,
                                unsigned long long *tile_res, unsigned long long *tile_diff,
                                CompactTotals *tot, uint32_t *first_result,
                                unsigned long long *diff_first, cudaStream_t st, int *nlaunch,
                                const uint32_t *cig = nullptr, unsigned long long *tile_cig = nullptr,
                                unsigned long long *cig_first = nullptr);
cudaError_t launch_scan_counts(const uint32_t *in, int n, unsigned long long *out, unsigned long long *tile,
                               cudaStream_t st, int *nlaunch);
cudaError_t launch_compact_gather(const smb_ali_result *slots, const uint32_t *nres, const uint8_t *diff_slots,
                                  const uint64_t *diff_off, const uint32_t *dused, int n, int max_res,
                                  const uint32_t *first_result, const unsigned long long *diff_first,
                                  smb_ali_result *out_res, uint8_t *out_diff, cudaStream_t st, int *nlaunch,
                                  const GatherCigar *cg = nullptr);

constexpr int BAND_SMEM_WCAP_MAX = 512;  // 512 slots * 64 threads * 4 B = 128 KB of shared memory
struct BandPlan {
  struct Class { int wcap, start, count; };
  std::vector<int> order;       // thread-per-task classes first, then the warp-per-task tasks
  std::vector<Class> classes;
  int warp_start = 0, warp_count = 0;   // K3 tasks of band_warp_kernel<32> (band_warp.cu)
  int half_start = 0, half_count = 0;   // ... and of band_warp_kernel<16> (two tasks per warp)
  int pack_start = 0, pack_count = 0, pack_maxrows = 0, pack_maxread = 0;   // band_pack_kernel<16, 2> (four tasks per warp, s16x2)
  int pack8_start = 0, pack8_count = 0, pack8_maxrows = 0, pack8_maxread = 0;   // band_pack_kernel<8, 3> (eight tasks per warp)
  int wide_start = 0, wide_count = 0;   // band_wide_kernel (band_wide.cu: bands <= 128 diagonals, windows <= 512 rows)
  int long16_start = 0, long16_count = 0, long32_start = 0, long32_count = 0;   // band_long_kernel<16 | 32> (band_long.cu)
};
void plan_band(const smb_band_task *h_tasks, int ntasks, bool align, const Scoring &sc, BandPlan &plan);
size_t band_gring_words(const BandPlan &plan);
// side stream for the small launches of a batch (the few tasks the packed kernel does not take): they
// run beside the packed kernel instead of adding their latency behind it
struct BandSide {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
cudaError_t launch_band(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                        const BandPlan &plan, const int *d_order, bool align,
                        int32_t *d_scores, BandOut out, int max_res,
                        const uint64_t *d_dir_off, uint32_t *d_dirs,
                        const uint64_t *d_diff_off, const uint32_t *d_diff_cap,
                        uint32_t *d_gring, int *d_ticket, int sm_count, cudaStream_t st, int *nlaunch,
                        const BandSide *side = nullptr);
cudaError_t launch_band_warp(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                             const int *d_order, int ntasks, int lanes, int *d_ticket, BandOut out, int max_res,
                             const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch);
cudaError_t warm_band_warp();
cudaError_t launch_band_wide(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                             const int *d_order, int ntasks, int *d_ticket, BandOut out, int max_res,
                             const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch);
cudaError_t warm_band_wide();
cudaError_t launch_band_long(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks, const int *d_order,
                             int ntasks, int dpt, int *d_ticket, BandOut out, int max_res, const uint64_t *d_dir_off,
                             uint32_t *d_dirs, const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch);
cudaError_t warm_band_long();
cudaError_t launch_band_pack(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                             const int *d_order, int ntasks, int lanes, int max_rows, int max_read, int *d_ticket, BandOut out, int max_res,
                             const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch);
cudaError_t warm_band_pack();

// ---- K1 ----
struct Index {
  int typ, wordlen, nskip, nbits_key, nbits_lo;
  uint32_t nkeys, npos, nwords, keymod;
  uint64_t wordmask, wordmask_lo, wordmask_hi;
  const uint32_t *idx, *pos, *wordidx, *posidx;
};

// hashTableFetchHitPositions (hashidx.c:1193-1212): the positions of a k-mer word found by lookup()
__device__ __forceinline__ uint32_t fetch_positions(const Index &ix, uint32_t posidx, const uint32_t *&posp) {
  posp = nullptr;
  if (ix.typ == 0) {
    if (posidx < ix.nkeys) {
      const uint32_t s = __ldg(ix.idx + posidx);
      posp = ix.pos + s;
      return __ldg(ix.idx + posidx + 1) - s;
    }
  } else if (posidx < ix.npos) {
    const uint32_t s = __ldg(ix.posidx + posidx);
    posp = ix.pos + s;
    return __ldg(ix.posidx + posidx + 1) - s;
  }
  return 0;
}

// per-read tables of a multi-table seed batch (smb_seed_batch_tables): perfect-hash indexes that
// share word length and sampling step and differ in their arrays only
struct IndexTab {
  const uint32_t *idx, *pos;
  uint32_t npos, reserved;
};

// index construction (index_build.cu)
struct IndexBuildSeq { uint64_t start; uint32_t offs, n_k, tup_base, reserved; };   // == smb_index_seq
struct IndexBuildOut {
  uint32_t npos, nwords, nkeys;
  uint32_t *block;                       // one device allocation holding the four arrays (cudaFree it)
  uint32_t *idx, *pos, *wordidx, *posidx;
};
cudaError_t index_build(const uint32_t *d_packed, const IndexBuildSeq *h_seqs, int nseq, int k, int nskip, int typ,
                        int nbits_key, int nbits_lo, cudaStream_t st, IndexBuildOut *out, int *nlaunch);

struct SeedArgs {
  const IndexTab *tab;        // nullptr: every read uses the kernel's Index argument
  const uint32_t *read_tab;   // [nreads] table of each read
  const uint64_t *read_off;   // [nreads] arena offsets
  const uint32_t *read_len;   // [nreads]
  const uint64_t *slot_off;   // [nreads] first slot of the read (forward strand); reverse at +read_len
  const uint8_t *qual;        // arena-parallel quality bytes or nullptr
  int nreads;
  uint32_t maxhit_per_tuple, maxhit_total;
  int basq_thresh, is_short;
  uint32_t maxlen;            // longest read of the batch
  smb_seed_info *info;        // [2*nreads]
  uint32_t *posidx, *nhits, *qoffs, *sortkey, *sidx, *frame;  // slot arrays
  uint8_t *qmask, *qbuf;
};

struct HitArgs {
  SeedArgs seed;              // the device-resident seed tables of the last smb_seed_batch
  const smb_hit_req *req;     // [nreq]
  int nreq;
  uint32_t nhits_alloc;       // HashHitList.nhits_alloc (hashhit.c:1497)
  uint32_t *count;            // [nreq] hits per list
  uint32_t *maxhit_used;      // [nreq] per-seed cut-off that finally succeeded
  int32_t *errs;              // [nreq]
  const uint64_t *offset;     // [nreq] start of each list in sqdat (FILL pass)
  uint64_t *sqdat;
  uint8_t *list_qmask;        // HITQUAL mask of every list (read_len bytes each) or nullptr
  const uint64_t *qmask_off;  // [nreq] start of each list's mask
  const uint8_t *req_skip;    // [nreq] 1: the list is known to be empty (nullptr: no such knowledge)
};
cudaError_t launch_hits(const Index &ix, const HitArgs &a, bool fill, cudaStream_t st, int *nlaunch);

cudaError_t launch_seed(const Index &ix, const uint8_t *arena, const SeedArgs &a, cudaStream_t st,
                        int *nlaunch);

// forces the (lazily loaded) kernels of the library onto the device
cudaError_t warm_sw();
cudaError_t warm_band();
cudaError_t warm_seed();
cudaError_t warm_compact();

cudaError_t run_int_peak(int mode, int sm_count, int *d_out, int iters, cudaStream_t st, double *ops);

}  // namespace smb
