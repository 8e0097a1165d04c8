// block.cuh - device-resident block pipeline between K1 and K3 (block.cu, api_block.cu):
// candidate selection (segment.c), K2/K3 task lists, the score / threshold replay of rmap.c.
#pragma once
#include "common.cuh"

namespace smb {

// one candidate as the reference's SEGCAND (segment.c:240-266), what the sort and the offsets need
struct SegCand {
  uint32_t qs, qe, rs, re;
  short shiftoffs, shift2mm, srange;
  uint8_t flag, pad;
  uint32_t cover;
  int32_t seqidx;
};
static_assert(sizeof(SegCand) == 32, "layout");

// candidate with its window, dense over the block (RMAPCAND of rmap.c:112-130 + window geometry)
struct DCand {
  uint64_t refoff;        // window start in the packed reference
  uint64_t rs;            // window start inside sequence sqidx (RMAPCAND.rs)
  uint32_t reflen;
  uint32_t qs, qe;
  int32_t band_l, band_r;
  int32_t sqidx;
  uint32_t cover;
  uint8_t rev, simd, pad[2];
};
static_assert(sizeof(DCand) == 48, "layout");

constexpr int BLK_K2_BINS = 20;   // [1..8] 32-bit classes, [9..16] paired 16-bit classes, [17] band-fast tasks, [18] long reads paired
constexpr int BLK_K3_BINS = 32;   // BandPlan classes (band.h)

// counters of a block on the device (zeroed per block)
struct BlockCounters {
  unsigned int k2_hist[BLK_K2_BINS];
  unsigned int k2_cursor[BLK_K2_BINS];
  unsigned int k3_hist[BLK_K3_BINS];
  unsigned int k3_cursor[BLK_K3_BINS];
  unsigned int max_rlen_multi;          // longest window of a K2 task with a read > 256 bases
  unsigned int pack_maxrows, pack_maxread;
  unsigned int n_exceed;                // SIMD scores that hit ERRCODE_SWATEXCEED (host runs the K2' fallback)
  unsigned int bf_cursor;
  unsigned int pad;
  unsigned long long dir_words, diff_bytes;   // totals of the K3 tasks
  unsigned long long k2_cells, k2_cells_ref, k2_tasks_ref;  // all candidates / those the reference scores
};

struct BlockArgs {
  // inputs
  const smb_block_job *jobs;
  int njobs;
  const smb_block_ival *ivals;
  smb_block_params prm;
  int ktup, nskip, nseq, match, mismatch, gap_init, gap_ext;
  const uint64_t *seq_offs;         // [nseq + 1]
  SeedArgs seed;                    // seed tables of the last seed batch
  Index ix;                         // its index
  unsigned long long *seqmask;      // [2 * njobs] sequences that hold a seed position of job x strand (nullptr: not computed)
  uint8_t *req_skip;                // [nreq] requests known to be empty
  // hit lists
  smb_hit_req *req;                 // [nreq]
  const uint32_t *job_req;          // [njobs + 1] first request of every job
  int32_t *req_seqidx;              // [nreq]
  const uint64_t *hit_off;          // [nreq + 1]
  const uint64_t *sqdat;
  const int32_t *req_err;
  uint32_t *req_ncand;              // [nreq] candidates of every list
  // candidate scratch (indexed by hit offset of the job)
  uint64_t *sd_sqo; int32_t *sd_len;                      // seeds of the current hit region
  uint32_t *sg_ix; int32_t *sg_nseed; uint32_t *sg_cover; // segments of the current hit region
  SegCand *cand;
  uint32_t *sort_key, *sort_idx;
  uint32_t *mask; uint32_t mask_words;                    // coverage masks in HBM: 32 per job, mask_words each
  // per job
  smb_block_read *rd;
  uint32_t *n_sort;                 // [njobs]  (scanned into cand_first)
  uint32_t *cover_deficit;          // [2 * njobs]
  const unsigned long long *cand_first; // [njobs + 1]
  uint32_t *nk3;                    // [njobs]
  const unsigned long long *k3_first;   // [njobs + 1]
  // dense candidates
  DCand *dc;
  uint32_t *dc_job;                 // job of every dense candidate
  smb_sw_task *swt;
  smb_band_task *bft;               // band-fast tasks (same index as the candidate) or nullptr
  int32_t *score, *serr;            // per candidate
  int32_t *k3rank;                  // rank among the aligned candidates of the job or -1
  uint8_t *k3cls;
  int *k2_order;                    // K2 order lists, class c at k2_start[c]
  int *bf_order;
  unsigned int k2_start[BLK_K2_BINS];
  // K3
  smb_band_task *bat;
  smb_block_cand *k3c;
  uint32_t *dir_words_arr, *diff_cap;   // per K3 task (scanned into dir_off / diff_off)
  uint32_t *diff_stride;
  int *k3_order;
  unsigned int k3_start[BLK_K3_BINS];
  BlockCounters *cnt;
};

cudaError_t launch_block_seqmask(const BlockArgs &a, cudaStream_t st, int *nlaunch);
cudaError_t launch_block_reqs(const BlockArgs &a, cudaStream_t st, int *nlaunch);
cudaError_t launch_block_cands(const BlockArgs &a, int max_lists, double hits_per_job, cudaStream_t st, int *nlaunch);
cudaError_t launch_block_emit_k2(const BlockArgs &a, unsigned long long ncand, cudaStream_t st, int *nlaunch);
cudaError_t launch_block_exceed(const BlockArgs &a, unsigned long long ncand, cudaStream_t st, int *nlaunch);
cudaError_t launch_block_replay(const BlockArgs &a, cudaStream_t st, int *nlaunch);
cudaError_t launch_block_emit_k3(const BlockArgs &a, unsigned long long ncand, cudaStream_t st, int *nlaunch);
cudaError_t warm_block();

}  // namespace smb
