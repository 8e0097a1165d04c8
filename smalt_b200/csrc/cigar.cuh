// cigar.cuh - arguments of the CIGAR kernels (cigar.cu); private to csrc/.
#pragma once
#include "common.cuh"

namespace smb {

struct CigarArgs {
  // alignments: (a) result slots of K3 tasks (nres != null), (b) dense results (res != null), (c) explicit (x_off != null)
  const smb_ali_result *res;
  const uint32_t *nres;            // (a) results per task
  const uint32_t *first;           // (a) dense index of a task's first result
  const uint64_t *diff_off_task;   // (a) DiffStr slot of a task
  const smb_band_task *tasks;      // (a, b) read length of the task
  const uint8_t *diff;
  const uint32_t *x_off, *x_cs, *x_ce;   // (c) DiffStr offset, clip_start, clip_end per alignment
  int n;                           // threads: tasks * max_res (a) / alignments (b, c)
  int max_res;
  int flags;                       // SMB_CIGAR_*
  size_t ndense;                   // alignments (fill pass: first_out[ndense] = total)
  uint32_t *len;                   // count pass out, dense
  int32_t *nm;                     // count pass out, dense
  const unsigned long long *off;   // fill pass in: exclusive scan of len
  uint32_t *first_out;             // fill pass out [ndense + 1]
  char *text;                      // fill pass out
};

cudaError_t launch_cigar_count(const CigarArgs &a, size_t nscan, unsigned long long *off, unsigned long long *tile,
                               cudaStream_t st, int *nlaunch);
cudaError_t launch_cigar_fill(const CigarArgs &a, cudaStream_t st, int *nlaunch);

}  // namespace smb
