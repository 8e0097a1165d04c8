// cigar.cuh - the CIGAR / edit-distance walk over an alignment string and the arguments of the kernels that use
// it (cigar.cu, compact.cu); private to csrc/.  What is restated: writeDiffStrCIGAR
// (/root/reference/src/diffstr.c:298-367, formats DIFFSTRFORM_CIGEXT / _XMISMATCH) and
// diffStrGetLevenshteinDistance (diffstr.c:1496-1510).
#pragma once
#include "common.cuh"

namespace smb {

__host__ __device__ __forceinline__ int cg_put(char *out, const int pos, uint32_t count, const char op, const bool write) {
  int nd = 1;
  for (uint32_t v = count; v >= 10u; v /= 10u) ++nd;
  if (write) {
    for (int j = nd - 1; j >= 0; --j) { out[pos + j] = (char)('0' + count % 10u); count /= 10u; }
    out[pos + nd] = op;
  }
  return nd + 1;
}

// returns the text length; *nm = edit distance, or < 0 where the reference fails (then length 0):
// -1 = ERRCODE_FAILURE (empty string), -59 = -ERRCODE_DIFFSTR (the string does not end with an S byte)
// (__host__ too: tests/test_cigar_walk_host.py compiles this very function for the CPU and checks it against the
// reference's functions without a GPU)
template <bool WRITE>
__host__ __device__ __forceinline__ int cg_walk(const uint8_t *__restrict__ d, const uint32_t clip_start, const uint32_t clip_end,
                                       const int flags, char *out, int *nm) {
  const bool silent_mm = !(flags & SMB_CIGAR_XMISMATCH);
  const char clipc = (flags & SMB_CIGAR_SOFTCLIP) ? 'S' : 'H';
  if (!d[0]) { *nm = SMB_ERRCODE_FAILURE; return 0; }   // empty string (diffstr.c:319); ERRCODE_FAILURE is -1
  int pos = 0, ed = 0;
  if (clip_start > 0) pos += cg_put(out, pos, clip_start, clipc, WRITE);
  uint32_t prev_count = 0, typ = 0, prev_typ = 0;
  for (int i = 0; d[i]; ++i) {
    const uint32_t count = d[i] & 63u;
    typ = d[i] >> 6;
    if (typ != 0u) ++ed;
    const bool silent = typ == 0u || (typ == 3u && silent_mm);
    if (prev_typ == 0u) {
      prev_count += count;
      if (silent) { ++prev_count; continue; }
    } else if (typ == prev_typ && count < 1u) {
      ++prev_count;
      continue;
    }
    if (prev_count > 0u) pos += cg_put(out, pos, prev_count, "MDIX"[prev_typ], WRITE);
    if (silent) {
      prev_count = count + 1u;
      prev_typ = 0u;
    } else {
      if (count > 0u && prev_typ != 0u) pos += cg_put(out, pos, count, 'M', WRITE);
      prev_count = 1u;
      prev_typ = typ;
    }
  }
  if (typ != 3u) { *nm = -SMB_ERRCODE_DIFFSTR; return 0; }
  if (prev_count > 1u) pos += cg_put(out, pos, prev_count - 1u, silent_mm ? 'M' : 'X', WRITE);
  if (clip_end > 0u) pos += cg_put(out, pos, clip_end, clipc, WRITE);
  if (ed > 0) --ed;   // the terminating S does not count
  *nm = ed;
  return pos;
}

// alignments given as dense results (res != null) or explicitly (x_off != null): cigar.cu
struct CigarArgs {
  const smb_ali_result *res;
  const smb_band_task *tasks;      // read length of the task of a dense result
  const uint8_t *diff;
  const uint32_t *x_off, *x_cs, *x_ce;   // explicit: DiffStr offset, clip_start, clip_end per alignment
  int n;
  int flags;                       // SMB_CIGAR_*
  uint32_t *len;                   // count pass out
  int32_t *nm;                     // count pass out
  const unsigned long long *off;   // fill pass in: exclusive scan of len
  uint32_t *first_out;             // fill pass out [n + 1]
  char *text;                      // fill pass out
};
cudaError_t launch_cigar_count(const CigarArgs &a, unsigned long long *off, unsigned long long *tile, cudaStream_t st,
                               int *nlaunch);
cudaError_t launch_cigar_fill(const CigarArgs &a, cudaStream_t st, int *nlaunch);

// alignments in the result slots of K3 tasks (resident block): text bytes per task, counted before the output
// compaction scans them along with the result and DiffStr counts; the text itself is written by the gather
struct CigarSlots {
  const smb_ali_result *slots;
  const uint32_t *nres;
  const uint64_t *diff_off_task;
  const uint8_t *diff_slots;
  const smb_band_task *tasks;
  int n, max_res, flags;
};
cudaError_t launch_cigar_task_count(const CigarSlots &a, uint32_t *task_bytes, cudaStream_t st, int *nlaunch);

// output of the gather when the stage is on (compact.cu gather_results): one blob {first[nres + 1], nm[nres], text}
struct GatherCigar {
  const smb_band_task *tasks;
  const unsigned long long *cig_first;   // per task: offset of its first alignment's text
  uint32_t *first_out;
  int32_t *nm;
  char *text;                            // null: stage off
  int flags;
  uint32_t nres_total;
  unsigned long long ncig_total;
};

}  // namespace smb
