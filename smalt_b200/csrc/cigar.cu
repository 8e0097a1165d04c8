// cigar.cu - CIGAR text and edit distance of the alignments on the device.
//
// What the reference's output stage computes per reported alignment from the compressed alignment
// string (DiffStr, diffstr.h:28-77: one byte per run, type << 6 | matches in front of it, terminated
// by an S byte and a 0):
//   writeDiffStrCIGAR            /root/reference/src/diffstr.c:298-367, as called by diffStrPrintf /
//                                diffStrPrintfStr with DIFFSTRFORM_CIGEXT[_XMISMATCH] (:1066-1075,
//                                :1103-1114) from fprintREPALIsam (report.c:892-896): "%d%c" per
//                                operation, clips in front and behind
//   diffStrGetLevenshteinDistance   diffstr.c:1496-1510 (the NM:i: field, report.c:898)
// The clips follow from the alignment itself (report.c:832-843 with results.c:1897-1903): on either
// strand clip_start = qs and clip_end = qlen - 1 - qe in the coordinates of the strand that was aligned.
//
// Two passes over the alignments (one thread each; a DiffStr of a 150-base read has ~5-20 bytes):
// text lengths + edit distances, exclusive scan of the lengths, then the text itself.  The alignments
// are read where K3 left them: per-task result slots (resident block), dense results (multi-pass path),
// or explicit DiffStr offsets and clips (smb_cigar_batch).
#include "common.cuh"
#include "cigar.cuh"

namespace smb {

__device__ __forceinline__ int cg_put(char *out, const int pos, uint32_t count, const char op, const bool write) {
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
template <bool WRITE>
__device__ int cg_walk(const uint8_t *__restrict__ d, const uint32_t clip_start, const uint32_t clip_end, const int flags,
                       char *out, int *nm) {
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

// where alignment `i` of the launch is: DiffStr, clips, dense index
__device__ __forceinline__ bool cg_locate(const CigarArgs &a, const int i, const uint8_t **d, uint32_t *cs, uint32_t *ce,
                                          size_t *dense) {
  if (a.x_off) {   // explicit
    *d = a.diff + a.x_off[i];
    *cs = a.x_cs[i]; *ce = a.x_ce[i];
    *dense = (size_t)i;
    return true;
  }
  smb_ali_result r;
  uint32_t task;
  if (a.nres) {   // result slots of task t = i / max_res
    const int t = i / a.max_res, k = i - t * a.max_res;
    if ((uint32_t)k >= a.nres[t]) return false;
    r = a.res[i];
    task = (uint32_t)t;
    *d = a.diff + a.diff_off_task[t] + r.diff_off;
    *dense = (size_t)a.first[t] + (size_t)k;
  } else {        // dense results
    r = a.res[i];
    task = r.task;
    *d = a.diff + r.diff_off;
    *dense = (size_t)i;
  }
  const uint32_t qlen = a.tasks[task].read_len;
  *cs = (uint32_t)r.qs;
  *ce = qlen - 1u - (uint32_t)r.qe;
  return true;
}

__global__ void __launch_bounds__(128) cigar_count_kernel(const CigarArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const uint8_t *d; uint32_t cs, ce; size_t dense;
  if (!cg_locate(a, i, &d, &cs, &ce, &dense)) return;
  int nm;
  a.len[dense] = (uint32_t)cg_walk<false>(d, cs, ce, a.flags, nullptr, &nm);
  a.nm[dense] = nm;
}

__global__ void __launch_bounds__(128) cigar_fill_kernel(const CigarArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) a.first_out[a.ndense] = (uint32_t)a.off[a.ndense];
  if (i >= a.n) return;
  const uint8_t *d; uint32_t cs, ce; size_t dense;
  if (!cg_locate(a, i, &d, &cs, &ce, &dense)) return;
  const unsigned long long o = a.off[dense];
  a.first_out[dense] = (uint32_t)o;
  int nm;
  (void)cg_walk<true>(d, cs, ce, a.flags, a.text + o, &nm);
}

cudaError_t launch_cigar_count(const CigarArgs &a, size_t nscan, unsigned long long *off, unsigned long long *tile,
                               cudaStream_t st, int *nlaunch) {
  // len[0 .. nscan) was zeroed by the caller (slots that hold no alignment, the entry behind the last one)
  if (a.n > 0) {
    cigar_count_kernel<<<(a.n + 127) / 128, 128, 0, st>>>(a); ++*nlaunch;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return launch_scan_counts(a.len, (int)nscan, off, tile, st, nlaunch);
}

cudaError_t launch_cigar_fill(const CigarArgs &a, cudaStream_t st, int *nlaunch) {
  const int n = a.n > 0 ? a.n : 1;
  cigar_fill_kernel<<<(n + 127) / 128, 128, 0, st>>>(a); ++*nlaunch;
  return cudaGetLastError();
}

}  // namespace smb
