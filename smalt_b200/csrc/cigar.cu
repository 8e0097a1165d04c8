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
// Resident block: cigar_task_count_kernel sums the text bytes of a task's alignments right after K3; the
// output compaction (compact.cu) scans them together with the result and DiffStr counts, and its gather
// kernel writes the text, the offsets and the edit distances into one blob next to the dense results - one
// extra launch and one extra copy per block.  Dense results (multi-pass path) and explicit alignment strings
// (smb_cigar_batch): count pass, scan, fill pass below.  One thread per alignment / task: a DiffStr of a
// 150-base read has 5-20 bytes.
#include "cigar.cuh"

namespace smb {

__global__ void __launch_bounds__(128) cigar_task_count_kernel(const CigarSlots a, uint32_t *__restrict__ task_bytes) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n) return;
  const uint32_t nr = a.nres[t];
  const uint32_t qlen = a.tasks[t].read_len;
  const uint8_t *db = a.diff_slots + a.diff_off_task[t];
  uint32_t bytes = 0;
  for (uint32_t k = 0; k < nr && k < (uint32_t)a.max_res; ++k) {
    const smb_ali_result r = a.slots[(size_t)t * a.max_res + k];
    int nm;
    bytes += (uint32_t)cg_walk<false>(db + r.diff_off, (uint32_t)r.qs, qlen - 1u - (uint32_t)r.qe, a.flags, nullptr, &nm);
  }
  task_bytes[t] = bytes;
}

cudaError_t launch_cigar_task_count(const CigarSlots &a, uint32_t *task_bytes, cudaStream_t st, int *nlaunch) {
  if (a.n < 1) return cudaSuccess;
  cigar_task_count_kernel<<<(a.n + 127) / 128, 128, 0, st>>>(a, task_bytes); ++*nlaunch;
  return cudaGetLastError();
}

// where alignment `i` of a dense / explicit launch is: DiffStr and clips
__device__ __forceinline__ void cg_locate(const CigarArgs &a, const int i, const uint8_t **d, uint32_t *cs, uint32_t *ce) {
  if (a.x_off) {
    *d = a.diff + a.x_off[i];
    *cs = a.x_cs[i]; *ce = a.x_ce[i];
    return;
  }
  const smb_ali_result r = a.res[i];
  *d = a.diff + r.diff_off;
  *cs = (uint32_t)r.qs;
  *ce = a.tasks[r.task].read_len - 1u - (uint32_t)r.qe;
}

__global__ void __launch_bounds__(128) cigar_count_kernel(const CigarArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.n) return;
  if (i == a.n) { a.len[i] = 0; return; }   // the entry behind the last one: its offset is the total
  const uint8_t *d; uint32_t cs, ce;
  cg_locate(a, i, &d, &cs, &ce);
  int nm;
  a.len[i] = (uint32_t)cg_walk<false>(d, cs, ce, a.flags, nullptr, &nm);
  a.nm[i] = nm;
}

__global__ void __launch_bounds__(128) cigar_fill_kernel(const CigarArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.n) return;
  const unsigned long long o = a.off[i];
  a.first_out[i] = (uint32_t)o;
  if (i == a.n) return;
  const uint8_t *d; uint32_t cs, ce;
  cg_locate(a, i, &d, &cs, &ce);
  int nm;
  (void)cg_walk<true>(d, cs, ce, a.flags, a.text + o, &nm);
}

cudaError_t launch_cigar_count(const CigarArgs &a, unsigned long long *off, unsigned long long *tile, cudaStream_t st,
                               int *nlaunch) {
  cigar_count_kernel<<<(a.n + 1 + 127) / 128, 128, 0, st>>>(a); ++*nlaunch;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_scan_counts(a.len, a.n + 1, off, tile, st, nlaunch);
}

cudaError_t launch_cigar_fill(const CigarArgs &a, cudaStream_t st, int *nlaunch) {
  cigar_fill_kernel<<<(a.n + 1 + 127) / 128, 128, 0, st>>>(a); ++*nlaunch;
  return cudaGetLastError();
}

}  // namespace smb
