// The device's CIGAR / edit-distance walk (smalt_b200/csrc/cigar.cuh cg_walk, the function cigar_task_count_kernel
// and gather_results call) compiled for the HOST: tests/test_cigar_walk_host.py checks this very source against the
// reference's writeDiffStrCIGAR / diffStrGetLevenshteinDistance without a GPU.  Test infrastructure.
#include "cigar.cuh"

extern "C" int cgw_host(const unsigned char *d, unsigned clip_start, unsigned clip_end, int flags, char *out, int maxout,
                        int *nm) {
  int nm_count = 0, nm_write = 0;
  const int n = smb::cg_walk<false>(d, clip_start, clip_end, flags, nullptr, &nm_count);   // the count pass
  if (n > maxout) return -2;
  const int m = smb::cg_walk<true>(d, clip_start, clip_end, flags, out, &nm_write);        // the fill pass
  if (m != n || nm_write != nm_count) return -3;   // both passes must agree: the offsets come from the first
  *nm = nm_count;
  return n;
}
