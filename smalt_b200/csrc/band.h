// band.h - band geometry shared by host planning code and the device kernels.
// Restates initALIBAND (/root/reference/src/alignment.c:310-396).
#pragma once
#ifdef __CUDACC__
#define SMB_HD __host__ __device__ __forceinline__
#else
#define SMB_HD inline
#endif

namespace smb {

struct Band {
  int band_width, l_edge, r_edge, l_edge_orig, r_edge_orig;
  int s_left, s_len, q_left, q_len;
};

// returns 0 on success, -1 (ERRCODE_FAILURE) when the band never meets the read segment
SMB_HD int band_init(Band &b, int l_edge, int r_edge, int q_left, int q_right, int q_len,
                     int s_left, int s_right, int s_len) {
  b.s_len = (s_right < 0 || s_right >= s_len) ? s_len : s_right + 1;
  b.q_len = (q_right < 0 || q_right >= q_len) ? q_len : q_right + 1;
  b.s_left = (s_left > 0 && s_left < b.s_len) ? s_left : 0;
  b.q_left = (q_left > 0 && q_left < b.q_len) ? q_left : 0;
  b.l_edge_orig = b.l_edge = l_edge;
  b.r_edge_orig = b.r_edge = r_edge;
  b.band_width = r_edge - l_edge + 1;
  if (b.band_width <= 0) {
    b.band_width = 0;
    b.l_edge = b.q_left;
    b.r_edge = b.q_len - 1;
  } else {
    if (b.l_edge_orig + b.s_len > b.q_len) b.s_len = b.q_len - b.l_edge_orig;
    b.l_edge += b.s_left;
    if (b.l_edge >= b.q_len || b.r_edge_orig + b.s_len <= b.q_left) return -1;
    b.r_edge += b.s_left;
    if (b.r_edge < b.q_left) {
      const int d = b.q_left - b.r_edge;
      b.s_left += d;
      b.l_edge += d;
      b.r_edge = b.q_left;
    }
    if (b.r_edge > b.q_len - 1) b.r_edge = b.q_len - 1;
  }
  b.band_width = b.r_edge - b.l_edge + 1;
  return (b.band_width >= 0) ? 0 : -1;
}

}  // namespace smb

// ---- warp-per-task K3 kernel (band_warp.cu): which tasks it takes ----
namespace smb {
constexpr int BW_MAXROWS = 256;   // window rows: staged window bytes, direction words per diagonal
constexpr int BW_MAXREAD = 384;   // staged read bases
constexpr int BW_MAXDIAG = 64;    // band diagonals: two per lane

// The wavefront kernel handles a task when its window, read and band fit the per-warp
// shared-memory staging.  Recursion sub-ranges only shrink rows and band, so the test on the
// initial geometry covers every DP pass of the task.
// returns 0 (not eligible), 16 (half a warp per task: band <= 32 diagonals) or 32
SMB_HD int band_warp_lanes(int l_edge, int r_edge, int p_left, int p_right, int read_len,
                           int u_left, int u_right, int ref_len) {
  if (ref_len > BW_MAXROWS || read_len > BW_MAXREAD || ref_len < 1 || read_len < 1) return 0;
  Band b;
  if (band_init(b, l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len)) return 16;
  // the clipped width can grow by at most the rows skipped when a sub-range starts further down:
  // bound it by the unclipped width
  const int bw0 = r_edge - l_edge + 1;
  const int bw = (bw0 <= 0) ? (b.q_len - b.q_left) : bw0;
  return bw <= BW_MAXDIAG / 2 ? 16 : (bw <= BW_MAXDIAG ? 32 : 0);
}
constexpr int BP8_MAXDIAG = 24;   // band_pack_kernel<8, 3>: eight lanes, three diagonals each
// upper bound of the band width of every DP pass of a task (same bound as band_warp_lanes)
SMB_HD int band_width_bound(int l_edge, int r_edge, int p_left, int p_right, int read_len,
                            int u_left, int u_right, int ref_len) {
  Band b;
  if (band_init(b, l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len)) return 0;
  const int bw0 = r_edge - l_edge + 1;
  return (bw0 <= 0) ? (b.q_len - b.q_left) : bw0;
}
// band_wide_kernel (band_wide.cu): four diagonals per lane, longer windows
constexpr int BWD_MAXROWS = 512;
constexpr int BWD_MAXREAD = 512;
constexpr int BWD_MAXDIAG = 128;
SMB_HD bool band_wide_eligible(int l_edge, int r_edge, int p_left, int p_right, int read_len,
                               int u_left, int u_right, int ref_len) {
  if (ref_len > BWD_MAXROWS || read_len > BWD_MAXREAD || ref_len < 1 || read_len < 1) return false;
  Band b;
  if (band_init(b, l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len)) return true;
  const int bw0 = r_edge - l_edge + 1;
  const int bw = (bw0 <= 0) ? (b.q_len - b.q_left) : bw0;
  return bw <= BWD_MAXDIAG;
}
SMB_HD bool band_warp_eligible(int l_edge, int r_edge, int p_left, int p_right, int read_len,
                               int u_left, int u_right, int ref_len) {
  return band_warp_lanes(l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len) != 0;
}
}  // namespace smb

// ---- thread-per-task band_kernel (band_dp.cu): ring capacity classes ----
namespace smb {
// band cells a row of the task can hold (the H/E ring of band_kernel); `fast` = score-only variant
SMB_HD int band_ring_need(int l_edge, int r_edge, int p_left, int p_right, int read_len, int u_left, int u_right,
                          int ref_len, bool fast) {
  Band b;
  if (band_init(b, l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len)) return 1;
  const int bw0 = r_edge - l_edge + 1;
  int full = b.q_len - b.q_left;
  if (full < 1) full = 1;
  const int w = (bw0 <= 0 || (fast && b.q_left > b.l_edge)) ? full : (bw0 < full ? bw0 : full);
  return w < 1 ? 1 : w;
}
SMB_HD int band_pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}
// class of a thread-per-task launch: ring capacity 32 << class
SMB_HD int band_ring_class(int need) {
  const int wcap = band_pow2_at_least(need + 1);
  int c = 0;
  while ((32 << c) < wcap) ++c;
  return c;
}
// pseudo classes of the warp / CTA kernels (last in BandPlan.order)
constexpr int BAND_CLS_WARP = 31, BAND_CLS_HALF = 30, BAND_CLS_PACK = 29, BAND_CLS_WIDE = 28, BAND_CLS_PACK8 = 27,
              BAND_CLS_LONG16 = 26, BAND_CLS_LONG32 = 25;
constexpr int BAND_CLS_THREAD_END = 25;   // classes below: thread-per-task launches by ring capacity
// band_long_kernel<D> (band_long.cu): one CTA of BAND_LONG_THREADS threads per task, D diagonals per thread
constexpr int BAND_LONG_THREADS = 128;
constexpr int BAND_LONG_MAXDIM = 1 << 20;   // rows / columns (21-bit fields of the maximum key)
// 0: not a task of band_long_kernel, else the diagonals per thread (16 or 32)
SMB_HD int band_long_dpt(int l_edge, int r_edge, int p_left, int p_right, int read_len, int u_left, int u_right,
                         int ref_len) {
  if (ref_len > BAND_LONG_MAXDIM || read_len > BAND_LONG_MAXDIM || ref_len < 1 || read_len < 1) return 0;
  Band b;
  if (band_init(b, l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len)) return 0;
  const int bw0 = r_edge - l_edge + 1;
  const int bw = (bw0 <= 0) ? (b.q_len - b.q_left) : bw0;     // bound of every DP pass of the task
  if (bw <= 16 * BAND_LONG_THREADS) return 16;
  if (bw <= 32 * BAND_LONG_THREADS) return 32;
  return 0;
}
// direction words a task needs in the HBM strip of band_kernel<true> (2 for the warp kernels)
SMB_HD unsigned long long band_dir_words(int l_edge, int r_edge, int p_left, int p_right, int read_len, int u_left,
                                         int u_right, int ref_len) {
  Band b;
  if (band_warp_eligible(l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len) ||
      band_wide_eligible(l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len) ||
      band_init(b, l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len))
    return 2;
  {   // band_long_kernel: one unit of dpt / 16 words per thread and iteration, rows + threads iterations
    const int dpt = band_long_dpt(l_edge, r_edge, p_left, p_right, read_len, u_left, u_right, ref_len);
    if (dpt)
      return ((unsigned long long)ref_len + BAND_LONG_THREADS) * BAND_LONG_THREADS * (unsigned long long)(dpt / 16) + 8u;
  }
  const int bw0 = r_edge - l_edge + 1;
  int w = (bw0 <= 0) ? (b.q_len - b.q_left) : bw0;
  if (w < 1) w = 1;
  return ((unsigned long long)w * (unsigned long long)ref_len) / 16u + 4u;
}
}  // namespace smb
