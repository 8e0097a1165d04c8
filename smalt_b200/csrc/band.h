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
