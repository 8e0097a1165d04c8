// sw2_lut.cuh - the constants of the K2 score lookup (sw_score.cu: sw_score2_kernel, sw_long2_kernel) as
// host + device functions: tests/test_k2_rowtable_model.py compiles them for the CPU and checks, with an emulated
// PRMT, that one byte permute gives the substitution scores of both tasks of a packed cell pair for every pair of
// bases.  Private to csrc/.
#pragma once
#include <cstdint>

namespace smb {

// Entry numbers of the row table: a * 4 + b for two standard bases (the 16 entries every in-window step reads lie
// in 32 different banks as 8-byte elements: lanes that read different entries never conflict), 16 + a * 8 + b
// otherwise.
constexpr int SW2_LUT_N = 16 + 64;
__host__ __device__ __forceinline__ uint32_t sw2_lut_index(uint32_t a, uint32_t b) {
  return (a < 4u && b < 4u) ? a * 4u + b : 16u + a * 8u + b;
}
// score bytes {s(q = 0..3, x)} of a window base x (4 = X: unused, the pair takes the per-cell path; 5..7: N / padding)
__host__ __device__ __forceinline__ uint32_t sw2_tab_word(uint32_t x, int match, int mismatch) {
  uint32_t t = 0;
  for (uint32_t q = 0; q < 4u; ++q) {
    const int v = x < 4u ? (q == x ? match : mismatch) : (x == 4u ? mismatch : 0);
    t |= (uint32_t)(v & 0xff) << (8u * q);
  }
  return t;
}
// the row's selector nibbles (low half) and N / padding masks (high half) of the masked form
__host__ __device__ __forceinline__ uint32_t sw2_wsel_word(uint32_t a, uint32_t b) {
  return (a < 4u ? a * 0x11u : 0x00440000u) | (b < 4u ? b * 0x1100u : 0x44000000u);
}
// the per-column PRMT selector of the row-table form; code 8 = padding column
__host__ __device__ __forceinline__ uint32_t sw2_qsel_tab(uint32_t qa, uint32_t qb) {
  return (qa < 4u ? (qa | ((qa | 8u) << 4)) : 0x88u) | ((qb < 4u ? ((qb | 4u) | ((qb | 12u) << 4)) : 0xCCu) << 8);
}
// ... of the masked form {qA, qA|8, qB, qB|8} (N, padding: 4), bit 2 flipped: the table is the SECOND source of
// the PRMT (as first source ptxas overwrites it with the result and copies it afresh for every cell):
// idx' = ((q ^ 4) ^ r) & ~mask
__host__ __device__ __forceinline__ uint32_t sw2_qsel_masked(uint32_t qa, uint32_t qb) {
  const uint32_t ia = qa < 4u ? qa : 4u, ib = qb < 4u ? qb : 4u;
  return (ia | ((ia | 8u) << 4) | (ib << 8) | ((ib | 8u) << 12)) ^ 0x4444u;
}

}  // namespace smb
