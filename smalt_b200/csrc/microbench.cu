// microbench.cu - measures the integer / DPX issue peak of the device the context runs on.
// Used as the roofline denominator for the Smith-Waterman kernels (no tensor cores on this
// path): independent chains of VIADDMNMX (__viaddmax_s32) resp. plain IADD3/IMNMX per thread,
// enough warps per SM to saturate the issue ports.  Reported in giga thread-operations/s.
#include "common.cuh"

namespace smb {

template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(int *out, int iters, int seed) {
  int a0 = threadIdx.x + seed, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
      a7 = a0 + 7;
  const int b = seed | 1, c = seed + 3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {  // DPX: max(a + b, c)
        a0 = __viaddmax_s32(a0, b, c); a1 = __viaddmax_s32(a1, b, c); a2 = __viaddmax_s32(a2, b, c);
        a3 = __viaddmax_s32(a3, b, c); a4 = __viaddmax_s32(a4, b, c); a5 = __viaddmax_s32(a5, b, c);
        a6 = __viaddmax_s32(a6, b, c); a7 = __viaddmax_s32(a7, b, c);
      } else if (MODE == 1) {  // three-input max
        a0 = __vimax3_s32(a0, a1, c); a1 = __vimax3_s32(a1, a2, b); a2 = __vimax3_s32(a2, a3, c);
        a3 = __vimax3_s32(a3, a4, b); a4 = __vimax3_s32(a4, a5, c); a5 = __vimax3_s32(a5, a6, b);
        a6 = __vimax3_s32(a6, a7, c); a7 = __vimax3_s32(a7, a0, b);
      } else if (MODE == 3) {  // DPX, two 16-bit lanes per register: max(a + b, c) per half
        a0 = (int)__viaddmax_s16x2((unsigned)a0, (unsigned)b, (unsigned)c); a1 = (int)__viaddmax_s16x2((unsigned)a1, (unsigned)b, (unsigned)c);
        a2 = (int)__viaddmax_s16x2((unsigned)a2, (unsigned)b, (unsigned)c); a3 = (int)__viaddmax_s16x2((unsigned)a3, (unsigned)b, (unsigned)c);
        a4 = (int)__viaddmax_s16x2((unsigned)a4, (unsigned)b, (unsigned)c); a5 = (int)__viaddmax_s16x2((unsigned)a5, (unsigned)b, (unsigned)c);
        a6 = (int)__viaddmax_s16x2((unsigned)a6, (unsigned)b, (unsigned)c); a7 = (int)__viaddmax_s16x2((unsigned)a7, (unsigned)b, (unsigned)c);
      } else if (MODE == 4) {  // three-input max per 16-bit half
        a0 = (int)__vimax3_s16x2((unsigned)a0, (unsigned)a1, (unsigned)c); a1 = (int)__vimax3_s16x2((unsigned)a1, (unsigned)a2, (unsigned)b);
        a2 = (int)__vimax3_s16x2((unsigned)a2, (unsigned)a3, (unsigned)c); a3 = (int)__vimax3_s16x2((unsigned)a3, (unsigned)a4, (unsigned)b);
        a4 = (int)__vimax3_s16x2((unsigned)a4, (unsigned)a5, (unsigned)c); a5 = (int)__vimax3_s16x2((unsigned)a5, (unsigned)a6, (unsigned)b);
        a6 = (int)__vimax3_s16x2((unsigned)a6, (unsigned)a7, (unsigned)c); a7 = (int)__vimax3_s16x2((unsigned)a7, (unsigned)a0, (unsigned)b);
      } else {  // plain integer add + max (2 ops)
        a0 = max(a0 + b, c); a1 = max(a1 + b, c); a2 = max(a2 + b, c); a3 = max(a3 + b, c);
        a4 = max(a4 + b, c); a5 = max(a5 + b, c); a6 = max(a6 + b, c); a7 = max(a7 + b, c);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

cudaError_t run_int_peak(int mode, int sm_count, int *d_out, int iters, cudaStream_t st, double *ops) {
  const int grid = sm_count * 8;
  if (mode == 0) int_peak_kernel<0><<<grid, 256, 0, st>>>(d_out, iters, 12345);
  else if (mode == 1) int_peak_kernel<1><<<grid, 256, 0, st>>>(d_out, iters, 12345);
  else if (mode == 3) int_peak_kernel<3><<<grid, 256, 0, st>>>(d_out, iters, 12345);
  else if (mode == 4) int_peak_kernel<4><<<grid, 256, 0, st>>>(d_out, iters, 12345);
  else int_peak_kernel<2><<<grid, 256, 0, st>>>(d_out, iters, 12345);
  *ops = (double)grid * 256.0 * iters * 64.0 * (mode == 2 ? 2.0 : 1.0);
  return cudaGetLastError();
}

}  // namespace smb
