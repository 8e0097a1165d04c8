// sw_score.cu - K2: batched Smith-Waterman score kernel for sm_100a.
//
// Replaces swSIMDAlignStriped (/root/reference/src/swsimd.c:868-933): the maximum local
// alignment score of a read (profiled sequence) against a reference window, unbanded,
// canonical affine gaps:
//     h = max(0, Hdiag + S);  H = max(h, E, F);
//     E' = max(E - ext, H - init);  F' = max(F - ext, H - init);   result = max H
// (swsimd.c:745-781; the reference's 8-bit pass / 16-bit retry / lazy-F loop are SSE2
// mechanics of the same recurrence - the returned integer is the exact maximum, and
// ERRCODE_SWATEXCEED when it reaches 65535, swsimd.c:644).
//
// B200 mapping (not a port of the striped layout, which exists to feed 128-bit SIMD):
//  * inter-task parallel: ONE WARP PER read x window task, persistent warps pulling tasks
//    from an atomic counter (grid = multiple of the SM count);
//  * inside a task the 32 lanes form a systolic array over the read: lane l owns C
//    consecutive read columns in REGISTERS (H, E, the read bases), reference rows stream
//    through the lanes one step apart (row i is in lane l at step i+l).  The only
//    inter-lane traffic is two warp shuffles per step: H of the lane's last column and
//    (F << 3 | reference base) - no shared memory, no global traffic in the inner loop;
//  * the recurrence is issued with the DPX integer instructions (VIADDMNMX / VIMNMX3 via
//    __viaddmax_s32 / __vimax3_s32 / __viaddmax_s32_relu): 5 DPX/ALU ops + a compare-select
//    for the substitution score per cell;
//  * reads longer than 32*C columns are processed in column blocks; the block's right
//    boundary column (H, F per row) goes through a per-warp L2-resident scratch strip;
//  * the reference window is read straight from the 3-bit packed reference (10 bases per
//    32-bit word, the .sma layout) or from explicit bytes, 32 rows per coalesced load.
#include "common.cuh"
#include <algorithm>
#include <vector>

#include "sw2_lut.cuh"

namespace smb {

constexpr int SW_WARPS = 4;  // warps per CTA

struct SwClassArgs {
  const int *order;  // task indices of this class
  int ntasks;
  int *counter;      // persistent-scheduler ticket
};

template <int C>
__global__ void __launch_bounds__(SW_WARPS * 32)
sw_score_kernel(const Scoring sc, const SeqSrc src, const smb_sw_task *__restrict__ tasks,
                const SwClassArgs cls, int32_t *__restrict__ scores, int32_t *__restrict__ errs,
                int2 *__restrict__ bscratch, const uint32_t bstride) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
  int2 *const strip0 = bscratch + (size_t)gwarp * 2u * bstride;
  const int gi = sc.gap_init, ge = sc.gap_ext, smatch = sc.match;

  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(cls.counter, 1);
    k = __shfl_sync(FULL, k, 0);
    if (k >= cls.ntasks) break;
    const int tix = __ldg(cls.order + k);
    const smb_sw_task tk = tasks[tix];
    const int qlen = (int)tk.read_len, rlen = (int)tk.ref_len;
    const bool rc = (tk.flags & SMB_TASK_READ_REVCOMP) != 0;
    const bool packed = (tk.flags & SMB_TASK_REF_PACKED) != 0;
    const int nblk = (qlen + 32 * C - 1) / (32 * C);
    const int nsteps = rlen + 31;
    int best = 0;

    for (int b = 0; b < nblk; ++b) {
      const int j0 = b * 32 * C + lane * C;
      int qcode[C], smis[C], H[C], E[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = j0 + c;
        int q = 7, sm = -30000;  // padding column: never scores
        if (j < qlen) {
          q = (int)read_base(src.arena, tk.read_off, (uint32_t)qlen, rc, (uint32_t)j);
          sm = (q < 4) ? sc.mismatch : (int)sc.S[q];  // non-standard read base: S[A][q]
        }
        qcode[c] = q;
        smis[c] = sm;
        H[c] = 0;
        E[c] = 0;
      }
      int hdiag = 0, hout = 0, frout = 0;
      uint32_t rbuf = 7u;
      int2 bin = make_int2(0, 0), keep = make_int2(0, 0);
      const int2 *bin_strip = strip0 + (size_t)((b & 1) ^ 1) * bstride;
      int2 *bout_strip = strip0 + (size_t)(b & 1) * bstride;
      const bool has_in = b > 0, has_out = b < nblk - 1;

      for (int t = 0; t < nsteps; ++t) {
        if ((t & 31) == 0) {
          const int i = t + lane;
          rbuf = (i < rlen) ? ref_base(src, packed, tk.ref_off, (uint32_t)i) : 7u;
          if (has_in) bin = (i < rlen) ? __ldcg(bin_strip + i) : make_int2(0, 0);
        }
        int hl = __shfl_up_sync(FULL, hout, 1);
        const int fr = __shfl_up_sync(FULL, frout, 1);
        const int r0 = (int)__shfl_sync(FULL, rbuf, t & 31);
        int r = fr & 7, F = fr >> 3;
        if (has_in) {
          const int hb = __shfl_sync(FULL, bin.x, t & 31);
          const int fb = __shfl_sync(FULL, bin.y, t & 31);
          if (lane == 0) { hl = hb; F = fb; }
        } else if (lane == 0) {
          hl = 0;
          F = 0;
        }
        if (lane == 0) r = r0;
        const int i = t - lane;
        if (i >= 0 && i < rlen) {
          int diag = hdiag;
          hdiag = hl;
          if (r < 4) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const int s = (qcode[c] == r) ? smatch : smis[c];
              const int h = __viaddmax_s32(diag, s, 0);
              diag = H[c];
              const int hn = __vimax3_s32(h, E[c], F);
              best = max(best, hn);
              const int tt = hn - gi;
              E[c] = __viaddmax_s32(E[c], -ge, tt);
              F = __viaddmax_s32_relu(F, -ge, tt);
              H[c] = hn;
            }
          } else {  // non-standard reference base (N, X, terminator): table row
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const int s = (qcode[c] == 7 && smis[c] < -1000) ? smis[c] : (int)sc.S[r * 8 + qcode[c]];
              const int h = __viaddmax_s32(diag, s, 0);
              diag = H[c];
              const int hn = __vimax3_s32(h, E[c], F);
              best = max(best, hn);
              const int tt = hn - gi;
              E[c] = __viaddmax_s32(E[c], -ge, tt);
              F = __viaddmax_s32_relu(F, -ge, tt);
              H[c] = hn;
            }
          }
          hout = H[C - 1];
          frout = (F << 3) | r;
        }
        if (has_out) {  // hand the block's last column (lane 31) to the next column block
          const int ho = __shfl_sync(FULL, hout, 31);
          const int fo = __shfl_sync(FULL, frout, 31) >> 3;
          const int i31 = t - 31;
          if (i31 >= 0) {
            if (lane == (i31 & 31)) keep = make_int2(ho, fo);
            if ((i31 & 31) == 31 || i31 == rlen - 1) {
              const int base = i31 & ~31;
              if (base + lane <= i31) __stcg(bout_strip + base + lane, keep);
            }
          }
        }
      }
      __syncwarp();
    }
    best = __reduce_max_sync(FULL, best);
    if (lane == 0) {
      const bool exceed = best >= 65535;  // swsimd.c:644
      scores[tix] = exceed ? 0 : best;
      errs[tix] = exceed ? SMB_ERRCODE_SWATEXCEED : SMB_OK;
    }
  }
}

template <int C>
static cudaError_t launch_class(const Scoring &sc, const SeqSrc &src, const smb_sw_task *d_tasks,
                                const SwClassArgs &cls, int32_t *d_scores, int32_t *d_errs,
                                int2 *bscratch, uint32_t bstride, int grid, cudaStream_t st) {
  sw_score_kernel<C><<<grid, SW_WARPS * 32, 0, st>>>(sc, src, d_tasks, cls, d_scores, d_errs,
                                                      bscratch, bstride);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------
// sw_score2_kernel: the same recurrence with TWO TASKS PER WARP in the two 16-bit halves of
// every register (DPX VIADDMNMX.S16x2 / VIMNMX3.S16x2): short reads cannot exceed a score of
// qlen * match, so 15 bits are enough and every integer instruction advances two cells.
//   * task A lives in the low, task B in the high half-words of H, E, F, best;
//   * the substitution score of both cells comes from ONE byte permute: an 8-byte table
//     T = {match, mismatch x3, 0 x4} indexed by (read code ^ window code) for A,C,G,T and by
//     4..7 for N / padding columns (PRMT sign-replication widens the byte to 16 bits), with the
//     selector nibbles of the read precomputed per column and the window's per row;
//   * window codes of both tasks are staged interleaved in shared memory (one 16-bit load per
//     row instead of a shuffle); H and F still travel by one shuffle each per step;
//   * rows with a non-standard window base (N, X, terminator / padding behind the shorter
//     window) and task pairs with an X in the read take the general per-cell table path.
// Padding columns and padding rows score 0: values there never exceed the maximum already
// recorded (h = Hdiag + s <= Hdiag), so they cannot change the result.
// ------------------------------------------------------------------------------------
constexpr int SW2_MAXROWS = 512;    // staged window rows per task
constexpr int SW2_PAD = 32;         // padding rows staged before and behind the window

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// H, E and F are kept BIASED by gap_init (the stored value is the score + gap_init): the floor of H
// is then gap_init instead of 0, max(E - ext, H - gap_init) becomes max(E' - ext, H' - gap_init) with
// H' - gap_init = H >= 0 in both half-words - so that one subtraction is an ordinary 32-bit one that
// cannot borrow across the halves and needs no DPX instruction (it leaves the ALU pipe, which bounds
// this kernel, for the FMA pipe) - and E', F' >= 0 need no RELU.
// One systolic step of C packed cell pairs.  The table index (read selector ^ row selector) with the
// row mask applied is ONE LOP3: rows with N / padding carry selector 0 and mask 4 in their task's
// nibbles, so that every column reads a zero there (a plain xor would turn N against N into a match).
#define SW2_STEP(SEL_EXPR)                                                                  \
  do {                                                                                      \
    _Pragma("unroll")                                                                       \
    for (int c = 0; c < C; ++c) {                                                           \
      const uint32_t s2 = (SEL_EXPR);                                                       \
      const uint32_t h = __viaddmax_s16x2(diag, s2, gi2);     /* floor: H = 0 */            \
      diag = H[c];                                                                          \
      const uint32_t hn = __vimax3_s16x2(h, E[c], F);                                       \
      if (c & 1) best = __vimax3_s16x2(best, H[c - 1], hn);   /* two columns per VIMNMX3 */  \
      else if (c == C - 1) best = __vmaxs2(best, hn);                                       \
      const uint32_t tt = hn - gi2;   /* plain 32-bit subtract (FMA pipe): no half borrows */ \
      E[c] = __viaddmax_s16x2(E[c], nge2, tt);                                              \
      F = __viaddmax_s16x2(F, nge2, tt);                                                    \
      H[c] = hn;                                                                            \
    }                                                                                       \
  } while (0)

// LANES = 32: one task pair per warp, lane owns C columns.  LANES = 16: TWO task pairs per warp (four
// tasks), one per half-warp, lane owns C columns of its pair (C = 2 x the 32-lane class): the skew of
// the systolic array - steps in which a lane runs over padding rows - halves (15 instead of 31 of
// ~200 steps) and the per-step instructions (two shuffles, two selects, one LDS) are shared by twice as
// many cells.  The half-warps run in lockstep: full-warp shuffles of width 16, trip counts and staging
// bounds are the maxima over the two halves.
// What a window row contributes to a step, looked up by the pair of base codes (a, b) of the two tasks' rows
// (the staged row is the BYTE OFFSET of its entry, so the two loads of a step need no arithmetic):
//   tab   score bytes {s(q = 0..3, a)} of task A in .x, of task B in .y: the two sources of ONE PRMT per cell
//         pair whose selector is a per-column constant {q_A, q_A | 8, 4 + q_B, (4 + q_B) | 8} (| 8: the sign of
//         the selected byte, i.e. the high byte of the 16-bit score).  N / padding rows hold zeros.  This form
//         has no spare byte for a read base that must score 0 against every row, so pairs with an N in a READ
//         take the masked form below; padding COLUMNS (behind the read) only need a score <= 0 in every row
//         - values there then never exceed the cell they came from - and select the sign of byte 0 twice
//         (0 or -1).
//   wsel  the row's selector nibbles and N / padding masks of the masked form: index into the table
//         {match, mismatch x3 | 0 x4} = (read selector ^ row selector) & ~mask, one LOP3 + one PRMT per cell pair
//   raw   the base codes for the per-cell table path (X anywhere in the pair's reads or windows)
// Entry numbers: a * 4 + b for two standard bases (the 16 entries every in-window step reads lie in 32 different
// banks as 8-byte elements: lanes that read different entries never conflict), 16 + a * 8 + b otherwise.
// (sw2_lut.cuh: sw2_lut_index, sw2_tab_word, sw2_wsel_word, sw2_qsel_tab, sw2_qsel_masked - host + device)
__device__ __forceinline__ void sw2_build_lut(const Scoring &sc, uint2 *s_tab, uint32_t *s_wsel, uint32_t *s_rawc) {
  if (threadIdx.x < 64) {   // (the caller synchronises the CTA)
    const uint32_t a = threadIdx.x >> 3, b = threadIdx.x & 7u;
    const uint32_t e = sw2_lut_index(a, b);
    s_tab[e] = make_uint2(sw2_tab_word(a, sc.match, sc.mismatch), sw2_tab_word(b, sc.match, sc.mismatch));
    s_wsel[e] = sw2_wsel_word(a, b);
    s_rawc[e] = a | (b << 8);
  }
}

template <int C, int LANES>
__global__ void __launch_bounds__(SW_WARPS * 32)
sw_score2_kernel(const Scoring sc, const SeqSrc src, const smb_sw_task *__restrict__ tasks,
                 const SwClassArgs cls, int32_t *__restrict__ scores, int32_t *__restrict__ errs) {
  constexpr int GROUPS = 32 / LANES;
  // per window row (SW2_PAD padding rows on either side): byte offset of the row's entry in s_tab
  __shared__ unsigned short s_po[SW_WARPS * GROUPS][SW2_MAXROWS + 2 * SW2_PAD];
  __shared__ __align__(16) uint2 s_tab[SW2_LUT_N];
  __shared__ uint32_t s_wsel[SW2_LUT_N], s_rawc[SW2_LUT_N];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & (LANES - 1);
  unsigned short *const spo = s_po[threadIdx.x / LANES];
  const uint32_t nge2 = (uint32_t)((-sc.gap_ext) & 0xffff) * 0x10001u;
  // The table {match, mismatch x3 | 0 x4} and the bias are read back from shared memory so
  // that they live in (vector) registers: as kernel-uniform values ptxas keeps them in uniform
  // registers / as immediates and re-materialises both for every cell (IMAD.U32 from UR + PRMT of
  // RZ: two extra instructions per cell pair).
  __shared__ uint32_t s_konst[2];
  if (threadIdx.x == 0) {
    s_konst[0] = (uint32_t)(sc.match & 0xff) | ((uint32_t)(sc.mismatch & 0xff) * 0x01010100u);
    s_konst[1] = (uint32_t)(sc.gap_init & 0xffff) * 0x10001u;
  }
  sw2_build_lut(sc, s_tab, s_wsel, s_rawc);
  __syncthreads();
  const uint32_t T0 = ((volatile uint32_t *)s_konst)[0], gi2 = ((volatile uint32_t *)s_konst)[1];
  const char *const lutb = (const char *)s_tab, *const wselb = (const char *)s_wsel, *const rawb = (const char *)s_rawc;
  const int npairs = (cls.ntasks + 1) >> 1;

  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(cls.counter, 1);
    k = __shfl_sync(FULL, k, 0, LANES);
    const bool alive = k < npairs;
    if (!__any_sync(FULL, alive)) break;
    const int tixA = alive ? __ldg(cls.order + 2 * k) : 0;
    const bool haveB = alive && 2 * k + 1 < cls.ntasks;
    const int tixB = haveB ? __ldg(cls.order + 2 * k + 1) : tixA;
    const smb_sw_task ta = tasks[tixA], tb = tasks[tixB];
    const int qlenA = alive ? (int)ta.read_len : 0, qlenB = alive ? (int)tb.read_len : 0;
    const int rlenA = alive ? (int)ta.ref_len : 0, rlenB = alive ? (int)tb.ref_len : 0;
    int rlen = max(rlenA, rlenB);                                      // of this group
    if (GROUPS > 1) rlen = max(rlen, __shfl_xor_sync(FULL, rlen, 16)); // ... of the warp: common trip count
    const bool rcA = (ta.flags & SMB_TASK_READ_REVCOMP) != 0, rcB = (tb.flags & SMB_TASK_READ_REVCOMP) != 0;
    const bool pkA = (ta.flags & SMB_TASK_REF_PACKED) != 0, pkB = (tb.flags & SMB_TASK_REF_PACKED) != 0;
    bool hasX = false, hasN = false;
    __syncwarp();
    for (int x = lane; x < rlen + 2 * SW2_PAD; x += LANES) {
      const int i = x - SW2_PAD;
      const uint32_t a = (i >= 0 && i < rlenA) ? ref_base(src, pkA, ta.ref_off, (uint32_t)i) : 7u;
      const uint32_t b = (i >= 0 && i < rlenB) ? ref_base(src, pkB, tb.ref_off, (uint32_t)i) : 7u;
      hasX |= (a == 4u) | (b == 4u);
      spo[x] = (unsigned short)(sw2_lut_index(a, b) * 8u);
    }
    // per column: the PRMT selector of the row-table form {qA, qA | 8, 4 + qB, (4 + qB) | 8}; padding columns
    // select the sign of byte 0 / byte 4 twice (a score of 0 or -1 in every row)
    uint32_t qsel[C], H[C], E[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = lane * C + c;
      const uint32_t qa = (j < qlenA) ? read_base(src.arena, ta.read_off, (uint32_t)qlenA, rcA, (uint32_t)j) : 8u;
      const uint32_t qb = (j < qlenB) ? read_base(src.arena, tb.read_off, (uint32_t)qlenB, rcB, (uint32_t)j) : 8u;
      hasX |= (qa == 4u) | (qb == 4u);
      hasN |= ((qa > 4u) & (qa < 8u)) | ((qb > 4u) & (qb < 8u));
      qsel[c] = sw2_qsel_tab(qa, qb);
      H[c] = gi2;
      E[c] = gi2;
    }
    // X (the mismatch-against-everything code, score.c:138-173) in a read or a window: the warp's
    // pairs take the per-cell table path; an N in a read: the masked form
    const bool general = __any_sync(FULL, hasX);
    const bool masked = __any_sync(FULL, hasN);
    __syncwarp();
    uint32_t hdiag = gi2, hout = gi2, fout = gi2, best = gi2;
    const int nsteps = rlen + LANES - 1;
    const unsigned short *rowp = spo + (SW2_PAD - lane);
    if (!general && !masked) {
      // Every lane computes in every step: before its first and behind its last window row it
      // runs over the padding rows, which score 0 in every column - H stays 0 in front of the
      // window and cannot rise behind it, E and F only matter where they are positive - so the
      // maximum is that of the window rows alone and the loop needs no activity predicate.
      // (the row's table is fetched one step ahead: two dependent loads in front of the first PRMT otherwise;
      // rowp[nsteps] is a staged padding row)
      uint2 rt = *(const uint2 *)(lutb + rowp[0]);
      for (int t = 0; t < nsteps; ++t) {
        const uint2 rtn = *(const uint2 *)(lutb + rowp[t + 1]);
        uint32_t hl = __shfl_up_sync(FULL, hout, 1, LANES);
        uint32_t F = __shfl_up_sync(FULL, fout, 1, LANES);
        if (lane == 0) { hl = gi2; F = gi2; }
        uint32_t diag = hdiag;
        hdiag = hl;
        SW2_STEP(prmt(rt.x, rt.y, qsel[c]));
        hout = H[C - 1];
        fout = F;
        rt = rtn;
      }
    } else if (!general) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = lane * C + c;
        const uint32_t qa = (j < qlenA) ? read_base(src.arena, ta.read_off, (uint32_t)qlenA, rcA, (uint32_t)j) : 7u;
        const uint32_t qb = (j < qlenB) ? read_base(src.arena, tb.read_off, (uint32_t)qlenB, rcB, (uint32_t)j) : 7u;
        qsel[c] = sw2_qsel_masked(qa, qb);
      }
      for (int t = 0; t < nsteps; ++t) {
        uint32_t hl = __shfl_up_sync(FULL, hout, 1, LANES);
        uint32_t F = __shfl_up_sync(FULL, fout, 1, LANES);
        if (lane == 0) { hl = gi2; F = gi2; }
        const uint32_t w = *(const uint32_t *)(wselb + (rowp[t] >> 1));
        const uint32_t wm = w >> 16;
        uint32_t diag = hdiag;
        hdiag = hl;
        SW2_STEP(prmt(0u, T0, (qsel[c] ^ w) & ~wm));
        hout = H[C - 1];
        fout = F;
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {   // raw base codes of the columns
        const int j = lane * C + c;
        const uint32_t qa = (j < qlenA) ? read_base(src.arena, ta.read_off, (uint32_t)qlenA, rcA, (uint32_t)j) : 7u;
        const uint32_t qb = (j < qlenB) ? read_base(src.arena, tb.read_off, (uint32_t)qlenB, rcB, (uint32_t)j) : 7u;
        qsel[c] = qa | (qb << 8);
      }
      for (int t = 0; t < nsteps; ++t) {
        uint32_t hl = __shfl_up_sync(FULL, hout, 1, LANES);
        uint32_t F = __shfl_up_sync(FULL, fout, 1, LANES);
        if (lane == 0) { hl = gi2; F = gi2; }
        const int i = t - lane;
        if (i >= 0 && i < rlen) {
          const uint32_t r2 = *(const uint32_t *)(rawb + (spo[i + SW2_PAD] >> 1));
          const uint32_t ra = r2 & 0xffu, rb = r2 >> 8;
          uint32_t diag = hdiag;
          hdiag = hl;
          SW2_STEP((((uint32_t)(int)sc.S[ra * 8u + (qsel[c] & 0xffu)]) & 0xffffu) |
                   ((uint32_t)(int)sc.S[rb * 8u + (qsel[c] >> 8)] << 16));
          hout = H[C - 1];
          fout = F;
        }
      }
    }
    int bA = (int)(short)(best & 0xffffu) - sc.gap_init, bB = (int)(short)(best >> 16) - sc.gap_init;
    for (int o = LANES / 2; o > 0; o >>= 1) {
      bA = max(bA, __shfl_xor_sync(FULL, bA, o, LANES));
      bB = max(bB, __shfl_xor_sync(FULL, bB, o, LANES));
    }
    if (lane == 0 && alive) {
      scores[tixA] = bA;
      errs[tixA] = SMB_OK;
      if (haveB) { scores[tixB] = bB; errs[tixB] = SMB_OK; }
    }
  }
}

// ------------------------------------------------------------------------------------
// sw_long2_kernel: the packed recurrence of sw_score2_kernel for LONG reads (more columns than a warp
// holds in registers): two tasks per warp in the 16-bit halves (scores <= qlen * match <= 16000), the
// read cut into column blocks of 32 lanes x C columns, every block streaming all window rows through
// the systolic array.  What a block's last column leaves for the next one - H and F of every row,
// both tasks packed - goes through a strip in HBM / L2 (8 bytes per row, written and read 32 rows at a
// time).  Window rows are staged in shared memory RCH steps at a time as PRMT selectors, with the
// zero-scoring padding rows in front of and behind the window that let every lane compute in every
// step (see sw_score2_kernel).  swSIMDAlignStriped (swsimd.c:868-933) on 5-10 kb reads: ~36 such tasks
// per read with 7 * 10^7 cells each.
// ------------------------------------------------------------------------------------
constexpr int SWL_RCH = 512;     // steps per staged chunk of window rows

template <int C>
__global__ void __launch_bounds__(SW_WARPS * 32)
sw_long2_kernel(const Scoring sc, const SeqSrc src, const smb_sw_task *__restrict__ tasks,
                const SwClassArgs cls, int32_t *__restrict__ scores, int32_t *__restrict__ errs,
                uint2 *__restrict__ bscratch, const uint32_t bstride) {
  __shared__ unsigned short s_po[SW_WARPS][SWL_RCH + 32];   // rows as byte offsets into s_tab (see sw_score2_kernel)
  __shared__ __align__(16) uint2 s_tab[SW2_LUT_N];
  __shared__ uint32_t s_wsel[SW2_LUT_N], s_rawc[SW2_LUT_N];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  unsigned short *const spo = s_po[threadIdx.x >> 5];
  const int gwarp = blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
  uint2 *const strip0 = bscratch + (size_t)gwarp * 2u * bstride;
  const uint32_t nge2 = (uint32_t)((-sc.gap_ext) & 0xffff) * 0x10001u;
  __shared__ uint32_t s_konst[2];
  if (threadIdx.x == 0) {
    s_konst[0] = (uint32_t)(sc.match & 0xff) | ((uint32_t)(sc.mismatch & 0xff) * 0x01010100u);
    s_konst[1] = (uint32_t)(sc.gap_init & 0xffff) * 0x10001u;
  }
  sw2_build_lut(sc, s_tab, s_wsel, s_rawc);
  __syncthreads();
  const uint32_t T0 = ((volatile uint32_t *)s_konst)[0], gi2 = ((volatile uint32_t *)s_konst)[1];
  const char *const lutb = (const char *)s_tab, *const wselb = (const char *)s_wsel, *const rawb = (const char *)s_rawc;
  const int npairs = (cls.ntasks + 1) >> 1;

  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(cls.counter, 1);
    k = __shfl_sync(FULL, k, 0);
    if (k >= npairs) break;
    const int tixA = __ldg(cls.order + 2 * k);
    const bool haveB = 2 * k + 1 < cls.ntasks;
    const int tixB = haveB ? __ldg(cls.order + 2 * k + 1) : tixA;
    const smb_sw_task ta = tasks[tixA], tb = tasks[tixB];
    const int qlenA = (int)ta.read_len, qlenB = (int)tb.read_len;
    const int rlenA = (int)ta.ref_len, rlenB = (int)tb.ref_len;
    const int rlen = max(rlenA, rlenB), qlen = max(qlenA, qlenB);
    const bool rcA = (ta.flags & SMB_TASK_READ_REVCOMP) != 0, rcB = (tb.flags & SMB_TASK_READ_REVCOMP) != 0;
    const bool pkA = (ta.flags & SMB_TASK_REF_PACKED) != 0, pkB = (tb.flags & SMB_TASK_REF_PACKED) != 0;
    const int nblk = (qlen + 32 * C - 1) / (32 * C);
    const int nsteps = rlen + 31;
    uint32_t best = gi2;
    // X (the mismatch-against-everything code) anywhere in the reads or windows: per-cell table path
    // ... an N in a read: the masked form; else one PRMT over the row's score table per cell pair
    bool hasX = false, hasN = false;
    for (int x = lane; x < max(rlen, qlen); x += 32) {
      if (x < rlenA) hasX |= ref_base(src, pkA, ta.ref_off, (uint32_t)x) == 4u;
      if (x < rlenB) hasX |= ref_base(src, pkB, tb.ref_off, (uint32_t)x) == 4u;
      if (x < qlenA) {
        const uint32_t q = read_base(src.arena, ta.read_off, (uint32_t)qlenA, rcA, (uint32_t)x);
        hasX |= q == 4u; hasN |= q > 4u;
      }
      if (x < qlenB) {
        const uint32_t q = read_base(src.arena, tb.read_off, (uint32_t)qlenB, rcB, (uint32_t)x);
        hasX |= q == 4u; hasN |= q > 4u;
      }
    }
    const bool general = __any_sync(FULL, hasX);
    const int mode = general ? 2 : (__any_sync(FULL, hasN) ? 1 : 0);

    for (int b = 0; b < nblk; ++b) {
      uint32_t qsel[C], H[C], E[C];   // qsel: the column's selector in the form of `mode` (raw codes for mode 2)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = b * 32 * C + lane * C + c;
        const uint32_t qa = (j < qlenA) ? read_base(src.arena, ta.read_off, (uint32_t)qlenA, rcA, (uint32_t)j) : 8u;
        const uint32_t qb = (j < qlenB) ? read_base(src.arena, tb.read_off, (uint32_t)qlenB, rcB, (uint32_t)j) : 8u;
        qsel[c] = mode == 0 ? sw2_qsel_tab(qa, qb)
                            : (mode == 1 ? sw2_qsel_masked(qa, qb) : ((qa < 8u ? qa : 7u) | ((qb < 8u ? qb : 7u) << 8)));
        H[c] = gi2;
        E[c] = gi2;
      }
      uint32_t hdiag = gi2, hout = gi2, fout = gi2;
      uint2 bin = make_uint2(gi2, gi2), keep = make_uint2(gi2, gi2);
      const uint2 *bin_strip = strip0 + (size_t)((b & 1) ^ 1) * bstride;
      uint2 *bout_strip = strip0 + (size_t)(b & 1) * bstride;
      const bool has_in = b > 0, has_out = b < nblk - 1;

      for (int base = 0; base < nsteps; base += SWL_RCH) {
        // rows base - 31 .. base + RCH - 1 of both windows as selectors (padding outside the windows)
        __syncwarp();
        for (int x = lane; x < SWL_RCH + 31; x += 32) {
          const int i = base - 31 + x;
          const uint32_t a = (i >= 0 && i < rlenA) ? ref_base(src, pkA, ta.ref_off, (uint32_t)i) : 7u;
          const uint32_t bb = (i >= 0 && i < rlenB) ? ref_base(src, pkB, tb.ref_off, (uint32_t)i) : 7u;
          spo[x] = (unsigned short)(sw2_lut_index(a, bb) * 8u);
        }
        __syncwarp();
        const int tend = min(nsteps, base + SWL_RCH);
        const unsigned short *rowp = spo + (31 - lane) - base;
        for (int t = base; t < tend; ++t) {
          if ((t & 31) == 0 && has_in) {
            const int i = t + lane;
            bin = (i < rlen) ? __ldcg(bin_strip + i) : make_uint2(gi2, gi2);
          }
          uint32_t hl = __shfl_up_sync(FULL, hout, 1);
          uint32_t F = __shfl_up_sync(FULL, fout, 1);
          if (has_in) {
            const uint32_t hb = __shfl_sync(FULL, bin.x, t & 31), fb = __shfl_sync(FULL, bin.y, t & 31);
            if (lane == 0) { hl = hb; F = fb; }
          } else if (lane == 0) {
            hl = gi2;
            F = gi2;
          }
          uint32_t diag = hdiag;
          hdiag = hl;
          if (mode == 0) {
            const uint2 rt = *(const uint2 *)(lutb + rowp[t]);
            SW2_STEP(prmt(rt.x, rt.y, qsel[c]));
          } else if (mode == 1) {
            const uint32_t w = *(const uint32_t *)(wselb + (rowp[t] >> 1));
            const uint32_t wm = w >> 16;
            SW2_STEP(prmt(0u, T0, (qsel[c] ^ w) & ~wm));
          } else {
            const uint32_t r2 = *(const uint32_t *)(rawb + (rowp[t] >> 1));
            const uint32_t ra = r2 & 0xffu, rb = r2 >> 8;
            SW2_STEP((((uint32_t)(int)sc.S[ra * 8u + (qsel[c] & 0xffu)]) & 0xffffu) |
                     ((uint32_t)(int)sc.S[rb * 8u + (qsel[c] >> 8)] << 16));
          }
          hout = H[C - 1];
          fout = F;
          if (has_out) {   // the block's last column (lane 31, row t - 31) for the next column block
            const uint32_t ho = __shfl_sync(FULL, hout, 31), fo = __shfl_sync(FULL, fout, 31);
            const int i31 = t - 31;
            if (i31 >= 0) {
              if (lane == (i31 & 31)) keep = make_uint2(ho, fo);
              if ((i31 & 31) == 31 || i31 == rlen - 1) {
                const int rb0 = i31 & ~31;
                if (rb0 + lane <= i31) __stcg(bout_strip + rb0 + lane, keep);
              }
            }
          }
        }
      }
      __syncwarp();
    }
    int bA = (int)(short)(best & 0xffffu) - sc.gap_init, bB = (int)(short)(best >> 16) - sc.gap_init;
    bA = __reduce_max_sync(FULL, bA);
    bB = __reduce_max_sync(FULL, bB);
    if (lane == 0) {
      scores[tixA] = bA;
      errs[tixA] = SMB_OK;
      if (haveB) { scores[tixB] = bB; errs[tixB] = SMB_OK; }
    }
  }
}

// C: columns per lane of the 32-lane form (ceil(qlen / 32)); the half-warp form owns 2C per lane
template <int C>
static cudaError_t launch_class2(const Scoring &sc, const SeqSrc &src, const smb_sw_task *d_tasks,
                                 const SwClassArgs &cls, int32_t *d_scores, int32_t *d_errs, int grid,
                                 int sm_count, int ntasks_cls, cudaStream_t st) {
  static const bool lanes32 = getenv("SMB_SW_LANES32") != nullptr;
  if (lanes32) sw_score2_kernel<C, 32><<<grid, SW_WARPS * 32, 0, st>>>(sc, src, d_tasks, cls, d_scores, d_errs);
  else {
    // the half-warp form: 64 registers and 10 KB of shared memory per CTA leave room for 8 CTAs per SM; 6 measured
    // best (26.4 ms per 1 M C2 reads; 4: 27.3, 8: 26.9 - more warps hide the step's loads, fewer end more evenly)
    static const int ctas_per_sm = getenv("SMB_SW_CTAS_PER_SM") ? atoi(getenv("SMB_SW_CTAS_PER_SM")) : 6;
    const int cap = sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 6);
    const int want = (ntasks_cls + 3) / 4 / SW_WARPS + 1;   // four tasks per warp
    sw_score2_kernel<2 * C, 16><<<want < cap ? want : cap, SW_WARPS * 32, 0, st>>>(sc, src, d_tasks, cls, d_scores, d_errs);
  }
  return cudaGetLastError();
}

// Host-side plan: tasks bucketed by columns-per-lane class (one launch per class).
void plan_sw(const smb_sw_task *h_tasks, int ntasks, int sm_count, const Scoring &sc, SwPlan &plan) {
  constexpr int NCLS = 8;
  plan.order.resize((size_t)ntasks);
  uint32_t max_rlen_multi = 0;
  auto cls_of = [](uint32_t qlen) {
    int c = (int)((qlen + 31) / 32);
    return c < 1 ? 1 : (c > NCLS ? NCLS : c);
  };
  // two-tasks-per-warp 16-bit kernel: single column block, staged window, scores that fit
  const bool pen16 = sc.match > 0 && sc.match < 128 && sc.mismatch <= 0 && sc.mismatch > -128 && sc.gap_init >= 0 &&
                     sc.gap_init < 8000 && sc.gap_ext >= 0 && sc.gap_ext < 8000 && sc.S[5] == 0 && sc.S[5 * 8] == 0;
  auto pair16 = [&](const smb_sw_task &t) {
    return pen16 && t.read_len <= 32u * NCLS && t.ref_len <= (uint32_t)SW2_MAXROWS &&
           (long long)t.read_len * sc.match <= 16000;
  };
  // long reads in 16-bit pairs (sw_long2_kernel)
  auto long16 = [&](const smb_sw_task &t) {
    return pen16 && t.read_len > 32u * NCLS && (long long)t.read_len * sc.match <= 16000;
  };
  // slots 1..8: 32-bit classes, 9..16: paired 16-bit classes, 17: long reads, paired
  auto slot_of = [&](const smb_sw_task &t) {
    const int c = cls_of(t.read_len);
    return long16(t) ? SW_SLOT_LONG2 : (pair16(t) ? c + NCLS : c);
  };
  for (int c = 0; c < SW_SLOTS; ++c) plan.count[c] = plan.start[c] = 0;
  for (int i = 0; i < ntasks; ++i) {
    plan.count[slot_of(h_tasks[i])]++;
    if (h_tasks[i].read_len > 32u * NCLS) max_rlen_multi = std::max(max_rlen_multi, h_tasks[i].ref_len);
  }
  for (int c = 1; c < SW_SLOTS; ++c) plan.start[c] = plan.start[c - 1] + plan.count[c - 1];
  int fill[SW_SLOTS];
  for (int c = 0; c < SW_SLOTS; ++c) fill[c] = plan.start[c];
  for (int i = 0; i < ntasks; ++i) plan.order[(size_t)fill[slot_of(h_tasks[i])]++] = i;
  // long reads first inside the multi-block classes (largest tasks start earliest)
  if (max_rlen_multi)
    for (const int c : {NCLS, (int)SW_SLOT_LONG2})
      std::stable_sort(plan.order.begin() + plan.start[c], plan.order.begin() + plan.start[c] + plan.count[c],
                       [&](int a, int b) {
                         return (uint64_t)h_tasks[a].read_len * h_tasks[a].ref_len >
                                (uint64_t)h_tasks[b].read_len * h_tasks[b].ref_len;
                       });
  plan.max_grid = sm_count * 8;
  plan.bstride = (max_rlen_multi + 31u) & ~31u;
  plan.strip_bytes = (size_t)plan.max_grid * SW_WARPS * 2u * plan.bstride * sizeof(int2);
}

// d_counters: 16 ints (zeroed here); d_order: plan.order on the device; d_strips: plan.strip_bytes
cudaError_t launch_sw_score(const Scoring &sc, const SeqSrc &src, const smb_sw_task *d_tasks,
                            const SwPlan &plan, int *d_counters, const int *d_order, void *d_strips,
                            int32_t *d_scores, int32_t *d_errs, cudaStream_t st, int *nlaunch) {
  constexpr int NCLS = 8;
  cudaError_t e;
  if ((e = cudaMemsetAsync(d_counters, 0, 32 * sizeof(int), st)) != cudaSuccess) return e;
  if (plan.count[SW_SLOT_LONG2]) {    // long reads, two tasks per warp
    const int n = plan.count[SW_SLOT_LONG2];
    SwClassArgs cls{d_order + plan.start[SW_SLOT_LONG2], n, d_counters + SW_SLOT_LONG2};
    int grid = ((n + 1) / 2 + SW_WARPS - 1) / SW_WARPS;
    if (grid > plan.max_grid) grid = plan.max_grid;
    sw_long2_kernel<16><<<grid, SW_WARPS * 32, 0, st>>>(sc, src, d_tasks, cls, d_scores, d_errs, (uint2 *)d_strips, plan.bstride);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*nlaunch;
  }
  for (int c = 1; c <= NCLS; ++c) {   // paired 16-bit classes
    const int n = plan.count[c + NCLS];
    if (!n) continue;
    SwClassArgs cls{d_order + plan.start[c + NCLS], n, d_counters + c + NCLS};
    int grid = ((n + 1) / 2 + SW_WARPS - 1) / SW_WARPS;
    if (grid > plan.max_grid) grid = plan.max_grid;
    switch (c) {
      case 1: e = launch_class2<1>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
      case 2: e = launch_class2<2>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
      case 3: e = launch_class2<3>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
      case 4: e = launch_class2<4>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
      case 5: e = launch_class2<5>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
      case 6: e = launch_class2<6>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
      case 7: e = launch_class2<7>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
      default: e = launch_class2<8>(sc, src, d_tasks, cls, d_scores, d_errs, grid, plan.max_grid / 8, n, st); break;
    }
    if (e != cudaSuccess) return e;
    ++*nlaunch;
  }
  for (int c = 1; c <= NCLS; ++c) {
    if (!plan.count[c]) continue;
    SwClassArgs cls{d_order + plan.start[c], plan.count[c], d_counters + c};
    int grid = (plan.count[c] + SW_WARPS - 1) / SW_WARPS;
    if (grid > plan.max_grid) grid = plan.max_grid;
    int2 *strips = (int2 *)d_strips;
    const uint32_t bstride = plan.bstride;
    switch (c) {
      case 1: e = launch_class<1>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
      case 2: e = launch_class<2>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
      case 3: e = launch_class<3>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
      case 4: e = launch_class<4>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
      case 5: e = launch_class<5>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
      case 6: e = launch_class<6>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
      case 7: e = launch_class<7>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
      default: e = launch_class<8>(sc, src, d_tasks, cls, d_scores, d_errs, strips, bstride, grid, st); break;
    }
    if (e != cudaSuccess) return e;
    ++*nlaunch;
  }
  return cudaSuccess;
}

cudaError_t warm_sw_long() {
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, sw_long2_kernel<16>);
}

cudaError_t warm_sw() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, sw_score_kernel<5>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, sw_score_kernel<4>);
  return e;
}

}  // namespace smb
