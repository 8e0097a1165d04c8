// band_pack.cu - K3 for short reads, FOUR TASKS PER WARP: the wavefront of band_warp.cu with
// two tasks packed into the 16-bit halves of every register (DPX / SIMD-in-register ops).
//
// Same function as band_warp_kernel / band_kernel<true> (aliSmiWatInBand,
// /root/reference/src/alignment.c:1548-1601), same results bit for bit.  Scores of short reads
// fit 15 bits (<= qlen * match), so H, E, F and the running maxima of TWO tasks share a
// register: task 0 in the low, task 1 in the high half-word.  A half-warp (16 lanes, two
// diagonals per lane, bands of at most 32 diagonals) advances both tasks with one instruction
// stream; the two half-warps of a warp run in lockstep (full-warp shuffles of width 16).
//
//  * restricted recurrence on packed values (alignment.c:885-982, restated in DESIGN.md):
//      h = diag + s;  m = max(E, F, 0);  dia = h > m;  H = max(h, m)
//      open = dia & (h > gap_init);  t = open ? h - gap_init : 0
//      E' = max(E - ext, t, 0);  F' = max(F - ext, t, 0)        (VIADDMNMX.S16x2.RELU)
//    Non-positive gap states mean "no gap" everywhere in the reference, so they are kept as 0;
//    comparisons become masks by sign replication of a packed difference (VIADD.16x2 + PRMT),
//    which is exact because all values stay below 2^14 in magnitude;
//  * the substitution score of both tasks comes from one PRMT over the byte table
//    {match, mismatch x3, 0 x4}: selector = (read selector) xor (window selector), where N and
//    padding select a zero entry through the sign-replication mode (see sel_read / sel_ref);
//  * per DP pass the window rows and read columns of both tasks are staged in shared memory as
//    PRMT selectors (decoded from whole source words); rows and columns outside the window / read
//    segment are staged as padding (score 0) and need no validity test (see the DP loop);
//  * direction codes: 4 bits per row and task (two diagonals of the lane), four rows per word;
//  * argmax, backtrace, DiffStr reversal, result emission and the recursion are per task as in
//    band_warp.cu, except that the eight lanes 8t..8t+7 of the half-warp walk the path of task t together.
// Tasks with an X base in the read or the window use the general table path for their scores.
#include "common.cuh"
#include "band.h"

namespace smb {

constexpr int BPK_WARPS = 2;          // warps per CTA (4 half-warp groups, 8 tasks)
// group geometry: LANES lanes advance two tasks, ND diagonals per lane (bands of at most LANES * ND
// diagonals): <16, 2> or <8, 3> (eight tasks per warp, for bands of at most 24 diagonals - the
// usual short-read band of 19 then keeps 7 of 8 lanes busy instead of 10 of 16)
constexpr int BPK_STACK = 24;
constexpr int BPK_ROWPAD = 16;        // rowarr index = r + BPK_ROWPAD

struct PackLayout {                   // per-group shared memory for windows of at most R rows, reads of at most Q bases
  int R;                              // multiple of 32
  int Q;                              // multiple of 16
  int lanes;                          // lanes per group
  int rpw;                            // rows per direction word (and task): 16 / (2 * ND) = 4 or 2
  __host__ __device__ int rowarr_n() const { return R + 64; }
  __host__ __device__ int colarr_n() const { return R + 64; }
  __host__ __device__ int dirw() const { return (R + lanes) / rpw + 1; }   // words per lane: one row per DP iteration
  __host__ __device__ int rev_n() const { return R + Q + 16; }
  // 16-bit entries {PRMT selector of task 0, of task 1}
  __host__ __device__ size_t rowarr_off() const { return 0; }
  __host__ __device__ size_t colarr_off() const { return (size_t)rowarr_n() * 2; }
  __host__ __device__ size_t dirs_off() const { return (colarr_off() + (size_t)colarr_n() * 2 + 15) & ~(size_t)15; }
  __host__ __device__ size_t stk_off() const { return dirs_off() + (size_t)lanes * dirw() * 4; }
  __host__ __device__ size_t rev_off() const { return stk_off() + (size_t)2 * 2 * BPK_STACK * 4; }
  __host__ __device__ size_t bytes() const { return (rev_off() + (size_t)2 * rev_n() + 15) & ~(size_t)15; }
};

__device__ __forceinline__ uint32_t bp_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// 0xffff in every half-word whose value is negative
__device__ __forceinline__ uint32_t bp_neg(uint32_t x) { return bp_prmt(x, 0u, 0xbb99u); }
// per half-word a > b (exact while |a - b| < 2^15)
__device__ __forceinline__ uint32_t bp_gt(uint32_t a, uint32_t b) { return bp_neg(__vsub2(b, a)); }

// PRMT selector byte (two nibbles: low byte, high byte of the 16-bit score) of a read base ...
// (bit 2 of both nibbles flipped: the table is the SECOND source of the PRMT, bytes 4..7, because as
// first source ptxas overwrites it with the result and copies it afresh for every cell)
// (X gets a code of its own, 0x81 / 0x4D: tasks with an X take the table path, which reads the base
// codes back from the selectors)
__device__ __forceinline__ uint32_t sel_read(uint32_t q) { return q < 4u ? ((q | ((q | 8u) << 4)) ^ 0x44u) : (q == 4u ? 0x81u : 0x80u); }
// ... and of a window base; N / padding: sign-replicate a non-negative table entry -> 0
__device__ __forceinline__ uint32_t sel_ref(uint32_t r) { return r < 4u ? (r | (r << 4)) : (r == 4u ? 0x4Du : 0x4Cu); }
// base codes from selector bytes (N and padding both come back as 7: the same zero row of the table)
__device__ __forceinline__ uint32_t code_read(uint32_t s) { return (s & 0x40u) ? (s & 3u) : ((s & 1u) ? 4u : 7u); }
__device__ __forceinline__ uint32_t code_ref(uint32_t s) { return (s & 0x40u) ? ((s & 1u) ? 4u : 7u) : (s & 3u); }

#define DIFFB(count, typ) ((uint8_t)((count) + ((typ) << 6)))

// two cells (task 0 / task 1) of the restricted recurrence; ok2 = 0xffff per half-word whose diagonal lies in the band.
// The ALU pipe (16 lanes per scheduler) bounds this kernel, so the cell is written for the fewest
// ALU instructions, with shifts and constant adds as IMADs (FMA pipe):
//  * there is no packed subtract or compare: h > m is the sign of h + ~m, made a mask by PRMT sign
//    replication;
//  * t = (dia && h > gap_init) ? h - gap_init : 0 = relu(h - gap_init) & dia & valid, one VIADDMNMX.RELU;
//  * only t is masked (by the band).  H, E and F of cells outside need no masks: padding scores 0,
//    E moves down a read column, F along a window row, H along a diagonal, and on each of those lines
//    the invalid cells come first (before the read segment / band) - where nothing but zeros can
//    arise without a t - or last, where no valid cell reads them;
//  * the running maximum of a diagonal is kept on t (same order as h, and maxima count only when
//    h > gap_init anyway, alignment.c:826-830) as the key (t << 8 | 255 - row): an unsigned packed max
//    keeps the FIRST row of the largest score - scores and rows are below 256 here;
//  * direction: DIA (3) if h > m, else 0 if m == 0, else COL (1) if E >= F, else ROW (2)
//    = min(m + m, min((m ^ E) + 1, 2)) | (dia ? 3 : 0), two VIADDMNMX.U16x2.  The backtrace reads the codes of valid cells only.
#define PACK_CELL(ok2, diag, ein, fin, s2, Hout, Eout, Fout, best, rinv2, dcode)                      \
  do {                                                                                                \
    const uint32_t h_ = __vadd2((diag), (s2));                                                        \
    const uint32_t m_ = __vmaxs2((ein), (fin));               /* E, F >= 0 */                          \
    const uint32_t ndia_ = bp_neg(__vadd2(h_, ~m_));          /* h <= m */                             \
    const uint32_t t_ = __viaddmax_s16x2_relu(h_, ngi2, ngi2) & ~ndia_ & (ok2);   /* (ngi2 < 0: any operand the RELU removes) */                         \
    const uint32_t sel_ = __viaddmin_u16x2(m_ ^ (ein), 0x00010001u, 0x00020002u);  /* E >= F ? 1 : 2 */  \
    (Eout) = __viaddmax_s16x2_relu((ein), nge2, t_);                                                  \
    (Fout) = __viaddmax_s16x2_relu((fin), nge2, t_);                                                  \
    (best) = __vmaxu2((best), t_ * 256u + (rinv2));                                                   \
    (dcode) = __viaddmin_u16x2(m_, m_, sel_) | (~ndia_ & 0x00030003u);                                \
    (Hout) = __vmaxs2(h_, m_);                                                                        \
  } while (0)

struct PackTask {           // group-uniform state of one of the two tasks of a group
  int tix, alive, err, sp, minscore, minscorlen, qlen, rlen;
  uint32_t nres, diff_used, dcap;
  uint64_t read_off, ref_off;
  int rc, packed;
  int l_edge0, r_edge0, p_left, p_right;
};

template <int LANES, int ND>
__global__ void __launch_bounds__(BPK_WARPS * 32)
band_pack_kernel(const Scoring sc, const SeqSrc src, const smb_band_task *__restrict__ tasks,
                 const int *__restrict__ order, const int ntasks, int *__restrict__ ticket,
                 BandOut out, const int max_res, const uint64_t *__restrict__ diff_off,
                 const uint32_t *__restrict__ diff_cap, const PackLayout lay) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr unsigned ALL = 0xffffffffu;
  constexpr int TEAM = LANES / 2;              // lanes walking one task's path
  constexpr int RPW = 8 / ND, RSH = RPW == 4 ? 2 : 1, BITS = 2 * ND;   // direction words: rows per half-word, bits per row
  static_assert((LANES == 16 && ND == 2) || (LANES == 8 && ND == 3), "group geometry");
  const int lane = threadIdx.x & (LANES - 1);
  unsigned char *base = s_raw + (size_t)(threadIdx.x / LANES) * lay.bytes();
  uint16_t *rowarr = (uint16_t *)(base + lay.rowarr_off());
  uint16_t *colarr = (uint16_t *)(base + lay.colarr_off());
  uint32_t *dirs = (uint32_t *)(base + lay.dirs_off());
  int *stk = (int *)(base + lay.stk_off());               // [task][l/r][BPK_STACK]
  uint8_t *revbase = base + lay.rev_off();
  const int DIRW = lay.dirw();
  const uint32_t ngi2 = (uint32_t)((-sc.gap_init) & 0xffff) * 0x10001u;
  const uint32_t nge2 = (uint32_t)((-sc.gap_ext) & 0xffff) * 0x10001u;
  const uint32_t T0 = (uint32_t)(sc.match & 0xff) | ((uint32_t)(sc.mismatch & 0xff) * 0x01010100u);
  const int npairs = (ntasks + 1) >> 1;
  unsigned long long ncell_tot = 0;

  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(ticket, 1);
    k = __shfl_sync(ALL, k, 0, LANES);
    if (!__any_sync(ALL, k < npairs)) break;
    PackTask T[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      PackTask &p = T[t];
      const int idx = 2 * k + t;
      p.alive = k < npairs && idx < ntasks;
      p.tix = p.alive ? __ldg(order + idx) : 0;
      const smb_band_task tk = tasks[p.tix];
      p.qlen = (int)tk.read_len; p.rlen = (int)tk.ref_len;
      p.read_off = tk.read_off; p.ref_off = tk.ref_off;
      p.rc = (tk.flags & SMB_TASK_READ_REVCOMP) != 0;
      p.packed = (tk.flags & SMB_TASK_REF_PACKED) != 0;
      p.l_edge0 = tk.l_edge; p.r_edge0 = tk.r_edge; p.p_left = tk.p_left; p.p_right = tk.p_right;
      p.minscore = tk.minscore; p.minscorlen = tk.minscorlen;
      p.err = SMB_OK; p.nres = 0; p.diff_used = 0; p.sp = 0;
      p.dcap = p.alive ? diff_cap[p.tix] : 0u;
      if (p.minscore < 1 || sc.match <= 0) p.err = SMB_ERRCODE_ASSERT;         // alignment.c:1569
      else {
        if (p.minscorlen * sc.match < p.minscore) p.minscorlen = p.minscore / sc.match;  // :1572
        if (p.minscorlen < 5) p.err = SMB_ERRCODE_ASSERT;                       // ALILEN_MIN :1574
      }
      if (p.alive && !p.err) {
        if (lane == 0) { stk[(t * 2 + 0) * BPK_STACK] = tk.u_left; stk[(t * 2 + 1) * BPK_STACK] = tk.u_right; }
        p.sp = 1;
      }
    }
    __syncwarp();

    // one round = one DP pass of every task that still has a row range on its stack
    for (;;) {
      bool on[2];
      on[0] = T[0].alive && T[0].sp > 0 && !T[0].err;
      on[1] = T[1].alive && T[1].sp > 0 && !T[1].err;
      if (!__any_sync(ALL, on[0] || on[1])) break;
      Band B[2];
      int s_left[2], s_right[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        PackTask &p = T[t];
        Band &b = B[t];
        b.band_width = 0; b.l_edge = 0; b.r_edge = 0; b.l_edge_orig = 0; b.r_edge_orig = 0;
        b.s_left = 0; b.s_len = 0; b.q_left = 0; b.q_len = 0;
        s_left[t] = s_right[t] = 0;
        if (on[t]) {
          --p.sp;
          s_left[t] = stk[(t * 2 + 0) * BPK_STACK + p.sp];
          s_right[t] = stk[(t * 2 + 1) * BPK_STACK + p.sp];
          if (band_init(b, p.l_edge0, p.r_edge0, p.p_left, p.p_right, p.qlen, s_left[t], s_right[t], p.rlen))
            on[t] = false;                                                   // :1333-1338
          else if (b.s_left >= b.s_len || b.band_width < 0) { p.err = SMB_ERRCODE_ASSERT; on[t] = false; }  // :459
          else if (b.band_width > ND * LANES || b.s_len - b.s_left > lay.R) { p.err = SMB_ERR_ARG; on[t] = false; }
        }
      }
      const int nrows0 = on[0] ? B[0].s_len - B[0].s_left : 0, nrows1 = on[1] ? B[1].s_len - B[1].s_left : 0;
      const int bw0 = on[0] ? B[0].band_width : 0, bw1 = on[1] ? B[1].band_width : 0;
      int iters = max(on[0] ? nrows0 + (bw0 + ND - 1) / ND - 1 : 0, on[1] ? nrows1 + (bw1 + ND - 1) / ND - 1 : 0);
#pragma unroll
      for (int o = LANES; o < 32; o <<= 1) iters = max(iters, __shfl_xor_sync(ALL, iters, o));
      iters = (iters + RPW - 1) & ~(RPW - 1);                  // whole direction words (the extra rows are padding)
      // ---- stage rows and columns: PRMT selectors, task t in byte t of a 16-bit entry ----
      // Everything the trip count (the maximum over the groups of the warp) can touch is padding first;
      // then whole source words are decoded: ten window bases per word of the packed reference, four
      // read bases per aligned word of the arena.
      bool hasx = false;
      const int nrow_e = min(iters + 2 * BPK_ROWPAD, lay.rowarr_n()), ncol_e = min(iters + (ND - 1) * LANES + 2, lay.colarr_n());
      uint8_t *const rowsel = (uint8_t *)rowarr, *const colsel = (uint8_t *)colarr;
      __syncwarp();
      for (int e = lane; e < (nrow_e + 1) >> 1; e += LANES) ((uint32_t *)rowarr)[e] = 0x4C4C4C4Cu;   // sel_ref(7)
      for (int e = lane; e < (ncol_e + 1) >> 1; e += LANES) ((uint32_t *)colarr)[e] = 0x80808080u;   // sel_read(7)
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (!on[t]) continue;
        const PackTask &p = T[t];
        const Band &b = B[t];
        const int nr = t ? nrows1 : nrows0;
        if (p.packed) {
          const uint64_t b0 = p.ref_off + (uint64_t)b.s_left, w0 = b0 / 10u;
          const int r0 = (int)(b0 - w0 * 10u);
          const int nw = (r0 + nr + 9) / 10;
          for (int k = lane; k < nw; k += LANES) {
            const uint32_t w = __ldg(src.packed + w0 + (uint64_t)k);
            const int rb = k * 10 - r0;
#pragma unroll
            for (int q = 0; q < 10; ++q) {
              const uint32_t c = (w >> (27 - 3 * q)) & 7u;
              if ((unsigned)(rb + q) < (unsigned)nr) {
                rowsel[2 * (rb + q + BPK_ROWPAD) + t] = (uint8_t)sel_ref(c);
                hasx |= c == 4u;
              }
            }
          }
        } else {
          for (int r = lane; r < nr; r += LANES) {
            const uint32_t c = (uint32_t)(__ldg(src.arena + p.ref_off + (uint64_t)(b.s_left + r)) & 7u);
            rowsel[2 * (r + BPK_ROWPAD) + t] = (uint8_t)sel_ref(c);
            hasx |= c == 4u;
          }
        }
        // read columns j in [jlo, jhi): source bytes s = j, or qlen - 1 - j for the reverse complement
        const int jlo = max(b.q_left, b.l_edge), jhi = min(b.q_len, b.l_edge + ncol_e);
        if (jhi > jlo) {
          const int s_lo = p.rc ? p.qlen - jhi : jlo, s_hi = p.rc ? p.qlen - jlo : jhi;
          const uint64_t a0 = (p.read_off + (uint64_t)s_lo) & ~(uint64_t)3;
          const int sb = (int)((long long)a0 - (long long)p.read_off);      // source index of byte 0 of word 0
          const int nwd = (s_hi - sb + 3) >> 2;
          const uint32_t *const wp = (const uint32_t *)(src.arena + a0);
          for (int k = lane; k < nwd; k += LANES) {
            const uint32_t w = __ldg(wp + k);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int si = sb + 4 * k + q;
              uint32_t c = (w >> (8 * q)) & 7u;
              if (si >= s_lo && si < s_hi) {
                const int j = p.rc ? p.qlen - 1 - si : si;
                if (p.rc && c < 4u) c = 3u - c;
                colsel[2 * (j - b.l_edge) + t] = (uint8_t)sel_read(c);
                hasx |= c == 4u;
              }
            }
          }
        }
      }
      const bool general = __any_sync(ALL, hasx);
      __syncwarp();

      // ---------------- packed wavefront DP ----------------
      // lane l owns diagonals ND*l .. ND*l + ND-1 and computes row it - l of each in iteration it:
      // cell (r, d) reads H(r-1, d) (own register), E(r-1, d+1) (own previous value of the next diagonal,
      // or the right neighbour's first diagonal of this iteration) and F(r, d-1) (own value just computed,
      // or the left neighbour's last diagonal of the previous iteration)
      // Cells outside the window rows or the read segment need no masks: they are staged as padding
      // (score 0).  Above and to the left of the valid cells nothing but zeros can arise (t = relu(0 - gap_init));
      // below and to the right a t can arise, but E, F and H there only flow on to other such cells, and
      // their keys lose against the valid cell their score came from (same t at an earlier row of the
      // same diagonal, or a larger t if a gap lies between: gap_init > 0).  Only the diagonals beyond the
      // band (has) must not open gaps, because E moves from diagonal d + 1 to d.
      uint32_t has[ND], H[ND], E[ND], F[ND], best[ND], col[ND];
#pragma unroll
      for (int x = 0; x < ND; ++x) {
        const int dx = ND * lane + x;
        has[x] = (dx < bw0 ? 0xffffu : 0u) | (dx < bw1 ? 0xffff0000u : 0u);
        H[x] = E[x] = F[x] = best[x] = 0;
        col[x] = 0;
      }
      uint32_t *const dirp = dirs + lane * DIRW;
      const uint32_t not_first = lane == 0 ? 0u : 0xffffffffu, not_last = lane == LANES - 1 ? 0u : 0xffffffffu;
#pragma unroll
      for (int x = 0; x + 1 < ND; ++x) col[x] = colarr[(ND - 1) * lane + x];
      // direction words hold the RPW rows of RPW consecutive ITERATIONS (row it - lane): the position of a
      // row in its word is then the same for all lanes and known at compile time
      for (int it0 = 0; it0 < iters; it0 += RPW) {
        uint32_t wdir = 0;
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
          const int it = it0 + k;
          const int r = it - lane;
          const uint32_t Fin = __shfl_up_sync(ALL, F[ND - 1], 1, LANES) & not_first;
          const uint32_t rw = rowarr[r + BPK_ROWPAD];
          col[ND - 1] = colarr[it + (ND - 1) * lane + ND - 1];
          uint32_t sx[ND], dc[ND];
          if (!general) {   // (PRMT reads the low 16 bits of the selector only)
#pragma unroll
            for (int x = 0; x < ND; ++x) sx[x] = bp_prmt(0u, T0, col[x] ^ rw);
          } else {   // X bases: per-cell table look-ups
            const uint32_t r0 = code_ref(rw & 0xffu), r1 = code_ref((rw >> 8) & 0xffu);
#pragma unroll
            for (int x = 0; x < ND; ++x) {
              const int a0 = sc.S[r0 * 8u + code_read(col[x] & 0xffu)], a1 = sc.S[r1 * 8u + code_read((col[x] >> 8) & 0xffu)];
              sx[x] = ((uint32_t)a0 & 0xffffu) | ((uint32_t)a1 << 16);
            }
          }
          // key of the row: 255 - r (rows beyond 255 do not exist; the drain iterations get 0)
          const uint32_t rinv2 = (uint32_t)__viaddmin_s32_relu(lane + 255 - it0, -k, 255) * 0x10001u;
#pragma unroll
          for (int x = 0; x < ND; ++x) {
            uint32_t ein, fin;
            if (x + 1 < ND) ein = E[x + 1];
            else ein = __shfl_down_sync(ALL, E[0], 1, LANES) & not_last;
            if (x == 0) fin = Fin;
            else fin = F[x - 1];
            PACK_CELL(has[x], H[x], ein, fin, sx[x], H[x], E[x], F[x], best[x], rinv2, dc[x]);
          }
#pragma unroll
          for (int x = 0; x + 1 < ND; ++x) col[x] = col[x + 1];
          uint32_t code = dc[0];
#pragma unroll
          for (int x = 1; x < ND; ++x) code |= dc[x] << (2 * x);
          wdir |= code << (k * BITS);
        }
        dirp[it0 >> RSH] = wdir;
      }
      __syncwarp();
      // cells of this pass (statistics): per diagonal the rows whose read column lies in the segment
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (on[t]) {
          const Band &b = B[t];
          const int nr = t ? nrows1 : nrows0;
#pragma unroll
          for (int x = 0; x < ND; ++x) {
            const int dx = ND * lane + x;
            const int lo = max(0, b.q_left - b.l_edge - dx), hi = min(nr, b.q_len - b.l_edge - dx);
            if (dx < b.band_width && hi > lo) ncell_tot += (unsigned)(hi - lo);
          }
        }
      __syncwarp();

      // ---- per task: argmax, backtrace, result, recursion (lane t serves task t) ----
      int max_scor[2], max_i[2], max_j[2], max_r[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const Band &b = B[t];
        const int sh = 16 * t;
        int bsc = 0, bestr = 0, bestd = 0;
#pragma unroll
        for (int x = 0; x < ND; ++x) {
          const uint32_t kx = (best[x] >> sh) & 0xffffu;
          const int bx = (kx >> 8) ? (int)(kx >> 8) + sc.gap_init : 0, rx = 255 - (int)(kx & 0xffu);
          if (x == 0 || bx > bsc || (bx == bsc && bx > 0 && rx < bestr)) { bsc = bx; bestr = rx; bestd = ND * lane + x; }
        }
        unsigned long long key = 0;
        if (on[t] && bsc > 0)
          key = ((unsigned long long)(unsigned)bsc << 32) | ((unsigned long long)(0xffffu - (unsigned)bestr) << 16) |
                (unsigned long long)(0xffffu - (unsigned)(b.l_edge + bestr + bestd - b.q_left));
        for (int o = LANES / 2; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(ALL, key, o, LANES);
          key = other > key ? other : key;
        }
        max_scor[t] = (int)(key >> 32);
        max_r[t] = (int)(0xffffu - (unsigned)((key >> 16) & 0xffffu));
        max_j[t] = (int)(0xffffu - (unsigned)(key & 0xffffu)) + b.q_left;
        max_i[t] = b.s_left + max_r[t];
        if (max_scor[t] < T[t].minscore) on[t] = false;                    // :1364
      }
      // makeMetaFromTrack (alignment.c:628-781): the TEAM lanes t*TEAM.. of the group walk the path of task t
      // together.  Sub-lane L looks at the cell L steps up the current diagonal; the run of cells
      // that continue as "DIA move onto a match" (ballot) is consumed at once - the reference's
      // per-match counter (`nmatch > 61 ? emit 61 : ++nmatch`, i.e. 62 rolls over to 1 with one
      // emission) has the closed form below - and the first cell that is anything else takes the
      // reference's step, with the values of the sub-lane that looked at it.
      int bi = 0, bj = 0, flag = 0;
      uint32_t bn = 0;
      {
        const int t = lane / TEAM, L = lane & (TEAM - 1);
        const int gbase = (int)(threadIdx.x & 31u & ~(unsigned)(TEAM - 1));   // first lane of this team within the warp
        const Band &b = t ? B[1] : B[0];
        const int mscor = t ? max_scor[1] : max_scor[0];
        uint8_t *rev = revbase + (size_t)t * lay.rev_n();
        const uint32_t revcap = (uint32_t)lay.rev_n();
        const int sh = 16 * t;
        int i = t ? max_i[1] : max_i[0], j = t ? max_j[1] : max_j[0];
        int r = t ? max_r[1] : max_r[0], d = j - b.l_edge - r;
        bool act = t ? on[1] : on[0];
        const bool mine = act;
        bool gap_open = false, ovf = false;
        unsigned nmatch = 0;
        int checksum = 0;
        uint32_t n = 0;
#define EMIT(c, ty) do { if (n < revcap) { if (L == 0) rev[n] = DIFFB(c, ty); } else ovf = true; ++n; } while (0)
        while (__any_sync(ALL, act)) {
          const int rr = r - L, itw = rr + d / ND;          // direction words are laid out by DP iteration
          const bool valid = act && i - L >= b.s_left && j - L >= b.q_left;
          uint32_t w = 0, csel = 0, rsel = 0;
          if (valid) {
            w = dirs[(d / ND) * DIRW + (itw >> RSH)];
            csel = colarr[j - L - b.l_edge];
            rsel = rowarr[rr + BPK_ROWPAD];
          }
          const uint32_t dir = (w >> (sh + (uint32_t)(itw & (RPW - 1)) * BITS + ((uint32_t)(d % ND) << 1))) & 3u;
          int s = (int)(short)(bp_prmt(0u, T0, csel ^ rsel) >> sh);
          if (general && valid)
            s = (int)sc.S[code_ref((rsel >> (8 * t)) & 0xffu) * 8u + code_read((csel >> (8 * t)) & 0xffu)];
          const bool fast = valid && dir == 3u && s > 0 && !general;
          const uint32_t m8 = (__ballot_sync(ALL, fast) >> gbase) & ((1u << TEAM) - 1u);
          const int run = __ffs((int)(~m8)) - 1;               // 0..TEAM leading cells of the run
          const int src = gbase + (run & (TEAM - 1));
          const uint32_t dir_g = __shfl_sync(ALL, dir, src);
          const int s_g = __shfl_sync(ALL, s, src);
          const bool valid_g = __shfl_sync(ALL, (int)valid, src) != 0;
          if (act) {
            if (run > 0) {
              nmatch += (unsigned)run;
              if (nmatch > 62u) { EMIT(61u, 0u); nmatch -= 62u; }
              checksum += run * (int)sc.match;
              gap_open = false;
              i -= run; j -= run; r -= run;
            }
            if (run < TEAM) {
              if (!valid_g || !dir_g) act = false;               // left the segment / start of the path
              else if (dir_g == 3u) {
                if (s_g > 0) {
                  if (nmatch > 61u) { EMIT(61u, 0u); nmatch -= 61u; }
                  else ++nmatch;
                } else {
                  EMIT(nmatch, 3u);
                  nmatch = 0;
                }
                checksum += s_g;
                gap_open = false;
                --i; --j; --r;
              } else {
                if (gap_open) checksum -= sc.gap_ext;
                else { checksum -= sc.gap_init; gap_open = true; }
                if (dir_g & 1u) {
                  EMIT(nmatch, 1u);
                  nmatch = 0;
                  --i; --r; ++d;
                } else {
                  EMIT(nmatch, 2u);
                  nmatch = 0;
                  --j; --d;
                }
              }
            }
          }
        }
        if (mine) {
          EMIT(nmatch, 3u);
          EMIT(0u, 0u);
          if (ovf) flag = SMB_ERR_CAPACITY;
          else if (checksum != mscor) flag = SMB_ERRCODE_SWATSCOR;        // :767
          bi = i; bj = j; bn = n;
        }
#undef EMIT
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        PackTask &p = T[t];
        const int fl = __shfl_sync(ALL, flag, TEAM * t, LANES);
        const int i = __shfl_sync(ALL, bi, TEAM * t, LANES), j = __shfl_sync(ALL, bj, TEAM * t, LANES);
        const uint32_t n = __shfl_sync(ALL, bn, TEAM * t, LANES);
        if (on[t] && fl) { p.err = fl; on[t] = false; }
        const int prof_start = j + 1, prof_end = max_j[t], np_start = i + 1, np_end = max_i[t];
        if (on[t] && prof_start + p.minscorlen > prof_end + 1) on[t] = false;        // :1379
        if (on[t] && (int)p.nres >= max_res) { p.err = SMB_ERR_CAPACITY; on[t] = false; }
        int f2 = 0;
        uint32_t u = p.diff_used;
        if (on[t] && lane == TEAM * t) {
          // diffStrReverse (diffstr.c:850-896) into the task's DiffStr area, then the result record
          const uint8_t *rev = revbase + (size_t)t * lay.rev_n();
          uint8_t *dfinal = out.diff + diff_off[p.tix];
          int l = (int)n - 2;
          if (l >= 32767) f2 = SMB_ERRCODE_OVERFLOW;
          else if ((rev[l] >> 6) != 3u) f2 = SMB_ERRCODE_DIFFSTR;
          else {
            unsigned count_prev = rev[l] & 0x3Fu;
            bool dovf = false;
#define PUT(v) do { if (u < p.dcap) dfinal[u] = (v); else dovf = true; ++u; } while (0)
            for (--l; l >= 0; --l) {
              const unsigned count = rev[l] & 0x3Fu, typ = rev[l] >> 6;
              if (typ == 0u) {
                count_prev = (count_prev + count + 1u) & 0xffu;
                if (count_prev > 61u) { PUT(DIFFB(61u, 0u)); count_prev -= 62u; }
              } else {
                PUT(DIFFB(count_prev, typ));
                count_prev = count;
              }
            }
            PUT(DIFFB(count_prev, 3u));
            PUT(DIFFB(0u, 0u));
#undef PUT
            if (dovf) f2 = SMB_ERR_CAPACITY;
            else {
              smb_ali_result rr;
              rr.score = max_scor[t]; rr.qs = prof_start; rr.qe = prof_end; rr.rs = np_start; rr.re = np_end;
              rr.diff_off = p.diff_used; rr.diff_len = u - p.diff_used; rr.task = (uint32_t)p.tix;
              out.results[(size_t)p.tix * max_res + p.nres] = rr;
            }
          }
        }
        f2 = __shfl_sync(ALL, f2, TEAM * t, LANES);
        u = __shfl_sync(ALL, u, TEAM * t, LANES);
        if (on[t] && f2) { p.err = f2; on[t] = false; }
        if (on[t]) {
          p.diff_used = u;
          ++p.nres;
          // pre-order recursion: left part first, so push right then left (:1389, :1411)
          const bool go_left = s_left[t] + p.minscorlen < np_start;
          const bool go_right = s_right[t] > np_end + p.minscorlen;
          if (p.sp + 2 > BPK_STACK && (go_left || go_right)) p.err = SMB_ERR_CAPACITY;
          else {
            if (go_right) {
              if (lane == 0) { stk[(t * 2 + 0) * BPK_STACK + p.sp] = np_end + 1; stk[(t * 2 + 1) * BPK_STACK + p.sp] = s_right[t]; }
              ++p.sp;
            }
            if (go_left) {
              if (lane == 0) { stk[(t * 2 + 0) * BPK_STACK + p.sp] = s_left[t]; stk[(t * 2 + 1) * BPK_STACK + p.sp] = np_start - 1; }
              ++p.sp;
            }
          }
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (T[t].alive && lane == 0) {
        out.nres[T[t].tix] = T[t].nres;
        out.errs[T[t].tix] = T[t].err;
        if (out.dused) out.dused[T[t].tix] = T[t].diff_used;
      }
  }
  for (int o = LANES / 2; o > 0; o >>= 1) ncell_tot += __shfl_down_sync(ALL, ncell_tot, o, LANES);
  if (lane == 0 && ncell_tot) atomicAdd(out.cells, ncell_tot);
}

template <int LANES, int ND>
static cudaError_t launch_pack_t(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks, const int *d_order,
                                 int ntasks, int max_rows, int max_read, int *d_ticket, BandOut out, int max_res,
                                 const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count, cudaStream_t st,
                                 int *nlaunch) {
  static std::atomic<unsigned long long> smem_done{0};
  cudaError_t e = ensure_dyn_smem(band_pack_kernel<LANES, ND>, 200 * 1024, smem_done);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(d_ticket, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  PackLayout lay{(max_rows + 31) & ~31, (max_read + 15) & ~15, LANES, 8 / ND};
  if (lay.R < 32) lay.R = 32;
  if (lay.Q < 16 || lay.Q > BW_MAXREAD) lay.Q = BW_MAXREAD;
  constexpr int groups = BPK_WARPS * 32 / LANES;
  const size_t smem = lay.bytes() * groups;
  const int per_cta = groups * 2;   // tasks per CTA
  int grid = (ntasks + per_cta - 1) / per_cta;
  const int cap = sm_count * 16;
  if (grid > cap) grid = cap;
  band_pack_kernel<LANES, ND><<<grid, BPK_WARPS * 32, smem, st>>>(sc, src, d_tasks, d_order, ntasks, d_ticket, out,
                                                                  max_res, d_diff_off, d_diff_cap, lay);
  ++*nlaunch;
  return cudaGetLastError();
}

// lanes = 16: bands of at most 32 diagonals, four tasks per warp; lanes = 8: at most 24, eight tasks per warp
cudaError_t launch_band_pack(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                             const int *d_order, int ntasks, int lanes, int max_rows, int max_read, int *d_ticket,
                             BandOut out, int max_res, const uint64_t *d_diff_off, const uint32_t *d_diff_cap,
                             int sm_count, cudaStream_t st, int *nlaunch) {
  if (ntasks <= 0) return cudaSuccess;
  if (lanes == 8)
    return launch_pack_t<8, 3>(sc, src, d_tasks, d_order, ntasks, max_rows, max_read, d_ticket, out, max_res, d_diff_off,
                               d_diff_cap, sm_count, st, nlaunch);
  return launch_pack_t<16, 2>(sc, src, d_tasks, d_order, ntasks, max_rows, max_read, d_ticket, out, max_res, d_diff_off,
                              d_diff_cap, sm_count, st, nlaunch);
}

cudaError_t warm_band_pack() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, band_pack_kernel<16, 2>);
  if (e != cudaSuccess) return e;
  return cudaFuncGetAttributes(&a, band_pack_kernel<8, 3>);
}

}  // namespace smb
