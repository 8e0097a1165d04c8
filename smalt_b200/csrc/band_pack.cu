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
//    64-bit entries {validity mask of the two halves, PRMT selector, raw codes}; cell validity
//    (band, read segment, row range - different for the two tasks) is one AND of three masks;
//  * direction codes: 4 bits per row and task (two diagonals of the lane), four rows per word;
//  * argmax, backtrace, DiffStr reversal, result emission and the recursion are per task as in
//    band_warp.cu, except that the eight lanes 8t..8t+7 of the half-warp walk the path of task t together.
// Tasks with an X base in the read or the window use the general table path for their scores.
#include "common.cuh"
#include "band.h"

namespace smb {

constexpr int BPK_WARPS = 2;          // warps per CTA (4 half-warp groups, 8 tasks)
constexpr int BPK_LANES = 16;
constexpr int BPK_STACK = 24;
constexpr int BPK_ROWPAD = 16;        // rowarr index = r + BPK_ROWPAD

struct PackLayout {                   // per-group shared memory for windows of at most R rows, reads of at most Q bases
  int R;                              // multiple of 32
  int Q;                              // multiple of 16
  __host__ __device__ int rowarr_n() const { return R + 64; }
  __host__ __device__ int colarr_n() const { return R + 64; }
  __host__ __device__ int dirw() const { return R / 4; }          // words per lane
  __host__ __device__ int rev_n() const { return R + Q + 16; }
  // 32-bit entries {PRMT selector of task 0, of task 1, validity of task 0 (0x80 / 0), of task 1};
  // the raw codes (general path only) in byte arrays behind them
  __host__ __device__ size_t rowarr_off() const { return 0; }
  __host__ __device__ size_t colarr_off() const { return (size_t)rowarr_n() * 4; }
  __host__ __device__ size_t rowraw_off() const { return colarr_off() + (size_t)colarr_n() * 4; }
  __host__ __device__ size_t colraw_off() const { return rowraw_off() + (size_t)rowarr_n(); }
  __host__ __device__ size_t dirs_off() const { return (colraw_off() + (size_t)colarr_n() + 15) & ~(size_t)15; }
  __host__ __device__ size_t stk_off() const { return dirs_off() + (size_t)BPK_LANES * dirw() * 4; }
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
__device__ __forceinline__ uint32_t sel_read(uint32_t q) { return (q < 4u ? (q | ((q | 8u) << 4)) : 0xC4u) ^ 0x44u; }
// ... and of a window base; N / padding: sign-replicate a non-negative table entry -> 0
__device__ __forceinline__ uint32_t sel_ref(uint32_t r) { return r < 4u ? (r | (r << 4)) : 0x4Cu; }

#define DIFFB(count, typ) ((uint8_t)((count) + ((typ) << 6)))

// two cells (task 0 / task 1) of the restricted recurrence; ok2 = validity mask per half-word,
// ok3 = ok2 & 0x00030003.  There is no packed subtract or compare: a > b is the sign of
// b + ~a (= b - a - 1), turned into a half-word mask by PRMT sign replication; the complement
// masks it yields are folded into the LOP3s that consume them.  The running maximum of a
// diagonal is kept as the key (score << 8 | 255 - row): an unsigned packed max then keeps the
// FIRST row of the largest score (alignment.c:826-830) - scores and rows are below 256 here.
#define PACK_CELL(ok2, ok3, diag, ein, fin, s2, Hout, Eout, Fout, best, rinv2, dcode)                \
  do {                                                                                                \
    const uint32_t h_ = __vadd2((diag), (s2));                                                        \
    const uint32_t m_ = __vmaxs2((ein), (fin));               /* E, F >= 0 */                          \
    const uint32_t nm_ = ~m_;                                                                         \
    const uint32_t ndia_ = bp_neg(__vadd2(h_, nm_));          /* h <= m */                             \
    const uint32_t hn_ = __vmaxs2(h_, m_);                                                            \
    const uint32_t x_ = __vadd2(h_, ngi2);                    /* h - gap_init */                       \
    const uint32_t nx_ = bp_neg(__vadd2(h_, ngi2m1));         /* h - gap_init <= 0 */                  \
    const uint32_t open_ = ~ndia_ & ~nx_ & (ok2);                                                     \
    const uint32_t t_ = x_ & open_;                                                                   \
    const uint32_t e_ = __viaddmax_s16x2_relu((ein), nge2, t_);                                       \
    const uint32_t f_ = __viaddmax_s16x2_relu((fin), nge2, t_);                                       \
    (best) = __vmaxu2((best), ((h_ & open_) << 8) | (rinv2));                                         \
    const uint32_t nfgt_ = bp_neg(__vadd2((fin), ~(ein)));    /* F <= E: COL, else ROW */              \
    const uint32_t pos_ = bp_neg(__vadd2(nm_, 0x00010001u));  /* m > 0 */                              \
    const uint32_t b0_ = (pos_ & nfgt_) | ~ndia_, b1_ = (pos_ & ~nfgt_) | ~ndia_;                     \
    (dcode) = ((b0_ & 0x00010001u) | (b1_ & ~0x00010001u)) & (ok3);                                   \
    (Hout) = hn_ & (ok2);                                                                             \
    (Eout) = e_ & (ok2);                                                                              \
    (Fout) = f_ & (ok2);                                                                              \
  } while (0)

struct PackTask {           // group-uniform state of one of the two tasks of a group
  int tix, alive, err, sp, minscore, minscorlen, qlen, rlen;
  uint32_t nres, diff_used, dcap;
  uint64_t read_off, ref_off;
  int rc, packed;
  int l_edge0, r_edge0, p_left, p_right;
};

__global__ void __launch_bounds__(BPK_WARPS * 32)
band_pack_kernel(const Scoring sc, const SeqSrc src, const smb_band_task *__restrict__ tasks,
                 const int *__restrict__ order, const int ntasks, int *__restrict__ ticket,
                 BandOut out, const int max_res, const uint64_t *__restrict__ diff_off,
                 const uint32_t *__restrict__ diff_cap, const PackLayout lay) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr unsigned ALL = 0xffffffffu;
  constexpr int LANES = BPK_LANES;
  const int lane = threadIdx.x & (LANES - 1);
  unsigned char *base = s_raw + (size_t)(threadIdx.x / LANES) * lay.bytes();
  uint32_t *rowarr = (uint32_t *)(base + lay.rowarr_off());
  uint32_t *colarr = (uint32_t *)(base + lay.colarr_off());
  uint8_t *rowraw = base + lay.rowraw_off(), *colraw = base + lay.colraw_off();
  uint32_t *dirs = (uint32_t *)(base + lay.dirs_off());
  int *stk = (int *)(base + lay.stk_off());               // [task][l/r][BPK_STACK]
  uint8_t *revbase = base + lay.rev_off();
  const int DIRW = lay.dirw();
  const uint32_t ngi2 = (uint32_t)((-sc.gap_init) & 0xffff) * 0x10001u;
  const uint32_t nge2 = (uint32_t)((-sc.gap_ext) & 0xffff) * 0x10001u;
  const uint32_t ngi2m1 = (uint32_t)((-sc.gap_init - 1) & 0xffff) * 0x10001u;
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
          else if (b.band_width > 2 * LANES || b.s_len - b.s_left > lay.R) { p.err = SMB_ERR_ARG; on[t] = false; }
        }
      }
      const int nrows0 = on[0] ? B[0].s_len - B[0].s_left : 0, nrows1 = on[1] ? B[1].s_len - B[1].s_left : 0;
      const int bw0 = on[0] ? B[0].band_width : 0, bw1 = on[1] ? B[1].band_width : 0;
      int iters = max(on[0] ? nrows0 + ((bw0 + 1) >> 1) - 1 : 0, on[1] ? nrows1 + ((bw1 + 1) >> 1) - 1 : 0);
      iters = max(iters, __shfl_xor_sync(ALL, iters, 16));
      // ---- stage rows and columns of both tasks: {mask32, selector16, raw codes} ----
      bool hasx = false;
      // the trip count is the maximum over both half-warps: stage (as invalid) everything it can touch
      const int nrow_e = iters + 2 * BPK_ROWPAD, ncol_e = min(iters + LANES + 2, lay.colarr_n());
      __syncwarp();
      for (int e = lane; e < min(nrow_e, lay.rowarr_n()); e += LANES) {
        const int r = e - BPK_ROWPAD;
        uint32_t mask = 0, sel = 0, raw = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int nr = t ? nrows1 : nrows0;
          uint32_t c = 7u;
          if (on[t] && r >= 0 && r < nr) {
            c = ref_base(src, T[t].packed != 0, T[t].ref_off, (uint32_t)(B[t].s_left + r));
            mask |= 0x800000u << (8 * t);
          }
          hasx |= c == 4u;
          sel |= sel_ref(c) << (8 * t);
          raw |= c << (4 * t);
        }
        rowarr[e] = mask | sel;
        rowraw[e] = (uint8_t)raw;
      }
      for (int x = lane; x < ncol_e; x += LANES) {
        uint32_t mask = 0, sel = 0, raw = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int j = B[t].l_edge + x;
          uint32_t c = 7u;
          if (on[t] && j >= B[t].q_left && j < B[t].q_len) {
            c = read_base(src.arena, T[t].read_off, (uint32_t)T[t].qlen, T[t].rc != 0, (uint32_t)j);
            mask |= 0x800000u << (8 * t);
          }
          hasx |= c == 4u;
          sel |= sel_read(c) << (8 * t);
          raw |= c << (4 * t);
        }
        colarr[x] = mask | sel;
        colraw[x] = (uint8_t)raw;
      }
      const bool general = __any_sync(ALL, hasx);
      __syncwarp();

      // ---------------- packed wavefront DP ----------------
      const int dA = 2 * lane, dB = dA + 1;
      const uint32_t hasA2 = (dA < bw0 ? 0xffffu : 0u) | (dA < bw1 ? 0xffff0000u : 0u);
      const uint32_t hasB2 = (dB < bw0 ? 0xffffu : 0u) | (dB < bw1 ? 0xffff0000u : 0u);
      uint32_t HA = 0, HB = 0, eA = 0, eB = 0, FA = 0, FB = 0;
      uint32_t bestA = 0, bestB = 0, wdir = 0, cnt2 = 0;
      uint32_t *const dirp = dirs + lane * DIRW;
      const int maxrows = max(nrows0, nrows1);
      // validity bytes -> half-word masks (sign replication of bytes 2 and 3)
#define BP_VMASK(e) bp_prmt((e), 0u, 0xbbaau)
      uint32_t colA = colarr[lane], cmaskA = BP_VMASK(colA);
      for (int it = 0; it < iters; ++it) {
        const int r = it - lane;
        const uint32_t Fin = __shfl_up_sync(ALL, FB, 1, LANES);
        const uint32_t rw = rowarr[r + BPK_ROWPAD];
        const uint32_t colB = colarr[it + lane + 1];
        const uint32_t rmask = BP_VMASK(rw), cmaskB = BP_VMASK(colB);
        const uint32_t okA = rmask & cmaskA & hasA2, okB = rmask & cmaskB & hasB2;
        const uint32_t okA3 = okA & 0x00030003u, okB3 = okB & 0x00030003u;
        uint32_t sA, sB;
        if (!general) {   // (PRMT reads the low 16 bits of the selector only)
          sA = bp_prmt(0u, T0, colA ^ rw);
          sB = bp_prmt(0u, T0, colB ^ rw);
        } else {   // X bases: per-cell table look-ups
          const uint32_t rr = rowraw[r + BPK_ROWPAD], ca = colraw[it + lane], cb = colraw[it + lane + 1];
          const uint32_t r0 = rr & 7u, r1 = (rr >> 4) & 7u;
          const int a0 = sc.S[r0 * 8u + (ca & 7u)], a1 = sc.S[r1 * 8u + ((ca >> 4) & 7u)];
          const int b0 = sc.S[r0 * 8u + (cb & 7u)], b1 = sc.S[r1 * 8u + ((cb >> 4) & 7u)];
          sA = ((uint32_t)a0 & 0xffffu) | ((uint32_t)a1 << 16);
          sB = ((uint32_t)b0 & 0xffffu) | ((uint32_t)b1 << 16);
        }
        const uint32_t rinv2 = ((unsigned)r < 256u) ? (uint32_t)(255 - r) * 0x10001u : 0u;
        uint32_t dcA, dcB;
        PACK_CELL(okA, okA3, HA, eB, (lane == 0 ? 0u : Fin), sA, HA, eA, FA, bestA, rinv2, dcA);
        const uint32_t Ein = __shfl_down_sync(ALL, eA, 1, LANES);
        PACK_CELL(okB, okB3, HB, (lane == LANES - 1 ? 0u : Ein), FA, sB, HB, eB, FB, bestB, rinv2, dcB);
        cnt2 += (okA & 0x00010001u) + (okB & 0x00010001u);
        colA = colB;
        cmaskA = cmaskB;
        if (r >= 0 && r < maxrows) {   // (the trip count may exceed this group's rows: never store beyond them)
          wdir |= (dcA | (dcB << 2)) << ((uint32_t)(r & 3) << 2);
          if ((r & 3) == 3) { dirp[r >> 2] = wdir; wdir = 0; }
        }
      }
      {
        const int rlast = min(iters - 1 - lane, maxrows - 1);
        if (rlast >= 0 && (rlast & 3) != 3) dirp[rlast >> 2] = wdir;
      }
      ncell_tot += (cnt2 & 0xffffu) + (cnt2 >> 16);
      __syncwarp();

      // ---- per task: argmax, backtrace, result, recursion (lane t serves task t) ----
      int max_scor[2], max_i[2], max_j[2], max_r[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const Band &b = B[t];
        const int sh = 16 * t;
        const uint32_t kA = (bestA >> sh) & 0xffffu, kB = (bestB >> sh) & 0xffffu;
        const int bA = (int)(kA >> 8), bB = (int)(kB >> 8);
        const int rA = 255 - (int)(kA & 0xffu), rB = 255 - (int)(kB & 0xffu);
        int best = bA, bestr = rA, bestd = dA;
        if (bB > best || (bB == best && bB > 0 && rB < bestr)) { best = bB; bestr = rB; bestd = dB; }
        unsigned long long key = 0;
        if (on[t] && best > 0)
          key = ((unsigned long long)(unsigned)best << 32) | ((unsigned long long)(0xffffu - (unsigned)bestr) << 16) |
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
      // makeMetaFromTrack (alignment.c:628-781): lanes 8t..8t+7 of the group walk the path of task t
      // together.  Sub-lane L looks at the cell L steps up the current diagonal; the run of cells
      // that continue as "DIA move onto a match" (ballot) is consumed at once - the reference's
      // per-match counter (`nmatch > 61 ? emit 61 : ++nmatch`, i.e. 62 rolls over to 1 with one
      // emission) has the closed form below - and the first cell that is anything else takes the
      // reference's step, with the values of the sub-lane that looked at it.
      int bi = 0, bj = 0, flag = 0;
      uint32_t bn = 0;
      {
        const int t = (lane >> 3) & 1, L = lane & 7;
        const int gbase = (int)(threadIdx.x & 24u);          // first lane of this 8-lane team within the warp
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
          const int rr = r - L;
          const bool valid = act && i - L >= b.s_left && j - L >= b.q_left;
          uint32_t w = 0, csel = 0, rsel = 0;
          if (valid) {
            w = dirs[(d >> 1) * DIRW + (rr >> 2)];
            csel = colarr[j - L - b.l_edge];
            rsel = rowarr[rr + BPK_ROWPAD];
          }
          const uint32_t dir = (w >> (sh + ((uint32_t)(rr & 3) << 2) + ((uint32_t)(d & 1) << 1))) & 3u;
          int s = (int)(short)(bp_prmt(0u, T0, csel ^ rsel) >> sh);
          if (general && valid)
            s = (int)sc.S[((rowraw[rr + BPK_ROWPAD] >> (4 * t)) & 7u) * 8u + ((colraw[j - L - b.l_edge] >> (4 * t)) & 7u)];
          const bool fast = valid && dir == 3u && s > 0 && !general;
          const uint32_t m8 = (__ballot_sync(ALL, fast) >> gbase) & 0xffu;
          const int run = __ffs((int)(~m8)) - 1;               // 0..8 leading cells of the run
          const int src = gbase + (run & 7);
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
            if (run < 8) {
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
        const int fl = __shfl_sync(ALL, flag, 8 * t, LANES);
        const int i = __shfl_sync(ALL, bi, 8 * t, LANES), j = __shfl_sync(ALL, bj, 8 * t, LANES);
        const uint32_t n = __shfl_sync(ALL, bn, 8 * t, LANES);
        if (on[t] && fl) { p.err = fl; on[t] = false; }
        const int prof_start = j + 1, prof_end = max_j[t], np_start = i + 1, np_end = max_i[t];
        if (on[t] && prof_start + p.minscorlen > prof_end + 1) on[t] = false;        // :1379
        if (on[t] && (int)p.nres >= max_res) { p.err = SMB_ERR_CAPACITY; on[t] = false; }
        int f2 = 0;
        uint32_t u = p.diff_used;
        if (on[t] && lane == 8 * t) {
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
        f2 = __shfl_sync(ALL, f2, 8 * t, LANES);
        u = __shfl_sync(ALL, u, 8 * t, LANES);
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

cudaError_t launch_band_pack(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                             const int *d_order, int ntasks, int max_rows, int max_read, int *d_ticket, BandOut out,
                             int max_res, const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch) {
  if (ntasks <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(band_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  PackLayout lay{(max_rows + 31) & ~31, (max_read + 15) & ~15};
  if (lay.R < 32) lay.R = 32;
  if (lay.Q < 16 || lay.Q > BW_MAXREAD) lay.Q = BW_MAXREAD;
  const size_t smem = lay.bytes() * BPK_WARPS * 2;
  const int per_cta = BPK_WARPS * 4;   // tasks per CTA
  int grid = (ntasks + per_cta - 1) / per_cta;
  const int cap = sm_count * 16;
  if (grid > cap) grid = cap;
  band_pack_kernel<<<grid, BPK_WARPS * 32, smem, st>>>(sc, src, d_tasks, d_order, ntasks, d_ticket, out, max_res,
                                                       d_diff_off, d_diff_cap, lay);
  ++*nlaunch;
  return cudaGetLastError();
}

cudaError_t warm_band_pack() {
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, band_pack_kernel);
}

}  // namespace smb
