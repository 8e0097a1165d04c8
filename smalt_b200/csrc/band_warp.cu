// band_warp.cu - K3 for short-read bands: ONE WARP PER TASK, anti-diagonal wavefront.
//
// Same function as band_kernel<true> of band_dp.cu (aliSmiWatInBand, /root/reference/src/
// alignment.c:1548-1601: alignSmiWatBandRecursive :1300-1434 -> alignSmiWatBand :788-1027 +
// makeMetaFromTrack :628-781 + diffStrReverse diffstr.c:850-896), different mapping.  A batch of
// a few thousand short reads yields ~20 k tasks of ~3 k cells: one THREAD per task leaves the
// GPU at 5 % occupancy (profiles/r1_ncu_full_k3_details.csv), so here a warp works on one task:
//
//  * band coordinates: cell (r, d) = window row r (from the first row of the pass), diagonal
//    d = j - (l_edge + r).  Its inputs are H(r-1, d) [same diagonal], E(r-1, d+1) [vertical
//    gap] and F(r, d-1) [horizontal gap] - so cells with equal 2r + d are independent;
//  * lane l owns the diagonals 2l and 2l+1 and, in iteration `it`, computes row r = it - l of
//    both, one after the other: every lane does useful work in every iteration (no parity
//    bubbles) and the exchange is two warp shuffles per iteration (F to the right neighbour, E
//    to the left one).  H, E, F live in registers; nothing but the direction bits leaves them;
//  * window and read bytes of the task are staged once in shared memory; direction codes are
//    accumulated 2 bits/cell in a register word per lane and stored to shared memory every
//    8 rows ([lane][row/8] layout, private to this kernel);
//  * the reference's "first strict maximum in row-major order" (alignment.c:826-830) is kept
//    exactly: every diagonal records its first strict maximum (rows ascend along a diagonal),
//    the warp then takes the maximum score with ties broken towards the smaller (row, column);
//  * backtrace, DiffStr reversal, result emission and the pre-order recursion are executed by
//    lane 0 out of shared memory, exactly as in band_dp.cu.
// Bands of at most 32 diagonals (every C1-C4 short-read task) use HALF a warp per task
// (template LANES = 16): the two halves of a warp hold one task each and run in lockstep
// (full-warp shuffles of width 16, trip counts = maximum over the halves, a half with nothing
// to do in a phase predicated off), which doubles the busy lanes at band width ~20.  The cell
// update is branch free; the direction codes of a lane's two diagonals share one word
// (4 bits per row).  Tasks whose window/read/band exceed the staging (long reads) stay with
// band_kernel<true>.
#include "common.cuh"
#include "band.h"

namespace smb {

constexpr int BWK_WARPS = 4;
constexpr int BWK_ROWW = BW_MAXROWS / 16;
constexpr int BWK_STACK = 48;
constexpr int BWK_REV = BW_MAXROWS + BW_MAXREAD + 16;

template <int LANES>
struct WarpSmem {   // one per task group (LANES lanes, 2*LANES diagonals)
  uint32_t dirs[LANES * (BW_MAXROWS / 8)];   // [lane][row/8]: 4 bits per row (2 per diagonal of the lane)
  uint8_t ref[BW_MAXROWS];
  uint8_t read[BW_MAXREAD];
  uint8_t rev[BWK_REV];
  int stk_l[BWK_STACK], stk_r[BWK_STACK];
};

#include "band_cell.cuh"

template <int LANES>
__global__ void __launch_bounds__(BWK_WARPS * 32)
band_warp_kernel(const Scoring sc, const SeqSrc src, const smb_band_task *__restrict__ tasks,
                 const int *__restrict__ order, const int ntasks, int *__restrict__ ticket,
                 BandOut out, const int max_res, const uint64_t *__restrict__ diff_off,
                 const uint32_t *__restrict__ diff_cap) {
  constexpr int GROUPS = 32 / LANES;          // task groups per warp
  constexpr int MAXDIAG = 2 * LANES;
  constexpr int DIRW = BW_MAXROWS / 8;
  constexpr unsigned ALL = 0xffffffffu;
  __shared__ WarpSmem<LANES> s_w[BWK_WARPS * GROUPS];
  __shared__ unsigned long long s_S64[8];
  if (threadIdx.x < 8) {
    unsigned long long v = 0;
    for (int q = 0; q < 8; ++q) v |= (unsigned long long)(unsigned char)sc.S[threadIdx.x * 8 + q] << (q * 8);
    s_S64[threadIdx.x] = v;
  }
  __syncthreads();
  // The groups of a warp (two half-warps for LANES = 16) work in LOCKSTEP on one task each:
  // every shuffle and vote is a full-warp operation of width LANES, loop trip counts are the
  // maximum over the groups and a group that has nothing to do in a phase is predicated off.
  const int lane = threadIdx.x & (LANES - 1);
  WarpSmem<LANES> &sm = s_w[threadIdx.x / LANES];
  const int gi = sc.gap_init, ge = sc.gap_ext;
  unsigned long long ncell_tot = 0;

  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(ticket, 1);
    k = __shfl_sync(ALL, k, 0, LANES);
    const bool alive = k < ntasks;
    if (!__any_sync(ALL, alive)) break;
    const int tix = alive ? __ldg(order + k) : 0;
    smb_band_task tk = tasks[tix];
    const bool rc = (tk.flags & SMB_TASK_READ_REVCOMP) != 0;
    const bool packed = (tk.flags & SMB_TASK_REF_PACKED) != 0;
    const int qlen = (int)tk.read_len, rlen = (int)tk.ref_len;
    __syncwarp();
    if (alive) {
      for (int x = lane; x < rlen; x += LANES) sm.ref[x] = (uint8_t)ref_base(src, packed, tk.ref_off, (uint32_t)x);
      for (int x = lane; x < qlen; x += LANES)
        sm.read[x] = (uint8_t)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)x);
    }
    int err = SMB_OK;
    uint32_t nres = 0, diff_used = 0;
    int minscore = tk.minscore, minscorlen = tk.minscorlen;
    uint8_t *dfinal = out.diff + (alive ? diff_off[tix] : 0);
    const uint32_t dcap = alive ? diff_cap[tix] : 0u;
    smb_ali_result *res = out.results + (size_t)tix * max_res;
    if (minscore < 1 || sc.match <= 0) err = SMB_ERRCODE_ASSERT;         // alignment.c:1569
    else {
      if (minscorlen * sc.match < minscore) minscorlen = minscore / sc.match;  // :1572
      if (minscorlen < 5) err = SMB_ERRCODE_ASSERT;                       // ALILEN_MIN :1574
    }
    int sp = 0;
    if (alive && !err) {
      if (lane == 0) { sm.stk_l[0] = tk.u_left; sm.stk_r[0] = tk.u_right; }
      sp = 1;
    }
    __syncwarp();

    // one round = one DP pass of every group that still has a row range on its stack
    while (__any_sync(ALL, alive && sp > 0 && !err)) {
      bool on = alive && sp > 0 && !err;          // group-uniform
      int s_left = 0, s_right = 0;
      Band b;
      b.band_width = 0; b.l_edge = 0; b.r_edge = 0; b.l_edge_orig = 0; b.r_edge_orig = 0;
      b.s_left = 0; b.s_len = 0; b.q_left = 0; b.q_len = 0;
      if (on) {
        --sp;
        s_left = sm.stk_l[sp];
        s_right = sm.stk_r[sp];
        if (band_init(b, tk.l_edge, tk.r_edge, tk.p_left, tk.p_right, qlen, s_left, s_right, rlen))
          on = false;                                                      // :1333-1338
        else if (b.s_left >= b.s_len || b.band_width < 0) { err = SMB_ERRCODE_ASSERT; on = false; }  // :459
        else if (b.band_width > MAXDIAG || b.s_len - b.s_left > BW_MAXROWS) { err = SMB_ERR_ARG; on = false; }
      }
      const int nrows = on ? b.s_len - b.s_left : 0, bw = on ? b.band_width : 0;

      // ---------------- wavefront DP ----------------
      const int dA = 2 * lane, dB = dA + 1;
      const bool hasA = dA < bw, hasB = dB < bw;
      int HA = 0, HB = 0, eA = 0, eB = 0, FA = 0, FB = 0;
      int bestA = 0, bestB = 0, bestAr = 0, bestBr = 0;
      uint32_t wdir = 0;
      unsigned ncell = 0;
      int iters = on ? nrows + ((bw + 1) >> 1) - 1 : 0;
      if (GROUPS > 1) iters = max(iters, __shfl_xor_sync(ALL, iters, 16));
      // column of diagonal A in row r = it - lane: j = l_edge + it + lane (advances with it)
      const int jbase = b.l_edge + lane;
      int qnext = 0;   // read base of column jA of the next iteration (= column jB of this one)
      {
        const int j0 = jbase;
        qnext = (on && j0 >= 0 && j0 < qlen) ? (int)sm.read[j0] : 0;
      }
      uint32_t *const dirp = sm.dirs + lane * DIRW;
      for (int it = 0; it < iters; ++it) {
        const int r = it - lane;
        const bool rowok = on && r >= 0 && r < nrows;
        const int Fin = __shfl_up_sync(ALL, FB, 1, LANES);     // F(r, dA-1): left neighbour's last step
        const int jA = jbase + it, jB = jA + 1;
        const int refc = rowok ? (int)sm.ref[b.s_left + r] : 0;
        const unsigned long long srow = s_S64[refc];
        const int qA = qnext;
        const int qB = (on && jB >= 0 && jB < qlen) ? (int)sm.read[jB] : 0;
        qnext = qB;
        const bool okA = rowok && hasA && jA >= b.q_left && jA < b.q_len;
        const bool okB = rowok && hasB && jB >= b.q_left && jB < b.q_len;
        const int sA = (int)(signed char)(srow >> (qA << 3));
        const int sB = (int)(signed char)(srow >> (qB << 3));
        uint32_t dcA, dcB;
        // diagonal A: E(r-1, dA+1) is this lane's diagonal B of the previous iteration
        BAND_CELL(okA, HA, eB, (lane == 0 ? 0 : Fin), sA, HA, eA, FA, bestA, bestAr, r, dcA);
        const int Ein = __shfl_down_sync(ALL, eA, 1, LANES);   // E(r-1, dB+1): right neighbour, this iteration
        // diagonal B: F(r, dA) was just computed
        BAND_CELL(okB, HB, (lane == LANES - 1 ? 0 : Ein), FA, sB, HB, eB, FB, bestB, bestBr, r, dcB);
        ncell += (unsigned)okA + (unsigned)okB;
        if (rowok) {
          wdir |= (dcA | (dcB << 2)) << ((uint32_t)(r & 7) << 2);
          if ((r & 7) == 7 || r == nrows - 1) { dirp[r >> 3] = wdir; wdir = 0; }
        }
      }
      ncell_tot += ncell;
      // first strict maximum in row-major order: max score, then smaller row, then smaller column
      int best = bestA, bestr = bestAr, bestd = dA;
      if (bestB > best || (bestB == best && bestB > 0 && bestBr < bestr)) { best = bestB; bestr = bestBr; bestd = dB; }
      unsigned long long key = 0;
      if (on && best > 0)
        key = ((unsigned long long)(unsigned)best << 32) | ((unsigned long long)(0xffffu - (unsigned)bestr) << 16) |
              (unsigned long long)(0xffffu - (unsigned)(b.l_edge + bestr + bestd - b.q_left));
      for (int o = LANES / 2; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(ALL, key, o, LANES);
        key = other > key ? other : key;
      }
      const int max_scor = (int)(key >> 32);
      const int max_r = (int)(0xffffu - (unsigned)((key >> 16) & 0xffffu));
      const int max_j = (int)(0xffffu - (unsigned)(key & 0xffffu)) + b.q_left;
      const int max_i = b.s_left + max_r;
      __syncwarp();
      if (max_scor < minscore) on = false;                                  // :1364

      // ---------------- makeMetaFromTrack (alignment.c:628-781), lane 0 of the group ----------------
      int i = max_i, j = max_j, flag = 0;
      uint32_t n = 0;
      if (on && lane == 0) {
        bool gap_open = false, ovf = false;
        unsigned nmatch = 0;
        int checksum = 0;
        int r = max_r, d = max_j - b.l_edge - max_r;
#define EMIT(c, t) do { if (n < (uint32_t)BWK_REV) sm.rev[n] = DIFFB(c, t); else ovf = true; ++n; } while (0)
        while (i >= b.s_left && j >= b.q_left) {
          const uint32_t dir = (sm.dirs[(d >> 1) * DIRW + (r >> 3)] >> (((uint32_t)(r & 7) << 2) + ((uint32_t)(d & 1) << 1))) & 3u;
          if (!dir) break;
          if (dir == 3u) {
            const int s = (int)(signed char)(s_S64[sm.ref[i]] >> ((int)sm.read[j] << 3));
            if (s > 0) {
              if (nmatch > 61u) { EMIT(61u, 0u); nmatch -= 61u; }
              else ++nmatch;
            } else {
              EMIT(nmatch, 3u);
              nmatch = 0;
            }
            checksum += s;
            gap_open = false;
            --i; --j; --r;
            continue;
          }
          if (gap_open) checksum -= sc.gap_ext;
          else { checksum -= sc.gap_init; gap_open = true; }
          if (dir & 1u) {
            EMIT(nmatch, 1u);
            nmatch = 0;
            --i; --r; ++d;
            continue;
          }
          EMIT(nmatch, 2u);
          nmatch = 0;
          --j; --d;
        }
        EMIT(nmatch, 3u);
        EMIT(0u, 0u);
#undef EMIT
        if (ovf) flag = SMB_ERR_CAPACITY;
        else if (checksum != max_scor) flag = SMB_ERRCODE_SWATSCOR;        // :767
      }
      flag = __shfl_sync(ALL, flag, 0, LANES);
      i = __shfl_sync(ALL, i, 0, LANES);
      j = __shfl_sync(ALL, j, 0, LANES);
      n = __shfl_sync(ALL, n, 0, LANES);
      if (on && flag) { err = flag; on = false; }
      const int prof_start = j + 1, prof_end = max_j, np_start = i + 1, np_end = max_i;
      if (on && prof_start + minscorlen > prof_end + 1) on = false;        // :1379
      // :1384 addALIMETAtoRsltSet (max_scor >= minscore holds here)
      if (on && (int)nres >= max_res) { err = SMB_ERR_CAPACITY; on = false; }
      int f2 = 0;
      uint32_t u = diff_used;
      if (on && lane == 0) {
        // diffStrReverse (diffstr.c:850-896)
        int l = (int)n - 2;
        if (l >= 32767) f2 = SMB_ERRCODE_OVERFLOW;
        else if ((sm.rev[l] >> 6) != 3u) f2 = SMB_ERRCODE_DIFFSTR;
        else {
          unsigned count_prev = sm.rev[l] & 0x3Fu;
          bool dovf = false;
#define PUT(v) do { if (u < dcap) dfinal[u] = (v); else dovf = true; ++u; } while (0)
          for (--l; l >= 0; --l) {
            const unsigned count = sm.rev[l] & 0x3Fu, typ = sm.rev[l] >> 6;
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
            rr.score = max_scor; rr.qs = prof_start; rr.qe = prof_end; rr.rs = np_start; rr.re = np_end;
            rr.diff_off = diff_used; rr.diff_len = u - diff_used; rr.task = (uint32_t)tix;
            res[nres] = rr;
          }
        }
      }
      f2 = __shfl_sync(ALL, f2, 0, LANES);
      u = __shfl_sync(ALL, u, 0, LANES);
      if (on && f2) { err = f2; on = false; }
      if (on) {
        diff_used = u;
        ++nres;
        // pre-order recursion: left part first, so push right then left (:1389, :1411)
        const bool go_left = s_left + minscorlen < np_start;
        const bool go_right = s_right > np_end + minscorlen;
        if (sp + 2 > BWK_STACK && (go_left || go_right)) err = SMB_ERR_CAPACITY;
        else {
          if (go_right) { if (lane == 0) { sm.stk_l[sp] = np_end + 1; sm.stk_r[sp] = s_right; } ++sp; }
          if (go_left) { if (lane == 0) { sm.stk_l[sp] = s_left; sm.stk_r[sp] = np_start - 1; } ++sp; }
        }
      }
      __syncwarp();
    }
    if (alive && lane == 0) {
      out.nres[tix] = nres;
      out.errs[tix] = err;
      if (out.dused) out.dused[tix] = diff_used;
    }
  }
  for (int o = LANES / 2; o > 0; o >>= 1) ncell_tot += __shfl_down_sync(ALL, ncell_tot, o, LANES);
  if (lane == 0 && ncell_tot) atomicAdd(out.cells, ncell_tot);
}

cudaError_t launch_band_warp(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                             const int *d_order, int ntasks, int lanes, int *d_ticket, BandOut out, int max_res,
                             const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch) {
  if (ntasks <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const int per_cta = BWK_WARPS * (32 / lanes);
  int grid = (ntasks + per_cta - 1) / per_cta;
  const int cap = sm_count * 8;
  if (grid > cap) grid = cap;
  if (lanes == 16)
    band_warp_kernel<16><<<grid, BWK_WARPS * 32, 0, st>>>(sc, src, d_tasks, d_order, ntasks, d_ticket, out, max_res,
                                                          d_diff_off, d_diff_cap);
  else
    band_warp_kernel<32><<<grid, BWK_WARPS * 32, 0, st>>>(sc, src, d_tasks, d_order, ntasks, d_ticket, out, max_res,
                                                          d_diff_off, d_diff_cap);
  ++*nlaunch;
  return cudaGetLastError();
}

cudaError_t warm_band_warp() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, band_warp_kernel<16>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, band_warp_kernel<32>);
  return e;
}

}  // namespace smb
