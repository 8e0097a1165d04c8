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
//    accumulated 2 bits/cell in a register word per diagonal and stored to shared memory every
//    16 rows ([diagonal][row/16] layout, private to this kernel);
//  * the reference's "first strict maximum in row-major order" (alignment.c:826-830) is kept
//    exactly: every diagonal records its first strict maximum (rows ascend along a diagonal),
//    the warp then takes the maximum score with ties broken towards the smaller (row, column);
//  * backtrace, DiffStr reversal, result emission and the pre-order recursion are executed by
//    lane 0 out of shared memory, exactly as in band_dp.cu.
// Bands of at most 32 diagonals (every C1-C4 short-read task) use HALF a warp per task
// (template LANES = 16): the two halves of a warp are independent 16-lane groups with their
// own task, shared-memory slice, shuffle width and sync mask, which doubles the busy lanes at
// band width ~20.  Tasks whose window/read/band exceed the staging (long reads) stay with
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
  uint32_t dirs[2 * LANES * BWK_ROWW];
  uint8_t ref[BW_MAXROWS];
  uint8_t read[BW_MAXREAD];
  uint8_t rev[BWK_REV];
  int stk_l[BWK_STACK], stk_r[BWK_STACK];
};

#define DIFFB(count, typ) ((uint8_t)((count) + ((typ) << 6)))

template <int LANES>
__global__ void __launch_bounds__(BWK_WARPS * 32)
band_warp_kernel(const Scoring sc, const SeqSrc src, const smb_band_task *__restrict__ tasks,
                 const int *__restrict__ order, const int ntasks, int *__restrict__ ticket,
                 BandOut out, const int max_res, const uint64_t *__restrict__ diff_off,
                 const uint32_t *__restrict__ diff_cap) {
  constexpr int GROUPS = 32 / LANES;          // task groups per warp
  constexpr int MAXDIAG = 2 * LANES;
  __shared__ WarpSmem<LANES> s_w[BWK_WARPS * GROUPS];
  __shared__ unsigned long long s_S64[8];
  if (threadIdx.x < 8) {
    unsigned long long v = 0;
    for (int q = 0; q < 8; ++q) v |= (unsigned long long)(unsigned char)sc.S[threadIdx.x * 8 + q] << (q * 8);
    s_S64[threadIdx.x] = v;
  }
  __syncthreads();
  // a group is an independent LANES-wide "warp": own mask for shuffles and syncs
  const int lane = threadIdx.x & (LANES - 1);
  const unsigned FULL = (LANES == 32) ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
  WarpSmem<LANES> &sm = s_w[threadIdx.x / LANES];
  const int gi = sc.gap_init, ge = sc.gap_ext;
  unsigned long long ncell_tot = 0;

  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(ticket, 1);
    k = __shfl_sync(FULL, k, 0, LANES);
    if (k >= ntasks) break;
    const int tix = __ldg(order + k);
    const smb_band_task tk = tasks[tix];
    const bool rc = (tk.flags & SMB_TASK_READ_REVCOMP) != 0;
    const bool packed = (tk.flags & SMB_TASK_REF_PACKED) != 0;
    const int qlen = (int)tk.read_len, rlen = (int)tk.ref_len;
    __syncwarp(FULL);
    for (int x = lane; x < rlen; x += LANES) sm.ref[x] = (uint8_t)ref_base(src, packed, tk.ref_off, (uint32_t)x);
    for (int x = lane; x < qlen; x += LANES)
      sm.read[x] = (uint8_t)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)x);
    __syncwarp(FULL);

    int err = SMB_OK;
    uint32_t nres = 0, diff_used = 0;
    int minscore = tk.minscore, minscorlen = tk.minscorlen;
    uint8_t *dfinal = out.diff + diff_off[tix];
    const uint32_t dcap = diff_cap[tix];
    smb_ali_result *res = out.results + (size_t)tix * max_res;
    if (minscore < 1 || sc.match <= 0) err = SMB_ERRCODE_ASSERT;         // alignment.c:1569
    else {
      if (minscorlen * sc.match < minscore) minscorlen = minscore / sc.match;  // :1572
      if (minscorlen < 5) err = SMB_ERRCODE_ASSERT;                       // ALILEN_MIN :1574
    }
    int sp = 0;
    if (!err) {
      if (lane == 0) { sm.stk_l[0] = tk.u_left; sm.stk_r[0] = tk.u_right; }
      sp = 1;
    }
    while (sp > 0 && !err) {   // every variable tested here is warp-uniform
      --sp;
      __syncwarp(FULL);
      const int s_left = sm.stk_l[sp], s_right = sm.stk_r[sp];
      Band b;
      if (band_init(b, tk.l_edge, tk.r_edge, tk.p_left, tk.p_right, qlen, s_left, s_right, rlen))
        continue;                                                          // :1333-1338
      if (b.s_left >= b.s_len || b.band_width < 0) { err = SMB_ERRCODE_ASSERT; break; }  // :459
      const int nrows = b.s_len - b.s_left, bw = b.band_width;
      if (bw > MAXDIAG || nrows > BW_MAXROWS) { err = SMB_ERR_ARG; break; }  // host planner bug

      // ---------------- wavefront DP ----------------
      const int dA = 2 * lane, dB = dA + 1;
      const bool hasA = dA < bw, hasB = dB < bw;
      int HA = 0, HB = 0, eA = 0, eB = 0, FA = 0, FB = 0;
      int bestA = 0, bestB = 0, bestAr = 0, bestBr = 0;
      uint32_t wA = 0, wB = 0;
      unsigned ncell = 0;
      const int nlanes = (bw + 1) >> 1;
      const int iters = nrows + nlanes - 1;
      for (int it = 0; it < iters; ++it) {
        const int r = it - lane;
        const bool rowok = r >= 0 && r < nrows;
        const int Fin = __shfl_up_sync(FULL, FB, 1, LANES);    // F(r, dA-1) from the left neighbour's last step
        int jA = b.l_edge + r + dA;
        unsigned long long srow = 0;
        if (rowok) srow = s_S64[sm.ref[b.s_left + r]];
        // ---- diagonal A ----
        {
          const bool ok = rowok && hasA && jA >= b.q_left && jA < b.q_len;
          int hn = 0, e = 0, F = 0;
          uint32_t d = 0;
          if (ok) {
            const int q = sm.read[jA];
            const int h = HA + (int)(signed char)(srow >> (q << 3));
            e = eB;                                   // E(r-1, dA+1): own diagonal B, previous iteration
            F = (lane == 0) ? 0 : Fin;
            const int ep = max(e, 0), fp = max(F, 0), m = max(ep, fp);
            const bool dia = h > m;
            hn = dia ? h : m;
            e -= (e > 0) ? ge : 0;
            F -= (F > 0) ? ge : 0;
            if (dia && h > gi) {
              const int t = h - gi;
              if (h > bestA) { bestA = h; bestAr = r; }
              e = max(e, t);
              F = max(F, t);
            }
            d = dia ? 3u : (m == 0 ? 0u : (ep >= fp ? 1u : 2u));
            ++ncell;
          }
          HA = hn; eA = e; FA = F;
          if (rowok && hasA) {
            wA |= d << ((uint32_t)(r & 15) * 2u);
            if ((r & 15) == 15 || r == nrows - 1) { sm.dirs[dA * BWK_ROWW + (r >> 4)] = wA; wA = 0; }
          }
        }
        const int Ein = __shfl_down_sync(FULL, eA, 1, LANES);   // E(r-1, dB+1) from the right neighbour, this iteration
        // ---- diagonal B ----
        {
          const int jB = jA + 1;
          const bool ok = rowok && hasB && jB >= b.q_left && jB < b.q_len;
          int hn = 0, e = 0, F = 0;
          uint32_t d = 0;
          if (ok) {
            const int q = sm.read[jB];
            const int h = HB + (int)(signed char)(srow >> (q << 3));
            e = (lane == LANES - 1) ? 0 : Ein;
            F = FA;                                   // F(r, dA): just computed
            const int ep = max(e, 0), fp = max(F, 0), m = max(ep, fp);
            const bool dia = h > m;
            hn = dia ? h : m;
            e -= (e > 0) ? ge : 0;
            F -= (F > 0) ? ge : 0;
            if (dia && h > gi) {
              const int t = h - gi;
              if (h > bestB) { bestB = h; bestBr = r; }
              e = max(e, t);
              F = max(F, t);
            }
            d = dia ? 3u : (m == 0 ? 0u : (ep >= fp ? 1u : 2u));
            ++ncell;
          }
          HB = hn; eB = e; FB = F;
          if (rowok && hasB) {
            wB |= d << ((uint32_t)(r & 15) * 2u);
            if ((r & 15) == 15 || r == nrows - 1) { sm.dirs[dB * BWK_ROWW + (r >> 4)] = wB; wB = 0; }
          }
        }
      }
      ncell_tot += ncell;
      // first strict maximum in row-major order: max score, then smaller row, then smaller column
      int best = bestA, bestr = bestAr, bestd = dA;
      if (bestB > best || (bestB == best && bestB > 0 && bestBr < bestr)) { best = bestB; bestr = bestBr; bestd = dB; }
      unsigned long long key = 0;
      if (best > 0)
        key = ((unsigned long long)(unsigned)best << 32) | ((unsigned long long)(0xffffu - (unsigned)bestr) << 16) |
              (unsigned long long)(0xffffu - (unsigned)(b.l_edge + bestr + bestd - b.q_left));
      for (int o = LANES / 2; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, key, o, LANES);
        key = other > key ? other : key;
      }
      const int max_scor = (int)(key >> 32);
      const int max_r = (int)(0xffffu - (unsigned)((key >> 16) & 0xffffu));
      const int max_j = (int)(0xffffu - (unsigned)(key & 0xffffu)) + b.q_left;
      const int max_i = b.s_left + max_r;
      __syncwarp(FULL);
      if (max_scor < minscore) continue;                                   // :1364

      // ---------------- makeMetaFromTrack (alignment.c:628-781), lane 0 ----------------
      int i = max_i, j = max_j, checksum = 0, flag = 0;
      uint32_t n = 0;
      if (lane == 0) {
        bool gap_open = false, ovf = false;
        unsigned nmatch = 0;
        int r = max_r, d = max_j - b.l_edge - max_r;
#define EMIT(c, t) do { if (n < (uint32_t)BWK_REV) sm.rev[n] = DIFFB(c, t); else ovf = true; ++n; } while (0)
        while (i >= b.s_left && j >= b.q_left) {
          const uint32_t dir = (sm.dirs[d * BWK_ROWW + (r >> 4)] >> ((uint32_t)(r & 15) * 2u)) & 3u;
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
      flag = __shfl_sync(FULL, flag, 0, LANES);
      if (flag) { err = flag; break; }
      i = __shfl_sync(FULL, i, 0, LANES);
      j = __shfl_sync(FULL, j, 0, LANES);
      n = __shfl_sync(FULL, n, 0, LANES);
      const int prof_start = j + 1, prof_end = max_j, np_start = i + 1, np_end = max_i;
      if (prof_start + minscorlen > prof_end + 1) continue;                // :1379
      if (max_scor >= minscore) {                                          // :1384 addALIMETAtoRsltSet
        if ((int)nres >= max_res) { err = SMB_ERR_CAPACITY; break; }
        int f2 = 0;
        uint32_t u = diff_used;
        if (lane == 0) {
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
        f2 = __shfl_sync(FULL, f2, 0, LANES);
        if (f2) { err = f2; break; }
        diff_used = __shfl_sync(FULL, u, 0, LANES);
        ++nres;
      }
      // pre-order recursion: left part first, so push right then left (:1389, :1411)
      const bool go_left = s_left + minscorlen < np_start;
      const bool go_right = s_right > np_end + minscorlen;
      if (sp + 2 > BWK_STACK && (go_left || go_right)) { err = SMB_ERR_CAPACITY; break; }
      __syncwarp(FULL);
      if (go_right) { if (lane == 0) { sm.stk_l[sp] = np_end + 1; sm.stk_r[sp] = s_right; } ++sp; }
      if (go_left) { if (lane == 0) { sm.stk_l[sp] = s_left; sm.stk_r[sp] = np_start - 1; } ++sp; }
    }
    if (lane == 0) {
      out.nres[tix] = nres;
      out.errs[tix] = err;
      if (out.dused) out.dused[tix] = diff_used;
    }
  }
  for (int o = LANES / 2; o > 0; o >>= 1) ncell_tot += __shfl_down_sync(FULL, ncell_tot, o, LANES);
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
