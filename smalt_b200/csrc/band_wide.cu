// band_wide.cu - K3 for bands and windows beyond the staging of band_warp.cu / band_pack.cu:
// one warp per task, FOUR diagonals per lane (bands of up to 128 diagonals), windows of up to 512
// rows, reads of up to 512 bases.
//
// Same function and same wavefront as band_warp_kernel<32> (see the header of band_warp.cu;
// aliSmiWatInBand, /root/reference/src/alignment.c:1548-1601, :1300-1434, :788-1027, :628-781):
// lane l owns the diagonals 4l .. 4l+3 and computes row r = it - l of all four in iteration
// `it`, one after the other.  Inputs of cell (r, d): H(r-1, d) own register; E(r-1, d+1) = the
// lane's next diagonal as left by the previous iteration, for the lane's last diagonal the right
// neighbour's first diagonal of THIS iteration (it is one row behind); F(r, d-1) = the diagonal
// just computed, for the lane's first diagonal the left neighbour's last one of the previous
// iteration (it is one row ahead).  Two shuffles per iteration, four cells per lane.
//
// Who lands here: the restricted searches of paired-end mapping in the on-the-fly k=5 index
// (rmap.c:2032-2046) and against weak first hits, whose bands are widened to
// (perfect score - best score) / gap extension (rmap.c:888-896) - ~100 diagonals for a mate that
// does not really map; single-end reads of 230-480 bases (windows of more than 256 rows).  With
// one THREAD per task (band_kernel<true>) a block's ~100 such tasks took 8-10 ms of latency.
#include "common.cuh"
#include "band.h"
#include "band_cell.cuh"

namespace smb {

constexpr int BWD_WARPS = 2;
constexpr int BWD_DPL = 4;                         // diagonals per lane
constexpr int BWD_DIRW = BWD_MAXROWS / 4;          // direction words per lane: 4 rows x (4 x 2 bits) per word
constexpr int BWD_STACK = 48;
constexpr int BWD_REV = BWD_MAXROWS + BWD_MAXREAD + 16;

struct WideSmem {
  uint32_t dirs[32 * BWD_DIRW];   // [lane][row/4]
  uint8_t ref[BWD_MAXROWS];
  uint8_t read[BWD_MAXREAD];
  uint8_t rev[BWD_REV];
  int stk_l[BWD_STACK], stk_r[BWD_STACK];
};

__global__ void __launch_bounds__(BWD_WARPS * 32)
band_wide_kernel(const Scoring sc, const SeqSrc src, const smb_band_task *__restrict__ tasks,
                 const int *__restrict__ order, const int ntasks, int *__restrict__ ticket,
                 BandOut out, const int max_res, const uint64_t *__restrict__ diff_off,
                 const uint32_t *__restrict__ diff_cap) {
  constexpr unsigned ALL = 0xffffffffu;
  constexpr int D = BWD_DPL;
  __shared__ WideSmem s_w[BWD_WARPS];
  __shared__ unsigned long long s_S64[8];
  if (threadIdx.x < 8) {
    unsigned long long v = 0;
    for (int q = 0; q < 8; ++q) v |= (unsigned long long)(unsigned char)sc.S[threadIdx.x * 8 + q] << (q * 8);
    s_S64[threadIdx.x] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  WideSmem &sm = s_w[threadIdx.x >> 5];
  const int gi = sc.gap_init, ge = sc.gap_ext;
  unsigned long long ncell_tot = 0;

  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(ticket, 1);
    k = __shfl_sync(ALL, k, 0);
    if (k >= ntasks) break;
    const int tix = __ldg(order + k);
    const smb_band_task tk = tasks[tix];
    const bool rc = (tk.flags & SMB_TASK_READ_REVCOMP) != 0;
    const bool packed = (tk.flags & SMB_TASK_REF_PACKED) != 0;
    const int qlen = (int)tk.read_len, rlen = (int)tk.ref_len;
    __syncwarp();
    for (int x = lane; x < rlen; x += 32) sm.ref[x] = (uint8_t)ref_base(src, packed, tk.ref_off, (uint32_t)x);
    for (int x = lane; x < qlen; x += 32)
      sm.read[x] = (uint8_t)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)x);
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
    __syncwarp();

    while (sp > 0 && !err) {   // one DP pass per row range of the recursion (warp-uniform)
      bool on = true;
      Band b;
      --sp;
      const int s_left = sm.stk_l[sp], s_right = sm.stk_r[sp];
      if (band_init(b, tk.l_edge, tk.r_edge, tk.p_left, tk.p_right, qlen, s_left, s_right, rlen)) on = false;   // :1333-1338
      else if (b.s_left >= b.s_len || b.band_width < 0) { err = SMB_ERRCODE_ASSERT; on = false; }  // :459
      else if (b.band_width > BWD_MAXDIAG || b.s_len - b.s_left > BWD_MAXROWS) { err = SMB_ERR_ARG; on = false; }
      if (!on) { __syncwarp(); continue; }
      const int nrows = b.s_len - b.s_left, bw = b.band_width;

      // ---------------- wavefront DP ----------------
      const int d0 = D * lane;
      int H[D], e[D], F[D], best[D], bestr[D], q[D];
#pragma unroll
      for (int c = 0; c < D; ++c) { H[c] = e[c] = F[c] = best[c] = bestr[c] = 0; }
      uint32_t wdir = 0;
      unsigned ncell = 0;
      const int iters = nrows + ((bw + D - 1) / D) - 1;
      // column of diagonal d0 + c in row r = it - lane: j = l_edge + it + (D-1)*lane + c
      const int jbase = b.l_edge + (D - 1) * lane;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const int j = jbase + c;
        q[c] = (j >= 0 && j < qlen) ? (int)sm.read[j] : 0;
      }
      uint32_t *const dirp = sm.dirs + lane * BWD_DIRW;
      for (int it = 0; it < iters; ++it) {
        const int r = it - lane;
        const bool rowok = r >= 0 && r < nrows;
        const int Fin = __shfl_up_sync(ALL, F[D - 1], 1);     // F(r, d0-1): left neighbour, previous iteration
        const int j0 = jbase + it;
        const int refc = rowok ? (int)sm.ref[b.s_left + r] : 0;
        const unsigned long long srow = s_S64[refc];
        uint32_t dc[D];
        int Ein = 0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const int j = j0 + c;
          const bool ok = rowok && (d0 + c) < bw && j >= b.q_left && j < b.q_len;
          const int s = (int)(signed char)(srow >> (q[c] << 3));
          const int ein = (c == D - 1) ? (lane == 31 ? 0 : Ein) : e[c + 1];   // E(r-1, d+1)
          const int fin = (c == 0) ? (lane == 0 ? 0 : Fin) : F[c - 1];         // F(r, d-1)
          BAND_CELL(ok, H[c], ein, fin, s, H[c], e[c], F[c], best[c], bestr[c], r, dc[c]);
          if (c == 0) Ein = __shfl_down_sync(ALL, e[0], 1);   // right neighbour's first diagonal, this iteration
          ncell += (unsigned)ok;
        }
        if (rowok) {
          wdir |= (dc[0] | (dc[1] << 2) | (dc[2] << 4) | (dc[3] << 6)) << ((uint32_t)(r & 3) << 3);
          if ((r & 3) == 3 || r == nrows - 1) { dirp[r >> 2] = wdir; wdir = 0; }
        }
        // read bases of the next iteration: every column moves one to the right
#pragma unroll
        for (int c = 0; c < D - 1; ++c) q[c] = q[c + 1];
        {
          const int jn = j0 + D;
          q[D - 1] = (jn >= 0 && jn < qlen) ? (int)sm.read[jn] : 0;
        }
      }
      ncell_tot += ncell;
      // first strict maximum in row-major order: max score, then smaller row, then smaller column
      int bst = best[0], bstr = bestr[0], bstd = d0;
#pragma unroll
      for (int c = 1; c < D; ++c)
        if (best[c] > bst || (best[c] == bst && best[c] > 0 && bestr[c] < bstr)) { bst = best[c]; bstr = bestr[c]; bstd = d0 + c; }
      unsigned long long key = 0;
      if (bst > 0)
        key = ((unsigned long long)(unsigned)bst << 32) | ((unsigned long long)(0xffffu - (unsigned)bstr) << 16) |
              (unsigned long long)(0xffffu - (unsigned)(b.l_edge + bstr + bstd - b.q_left));
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(ALL, key, o);
        key = other > key ? other : key;
      }
      const int max_scor = (int)(key >> 32);
      const int max_r = (int)(0xffffu - (unsigned)((key >> 16) & 0xffffu));
      const int max_j = (int)(0xffffu - (unsigned)(key & 0xffffu)) + b.q_left;
      const int max_i = b.s_left + max_r;
      __syncwarp();
      if (max_scor < minscore) continue;                                    // :1364

      // ---------------- makeMetaFromTrack (alignment.c:628-781), lane 0 ----------------
      int i = max_i, j = max_j, flag = 0;
      uint32_t n = 0;
      if (lane == 0) {
        bool gap_open = false, ovf = false;
        unsigned nmatch = 0;
        int checksum = 0;
        int r = max_r, d = max_j - b.l_edge - max_r;
#define EMIT(c, t) do { if (n < (uint32_t)BWD_REV) sm.rev[n] = DIFFB(c, t); else ovf = true; ++n; } while (0)
        while (i >= b.s_left && j >= b.q_left) {
          const uint32_t dir = (sm.dirs[(d >> 2) * BWD_DIRW + (r >> 2)] >> (((uint32_t)(r & 3) << 3) + ((uint32_t)(d & 3) << 1))) & 3u;
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
      flag = __shfl_sync(ALL, flag, 0);
      i = __shfl_sync(ALL, i, 0);
      j = __shfl_sync(ALL, j, 0);
      n = __shfl_sync(ALL, n, 0);
      if (flag) { err = flag; continue; }
      const int prof_start = j + 1, prof_end = max_j, np_start = i + 1, np_end = max_i;
      if (prof_start + minscorlen > prof_end + 1) continue;                // :1379
      if ((int)nres >= max_res) { err = SMB_ERR_CAPACITY; continue; }      // :1384 addALIMETAtoRsltSet
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
      f2 = __shfl_sync(ALL, f2, 0);
      u = __shfl_sync(ALL, u, 0);
      if (f2) { err = f2; continue; }
      diff_used = u;
      ++nres;
      // pre-order recursion: left part first, so push right then left (:1389, :1411)
      const bool go_left = s_left + minscorlen < np_start;
      const bool go_right = s_right > np_end + minscorlen;
      if (sp + 2 > BWD_STACK && (go_left || go_right)) err = SMB_ERR_CAPACITY;
      else {
        if (go_right) { if (lane == 0) { sm.stk_l[sp] = np_end + 1; sm.stk_r[sp] = s_right; } ++sp; }
        if (go_left) { if (lane == 0) { sm.stk_l[sp] = s_left; sm.stk_r[sp] = np_start - 1; } ++sp; }
      }
      __syncwarp();
    }
    if (lane == 0) {
      out.nres[tix] = nres;
      out.errs[tix] = err;
      if (out.dused) out.dused[tix] = diff_used;
    }
  }
  for (int o = 16; o > 0; o >>= 1) ncell_tot += __shfl_down_sync(ALL, ncell_tot, o);
  if (lane == 0 && ncell_tot) atomicAdd(out.cells, ncell_tot);
}

cudaError_t launch_band_wide(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                             const int *d_order, int ntasks, int *d_ticket, BandOut out, int max_res,
                             const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch) {
  if (ntasks <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  int grid = (ntasks + BWD_WARPS - 1) / BWD_WARPS;
  const int cap = sm_count * 5;
  if (grid > cap) grid = cap;
  band_wide_kernel<<<grid, BWD_WARPS * 32, 0, st>>>(sc, src, d_tasks, d_order, ntasks, d_ticket, out, max_res,
                                                     d_diff_off, d_diff_cap);
  ++*nlaunch;
  return cudaGetLastError();
}

cudaError_t warm_band_wide() {
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, band_wide_kernel);
}

}  // namespace smb
