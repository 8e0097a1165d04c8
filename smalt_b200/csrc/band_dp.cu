// band_dp.cu - K2' and K3: banded "restricted" Smith-Waterman for sm_100a.
//
// Replaces aliSmiWatInBandFast (/root/reference/src/alignment.c:1603-1638 ->
// alignSmiWatBandFast :1029-1233, score only) and aliSmiWatInBand (:1548-1601 ->
// alignSmiWatBandRecursive :1300-1434 -> alignSmiWatBand :788-1027 + makeMetaFromTrack
// :628-781 + diffStrReverse diffstr.c:850-896): banded DP with one direction code per cell,
// first-strict-maximum argmax, backtrace to a DiffStr and recursion on the window rows left
// and right of each local alignment.
//
// The recurrence is NOT the canonical one of K2: a gap state is only opened from a cell
// whose H came from the diagonal, non-positive E/F mean "no gap" and a maximum counts only
// when H > gap_init (alignment.c:885-982).  The four-way case analysis of the reference is
// restated branch-free (derivation in DESIGN.md):
//     Ep = max(E,0); Fp = max(F,0); M = max(Ep,Fp); dia = h > M; H = dia ? h : M
//     E -= (E>0)*ext; F -= (F>0)*ext; if (dia && h > init) { E = max(E,h-init); F = max(F,h-init) }
//     dir = dia ? DIA : (M == 0 ? 0 : (Ep >= Fp ? COL : ROW))
//
// B200 mapping: short-read bands are narrow (tens of columns) and tasks are plentiful
// (millions per batch), so the kernel is inter-task parallel with ONE THREAD PER TASK:
//  * the H/E band row lives in shared memory as a ring of packed s16x2 words, laid out
//    [slot][thread] so that a warp's accesses are bank-conflict free;
//  * row-major evaluation order inside a thread makes the reference's "first strict
//    maximum in row-major order" argmax (alignment.c:826-830) fall out naturally;
//  * direction codes are packed 2 bits/cell into 32-bit words in registers and stored to a
//    per-task strip in HBM (16 cells per store), addressed exactly like the reference's byte
//    matrix (cell (r,j) at r*(bw-1) + j - l_edge, alignment.c:676);
//  * backtrace, DiffStr reversal and the recursion (explicit stack, pre-order like the
//    reference) run in the same thread right after the DP, so a task is one kernel pass.
// Tasks are bucketed by ring width on the host and sorted by size so that the threads of a
// warp work on similar problems.  Bands wider than the shared-memory ring (long reads) use
// the same code with the ring in an HBM scratch strip.
#include "common.cuh"
#include "band.h"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace smb {

constexpr int BAND_THREADS = 64;
constexpr int MAX_STACK = 48;

struct BandArgs {
  const int *order;
  int ntasks;
  int wcap;             // ring capacity (power of two)
  uint32_t *gring;      // global ring (nullptr: shared memory)
  int max_res;
};

struct Track { int max_i, max_j, max_scor; };

// One banded DP pass (alignSmiWatBand / alignSmiWatBandFast).
template <bool ALIGN>
__device__ __forceinline__ void band_dp(const Scoring &sc, const Band &b, const SeqSrc &src,
                                        const smb_band_task &tk, bool rc, bool packed,
                                        uint32_t *ring, int rstride, int wmask,
                                        const unsigned long long *s_S64,
                                        uint32_t *dirs, Track &tr, unsigned long long &ncell) {
  const int gi = sc.gap_init, ge = sc.gap_ext;
  int dstart, jstart;
  if (b.q_left > b.l_edge) { dstart = b.q_left - b.l_edge; jstart = b.q_left; }
  else { dstart = 0; jstart = b.l_edge; }
  int jlen = b.r_edge + 1;
  int max_i = 0, max_j = 0, max_scor = 0, currH = 0;
  // direction word accumulator
  long long ci = dstart;  // running cell index (reference's dirp - bdp)
  long long cur_w = -1;
  uint32_t cur_bits = 0;
  int dend = 0;
  for (int j = jstart; j < jlen; ++j) ring[(j & wmask) * rstride] = 0u;

  for (int i = b.s_left; i < b.s_len; ++i) {
    const int r = (int)ref_base(src, packed, tk.ref_off, (uint32_t)i);
    const unsigned long long srow = s_S64[r];  // the 8 substitution scores of this window base
    int F = 0;
    for (int j = jstart; j < jlen; ++j) {
      const int q = (int)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)j);
      uint32_t *cellp = ring + (j & wmask) * rstride;
      const uint32_t he = *cellp;
      const int hprev = (int)(short)(he & 0xffffu);
      int e = (int)(short)(he >> 16);
      const int h = currH + (int)(signed char)(srow >> (q << 3));
      currH = hprev;
      const int ep = max(e, 0), fp = max(F, 0);
      const int m = max(ep, fp);
      const bool dia = h > m;
      const int hn = dia ? h : m;
      e -= (e > 0) ? ge : 0;
      F -= (F > 0) ? ge : 0;
      if (dia && h > gi) {
        const int t = h - gi;
        if (h > max_scor) { max_scor = h; max_i = i; max_j = j; }
        e = max(e, t);
        F = max(F, t);
      }
      *cellp = ((uint32_t)(unsigned short)hn) | ((uint32_t)(unsigned short)e << 16);
      if (ALIGN) {
        const uint32_t d = dia ? 3u : (m == 0 ? 0u : (ep >= fp ? 1u : 2u));
        const long long w = ci >> 4;
        if (w != cur_w) {
          if (cur_w >= 0) dirs[cur_w] = cur_bits;
          cur_w = w;
          cur_bits = 0;
        }
        cur_bits |= d << ((uint32_t)(ci & 15) * 2u);
        ++ci;
      }
    }
    if (jlen > jstart) ncell += (unsigned long long)(jlen - jstart);
    if (dstart > 0) {
      currH = 0;
      if (ALIGN) ci += --dstart;  // the fast variant never releases the clipped start (:1213-1218)
    } else {
      currH = (int)(short)(ring[(jstart & wmask) * rstride] & 0xffffu);
      ++jstart;
    }
    if (jlen < b.q_len) {
      ring[(jlen & wmask) * rstride] = 0u;  // column entering the band: H = E = 0
      ++jlen;
    } else if (ALIGN) {
      ci += dend++;
    }
  }
  if (ALIGN && cur_w >= 0) dirs[cur_w] = cur_bits;
  tr.max_i = max_i; tr.max_j = max_j; tr.max_scor = max_scor;
}

__device__ __forceinline__ uint32_t dir_at(const uint32_t *dirs, long long ci) {
  return (dirs[ci >> 4] >> ((uint32_t)(ci & 15) * 2u)) & 3u;
}

#define DIFFB(count, typ) ((uint8_t)((count) + ((typ) << 6)))

template <bool ALIGN>
__global__ void __launch_bounds__(BAND_THREADS)
band_kernel(const Scoring sc, const SeqSrc src, const smb_band_task *__restrict__ tasks,
            const BandArgs a, int32_t *__restrict__ scores, BandOut out,
            const uint64_t *__restrict__ dir_off, uint32_t *__restrict__ dirs_base,
            const uint64_t *__restrict__ diff_off, const uint32_t *__restrict__ diff_cap) {
  extern __shared__ uint32_t s_ring[];
  __shared__ unsigned long long s_S64[8];  // substitution matrix rows, one 64-bit word per window base
  if (threadIdx.x < 8) {
    unsigned long long v = 0;
    for (int q = 0; q < 8; ++q) v |= (unsigned long long)(unsigned char)sc.S[threadIdx.x * 8 + q] << (q * 8);
    s_S64[threadIdx.x] = v;
  }
  __syncthreads();
  const int g = blockIdx.x * BAND_THREADS + threadIdx.x;
  if (g >= a.ntasks) return;
  const int tix = a.order[g];
  const smb_band_task tk = tasks[tix];
  const bool rc = (tk.flags & SMB_TASK_READ_REVCOMP) != 0;
  const bool packed = (tk.flags & SMB_TASK_REF_PACKED) != 0;
  const int qlen = (int)tk.read_len, rlen = (int)tk.ref_len;
  uint32_t *ring;
  int rstride;
  if (a.gring) {
    ring = a.gring + (size_t)blockIdx.x * BAND_THREADS * a.wcap + threadIdx.x;
    rstride = BAND_THREADS;
  } else {
    ring = s_ring + threadIdx.x;
    rstride = BAND_THREADS;
  }
  const int wmask = a.wcap - 1;
  unsigned long long ncell = 0;
  Band b;
  Track tr;

  if (!ALIGN) {
    int err = SMB_OK, score = 0;
    if (band_init(b, tk.l_edge, tk.r_edge, tk.p_left, tk.p_right, qlen, tk.u_left, tk.u_right, rlen)) {
      err = SMB_ERRCODE_FAILURE;  // alignment.c:1622-1627
    } else {
      band_dp<false>(sc, b, src, tk, rc, packed, ring, rstride, wmask, s_S64, nullptr, tr, ncell);
      score = tr.max_scor;
    }
    scores[tix] = score;
    out.errs[tix] = err;
    if (ncell) atomicAdd(out.cells, ncell);
    return;
  }

  // ---- K3: aliSmiWatInBand ----
  int err = SMB_OK;
  uint32_t nres = 0, diff_used = 0;
  int minscore = tk.minscore, minscorlen = tk.minscorlen;
  uint32_t *dirs = dirs_base + dir_off[tix];
  uint8_t *dfinal = out.diff + diff_off[tix];
  const uint32_t dcap = diff_cap[tix];          // final area; the reversed scratch follows it
  uint8_t *rev = dfinal + dcap;
  const uint32_t revcap = (uint32_t)(qlen + rlen + 8);
  smb_ali_result *res = out.results + (size_t)tix * a.max_res;

  if (minscore < 1 || sc.match <= 0) err = SMB_ERRCODE_ASSERT;         // alignment.c:1569
  else {
    if (minscorlen * sc.match < minscore) minscorlen = minscore / sc.match;  // :1572
    if (minscorlen < 5) err = SMB_ERRCODE_ASSERT;                       // ALILEN_MIN :1574
  }
  int stk_l[MAX_STACK], stk_r[MAX_STACK];
  int sp = 0;
  if (!err) { stk_l[0] = tk.u_left; stk_r[0] = tk.u_right; sp = 1; }
  while (sp > 0 && !err) {
    --sp;
    const int s_left = stk_l[sp], s_right = stk_r[sp];
    if (band_init(b, tk.l_edge, tk.r_edge, tk.p_left, tk.p_right, qlen, s_left, s_right, rlen))
      continue;                                                          // :1333-1338
    if (b.s_left >= b.s_len || b.band_width < 0) { err = SMB_ERRCODE_ASSERT; break; }  // :459
    band_dp<true>(sc, b, src, tk, rc, packed, ring, rstride, wmask, s_S64, dirs, tr, ncell);
    if (tr.max_scor < minscore) continue;                                // :1364
    // ---- makeMetaFromTrack (alignment.c:628-781) ----
    int i = tr.max_i, j = tr.max_j, checksum = 0;
    uint32_t n = 0;
    bool gap_open = false, ovf = false;
    unsigned nmatch = 0;
    long long ci = (long long)(tr.max_i - b.s_left) * (b.band_width - 1) + tr.max_j - b.l_edge;
#define EMIT(c, t) do { if (n < revcap) rev[n] = DIFFB(c, t); else ovf = true; ++n; } while (0)
    while (i >= b.s_left && j >= b.q_left) {
      const uint32_t d = dir_at(dirs, ci);
      if (!d) break;
      if (d == 3u) {
        const int r = (int)ref_base(src, packed, tk.ref_off, (uint32_t)i);
        const int q = (int)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)j);
        const int s = (int)(signed char)(s_S64[r] >> (q << 3));
        if (s > 0) {
          if (nmatch > 61u) { EMIT(61u, 0u); nmatch -= 61u; }
          else ++nmatch;
        } else {
          EMIT(nmatch, 3u);
          nmatch = 0;
        }
        checksum += s;
        gap_open = false;
        ci -= b.band_width;
        --i; --j;
        continue;
      }
      if (gap_open) checksum -= sc.gap_ext;
      else { checksum -= sc.gap_init; gap_open = true; }
      if (d & 1u) {
        EMIT(nmatch, 1u);
        nmatch = 0;
        ci -= b.band_width - 1;
        --i;
        continue;
      }
      EMIT(nmatch, 2u);
      nmatch = 0;
      --ci;
      --j;
    }
    EMIT(nmatch, 3u);
    EMIT(0u, 0u);
#undef EMIT
    if (ovf) { err = SMB_ERR_CAPACITY; break; }
    const int prof_start = j + 1, prof_end = tr.max_j, np_start = i + 1, np_end = tr.max_i;
    if (checksum != tr.max_scor) { err = SMB_ERRCODE_SWATSCOR; break; }  // :767
    if (prof_start + minscorlen > prof_end + 1) continue;                // :1379
    if (checksum >= minscore) {                                          // :1384 addALIMETAtoRsltSet
      if ((int)nres >= a.max_res) { err = SMB_ERR_CAPACITY; break; }
      // diffStrReverse (diffstr.c:850-896): rev[0..n-2] + terminator -> forward string
      int l = (int)n - 2;  // last non-terminator byte
      if (l >= 32767) { err = SMB_ERRCODE_OVERFLOW; break; }
      unsigned count_prev = rev[l] & 0x3Fu;
      if ((rev[l] >> 6) != 3u) { err = SMB_ERRCODE_DIFFSTR; break; }
      uint32_t u = diff_used;
      bool dovf = false;
#define PUT(v) do { if (u < dcap) dfinal[u] = (v); else dovf = true; ++u; } while (0)
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
      if (dovf) { err = SMB_ERR_CAPACITY; break; }
      smb_ali_result rr;
      rr.score = checksum; rr.qs = prof_start; rr.qe = prof_end; rr.rs = np_start; rr.re = np_end;
      rr.diff_off = diff_used; rr.diff_len = u - diff_used; rr.task = (uint32_t)tix;
      res[nres++] = rr;
      diff_used = u;
    }
    // pre-order recursion: left part first, so push right then left (:1389, :1411)
    const bool go_left = s_left + minscorlen < np_start;
    const bool go_right = s_right > np_end + minscorlen;
    if (sp + 2 > MAX_STACK && (go_left || go_right)) { err = SMB_ERR_CAPACITY; break; }
    if (go_right) { stk_l[sp] = np_end + 1; stk_r[sp] = s_right; ++sp; }
    if (go_left) { stk_l[sp] = s_left; stk_r[sp] = np_start - 1; ++sp; }
  }
  out.nres[tix] = nres;
  out.errs[tix] = err;
  if (out.dused) out.dused[tix] = diff_used;
  if (ncell) atomicAdd(out.cells, ncell);
}

// ------------------------------------------------------------------------------------
// host side planning
// ------------------------------------------------------------------------------------
// Host-side plan: tasks bucketed by ring capacity (one launch per class), input order kept
// inside a class (neighbouring tasks come from the same read / candidate list and have
// similar geometry, so the threads of a warp stay balanced without a full sort).
void plan_band(const smb_band_task *h_tasks, int ntasks, bool align, const Scoring &sc, BandPlan &plan) {
  plan.order.resize((size_t)ntasks);
  plan.classes.clear();
  std::vector<int> wc((size_t)ntasks);
  int count[32] = {0};
  constexpr int WARP_CLS = BAND_CLS_WARP, HALF_CLS = BAND_CLS_HALF, PACK_CLS = BAND_CLS_PACK, WIDE_CLS = BAND_CLS_WIDE,
                PACK8_CLS = BAND_CLS_PACK8;
  // packed 16-bit kernel: scores must stay far below 2^15 (band_pack.cu)
  const bool pen16 = sc.match > 0 && sc.match < 128 && sc.mismatch <= 0 && sc.mismatch > -128 && sc.gap_init > 0 &&
                     sc.gap_init < 4000 && sc.gap_ext >= 0 && sc.gap_ext < 4000 && sc.S[5] == 0 && sc.S[5 * 8] == 0 &&
                     !getenv("SMB_NO_PACK");
  const bool no_wide = getenv("SMB_NO_WIDE") != nullptr;
  // the eight-tasks-per-warp geometry is measured slower on B200: opt-in
  const bool no_pack8 = getenv("SMB_PACK8") == nullptr;
  plan.pack_maxrows = plan.pack8_maxrows = 0;
  plan.pack_maxread = plan.pack8_maxread = 0;
  for (int i = 0; i < ntasks; ++i) {
    const smb_band_task &t = h_tasks[i];
    const int wl = align ? band_warp_lanes(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left,
                                           t.u_right, (int)t.ref_len) : 0;
    if (wl == 16 && pen16 && (long long)t.read_len * sc.match <= 255) {   // (maximum keys of band_pack.cu: score < 256)
      if (!no_pack8 && band_width_bound(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                        (int)t.ref_len) <= BP8_MAXDIAG) {
        wc[(size_t)i] = PACK8_CLS;
        plan.pack8_maxrows = std::max(plan.pack8_maxrows, (int)t.ref_len);
        plan.pack8_maxread = std::max(plan.pack8_maxread, (int)t.read_len);
      } else {
        wc[(size_t)i] = PACK_CLS;
        plan.pack_maxrows = std::max(plan.pack_maxrows, (int)t.ref_len);
        plan.pack_maxread = std::max(plan.pack_maxread, (int)t.read_len);
      }
    } else if (wl) {
      wc[(size_t)i] = wl == 16 ? HALF_CLS : WARP_CLS;
    } else if (align && !no_wide &&
               band_wide_eligible(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                  (int)t.ref_len)) {
      wc[(size_t)i] = WIDE_CLS;
    } else if (const int dpt = align ? band_long_dpt(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                                     (int)t.ref_len) : 0) {
      wc[(size_t)i] = dpt == 16 ? BAND_CLS_LONG16 : BAND_CLS_LONG32;
    } else {
      wc[(size_t)i] = band_ring_class(band_ring_need(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left,
                                                      t.u_right, (int)t.ref_len, !align));
    }
    count[wc[(size_t)i]]++;
  }
  int start[33];
  start[0] = 0;
  for (int c = 0; c < 32; ++c) start[c + 1] = start[c] + count[c];
  int fill[32];
  for (int c = 0; c < 32; ++c) fill[c] = start[c];
  for (int i = 0; i < ntasks; ++i) plan.order[(size_t)fill[wc[(size_t)i]]++] = i;
  for (int c = 0; c < BAND_CLS_THREAD_END; ++c)
    if (count[c]) plan.classes.push_back(BandPlan::Class{32 << c, start[c], count[c]});
  plan.long16_start = start[BAND_CLS_LONG16];
  plan.long16_count = count[BAND_CLS_LONG16];
  plan.long32_start = start[BAND_CLS_LONG32];
  plan.long32_count = count[BAND_CLS_LONG32];
  plan.wide_start = start[WIDE_CLS];
  plan.wide_count = count[WIDE_CLS];
  plan.pack8_start = start[PACK8_CLS];
  plan.pack8_count = count[PACK8_CLS];
  plan.pack_start = start[PACK_CLS];
  plan.pack_count = count[PACK_CLS];
  plan.half_start = start[HALF_CLS];
  plan.half_count = count[HALF_CLS];
  plan.warp_start = start[WARP_CLS];
  plan.warp_count = count[WARP_CLS];
}

cudaError_t launch_band(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks,
                        const BandPlan &plan, const int *d_order, bool align,
                        int32_t *d_scores, BandOut out, int max_res,
                        const uint64_t *d_dir_off, uint32_t *d_dirs,
                        const uint64_t *d_diff_off, const uint32_t *d_diff_cap,
                        uint32_t *d_gring, int *d_ticket, int sm_count, cudaStream_t main_st, int *nlaunch,
                        const BandSide *side) {
  static std::atomic<unsigned long long> smem_done_a{0}, smem_done_f{0};
  cudaError_t e;
  if ((e = ensure_dyn_smem(band_kernel<true>, 200 * 1024, smem_done_a)) != cudaSuccess) return e;
  if ((e = ensure_dyn_smem(band_kernel<false>, 200 * 1024, smem_done_f)) != cudaSuccess) return e;
  // the packed kernels take (nearly) all tasks of a short-read batch; whatever else the batch holds goes
  // to the side stream
  const bool packed_any = align && (plan.pack_count || plan.pack8_count);
  const bool other_any = !plan.classes.empty() || (align && (plan.wide_count || plan.half_count || plan.warp_count || plan.long16_count || plan.long32_count));
  static const bool no_side = getenv("SMB_NO_SIDE") != nullptr;
  const bool fork = side && side->stream && packed_any && other_any && !no_side;
  cudaStream_t st = main_st;
  if (fork) {
    if ((e = cudaEventRecord(side->fork, main_st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(side->stream, side->fork, 0)) != cudaSuccess) return e;
    st = side->stream;
  }
  for (const BandPlan::Class &c : plan.classes) {
    const int grid = (c.count + BAND_THREADS - 1) / BAND_THREADS;
    BandArgs a{d_order + c.start, c.count, c.wcap, nullptr, max_res};
    size_t smem = (size_t)c.wcap * BAND_THREADS * sizeof(uint32_t);
    if (c.wcap > BAND_SMEM_WCAP_MAX) {  // long-read bands: ring in an HBM strip
      a.gring = d_gring;
      smem = 0;
    }
    if (align)
      band_kernel<true><<<grid, BAND_THREADS, smem, st>>>(sc, src, d_tasks, a, d_scores, out, d_dir_off,
                                                          d_dirs, d_diff_off, d_diff_cap);
    else
      band_kernel<false><<<grid, BAND_THREADS, smem, st>>>(sc, src, d_tasks, a, d_scores, out, d_dir_off,
                                                           d_dirs, d_diff_off, d_diff_cap);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*nlaunch;
  }
  if (align && plan.long32_count &&
      (e = launch_band_long(sc, src, d_tasks, d_order + plan.long32_start, plan.long32_count, 32, d_ticket + 5, out, max_res,
                            d_dir_off, d_dirs, d_diff_off, d_diff_cap, sm_count, st, nlaunch)) != cudaSuccess)
    return e;
  if (align && plan.long16_count &&
      (e = launch_band_long(sc, src, d_tasks, d_order + plan.long16_start, plan.long16_count, 16, d_ticket + 6, out, max_res,
                            d_dir_off, d_dirs, d_diff_off, d_diff_cap, sm_count, st, nlaunch)) != cudaSuccess)
    return e;
  if (align && plan.wide_count &&
      (e = launch_band_wide(sc, src, d_tasks, d_order + plan.wide_start, plan.wide_count, d_ticket + 3, out, max_res,
                            d_diff_off, d_diff_cap, sm_count, st, nlaunch)) != cudaSuccess)
    return e;
  if (align && plan.half_count &&
      (e = launch_band_warp(sc, src, d_tasks, d_order + plan.half_start, plan.half_count, 16, d_ticket, out, max_res,
                            d_diff_off, d_diff_cap, sm_count, st, nlaunch)) != cudaSuccess)
    return e;
  if (align && plan.warp_count &&
      (e = launch_band_warp(sc, src, d_tasks, d_order + plan.warp_start, plan.warp_count, 32, d_ticket + 1, out,
                            max_res, d_diff_off, d_diff_cap, sm_count, st, nlaunch)) != cudaSuccess)
    return e;
  if (fork && (e = cudaEventRecord(side->join, side->stream)) != cudaSuccess) return e;
  st = main_st;
  if (align && plan.pack_count &&
      (e = launch_band_pack(sc, src, d_tasks, d_order + plan.pack_start, plan.pack_count, 16, plan.pack_maxrows, plan.pack_maxread, d_ticket + 2,
                            out, max_res, d_diff_off, d_diff_cap, sm_count, st, nlaunch)) != cudaSuccess)
    return e;
  if (align && plan.pack8_count &&
      (e = launch_band_pack(sc, src, d_tasks, d_order + plan.pack8_start, plan.pack8_count, 8, plan.pack8_maxrows, plan.pack8_maxread, d_ticket + 4,
                            out, max_res, d_diff_off, d_diff_cap, sm_count, st, nlaunch)) != cudaSuccess)
    return e;
  if (fork && (e = cudaStreamWaitEvent(main_st, side->join, 0)) != cudaSuccess) return e;
  return cudaSuccess;
}

size_t band_gring_words(const BandPlan &plan) {
  size_t need = 0;
  for (const BandPlan::Class &c : plan.classes)
    if (c.wcap > BAND_SMEM_WCAP_MAX) {
      const size_t grid = (size_t)(c.count + BAND_THREADS - 1) / BAND_THREADS;
      need = std::max(need, grid * BAND_THREADS * (size_t)c.wcap);
    }
  return need;
}

cudaError_t warm_band() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, band_kernel<true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, band_kernel<false>);
  return e;
}

}  // namespace smb
