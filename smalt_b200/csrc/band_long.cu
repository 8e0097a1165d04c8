// band_long.cu - K3 for long reads: bands of hundreds to thousands of diagonals over windows of
// thousands of rows (aliSmiWatInBand, /root/reference/src/alignment.c:1548-1601; DP :788-1027,
// backtrace :628-781, recursion :1300-1434).  With 5-10 kb reads at 12 % error the band is
// (perfect score - best score) / gap extension ~ 1000-2500 diagonals (rmap.c:888-896) and a task has
// ~10^7 cells; one thread per task (band_kernel<true>) takes ~0.1 s for such a task.
//
// ONE CTA PER TASK: the 128 threads form a systolic array over the band - thread t owns the D
// diagonals D*t .. D*t+D-1 (H and E in registers) and computes row r = it - t of all of them in
// iteration `it`, so that the three inputs of a cell are at hand:
//   H(r-1, d)     own register
//   E(r-1, d+1)   the thread's next diagonal as left by the previous iteration; for its last diagonal
//                 the right neighbour's first diagonal of THIS iteration (that thread is one row behind)
//   F(r, d-1)     the diagonal just computed; for the first diagonal the left neighbour's last one of
//                 the previous iteration (that thread is one row ahead)
// The two neighbour values cross through shared memory (two barriers per iteration against D x ~25
// instructions of cell work per thread).  Read and window bases stream from HBM / L1 (the read one new
// base per iteration and thread, the window one row per iteration), direction codes go to an HBM strip,
// 2 bits per cell, laid out by ITERATION (unit (it, t) = the D cells thread t computed in iteration it)
// so that the stores of an iteration coalesce.  The running maximum is the reference's first strict
// maximum in row-major order: per thread the first strict improvement (rows ascend with the
// iterations, columns within an iteration), then a reduction with ties to the smaller (row, column).
// Backtrace: warp 0 walks the path; the direction units of the next 32 rows around the current
// diagonal are fetched by the 32 lanes at once, lane 0 then walks inside that tile from shared memory
// (one DRAM round trip per ~32 steps instead of one per step).
#include "common.cuh"
#include "band.h"
#include "band_cell.cuh"

namespace smb {

constexpr int BL_T = BAND_LONG_THREADS;
constexpr int BL_STACK = 64;

constexpr int BL_CH = 1024;          // window rows per staged chunk of the packed reference
constexpr int BL_CHW = 128;          // words per chunk buffer (1024 bases = 103 words + alignment), one 512-byte bulk copy

// ---- TMA (bulk async copy) staging of the packed reference window: cp.async.bulk global -> shared, completion
// on an mbarrier (one elected thread issues, every thread waits on the barrier's phase) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  } while (!ok);
}

template <int D>
struct LongSmem {
  alignas(16) uint32_t wbuf[2][BL_CHW];   // two chunks of the packed window (TMA destinations)
  alignas(8) unsigned long long bar[2];   // their mbarriers
  int xF[BL_T];                 // F of every thread's last diagonal, previous iteration
  int xE[BL_T];                 // E of every thread's first diagonal, this iteration
  unsigned long long red[BL_T / 32];
  unsigned long long key;
  uint32_t tile[32][3 * (D / 16)];
  int stk_l[BL_STACK], stk_r[BL_STACK];
  int task;
  int bt_i, bt_j, bt_flag;
  uint32_t bt_n;
};

template <int D>
__global__ void __launch_bounds__(BL_T)
band_long_kernel(const Scoring sc, const SeqSrc src, const smb_band_task *__restrict__ tasks,
                 const int *__restrict__ order, const int ntasks, int *__restrict__ ticket, BandOut out,
                 const int max_res, const uint64_t *__restrict__ dir_off, uint32_t *__restrict__ dirs_base,
                 const uint64_t *__restrict__ diff_off, const uint32_t *__restrict__ diff_cap) {
  constexpr int W = D / 16;     // direction words per thread and iteration
  constexpr unsigned ALL = 0xffffffffu;
  __shared__ LongSmem<D> sm;
  __shared__ unsigned long long s_S64[8];
  const int t = threadIdx.x, lane = t & 31;
  if (t < 8) {
    unsigned long long v = 0;
    for (int q = 0; q < 8; ++q) v |= (unsigned long long)(unsigned char)sc.S[t * 8 + q] << (q * 8);
    s_S64[t] = v;
  }
  if (t == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int gi = sc.gap_init, ge = sc.gap_ext;
  unsigned long long ncell_tot = 0;
  uint32_t phase[2] = {0u, 0u};     // parity of the next completion of each chunk barrier (uniform over the CTA)
  const uint64_t packed_words = src.packed_nbases / 10u + 1u;

  for (;;) {
    if (t == 0) sm.task = atomicAdd(ticket, 1);
    __syncthreads();
    const int k = sm.task;
    if (k >= ntasks) break;
    const int tix = __ldg(order + k);
    const smb_band_task tk = tasks[tix];
    const bool rc = (tk.flags & SMB_TASK_READ_REVCOMP) != 0;
    const bool packed = (tk.flags & SMB_TASK_REF_PACKED) != 0;
    const int qlen = (int)tk.read_len, rlen = (int)tk.ref_len;
    int err = SMB_OK;
    uint32_t nres = 0, diff_used = 0;
    int minscore = tk.minscore, minscorlen = tk.minscorlen;
    uint32_t *const dirs = dirs_base + dir_off[tix];
    uint8_t *dfinal = out.diff + diff_off[tix];
    const uint32_t dcap = diff_cap[tix];
    uint8_t *const rev = dfinal + dcap;                 // reversed DiffStr scratch behind the final area
    const uint32_t revcap = (uint32_t)(qlen + rlen + 8);
    smb_ali_result *res = out.results + (size_t)tix * max_res;
    if (minscore < 1 || sc.match <= 0) err = SMB_ERRCODE_ASSERT;         // alignment.c:1569
    else {
      if (minscorlen * sc.match < minscore) minscorlen = minscore / sc.match;  // :1572
      if (minscorlen < 5) err = SMB_ERRCODE_ASSERT;                       // ALILEN_MIN :1574
    }
    int sp = 0;
    if (!err) {
      if (t == 0) { sm.stk_l[0] = tk.u_left; sm.stk_r[0] = tk.u_right; }
      sp = 1;
    }
    __syncthreads();

    while (sp > 0 && !err) {   // one DP pass per row range of the recursion (CTA-uniform)
      bool on = true;
      Band b;
      --sp;
      const int s_left = sm.stk_l[sp], s_right = sm.stk_r[sp];
      if (band_init(b, tk.l_edge, tk.r_edge, tk.p_left, tk.p_right, qlen, s_left, s_right, rlen)) on = false;   // :1333-1338
      else if (b.s_left >= b.s_len || b.band_width < 0) { err = SMB_ERRCODE_ASSERT; on = false; }  // :459
      else if (b.band_width > D * BL_T) { err = SMB_ERR_ARG; on = false; }
      if (!on) { __syncthreads(); continue; }
      const int nrows = b.s_len - b.s_left, bw = b.band_width;

      // ---------------- wavefront DP ----------------
      const int d0 = D * t;
      int H[D], e[D], q[D];
#pragma unroll
      for (int c = 0; c < D; ++c) { H[c] = 0; e[c] = 0; }
      int Flast = 0, best = 0, bestr = 0, bestc = 0;
      unsigned ncell = 0;
      const int nthr = (bw + D - 1) / D;                 // threads that own band diagonals
      const int iters = nrows + nthr - 1;
      const int jbase = b.l_edge + (D - 1) * t;          // column of diagonal d0 + c in row it - t: jbase + it + c
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const int j = jbase + c;
        q[c] = (j >= 0 && j < qlen) ? (int)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)j) : 0;
      }
      sm.xF[t] = 0;
      sm.xE[t] = 0;
      // packed windows are staged by TMA, BL_CH rows per chunk, chunk c in buffer c & 1: rows it - 127 .. it are
      // in use in iteration it, so chunk c + 1 is requested at it = c * BL_CH + 128 (its buffer, that of chunk
      // c - 1, is dead by then) and awaited at it = (c + 1) * BL_CH
      const uint64_t B0 = tk.ref_off + (uint64_t)b.s_left;         // packed base index of window row 0 of this pass
      const uint32_t w0 = (uint32_t)(B0 / 10u), dg0 = (uint32_t)(B0 - (uint64_t)w0 * 10u);
      const int nchunks = packed ? (nrows + BL_CH - 1) / BL_CH : 0;
      auto chunk_word = [&](int c) { return (w0 + (dg0 + (uint32_t)c * BL_CH) / 10u) & ~3u; };
      auto chunk_issue = [&](int c) {
        const uint32_t wl = chunk_word(c);
        uint64_t nw = packed_words > wl ? packed_words - wl : 0;
        if (nw > BL_CHW) nw = BL_CHW;
        nw &= ~(uint64_t)3;                                           // whole 16-byte units
        if (nw < 4) nw = 4;
        bulk_load(sm.wbuf[c & 1], src.packed + wl, (uint32_t)nw * 4u, &sm.bar[c & 1]);
      };
      if (t == 0 && nchunks > 0) {
        chunk_issue(0);
        if (nchunks > 1) chunk_issue(1);
      }
      __syncthreads();
      for (int it = 0; it < iters; ++it) {
        const int r = it - t;
        const bool rowok = r >= 0 && r < nrows && t < nthr;
        if (nchunks > 0) {
          const int k = it / BL_CH, ph = it - k * BL_CH;
          if (ph == 0 && k < nchunks) { mbar_wait(&sm.bar[k & 1], phase[k & 1]); phase[k & 1] ^= 1u; }
          if (t == 0 && ph == 128 && k >= 1 && k + 1 < nchunks) chunk_issue(k + 1);
        }
        const int Fin = (t > 0) ? sm.xF[t - 1] : 0;       // F(r, d0-1): left neighbour, previous iteration
        const int j0 = jbase + it;
        int refc = 0;
        if (rowok) {
          if (packed) {
            const uint32_t x = dg0 + (uint32_t)r, wq = x / 10u, dig = x - wq * 10u;
            const int c = r / BL_CH;
            const uint32_t word = sm.wbuf[c & 1][w0 + wq - chunk_word(c)];
            refc = (int)((word >> (3u * (9u - dig))) & 7u);
          } else {
            refc = (int)ref_base(src, false, tk.ref_off, (uint32_t)(b.s_left + r));
          }
        }
        const unsigned long long srow = s_S64[refc];
        uint32_t dw[W];
#pragma unroll
        for (int w = 0; w < W; ++w) dw[w] = 0u;
        int F = 0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          if (c == D - 1) {                                // E(r-1, d+1) of the last diagonal: right neighbour, this iteration
            __syncthreads();
          }
          const int j = j0 + c;
          const bool ok = rowok && (d0 + c) < bw && j >= b.q_left && j < b.q_len;
          const int s = (int)(signed char)(srow >> (q[c] << 3));
          const int ein = (c == D - 1) ? ((t == BL_T - 1) ? 0 : sm.xE[t + 1]) : e[c + 1];
          const int fin = (c == 0) ? Fin : F;
          // the cell of the restricted recurrence (alignment.c:885-982; band_cell.cuh)
          const int h = H[c] + s;
          const int m = __vimax3_s32(ein, fin, 0);
          const bool dia = h > m;
          int en = ein - ((ein > 0) ? ge : 0);
          int fn = fin - ((fin > 0) ? ge : 0);
          const bool open = dia && h > gi;
          const int tt = open ? h - gi : (int)0x80000000;
          en = max(en, tt);
          fn = max(fn, tt);
          if (ok && open && h > best) { best = h; bestr = r; bestc = c; }
          const uint32_t dcode = dia ? 3u : (m == 0 ? 0u : (ein >= fin ? 1u : 2u));
          H[c] = ok ? max(h, m) : 0;
          e[c] = ok ? en : 0;
          F = ok ? fn : 0;
          if (ok) dw[c >> 4] |= dcode << ((uint32_t)(c & 15) << 1);
          ncell += (unsigned)ok;
          if (c == 0) sm.xE[t] = e[0];                     // for the left neighbour's last diagonal
        }
        Flast = F;
        if (rowok) {
#pragma unroll
          for (int w = 0; w < W; ++w) dirs[((size_t)it * BL_T + t) * W + w] = dw[w];
        }
        // read bases of the next iteration: every column moves one to the right
#pragma unroll
        for (int c = 0; c < D - 1; ++c) q[c] = q[c + 1];
        {
          const int jn = j0 + D;
          q[D - 1] = (jn >= 0 && jn < qlen) ? (int)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)jn) : 0;
        }
        sm.xF[t] = Flast;
        __syncthreads();
      }
      ncell_tot += ncell;
      // first strict maximum in row-major order: max score, then smaller row, then smaller column
      unsigned long long key = 0;
      if (best > 0)
        key = ((unsigned long long)(unsigned)best << 42) | ((unsigned long long)(0x1fffffu - (unsigned)bestr) << 21) |
              (unsigned long long)(0x1fffffu - (unsigned)(b.l_edge + bestr + d0 + bestc - b.q_left));
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(ALL, key, o);
        key = other > key ? other : key;
      }
      if (lane == 0) sm.red[t >> 5] = key;
      __syncthreads();
      if (t == 0) {
        unsigned long long kk = sm.red[0];
        for (int w = 1; w < BL_T / 32; ++w) kk = sm.red[w] > kk ? sm.red[w] : kk;
        sm.key = kk;
      }
      __syncthreads();
      key = sm.key;
      const int max_scor = (int)(key >> 42);
      const int max_r = (int)(0x1fffffu - (unsigned)((key >> 21) & 0x1fffffu));
      const int max_j = (int)(0x1fffffu - (unsigned)(key & 0x1fffffu)) + b.q_left;
      const int max_i = b.s_left + max_r;
      if (max_scor < minscore) { __syncthreads(); continue; }                // :1364

      // ---------------- makeMetaFromTrack (alignment.c:628-781): warp 0 ----------------
      if (t < 32) {
        int i = max_i, j = max_j, flag = 0;
        uint32_t n = 0;
        int r = max_r, d = max_j - b.l_edge - max_r;
        bool gap_open = false, ovf = false, done = false;
        unsigned nmatch = 0;
        int checksum = 0;
        while (!done) {
          // tile: rows r0 - lane, units of the threads tc-1 .. tc+1
          const int r0 = r, tc = d / D;
          {
            const int rr = r0 - lane;
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              const int tu = tc - 1 + u;
#pragma unroll
              for (int w = 0; w < W; ++w) {
                uint32_t v = 0;
                if (rr >= 0 && tu >= 0 && tu < BL_T) v = dirs[((size_t)(rr + tu) * BL_T + tu) * W + w];
                sm.tile[lane][u * W + w] = v;
              }
            }
          }
          __syncwarp();
          if (lane == 0) {
#define EMIT(c, ty) do { if (n < revcap) rev[n] = DIFFB(c, ty); else ovf = true; ++n; } while (0)
            for (;;) {
              if (!(i >= b.s_left && j >= b.q_left)) { done = true; break; }
              const int tu = d / D - (tc - 1);
              if (r > r0 || r < r0 - 31 || r < 0 || tu < 0 || tu > 2) break;      // next tile
              const int c = d % D;
              const uint32_t dir = (sm.tile[r0 - r][tu * W + (c >> 4)] >> ((uint32_t)(c & 15) << 1)) & 3u;
              if (!dir) { done = true; break; }
              if (dir == 3u) {
                const int rb = (int)ref_base(src, packed, tk.ref_off, (uint32_t)i);
                const int qb = (int)read_base(src.arena, tk.read_off, tk.read_len, rc, (uint32_t)j);
                const int s = (int)(signed char)(s_S64[rb] >> (qb << 3));
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
          }
          done = __shfl_sync(ALL, (int)done, 0) != 0;
          r = __shfl_sync(ALL, r, 0);
          d = __shfl_sync(ALL, d, 0);
          if (!done && (r < 0 || d < 0)) done = true;      // (cannot happen: the walk ends at the segment edge first)
          __syncwarp();
        }
        if (lane == 0) {
          EMIT(nmatch, 3u);
          EMIT(0u, 0u);
#undef EMIT
          if (ovf) flag = SMB_ERR_CAPACITY;
          else if (checksum != max_scor) flag = SMB_ERRCODE_SWATSCOR;        // :767
          sm.bt_i = i; sm.bt_j = j; sm.bt_flag = flag; sm.bt_n = n;
        }
      }
      __syncthreads();
      const int flag = sm.bt_flag, i = sm.bt_i, j = sm.bt_j;
      const uint32_t n = sm.bt_n;
      __syncthreads();
      if (flag) { err = flag; continue; }
      const int prof_start = j + 1, prof_end = max_j, np_start = i + 1, np_end = max_i;
      if (prof_start + minscorlen > prof_end + 1) continue;                // :1379
      if ((int)nres >= max_res) { err = SMB_ERR_CAPACITY; continue; }      // :1384 addALIMETAtoRsltSet
      if (t == 0) {
        int f2 = 0;
        uint32_t u = diff_used;
        // diffStrReverse (diffstr.c:850-896)
        int l = (int)n - 2;
        if (l >= 32767) f2 = SMB_ERRCODE_OVERFLOW;
        else if ((rev[l] >> 6) != 3u) f2 = SMB_ERRCODE_DIFFSTR;
        else {
          unsigned count_prev = rev[l] & 0x3Fu;
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
          if (dovf) f2 = SMB_ERR_CAPACITY;
          else {
            smb_ali_result rr;
            rr.score = max_scor; rr.qs = prof_start; rr.qe = prof_end; rr.rs = np_start; rr.re = np_end;
            rr.diff_off = diff_used; rr.diff_len = u - diff_used; rr.task = (uint32_t)tix;
            res[nres] = rr;
          }
        }
        sm.bt_flag = f2;
        sm.bt_n = u;
      }
      __syncthreads();
      const int f2 = sm.bt_flag;
      const uint32_t u = sm.bt_n;
      __syncthreads();
      if (f2) { err = f2; continue; }
      diff_used = u;
      ++nres;
      // pre-order recursion: left part first, so push right then left (:1389, :1411)
      const bool go_left = s_left + minscorlen < np_start;
      const bool go_right = s_right > np_end + minscorlen;
      if (sp + 2 > BL_STACK && (go_left || go_right)) err = SMB_ERR_CAPACITY;
      else {
        if (go_right) { if (t == 0) { sm.stk_l[sp] = np_end + 1; sm.stk_r[sp] = s_right; } ++sp; }
        if (go_left) { if (t == 0) { sm.stk_l[sp] = s_left; sm.stk_r[sp] = np_start - 1; } ++sp; }
      }
      __syncthreads();
    }
    if (t == 0) {
      out.nres[tix] = nres;
      out.errs[tix] = err;
      if (out.dused) out.dused[tix] = diff_used;
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) ncell_tot += __shfl_down_sync(ALL, ncell_tot, o);
  if (lane == 0 && ncell_tot) atomicAdd(out.cells, ncell_tot);
}

cudaError_t launch_band_long(const Scoring &sc, const SeqSrc &src, const smb_band_task *d_tasks, const int *d_order,
                             int ntasks, int dpt, int *d_ticket, BandOut out, int max_res, const uint64_t *d_dir_off,
                             uint32_t *d_dirs, const uint64_t *d_diff_off, const uint32_t *d_diff_cap, int sm_count,
                             cudaStream_t st, int *nlaunch) {
  if (ntasks <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  int grid = ntasks;
  const int cap = sm_count * 4;
  if (grid > cap) grid = cap;
  if (dpt == 16)
    band_long_kernel<16><<<grid, BL_T, 0, st>>>(sc, src, d_tasks, d_order, ntasks, d_ticket, out, max_res, d_dir_off, d_dirs,
                                                 d_diff_off, d_diff_cap);
  else
    band_long_kernel<32><<<grid, BL_T, 0, st>>>(sc, src, d_tasks, d_order, ntasks, d_ticket, out, max_res, d_dir_off, d_dirs,
                                                 d_diff_off, d_diff_cap);
  ++*nlaunch;
  return cudaGetLastError();
}

cudaError_t warm_band_long() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, band_long_kernel<16>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, band_long_kernel<32>);
  return e;
}

}  // namespace smb
