// seed.cu - K1: k-mer seed lookup, seed ranking and hit-list construction for sm_100a.
//
// Replaces, for a whole batch of reads and both strands,
//   collectHitInfo            /root/reference/src/hashhit.c:480-657
//   hashTableGetKtupleHits    hashidx.c:1146-1191 (+ MAKE_HASHKEY :155-158, hash32mix :163-172)
//   hashCollectHitInfoShort   hashhit.c:1007-1080 (quicksort sort.c:233-330 + getHitInfoMaxRank
//                             hashhit.c:769-891)
//   hashCalcHitInfoCoverDeficit :1096, hashHitInfoCalcHitNumbers :1200,
//   hashCalcHitInfoNumberOfHits :1171
//   hashCollectHitsForSegment :1691-1769 (fillHitListFromHitInfoSegment :1416-1546) and
//   hashCollectHitsUsingCutoff :1593-1689, incl. the ascending sort of the packed hits.
//
// B200 mapping.  The index (idx / wordidx / posidx / pos) is uploaded once and stays in HBM;
// for bacterial-size genomes it is L2 resident (126 MB L2).  A lookup is a chain of dependent
// 4-byte loads (idx pair -> binary search in wordidx -> posidx pair), i.e. latency bound, so
// the kernel exposes memory-level parallelism across reads: ONE THREAD PER read x strand,
// tens of thousands of independent chains in flight per SM wave.  The per-read sequential
// parts of the reference (non-standard countdown, 4-deep tandem-repeat filter, the UNSTABLE
// quicksort whose tie order decides the seed_rank cut, the per-frame coverage scan) run
// unchanged inside that thread, which keeps them bit-identical by construction.
// Per-strand tables live in HBM in SoA form (read_len slots per strand).
#include "common.cuh"
#include "sort2.cuh"
#include <vector>

namespace smb {

__device__ __forceinline__ uint32_t hash32mix(uint32_t a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

// hashTableGetKtupleHits
__device__ __forceinline__ uint32_t lookup(const Index &ix, uint64_t word, uint32_t &posidx) {
  if (ix.typ == 0) {
    const uint32_t key = (uint32_t)(word & ix.wordmask);
    posidx = key;
    return (key < ix.nkeys) ? __ldg(ix.idx + key + 1) - __ldg(ix.idx + key) : 0u;
  }
  const uint32_t word_hi = (uint32_t)((word & ix.wordmask_hi) >> ix.nbits_lo);
  const uint32_t key = ((hash32mix(word_hi) % ix.keymod) << ix.nbits_lo) + (uint32_t)(word & ix.wordmask_lo);
  uint32_t b = __ldg(ix.idx + key + 1);
  if (b < 1) return 0;
  uint32_t a = __ldg(ix.idx + key);
  --b;
  while (a < b) {
    const uint32_t pivot = (a + b) >> 1;
    if (__ldg(ix.wordidx + pivot) < word_hi) a = pivot + 1; else b = pivot;
  }
  if (a == b && __ldg(ix.wordidx + b) == word_hi) {
    posidx = b;
    return __ldg(ix.posidx + b + 1) - __ldg(ix.posidx + b);
  }
  return 0;
}


// The same sort by all lanes of a warp.  The result of sort2 for a sub-array depends on nothing but
// the sub-array, and the two parts a partitioning step leaves are disjoint - so they can be sorted
// in any order, or at the same time, and the arrangement (ties included) stays that of the
// reference's sequential run.  Level by level: lane i takes the i-th pending range and either
// finishes it (straight insertion below 7 elements, sort.c:251-262) or partitions it exactly like
// the reference (sort.c:264-316) and appends the two parts to the next level's list.  The critical
// path is n + n/2 + n/4 + ... element steps instead of n log n.
// `ranges`: room for 2 lists of (cap) ranges = 4 * cap ints.
__device__ void sort2_warp(int n, uint32_t *key, uint32_t *val, int *ranges, int cap, int lane) {
  const unsigned FULL = 0xffffffffu;
  int *cur = ranges, *nxt = ranges + 2 * cap;
  int ncur = 1;
  if (lane == 0) { cur[0] = 0; cur[1] = n - 1; }
  __syncwarp();
#define XC(a, b) do { uint32_t t_ = (a); (a) = (b); (b) = t_; } while (0)
  while (ncur > 0) {
    int nnext = 0;
    for (int base = 0; base < ncur; base += 32) {
      const int idx = base + lane;
      int c0lo = 0, c0hi = -1, c1lo = 0, c1hi = -1;
      if (idx < ncur) {
        const int lo = cur[2 * idx], hi = cur[2 * idx + 1];
        if (hi - lo < 7) {
          for (int j = lo + 1; j <= hi; ++j) {
            const uint32_t k = key[j], v = val[j];
            int i;
            for (i = j - 1; i >= lo && key[i] > k; --i) { key[i + 1] = key[i]; val[i + 1] = val[i]; }
            key[i + 1] = k; val[i + 1] = v;
          }
        } else {
          const int mid = (lo + hi) >> 1;
          XC(key[mid], key[lo + 1]); XC(val[mid], val[lo + 1]);
          if (key[lo] > key[hi]) { XC(key[lo], key[hi]); XC(val[lo], val[hi]); }
          if (key[lo + 1] > key[hi]) { XC(key[lo + 1], key[hi]); XC(val[lo + 1], val[hi]); }
          if (key[lo] > key[lo + 1]) { XC(key[lo], key[lo + 1]); XC(val[lo], val[lo + 1]); }
          int i = lo + 1, j = hi;
          const uint32_t pk = key[lo + 1], pv = val[lo + 1];
          for (;;) {
            do ++i; while (key[i] < pk);
            do --j; while (key[j] > pk);
            if (j < i) break;
            XC(key[i], key[j]); XC(val[i], val[j]);
          }
          key[lo + 1] = key[j]; val[lo + 1] = val[j];
          key[j] = pk; val[j] = pv;
          c0lo = lo; c0hi = j - 1;      // the two parts (the reference pushes the larger, goes on with the smaller)
          c1lo = i; c1hi = hi;
        }
      }
      const int n0 = c0hi > c0lo, n1 = c1hi > c1lo;   // parts of fewer than two elements are done
      int incl = n0 + n1;
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      int off = nnext + incl - (n0 + n1);
      if (n0 && off < cap) { nxt[2 * off] = c0lo; nxt[2 * off + 1] = c0hi; }
      off += n0;
      if (n1 && off < cap) { nxt[2 * off] = c1lo; nxt[2 * off + 1] = c1hi; }
      nnext += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    int *t = cur; cur = nxt; nxt = t;
    ncur = nnext < cap ? nnext : cap;
  }
#undef XC
}

enum { HQ_TERM = 0, HQ_NORMHIT = 1, HQ_MULTIHIT = 2, HQ_REPEAT = 3, HQ_NOHIT = 4, HQ_NONSTDNT = 5 };
enum { HI_REVERSE = 1, HI_SORTED = 2, HI_RANK = 4 };

__global__ void __launch_bounds__(128)
seed_kernel(const Index ix0, const uint8_t *__restrict__ arena, const SeedArgs a) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 2 * a.nreads) return;
  const int rd = g >> 1, is_reverse = g & 1;
  Index ix = ix0;
  if (a.tab) { const IndexTab t = a.tab[a.read_tab[rd]]; ix.idx = t.idx; ix.pos = t.pos; ix.npos = t.npos; }
  const uint32_t qlen = a.read_len[rd];
  const uint8_t *read = arena + a.read_off[rd];
  const uint8_t *qual = a.qual ? a.qual + a.read_off[rd] : nullptr;
  const uint64_t slot = a.slot_off[rd] + (is_reverse ? qlen : 0u);
  uint32_t *posidx = a.posidx + slot, *nhits = a.nhits + slot, *qoffs = a.qoffs + slot;
  uint32_t *sortkey = a.sortkey + slot, *sidx = a.sidx + slot;
  uint8_t *qmask = a.qmask + slot, *qbuf = a.qbuf + slot;
  const int ktup = ix.wordlen, nskip = ix.nskip;
  smb_seed_info inf;
  inf.n_seeds = inf.seed_rank = inf.cover_deficit = inf.nhit_rank = inf.nhit_tot = inf.nhit_all = 0;
  inf.status = is_reverse ? HI_REVERSE : 0;
  inf.err = 0;
  if (qlen < (uint32_t)ktup) {
    inf.err = SMB_ERRCODE_SHORTSEQ;
    a.info[g] = inf;
    return;
  }
  // ---- collectHitInfo ----
  const uint8_t minqval = (uint8_t)(a.basq_thresh + 0x21);
  const int rc_addpos = (ktup - 1) << 1;
  const uint64_t wordmask = (ktup >= 32) ? ~0ull : ((1ull << (ktup << 1)) - 1ull);
  const uint32_t maxhit = a.is_short ? a.maxhit_per_tuple : 0u;
  uint64_t word = 0;
  long long h0 = -1, h1 = -2, h2 = -3, h3 = -4;  // initRepeatFilter (hashhit.c:342-346)
  uint32_t tuplectr = 0, seedctr = 0, non_std = 0;
  for (uint32_t s = 0; s < qlen; ++s) {
    const uint32_t c = __ldg(read + s);
    if ((c & 4u) || (qual && __ldg(qual + s) < minqval)) non_std = (uint32_t)ktup;
    else if (non_std) --non_std;
    if (is_reverse) word = (word >> 2) + ((uint64_t)((c ^ 3u) & 3u) << rc_addpos);
    else word = (word << 2) + (c & 3u);
    if (s + 1 < (uint32_t)ktup) continue;
    if (non_std) { qmask[tuplectr++] = HQ_NONSTDNT; continue; }
    const long long w = (long long)(word & wordmask);
    const bool rep = (w == h0) | (w == h1) | (w == h2) | (w == h3);
    h3 = h2; h2 = h1; h1 = h0; h0 = w;
    if (rep) { qmask[tuplectr++] = HQ_REPEAT; continue; }
    uint32_t px = 0;
    const uint32_t nh = lookup(ix, word, px);
    if (nh < 1) { qmask[tuplectr++] = HQ_NOHIT; continue; }
    if (maxhit > 0 && nh > maxhit) { qmask[tuplectr++] = HQ_MULTIHIT; continue; }
    sortkey[seedctr] = nh;
    qmask[tuplectr] = HQ_NORMHIT;
    posidx[seedctr] = px;
    nhits[seedctr] = nh;
    qoffs[seedctr] = tuplectr;
    sidx[seedctr] = seedctr;
    ++seedctr;
    ++tuplectr;
  }
  for (; tuplectr < qlen; ++tuplectr) qmask[tuplectr] = HQ_TERM;
  const uint32_t n_seeds = seedctr;
  inf.n_seeds = n_seeds;

  // ---- hashCollectHitInfoShort: sort + rank ----
  uint32_t *frame_cnt = nullptr;
  uint32_t fcnt[32];  // nskip <= 31 in practice (nskip is a uint8, k <= 20): counts per frame
  (void)frame_cnt;
  uint32_t *frame = a.frame + slot;  // rank lists of the frames, back to back
  uint32_t fstart[33];
  if (a.is_short) {
    if (n_seeds <= 1) {
      inf.status |= HI_SORTED;
      inf.seed_rank = n_seeds;
    } else if (nskip > 32) {
      inf.err = SMB_ERR_ARG;
    } else {
      int e = sort2((int)n_seeds, sortkey, sidx);
      if (e) inf.err = e;
      inf.status |= HI_SORTED;
      uint32_t mincover = 2u * (uint32_t)ktup + (uint32_t)nskip;
      uint32_t maxcover = qlen * 80u / 100u;
      if (maxcover < (uint32_t)(ktup + nskip)) maxcover = (uint32_t)(ktup + nskip);
      else if (maxcover > qlen - (uint32_t)nskip) maxcover = qlen - (uint32_t)nskip;
      if (mincover > maxcover) { mincover = 0; maxcover = qlen; }
      // getHitInfoMaxRank (hashhit.c:769-891)
      for (int f = 0; f < nskip; ++f) fcnt[f] = 0;
      for (uint32_t i = 0; i < n_seeds; ++i) fcnt[qoffs[sidx[i]] % (uint32_t)nskip]++;
      fstart[0] = 0;
      for (int f = 0; f < nskip; ++f) { fstart[f + 1] = fstart[f] + fcnt[f]; fcnt[f] = 0; }
      for (uint32_t i = 0; i < n_seeds; ++i) {
        const int f = (int)(qoffs[sidx[i]] % (uint32_t)nskip);
        frame[fstart[f] + fcnt[f]++] = i;  // the rank
      }
      // the reference reads sortkey[n_seeds] here (one past the end, hashhit.c:823); the
      // value never changes the result because the loop ends at i == n_seeds + 1
      uint32_t ntot = sortkey[0], i;
      for (i = 1; i <= n_seeds && ntot <= a.maxhit_total; ++i) ntot += (i < n_seeds) ? sortkey[i] : 0u;
      uint32_t n = i - 1, nmax = n;
      for (int f = 0; f < nskip; ++f) {
        const uint32_t imax = fcnt[f];
        const uint32_t *ixp = frame + fstart[f];
        uint32_t cover = 0;
        if (!imax) continue;
        for (uint32_t q = 0; q < qlen; ++q) qbuf[q] = 0;
        for (i = 0; i < imax && cover <= maxcover && (cover < mincover || ixp[i] <= n); ++i) {
          const uint32_t q0 = qoffs[sidx[ixp[i]]];
          for (uint32_t q = q0; q < q0 + (uint32_t)ktup - 1u; ++q)
            if (!qbuf[q]) { qbuf[q] = 1; ++cover; }
        }
        if (i > 0 && ixp[i - 1] > nmax) nmax = ixp[i - 1];
      }
      inf.seed_rank = (nmax < 3u) ? (3u < n_seeds ? 3u : n_seeds) : nmax;  // HITINFO_MINSEEDNUM
      inf.status |= HI_RANK;
    }
  }
  // ---- hashCalcHitInfoCoverDeficit (hashhit.c:1096-1169) ----
  if (inf.status & HI_RANK) {
    uint32_t d = qlen, maxc = 0;
    for (int f = 0; f < nskip; ++f) {
      const uint32_t imax = fcnt[f];
      const uint32_t *ixp = frame + fstart[f];
      uint32_t cover = 0;
      if (!imax) continue;
      for (uint32_t q = 0; q < qlen; ++q) qbuf[q] = 0;
      for (uint32_t i = 0; i < imax && ixp[i] < inf.seed_rank; ++i) {
        const uint32_t q0 = qoffs[sidx[ixp[i]]];
        for (uint32_t q = q0; q < q0 + (uint32_t)ktup; ++q)
          if (!qbuf[q]) { qbuf[q] = 1; ++cover; }
      }
      if (cover < d) d = cover;
      if (cover > maxc) maxc = cover;
    }
    inf.cover_deficit = maxc - d + 1u;
  } else {
    uint32_t k = (uint32_t)(ktup / nskip), deficit = 0;
    if (k > 0) --k;
    k &= 0xffu;
    for (int f = 0; f < nskip; ++f) {
      uint32_t d = 0, ctr = 0;
      for (uint32_t i = (uint32_t)f; i < qlen; i += (uint32_t)nskip) {
        if (qmask[i] == HQ_NORMHIT) ctr = k;
        else if (ctr) --ctr;
        else d += (uint32_t)nskip;
      }
      if (d > deficit) deficit = d;
    }
    inf.cover_deficit = deficit;
  }
  // ---- hashHitInfoCalcHitNumbers (:1200) / hashCalcHitInfoNumberOfHits (:1171) ----
  {
    const uint32_t ns = inf.seed_rank > 0 ? inf.seed_rank : n_seeds;
    uint32_t nr = 0, i, hnum = 0;
    for (i = 0; i < ns && i < n_seeds; ++i) nr += sortkey[i];
    inf.nhit_rank = nr;
    for (; i < n_seeds; ++i) nr += sortkey[i];
    inf.nhit_tot = nr;
    for (i = 0; i < n_seeds; ++i)
      if (a.maxhit_per_tuple < 1u || sortkey[i] <= a.maxhit_per_tuple) hnum += sortkey[i];
    inf.nhit_all = hnum;
  }
  a.info[g] = inf;
}


// ------------------------------------------------------------------------------------
// seed_warp_kernel: the same function as seed_kernel (hashCollectHitInfoShort of one read x
// strand), ONE WARP per read x strand.  A block of a few thousand reads gives only ~16 k
// read-strands: one thread each leaves the GPU at 6 % occupancy and serialises 32 divergent
// quicksorts per warp (profiles/r1_ncu_full_k1_*).  Here
//   * the k-mer words of all read offsets are formed by the lanes in parallel from the read
//     staged in shared memory; validity (non-standard base / low quality inside the word),
//     the 4-deep repeat filter (previous four VALID words, hashhit.c:342-346, :584-600) and the
//     seed compaction are warp ballots / prefix sums - same decisions, order-preserving;
//   * the index probes of 32 words are in flight at once (the dependent idx -> wordidx ->
//     posidx chains of different words are independent);
//   * the reference's unstable quicksort (sort.c:233-330) runs unchanged in lane 0 on shared
//     memory (its exchange sequence decides the tie order, so it is not parallelised);
//   * the per-frame coverage scans of getHitInfoMaxRank / hashCalcHitInfoCoverDeficit run one
//     frame per lane on bit masks.
// Reads longer than the shared-memory staging or nskip > 32 stay with seed_kernel.
// ------------------------------------------------------------------------------------
constexpr int SEEDW_WARPS = 4;

struct SeedWarpLayout {   // per-warp shared memory carve-up for reads of at most qmax bases
  int qmax;
  __host__ __device__ size_t words_off() const { return 0; }                       // u64[qmax]
  __host__ __device__ size_t sortkey_off() const { return (size_t)qmax * 8; }      // u32[qmax+1]
  __host__ __device__ size_t sidx_off() const { return sortkey_off() + ((size_t)qmax + 1) * 4; }  // u32[qmax]
  __host__ __device__ size_t cov_off() const { return sidx_off() + (size_t)qmax * 4; }   // u32[32 * covw]
  __host__ __device__ int covw() const { return (qmax + 31) / 32; }
  __host__ __device__ size_t qoffs_off() const { return cov_off() + (size_t)32 * covw() * 4; }  // u16[qmax]
  __host__ __device__ size_t frame_off() const { return qoffs_off() + (size_t)qmax * 2; }       // u16[qmax]
  __host__ __device__ size_t vt_off() const { return frame_off() + (size_t)qmax * 2; }          // u16[qmax]
  __host__ __device__ size_t codes_off() const { return vt_off() + (size_t)qmax * 2; }          // u8[qmax]
  __host__ __device__ size_t qmask_off() const { return codes_off() + (size_t)qmax; }           // u8[qmax]
  __host__ __device__ size_t bytes() const { return (qmask_off() + (size_t)qmax + 15) & ~(size_t)15; }
};

// marks read positions [q0, q0 + len) in a bit mask and returns how many of them were not marked
// before (the per-base loops of hashhit.c:838-850 and :1120-1131, a word at a time)
__device__ __forceinline__ uint32_t cov_add(uint32_t *cov, uint32_t q0, uint32_t len) {
  uint32_t added = 0, w = q0 >> 5, sft = q0 & 31u;
  while (len) {
    const uint32_t take = min(len, 32u - sft);
    const uint32_t m = (take >= 32u ? 0xffffffffu : ((1u << take) - 1u)) << sft;
    const uint32_t old = cov[w];
    added += __popc(m & ~old);
    cov[w] = old | m;
    len -= take;
    sft = 0;
    ++w;
  }
  return added;
}

__global__ void __launch_bounds__(SEEDW_WARPS * 32)
seed_warp_kernel(const Index ix0, const uint8_t *__restrict__ arena, const SeedArgs a, const SeedWarpLayout lay) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * SEEDW_WARPS + (threadIdx.x >> 5);
  if (g >= 2 * a.nreads) return;
  unsigned char *base = s_raw + (size_t)(threadIdx.x >> 5) * lay.bytes();
  unsigned long long *s_word = (unsigned long long *)(base + lay.words_off());
  uint32_t *s_key = (uint32_t *)(base + lay.sortkey_off());
  uint32_t *s_sidx = (uint32_t *)(base + lay.sidx_off());
  uint32_t *s_cov = (uint32_t *)(base + lay.cov_off());
  unsigned short *s_qoffs = (unsigned short *)(base + lay.qoffs_off());
  unsigned short *s_frame = (unsigned short *)(base + lay.frame_off());
  unsigned short *s_vt = (unsigned short *)(base + lay.vt_off());
  uint8_t *s_code = base + lay.codes_off();
  uint8_t *s_qmask = base + lay.qmask_off();

  const int rd = g >> 1, is_reverse = g & 1;
  Index ix = ix0;
  if (a.tab) { const IndexTab t = a.tab[a.read_tab[rd]]; ix.idx = t.idx; ix.pos = t.pos; ix.npos = t.npos; }
  const uint32_t qlen = a.read_len[rd];
  const uint8_t *read = arena + a.read_off[rd];
  const uint8_t *qual = a.qual ? a.qual + a.read_off[rd] : nullptr;
  const uint64_t slot = a.slot_off[rd] + (is_reverse ? qlen : 0u);
  uint32_t *g_posidx = a.posidx + slot, *g_nhits = a.nhits + slot, *g_qoffs = a.qoffs + slot;
  uint32_t *g_sortkey = a.sortkey + slot, *g_sidx = a.sidx + slot;
  uint8_t *g_qmask = a.qmask + slot;
  const int ktup = ix.wordlen, nskip = ix.nskip;
  smb_seed_info inf;
  inf.n_seeds = inf.seed_rank = inf.cover_deficit = inf.nhit_rank = inf.nhit_tot = inf.nhit_all = 0;
  inf.status = is_reverse ? HI_REVERSE : 0;
  inf.err = 0;
  if (qlen < (uint32_t)ktup) {
    inf.err = SMB_ERRCODE_SHORTSEQ;
    if (lane == 0) a.info[g] = inf;
    return;
  }
  // ---- stage the read: code in bits 0-2, bit 7 = base that invalidates its k-mers ----
  const uint8_t minqval = (uint8_t)(a.basq_thresh + 0x21);
  for (uint32_t s = lane; s < qlen; s += 32) {
    const uint32_t c = __ldg(read + s);
    const bool bad = (c & 4u) || (qual && __ldg(qual + s) < minqval);
    s_code[s] = (uint8_t)((c & 7u) | (bad ? 0x80u : 0u));
  }
  __syncwarp();
  // ---- k-mer word of every read offset (collectHitInfo, hashhit.c:559-587) ----
  const uint32_t ntup = qlen - (uint32_t)ktup + 1u;
  for (uint32_t t = lane; t < ntup; t += 32) {
    unsigned long long w = 0;
    uint32_t bad = 0;
    if (is_reverse) {
      for (int m = 0; m < ktup; ++m) {
        const uint32_t c = s_code[t + m];
        bad |= c;
        w |= (unsigned long long)((c ^ 3u) & 3u) << (2 * m);
      }
    } else {
      for (int m = 0; m < ktup; ++m) {
        const uint32_t c = s_code[t + m];
        bad |= c;
        w = (w << 2) | (c & 3u);
      }
    }
    s_word[t] = (bad & 0x80u) ? ~0ull : w;   // 2k <= 62 bits: ~0 is never a word
  }
  __syncwarp();
  // ---- compaction of the valid words, in read order ----
  uint32_t nvalid = 0;
  for (uint32_t t0 = 0; t0 < ntup; t0 += 32) {
    const uint32_t t = t0 + lane;
    const bool v = t < ntup && s_word[t] != ~0ull;
    const unsigned m = __ballot_sync(FULL, v);
    if (v) s_vt[nvalid + __popc(m & ((1u << lane) - 1u))] = (unsigned short)t;
    if (t < ntup && !v) s_qmask[t] = HQ_NONSTDNT;
    nvalid += __popc(m);
  }
  for (uint32_t t = ntup + lane; t < qlen; t += 32) s_qmask[t] = HQ_TERM;
  __syncwarp();
  // ---- repeat filter, index probes, seed compaction: 32 valid words per round ----
  const uint32_t maxhit = a.maxhit_per_tuple;   // is_short
  uint32_t n_seeds = 0;
  for (uint32_t v0 = 0; v0 < nvalid; v0 += 32) {
    const uint32_t vi = v0 + lane;
    bool seed = false;
    uint32_t nh = 0, px = 0, t = 0;
    if (vi < nvalid) {
      t = s_vt[vi];
      const unsigned long long w = s_word[t];
      bool rep = false;
      for (uint32_t j = 1; j <= 4u && j <= vi; ++j) rep |= (s_word[s_vt[vi - j]] == w);
      uint8_t code;
      if (rep) code = HQ_REPEAT;
      else {
        nh = lookup(ix, w, px);
        if (nh < 1) code = HQ_NOHIT;
        else if (maxhit > 0 && nh > maxhit) code = HQ_MULTIHIT;
        else { code = HQ_NORMHIT; seed = true; }
      }
      s_qmask[t] = code;
    }
    const unsigned m = __ballot_sync(FULL, seed);
    if (seed) {
      const uint32_t k = n_seeds + __popc(m & ((1u << lane) - 1u));
      s_key[k] = nh;
      s_sidx[k] = k;
      s_qoffs[k] = (unsigned short)t;
      g_posidx[k] = px;
      g_nhits[k] = nh;
      g_qoffs[k] = t;
    }
    n_seeds += __popc(m);
  }
  __syncwarp();
  for (uint32_t t = lane; t < qlen; t += 32) g_qmask[t] = s_qmask[t];
  inf.n_seeds = n_seeds;

  // ---- hashCollectHitInfoShort: sort + rank ----
  int fcnt = 0, fstart = 0;   // lane f: seeds in frame f, start of its rank list
  if (n_seeds <= 1) {
    inf.status |= HI_SORTED;
    inf.seed_rank = n_seeds;
  } else {
    // (s_word is free by now: it serves as the range lists; a part has at least two elements, and
    // the parts of a level are disjoint: never more than n / 2 <= qmax / 2 of them)
    sort2_warp((int)n_seeds, s_key, s_sidx, (int *)s_word, lay.qmax / 2, lane);
    inf.status |= HI_SORTED;
    __syncwarp();
    uint32_t mincover = 2u * (uint32_t)ktup + (uint32_t)nskip;
    uint32_t maxcover = qlen * 80u / 100u;
    if (maxcover < (uint32_t)(ktup + nskip)) maxcover = (uint32_t)(ktup + nskip);
    else if (maxcover > qlen - (uint32_t)nskip) maxcover = qlen - (uint32_t)nskip;
    if (mincover > maxcover) { mincover = 0; maxcover = qlen; }
    // getHitInfoMaxRank (hashhit.c:769-891): rank lists of the frames, one frame per lane
    // (frame of every seed once, by all lanes; s_word is free again after the sort)
    uint8_t *const s_fr = (uint8_t *)s_word;
    for (uint32_t i = lane; i < n_seeds; i += 32) s_fr[i] = (uint8_t)(s_qoffs[s_sidx[i]] % (uint32_t)nskip);
    __syncwarp();
    if (lane < nskip)
      for (uint32_t i = 0; i < n_seeds; ++i) fcnt += s_fr[i] == (uint8_t)lane;
    int incl = fcnt;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    fstart = incl - fcnt;
    if (lane < nskip) {
      int c = 0;
      for (uint32_t i = 0; i < n_seeds; ++i)
        if (s_fr[i] == (uint8_t)lane) s_frame[fstart + c++] = (unsigned short)i;
    }
    // seeds whose hits sum up to at most maxhit_total (the reference reads one element past the
    // end here, hashhit.c:823; its value cannot change the result)
    uint32_t n = 0;
    if (lane == 0) {
      uint32_t ntot = s_key[0], i;
      for (i = 1; i <= n_seeds && ntot <= a.maxhit_total; ++i) ntot += (i < n_seeds) ? s_key[i] : 0u;
      n = i - 1;
    }
    n = __shfl_sync(FULL, n, 0);
    __syncwarp();
    uint32_t nmax = n;
    const int covw = lay.covw();
    uint32_t *cov = s_cov + lane * covw;
    if (lane < nskip && fcnt > 0) {
      for (int wq = 0; wq < covw; ++wq) cov[wq] = 0;
      uint32_t cover = 0;
      int i = 0;
      for (; i < fcnt && cover <= maxcover && (cover < mincover || s_frame[fstart + i] <= n); ++i) {
        const uint32_t q0 = s_qoffs[s_sidx[s_frame[fstart + i]]];
        cover += cov_add(cov, q0, (uint32_t)ktup - 1u);
      }
      if (i > 0 && s_frame[fstart + i - 1] > nmax) nmax = s_frame[fstart + i - 1];
    }
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));
    inf.seed_rank = (nmax < 3u) ? (3u < n_seeds ? 3u : n_seeds) : nmax;  // HITINFO_MINSEEDNUM
    inf.status |= HI_RANK;
  }
  // ---- hashCalcHitInfoCoverDeficit (hashhit.c:1096-1169) ----
  if (inf.status & HI_RANK) {
    uint32_t dmin = 0xffffffffu, maxc = 0;
    const int covw = lay.covw();
    uint32_t *cov = s_cov + lane * covw;
    if (lane < nskip && fcnt > 0) {
      for (int wq = 0; wq < covw; ++wq) cov[wq] = 0;
      uint32_t cover = 0;
      for (int i = 0; i < fcnt && s_frame[fstart + i] < inf.seed_rank; ++i) {
        const uint32_t q0 = s_qoffs[s_sidx[s_frame[fstart + i]]];
        cover += cov_add(cov, q0, (uint32_t)ktup);
      }
      dmin = cover;
      maxc = cover;
    }
    for (int o = 16; o > 0; o >>= 1) {
      dmin = min(dmin, __shfl_xor_sync(FULL, dmin, o));
      maxc = max(maxc, __shfl_xor_sync(FULL, maxc, o));
    }
    if (dmin > qlen) dmin = qlen;
    inf.cover_deficit = maxc - dmin + 1u;
  } else {
    uint32_t k = (uint32_t)(ktup / nskip), deficit = 0;
    if (k > 0) --k;
    k &= 0xffu;
    if (lane < nskip) {
      uint32_t d = 0, ctr = 0;
      for (uint32_t i = (uint32_t)lane; i < qlen; i += (uint32_t)nskip) {
        if (s_qmask[i] == HQ_NORMHIT) ctr = k;
        else if (ctr) --ctr;
        else d += (uint32_t)nskip;
      }
      deficit = d;
    }
    for (int o = 16; o > 0; o >>= 1) deficit = max(deficit, __shfl_xor_sync(FULL, deficit, o));
    inf.cover_deficit = deficit;
  }
  // ---- hashHitInfoCalcHitNumbers (:1200) / hashCalcHitInfoNumberOfHits (:1171), outputs ----
  {
    const uint32_t ns = inf.seed_rank > 0 ? inf.seed_rank : n_seeds;
    uint32_t nr = 0, nt = 0, hnum = 0;
    for (uint32_t i = lane; i < n_seeds; i += 32) {
      const uint32_t kx = s_key[i];
      if (i < ns) nr += kx;
      nt += kx;
      if (a.maxhit_per_tuple < 1u || kx <= a.maxhit_per_tuple) hnum += kx;
      g_sortkey[i] = kx;
      g_sidx[i] = s_sidx[i];
    }
    for (int o = 16; o > 0; o >>= 1) {
      nr += __shfl_xor_sync(FULL, nr, o);
      nt += __shfl_xor_sync(FULL, nt, o);
      hnum += __shfl_xor_sync(FULL, hnum, o);
    }
    inf.nhit_rank = nr;
    inf.nhit_tot = nt;
    inf.nhit_all = hnum;
  }
  if (lane == 0) a.info[g] = inf;
}

cudaError_t launch_seed(const Index &ix, const uint8_t *arena, const SeedArgs &a, cudaStream_t st,
                        int *nlaunch) {
  const int n = 2 * a.nreads;
  if (n <= 0) return cudaSuccess;
  // warp per read x strand when the reads fit the shared-memory staging (a.maxlen), the
  // ranked short mode is asked for and a frame fits a lane
  SeedWarpLayout lay{(int)((a.maxlen + 63u) & ~63u)};
  const size_t smem = lay.bytes() * SEEDW_WARPS;
  if (a.is_short && a.maxlen > 0 && a.maxlen <= 2048u && ix.nskip <= 32 && ix.wordlen <= 31 && smem <= 200 * 1024) {
    static std::atomic<unsigned long long> smem_done{0};
    const cudaError_t ea = ensure_dyn_smem(seed_warp_kernel, 200 * 1024, smem_done);
    if (ea != cudaSuccess) return ea;
    seed_warp_kernel<<<(n + SEEDW_WARPS - 1) / SEEDW_WARPS, SEEDW_WARPS * 32, smem, st>>>(ix, arena, a, lay);
  } else {
    seed_kernel<<<(n + 127) / 128, 128, 0, st>>>(ix, arena, a);
  }
  ++*nlaunch;
  return cudaGetLastError();
}

}  // namespace smb

// ------------------------------------------------------------------------------------
// Hit lists: fillHitListFromHitInfoSegment (hashhit.c:1416-1546, unfiltered branch) with the
// retry loop of hashCollectHitsForSegment (:1739-1741), and hashCollectHitsUsingCutoff
// (:1593-1689).  Two passes: COUNT (sizes every list, resolves the halving of the per-seed
// cut-off) and FILL (writes the packed hits and sorts them ascending).
// The reference's resume pointer SEED.cix only accelerates the scan for the first position
// >= lo (positions of a word are ascending); a binary search gives the same element.
// ------------------------------------------------------------------------------------
namespace smb {

__device__ __forceinline__ uint32_t lower_bound_pos(const uint32_t *p, uint32_t n, uint32_t v) {
  uint32_t a = 0, b = n;
  while (a < b) {
    const uint32_t m = (a + b) >> 1;
    if (__ldg(p + m) < v) a = m + 1; else b = m;
  }
  return a;
}

__device__ __forceinline__ uint64_t pack_hit(bool is_reverse, uint32_t pos, uint32_t q, uint32_t nskip) {
  // SET_NEXT_SHIFT (hashhit.c:283-288), HASHHIT_HALFBIT = 31
  if (is_reverse) return (((uint64_t)pos + q / nskip) << 31) + q;
  return ((((uint64_t)pos | (1ull << 32)) - q / nskip) << 31) + q;
}

// one pass over the seeds of a request; FILL writes hits, otherwise only counts
template <bool FILL>
__device__ int hits_segment_pass(const Index &ix, const HitArgs &a, const smb_hit_req &rq, uint64_t slot,
                                 uint32_t n_seeds_tot, uint32_t seed_rank, uint32_t maxhit,
                                 uint32_t pos_lo, uint32_t pos_hi, uint64_t *out, uint32_t &total) {
  const uint32_t *posidx = a.seed.posidx + slot, *qoffs = a.seed.qoffs + slot;
  const uint32_t *sortkey = a.seed.sortkey + slot, *sidx = a.seed.sidx + slot;
  uint8_t *qmask = a.seed.qmask + slot;
  const bool use_short = rq.use_short != 0;
  const uint32_t n_seeds = (use_short && seed_rank > 0) ? seed_rank : n_seeds_tot;
  const bool is_reverse = rq.strand != 0;
  uint32_t nh_tot = 0;
  for (uint32_t n = 0; n < n_seeds; ++n) {
    const uint32_t sd = use_short ? sidx[n] : n;
    if (maxhit > 0 && sortkey[n] > maxhit) {
      if (FILL) qmask[qoffs[sd]] = HQ_MULTIHIT;
      continue;
    }
    const uint32_t *posp;
    const uint32_t nhits = fetch_positions(ix, posidx[sd], posp);
    if (!posp || nhits == 0) continue;
    const uint32_t first = lower_bound_pos(posp, nhits, pos_lo);
    if (first >= nhits) continue;  // all positions below the segment
    const uint32_t nh = nhits - first;
    if (nh_tot + nh > a.nhits_alloc) {
      if (maxhit > 0) { total = nh_tot; return SMB_ERRCODE_ALLOCBOUNDARY; }
      if (FILL) qmask[qoffs[sd]] = HQ_MULTIHIT;
      continue;
    }
    const uint32_t q = qoffs[sd];
    uint32_t i = 0;
    for (; i < nh; ++i) {
      const uint32_t p = __ldg(posp + first + i);
      if (p >= pos_hi) break;
      if (FILL) out[nh_tot + i] = pack_hit(is_reverse, p, q, (uint32_t)ix.nskip);
    }
    nh_tot += i;
  }
  total = nh_tot;
  return 0;
}

// in-place ascending sort of one list by a single thread (keys are unique)
__device__ void sort_u64(uint64_t *a, int n) {
  if (n < 2) return;
  if (n <= 24) {
    for (int j = 1; j < n; ++j) {
      const uint64_t k = a[j];
      int i = j - 1;
      for (; i >= 0 && a[i] > k; --i) a[i + 1] = a[i];
      a[i + 1] = k;
    }
    return;
  }
  // heapsort: O(n log n) worst case, no stack
  for (int start = n / 2 - 1; start >= 0; --start) {
    int root = start;
    const uint64_t v = a[root];
    for (;;) {
      int child = 2 * root + 1;
      if (child >= n) break;
      if (child + 1 < n && a[child + 1] > a[child]) ++child;
      if (a[child] <= v) break;
      a[root] = a[child];
      root = child;
    }
    a[root] = v;
  }
  for (int end = n - 1; end > 0; --end) {
    const uint64_t v = a[end];
    a[end] = a[0];
    int root = 0;
    for (;;) {
      int child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && a[child + 1] > a[child]) ++child;
      if (a[child] <= v) break;
      a[root] = a[child];
      root = child;
    }
    a[root] = v;
  }
}

// ------------------------------------------------------------------------------------
// hits_warp_kernel: ONE WARP PER REQUEST (read x strand x reference segment).  The lanes take
// the seeds of the request in parallel (position-range search of each seed is independent),
// list offsets come from a warp prefix sum, and the final ascending sort of the (unique)
// packed hits is a bitonic sort in shared memory.  The reference's sequential overflow
// handling (list longer than nhits_alloc: halve the per-seed cut-off and retry,
// hashhit.c:1730-1741) cannot trigger when even the sum over ALL seeds of the hits at or
// beyond the segment start fits nhits_alloc - the test that selects this path; otherwise
// lane 0 runs the sequential code (hits_segment_pass): COUNT resolves the halving of the
// per-seed cut-off, FILL replays the failed attempts (they mark MULTIHIT seeds in the read's
// qmask as a side effect) and the final one.
// ------------------------------------------------------------------------------------
constexpr int HITW_WARPS = 4;
constexpr int HITW_SORTCAP = 512;   // hits sorted in shared memory (larger lists: heapsort by lane 0)

__device__ __forceinline__ uint32_t upper_bound_pos(const uint32_t *p, uint32_t n, uint32_t from, uint32_t v) {
  uint32_t a = from, b = n;   // first index >= from with p[index] >= v
  while (a < b) {
    const uint32_t m = (a + b) >> 1;
    if (__ldg(p + m) < v) a = m + 1; else b = m;
  }
  return a;
}

template <bool FILL>
__global__ void __launch_bounds__(HITW_WARPS * 32) hits_warp_kernel(const Index ix0, const HitArgs a) {
  __shared__ unsigned long long s_sort[HITW_WARPS][HITW_SORTCAP];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * HITW_WARPS + (threadIdx.x >> 5);
  if (g >= a.nreq) return;
  if (a.req_skip && a.req_skip[g]) {   // no seed position of this read x strand lies in the request's sequence (block.cu)
    if (!FILL && lane == 0) { a.count[g] = 0; a.maxhit_used[g] = 0x80000000u; a.errs[g] = 0; }
    return;
  }
  unsigned long long *srt = s_sort[threadIdx.x >> 5];
  const smb_hit_req rq = a.req[g];
  const uint32_t rd = rq.read;
  Index ix = ix0;
  if (a.seed.tab) { const IndexTab t = a.seed.tab[a.seed.read_tab[rd]]; ix.idx = t.idx; ix.pos = t.pos; ix.npos = t.npos; }
  const uint32_t qlen = a.seed.read_len[rd];
  const uint64_t slot = a.seed.slot_off[rd] + (rq.strand ? qlen : 0u);
  const smb_seed_info inf = a.seed.info[2 * rd + (rq.strand ? 1 : 0)];
  uint64_t lo = rq.lo / (uint64_t)ix.nskip, hi = rq.hi / (uint64_t)ix.nskip;
  if (inf.err || lo > 0xFFFFFFFFull) {
    if (!FILL && lane == 0) { a.count[g] = 0; a.maxhit_used[g] = 0; a.errs[g] = inf.err ? inf.err : SMB_ERRCODE_ARGRANGE; }
    return;
  }
  if (hi > 0xFFFFFFFFull) hi = 0xFFFFFFFFull;
  const uint32_t pos_lo = (uint32_t)lo, pos_hi = (uint32_t)hi;
  const uint32_t *posidx = a.seed.posidx + slot, *qoffs = a.seed.qoffs + slot;
  const uint32_t *sortkey = a.seed.sortkey + slot, *sidx = a.seed.sidx + slot;
  uint8_t *qmask = a.seed.qmask + slot;
  const bool use_short = rq.use_short != 0;
  const uint32_t n_seeds = (use_short && inf.seed_rank > 0) ? inf.seed_rank : inf.n_seeds;
  const bool is_reverse = rq.strand != 0;
  const uint32_t maxhit = rq.nhit_max;
  constexpr uint32_t FASTFLAG = 0x80000000u;

  if (rq.use_short == 2) {   // hashCollectHitsUsingCutoff (hashhit.c:1593-1689): whole set, rank order
    const uint32_t ns = inf.seed_rank ? inf.seed_rank : inf.n_seeds;
    uint32_t nhits_max = rq.nhits_max;
    if (!nhits_max) {
      double t = qlen > 1 ? (double)qlen * log((double)qlen) * 32.0 : 0.0;
      nhits_max = t > 2147483647.0 ? 2147483647u : (t < 8192.0 ? 8192u : (uint32_t)t);
    }
    uint8_t *lq = a.list_qmask ? a.list_qmask + a.qmask_off[g] : nullptr;
    if (!FILL) {
      unsigned long long tot = 0;
      for (uint32_t n = lane; n < ns; n += 32) {
        const uint32_t nh = sortkey[n];
        if (nh >= 1 && !(maxhit > 0 && nh > maxhit)) tot += nh;
      }
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
      if (tot <= (unsigned long long)nhits_max) {   // the ceiling cannot be reached: first attempt, every seed
        if (lane == 0) { a.count[g] = (uint32_t)tot; a.maxhit_used[g] = maxhit | FASTFLAG; a.errs[g] = 0; }
        return;
      }
      if (lane == 0) {   // the reference's retry loop: halve the cut-off while the ceiling is hit
        uint32_t mh = maxhit, used, nh_tot;
        bool ceiling;
        do {
          ceiling = false;
          used = mh;
          nh_tot = 0;
          for (uint32_t n = 0; n < ns; ++n) {
            const uint32_t nh = sortkey[n];
            if (nh < 1) continue;
            if (mh > 0 && nh > mh) continue;
            if (nh_tot + nh > nhits_max) { ceiling = true; break; }
            nh_tot += nh;
          }
          mh /= 2;
        } while (ceiling && mh > 16u);   // MINHIT_PER_TUPLE
        a.count[g] = nh_tot;
        a.maxhit_used[g] = used & ~FASTFLAG;
        a.errs[g] = 0;
      }
      return;
    }
    // FILL: the final attempt (cut-off a.maxhit_used), seeds in rank order up to the ceiling
    const uint32_t usedflag = a.maxhit_used[g];
    const uint32_t mh = usedflag & ~FASTFLAG;
    uint64_t *out = a.sqdat + a.offset[g];
    const uint32_t count = a.count[g];
    if (lq) for (uint32_t q = lane; q < qlen; q += 32) lq[q] = HQ_NOHIT;
    __syncwarp();
    uint32_t base = 0;
    bool stop = false;
    for (uint32_t n0 = 0; n0 < ns && !stop; n0 += 32) {
      const uint32_t n = n0 + lane;
      uint32_t nh = 0, q = 0, sd = 0;
      bool multi = false;
      if (n < ns) {
        nh = sortkey[n];
        sd = sidx[n];
        q = qoffs[sd];
        if (nh >= 1 && mh > 0 && nh > mh) { multi = true; nh = 0; }
      }
      uint32_t incl = nh;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      // ceiling: the first seed whose hits would not fit ends the list (nothing after it is looked at)
      const bool over = nh > 0 && base + incl > nhits_max;
      const unsigned overm = __ballot_sync(FULL, over);
      const int firstover = overm ? __ffs(overm) - 1 : 32;
      if (lane < firstover) {
        if (multi) { if (lq) lq[q] = HQ_MULTIHIT; }
        else if (nh > 0) {
          const uint32_t *posp;
          const uint32_t nhits = fetch_positions(ix, posidx[sd], posp);
          const uint32_t off = base + incl - nh;
          if (lq) lq[q] = HQ_NORMHIT;
          for (uint32_t i = 0; i < nh && i < nhits && off + i < count; ++i)
            out[off + i] = pack_hit(is_reverse, __ldg(posp + i), q, (uint32_t)ix.nskip);
        }
      }
      if (overm) stop = true;
      base += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    if (count <= (uint32_t)HITW_SORTCAP) {
      for (uint32_t i = lane; i < count; i += 32) srt[i] = out[i];
      int np2 = 1;
      while (np2 < (int)count) np2 <<= 1;
      for (int i = (int)count + lane; i < np2; i += 32) srt[i] = ~0ull;
      __syncwarp();
      for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < np2; i += 32) {
            const int l = i ^ j;
            if (l > i) {
              const unsigned long long x = srt[i], y = srt[l];
              const bool up = (i & k) == 0;
              if ((x > y) == up) { srt[i] = y; srt[l] = x; }
            }
          }
          __syncwarp();
        }
      for (uint32_t i = lane; i < count; i += 32) out[i] = srt[i];
    } else if (lane == 0) {
      sort_u64(out, (int)count);
    }
    return;
  }
  if (FILL && a.list_qmask)   // segment lists never mark seeds: all NOHIT (hashhit.c:1224-1231)
    for (uint32_t q = lane; q < qlen; q += 32) a.list_qmask[a.qmask_off[g] + q] = HQ_NOHIT;

  bool fast;
  if (!FILL) {
    // can the sequential overflow logic trigger at all?  bound: hits at or beyond the segment start
    unsigned long long bound = 0;
    uint32_t total = 0;
    for (uint32_t n = lane; n < n_seeds; n += 32) {
      if (maxhit > 0 && sortkey[n] > maxhit) continue;
      const uint32_t sd = use_short ? sidx[n] : n;
      const uint32_t *posp;
      const uint32_t nhits = fetch_positions(ix, posidx[sd], posp);
      if (!posp || nhits == 0) continue;
      const uint32_t first = lower_bound_pos(posp, nhits, pos_lo);
      bound += nhits - first;
      total += upper_bound_pos(posp, nhits, first, pos_hi) - first;
    }
    for (int o = 16; o > 0; o >>= 1) {
      bound += __shfl_xor_sync(FULL, bound, o);
      total += __shfl_xor_sync(FULL, total, o);
    }
    fast = bound <= (unsigned long long)a.nhits_alloc;
    if (fast) {
      if (lane == 0) { a.count[g] = total; a.maxhit_used[g] = maxhit | FASTFLAG; a.errs[g] = 0; }
      return;
    }
    if (lane == 0) {   // sequential path: hashCollectHitsForSegment retry loop (hashhit.c:1730-1741)
      uint32_t mh = maxhit, tot = 0, used = 0;
      int err;
      do {
        used = mh;
        err = hits_segment_pass<false>(ix, a, rq, slot, inf.n_seeds, inf.seed_rank, mh, pos_lo, pos_hi, nullptr, tot);
        mh /= 2;
      } while (err == SMB_ERRCODE_ALLOCBOUNDARY && mh > 16u);
      a.errs[g] = (err == SMB_ERRCODE_ALLOCBOUNDARY) ? SMB_ERRCODE_ALLOCBOUNDARY : 0;
      a.count[g] = tot;
      a.maxhit_used[g] = used & ~FASTFLAG;
    }
    return;
  }

  // ---- FILL ----
  const uint32_t usedflag = a.maxhit_used[g];
  uint64_t *out = a.sqdat + a.offset[g];
  const int count = (int)a.count[g];
  fast = (usedflag & FASTFLAG) != 0;
  if (!fast) {
    if (lane == 0) {   // sequential path
      uint32_t mh = maxhit, total = 0;
      const uint32_t used = usedflag;
      while (mh != used && mh > 16u) {
        uint32_t t2 = 0;
        hits_segment_pass<true>(ix, a, rq, slot, inf.n_seeds, inf.seed_rank, mh, pos_lo, pos_hi, out, t2);
        mh /= 2;
      }
      hits_segment_pass<true>(ix, a, rq, slot, inf.n_seeds, inf.seed_rank, used, pos_lo, pos_hi, out, total);
      sort_u64(out, count);
    }
    return;
  }
  const bool in_smem = count <= HITW_SORTCAP;
  uint32_t base = 0;
  for (uint32_t n0 = 0; n0 < n_seeds; n0 += 32) {
    const uint32_t n = n0 + lane;
    uint32_t cnt = 0, first = 0, q = 0;
    const uint32_t *posp = nullptr;
    if (n < n_seeds) {
      const uint32_t sd = use_short ? sidx[n] : n;
      q = qoffs[sd];
      if (maxhit > 0 && sortkey[n] > maxhit) {
        qmask[q] = HQ_MULTIHIT;
      } else {
        const uint32_t nhits = fetch_positions(ix, posidx[sd], posp);
        if (posp && nhits) {
          first = lower_bound_pos(posp, nhits, pos_lo);
          cnt = upper_bound_pos(posp, nhits, first, pos_hi) - first;
        }
      }
    }
    uint32_t incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t off = base + incl - cnt;
    for (uint32_t i = 0; i < cnt; ++i) {
      const unsigned long long h = pack_hit(is_reverse, __ldg(posp + first + i), q, (uint32_t)ix.nskip);
      if (in_smem) srt[off + i] = h; else out[off + i] = h;
    }
    base += __shfl_sync(FULL, incl, 31);
  }
  __syncwarp();
  if (!in_smem) {
    if (lane == 0) sort_u64(out, count);
    return;
  }
  // bitonic sort of `count` keys padded to a power of two
  int np2 = 1;
  while (np2 < count) np2 <<= 1;
  for (int i = count + lane; i < np2; i += 32) srt[i] = ~0ull;
  __syncwarp();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < np2; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = srt[i], y = srt[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { srt[i] = y; srt[l] = x; }
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < count; i += 32) out[i] = srt[i];
}

cudaError_t launch_hits(const Index &ix, const HitArgs &a, bool fill, cudaStream_t st, int *nlaunch) {
  if (a.nreq <= 0) return cudaSuccess;
  const int grid = (a.nreq + HITW_WARPS - 1) / HITW_WARPS;
  if (fill) hits_warp_kernel<true><<<grid, HITW_WARPS * 32, 0, st>>>(ix, a);
  else hits_warp_kernel<false><<<grid, HITW_WARPS * 32, 0, st>>>(ix, a);
  ++*nlaunch;
  return cudaGetLastError();
}

cudaError_t warm_seed() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, seed_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, seed_warp_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, hits_warp_kernel<true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, hits_warp_kernel<false>);
  return e;
}

}  // namespace smb
