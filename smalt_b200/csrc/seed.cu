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

// hashTableFetchHitPositions
__device__ __forceinline__ uint32_t fetch_positions(const Index &ix, uint32_t posidx, const uint32_t *&posp) {
  posp = nullptr;
  if (ix.typ == 0) {
    if (posidx < ix.nkeys) {
      const uint32_t s = __ldg(ix.idx + posidx);
      posp = ix.pos + s;
      return __ldg(ix.idx + posidx + 1) - s;
    }
  } else if (posidx < ix.npos) {
    const uint32_t s = __ldg(ix.posidx + posidx);
    posp = ix.pos + s;
    return __ldg(ix.posidx + posidx + 1) - s;
  }
  return 0;
}

// sort2UINTarraysByQuickSort (sort.c:233-330): median-of-three quicksort, insertion sort
// below 7 elements, explicit stack, smaller partition first.  Unstable; the exchange
// sequence is reproduced so that ties end up in the reference's order.
__device__ int sort2(int n, uint32_t *key, uint32_t *val) {
  int lo = 0, hi = n - 1, sp = 0, i, j;
  int stack[62];
#define XC(a, b) do { uint32_t t_ = (a); (a) = (b); (b) = t_; } while (0)
  for (;;) {
    if (hi - lo < 7) {
      for (j = lo + 1; j <= hi; ++j) {
        const uint32_t k = key[j], v = val[j];
        for (i = j - 1; i >= lo && key[i] > k; --i) { key[i + 1] = key[i]; val[i + 1] = val[i]; }
        key[i + 1] = k; val[i + 1] = v;
      }
      if (!sp) return 0;
      hi = stack[sp--];
      lo = stack[sp--];
    } else {
      const int mid = (lo + hi) >> 1;
      XC(key[mid], key[lo + 1]); XC(val[mid], val[lo + 1]);
      if (key[lo] > key[hi]) { XC(key[lo], key[hi]); XC(val[lo], val[hi]); }
      if (key[lo + 1] > key[hi]) { XC(key[lo + 1], key[hi]); XC(val[lo + 1], val[hi]); }
      if (key[lo] > key[lo + 1]) { XC(key[lo], key[lo + 1]); XC(val[lo], val[lo + 1]); }
      i = lo + 1; j = hi;
      const uint32_t pk = key[lo + 1], pv = val[lo + 1];
      for (;;) {
        do ++i; while (key[i] < pk);
        do --j; while (key[j] > pk);
        if (j < i) break;
        XC(key[i], key[j]); XC(val[i], val[j]);
      }
      key[lo + 1] = key[j]; val[lo + 1] = val[j];
      key[j] = pk; val[j] = pv;
      sp += 2;
      if (sp > 60) return 34;  // ERRCODE_SORTSTACK
      if (hi - i + 1 >= j - lo) { stack[sp] = hi; stack[sp - 1] = i; hi = j - 1; }
      else { stack[sp] = j - 1; stack[sp - 1] = lo; lo = i; }
    }
  }
#undef XC
}

enum { HQ_TERM = 0, HQ_NORMHIT = 1, HQ_MULTIHIT = 2, HQ_REPEAT = 3, HQ_NOHIT = 4, HQ_NONSTDNT = 5 };
enum { HI_REVERSE = 1, HI_SORTED = 2, HI_RANK = 4 };

__global__ void __launch_bounds__(128)
seed_kernel(const Index ix, const uint8_t *__restrict__ arena, const SeedArgs a) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 2 * a.nreads) return;
  const int rd = g >> 1, is_reverse = g & 1;
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

cudaError_t launch_seed(const Index &ix, const uint8_t *arena, const SeedArgs &a, cudaStream_t st,
                        int *nlaunch) {
  const int n = 2 * a.nreads;
  if (n <= 0) return cudaSuccess;
  seed_kernel<<<(n + 127) / 128, 128, 0, st>>>(ix, arena, a);
  ++*nlaunch;
  return cudaGetLastError();
}

}  // namespace smb
