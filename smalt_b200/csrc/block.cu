// block.cu - the stages between K1 and K3 on the device, for a whole block of reads:
//
//   block_reqs_kernel     the hit-list requests of every job (read x strand x sequence | interval),
//                         what collectHits / collectHitsFromInterVal ask for (rmap.c:283-318, :438-493)
//   block_cands_kernel    candidate selection, one warp per job and one lane per hit list: hit regions, seeds, constant-shift
//                         segments (segLstFillHits, segment.c:763-810 with :396-584), candidates by coverage
//                         (segAliCandsAddFast -> addCandsFast / derriveSEGCAND, :1140-1223, :929-1059),
//                         threshold + sort + depth cut (segAliCandsStats, :1616-1785; the reference's
//                         unstable quicksort decides the order of equal covers)
//   block_emit_k2_kernel  windows and bands (segAliCandsCalcSegmentOffsets, :1861-1985), the SIMD predicate
//                         of rmap.c:715-718, K2 / K2' task lists grouped by kernel class
//   block_replay_kernel   the sequential part of scoreRMAPCAND (rmap.c:745-786) and the thresholds of
//                         mapSingleRead (:1373-1400) on the scores, which candidates go to K3
//   block_emit_k3_kernel  K3 tasks (widened bands, rmap.c:888-896) grouped by kernel class
//
// Integer work on small per-read lists (a few dozen hits, a handful of candidates): one lane walks
// one hit list exactly in the reference's order, 32-bit and 64-bit quantities as there.
#include "block.cuh"
#include <cstdlib>
#include "band.h"
#include "sort2.cuh"

namespace smb {

namespace {

constexpr int HALFBIT = 31;                         // HASHHIT_HALFBIT (hashhit.h:67)
constexpr uint32_t HALFMASK = 0x7FFFFFFFu;          // HASHHIT_HALFMASK
constexpr unsigned long long SOFFSMASK = 0xFFFFFFFFull;
constexpr int SEGMENTING_DIFFSHIFT = 3, MAXIMUM_DEPTH = 8000, DEFAULT_TARGET_DEPTH = 200, EDGE_BAND_FACTOR = 4,
              MAX_BANDEDGE_2POW = 4;                // segment.c:118-143
constexpr uint8_t CANDFLG_REVERSE = 1, CANDFLG_MMALI = 4;
constexpr int ERR_ASSERT = SMB_ERRCODE_ASSERT, ERR_OVERFLOW = SMB_ERRCODE_OVERFLOW;
constexpr int MINLEN_QUERY_STRIPED = 32, BWSCAL_QLEN = 48;   // rmap.c:83-84
constexpr int SW2_MAXROWS_ = 512;                   // sw_score.cu: staged window rows of the paired kernel

__device__ __forceinline__ uint64_t shiftpart(uint64_t x) { return x & ~((uint64_t)HALFMASK); }

// Division by the sampling step.  Every dividend on this path is a read offset or a seed length; for
// x < 2^24 and d <= 255, floor(x / d) = umulhi(x, floor(2^32 / d) + 1) exactly (the error term x * e / 2^32 with
// e <= 1 stays below 1 / d; checked exhaustively in tests/test_stepdiv_model.py - beyond 2^24 the first wrong
// quotient is at x = 16 909 559, d = 255), anything larger takes the hardware divide; a replacement of the
// ~20-instruction runtime divide that sits in the innermost hit loops.
struct StepDiv {
  uint32_t d, m;
  __device__ __forceinline__ explicit StepDiv(int step) : d((uint32_t)step), m(step > 1 ? (uint32_t)(0x100000000ull / (uint32_t)step) + 1u : 0u) {}
  __device__ __forceinline__ uint32_t div(uint32_t x) const { return d > 1u ? (x < (1u << 24) ? __umulhi(x, m) : x / d) : x; }
  __device__ __forceinline__ uint32_t mod(uint32_t x) const { return x - div(x) * d; }
  __device__ __forceinline__ int divs(int x) const { return x >= 0 ? (int)div((uint32_t)x) : -(int)div((uint32_t)(-x)); }   // C truncation
};

// sets bits [q0, q0 + len) of the coverage mask, returns how many of them were clear
__device__ __forceinline__ uint32_t mask_cover(uint32_t *mask, uint32_t q0, uint32_t len) {
  uint32_t fresh = 0;
  uint32_t q = q0;
  const uint32_t end = q0 + len;
  while (q < end) {
    const uint32_t w = q >> 5, b = q & 31u;
    const uint32_t n = min(32u - b, end - q);
    const uint32_t bits = (n == 32u ? 0xffffffffu : ((1u << n) - 1u)) << b;
    const uint32_t old = mask[w];
    fresh += __popc(bits & ~old);
    mask[w] = old | bits;
    q += n;
  }
  return fresh;
}

struct SegView {   // the seeds / segments of the current hit region (per job scratch)
  uint64_t *sd_sqo;
  int32_t *sd_len;
  uint32_t *sg_ix;
  int32_t *sg_nseed;
  uint32_t *sg_cover;
};

// calcSegmentBoundaries (segment.c:635-668)
__device__ __forceinline__ void seg_bounds(uint32_t &qs, uint32_t &qe, uint32_t &rs, uint32_t &re, const SegView &v,
                                           int sg, int ktup, const StepDiv &nskip, bool is_reverse) {
  const uint32_t i0 = v.sg_ix[sg], i1 = i0 + (uint32_t)v.sg_nseed[sg] - 1u;
  const uint64_t s0 = v.sd_sqo[i0], s1 = v.sd_sqo[i1];
  const int32_t l1 = v.sd_len[i1];
  qs = (uint32_t)(s0 & HALFMASK);
  qe = (uint32_t)(s1 & HALFMASK) + (uint32_t)l1 - 1u;
  if (is_reverse) {
    rs = (uint32_t)(((s1 >> HALFBIT) - nskip.div((uint32_t)(s1 & HALFMASK))) & SOFFSMASK);
    rs -= (uint32_t)nskip.divs(l1 - ktup);
    re = (uint32_t)(((s0 >> HALFBIT) - nskip.div(qs)) & SOFFSMASK);
  } else {
    rs = (uint32_t)(((s0 >> HALFBIT) + nskip.div(qs)) & SOFFSMASK);
    re = (uint32_t)(((s1 >> HALFBIT) + nskip.div((uint32_t)(s1 & HALFMASK))) & SOFFSMASK);
    re += (uint32_t)nskip.divs(l1 - ktup);
  }
}

// derriveSEGCAND (segment.c:929-1059) for segments [first, first + nseg) of the region
__device__ int derive_cand(SegCand &cd, int first, int nseg, const SegView &v, int ktup, const StepDiv &nskip, uint32_t cover,
                           uint32_t mincover_noindel, bool is_reverse) {
  if (v.sg_nseed[first] < 0) return ERR_ASSERT;
  uint32_t cqs, cqe, crs, cre;
  seg_bounds(cqs, cqe, crs, cre, v, first, ktup, nskip, is_reverse);
  v.sg_nseed[first] *= -1;
  long long shift_min = (long long)(v.sd_sqo[v.sg_ix[first]] >> HALFBIT), shift_2mm = shift_min;
  uint32_t maxcover = v.sg_cover[first];
  int last = first;
  for (int n = 1; n < nseg; ++n) {
    const int sg = first + n;
    if (v.sg_nseed[sg] < 0) return ERR_ASSERT;
    uint32_t qs, qe, rs, re;
    seg_bounds(qs, qe, rs, re, v, sg, ktup, nskip, is_reverse);
    if (v.sg_cover[sg] > maxcover) {
      shift_2mm = (long long)(v.sd_sqo[v.sg_ix[sg]] >> HALFBIT);
      maxcover = v.sg_cover[sg];
    }
    v.sg_nseed[sg] *= -1;
    if (qs < cqs) cqs = qs;
    if (qe > cqe) cqe = qe;
    if (rs < crs) crs = rs;
    if (re > cre) cre = re;
    last = sg;
  }
  uint8_t flag = 0;
  long long shift_start;
  if (is_reverse) {
    flag |= CANDFLG_REVERSE;
    shift_start = ((long long)crs) + (long long)nskip.div(cqe - (uint32_t)ktup + 1u);
  } else {
    shift_start = (long long)(((unsigned long long)crs) | (1ull << (HALFBIT + 1))) - (long long)nskip.div(cqs);
  }
  const unsigned long long shift_range =
      (unsigned long long)(((long long)(v.sd_sqo[v.sg_ix[last]] >> HALFBIT)) - shift_min);
  const long long diff_shift = shift_min - shift_start;
  if (shift_range > 32767ull) return ERR_OVERFLOW;
  if (diff_shift < -32768ll || diff_shift > 32767ll) return ERR_OVERFLOW;
  cd.shiftoffs = (short)diff_shift;
  if (maxcover >= mincover_noindel) {
    const long long ds = shift_2mm - shift_start;
    flag |= CANDFLG_MMALI;
    if (ds < -32768ll || ds > 32767ll) return ERR_OVERFLOW;
    cd.shift2mm = (short)ds;
  } else {
    cd.shift2mm = 0;
  }
  cd.qs = cqs; cd.qe = cqe; cd.rs = crs; cd.re = cre;
  cd.flag = flag;
  cd.pad = 0;
  cd.srange = (short)shift_range;
  cd.cover = cover;
  cd.seqidx = -1;
  return 0;
}

// One hit list -> candidates appended to cand[ncand...].
__device__ int add_list(const uint64_t *__restrict__ sqdat, int nhits, bool is_reverse, uint32_t qlen, int ktup,
                        int nskip_, uint32_t min_ktup, uint32_t mincover, int seqidx, const SegView &v,
                        uint32_t *mask, uint32_t mask_words, SegCand *cand, uint32_t &ncand, uint32_t &max_cover,
                        uint32_t &max2nd_cover) {
  if (nhits < 1) return 0;
  const StepDiv nskip(nskip_);
  // defineHitRegions (segment.c:396-453)
  uint32_t max_dshift = (uint32_t)(ktup * SEGMENTING_DIFFSHIFT / nskip_) & 0xffffu;
  const uint32_t ds = (qlen - (uint32_t)ktup) / (uint32_t)nskip_ + 1u;
  if (ds < max_dshift) max_dshift = ds & 0xffffu;
  const uint64_t dsthresh = ((uint64_t)max_dshift) << HALFBIT;
  int i = 0;
  while (i < nhits) {
    int j = i + 1;
    uint64_t prev = sqdat[i];
    for (; j < nhits; ++j) {
      const uint64_t cur = sqdat[j];
      if (cur - prev >= dsthresh) break;
      prev = cur;
    }
    if ((uint32_t)(j - i) >= min_ktup) {
      // makeSeedsFromHits (segment.c:455-533) of this region
      int nseed = 0;
      for (int a = i; a < j;) {
        const uint64_t sqo = sqdat[a];
        const uint64_t shift = shiftpart(sqo);
        const uint32_t qoffs = (uint32_t)(sqo & HALFMASK);
        uint32_t lastq = qoffs + (uint32_t)ktup;
        int b = a + 1;
        for (; b < j; ++b) {
          const uint64_t h = sqdat[b];
          if (shiftpart(h) != shift) break;
          const uint32_t qo = (uint32_t)(h & HALFMASK);
          if (qo > lastq || nskip.mod(qo - qoffs)) break;
          lastq = qo + (uint32_t)ktup;
        }
        v.sd_sqo[nseed] = sqo;
        v.sd_len[nseed] = (int32_t)(lastq - qoffs);
        ++nseed;
        a = b;
      }
      // makeSegmentsFromSeeds (segment.c:535-584)
      int nsegm = 0;
      for (int a = 0; a < nseed;) {
        const uint64_t sqo = v.sd_sqo[a];
        const uint64_t shift = shiftpart(sqo);
        const uint32_t qoffs = (uint32_t)(sqo & HALFMASK);
        uint32_t cover = (uint32_t)v.sd_len[a];
        int b = a + 1;
        for (; b < nseed; ++b) {
          const uint64_t s2 = v.sd_sqo[b];
          if (shiftpart(s2) != shift || nskip.mod(((uint32_t)(s2 & HALFMASK)) - qoffs)) break;
          cover += (uint32_t)v.sd_len[b];
        }
        v.sg_ix[nsegm] = (uint32_t)a;
        v.sg_nseed[nsegm] = b - a;
        v.sg_cover[nsegm] = cover;
        ++nsegm;
        a = b;
      }
      // addCandsFast (segment.c:1140-1223) on this region
      for (int s = 0; s < nsegm;) {
        for (uint32_t w = 0; w < mask_words; ++w) mask[w] = 0u;
        {
          const uint32_t i0 = v.sg_ix[s];
          for (int l = 0; l < v.sg_nseed[s]; ++l)
            mask_cover(mask, (uint32_t)(v.sd_sqo[i0 + l] & HALFMASK), (uint32_t)v.sd_len[i0 + l]);
        }
        uint32_t cover = v.sg_cover[s];
        int t = s + 1;
        for (; t < nsegm; ++t) {
          if (v.sg_nseed[t] < 0) break;
          uint32_t cover_new = 0;
          const uint32_t i0 = v.sg_ix[t];
          for (int l = 0; l < v.sg_nseed[t]; ++l)
            cover_new += mask_cover(mask, (uint32_t)(v.sd_sqo[i0 + l] & HALFMASK), (uint32_t)v.sd_len[i0 + l]);
          if ((cover_new << 1) < v.sg_cover[t] && cover >= mincover) break;
          cover += cover_new;
        }
        if (cover >= mincover) {
          SegCand cd;
          const int e = derive_cand(cd, s, t - s, v, ktup, nskip, cover, mincover, is_reverse);
          if (e) return e;
          cd.seqidx = seqidx;
          cand[ncand++] = cd;
          if (cover > max2nd_cover) {
            if (cover > max_cover) { max2nd_cover = max_cover; max_cover = cover; }
            else if (cover != max_cover) max2nd_cover = cover;
          }
        }
        s = t;
      }
    }
    i = j;
  }
  return 0;
}

struct Offsets {
  uint32_t qs, qe;
  uint64_t rs, re;
  int band_l, band_r;
};

// segAliCandsCalcSegmentOffsets (segment.c:1861-1985) with edgelen 0 (makeRMAPCANDfromSegment under
// SCORE_SIMD, rmap.c:535-556); seqidx is always a valid sequence on this path
__device__ int cand_offsets(Offsets &o, const SegCand &sc, uint32_t qlen, int ktup, int nskip,
                            const uint64_t *__restrict__ seq_offs, int nseq, bool termchar) {
  if (sc.seqidx < 0 || sc.seqidx >= nseq) return ERR_ASSERT;
  const uint64_t roffs = seq_offs[sc.seqidx];
  uint64_t rlen = seq_offs[sc.seqidx + 1] - roffs;
  if (termchar && rlen > 0) --rlen;
  rlen = (uint32_t)rlen;
  uint64_t rs = ((uint64_t)sc.rs) * (uint64_t)nskip;
  uint64_t re = ((uint64_t)sc.re) * (uint64_t)nskip + (uint64_t)ktup - 1u;
  if (rs < roffs || re < rs) return ERR_ASSERT;
  rs -= roffs;
  re -= roffs;
  if (re >= rlen) return ERR_ASSERT;
  if (sc.qe < sc.qs || sc.qs >= qlen) return ERR_ASSERT;
  uint32_t qs, qe;
  if (sc.flag & CANDFLG_REVERSE) { qs = qlen - sc.qe - 1u; qe = qlen - sc.qs - 1u; }
  else { qs = sc.qs; qe = sc.qe; }
  int edge_band = (int)(qlen - sc.cover) / EDGE_BAND_FACTOR;
  if (edge_band > nskip) {
    if (edge_band > (int)(qlen >> MAX_BANDEDGE_2POW)) edge_band = (int)(qlen >> MAX_BANDEDGE_2POW);
    edge_band -= nskip - 1;
  } else {
    edge_band = 0;
  }
  const int br = (-sc.shiftoffs + 1) * nskip + edge_band + 1;
  const int bl = br - (sc.srange + 2) * nskip - 2 * edge_band - 2;
  const int q_edge_l = (int)qs, q_edge_r = (int)(qlen - qe - 1u);
  qs -= (uint32_t)q_edge_l;
  qe += (uint32_t)q_edge_r;
  int r_edge_l = q_edge_l + br;
  const int r_edge_r = q_edge_r - bl;
  if (r_edge_l > 0 && rs < (uint64_t)r_edge_l) { r_edge_l = (int)rs; rs = 0; }
  else rs -= (uint64_t)(long long)r_edge_l;
  if (re + (uint64_t)(long long)r_edge_r >= rlen) re = rlen - 1u;
  else re += (uint64_t)(long long)r_edge_r;
  if (re < rs) return ERR_ASSERT;
  const int band_offs = q_edge_l - r_edge_l;
  o.band_l = bl + band_offs + (int)qs;
  o.band_r = br + band_offs + (int)qs;
  o.qs = qs; o.qe = qe; o.rs = rs; o.re = re;
  // makeRMAPCANDfromSegment (rmap.c:553-555)
  if (qe > 0x7fffffffu || re - rs > 0x7fffffffull) return ERR_OVERFLOW;
  return 0;
}

__device__ __forceinline__ bool pen16_k2(const BlockArgs &a) {   // plan_sw (sw_score.cu)
  return a.match > 0 && a.match < 128 && a.mismatch <= 0 && a.mismatch > -128 && a.gap_init >= 0 && a.gap_init < 8000 &&
         a.gap_ext >= 0 && a.gap_ext < 8000;
}
__device__ __forceinline__ int k2_bin(const BlockArgs &a, uint32_t qlen, uint32_t reflen) {
  int c = (int)((qlen + 31u) / 32u);
  c = c < 1 ? 1 : (c > 8 ? 8 : c);
  const bool fits16 = pen16_k2(a) && (long long)qlen * a.match <= 16000;
  if (fits16 && qlen > 256u) return 18;   // long reads, paired (sw_long2_kernel)
  const bool pair16 = fits16 && qlen <= 256u && reflen <= (uint32_t)SW2_MAXROWS_;
  return pair16 ? c + 8 : c;
}
__device__ __forceinline__ bool simd_pred(uint32_t qlen, const Offsets &o) {   // rmap.c:715-718
  return qlen >= (uint32_t)MINLEN_QUERY_STRIPED && ((uint32_t)(o.band_r - o.band_l) * (uint32_t)BWSCAL_QLEN) > qlen &&
         o.qs == 0u && o.qe >= qlen - 1u;
}

// a slot of list `bin` for every calling thread, one atomic per warp and bin, lane order kept
__device__ __forceinline__ unsigned int warp_ticket(unsigned int *cursor, int bin, bool want) {
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, want ? bin : -1);
  unsigned int base = 0;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  if (want && lane == leader) base = atomicAdd(cursor + bin, (unsigned int)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + (unsigned int)__popc(peers & ((1u << lane) - 1u));
}

}  // namespace

// ------------------------------------------------------------------------------------
// With many reference sequences nearly all hit lists of a read are empty (collectHits asks for one list
// per sequence and strand, rmap.c:293-318: 48 requests per read on a 24-sequence genome, ~1 of them with
// hits).  One warp per job x strand marks the sequences that hold a position of any ranked seed; the request
// of every other sequence is answered "empty" without looking at the seeds again.
__global__ void __launch_bounds__(128) block_seqmask_kernel(const BlockArgs a) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (g >= 2 * a.njobs) return;
  const smb_block_job jb = a.jobs[g >> 1];
  const uint32_t rd = jb.seed_read, st = (uint32_t)(g & 1);
  const uint32_t qlen = a.seed.read_len[rd];
  const uint64_t slot = a.seed.slot_off[rd] + (st ? qlen : 0u);
  const smb_seed_info inf = a.seed.info[2 * rd + st];
  unsigned long long m = 0;
  if (jb.niv >= 0 || inf.err) m = ~0ull;   // (interval searches take all seeds: not pruned)
  else {
    const uint32_t ns = inf.seed_rank > 0 ? inf.seed_rank : inf.n_seeds;
    const uint32_t *posidx = a.seed.posidx + slot, *sidx = a.seed.sidx + slot;
    for (uint32_t n = lane; n < ns; n += 32) {
      const uint32_t *posp;
      const uint32_t nh = fetch_positions(a.ix, posidx[sidx[n]], posp);
      if (!posp) continue;
      if (nh > 256u) { m = ~0ull; continue; }   // very frequent word: every sequence may hold it
      for (uint32_t i = 0; i < nh; ++i) {
        const uint64_t p = __ldg(posp + i);
        int lo = 0, hi = a.nseq;                // last s with seq_offs[s] / nskip <= p
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (a.seq_offs[mid] / (uint64_t)a.nskip <= p) lo = mid; else hi = mid;
        }
        m |= 1ull << lo;
        // a position on the boundary may count for the previous sequence too (hi of s == lo of s + 1 after the division)
        if (lo > 0 && a.seq_offs[lo] / (uint64_t)a.nskip == p) m |= 1ull << (lo - 1);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
  if (lane == 0) a.seqmask[g] = m;
}

__global__ void __launch_bounds__(128) block_reqs_kernel(const BlockArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.njobs) return;
  const smb_block_job jb = a.jobs[j];
  const int nlist = jb.niv < 0 ? a.nseq : jb.niv;
  uint32_t r = a.job_req[j];
  for (int st = 0; st < 2; ++st)
    for (int c = 0; c < nlist; ++c, ++r) {
      smb_hit_req rq;
      int sx;
      if (jb.niv < 0) { rq.lo = a.seq_offs[c]; rq.hi = a.seq_offs[c + 1]; sx = c; rq.use_short = 1; }
      else {
        const smb_block_ival iv = a.ivals[jb.iv_first + c];
        rq.lo = iv.lo; rq.hi = iv.hi; sx = iv.seqidx; rq.use_short = 0;
      }
      rq.read = jb.seed_read;
      rq.nhit_max = a.prm.nhit_max;
      rq.strand = (uint8_t)st;
      rq.reserved[0] = rq.reserved[1] = 0;
      rq.nhits_max = 0;
      a.req[r] = rq;
      a.req_seqidx[r] = sx;
      if (a.req_skip) a.req_skip[r] = (a.seqmask && jb.niv < 0 && !((a.seqmask[2 * j + st] >> c) & 1ull)) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------
// ONE WARP PER JOB, ONE LANE PER HIT LIST.  The hit lists of a job (read x strand x sequence | interval)
// are independent until segAliCandsStats: every lane runs the reference's sequential segmentation on its
// own list (regions, seeds, segments, candidates by coverage) with its scratch in shared memory, the
// candidates of the lists are then numbered in list order (the order candr has in the reference), the two
// largest distinct covers are merged, the threshold filter keeps that order, lane 0 runs the reference's
// quicksort and depth cut on the (few) keys, and the windows of the selected candidates are checked by all
// lanes.  Jobs whose lists do not fit the shared scratch (more than CW_HITS hits, reads beyond 32 * CW_MASKW
// bases) use the per-job scratch in HBM with the same code.
// Few lists per job (one or two sequences): G = 4 .. 16 lanes per job and 32 / G jobs per warp, otherwise most
// lanes of the warp would idle (5.6 of 32 busy on a one-sequence genome, profiles/r2_ncu_full_cands_*).
constexpr int CW_WARPS = 2;      // warps per CTA
constexpr int CW_MASKW = 8;      // coverage mask words per lane in shared memory (reads of <= 256 bases)
// hits of a job whose seeds / segments fit the shared scratch, by group size
__host__ __device__ constexpr int cw_hits(int G) { return G >= 32 ? 320 : G >= 16 ? 256 : G >= 8 ? 192 : 112; }
__host__ __device__ constexpr size_t cw_job_bytes(int G) { return (size_t)cw_hits(G) * 24u; }     // u64 + 4 x 32 bit per hit
__host__ __device__ constexpr size_t cw_warp_bytes(int G) { return (size_t)(32 / G) * cw_job_bytes(G) + 32u * CW_MASKW * 4u; }

// the two largest distinct values of two (max, second) pairs
__device__ __forceinline__ void merge_top2(uint32_t &m, uint32_t &s, uint32_t m2, uint32_t s2) {
  const uint32_t M = max(m, m2);
  uint32_t S = 0;
  if (m < M && m > S) S = m;
  if (m2 < M && m2 > S) S = m2;
  if (s < M && s > S) S = s;
  if (s2 < M && s2 > S) S = s2;
  m = M;
  s = S;
}

template <int G>
__global__ void __launch_bounds__(CW_WARPS * 32) block_cands_kernel(const BlockArgs a) {
  extern __shared__ __align__(16) unsigned char s_cw[];
  constexpr int CW_HITS = cw_hits(G);
  constexpr int JPW = 32 / G;
  const int wlane = threadIdx.x & 31;                 // lane of the warp
  const int lane = wlane & (G - 1);                   // lane of the job's group
  const unsigned FULL = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (wlane - lane));   // the group's lanes
  const int j = (blockIdx.x * CW_WARPS + (threadIdx.x >> 5)) * JPW + wlane / G;
  if (j >= a.njobs) return;
  unsigned char *const wbase = s_cw + (size_t)(threadIdx.x >> 5) * cw_warp_bytes(G);
  unsigned char *const jbase = wbase + (size_t)(wlane / G) * cw_job_bytes(G);
  struct {
    unsigned long long *sd_sqo; int *sd_len; unsigned int *sg_ix; int *sg_nseed; unsigned int *sg_cover; unsigned int *mask;
  } sm;
  sm.sd_sqo = (unsigned long long *)jbase;
  sm.sd_len = (int *)(jbase + (size_t)CW_HITS * 8u);
  sm.sg_ix = (unsigned int *)(jbase + (size_t)CW_HITS * 12u);
  sm.sg_nseed = (int *)(jbase + (size_t)CW_HITS * 16u);
  sm.sg_cover = (unsigned int *)(jbase + (size_t)CW_HITS * 20u);
  sm.mask = (unsigned int *)(wbase + (size_t)JPW * cw_job_bytes(G));
  const smb_block_job jb = a.jobs[j];
  smb_block_read rd;
  memset(&rd, 0, sizeof rd);
  const uint32_t r = jb.seed_read;
  const uint32_t qlen = a.seed.read_len[r];
  const smb_seed_info inf0 = a.seed.info[2 * r], inf1 = a.seed.info[2 * r + 1];
  rd.errcode = inf0.err ? inf0.err : inf1.err;
  if (lane == 0) {
    a.n_sort[j] = 0;
    a.nk3[j] = 0;
    a.cover_deficit[2 * j] = inf0.cover_deficit;
    a.cover_deficit[2 * j + 1] = inf1.cover_deficit;
  }
  if (rd.errcode) { if (lane == 0) a.rd[j] = rd; return; }
  const int ktup = a.ktup, nskip = a.nskip;
  // calcMinKtup (rmap.c:240-247) and the prelude of mapSingleRead (:1282-1288)
  uint32_t min_cover = jb.min_cover;
  uint32_t min_ktup = (min_cover >= (uint32_t)(ktup + nskip)) ? (min_cover - (uint32_t)ktup) / (uint32_t)nskip : 1u;
  min_cover = (min_ktup - 1u) * (uint32_t)nskip + (uint32_t)ktup;
  uint32_t below;
  if (a.prm.min_swatscor_below_max < 0) {
    below = qlen - 1u;
  } else {
    const short mismatchdiff = (short)(a.match - a.mismatch);
    below = ((uint32_t)(a.prm.min_swatscor_below_max / mismatchdiff)) * (uint32_t)nskip;
    if (below < (uint32_t)ktup || a.prm.best) below = (uint32_t)(ktup + 2 * (nskip - 1));
  }
  // segLstFillHits (segment.c:781-788): one k-tuple less per entry of the list's mask that is not a
  // NORMHIT - the masks of segment lists are qlen x HITQUAL_NOHIT (hashhit.c:1224-1230)
  if (min_ktup >= 2u) min_ktup = (min_ktup - 1u > qlen) ? min_ktup - qlen : 1u;

  const int nlist = jb.niv < 0 ? a.nseq : jb.niv;
  const int nl = 2 * nlist;
  const uint32_t rq0 = a.job_req[j];
  const uint64_t base = a.hit_off[rq0];
  const uint64_t H = a.hit_off[rq0 + nl] - base;
  const uint32_t mask_words = (qlen + 31u) / 32u;
  const bool in_smem = H <= (uint64_t)CW_HITS && mask_words <= (uint32_t)CW_MASKW;
  uint32_t *mask = mask_words <= (uint32_t)CW_MASKW ? sm.mask + wlane * CW_MASKW : a.mask + ((size_t)j * 32u + lane) * a.mask_words;

  // ---- phase 1: the lists, 32 at a time ----
  uint32_t max_cover = 0, max2nd_cover = 0, ncand_all = 0;
  for (int l0 = 0; l0 < nl && !rd.errcode; l0 += G) {
    const int l = l0 + lane;
    int e = 0;
    uint32_t nc = 0, mc = 0, m2 = 0;
    if (l < nl) {
      const uint32_t rq = rq0 + (uint32_t)l;
      e = a.req_err[rq];
      if (e == SMB_ERRCODE_ALLOCBOUNDARY) e = 0;
      if (!e) {
        const uint64_t f0 = a.hit_off[rq], f1 = a.hit_off[rq + 1];
        const uint64_t o = f0 - base;
        SegView v;
        if (in_smem) v = SegView{(uint64_t *)sm.sd_sqo + o, sm.sd_len + o, sm.sg_ix + o, sm.sg_nseed + o, sm.sg_cover + o};
        else v = SegView{a.sd_sqo + f0, a.sd_len + f0, a.sg_ix + f0, a.sg_nseed + f0, a.sg_cover + f0};
        e = add_list(a.sqdat + f0, (int)(f1 - f0), l >= nlist, qlen, ktup, nskip, min_ktup, min_cover, a.req_seqidx[rq], v,
                     mask, mask_words, a.cand + f0, nc, mc, m2);
      }
      a.req_ncand[rq] = nc;
    }
    // the first list (in order) with an error ends the job
    const unsigned bad = __ballot_sync(FULL, e != 0);
    if (bad) { rd.errcode = __shfl_sync(FULL, e, __ffs(bad) - 1); break; }
    for (int o = G / 2; o > 0; o >>= 1) {
      const uint32_t om = __shfl_xor_sync(FULL, mc, o), o2 = __shfl_xor_sync(FULL, m2, o);
      merge_top2(mc, m2, om, o2);
      nc += __shfl_xor_sync(FULL, nc, o);
    }
    merge_top2(max_cover, max2nd_cover, mc, m2);
    ncand_all += nc;
  }
  if (rd.errcode) { if (lane == 0) a.rd[j] = rd; return; }

  // ---- phase 2: segAliCandsStats (segment.c:1616-1785): threshold filter in list order ----
  uint32_t max_depth = (uint32_t)a.prm.max_depth, target_depth = (uint32_t)a.prm.target_depth;   // SEGNUM_t
  if (max_depth < 1u || max_depth > (uint32_t)MAXIMUM_DEPTH) max_depth = MAXIMUM_DEPTH;
  if (target_depth < 1u) target_depth = DEFAULT_TARGET_DEPTH;
  if (target_depth > max_depth) target_depth = max_depth;
  uint32_t thr = (below > max_cover) ? 0u : max_cover - below, cdf = 0;
  if (thr > max2nd_cover) { cdf = thr - max2nd_cover; thr = max2nd_cover; }
  const uint32_t cda = inf0.cover_deficit > cdf ? inf0.cover_deficit - cdf : 0u;   // both strands: FORWARD deficit (:1674)
  uint32_t *skey = a.sort_key + base, *sidx = a.sort_idx + base;
  uint32_t nk = 0;
  for (int l0 = 0; l0 < nl; l0 += G) {
    const int l = l0 + lane;
    uint32_t cnt = 0, nc = 0;
    uint64_t f0 = 0;
    if (l < nl) {
      const uint32_t rq = rq0 + (uint32_t)l;
      nc = a.req_ncand[rq];
      f0 = a.hit_off[rq];
      for (uint32_t i = 0; i < nc; ++i) cnt += (a.cand[f0 + i].cover + cda >= thr);
    }
    uint32_t pos = cnt;   // inclusive scan over the lanes
    for (int o = 1; o < G; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, pos, o, G);
      if (lane >= o) pos += t;
    }
    const uint32_t tot = __shfl_sync(FULL, pos, G - 1, G);
    pos = nk + pos - cnt;
    for (uint32_t i = 0; i < nc; ++i) {
      const uint32_t cov = a.cand[f0 + i].cover;
      if (cov + cda < thr) continue;
      skey[pos] = max_cover - cov;
      sidx[pos] = (uint32_t)(f0 - base) + i;
      ++pos;
    }
    nk += tot;
  }
  __syncwarp(FULL);
  // ---- phase 3: the reference's quicksort and the depth cut, lane 0 (in shared memory if the keys fit) ----
  const uint32_t n_mincover = nk;
  uint32_t *wk = skey, *wi = sidx;
  const bool sort_smem = nk <= (uint32_t)CW_HITS;   // (the seed scratch is free again)
  if (sort_smem) {
    wk = (uint32_t *)sm.sd_sqo;
    wi = wk + CW_HITS;
    for (uint32_t i = lane; i < nk; i += G) { wk[i] = skey[i]; wi[i] = sidx[i]; }
    __syncwarp(FULL);
  }
  int serr = 0;
  if (lane == 0) {
    serr = sort2((int)nk, wk, wi);
    if (!serr && nk > target_depth) {
      const uint32_t maxj = (nk < max_depth) ? nk : max_depth;
      uint32_t t = target_depth;
      if (a.prm.sensitive) {
        for (; t < maxj; ++t)
          if (wk[t] >= cda) break;            // (the reference indexes candr by t here, both strands share cda)
        for (; t < n_mincover && wk[t] < (uint32_t)nskip; ++t);
      } else {
        uint32_t cov = wk[nk / 2];
        if (cov < (uint32_t)nskip) cov = (uint32_t)nskip;
        for (; t < maxj && wk[t] < cov; ++t);
      }
      nk = t;
    }
  }
  serr = __shfl_sync(FULL, serr, 0, G);
  nk = __shfl_sync(FULL, nk, 0, G);
  if (sort_smem) {
    __syncwarp(FULL);
    for (uint32_t i = lane; i < n_mincover; i += G) { skey[i] = wk[i]; sidx[i] = wi[i]; }
  }
  if (!serr && (nk > 0x7fffffffu || n_mincover > 0x7fffffffu)) serr = ERR_ASSERT;
  if (serr) { rd.errcode = serr; if (lane == 0) a.rd[j] = rd; return; }
  rd.nseg = (int32_t)nk;
  rd.nseg_tot = (int32_t)n_mincover;
  rd.nhit = inf0.nhit_rank + inf1.nhit_rank;      // calcTotalHitNumStats (rmap.c:1086-1094)
  rd.nhit_tot = inf0.nhit_tot + inf1.nhit_tot;
  rd.reached_stats = 1;
  // ---- phase 4: windows of the selected candidates (errors end the read, rmap.c:669-671), K2 launch bins ----
  // (two passes: the bins are only counted for reads all of whose windows are valid)
  unsigned int multi = 0;
  unsigned long long cells = 0;
  for (int pass = 0; pass < 2 && !rd.errcode; ++pass)
    for (uint32_t c0 = 0; c0 < nk; c0 += G) {
      const uint32_t c = c0 + lane;
      int e = 0, bin = -1;
      if (c < nk) {
        Offsets o;
        const uint32_t ix = sort_smem ? wi[c] : sidx[c];
        e = cand_offsets(o, a.cand[base + ix], qlen, ktup, nskip, a.seq_offs, a.nseq, a.prm.termchar != 0);
        if (!e && pass) {
          const uint32_t reflen = (uint32_t)(o.re - o.rs + 1u);
          if (simd_pred(qlen, o)) {
            bin = k2_bin(a, qlen, reflen);
            if (qlen > 256u && reflen > multi) multi = reflen;
            cells += (unsigned long long)qlen * reflen;
          } else {
            bin = 17;
          }
        }
      }
      if (!pass) {
        const unsigned bad = __ballot_sync(FULL, e != 0);
        if (bad) { rd.errcode = __shfl_sync(FULL, e, __ffs(bad) - 1); break; }
      } else {   // one atomic per warp and bin
        const unsigned peers = __match_any_sync(FULL, bin);
        if (bin >= 0 && wlane == __ffs(peers) - 1) atomicAdd(&a.cnt->k2_hist[bin], (unsigned int)__popc(peers));
      }
    }
  if (rd.errcode) { if (lane == 0) a.rd[j] = rd; return; }   // (no candidates: the wave driver drops them too)
  for (int o = G / 2; o > 0; o >>= 1) {
    cells += __shfl_xor_sync(FULL, cells, o);
    multi = max(multi, __shfl_xor_sync(FULL, multi, o));
  }
  if (lane == 0) {
    rd.ncand = nk;
    a.n_sort[j] = nk;
    if (multi) atomicMax(&a.cnt->max_rlen_multi, multi);
    if (cells) atomicAdd(&a.cnt->k2_cells, cells);
    a.rd[j] = rd;
  }
}

// ------------------------------------------------------------------------------------
// one thread per selected candidate of the block (dense index ci)
__global__ void __launch_bounds__(128) block_emit_k2_kernel(const BlockArgs a, const unsigned long long ncand) {
  const unsigned long long ci = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = ci < ncand;
  int bin = -1;
  bool simd = false;
  if (live) {
    // job of the candidate: last j with cand_first[j] <= ci
    int lo = 0, hi = a.njobs;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.cand_first[mid] <= ci) lo = mid; else hi = mid;
    }
    const int j = lo;
    const uint32_t c = (uint32_t)(ci - a.cand_first[j]);
    const smb_block_job jb = a.jobs[j];
    const uint32_t qlen = a.seed.read_len[jb.seed_read];
    const uint64_t base = a.hit_off[a.job_req[j]];
    const SegCand sc = a.cand[base + a.sort_idx[base + c]];
    Offsets o;
    cand_offsets(o, sc, qlen, a.ktup, a.nskip, a.seq_offs, a.nseq, a.prm.termchar != 0);   // checked by block_cands_kernel
    DCand d;
    d.rs = o.rs;
    d.reflen = (uint32_t)(o.re - o.rs + 1u);
    d.refoff = a.seq_offs[sc.seqidx] + o.rs;
    d.qs = o.qs; d.qe = o.qe;
    d.band_l = o.band_l; d.band_r = o.band_r;
    d.sqidx = sc.seqidx;
    d.cover = sc.cover;
    d.rev = (sc.flag & CANDFLG_REVERSE) ? 1 : 0;
    simd = simd_pred(qlen, o);
    d.simd = simd ? 1 : 0;
    d.pad[0] = d.pad[1] = 0;
    a.dc[ci] = d;
    a.dc_job[ci] = (uint32_t)j;
    a.k3rank[ci] = -1;
    a.score[ci] = 0;
    a.serr[ci] = 0;
    const uint32_t flags = SMB_TASK_REF_PACKED | (d.rev ? SMB_TASK_READ_REVCOMP : 0u);
    const uint64_t read_off = a.seed.read_off[jb.seed_read];
    if (simd) {
      smb_sw_task t;
      t.read_off = read_off; t.ref_off = d.refoff; t.read_len = qlen; t.ref_len = d.reflen; t.flags = flags; t.reserved = 0;
      a.swt[ci] = t;
      bin = k2_bin(a, qlen, d.reflen);
    } else {
      smb_band_task t;
      t.read_off = read_off; t.ref_off = d.refoff; t.read_len = qlen; t.ref_len = d.reflen; t.flags = flags;
      t.l_edge = d.band_l; t.r_edge = d.band_r; t.p_left = (int)d.qs; t.p_right = (int)d.qe;
      t.u_left = 0; t.u_right = (int)d.reflen - 1; t.minscore = 0; t.minscorlen = 0;
      a.bft[ci] = t;
      bin = 17;
    }
  }
  const unsigned int slot = warp_ticket(a.cnt->k2_cursor, bin < 0 ? 0 : bin, live);
  if (live) {
    if (simd) a.k2_order[a.k2_start[bin] + slot] = (int)ci;
    else a.bf_order[slot] = (int)ci;
  }
}

// SIMD scores that overflowed 16 bits (ERRCODE_SWATEXCEED) are scored again by the banded kernel
// (rmap.c:730-744): builds those K2' tasks
__global__ void __launch_bounds__(128) block_exceed_kernel(const BlockArgs a, const unsigned long long ncand) {
  const unsigned long long ci = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool hit = ci < ncand && a.dc[ci].simd && a.serr[ci] == SMB_ERRCODE_SWATEXCEED;
  if (hit) {
    const DCand d = a.dc[ci];
    const smb_block_job jb = a.jobs[a.dc_job[ci]];
    smb_band_task t;
    t.read_off = a.seed.read_off[jb.seed_read]; t.ref_off = d.refoff; t.read_len = a.seed.read_len[jb.seed_read];
    t.ref_len = d.reflen; t.flags = SMB_TASK_REF_PACKED | (d.rev ? SMB_TASK_READ_REVCOMP : 0u);
    t.l_edge = d.band_l; t.r_edge = d.band_r; t.p_left = (int)d.qs; t.p_right = (int)d.qe;
    t.u_left = 0; t.u_right = (int)d.reflen - 1; t.minscore = 0; t.minscorlen = 0;
    a.bft[ci] = t;
    a.dc[ci].simd = 0;
    a.bf_order[atomicAdd(&a.cnt->bf_cursor, 1u)] = (int)ci;
  }
}

// ------------------------------------------------------------------------------------
// K3 class of a task (plan_band, band_dp.cu)
__device__ __forceinline__ int k3_class(const BlockArgs &a, const smb_band_task &t) {
  const bool pen16 = a.match > 0 && a.match < 128 && a.mismatch <= 0 && a.mismatch > -128 && a.gap_init > 0 &&
                     a.gap_init < 4000 && a.gap_ext >= 0 && a.gap_ext < 4000;
  const int wl = band_warp_lanes(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                 (int)t.ref_len);
  if (wl == 16 && pen16 && (long long)t.read_len * a.match <= 255) return BAND_CLS_PACK;
  if (wl) return wl == 16 ? BAND_CLS_HALF : BAND_CLS_WARP;
  if (band_wide_eligible(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right, (int)t.ref_len))
    return BAND_CLS_WIDE;
  {
    const int dpt = band_long_dpt(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right, (int)t.ref_len);
    if (dpt) return dpt == 16 ? BAND_CLS_LONG16 : BAND_CLS_LONG32;
  }
  return band_ring_class(band_ring_need(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                        (int)t.ref_len, false));
}

__device__ __forceinline__ void k3_task_of(const BlockArgs &a, const DCand &d, const smb_block_read &rd, uint64_t read_off,
                                           uint32_t qlen, smb_band_task &t) {
  int bw = d.band_r - d.band_l, band_l, band_r;      // alignRMAPCANDFull, rmap.c:888-896
  if (bw < rd.bandwidth_min) {
    bw = (rd.bandwidth_min - bw + 1) / 2;
    band_l = d.band_l - bw;
    band_r = d.band_r + bw;
  } else {
    band_l = d.band_l;
    band_r = d.band_r;
  }
  t.read_off = read_off; t.ref_off = d.refoff; t.read_len = qlen; t.ref_len = d.reflen;
  t.flags = SMB_TASK_REF_PACKED | (d.rev ? SMB_TASK_READ_REVCOMP : 0u);
  t.l_edge = band_l; t.r_edge = band_r; t.p_left = (int)d.qs; t.p_right = (int)d.qe;
  t.u_left = 0; t.u_right = (int)d.reflen - 1;
  t.minscore = rd.min_swatscor; t.minscorlen = rd.scorlen_min;
}

// one thread per job: scoreRMAPCAND's bookkeeping (rmap.c:745-786) + thresholds (rmap.c:1373-1400)
__global__ void __launch_bounds__(64) block_replay_kernel(const BlockArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.njobs) return;
  smb_block_read rd = a.rd[j];
  a.nk3[j] = 0;
  if (rd.errcode || !rd.reached_stats) return;
  const smb_block_job jb = a.jobs[j];
  const uint32_t qlen = a.seed.read_len[jb.seed_read];
  const unsigned long long c0 = a.cand_first[j];
  const uint32_t ncand = rd.ncand;
  const short mmscordiff = (short)(a.match - a.mismatch);
  const uint32_t cdfs[2] = {a.cover_deficit[2 * j], a.cover_deficit[2 * j + 1]};
  uint32_t max_cover = 0, min_cover = 0;
  int max1 = 0, max2 = 0;
  uint32_t c = 0;
  unsigned long long cells_ref = 0, tasks_ref = 0;
  bool exceed = false;
  for (; c < ncand; ++c) {
    const DCand d = a.dc[c0 + c];
    const int e = a.serr[c0 + c];
    const int sc = a.score[c0 + c];
    if (e) {
      if (e == SMB_ERRCODE_SWATEXCEED && d.simd) exceed = true;   // the host runs the K2' fallback and this kernel again
      rd.errcode = e;
      break;
    }
    if (d.simd) { cells_ref += (unsigned long long)qlen * d.reflen; ++tasks_ref; }
    const uint32_t cdf = cdfs[d.rev ? 1 : 0];
    if (a.prm.best && (d.cover + cdf < min_cover)) break;
    if (sc > max2) {
      if (sc > max1) {
        max2 = max1;
        max1 = sc;
        if (d.cover + cdf > max_cover) max_cover = (d.cover > cdf) ? d.cover - cdf : 0u;
      } else {
        max2 = sc;
      }
      const uint32_t dcov = (uint32_t)(((int)((max1 - max2) / mmscordiff) + 1) * a.nskip);
      if (dcov + cdf + min_cover < max_cover) min_cover = max_cover - dcov;
    }
  }
  if (exceed) { atomicAdd(&a.cnt->n_exceed, 1u); return; }   // rd stays as block_cands_kernel left it
  rd.nscored = c;
  rd.max1scor = max1;
  rd.max2scor = max2;
  const int max_possible = (int)(qlen * (uint32_t)a.match);
  if (!rd.errcode && max1 > max_possible) rd.errcode = ERR_ASSERT;
  if (rd.errcode || max1 < 1) { a.rd[j] = rd; atomicAdd(&a.cnt->k2_cells_ref, cells_ref); atomicAdd(&a.cnt->k2_tasks_ref, tasks_ref); return; }
  rd.do_align = 1;
  int scorlen_min = a.ktup + a.nskip;
  rd.bandwidth_min = (max_possible - max1) / a.gap_ext;     // (-1 * gapextscor, the penalty as a positive cost)
  int below = a.prm.min_swatscor_below_max, min_swatscor = jb.min_swatscor;
  if (below >= max1) below = max1;
  if (min_swatscor > max2 && max2 > 0) min_swatscor = max2;
  if (below >= 0) {
    const int minswc = (max2 > 0) ? max2 : max1;
    if (a.prm.best) {
      if (minswc > min_swatscor) min_swatscor = minswc;
    } else if (min_swatscor + below < max1) {
      min_swatscor = max1 - below;
      if (min_swatscor > minswc) min_swatscor = minswc;
    }
  }
  if (min_swatscor > scorlen_min * a.match && a.match > 0) scorlen_min = min_swatscor / a.match;
  rd.min_swatscor = min_swatscor;
  rd.scorlen_min = scorlen_min;
  // every scored candidate that passes the INITIAL threshold goes to K3 (rmap.c:833-835)
  uint32_t nk3 = 0;
  unsigned int hist[BLK_K3_BINS];
  for (int b = 0; b < BLK_K3_BINS; ++b) hist[b] = 0;
  unsigned long long dir_words = 0, diff_bytes = 0;
  unsigned int mrows = 0, mread = 0;
  const uint64_t read_off = a.seed.read_off[jb.seed_read];
  for (uint32_t k = 0; k < rd.nscored; ++k) {
    if (a.score[c0 + k] < min_swatscor) continue;
    const DCand d = a.dc[c0 + k];
    smb_band_task t;
    k3_task_of(a, d, rd, read_off, qlen, t);
    const int cls = k3_class(a, t);
    a.k3rank[c0 + k] = (int32_t)nk3++;
    a.k3cls[c0 + k] = (uint8_t)cls;
    ++hist[cls];
    if (cls == BAND_CLS_PACK) { mrows = max(mrows, t.ref_len); mread = max(mread, t.read_len); }
    dir_words += band_dir_words(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right, (int)t.ref_len);
    diff_bytes += 2ull * (t.read_len + t.ref_len) + 72ull;
  }
  rd.nk3 = nk3;
  a.nk3[j] = nk3;
  a.rd[j] = rd;
  for (int b = 0; b < BLK_K3_BINS; ++b)
    if (hist[b]) atomicAdd(&a.cnt->k3_hist[b], hist[b]);
  if (mrows) { atomicMax(&a.cnt->pack_maxrows, mrows); atomicMax(&a.cnt->pack_maxread, mread); }
  if (nk3) { atomicAdd(&a.cnt->dir_words, dir_words); atomicAdd(&a.cnt->diff_bytes, diff_bytes); }
  atomicAdd(&a.cnt->k2_cells_ref, cells_ref);
  atomicAdd(&a.cnt->k2_tasks_ref, tasks_ref);
}

// one thread per dense candidate: the K3 task of every aligned one
__global__ void __launch_bounds__(128) block_emit_k3_kernel(const BlockArgs a, const unsigned long long ncand) {
  const unsigned long long ci = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = ci < ncand && a.k3rank[ci] >= 0;
  int cls = 0;
  unsigned long long k = 0;
  if (live) {
    const uint32_t j = a.dc_job[ci];
    const smb_block_read rd = a.rd[j];
    const smb_block_job jb = a.jobs[j];
    const DCand d = a.dc[ci];
    k = a.k3_first[j] + (unsigned long long)a.k3rank[ci];
    if (a.k3rank[ci] == 0) a.rd[j].k3_first = (uint32_t)a.k3_first[j];
    smb_band_task t;
    k3_task_of(a, d, rd, a.seed.read_off[jb.seed_read], a.seed.read_len[jb.seed_read], t);
    a.bat[k] = t;
    smb_block_cand o;
    o.rs = d.rs; o.sqidx = d.sqidx; o.swscor = a.score[ci]; o.reflen = d.reflen; o.band_l = t.l_edge; o.band_r = t.r_edge;
    o.reverse = d.rev; o.reserved[0] = o.reserved[1] = o.reserved[2] = 0;
    a.k3c[k] = o;
    cls = a.k3cls[ci];
    a.dir_words_arr[k] = (uint32_t)band_dir_words(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left,
                                                  t.u_right, (int)t.ref_len);
    a.diff_cap[k] = t.read_len + t.ref_len + 64u;
    a.diff_stride[k] = 2u * (t.read_len + t.ref_len) + 72u;
  }
  const unsigned int slot = warp_ticket(a.cnt->k3_cursor, cls, live);
  if (live) a.k3_order[a.k3_start[cls] + slot] = (int)k;
}

// ------------------------------------------------------------------------------------
cudaError_t launch_block_seqmask(const BlockArgs &a, cudaStream_t st, int *nlaunch) {
  if (a.njobs <= 0 || !a.seqmask) return cudaSuccess;
  block_seqmask_kernel<<<(2 * a.njobs + 3) / 4, 128, 0, st>>>(a);
  ++*nlaunch;
  return cudaGetLastError();
}
cudaError_t launch_block_reqs(const BlockArgs &a, cudaStream_t st, int *nlaunch) {
  if (a.njobs <= 0) return cudaSuccess;
  block_reqs_kernel<<<(a.njobs + 127) / 128, 128, 0, st>>>(a);
  ++*nlaunch;
  return cudaGetLastError();
}
template <int G>
static cudaError_t launch_block_cands_g(const BlockArgs &a, cudaStream_t st) {
  static std::atomic<unsigned long long> smem_done{0};
  const size_t smem = cw_warp_bytes(G) * CW_WARPS;
  const cudaError_t e = ensure_dyn_smem(block_cands_kernel<G>, (int)smem, smem_done);
  if (e != cudaSuccess) return e;
  const int per_cta = CW_WARPS * (32 / G);
  block_cands_kernel<G><<<(a.njobs + per_cta - 1) / per_cta, CW_WARPS * 32, smem, st>>>(a);
  return cudaGetLastError();
}
// max_lists: the largest number of hit lists (2 x sequences or intervals) of a job of the block; hits_per_job: mean
// number of hits of a job.  Group size: a lane per list, at least four (C2, 32000 jobs: 2 lanes 0.36, 4 lanes 0.34,
// 8 lanes 0.38, 32 lanes 0.44 ms for the candidate stage), and more when the jobs' hits would not fit the shared
// scratch of a smaller group (the restricted passes of paired reads take all seeds: long lists, 2.7 x slower
// with 2 lanes per job than with a warp per job).
cudaError_t launch_block_cands(const BlockArgs &a, int max_lists, double hits_per_job, cudaStream_t st, int *nlaunch) {
  if (a.njobs <= 0) return cudaSuccess;
  ++*nlaunch;
  static const int force = getenv("SMALT_B200_CANDG") ? atoi(getenv("SMALT_B200_CANDG")) : 0;   // (A/B measurements)
  int G = 4;
  while (G < 32 && (G < max_lists || G < force || (double)cw_hits(G) < 2.0 * hits_per_job)) G *= 2;
  if (G <= 4) return launch_block_cands_g<4>(a, st);
  if (G <= 8) return launch_block_cands_g<8>(a, st);
  if (G <= 16) return launch_block_cands_g<16>(a, st);
  return launch_block_cands_g<32>(a, st);
}
cudaError_t launch_block_emit_k2(const BlockArgs &a, unsigned long long ncand, cudaStream_t st, int *nlaunch) {
  if (!ncand) return cudaSuccess;
  block_emit_k2_kernel<<<(unsigned int)((ncand + 127) / 128), 128, 0, st>>>(a, ncand);
  ++*nlaunch;
  return cudaGetLastError();
}
cudaError_t launch_block_exceed(const BlockArgs &a, unsigned long long ncand, cudaStream_t st, int *nlaunch) {
  if (!ncand) return cudaSuccess;
  block_exceed_kernel<<<(unsigned int)((ncand + 127) / 128), 128, 0, st>>>(a, ncand);
  ++*nlaunch;
  return cudaGetLastError();
}
cudaError_t launch_block_replay(const BlockArgs &a, cudaStream_t st, int *nlaunch) {
  if (a.njobs <= 0) return cudaSuccess;
  block_replay_kernel<<<(a.njobs + 63) / 64, 64, 0, st>>>(a);
  ++*nlaunch;
  return cudaGetLastError();
}
cudaError_t launch_block_emit_k3(const BlockArgs &a, unsigned long long ncand, cudaStream_t st, int *nlaunch) {
  if (!ncand) return cudaSuccess;
  block_emit_k3_kernel<<<(unsigned int)((ncand + 127) / 128), 128, 0, st>>>(a, ncand);
  ++*nlaunch;
  return cudaGetLastError();
}

cudaError_t warm_block() {
  cudaFuncAttributes f;
  cudaError_t e = cudaFuncGetAttributes(&f, block_reqs_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_seqmask_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_cands_kernel<4>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_cands_kernel<8>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_cands_kernel<32>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_emit_k2_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_exceed_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_replay_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&f, block_emit_k3_kernel);
  return e;
}

}  // namespace smb
