/* smalt_oracle_cand.c - CPU restatement of SMALT's candidate selection (segment.c) and of the
 * score / threshold bookkeeping of rmap.c that sits between K1, K2 and K3.
 *
 * TEST INFRASTRUCTURE ONLY (see smalt_oracle.h).  Pinned against the UNMODIFIED reference:
 * tests/test_oracle_cand_vs_ref.py runs the reference's own segLstFillHits / segAliCandsAddFast /
 * segAliCandsStats / segAliCandsCalcSegmentOffsets (oracle/_ref/libsmalt_ref.so, refh_candidates)
 * on the same hit lists and compares every output.  Citations are file:line in
 * /root/reference/src.
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "smalt_oracle.h"

#define HALFBIT 31                       /* HASHHIT_HALFBIT, hashhit.h:67 */
#define HALFMASK 0x7FFFFFFFu             /* HASHHIT_HALFMASK */
#define SOFFSMASK 0xFFFFFFFFull          /* HASHHIT_SOFFSMASK */
#define SHIFTPART(x) ((x) & ~((uint64_t) HALFMASK))

enum { SEGMENTING_DIFFSHIFT = 3, MAXIMUM_DEPTH = 8000, DEFAULT_TARGET_DEPTH = 200,
       EDGE_BAND_FACTOR = 4, MAX_BANDEDGE_2POW = 4 };   /* segment.c:118-143 */
enum { CANDFLG_REVERSE = 1, CANDFLG_MMALI = 4 };         /* segment.h:51-55 */
enum { HITQUAL_NORMHIT = 1 };

typedef struct { uint64_t sqo; int32_t len; } seed_t;                     /* SEED, segment.c:160-194 */
typedef struct { uint32_t idx; int32_t num; } hreg_t;                     /* HITREGION :196-205 */
typedef struct { uint32_t ix; int32_t nseed; uint32_t cover; } segm_t;    /* SEGMENT :207-217 */
typedef struct {                                                          /* SEGCAND :240-266 */
  uint32_t qs, qe, rs, re;
  short shiftoffs, shift2mm, srange;
  uint32_t cover;
  uint8_t flag;
  int32_t nseg;
  uint32_t hregix;
  int32_t seqidx;
} segcand_t;

struct so_cands_ {
  hreg_t *hreg; seed_t *seed; segm_t *segm; uint8_t *mask;
  size_t hreg_n, seed_n, segm_n, list_alloc, mask_alloc;
  segcand_t *cand; size_t ncand, cand_alloc;
  uint32_t *sort_keys, *sort_idx; size_t sort_alloc;
  uint32_t n_sort, n_mincover, max_cover, max2nd_cover, cover_deficit[2];
  uint8_t nskip, ktup;
};

so_cands *so_cands_create(void) { return (so_cands *) calloc(1, sizeof(so_cands)); }

void so_cands_delete(so_cands *c)
{
  if (!c) return;
  free(c->hreg); free(c->seed); free(c->segm); free(c->mask); free(c->cand);
  free(c->sort_keys); free(c->sort_idx);
  free(c);
}

/* segAliCandsBlank, segment.c:1515-1528 */
void so_cands_blank(so_cands *c)
{
  c->ncand = 0; c->n_sort = 0; c->max_cover = c->max2nd_cover = 0; c->n_mincover = 0;
  c->nskip = c->ktup = 0; c->cover_deficit[0] = c->cover_deficit[1] = 0;
}

/* calcSegmentBoundaries, segment.c:635-668 (all quantities 32-bit unsigned as there) */
static void seg_bounds(uint32_t *qs, uint32_t *qe, uint32_t *rs, uint32_t *re, const segm_t *sg,
		       const seed_t *seedr, int ktup, int nskip, int is_reverse)
{
  const seed_t *s0 = seedr + sg->ix, *s1 = s0 + sg->nseed - 1;
  *qs = (uint32_t) (s0->sqo & HALFMASK);
  *qe = (uint32_t) (s1->sqo & HALFMASK) + (uint32_t) s1->len - 1;
  if (is_reverse) {
    *rs = (uint32_t) (((s1->sqo >> HALFBIT) - (s1->sqo & HALFMASK) / (uint32_t) nskip) & SOFFSMASK);
    *rs -= (uint32_t) ((s1->len - ktup) / nskip);
    *re = (uint32_t) (((s0->sqo >> HALFBIT) - (*qs) / (uint32_t) nskip) & SOFFSMASK);
  } else {
    *rs = (uint32_t) (((s0->sqo >> HALFBIT) + (*qs) / (uint32_t) nskip) & SOFFSMASK);
    *re = (uint32_t) (((s1->sqo >> HALFBIT) + (s1->sqo & HALFMASK) / (uint32_t) nskip) & SOFFSMASK);
    *re += (uint32_t) ((s1->len - ktup) / nskip);
  }
}

/* derriveSEGCAND, segment.c:929-1059 */
static int derive_cand(segcand_t *cd, int segix_start, int nseg, segm_t *segmbas, const seed_t *seedr,
		       int ktup, int nskip, uint32_t cover, uint32_t mincover_noindel, uint32_t hregix,
		       int is_reverse)
{
  int n;
  uint32_t qs, qe, rs, re, maxcover;
  uint64_t shift_range;
  uint8_t flag = 0;
  int64_t shift_min, shift_start, diff_shift, shift_2mm;
  segm_t *sg0 = segmbas + segix_start, *sg;
  const uint64_t offbit = ((uint64_t) 1) << (HALFBIT + 1);

  if (sg0->nseed < 0) return SO_ASSERT;
  seg_bounds(&cd->qs, &cd->qe, &cd->rs, &cd->re, sg0, seedr, ktup, nskip, is_reverse);
  sg0->nseed *= -1;
  shift_2mm = shift_min = (int64_t) (seedr[sg0->ix].sqo >> HALFBIT);
  maxcover = sg0->cover;
  sg = sg0 + 1;
  for (n = 1; n < nseg; n++, sg++) {
    if (sg->nseed < 0) return SO_ASSERT;
    seg_bounds(&qs, &qe, &rs, &re, sg, seedr, ktup, nskip, is_reverse);
    if (sg->cover > maxcover) {
      shift_2mm = (int64_t) (seedr[sg->ix].sqo >> HALFBIT);
      maxcover = sg->cover;
    }
    sg->nseed *= -1;
    if (qs < cd->qs) cd->qs = qs;
    if (qe > cd->qe) cd->qe = qe;
    if (rs < cd->rs) cd->rs = rs;
    if (re > cd->re) cd->re = re;
  }
  sg--;
  if (is_reverse) {
    flag |= CANDFLG_REVERSE;
    shift_start = ((int64_t) cd->rs) + (cd->qe - (uint32_t) ktup + 1) / (uint32_t) nskip;
  } else {
    shift_start = (int64_t) (((uint64_t) cd->rs) | offbit) - cd->qs / (uint32_t) nskip;
  }
  shift_range = (uint64_t) (((int64_t) (seedr[sg->ix].sqo >> HALFBIT)) - shift_min);
  diff_shift = shift_min - shift_start;
  if (shift_range > SHRT_MAX) return SO_OVERFLOW;
  if (diff_shift < SHRT_MIN || diff_shift > SHRT_MAX) return SO_OVERFLOW;
  cd->shiftoffs = (short) diff_shift;
  if (maxcover >= mincover_noindel) {
    const int64_t ds_2mm = shift_2mm - shift_start;
    flag |= CANDFLG_MMALI;
    if (ds_2mm < SHRT_MIN || ds_2mm > SHRT_MAX) return SO_OVERFLOW;
    cd->shift2mm = (short) ds_2mm;
  } else {
    cd->shift2mm = 0;
  }
  cd->flag = flag;
  cd->srange = (short) shift_range;
  cd->cover = cover;
  cd->nseg = nseg;
  cd->hregix = hregix;
  cd->seqidx = -1;
  return SO_SUCCESS;
}

static int grow_lists(so_cands *c, size_t nhits, uint32_t qlen)
{
  if (nhits + 1 > c->list_alloc) {
    const size_t na = nhits + nhits / 2 + 64;
    free(c->hreg); free(c->seed); free(c->segm);
    c->hreg = (hreg_t *) malloc(na * sizeof(hreg_t));
    c->seed = (seed_t *) malloc(na * sizeof(seed_t));
    c->segm = (segm_t *) malloc(na * sizeof(segm_t));
    if (!c->hreg || !c->seed || !c->segm) return SO_NOMEM;
    c->list_alloc = na;
  }
  if ((size_t) qlen + 1 > c->mask_alloc) {
    free(c->mask);
    c->mask_alloc = (size_t) qlen + 1024;
    if (!(c->mask = (uint8_t *) malloc(c->mask_alloc))) return SO_NOMEM;
  }
  return SO_SUCCESS;
}

/* One hit list -> candidates: segLstFillHits (segment.c:763-810: defineHitRegions :396-453,
 * makeSeedsFromHits :455-533, makeSegmentsFromSeeds :535-584) followed by segAliCandsAddFast
 * (:1530-1557 -> addCandsFast :1140-1223).  qmask: the 0-terminated HITQUAL string of the hit
 * list (hashGetHitListData), NULL = qlen x HITQUAL_NOHIT as hashCollectHitsForSegment leaves it. */
int so_cands_add_list(so_cands *c, const uint64_t *sqdat, int nhits, int is_reverse, uint32_t qlen,
		      int ktup, int nskip, const uint8_t *qmask, uint32_t min_ktup, uint32_t mincover,
		      int seqidx)
{
  int errcode, i, j;
  uint32_t r, q;

  if ((errcode = grow_lists(c, (size_t) (nhits > 0 ? nhits : 0), qlen))) return errcode;
  c->hreg_n = c->seed_n = c->segm_n = 0;

  /* segment.c:781-788: the minimum number of k-tuples shrinks by one per k-tuple that is not a
   * NORMHIT in the list's mask */
  if (qmask) {
    for (; *qmask; qmask++) {
      if (*qmask == HITQUAL_NORMHIT) continue;
      if (min_ktup < 2) break;
      min_ktup--;
    }
  } else {
    for (q = 0; q < qlen; q++) {
      if (min_ktup < 2) break;
      min_ktup--;
    }
  }

  /* defineHitRegions */
  if (nhits >= 1) {
    uint32_t max_dshift = (uint32_t) (ktup * SEGMENTING_DIFFSHIFT / nskip) & 0xffffu;
    const uint32_t ds = (qlen - (uint32_t) ktup) / (uint32_t) nskip + 1;
    uint64_t dsthresh;
    if (ds < max_dshift) max_dshift = ds & 0xffffu;
    dsthresh = ((uint64_t) max_dshift) << HALFBIT;
    for (i = 0; i < nhits;) {
      for (j = i + 1; j < nhits; j++)
	if (sqdat[j] - sqdat[j - 1] >= dsthresh) break;
      if ((uint32_t) (j - i) >= min_ktup) {
	c->hreg[c->hreg_n].idx = (uint32_t) i;
	c->hreg[c->hreg_n].num = j - i;
	c->hreg_n++;
      }
      i = j;
    }
  }
  /* makeSeedsFromHits */
  for (r = 0; r < c->hreg_n; r++) {
    uint32_t a = c->hreg[r].idx, b, end = a + (uint32_t) c->hreg[r].num;
    c->hreg[r].idx = (uint32_t) c->seed_n;
    while (a < end) {
      seed_t *sd = c->seed + c->seed_n++;
      uint64_t shift;
      uint32_t qoffs, lastq, qo;
      sd->sqo = sqdat[a];
      shift = SHIFTPART(sd->sqo);
      qoffs = (uint32_t) (sd->sqo & HALFMASK);
      lastq = qoffs + (uint32_t) ktup;
      for (b = a + 1; b < end; b++) {
	if (SHIFTPART(sqdat[b]) != shift) break;
	qo = (uint32_t) (sqdat[b] & HALFMASK);
	if (qo > lastq || ((qo - qoffs) % (uint32_t) nskip)) break;
	lastq = qo + (uint32_t) ktup;
      }
      sd->len = (int32_t) (lastq - qoffs);
      a = b;
    }
    c->hreg[r].num = (int32_t) (c->seed_n - c->hreg[r].idx);
  }
  /* makeSegmentsFromSeeds */
  for (r = 0; r < c->hreg_n; r++) {
    uint32_t a = c->hreg[r].idx, b, end = a + (uint32_t) c->hreg[r].num;
    c->hreg[r].idx = (uint32_t) c->segm_n;
    c->hreg[r].num = 0;
    while (a < end) {
      segm_t *sg = c->segm + c->segm_n++;
      const uint64_t shift = SHIFTPART(c->seed[a].sqo);
      const uint32_t qoffs = (uint32_t) (c->seed[a].sqo & HALFMASK);
      c->hreg[r].num++;
      sg->ix = a;
      sg->cover = (uint32_t) c->seed[a].len;
      for (b = a + 1; b < end; b++) {
	if (SHIFTPART(c->seed[b].sqo) != shift ||
	    (((uint32_t) (c->seed[b].sqo & HALFMASK)) - qoffs) % (uint32_t) nskip)
	  break;
	sg->cover += (uint32_t) c->seed[b].len;
      }
      sg->nseed = (int32_t) (b - a);
      a = b;
    }
  }

  /* transferParamFromSegLst, segment.c:1457-1468 */
  if (c->ncand == 0) { c->ktup = (uint8_t) ktup; c->nskip = (uint8_t) nskip; }
  else if (c->ktup != ktup || c->nskip != nskip) return SO_ASSERT;

  /* addCandsFast (mincover_noindel == mincover, segment.c:1547-1553) */
  for (r = 0; r < c->hreg_n; r++) {
    const hreg_t *hr = c->hreg + r;
    segm_t *segbas = c->segm + hr->idx;
    for (i = 0; i < hr->num;) {
      segm_t *sg = segbas + i;
      uint32_t cover, cover_new;
      int l;
      const seed_t *sp;
      /* INIT_COVERAGE_CALC */
      memset(c->mask, 0, qlen);
      for (l = sg->nseed, sp = c->seed + sg->ix; l > 0; l--, sp++) {
	uint8_t *u = c->mask + (sp->sqo & HALFMASK);
	for (q = 0; (int32_t) q < sp->len; q++) u[q] = 1;
      }
      cover = sg->cover;
      sg++;
      for (j = i + 1; j < hr->num; j++, sg++) {
	if (sg->nseed < 0) break;
	/* CALC_COVERAGE */
	cover_new = 0;
	for (l = sg->nseed, sp = c->seed + sg->ix; l > 0; l--, sp++) {
	  uint8_t *u = c->mask + (sp->sqo & HALFMASK);
	  for (q = 0; (int32_t) q < sp->len; q++)
	    if (!u[q]) { cover_new++; u[q] = 1; }
	}
	if ((cover_new << 1) < sg->cover && cover >= mincover) break;
	cover += cover_new;
      }
      if (cover >= mincover) {
	segcand_t *cd;
	if (c->ncand + 1 > c->cand_alloc) {
	  const size_t na = c->cand_alloc * 2 + 256;
	  void *hp = realloc(c->cand, na * sizeof(segcand_t));
	  if (!hp) return SO_NOMEM;
	  c->cand = (segcand_t *) hp;
	  c->cand_alloc = na;
	}
	cd = c->cand + c->ncand++;
	if ((errcode = derive_cand(cd, i, j - i, segbas, c->seed, ktup, nskip, cover, mincover, r, is_reverse)))
	  return errcode;
	cd->seqidx = seqidx;
	if (cover > c->max2nd_cover) {
	  if (cover > c->max_cover) { c->max2nd_cover = c->max_cover; c->max_cover = cover; }
	  else if (cover != c->max_cover) c->max2nd_cover = cover;
	}
      }
      i = j;
    }
  }
  return SO_SUCCESS;
}

/* segAliCandsStats, segment.c:1616-1785 */
int so_cands_stats(so_cands *c, uint32_t min_cover_below_max, uint32_t cover_deficit_f,
		   uint32_t cover_deficit_r, int target_depth_arg, int max_depth_arg, int is_sensitive)
{
  int errcode;
  uint32_t i, j, n_cands = (uint32_t) c->ncand;
  uint32_t target_depth = (uint32_t) target_depth_arg, max_depth = (uint32_t) max_depth_arg; /* SEGNUM_t */
  const uint32_t nskip = c->nskip;
  uint32_t min_cover, cdf = 0, cda[2];
  const segcand_t *scp = c->cand;

  if (max_depth < 1 || max_depth > MAXIMUM_DEPTH) max_depth = MAXIMUM_DEPTH;
  if (target_depth < 1) target_depth = DEFAULT_TARGET_DEPTH;
  if (target_depth > max_depth) target_depth = max_depth;
  min_cover = (min_cover_below_max > c->max_cover) ? 0 : c->max_cover - min_cover_below_max;
  if (min_cover > c->max2nd_cover) { cdf = min_cover - c->max2nd_cover; min_cover = c->max2nd_cover; }
  c->cover_deficit[0] = cover_deficit_f;
  c->cover_deficit[1] = cover_deficit_r;
  for (i = 0; i < 2; i++) {            /* both strands use the FORWARD deficit, segment.c:1674 */
    cda[i] = c->cover_deficit[0];
    cda[i] = (cda[i] > cdf) ? cda[i] - cdf : 0;
  }
  if (n_cands + 1 > c->sort_alloc) {
    free(c->sort_keys); free(c->sort_idx);
    c->sort_alloc = (size_t) n_cands + n_cands / 2 + 64;
    c->sort_keys = (uint32_t *) malloc(c->sort_alloc * sizeof(uint32_t));
    c->sort_idx = (uint32_t *) malloc(c->sort_alloc * sizeof(uint32_t));
    if (!c->sort_keys || !c->sort_idx) return SO_NOMEM;
  }
  for (i = j = 0; i < n_cands; i++) {
    const int is_rev = (scp[i].flag & CANDFLG_REVERSE) ? 1 : 0;
    if (scp[i].cover + cda[is_rev] < min_cover) continue;
    if (scp[i].cover > c->max_cover) return SO_ASSERT;
    c->sort_keys[j] = c->max_cover - scp[i].cover;
    c->sort_idx[j] = i;
    j++;
  }
  if ((errcode = so_sort2(j, c->sort_keys, c->sort_idx))) return errcode;
  c->n_mincover = j;
  if (j > target_depth) {
    const uint32_t maxj = (j < max_depth) ? j : max_depth;
    if (is_sensitive) {
      for (j = target_depth; j < maxj; j++)      /* indexes candr by j, not sort_idx[j]: segment.c:1757-1759 */
	if (c->sort_keys[j] >= cda[(scp[j].flag & CANDFLG_REVERSE) ? 1 : 0]) break;
      for (; j < c->n_mincover && c->sort_keys[j] < nskip; j++);
    } else {
      uint32_t cov = c->sort_keys[j / 2];
      if (cov < nskip) cov = nskip;
      for (j = target_depth; j < maxj && c->sort_keys[j] < cov; j++);
    }
  }
  c->n_sort = j;
  return SO_SUCCESS;
}

/* segAliCandsGetNumberOfSegments, segment.c:1817-1830 */
uint32_t so_cands_count(const so_cands *c, uint32_t *max_cover, uint32_t *max2nd_cover, uint32_t *n_mincover,
			uint32_t *n_all)
{
  if (max_cover) *max_cover = c->max_cover;
  if (max2nd_cover) *max2nd_cover = c->max2nd_cover;
  if (n_mincover) *n_mincover = c->n_mincover;
  if (n_all) *n_all = (uint32_t) c->ncand;
  return c->n_sort;
}

/* segAliCandsCalcSegmentOffsets, segment.c:1861-1985.  soffs[nseq+1]: offsets of the sequences in
 * the concatenated set; termchar: a terminator follows every sequence (seqSetGetSeqDatByIndex,
 * sequence.c:2805-2817). */
int so_cands_offsets(const so_cands *c, uint32_t scidx, int edgelen_arg, uint32_t qlen, const uint64_t *soffs,
		     int nseq, int termchar, so_cand *out)
{
  const short edgelen = (short) edgelen_arg;
  int bl, br, band_offs, ds, q_edge_l, q_edge_r, r_edge_l, r_edge_r, edge_band;
  const int nskip = c->nskip, ktup = c->ktup;
  uint64_t roffs, rlen, rs, re;
  uint32_t qs, qe;
  const segcand_t *sc;

  if (scidx >= c->n_sort) return SO_FAILURE;
  sc = c->cand + c->sort_idx[scidx];
  out->sqidx = sc->seqidx;
  out->flags = sc->flag;
  out->cover = sc->cover;
  if (sc->seqidx < 0 || sc->seqidx >= nseq) { roffs = 0; rlen = soffs[nseq]; }
  else {
    roffs = soffs[sc->seqidx];
    rlen = soffs[sc->seqidx + 1] - soffs[sc->seqidx];
    if (termchar && rlen > 0) rlen--;
    rlen = (uint32_t) rlen;   /* SEQLEN_t */
  }
  rs = ((uint64_t) sc->rs) * (uint64_t) nskip;
  re = ((uint64_t) sc->re) * (uint64_t) nskip + (uint64_t) ktup - 1;
  if (rs < roffs || re < rs) return SO_ASSERT;
  rs -= roffs; re -= roffs;
  if (re >= rlen) return SO_ASSERT;
  if (sc->qe < sc->qs || sc->qs >= qlen) return SO_ASSERT;
  if (sc->flag & CANDFLG_REVERSE) { qs = qlen - sc->qe - 1; qe = qlen - sc->qs - 1; }
  else { qs = sc->qs; qe = sc->qe; }
  edge_band = (int) (qlen - sc->cover) / EDGE_BAND_FACTOR;
  if (edge_band > nskip) {
    if (edge_band > (int) (qlen >> MAX_BANDEDGE_2POW)) edge_band = (int) (qlen >> MAX_BANDEDGE_2POW);
    edge_band -= nskip - 1;
  } else edge_band = 0;
  br = (-sc->shiftoffs + 1) * nskip + edge_band + 1;
  bl = br - (sc->srange + 2) * nskip - 2 * edge_band - 2;
  q_edge_l = (qs >= ((uint32_t) edgelen) && edgelen > 0) ? edgelen : (int) qs;
  q_edge_r = (qe + edgelen + 1 <= qlen && edgelen > 0) ? edgelen : (int) (qlen - qe - 1);
  qs -= (uint32_t) q_edge_l;
  qe += (uint32_t) q_edge_r;
  r_edge_l = q_edge_l + br;
  r_edge_r = q_edge_r - bl;
  if (r_edge_l > 0 && rs < (uint64_t) r_edge_l) { r_edge_l = (int) rs; rs = 0; }
  else rs -= (uint64_t) (int64_t) r_edge_l;
  if (re + (uint64_t) (int64_t) r_edge_r >= rlen) { r_edge_r = (int) (rlen - re - 1); re = rlen - 1; }
  else re += (uint64_t) (int64_t) r_edge_r;
  if (re < rs) return SO_ASSERT;
  band_offs = q_edge_l - r_edge_l;
  ds = sc->shift2mm * nskip + band_offs;
  out->band_l = bl + band_offs + (int) qs;
  out->band_r = br + band_offs + (int) qs;
  if (ds < 0) { out->dqo = qs - (uint32_t) ds; out->dro = 0; }
  else { out->dqo = qs; out->dro = ds; }
  (void) r_edge_r;
  out->qs = qs; out->qe = qe; out->rs = rs; out->re = re;
  return SO_SUCCESS;
}

/* The sequential bookkeeping of scoreRMAPCAND (rmap.c:745-786: which candidates count as scored
 * before the loop breaks - ARRLEN(*csr) = i excludes the candidate the break happens on - and the
 * two best scores) and the threshold logic of mapSingleRead (rmap.c:1373-1400) on given scores.
 * cover / rev / score: per candidate in list order; best = RMAPFLG_BEST.  scorlen_min starts as
 * ktup + nskip (rmap.c:1261).  When max1 >= 1, align[c] = 1 marks the scored candidates that
 * alignRMAPCANDFull hands to aliSmiWatInBand with the INITIAL threshold (rmap.c:833-835) and
 * band_l/band_r receive the widened band of rmap.c:888-896. */
int so_score_replay(int ncand, const uint32_t *cover, const uint8_t *rev, const int32_t *score,
		    const int32_t *cand_band_l, const int32_t *cand_band_r,
		    const uint32_t cover_deficit[2], uint32_t qlen, int ktup, int nskip, int matchscor,
		    int mismatchscor, int gapinitscor, int gapextscor, int min_swatscor, int min_swatscor_below_max,
		    int best, int *nscored, int *max1scor, int *max2scor, int *min_swatscor_out,
		    int *scorlen_min_out, int *bandwidth_min_out, uint8_t *align, int32_t *band_l, int32_t *band_r)
{
  const short mmscordiff = (short) (matchscor - mismatchscor);
  uint32_t max_cover = 0, min_cover = 0, dcov, cdf;
  int max1 = 0, max2 = 0, c, n, scorlen_min = ktup + nskip, bandwidth_min;
  const int max_possible = (int) (qlen * (uint32_t) matchscor);
  (void) gapinitscor;
  *min_swatscor_out = *scorlen_min_out = *bandwidth_min_out = 0;
  if (mmscordiff < 1) return SO_ASSERT;
  for (c = 0; c < ncand; c++) {
    align[c] = 0;
    cdf = cover_deficit[rev[c] ? 1 : 0];
    if (best && (cover[c] + cdf < min_cover)) break;
    if (score[c] > max2) {
      if (score[c] > max1) {
	max2 = max1; max1 = score[c];
	if (cover[c] + cdf > max_cover) max_cover = (cover[c] > cdf) ? cover[c] - cdf : 0;
      } else max2 = score[c];
      dcov = (uint32_t) (((int) ((max1 - max2) / mmscordiff) + 1) * nskip);
      if (dcov + cdf + min_cover < max_cover) min_cover = max_cover - dcov;
    }
  }
  n = c;
  for (; c < ncand; c++) align[c] = 0;
  *nscored = n; *max1scor = max1; *max2scor = max2;
  if (max1 > max_possible) return SO_ASSERT;
  if (max1 < 1) return SO_SUCCESS;
  bandwidth_min = (max_possible - max1) / (-1 * gapextscor);
  if (min_swatscor_below_max >= max1) min_swatscor_below_max = max1;
  if (min_swatscor > max2 && max2 > 0) min_swatscor = max2;
  if (min_swatscor_below_max >= 0) {
    const int minswc = (max2 > 0) ? max2 : max1;
    if (best) {
      if (minswc > min_swatscor) min_swatscor = minswc;
    } else if (min_swatscor + min_swatscor_below_max < max1) {
      min_swatscor = max1 - min_swatscor_below_max;
      if (min_swatscor > minswc) min_swatscor = minswc;
    }
  }
  if (min_swatscor > scorlen_min * matchscor && matchscor > 0) scorlen_min = min_swatscor / matchscor;
  *min_swatscor_out = min_swatscor; *scorlen_min_out = scorlen_min; *bandwidth_min_out = bandwidth_min;
  for (c = 0; c < n; c++) {
    int bw = cand_band_r[c] - cand_band_l[c];
    if (score[c] < min_swatscor) continue;
    align[c] = 1;
    if (bw < bandwidth_min) {
      bw = (bandwidth_min - bw + 1) / 2;
      band_l[c] = cand_band_l[c] - bw; band_r[c] = cand_band_r[c] + bw;
    } else { band_l[c] = cand_band_l[c]; band_r[c] = cand_band_r[c]; }
  }
  return SO_SUCCESS;
}
