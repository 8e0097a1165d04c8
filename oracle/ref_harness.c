/* ref_harness.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C-ABI shims around the UNMODIFIED reference objects compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/.  They let the
 * Python tests (ctypes) call the reference's own hot-path functions one call
 * at a time with plain arrays, so that
 *   (1) the C restatement in oracle/smalt_oracle.c can be pinned against the
 *       real reference, and
 *   (2) the CUDA path can be compared with the real reference on the GPU box
 *       (oracle/_ref/ travels with the snapshot, /root/reference does not).
 *
 * Sequences cross this boundary as 3-bit alphabet codes (A0 C1 G2 T3 X4 N5,
 * sequence.c:101), one byte per base; they are converted to the letters
 * "ACGTXN" and pushed through the reference's own codec (seqFastqSetAscii +
 * seqFastqEncode) so the reference sees exactly what `smalt map` would feed it.
 *
 * Reference entry points exercised:
 *   swSIMDAlignStriped       swsimd.c:868
 *   aliSmiWatInBandFast      alignment.c:1603
 *   aliSmiWatInBand          alignment.c:1548  (+ aliRsltSetFetchData :1518)
 *   hashTableRead            hashidx.c:1257,  seqSetReadBinFil sequence.c
 *   hashTableGetKtupleHits   hashidx.c:1146
 *   hashCollectHitInfoShort  hashhit.c:1007 / hashCollectHitInfo :987
 *   hashCollectHitsForSegment hashhit.c:1691, hashCollectHitsUsingCutoff :1593
 *   hashCalcHitInfoCoverDeficit :1096, hashHitInfoCalcHitNumbers :1200,
 *   hashCalcHitInfoNumberOfHits :1171
 *   seqSetFetchSegmentBySequence sequence.c:2741
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "elib.h"
#include "sequence.h"
#include "score.h"
#include "alibuffer.h"
#include "alignment.h"
#include "swsimd.h"
#include "diffstr.h"
#include "hashidx.h"
#include "hashhit.h"

static SeqCodec *g_codec;
static ScorePenalties *g_pen;
static ScoreMatrix *g_mtx;
static ScoreProfile *g_prof;
static AliBuffer *g_abuf;
static AliRsltSet *g_rset;
static SeqFastq *g_read, *g_ref, *g_tmp;
/* reads as smalt.c:744 creates them (SEQTYP_UNKNOWN): the quality container only exists
 * once qualities were loaded, so FASTA-like and FASTQ-like reads use separate objects */
static SeqFastq *g_read_q, *g_read_nq;
static HashTable *g_ht;
static SeqSet *g_ss;
static HashHitInfo *g_hhi[2];
static HashHitList *g_hhl;

static const char LETTERS[] = "ACGTXN";

static char *codes_to_ascii(const unsigned char *codes, int n)
{
  char *s = (char *) malloc((size_t) n + 1);
  int i;
  for (i = 0; i < n; i++) s[i] = LETTERS[codes[i] > 5 ? 5 : codes[i]];
  s[n] = '\0';
  return s;
}

static int load_seq(SeqFastq *sq, const unsigned char *codes, int n, const char *qual)
{
  int errcode;
  char *s = codes_to_ascii(codes, n);
  seqFastqBlank(sq); /* resets the code to ASCII (setSeq, sequence.c:780, does not) */
  errcode = seqFastqSetAscii(sq, "s", s, qual ? "s" : NULL, qual);
  free(s);
  if (errcode) return errcode;
  return seqFastqEncode(sq, g_codec);
}

int refh_init(int match, int mismatch, int gapopen, int gapext)
{
  int errcode = 0;
  if (g_codec) return 0;
  g_codec = seqCodecCreate();
  g_pen = scorePenaltiesCreate();
  if (!g_codec || !g_pen) return ERRCODE_NOMEM;
  if ((errcode = scoreSetPenalty(g_pen, SCORPNLTYP_MATCH, (short) match)) ||
      (errcode = scoreSetPenalty(g_pen, SCORPNLTYP_MISMATCH, (short) mismatch)) ||
      (errcode = scoreSetPenalty(g_pen, SCORPNLTYP_GAPOPEN, (short) gapopen)) ||
      (errcode = scoreSetPenalty(g_pen, SCORPNLTYP_GAPEXT, (short) gapext)))
    return errcode;
  g_mtx = scoreCreateMatrix(g_codec, g_pen);
  g_prof = scoreCreateProfile(0, g_codec,
			      SCORPROF_SCALAR | SCORPROF_STRIPED_8 | SCORPROF_STRIPED_16);
  g_abuf = aliBufferCreate(0);
  g_rset = aliRsltSetCreate(NULL, 0, 0, 0, 0);
  g_read = seqFastqCreate(0, SEQTYP_FASTQ);
  g_ref = seqFastqCreate(0, SEQTYP_FASTQ);
  g_tmp = seqFastqCreate(0, SEQTYP_FASTQ);
  g_read_q = seqFastqCreate(0, SEQTYP_UNKNOWN);
  g_read_nq = seqFastqCreate(0, SEQTYP_UNKNOWN);
  if (!g_mtx || !g_prof || !g_abuf || !g_rset || !g_read || !g_ref || !g_tmp)
    return ERRCODE_NOMEM;
  return 0;
}

static int setup_pair(const unsigned char *read, int qlen,
		      const unsigned char *ref, int rlen)
{
  int errcode;
  if (!g_codec) return ERRCODE_ASSERT;
  if ((errcode = load_seq(g_read, read, qlen, NULL))) return errcode;
  if ((errcode = load_seq(g_ref, ref, rlen, NULL))) return errcode;
  if ((errcode = scoreMakeProfileFromSequence(g_prof, g_read, g_mtx))) return errcode;
  return aliBufferInit(g_abuf, (unsigned int) qlen);
}

/* swSIMDAlignStriped (swsimd.c:868).  Returns the reference's error code. */
int refh_sw_striped(const unsigned char *read, int qlen,
		    const unsigned char *ref, int rlen, int *score)
{
  int errcode;
  SEQLEN_t l;
  const char *refp;
  if ((errcode = setup_pair(read, qlen, ref, rlen))) return errcode;
  refp = seqFastqGetConstSequence(g_ref, &l, NULL);
  return swSIMDAlignStriped(score, g_abuf, g_prof, refp, (int) l);
}

/* aliSmiWatInBandFast (alignment.c:1603). */
int refh_band_fast(const unsigned char *read, int qlen,
		   const unsigned char *ref, int rlen,
		   int l_edge, int r_edge, int p_left, int p_right,
		   int u_left, int u_right, int *score)
{
  int errcode;
  SEQLEN_t l;
  const char *refp;
  if ((errcode = setup_pair(read, qlen, ref, rlen))) return errcode;
  refp = seqFastqGetConstSequence(g_ref, &l, NULL);
  *score = 0;
  return aliSmiWatInBandFast(score, g_abuf, g_prof, refp, (int) l,
			     l_edge, r_edge, p_left, p_right, u_left, u_right);
}

/* aliSmiWatInBand (alignment.c:1548).  Results are returned as 5 ints each
 * (score, qs, qe, rs, re) in out5[] and the DiffStr bytes (including the
 * terminating 0) concatenated in diffbuf with per-result lengths in difflen[]. */
int refh_band_align(const unsigned char *read, int qlen,
		    const unsigned char *ref, int rlen,
		    int l_edge, int r_edge, int p_left, int p_right,
		    int u_left, int u_right, int minscore, int minscorlen,
		    int maxres, int *nres, int *out5,
		    int maxdiff, unsigned char *diffbuf, int *difflen)
{
  int errcode, n, i, used = 0;
  SEQLEN_t l;
  const char *refp;
  *nres = 0;
  if ((errcode = setup_pair(read, qlen, ref, rlen))) return errcode;
  refp = seqFastqGetConstSequence(g_ref, &l, NULL);
  aliRsltSetReset(g_rset);
  errcode = aliSmiWatInBand(g_rset, g_abuf, g_prof, refp, (int) l,
			    l_edge, r_edge, p_left, p_right, u_left, u_right,
			    minscore, minscorlen);
  if (errcode) return errcode;
  n = aliRsltSetGetSize(g_rset);
  for (i = 0; i < n && i < maxres; i++) {
    const DiffStr *dfs;
    int *o = out5 + 5 * i;
    aliRsltSetFetchData(g_rset, (short) i, o, o + 1, o + 2, o + 3, o + 4, &dfs);
    if (used + dfs->len > maxdiff) return ERRCODE_OVERFLOW;
    memcpy(diffbuf + used, dfs->dstrp, (size_t) dfs->len);
    difflen[i] = dfs->len;
    used += dfs->len;
  }
  *nres = n;
  return 0;
}

/* ---------------------------- K1: index + seeds ------------------------- */

int refh_index_load(const char *prefix)
{
  int errcode = 0;
  if (!g_codec) return ERRCODE_ASSERT;
  if (g_ht) { hashTableDelete(g_ht); g_ht = NULL; }
  if (g_ss) { seqSetDelete(g_ss); g_ss = NULL; }
  if (g_hhi[0]) { hashDeleteHitInfo(g_hhi[0]); hashDeleteHitInfo(g_hhi[1]); g_hhi[0] = g_hhi[1] = NULL; }
  if (g_hhl) { hashDeleteHitList(g_hhl); g_hhl = NULL; }
  g_ss = seqSetReadBinFil(&errcode, prefix);
  if (errcode) return errcode;
  g_ht = hashTableRead(&errcode, prefix);
  if (errcode) return errcode;
  g_hhi[0] = hashCreateHitInfo(0, g_ht);
  g_hhi[1] = hashCreateHitInfo(0, g_ht);
  g_hhl = hashCreateHitList(16384); /* HASH_MAXNHITS rmap.c:50, rmap.c:1121 */
  return (g_hhi[0] && g_hhi[1] && g_hhl) ? 0 : ERRCODE_NOMEM;
}

int refh_index_params(int *ktup, int *nskip, long long *nseq, long long *totlen)
{
  uint8_t ns;
  SETSIZ_t tl;
  if (!g_ht) return ERRCODE_ASSERT;
  *ktup = hashTableGetKtupLen(g_ht, &ns);
  *nskip = ns;
  *nseq = seqSetGetSeqNumAndTotLen(&tl, g_ss);
  *totlen = (long long) tl;
  return 0;
}

/* hashTableGetKtupleHits (hashidx.c:1146) for a batch of words. */
int refh_lookup(const uint64_t *words, int n, uint32_t *nhits, uint32_t *posidx)
{
  int i;
  if (!g_ht) return ERRCODE_ASSERT;
  for (i = 0; i < n; i++) {
    HASHNUM_t px = 0;
    nhits[i] = hashTableGetKtupleHits(NULL, &px, g_ht, words[i]);
    posidx[i] = px;
  }
  return 0;
}

/* Reference window fetch (what makeRMAPCANDfromSegment does, rmap.c:535):
 * returns 3-bit codes of sequence `seqidx`, [offs, offs+len). */
int refh_fetch(long long seqidx, unsigned int offs, unsigned int len,
	       unsigned char *codes, unsigned int *outlen)
{
  int errcode;
  SEQLEN_t l, i;
  char cod;
  const char *p;
  if (!g_ss) return ERRCODE_ASSERT;
  if ((errcode = seqSetFetchSegmentBySequence(g_tmp, seqidx, offs, len, g_ss, g_codec)))
    return errcode;
  p = seqFastqGetConstSequence(g_tmp, &l, &cod);
  if (cod == SEQCOD_ASCII) {
    if ((errcode = seqFastqEncode(g_tmp, g_codec))) return errcode;
    p = seqFastqGetConstSequence(g_tmp, &l, &cod);
  }
  for (i = 0; i < l; i++) codes[i] = (unsigned char) (p[i] & SEQCOD_ALPHA_MASK);
  *outlen = l;
  return 0;
}

/* Accessors into the opaque HashHitInfo live in ref_harness_hashhit.c (which
 * compiles hashhit.c in the same translation unit to see the private struct). */
extern int refh_hitinfo_dump(const HashHitInfo *hip, int maxn, uint32_t *n_seeds,
			     uint32_t *seed_rank, uint32_t *posidx, uint32_t *nhits,
			     uint32_t *qoffs, uint32_t *sortkey, uint32_t *sidx,
			     unsigned char *qmask, unsigned char *status);

/* hashCollectHitInfoShort (hashhit.c:1007) when `is_short` else
 * hashCollectHitInfo (:987) on the whole read. */
int refh_hitinfo(const unsigned char *read, int qlen, const char *qual,
		 int is_reverse, int is_short,
		 unsigned int maxhit_per_tuple, unsigned int maxhit_total,
		 int basq_thresh, int maxn,
		 uint32_t *n_seeds, uint32_t *seed_rank,
		 uint32_t *posidx, uint32_t *nhits, uint32_t *qoffs,
		 uint32_t *sortkey, uint32_t *sidx, unsigned char *qmask,
		 unsigned char *status, uint32_t *cover_deficit,
		 uint32_t *nhit_rank, uint32_t *nhit_tot, uint32_t *nhit_all)
{
  int errcode;
  HashHitInfo *hip;
  if (!g_ht) return ERRCODE_ASSERT;
  SeqFastq *rd = qual ? g_read_q : g_read_nq;
  hip = g_hhi[is_reverse ? 1 : 0];
  if ((errcode = load_seq(rd, read, qlen, qual))) return errcode;
  if (is_short)
    errcode = hashCollectHitInfoShort(hip, (unsigned char) is_reverse, maxhit_per_tuple,
				      maxhit_total, (unsigned char) basq_thresh, rd, g_ht);
  else
    errcode = hashCollectHitInfo(hip, (unsigned char) is_reverse,
				 (unsigned char) basq_thresh, 0, 0, rd, g_ht);
  if (errcode) return errcode;
  *cover_deficit = hashCalcHitInfoCoverDeficit(hip);
  *nhit_tot = hashHitInfoCalcHitNumbers(hip, nhit_rank);
  *nhit_all = hashCalcHitInfoNumberOfHits(hip, maxhit_per_tuple);
  return refh_hitinfo_dump(hip, maxn, n_seeds, seed_rank, posidx, nhits, qoffs,
			   sortkey, sidx, qmask, status);
}

/* Hit list for the HashHitInfo filled by the last refh_hitinfo() call of that
 * strand.  seqidx >= 0: hashCollectHitsForSegment (hashhit.c:1691) restricted
 * to that reference sequence as in collectHits SEQBYSEQ (rmap.c:273);
 * seqidx < 0: hashCollectHitsUsingCutoff (hashhit.c:1593). */
int refh_hitlist(int is_reverse, long long seqidx, unsigned int maxhit_per_tuple,
		 int use_short, int maxhits, int *nhits, uint64_t *sqdat,
		 int maxq, char *qmask_out)
{
  int errcode, n, i;
  const uint64_t *dat;
  uint32_t qlen;
  const char *qmask;
  HashHitInfo *hip;
  if (!g_ht) return ERRCODE_ASSERT;
  hip = g_hhi[is_reverse ? 1 : 0];
  if (seqidx >= 0) {
    const SETSIZ_t *soffsp;
    const SEQNUM_t nseq = seqSetGetOffsets(g_ss, &soffsp);
    if (seqidx >= nseq) return ERRCODE_ARGRANGE;
    hashBlankHitList(g_hhl);
    errcode = hashCollectHitsForSegment(g_hhl, soffsp[seqidx], soffsp[seqidx + 1],
					maxhit_per_tuple, (unsigned char) use_short,
					hip, g_ht, NULL);
  } else {
    errcode = hashCollectHitsUsingCutoff(g_hhl, maxhit_per_tuple, g_ht, hip);
  }
  if (errcode) return errcode;
  dat = hashGetHitListData(&n, NULL, &qlen, NULL, NULL, &qmask, g_hhl);
  *nhits = n;
  for (i = 0; i < n && i < maxhits; i++) sqdat[i] = dat[i];
  for (i = 0; i < (int) qlen && i < maxq; i++) qmask_out[i] = qmask[i];
  return 0;
}

/* ---------------------- candidate selection (segment.c) ------------------ */
#include "segment.h"

static SegLst *g_sgl;
static SegAliCands *g_sac;
static SegQMask *g_qm;

/* What mapSingleRead does between the seed tables and the scoring (rmap.c:1258-1337 with
 * fillRMAPBUFF/collectHits, rmap.c:273-318, in the sequence-by-sequence mode): for both strands
 * and every reference sequence hashCollectHitsForSegment -> segLstFillHits -> segAliCandsAddFast,
 * then segAliCandsStats and, for every selected candidate, segAliCandsCalcSegmentOffsets with
 * edgelen 0 (makeRMAPCANDfromSegment, rmap.c:535-556).  Uses the seed tables left by the last
 * refh_hitinfo() calls of the two strands (same read).  `out` rows: qs qe rs re band_l band_r dqo
 * dro sqidx flags cover (11 x int64). */
int refh_candidates(unsigned int maxhit_per_tuple, unsigned int min_cover, int min_swatscor_below_max,
		    int mismatchdiff, int is_best, int target_depth, int max_depth, int is_sensitive,
		    unsigned int qlen, int maxcand, unsigned int *stats /* n_sort n_mincover max_cover
		    max2nd_cover cdF cdR nhit nhit_tot */, long long *out)
{
  int errcode = 0, st;
  uint8_t nskip;
  const uint8_t ktup = hashTableGetKtupLen(g_ht, &nskip);
  const SETSIZ_t *soffsp;
  const SEQNUM_t nseq = seqSetGetOffsets(g_ss, &soffsp);
  uint32_t min_ktup, mincov_below_max, n, c, r0, r1;
  SEQNUM_t s;
  if (!g_ht) return ERRCODE_ASSERT;
  if (!g_sgl) {
    g_sgl = segLstCreate(0);
    g_sac = segAliCandsCreate(0);
    g_qm = segQMaskCreate(0);
    if (!g_sgl || !g_sac || !g_qm) return ERRCODE_NOMEM;
  }
  /* calcMinKtup, rmap.c:240-247 */
  min_ktup = (min_cover >= (uint32_t) ktup + nskip) ? (min_cover - ktup) / nskip : 1;
  min_cover = (min_ktup - 1) * nskip + ktup;
  if (min_swatscor_below_max < 0) mincov_below_max = qlen - 1;
  else {
    mincov_below_max = ((uint32_t) (min_swatscor_below_max / mismatchdiff)) * nskip;
    if (mincov_below_max < ktup || is_best) mincov_below_max = ktup + 2 * (nskip - 1);
  }
  segAliCandsBlank(g_sac);
  for (st = 0; st < 2 && !errcode; st++)
    for (s = 0; s < nseq; s++) {
      hashBlankHitList(g_hhl);
      if ((errcode = hashCollectHitsForSegment(g_hhl, soffsp[s], soffsp[s + 1], maxhit_per_tuple, 1,
					       g_hhi[st], g_ht, NULL)))
	break;
      segLstBlank(g_sgl);
      if ((errcode = segLstFillHits(g_sgl, min_ktup, g_hhl))) break;
      if ((errcode = segAliCandsAddFast(g_sac, g_qm, g_sgl, min_cover, s))) break;
    }
  if (errcode) return errcode;
  if ((errcode = segAliCandsStats(g_sac, mincov_below_max, g_hhi[0], g_hhi[1], (SEGNUM_t) (short) target_depth,
				  (SEGNUM_t) (short) max_depth, (uint8_t) is_sensitive)))
    return errcode;
  n = segAliCandsGetNumberOfSegments(g_sac, &stats[2], &stats[3], &stats[4], &stats[5], &stats[1]);
  stats[0] = n;
  r0 = hashHitInfoCalcHitNumbers(g_hhi[0], &c);
  r1 = hashHitInfoCalcHitNumbers(g_hhi[1], &stats[6]);
  stats[6] += c;
  stats[7] = r0 + r1;
  for (c = 0; c < n && (int) c < maxcand; c++) {
    SEQLEN_t qs, qe, dqo;
    SETSIZ_t rs, re;
    int bl, br, dro;
    SEQNUM_t sx;
    SEGBITFLG_t fl;
    SEGCOV_t cov;
    if ((errcode = segAliCandsCalcSegmentOffsets(&qs, &qe, &rs, &re, &bl, &br, &dqo, &dro, &sx, &fl, &cov, 0, qlen,
						 g_ss, c, g_sac)))
      return errcode;
    out[11 * c + 0] = qs; out[11 * c + 1] = qe; out[11 * c + 2] = (long long) rs; out[11 * c + 3] = (long long) re;
    out[11 * c + 4] = bl; out[11 * c + 5] = br; out[11 * c + 6] = dqo; out[11 * c + 7] = dro;
    out[11 * c + 8] = sx; out[11 * c + 9] = fl; out[11 * c + 10] = cov;
  }
  return 0;
}
