/* smalt_oracle.h - CPU restatement of the SMALT 0.7.6 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so.  The product path
 * (smalt_b200/) never links, imports or calls it.
 *
 * Parity pinning: every function here is checked against the UNMODIFIED
 * reference compiled from /root/reference/src (oracle/_ref/libsmalt_ref.so,
 * tests/test_oracle_vs_ref.py) and against golden vectors dumped from the
 * reference (tests/golden/).  Citations are file:line in /root/reference/src.
 *
 * Sequences are arrays of 3-bit alphabet codes A0 C1 G2 T3 X4 N5
 * (sequence.c:101, :287-318; only `code & 7` reaches the DP, swsimd.c:725,
 * alignment.c:876).
 */
#ifndef SMALT_ORACLE_H
#define SMALT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes used on this path (elib.h:49-139) */
enum {
  SO_SUCCESS = 0,
  SO_FAILURE = -1,
  SO_NOMEM = 2,
  SO_ARGRANGE = 29,
  SO_SHORTSEQ = 30,
  SO_ALLOCBOUNDARY = 32,
  SO_SWATEXCEED = 41,
  SO_SWATSCOR = 44,
  SO_ASSERT = 47,
  SO_OVERFLOW = 48,
  SO_DIFFSTR = 59,
  SO_HITINFO = 64
};

typedef struct {
  int match, mismatch;   /* +1, -2 (score.c:41-47) */
  int gap_init, gap_ext; /* positive costs 4, 3 (score.c:708-709) */
  signed char S[8][8];   /* substitution matrix over the 3-bit alphabet (score.c:138-173) */
} so_scoring;

/* gapopen/gapext are given as the (negative) penalties of `smalt map -S` */
void so_scoring_init(so_scoring *sc, int match, int mismatch, int gapopen, int gapext);

/* K2: swSIMDAlignStriped (swsimd.c:868-933). */
int so_sw_striped(const so_scoring *sc, const uint8_t *read, int qlen,
		  const uint8_t *ref, int rlen, int *score);

/* band geometry: initALIBAND (alignment.c:310-396) */
typedef struct {
  int band_width, l_edge, r_edge, l_edge_orig, r_edge_orig;
  int s_left, s_left_orig, s_len, s_totlen;
  int q_left, q_left_orig, q_len, q_totlen;
} so_band;
int so_band_init(so_band *b, int l_edge, int r_edge, int q_left, int q_right, int q_len,
		 int s_left, int s_right, int s_len);

/* K2': aliSmiWatInBandFast (alignment.c:1603-1638, :1029-1233). `cells` (may be
 * NULL) receives the number of DP cells visited. */
int so_band_fast(const so_scoring *sc, const uint8_t *read, int qlen,
		 const uint8_t *ref, int rlen, int l_edge, int r_edge,
		 int p_left, int p_right, int u_left, int u_right,
		 int *maxscore, long long *cells);

/* K3: aliSmiWatInBand (alignment.c:1548-1601) incl. recursion, backtrace and
 * diffStrReverse.  Results in discovery (pre-)order: out5[5*i..] = score, qs,
 * qe, rs, re; DiffStr bytes incl. terminating 0 concatenated in diffbuf,
 * per-result byte counts in difflen. */
int so_band_align(const so_scoring *sc, const uint8_t *read, int qlen,
		  const uint8_t *ref, int rlen, int l_edge, int r_edge,
		  int p_left, int p_right, int u_left, int u_right,
		  int minscore, int minscorlen,
		  int maxres, int *nres, int *out5,
		  int maxdiff, uint8_t *diffbuf, int *difflen, long long *cells);

/* diffStrReverse (diffstr.c:850-896) on a 0-terminated reversed diff string;
 * returns the output length (incl. terminator) or a negative error. */
int so_diffstr_reverse(const uint8_t *in, uint8_t *out, int maxout);

/* ------------------------------- K1 -------------------------------------- */

/* The `.smi` hash index (hashidx.c:105-146) as read by hashTableRead
 * (hashidx.c:1257-1366).  typ 0 = perfect, 1 = with collisions (hashidx.h:46-49). */
typedef struct {
  int typ, wordlen, nskip, nbits_key, nbits_lo;
  uint32_t nkeys, npos, nwords, maxpos;
  uint32_t keymod;
  uint64_t wordmask, wordmask_lo, wordmask_hi;
  const uint32_t *idx, *pos, *wordidx, *posidx;
} so_index;

void so_index_setup(so_index *ix, int typ, int wordlen, int nskip, int nbits_key,
		    int nbits_lo, uint32_t npos, uint32_t nwords,
		    const uint32_t *idx, const uint32_t *pos,
		    const uint32_t *wordidx, const uint32_t *posidx);

uint32_t so_hash32mix(uint32_t a);

/* hashTableGetKtupleHits (hashidx.c:1146-1191) */
uint32_t so_lookup(const so_index *ix, uint64_t word, uint32_t *posidx);

/* per read x strand seed table, HashHitInfo (hashhit.c:164-213) */
typedef struct {
  uint32_t qlen, n_seeds, seed_rank;
  uint8_t status;
  uint32_t *posidx, *nhits, *cix, *qoffs; /* SEED arrays, index = seed number */
  uint32_t *sortkey, *sidx;              /* nhitqual_sortkeyp, sidxp */
  uint8_t *qmask, *qbuf;                 /* [qlen] */
  uint32_t *frame_cnt;                   /* countp[nskip] */
  uint32_t *frame_ix;                    /* framep[nskip][..] flattened, stride n_alloc */
  uint32_t n_alloc;
} so_hitinfo;

so_hitinfo *so_hitinfo_create(uint32_t maxlen, int nskip);
void so_hitinfo_delete(so_hitinfo *h);

/* collectHitInfo (hashhit.c:480-657) + hashCollectHitInfoShort (:1007-1080)
 * when is_short, hashCollectHitInfo (:987) otherwise. qual may be NULL. */
int so_collect_hitinfo(so_hitinfo *h, const so_index *ix, const uint8_t *read,
		       const uint8_t *qual, uint32_t qlen, int is_reverse, int is_short,
		       uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq_thresh);

/* sort2UINTarraysByQuickSort (sort.c:233-330) - unstable, tie order observable */
int so_sort2(uint32_t n, uint32_t *key, uint32_t *val);
/* sortUINT64arrayByQuickSort (sort.c:415-497) */
int so_sort64(uint32_t n, uint64_t *a);

uint32_t so_cover_deficit(const so_hitinfo *h, int ktup, int nskip);            /* hashhit.c:1096 */
uint32_t so_number_of_hits(const so_hitinfo *h, uint32_t maxhit_per_tuple);     /* hashhit.c:1171 */
uint32_t so_hit_numbers(const so_hitinfo *h, uint32_t *nhit_rank);              /* hashhit.c:1200 */

/* HashHitList (hashhit.c:215-236) */
typedef struct {
  int nhits, nhits_max, nhits_alloc;
  uint8_t status;
  uint64_t *sqdat;
  uint8_t *qmask;
  uint32_t qlen;
} so_hitlist;
so_hitlist *so_hitlist_create(int maxnhits);
void so_hitlist_delete(so_hitlist *l);

/* hashCollectHitsForSegment (hashhit.c:1691-1769); lo/hi are base offsets in
 * the concatenated reference (soffs[s], soffs[s+1], rmap.c:296-308). */
int so_collect_hits_segment(so_hitlist *l, so_hitinfo *h, const so_index *ix,
			    uint64_t lo, uint64_t hi, uint32_t nhit_max, int use_short);
/* hashCollectHitsUsingCutoff (hashhit.c:1593-1689) */
int so_collect_hits_cutoff(so_hitlist *l, const so_hitinfo *h, const so_index *ix,
			   uint32_t max_nhit_per_tup);

/* ------------------- candidate selection (segment.c) and score replay (rmap.c) ----------- */
/* see smalt_oracle_cand.c */
typedef struct so_cands_ so_cands;
typedef struct {
  uint32_t qs, qe;       /* read segment in the profiled orientation */
  uint64_t rs, re;       /* window inside reference sequence sqidx */
  int32_t band_l, band_r;
  uint32_t dqo;
  int32_t dro;
  int32_t sqidx;
  uint32_t cover;
  uint8_t flags;         /* SEGCANDFLG_* (segment.h:51-55) */
} so_cand;
so_cands *so_cands_create(void);
void so_cands_delete(so_cands *c);
void so_cands_blank(so_cands *c);
int so_cands_add_list(so_cands *c, const uint64_t *sqdat, int nhits, int is_reverse, uint32_t qlen,
		      int ktup, int nskip, const uint8_t *qmask, uint32_t min_ktup, uint32_t mincover,
		      int seqidx);
int so_cands_stats(so_cands *c, uint32_t min_cover_below_max, uint32_t cover_deficit_f,
		   uint32_t cover_deficit_r, int target_depth, int max_depth, int is_sensitive);
uint32_t so_cands_count(const so_cands *c, uint32_t *max_cover, uint32_t *max2nd_cover, uint32_t *n_mincover,
			uint32_t *n_all);
int so_cands_offsets(const so_cands *c, uint32_t scidx, int edgelen, uint32_t qlen, const uint64_t *soffs,
		     int nseq, int termchar, so_cand *out);
int so_score_replay(int ncand, const uint32_t *cover, const uint8_t *rev, const int32_t *score,
		    const int32_t *cand_band_l, const int32_t *cand_band_r,
		    const uint32_t cover_deficit[2], uint32_t qlen, int ktup, int nskip, int matchscor,
		    int mismatchscor, int gapinitscor, int gapextscor, int min_swatscor, int min_swatscor_below_max,
		    int best, int *nscored, int *max1scor, int *max2scor, int *min_swatscor_out,
		    int *scorlen_min_out, int *bandwidth_min_out, uint8_t *align, int32_t *band_l, int32_t *band_r);

/* smalt_oracle_cigar.c: CIGAR text + edit distance of an alignment string (diffstr.c:298-367, :1496-1510) */
int so_cigar(const unsigned char *diffstr, int clip_start, int clip_end, int flags, char *out, int maxout, int *nm);

#ifdef __cplusplus
}
#endif
#endif
