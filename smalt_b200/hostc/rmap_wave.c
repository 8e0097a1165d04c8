/* rmap_wave.c - batched ("wave") single-end read mapping for the smalt_b200 driver build.
 *
 * The reference maps one read at a time (rmapSingle, /root/reference/src/rmap.c:1648) and
 * calls the hot-path functions once per read x strand (seed lookup), once per reference
 * sequence (hit list) and once per candidate segment (SW score, banded alignment).  One
 * synchronous GPU call per such function call would serialise the device, so this file
 * replays the same per-read logic in WAVES over a whole block of reads:
 *
 *   wave 1  GPU  K1  seed tables of all reads, both strands        (smb_seed_batch)
 *           GPU  K1  hit lists for every read x strand x sequence  (smb_hits_batch)
 *           host     candidate selection per read: the reference's own segment.c
 *                    (segLstFillHits, segAliCandsAddFast, segAliCandsStats,
 *                    segAliCandsCalcSegmentOffsets) on the GPU hit lists
 *   wave 2  GPU  K2/K2'  score of EVERY candidate of every read    (smb_sw_score_batch,
 *                    smb_band_score_batch) - the reference stops scoring early
 *                    (rmap.c:756-783); the early exits are replayed on the host afterwards,
 *                    so over-computed scores are simply never looked at
 *           host     replay of scoreRMAPCAND (rmap.c:660-786) and of the threshold logic of
 *                    mapSingleRead (rmap.c:1379-1400)
 *   wave 3  GPU  K3  banded alignment + backtrace of every candidate that passes the
 *                    INITIAL score threshold                        (smb_band_align_batch)
 *           host     replay of alignRMAPCANDFull (rmap.c:790-928): in BEST mode the
 *                    threshold rises while results are added (rmap.c:881-885); a K3 task
 *                    computed with the initial threshold contains the result tree of any
 *                    higher threshold as a prefix-closed subtree, which prune_results()
 *                    extracts exactly ("threshold replay")
 *           host     the reference's results.c / report.c produce MAPQ and SAM as before.
 *
 * This translation unit compiles the reference's rmap.c in place (from the read-only tree on
 * the include path; nothing is copied) so that the reference's own RMap buffers and static
 * helpers are reused and rmapCreate/rmapSingle/rmapPair keep working through the shim
 * (one GPU call per function call: slow, but still no CPU hot path).
 */
#include <time.h>
#include <string.h>
#include "rmap.c"
#include "shim.h"
#include "rmap_wave.h"

typedef struct {
  RMAPCAND c;
  COVERAGE cover;
  uint32_t reflen;   /* length of the (clipped) reference window */
  uint64_t refoff;   /* offset of the window in the packed reference */
  int32_t task;      /* index into the K2 / K2' task list */
  int32_t k3task;    /* index into the K3 task list or -1 */
  uint8_t simd;
} WCAND;

/* one mapSingleRead pass (rmap.c:1228-1433) of one read of the block */
typedef struct {
  uint32_t read;         /* index of the read in the block (arena) */
  uint32_t seed_read;    /* index of the read in the seed batch that is on the device */
  uint32_t min_cover;
  int min_swatscor;      /* absolute score threshold argument */
  SeqFastq *readp;
  ResultSet *rsp;        /* results are ADDED to this set */
  int niv;               /* < 0: unrestricted search; else number of intervals (rmap.c:438-493) */
  uint32_t iv_first;     /* first interval in RmapWave.iv */
  uint8_t blank;         /* blank rsp first (rmapSingle) */
} WJOB;
typedef struct { uint64_t lo, hi; int32_t sx; } WIVAL;   /* [lo, hi) in bases of the concatenated set */
typedef int (WAVE_DONEF)(void *user, int job, int errcode, ResultSet *rsp);

typedef struct {
  uint32_t qlen;
  int errcode;           /* error that ends the mapping of this read */
  uint8_t reached_stats; /* mapSingleRead got as far as resultSetAlignmentStats */
  uint8_t do_align;      /* max1scor >= 1 */
  int nseg, nseg_tot;
  uint32_t nhit, nhit_tot;
  uint32_t cand_first, ncand, nscored;
  COVERAGE curr_min_cover, cover_deficit[2];
  SWATSCOR max1scor, max2scor;
  int min_swatscor, scorlen_min, bandwidth_min;
} WREAD;

/* paired mode (rmapPairWave): what is kept of a pair between the passes */
struct PairState_ {
  ResultSet *rs[2];      /* [0] results of the read, [1] of its mate */
  unsigned char status, rare_mate;
  RSLTPAIRFLG_t pairflg;
  int mapq1, swscor1, swscor2_restricted, n_proper, swscor1_2ndbest;
  unsigned char stage;   /* PST_* */
};
/* Growth step of the per-pair result sets (results.c:1781-1817 takes it as a parameter; the
 * default of 4096 results = 112 KB per set is meant for one set per thread): most reads have one
 * or two results, and a block keeps two sets per pair. */
/* SMALT_B200_DEBUG, looked up once (the test sits in per-read and per-candidate loops) */
static int wave_debug(void)
{
  static int state = -1;
  if (state < 0) state = getenv("SMALT_B200_DEBUG") != NULL;
  return state;
}

enum { PAIR_RESULT_BLKSZ = 4 };
enum { PST_END = 0, PST_SECOND_UNRESTRICTED = 3, PST_RESCUE_MAIN = 4, PST_RESCUE_FINE = 5 };

/* page-locked staging buffers (smb_host_alloc): what crosses the C ABI is copied by DMA */
typedef struct { void *p; size_t cap; } WBUF;
enum { WB_ARENA, WB_QUAL, WB_READ_OFF, WB_READ_LEN, WB_INFO, WB_INFO4, WB_REQ, WB_LIST_FIRST, WB_REQ_ERR, WB_SQDAT,
       WB_SWT, WB_SW_SCORE, WB_SW_ERR, WB_BFT, WB_BF_SCORE, WB_BF_ERR, WB_BAT, WB_BA_ERR, WB_RES,
       WB_RES_FIRST, WB_DIFF, WB_BJOB, WB_BIVAL, WB_BRD, WB_BCAND, WB_CIG_BLOB, WB_COUNT };

struct RmapWave_ {
  smb_ctx *ctx;
  WBUF wb[WB_COUNT];
  /* host staging, grown on demand */
  uint8_t *arena, *qual;
  size_t arena_alloc;
  uint64_t *read_off;
  uint32_t *read_len;
  smb_seed_info *info;
  WREAD *rd;
  size_t n_alloc;
  smb_hit_req *req;
  uint64_t *list_first;
  int32_t *req_err;
  size_t req_alloc;
  uint64_t *sqdat;
  size_t sqdat_alloc;
  WCAND *cand;
  size_t cand_alloc, ncand;
  smb_sw_task *swt;
  int32_t *sw_score, *sw_err;
  size_t swt_alloc;
  smb_band_task *bft, *bat;
  int32_t *bf_score, *bf_err, *ba_err;
  size_t bft_alloc, bat_alloc;
  smb_ali_result *res;
  uint32_t *res_first;
  uint8_t *diff;
  size_t res_alloc, diff_alloc;
  /* output stage on the device (csrc/cigar.cu): CIGAR text + NM of every alignment in `res` */
  int cigar_mode;        /* SMB_CIGAR_* flags of the next block run, 0 = off */
  int have_cigar;        /* the arrays below describe the current `res` */
  uint8_t *cig_blob;     /* SMB_CIGAR_BLOB_BYTES: first[], nm[], text */
  const uint32_t *cig_first;
  const int32_t *cig_nm;
  const char *cig_text;
  ScoreProfile *prof, *profRC;
  SeqFastq *readRC;
  WJOB *jobs, *pjob;
  size_t jobs_alloc, pjob_alloc;
  uint64_t n_pairs_pass3, n_pairs_pass4;
  WIVAL *iv;
  size_t iv_alloc, niv;
  int nreads;            /* reads of the seed batch on the device */
  int any_qual;
  smb_seed_info *info4;  /* seed tables against the on-the-fly indexes (pass 4) */
  uint32_t *ftab;        /* arrays of the on-the-fly indexes of a block: idx[nkeys+1], pos[npos] each */
  size_t ftab_alloc, nftab;
  smb_small_index *ftabs;
  size_t *ftab_off;
  uint32_t *ftab_read;
  uint64_t *f_off;
  uint32_t *f_len;
  size_t ftabs_alloc;
  struct PairState_ *ps; /* paired mode: per-pair result sets of the block */
  size_t ps_alloc;
  uint64_t n_pairs, n_pairs_fallback;
  /* statistics */
  double ms_k1, ms_k2, ms_k3;
  double wall[8]; /* host wall seconds: stage, seed, hits, candidates, score, replay, align, results */
  double wall_res[3]; /* inside results: add alignments, sort/MAPQ/filter, emit (report + format) */
  double cpu[8];      /* thread CPU seconds of the same eight stages (wall minus waiting for the GPU) */
  uint64_t cells_k2, cells_k3, n_k2, n_k3, n_reads;
  double ms_cand;     /* part of ms_k1: candidate selection + task lists on the device */
};

static void *wbuf_need(WBUF *b, size_t bytes)
{
  if (bytes > b->cap) {
    smb_host_free(b->p);
    b->cap = bytes + bytes / 2 + 4096;
    if (!(b->p = smb_host_alloc(b->cap))) b->cap = 0;
  }
  return b->p;
}
#define WPIN(ptr, which, need, type)						\
  do { if (!((ptr) = (type *) wbuf_need(&w->wb[which], (size_t) (need) * sizeof(type)))) return ERRCODE_NOMEM; } while (0)

#define WGROW(ptr, alloc, need, type)						\
  do { if ((size_t) (need) > (alloc)) {						\
      size_t na_ = (size_t) (need) + (size_t) (need) / 2 + 64;			\
      void *hp_ = realloc((ptr), na_ * sizeof(type));				\
      if (!hp_) return ERRCODE_NOMEM;						\
      (ptr) = (type *) hp_; (alloc) = na_; } } while (0)

RmapWave *rmapWaveCreate(const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
			 const ScoreMatrix *scormtxp)
{
  RmapWave *w;
  if (smbShimInit(htp, ssp, codecp, scormtxp)) return NULL;
  w = (RmapWave *) calloc(1, sizeof(*w));
  if (!w) return NULL;
  if (smbShimWorkerCtx(&w->ctx, scormtxp)) { free(w); return NULL; }
  w->prof = scoreCreateProfile(0, codecp, SCORPROF_SCALAR);
  w->profRC = scoreCreateProfile(0, codecp, SCORPROF_SCALAR);
  w->readRC = seqFastqCreate(0, SEQTYP_FASTQ);
  if (!w->prof || !w->profRC || !w->readRC) { free(w); return NULL; }
  return w;
}

void rmapWaveDelete(RmapWave *w)
{
  if (!w) return;
  {
    int k;
    for (k = 0; k < WB_COUNT; k++) smb_host_free(w->wb[k].p);
  }
  smb_ctx_destroy(w->ctx);
  free(w->rd); free(w->cand); free(w->jobs); free(w->pjob); free(w->iv);
  free(w->ftab); free(w->ftabs); free(w->ftab_off); free(w->ftab_read); free(w->f_off); free(w->f_len);
  if (w->ps) {
    size_t k;
    for (k = 0; k < w->ps_alloc; k++) { resultSetDelete(w->ps[k].rs[0]); resultSetDelete(w->ps[k].rs[1]); }
    free(w->ps);
  }
  scoreDeleteProfile(w->prof); scoreDeleteProfile(w->profRC); seqFastqDelete(w->readRC);
  free(w);
}

void rmapWaveGetWall(const RmapWave *w, double wall[11]) { memcpy(wall, w->wall, sizeof(w->wall)); memcpy(wall + 8, w->wall_res, sizeof(w->wall_res)); }

void rmapWaveGetStats(const RmapWave *w, double ms[3], uint64_t counts[5])
{
  ms[0] = w->ms_k1; ms[1] = w->ms_k2; ms[2] = w->ms_k3;
  counts[0] = w->n_reads; counts[1] = w->n_k2; counts[2] = w->cells_k2; counts[3] = w->n_k3;
  counts[4] = w->cells_k3;
}

static double wnow(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
/* per-READ stop watch (the shares inside results): five clock reads per read are 8 % of the host time of a
 * 150 bp read, so they only run with SMALT_B200_TIMING */
static int wave_timing(void);
static double rnow(void)
{
  return wave_timing() ? wnow() : 0.0;
}
static double cnow(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define WTICK(k) do { const double t_ = wnow(), c_ = cnow(); w->wall[k] += t_ - tw; tw = t_; \
    w->cpu[k] += c_ - tc; tc = c_; } while (0)
void rmapWaveGetCpu(const RmapWave *w, double cpu[8]) { memcpy(cpu, w->cpu, sizeof(w->cpu)); }
double rmapWaveGetCandMs(const RmapWave *w) { return w->ms_cand; }

static int gpu_fail(ErrMsg *errmsgp, const RmapWave *w, int rc)
{
  fprintf(stderr, "smalt_b200: GPU call failed (%d): %s\n", rc, smb_last_error(w->ctx));
  ERRMSGNO(errmsgp, ERRCODE_FAILURE);
  return ERRCODE_FAILURE;
}

/* The results of a K3 task computed with thresholds (ms0, msl0), in the reference's
 * discovery (pre-)order, contain the results for any (ms >= ms0, msl >= msl0): walk the
 * recursion tree of alignSmiWatBandRecursive (alignment.c:1300-1434) implied by the row
 * ranges and apply the cuts the reference would apply with the higher thresholds. */
static int prune_results(AliRsltSet *out, const smb_ali_result *res, const uint8_t *diff,
			 uint32_t *pos, uint32_t end, int s_left, int s_right,
			 int minscore, int minscorlen, int keep)
{
  int errcode = 0;
  const smb_ali_result *r;
  int s_start, s_end, k;
  if (*pos >= end) return 0;
  r = res + *pos;
  if (r->rs < s_left || r->re > s_right) return 0; /* nothing was found in this row range */
  ++*pos;
  s_start = r->rs; s_end = r->re;
  k = keep && r->score >= minscore && !(r->qs + minscorlen > r->qe + 1);
  if (k && (errcode = smbShimAliRsltSetAdd(out, r->score, r->qs, r->qe, r->rs, r->re,
					   diff + r->diff_off, (int) r->diff_len)))
    return errcode;
  /* children exist in the list iff the initial thresholds explored them; they are kept only
   * if this node survives and the (possibly longer) minimum length still allows the range */
  if ((errcode = prune_results(out, res, diff, pos, end, s_left, s_start - 1, minscore, minscorlen,
			       k && (s_left + minscorlen < s_start))))
    return errcode;
  return prune_results(out, res, diff, pos, end, s_end + 1, s_right, minscore, minscorlen,
		       k && (s_right > s_end + minscorlen));
}

/* wave 1a: seed tables of all reads of a block, both strands (one K1 batch) */
static int wave_seed(ErrMsg *errmsgp, RmapWave *w, int n, SeqFastq **reads, int ktuple_maxhit, UCHAR min_basqval)
{
  int rc, i, any_qual = 0;
  size_t tot = 0;
  double tw = wnow(), tc = cnow();
  w->nreads = 0;
  WPIN(w->read_off, WB_READ_OFF, n, uint64_t);
  WPIN(w->read_len, WB_READ_LEN, n, uint32_t);
  WPIN(w->info, WB_INFO, 2 * (size_t) n, smb_seed_info);
  for (i = 0; i < n; i++) {
    SEQLEN_t len;
    char cod;
    seqFastqGetConstSequence(reads[i], &len, &cod);
    if (cod != SEQCOD_MANGLED) return ERRCODE_SEQCODE;
    w->read_off[i] = tot;
    w->read_len[i] = len;
    tot += len;
    if (seqFastqGetConstQualityFactors(reads[i], NULL, NULL)) any_qual = 1;
  }
  WPIN(w->arena, WB_ARENA, tot + 16, uint8_t);
  WPIN(w->qual, WB_QUAL, tot + 16, uint8_t);
  for (i = 0; i < n; i++) {
    SEQLEN_t len;
    const char *p = seqFastqGetConstSequence(reads[i], &len, NULL);
    const char *q = seqFastqGetConstQualityFactors(reads[i], NULL, NULL);
    memcpy(w->arena + w->read_off[i], p, len);
    if (any_qual) {
      if (q) memcpy(w->qual + w->read_off[i], q, len);
      else memset(w->qual + w->read_off[i], 0xff, len); /* FASTA read: never below the threshold */
    }
  }
  WTICK(0);
  if ((rc = smb_arena_upload(w->ctx, w->arena, tot))) return gpu_fail(errmsgp, w, rc);
  if ((rc = smb_seed_batch(w->ctx, w->read_off, w->read_len, n, any_qual ? w->qual : NULL,
			   (uint32_t) ktuple_maxhit, HASH_MAXNHITS, min_basqval, 1, w->info,
			   NULL, NULL, NULL, NULL, NULL, NULL)))
    return gpu_fail(errmsgp, w, rc);
  w->ms_k1 += smb_last_kernel_ms(w->ctx);
  WTICK(1);
  w->nreads = n;
  w->any_qual = any_qual;
  return ERRCODE_SUCCESS;
}

/* SMALT_B200_HOSTCAND: candidate selection with the reference's segment.c on the host (the first
 * version of this driver) instead of on the device - kept for comparing the two */
static int wave_host_cand(void)
{
  static int state = -1;
  if (state < 0) state = getenv("SMALT_B200_HOSTCAND") != NULL;
  return state;
}

/* SMB_CIGAR_* flags for the blocks of single-end reads (set by the driver from the report writer: SAM output) */
static int g_cigar_mode = 0;
void rmapWaveSetCigarMode(int flags) { g_cigar_mode = flags; }

static int wave_timing(void)
{
  static int state = -1;
  if (state < 0) state = getenv("SMALT_B200_TIMING") != NULL;
  return state;
}

static int wave_skip_results(void)
{
  static int state = -1;
  if (state < 0) state = getenv("SMALT_B200_SKIP_RESULTS") != NULL;
  return state;
}

/* host side of a block whose device outputs are in `bw` (the worker's own wave or the shared wave of a
 * combined batch): replay of alignRMAPCANDFull (rmap.c:820-926) on the alignments of reads first .. first+n-1
 * of the batch, then results.c as in the reference.  `w` = the calling worker's wave (profiles, timers). */
static int dev_results(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, WREAD *rdarr, const RmapWave *bw, const smb_block_read *brd,
		       const smb_block_cand *bc, size_t first, int n, const WJOB *jobs, short max_depth, RMAPFLG_t rmapflg,
		       short matchscor, const SeqSet *ssp, const SeqCodec *codecp, WAVE_DONEF *donef, void *user)
{
  int errcode = ERRCODE_SUCCESS, i;
  SmbCigarSource src;
  RMAPBUFF *bufp = rmp->bfp;
  double tres;
  tres = rnow();
  if (wave_skip_results()) {   /* diagnostic: device + transfer side alone, every read reported unmapped */
    for (i = 0; i < n; i++) {
      if (rdarr) memset(rdarr + i, 0, sizeof(WREAD));
      if (jobs[i].blank) resultSetBlank(jobs[i].rsp);
      if (donef && (errcode = (*donef)(user, i, 0, jobs[i].rsp))) return errcode;
    }
    return ERRCODE_SUCCESS;
  }
  for (i = 0; i < n; i++) {
    WREAD rd_local, *rd = rdarr ? rdarr + i : &rd_local;
    const smb_block_read *b = brd + first + i;
    ResultSet *rsp = jobs[i].rsp;
    SeqFastq *readp = jobs[i].readp;
    memset(rd, 0, sizeof(*rd));
    rd->qlen = bw->read_len[jobs[i].read];
    rd->errcode = b->errcode;
    rd->reached_stats = b->reached_stats;
    rd->do_align = b->do_align;
    rd->nseg = b->nseg; rd->nseg_tot = b->nseg_tot; rd->nhit = b->nhit; rd->nhit_tot = b->nhit_tot;
    rd->ncand = b->ncand; rd->nscored = b->nscored;
    rd->max1scor = b->max1scor; rd->max2scor = b->max2scor;
    rd->min_swatscor = b->min_swatscor; rd->scorlen_min = b->scorlen_min; rd->bandwidth_min = b->bandwidth_min;
    if (jobs[i].blank) resultSetBlank(rsp);
    if (rd->errcode == ERRCODE_SHORTSEQ) { /* too short to be hashed (rmap.c:1273-1275) */
      if (donef && (errcode = (*donef)(user, i, ERRCODE_SHORTSEQ, rsp))) return errcode;
      continue;
    }
    if (rd->errcode) ERRMSGNO(errmsgp, rd->errcode);
    if (rd->reached_stats)
      resultSetAlignmentStats(rsp, rd->nseg, rd->nseg_tot, max_depth, rd->nhit, rd->nhit_tot);
    if (wave_debug())
      fprintf(stderr, "DBG read %d ncand %u nscored %u max1 %d max2 %d min_swatscor %d scorlen_min %d bw_min %d nseg %d/%d nhit %u/%u\n",
	      i, rd->ncand, rd->nscored, rd->max1scor, rd->max2scor, rd->min_swatscor, rd->scorlen_min, rd->bandwidth_min,
	      rd->nseg, rd->nseg_tot, rd->nhit, rd->nhit_tot);
    if (!rd->errcode && rd->do_align) {
      int min_swatscor = rd->min_swatscor;
      SWATSCOR swatscor_2ndmax = 0;
      uint32_t c;
      for (c = 0; c < b->nk3 && !rd->errcode; c++) {
	const size_t t = (size_t) b->k3_first + c;
	const smb_block_cand *cp = bc + t;
	int minscorlen = rd->scorlen_min;
	uint32_t pos, end;
	if (rmapflg & RMAPFLG_BEST) {
	  resultSetGetMaxSwat(rsp, &swatscor_2ndmax);
	  if (swatscor_2ndmax > min_swatscor) min_swatscor = swatscor_2ndmax;
	}
	if (bw->ba_err[t]) { rd->errcode = bw->ba_err[t]; break; }
	aliRsltSetReset(bufp->alirsltp);
	/* aliSmiWatInBand (alignment.c:1569-1575) with the current threshold */
	if (min_swatscor < 1 || matchscor <= 0) { rd->errcode = ERRCODE_ASSERT; break; }
	if (minscorlen * matchscor < min_swatscor) minscorlen = min_swatscor / matchscor;
	if (minscorlen < 5) { rd->errcode = ERRCODE_ASSERT; break; }
	pos = bw->res_first[t];
	end = bw->res_first[t + 1];
	if ((errcode = prune_results(bufp->alirsltp, bw->res, bw->diff, &pos, end, 0, (int) cp->reflen - 1,
				     min_swatscor, minscorlen, 1)))
	  return errcode;
	if (wave_debug()) {
	  short k_, n_ = aliRsltSetGetSize(bufp->alirsltp);
	  fprintf(stderr, "DBG  cand swscor %d rev %d rs %llu band %d %d minscore %d minscorlen %d raw %u kept %d:",
		  cp->swscor, (int) cp->reverse, (unsigned long long) cp->rs, cp->band_l, cp->band_r, min_swatscor,
		  minscorlen, bw->res_first[t + 1] - bw->res_first[t], (int) n_);
	  for (k_ = 0; k_ < n_; k_++) {
	    int sc_, a_, b_, c_, d_;
	    aliRsltSetFetchData(bufp->alirsltp, k_, &sc_, &a_, &b_, &c_, &d_, NULL);
	    fprintf(stderr, " (%d q%d-%d r%d-%d)", sc_, a_, b_, c_, d_);
	  }
	  fputc('\n', stderr);
	}
	errcode = resultSetAddFromAli(rsp, bufp->alirsltp, (SETSIZ_t) cp->rs, 0, rd->qlen, (SEQNUM_t) cp->sqidx,
				      (char) (cp->reverse ? RMAPCANDFLG_REVERSE : 0));
	if (errcode) { rd->errcode = errcode; break; }
      }
      { const double t_ = rnow(); w->wall_res[0] += t_ - tres; tres = t_; }
      if (rd->errcode) ERRMSGNO(errmsgp, rd->errcode);
      else {
	/* (the read's profiles are only dereferenced for results without a sequence index, which the
	 * sequence-by-sequence mode never produces: results.c:1715, :1742-1756) */
	errcode = resultSetSortAndAssignSequence(rsp, bufp->sqbfp, 0, readp, w->prof, w->profRC, ssp, codecp);
	if (errcode) { rd->errcode = errcode; ERRMSGNO(errmsgp, errcode); }
      }
    }
    if (bw->have_cigar && b->nk3) {   /* the output stage's text for the alignments of this read (shim_report.c) */
      /* (`src` lives at function scope: the writer keeps a pointer to it until it is cleared below) */
      src.cands = bc + b->k3_first; src.res_first = bw->res_first + b->k3_first; src.nk3 = b->nk3;
      src.res = bw->res; src.diff = bw->diff;
      src.cig_first = bw->cig_first; src.cig_nm = bw->cig_nm; src.cig_text = bw->cig_text;
      src.qlen = rd->qlen;
      smbShimSetCigarSource(&src);
    }
    errcode = donef ? (*donef)(user, i, rd->errcode, rsp) : ERRCODE_SUCCESS;
    smbShimSetCigarSource(NULL);
    if (errcode) break;
    tres = rnow();
  }
  smbShimCigarFlush();
  return errcode;
}

/* waves 1b-3 with the block resident on the device (smb_block_run / smb_block_fetch): hit lists,
 * candidate selection (segment.c), K2, the score replay (rmap.c:745-786, :1373-1400) and K3 run
 * back to back on the GPU; the host gets the per-read summaries, the aligned candidates and their
 * alignments and replays alignRMAPCANDFull's result handling (rmap.c:881-926) with results.c. */
static int wave_pass_dev(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int n, const WJOB *jobs, const smb_seed_info *info,
			 int ktuple_maxhit, int min_swatscor_below_max_arg, short target_depth, short max_depth,
			 RMAPFLG_t rmapflg, const ScoreMatrix *scormtxp,
			 const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
			 WAVE_DONEF *donef, void *user)
{
  int errcode = ERRCODE_SUCCESS, rc, i;
  UCHAR nskip;
  const UCHAR ktup = hashTableGetKtupLen(htp, &nskip);
  const SETSIZ_t *soffs;
  const SEQNUM_t nseq = seqSetGetOffsets(ssp, &soffs);
  RMAPBUFF *bufp = rmp->bfp;
  short matchscor = 0, mismatchscor = 0, gapinitscor = 0, gapextscor = 0;
  smb_block_job *bj;
  smb_block_ival *biv;
  smb_block_read *brd;
  smb_block_cand *bc;
  smb_block_params prm;
  smb_block_sizes sz;
  double tw = wnow(), tc = cnow(), tres;
  size_t k;

  (void) ktup;
  if (n < 1) return ERRCODE_SUCCESS;
  WGROW(w->rd, w->n_alloc, n, WREAD);
  /* penalties are those of the score matrix, identical for every read (rmap.c:1258-1266) */
  for (i = 0; i < n; i++)
    if (!info[2 * jobs[i].seed_read].err && !info[2 * jobs[i].seed_read + 1].err) break;
  if (i < n) {
    short mismatchdiff;
    if ((errcode = scoreMakeProfileFromSequence(w->prof, jobs[i].readp, scormtxp))) return errcode;
    matchscor = scoreProfileGetAvgPenalties(&mismatchscor, &gapinitscor, &gapextscor, w->prof);
    if ((rc = smbShimSetScoring(w->ctx, w->prof))) return gpu_fail(errmsgp, w, rc);
    mismatchdiff = (short) (matchscor - mismatchscor);
    if (mismatchdiff < 0 || gapextscor >= 0 || mismatchscor >= 0) return ERRCODE_ASSERT;
    if ((short) (matchscor - mismatchscor) < 1 || (short) (matchscor - gapinitscor) < 1) return ERRCODE_ASSERT;
  }
  WPIN(bj, WB_BJOB, n, smb_block_job);
  WPIN(biv, WB_BIVAL, w->niv + 1, smb_block_ival);
  for (i = 0; i < n; i++) {
    const WJOB *jb = jobs + i;
    memset(bj + i, 0, sizeof(*bj));
    bj[i].seed_read = jb->seed_read;
    bj[i].niv = jb->niv < 0 ? -1 : jb->niv;
    bj[i].iv_first = jb->niv < 0 ? 0 : jb->iv_first;
    bj[i].min_cover = jb->min_cover;
    bj[i].min_swatscor = jb->min_swatscor;
  }
  for (k = 0; k < w->niv; k++) {
    biv[k].lo = w->iv[k].lo; biv[k].hi = w->iv[k].hi; biv[k].seqidx = w->iv[k].sx; biv[k].reserved = 0;
  }
  memset(&prm, 0, sizeof(prm));
  prm.nhit_max = (uint32_t) ktuple_maxhit;
  prm.min_swatscor_below_max = min_swatscor_below_max_arg;
  prm.target_depth = (int32_t) target_depth;
  prm.max_depth = (int32_t) max_depth;
  prm.best = (uint8_t) ((rmapflg & RMAPFLG_BEST) != 0);
  prm.sensitive = (uint8_t) ((rmapflg & RMAPFLG_SENSITIVE) != 0);
  prm.cigar = (uint8_t) w->cigar_mode;
  {
    SETSIZ_t roffs;
    const SEQLEN_t rlen0 = nseq > 0 ? seqSetGetSeqDatByIndex(&roffs, NULL, 0, ssp) : 0;
    prm.termchar = (uint8_t) (nseq > 0 && (SETSIZ_t) rlen0 != soffs[1] - soffs[0]);
  }
  WTICK(3);
  if ((rc = smb_block_run(w->ctx, &prm, bj, n, biv, (int) w->niv, &sz))) return gpu_fail(errmsgp, w, rc);
  w->ms_k1 += sz.ms_hits + sz.ms_cand;
  w->ms_cand += sz.ms_cand;
  w->ms_k2 += sz.ms_k2;
  w->ms_k3 += sz.ms_k3;
  w->n_k2 += sz.k2_tasks_ref;
  w->cells_k2 += sz.k2_cells_ref;
  w->n_k3 += sz.nk3;
  w->cells_k3 += sz.k3_cells;
  WPIN(brd, WB_BRD, n, smb_block_read);
  WPIN(bc, WB_BCAND, sz.nk3 + 1, smb_block_cand);
  WPIN(w->ba_err, WB_BA_ERR, sz.nk3 + 1, int32_t);
  WPIN(w->res_first, WB_RES_FIRST, sz.nk3 + 2, uint32_t);
  if (w->res_alloc < sz.nresults + 1) {
    WPIN(w->res, WB_RES, sz.nresults + sz.nresults / 2 + 64, smb_ali_result);
    w->res_alloc = w->wb[WB_RES].cap / sizeof(smb_ali_result);
  }
  if (w->diff_alloc < sz.ndiffbytes + 1) {
    WPIN(w->diff, WB_DIFF, sz.ndiffbytes + sz.ndiffbytes / 2 + 4096, uint8_t);
    w->diff_alloc = w->wb[WB_DIFF].cap;
  }
  w->have_cigar = 0;
  if (prm.cigar) {
    WPIN(w->cig_blob, WB_CIG_BLOB, SMB_CIGAR_BLOB_BYTES(sz.nresults, sz.ncigarbytes) + 64, uint8_t);
    if ((rc = smb_block_fetch_cigar(w->ctx, brd, bc, w->ba_err, w->res_first, w->res, w->diff, w->cig_blob)))
      return gpu_fail(errmsgp, w, rc);
    w->cig_first = SMB_CIGAR_FIRST(w->cig_blob);
    w->cig_nm = SMB_CIGAR_NM(w->cig_blob, sz.nresults);
    w->cig_text = SMB_CIGAR_TEXT(w->cig_blob, sz.nresults);
    w->have_cigar = 1;
  } else if ((rc = smb_block_fetch(w->ctx, brd, bc, w->ba_err, w->res_first, w->res, w->diff))) return gpu_fail(errmsgp, w, rc);
  WTICK(6);

  /* host: replay of alignRMAPCANDFull (rmap.c:820-926) on the alignments, then results.c as in the reference */
  if ((errcode = dev_results(errmsgp, rmp, w, w->rd, w, brd, bc, 0, n, jobs, max_depth, rmapflg, matchscor, ssp, codecp, donef, user)))
    return errcode;
  WTICK(7);
  w->n_reads += (uint64_t) n;
  return ERRCODE_SUCCESS;
}

/* waves 1b-3 for a list of jobs on the reads of the current seed batch */
static int wave_pass(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int n, const WJOB *jobs, const smb_seed_info *info,
		     int ktuple_maxhit, int min_swatscor_below_max_arg, short target_depth, short max_depth,
		     RMAPFLG_t rmapflg, const ScoreMatrix *scormtxp,
		     const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
		     WAVE_DONEF *donef, void *user)
{
  int errcode = ERRCODE_SUCCESS, rc, i;
  UCHAR nskip;
  const UCHAR ktup = hashTableGetKtupLen(htp, &nskip);
  const SETSIZ_t *soffs;
  const SEQNUM_t nseq = seqSetGetOffsets(ssp, &soffs);
  RMAPBUFF *bufp = rmp->bfp;
  size_t nreq = 0, nreq_max = 0, nsw = 0, nbf = 0, nba = 0;
  int have_pen = 0;
  short matchscor = 0, mismatchscor = 0, gapinitscor = 0, gapextscor = 0;
  size_t nres = 0, ndiff = 0;
  uint64_t cells = 0;
  double tw = wnow(), tc = cnow(), tres;

  if (n < 1) return ERRCODE_SUCCESS;
  if (!wave_host_cand())
    return wave_pass_dev(errmsgp, rmp, w, n, jobs, info, ktuple_maxhit, min_swatscor_below_max_arg, target_depth, max_depth,
			 rmapflg, scormtxp, htp, ssp, codecp, donef, user);
  WGROW(w->rd, w->n_alloc, n, WREAD);
  for (i = 0; i < n; i++) nreq_max += 2 * (size_t) (jobs[i].niv < 0 ? nseq : jobs[i].niv);

  /* hit lists: every read x strand x reference sequence (collectHits, rmap.c:283-318) */
  /* or, restricted: read x strand x interval with the full seed table (collectHitsFromInterVal,
   * rmap.c:438-493) */
  WPIN(w->req, WB_REQ, nreq_max + 1, smb_hit_req);
  WPIN(w->list_first, WB_LIST_FIRST, nreq_max + 2, uint64_t);
  WPIN(w->req_err, WB_REQ_ERR, nreq_max + 1, int32_t);
  for (i = 0; i < n; i++) {
    WREAD *rd = w->rd + i;
    const WJOB *jb = jobs + i;
    const uint32_t r = jb->seed_read;
    int st, k;
    SEQNUM_t s;
    memset(rd, 0, sizeof(*rd));
    rd->qlen = w->read_len[jb->read];
    rd->errcode = info[2 * r].err ? info[2 * r].err : info[2 * r + 1].err;
    rd->cand_first = 0;
    if (rd->errcode) continue;
    for (st = 0; st < 2; st++) {
      if (jb->niv < 0) {
	for (s = 0; s < nseq; s++) {
	  smb_hit_req *rq = w->req + nreq++;
	  memset(rq, 0, sizeof(*rq));
	  rq->lo = soffs[s]; rq->hi = soffs[s + 1];
	  rq->read = r; rq->nhit_max = (uint32_t) ktuple_maxhit;
	  rq->strand = (uint8_t) st; rq->use_short = 1;
	}
      } else {
	for (k = 0; k < jb->niv; k++) {
	  const WIVAL *iv = w->iv + jb->iv_first + k;
	  smb_hit_req *rq = w->req + nreq++;
	  memset(rq, 0, sizeof(*rq));
	  rq->lo = iv->lo; rq->hi = iv->hi;
	  rq->read = r; rq->nhit_max = (uint32_t) ktuple_maxhit;
	  rq->strand = (uint8_t) st; rq->use_short = 0;
	}
      }
    }
  }
  if (w->sqdat_alloc < 32 * (size_t) n) { /* typical short-read blocks need no second counting pass */
    WPIN(w->sqdat, WB_SQDAT, 48 * (size_t) n + 1024, uint64_t);
    w->sqdat_alloc = w->wb[WB_SQDAT].cap / sizeof(uint64_t);
  }
  for (;;) {
    size_t need = 0;
    rc = smb_hits_batch(w->ctx, w->req, (int) nreq, 0, w->sqdat, w->sqdat_alloc, &need, w->list_first,
			w->req_err);
    if (rc == SMB_ERR_CAPACITY && need > w->sqdat_alloc) {
      WPIN(w->sqdat, WB_SQDAT, need + need / 4 + 1024, uint64_t);
      w->sqdat_alloc = w->wb[WB_SQDAT].cap / sizeof(uint64_t);
      continue;
    }
    if (rc) return gpu_fail(errmsgp, w, rc);
    break;
  }
  w->ms_k1 += smb_last_kernel_ms(w->ctx);
  WTICK(2);

  /* host: candidate selection with the reference's segment.c, per read */
  w->ncand = 0;
  nreq = 0;
  for (i = 0; i < n; i++) {
    WREAD *rd = w->rd + i;
    const WJOB *jb = jobs + i;
    uint32_t min_cover = jb->min_cover, mincov_below_max, min_ktup, n_candseg, c;
    int st, nlist;
    short mismatchdiff;
    rd->cand_first = (uint32_t) w->ncand;
    if (rd->errcode) continue;
    /* prelude of mapSingleRead (rmap.c:1258-1290) */
    if (!have_pen) { /* penalties are those of the score matrix, identical for every read */
      if ((errcode = scoreMakeProfileFromSequence(w->prof, jb->readp, scormtxp))) return errcode;
      matchscor = scoreProfileGetAvgPenalties(&mismatchscor, &gapinitscor, &gapextscor, w->prof);
      if ((rc = smbShimSetScoring(w->ctx, w->prof))) return gpu_fail(errmsgp, w, rc);
      have_pen = 1;
    }
    mismatchdiff = (short) (matchscor - mismatchscor);
    if (mismatchdiff < 0 || gapextscor >= 0 || mismatchscor >= 0) return ERRCODE_ASSERT;
    min_ktup = calcMinKtup(&min_cover, htp);
    if (min_swatscor_below_max_arg < 0) {
      mincov_below_max = rd->qlen - 1;
    } else {
      mincov_below_max = ((uint32_t) (min_swatscor_below_max_arg / mismatchdiff)) * nskip;
      if (mincov_below_max < ktup || (rmapflg & RMAPFLG_BEST))
	mincov_below_max = ktup + 2 * (nskip - 1);
    }
    smbShimHitInfoSet(rmp->mrp->hhiFp, info + 2 * jb->seed_read);
    smbShimHitInfoSet(rmp->mrp->hhiRp, info + 2 * jb->seed_read + 1);
    blankRMAPBUFF(bufp);
    nlist = (jb->niv < 0) ? (int) nseq : jb->niv;
    for (st = 0; st < 2; st++)
      for (c = 0; c < (uint32_t) nlist; c++, nreq++) {
	const uint64_t f0 = w->list_first[nreq], f1 = w->list_first[nreq + 1];
	const SEQNUM_t s = (jb->niv < 0) ? (SEQNUM_t) c : (SEQNUM_t) w->iv[jb->iv_first + c].sx;
	if (rd->errcode) continue;
	if (w->req_err[nreq] && w->req_err[nreq] != SMB_ERRCODE_ALLOCBOUNDARY) { rd->errcode = w->req_err[nreq]; continue; }
	hashBlankHitList(bufp->hhlp);
	if ((errcode = smbShimHitListSet(bufp->hhlp, w->sqdat + f0, (int) (f1 - f0), st, rd->qlen, ktup, nskip)))
	  return errcode;
	segLstBlank(bufp->sglp);
	if ((errcode = segLstFillHits(bufp->sglp, min_ktup, bufp->hhlp)) ||
	    (errcode = segAliCandsAddFast(bufp->sacp, bufp->qmp, bufp->sglp, min_cover, (int) s)))
	  rd->errcode = errcode;
      }
    if (rd->errcode) { ERRMSGNO(errmsgp, rd->errcode); continue; }
    if ((errcode = segAliCandsStats(bufp->sacp, mincov_below_max, rmp->mrp->hhiFp, rmp->mrp->hhiRp,
				    target_depth, max_depth, (uint8_t) (rmapflg & RMAPFLG_SENSITIVE)))) {
      rd->errcode = errcode;
      ERRMSGNO(errmsgp, errcode);
      continue;
    }
    {
      uint32_t nseg_tot;
      const uint32_t nseg = segAliCandsGetNumberOfSegments(bufp->sacp, NULL, NULL, NULL, NULL, &nseg_tot);
      if (nseg > INT_MAX || nseg_tot > INT_MAX) { rd->errcode = ERRCODE_ASSERT; continue; }
      rd->nseg = (int) nseg;
      rd->nseg_tot = (int) nseg_tot;
      rd->nhit = calcTotalHitNumStats(rmp->mrp, &rd->nhit_tot);
      rd->reached_stats = 1;
    }
    /* candidate windows: what scoreRMAPCAND would fetch one by one (rmap.c:660-692) */
    n_candseg = segAliCandsGetNumberOfSegments(bufp->sacp, &rd->curr_min_cover, NULL,
					       rd->cover_deficit, rd->cover_deficit + 1, NULL);
    WGROW(w->cand, w->cand_alloc, w->ncand + n_candseg + 1, WCAND);
    for (c = 0; c < n_candseg; c++) {
      WCAND *wc = w->cand + w->ncand;
      RMAPCAND *cp = &wc->c;
      uint8_t bitflags;
      SEQLEN_t slen;
      memset(wc, 0, sizeof(*wc));
      errcode = segAliCandsCalcSegmentOffsets(&cp->qs, &cp->qe, &cp->rs, &cp->re, &cp->band_l, &cp->band_r,
					      &cp->dqo, &cp->dro, &cp->sqidx, &bitflags, &wc->cover,
					      0, rd->qlen, ssp, c, bufp->sacp);
      if (!errcode && (cp->qe > INT_MAX || cp->re < cp->rs || cp->re - cp->rs > INT_MAX))
	errcode = ERRCODE_OVERFLOW;
      if (!errcode && (cp->sqidx < 0 || cp->sqidx >= nseq)) errcode = ERRCODE_ARGRANGE;
      if (!errcode) {
	slen = (SEQLEN_t) (soffs[cp->sqidx + 1] - soffs[cp->sqidx]);
	if (cp->rs >= slen) errcode = ERRCODE_SEQOFFS; /* seqSetFetchSegmentBySequence, sequence.c:2756 */
	else {
	  uint64_t len = cp->re - cp->rs + 1;
	  if (cp->rs + len > slen) len = slen - cp->rs;
	  wc->reflen = (uint32_t) len;
	  wc->refoff = soffs[cp->sqidx] + cp->rs;
	}
      }
      if (errcode) { rd->errcode = errcode; ERRMSGNO(errmsgp, errcode); break; }
      cp->flags = (RMAPFLG_t) ((bitflags & SEGCANDFLG_REVERSE) ? RMAPCANDFLG_REVERSE : 0);
      cp->swscor = 0;
      wc->k3task = -1;
      wc->simd = (uint8_t) (rd->qlen >= MINLEN_QUERY_STRIPED &&
			    ((SEQLEN_t) (cp->band_r - cp->band_l) * BWSCAL_QLEN) > rd->qlen &&
			    cp->qs == 0 && cp->qe >= rd->qlen - 1);
      w->ncand++;
      rd->ncand++;
    }
    if (rd->errcode) { w->ncand = rd->cand_first; rd->ncand = 0; }
  }

  WTICK(3);
  /* ------------------------------ wave 2: scores ------------------------------------- */
  WPIN(w->swt, WB_SWT, w->ncand + 1, smb_sw_task);
  WPIN(w->sw_score, WB_SW_SCORE, w->ncand + 1, int32_t);
  WPIN(w->sw_err, WB_SW_ERR, w->ncand + 1, int32_t);
  WPIN(w->bft, WB_BFT, w->ncand + 1, smb_band_task);
  WPIN(w->bf_score, WB_BF_SCORE, w->ncand + 1, int32_t);
  WPIN(w->bf_err, WB_BF_ERR, w->ncand + 1, int32_t);
  for (i = 0; i < n; i++) {
    WREAD *rd = w->rd + i;
    uint32_t c;
    for (c = 0; c < rd->ncand; c++) {
      WCAND *wc = w->cand + rd->cand_first + c;
      const uint32_t flags = SMB_TASK_REF_PACKED | ((wc->c.flags & RMAPCANDFLG_REVERSE) ? SMB_TASK_READ_REVCOMP : 0);
      if (wc->simd) {
	smb_sw_task *t = w->swt + nsw;
	memset(t, 0, sizeof(*t));
	t->read_off = w->read_off[jobs[i].read]; t->read_len = rd->qlen;
	t->ref_off = wc->refoff; t->ref_len = wc->reflen; t->flags = flags;
	wc->task = (int32_t) nsw++;
	w->cells_k2 += (uint64_t) rd->qlen * wc->reflen;
      } else {
	smb_band_task *t = w->bft + nbf;
	memset(t, 0, sizeof(*t));
	t->read_off = w->read_off[jobs[i].read]; t->read_len = rd->qlen;
	t->ref_off = wc->refoff; t->ref_len = wc->reflen; t->flags = flags;
	t->l_edge = wc->c.band_l; t->r_edge = wc->c.band_r;
	t->p_left = (int) wc->c.qs; t->p_right = (int) wc->c.qe;
	t->u_left = 0; t->u_right = (int) wc->reflen - 1;
	wc->task = (int32_t) nbf++;
      }
    }
  }
  if (nsw) {
    if ((rc = smb_sw_score_batch(w->ctx, w->swt, (int) nsw, w->sw_score, w->sw_err))) return gpu_fail(errmsgp, w, rc);
    w->ms_k2 += smb_last_kernel_ms(w->ctx);
    w->n_k2 += nsw;
    /* ERRCODE_SWATEXCEED -> banded fast variant (rmap.c:730-744) */
    for (i = 0; i < n; i++) {
      WREAD *rd = w->rd + i;
      uint32_t c;
      for (c = 0; c < rd->ncand; c++) {
	WCAND *wc = w->cand + rd->cand_first + c;
	if (wc->simd && w->sw_err[wc->task] == ERRCODE_SWATEXCEED) {
	  smb_band_task *t = w->bft + nbf;
	  memset(t, 0, sizeof(*t));
	  t->read_off = w->read_off[jobs[i].read]; t->read_len = rd->qlen;
	  t->ref_off = wc->refoff; t->ref_len = wc->reflen;
	  t->flags = SMB_TASK_REF_PACKED | ((wc->c.flags & RMAPCANDFLG_REVERSE) ? SMB_TASK_READ_REVCOMP : 0);
	  t->l_edge = wc->c.band_l; t->r_edge = wc->c.band_r;
	  t->p_left = (int) wc->c.qs; t->p_right = (int) wc->c.qe;
	  t->u_left = 0; t->u_right = (int) wc->reflen - 1;
	  wc->simd = 0;
	  wc->task = (int32_t) nbf++;
	}
      }
    }
  }
  if (nbf) {
    if ((rc = smb_band_score_batch(w->ctx, w->bft, (int) nbf, w->bf_score, w->bf_err))) return gpu_fail(errmsgp, w, rc);
    w->ms_k2 += smb_last_kernel_ms(w->ctx);
  }

  WTICK(4);
  /* host: replay of scoreRMAPCAND (rmap.c:646-786) and of mapSingleRead (rmap.c:1366-1400) */
  WPIN(w->bat, WB_BAT, w->ncand + 1, smb_band_task);
  WPIN(w->ba_err, WB_BA_ERR, w->ncand + 1, int32_t);
  for (i = 0; i < n; i++) {
    WREAD *rd = w->rd + i;
    const short mmscordiff = (short) (matchscor - mismatchscor), gapscordiff = (short) (matchscor - gapinitscor);
    COVERAGE max_cover = 0, min_cover = 0, dcov, cdf;
    SWATSCOR max1 = 0, max2 = 0;
    const SWATSCOR max_possible_swscor = (SWATSCOR) (rd->qlen * matchscor);
    uint32_t c;
    int min_swatscor = jobs[i].min_swatscor, min_swatscor_below_max = min_swatscor_below_max_arg;
    if (rd->errcode || !rd->reached_stats) continue;
    if (mmscordiff < 1 || gapscordiff < 1) return ERRCODE_ASSERT;
    for (c = 0; c < rd->ncand; c++) {
      WCAND *wc = w->cand + rd->cand_first + c;
      RMAPCAND *cp = &wc->c;
      const COVERAGE cover = wc->cover;
      int e;
      /* (the cover-rank early stop of rmap.c:673-679 is compiled out in the reference:
       * rmap_stop_candlist_early is not defined, rmap.h:39) */
      if (wc->simd) { e = w->sw_err[wc->task]; cp->swscor = w->sw_score[wc->task]; }
      else { e = w->bf_err[wc->task]; cp->swscor = w->bf_score[wc->task]; }
      if (e) { rd->errcode = e; break; }
      cp->flags |= RMAPCANDFLG_SCORED;
      cdf = rd->cover_deficit[(cp->flags & RMAPCANDFLG_REVERSE) ? 1 : 0];
      if ((rmapflg & RMAPFLG_BEST) && (cover + cdf < min_cover)) break;
      if (cp->swscor > max2) {
	if (cp->swscor > max1) {
	  max2 = max1;
	  max1 = cp->swscor;
	  if (cover + cdf > max_cover) max_cover = (cover > cdf) ? cover - cdf : 0;
	} else {
	  max2 = cp->swscor;
	}
	dcov = ((int) ((max1 - max2) / mmscordiff) + 1) * nskip;
	if (dcov + cdf + min_cover < max_cover) min_cover = max_cover - dcov;
      }
    }
    rd->nscored = c;
    rd->max1scor = max1;
    rd->max2scor = max2;
    if (rd->errcode) { ERRMSGNO(errmsgp, rd->errcode); continue; }
    if (max1 > max_possible_swscor) { rd->errcode = ERRCODE_ASSERT; ERRMSGNO(errmsgp, ERRCODE_ASSERT); continue; }
    if (max1 < 1) continue;
    rd->do_align = 1;
    rd->scorlen_min = ktup + nskip;
    rd->bandwidth_min = (max_possible_swscor - max1) / (-1 * gapextscor);
    if (min_swatscor_below_max >= max1) min_swatscor_below_max = max1;
    if (min_swatscor > max2 && max2 > 0) min_swatscor = max2;
    if (min_swatscor_below_max >= 0) {
      const SWATSCOR minswc = (max2 > 0) ? max2 : max1;
      if (rmapflg & RMAPFLG_BEST) {
	if (minswc > min_swatscor) min_swatscor = minswc;
      } else if (min_swatscor + min_swatscor_below_max < max1) {
	min_swatscor = max1 - min_swatscor_below_max;
	if (min_swatscor > minswc) min_swatscor = minswc;
      }
    }
    if (min_swatscor > rd->scorlen_min * matchscor && matchscor > 0) rd->scorlen_min = min_swatscor / matchscor;
    rd->min_swatscor = min_swatscor;
    if (wave_debug())
      fprintf(stderr, "DBG read %d ncand %u nscored %u max1 %d max2 %d min_swatscor %d scorlen_min %d bw_min %d nseg %d/%d nhit %u/%u\n",
	      i, rd->ncand, rd->nscored, max1, max2, min_swatscor, rd->scorlen_min, rd->bandwidth_min, rd->nseg,
	      rd->nseg_tot, rd->nhit, rd->nhit_tot);
    /* K3 tasks: every scored candidate that passes the INITIAL threshold (rmap.c:833-835) */
    for (c = 0; c < rd->nscored; c++) {
      WCAND *wc = w->cand + rd->cand_first + c;
      const RMAPCAND *cp = &wc->c;
      smb_band_task *t;
      int bw, band_l, band_r;
      if ((cp->flags & RMAPCANDFLG_SCORED) && cp->swscor < min_swatscor) continue;
      bw = cp->band_r - cp->band_l;
      if (bw < rd->bandwidth_min) {
	bw = (rd->bandwidth_min - bw + 1) / 2;
	band_l = cp->band_l - bw;
	band_r = cp->band_r + bw;
      } else {
	band_l = cp->band_l;
	band_r = cp->band_r;
      }
      t = w->bat + nba;
      memset(t, 0, sizeof(*t));
      t->read_off = w->read_off[jobs[i].read]; t->read_len = rd->qlen;
      t->ref_off = wc->refoff; t->ref_len = wc->reflen;
      t->flags = SMB_TASK_REF_PACKED | ((cp->flags & RMAPCANDFLG_REVERSE) ? SMB_TASK_READ_REVCOMP : 0);
      t->l_edge = band_l; t->r_edge = band_r;
      t->p_left = (int) cp->qs; t->p_right = (int) cp->qe;
      t->u_left = 0; t->u_right = (int) wc->reflen - 1;
      t->minscore = min_swatscor; t->minscorlen = rd->scorlen_min;
      wc->k3task = (int32_t) nba++;
    }
  }

  WTICK(5);
  /* ------------------------------ wave 3: alignments --------------------------------- */
  if (nba) {
    for (;;) {
      WPIN(w->res_first, WB_RES_FIRST, nba + 2, uint32_t);
      if (w->res_alloc < nba + 16) {
	WPIN(w->res, WB_RES, nba + nba / 2 + 64, smb_ali_result);
	w->res_alloc = w->wb[WB_RES].cap / sizeof(smb_ali_result);
      }
      if (w->diff_alloc < 32 * nba) {
	WPIN(w->diff, WB_DIFF, 48 * nba + 4096, uint8_t);
	w->diff_alloc = w->wb[WB_DIFF].cap;
      }
      rc = smb_band_align_batch(w->ctx, w->bat, (int) nba, w->res, w->res_alloc, &nres, w->res_first,
				w->diff, w->diff_alloc, &ndiff, w->ba_err, &cells);
      if (rc == SMB_ERR_CAPACITY && (nres > w->res_alloc || ndiff > w->diff_alloc)) {
	if (nres > w->res_alloc) {
	  WPIN(w->res, WB_RES, nres + 64, smb_ali_result);
	  w->res_alloc = w->wb[WB_RES].cap / sizeof(smb_ali_result);
	}
	if (ndiff > w->diff_alloc) {
	  WPIN(w->diff, WB_DIFF, ndiff + 4096, uint8_t);
	  w->diff_alloc = w->wb[WB_DIFF].cap;
	}
	continue;
      }
      if (rc) return gpu_fail(errmsgp, w, rc);
      break;
    }
    w->ms_k3 += smb_last_kernel_ms(w->ctx);
    w->n_k3 += nba;
    w->cells_k3 += cells;
  }

  WTICK(6);
  /* host: replay of alignRMAPCANDFull (rmap.c:820-926), then results.c as in the reference */
  tres = rnow();
  for (i = 0; i < n; i++) {
    WREAD *rd = w->rd + i;
    ResultSet *rsp = jobs[i].rsp;
    SeqFastq *readp = jobs[i].readp;
    uint32_t c;
    if (jobs[i].blank) resultSetBlank(rsp);
    if (rd->errcode == ERRCODE_SHORTSEQ) { /* too short to be hashed (rmap.c:1273-1275) */
      if (donef && (errcode = (*donef)(user, i, ERRCODE_SHORTSEQ, rsp))) return errcode;
      continue;
    }
    if (rd->reached_stats)
      resultSetAlignmentStats(rsp, rd->nseg, rd->nseg_tot, max_depth, rd->nhit, rd->nhit_tot);
    if (!rd->errcode && rd->do_align) {
      int min_swatscor = rd->min_swatscor;
      SWATSCOR swatscor_2ndmax = 0;
      int need_profiles = 0;
      for (c = 0; c < rd->nscored && !rd->errcode; c++) {
	WCAND *wc = w->cand + rd->cand_first + c;
	const RMAPCAND *cp = &wc->c;
	int minscorlen = rd->scorlen_min;
	uint32_t pos, end;
	if ((cp->flags & RMAPCANDFLG_SCORED) && cp->swscor < min_swatscor) continue;
	if (rmapflg & RMAPFLG_BEST) {
	  resultSetGetMaxSwat(rsp, &swatscor_2ndmax);
	  if (swatscor_2ndmax > min_swatscor) min_swatscor = swatscor_2ndmax;
	}
	if (wc->k3task < 0) { rd->errcode = ERRCODE_ASSERT; break; }
	if (w->ba_err[wc->k3task]) { rd->errcode = w->ba_err[wc->k3task]; break; }
	aliRsltSetReset(bufp->alirsltp);
	/* aliSmiWatInBand (alignment.c:1569-1575) with the current threshold */
	if (min_swatscor < 1 || matchscor <= 0) { rd->errcode = ERRCODE_ASSERT; break; }
	if (minscorlen * matchscor < min_swatscor) minscorlen = min_swatscor / matchscor;
	if (minscorlen < 5) { rd->errcode = ERRCODE_ASSERT; break; }
	pos = w->res_first[wc->k3task];
	end = w->res_first[wc->k3task + 1];
	if ((errcode = prune_results(bufp->alirsltp, w->res, w->diff, &pos, end, 0, (int) wc->reflen - 1,
				     min_swatscor, minscorlen, 1)))
	  return errcode;
	if (wave_debug()) {
	  short k_, n_ = aliRsltSetGetSize(bufp->alirsltp);
	  fprintf(stderr, "DBG  cand %u swscor %d cover %u rev %d rs %llu band %d %d minscore %d minscorlen %d raw %u kept %d:",
		  c, cp->swscor, wc->cover, (int) (cp->flags & RMAPCANDFLG_REVERSE), (unsigned long long) cp->rs,
		  w->bat[wc->k3task].l_edge, w->bat[wc->k3task].r_edge, min_swatscor, minscorlen,
		  w->res_first[wc->k3task + 1] - w->res_first[wc->k3task], (int) n_);
	  for (k_ = 0; k_ < n_; k_++) {
	    int sc_, a_, b_, c_, d_;
	    aliRsltSetFetchData(bufp->alirsltp, k_, &sc_, &a_, &b_, &c_, &d_, NULL);
	    fprintf(stderr, " (%d q%d-%d r%d-%d)", sc_, a_, b_, c_, d_);
	  }
	  fputc('\n', stderr);
	}
	if (cp->sqidx == SEGCAND_UNKNOWN_SEQIDX) need_profiles = 1;
	errcode = resultSetAddFromAli(rsp, bufp->alirsltp, cp->rs, 0, rd->qlen,
				      (cp->sqidx == SEGCAND_UNKNOWN_SEQIDX) ? RESULTSET_UNKNOWN_SEQIDX : cp->sqidx,
				      (char) (cp->flags & RMAPCANDFLG_REVERSE));
	if (errcode) { rd->errcode = errcode; break; }
      }
      { const double t_ = rnow(); w->wall_res[0] += t_ - tres; tres = t_; }
      if (rd->errcode) ERRMSGNO(errmsgp, rd->errcode);
      else {
	/* The profiles of the read (rmap.c:1419-1428) are only dereferenced by
	 * resultSetSortAndAssignSequence for results without a sequence index (results.c:1715,
	 * :1742-1756: alignments spanning reference sequences in the lumped mode) - never in the
	 * sequence-by-sequence mode this path runs in; they are built only if such a result exists. */
	if (need_profiles) {
	  seqFastqBlank(w->readRC);
	  if ((errcode = seqFastqAppendSegment(w->readRC, readp, 0, 0, 1, codecp)) ||
	      (errcode = scoreMakeProfileFromSequence(w->prof, readp, scormtxp)) ||
	      (errcode = scoreMakeProfileFromSequence(w->profRC, w->readRC, scormtxp)))
	    return errcode;
	}
	errcode = resultSetSortAndAssignSequence(rsp, bufp->sqbfp, 0, readp, w->prof, w->profRC, ssp, codecp);
	if (errcode) { rd->errcode = errcode; ERRMSGNO(errmsgp, errcode); }
      }
    }
    if (donef && (errcode = (*donef)(user, i, rd->errcode, rsp))) return errcode;
    tres = rnow();
  }
  WTICK(7);
  w->n_reads += (uint64_t) n;
  return ERRCODE_SUCCESS;
}

/* ------------------------------------------------------------------------------------ */
/* single-end blocks                                                                      */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  RmapWave *w;
  ErrMsg *errmsgp;
  const ResultFilter *rsfp;
  const WJOB *jobs;
  RMAPWAVE_EMITF *emitf;
  void *user;
} SINGLEDONE;

static int single_done(void *user, int i, int errcode, ResultSet *rsp)
{
  SINGLEDONE *sd = (SINGLEDONE *) user;
  double t0 = rnow(), t1;
  int e;
  if (errcode != ERRCODE_SHORTSEQ && !errcode && (e = resultSetFilterResults(rsp, sd->rsfp, sd->jobs[i].readp)))
    ERRMSGNO(sd->errmsgp, e);
  t1 = rnow();
  sd->w->wall_res[1] += t1 - t0;
  e = (*sd->emitf)(sd->user, i, rsp);
  sd->w->wall_res[2] += rnow() - t1;
  return e;
}

int rmapSingleWave(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int n, SeqFastq **reads,
		   const uint32_t *min_cover_arr, int ktuple_maxhit, int min_swatscor_arg,
		   int min_swatscor_below_max_arg, UCHAR min_basqval, short target_depth, short max_depth,
		   RMAPFLG_t rmapflg, const ScoreMatrix *scormtxp, const ResultFilter *rsfp,
		   const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
		   RMAPWAVE_EMITF *emitf, void *user)
{
  int errcode, i;
  SINGLEDONE sd;
  if (n < 1) return ERRCODE_SUCCESS;
  if (!(rmapflg & RMAPFLG_SEQBYSEQ) || (rmapflg & (RMAPFLG_NOSHRTINFO | RMAPFLG_SPLIT | RMAPFLG_CMPLXW)))
    return ERRCODE_ARGINVAL; /* caller runs the reference's own per-read code on fibers (still GPU) */
  w->cigar_mode = g_cigar_mode;
  if ((errcode = wave_seed(errmsgp, w, n, reads, ktuple_maxhit, min_basqval))) return errcode;
  WGROW(w->jobs, w->jobs_alloc, n, WJOB);
  for (i = 0; i < n; i++) {
    WJOB *jb = w->jobs + i;
    jb->read = jb->seed_read = (uint32_t) i; jb->min_cover = min_cover_arr[i]; jb->min_swatscor = min_swatscor_arg;
    jb->readp = reads[i]; jb->rsp = rmp->rsrp; jb->niv = -1; jb->iv_first = 0; jb->blank = 1;
  }
  sd.w = w; sd.errmsgp = errmsgp; sd.rsfp = rsfp; sd.jobs = w->jobs; sd.emitf = emitf; sd.user = user;
  return wave_pass(errmsgp, rmp, w, n, w->jobs, w->info, ktuple_maxhit, min_swatscor_below_max_arg, target_depth, max_depth,
		   rmapflg, scormtxp, htp, ssp, codecp, single_done, &sd);
}

/* ------------------------------------------------------------------------------------ */
/* combined device batches                                                                */
/* ------------------------------------------------------------------------------------ */
/* The host stages want small blocks (parse and results.c of ~8000 reads per worker thread keep all cores
 * busy and the output in order), the device wants large batches: the kernels of an 8192-read block are
 * short and latency bound, and since the persistent K2 / K3 kernels fill the machine the blocks of different
 * worker streams run one after the other (104 ms of device time per 1 M reads at 8192-read launches, 74 ms
 * at 32000).  A WaveCombiner joins the blocks of several workers into ONE device batch: a worker copies its
 * reads into the shared staging of the open batch and waits; the batch closes when it is full or as soon as
 * the device is free, its last member to arrive drives the GPU for everybody (seed tables, smb_block_run,
 * smb_block_fetch), then every member runs results.c on its own reads.  Batch size follows the load: a lone
 * worker maps its block at once, a saturated device gets batches of `target` reads. */
#include <pthread.h>
enum { CS_FREE = 0, CS_OPEN, CS_CLOSED, CS_RUNNING, CS_DONE };
enum { COMB_MAXSLOTS = 6 };
typedef struct {
  RmapWave *bw;              /* device context + page-locked staging / outputs of the batch */
  int state, nmembers, ncopied, ndone, nreads, leader, any_qual, errcode;
  size_t bytes;
  smb_block_job *jobs;
  smb_block_read *brd;
  smb_block_cand *bc;
  int have_pen, have_prof;
} CombSlot;
struct WaveCombiner_ {
  pthread_mutex_t lock;
  pthread_cond_t cond;
  int nslots, target, cap_reads, open, nrunning, max_running, flush, stop;
  size_t cap_bytes;
  CombSlot slot[COMB_MAXSLOTS];
  uint64_t nbatches, nbatch_reads;
};

WaveCombiner *waveCombinerCreate(const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
				 const ScoreMatrix *scormtxp, int nslots, int target_reads, int spin)
{
  WaveCombiner *wc = (WaveCombiner *) calloc(1, sizeof(*wc));
  int k;
  if (!wc) return NULL;
  if (nslots < 2) nslots = 2;
  if (nslots > COMB_MAXSLOTS) nslots = COMB_MAXSLOTS;
  if (target_reads < 1024) target_reads = 1024;
  wc->nslots = nslots; wc->target = target_reads; wc->cap_reads = 2 * target_reads;
  wc->cap_bytes = (size_t) wc->cap_reads * 320;
  wc->open = -1;
  wc->max_running = getenv("SMALT_B200_BATCHES") ? atoi(getenv("SMALT_B200_BATCHES")) : 3;
  if (wc->max_running < 1) wc->max_running = 1;
  pthread_mutex_init(&wc->lock, NULL);
  pthread_cond_init(&wc->cond, NULL);
  for (k = 0; k < nslots; k++) {
    RmapWave *bw = rmapWaveCreate(htp, ssp, codecp, scormtxp);
    CombSlot *b = wc->slot + k;
    if (!bw) { waveCombinerDelete(wc); return NULL; }
    b->bw = bw;
    /* the leader of a batch drives the device for several workers: it polls instead of sleeping, so that it
     * does not wait for a time slice after every synchronisation while the cores run results.c */
    if (spin && !getenv("SMALT_B200_NOSPIN")) smb_ctx_set_spin(bw->ctx, spin);
    bw->read_off = (uint64_t *) wbuf_need(&bw->wb[WB_READ_OFF], (size_t) wc->cap_reads * sizeof(uint64_t));
    bw->read_len = (uint32_t *) wbuf_need(&bw->wb[WB_READ_LEN], (size_t) wc->cap_reads * sizeof(uint32_t));
    bw->info = (smb_seed_info *) wbuf_need(&bw->wb[WB_INFO], 2 * (size_t) wc->cap_reads * sizeof(smb_seed_info));
    bw->arena = (uint8_t *) wbuf_need(&bw->wb[WB_ARENA], wc->cap_bytes + 16);
    bw->qual = (uint8_t *) wbuf_need(&bw->wb[WB_QUAL], wc->cap_bytes + 16);
    b->jobs = (smb_block_job *) wbuf_need(&bw->wb[WB_BJOB], (size_t) wc->cap_reads * sizeof(smb_block_job));
    if (!bw->read_off || !bw->read_len || !bw->info || !bw->arena || !bw->qual || !b->jobs) { waveCombinerDelete(wc); return NULL; }
  }
  return wc;
}

void waveCombinerDelete(WaveCombiner *wc)
{
  int k;
  if (!wc) return;
  for (k = 0; k < COMB_MAXSLOTS; k++) rmapWaveDelete(wc->slot[k].bw);
  pthread_mutex_destroy(&wc->lock);
  pthread_cond_destroy(&wc->cond);
  free(wc);
}

int waveCombinerSlots(const WaveCombiner *wc, RmapWave **waves, uint64_t counts[2])
{
  int k;
  for (k = 0; k < wc->nslots; k++) waves[k] = wc->slot[k].bw;
  counts[0] = wc->nbatches; counts[1] = wc->nbatch_reads;
  return wc->nslots;
}

/* batches on the device at a time = device threads of the caller (unless SMALT_B200_BATCHES says otherwise) */
void waveCombinerSetRunning(WaveCombiner *wc, int n)
{
  if (wc && n >= 1 && !getenv("SMALT_B200_BATCHES")) wc->max_running = n;
}

/* closes the open batch if fewer than max_running batches are on the device (two, so that the copies and
 * host synchronisations of one overlap the kernels of the other) and every member has delivered its reads
 * (lock held) */
static void comb_try_close(WaveCombiner *wc)
{
  if (wc->open >= 0 && wc->nrunning < wc->max_running) {
    CombSlot *b = wc->slot + wc->open;
    if (b->state == CS_OPEN && b->nmembers > 0 && b->ncopied == b->nmembers) {
      b->state = CS_CLOSED;
      wc->open = -1;
      pthread_cond_broadcast(&wc->cond);
    }
  }
}

/* the leader's part: the whole batch through the device */
static int comb_run_gpu(ErrMsg *errmsgp, RmapWave *lw, CombSlot *b, int ktuple_maxhit, int min_swatscor_below_max_arg,
			UCHAR min_basqval, short target_depth, short max_depth, RMAPFLG_t rmapflg,
			const ScoreMatrix *scormtxp, SeqFastq *any_read, const SeqSet *ssp)
{
  RmapWave *w = b->bw;
  const int n = b->nreads;
  int rc, errcode;
  smb_block_params prm;
  smb_block_sizes sz;
  const SETSIZ_t *soffs;
  const SEQNUM_t nseq = seqSetGetOffsets(ssp, &soffs);
  (void) lw;
  if (!b->have_pen) {   /* (the profile was made by the first worker that delivered reads to this slot) */
    if (!b->have_prof) return ERRCODE_ASSERT;
    if ((rc = smbShimSetScoring(w->ctx, w->prof))) return gpu_fail(errmsgp, w, rc);
    b->have_pen = 1;
  }
  (void) any_read; (void) scormtxp; (void) errcode;
  if ((rc = smb_arena_upload(w->ctx, w->arena, b->bytes))) return gpu_fail(errmsgp, w, rc);
  if ((rc = smb_seed_batch(w->ctx, w->read_off, w->read_len, n, b->any_qual ? w->qual : NULL,
			   (uint32_t) ktuple_maxhit, HASH_MAXNHITS, min_basqval, 1, w->info,
			   NULL, NULL, NULL, NULL, NULL, NULL)))
    return gpu_fail(errmsgp, w, rc);
  w->ms_k1 += smb_last_kernel_ms(w->ctx);
  memset(&prm, 0, sizeof(prm));
  prm.nhit_max = (uint32_t) ktuple_maxhit;
  prm.min_swatscor_below_max = min_swatscor_below_max_arg;
  prm.target_depth = (int32_t) target_depth;
  prm.max_depth = (int32_t) max_depth;
  prm.best = (uint8_t) ((rmapflg & RMAPFLG_BEST) != 0);
  prm.sensitive = (uint8_t) ((rmapflg & RMAPFLG_SENSITIVE) != 0);
  prm.cigar = (uint8_t) g_cigar_mode;   /* (combined batches are single-end reads) */
  {
    SETSIZ_t roffs;
    const SEQLEN_t rlen0 = nseq > 0 ? seqSetGetSeqDatByIndex(&roffs, NULL, 0, ssp) : 0;
    prm.termchar = (uint8_t) (nseq > 0 && (SETSIZ_t) rlen0 != soffs[1] - soffs[0]);
  }
  if ((rc = smb_block_run(w->ctx, &prm, b->jobs, n, NULL, 0, &sz))) return gpu_fail(errmsgp, w, rc);
  w->ms_k1 += sz.ms_hits + sz.ms_cand;
  w->ms_cand += sz.ms_cand;
  w->ms_k2 += sz.ms_k2;
  w->ms_k3 += sz.ms_k3;
  w->n_k2 += sz.k2_tasks_ref;
  w->cells_k2 += sz.k2_cells_ref;
  w->n_k3 += sz.nk3;
  w->cells_k3 += sz.k3_cells;
  WPIN(b->brd, WB_BRD, n, smb_block_read);
  WPIN(b->bc, WB_BCAND, sz.nk3 + 1, smb_block_cand);
  WPIN(w->ba_err, WB_BA_ERR, sz.nk3 + 1, int32_t);
  WPIN(w->res_first, WB_RES_FIRST, sz.nk3 + 2, uint32_t);
  if (w->res_alloc < sz.nresults + 1) {
    WPIN(w->res, WB_RES, sz.nresults + sz.nresults / 2 + 64, smb_ali_result);
    w->res_alloc = w->wb[WB_RES].cap / sizeof(smb_ali_result);
  }
  if (w->diff_alloc < sz.ndiffbytes + 1) {
    WPIN(w->diff, WB_DIFF, sz.ndiffbytes + sz.ndiffbytes / 2 + 4096, uint8_t);
    w->diff_alloc = w->wb[WB_DIFF].cap;
  }
  w->have_cigar = 0;
  if (prm.cigar) {
    WPIN(w->cig_blob, WB_CIG_BLOB, SMB_CIGAR_BLOB_BYTES(sz.nresults, sz.ncigarbytes) + 64, uint8_t);
    if ((rc = smb_block_fetch_cigar(w->ctx, b->brd, b->bc, w->ba_err, w->res_first, w->res, w->diff, w->cig_blob)))
      return gpu_fail(errmsgp, w, rc);
    w->cig_first = SMB_CIGAR_FIRST(w->cig_blob);
    w->cig_nm = SMB_CIGAR_NM(w->cig_blob, sz.nresults);
    w->cig_text = SMB_CIGAR_TEXT(w->cig_blob, sz.nresults);
    w->have_cigar = 1;
  } else if ((rc = smb_block_fetch(w->ctx, b->brd, b->bc, w->ba_err, w->res_first, w->res, w->diff))) return gpu_fail(errmsgp, w, rc);
  w->n_reads += (uint64_t) n;
  return ERRCODE_SUCCESS;
}

/* can this block go through the combiner?  (flags of the wave path, sizes inside the staging of a batch) */
int waveCombinerTakes(const WaveCombiner *wc, int n, SeqFastq **reads, RMAPFLG_t rmapflg, const HashTable *htp)
{
  size_t tot = 0;
  int i, valid = 0;
  UCHAR nskip;
  const UCHAR ktup = hashTableGetKtupLen(htp, &nskip);
  if (!wc || wave_host_cand() || n < 1 || !(rmapflg & RMAPFLG_SEQBYSEQ) ||
      (rmapflg & (RMAPFLG_NOSHRTINFO | RMAPFLG_SPLIT | RMAPFLG_CMPLXW)))
    return 0;
  for (i = 0; i < n; i++) {
    SEQLEN_t len;
    char cod;
    seqFastqGetConstSequence(reads[i], &len, &cod);
    if (cod != SEQCOD_MANGLED) return 0;
    if (len >= ktup) valid = 1;
    tot += len;
  }
  return valid && n <= wc->cap_reads / 2 && tot <= wc->cap_bytes / 2;
}

/* A worker hands the reads of its block to the open batch (never blocks on the device): 0 and the ticket,
 * or 1 when no batch can take the block right now (all slots busy: the caller finishes other work first). */
int waveCombinerDeliver(WaveCombiner *wc, int n, SeqFastq **reads, const uint32_t *min_cover_arr, int min_swatscor_arg,
			const ScoreMatrix *scormtxp, const HashTable *htp, WaveTicket *tk)
{
  int i, any_qual = 0, s, first, valid = -1;
  size_t tot = 0, off;
  CombSlot *b;
  UCHAR nskip;
  const UCHAR ktup = hashTableGetKtupLen(htp, &nskip);
  for (i = 0; i < n; i++) {
    SEQLEN_t len;
    seqFastqGetConstSequence(reads[i], &len, NULL);
    if (valid < 0 && len >= ktup) valid = i;
    tot += len;
    if (seqFastqGetConstQualityFactors(reads[i], NULL, NULL)) any_qual = 1;
  }
  pthread_mutex_lock(&wc->lock);
  for (;;) {
    s = wc->open;
    if (s >= 0) {
      b = wc->slot + s;
      if (b->nreads + n <= wc->cap_reads && b->bytes + tot <= wc->cap_bytes) break;
      b->state = CS_CLOSED;   /* no room for this block: the batch goes as it is */
      wc->open = -1;
      pthread_cond_broadcast(&wc->cond);
      continue;
    }
    for (s = 0; s < wc->nslots; s++)
      if (wc->slot[s].state == CS_FREE) break;
    if (s >= wc->nslots) { pthread_mutex_unlock(&wc->lock); return 1; }
    b = wc->slot + s;
    b->state = CS_OPEN;
    b->nmembers = b->ncopied = b->ndone = b->nreads = b->leader = b->any_qual = b->errcode = 0;
    b->bytes = 0;
    wc->open = s;
    break;
  }
  if (!b->have_prof && valid >= 0) {   /* penalties are those of the score matrix (rmap.c:1258-1266): any read gives them */
    if (scoreMakeProfileFromSequence(b->bw->prof, reads[valid], scormtxp)) { pthread_mutex_unlock(&wc->lock); return ERRCODE_ASSERT; }
    b->have_prof = 1;
  }
  first = b->nreads;
  off = b->bytes;
  b->nreads += n;
  b->bytes += tot;
  b->nmembers++;
  if (b->nreads >= wc->target) { b->state = CS_CLOSED; wc->open = -1; }
  tk->slot = (int) (b - wc->slot); tk->first = first; tk->n = n;   /* (before the batch can possibly run) */
  pthread_mutex_unlock(&wc->lock);
  {
    RmapWave *bw = b->bw;
    size_t o = off;
    for (i = 0; i < n; i++) {
      SEQLEN_t len;
      const char *p = seqFastqGetConstSequence(reads[i], &len, NULL);
      const char *q = seqFastqGetConstQualityFactors(reads[i], NULL, NULL);
      smb_block_job *jb = b->jobs + first + i;
      bw->read_off[first + i] = o;
      bw->read_len[first + i] = len;
      memcpy(bw->arena + o, p, len);
      if (q) memcpy(bw->qual + o, q, len);
      else memset(bw->qual + o, 0xff, len); /* FASTA read: never below the threshold */
      o += len;
      memset(jb, 0, sizeof(*jb));
      jb->seed_read = (uint32_t) (first + i);
      jb->niv = -1;
      jb->min_cover = min_cover_arr[i];
      jb->min_swatscor = min_swatscor_arg;
    }
  }
  pthread_mutex_lock(&wc->lock);
  b->ncopied++;
  if (any_qual) b->any_qual = 1;
  comb_try_close(wc);
  pthread_cond_broadcast(&wc->cond);
  pthread_mutex_unlock(&wc->lock);
  return 0;
}

/* end of the input: whatever the open batch holds goes to the device */
void waveCombinerFlush(WaveCombiner *wc, int stop)
{
  pthread_mutex_lock(&wc->lock);
  wc->flush = 1;
  if (stop) wc->stop = 1;
  if (wc->open >= 0 && wc->slot[wc->open].nmembers > 0) {
    wc->slot[wc->open].state = CS_CLOSED;
    wc->open = -1;
  }
  pthread_cond_broadcast(&wc->cond);
  pthread_mutex_unlock(&wc->lock);
}

void waveCombinerRestart(WaveCombiner *wc)
{
  pthread_mutex_lock(&wc->lock);
  wc->flush = wc->stop = 0;
  pthread_mutex_unlock(&wc->lock);
}

/* A device thread: waits for the next closed batch whose members have all delivered, runs it (seed tables,
 * smb_block_run, smb_block_fetch) and returns its slot (>= 0), or -1 when the combiner is stopped.  *errcode
 * = the error of the batch (its members see it in waveCombinerResults too). */
int waveCombinerRunNext(ErrMsg *errmsgp, WaveCombiner *wc, int ktuple_maxhit, int min_swatscor_below_max_arg,
			UCHAR min_basqval, short target_depth, short max_depth, RMAPFLG_t rmapflg,
			const ScoreMatrix *scormtxp, SeqFastq *any_read, const SeqSet *ssp, int *errcode)
{
  int s, rc;
  CombSlot *b = NULL;
  pthread_mutex_lock(&wc->lock);
  for (;;) {
    for (s = 0; s < wc->nslots; s++) {
      b = wc->slot + s;
      if (b->state == CS_CLOSED && b->ncopied == b->nmembers && !b->leader) break;
    }
    if (s < wc->nslots) break;
    if (wc->stop) { pthread_mutex_unlock(&wc->lock); return -1; }
    /* an open batch with all members in and an idle device thread: take it now (latency), do not wait for it to fill */
    if (wc->open >= 0 && wc->slot[wc->open].nmembers > 0 && wc->slot[wc->open].ncopied == wc->slot[wc->open].nmembers) {
      wc->slot[wc->open].state = CS_CLOSED;
      wc->open = -1;
      continue;
    }
    pthread_cond_wait(&wc->cond, &wc->lock);
  }
  b->leader = 1;
  b->state = CS_RUNNING;
  wc->nrunning++;
  wc->nbatches++;
  wc->nbatch_reads += (uint64_t) b->nreads;
  pthread_mutex_unlock(&wc->lock);
  {
    const double t0_ = wnow();
    rc = comb_run_gpu(errmsgp, NULL, b, ktuple_maxhit, min_swatscor_below_max_arg, min_basqval, target_depth, max_depth,
		      rmapflg, scormtxp, any_read, ssp);
    if (wave_timing())
      fprintf(stderr, "smalt_b200 timing: batch slot %d: %d blocks, %d reads, device round trip %.3f ms, at %.3f s\n",
	      s, b->nmembers, b->nreads, 1e3 * (wnow() - t0_), wnow());
  }
  pthread_mutex_lock(&wc->lock);
  wc->nrunning--;
  b->errcode = rc;
  b->state = CS_DONE;
  pthread_cond_broadcast(&wc->cond);
  pthread_mutex_unlock(&wc->lock);
  *errcode = rc;
  return s;
}

/* A worker turns the device outputs of ITS reads of a finished batch into results (results.c, emitf per read
 * in order) and gives its share of the batch back. */
int waveCombinerResults(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, WaveCombiner *wc, const WaveTicket *tk, SeqFastq **reads,
			const uint32_t *min_cover_arr, int min_swatscor_arg, short max_depth, RMAPFLG_t rmapflg,
			const ScoreMatrix *scormtxp, const ResultFilter *rsfp, const SeqSet *ssp, const SeqCodec *codecp,
			RMAPWAVE_EMITF *emitf, void *user)
{
  CombSlot *b = wc->slot + tk->slot;
  const int n = tk->n, first = tk->first;
  int errcode = b->errcode, i;
  SINGLEDONE sd;
  double tw = wnow(), tc = cnow();
  if (!errcode) {
    short matchscor, mismatchscor = 0, gapinitscor = 0, gapextscor = 0;
    /* penalties are those of the score matrix (rmap.c:1258-1266): the batch's profile was made from one of its reads */
    matchscor = scoreProfileGetAvgPenalties(&mismatchscor, &gapinitscor, &gapextscor, b->bw->prof);
    if ((short) (matchscor - mismatchscor) < 1 || gapextscor >= 0 || mismatchscor >= 0 || (short) (matchscor - gapinitscor) < 1)
      errcode = ERRCODE_ASSERT;
    (void) scormtxp;
    if (!errcode) {
      WGROW(w->jobs, w->jobs_alloc, n, WJOB);
      for (i = 0; i < n; i++) {
	WJOB *jb = w->jobs + i;
	jb->read = jb->seed_read = (uint32_t) (first + i); jb->min_cover = min_cover_arr[i]; jb->min_swatscor = min_swatscor_arg;
	jb->readp = reads[i]; jb->rsp = rmp->rsrp; jb->niv = -1; jb->iv_first = 0; jb->blank = 1;
      }
      sd.w = w; sd.errmsgp = errmsgp; sd.rsfp = rsfp; sd.jobs = w->jobs; sd.emitf = emitf; sd.user = user;
      errcode = dev_results(errmsgp, rmp, w, NULL, b->bw, b->brd, b->bc, (size_t) first, n, w->jobs, max_depth, rmapflg,
			    matchscor, ssp, codecp, single_done, &sd);
    }
  }
  WTICK(7);
  pthread_mutex_lock(&wc->lock);
  if (++b->ndone == b->nmembers) {
    b->state = CS_FREE;
    pthread_cond_broadcast(&wc->cond);
  }
  pthread_mutex_unlock(&wc->lock);
  return errcode;
}

/* ------------------------------------------------------------------------------------ */
/* paired-end blocks                                                                      */
/* ------------------------------------------------------------------------------------ */
/* rmapPair (rmap.c:1744-2112) for a block of pairs.  The common course of a pair -
 *   seeds of both mates; the mate with fewer seed hits ("mate 1") mapped without restriction;
 *   insert-size intervals from its best hits (setupInterValFromResultSet, rmap.c:354-436);
 *   the other mate mapped inside those intervals; proper pairs found, mate 1 confidently
 *   placed (rmap.c:1950-1969)
 * - runs as two wave passes over all pairs of the block (the same K1/K2/K3 batches as
 * single-end reads, hit lists restricted to the intervals in the second pass).  A pair that
 * leaves that course (no proper pair, weak mate 1: unrestricted second search and the rescue
 * with the on-the-fly k=5 index, rmap.c:1970-2061; a mate too short to be hashed) is marked
 * RMAPPAIR_FALLBACK and mapped from scratch by the reference's own rmapPair on a fiber
 * (shim_fiber.inc.c) - the course of a pair is a deterministic function of the pair, so the
 * replay gives the reference's result.  Every decision is made by the reference's own
 * functions (results.c / resultpairs.c / interval.c); nothing of their logic is restated. */

void rmapWaveGetPairStats(const RmapWave *w, uint64_t counts[4])
{
  counts[0] = w->n_pairs; counts[1] = w->n_pairs_fallback; counts[2] = w->n_pairs_pass3; counts[3] = w->n_pairs_pass4;
}

/* job of a restricted search (intervals copied from ivr): job slot j >= 0 in w->jobs, or the
 * per-pair slot of pair -1-j in w->pjob */
static int pair_job_restricted(RmapWave *w, int j, int read, int seed_read, uint32_t min_cover, int min_swatscor,
			       SeqFastq *readp, ResultSet *rsp, const InterVal *ivr, const SETSIZ_t *soffs)
{
  const int niv = interValNum(ivr);
  WJOB *jb = (j >= 0) ? w->jobs + j : w->pjob + (-1 - j);
  int k;
  WGROW(w->iv, w->iv_alloc, w->niv + (size_t) niv + 1, WIVAL);
  jb->read = (uint32_t) read; jb->seed_read = (uint32_t) seed_read;
  jb->min_cover = min_cover; jb->min_swatscor = min_swatscor;
  jb->readp = readp; jb->rsp = rsp;
  jb->niv = niv; jb->iv_first = (uint32_t) w->niv; jb->blank = 0;
  for (k = 0; k < niv; k++) {
    SEQLEN_t lo, hi;
    SEQNUM_t sx;
    WIVAL *iv = w->iv + w->niv++;
    if (interValGet(&lo, &hi, &sx, NULL, k, ivr)) return ERRCODE_ASSERT;
    iv->lo = soffs[sx] + lo; iv->hi = soffs[sx] + hi + 1; iv->sx = (int32_t) sx;
  }
  return ERRCODE_SUCCESS;
}

int rmapPairWave(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int npairs, SeqFastq **reads, const uint32_t *mincov,
		 int d_min, int d_max, RSLTPAIRLIB_t pairlibcode, int ktuple_maxhit, int min_swatscor,
		 UCHAR min_basqval, short target_depth, short max_depth, RMAPFLG_t rmapflg,
		 const ScoreMatrix *scormtxp, const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
		 unsigned char *status)
{
  int errcode, p, nj;
  const SETSIZ_t *soffs;
  seqSetGetOffsets(ssp, &soffs);
  if (npairs < 1) return ERRCODE_SUCCESS;
  if (!(rmapflg & RMAPFLG_SEQBYSEQ) ||
      (rmapflg & (RMAPFLG_NOSHRTINFO | RMAPFLG_SPLIT | RMAPFLG_CMPLXW | RMAPFLG_ALLPAIR)) ||
      !rmp->mmp || !rmp->ivr || !rmp->pairp)
    return ERRCODE_ARGINVAL;
  w->cigar_mode = 0;   /* pairs are reported by the reference's reportWrite */
  if ((size_t) npairs > w->ps_alloc) {
    const size_t na = (size_t) npairs + 64;
    struct PairState_ *hp = (struct PairState_ *) realloc(w->ps, na * sizeof(*hp));
    if (!hp) return ERRCODE_NOMEM;
    memset(hp + w->ps_alloc, 0, (na - w->ps_alloc) * sizeof(*hp));
    w->ps = hp;
    w->ps_alloc = na;
  }
  if ((errcode = wave_seed(errmsgp, w, 2 * npairs, reads, ktuple_maxhit, min_basqval))) return errcode;
  WGROW(w->jobs, w->jobs_alloc, npairs, WJOB);
  WGROW(w->pjob, w->pjob_alloc, npairs, WJOB);

  /* pass 1: the mate with fewer hits, unrestricted (rmap.c:1870-1919) */
  for (p = 0, nj = 0; p < npairs; p++) {
    struct PairState_ *ps = w->ps + p;
    const smb_seed_info *inf = w->info + 4 * (size_t) p;
    const int err_read = inf[0].err ? inf[0].err : inf[1].err, err_mate = inf[2].err ? inf[2].err : inf[3].err;
    WJOB *jb;
    if (!ps->rs[0] && (!(ps->rs[0] = resultSetCreate(PAIR_RESULT_BLKSZ, 64)) || !(ps->rs[1] = resultSetCreate(PAIR_RESULT_BLKSZ, 64))))
      return ERRCODE_NOMEM;
    ps->status = RMAPPAIR_DONE;
    if (err_read || err_mate) { ps->status = RMAPPAIR_FALLBACK; continue; }
    resultSetBlank(ps->rs[0]);
    resultSetBlank(ps->rs[1]);
    /* calcTotalNumberOfHits (rmap.c:1079-1084) of both mates */
    ps->rare_mate = (unsigned char) ((uint32_t) (inf[0].nhit_all + inf[1].nhit_all) > (uint32_t) (inf[2].nhit_all + inf[3].nhit_all));
    ps->pairflg = (RSLTPAIRFLG_t) (RSLTPAIRFLG_PAIRED | (ps->rare_mate ? RSLTPAIRFLG_RAREMATE : 0));
    jb = w->jobs + nj++;
    jb->read = jb->seed_read = (uint32_t) (2 * p + ps->rare_mate);
    jb->min_cover = mincov[jb->read]; jb->min_swatscor = min_swatscor;
    jb->readp = reads[jb->read]; jb->rsp = ps->rs[ps->rare_mate];
    jb->niv = -1; jb->iv_first = 0; jb->blank = 0;
  }
  if ((errcode = wave_pass(errmsgp, rmp, w, nj, w->jobs, w->info, ktuple_maxhit, MINSCOR_BELOW_MAX_BEST, target_depth,
			   max_depth, rmapflg, scormtxp, htp, ssp, codecp, NULL, NULL)))
    return errcode;
  for (p = 0, nj = 0; p < npairs; p++) /* errors of pass 1 (the reference exits on them) -> reference code */
    if (w->ps[p].status == RMAPPAIR_DONE && w->rd[nj++].errcode) w->ps[p].status = RMAPPAIR_FALLBACK;

  /* pass 2: the other mate inside the insert-size intervals around mate 1 (rmap.c:1921-1948) */
  w->niv = 0;
  for (p = 0, nj = 0; p < npairs; p++) {
    struct PairState_ *ps = w->ps + p;
    const int i1 = 2 * p + ps->rare_mate, i2 = 2 * p + !ps->rare_mate;
    if (ps->status != RMAPPAIR_DONE) continue;
    ps->mapq1 = resultSetGetMappingScore(ps->rs[ps->rare_mate], &ps->swscor1);
    if (setupInterValFromResultSet(rmp->ivr, d_min, d_max, reads[i1], reads[i2], htp, ssp, ps->rs[ps->rare_mate])) {
      ps->status = RMAPPAIR_FALLBACK;
      continue;
    }
    interValPrune(rmp->ivr);
    if ((errcode = pair_job_restricted(w, nj++, i2, i2, mincov[i2], min_swatscor, reads[i2], ps->rs[!ps->rare_mate],
				       rmp->ivr, soffs)))
      return errcode;
  }
  if ((errcode = wave_pass(errmsgp, rmp, w, nj, w->jobs, w->info, ktuple_maxhit, MINSCOR_BELOW_MAX_BEST, target_depth,
			   max_depth, rmapflg, scormtxp, htp, ssp, codecp, NULL, NULL)))
    return errcode;

  /* does the pair end here? (rmap.c:1950-1969) */
  for (p = 0, nj = 0; p < npairs; p++) {
    struct PairState_ *ps = w->ps + p;
    const int i1 = 2 * p + ps->rare_mate, i2 = 2 * p + !ps->rare_mate;
    ps->stage = PST_END;
    if (ps->status != RMAPPAIR_DONE) continue;
    if (w->rd[nj++].errcode) { ps->status = RMAPPAIR_FALLBACK; continue; }
    errcode = resultSetFindProperPairs(rmp->pairp, d_min, d_max, MAXNUM_PAIRS_TOTAL, 0, pairlibcode, ps->rs[0], ps->rs[1]);
    if (errcode && errcode != ERRCODE_PAIRNUM) { ps->status = RMAPPAIR_FALLBACK; continue; }
    ps->swscor2_restricted = 0;
    ps->n_proper = 0;
    resultSetGetMappingScore(ps->rs[!ps->rare_mate], &ps->swscor2_restricted);
    resultSetGetNumberOfPairs(&ps->n_proper, rmp->pairp);
    if (ps->n_proper < 1 || ps->mapq1 < MAPSCORE_UNIQUE_MAPPED_1ST ||
	!scorIsAboveFractMax(ps->swscor2_restricted, ps->swscor1, MINFRACT_MAXSCOR_2ND, reads[i2], reads[i1]))
      ps->stage = PST_SECOND_UNRESTRICTED;
    else
      ps->pairflg = (RSLTPAIRFLG_t) (ps->pairflg | (ps->rare_mate == 0 ? RSLTPAIRFLG_RESTRICT_2nd : RSLTPAIRFLG_RESTRICT_1st));
  }

  /* pass 3: no proper pair or mate 1 not placed with confidence -> unrestricted search for the
   * other mate (rmap.c:1970-1996) */
  for (p = 0, nj = 0; p < npairs; p++) {
    struct PairState_ *ps = w->ps + p;
    const int i2 = 2 * p + !ps->rare_mate;
    WJOB *jb;
    if (ps->status != RMAPPAIR_DONE || ps->stage != PST_SECOND_UNRESTRICTED) continue;
    if (ps->n_proper < 1) resultSetBlank(ps->rs[!ps->rare_mate]);
    jb = w->jobs + nj++;
    jb->read = jb->seed_read = (uint32_t) i2;
    jb->min_cover = mincov[i2]; jb->min_swatscor = min_swatscor;
    jb->readp = reads[i2]; jb->rsp = ps->rs[!ps->rare_mate];
    jb->niv = -1; jb->iv_first = 0; jb->blank = 0;
  }
  w->n_pairs_pass3 += (uint64_t) nj;
  if (nj && (errcode = wave_pass(errmsgp, rmp, w, nj, w->jobs, w->info, ktuple_maxhit, MINSCOR_BELOW_MAX_BEST,
				 target_depth, max_depth, rmapflg, scormtxp, htp, ssp, codecp, NULL, NULL)))
    return errcode;

  /* something better for mate 1 near the new hits of the other mate?  (rmap.c:1998-2061):
   * intervals around them; the k=5 index of those intervals is built by the reference's own
   * hashTableSetUp on the host (setupFineHashTable, rmap.c:495-517) and copied out per pair */
  w->niv = 0;
  w->nftab = 0;
  {
    int n4m = 0, n4f = 0, fine_k = 0, fine_s = 0;
    UCHAR nskip_main;
    const UCHAR ktup_main = hashTableGetKtupLen(htp, &nskip_main);
    for (p = 0, nj = 0; p < npairs; p++) {
      struct PairState_ *ps = w->ps + p;
      const int i1 = 2 * p + ps->rare_mate, i2 = 2 * p + !ps->rare_mate;
      int mapq2, swscor2 = 0;
      SEQLEN_t rlen;
      if (ps->status != RMAPPAIR_DONE || ps->stage != PST_SECOND_UNRESTRICTED) continue;
      ps->stage = PST_END;
      if (w->rd[nj++].errcode) { ps->status = RMAPPAIR_FALLBACK; continue; }
      mapq2 = resultSetGetMappingScore(ps->rs[!ps->rare_mate], &swscor2);
      if (!(mapq2 > MAPSCORE_UNIQUE_MAPPED_1ST || swscor2 > ps->swscor2_restricted || swscor2 > ps->swscor1)) continue;
      ps->swscor1_2ndbest = 0;
      resultSetGetScorStats(ps->rs[ps->rare_mate], NULL, NULL, &ps->swscor1_2ndbest, NULL);
      if (setupInterValFromResultSet(rmp->ivr, d_min, d_max, reads[i2], reads[i1], htp, ssp, ps->rs[!ps->rare_mate])) {
	ps->status = RMAPPAIR_FALLBACK;
	continue;
      }
      interValPrune(rmp->ivr);
      seqFastqGetConstSequence(reads[i1], &rlen, NULL);
      if (ktup_main > rlen) continue;
      errcode = setupFineHashTable(rmp->htflyp, rmp->bfp->sqbfp, ssp, rmp->ivr, htp, codecp);
      if (errcode && errcode != ERRCODE_MAXKPOS) { ps->status = RMAPPAIR_FALLBACK; continue; }
      if (errcode) {
	ps->stage = PST_RESCUE_MAIN; /* too many positions for the on-the-fly index: main index, restricted */
	n4m++;
      } else {
	int typ, wordlen, nsk, nbk, nbl;
	uint32_t npos, nwords;
	const uint32_t *idx, *pos, *widx, *pidx;
	size_t nkeys, need;
	smbShimHashTableArrays(rmp->htflyp, &typ, &wordlen, &nsk, &nbk, &nbl, &npos, &nwords, &idx, &pos, &widx, &pidx);
	if (!fine_k) { fine_k = wordlen; fine_s = nsk; }
	if (typ != HASHIDXTYP_PERFECT || wordlen != fine_k || nsk != fine_s || wordlen > 12) {
	  ps->status = RMAPPAIR_FALLBACK; /* (sampling step changed by hashTableReset, rmap.c:503-510) */
	  continue;
	}
	nkeys = (size_t) 1 << (2 * wordlen);
	need = w->nftab + nkeys + 1 + npos + 2;
	WGROW(w->ftab, w->ftab_alloc, need, uint32_t);
	if ((size_t) n4f + 1 > w->ftabs_alloc) {
	  const size_t na = (size_t) n4f + 256;
	  w->ftabs = (smb_small_index *) realloc(w->ftabs, na * sizeof(smb_small_index));
	  w->ftab_off = (size_t *) realloc(w->ftab_off, 2 * na * sizeof(size_t));
	  w->ftab_read = (uint32_t *) realloc(w->ftab_read, na * sizeof(uint32_t));
	  w->f_off = (uint64_t *) realloc(w->f_off, na * sizeof(uint64_t));
	  w->f_len = (uint32_t *) realloc(w->f_len, na * sizeof(uint32_t));
	  if (!w->ftabs || !w->ftab_off || !w->ftab_read || !w->f_off || !w->f_len) return ERRCODE_NOMEM;
	  w->ftabs_alloc = na;
	}
	memcpy(w->ftab + w->nftab, idx, (nkeys + 1) * sizeof(uint32_t));
	if (npos) memcpy(w->ftab + w->nftab + nkeys + 1, pos, (size_t) npos * sizeof(uint32_t));
	w->ftab_off[2 * n4f] = w->nftab; w->ftab_off[2 * n4f + 1] = w->nftab + nkeys + 1;
	w->ftabs[n4f].npos = npos;
	w->nftab = need;
	ps->stage = PST_RESCUE_FINE;
	n4f++;
      }
      /* the job of this pair goes after those of the same kind: intervals are kept per pair */
      if ((errcode = pair_job_restricted(w, -1 - p, i1, i1, mincov[i1], ps->swscor1_2ndbest, reads[i1],
					 ps->rs[ps->rare_mate], rmp->ivr, soffs)))
	return errcode;
    }
    /* pass 4 (main index): restricted search for mate 1 with its block seed tables (rmap.c:2048-2060) */
    if (n4m) {
      for (p = 0, nj = 0; p < npairs; p++)
	if (w->ps[p].status == RMAPPAIR_DONE && w->ps[p].stage == PST_RESCUE_MAIN) w->jobs[nj++] = w->pjob[p];
      if ((errcode = wave_pass(errmsgp, rmp, w, nj, w->jobs, w->info, ktuple_maxhit, MINSCOR_BELOW_MAX_BEST, target_depth,
			       max_depth, rmapflg, scormtxp, htp, ssp, codecp, NULL, NULL)))
	return errcode;
      for (p = 0, nj = 0; p < npairs; p++)
	if (w->ps[p].status == RMAPPAIR_DONE && w->ps[p].stage == PST_RESCUE_MAIN && w->rd[nj++].errcode)
	  w->ps[p].status = RMAPPAIR_FALLBACK;
    }
    /* pass 4 (on-the-fly indexes): full seed tables of mate 1 against the index of its pair
     * (initRMAPINFO on htflyp, rmap.c:2026-2030) - one K1 batch with per-read tables - then the
     * restricted search with those (rmap.c:2032-2046) */
    if (n4f) {
      int rc, k = 0;
      smb_seed_info *info4;
      WPIN(info4, WB_INFO4, 2 * (size_t) n4f, smb_seed_info);
      w->info4 = info4;
      for (p = 0, nj = 0; p < npairs; p++) {
	struct PairState_ *ps = w->ps + p;
	if (ps->status != RMAPPAIR_DONE || ps->stage != PST_RESCUE_FINE) continue;
	w->ftabs[k].idx = w->ftab + w->ftab_off[2 * k];
	w->ftabs[k].pos = w->ftab + w->ftab_off[2 * k + 1];
	w->ftab_read[k] = (uint32_t) k;
	w->jobs[nj] = w->pjob[p];
	w->jobs[nj].seed_read = (uint32_t) k;
	w->f_off[k] = w->read_off[w->jobs[nj].read];
	w->f_len[k] = w->read_len[w->jobs[nj].read];
	nj++; k++;
      }
      if ((rc = smb_seed_batch_tables(w->ctx, fine_k, fine_s, w->ftabs, n4f, w->ftab_read, w->f_off, w->f_len, n4f,
				      w->any_qual ? w->qual : NULL, 0, 0, min_basqval, 0, info4)))
	return gpu_fail(errmsgp, w, rc);
      w->ms_k1 += smb_last_kernel_ms(w->ctx);
      if ((errcode = wave_pass(errmsgp, rmp, w, nj, w->jobs, info4, ktuple_maxhit, MINSCOR_BELOW_MAX_BEST, target_depth,
			       max_depth, rmapflg, scormtxp, rmp->htflyp, ssp, codecp, NULL, NULL)))
	return errcode;
      for (p = 0, nj = 0; p < npairs; p++)
	if (w->ps[p].status == RMAPPAIR_DONE && w->ps[p].stage == PST_RESCUE_FINE && w->rd[nj++].errcode)
	  w->ps[p].status = RMAPPAIR_FALLBACK;
      w->nreads = 0; /* the block seed tables are gone from the device */
    }
    w->n_pairs_pass4 += (uint64_t) (n4m + n4f);
  }
  for (p = 0; p < npairs; p++) {
    status[p] = w->ps[p].status;
    if (status[p] != RMAPPAIR_DONE) w->n_pairs_fallback++;
  }
  w->n_pairs += (uint64_t) npairs;
  return ERRCODE_SUCCESS;
}

/* tail of rmapPair (rmap.c:2095-2109) for a pair that rmapPairWave completed; the result sets
 * stay valid until the next rmapPairWave */
int rmapPairWaveFinish(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int p, int d_min, int d_max,
		       RSLTPAIRLIB_t pairlibcode, const ResultFilter *rsfp, SeqFastq *readp, SeqFastq *matep,
		       const ResultSet **rsltp, const ResultSet **rslt_matep, const ResultPairs **pairp,
		       RSLTPAIRFLG_t *pairflg)
{
  struct PairState_ *ps = w->ps + p;
  int errcode;
  if ((size_t) p >= w->ps_alloc || ps->status != RMAPPAIR_DONE) return ERRCODE_ASSERT;
  /* same sequence of calls on the ResultPairs object as in rmapPair */
  errcode = resultSetFindProperPairs(rmp->pairp, d_min, d_max, MAXNUM_PAIRS_TOTAL, 0, pairlibcode, ps->rs[0], ps->rs[1]);
  if (errcode && errcode != ERRCODE_PAIRNUM) ERRMSGNO(errmsgp, errcode);
  if ((errcode = resultSetFindPairs(rmp->pairp, ps->pairflg, pairlibcode, d_min, d_max, ps->rs[0], ps->rs[1])))
    ERRMSGNO(errmsgp, errcode);
  if ((errcode = resultSetFilterResults(ps->rs[0], rsfp, readp))) ERRMSGNO(errmsgp, errcode);
  if ((errcode = resultSetFilterResults(ps->rs[1], rsfp, matep))) ERRMSGNO(errmsgp, errcode);
  *rsltp = ps->rs[0]; *rslt_matep = ps->rs[1]; *pairp = rmp->pairp; *pairflg = ps->pairflg;
  return errcode;
}
