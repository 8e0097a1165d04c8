/* shim_hot.c - the reference's hot-path C interfaces on top of libsmalt_b200.so.
 *
 * This is the drop-in boundary of SURVEY.md section 8b: the objects swsimd.o, alignment.o
 * and hashhit.o of the reference's libsgm.a are replaced by this file, which implements the
 * same symbols with the same argument meaning and error codes -
 *     swsimd.h:47-55      swSIMDAlignStriped
 *     alignment.h:67-175  aliScoreDiffStr, aliRsltSetCreate/Delete/Reset/GetSize/FetchData,
 *                         aliSmiWatInBand, aliSmiWatInBandFast
 *     hashhit.h:90-322    hashCreateHitInfo/DeleteHitInfo, hashCollectHitInfo(Short),
 *                         hashCalcHitInfoCoverDeficit/NumberOfHits, hashHitInfoCalcHitNumbers,
 *                         hashCreateHitList/DeleteHitList/BlankHitList,
 *                         hashCollectHitsForSegment, hashGetHitListData
 * - by calling the batched C ABI (include/smalt_b200.h).  There is NO CPU implementation of
 * the DP or of the seed lookup in here: every function either runs the CUDA kernels (as a
 * batch of one when called through the reference's one-call-at-a-time API) or serves results
 * that the wave orchestrator (rmap_wave.c) computed on the GPU for a whole block of reads and
 * injected with the smbShim* functions.  Anything not supported fails loudly.
 *
 * The opaque types (HashHitInfo, HashHitList, AliRsltSet) are defined here; callers only use
 * them through the accessors above (segment.c:412,:475,:775,:1671; results.c:1867-1894).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <limits.h>
#include <time.h>

#include "elib.h"
#include "sequence.h"
#include "score.h"
#include "alibuffer.h"
#include "alignment.h"
#include "swsimd.h"
#include "diffstr.h"
#include "hashidx.h"
#include "hashhit.h"
#include "shim.h"

/* ------------------------------------------------------------------------------------ */
/* global GPU state: one "root" context per process holding the index + packed reference */
/* ------------------------------------------------------------------------------------ */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static smb_ctx *g_root;             /* owns index + packed reference */
static const HashTable *g_root_htp; /* table uploaded to g_root */
static smb_ctx *g_aux;              /* one-call context for any OTHER table (the on-the-fly k=5 index of
				     * rmap.c:495-517): its index is uploaded per seed call, so that the
				     * arrays shared by the worker contexts are never overwritten */
static int g_device = -1;

/* fiber scheduler (shim_fiber.inc.c): batches the one-call API over many reads */
static struct SmbFiberPool_ *fiber_pool_current(void);
static int fiber_seed_lookup(HashHitInfo *h, int is_reverse, int is_short, uint32_t maxhit_per_tuple,
			     uint32_t maxhit_total, int basq, const SeqFastq *seqp, const HashTable *htp);
static int fiber_hits(HashHitList *hlp, uint64_t lo, uint64_t hi, uint32_t nhit_max, int mode, HashHitInfo *h,
		      int *done);
static int fiber_dp(int kind, int *score, AliRsltSet *rssp, const ScoreProfile *profp, const char *useq, int uslen,
		    int l_edge, int r_edge, int pl, int pr, int ul, int ur, int minscore, int minscorlen, int *done);
enum { FOP_NONE, FOP_NOP, FOP_WAITORDER, FOP_HITS, FOP_SW, FOP_BANDFAST, FOP_BANDALI };

static double shim_now(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define SHIM_TIMING(what, t0) do { if (getenv("SMALT_B200_TIMING")) fprintf(stderr, "smalt_b200 timing: %s %.3f s\n", what, shim_now() - (t0)); } while (0)

/* a feature of the reference interface that the device path does not implement: one message, then the
 * call fails with an error code the caller handles like any other failure (no abort, no CPU fallback) */
static void shim_unsupported(const char *what)
{
  static int said;
  if (!said++) fprintf(stderr, "smalt_b200: %s is not implemented on the B200 path (include/smalt_b200.h, \"not covered\")\n", what);
}

static void shim_die(const char *what)
{
  fprintf(stderr, "smalt_b200: %s\n", what);
  abort();
}

static int shim_device(void)
{
  if (g_device < 0) {
    const char *e = getenv("SMALT_B200_DEVICE");
    const char *lr = getenv("LOCAL_RANK");
    g_device = e ? atoi(e) : (lr ? atoi(lr) : 0);
  }
  return g_device;
}

int smbShimDevice(void) { return shim_device(); }
double smbShimNow(void) { return shim_now(); }

extern void smbShimHashTableArrays(const HashTable *htp, int *typ, int *wordlen, int *nskip,
				   int *nbits_key, int *nbits_lo, uint32_t *npos, uint32_t *nwords,
				   const uint32_t **idx, const uint32_t **pos,
				   const uint32_t **wordidx, const uint32_t **posidx);

static int upload_index(smb_ctx *ctx, const HashTable *htp)
{
  int typ, wordlen, nskip, nbk, nbl;
  uint32_t npos, nwords;
  const uint32_t *idx, *pos, *widx, *pidx;
  smbShimHashTableArrays(htp, &typ, &wordlen, &nskip, &nbk, &nbl, &npos, &nwords, &idx, &pos, &widx, &pidx);
  return smb_index_upload(ctx, typ, wordlen, nskip, nbk, nbl, npos, nwords, idx, pos, widx, pidx);
}

/* pack the reference set 3 bits/base exactly as the .sma store (sequence.c:1360-1424) */
static int upload_refseq(smb_ctx *ctx, const SeqSet *ssp, const SeqCodec *codecp)
{
  const SETSIZ_t *soffs;
  const SEQNUM_t nseq = seqSetGetOffsets(ssp, &soffs);
  const uint64_t nbases = soffs[nseq] + 1;
  const size_t nwords = nbases / 10 + 1;
  uint32_t *words = (uint32_t *) calloc(nwords, sizeof(uint32_t));
  uint64_t *so = (uint64_t *) malloc(((size_t) nseq + 1) * sizeof(uint64_t));
  SeqFastq *buf = seqFastqCreate(0, SEQTYP_FASTA);
  SEQNUM_t s;
  int errcode = 0;
  if (!words || !so || !buf) return ERRCODE_NOMEM;
  for (s = 0; s <= nseq; s++) so[s] = soffs[s];
  for (s = 0; s < nseq && !errcode; s++) {
    SEQLEN_t len, i;
    char cod;
    const char *p;
    if ((errcode = seqSetFetchSegmentBySequence(buf, s, 0, 0, ssp, codecp))) break;
    p = seqFastqGetConstSequence(buf, &len, &cod);
    if (cod == SEQCOD_ASCII) {
      if ((errcode = seqFastqEncode(buf, codecp))) break;
      p = seqFastqGetConstSequence(buf, &len, &cod);
    }
    for (i = 0; i < len; i++) {
      const uint64_t b = soffs[s] + i;
      words[b / 10] |= ((uint32_t) (p[i] & 7)) << (3 * (9 - (unsigned) (b % 10)));
    }
  }
  if (!errcode) {
    const uint64_t b = nbases - 1; /* final terminator */
    words[b / 10] |= 7u << (3 * (9 - (unsigned) (b % 10)));
    errcode = smb_refseq_upload(ctx, words, nwords, nbases, so, (int) nseq);
  }
  seqFastqDelete(buf);
  free(words);
  free(so);
  return errcode;
}

/* Index construction on the GPU for `smalt_b200 index` (hashTableSetUp of a whole sequence set,
 * called from shim_hashidx.c): the k-mer grid bookkeeping of doWordsInSeq (hashidx.c:465-531) per
 * sequence on the host, everything else in csrc/index_build.cu.  Returns < 0 when the GPU path
 * does not apply (no device, a sequence shorter than k, ...): the caller then runs the
 * reference's CPU builder - index construction is not the mapping hot path. */
int smbShimIndexBuild(const SeqSet *ssp, const SeqCodec *codecp, int k, int nskip, int typ, int nbits_key,
		      int nbits_lo, uint32_t *npos, uint32_t *nwords, uint32_t *tuplectr_out, float *ms)
{
  const SETSIZ_t *soffs;
  const SEQNUM_t nseq = seqSetGetOffsets(ssp, &soffs);
  smb_index_seq *grid;
  smb_index_info info;
  long long tuplectr = 0, offs = 0;
  SEQNUM_t s;
  int rc;
  if (nseq < 1 || nseq > INT_MAX) return -1;
  if (!(grid = (smb_index_seq *) calloc((size_t) nseq, sizeof(*grid)))) return -1;
  for (s = 0; s < nseq; s++) {
    const long long L = (long long) (soffs[s + 1] - soffs[s]);
    long long n_k, ktup_i, d;
    if (L < k) { free(grid); return -1; }
    n_k = (L - k - offs >= 0) ? (L - k - offs) / nskip + 1 : 0;
    grid[s].start = soffs[s]; grid[s].offs = (uint32_t) offs; grid[s].n_k = (uint32_t) n_k;
    grid[s].tup_base = (uint32_t) tuplectr;
    if (n_k > 0) ktup_i = nskip - (L - 1 - (offs + (n_k - 1) * nskip + k - 1));
    else ktup_i = k + offs - L;
    tuplectr += n_k;
    d = k - ktup_i;
    offs = d % nskip;
    if (offs) offs = nskip - offs;
    tuplectr += (k - ktup_i + offs) / nskip;
    if (tuplectr > 0xFFFFFFFFll || n_k > 0xFFFFFFFFll) { free(grid); return -1; }
  }
  pthread_mutex_lock(&g_lock);
  rc = 0;
  if (!g_root && smb_ctx_create(&g_root, shim_device())) rc = -1;
  if (!rc && upload_refseq(g_root, ssp, codecp)) rc = -1;
  if (!rc && smb_index_build(g_root, k, nskip, typ, nbits_key, nbits_lo, grid, (int) nseq, &info)) {
    fprintf(stderr, "smalt_b200: GPU index construction failed: %s\n", smb_last_error(g_root));
    rc = -1;
  }
  pthread_mutex_unlock(&g_lock);
  free(grid);
  if (rc) return rc;
  *npos = info.npos; *nwords = info.nwords; *tuplectr_out = (uint32_t) tuplectr;
  if (ms) *ms = info.kernel_ms;
  return 0;
}

int smbShimIndexFetch(uint32_t *idx, uint32_t *pos, uint32_t *wordidx, uint32_t *posidx)
{
  int rc;
  pthread_mutex_lock(&g_lock);
  rc = g_root ? smb_index_fetch(g_root, idx, pos, wordidx, posidx) : -1;
  pthread_mutex_unlock(&g_lock);
  return rc ? ERRCODE_FAILURE : ERRCODE_SUCCESS;
}

int smbShimInit(const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
		const ScoreMatrix *scormtxp)
{
  int errcode = 0;
  pthread_mutex_lock(&g_lock);
  if (!g_root) {
    const double t0 = shim_now();
    errcode = smb_ctx_create(&g_root, shim_device());
    SHIM_TIMING("root context (CUDA init)", t0);
    if (errcode) {
      pthread_mutex_unlock(&g_lock);
      fprintf(stderr, "smalt_b200: cannot create a CUDA context on device %d (error %d); "
	      "there is no CPU fallback\n", shim_device(), errcode);
      return ERRCODE_FAILURE;
    }
    (void) scormtxp;
  }
  if (htp && g_root_htp != htp && (ssp || !g_root_htp)) {
    double t0 = shim_now();
    if (!(errcode = upload_index(g_root, htp))) g_root_htp = htp;
    SHIM_TIMING("index upload", t0);
    t0 = shim_now();
    if (!errcode && ssp) errcode = upload_refseq(g_root, ssp, codecp);
    SHIM_TIMING("reference pack + upload", t0);
  }
  pthread_mutex_unlock(&g_lock);
  if (errcode) fprintf(stderr, "smalt_b200: GPU upload failed: %s\n", smb_last_error(g_root));
  return errcode;
}

smb_ctx *smbShimRootCtx(void) { return g_root; }

/* The uploaded index is identified by the HashTable pointer; a driver that deletes the table
 * (end of a mapping session) must say so, or a new table allocated at the same address would
 * be taken for the one already on the device. */
void smbShimForgetIndex(void)
{
  pthread_mutex_lock(&g_lock);
  g_root_htp = NULL;
  pthread_mutex_unlock(&g_lock);
}

int smbShimWorkerCtx(smb_ctx **ctxp, const ScoreMatrix *scormtxp)
{
  int errcode;
  if (!g_root) return ERRCODE_ASSERT;
  {
    const double t0 = shim_now();
    errcode = smb_ctx_create(ctxp, shim_device());
    SHIM_TIMING("worker context", t0);
  }
  if (errcode) return ERRCODE_FAILURE;
  if ((errcode = smb_ctx_share_index(*ctxp, g_root))) return errcode;
  if (getenv("SMALT_B200_WSPIN")) smb_ctx_set_spin(*ctxp, atoi(getenv("SMALT_B200_WSPIN")));   /* how worker threads wait */
  (void) scormtxp; /* penalties are set per block from the read profiles (smbShimSetScoring) */
  return ERRCODE_SUCCESS;
}

/* scoring of a context from a profile's penalties (score.c:682-709) */
static int ctx_scoring_from_profile(smb_ctx *ctx, const ScoreProfile *profp)
{
  short mismatch, gapinit, gapext;
  const short match = scoreProfileGetAvgPenalties(&mismatch, &gapinit, &gapext, profp);
  return smb_set_scoring(ctx, match, mismatch, gapinit, gapext);
}
int smbShimSetScoring(smb_ctx *ctx, const ScoreProfile *profp) { return ctx_scoring_from_profile(ctx, profp); }

/* ------------------------------------------------------------------------------------ */
/* AliRsltSet (alignment.c:149-170) and its accessors                                    */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  int score, qs, qe, rs, re;
  DiffStr diffstr;
} SHIMRESULT;

struct _AliRsltSet {
  SHIMRESULT *rsp;
  short nres;
  short n_alloc;
};

AliRsltSet *aliRsltSetCreate(const ScoreMatrix *smp, short blksz, short diffblksz,
			     int track_blksz, int track_thresh)
{
  AliRsltSet *p;
  (void) blksz; (void) diffblksz; (void) track_blksz; (void) track_thresh;
  if (smp != NULL) {   /* scaleALICPLX (alignment.c:268-305): floating-point rescoring inside K3 - not on the device */
    shim_unsupported("complexity-weighted Smith-Waterman scores (map -w)");
    return NULL;       /* constructors report failure as NULL; the caller ends with its own error message */
  }
  p = (AliRsltSet *) calloc(1, sizeof(*p));
  return p;
}

void aliRsltSetDelete(AliRsltSet *p)
{
  short i;
  if (p) {
    for (i = 0; i < p->n_alloc; i++) free(p->rsp[i].diffstr.dstrp);
    free(p->rsp);
  }
  free(p);
}

void aliRsltSetReset(AliRsltSet *p) { p->nres = 0; }
short aliRsltSetGetSize(const AliRsltSet *arp) { return arp->nres; }

int aliRsltSetFetchData(const AliRsltSet *arp, short idx, int *score, int *ps_start, int *ps_end,
			int *us_start, int *us_end, const DiffStr **dfsp)
{
  const SHIMRESULT *rp;
  if (idx >= arp->nres) return ERRCODE_FAILURE;
  rp = arp->rsp + idx;
  if (score) *score = rp->score;
  if (ps_start) *ps_start = rp->qs;
  if (ps_end) *ps_end = rp->qe;
  if (us_start) *us_start = rp->rs;
  if (us_end) *us_end = rp->re;
  if (dfsp) *dfsp = &rp->diffstr;
  return ERRCODE_SUCCESS;
}

int smbShimAliRsltSetAdd(AliRsltSet *p, int score, int qs, int qe, int rs, int re,
			 const unsigned char *diffstr, int difflen)
{
  SHIMRESULT *rp;
  if (p->nres >= p->n_alloc) {
    const int na = p->n_alloc ? 2 * p->n_alloc : 16;
    SHIMRESULT *hp;
    if (na > 32767) return ERRCODE_OVERFLOW; /* nres is a short (alignment.c:1251) */
    hp = (SHIMRESULT *) realloc(p->rsp, (size_t) na * sizeof(SHIMRESULT));
    if (!hp) return ERRCODE_NOMEM;
    memset(hp + p->n_alloc, 0, (size_t) (na - p->n_alloc) * sizeof(SHIMRESULT));
    p->rsp = hp;
    p->n_alloc = (short) na;
  }
  rp = p->rsp + p->nres;
  if (difflen > rp->diffstr.n_alloc) {
    unsigned char *hp = (unsigned char *) realloc(rp->diffstr.dstrp, (size_t) difflen + 16);
    if (!hp) return ERRCODE_NOMEM;
    rp->diffstr.dstrp = hp;
    rp->diffstr.n_alloc = difflen + 16;
    rp->diffstr.blksz = 64;
  }
  memcpy(rp->diffstr.dstrp, diffstr, (size_t) difflen);
  rp->diffstr.len = difflen;
  rp->score = score; rp->qs = qs; rp->qe = qe; rp->rs = rs; rp->re = re;
  p->nres++;
  return ERRCODE_SUCCESS;
}

/* aliScoreDiffStr (alignment.c:179-225): score of a given alignment string against the
 * profile; host-side helper of results.c:1638 (not a kernel). */
int aliScoreDiffStr(int *swscor, const char *unprofiled_seqp, int unprofiled_seqlen,
		    unsigned int profiled_offs, const DIFFSTR_T *diffstrp, int diffstrlen,
		    const ScoreProfile *scpp)
{
  unsigned int plen;
  signed char gap_init, gap_ext;
  signed char *const *sc = scoreGetProfile(NULL, &plen, &gap_init, &gap_ext, scpp);
  int i, rs = 0, open = 0;
  *swscor = 0;
  for (i = 0; i < diffstrlen && diffstrp[i]; i++) {
    unsigned count = diffstrp[i] & DIFFSTR_COUNTMASK, typ = diffstrp[i] >> DIFFSTR_TYPSHIFT, j;
    if (typ == DIFFCOD_M || (typ == DIFFCOD_S && diffstrp[i + 1])) count++;
    if (count > 0) {
      open = 0;
      for (j = 0; j < count; j++) {
	*swscor += sc[unprofiled_seqp[rs++] & SEQCOD_ALPHA_MASK][profiled_offs++];
	if (profiled_offs > plen || rs > unprofiled_seqlen) return ERRCODE_ASSERT;
      }
    }
    if (typ == DIFFCOD_I || typ == DIFFCOD_D) {
      if (open) *swscor -= gap_ext;
      else { *swscor -= gap_init; open = 1; }
      if (typ == DIFFCOD_I) {
	if (++profiled_offs > plen) return ERRCODE_ASSERT;
      } else if (++rs > unprofiled_seqlen) return ERRCODE_ASSERT;
    }
  }
  return diffstrp[i] ? ERRCODE_DIFFSTR : ERRCODE_SUCCESS;
}

/* ------------------------------------------------------------------------------------ */
/* one-call-at-a-time DP entry points: a GPU batch of one                                */
/* ------------------------------------------------------------------------------------ */
static unsigned char *profile_codes(const ScoreProfile *profp, unsigned int *qlen)
{
  short asiz;
  SEQLEN_t n, j;
  signed char *const *sc = scoreGetProfile(&asiz, &n, NULL, NULL, profp);
  const short match = scoreProfileGetAvgPenalties(NULL, NULL, NULL, profp);
  unsigned char *codes = (unsigned char *) malloc((size_t) n + 1);
  if (!codes) return NULL;
  for (j = 0; j < n; j++) {
    int c, code = -1;
    for (c = 0; c < 4; c++) if (sc[c][j] == match) { code = c; break; }
    if (code < 0) code = (sc[0][j] == 0) ? 5 : 4;
    codes[j] = (unsigned char) code;
  }
  *qlen = n;
  return codes;
}

static int single_arena(smb_ctx *ctx, const ScoreProfile *profp, const char *useq, int uslen,
			unsigned int *qlen)
{
  unsigned char *codes = profile_codes(profp, qlen), *arena;
  int errcode;
  if (!codes) return ERRCODE_NOMEM;
  arena = (unsigned char *) malloc((size_t) *qlen + (size_t) uslen + 1);
  if (!arena) { free(codes); return ERRCODE_NOMEM; }
  memcpy(arena, codes, *qlen);
  memcpy(arena + *qlen, useq, (size_t) uslen);
  errcode = ctx_scoring_from_profile(ctx, profp);
  if (!errcode) errcode = smb_arena_upload(ctx, arena, (size_t) *qlen + (size_t) uslen);
  free(codes);
  free(arena);
  return errcode ? ERRCODE_FAILURE : 0;
}

int swSIMDAlignStriped(int *maxscor, const AliBuffer *abp, const ScoreProfile *profp,
		       const char *unprofiled_seqp, int unprofiled_seqlen)
{
  smb_sw_task t;
  int32_t score = 0, err = 0;
  unsigned int qlen;
  int errcode;
  (void) abp;
  *maxscor = 0;
  if (fiber_pool_current()) {
    int done = 0;
    errcode = fiber_dp(FOP_SW, maxscor, NULL, profp, unprofiled_seqp, unprofiled_seqlen, 0, 0, 0, 0, 0, 0, 0, 0, &done);
    if (done) return errcode;
  }
  if (smbShimInit(NULL, NULL, NULL, NULL)) return ERRCODE_FAILURE;
  pthread_mutex_lock(&g_lock);
  if (!(errcode = single_arena(g_root, profp, unprofiled_seqp, unprofiled_seqlen, &qlen))) {
    memset(&t, 0, sizeof t);
    t.read_off = 0; t.read_len = qlen; t.ref_off = qlen; t.ref_len = (uint32_t) unprofiled_seqlen;
    errcode = smb_sw_score_batch(g_root, &t, 1, &score, &err);
  }
  pthread_mutex_unlock(&g_lock);
  if (errcode) return ERRCODE_FAILURE;
  if (!err) *maxscor = score;
  return err;
}

static void band_task(smb_band_task *t, unsigned int qlen, int uslen, int l_edge, int r_edge,
		      int pl, int pr, int ul, int ur, int minscore, int minscorlen)
{
  memset(t, 0, sizeof *t);
  t->read_off = 0; t->read_len = qlen; t->ref_off = qlen; t->ref_len = (uint32_t) uslen;
  t->l_edge = l_edge; t->r_edge = r_edge; t->p_left = pl; t->p_right = pr;
  t->u_left = ul; t->u_right = ur; t->minscore = minscore; t->minscorlen = minscorlen;
}

int aliSmiWatInBandFast(int *maxswscor, AliBuffer *bufp, const ScoreProfile *profp,
			const char *unprofiled_seqp, int unprofiled_seqlen, int l_edge, int r_edge,
			int profiled_left, int profiled_right, int unprofiled_left, int unprofiled_right)
{
  smb_band_task t;
  int32_t score = 0, err = 0;
  unsigned int qlen;
  int errcode;
  (void) bufp;
  if (fiber_pool_current()) {
    int done = 0;
    errcode = fiber_dp(FOP_BANDFAST, maxswscor, NULL, profp, unprofiled_seqp, unprofiled_seqlen, l_edge, r_edge,
		       profiled_left, profiled_right, unprofiled_left, unprofiled_right, 0, 0, &done);
    if (done) return errcode;
  }
  if (smbShimInit(NULL, NULL, NULL, NULL)) return ERRCODE_FAILURE;
  pthread_mutex_lock(&g_lock);
  if (!(errcode = single_arena(g_root, profp, unprofiled_seqp, unprofiled_seqlen, &qlen))) {
    band_task(&t, qlen, unprofiled_seqlen, l_edge, r_edge, profiled_left, profiled_right,
	      unprofiled_left, unprofiled_right, 0, 0);
    errcode = smb_band_score_batch(g_root, &t, 1, &score, &err);
  }
  pthread_mutex_unlock(&g_lock);
  if (errcode) return ERRCODE_FAILURE;
  if (!err) *maxswscor = score;
  return err;
}

int aliSmiWatInBand(AliRsltSet *rssp, AliBuffer *bufp, const ScoreProfile *profp,
		    const char *unprofiled_seqp, int unprofiled_seqlen, int l_edge, int r_edge,
		    int profiled_left, int profiled_right, int unprofiled_left, int unprofiled_right,
		    int minscore, int minscorlen)
{
  smb_band_task t;
  unsigned int qlen;
  int errcode;
  int32_t err = 0;
  size_t nres = 0, ndiff = 0, cap_r = 64, cap_d = 1 << 16, i;
  smb_ali_result *res = NULL;
  uint8_t *diff = NULL;
  uint32_t first[2];
  (void) bufp;
  if (fiber_pool_current()) {
    int done = 0;
    errcode = fiber_dp(FOP_BANDALI, NULL, rssp, profp, unprofiled_seqp, unprofiled_seqlen, l_edge, r_edge,
		       profiled_left, profiled_right, unprofiled_left, unprofiled_right, minscore, minscorlen, &done);
    if (done) return errcode;
  }
  if (smbShimInit(NULL, NULL, NULL, NULL)) return ERRCODE_FAILURE;
  pthread_mutex_lock(&g_lock);
  if (!(errcode = single_arena(g_root, profp, unprofiled_seqp, unprofiled_seqlen, &qlen))) {
    band_task(&t, qlen, unprofiled_seqlen, l_edge, r_edge, profiled_left, profiled_right,
	      unprofiled_left, unprofiled_right, minscore, minscorlen);
    for (;;) {
      res = (smb_ali_result *) realloc(res, cap_r * sizeof(*res));
      diff = (uint8_t *) realloc(diff, cap_d);
      errcode = smb_band_align_batch(g_root, &t, 1, res, cap_r, &nres, first, diff, cap_d, &ndiff, &err, NULL);
      if (errcode == SMB_ERR_CAPACITY && (nres > cap_r || ndiff > cap_d)) {
	if (nres > cap_r) cap_r = nres;
	if (ndiff > cap_d) cap_d = ndiff;
	continue;
      }
      break;
    }
  }
  pthread_mutex_unlock(&g_lock);
  if (!errcode && !err)
    for (i = 0; i < nres && !errcode; i++)
      errcode = smbShimAliRsltSetAdd(rssp, res[i].score, res[i].qs, res[i].qe, res[i].rs, res[i].re,
				     diff + res[i].diff_off, (int) res[i].diff_len);
  free(res);
  free(diff);
  if (errcode) return (errcode < 100) ? errcode : ERRCODE_FAILURE;
  return err;
}

/* ------------------------------------------------------------------------------------ */
/* HashHitInfo / HashHitList                                                              */
/* ------------------------------------------------------------------------------------ */
struct _HashHitInfo {
  smb_seed_info info;      /* computed on the GPU */
  unsigned char ktup, nskip;
  /* for the one-call API: what is needed to (re)build the device tables of this read */
  unsigned char *codes, *qual;
  uint32_t qlen, n_alloc;
  uint32_t maxhit_per_tuple, maxhit_total;
  int basq, is_short, has_qual, is_reverse;
  unsigned long serial;
  const HashTable *htp;    /* table of the last collecting call */
  void *pool;              /* fiber pool whose block seed batch holds this read's tables, or NULL */
  int pool_read;           /* read index in that batch */
};

struct _HashHitList {
  const uint64_t *sqdat;   /* borrowed (wave path) or == own */
  uint64_t *own;
  size_t own_alloc;
  int nhits;
  char is_reverse;
  uint32_t qlen;
  unsigned char ktup, nskip;
  char *qmask;
  size_t qmask_alloc;
};

static unsigned long g_serial;          /* id of the read whose tables are on g_root / g_aux */
static unsigned long g_resident_serial, g_aux_resident_serial;


HashHitInfo *hashCreateHitInfo(int blksz, const HashTable *htp)
{
  HashHitInfo *p = (HashHitInfo *) calloc(1, sizeof(*p));
  (void) blksz;
  if (p) p->ktup = hashTableGetKtupLen(htp, &p->nskip);
  return p;
}

void hashDeleteHitInfo(HashHitInfo *p)
{
  if (p) { free(p->codes); free(p->qual); }
  free(p);
}

void smbShimHitInfoSet(HashHitInfo *p, const smb_seed_info *info) { p->info = *info; p->serial = 0; p->pool = NULL; }
const smb_seed_info *smbShimHitInfoGet(const HashHitInfo *p) { return &p->info; }

uint32_t hashCalcHitInfoCoverDeficit(const HashHitInfo *hip) { return hip->info.cover_deficit; }

uint32_t hashCalcHitInfoNumberOfHits(const HashHitInfo *hhip, HASHNUM_t maxhit_per_tuple)
{
  (void) maxhit_per_tuple; /* evaluated on the device with the cut-off of the collecting call */
  return hhip->info.nhit_all;
}

uint32_t hashHitInfoCalcHitNumbers(const HashHitInfo *hhip, uint32_t *nhit_rank)
{
  *nhit_rank = hhip->info.nhit_rank;
  return hhip->info.nhit_tot;
}

static smb_ctx *hitinfo_ctx(const HashHitInfo *h) { return (h->htp == g_root_htp) ? g_root : g_aux; }

/* (re)builds the device seed tables of one read on the one-call context of its table; g_lock held */
static int seed_one(HashHitInfo *h)
{
  uint64_t off = 0;
  smb_seed_info info[2];
  int errcode;
  smb_ctx *c;
  if (h->htp != g_root_htp) {
    if (!g_aux && smb_ctx_create(&g_aux, shim_device())) return ERRCODE_FAILURE;
    if (upload_index(g_aux, h->htp)) return ERRCODE_FAILURE;
    g_aux_resident_serial = 0;
  }
  c = hitinfo_ctx(h);
  if ((errcode = smb_arena_upload(c, h->codes, h->qlen))) return ERRCODE_FAILURE;
  errcode = smb_seed_batch(c, &off, &h->qlen, 1, h->has_qual ? h->qual : NULL, h->maxhit_per_tuple,
			   h->maxhit_total, h->basq, h->is_short, info, NULL, NULL, NULL, NULL, NULL, NULL);
  if (errcode) return ERRCODE_FAILURE;
  h->info = info[h->is_reverse ? 1 : 0];
  if (c == g_root) g_resident_serial = h->serial; else g_aux_resident_serial = h->serial;
  return 0;
}
static int hitinfo_resident(const HashHitInfo *h)
{
  return (h->htp == g_root_htp) ? g_resident_serial == h->serial : g_aux_resident_serial == h->serial;
}

static int collect_info(HashHitInfo *h, int is_reverse, int is_short, uint32_t maxhit_per_tuple,
			uint32_t maxhit_total, int basq, const SeqFastq *seqp, const HashTable *htp)
{
  SEQLEN_t len, i;
  char cod;
  const char *p = seqFastqGetConstSequence(seqp, &len, &cod);
  const char *q = seqFastqGetConstQualityFactors(seqp, NULL, NULL);
  int errcode;
  if (cod != SEQCOD_MANGLED) return ERRCODE_SEQCODE;
  if (smbShimInit(htp, NULL, NULL, NULL)) return ERRCODE_FAILURE;
  h->htp = htp;
  h->pool = NULL;
  h->ktup = hashTableGetKtupLen(htp, &h->nskip);
  if (fiber_pool_current() &&
      fiber_seed_lookup(h, is_reverse, is_short, maxhit_per_tuple, maxhit_total, basq, seqp, htp))
    return h->info.err;
  if (len + 1 > h->n_alloc) {
    h->codes = (unsigned char *) realloc(h->codes, (size_t) len + 64);
    h->qual = (unsigned char *) realloc(h->qual, (size_t) len + 64);
    h->n_alloc = len + 64;
    if (!h->codes || !h->qual) return ERRCODE_NOMEM;
  }
  for (i = 0; i < len; i++) h->codes[i] = (unsigned char) p[i];
  h->has_qual = (q != NULL);
  if (q) memcpy(h->qual, q, len);
  h->qlen = len; h->is_reverse = is_reverse; h->is_short = is_short;
  h->maxhit_per_tuple = maxhit_per_tuple; h->maxhit_total = maxhit_total; h->basq = basq;
  pthread_mutex_lock(&g_lock);
  h->serial = ++g_serial;
  errcode = seed_one(h);
  pthread_mutex_unlock(&g_lock);
  if (errcode) return errcode;
  return h->info.err;
}

int hashCollectHitInfo(HashHitInfo *hhip, unsigned char is_reverse, unsigned char basq_thresh,
		       SEQLEN_t seq_start, SEQLEN_t seq_end, const SeqFastq *seqp, const HashTable *htp)
{
  if (seq_start != 0 || seq_end != 0) {   /* seed tables of a read segment (hashhit.c:538-548): split reads, map -p */
    shim_unsupported("seed tables of a read segment (split-read mode, map -p)");
    return ERRCODE_ARGINVAL;
  }
  return collect_info(hhip, is_reverse, 0, 0, 0, basq_thresh, seqp, htp);
}

int hashCollectHitInfoShort(HashHitInfo *hhip, unsigned char is_reverse, HASHNUM_t maxhit_per_tuple,
			    HASHNUM_t maxhit_total, unsigned char basq_thresh, const SeqFastq *seqp,
			    const HashTable *htp)
{
  return collect_info(hhip, is_reverse, 1, maxhit_per_tuple, maxhit_total, basq_thresh, seqp, htp);
}

HashHitList *hashCreateHitList(int maxnhits)
{
  HashHitList *p = (HashHitList *) calloc(1, sizeof(*p));
  (void) maxnhits;
  return p;
}

void hashDeleteHitList(HashHitList *p)
{
  if (p) { free(p->own); free(p->qmask); }
  free(p);
}

void hashBlankHitList(HashHitList *p)
{
  if (p) {
    p->nhits = 0;
    p->sqdat = p->own;
    if (p->qmask) memset(p->qmask, HITQUAL_NOHIT, p->qlen);
  }
}

static int hitlist_qmask(HashHitList *p, uint32_t qlen)
{
  if ((size_t) qlen + 1 > p->qmask_alloc) {
    char *hp = (char *) realloc(p->qmask, (size_t) qlen + 512);
    if (!hp) return ERRCODE_NOMEM;
    p->qmask = hp;
    p->qmask_alloc = (size_t) qlen + 512;
  }
  /* initHitList -> blankHitList (hashhit.c:1224-1231): all NOHIT; the segment path never sets
   * NORMHIT (hashhit.c:1416-1546), which segLstFillHits relies on (segment.c:782-788 scans
   * until a 0 byte, so the mask is 0-terminated like the reference's calloc'ed block) */
  memset(p->qmask, HITQUAL_NOHIT, qlen);
  memset(p->qmask + qlen, 0, p->qmask_alloc - qlen);
  p->qlen = qlen;
  return 0;
}

/* wave path: serve a list computed by smb_hits_batch for a whole block of reads */
int smbShimHitListSet(HashHitList *p, const uint64_t *sqdat, int nhits, int is_reverse, uint32_t qlen,
		      unsigned char ktup, unsigned char nskip)
{
  int errcode = hitlist_qmask(p, qlen);
  if (errcode) return errcode;
  p->sqdat = sqdat;
  p->nhits = nhits;
  p->is_reverse = (char) (is_reverse != 0);
  p->ktup = ktup;
  p->nskip = nskip;
  return 0;
}

/* one hit list through the one-call contexts (mode: smb_hit_req.use_short) */
static int collect_hits(HashHitList *hlp, uint64_t lo, uint64_t hi, uint32_t nhit_max, int mode, HashHitInfo *h)
{
  smb_hit_req rq;
  uint64_t first[2], qfirst[2];
  int32_t err = 0;
  size_t tot = 0;
  int errcode, done = 0;
  smb_ctx *c;
  if ((errcode = hitlist_qmask(hlp, h->qlen))) return errcode;
  if (h->pool) {
    errcode = fiber_hits(hlp, lo, hi, nhit_max, mode, h, &done);
    if (done) return errcode;
  }
  if (!h->serial) shim_die("hit list requested for a HashHitInfo whose seed tables are not on the device");
  memset(&rq, 0, sizeof rq);
  rq.lo = lo; rq.hi = hi; rq.read = 0; rq.nhit_max = nhit_max;
  rq.strand = (uint8_t) (h->is_reverse != 0); rq.use_short = (uint8_t) mode;
  pthread_mutex_lock(&g_lock);
  if (!hitinfo_resident(h)) errcode = seed_one(h);
  c = hitinfo_ctx(h);
  if (!errcode) {
    errcode = smb_hits_batch(c, &rq, 1, 0, hlp->own, hlp->own_alloc, &tot, first, &err);
    if (errcode == SMB_ERR_CAPACITY && tot > hlp->own_alloc) {
      hlp->own = (uint64_t *) realloc(hlp->own, (tot + 1024) * sizeof(uint64_t));
      hlp->own_alloc = tot + 1024;
      errcode = hlp->own ? smb_hits_batch(c, &rq, 1, 0, hlp->own, hlp->own_alloc, &tot, first, &err)
	: ERRCODE_NOMEM;
    }
    if (!errcode && mode == 2) /* seeds marked NORMHIT / MULTIHIT (hashhit.c:1632-1650) */
      errcode = smb_hits_qmask(c, (uint8_t *) hlp->qmask, h->qlen, qfirst);
  }
  pthread_mutex_unlock(&g_lock);
  if (errcode) return ERRCODE_FAILURE;
  hlp->sqdat = hlp->own;
  hlp->nhits = (int) tot;
  hlp->is_reverse = (char) (h->is_reverse != 0);
  hlp->ktup = h->ktup;
  hlp->nskip = h->nskip;
  return (err == SMB_ERRCODE_ALLOCBOUNDARY) ? ERRCODE_SUCCESS : err;
}

int hashCollectHitsForSegment(HashHitList *hlp, SETSIZ_t segmoffs_lo, SETSIZ_t segmoffs_hi,
			      HASHNUM_t nhit_max, unsigned char use_short_hitinfo,
			      const HashHitInfo *hhip, const HashTable *htp, const HashHitFilter *hhfp)
{
  (void) htp;
  if (hhfp) {   /* (no caller in the smalt driver passes a filter) */
    shim_unsupported("hit filters (HashHitFilter)");
    return ERRCODE_ARGINVAL;
  }
  return collect_hits(hlp, segmoffs_lo, segmoffs_hi, nhit_max, use_short_hitinfo ? 1 : 0, (HashHitInfo *) hhip);
}

/* whole-set list (rmap.c:320-346, >= 512 reference sequences): K1 mode 2 */
int hashCollectHitsUsingCutoff(HashHitList *hlp, HASHNUM_t max_nhit_per_tup, const HashTable *htp,
			       const HashHitInfo *hip)
{
  (void) htp;
  return collect_hits(hlp, 0, 0, max_nhit_per_tup, 2, (HashHitInfo *) hip);
}

const uint64_t *hashGetHitListData(int *nhits, char *is_reverse, uint32_t *qlen, unsigned char *ktup,
				   unsigned char *nskip, const char **qmask, const HashHitList *hlp)
{
  if (nhits) *nhits = hlp->nhits;
  if (is_reverse) *is_reverse = hlp->is_reverse;
  if (qlen) *qlen = hlp->qlen;
  if (ktup) *ktup = hlp->ktup;
  if (qmask) *qmask = hlp->qmask;
  if (nskip) *nskip = hlp->nskip;
  return hlp->sqdat;
}

#include "shim_fiber.inc.c"
