/* shim_fiber.inc.c - fiber scheduler behind the one-call hot-path API (included by shim_hot.c).
 *
 * The reference maps a read pair with up to five dependent mapSingleRead passes whose control
 * flow lives in rmapPair (rmap.c:1744-2112); each pass calls the hot-path functions one at a
 * time (hashCollectHitsForSegment per strand x sequence or interval, swSIMDAlignStriped /
 * aliSmiWatInBandFast per candidate, aliSmiWatInBand per surviving candidate).  Run as GPU
 * batches of one that serialises the device.  Instead of restating that control flow, the
 * driver runs the reference's UNMODIFIED per-item function (processMapArgs, smalt.c:1083) for
 * many items at once, each on its own small stack ("fiber") with its own RMap: a hot-path call
 * made on a fiber records its arguments and switches back to the scheduler; when every
 * runnable fiber is parked on a call, the scheduler executes all parked calls as ONE batch per
 * kernel type (smb_hits_batch, smb_sw_score_batch, smb_band_score_batch,
 * smb_band_align_batch) on the worker's CUDA stream and resumes the fibers with their results.
 * Fibers that finish an item pick the next item of the block, so the batches stay full until
 * the block runs out.
 *
 * Seed tables: hashCollectHitInfo[Short] of the reads of a block is computed up front for the
 * whole block (smbFiberPoolSeed = one smb_seed_batch) and served to the fibers without a
 * switch; hit-list requests then address those device-resident tables by read index.  Calls
 * that do not match the block batch (the on-the-fly k=5 index of rmap.c:495-517, other
 * parameters) take the one-call contexts (g_root / g_aux) as before.
 *
 * Order of random draws: the reference draws among equally good hits with drand48()
 * (results.c:2298, :2532, resultpairs.c:737) in read order.  __wrap_drand48() parks a fiber
 * until every earlier item of the block is finished, so one worker reproduces the sequence of
 * the reference's single worker thread (the driver wraps drand48 with it).
 *
 * No compute happens here: every parked call ends in a CUDA kernel.
 */
#include <sys/mman.h>
#include <stdint.h>

#if defined(__x86_64__)
/* smb_fiber_swap(void **save_sp, void *load_sp): callee-saved registers on the old stack,
 * switch stack pointers, restore from the new one */
__asm__(".text\n"
	".align 16\n"
	".type smb_fiber_swap,@function\n"
	"smb_fiber_swap:\n"
	"  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
	"  movq %rsp, (%rdi)\n"
	"  movq %rsi, %rsp\n"
	"  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
	"  ret\n"
	".size smb_fiber_swap,.-smb_fiber_swap\n");
void smb_fiber_swap(void **save_sp, void *load_sp) __attribute__((visibility("hidden")));
#define FIBER_NATIVE 1
#else
#include <ucontext.h>
#define FIBER_NATIVE 0
#endif

enum { FS_IDLE, FS_RUNNABLE, FS_BLOCKED, FS_DONE };

typedef struct {
  int kind;
  /* FOP_HITS */
  HashHitList *hlp;
  HashHitInfo *hip;
  uint64_t lo, hi;
  uint32_t nhit_max;
  int mode;
  /* DP ops */
  const ScoreProfile *profp;
  const char *useq;
  int uslen;
  int l_edge, r_edge, pl, pr, ul, ur, minscore, minscorlen;
  AliRsltSet *rssp;
  /* results */
  int score, err;
} FiberOp;

typedef struct {
#if FIBER_NATIVE
  void *sp;
#else
  ucontext_t uc;
#endif
  char *stack;
  int state, item;
  FiberOp op;
} Fiber;

typedef struct { void *p; size_t cap; } FBUF;
enum { FB_ARENA, FB_QUAL, FB_OFF, FB_LEN, FB_INFO, FB_REQ, FB_FIRST, FB_ERR, FB_SQ, FB_QM, FB_QMFIRST,
       FB_SWT, FB_SWS, FB_SWE, FB_BFT, FB_BFS, FB_BFE, FB_BAT, FB_BAE, FB_RES, FB_RESFIRST, FB_DIFF, FB_COUNT };

struct SmbFiberPool_ {
  smb_ctx *ctx;
  Fiber *fib;
  int nfib;
  size_t stack_sz;
#if FIBER_NATIVE
  void *sched_sp;
#else
  ucontext_t sched_uc;
#endif
  Fiber *cur;
  SMBFIBER_ITEMF *itemf;
  void *user;
  /* seed batch of the block */
  int nreads, reads_per_item;
  const SeqFastq **reads;
  size_t reads_alloc;
  smb_seed_info *info;
  int pre_short, pre_basq;
  uint32_t pre_mpt, pre_mtot;
  const HashTable *pre_htp;
  /* items */
  int n_items, next_item, done_upto;
  unsigned char *item_done;
  size_t item_alloc;
  int *opidx;               /* scratch: fiber index per op of a wave */
  int have_scoring;
  FBUF fb[FB_COUNT];
  smbFiberStats st;
};

static __thread struct SmbFiberPool_ *t_pool; /* pool whose fibers run on this thread right now */

static struct SmbFiberPool_ *fiber_pool_current(void) { return (t_pool && t_pool->cur) ? t_pool : NULL; }

static void *fbuf(struct SmbFiberPool_ *p, int which, size_t bytes)
{
  FBUF *b = &p->fb[which];
  if (bytes > b->cap) {
    smb_host_free(b->p);
    b->cap = bytes + bytes / 2 + 4096;
    if (!(b->p = smb_host_alloc(b->cap))) { b->cap = 0; shim_die("out of page-locked host memory"); }
  }
  return b->p;
}
/* like fbuf, but keeps the first `keep` bytes */
static void *fbuf_grow(struct SmbFiberPool_ *p, int which, size_t bytes, size_t keep)
{
  FBUF *b = &p->fb[which];
  if (bytes > b->cap) {
    const size_t nc = bytes + bytes / 2 + 4096;
    void *np = smb_host_alloc(nc);
    if (!np) shim_die("out of page-locked host memory");
    if (keep) memcpy(np, b->p, keep);
    smb_host_free(b->p);
    b->p = np;
    b->cap = nc;
  }
  return b->p;
}

/* ---------------------------------------------------------------------------------------- */
/* context switching                                                                        */
/* ---------------------------------------------------------------------------------------- */
static void fiber_entry(void)
{
  struct SmbFiberPool_ *p = t_pool;
  Fiber *f = p->cur;
  (*p->itemf)(p->user, f->item, (int) (f - p->fib));
  f->state = FS_DONE;
#if FIBER_NATIVE
  smb_fiber_swap(&f->sp, p->sched_sp);
#else
  swapcontext(&f->uc, &p->sched_uc);
#endif
  abort(); /* a finished fiber is never resumed */
}

static void fiber_start(struct SmbFiberPool_ *p, Fiber *f, int item)
{
  f->item = item;
  f->state = FS_RUNNABLE;
  f->op.kind = FOP_NONE;
#if FIBER_NATIVE
  {
    void **top = (void **) (f->stack + p->stack_sz);
    top[-1] = NULL;                    /* return address slot of fiber_entry: keeps rsp = 8 mod 16 at entry */
    top[-2] = (void *) fiber_entry;
    memset(top - 8, 0, 6 * sizeof(void *));
    f->sp = (void *) (top - 8);
  }
#else
  getcontext(&f->uc);
  f->uc.uc_stack.ss_sp = f->stack;
  f->uc.uc_stack.ss_size = p->stack_sz;
  f->uc.uc_link = NULL;
  makecontext(&f->uc, fiber_entry, 0);
#endif
}

static void fiber_resume(struct SmbFiberPool_ *p, Fiber *f)
{
  p->cur = f;
#if FIBER_NATIVE
  smb_fiber_swap(&p->sched_sp, f->sp);
#else
  swapcontext(&p->sched_uc, &f->uc);
#endif
  p->cur = NULL;
}

/* called on a fiber: park it on f->op and run the scheduler */
static void fiber_park(struct SmbFiberPool_ *p)
{
  Fiber *f = p->cur;
  f->state = FS_BLOCKED;
#if FIBER_NATIVE
  smb_fiber_swap(&f->sp, p->sched_sp);
#else
  swapcontext(&f->uc, &p->sched_uc);
#endif
}

/* ---------------------------------------------------------------------------------------- */
/* the hooks of the one-call API                                                            */
/* ---------------------------------------------------------------------------------------- */
static int fiber_seed_lookup(HashHitInfo *h, int is_reverse, int is_short, uint32_t maxhit_per_tuple,
			     uint32_t maxhit_total, int basq, const SeqFastq *seqp, const HashTable *htp)
{
  struct SmbFiberPool_ *p = t_pool;
  const Fiber *f = p->cur;
  int k;
  if (!p->nreads || htp != p->pre_htp || (is_short != 0) != p->pre_short || basq != p->pre_basq ||
      (is_short && (maxhit_per_tuple != p->pre_mpt || maxhit_total != p->pre_mtot)))
    return 0;
  for (k = 0; k < p->reads_per_item; k++) {
    const int r = f->item * p->reads_per_item + k;
    if (r < p->nreads && p->reads[r] == seqp) {
      SEQLEN_t len;
      seqFastqGetConstSequence(seqp, &len, NULL);
      h->info = p->info[2 * r + (is_reverse ? 1 : 0)];
      h->qlen = len;
      h->is_reverse = is_reverse; h->is_short = is_short;
      h->maxhit_per_tuple = maxhit_per_tuple; h->maxhit_total = maxhit_total; h->basq = basq;
      h->serial = 0;
      h->pool = p;
      h->pool_read = r;
      p->st.seeds_served++;
      return 1;
    }
  }
  return 0;
}

static int fiber_hits(HashHitList *hlp, uint64_t lo, uint64_t hi, uint32_t nhit_max, int mode, HashHitInfo *h,
		      int *done)
{
  struct SmbFiberPool_ *p = fiber_pool_current();
  FiberOp *op;
  *done = 0;
  if (!p || h->pool != p) {
    if (!h->serial) shim_die("hit list requested outside the fiber pool that holds the read's seed tables");
    return 0;
  }
  op = &p->cur->op;
  op->kind = FOP_HITS;
  op->hlp = hlp; op->hip = h; op->lo = lo; op->hi = hi; op->nhit_max = nhit_max; op->mode = mode;
  op->err = 0;
  fiber_park(p);
  *done = 1;
  return op->err;
}

static int fiber_dp(int kind, int *score, AliRsltSet *rssp, const ScoreProfile *profp, const char *useq, int uslen,
		    int l_edge, int r_edge, int pl, int pr, int ul, int ur, int minscore, int minscorlen, int *done)
{
  struct SmbFiberPool_ *p = fiber_pool_current();
  FiberOp *op;
  *done = 0;
  if (!p) return 0;
  op = &p->cur->op;
  op->kind = kind;
  op->profp = profp; op->useq = useq; op->uslen = uslen;
  op->l_edge = l_edge; op->r_edge = r_edge; op->pl = pl; op->pr = pr; op->ul = ul; op->ur = ur;
  op->minscore = minscore; op->minscorlen = minscorlen;
  op->rssp = rssp;
  op->score = 0; op->err = 0;
  fiber_park(p);
  *done = 1;
  if (score && !op->err) *score = op->score;
  return op->err;
}

int smbFiberYield(void)
{
  struct SmbFiberPool_ *p = fiber_pool_current();
  if (!p) return 0;
  p->cur->op.kind = FOP_NOP;
  fiber_park(p);
  return 1;
}

/* parks the calling fiber until all earlier items of the block are finished */
void smbFiberWaitOrder(void)
{
  struct SmbFiberPool_ *p = fiber_pool_current();
  if (!p) return;
  while (p->done_upto < p->cur->item) {
    p->cur->op.kind = FOP_WAITORDER;
    p->st.order_waits++;
    fiber_park(p);
  }
}

/* ---------------------------------------------------------------------------------------- */
/* execution of the parked calls of a wave                                                  */
/* ---------------------------------------------------------------------------------------- */
static void flush_fail(struct SmbFiberPool_ *p, const char *what, int rc)
{
  fprintf(stderr, "smalt_b200: GPU batch (%s) failed (%d): %s\n", what, rc, smb_last_error(p->ctx));
  abort(); /* no CPU fallback */
}

static void flush_hits(struct SmbFiberPool_ *p, int n)
{
  smb_hit_req *req = (smb_hit_req *) fbuf(p, FB_REQ, (size_t) n * sizeof(smb_hit_req));
  uint64_t *first = (uint64_t *) fbuf(p, FB_FIRST, ((size_t) n + 2) * sizeof(uint64_t));
  int32_t *errs = (int32_t *) fbuf(p, FB_ERR, ((size_t) n + 1) * sizeof(int32_t));
  uint64_t *sq;
  size_t need = 0, qbytes = 0;
  int i, rc, any2 = 0;
  for (i = 0; i < n; i++) {
    const FiberOp *op = &p->fib[p->opidx[i]].op;
    smb_hit_req *rq = req + i;
    memset(rq, 0, sizeof(*rq));
    rq->lo = op->lo; rq->hi = op->hi; rq->read = (uint32_t) op->hip->pool_read; rq->nhit_max = op->nhit_max;
    rq->strand = (uint8_t) (op->hip->is_reverse != 0); rq->use_short = (uint8_t) op->mode;
    if (op->mode == 2) any2 = 1;
    qbytes += op->hip->qlen;
  }
  if (p->fb[FB_SQ].cap < 64 * (size_t) n * sizeof(uint64_t)) fbuf(p, FB_SQ, 96 * (size_t) n * sizeof(uint64_t));
  for (;;) {
    sq = (uint64_t *) p->fb[FB_SQ].p;
    rc = smb_hits_batch(p->ctx, req, n, 0, sq, p->fb[FB_SQ].cap / sizeof(uint64_t), &need, first, errs);
    if (rc == SMB_ERR_CAPACITY && need > p->fb[FB_SQ].cap / sizeof(uint64_t)) {
      fbuf(p, FB_SQ, (need + need / 4 + 1024) * sizeof(uint64_t));
      continue;
    }
    if (rc) flush_fail(p, "hit lists", rc);
    break;
  }
  p->st.ms_k1 += smb_last_kernel_ms(p->ctx);
  {
    uint8_t *qm = NULL;
    uint64_t *qfirst = NULL;
    if (any2) {
      qm = (uint8_t *) fbuf(p, FB_QM, qbytes + 16);
      qfirst = (uint64_t *) fbuf(p, FB_QMFIRST, ((size_t) n + 2) * sizeof(uint64_t));
      if ((rc = smb_hits_qmask(p->ctx, qm, qbytes, qfirst))) flush_fail(p, "hit list masks", rc);
    }
    for (i = 0; i < n; i++) {
      FiberOp *op = &p->fib[p->opidx[i]].op;
      HashHitList *hlp = op->hlp;
      const HashHitInfo *h = op->hip;
      const size_t cnt = (size_t) (first[i + 1] - first[i]);
      if (cnt > hlp->own_alloc) {
	free(hlp->own);
	hlp->own_alloc = cnt + cnt / 2 + 256;
	if (!(hlp->own = (uint64_t *) malloc(hlp->own_alloc * sizeof(uint64_t)))) shim_die("out of memory");
      }
      if (cnt) memcpy(hlp->own, sq + first[i], cnt * sizeof(uint64_t));
      if (op->mode == 2) memcpy(hlp->qmask, qm + qfirst[i], h->qlen);
      hlp->sqdat = hlp->own;
      hlp->nhits = (int) cnt;
      hlp->is_reverse = (char) (h->is_reverse != 0);
      hlp->ktup = h->ktup;
      hlp->nskip = h->nskip;
      op->err = (errs[i] == SMB_ERRCODE_ALLOCBOUNDARY) ? ERRCODE_SUCCESS : errs[i];
    }
  }
  p->st.n_hits += (uint64_t) n;
}

static void band_task_at(smb_band_task *t, const FiberOp *op, uint64_t read_off, uint32_t qlen, uint64_t ref_off)
{
  memset(t, 0, sizeof(*t));
  t->read_off = read_off; t->read_len = qlen; t->ref_off = ref_off; t->ref_len = (uint32_t) op->uslen;
  t->l_edge = op->l_edge; t->r_edge = op->r_edge; t->p_left = op->pl; t->p_right = op->pr;
  t->u_left = op->ul; t->u_right = op->ur; t->minscore = op->minscore; t->minscorlen = op->minscorlen;
}

/* read codes of a profile appended to the wave arena (cf. profile_codes) */
static uint32_t arena_put_profile(uint8_t *dst, const ScoreProfile *profp)
{
  short asiz;
  SEQLEN_t n, j;
  signed char *const *sc = scoreGetProfile(&asiz, &n, NULL, NULL, profp);
  const short match = scoreProfileGetAvgPenalties(NULL, NULL, NULL, profp);
  const signed char *a = sc[0], *c = sc[1], *g = sc[2], *t = sc[3];
  for (j = 0; j < n; j++)
    dst[j] = (uint8_t) (a[j] == match ? 0 : c[j] == match ? 1 : g[j] == match ? 2 : t[j] == match ? 3 : (a[j] == 0 ? 5 : 4));
  return n;
}

static void flush_dp(struct SmbFiberPool_ *p, int n, int nsw, int nbf, int nba)
{
  size_t bytes = 0, pos = 0, nres = 0, ndiff = 0;
  uint8_t *arena;
  smb_sw_task *swt = (smb_sw_task *) fbuf(p, FB_SWT, ((size_t) nsw + 1) * sizeof(smb_sw_task));
  smb_band_task *bft = (smb_band_task *) fbuf(p, FB_BFT, ((size_t) nbf + 1) * sizeof(smb_band_task));
  smb_band_task *bat = (smb_band_task *) fbuf(p, FB_BAT, ((size_t) nba + 1) * sizeof(smb_band_task));
  int32_t *sws = (int32_t *) fbuf(p, FB_SWS, ((size_t) nsw + 1) * sizeof(int32_t));
  int32_t *swe = (int32_t *) fbuf(p, FB_SWE, ((size_t) nsw + 1) * sizeof(int32_t));
  int32_t *bfs = (int32_t *) fbuf(p, FB_BFS, ((size_t) nbf + 1) * sizeof(int32_t));
  int32_t *bfe = (int32_t *) fbuf(p, FB_BFE, ((size_t) nbf + 1) * sizeof(int32_t));
  int32_t *bae = (int32_t *) fbuf(p, FB_BAE, ((size_t) nba + 1) * sizeof(int32_t));
  int i, rc, isw = 0, ibf = 0, iba = 0;
  double t_dp0 = shim_now();
  for (i = 0; i < n; i++) {
    const FiberOp *op = &p->fib[p->opidx[i]].op;
    SEQLEN_t ql;
    scoreGetProfile(NULL, &ql, NULL, NULL, op->profp);
    bytes += (size_t) ql + (size_t) op->uslen;
  }
  arena = (uint8_t *) fbuf(p, FB_ARENA, bytes + 64);
  for (i = 0; i < n; i++) {
    FiberOp *op = &p->fib[p->opidx[i]].op;
    const uint64_t roff = pos;
    const uint32_t ql = arena_put_profile(arena + pos, op->profp);
    pos += ql;
    memcpy(arena + pos, op->useq, (size_t) op->uslen);
    if (op->kind == FOP_SW) {
      smb_sw_task *t = swt + isw;
      memset(t, 0, sizeof(*t));
      t->read_off = roff; t->read_len = ql; t->ref_off = pos; t->ref_len = (uint32_t) op->uslen;
      op->score = isw++;
      p->st.cells_k2 += (uint64_t) ql * (uint64_t) op->uslen;
    } else if (op->kind == FOP_BANDFAST) {
      band_task_at(bft + ibf, op, roff, ql, pos);
      op->score = ibf++;
    } else {
      band_task_at(bat + iba, op, roff, ql, pos);
      op->score = iba++;
    }
    pos += (size_t) op->uslen;
  }
  if (!p->have_scoring) {
    if ((rc = ctx_scoring_from_profile(p->ctx, p->fib[p->opidx[0]].op.profp))) flush_fail(p, "scoring", rc);
    p->have_scoring = 1;
  }
  { const double t_ = shim_now(); p->st.wall_stage += t_ - t_dp0; t_dp0 = t_; }
  if ((rc = smb_arena_upload(p->ctx, arena, pos))) flush_fail(p, "arena", rc);
  { const double t_ = shim_now(); p->st.wall_arena += t_ - t_dp0; t_dp0 = t_; }
  if (nsw) {
    if ((rc = smb_sw_score_batch(p->ctx, swt, nsw, sws, swe))) flush_fail(p, "SW scores", rc);
    p->st.ms_k2 += smb_last_kernel_ms(p->ctx);
    p->st.n_sw += (uint64_t) nsw;
    { const double t_ = shim_now(); p->st.wall_sw += t_ - t_dp0; t_dp0 = t_; }
  }
  if (nbf) {
    if ((rc = smb_band_score_batch(p->ctx, bft, nbf, bfs, bfe))) flush_fail(p, "band scores", rc);
    p->st.ms_k2 += smb_last_kernel_ms(p->ctx);
    p->st.n_bandfast += (uint64_t) nbf;
  }
  if (nba) {
    uint32_t *rfirst = (uint32_t *) fbuf(p, FB_RESFIRST, ((size_t) nba + 2) * sizeof(uint32_t));
    uint64_t cells = 0;
    if (p->fb[FB_RES].cap < ((size_t) nba + 16) * sizeof(smb_ali_result))
      fbuf(p, FB_RES, ((size_t) nba * 2 + 64) * sizeof(smb_ali_result));
    if (p->fb[FB_DIFF].cap < 48 * (size_t) nba) fbuf(p, FB_DIFF, 64 * (size_t) nba + 4096);
    for (;;) {
      const size_t cap_r = p->fb[FB_RES].cap / sizeof(smb_ali_result), cap_d = p->fb[FB_DIFF].cap;
      rc = smb_band_align_batch(p->ctx, bat, nba, (smb_ali_result *) p->fb[FB_RES].p, cap_r, &nres, rfirst,
				(uint8_t *) p->fb[FB_DIFF].p, cap_d, &ndiff, bae, &cells);
      if (rc == SMB_ERR_CAPACITY && (nres > cap_r || ndiff > cap_d)) {
	if (nres > cap_r) fbuf(p, FB_RES, (nres + 64) * sizeof(smb_ali_result));
	if (ndiff > cap_d) fbuf(p, FB_DIFF, ndiff + 4096);
	continue;
      }
      if (rc) flush_fail(p, "band alignments", rc);
      break;
    }
    p->st.ms_k3 += smb_last_kernel_ms(p->ctx);
    p->st.n_bandali += (uint64_t) nba;
    p->st.cells_k3 += cells;
    { const double t_ = shim_now(); p->st.wall_ba += t_ - t_dp0; t_dp0 = t_; }
    for (i = 0; i < n; i++) {
      FiberOp *op = &p->fib[p->opidx[i]].op;
      const smb_ali_result *res = (const smb_ali_result *) p->fb[FB_RES].p;
      const uint8_t *diff = (const uint8_t *) p->fb[FB_DIFF].p;
      uint32_t k;
      int e = 0;
      if (op->kind != FOP_BANDALI) continue;
      if (!bae[op->score])
	for (k = rfirst[op->score]; k < rfirst[op->score + 1] && !e; k++)
	  e = smbShimAliRsltSetAdd(op->rssp, res[k].score, res[k].qs, res[k].qe, res[k].rs, res[k].re,
				   diff + res[k].diff_off, (int) res[k].diff_len);
      op->err = e ? e : bae[op->score];
    }
  }
  for (i = 0; i < n; i++) {
    FiberOp *op = &p->fib[p->opidx[i]].op;
    if (op->kind == FOP_SW) { const int k = op->score; op->err = swe[k]; op->score = swe[k] ? 0 : sws[k]; }
    else if (op->kind == FOP_BANDFAST) { const int k = op->score; op->err = bfe[k]; op->score = bfe[k] ? 0 : bfs[k]; }
  }
}

/* executes every parked call; returns the number of fibers made runnable */
static int fiber_flush(struct SmbFiberPool_ *p)
{
  int i, n, released = 0, nsw = 0, nbf = 0, nba = 0;
  double t0 = shim_now(), t1;
  /* 1. hit lists */
  for (i = 0, n = 0; i < p->nfib; i++)
    if (p->fib[i].state == FS_BLOCKED && p->fib[i].op.kind == FOP_HITS) p->opidx[n++] = i;
  if (n) flush_hits(p, n);
  t1 = shim_now(); p->st.wall_hits += t1 - t0; t0 = t1;
  /* 2. DP calls share one arena upload */
  for (i = 0, n = 0; i < p->nfib; i++) {
    const Fiber *f = p->fib + i;
    if (f->state != FS_BLOCKED) continue;
    if (f->op.kind == FOP_SW) nsw++;
    else if (f->op.kind == FOP_BANDFAST) nbf++;
    else if (f->op.kind == FOP_BANDALI) nba++;
    else continue;
    p->opidx[n++] = i;
  }
  if (n) flush_dp(p, n, nsw, nbf, nba);
  t1 = shim_now(); p->st.wall_dp += t1 - t0;
  /* 3. release */
  for (i = 0; i < p->nfib; i++) {
    Fiber *f = p->fib + i;
    if (f->state != FS_BLOCKED) continue;
    if (f->op.kind == FOP_WAITORDER && p->done_upto < f->item) continue;
    f->state = FS_RUNNABLE;
    released++;
  }
  p->st.n_waves++;
  return released;
}

/* ---------------------------------------------------------------------------------------- */
/* pool                                                                                     */
/* ---------------------------------------------------------------------------------------- */
SmbFiberPool *smbFiberPoolCreate(int nfibers, size_t stack_bytes)
{
  struct SmbFiberPool_ *p;
  int i;
  const size_t pg = 4096;
  if (nfibers < 1) nfibers = 1;
  if (stack_bytes < 65536) stack_bytes = 65536;
  stack_bytes = (stack_bytes + pg - 1) / pg * pg;
  p = (struct SmbFiberPool_ *) calloc(1, sizeof(*p));
  if (!p) return NULL;
  p->nfib = nfibers;
  p->stack_sz = stack_bytes;
  p->fib = (Fiber *) calloc((size_t) nfibers, sizeof(Fiber));
  p->opidx = (int *) calloc((size_t) nfibers, sizeof(int));
  if (!p->fib || !p->opidx) { free(p->fib); free(p->opidx); free(p); return NULL; }
  for (i = 0; i < nfibers; i++) {
    /* lazily committed stack with a guard page below it */
    char *m = (char *) mmap(NULL, stack_bytes + pg, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == (char *) MAP_FAILED) { smbFiberPoolDelete(p); return NULL; }
    mprotect(m, pg, PROT_NONE);
    p->fib[i].stack = m + pg;
  }
  return p;
}

void smbFiberPoolDelete(SmbFiberPool *p)
{
  int i;
  if (!p) return;
  for (i = 0; i < p->nfib; i++)
    if (p->fib[i].stack) munmap(p->fib[i].stack - 4096, p->stack_sz + 4096);
  for (i = 0; i < FB_COUNT; i++) smb_host_free(p->fb[i].p);
  if (p->ctx) smb_ctx_destroy(p->ctx);
  free(p->fib); free(p->opidx); free(p->reads); free(p->item_done);
  free(p);
}

int smbFiberPoolSize(const SmbFiberPool *p) { return p->nfib; }
void smbFiberPoolGetStats(const SmbFiberPool *p, smbFiberStats *st) { *st = p->st; }

static int pool_ctx(struct SmbFiberPool_ *p)
{
  if (p->ctx) return 0;
  if (smbShimInit(NULL, NULL, NULL, NULL)) return ERRCODE_FAILURE;
  return smbShimWorkerCtx(&p->ctx, NULL);
}

/* Seed tables of the reads of a block (reads_per_item consecutive entries per item), computed
 * with the parameters the fibers' hashCollectHitInfo[Short] calls will use. */
int smbFiberPoolSeed(SmbFiberPool *p, int nreads, SeqFastq *const *reads, int reads_per_item, int is_short,
		     uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq, const HashTable *htp)
{
  uint64_t *off;
  uint32_t *len;
  uint8_t *arena, *qual;
  size_t tot = 0;
  int i, rc, any_qual = 0;
  p->nreads = 0;
  if (nreads < 1) return 0;
  if (htp != g_root_htp) return 0; /* not the uploaded table: the fibers take the one-call path */
  if (pool_ctx(p)) return ERRCODE_FAILURE;
  if ((size_t) nreads > p->reads_alloc) {
    free(p->reads);
    p->reads_alloc = (size_t) nreads + 256;
    if (!(p->reads = (const SeqFastq **) malloc(p->reads_alloc * sizeof(*p->reads)))) return ERRCODE_NOMEM;
  }
  off = (uint64_t *) fbuf(p, FB_OFF, (size_t) nreads * sizeof(uint64_t));
  len = (uint32_t *) fbuf(p, FB_LEN, (size_t) nreads * sizeof(uint32_t));
  p->info = (smb_seed_info *) fbuf(p, FB_INFO, 2 * (size_t) nreads * sizeof(smb_seed_info));
  for (i = 0; i < nreads; i++) {
    SEQLEN_t l;
    char cod;
    seqFastqGetConstSequence(reads[i], &l, &cod);
    if (cod != SEQCOD_MANGLED) return ERRCODE_SEQCODE;
    p->reads[i] = reads[i];
    off[i] = tot; len[i] = l; tot += l;
    if (seqFastqGetConstQualityFactors(reads[i], NULL, NULL)) any_qual = 1;
  }
  arena = (uint8_t *) fbuf(p, FB_ARENA, tot + 64);
  qual = (uint8_t *) fbuf(p, FB_QUAL, tot + 64);
  for (i = 0; i < nreads; i++) {
    SEQLEN_t l;
    const char *s = seqFastqGetConstSequence(reads[i], &l, NULL);
    const char *q = seqFastqGetConstQualityFactors(reads[i], NULL, NULL);
    memcpy(arena + off[i], s, l);
    if (any_qual) {
      if (q) memcpy(qual + off[i], q, l);
      else memset(qual + off[i], 0xff, l);
    }
  }
  if ((rc = smb_arena_upload(p->ctx, arena, tot)) ||
      (rc = smb_seed_batch(p->ctx, off, len, nreads, any_qual ? qual : NULL, is_short ? maxhit_per_tuple : 0,
			   is_short ? maxhit_total : 0, basq, is_short, p->info, NULL, NULL, NULL, NULL, NULL, NULL))) {
    fprintf(stderr, "smalt_b200: seed batch failed (%d): %s\n", rc, smb_last_error(p->ctx));
    return ERRCODE_FAILURE;
  }
  p->st.ms_k1 += smb_last_kernel_ms(p->ctx);
  p->nreads = nreads;
  p->reads_per_item = reads_per_item < 1 ? 1 : reads_per_item;
  p->pre_short = is_short != 0; p->pre_mpt = maxhit_per_tuple; p->pre_mtot = maxhit_total;
  p->pre_basq = basq; p->pre_htp = htp;
  p->st.n_seeded += (uint64_t) nreads;
  return 0;
}

/* Runs itemf(user, item, slot) for item = 0..nitems-1, each on a fiber (slot = index of the
 * fiber, < pool size; at most one item runs on a slot at a time). */
int smbFiberPoolRun(SmbFiberPool *p, int nitems, SMBFIBER_ITEMF *itemf, void *user)
{
  int i, nblocked, active;
  if (nitems < 1) return 0;
  if (t_pool) return ERRCODE_ASSERT; /* no nesting */
  if ((size_t) nitems > p->item_alloc) {
    free(p->item_done);
    p->item_alloc = (size_t) nitems + 1024;
    if (!(p->item_done = (unsigned char *) malloc(p->item_alloc))) return ERRCODE_NOMEM;
  }
  memset(p->item_done, 0, (size_t) nitems);
  p->n_items = nitems; p->next_item = 0; p->done_upto = 0;
  p->itemf = itemf; p->user = user;
  p->have_scoring = 0;
  for (i = 0; i < p->nfib; i++) p->fib[i].state = FS_IDLE;
  t_pool = p;
  for (i = 0; i < p->nfib && p->next_item < nitems; i++) fiber_start(p, p->fib + i, p->next_item++);
  for (;;) {
    const double t0 = shim_now();
    struct timespec c0_, c1_;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &c0_);
    nblocked = 0; active = 0;
    for (i = 0; i < p->nfib; i++) {
      Fiber *f = p->fib + i;
      while (f->state == FS_RUNNABLE) {
	fiber_resume(p, f);
	if (f->state == FS_DONE) {
	  p->item_done[f->item] = 1;
	  while (p->done_upto < nitems && p->item_done[p->done_upto]) p->done_upto++;
	  if (p->next_item < nitems) fiber_start(p, f, p->next_item++);
	  else f->state = FS_IDLE;
	}
      }
      if (f->state == FS_BLOCKED) nblocked++;
    }
    p->st.wall_host += shim_now() - t0;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &c1_);
    p->st.wall_stage -= 0; p->st.cpu_host += (c1_.tv_sec - c0_.tv_sec) + 1e-9 * (c1_.tv_nsec - c0_.tv_nsec);
    if (!nblocked) break;
    if (!fiber_flush(p)) { t_pool = NULL; shim_die("fiber scheduler: every fiber is parked and none can be released"); }
    (void) active;
  }
  t_pool = NULL;
  p->nreads = 0;
  p->st.n_items += (uint64_t) nitems;
  return 0;
}

/* drand48 of a mapping run (ld --wrap=drand48): draws happen in item order (see the header) */
extern double __real_drand48(void);
double __wrap_drand48(void)
{
  smbFiberWaitOrder();
  return __real_drand48();
}

/* CPU-only self test of the scheduler (no GPU call): every item yields a few times, some wait
 * for their turn; returns a checksum of the completion order and the order of the "draws" */
typedef struct { int *order, *draws; int n_order, n_draws; } FiberSelfTest;
static void selftest_item(void *user, int item, int slot)
{
  FiberSelfTest *t = (FiberSelfTest *) user;
  volatile char pad[2048];
  int k;
  (void) slot;
  for (k = 0; k < (item * 7) % 5; k++) { pad[(k * 131) % 2048] = (char) item; smbFiberYield(); }
  if (item % 3 == 0) { smbFiberWaitOrder(); t->draws[t->n_draws++] = item; }
  for (k = 0; k < item % 3; k++) smbFiberYield();
  t->order[t->n_order++] = item;
}
int smbFiberSelfTest(int nfibers, int nitems, int *order_out, int *draws_out, int *ndraws)
{
  FiberSelfTest t;
  SmbFiberPool *p = smbFiberPoolCreate(nfibers, 65536);
  int rc;
  if (!p) return -1;
  t.order = order_out; t.draws = draws_out; t.n_order = 0; t.n_draws = 0;
  rc = smbFiberPoolRun(p, nitems, selftest_item, &t);
  *ndraws = t.n_draws;
  smbFiberPoolDelete(p);
  return rc ? rc : t.n_order;
}
