/* fastmap.inc.c - block-parallel single-end mapping pipeline of the smalt_b200 driver
 * (included by smalt_main.c, which compiles the reference's smalt.c in place and therefore sees
 * its SmaltMapConst / SmaltMapArgs types).
 *
 * The reference pipeline is INPUT thread -> PROC threads -> OUTPUT thread over blocks of
 * nthreads*32 reads (smalt.c:1353-1386, threads.c): one thread parses FASTQ, one thread formats
 * and writes SAM.  With the hot path on the GPU those two serial stages cap the whole program at
 * a few hundred thousand reads/s.  Reads are independent, so this pipeline cuts the
 * input TEXT into byte ranges and lets every worker thread do, for its range,
 *     parse  (the reference's own seqFastqRead, through a memfd so that header/sequence parsing
 *             rules stay the reference's)
 *  -> encode + GPU waves (rmapSingleWave: K1 -> segment.c -> K2 -> replay -> K3 -> results.c)
 *  -> format (the reference's own reportWrite into a memory stream, shim_report.c)
 * and writes the formatted blocks in input order.  Only the boundary search between blocks and
 * the final fwrite are new code.
 *
 * Eligible: `map` of ONE plain-text 4-line FASTQ (or FASTA) file or memory buffer, text output
 * formats, modes covered by the wave path.  Everything else runs the reference pipeline
 * (smb_processArgBlock).  SMALT_B200_REFIO=1 forces the reference pipeline.
 */
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "fastprintf.h"

typedef int (FASTMAP_SINKF)(void *user, const char *buf, size_t len);

typedef struct {
  char *buf;
  size_t len;
  int done;
} FmBlockOut;

/* a block of reads on its way through the pipelined single-end path: parsed by one worker, mapped in a
 * device batch together with the blocks of other workers, turned into results by whichever worker is free */
typedef struct FmBlock_ {
  SeqFastq **reads;
  uint32_t *mincov;
  size_t n_alloc, n, chunk;
  WaveTicket tk;
  int inflight;
  struct FmBlock_ *next;
} FmBlock;

typedef struct FastMap_ {
  const SmaltMapConst *macop;
  SmaltMapArgs *maps;        /* one per worker (threadsGetMem(THRTASK_PROC)) */
  const ReportWriter *proto;
  const char *data;
  size_t len;
  int is_fasta;
  int refparse;              /* SMALT_B200_REFPARSE: every block through the reference's own parser */
  /* paired input: second file, cut at the same record numbers */
  const char *dataB;
  size_t lenB;
  struct FmLines_ *linesA, *linesB;
  size_t block_recs, nrecs;
  size_t chunk_bytes, nchunks;
  int nworkers;
  FASTMAP_SINKF *sinkf;
  void *sink_user;
  pthread_mutex_t lock;
  size_t next_chunk, next_out;
  int flushing;
  FmBlockOut *out;
  int errcode;
  uint64_t n_reads;
  /* pipelined path */
  pthread_cond_t cond;
  FmBlock *blocks, *free_blk, *ready;
  int nblocks, outstanding;
} FastMap;

typedef struct {
  FastMap *fm;
  int id;
  RmapWave *wave;            /* persists across fastmap runs (library mode) */
  ReportWriter *writer;
  Report *rep;
  SeqFastq **reads;
  uint32_t *mincov;
  size_t n_alloc;
  SeqFastq **mates;          /* paired input */
  uint32_t *mincov_m;
  size_t m_alloc;
  SeqFastq **pairs;          /* reads[i], mates[i] interleaved for pair_core */
  size_t pairs_alloc;
  PairWorker pw;
  int memfd;
  char fdpath[64];
  char *scratch, *scratch2;  /* names of a record whose header lines hold white space */
  size_t scratch_alloc, scratch2_alloc;
  FILE *keyfp;               /* stream handed to the report writer while its output is captured */
  char *keybuf;
  size_t keylen;
  ErrMsg *errmsgp;
  double ms_prev[3], wall_prev[11], cpu_prev[8], cand_prev;
  uint64_t counts_prev[5];
} FmWorker;

static FmWorker *g_fm_workers;
static int g_fm_nworkers;
/* device batches shared by the workers of the single-end pipeline (rmap_wave.c, WaveCombiner);
 * SMALT_B200_COMBINE=0: every worker maps its block on its own stream; SMALT_B200_BATCH: reads per batch */
static struct FmBlock_ *g_fm_blocks;   /* block objects of the pipelined path (their read buffers persist) */
static int g_fm_nblocks;
static WaveCombiner *g_fm_comb;
static FmWorker g_fm_comb_stats[8];   /* the batch slots' waves, for fm_collect_stats */

/* ---- block boundaries ---------------------------------------------------------------- */

static size_t fm_next_line(const char *d, size_t len, size_t p)
{ /* start of the first line that begins at or after p */
  const char *q;
  if (p == 0) return 0;
  if (p > len) return len;
  q = (const char *) memchr(d + p - 1, '\n', len - (p - 1));
  return q ? (size_t) (q - d) + 1 : len;
}

/* first record start at or after byte offset p */
static size_t fm_record_start(const FastMap *fm, size_t p)
{
  const char *d = fm->data;
  const size_t len = fm->len;
  size_t l0 = fm_next_line(d, len, p);
  int k;
  if (p == 0) return 0;
  if (fm->is_fasta) {
    while (l0 < len && d[l0] != '>') l0 = fm_next_line(d, len, l0 + 1);
    return l0;
  }
  /* 4-line FASTQ: a header line starts with '@' and the line two further down with '+'.
   * (a quality line may start with '@' too, but then the line two further down is the next
   * record's sequence line, which cannot start with '+') */
  for (k = 0; k < 5 && l0 < len; k++) {
    if (d[l0] == '@') {
      const size_t l1 = fm_next_line(d, len, l0 + 1);
      const size_t l2 = (l1 < len) ? fm_next_line(d, len, l1 + 1) : len;
      if (l2 < len && d[l2] == '+') return l0;
    }
    l0 = fm_next_line(d, len, l0 + 1);
  }
  return len;
}

/* checks that [start, end) is a whole number of 4-line records */
static int fm_check_fastq(const char *d, size_t start, size_t end, size_t *nrec)
{
  size_t p = start, line = 0;
  while (p < end) {
    const char *q = (const char *) memchr(d + p, '\n', end - p);
    const size_t e = q ? (size_t) (q - d) : end;
    if (e == p || (e == p + 1 && d[p] == '\r')) { /* blank line: only allowed between/after records */
      if (line & 3) return ERRCODE_FASTA;
      p = e + 1;
      continue;
    }
    if ((line & 3) == 0 && d[p] != '@') return ERRCODE_FASTA;
    if ((line & 3) == 2 && d[p] != '+') return ERRCODE_FASTA;
    line++;
    p = e + 1;
  }
  if (line & 3) return ERRCODE_FASTA;
  *nrec = line >> 2;
  return ERRCODE_SUCCESS;
}

/* ---- ordered output -------------------------------------------------------------------- */

static void fm_publish(FastMap *fm, size_t c, char *buf, size_t len, int errcode)
{
  smbShimCigarFlush();   /* this worker's device- / host-formatted record counts of the block */
  pthread_mutex_lock(&fm->lock);
  fm->out[c].buf = buf;
  fm->out[c].len = len;
  fm->out[c].done = 1;
  if (errcode && !fm->errcode) fm->errcode = errcode;
  while (!fm->flushing && fm->next_out < fm->nchunks && fm->out[fm->next_out].done) {
    FmBlockOut *o = fm->out + fm->next_out;
    int rc = 0;
    fm->flushing = 1;
    pthread_mutex_unlock(&fm->lock);
    if (o->len && !fm->errcode) rc = (*fm->sinkf)(fm->sink_user, o->buf, o->len);
    free(o->buf);
    o->buf = NULL;
    pthread_mutex_lock(&fm->lock);
    if (rc && !fm->errcode) fm->errcode = rc;
    fm->flushing = 0;
    fm->next_out++;
  }
  pthread_mutex_unlock(&fm->lock);
}

/* ---- worker ------------------------------------------------------------------------------ */

typedef struct {
  FmWorker *w;
  FILE *fp;
  size_t n;
} FmEmit;

static int fm_emit(void *user, int i, const ResultSet *rsltp)
{
  FmEmit *e = (FmEmit *) user;
  FmWorker *w = e->w;
  const SmaltMapConst *macop = w->fm->macop;
  int errcode;
  reportBlank(w->rep);
  if ((errcode = resultSetAddToReport(w->rep, macop->rsltouflg, rsltp))) return errcode;
  /* outputIOBuffArg (smalt.c:832-868) */
  if ((macop->menuflg & MENUFLAG_RELSCOR) &&
      (macop->outform == REPORTFMT_SAM || macop->outform == REPORTFMT_BAM))
    reportFixMultiplePrimary(w->rep);
  return smbShimReportWriteSAM(w->writer, w->reads[i], macop->ssp, macop->codecp, w->rep);
}

static int fm_worker_setup(FmWorker *w, FastMap *fm, int id)
{
  const SmaltMapConst *macop = fm->macop;
  w->fm = fm;
  w->id = id;
  if (!w->wave && !(w->wave = rmapWaveCreate(macop->htp, macop->ssp, macop->codecp, macop->scormtxp))) {
    fprintf(stderr, "smalt_b200: cannot set up the GPU context of a worker thread\n");
    return ERRCODE_FAILURE;
  }
  if (!w->writer && !(w->writer = smbShimReportWriterClone(fm->proto))) return ERRCODE_NOMEM;
  if (!w->rep && !(w->rep = reportCreate(0))) return ERRCODE_NOMEM;
  if (!w->errmsgp) ERRMSG_CREATE(w->errmsgp);
  if (w->memfd <= 0) {
    w->memfd = memfd_create("smalt_b200_block", 0);
    if (w->memfd < 0) return ERRCODE_NOFILE;
    snprintf(w->fdpath, sizeof(w->fdpath), "/proc/self/fd/%d", w->memfd);
  }
  return ERRCODE_SUCCESS;
}


/* header line -> name, the rule of readHeader (sequence.c:1056-1140): the prompt is dropped,
 * leading white space skipped, of every later run of white space the first character is kept,
 * one trailing white-space character is removed */
static size_t fm_header_name(char *dst, const char *line, size_t len)
{
  size_t i, n = 0;
  int was_space = 1;
  for (i = 1; i < len; i++) { /* line[0] is the prompt */
    const unsigned char c = (unsigned char) line[i];
    const int sp = (c == ' ' || (c >= 9 && c <= 13));
    if (was_space) {
      if (sp) continue;
      was_space = 0;
    } else if (sp) {
      was_space = 1;
    }
    dst[n++] = (char) c;
  }
  if (was_space && n > 0) n--;
  dst[n] = '\0';
  return n;
}

/* white space (' ', 9..13) in a sequence / quality line?  Eight bytes at a time: any byte below '!'
 * counts (control characters other than white space also send the block to the reference parser,
 * which is always right); with `high` also any byte beyond 0x7f. */
static int fm_has_space(const char *p, size_t len, int high)
{
  size_t i = 0;
  uint64_t acc = 0, hi = 0;
  for (; i + 8 <= len; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    acc |= (w - 0x2121212121212121ULL) & ~w;
    hi |= w;
  }
  for (; i < len; i++) {
    acc |= (uint64_t) ((unsigned char) p[i] < 0x21u) << 7;
    hi |= (unsigned char) p[i];
  }
  return ((acc | (high ? hi : 0)) & 0x8080808080808080ULL) != 0;
}

static int fm_reads_reserve(FmWorker *w, size_t n)
{
  if (n >= w->n_alloc) {
    const size_t na = w->n_alloc ? 2 * w->n_alloc : 1024;
    SeqFastq **hp = (SeqFastq **) realloc(w->reads, na * sizeof(SeqFastq *));
    uint32_t *mp = (uint32_t *) realloc(w->mincov, na * sizeof(uint32_t));
    if (hp) w->reads = hp;
    if (mp) w->mincov = mp;
    if (!hp || !mp) return ERRCODE_NOMEM;
    memset(w->reads + w->n_alloc, 0, (na - w->n_alloc) * sizeof(SeqFastq *));
    w->n_alloc = na;
  }
  if (!w->reads[n] && !(w->reads[n] = seqFastqCreate(0, SEQTYP_UNKNOWN))) return ERRCODE_NOMEM;
  return ERRCODE_SUCCESS;
}

/* Plain 4-line FASTQ records without white space inside the sequence / quality lines are
 * loaded straight from the text with seqFastqSetAscii (what infmtRead does for SAM/BAM input,
 * infmt.c:250-263), reproducing readHeader's name rule.  Returns 1 when the block contains
 * anything else (blank lines, CR, wrapped or ragged records): the caller then uses the
 * reference's own parser for the whole block. */
static int fm_parse_block_fast(FmWorker *w, const char *d, size_t start, size_t end, size_t *nreads)
{
  const SmaltMapConst *macop = w->fm->macop;
  size_t p = start, n = 0;
  int errcode;
  *nreads = 0;
  while (p < end) {
    const char *l[4], *nl;
    size_t ll[4], need;
    char *name, *qnam;
    size_t nl_, ql_;
    /* header and bases: up to the next line feed; '+' line: usually bare; qualities: as long as the bases */
    if (!(nl = (const char *) memchr(d + p, '\n', end - p))) return 1;
    l[0] = d + p; ll[0] = (size_t) (nl - l[0]); p += ll[0] + 1;
    if (p >= end || !(nl = (const char *) memchr(d + p, '\n', end - p))) return 1;
    l[1] = d + p; ll[1] = (size_t) (nl - l[1]); p += ll[1] + 1;
    if (p + 1 < end && d[p] == '+' && d[p + 1] == '\n') nl = d + p + 1;
    else if (p >= end || !(nl = (const char *) memchr(d + p, '\n', end - p))) return 1;
    l[2] = d + p; ll[2] = (size_t) (nl - l[2]); p += ll[2] + 1;
    l[3] = d + p; ll[3] = ll[1];
    if (p + ll[3] >= end || d[p + ll[3]] != '\n') return 1;      /* (a shorter line shows as a line feed inside, below) */
    p += ll[3] + 1;
    if (ll[0] < 1 || l[0][0] != '@' || ll[2] < 1 || l[2][0] != '+' || ll[1] < 1 ||
	fm_has_space(l[1], ll[1], 1) || fm_has_space(l[3], ll[3], 0))
      return 1;
    if ((errcode = fm_reads_reserve(w, n))) return errcode;
    if (!fm_has_space(l[0] + 1, ll[0] - 1, 0)) {   /* the usual header: one word */
      name = (char *) l[0] + 1; nl_ = ll[0] - 1;
    } else {
      need = ll[0] + 8;
      if (need > w->scratch_alloc) {
	char *hp = (char *) realloc(w->scratch, 2 * need);
	if (!hp) return ERRCODE_NOMEM;
	w->scratch = hp;
	w->scratch_alloc = 2 * need;
      }
      name = w->scratch;
      nl_ = fm_header_name(name, l[0], ll[0]);
      /* setSeq (sequence.c:780-803) also strips white space at both ends of what it is given */
      while (nl_ > 0 && isspace((unsigned char) name[nl_ - 1])) nl_--;
    }
    if (ll[2] == 1) { qnam = (char *) l[2]; ql_ = 0; }
    else {
      need = ll[2] + 8;
      if (need > w->scratch2_alloc) {
	char *hp = (char *) realloc(w->scratch2, 2 * need);
	if (!hp) return ERRCODE_NOMEM;
	w->scratch2 = hp;
	w->scratch2_alloc = 2 * need;
      }
      qnam = w->scratch2;
      ql_ = fm_header_name(qnam, l[2], ll[2]);
      while (ql_ > 0 && isspace((unsigned char) qnam[ql_ - 1])) ql_--;
    }
    /* bases stored encoded (seqFastqEncode of the per-read preparation, smalt.c:1106-1127) */
    if ((errcode = smbShimSeqFastqLoad(w->reads[n], name, nl_, l[1], ll[1], qnam, ql_, l[3], ll[3], macop->codecp)))
      return errcode;
    n++;
  }
  *nreads = n;
  return ERRCODE_SUCCESS;
}

static int fm_parse_block(FmWorker *w, const char *data, size_t start, size_t end, size_t *nreads)
{
  FastMap *fm = w->fm;
  int errcode = ERRCODE_SUCCESS;
  size_t n = 0, nrec_expect = 0, off;
  SeqIO *sio;
  *nreads = 0;
  if (end <= start) return ERRCODE_SUCCESS;
  if (!fm->is_fasta && !fm->refparse) {
    /* the fast parser accepts whole 4-line records only, so a block it accepts needs no other check */
    errcode = fm_parse_block_fast(w, data, start, end, &n);
    if (errcode != 1) {
      *nreads = n;
      return errcode;
    }
    errcode = ERRCODE_SUCCESS;
    n = 0;
  }
  if (!fm->is_fasta && (errcode = fm_check_fastq(data, start, end, &nrec_expect))) {
    fprintf(stderr, "smalt_b200: the read file is not plain 4-line FASTQ near byte %zu; "
	    "rerun with SMALT_B200_REFIO=1 (the reference's own reader)\n", start);
    return errcode;
  }
  if (ftruncate(w->memfd, 0)) return ERRCODE_FILEIO;
  for (off = start; off < end;) {
    const ssize_t k = pwrite(w->memfd, data + off, end - off, (off_t) (off - start));
    if (k <= 0) return ERRCODE_FILEIO;
    off += (size_t) k;
  }
  sio = seqIOopen(&errcode, w->fdpath, SEQIO_READ, 0);
  if (!sio) return errcode ? errcode : ERRCODE_NOFILE;
  while (!seqIOstatus(sio)) { /* loadIOBuffArg / infmtRead (smalt.c:795-830, infmt.c:197-240) */
    if ((errcode = fm_reads_reserve(w, n))) break;
    seqFastqBlank(w->reads[n]);
    if ((errcode = seqFastqRead(w->reads[n], sio))) break;
    n++;
  }
  if (errcode == ERRCODE_EOF) errcode = ERRCODE_SUCCESS;
  if (!errcode && seqIOstatus(sio) != ERRCODE_EOF) errcode = seqIOstatus(sio);
  seqIOclose(sio);
  if (!errcode && !fm->is_fasta && n != nrec_expect) errcode = ERRCODE_FASTA;
  *nreads = n;
  return errcode;
}

static int fm_map_parsed(FmWorker *w, size_t c, size_t n);
static int fm_map_block(FmWorker *w, size_t c)
{
  FastMap *fm = w->fm;
  const SmaltMapConst *macop = fm->macop;
  const size_t start = fm_record_start(fm, c * fm->chunk_bytes);
  const size_t end = (c + 1 == fm->nchunks) ? fm->len : fm_record_start(fm, (c + 1) * fm->chunk_bytes);
  size_t n = 0, i, pos, buflen = 0;
  char *buf = NULL;
  int errcode;
  FmEmit em;
  FILE *fp;
  struct timespec t0, t1;

  clock_gettime(CLOCK_MONOTONIC, &t0);
  errcode = fm_parse_block(w, fm->data, start, end, &n);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_mutex_lock(&g_stats_lock);
  g_fm_parse_s += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  pthread_mutex_unlock(&g_stats_lock);
  if (errcode || !n) { fm_publish(fm, c, NULL, 0, errcode); return errcode; }

  for (i = 0; i < n; i++) { /* per-read preparation of processMapArgs (smalt.c:1106-1127) */
    uint32_t covermin_tuple;
    if ((errcode = seqFastqEncode(w->reads[i], macop->codecp))) break;
    if (macop->tupcovmin < 1.01) {
      uint32_t readlen;
      seqFastqGetConstSequence(w->reads[i], &readlen, NULL);
      covermin_tuple = (uint32_t) (macop->tupcovmin * readlen);
      if (covermin_tuple > readlen) covermin_tuple = readlen;
    } else {
      covermin_tuple = (uint32_t) macop->tupcovmin;
    }
    w->mincov[i] = covermin_tuple;
  }
  if (errcode) { fm_publish(fm, c, NULL, 0, errcode); return errcode; }
  return fm_map_parsed(w, c, n);
}

/* maps the n parsed reads of w->reads (block c) on the worker's own stream and publishes the block */
static int fm_map_parsed(FmWorker *w, size_t c, size_t n)
{
  FastMap *fm = w->fm;
  const SmaltMapConst *macop = fm->macop;
  size_t pos, buflen = 0;
  char *buf = NULL;
  int errcode = ERRCODE_SUCCESS;
  FmEmit em;
  FILE *fp;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  /* Formatted records are captured from the reference's fprintf calls (fastprintf.h) for the
   * line-oriented formats; explicit alignment output (-a) also writes by other means and
   * goes through a memory stream. */
  {
    const int capture = !(macop->oumodflg & REPORTMODIF_ALIOUT) &&
      (macop->outform == REPORTFMT_SAM || macop->outform == REPORTFMT_CIGAR || macop->outform == REPORTFMT_SSAHA);
    if (capture) {
      if (!w->keyfp && !(w->keyfp = open_memstream(&w->keybuf, &w->keylen))) {
	fm_publish(fm, c, NULL, 0, ERRCODE_NOMEM);
	return ERRCODE_NOMEM;
      }
      fp = NULL;
      smbFastCaptureBegin(w->keyfp);
      smbShimReportWriterSetStream(w->writer, w->keyfp);
    } else {
      if (!(fp = open_memstream(&buf, &buflen))) { fm_publish(fm, c, NULL, 0, ERRCODE_NOMEM); return ERRCODE_NOMEM; }
      smbShimReportWriterSetStream(w->writer, fp);
    }
  }
  em.w = w; em.fp = fp; em.n = n;
  /* rmapSingleWave takes at most INT_MAX reads; blocks are far smaller */
  for (pos = 0; pos < n && !errcode; pos += 32000) {
    const int nb = (int) ((n - pos < 32000) ? n - pos : 32000);
    SeqFastq **save = w->reads;
    w->reads += pos; /* fm_emit indexes relative to the sub-block */
    errcode = rmapSingleWave(w->errmsgp, fm->maps[w->id].rmp, w->wave, nb, w->reads, w->mincov + pos,
			     macop->nhitmax_tuple, (int) macop->min_swatscor, macop->swatscordiff, macop->minbasq,
			     SMALT_TARGET_DEPTH, SMALT_MAX_DEPTH, (RMAPFLG_t) (macop->rmapflg & ~RMAPFLG_ALLPAIR),
			     macop->scormtxp, macop->rfp, macop->htp, macop->ssp, macop->codecp, fm_emit, &em);
    w->reads = save;
  }
  smbShimReportWriterSetStream(w->writer, NULL);
  if (fp) {
    if (fclose(fp) && !errcode) errcode = ERRCODE_FILEIO;
  } else {
    if (smbFastCaptureEnd(&buf, &buflen) && !errcode) errcode = ERRCODE_NOMEM;
    /* nothing may have reached the key stream itself (it would be out of order) */
    if (fflush(w->keyfp) || w->keylen != 0) { if (!errcode) errcode = ERRCODE_ASSERT; }
  }
  pthread_mutex_lock(&g_stats_lock);
  fm->n_reads += n;
  pthread_mutex_unlock(&g_stats_lock);
  fm_publish(fm, c, buf, buflen, errcode);
  if (getenv("SMALT_B200_TIMING")) {
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fprintf(stderr, "smalt_b200 timing: worker %d block %zu (%zu reads) %.3f s, done at %.3f s\n", w->id, c, n,
	    (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), t1.tv_sec + 1e-9 * t1.tv_nsec - g_t0);
  }
  return errcode;
}

/* ---- pipelined single-end path ------------------------------------------------------------ */
/* Workers never wait for the device: a worker parses a block and delivers its reads to the open device batch
 * (rmap_wave.c, WaveCombiner), then takes the next piece of work - preferably the results of a block whose
 * batch has come back, else the next block to parse.  Two device threads run the batches.  Blocks are
 * published in input order as before. */
static void fm_block_swap_in(FmWorker *w, FmBlock *b, SeqFastq ***sr, uint32_t **sm, size_t *sa)
{
  *sr = w->reads; *sm = w->mincov; *sa = w->n_alloc;
  w->reads = b->reads; w->mincov = b->mincov; w->n_alloc = b->n_alloc;
}
static void fm_block_swap_out(FmWorker *w, FmBlock *b, SeqFastq **sr, uint32_t *sm, size_t sa)
{
  b->reads = w->reads; b->mincov = w->mincov; b->n_alloc = w->n_alloc;
  w->reads = sr; w->mincov = sm; w->n_alloc = sa;
}

static void fm_block_release(FastMap *fm, FmBlock *b)
{
  pthread_mutex_lock(&fm->lock);
  b->inflight = 0;
  b->tk.slot = -1;
  b->next = fm->free_blk;
  fm->free_blk = b;
  fm->outstanding--;
  pthread_cond_broadcast(&fm->cond);
  pthread_mutex_unlock(&fm->lock);
}

/* results + formatting of a block whose batch is back */
static int fm_block_results(FmWorker *w, FmBlock *b)
{
  FastMap *fm = w->fm;
  const SmaltMapConst *macop = fm->macop;
  SeqFastq **sr; uint32_t *sm; size_t sa, buflen = 0;
  char *buf = NULL;
  FILE *fp = NULL;
  FmEmit em;
  int errcode = ERRCODE_SUCCESS;
  const int capture = !(macop->oumodflg & REPORTMODIF_ALIOUT) &&
    (macop->outform == REPORTFMT_SAM || macop->outform == REPORTFMT_CIGAR || macop->outform == REPORTFMT_SSAHA);
  fm_block_swap_in(w, b, &sr, &sm, &sa);
  if (capture) {
    if (!w->keyfp && !(w->keyfp = open_memstream(&w->keybuf, &w->keylen))) errcode = ERRCODE_NOMEM;
    else { smbFastCaptureBegin(w->keyfp); smbShimReportWriterSetStream(w->writer, w->keyfp); }
  } else {
    if (!(fp = open_memstream(&buf, &buflen))) errcode = ERRCODE_NOMEM;
    else smbShimReportWriterSetStream(w->writer, fp);
  }
  if (!errcode) {
    em.w = w; em.fp = fp; em.n = b->n;
    errcode = waveCombinerResults(w->errmsgp, fm->maps[w->id].rmp, w->wave, g_fm_comb, &b->tk, w->reads, w->mincov,
				  (int) macop->min_swatscor, SMALT_MAX_DEPTH, (RMAPFLG_t) (macop->rmapflg & ~RMAPFLG_ALLPAIR),
				  macop->scormtxp, macop->rfp, macop->ssp, macop->codecp, fm_emit, &em);
    smbShimReportWriterSetStream(w->writer, NULL);
    if (fp) {
      if (fclose(fp) && !errcode) errcode = ERRCODE_FILEIO;
    } else {
      if (smbFastCaptureEnd(&buf, &buflen) && !errcode) errcode = ERRCODE_NOMEM;
      if (fflush(w->keyfp) || w->keylen != 0) { if (!errcode) errcode = ERRCODE_ASSERT; }
    }
  }
  pthread_mutex_lock(&g_stats_lock);
  fm->n_reads += b->n;
  pthread_mutex_unlock(&g_stats_lock);
  fm_block_swap_out(w, b, sr, sm, sa);
  fm_publish(fm, b->chunk, buf, buflen, errcode);
  fm_block_release(fm, b);
  return errcode;
}

/* parse block b->chunk and deliver it to the device batches (or map it directly when the combiner does not
 * take it); while no batch slot is free the worker finishes blocks that are back */
static int fm_block_parse_deliver(FmWorker *w, FmBlock *b)
{
  FastMap *fm = w->fm;
  const SmaltMapConst *macop = fm->macop;
  const size_t c = b->chunk;
  const size_t start = fm_record_start(fm, c * fm->chunk_bytes);
  const size_t end = (c + 1 == fm->nchunks) ? fm->len : fm_record_start(fm, (c + 1) * fm->chunk_bytes);
  SeqFastq **sr; uint32_t *sm; size_t sa, n = 0, i;
  int errcode, takes;
  struct timespec t0, t1;
  fm_block_swap_in(w, b, &sr, &sm, &sa);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  errcode = fm_parse_block(w, fm->data, start, end, &n);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_mutex_lock(&g_stats_lock);
  g_fm_parse_s += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  pthread_mutex_unlock(&g_stats_lock);
  for (i = 0; i < n && !errcode; i++) { /* per-read preparation of processMapArgs (smalt.c:1106-1127) */
    uint32_t covermin_tuple;
    if ((errcode = seqFastqEncode(w->reads[i], macop->codecp))) break;
    if (macop->tupcovmin < 1.01) {
      uint32_t readlen;
      seqFastqGetConstSequence(w->reads[i], &readlen, NULL);
      covermin_tuple = (uint32_t) (macop->tupcovmin * readlen);
      if (covermin_tuple > readlen) covermin_tuple = readlen;
    } else {
      covermin_tuple = (uint32_t) macop->tupcovmin;
    }
    w->mincov[i] = covermin_tuple;
  }
  b->n = n;
  takes = !errcode && n > 0 && n <= 32000 &&
    waveCombinerTakes(g_fm_comb, (int) n, w->reads, (RMAPFLG_t) (macop->rmapflg & ~RMAPFLG_ALLPAIR), macop->htp);
  fm_block_swap_out(w, b, sr, sm, sa);
  if (errcode || !n) {
    fm_publish(fm, c, NULL, 0, errcode);
    fm_block_release(fm, b);
    return errcode;
  }
  if (!takes) {   /* long reads, other modes: the block is mapped by this worker on its own stream */
    FmBlock *keep = b;
    SeqFastq **r2; uint32_t *m2; size_t a2;
    fm_block_swap_in(w, keep, &r2, &m2, &a2);
    errcode = fm_map_parsed(w, c, n);
    fm_block_swap_out(w, keep, r2, m2, a2);
    fm_block_release(fm, keep);
    return errcode;
  }
  b->tk.slot = -1;
  b->inflight = 1;
  for (;;) {
    const int rc = waveCombinerDeliver(g_fm_comb, (int) n, b->reads, b->mincov, (int) macop->min_swatscor, macop->scormtxp,
				       macop->htp, &b->tk);
    FmBlock *other = NULL;
    if (rc == 0) break;
    if (rc != 1) { fm_publish(fm, c, NULL, 0, rc); fm_block_release(fm, b); return rc; }
    /* every batch slot is busy: finish a block that is back, or wait for one */
    pthread_mutex_lock(&fm->lock);
    if (fm->ready) { other = fm->ready; fm->ready = other->next; }
    else {
      struct timespec ts;
      clock_gettime(CLOCK_REALTIME, &ts);
      ts.tv_nsec += 500000;
      if (ts.tv_nsec >= 1000000000) { ts.tv_sec++; ts.tv_nsec -= 1000000000; }
      pthread_cond_timedwait(&fm->cond, &fm->lock, &ts);
    }
    pthread_mutex_unlock(&fm->lock);
    if (other && (errcode = fm_block_results(w, other))) return errcode;
  }
  return ERRCODE_SUCCESS;
}

static void *fm_worker_pipe(void *arg)
{
  FmWorker *w = (FmWorker *) arg;
  FastMap *fm = w->fm;
  int rc = (pregrow_arena(), fm_worker_setup(w, fm, w->id));
  if (rc) {
    pthread_mutex_lock(&fm->lock);
    if (!fm->errcode) fm->errcode = rc;
    pthread_cond_broadcast(&fm->cond);
    pthread_mutex_unlock(&fm->lock);
    return NULL;
  }
  for (;;) {
    FmBlock *b = NULL;
    int mode = 0;   /* 1 results, 2 parse, 3 exit */
    pthread_mutex_lock(&fm->lock);
    for (;;) {
      if (fm->ready) { b = fm->ready; fm->ready = b->next; mode = 1; break; }
      if (!fm->errcode && fm->next_chunk < fm->nchunks && fm->free_blk) {
	b = fm->free_blk; fm->free_blk = b->next;
	b->chunk = fm->next_chunk++;
	fm->outstanding++;
	mode = 2;
	break;
      }
      if ((fm->errcode || fm->next_chunk >= fm->nchunks) && fm->outstanding == 0) { mode = 3; break; }
      pthread_cond_wait(&fm->cond, &fm->lock);
    }
    pthread_mutex_unlock(&fm->lock);
    if (mode == 3) break;
    rc = (mode == 1) ? fm_block_results(w, b) : fm_block_parse_deliver(w, b);
    if (rc) {
      pthread_mutex_lock(&fm->lock);
      if (!fm->errcode) fm->errcode = rc;
      pthread_cond_broadcast(&fm->cond);
      pthread_mutex_unlock(&fm->lock);
    }
  }
  pthread_mutex_lock(&fm->lock);
  pthread_cond_broadcast(&fm->cond);
  pthread_mutex_unlock(&fm->lock);
  return NULL;
}

/* a device thread: runs closed batches, hands their blocks to the workers */
static void *fm_device_main(void *arg)
{
  FastMap *fm = (FastMap *) arg;
  const SmaltMapConst *macop = fm->macop;
  ErrMsg *errmsgp;
  ERRMSG_CREATE(errmsgp);
#if defined(__linux__)
  /* this thread polls the device with sleeps of 50 us (smb_ctx_set_spin): with the default timer slack of 50 us such
   * a sleep lasts 100 us and more, which costs a quarter of the throughput on few cores */
  prctl(PR_SET_TIMERSLACK, 1000UL, 0UL, 0UL, 0UL);
#endif
  for (;;) {
    int errcode = 0, k;
    const int s = waveCombinerRunNext(errmsgp, g_fm_comb, macop->nhitmax_tuple, macop->swatscordiff, macop->minbasq,
				      SMALT_TARGET_DEPTH, SMALT_MAX_DEPTH, (RMAPFLG_t) (macop->rmapflg & ~RMAPFLG_ALLPAIR),
				      macop->scormtxp, NULL, macop->ssp, &errcode);
    if (s < 0) break;
    pthread_mutex_lock(&fm->lock);
    if (errcode && !fm->errcode) fm->errcode = errcode;
    for (k = 0; k < fm->nblocks; k++) {
      FmBlock *b = fm->blocks + k;
      if (b->inflight == 1 && b->tk.slot == s) {
	FmBlock **pp = &fm->ready;
	b->inflight = 2;
	while (*pp && (*pp)->chunk < b->chunk) pp = &(*pp)->next;   /* earlier blocks first: the output is flushed in order */
	b->next = *pp;
	*pp = b;
      }
    }
    pthread_cond_broadcast(&fm->cond);
    pthread_mutex_unlock(&fm->lock);
  }
  ERRMSG_END(errmsgp);
  return NULL;
}

/* ---- paired input ---------------------------------------------------------------------- */
/* Two files hold the mates of pair i as their i-th records, so both are cut at the same RECORD
 * numbers: the newlines of fixed-size segments are counted in parallel once, the byte offset of
 * a line is then found by a search over the segment sums and a scan inside one segment. */
typedef struct FmLines_ {
  const char *d;
  size_t len, seg_bytes, nseg;
  uint64_t *cum;        /* cum[s] = newlines in [0, s*seg_bytes), nseg+1 entries */
  uint64_t nlines;      /* lines of the text (a last line without newline counts) */
} FmLines;

typedef struct { FmLines *L; size_t s0, s1; } FmCountJob;

static void *fm_count_main(void *arg)
{
  FmCountJob *j = (FmCountJob *) arg;
  FmLines *L = j->L;
  size_t s;
  for (s = j->s0; s < j->s1; s++) {
    const size_t a = s * L->seg_bytes, b = (a + L->seg_bytes < L->len) ? a + L->seg_bytes : L->len;
    const char *p = L->d + a, *e = L->d + b;
    uint64_t c = 0;
    while (p < e && (p = (const char *) memchr(p, '\n', (size_t) (e - p)))) { c++; p++; }
    L->cum[s + 1] = c;
  }
  return NULL;
}

static FmLines *fm_lines_build(const char *d, size_t len, int nthreads)
{
  FmLines *L = (FmLines *) calloc(1, sizeof(*L));
  pthread_t *tid;
  FmCountJob *jobs;
  size_t s;
  int t;
  if (!L) return NULL;
  L->d = d; L->len = len;
  L->seg_bytes = (size_t) 1 << 18;
  L->nseg = (len + L->seg_bytes - 1) / L->seg_bytes;
  if (nthreads < 1) nthreads = 1;
  if ((size_t) nthreads > L->nseg) nthreads = L->nseg ? (int) L->nseg : 1;
  L->cum = (uint64_t *) calloc(L->nseg + 2, sizeof(uint64_t));
  tid = (pthread_t *) calloc((size_t) nthreads, sizeof(pthread_t));
  jobs = (FmCountJob *) calloc((size_t) nthreads, sizeof(FmCountJob));
  if (!L->cum || !tid || !jobs) { free(L->cum); free(L); free(tid); free(jobs); return NULL; }
  for (t = 0; t < nthreads; t++) {
    jobs[t].L = L;
    jobs[t].s0 = L->nseg * (size_t) t / (size_t) nthreads;
    jobs[t].s1 = L->nseg * (size_t) (t + 1) / (size_t) nthreads;
    if (nthreads == 1) fm_count_main(jobs + t);
    else pthread_create(tid + t, NULL, fm_count_main, jobs + t);
  }
  if (nthreads > 1) for (t = 0; t < nthreads; t++) pthread_join(tid[t], NULL);
  for (s = 0; s < L->nseg; s++) L->cum[s + 1] += L->cum[s];
  L->nlines = L->cum[L->nseg] + ((len > 0 && d[len - 1] != '\n') ? 1 : 0);
  free(tid); free(jobs);
  return L;
}

static void fm_lines_free(FmLines *L) { if (L) { free(L->cum); free(L); } }

/* byte offset of the start of line `line` (0-based); len for line >= nlines */
static size_t fm_line_offset(const FmLines *L, uint64_t line)
{
  size_t lo = 0, hi = L->nseg, a, b;
  const char *p, *e;
  uint64_t need;
  if (line == 0) return 0;
  if (line > L->cum[L->nseg]) return L->len;
  /* the line starts behind the line-th newline: segment s with cum[s] < line <= cum[s+1] */
  while (lo + 1 < hi) {
    const size_t mid = (lo + hi) / 2;
    if (L->cum[mid] < line) lo = mid; else hi = mid;
  }
  a = lo * L->seg_bytes;
  b = (a + L->seg_bytes < L->len) ? a + L->seg_bytes : L->len;
  need = line - L->cum[lo];
  p = L->d + a; e = L->d + b;
  while (need && p < e && (p = (const char *) memchr(p, '\n', (size_t) (e - p)))) { p++; need--; }
  return (need || !p) ? L->len : (size_t) (p - L->d);
}

typedef struct { FmWorker *w; } FmPairEmit;

static int fm_emit_pair(void *user, int i, const ResultSet *rsltp, const ResultSet *rslt_matep,
			const ResultPairs *pairp, RSLTPAIRFLG_t pairflg)
{ /* tail of processMapArgs (smalt.c:1168-1176) + outputIOBuffArg (smalt.c:852-866) */
  FmWorker *w = ((FmPairEmit *) user)->w;
  const SmaltMapConst *macop = w->fm->macop;
  int errcode;
  reportBlank(w->rep);
  if ((errcode = resultSetAddPairToReport(w->rep, macop->ihp, pairp, pairflg, macop->rsltouflg, rsltp, rslt_matep)))
    return errcode;
  if ((macop->menuflg & MENUFLAG_RELSCOR) &&
      (macop->outform == REPORTFMT_SAM || macop->outform == REPORTFMT_BAM))
    reportFixMultiplePrimary(w->rep);
  return reportWrite(w->writer, w->pairs[2 * i], w->pairs[2 * i + 1], macop->ssp, macop->codecp, w->rep);
}

static int fm_map_block_pairs(FmWorker *w, size_t c)
{
  FastMap *fm = w->fm;
  const SmaltMapConst *macop = fm->macop;
  const uint64_t r0 = (uint64_t) c * fm->block_recs;
  const uint64_t r1 = (r0 + fm->block_recs < fm->nrecs) ? r0 + fm->block_recs : fm->nrecs;
  const size_t sA = fm_line_offset(fm->linesA, 4 * r0), eA = fm_line_offset(fm->linesA, 4 * r1);
  const size_t sB = fm_line_offset(fm->linesB, 4 * r0), eB = fm_line_offset(fm->linesB, 4 * r1);
  size_t n = 0, nB = 0, i, buflen = 0;
  char *buf = NULL;
  int errcode;
  FILE *fp = NULL;
  FmPairEmit em;
  struct timespec t0, t1;

  clock_gettime(CLOCK_MONOTONIC, &t0);
  errcode = fm_parse_block(w, fm->data, sA, eA, &n);
  if (!errcode) { /* the mates with the same parser into their own array */
    SeqFastq **r = w->reads; uint32_t *m = w->mincov; size_t a = w->n_alloc;
    w->reads = w->mates; w->mincov = w->mincov_m; w->n_alloc = w->m_alloc;
    errcode = fm_parse_block(w, fm->dataB, sB, eB, &nB);
    w->mates = w->reads; w->mincov_m = w->mincov; w->m_alloc = w->n_alloc;
    w->reads = r; w->mincov = m; w->n_alloc = a;
  }
  if (!errcode && (n != nB || n != (size_t) (r1 - r0))) {
    fprintf(stderr, "smalt_b200: the two read files do not hold the same records near pair %llu; "
	    "rerun with SMALT_B200_REFIO=1 (the reference's own reader)\n", (unsigned long long) r0);
    errcode = ERRCODE_FASTA;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_mutex_lock(&g_stats_lock);
  g_fm_parse_s += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  pthread_mutex_unlock(&g_stats_lock);
  if (errcode || !n) { fm_publish(fm, c, NULL, 0, errcode); return errcode; }
  if (n > w->pairs_alloc) {
    free(w->pairs);
    w->pairs_alloc = n + 256;
    if (!(w->pairs = (SeqFastq **) malloc(2 * w->pairs_alloc * sizeof(SeqFastq *)))) {
      w->pairs_alloc = 0;
      fm_publish(fm, c, NULL, 0, ERRCODE_NOMEM);
      return ERRCODE_NOMEM;
    }
  }
  for (i = 0; i < n; i++) { w->pairs[2 * i] = w->reads[i]; w->pairs[2 * i + 1] = w->mates[i]; }
  {
    const int capture = !(macop->oumodflg & REPORTMODIF_ALIOUT) &&
      (macop->outform == REPORTFMT_SAM || macop->outform == REPORTFMT_CIGAR || macop->outform == REPORTFMT_SSAHA);
    if (capture) {
      if (!w->keyfp && !(w->keyfp = open_memstream(&w->keybuf, &w->keylen))) {
	fm_publish(fm, c, NULL, 0, ERRCODE_NOMEM);
	return ERRCODE_NOMEM;
      }
      smbFastCaptureBegin(w->keyfp);
      smbShimReportWriterSetStream(w->writer, w->keyfp);
    } else {
      if (!(fp = open_memstream(&buf, &buflen))) { fm_publish(fm, c, NULL, 0, ERRCODE_NOMEM); return ERRCODE_NOMEM; }
      smbShimReportWriterSetStream(w->writer, fp);
    }
  }
  em.w = w;
  errcode = pair_core(w->errmsgp, fm->maps[w->id].rmp, w->wave, &w->pw, macop, (short) w->id, (int) n, w->pairs,
		      fm_emit_pair, &em, 0);
  smbShimReportWriterSetStream(w->writer, NULL);
  if (fp) {
    if (fclose(fp) && !errcode) errcode = ERRCODE_FILEIO;
  } else {
    if (smbFastCaptureEnd(&buf, &buflen) && !errcode) errcode = ERRCODE_NOMEM;
    if (fflush(w->keyfp) || w->keylen != 0) { if (!errcode) errcode = ERRCODE_ASSERT; }
  }
  pthread_mutex_lock(&g_stats_lock);
  fm->n_reads += 2 * n;
  pthread_mutex_unlock(&g_stats_lock);
  fm_publish(fm, c, buf, buflen, errcode);
  if (getenv("SMALT_B200_TIMING")) {
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fprintf(stderr, "smalt_b200 timing: worker %d block %zu (%zu pairs) %.3f s, done at %.3f s\n", w->id, c, n,
	    (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), t1.tv_sec + 1e-9 * t1.tv_nsec - g_t0);
  }
  return errcode;
}

static void *fm_worker_main(void *arg)
{
  FmWorker *w = (FmWorker *) arg;
  FastMap *fm = w->fm;
  int rc = (pregrow_arena(), fm_worker_setup(w, fm, w->id));
  if (rc) {
    pthread_mutex_lock(&fm->lock);
    if (!fm->errcode) fm->errcode = rc;
    pthread_mutex_unlock(&fm->lock);
    return NULL;
  }
  for (;;) {
    size_t c;
    int stop;
    pthread_mutex_lock(&fm->lock);
    c = fm->next_chunk;
    stop = fm->errcode != 0 || c >= fm->nchunks;
    if (!stop) fm->next_chunk++;
    pthread_mutex_unlock(&fm->lock);
    if (stop) break;
    if (fm->dataB ? fm_map_block_pairs(w, c) : fm_map_block(w, c)) break;
  }
  /* chunks claimed by nobody after an error must still be marked done for the flusher */
  return NULL;
}

static void fm_collect_stats(FmWorker *w)
{
  double ms[3], wall[11];
  uint64_t counts[5];
  int i;
  if (!w->wave) return;
  rmapWaveGetStats(w->wave, ms, counts);
  rmapWaveGetWall(w->wave, wall);
  {
    double cpu[8];
    rmapWaveGetCpu(w->wave, cpu);
    pthread_mutex_lock(&g_stats_lock);
    for (i = 0; i < 8; i++) { g_cpu[i] += cpu[i] - w->cpu_prev[i]; w->cpu_prev[i] = cpu[i]; }
    pthread_mutex_unlock(&g_stats_lock);
  }
  pthread_mutex_lock(&g_stats_lock);
  for (i = 0; i < 11; i++) { g_wall[i] += wall[i] - w->wall_prev[i]; w->wall_prev[i] = wall[i]; }
  for (i = 0; i < 3; i++) { g_ms[i] += ms[i] - w->ms_prev[i]; w->ms_prev[i] = ms[i]; }
  { const double c = rmapWaveGetCandMs(w->wave); g_ms_cand += c - w->cand_prev; w->cand_prev = c; }
  for (i = 0; i < 5; i++) { g_counts[i] += counts[i] - w->counts_prev[i]; w->counts_prev[i] = counts[i]; }
  pthread_mutex_unlock(&g_stats_lock);
}

/* Maps the reads in data[0..len) and hands the formatted output to sinkf in input order. */
static int fastmap_run(const SmaltMapConst *macop, SmaltMapArgs *maps, int nworkers, const ReportWriter *proto,
		       const char *data, size_t len, const char *dataB, size_t lenB,
		       FASTMAP_SINKF *sinkf, void *sink_user, uint64_t *n_reads)
{
  FastMap fm;
  pthread_t *tid;
  int i, errcode = ERRCODE_SUCCESS;
  size_t p = 0, block = 8192, rec_bytes;
  const char *e = getenv("SMALT_B200_BLOCK");
  {   /* pipelined single-end path (device batches are made of several blocks): smaller host blocks */
    const char *ce = getenv("SMALT_B200_COMBINE");
    if (!dataB && nworkers > 1 && (!ce || atoi(ce) != 0)) block = nworkers >= 8 ? 2048 : 4096;
  }

  memset(&fm, 0, sizeof(fm));
  while (p < len && isspace((unsigned char) data[p])) p++;
  if (p >= len) { if (n_reads) *n_reads = 0; return ERRCODE_SUCCESS; }
  fm.macop = macop; fm.maps = maps; fm.proto = proto;
  /* Single-end SAM: the device can also emit CIGAR and NM of every alignment (csrc/cigar.cu).  The stage follows
   * the bottleneck: it takes ~5 % of the workers' CPU time away and lengthens every device batch by one small
   * kernel and one copy - measured on one box (tools/core_sweep.py, 1 M C2 reads per call): 4 cores 5.5 M reads/s
   * with it, 5.2 M without; 16 cores 11.7 M with it, 12.2 M without (there the device batches are the critical path).
   * Default: on up to 8 workers (host-bound), off above; SMALT_B200_DEVCIGAR=1 / 0 forces it. */
  {
    const char *dc = getenv("SMALT_B200_DEVCIGAR");
    const int on = dc ? atoi(dc) != 0 : (nworkers <= 8 && !getenv("SMALT_B200_HOSTCIGAR"));
    rmapWaveSetCigarMode((dataB || !on) ? 0 : smbShimReportCigarFlags(proto));
  }
  fm.data = data + p; fm.len = len - p;
  fm.is_fasta = data[p] == '>';
  fm.refparse = getenv("SMALT_B200_REFPARSE") != NULL;
  if (dataB) g_pregrow_mb = 96;
  fm.nworkers = nworkers;
  fm.sinkf = sinkf; fm.sink_user = sink_user;
  if (dataB) { /* paired: blocks of whole records, the same record numbers in both files */
    size_t b = 4096;  /* pairs per block: the four passes of a block are four rounds of launches, and larger rounds use the device better (C3, 16 workers: 1024 pairs 1.20, 2048 1.65, 4096 1.97, 8192 1.99 M reads/s); two result sets of >= 24 KB (six 4 KB arrays, array.c:52-76) stay allocated per pair = 200 MB per worker */
    if (p || fm.is_fasta || data[0] != '@' || !lenB || dataB[0] != '@') return ERRCODE_ARGINVAL;
    fm.dataB = dataB; fm.lenB = lenB;
    fm.linesA = fm_lines_build(fm.data, fm.len, nworkers);
    fm.linesB = fm_lines_build(dataB, lenB, nworkers);
    if (!fm.linesA || !fm.linesB) { fm_lines_free(fm.linesA); fm_lines_free(fm.linesB); return ERRCODE_NOMEM; }
    if (fm.linesA->nlines != fm.linesB->nlines || (fm.linesA->nlines & 3)) {
      fm_lines_free(fm.linesA); fm_lines_free(fm.linesB);
      return ERRCODE_ARGINVAL; /* not two plain 4-line FASTQ files of equal length: reference reader */
    }
    fm.nrecs = (size_t) (fm.linesA->nlines >> 2);
    if (e && atol(e) > 0) b = (size_t) atol(e);
    else {
      const size_t bb = fm.nrecs / ((size_t) nworkers * 6) + 1;
      if (bb < b) b = bb;
      if (b < 128) b = 128;
    }
    if (b > 16000) b = 16000;
    fm.block_recs = block = b;
    fm.nchunks = (fm.nrecs + b - 1) / b;
    if (!fm.nchunks) { fm_lines_free(fm.linesA); fm_lines_free(fm.linesB); if (n_reads) *n_reads = 0; return ERRCODE_SUCCESS; }
  } else
  /* block size: reads per block -> bytes per block from the first records */
  {
    size_t q = 0, lines = 0, want = fm.is_fasta ? 128 : 256;
    while (q < fm.len && lines < want) {
      const char *nl = (const char *) memchr(fm.data + q, '\n', fm.len - q);
      if (!nl) { q = fm.len; lines++; break; }
      q = (size_t) (nl - fm.data) + 1;
      lines++;
    }
    rec_bytes = q / ((lines + (fm.is_fasta ? 1 : 3)) / (fm.is_fasta ? 2 : 4) + (lines < 4));
    if (rec_bytes < 16) rec_bytes = 16;
  }
  if (dataB) ;
  else if (e && atol(e) > 0) block = (size_t) atol(e);
  else { /* at least ~6 blocks per worker for load balance, at least 512 reads per block */
    const size_t est_reads = fm.len / rec_bytes + 1;
    size_t b = est_reads / ((size_t) nworkers * 6) + 1;
    if (b < 512) b = 512;
    if (b < block) block = b;
  }
  if (!dataB) {
    if (block > 32000) block = 32000;
    fm.chunk_bytes = block * rec_bytes;
    fm.nchunks = (fm.len + fm.chunk_bytes - 1) / fm.chunk_bytes;
  }
  if (!(fm.out = (FmBlockOut *) calloc(fm.nchunks, sizeof(FmBlockOut)))) return ERRCODE_NOMEM;
  pthread_mutex_init(&fm.lock, NULL);

  if (g_fm_nworkers < nworkers) {
    FmWorker *hp = (FmWorker *) realloc(g_fm_workers, (size_t) nworkers * sizeof(FmWorker));
    if (!hp) { free(fm.out); return ERRCODE_NOMEM; }
    memset(hp + g_fm_nworkers, 0, (size_t) (nworkers - g_fm_nworkers) * sizeof(FmWorker));
    g_fm_workers = hp;
    g_fm_nworkers = nworkers;
  }
  {
    struct timespec ts;
    double t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &ts); t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
    /* the process-wide part (CUDA context, index + reference upload) once, the per-worker part
     * (stream, report writer, memfd) by the workers themselves */
    if (smbShimInit(macop->htp, macop->ssp, macop->codecp, macop->scormtxp)) errcode = ERRCODE_FAILURE;
    for (i = 0; i < nworkers; i++) { g_fm_workers[i].fm = &fm; g_fm_workers[i].id = i; }
    clock_gettime(CLOCK_MONOTONIC, &ts); t1 = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if (getenv("SMALT_B200_TIMING"))
      fprintf(stderr, "smalt_b200 timing: fastmap set-up of %d workers %.3f s (at %.3f s); %zu blocks of ~%zu reads\n",
	      nworkers, t1 - t0, t1 - g_t0, fm.nchunks, block);
  }
  if (!errcode && !dataB && nworkers > 1 && !g_fm_comb) {
    const char *ce = getenv("SMALT_B200_COMBINE"), *be = getenv("SMALT_B200_BATCH");
    if (!ce || atoi(ce) != 0) {
      /* the device threads poll with 50 us sleeps instead of blocking in the driver (SMALT_B200_SPIN: 0 block, 1 spin,
       * 2 poll with 20 us sleeps, n > 2 with n us): as fast as spinning on 16 cores, 15-25 % faster on 4 and 8, where
       * spinning threads take the workers' cores; blocking costs a third of the throughput (tools/core_sweep.py) */
      const char *se = getenv("SMALT_B200_SPIN");
      g_fm_comb = waveCombinerCreate(macop->htp, macop->ssp, macop->codecp, macop->scormtxp, 8,
				     /* reads per device batch: smaller batches shorten the fill and drain of a call's pipeline
				      * when the device is the limit (16 cores: 12288 -> 11.7, 16384 -> 11.2-11.5, 24576 -> 10.3 M
				      * reads/s per 1 M-read call), larger ones save launches when the host is (4 cores: 16384 ->
				      * 5.3, 8192 -> 4.9 M reads/s); tools/core_sweep.py, tools/batch_sweep.sh */
				     (be && atoi(be) > 0) ? atoi(be) : (nworkers >= 8 ? 12288 : 16384),
				     /* poll interval: 20 us where cores are plenty (16 cores: 12.3 vs 12.0 M reads/s with 50 us) */
				     se ? atoi(se) : (nworkers >= 8 ? 2 : 50));
      if (!g_fm_comb) errcode = ERRCODE_FAILURE;
    }
  }
  if (!errcode && g_fm_comb && !dataB && nworkers > 1) {
    /* pipelined path: workers parse / deliver / finish blocks, two device threads run the combined batches */
    const char *de = getenv("SMALT_B200_DEVTHREADS");
    /* three device threads = three batches on the device: the third fills the gaps the host synchronisations of the
     * other two leave (1 M C2 reads, 14 workers: 91 ms with two, 85 ms with three) */
    const int nblk = 2 * nworkers + 8, ndev = (de && atoi(de) >= 1 && atoi(de) <= 4) ? atoi(de) : 3;
    pthread_t dev[4];
    int k;
    if (g_fm_nblocks < nblk) {
      FmBlock *hp = (FmBlock *) realloc(g_fm_blocks, (size_t) nblk * sizeof(FmBlock));
      if (!hp) errcode = ERRCODE_NOMEM;
      else {
	memset(hp + g_fm_nblocks, 0, (size_t) (nblk - g_fm_nblocks) * sizeof(FmBlock));
	g_fm_blocks = hp;
	g_fm_nblocks = nblk;
      }
    }
    if (!errcode) {
      pthread_cond_init(&fm.cond, NULL);
      fm.blocks = g_fm_blocks; fm.nblocks = g_fm_nblocks;
      fm.free_blk = fm.ready = NULL;
      for (k = g_fm_nblocks - 1; k >= 0; k--) {
	g_fm_blocks[k].inflight = 0; g_fm_blocks[k].tk.slot = -1;
	g_fm_blocks[k].next = fm.free_blk; fm.free_blk = g_fm_blocks + k;
      }
      waveCombinerRestart(g_fm_comb);
      waveCombinerSetRunning(g_fm_comb, ndev);
      tid = (pthread_t *) calloc((size_t) nworkers, sizeof(pthread_t));
      for (k = 0; k < ndev; k++) pthread_create(dev + k, NULL, fm_device_main, &fm);
      for (i = 0; i < nworkers; i++) pthread_create(tid + i, NULL, fm_worker_pipe, g_fm_workers + i);
      for (i = 0; i < nworkers; i++) pthread_join(tid[i], NULL);
      waveCombinerFlush(g_fm_comb, 1);
      for (k = 0; k < ndev; k++) pthread_join(dev[k], NULL);
      free(tid);
      pthread_cond_destroy(&fm.cond);
      errcode = fm.errcode;
    }
  } else if (!errcode) {
    if (nworkers == 1) {
      fm_worker_main(g_fm_workers);
    } else {
      tid = (pthread_t *) calloc((size_t) nworkers, sizeof(pthread_t));
      for (i = 0; i < nworkers; i++) pthread_create(tid + i, NULL, fm_worker_main, g_fm_workers + i);
      for (i = 0; i < nworkers; i++) pthread_join(tid[i], NULL);
      free(tid);
    }
    errcode = fm.errcode;
  }
  for (i = 0; i < nworkers; i++) fm_collect_stats(g_fm_workers + i);
  if (g_fm_comb) {
    RmapWave *waves[8];
    uint64_t cc[2];
    const int ns = waveCombinerSlots(g_fm_comb, waves, cc);
    for (i = 0; i < ns && i < 8; i++) { g_fm_comb_stats[i].wave = waves[i]; fm_collect_stats(g_fm_comb_stats + i); }
    if (getenv("SMALT_B200_TIMING"))
      fprintf(stderr, "smalt_b200 timing: %llu device batches, %.0f reads each\n", (unsigned long long) cc[0],
	      cc[0] ? (double) cc[1] / (double) cc[0] : 0.0);
  }
  for (p = 0; p < fm.nchunks; p++) free(fm.out[p].buf);
  free(fm.out);
  fm_lines_free(fm.linesA);
  fm_lines_free(fm.linesB);
  pthread_mutex_destroy(&fm.lock);
  if (n_reads) *n_reads = fm.n_reads;
  return errcode;
}

static void fastmap_cleanup(void)
{
  int i;
  for (i = 0; i < g_fm_nworkers; i++) {
    FmWorker *w = g_fm_workers + i;
    size_t k;
    rmapWaveDelete(w->wave);
    smbShimReportWriterDelete(w->writer);
    reportDelete(w->rep);
    for (k = 0; k < w->n_alloc; k++) seqFastqDelete(w->reads[k]);
    free(w->reads);
    free(w->mincov);
    for (k = 0; k < w->m_alloc; k++) seqFastqDelete(w->mates[k]);
    free(w->mates);
    free(w->mincov_m);
    free(w->pairs);
    if (w->memfd > 0) close(w->memfd);
    if (w->keyfp) { fclose(w->keyfp); free(w->keybuf); }
    free(w->scratch);
    free(w->scratch2);
    if (w->errmsgp) { ERRMSG_END(w->errmsgp); }
  }
  free(g_fm_workers);
  g_fm_workers = NULL;
  g_fm_nworkers = 0;
  waveCombinerDelete(g_fm_comb);
  g_fm_comb = NULL;
  for (i = 0; i < g_fm_nblocks; i++) {
    size_t k;
    for (k = 0; k < g_fm_blocks[i].n_alloc; k++) seqFastqDelete(g_fm_blocks[i].reads[k]);
    free(g_fm_blocks[i].reads);
    free(g_fm_blocks[i].mincov);
  }
  free(g_fm_blocks);
  g_fm_blocks = NULL;
  g_fm_nblocks = 0;
  memset(g_fm_comb_stats, 0, sizeof(g_fm_comb_stats));
}
