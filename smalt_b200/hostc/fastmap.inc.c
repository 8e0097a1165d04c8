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

typedef struct FastMap_ {
  const SmaltMapConst *macop;
  SmaltMapArgs *maps;        /* one per worker (threadsGetMem(THRTASK_PROC)) */
  const ReportWriter *proto;
  const char *data;
  size_t len;
  int is_fasta;
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
  int memfd;
  char fdpath[64];
  char *scratch;             /* NUL-terminated copies of the lines of one record */
  size_t scratch_alloc;
  FILE *keyfp;               /* stream handed to the report writer while its output is captured */
  char *keybuf;
  size_t keylen;
  ErrMsg *errmsgp;
  double ms_prev[3], wall_prev[11], cpu_prev[8];
  uint64_t counts_prev[5];
} FmWorker;

static FmWorker *g_fm_workers;
static int g_fm_nworkers;

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

static int fm_has_space(const char *p, size_t len)
{
  size_t i;
  unsigned acc = 0;
  for (i = 0; i < len; i++) {
    const unsigned char c = (unsigned char) p[i];
    acc |= (unsigned) (c == ' ') | (unsigned) (c - 9u <= 4u);
  }
  return acc != 0;
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
static int fm_parse_block_fast(FmWorker *w, size_t start, size_t end, size_t *nreads)
{
  const char *d = w->fm->data;
  size_t p = start, n = 0;
  int errcode;
  *nreads = 0;
  while (p < end) {
    const char *l[4];
    size_t ll[4], need;
    int k;
    char *name, *seq, *qnam, *qual;
    for (k = 0; k < 4; k++) {
      const char *nl = (p < end) ? (const char *) memchr(d + p, '\n', end - p) : NULL;
      if (!nl) return 1;
      l[k] = d + p;
      ll[k] = (size_t) (nl - (d + p));
      p = (size_t) (nl - d) + 1;
    }
    if (ll[0] < 1 || l[0][0] != '@' || ll[2] < 1 || l[2][0] != '+' || ll[1] < 1 || ll[1] != ll[3] ||
	fm_has_space(l[1], ll[1]) || fm_has_space(l[3], ll[3]))
      return 1;
    need = ll[0] + ll[1] + ll[2] + ll[3] + 8;
    if (need > w->scratch_alloc) {
      char *hp = (char *) realloc(w->scratch, 2 * need);
      if (!hp) return ERRCODE_NOMEM;
      w->scratch = hp;
      w->scratch_alloc = 2 * need;
    }
    name = w->scratch;
    seq = name + ll[0] + 1;
    qnam = seq + ll[1] + 1;
    qual = qnam + ll[2] + 1;
    {
      size_t nl_ = fm_header_name(name, l[0], ll[0]), ql_ = fm_header_name(qnam, l[2], ll[2]);
      /* setSeq (sequence.c:780-803) also strips white space at both ends of what it is given */
      while (nl_ > 0 && isspace((unsigned char) name[nl_ - 1])) nl_--;
      while (ql_ > 0 && isspace((unsigned char) qnam[ql_ - 1])) ql_--;
      (void) seq; (void) qual;
      if ((errcode = fm_reads_reserve(w, n))) return errcode;
      seqFastqBlank(w->reads[n]);
      if ((errcode = smbShimSeqFastqLoad(w->reads[n], name, nl_, l[1], ll[1], qnam, ql_, l[3], ll[3])))
	return errcode;
    }
    n++;
  }
  *nreads = n;
  return ERRCODE_SUCCESS;
}

static int fm_parse_block(FmWorker *w, size_t start, size_t end, size_t *nreads)
{
  FastMap *fm = w->fm;
  int errcode = ERRCODE_SUCCESS;
  size_t n = 0, nrec_expect = 0, off;
  SeqIO *sio;
  *nreads = 0;
  if (end <= start) return ERRCODE_SUCCESS;
  if (!fm->is_fasta && (errcode = fm_check_fastq(fm->data, start, end, &nrec_expect))) {
    fprintf(stderr, "smalt_b200: the read file is not plain 4-line FASTQ near byte %zu; "
	    "rerun with SMALT_B200_REFIO=1 (the reference's own reader)\n", start);
    return errcode;
  }
  if (!fm->is_fasta && !getenv("SMALT_B200_REFPARSE")) {
    errcode = fm_parse_block_fast(w, start, end, &n);
    if (errcode != 1) {
      if (!errcode && n != nrec_expect) errcode = ERRCODE_FASTA;
      *nreads = n;
      return errcode;
    }
    errcode = ERRCODE_SUCCESS;
    n = 0;
  }
  if (ftruncate(w->memfd, 0)) return ERRCODE_FILEIO;
  for (off = start; off < end;) {
    const ssize_t k = pwrite(w->memfd, fm->data + off, end - off, (off_t) (off - start));
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
  errcode = fm_parse_block(w, start, end, &n);
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

static void *fm_worker_main(void *arg)
{
  FmWorker *w = (FmWorker *) arg;
  FastMap *fm = w->fm;
  int rc = fm_worker_setup(w, fm, w->id);
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
    if (fm_map_block(w, c)) break;
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
  for (i = 0; i < 5; i++) { g_counts[i] += counts[i] - w->counts_prev[i]; w->counts_prev[i] = counts[i]; }
  pthread_mutex_unlock(&g_stats_lock);
}

/* Maps the reads in data[0..len) and hands the formatted output to sinkf in input order. */
static int fastmap_run(const SmaltMapConst *macop, SmaltMapArgs *maps, int nworkers, const ReportWriter *proto,
		       const char *data, size_t len, FASTMAP_SINKF *sinkf, void *sink_user, uint64_t *n_reads)
{
  FastMap fm;
  pthread_t *tid;
  int i, errcode = ERRCODE_SUCCESS;
  size_t p = 0, block = 8192, rec_bytes;
  const char *e = getenv("SMALT_B200_BLOCK");

  memset(&fm, 0, sizeof(fm));
  while (p < len && isspace((unsigned char) data[p])) p++;
  if (p >= len) { if (n_reads) *n_reads = 0; return ERRCODE_SUCCESS; }
  fm.macop = macop; fm.maps = maps; fm.proto = proto;
  fm.data = data + p; fm.len = len - p;
  fm.is_fasta = data[p] == '>';
  fm.nworkers = nworkers;
  fm.sinkf = sinkf; fm.sink_user = sink_user;
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
  if (e && atol(e) > 0) block = (size_t) atol(e);
  else { /* at least ~6 blocks per worker for load balance, at least 512 reads per block */
    const size_t est_reads = fm.len / rec_bytes + 1;
    size_t b = est_reads / ((size_t) nworkers * 6) + 1;
    if (b < 512) b = 512;
    if (b < block) block = b;
  }
  if (block > 32000) block = 32000;
  fm.chunk_bytes = block * rec_bytes;
  fm.nchunks = (fm.len + fm.chunk_bytes - 1) / fm.chunk_bytes;
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
  if (!errcode) {
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
  for (p = 0; p < fm.nchunks; p++) free(fm.out[p].buf);
  free(fm.out);
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
    if (w->memfd > 0) close(w->memfd);
    if (w->keyfp) { fclose(w->keyfp); free(w->keybuf); }
    free(w->scratch);
    if (w->errmsgp) { ERRMSG_END(w->errmsgp); }
  }
  free(g_fm_workers);
  g_fm_workers = NULL;
  g_fm_nworkers = 0;
}
