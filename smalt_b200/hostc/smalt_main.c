/* smalt_main.c - the `smalt_b200` driver: the reference's own driver with the hot path on a B200.
 *
 * The reference driver (command line, FASTQ reader, work queue, SAM writer;
 * /root/reference/src/smalt.c, menu.c, threads.c, infmt.c, report.c) is used UNCHANGED: this
 * translation unit compiles smalt.c in place (read-only tree on the include path, nothing is
 * copied) with its main() renamed, and hooks in at exactly one point - the registration of the
 * PROC task of the work queue (smalt.c:1369-1375):
 *   * the per-read-block worker processArgBlock (smalt.c:1221) is replaced by
 *     smb_processArgBlock below, which maps the single-end reads of a block in GPU waves
 *     (rmap_wave.c) instead of one rmapSingle call per read, and
 *   * the block size (smalt.c:466, 32 reads per thread) is raised so that a block is a useful
 *     GPU batch.
 * Paired reads and modes the wave path does not cover are passed to the reference's own
 * processArgBlock, whose hot-path calls then go through the shim one call at a time (GPU
 * batches of one - slow, but never a CPU hot path).
 * `smalt_b200 index` is the reference's CPU index builder (not on the hot path).
 */
#include <pthread.h>
#include <time.h>
#define main ref_smalt_main
#include "smalt.c"
#undef main

#include "rmap_wave.h"

static THREAD_PROCF *g_ref_procf;
static short g_blocksz = 8192;
static pthread_mutex_t g_stats_lock = PTHREAD_MUTEX_INITIALIZER;
static double g_ms[3];
static uint64_t g_counts[5];
static double g_wall[8], g_wall_enc, g_t0;

typedef struct {
  RmapWave *wave;
  SeqFastq **reads;
  uint32_t *mincov;
  short n_alloc;
  double ms_prev[3], wall_prev[8];
  uint64_t counts_prev[5];
} WorkerState;

static __thread WorkerState t_ws;

typedef struct {
  const SmaltMapConst *macop;
  SmaltArgBlock *blockp;
} EmitArg;

static int emitResult(void *user, int i, const ResultSet *rsltp)
{
  EmitArg *ea = (EmitArg *) user;
  return resultSetAddToReport(ea->blockp->iobfp[i].rep, ea->macop->rsltouflg, rsltp);
}

static void flushStats(void)
{
  const char *fn = getenv("SMALT_B200_STATS");
  FILE *fp;
  if (!fn) return;
  fp = fopen(fn, "w");
  if (!fp) return;
  fprintf(fp, "{\"k1_ms\": %.3f, \"k2_ms\": %.3f, \"k3_ms\": %.3f, \"reads\": %llu, \"k2_tasks\": %llu, "
	  "\"k2_cells\": %llu, \"k3_tasks\": %llu, \"k3_cells\": %llu}\n", g_ms[0], g_ms[1], g_ms[2],
	  (unsigned long long) g_counts[0], (unsigned long long) g_counts[1], (unsigned long long) g_counts[2],
	  (unsigned long long) g_counts[3], (unsigned long long) g_counts[4]);
  fprintf(fp, "{\"host_wall_s\": {\"staging\": %.3f, \"seed\": %.3f, \"hits\": %.3f, \"candidates\": %.3f, "
	  "\"score\": %.3f, \"replay\": %.3f, \"align\": %.3f, \"results\": %.3f, \"encode\": %.3f}}\n",
	  g_wall[0], g_wall[1], g_wall[2], g_wall[3], g_wall[4], g_wall[5], g_wall[6], g_wall[7], g_wall_enc);
  fclose(fp);
}

/* THREAD_PROCF replacing processArgBlock (smalt.c:1221) */
static int smb_processArgBlock(ErrMsg *errmsgp, void *targp, void *bufargp)
{
  int errcode = ERRCODE_SUCCESS;
  short i;
  SmaltMapArgs *map = (SmaltMapArgs *) targp;
  SmaltArgBlock *blockp = (SmaltArgBlock *) bufargp;
  const SmaltMapConst *macop = map->smconstp;
  const short n = blockp->n_iobf;
  EmitArg ea;
  double ms[3], wall[8], t_enc;
  uint64_t counts[5];
  struct timespec ts0, ts1;

  for (i = 0; i < n; i++)
    if (blockp->iobfp[i].isPaired)
      return (*g_ref_procf)(errmsgp, targp, bufargp);
  if (macop->tupcovmin < 0)
    return ERRCODE_ASSERT;
  if (getenv("SMALT_B200_TIMING")) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    fprintf(stderr, "smalt_b200 timing: block of %d reads enters PROC at %.3f s\n", (int) n, ts.tv_sec + 1e-9 * ts.tv_nsec - g_t0);
  }
  if (!t_ws.wave) {
    t_ws.wave = rmapWaveCreate(macop->htp, macop->ssp, macop->codecp, macop->scormtxp);
    if (getenv("SMALT_B200_TIMING")) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    fprintf(stderr, "smalt_b200 timing: block of %d reads enters PROC at %.3f s\n", (int) n, ts.tv_sec + 1e-9 * ts.tv_nsec - g_t0);
  }
  if (!t_ws.wave) {
      fprintf(stderr, "smalt_b200: cannot set up the GPU context of a worker thread\n");
      return ERRCODE_FAILURE;
    }
  }
  if (n > t_ws.n_alloc) {
    t_ws.reads = (SeqFastq **) realloc(t_ws.reads, (size_t) n * sizeof(SeqFastq *));
    t_ws.mincov = (uint32_t *) realloc(t_ws.mincov, (size_t) n * sizeof(uint32_t));
    if (!t_ws.reads || !t_ws.mincov) return ERRCODE_NOMEM;
    t_ws.n_alloc = n;
  }
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  for (i = 0; i < n; i++) { /* per-read preparation of processMapArgs (smalt.c:1106-1127) */
    SeqFastq *readp = blockp->iobfp[i].readp;
    uint32_t covermin_tuple;
    if ((errcode = seqFastqEncode(readp, macop->codecp))) {
      ERRMSGNO(errmsgp, errcode);
      return errcode;
    }
    if (macop->tupcovmin < 1.01) {
      uint32_t readlen;
      seqFastqGetConstSequence(readp, &readlen, NULL);
      if (readlen > INT_MAX) ERRMSGNO(errmsgp, ERRCODE_OVERFLOW);
      covermin_tuple = (uint32_t) (macop->tupcovmin * readlen);
      if (covermin_tuple > readlen) covermin_tuple = readlen;
    } else {
      covermin_tuple = (uint32_t) macop->tupcovmin;
    }
    t_ws.reads[i] = readp;
    t_ws.mincov[i] = covermin_tuple;
  }
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  t_enc = (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec);
  ea.macop = macop;
  ea.blockp = blockp;
  if (getenv("SMALT_B200_IOTEST")) { /* diagnostic: measure the driver's input/output stages alone */
    rmapBlank(map->rmp);
    for (i = 0; i < n; i++) {
      const ResultSet *rsltp;
      rmapGetData(&rsltp, NULL, NULL, NULL, NULL, map->rmp);
      if ((errcode = emitResult(&ea, i, rsltp))) return errcode;
    }
    return ERRCODE_SUCCESS;
  }
  errcode = rmapSingleWave(errmsgp, map->rmp, t_ws.wave, n, t_ws.reads, t_ws.mincov,
			   macop->nhitmax_tuple, (int) macop->min_swatscor, macop->swatscordiff, macop->minbasq,
			   SMALT_TARGET_DEPTH, SMALT_MAX_DEPTH, (RMAPFLG_t) (macop->rmapflg & ~RMAPFLG_ALLPAIR),
			   macop->scormtxp, macop->rfp, macop->htp, macop->ssp, macop->codecp,
			   emitResult, &ea);
  if (errcode == ERRCODE_ARGINVAL) /* mode not covered by the wave path */
    return (*g_ref_procf)(errmsgp, targp, bufargp);
  rmapWaveGetStats(t_ws.wave, ms, counts);
  rmapWaveGetWall(t_ws.wave, wall);
  pthread_mutex_lock(&g_stats_lock);
  g_wall_enc += t_enc;
  for (i = 0; i < 8; i++) { g_wall[i] += wall[i] - t_ws.wall_prev[i]; t_ws.wall_prev[i] = wall[i]; }
  for (i = 0; i < 3; i++) { g_ms[i] += ms[i] - t_ws.ms_prev[i]; t_ws.ms_prev[i] = ms[i]; }
  for (i = 0; i < 5; i++) { g_counts[i] += counts[i] - t_ws.counts_prev[i]; t_ws.counts_prev[i] = counts[i]; }
  pthread_mutex_unlock(&g_stats_lock);
  return errcode;
}

int __real_threadsSetTask(uint8_t task_typ, short n_threads, THREAD_INITF *initf, const void *initargp,
			  THREAD_PROCF *procf, THREAD_CLEANF *cleanf, THREAD_CHECKF *checkf,
			  THREAD_CMPF *cmpf, size_t argsz);

int __wrap_threadsSetTask(uint8_t task_typ, short n_threads, THREAD_INITF *initf, const void *initargp,
			  THREAD_PROCF *procf, THREAD_CLEANF *cleanf, THREAD_CHECKF *checkf,
			  THREAD_CMPF *cmpf, size_t argsz)
{
  if (task_typ == THRTASK_ARGBUF && argsz == sizeof(SmaltArgBlock)) {
    /* a block of reads is the GPU batch: raise smalt.c:466's 32 reads per thread */
    const char *e = getenv("SMALT_B200_BLOCK");
    long b = e ? atol(e) : g_blocksz;
    if (b < 1) b = 1;
    if (b > 32000) b = 32000;
    ((SmaltMapConst *) initargp)->threadblksz = (short) b;
  } else if (task_typ == THRTASK_PROC && argsz == sizeof(SmaltMapArgs)) {
    g_ref_procf = procf;
    procf = smb_processArgBlock;
  }
  return __real_threadsSetTask(task_typ, n_threads, initf, initargp, procf, cleanf, checkf, cmpf, argsz);
}

int main(int argc, char *argv[])
{
  struct timespec ts;
  int rv;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  g_t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
  atexit(flushStats);
  rv = ref_smalt_main(argc, argv);
  if (getenv("SMALT_B200_TIMING")) {
    clock_gettime(CLOCK_MONOTONIC, &ts);
    fprintf(stderr, "smalt_b200 timing: main %.3f s\n", ts.tv_sec + 1e-9 * ts.tv_nsec - g_t0);
  }
  return rv;
}
