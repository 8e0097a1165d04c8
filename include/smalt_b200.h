/* smalt_b200.h - C ABI of the B200-native SMALT hot path (libsmalt_b200.so).
 *
 * Plain C: opaque context, plain pointers and sizes, no C++/torch types.
 * Every entry point returns 0 (SMB_OK) or an error code; codes <  100 are the
 * reference's own ERRCODE_* values (/root/reference/src/elib.h:49-139) so a
 * caller that branches on them (rmap.c:730 ERRCODE_SWATEXCEED, rmap.c:748,
 * rmap.c:1695 ERRCODE_SHORTSEQ ...) keeps working; codes >= 100 are specific
 * to this library.  smb_last_error() gives a text for the last failure.
 *
 * There is no CPU fallback: without a CUDA device smb_ctx_create() fails with
 * SMB_ERR_NODEVICE and every compute entry point needs a context.
 *
 * Sequences cross the ABI as one byte per base holding the reference's 3-bit
 * alphabet code in the low 3 bits (A0 C1 G2 T3 X4 N5; SEQCOD_ALPHA_MASK,
 * sequence.h:98) - the reference's SEQCOD_MANGLED bytes can be passed as they
 * are, only `code & 7` is used (as in swsimd.c:725, alignment.c:876).
 *
 * Reference interfaces replaced (file:line under /root/reference/src):
 *   smb_sw_score_batch    <- swSIMDAlignStriped      swsimd.h:47-55  (swsimd.c:868)
 *   smb_band_score_batch  <- aliSmiWatInBandFast     alignment.h:147-175 (alignment.c:1603)
 *   smb_band_align_batch  <- aliSmiWatInBand + aliRsltSetGetSize/FetchData
 *                                                    alignment.h:67-145 (alignment.c:1548, :1513, :1518)
 *   smb_index_upload      <- hashTableRead           hashidx.h:186-190 (hashidx.c:1257)
 *   smb_refseq_upload     <- seqSetReadBinFil        sequence.h:444 (sequence.c:2521)
 *   smb_seed_batch        <- hashCollectHitInfoShort hashhit.h (hashhit.c:1007) for both strands
 *                            + hashCalcHitInfoCoverDeficit (:1096) + hashHitInfoCalcHitNumbers (:1200)
 *                            + hashCalcHitInfoNumberOfHits (:1171)
 *   smb_hits_batch        <- hashCollectHitsForSegment (hashhit.c:1691) / hashCollectHitsUsingCutoff
 *                            (:1593) + hashGetHitListData (:1867)
 * The reference-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Not covered (the reference's symbols on top of this ABI, hostc/shim_hot.c, print one message and
 * fail with ERRCODE_ARGINVAL / a NULL constructor result instead of computing on the CPU):
 *   map -w   complexity-weighted scores: scaleALICPLX (alignment.c:268-305) rescoring in floating point
 *            inside aliSmiWatInBand
 *   map -p   split reads: hashCollectHitInfo on a read segment (hashhit.c:987 with seq_start / seq_end)
 *   HashHitFilter arguments of hashCollectHitsForSegment (no caller in the smalt driver passes one)
 */
#ifndef SMALT_B200_H
#define SMALT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  SMB_OK = 0,
  /* reference codes (elib.h) */
  SMB_ERRCODE_FAILURE = -1,
  SMB_ERRCODE_NOMEM = 2,
  SMB_ERRCODE_ARGRANGE = 29,
  SMB_ERRCODE_SHORTSEQ = 30,
  SMB_ERRCODE_ALLOCBOUNDARY = 32,
  SMB_ERRCODE_SWATEXCEED = 41,
  SMB_ERRCODE_SWATSCOR = 44,
  SMB_ERRCODE_ASSERT = 47,
  SMB_ERRCODE_OVERFLOW = 48,
  SMB_ERRCODE_DIFFSTR = 59,
  /* library codes */
  SMB_ERR_NODEVICE = 100,   /* no CUDA device / driver: there is no CPU fallback */
  SMB_ERR_CUDA = 101,       /* a CUDA runtime call failed (see smb_last_error) */
  SMB_ERR_ARG = 102,        /* invalid argument */
  SMB_ERR_CAPACITY = 103,   /* per-task output capacity exceeded (results / diffstr / stack) */
  SMB_ERR_STATE = 104       /* call order: index / reference / reads not uploaded yet */
};

typedef struct smb_ctx smb_ctx;

/* ----------------------------- context ---------------------------------- */
int smb_ctx_create(smb_ctx **ctx, int device);
/* Optional: initialises CUDA on `device` and loads the kernels.  A driver calls it from a helper
 * thread at program start so that the ~0.7 s of CUDA start-up overlap its own index loading. */
int smb_device_warmup(int device);
void smb_ctx_destroy(smb_ctx *ctx);
const char *smb_last_error(const smb_ctx *ctx);
/* library version string, e.g. "smalt-b200 0.1 sm_100a" */
const char *smb_version(void);
/* Alignment penalties as given to `smalt map -S` (score.c:41-47 defaults
 * match=1 mismatch=-2 gapopen=-4 gapext=-3); builds the 8x8 matrix of
 * score.c:138-173 in constant memory. */
int smb_set_scoring(smb_ctx *ctx, int match, int mismatch, int gapopen, int gapext);
/* Device time in ms of the kernels launched by the last batch call, measured
 * with CUDA events on the context's stream (no host copies included), and the
 * number of kernel launches of that call. */
float smb_last_kernel_ms(const smb_ctx *ctx);
int smb_last_kernel_launches(const smb_ctx *ctx);
/* total kernels launched by this context since creation */
long long smb_total_kernel_launches(const smb_ctx *ctx);

/* process-wide totals over all contexts: kernels launched, bytes copied host->device and
 * device->host by this library */
void smb_process_counters(unsigned long long *launches, unsigned long long *h2d_bytes,
			  unsigned long long *d2h_bytes);

/* Page-locked host memory for the buffers that cross this ABI (reads, task lists, outputs).
 * Any host pointer is accepted by every entry point; buffers from smb_host_alloc make the
 * host<->device copies asynchronous DMA transfers instead of staged copies.  NULL on failure. */
void *smb_host_alloc(size_t nbytes);
void smb_host_free(void *p);

/* How the calling thread waits for the device inside the batch calls: 0 (default) = it sleeps on a
 * blocking-sync event (one context per worker thread, more threads than cores); 1 = it polls, for a thread
 * that drives the device for many others and must not wait for a time slice on a busy host; 2 = it polls
 * with sleeps of 20 us in between (n > 2: of n us), which costs next to no CPU time - call it from the
 * thread that will wait (the timer slack of that thread is lowered). */
int smb_ctx_set_spin(smb_ctx *ctx, int spin);

/* Makes `dst` use the index and packed reference already uploaded to `src`
 * (same device) instead of holding its own copy: one resident copy per GPU,
 * one context (stream + scratch buffers) per host worker thread. */
int smb_ctx_share_index(smb_ctx *dst, const smb_ctx *src);

/* Integer-issue micro-benchmark used as the roofline denominator of the DP kernels:
 * sustained giga thread-operations per second of (0) VIADDMNMX  max(a+b,c),
 * (1) VIMNMX3 max(a,b,c), (2) plain IADD+IMNMX pairs, (3) VIADDMNMX on two 16-bit halves
 * (instructions, i.e. two cell-operations each) and (4) VIMNMX3 on two 16-bit halves,
 * measured on this device. */
int smb_int_peak(smb_ctx *ctx, double gops[5]);

/* --------------------------- sequence arena ------------------------------ */
/* Uploads a block of concatenated sequences (reads and, for the *_batch calls
 * that take explicit windows, reference windows) to HBM.  Tasks address it by
 * byte offset.  A later upload replaces the previous one. */
int smb_arena_upload(smb_ctx *ctx, const uint8_t *codes, size_t nbytes);

/* The packed reference of a `.sma` file: 3 bits per base, 10 bases per 32-bit
 * word, base i in bits 3*(9 - i%10) of word i/10 (sequence.c:1360-1424);
 * seq_offs[nseq+1] are the base offsets of the sequences in the concatenated
 * set (each sequence is followed by one terminator, code 7).  Uploaded once per
 * GPU; tasks flagged SMB_TASK_REF_PACKED address it by base offset. */
int smb_refseq_upload(smb_ctx *ctx, const uint32_t *words, size_t nwords,
		      uint64_t nbases, const uint64_t *seq_offs, int nseq);

/* ------------------------------- tasks ----------------------------------- */
enum {
  SMB_TASK_READ_REVCOMP = 1, /* profile the reverse complement of the read (rmap.c:682-692) */
  SMB_TASK_REF_PACKED = 2    /* ref_off/ref_len address the packed reference, not the arena */
};

typedef struct {
  uint64_t read_off; /* byte offset of the read in the arena */
  uint64_t ref_off;  /* offset of the reference window (arena bytes or packed bases) */
  uint32_t read_len;
  uint32_t ref_len;
  uint32_t flags;
  uint32_t reserved;
} smb_sw_task;

typedef struct {
  uint64_t read_off;
  uint64_t ref_off;
  uint32_t read_len;
  uint32_t ref_len;
  uint32_t flags;
  int32_t l_edge, r_edge;   /* band on the read axis at window row 0 (alignment.h:96-100) */
  int32_t p_left, p_right;  /* read sub-range */
  int32_t u_left, u_right;  /* window sub-range */
  int32_t minscore;         /* smb_band_align_batch only */
  int32_t minscorlen;       /* smb_band_align_batch only */
} smb_band_task;

typedef struct {
  int32_t score, qs, qe, rs, re; /* ALIRESULT, alignment.c:149-158 (0-based, inclusive) */
  uint32_t diff_off;             /* offset of the DiffStr in the diffstr output buffer */
  uint32_t diff_len;             /* bytes including the terminating 0 */
  uint32_t task;                 /* index of the task that produced it */
} smb_ali_result;

/* K2: maximum local alignment score of each read x window pair, unbanded,
 * canonical affine gaps; scores[i] and errs[i] per task (errs[i] is
 * SMB_ERRCODE_SWATEXCEED where the reference would return it, swsimd.c:644). */
int smb_sw_score_batch(smb_ctx *ctx, const smb_sw_task *tasks, int ntasks,
		       int32_t *scores, int32_t *errs);

/* K2': maximum score of the banded "fast" DP (alignment.c:1029-1233);
 * errs[i] = SMB_ERRCODE_FAILURE where the band misses the read (alignment.c:1622-1627). */
int smb_band_score_batch(smb_ctx *ctx, const smb_band_task *tasks, int ntasks,
			 int32_t *scores, int32_t *errs);

/* K3: banded DP with backtrace and recursion (alignment.c:1300-1434).
 * Results of all tasks are appended to `results` in task order and, within a
 * task, in the reference's discovery order; first_result[i] .. first_result[i+1]
 * delimit task i (first_result has ntasks+1 entries).  DiffStr bytes
 * (diffstr.h:28-105, forward orientation as stored by addALIMETAtoRsltSet,
 * alignment.c:1294) go to `diffstr`.  Returns SMB_ERR_CAPACITY if the
 * caller's buffers are too small (nothing partial is hidden: *nresults and
 * *ndiffbytes then hold the required sizes). */
int smb_band_align_batch(smb_ctx *ctx, const smb_band_task *tasks, int ntasks,
			 smb_ali_result *results, size_t max_results, size_t *nresults,
			 uint32_t *first_result,
			 uint8_t *diffstr, size_t max_diffbytes, size_t *ndiffbytes,
			 int32_t *errs, uint64_t *ncells);

/* ------------------------------ K1: seeds -------------------------------- */
/* Hash index as stored in `.smi` (hashidx.c:1214-1255) and as hashTableRead
 * leaves it in memory (posidx[nwords] is NOT read from the file and stays 0,
 * hashidx.c:1334 - pass the arrays exactly as read).  typ 0 = perfect, 1 =
 * with collisions.  Uploaded once per GPU. */
int smb_index_upload(smb_ctx *ctx, int typ, int wordlen, int nskip, int nbits_key,
		     int nbits_lo, uint32_t npos, uint32_t nwords,
		     const uint32_t *idx, const uint32_t *pos,
		     const uint32_t *wordidx, const uint32_t *posidx);

/* Per read x strand seed table (HashHitInfo, hashhit.c:164-213) in SoA form. */
typedef struct {
  uint32_t n_seeds;       /* number of seeds (k-mers with 1..maxhit hits) */
  uint32_t seed_rank;     /* hashhit.c:769-891 */
  uint32_t cover_deficit; /* hashCalcHitInfoCoverDeficit */
  uint32_t nhit_rank;     /* hashHitInfoCalcHitNumbers: hits of the seeds below seed_rank */
  uint32_t nhit_tot;      /*   ... and of all seeds */
  uint32_t nhit_all;      /* hashCalcHitInfoNumberOfHits(maxhit_per_tuple) */
  uint32_t status;        /* HITINFO_STATUS_FLAGS */
  int32_t err;            /* 0 or SMB_ERRCODE_SHORTSEQ */
} smb_seed_info;

/* Seeds of `nreads` reads (arena offsets/lengths), both strands: entry
 * 2*r+s is read r, strand s (0 forward, 1 reverse).  Per-seed arrays are
 * written at seed_off[2*r+s] = (read offset counted in bases over all previous
 * reads + strand*read_len) - i.e. each strand of each read owns read_len slots:
 *   seed_posidx/nhits/qoffs: SEED fields in discovery order,
 *   sortkey/sidx: nhitqual_sortkeyp / sidxp after the reference's quicksort,
 *   qmask: HITQUAL codes per read offset (hashhit.h:57-65).
 * qual may be NULL (FASTA); basq_thresh as `smalt map -q`.  short_info != 0:
 * hashCollectHitInfoShort (sorted + ranked, the default mode of rmap.c:1686);
 * short_info == 0: hashCollectHitInfo (unsorted, no per-seed cut, `smalt map -x`). */
int smb_seed_batch(smb_ctx *ctx, const uint64_t *read_off, const uint32_t *read_len,
		   int nreads, const uint8_t *qual, uint32_t maxhit_per_tuple,
		   uint32_t maxhit_total, int basq_thresh, int short_info,
		   smb_seed_info *info, uint32_t *seed_posidx, uint32_t *seed_nhits,
		   uint32_t *seed_qoffs, uint32_t *sortkey, uint32_t *sidx, uint8_t *qmask);

/* Index construction on the GPU: the arrays hashTableSetUp (hashidx.c:829-998) builds for a whole
 * sequence set, byte for byte (the `.smi` file written from them equals `smalt index`'s).  The
 * 3-bit packed reference must have been uploaded (smb_refseq_upload).  typ / nbits_key / nbits_lo
 * as selectHashTyp (smalt.c:268-332) chooses them; seqs[i] describes the k-mer grid of sequence i
 * (doWordsInSeq, hashidx.c:465-531): `start` = offset of the sequence in the concatenated set,
 * `offs` = offset of its first grid position, `n_k` = grid positions in it, `tup_base` = serial
 * number of the first one.  smb_index_fetch copies the arrays out (wordidx / posidx: nwords + 1
 * entries, collision type only; NULL pointers are skipped) and releases the device copy. */
typedef struct { uint64_t start; uint32_t offs, n_k, tup_base, reserved; } smb_index_seq;
typedef struct { uint32_t npos, nwords, nkeys; float kernel_ms; } smb_index_info;
int smb_index_build(smb_ctx *ctx, int wordlen, int nskip, int typ, int nbits_key, int nbits_lo,
		    const smb_index_seq *seqs, int nseq, smb_index_info *info);
int smb_index_fetch(smb_ctx *ctx, uint32_t *idx, uint32_t *pos, uint32_t *wordidx, uint32_t *posidx);

/* Seed tables against per-read SMALL indexes: the reference builds a k=5, s=1 perfect-hash
 * table on the fly from the insert-size intervals around a mapped mate (setupFineHashTable,
 * rmap.c:495-517: hashTableSetUp with an InterVal, hashidx.c:549-575, :829-998) and collects
 * the seeds of the other mate in it (initRMAPINFO on rmp->htflyp, rmap.c:2026-2030).  One such
 * table exists per pair, so a batch carries `ntables` tables (HASHIDXTYP_PERFECT only: idx has
 * 4^wordlen + 1 entries, pos has npos; all tables share wordlen and nskip) and read_table[r]
 * names the table of read r.  Reads are addressed in the arena like in smb_seed_batch; the
 * seed tables stay on the device for smb_hits_batch, which then reads the per-read tables. */
typedef struct {
  const uint32_t *idx;
  const uint32_t *pos;
  uint32_t npos;
} smb_small_index;
int smb_seed_batch_tables(smb_ctx *ctx, int wordlen, int nskip, const smb_small_index *tables, int ntables,
			  const uint32_t *read_table, const uint64_t *read_off, const uint32_t *read_len, int nreads,
			  const uint8_t *qual, uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq_thresh,
			  int short_info, smb_seed_info *info);

/* One hit list to build from the seed tables left on the device by the last
 * smb_seed_batch: the hits of read `read`, strand `strand`, that fall into the
 * reference segment [lo, hi) given as base offsets in the concatenated set -
 * what collectHits (rmap.c:283-318) asks hashCollectHitsForSegment for with
 * lo = soffs[s], hi = soffs[s+1], nhit_max = ktuple_maxhit, use_short = 1. */
typedef struct {
  uint64_t lo, hi;       /* segment [lo, hi) in bases of the concatenated set (ignored for mode 2) */
  uint32_t read;
  uint32_t nhit_max;     /* per-seed cut-off (ktuple_maxhit) */
  uint8_t strand;
  uint8_t use_short;     /* 1: ranked seeds, 0: all seeds unsorted (hashCollectHitsForSegment);
			    2: whole set with cut-off = hashCollectHitsUsingCutoff (hashhit.c:1593-1689),
			       the path of rmap.c:320-346 for >= 512 reference sequences */
  uint8_t reserved[2];
  uint32_t nhits_max;    /* mode 2: HashHitList.nhits_max of this read, qlen*ln(qlen)*32 clamped to
			    [8192, INT_MAX] (hashhit.c:1266-1288; 0 = let the library derive it) */
} smb_hit_req;

/* Builds and sorts the requested hit lists (HashHitList.sqdat, hashhit.c:215-236:
 * shift << 31 | read offset, ascending).  list_first[i]..list_first[i+1]
 * delimit list i in sqdat (list_first has nreq+1 entries).  nhits_alloc is the
 * reference's allocation bound of the hit list (16384-blocks >= qlen*ln(qlen)*32,
 * hashhit.c:1262-1296); 0 = derive it from the longest read of the batch.
 * Returns SMB_ERR_CAPACITY with *nhits_total = required size if max_hits is
 * too small. */
int smb_hits_batch(smb_ctx *ctx, const smb_hit_req *req, int nreq, uint32_t nhits_alloc,
		   uint64_t *sqdat, size_t max_hits, size_t *nhits_total,
		   uint64_t *list_first, int32_t *errs);

/* The HITQUAL masks (hashhit.h:57-65) of the hit lists built by the last smb_hits_batch, one
 * byte per read offset: list i at qmask[qmask_first[i] .. qmask_first[i+1]) (read_len bytes).
 * Only mode-2 lists mark NORMHIT / MULTIHIT seeds (hashhit.c:1632-1650); segment lists are all
 * HITQUAL_NOHIT like the reference's.  segLstFillHits (segment.c:782-788) consumes the mask. */
int smb_hits_qmask(smb_ctx *ctx, uint8_t *qmask, size_t max_bytes, uint64_t *qmask_first);

/* ---------------- resident block: hit lists -> candidates -> K2 -> replay -> K3 ----------------
 * A block of reads stays on the device from the seed tables (smb_seed_batch / _tables) to the
 * alignments: nothing but the jobs goes up, nothing but the per-read summaries, the aligned
 * candidates and their alignments comes down.  Replaces, for every job (one mapSingleRead pass,
 * rmap.c:1228-1433):
 *   collectHits / collectHitsFromInterVal           rmap.c:273-318, :438-493   (hit lists, as smb_hits_batch)
 *   segLstFillHits + segAliCandsAddFast             segment.c:763-810, :1530-1557 (:396-584, :1140-1223)
 *   segAliCandsStats                                segment.c:1616-1785 (NR quicksort ties, sort.c:233-330)
 *   segAliCandsCalcSegmentOffsets                   segment.c:1861-1985 (makeRMAPCANDfromSegment, rmap.c:535)
 *   scoreRMAPCAND                                   rmap.c:660-786  (K2 / K2' of every candidate + the sequential
 *                                                   early-break bookkeeping)
 *   the thresholds of mapSingleRead                 rmap.c:1373-1400
 *   alignRMAPCANDFull up to aliSmiWatInBand         rmap.c:820-911 with the INITIAL threshold (the rising
 *                                                   threshold of BEST mode is replayed by the caller on the
 *                                                   results, see DESIGN.md "threshold replay")
 * calcTotalHitNumStats (rmap.c:1086) fills nhit / nhit_tot. */
typedef struct {
  uint32_t seed_read;    /* read in the last seed batch */
  int32_t niv;           /* < 0: hit lists of every reference sequence; else number of intervals */
  uint32_t iv_first;     /* first interval of the job in `ivals` */
  uint32_t min_cover;    /* min_cover argument of mapSingleRead (before calcMinKtup, rmap.c:240-247) */
  int32_t min_swatscor;  /* absolute score threshold argument */
  uint32_t reserved;
} smb_block_job;

typedef struct {         /* hit lists restricted to [lo, hi) of the concatenated set, candidates labelled seqidx */
  uint64_t lo, hi;
  int32_t seqidx;
  int32_t reserved;
} smb_block_ival;

typedef struct {
  uint32_t nhit_max;               /* ktuple_maxhit */
  int32_t min_swatscor_below_max;  /* relative threshold argument (< 0: none) */
  int32_t target_depth, max_depth; /* as passed to segAliCandsStats (shorts converted like SEGNUM_t) */
  uint8_t best;                    /* RMAPFLG_BEST */
  uint8_t sensitive;               /* RMAPFLG_SENSITIVE */
  uint8_t termchar;                /* a terminator follows every sequence in seq_offs (SEQSET_TERMCHAR) */
  uint8_t cigar;                   /* SMB_CIGAR_* flags: CIGAR text + edit distance of every alignment (0: none) */
} smb_block_params;

/* Output stage on the device: what fprintREPALIsam (report.c:832-898) derives from the alignment
 * string of a reported alignment - the CIGAR field (writeDiffStrCIGAR, diffstr.c:298-367, through
 * diffStrPrintf with DIFFSTRFORM_CIGEXT / _XMISMATCH, :1066-1075) and the NM:i: edit distance
 * (diffStrGetLevenshteinDistance, diffstr.c:1496-1510). */
enum smb_cigar_flags {
  SMB_CIGAR_ON = 1,          /* run the stage (smb_block_params.cigar) */
  SMB_CIGAR_SOFTCLIP = 2,    /* clips as 'S' (REPORTMODIF_SOFTCLIP) instead of 'H' */
  SMB_CIGAR_XMISMATCH = 4    /* mismatches as 'X' runs (REPORTMODIF_XMISMATCH) instead of inside 'M' */
};

typedef struct {                   /* per job */
  int32_t errcode;                 /* error that ends the mapping of this read (ERRCODE_*) or 0 */
  uint8_t reached_stats;           /* got as far as resultSetAlignmentStats (rmap.c:1338) */
  uint8_t do_align;                /* max1scor >= 1 */
  uint8_t reserved[2];
  int32_t nseg, nseg_tot;          /* n_sort, n_mincover */
  uint32_t nhit, nhit_tot;
  uint32_t ncand, nscored;
  int32_t max1scor, max2scor;
  int32_t min_swatscor, scorlen_min, bandwidth_min;
  uint32_t k3_first, nk3;          /* aligned candidates of this job in `cands` */
  uint32_t reserved2;
} smb_block_read;

typedef struct {                   /* one candidate that went to K3 (RMAPCAND fields resultSetAddFromAli needs) */
  uint64_t rs;                     /* window start inside the reference sequence */
  int32_t sqidx;
  int32_t swscor;
  uint32_t reflen;
  int32_t band_l, band_r;          /* the (widened) band of the K3 task */
  uint8_t reverse;
  uint8_t reserved[3];
} smb_block_cand;

typedef struct {
  uint64_t nhits, ncand, nk2, nk2_band, nk3, nresults, ndiffbytes;
  uint64_t k2_cells, k2_cells_ref, k2_tasks_ref, k3_cells;   /* cells of all candidates / of those the reference scores */
  float ms_hits, ms_cand, ms_k2, ms_k3;                      /* device times (CUDA events on the context's stream) */
  int32_t launches;
  int32_t reserved;
  uint64_t ncigarbytes;                                      /* CIGAR text of all alignments (smb_block_params.cigar) */
} smb_block_sizes;

/* Runs the block on the seed tables of the last seed batch and leaves the outputs on the device;
 * `sizes` tells the caller how much room smb_block_fetch needs. */
int smb_block_run(smb_ctx *ctx, const smb_block_params *prm, const smb_block_job *jobs, int njobs,
		  const smb_block_ival *ivals, int nivals, smb_block_sizes *sizes);
/* reads[njobs]; cands[nk3], errs[nk3], first_result[nk3 + 1]; results[nresults]; diffstr[ndiffbytes]
 * (results / diffstr as smb_band_align_batch, task = index into cands). */
int smb_block_fetch(smb_ctx *ctx, smb_block_read *reads, smb_block_cand *cands, int32_t *errs,
		    uint32_t *first_result, smb_ali_result *results, uint8_t *diffstr);
/* smb_block_fetch plus the output stage (smb_block_params.cigar & SMB_CIGAR_ON).  `cigar_blob` receives, in one
 * copy, for the n = sizes.nresults alignments of `results`:
 *     uint32_t first[n + 1]; int32_t nm[n]; char text[sizes.ncigarbytes];      (SMB_CIGAR_BLOB_BYTES)
 * alignment i has the CIGAR text[first[i] .. first[i + 1]) (no terminator) with clip_start = qs, clip_end = read
 * length - 1 - qe (report.c:832-843 for either strand) and the edit distance nm[i]; nm[i] < 0 where the
 * reference's function fails on the string (-1 = ERRCODE_FAILURE, -59 = -ERRCODE_DIFFSTR; no text then). */
#define SMB_CIGAR_BLOB_BYTES(n, ntext) ((2 * (size_t) (n) + 1) * 4 + (size_t) (ntext))
#define SMB_CIGAR_FIRST(blob) ((const uint32_t *) (blob))
#define SMB_CIGAR_NM(blob, n) ((const int32_t *) (blob) + (size_t) (n) + 1)
#define SMB_CIGAR_TEXT(blob, n) ((const char *) (blob) + (2 * (size_t) (n) + 1) * 4)
int smb_block_fetch_cigar(smb_ctx *ctx, smb_block_read *reads, smb_block_cand *cands, int32_t *errs,
			  uint32_t *first_result, smb_ali_result *results, uint8_t *diffstr, void *cigar_blob);
/* The same stage for a batch of n alignment strings given explicitly: string i starts at diffstr[diff_off[i]]
 * (0-terminated); the blob has the layout above with room for max_text bytes of text.  Returns SMB_ERR_CAPACITY
 * with *ntext = required text size if max_text is too small.  (Shares device buffers with the stage of a block:
 * call it after smb_block_fetch_cigar, not between smb_block_run and the fetch.) */
int smb_cigar_batch(smb_ctx *ctx, const uint8_t *diffstr, size_t ndiffbytes, const uint32_t *diff_off,
		    const uint32_t *clip_start, const uint32_t *clip_end, int n, int flags,
		    void *cigar_blob, size_t max_text, size_t *ntext);
/* Test access: the candidate list of the last smb_block_run, all jobs, in scoring order
 * (cand_first[njobs + 1]); swscor = K2 / K2' score of every candidate (also the over-computed ones). */
int smb_block_debug_cands(smb_ctx *ctx, uint64_t *cand_first, smb_block_cand *cands, uint32_t *cover,
			  uint32_t *qs_qe, size_t max_cands);

#ifdef __cplusplus
}
#endif
#endif /* SMALT_B200_H */
