/* shim.h - private interface between the hot-path shim (shim_hot.c) and the wave
 * orchestrator (rmap_wave.c) of the smalt_b200 driver build. */
#ifndef SMALT_B200_SHIM_H
#define SMALT_B200_SHIM_H
#include "../../include/smalt_b200.h"
#include "hashhit.h"
#include "alignment.h"
#include "score.h"

/* uploads index + packed reference to the process-wide root context (once) */
int smbShimInit(const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
		const ScoreMatrix *scormtxp);
smb_ctx *smbShimRootCtx(void);
int smbShimDevice(void);
void smbShimForgetIndex(void);
/* a per-thread context (own stream and scratch) that shares the root's index/reference */
int smbShimWorkerCtx(smb_ctx **ctxp, const ScoreMatrix *scormtxp);
int smbShimSetScoring(smb_ctx *ctx, const ScoreProfile *profp);

/* injectors: serve GPU results of a whole block through the reference's opaque types */
void smbShimHitInfoSet(HashHitInfo *p, const smb_seed_info *info);
const smb_seed_info *smbShimHitInfoGet(const HashHitInfo *p);
int smbShimHitListSet(HashHitList *p, const uint64_t *sqdat, int nhits, int is_reverse, uint32_t qlen,
		      unsigned char ktup, unsigned char nskip);
int smbShimAliRsltSetAdd(AliRsltSet *p, int score, int qs, int qe, int rs, int re,
			 const unsigned char *diffstr, int difflen);

/* arrays of a hash table (shim_hashidx.c) */
void smbShimHashTableArrays(const HashTable *htp, int *typ, int *wordlen, int *nskip,
			    int *nbits_key, int *nbits_lo, uint32_t *npos, uint32_t *nwords,
			    const uint32_t **idx, const uint32_t **pos,
			    const uint32_t **wordidx, const uint32_t **posidx);

/* fiber scheduler (shim_fiber.inc.c): runs a per-item function that uses the reference's
 * one-call hot-path API for many items at once and executes the calls as GPU batches */
typedef struct SmbFiberPool_ SmbFiberPool;
typedef void (SMBFIBER_ITEMF)(void *user, int item, int slot);
typedef struct {
  uint64_t n_items, n_waves, n_seeded, seeds_served, n_hits, n_sw, n_bandfast, n_bandali, order_waits;
  uint64_t cells_k2, cells_k3;
  double ms_k1, ms_k2, ms_k3;
  double wall_stage, wall_arena, wall_sw, wall_ba;
  double wall_host, wall_hits, wall_dp, cpu_host;   /* seconds: fibers running, hit-list batches, DP batches */
} smbFiberStats;
SmbFiberPool *smbFiberPoolCreate(int nfibers, size_t stack_bytes);
void smbFiberPoolDelete(SmbFiberPool *p);
int smbFiberPoolSize(const SmbFiberPool *p);
void smbFiberPoolGetStats(const SmbFiberPool *p, smbFiberStats *st);
int smbFiberPoolSeed(SmbFiberPool *p, int nreads, SeqFastq *const *reads, int reads_per_item, int is_short,
		     uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq, const HashTable *htp);
int smbFiberPoolRun(SmbFiberPool *p, int nitems, SMBFIBER_ITEMF *itemf, void *user);
int smbFiberYield(void);
void smbFiberWaitOrder(void);
int smbFiberSelfTest(int nfibers, int nitems, int *order_out, int *draws_out, int *ndraws);

/* FASTQ record -> SeqFastq by memcpy (shim_sequence.c) */
int smbShimSeqFastqLoad(SeqFastq *sqp, const char *name, size_t nlen, const char *seq, size_t slen,
			const char *qnam, size_t qnlen, const char *qual, size_t qlen, const SeqCodec *codep);

/* per-worker report writers of the block-parallel driver (shim_report.c) */
#include "report.h"
ReportWriter *smbShimReportWriterClone(const ReportWriter *proto);
void smbShimReportWriterSetStream(ReportWriter *p, FILE *fp);
void smbShimReportWriterDelete(ReportWriter *p);
/* reportWrite for single-end SAM records with the read decoded once (falls back to reportWrite) */
int smbShimReportWriteSAM(const ReportWriter *wrp, const SeqFastq *readp, const SeqSet *ssp,
			  const SeqCodec *codecp, const Report *rep);
/* CIGAR text + NM of the alignments of the read that is being reported, computed by the device's output stage
 * (csrc/cigar.cu); set per read by the wave orchestrator around the emit call (thread-local, NULL clears) */
typedef struct {
  const smb_block_cand *cands;       /* aligned candidates of the read [nk3] */
  const uint32_t *res_first;         /* their alignments in res: res_first[t] .. res_first[t + 1] */
  uint32_t nk3, qlen;
  const smb_ali_result *res;
  const uint8_t *diff;
  const uint32_t *cig_first;
  const int32_t *cig_nm;
  const char *cig_text;
} SmbCigarSource;
void smbShimSetCigarSource(const SmbCigarSource *src);
/* SMB_CIGAR_* flags that make the device's text equal what this writer prints (0: not a plain SAM writer) */
int smbShimReportCigarFlags(const ReportWriter *wrp);
void smbShimCigarCounters(unsigned long long *ndev, unsigned long long *nhost);
void smbShimCigarFlush(void);
int smbShimWriteSAMHeader(FILE *fp, const SeqSet *ssp, const char *prognam, const char *progversion,
			  int narg, char * const *argv);
#endif
