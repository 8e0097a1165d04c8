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

/* FASTQ record -> SeqFastq by memcpy (shim_sequence.c) */
int smbShimSeqFastqLoad(SeqFastq *sqp, const char *name, size_t nlen, const char *seq, size_t slen,
			const char *qnam, size_t qnlen, const char *qual, size_t qlen);

/* per-worker report writers of the block-parallel driver (shim_report.c) */
#include "report.h"
ReportWriter *smbShimReportWriterClone(const ReportWriter *proto);
void smbShimReportWriterSetStream(ReportWriter *p, FILE *fp);
void smbShimReportWriterDelete(ReportWriter *p);
/* reportWrite for single-end SAM records with the read decoded once (falls back to reportWrite) */
int smbShimReportWriteSAM(const ReportWriter *wrp, const SeqFastq *readp, const SeqSet *ssp,
			  const SeqCodec *codecp, const Report *rep);
int smbShimWriteSAMHeader(FILE *fp, const SeqSet *ssp, const char *prognam, const char *progversion,
			  int narg, char * const *argv);
#endif
