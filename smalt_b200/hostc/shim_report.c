/* shim_report.c - output formatting object for the smalt_b200 driver build.
 *
 * SAM/CIGAR/... formatting is NOT on the hot path and stays the reference's code: this
 * translation unit compiles the reference's report.c in place (read-only tree on the include
 * path, nothing is copied).  The reference formats every record in its single OUTPUT thread
 * straight into the output file (reportWrite -> fprintf on ReportWriter.oufp, report.c:156-173,
 * :1486-); the block-parallel driver (fastmap.inc.c) instead lets every worker thread format
 * the records of its block of reads into a memory stream and writes the blocks in input
 * order.  That needs a ReportWriter per worker whose stream can be pointed at a buffer - the
 * three small functions below; the formatting itself is the reference's, byte for byte.
 */
#include "report.c"
#include "shim.h"

/* a writer with the format settings of `proto` and private scratch buffers, no stream yet */
ReportWriter *smbShimReportWriterClone(const ReportWriter *proto)
{
  ReportWriter *p;
  EMALLOCP0(p);
  if (!p) return NULL;
  p->oufmt = proto->oufmt;
  p->modflg = proto->modflg;
  p->linwidth = proto->linwidth;
  memcpy(p->namext, proto->namext, sizeof(p->namext));
  memcpy(p->namext_mate, proto->namext_mate, sizeof(p->namext_mate));
  p->filnam = NULL;
  p->oufp = NULL;
  p->dfblkp = NULL;
  p->qbufp = seqFastqCreate(0, SEQTYP_FASTQ);
  p->sbufp = seqFastqCreate(0, SEQTYP_FASTA);
  p->nambufp = createREPNAMBUF();
  if (!p->qbufp || !p->sbufp || !p->nambufp) {
    smbShimReportWriterDelete(p);
    return NULL;
  }
  return p;
}

void smbShimReportWriterSetStream(ReportWriter *p, FILE *fp) { p->oufp = fp; }

void smbShimReportWriterDelete(ReportWriter *p)
{
  if (!p) return;
  p->oufp = NULL; /* the stream belongs to the caller */
  reportDeleteWriter(p);
}

int smbShimWriteSAMHeader(FILE *fp, const SeqSet *ssp, const char *prognam, const char *progversion,
			  int narg, char * const *argv)
{
  return writeSAMHeaderf(fp, ssp, prognam, progversion, narg, argv);
}
