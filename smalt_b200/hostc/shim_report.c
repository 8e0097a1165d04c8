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

/* ------------------------------------------------------------------------------------
 * Single-end SAM records without the per-character copies of fprintREPALIsam.
 *
 * reportWrite (report.c:1758) -> writeReportForRead -> writeREPALI -> fprintREPALIsam
 * (:763-905) copies the read into a scratch SeqFastq (seqFastqAppendSegment), decodes it in
 * place (seqFastqDecode) and prints it with "%s".  smbShimReportWriteSAM prints the same
 * record with the read decoded once, straight into a thread-local line buffer
 * (smbShimSeqFastqDecodeSegment).  It covers what the block-parallel driver emits for
 * single-end reads - SAM format, no mate, no explicit alignment output; for anything else,
 * and whenever a precondition does not hold, it calls the reference's reportWrite.
 * ------------------------------------------------------------------------------------ */
extern int smbShimSeqFastqDecodeSegment(char *seq, char *qual, int *has_qual, const SeqFastq *sqp,
					SEQLEN_t start, SEQLEN_t len, int reverse, const SeqCodec *codep);

/* length of a name as copyReadNamStrToREPSTR (report.c:434-461) copies it: up to the first white space,
 * without the "/1", "/2" mate extension if is_stripped */
static size_t samNameLen(const char *namp, int is_stripped)
{
  size_t i = 0;
  for (;; i++) {
    const unsigned char c = (unsigned char) namp[i];
    if (!c || c == ' ' || (c >= 9 && c <= 13)) break;
  }
  if (is_stripped && i > 2 && namp[i - 2] == OUFMT_NAMSTR_MATESEP &&
      (namp[i - 1] == OUFMT_NAMSTR_MATE1 || namp[i - 1] == OUFMT_NAMSTR_MATE2))
    i -= 2;
  return i;
}

/* ---- the device's output stage as the source of CIGAR and NM ---- */
static __thread const SmbCigarSource *t_cigp;   /* (the caller's struct stays valid until it clears the pointer) */
static unsigned long long g_cig_dev, g_cig_host;   /* records served from the device / formatted on the host */
static __thread unsigned long long t_cig_dev, t_cig_host;   /* ... of this thread since its last flush */

/* adds the calling thread's counts to the process totals (once per block, not per record: the totals share a line) */
void smbShimCigarFlush(void)
{
  if (t_cig_dev) { __atomic_fetch_add(&g_cig_dev, t_cig_dev, __ATOMIC_RELAXED); t_cig_dev = 0; }
  if (t_cig_host) { __atomic_fetch_add(&g_cig_host, t_cig_host, __ATOMIC_RELAXED); t_cig_host = 0; }
}

void smbShimSetCigarSource(const SmbCigarSource *src) { t_cigp = src; }

int smbShimReportCigarFlags(const ReportWriter *wrp)
{
  if (wrp->oufmt != REPORTFMT_SAM || (wrp->modflg & REPORTMODIF_ALIOUT)) return 0;
  return SMB_CIGAR_ON | ((wrp->modflg & REPORTMODIF_SOFTCLIP) ? SMB_CIGAR_SOFTCLIP : 0) |
    ((wrp->modflg & REPORTMODIF_XMISMATCH) ? SMB_CIGAR_XMISMATCH : 0);
}

void smbShimCigarCounters(unsigned long long *ndev, unsigned long long *nhost)
{
  *ndev = __atomic_load_n(&g_cig_dev, __ATOMIC_RELAXED);
  *nhost = __atomic_load_n(&g_cig_host, __ATOMIC_RELAXED);
}

/* The alignment of the device's result list that the report entry `rrp` was made from: same candidate
 * (sequence, strand), same coordinates (results.c:1897-1907) and score, and the same alignment string bytes
 * (report.c:1700 copies them unchanged).  -> index into res / cig_first / cig_nm, or -1. */
static long cigarLookup(const SmbCigarSource *cs, const REPALI *rrp, const DIFFSTR_T *diffstr, SEQLEN_t qlen)
{
  uint32_t t, i;
  const int is_rev = (rrp->status & REPMATEFLG_REVERSE) != 0;
  if (!cs || cs->qlen != (uint32_t) qlen) return -1;
  for (t = 0; t < cs->nk3; t++) {
    const smb_block_cand *cp = cs->cands + t;
    if ((SEQNUM_t) cp->sqidx != rrp->s_idx || (cp->reverse != 0) != is_rev) continue;
    for (i = cs->res_first[t]; i < cs->res_first[t + 1]; i++) {
      const smb_ali_result *r = cs->res + i;
      const long long qs1 = is_rev ? (long long) qlen - r->qe : (long long) r->qs + 1;
      const long long qe1 = is_rev ? (long long) qlen - r->qs : (long long) r->qe + 1;
      if (r->score != rrp->swatscor || qs1 != (long long) rrp->q_start || qe1 != (long long) rrp->q_end ||
	  (long long) cp->rs + r->rs + 1 != (long long) rrp->s_start ||
	  (long long) cp->rs + r->re + 1 != (long long) rrp->s_end)
	continue;
      if (r->diff_len && !memcmp(cs->diff + r->diff_off, diffstr, r->diff_len)) return (long) i;
    }
  }
  return -1;
}

static __thread char *t_seqbuf;
static __thread size_t t_seqbuf_alloc;

static int samRecordSingle(const ReportWriter *wrp, const REPALI *rrp, const DiffStr *rdfsp,
			   const SeqFastq *q_sqp, const SeqSet *ssp, const SeqCodec *codecp)
{
  int errcode = ERRCODE_SUCCESS;
  FILE *fp = wrp->oufp;
  const REPMODIFLG_t oumodiflg = wrp->modflg;
  const BOOL_t is_mapped = (BOOL_t) ((rrp->status & REPMATEFLG_MAPPED) != 0);
  const char *s_nam = OUFMT_SAM_NULLSTR, *seqstr, *qualstr;
  const DIFFSTR_T *diffstr = NULL;
  int editdist = 0, clip_start = 0, clip_end = 0, swatscor = 0, has_qual = 0;
  SAMFLAG_t samflg = 0;
  SEQLEN_t qlen, pos = 0, qseg_start = 0, qseg_len = 0;
  char cod, *seqbuf, *qualbuf;
  BOOL_t isReverse = 0, want_seq;

  if (is_mapped) {
    seqSetGetSeqDatByIndex(NULL, &s_nam, rrp->s_idx, ssp);
    diffstr = rdfsp->dstrp + rrp->dfo;
  }
  seqFastqGetConstSequence(q_sqp, &qlen, &cod);

  if (is_mapped) {
    isReverse = (BOOL_t) ((rrp->status & REPMATEFLG_REVERSE) ? 1 : 0);
    if (oumodiflg & REPORTMODIF_SOFTCLIP) { qseg_start = 0; qseg_len = qlen; }
    else { qseg_start = rrp->q_start - 1; qseg_len = rrp->q_end - rrp->q_start + 1; }
    want_seq = 1;
    pos = (SEQLEN_t) rrp->s_start;
    if (rrp->q_end > qlen) return ERRCODE_ASSERT;
    if (isReverse) {
      samflg |= SAMFLAG_STRAND;
      clip_start = qlen - rrp->q_end;
      clip_end = rrp->q_start - 1;
    } else {
      clip_start = rrp->q_start - 1;
      clip_end = qlen - rrp->q_end;
    }
    if (rrp->status & REPMATEFLG_PARTIAL) samflg |= SAMFLAG_NOTPRIMARY;
    swatscor = rrp->swatscor;
  } else {
    want_seq = (BOOL_t) ((oumodiflg & REPORTMODIF_SOFTCLIP) != 0);
    qseg_start = 0;
    qseg_len = qlen;
    samflg |= SAMFLAG_NOMAP;
  }
  if (want_seq) {
    /* appendSeqSegment: a zero length means "to the end of the sequence" */
    if (!qseg_len || qseg_start + qseg_len > qlen) qseg_len = qlen - qseg_start;
    if (2 * ((size_t) qseg_len + 1) > t_seqbuf_alloc) {
      char *hp = (char *) realloc(t_seqbuf, 4 * ((size_t) qseg_len + 1));
      if (!hp) return ERRCODE_NOMEM;
      t_seqbuf = hp;
      t_seqbuf_alloc = 4 * ((size_t) qseg_len + 1);
    }
    seqbuf = t_seqbuf;
    qualbuf = t_seqbuf + qseg_len + 1;
    if ((errcode = smbShimSeqFastqDecodeSegment(seqbuf, qualbuf, &has_qual, q_sqp, qseg_start, qseg_len,
						is_mapped && isReverse, codecp)))
      return errcode;
    seqstr = seqbuf;
    qualstr = has_qual ? qualbuf : OUFMT_SAM_NULLSTR;
  } else {
    seqstr = OUFMT_SAM_NULLSTR;
    qualstr = OUFMT_SAM_NULLSTR;
  }
  if (!qualstr[0]) qualstr = OUFMT_SAM_NULLSTR;

  /* OUFMT_SAM_BEFORE "%s\t%hu\t%s\t%i\t%hi\t" (report.c:192), written field by field when the stream
   * is this thread's capture buffer */
  {
    /* the names straight from their strings (the reference copies them into REPSTR buffers first) */
    const char *qn = q_sqp ? seqFastqGetSeqName(q_sqp) : OUFMT_SAM_NULLSTR, *rn = is_mapped ? s_nam : OUFMT_SAM_NULLSTR;
    const size_t lq = samNameLen(qn, 1), lr = samNameLen(rn, 0);
    char *o = smbFastReserve(fp, lq + lr + 64);
    if (o) {
      char *p = o;
      memcpy(p, qn, lq); p += lq; *p++ = '\t';
      p += smbFastPutInt(p, (unsigned short) samflg); *p++ = '\t';
      memcpy(p, rn, lr); p += lr; *p++ = '\t';
      p += smbFastPutInt(p, (int) pos); *p++ = '\t';
      p += smbFastPutInt(p, (short) rrp->mapscor); *p++ = '\t';
      smbFastCommit((size_t) (p - o));
    } else
      fprintf(fp, "%.*s\t%hu\t%.*s\t%i\t%hi\t", (int) lq, qn, samflg, (int) lr, rn, pos, rrp->mapscor);
  }
  if (is_mapped) {
    /* CIGAR and NM from the device's output stage; the reference's functions where the record does not come
     * from a resident block (other drivers of this writer) or the stage reported the reference's error */
    const SmbCigarSource *cs = t_cigp;
    const long ci = cigarLookup(cs, rrp, diffstr, qlen);
    if (ci >= 0 && cs->cig_nm[ci] >= 0) {
      const size_t lc = cs->cig_first[ci + 1] - cs->cig_first[ci];
      char *o = smbFastReserve(fp, lc + 8);
      if (o) {
	memcpy(o, cs->cig_text + cs->cig_first[ci], lc);
	smbFastCommit(lc);
      } else
	fwrite(cs->cig_text + cs->cig_first[ci], 1, lc, fp);
      editdist = cs->cig_nm[ci];
      t_cig_dev++;
    } else {
      errcode = diffStrPrintf(fp, diffstr,
			      (char) ((oumodiflg & REPORTMODIF_XMISMATCH) ? DIFFSTRFORM_CIGEXT_XMISMATCH : DIFFSTRFORM_CIGEXT),
			      clip_start, clip_end, (char) ((oumodiflg & REPORTMODIF_SOFTCLIP) != 0));
      if (!errcode) editdist = diffStrGetLevenshteinDistance(diffstr);
      t_cig_host++;
    }
  } else {
    fprintf(fp, OUFMT_SAM_NULLSTR);
  }
  /* OUFMT_SAM_AFTER "\t%s\t%i\t%i\t%s\t%s\tNM:i:%i\tAS:i:%i\n" (report.c:194) */
  {
    const size_t ls = (want_seq && qseg_len > 0) ? (size_t) qseg_len : strlen(seqstr);
    const size_t lq = (want_seq && has_qual && qseg_len > 0) ? (size_t) qseg_len : strlen(qualstr);
    char *o = smbFastReserve(fp, ls + lq + 96);
    if (o) {
      char *p = o;
      memcpy(p, "\t*\t0\t0\t", 7); p += 7;
      memcpy(p, seqstr, ls); p += ls; *p++ = '\t';
      memcpy(p, qualstr, lq); p += lq;
      memcpy(p, "\tNM:i:", 6); p += 6;
      p += smbFastPutInt(p, editdist);
      memcpy(p, "\tAS:i:", 6); p += 6;
      p += smbFastPutInt(p, swatscor);
      *p++ = '\n';
      smbFastCommit((size_t) (p - o));
    } else
      fprintf(fp, OUFMT_SAM_AFTER, OUFMT_SAM_NULLSTR, 0, 0, seqstr, qualstr, editdist, swatscor);
  }
  return errcode;
}

int smbShimReportWriteSAM(const ReportWriter *wrp, const SeqFastq *readp, const SeqSet *ssp,
			  const SeqCodec *codecp, const Report *rep)
{
  int errcode = ERRCODE_SUCCESS, n;
  const int na = ARRLEN(rep->arAr);
  char cod;
  seqFastqGetConstSequence(readp, NULL, &cod);
  if (wrp->oufmt != REPORTFMT_SAM || (wrp->modflg & REPORTMODIF_ALIOUT) || ARRLEN(rep->arBr) > 0 ||
      ARRLEN(rep->pairr) > 0 || cod != SEQCOD_MANGLED)
    return reportWrite(wrp, readp, NULL, ssp, codecp, rep);
  for (n = 0; n < na; n++)
    if (rep->arAr[n].status & REPMATEFLG_PAIRED)
      return reportWrite(wrp, readp, NULL, ssp, codecp, rep);
  for (n = 0; n < na && !errcode; n++) {
    rep->arAr[n].was_output = 0;
    errcode = samRecordSingle(wrp, rep->arAr + n, &rep->dfs, readp, ssp, codecp);
  }
  return errcode;
}
