/* shim_sequence.c - sequence objects for the smalt_b200 driver build.
 *
 * The reference's sequence.c compiled in place (read-only tree on the include path, nothing is
 * copied) plus ONE function: loading a SeqFastq from the four line segments of a FASTQ
 * record with memcpy.  The public seqFastqSetAscii (sequence.c:1860) does the same through
 * setSeq (:780-803), a per-character loop with a reallocation test per character - 0.6 us per
 * 150 bp read, a visible share of the host time once the hot path runs on the GPU.  The
 * resulting object state (strings, sizes, codes, type) is identical.
 */
#include "sequence.c"

static int shim_load(SEQSEQ *sp, const char *s, size_t len)
{
  if (len + 2 >= sp->alloc_size && reallocSeqBlocks(sp, len + 2)) return ERRCODE_NOMEM;
  memcpy(sp->basep, s, len);
  sp->basep[len] = '\0';
  sp->size = (SETSIZ_t) len;
  sp->code = SEQCOD_ASCII;
  sp->nbit_symb = NBITS_PER_BYTE;
  return ERRCODE_SUCCESS;
}

/* name / qual name: already white-space trimmed; seq / qual: no white space, equal lengths > 0
 * (the caller has checked all of this - everything else goes through the reference parser) */
int smbShimSeqFastqLoad(SeqFastq *sqp, const char *name, size_t nlen, const char *seq, size_t slen,
			const char *qnam, size_t qnlen, const char *qual, size_t qlen)
{
  int errcode;
  if (slen >= SEQ_MAXLEN || nlen >= SEQ_MAXLEN || slen != qlen || slen < 1) return ERRCODE_SEQLEN;
  if ((errcode = shim_load(sqp->headp, name, nlen)) || (errcode = shim_load(sqp->datap, seq, slen)))
    return errcode;
  if (!(sqp->qheadp || (sqp->qheadp = createSeq(BLOCKSIZE_HEADER)))) return ERRCODE_NOMEM;
  if ((errcode = shim_load(sqp->qheadp, qnam, qnlen))) return errcode;
  if (!(sqp->qualp || (sqp->qualp = createSeq(sqp->datap->block_size)))) return ERRCODE_NOMEM;
  if ((errcode = shim_load(sqp->qualp, qual, qlen))) return errcode;
  sqp->type = SEQTYP_FASTQ;
  return ERRCODE_SUCCESS;
}

/* Decoded (ASCII) bases and quality characters of read segment [start, start+len) in one pass,
 * as fprintREPALIsam obtains them with seqFastqAppendSegment (reverse complement for reads
 * mapped to the reverse strand, appendSeqSegment sequence.c:878-892) + seqFastqDecode
 * (decodeSeq :1552-1563).  seq/qual get len characters and a terminating 0; *has_qual = 0 when
 * the read carries no qualities.  Returns ERRCODE_SEQCODE if the read is not in the mangled
 * encoding (the caller then uses the reference's own functions). */
int smbShimSeqFastqDecodeSegment(char *seq, char *qual, int *has_qual, const SeqFastq *sqp,
				 SEQLEN_t start, SEQLEN_t len, int reverse, const SeqCodec *codep)
{
  const SEQSEQ *dp = sqp->datap, *qp = sqp->qualp;
  const unsigned char *cp;
  SEQLEN_t i;
  if (dp->code != SEQCOD_MANGLED) return ERRCODE_SEQCODE;
  if (start > dp->size || start + len > dp->size) return ERRCODE_ARGRANGE;
  cp = (const unsigned char *) dp->basep + start;
  if (reverse) {
    for (i = 0; i < len; i++) {
      const unsigned char c = cp[len - 1 - i];
      seq[i] = (char) codep->decodtab[(c & SEQCOD_STDNT_TESTBIT) ? c :
				      (unsigned char) codep->codtab_complement[c & SEQCOD_STDNT_MASK]];
    }
  } else {
    for (i = 0; i < len; i++) seq[i] = (char) codep->decodtab[cp[i]];
  }
  seq[len] = '\0';
  /* seqFastqAppendSegment (sequence.c:1934-1949) copies qualities whenever there are any */
  *has_qual = qp != NULL && qp->size >= 1;
  if (*has_qual && qp->size != dp->size) return ERRCODE_QUALLEN;
  if (*has_qual) {
    const char *qc = qp->basep + start;
    if (reverse) for (i = 0; i < len; i++) qual[i] = qc[len - 1 - i];
    else memcpy(qual, qc, len);
    qual[len] = '\0';
  } else {
    qual[0] = '\0';
  }
  return ERRCODE_SUCCESS;
}
