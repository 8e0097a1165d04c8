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

/* the bases of a record encoded while they are loaded: what seqFastqEncode (encodeSeq, sequence.c:1337-1357)
 * makes of the ASCII string - every character through codtab (the caller has made sure that there are
 * neither 0 characters nor characters beyond 0x7f, which the reference looks up with a negative index) */
static void shim_encode_bytes(char *dst, const char *src, size_t len, const UCHAR_t *codtab)
{
  size_t i;
  for (i = 0; i < len; i++) dst[i] = (char) codtab[(unsigned char) src[i]];
}

static int shim_load_encoded(SEQSEQ *sp, const char *s, size_t len, const SeqCodec *codep)
{
  if (len + 2 >= sp->alloc_size && reallocSeqBlocks(sp, len + 2)) return ERRCODE_NOMEM;
  shim_encode_bytes(sp->basep, s, len, codep->codtab);
  sp->basep[len] = '\0';
  sp->size = (SETSIZ_t) len;
  sp->code = SEQCOD_MANGLED;
  sp->nbit_symb = NBITS_PER_BYTE;
  return ERRCODE_SUCCESS;
}

/* name / qual name: already white-space trimmed; seq / qual: no white space, equal lengths > 0
 * (the caller has checked all of this - everything else goes through the reference parser).
 * All four strings of the object are set (no seqFastqBlank needed before).  With a codec the
 * bases are stored encoded (SEQCOD_MANGLED), as after seqFastqEncode. */
int smbShimSeqFastqLoad(SeqFastq *sqp, const char *name, size_t nlen, const char *seq, size_t slen,
			const char *qnam, size_t qnlen, const char *qual, size_t qlen, const SeqCodec *codep)
{
  int errcode;
  if (slen >= SEQ_MAXLEN || nlen >= SEQ_MAXLEN || slen != qlen || slen < 1) return ERRCODE_SEQLEN;
  if ((errcode = shim_load(sqp->headp, name, nlen)) ||
      (errcode = codep ? shim_load_encoded(sqp->datap, seq, slen, codep) : shim_load(sqp->datap, seq, slen)))
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
/* code -> character of the reverse strand for all 256 codes of a codec (built once per codec and thread) */
static const char *shim_rc_table(const SeqCodec *codep)
{
  static __thread const SeqCodec *t_codec;
  static __thread char t_tab[256];
  if (t_codec != codep) {
    int c;
    for (c = 0; c < 256; c++)
      t_tab[c] = codep->decodtab[(c & SEQCOD_STDNT_TESTBIT) ? c :
				 (unsigned char) codep->codtab_complement[c & SEQCOD_STDNT_MASK]];
    t_codec = codep;
  }
  return t_tab;
}

int smbShimSeqFastqDecodeSegment(char *seq, char *qual, int *has_qual, const SeqFastq *sqp,
				 SEQLEN_t start, SEQLEN_t len, int reverse, const SeqCodec *codep)
{
  const SEQSEQ *dp = sqp->datap, *qp = sqp->qualp;
  const unsigned char *cp;
  SEQLEN_t i;
  if (dp->code != SEQCOD_MANGLED) return ERRCODE_SEQCODE;
  if (start > dp->size || start + len > dp->size) return ERRCODE_ARGRANGE;
  if (SIZE_DECODTAB < 256) return ERRCODE_ASSERT;
  cp = (const unsigned char *) dp->basep + start;
  if (reverse) {
    const char *rct = shim_rc_table(codep);
    const unsigned char *ep = cp + len - 1;
    for (i = 0; i < len; i++) seq[i] = rct[ep[-(ptrdiff_t) i]];
  } else {
    const char *dt = codep->decodtab;
    for (i = 0; i < len; i++) seq[i] = dt[cp[i]];
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
