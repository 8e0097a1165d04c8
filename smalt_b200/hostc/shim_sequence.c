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
static void shim_encode_scalar(char *dst, const char *src, size_t len, const UCHAR_t *codtab)
{
  size_t i;
  for (i = 0; i < len; i++) dst[i] = (char) codtab[(unsigned char) src[i]];
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define SHIM_SIMD 1
/* what the SIMD paths assume of a codec, checked once per codec and thread: the code of A, C, G, T in either
 * case is (x ^ x >> 1) & 3 of bits 1, 2 of the letter plus (upper-case letter - 64) << 3 (make3BitMangledCodec,
 * sequence.c:287-318), every code the encoder can produce decodes to 64 + (code >> 3), and the complement
 * of a standard code c is codtab_complement[c & 3].  Anything else: the table loops. */
typedef struct {
  const SeqCodec *codec;
  int ok;
  char rc4[16];     /* letter of the complement of the standard code v (entries 0..3) */
} ShimSimd;
static __thread ShimSimd t_simd;

static const ShimSimd *shim_simd(const SeqCodec *codep)
{
  if (t_simd.codec != codep) {
    int i, ok = __builtin_cpu_supports("ssse3") != 0;
    static const char acgt[] = "ACGTacgt";
    for (i = 0; i < 8 && ok; i++) {
      const unsigned c = (unsigned char) acgt[i], up = c & 0xDFu;
      const unsigned code = (((up >> 1) ^ (up >> 2)) & 3u) | ((up - 64u) << 3);
      if (codep->codtab[c] != code) ok = 0;
    }
    for (i = 1; i < SIZE_CODTAB && ok; i++) {
      const unsigned code = codep->codtab[i];
      if ((unsigned char) codep->decodtab[code] != 64u + (code >> 3)) ok = 0;
    }
    memset(t_simd.rc4, 0, sizeof(t_simd.rc4));
    for (i = 0; i < 4; i++) t_simd.rc4[i] = codep->decodtab[codep->codtab_complement[i]];
    t_simd.ok = ok;
    t_simd.codec = codep;
  }
  return &t_simd;
}

__attribute__((target("ssse3")))
static void shim_encode_simd(char *dst, const char *src, size_t len, const UCHAR_t *codtab)
{
  const __m128i up_mask = _mm_set1_epi8((char) 0xDF), three = _mm_set1_epi8(3), c64 = _mm_set1_epi8(64);
  /* ACGT test: the letter must equal the table entry selected by its own low nibble ('A' 1, 'C' 3, 'D'..: 4 'T', 7 'G') */
  const __m128i letters = _mm_setr_epi8(0, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0);
  size_t i = 0;
  for (; i + 16 <= len; i += 16) {
    const __m128i x = _mm_loadu_si128((const __m128i *) (src + i));
    const __m128i up = _mm_and_si128(x, up_mask);
    const __m128i good = _mm_cmpeq_epi8(_mm_shuffle_epi8(letters, _mm_and_si128(up, _mm_set1_epi8(15))), up);
    if (_mm_movemask_epi8(good) != 0xffff) {
      shim_encode_scalar(dst + i, src + i, 16, codtab);
      continue;
    }
    {
      const __m128i h1 = _mm_and_si128(_mm_srli_epi16(up, 1), _mm_set1_epi8(0x7f));
      const __m128i h2 = _mm_and_si128(_mm_srli_epi16(up, 2), _mm_set1_epi8(0x3f));
      const __m128i a = _mm_and_si128(_mm_xor_si128(h1, h2), three);
      const __m128i offs = _mm_sub_epi8(up, c64);                                   /* 1 .. 20 */
      const __m128i hi = _mm_and_si128(_mm_slli_epi16(offs, 3), _mm_set1_epi8((char) 0xF8));
      _mm_storeu_si128((__m128i *) (dst + i), _mm_or_si128(a, hi));
    }
  }
  if (i < len) shim_encode_scalar(dst + i, src + i, len - i, codtab);
}

/* decoded letters of len codes; reverse: of the reverse complement (source read backwards) */
__attribute__((target("ssse3")))
static void shim_decode_simd(char *dst, const unsigned char *src, size_t len, int reverse, const ShimSimd *sd,
			     const char *fwd_tab, const char *rc_tab)
{
  const __m128i c64 = _mm_set1_epi8(64), m1f = _mm_set1_epi8(0x1f), three = _mm_set1_epi8(3), four = _mm_set1_epi8(4);
  const __m128i rev = _mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  const __m128i rc4 = _mm_loadu_si128((const __m128i *) sd->rc4);
  size_t i = 0;
  if (!reverse) {
    for (; i + 16 <= len; i += 16) {
      const __m128i x = _mm_loadu_si128((const __m128i *) (src + i));
      _mm_storeu_si128((__m128i *) (dst + i), _mm_add_epi8(_mm_and_si128(_mm_srli_epi16(x, 3), m1f), c64));
    }
    for (; i < len; i++) dst[i] = fwd_tab[src[i]];
  } else {
    for (; i + 16 <= len; i += 16) {
      const __m128i x = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *) (src + len - 16 - i)), rev);
      const __m128i own = _mm_add_epi8(_mm_and_si128(_mm_srli_epi16(x, 3), m1f), c64);
      const __m128i comp = _mm_shuffle_epi8(rc4, _mm_and_si128(x, three));
      const __m128i nonstd = _mm_cmpeq_epi8(_mm_and_si128(x, four), four);
      _mm_storeu_si128((__m128i *) (dst + i), _mm_or_si128(_mm_and_si128(nonstd, own), _mm_andnot_si128(nonstd, comp)));
    }
    for (; i < len; i++) dst[i] = rc_tab[src[len - 1 - i]];
  }
}

__attribute__((target("ssse3")))
static void shim_reverse_bytes(char *dst, const char *src, size_t len)
{
  const __m128i rev = _mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  size_t i = 0;
  for (; i + 16 <= len; i += 16)
    _mm_storeu_si128((__m128i *) (dst + i), _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *) (src + len - 16 - i)), rev));
  for (; i < len; i++) dst[i] = src[len - 1 - i];
}
#endif

static void shim_encode_bytes(char *dst, const char *src, size_t len, const SeqCodec *codep)
{
#ifdef SHIM_SIMD
  if (shim_simd(codep)->ok) { shim_encode_simd(dst, src, len, codep->codtab); return; }
#endif
  shim_encode_scalar(dst, src, len, codep->codtab);
}

static int shim_load_encoded(SEQSEQ *sp, const char *s, size_t len, const SeqCodec *codep)
{
  if (len + 2 >= sp->alloc_size && reallocSeqBlocks(sp, len + 2)) return ERRCODE_NOMEM;
  shim_encode_bytes(sp->basep, s, len, codep);
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
#ifdef SHIM_SIMD
  if (shim_simd(codep)->ok) {
    shim_decode_simd(seq, cp, len, reverse, shim_simd(codep), codep->decodtab, reverse ? shim_rc_table(codep) : NULL);
  } else
#endif
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
    if (!reverse) memcpy(qual, qc, len);
#ifdef SHIM_SIMD
    else if (shim_simd(codep)->ok) shim_reverse_bytes(qual, qc, len);
#endif
    else for (i = 0; i < len; i++) qual[i] = qc[len - 1 - i];
    qual[len] = '\0';
  } else {
    qual[0] = '\0';
  }
  return ERRCODE_SUCCESS;
}
