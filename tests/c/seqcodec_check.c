/* The SIMD paths of smalt_b200/hostc/shim_sequence.c against the reference's own functions:
 *   smbShimSeqFastqLoad with a codec          == seqFastqSetAscii + seqFastqEncode (sequence.c:1860, :2145)
 *   smbShimSeqFastqDecodeSegment fwd / reverse == seqFastqAppendSegment + seqFastqDecode (:1919, :2150)
 * on random reads of every length 1..200 over ACGT, lower case, N, IUPAC letters and every printable character. */
#include "shim_sequence.c"

static unsigned long long rng_state = 88172645463325252ULL;
static unsigned rnd(void)
{
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (unsigned) (rng_state >> 11);
}

int main(void)
{
  SeqCodec *codep = seqCodecCreate();
  SeqFastq *a = seqFastqCreate(0, SEQTYP_FASTQ), *b = seqFastqCreate(0, SEQTYP_FASTQ), *seg = seqFastqCreate(0, SEQTYP_FASTQ);
  char seq[512], qual[512], out[512], outq[512];
  int len, rep, has_qual, nsimd = 0;
  if (!codep || !a || !b || !seg) return 2;
  for (len = 1; len <= 200; len++)
    for (rep = 0; rep < 12; rep++) {
      int i, reverse;
      const int kind = rep % 4;
      for (i = 0; i < len; i++) {
        const unsigned r = rnd();
        if (kind == 0) seq[i] = "ACGT"[r & 3];
        else if (kind == 1) seq[i] = "ACGTacgt"[r & 7];
        else if (kind == 2) seq[i] = (r % 23) ? "ACGT"[r & 3] : "NnRYKMSWBDHVUuXx"[(r >> 8) & 15];
        else seq[i] = (char) (0x21 + (r >> 4) % (0x7f - 0x21));
        qual[i] = (char) (33 + (r >> 16) % 60);
      }
      seq[len] = qual[len] = '\0';
      seqFastqBlank(a); seqFastqBlank(b);
      if (seqFastqSetAscii(a, "name", seq, "", qual) || seqFastqEncode(a, codep)) { printf("reference load failed\n"); return 1; }
      if (smbShimSeqFastqLoad(b, "name", 4, seq, (size_t) len, "", 0, qual, (size_t) len, codep)) { printf("load failed\n"); return 1; }
      if (a->datap->size != b->datap->size || a->datap->code != b->datap->code ||
          memcmp(a->datap->basep, b->datap->basep, (size_t) len + 1)) { printf("encode differs: len %d kind %d\n", len, kind); return 1; }
      if (memcmp(a->qualp->basep, b->qualp->basep, (size_t) len + 1)) { printf("qual differs\n"); return 1; }
      for (reverse = 0; reverse < 2; reverse++) {
        const SEQLEN_t start = (SEQLEN_t) (rnd() % (unsigned) len), sl = (SEQLEN_t) (1 + rnd() % (unsigned) (len - (int) start));
        seqFastqBlank(seg);
        if (seqFastqAppendSegment(seg, a, start, sl, (char) reverse, codep) || seqFastqDecode(seg, codep)) { printf("reference segment failed\n"); return 1; }
        if (smbShimSeqFastqDecodeSegment(out, outq, &has_qual, b, start, sl, reverse, codep)) { printf("decode failed\n"); return 1; }
        if (strcmp(out, seg->datap->basep) || !has_qual || strcmp(outq, seg->qualp->basep)) {
          printf("decode differs: len %d kind %d reverse %d start %u seglen %u\n  %s\n  %s\n", len, kind, reverse, start, sl, out, seg->datap->basep);
          return 1;
        }
      }
    }
#ifdef SHIM_SIMD
  nsimd = shim_simd(codep)->ok;
#endif
  printf("ok simd=%d\n", nsimd);
  return 0;
}
