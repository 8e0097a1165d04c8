/* smalt_oracle_cigar.c - TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and nothing else).
 *
 * CPU restatement of what the reference's output stage derives from a compressed alignment
 * string (DiffStr): the CIGAR field and the NM:i: edit distance of a SAM record.
 *   writeDiffStrCIGAR               /root/reference/src/diffstr.c:298-367 (formats CIGEXT and
 *                                   CIGEXT_XMISMATCH of diffStrPrintfStr, :1103-1114)
 *   diffStrGetLevenshteinDistance   /root/reference/src/diffstr.c:1496-1510
 * Pinned against those two functions of the compiled reference (oracle/_ref/libsmalt_ref.so)
 * in tests/test_oracle_cigar_vs_ref.py and against the known answers of the reference's own
 * test/bam_cigar_test.py.
 *
 * Formulation (deliberately not the reference's state machine): a DiffStr byte (type << 6 | n)
 * stands for n matching columns followed by ONE column of its type (diffstr.h:28-77) - M: another
 * match, D / I: a gap column, S: a mismatch - except that the last byte must be of type S and
 * its own column does not exist.  The CIGAR is the run-length encoding of that column sequence
 * with match and mismatch columns merged into M (or kept apart as M / X), framed by the clips.
 */
#include "smalt_oracle.h"
#include <stdio.h>

typedef struct {
  char *out;
  int max, len, overflow;
  char op;       /* current run */
  unsigned n;
} CigRun;

static void run_flush(CigRun *r)
{
  char buf[16];
  int k, i;
  if (!r->n) return;
  k = snprintf(buf, sizeof buf, "%u%c", r->n, r->op);
  for (i = 0; i < k; i++) {
    if (r->len < r->max) r->out[r->len] = buf[i];
    else r->overflow = 1;
    r->len++;
  }
  r->n = 0;
}

static void run_push(CigRun *r, char op, unsigned n)
{
  if (!n) return;
  if (r->n && r->op != op) run_flush(r);
  r->op = op;
  r->n += n;
}

/* flags: 2 = soft clips ('S' instead of 'H'), 4 = mismatches as 'X'.  Returns the text length
 * (the text is NOT terminated), -1 for an empty string (ERRCODE_FAILURE), -59 if the string does
 * not end with an S byte (ERRCODE_DIFFSTR), -2 if `out` is too small.  *nm = edit distance. */
int so_cigar(const unsigned char *diffstr, int clip_start, int clip_end, int flags, char *out, int maxout, int *nm)
{
  CigRun r = {out, maxout, 0, 0, 'M', 0};
  const char clipc = (flags & 2) ? 'S' : 'H', mmc = (flags & 4) ? 'X' : 'M';
  int i, ed = 0, last_typ = 0;
  if (!diffstr || !diffstr[0]) return -1;
  if (clip_start > 0) { run_push(&r, clipc, (unsigned) clip_start); run_flush(&r); }
  for (i = 0; diffstr[i]; i++) {
    const unsigned n = diffstr[i] & 63u;
    const int typ = diffstr[i] >> 6, is_last = !diffstr[i + 1];
    last_typ = typ;
    run_push(&r, 'M', n);
    if (is_last) break;       /* the closing byte has no column of its own */
    if (typ == 0) run_push(&r, 'M', 1);
    else if (typ == 1) { run_push(&r, 'D', 1); ed++; }
    else if (typ == 2) { run_push(&r, 'I', 1); ed++; }
    else { run_push(&r, mmc, 1); ed++; }
  }
  if (last_typ != 3) return -59;
  run_flush(&r);
  if (clip_end > 0) { run_push(&r, clipc, (unsigned) clip_end); run_flush(&r); }
  if (nm) *nm = ed;
  return r.overflow ? -2 : r.len;
}
