/* ref_trace.c - TEST INFRASTRUCTURE ONLY.
 *
 * Link-time interposers (GNU ld --wrap) for the three DP entry points the
 * reference driver calls from rmap.c (rmap.c:720, :734, :898).  Linked with
 * the UNMODIFIED reference objects they give oracle/_ref/smalt_trace, a
 * `smalt` binary that behaves exactly like the reference but appends one text
 * record per DP call to the file named by $SMALT_TRACE when that variable is
 * set.  The records are the reference's own inputs/outputs at the hot-path
 * boundary and are the source of tests/golden/dp_trace_*.txt
 * (tests/golden/make_golden.py).
 *
 * Record formats (sequences as letters over "ACGTXN"):
 *  SW <err> <score> <qlen> <rlen> <read> <ref>
 *  BF <err> <score> <l_edge> <r_edge> <pl> <pr> <ul> <ur> <qlen> <rlen> <read> <ref>
 *  BA <err> <l_edge> <r_edge> <pl> <pr> <ul> <ur> <minscore> <minscorlen>
 *     <qlen> <rlen> <read> <ref> <nres> { <score> <qs> <qe> <rs> <re> <diffstr hex> }*
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sequence.h"
#include "score.h"
#include "alibuffer.h"
#include "alignment.h"
#include "swsimd.h"
#include "diffstr.h"

static FILE *g_fp;
static int g_tried;
static long g_max = -1, g_n;

static FILE *trace_fp(void)
{
  if (!g_tried) {
    const char *fn = getenv("SMALT_TRACE");
    const char *mx = getenv("SMALT_TRACE_MAX");
    g_tried = 1;
    if (fn) g_fp = fopen(fn, "w");
    if (mx) g_max = atol(mx);
  }
  if (g_fp && g_max >= 0 && g_n >= g_max) return NULL;
  return g_fp;
}

static void put_read(FILE *fp, const ScoreProfile *profp)
{
  static const char L[] = "ACGTXN";
  short asiz;
  SEQLEN_t qlen, j;
  signed char *const *sc = scoreGetProfile(&asiz, &qlen, NULL, NULL, profp);
  short match = scoreProfileGetAvgPenalties(NULL, NULL, NULL, profp);
  for (j = 0; j < qlen; j++) {
    int c, code = -1;
    for (c = 0; c < 4; c++) if (sc[c][j] == match) { code = c; break; }
    if (code < 0) code = (sc[0][j] == 0) ? 5 : 4;
    fputc(L[code], fp);
  }
}

static void put_ref(FILE *fp, const char *p, int n)
{
  static const char L[] = "ACGTXN??";
  int i;
  for (i = 0; i < n; i++) fputc(L[p[i] & 7], fp);
}

int __real_swSIMDAlignStriped(int *, const AliBuffer *, const ScoreProfile *, const char *, int);
int __real_aliSmiWatInBandFast(int *, AliBuffer *, const ScoreProfile *, const char *, int,
			       int, int, int, int, int, int);
int __real_aliSmiWatInBand(AliRsltSet *, AliBuffer *, const ScoreProfile *, const char *, int,
			   int, int, int, int, int, int, int, int);

int __wrap_swSIMDAlignStriped(int *maxscor, const AliBuffer *abp, const ScoreProfile *profp,
			      const char *usp, int uslen)
{
  int errcode = __real_swSIMDAlignStriped(maxscor, abp, profp, usp, uslen);
  FILE *fp = trace_fp();
  if (fp) {
    SEQLEN_t qlen;
    scoreGetProfile(NULL, &qlen, NULL, NULL, profp);
    fprintf(fp, "SW %d %d %u %d ", errcode, *maxscor, qlen, uslen);
    put_read(fp, profp); fputc(' ', fp); put_ref(fp, usp, uslen); fputc('\n', fp);
    g_n++;
  }
  return errcode;
}

int __wrap_aliSmiWatInBandFast(int *maxscor, AliBuffer *bufp, const ScoreProfile *profp,
			       const char *usp, int uslen, int l_edge, int r_edge,
			       int pl, int pr, int ul, int ur)
{
  int errcode = __real_aliSmiWatInBandFast(maxscor, bufp, profp, usp, uslen,
					   l_edge, r_edge, pl, pr, ul, ur);
  FILE *fp = trace_fp();
  if (fp) {
    SEQLEN_t qlen;
    scoreGetProfile(NULL, &qlen, NULL, NULL, profp);
    fprintf(fp, "BF %d %d %d %d %d %d %d %d %u %d ", errcode, *maxscor,
	    l_edge, r_edge, pl, pr, ul, ur, qlen, uslen);
    put_read(fp, profp); fputc(' ', fp); put_ref(fp, usp, uslen); fputc('\n', fp);
    g_n++;
  }
  return errcode;
}

int __wrap_aliSmiWatInBand(AliRsltSet *rssp, AliBuffer *bufp, const ScoreProfile *profp,
			   const char *usp, int uslen, int l_edge, int r_edge,
			   int pl, int pr, int ul, int ur, int minscore, int minscorlen)
{
  short n0 = aliRsltSetGetSize(rssp);
  int errcode = __real_aliSmiWatInBand(rssp, bufp, profp, usp, uslen,
				       l_edge, r_edge, pl, pr, ul, ur, minscore, minscorlen);
  FILE *fp = trace_fp();
  if (fp) {
    SEQLEN_t qlen;
    short i, n = aliRsltSetGetSize(rssp);
    scoreGetProfile(NULL, &qlen, NULL, NULL, profp);
    fprintf(fp, "BA %d %d %d %d %d %d %d %d %d %u %d ", errcode,
	    l_edge, r_edge, pl, pr, ul, ur, minscore, minscorlen, qlen, uslen);
    put_read(fp, profp); fputc(' ', fp); put_ref(fp, usp, uslen);
    fprintf(fp, " %d", (int) (n - n0));
    for (i = n0; i < n; i++) {
      int sc, qs, qe, rs, re, k;
      const DiffStr *dfs;
      aliRsltSetFetchData(rssp, i, &sc, &qs, &qe, &rs, &re, &dfs);
      fprintf(fp, " %d %d %d %d %d ", sc, qs, qe, rs, re);
      for (k = 0; k < dfs->len; k++) fprintf(fp, "%02x", dfs->dstrp[k]);
    }
    fputc('\n', fp);
    g_n++;
  }
  return errcode;
}
