/* smalt_oracle.c - plain-C restatement of the SMALT 0.7.6 hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see smalt_oracle.h).  Written from the behaviour of
 * the reference, each function citing the reference file:line it follows
 * (paths relative to /root/reference/src).  Parity is pinned against the real
 * reference by tests/test_oracle_vs_ref.py and tests/golden/.
 */
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>
#include "smalt_oracle.h"

/* ------------------------------------------------------------------------ */
/* scoring: setScoreMatrix (score.c:138-173) over the alphabet "ACGTXN"     */
/* ------------------------------------------------------------------------ */
void so_scoring_init(so_scoring *sc, int match, int mismatch, int gapopen, int gapext)
{
  int a, b;
  sc->match = match;
  sc->mismatch = mismatch;
  sc->gap_init = -gapopen; /* scoreGetProfile returns -1*gap (score.c:682-683) */
  sc->gap_ext = -gapext;
  for (a = 0; a < 8; a++)
    for (b = 0; b < 8; b++) {
      int s;
      if (a >= 6 || b >= 6 || a == 5 || b == 5) s = 0;        /* N or outside alphabet */
      else if (a == 4 || b == 4) s = mismatch - match;        /* X */
      else s = (a == b) ? match : mismatch;
      sc->S[a][b] = (signed char) s;
    }
}

/* ------------------------------------------------------------------------ */
/* K2: swSIMDAlignStriped (swsimd.c:868-933)                                */
/*                                                                          */
/* The striped 8-bit kernel (swsimd.c:656-861) and its 16-bit retry         */
/* (:443-654) compute, lane by lane, the canonical affine-gap local         */
/* alignment recurrence                                                     */
/*   h = max(0, Hdiag + S);  H = max(h, E, F)                               */
/*   E' = max(E - ext, H - init);  F' = max(F - ext, H - init)              */
/* (swsimd.c:745-781; the lazy-F loop :803-831 only completes F across      */
/* stripe boundaries), all values floored at 0 by the unsigned saturating   */
/* arithmetic, and return max h.  Saturation is detected, never returned:   */
/* 8-bit overflow (score + bias >= 255, :854) triggers the 16-bit pass,     */
/* whose own overflow (score >= 65535, :644) yields ERRCODE_SWATEXCEED.     */
/* So the result is the exact maximum unless it reaches 65535.              */
/* ------------------------------------------------------------------------ */
int so_sw_striped(const so_scoring *sc, const uint8_t *read, int qlen,
		  const uint8_t *ref, int rlen, int *score)
{
  int i, j, best = 0;
  const int gi = sc->gap_init, ge = sc->gap_ext;
  int *H, *E;
  *score = 0;
  if (qlen < 1) return SO_SUCCESS;
  H = (int *) calloc((size_t) qlen + 1, sizeof(int));
  E = (int *) calloc((size_t) qlen + 1, sizeof(int));
  if (!H || !E) { free(H); free(E); return SO_NOMEM; }
  for (i = 0; i < rlen; i++) {
    const signed char *row = sc->S[ref[i] & 7];
    int diag = 0, F = 0;
    for (j = 0; j < qlen; j++) {
      int h = diag + row[read[j] & 7], t;
      if (h < 0) h = 0;
      if (h > best) best = h;
      diag = H[j];
      if (E[j] > h) h = E[j];
      if (F > h) h = F;
      H[j] = h;
      t = h - gi;
      if (t < 0) t = 0;
      E[j] = (E[j] - ge > t) ? E[j] - ge : t;
      F = (F - ge > t) ? F - ge : t;
    }
  }
  free(H); free(E);
  if (best >= 65535) return SO_SWATEXCEED;
  *score = best;
  return SO_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* band geometry: initALIBAND (alignment.c:310-396)                         */
/* ------------------------------------------------------------------------ */
int so_band_init(so_band *b, int l_edge, int r_edge, int q_left, int q_right, int q_len,
		 int s_left, int s_right, int s_len)
{
  b->s_len = (s_right < 0 || s_right >= s_len) ? s_len : s_right + 1;
  b->q_len = (q_right < 0 || q_right >= q_len) ? q_len : q_right + 1;
  b->s_totlen = s_len;
  b->q_totlen = q_len;
  b->s_left = b->s_left_orig = (s_left > 0 && s_left < b->s_len) ? s_left : 0;
  b->q_left = b->q_left_orig = (q_left > 0 && q_left < b->q_len) ? q_left : 0;
  b->l_edge_orig = b->l_edge = l_edge;
  b->r_edge_orig = b->r_edge = r_edge;
  b->band_width = r_edge - l_edge + 1;
  if (b->band_width <= 0) {
    b->band_width = 0;
    b->l_edge = b->q_left;
    b->r_edge = b->q_len - 1;
  } else {
    if (b->l_edge_orig + b->s_len > b->q_len) b->s_len = b->q_len - b->l_edge_orig;
    b->l_edge += b->s_left;
    if (b->l_edge >= b->q_len || b->r_edge_orig + b->s_len <= b->q_left)
      return SO_FAILURE;
    b->r_edge += b->s_left;
    if (b->r_edge < b->q_left) {
      const int d = b->q_left - b->r_edge;
      b->s_left += d;
      b->l_edge += d;
      b->r_edge = b->q_left;
    }
    if (b->r_edge > b->q_len - 1) b->r_edge = b->q_len - 1;
  }
  b->band_width = b->r_edge - b->l_edge + 1;
  return (b->band_width >= 0) ? SO_SUCCESS : SO_FAILURE;
}

/* ------------------------------------------------------------------------ */
/* banded DP: alignSmiWatBand (alignment.c:788-1027, dirs != NULL) and      */
/* alignSmiWatBandFast (alignment.c:1029-1233, dirs == NULL).               */
/*                                                                          */
/* The "restricted" recurrence: a gap state is only (re)opened from a cell  */
/* whose H came from the diagonal; non-positive E/F mean "no gap"; a        */
/* maximum is recorded only for diagonal cells with H > gap_init, first     */
/* strict maximum in row-major order (alignment.c:826-830).                 */
/* Direction bytes (alignment.c:53-59): 0 stop, 1 COL, 2 ROW, 3 DIA, stored */
/* at dirs[r*(bw-1) + j - l_edge] (alignment.c:676, :866, :1006-1017).      */
/* Fast variant quirk: once the band start is clipped at q_left it is never */
/* released (alignment.c:1213-1218 lacks the `--delta_band_start` of :1008).*/
/* ------------------------------------------------------------------------ */
typedef struct { int max_i, max_j, max_scor; } so_track;

static int band_dp(const so_scoring *sc, const so_band *b, const uint8_t *read,
		   const uint8_t *ref, uint8_t *dirs, so_track *tk, long long *cells)
{
  const int gi = sc->gap_init, ge = sc->gap_ext;
  int i, j, jstart, jlen, dstart, dend = 0;
  int max_i = 0, max_j = 0, max_scor = 0;
  int currH = 0, F, h;
  long long ncell = 0;
  int *Hp, *Ep;
  uint8_t *dp;
  const int fast = (dirs == NULL);

  Hp = (int *) calloc((size_t) b->q_len + 1, sizeof(int));
  Ep = (int *) calloc((size_t) b->q_len + 1, sizeof(int));
  if (!Hp || !Ep) { free(Hp); free(Ep); return SO_NOMEM; }

  if (b->q_left > b->l_edge) { dstart = b->q_left - b->l_edge; jstart = b->q_left; }
  else { dstart = 0; jstart = b->l_edge; }
  jlen = b->r_edge + 1;
  dp = fast ? NULL : dirs + dstart;

  for (i = b->s_left; i < b->s_len; i++) {
    const signed char *row = sc->S[ref[i] & 7];
    F = 0;
    for (j = jstart; j < jlen; j++) {
      int e = Ep[j], d;
      h = currH + row[read[j] & 7];
      currH = Hp[j];
      ncell++;
      if (F > 0) {
	if (e > 0) {
	  if (h > e) {
	    if (h > F) {
	      Hp[j] = h; d = 3;
	      F -= ge; e -= ge;
	      if (h > gi) {
		const int t = h - gi;
		if (h > max_scor) { max_scor = h; max_i = i; max_j = j; }
		if (F < t) F = t;
		if (e < t) e = t;
	      }
	    } else { Hp[j] = F; d = 2; F -= ge; e -= ge; }
	  } else {
	    if (e >= F) { Hp[j] = e; d = 1; } else { Hp[j] = F; d = 2; }
	    e -= ge; F -= ge;
	  }
	} else {
	  if (h > F) {
	    Hp[j] = h; d = 3;
	    F -= ge;
	    if (h > gi) {
	      if (h > max_scor) { max_scor = h; max_i = i; max_j = j; }
	      e = h - gi;
	      if (F < e) F = e;
	    }
	  } else { Hp[j] = F; d = 2; F -= ge; }
	}
      } else if (e > 0) {
	if (h > e) {
	  Hp[j] = h; d = 3;
	  e -= ge;
	  if (h > gi) {
	    if (h > max_scor) { max_scor = h; max_i = i; max_j = j; }
	    F = h - gi;
	    if (e < F) e = F;
	  }
	} else { Hp[j] = e; d = 1; e -= ge; }
      } else {
	if (h > 0) {
	  Hp[j] = h; d = 3;
	  if (h > gi) {
	    if (h > max_scor) { max_scor = h; max_i = i; max_j = j; }
	    F = e = h - gi;
	  }
	} else { Hp[j] = 0; d = 0; }
      }
      Ep[j] = e;
      if (!fast) *dp++ = (uint8_t) d;
    }
    if (dstart > 0) {
      currH = 0;
      if (!fast) dp += --dstart;      /* alignment.c:1006-1008; fast: never released */
    } else {
      currH = Hp[jstart];
      jstart++;
    }
    if (jlen < b->q_len) jlen++;
    else if (!fast) dp += dend++;
  }
  free(Hp); free(Ep);
  tk->max_i = max_i; tk->max_j = max_j; tk->max_scor = max_scor;
  if (cells) *cells += ncell;
  return SO_SUCCESS;
}

int so_band_fast(const so_scoring *sc, const uint8_t *read, int qlen,
		 const uint8_t *ref, int rlen, int l_edge, int r_edge,
		 int p_left, int p_right, int u_left, int u_right,
		 int *maxscore, long long *cells)
{
  so_band b;
  so_track tk;
  int errcode;
  /* aliSmiWatInBandFast returns initALIBAND's failure to the caller (alignment.c:1622-1627) */
  if ((errcode = so_band_init(&b, l_edge, r_edge, p_left, p_right, qlen, u_left, u_right, rlen)))
    return errcode;
  if ((errcode = band_dp(sc, &b, read, ref, NULL, &tk, cells))) return errcode;
  *maxscore = tk.max_scor;
  return SO_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* DiffStr: 1 byte = type<<6 | count; M0 D1 I2 S3 (diffstr.h:28-105)        */
/* ------------------------------------------------------------------------ */
#define DIFF(count, typ) ((uint8_t) ((count) + ((typ) << 6)))
enum { DF_M = 0, DF_D = 1, DF_I = 2, DF_S = 3, DF_MAX = 61 };

/* diffStrReverse (diffstr.c:850-896) */
int so_diffstr_reverse(const uint8_t *in, uint8_t *out, int maxout)
{
  int l, u = 0;
  uint8_t typ, count, count_prev;
  for (l = 0; in[l]; l++)
    if (l >= SHRT_MAX) return -SO_OVERFLOW;
  if (l + 1 > maxout) return -SO_NOMEM;
  l--;
  if (l < 0) return -SO_DIFFSTR;
  count_prev = in[l] & 0x3F; typ = in[l] >> 6;
  if (typ != DF_S) return -SO_DIFFSTR;
  for (l--; l >= 0; l--) {
    count = in[l] & 0x3F; typ = in[l] >> 6;
    if (typ == DF_M) {
      count_prev = (uint8_t) (count_prev + count + 1);
      if (count_prev > DF_MAX) {
	out[u++] = DIFF(DF_MAX, DF_M);
	count_prev -= DF_MAX + 1;
      }
    } else {
      out[u++] = DIFF(count_prev, typ);
      count_prev = count;
    }
  }
  out[u++] = DIFF(count_prev, DF_S);
  out[u++] = DIFF(0, DF_M);
  return u;
}

/* makeMetaFromTrack (alignment.c:628-781): backtrace from (max_i,max_j),
 * emitting the *reversed* diff string; returns its length (without the
 * terminator) or a negative error code. */
typedef struct { int prof_start, prof_end, nonprof_start, nonprof_end, score; } so_meta;

static int backtrace(const so_scoring *sc, const so_band *b, const so_track *tk,
		     const uint8_t *dirs, const uint8_t *read, const uint8_t *ref,
		     uint8_t *rev, int maxrev, so_meta *m)
{
  int i = tk->max_i, j = tk->max_j, n = 0, checksum = 0, gap_open = 0;
  uint8_t nmatch = 0;
  const uint8_t *dp = dirs + (long) (tk->max_i - b->s_left) * (b->band_width - 1)
    + tk->max_j - b->l_edge;
#define EMIT(c, t) do { if (n >= maxrev) return -SO_NOMEM; rev[n++] = DIFF(c, t); } while (0)
  while (i >= b->s_left && j >= b->q_left && *dp) {
    if (*dp == 3) {
      const int s = sc->S[ref[i] & 7][read[j] & 7];
      if (s > 0) {
	if (nmatch > DF_MAX) { EMIT(DF_MAX, DF_M); nmatch -= DF_MAX; }
	else nmatch++;
      } else {
	EMIT(nmatch, DF_S);
	nmatch = 0;
      }
      checksum += s;
      gap_open = 0;
      dp -= b->band_width;
      i--; j--;
      continue;
    }
    if (gap_open) checksum -= sc->gap_ext;
    else { checksum -= sc->gap_init; gap_open = 1; }
    if (*dp & 1) {
      EMIT(nmatch, DF_D);
      nmatch = 0;
      dp -= b->band_width - 1;
      i--;
      continue;
    }
    EMIT(nmatch, DF_I);
    nmatch = 0;
    dp--;
    j--;
  }
  EMIT(nmatch, DF_S);
  EMIT(0, DF_M);
#undef EMIT
  m->nonprof_start = i + 1;
  m->nonprof_end = tk->max_i;
  m->prof_start = j + 1;
  m->prof_end = tk->max_j;
  m->score = checksum;
  if (checksum != tk->max_scor) return -SO_SWATSCOR;
  return n;
}

/* alignSmiWatBandRecursive (alignment.c:1300-1434) */
typedef struct {
  const so_scoring *sc;
  const uint8_t *read, *ref;
  int qlen, rlen, l_edge, r_edge, q_left, q_right, minscore, minscorlen;
  int maxres, nres, *out5;
  int maxdiff, useddiff, *difflen;
  uint8_t *diffbuf;
  long long *cells;
} so_rec;

static int band_recursive(so_rec *R, int s_left, int s_right)
{
  so_band b;
  so_track tk;
  so_meta m;
  int errcode, n, s_start, s_end;
  uint8_t *dirs, *rev;
  size_t ndir;

  if (R->minscorlen < 2) return SO_ASSERT;
  if (so_band_init(&b, R->l_edge, R->r_edge, R->q_left, R->q_right, R->qlen,
		   s_left, s_right, R->rlen))
    return SO_SUCCESS; /* inconsistent limits silently end the recursion (:1333-1338) */
  if (b.s_left >= b.s_len || b.band_width < 0) return SO_ASSERT; /* setMemALITRACK :459 */
  ndir = (size_t) b.band_width * (size_t) (b.s_len - b.s_left) + 1;
  dirs = (uint8_t *) malloc(ndir);
  if (!dirs) return SO_NOMEM;
  if ((errcode = band_dp(R->sc, &b, R->read, R->ref, dirs, &tk, R->cells))) {
    free(dirs); return errcode;
  }
  if (tk.max_scor < R->minscore) { free(dirs); return SO_SUCCESS; }
  rev = (uint8_t *) malloc((size_t) R->qlen + (size_t) R->rlen + 4);
  if (!rev) { free(dirs); return SO_NOMEM; }
  n = backtrace(R->sc, &b, &tk, dirs, R->read, R->ref, rev, R->qlen + R->rlen + 4, &m);
  free(dirs);
  if (n < 0) { free(rev); return -n; }
  if (m.prof_start + R->minscorlen > m.prof_end + 1) { free(rev); return SO_SUCCESS; }
  s_start = m.nonprof_start;
  s_end = m.nonprof_end;
  if (m.score >= R->minscore) {
    int *o, len;
    if (R->nres >= R->maxres) { free(rev); return SO_OVERFLOW; }
    len = so_diffstr_reverse(rev, R->diffbuf + R->useddiff, R->maxdiff - R->useddiff);
    if (len < 0) { free(rev); return -len; }
    o = R->out5 + 5 * R->nres;
    o[0] = m.score; o[1] = m.prof_start; o[2] = m.prof_end;
    o[3] = m.nonprof_start; o[4] = m.nonprof_end;
    R->difflen[R->nres] = len;
    R->useddiff += len;
    R->nres++;
  }
  free(rev);
  if (s_left + R->minscorlen < s_start &&
      (errcode = band_recursive(R, s_left, s_start - 1)))
    return errcode;
  if (s_right > s_end + R->minscorlen &&
      (errcode = band_recursive(R, s_end + 1, s_right)))
    return errcode;
  return SO_SUCCESS;
}

/* aliSmiWatInBand (alignment.c:1548-1601) */
int so_band_align(const so_scoring *sc, const uint8_t *read, int qlen,
		  const uint8_t *ref, int rlen, int l_edge, int r_edge,
		  int p_left, int p_right, int u_left, int u_right,
		  int minscore, int minscorlen,
		  int maxres, int *nres, int *out5,
		  int maxdiff, uint8_t *diffbuf, int *difflen, long long *cells)
{
  so_rec R;
  int errcode;
  *nres = 0;
  if (minscore < 1 || sc->match <= 0) return SO_ASSERT;
  if (minscorlen * sc->match < minscore) minscorlen = minscore / sc->match;
  if (minscorlen < 5) return SO_ASSERT; /* ALILEN_MIN alignment.c:50, :1574 */
  R.sc = sc; R.read = read; R.ref = ref; R.qlen = qlen; R.rlen = rlen;
  R.l_edge = l_edge; R.r_edge = r_edge; R.q_left = p_left; R.q_right = p_right;
  R.minscore = minscore; R.minscorlen = minscorlen;
  R.maxres = maxres; R.nres = 0; R.out5 = out5;
  R.maxdiff = maxdiff; R.useddiff = 0; R.difflen = difflen; R.diffbuf = diffbuf;
  R.cells = cells;
  errcode = band_recursive(&R, u_left, u_right);
  *nres = R.nres;
  return errcode;
}

/* ======================================================================== */
/* K1: index lookup, seed table, ranking, hit lists                          */
/* ======================================================================== */

/* hash32mix (hashidx.c:163-172) */
uint32_t so_hash32mix(uint32_t a)
{
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

/* derived masks as set by hashTableCreate (hashidx.c:640-760) */
void so_index_setup(so_index *ix, int typ, int wordlen, int nskip, int nbits_key,
		    int nbits_lo, uint32_t npos, uint32_t nwords,
		    const uint32_t *idx, const uint32_t *pos,
		    const uint32_t *wordidx, const uint32_t *posidx)
{
  memset(ix, 0, sizeof(*ix));
  ix->typ = typ; ix->wordlen = wordlen; ix->nskip = nskip;
  ix->nbits_key = nbits_key; ix->nbits_lo = nbits_lo;
  ix->npos = npos; ix->nwords = nwords;
  ix->idx = idx; ix->pos = pos; ix->wordidx = wordidx; ix->posidx = posidx;
  ix->wordmask = (wordlen >= 32) ? ~0ULL : ((1ULL << (2 * wordlen)) - 1);
  if (typ == 0) {
    ix->nkeys = (uint32_t) 1 << (2 * wordlen);
  } else {
    ix->nkeys = (uint32_t) 1 << nbits_key;
    ix->wordmask_lo = (1ULL << nbits_lo) - 1;
    ix->wordmask_hi = ix->wordmask & ~ix->wordmask_lo;
    ix->keymod = (uint32_t) 1 << (nbits_key - nbits_lo);
  }
}

/* hashTableGetKtupleHits (hashidx.c:1146-1191) */
uint32_t so_lookup(const so_index *ix, uint64_t word, uint32_t *posidx)
{
  uint32_t nhits = 0;
  if (ix->typ == 0) {
    const uint32_t key = (uint32_t) (word & ix->wordmask);
    if (posidx) *posidx = key;
    if (key < ix->nkeys) nhits = ix->idx[key + 1] - ix->idx[key];
  } else {
    const uint32_t word_hi = (uint32_t) ((word & ix->wordmask_hi) >> ix->nbits_lo);
    const uint32_t key_hi = so_hash32mix(word_hi) % ix->keymod;
    const uint32_t key = (key_hi << ix->nbits_lo) + (uint32_t) (word & ix->wordmask_lo);
    uint32_t a, b = ix->idx[key + 1];
    if (b < 1) return 0;
    a = ix->idx[key];
    b--;
    while (a < b) {
      const uint32_t pivot = (a + b) >> 1;
      if (ix->wordidx[pivot] < word_hi) a = pivot + 1; else b = pivot;
    }
    if (a == b && ix->wordidx[b] == word_hi) {
      nhits = ix->posidx[b + 1] - ix->posidx[b];
      if (posidx) *posidx = b;
    }
  }
  return nhits;
}

/* hashTableFetchHitPositions (hashidx.c:1193-1212) */
static uint32_t fetch_positions(const so_index *ix, uint32_t posidx, const uint32_t **posp)
{
  *posp = NULL;
  if (ix->typ == 0) {
    if (posidx < ix->nkeys) {
      *posp = ix->pos + ix->idx[posidx];
      return ix->idx[posidx + 1] - ix->idx[posidx];
    }
  } else if (posidx < ix->npos) {
    *posp = ix->pos + ix->posidx[posidx];
    return ix->posidx[posidx + 1] - ix->posidx[posidx];
  }
  return 0;
}

so_hitinfo *so_hitinfo_create(uint32_t maxlen, int nskip)
{
  so_hitinfo *h = (so_hitinfo *) calloc(1, sizeof(*h));
  const size_t n = (size_t) maxlen + 2;
  if (!h) return NULL;
  h->n_alloc = (uint32_t) n;
  h->posidx = (uint32_t *) calloc(n, 4); h->nhits = (uint32_t *) calloc(n, 4);
  h->cix = (uint32_t *) calloc(n, 4); h->qoffs = (uint32_t *) calloc(n, 4);
  h->sortkey = (uint32_t *) calloc(n, 4); h->sidx = (uint32_t *) calloc(n, 4);
  h->qmask = (uint8_t *) calloc(n, 1); h->qbuf = (uint8_t *) calloc(n, 1);
  h->frame_cnt = (uint32_t *) calloc((size_t) nskip, 4);
  h->frame_ix = (uint32_t *) calloc(n * (size_t) nskip, 4);
  return h;
}

void so_hitinfo_delete(so_hitinfo *h)
{
  if (!h) return;
  free(h->posidx); free(h->nhits); free(h->cix); free(h->qoffs);
  free(h->sortkey); free(h->sidx); free(h->qmask); free(h->qbuf);
  free(h->frame_cnt); free(h->frame_ix);
  free(h);
}

/* sort2UINTarraysByQuickSort (sort.c:233-330): median-of-three quicksort with
 * insertion sort for short partitions and an explicit stack, processing the
 * smaller partition first.  NOT stable; the resulting tie order is observable
 * downstream (seed_rank cut), so the same exchange sequence is reproduced. */
#define XCHG(T, x, y) do { T t_ = (x); (x) = (y); (y) = t_; } while (0)
int so_sort2(uint32_t n, uint32_t *key, uint32_t *val)
{
  enum { SMALL = 7, STACK = 60 };
  int lo = 0, hi = (int) n - 1, sp = 0, i, j;
  int stack[STACK + 2];
  for (;;) {
    if (hi - lo < SMALL) {
      for (j = lo + 1; j <= hi; j++) {
	const uint32_t k = key[j], v = val[j];
	for (i = j - 1; i >= lo && key[i] > k; i--) { key[i + 1] = key[i]; val[i + 1] = val[i]; }
	key[i + 1] = k; val[i + 1] = v;
      }
      if (!sp) return SO_SUCCESS;
      hi = stack[sp--];
      lo = stack[sp--];
    } else {
      const int mid = (lo + hi) >> 1;
      uint32_t pk, pv;
      XCHG(uint32_t, key[mid], key[lo + 1]); XCHG(uint32_t, val[mid], val[lo + 1]);
      if (key[lo] > key[hi]) { XCHG(uint32_t, key[lo], key[hi]); XCHG(uint32_t, val[lo], val[hi]); }
      if (key[lo + 1] > key[hi]) { XCHG(uint32_t, key[lo + 1], key[hi]); XCHG(uint32_t, val[lo + 1], val[hi]); }
      if (key[lo] > key[lo + 1]) { XCHG(uint32_t, key[lo], key[lo + 1]); XCHG(uint32_t, val[lo], val[lo + 1]); }
      i = lo + 1; j = hi;
      pk = key[lo + 1]; pv = val[lo + 1];
      for (;;) {
	do i++; while (key[i] < pk);
	do j--; while (key[j] > pk);
	if (j < i) break;
	XCHG(uint32_t, key[i], key[j]); XCHG(uint32_t, val[i], val[j]);
      }
      key[lo + 1] = key[j]; val[lo + 1] = val[j];
      key[j] = pk; val[j] = pv;
      sp += 2;
      if (sp > STACK) return 34; /* ERRCODE_SORTSTACK */
      if (hi - i + 1 >= j - lo) { stack[sp] = hi; stack[sp - 1] = i; hi = j - 1; }
      else { stack[sp] = j - 1; stack[sp - 1] = lo; lo = i; }
    }
  }
}

static int cmp_u64(const void *a, const void *b)
{
  const uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
  return (x > y) - (x < y);
}
/* sortUINT64arrayByQuickSort (sort.c:415-497): plain ascending sort of a
 * single array - any correct sort gives the same array. */
int so_sort64(uint32_t n, uint64_t *a)
{
  qsort(a, n, sizeof(uint64_t), cmp_u64);
  return SO_SUCCESS;
}

/* getHitInfoMaxRank (hashhit.c:769-891) */
static int max_rank(so_hitinfo *h, int ktup, int nskip, uint32_t mincover, uint32_t maxcover,
		    uint32_t maxhit)
{
  uint32_t i, n, nmax, ntot;
  int f;
  if (h->n_seeds < 1 || maxcover < mincover) return SO_ASSERT;
  for (f = 0; f < nskip; f++) h->frame_cnt[f] = 0;
  for (i = 0; i < h->n_seeds; i++) {
    const uint32_t s = h->sidx[i];
    f = (int) (h->qoffs[s] % (uint32_t) nskip);
    h->frame_ix[(size_t) f * h->n_alloc + h->frame_cnt[f]++] = i; /* the rank */
  }
  /* note: reads sortkey[n_seeds] (one past) exactly like hashhit.c:823; that
   * element never influences n because the loop ends at i == n_seeds + 1 */
  ntot = h->sortkey[0];
  for (i = 1; i <= h->n_seeds && ntot <= maxhit; i++) ntot += h->sortkey[i];
  n = nmax = i - 1;
  for (f = 0; f < nskip; f++) {
    const uint32_t imax = h->frame_cnt[f];
    const uint32_t *ixp = h->frame_ix + (size_t) f * h->n_alloc;
    uint32_t cover = 0;
    if (!imax) continue;
    memset(h->qbuf, 0, h->qlen);
    for (i = 0; i < imax && cover <= maxcover && (cover < mincover || ixp[i] <= n); i++) {
      const uint32_t s = h->sidx[ixp[i]];
      uint32_t q;
      for (q = h->qoffs[s]; q < h->qoffs[s] + (uint32_t) ktup - 1; q++)
	if (!h->qbuf[q]) { h->qbuf[q] = 1; cover++; }
    }
    if (i > 0 && ixp[i - 1] > nmax) nmax = ixp[i - 1];
  }
  if (nmax < 3) h->seed_rank = (3 < h->n_seeds) ? 3 : h->n_seeds; /* HITINFO_MINSEEDNUM */
  else h->seed_rank = nmax;
  return SO_SUCCESS;
}

enum { HQ_TERM = 0, HQ_NORMHIT = 1, HQ_MULTIHIT = 2, HQ_REPEAT = 3, HQ_NOHIT = 4, HQ_NONSTDNT = 5 };
enum { HI_REVERSE = 1, HI_SORTED = 2, HI_RANK = 4 };

/* collectHitInfo (hashhit.c:480-657) followed by the ranking of
 * hashCollectHitInfoShort (hashhit.c:1007-1080) */
int so_collect_hitinfo(so_hitinfo *h, const so_index *ix, const uint8_t *read,
		       const uint8_t *qual, uint32_t qlen, int is_reverse, int is_short,
		       uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq_thresh)
{
  const int ktup = ix->wordlen, nskip = ix->nskip;
  const uint8_t minqval = (uint8_t) (basq_thresh + 0x21);
  const int rc_addpos = (ktup - 1) << 1;
  const uint64_t wordmask = (1ULL << (ktup << 1)) - 1;
  uint64_t word = 0;
  int64_t hist[4] = { -1, -2, -3, -4 }; /* initRepeatFilter hashhit.c:342-346 */
  uint32_t s, tuplectr, seedctr = 0, non_std = 0, mincover, maxcover;
  const uint32_t maxhit = is_short ? maxhit_per_tuple : 0;
  int errcode;

  h->status = 0;
  if (qlen < (uint32_t) ktup) return SO_SHORTSEQ;
  if (qlen + 1 >= h->n_alloc) return SO_NOMEM;
  if (is_reverse) h->status |= HI_REVERSE;
  h->qlen = qlen;
  h->n_seeds = 0;
  for (s = 0, tuplectr = 0; s < qlen; s++) {
    const uint8_t c = read[s];
    if ((c & 4) || (qual && qual[s] < minqval)) non_std = (uint32_t) ktup;
    else if (non_std) non_std--;
    if (is_reverse) word = (word >> 2) + ((uint64_t) ((c ^ 3) & 3) << rc_addpos);
    else word = (word << 2) + (c & 3);
    if (s + 1 < (uint32_t) ktup) continue;
    /* k-mer starting at tuplectr = s - ktup + 1 is complete */
    {
      uint32_t posidx = 0, nhits;
      const int64_t w = (int64_t) (word & wordmask);
      int rep;
      if (non_std) { h->qmask[tuplectr++] = HQ_NONSTDNT; continue; }
      rep = (w == hist[0] || w == hist[1] || w == hist[2] || w == hist[3]);
      hist[3] = hist[2]; hist[2] = hist[1]; hist[1] = hist[0]; hist[0] = w;
      if (rep) { h->qmask[tuplectr++] = HQ_REPEAT; continue; }
      nhits = so_lookup(ix, word, &posidx);
      if (nhits < 1) { h->qmask[tuplectr++] = HQ_NOHIT; continue; }
      if (maxhit > 0 && nhits > maxhit) { h->qmask[tuplectr++] = HQ_MULTIHIT; continue; }
      h->sortkey[seedctr] = nhits;
      h->qmask[tuplectr] = HQ_NORMHIT;
      h->posidx[seedctr] = posidx;
      h->nhits[seedctr] = nhits;
      h->cix[seedctr] = 0;
      h->qoffs[seedctr] = tuplectr;
      h->sidx[seedctr] = seedctr;
      seedctr++;
      tuplectr++;
    }
  }
  for (; tuplectr < qlen; tuplectr++) h->qmask[tuplectr] = HQ_TERM;
  h->n_seeds = seedctr;
  h->seed_rank = 0;
  if (!is_short) return SO_SUCCESS;

  if (h->n_seeds <= 1) {
    h->status |= HI_SORTED;
    h->seed_rank = h->n_seeds;
    return SO_SUCCESS;
  }
  if ((errcode = so_sort2(h->n_seeds, h->sortkey, h->sidx))) return errcode;
  h->status |= HI_SORTED;
  mincover = 2 * (uint32_t) ktup + (uint32_t) nskip;     /* HITINFO_MINCOVER_KMER */
  maxcover = qlen * 80 / 100;                             /* HITINFO_MAXCOVER_PERCENT */
  if (maxcover < (uint32_t) (ktup + nskip)) maxcover = (uint32_t) (ktup + nskip);
  else if (maxcover > qlen - (uint32_t) nskip) maxcover = qlen - (uint32_t) nskip;
  if (mincover > maxcover) { mincover = 0; maxcover = qlen; }
  if ((errcode = max_rank(h, ktup, nskip, mincover, maxcover, maxhit_total))) return errcode;
  h->status |= HI_RANK;
  return SO_SUCCESS;
}

/* hashCalcHitInfoCoverDeficit (hashhit.c:1096-1169) */
uint32_t so_cover_deficit(const so_hitinfo *h, int ktup, int nskip)
{
  uint32_t deficit, d, i;
  int s;
  if (h->status & HI_RANK) {
    uint32_t maxcover = 0;
    d = h->qlen;
    for (s = 0; s < nskip; s++) {
      const uint32_t imax = h->frame_cnt[s];
      const uint32_t *ixp = h->frame_ix + (size_t) s * h->n_alloc;
      uint32_t cover = 0, q;
      if (!imax) continue;
      memset(h->qbuf, 0, h->qlen);
      for (i = 0; i < imax && ixp[i] < h->seed_rank; i++) {
	const uint32_t sd = h->sidx[ixp[i]];
	for (q = h->qoffs[sd]; q < h->qoffs[sd] + (uint32_t) ktup; q++)
	  if (!h->qbuf[q]) { h->qbuf[q] = 1; cover++; }
      }
      if (cover < d) d = cover;
      if (cover > maxcover) maxcover = cover;
    }
    deficit = maxcover - d + 1;
  } else {
    uint8_t k = (uint8_t) (ktup / nskip), ctr;
    if (k > 0) k--;
    deficit = 0;
    for (s = 0; s < nskip; s++) {
      d = 0;
      for (ctr = 0, i = (uint32_t) s; i < h->qlen; i += (uint32_t) nskip) {
	if (h->qmask[i] == HQ_NORMHIT) ctr = k;
	else if (ctr) ctr--;
	else d += (uint32_t) nskip;
      }
      if (d > deficit) deficit = d;
    }
  }
  return deficit;
}

/* hashCalcHitInfoNumberOfHits (hashhit.c:1171-1198) */
uint32_t so_number_of_hits(const so_hitinfo *h, uint32_t maxhit_per_tuple)
{
  uint32_t i, hnum = 0;
  for (i = 0; i < h->n_seeds; i++)
    if (maxhit_per_tuple < 1 || h->sortkey[i] <= maxhit_per_tuple) hnum += h->sortkey[i];
  return hnum;
}

/* hashHitInfoCalcHitNumbers (hashhit.c:1200-1220) */
uint32_t so_hit_numbers(const so_hitinfo *h, uint32_t *nhit_rank)
{
  const uint32_t ns = (h->seed_rank > 0) ? h->seed_rank : h->n_seeds;
  uint32_t i, nr = 0;
  for (i = 0; i < ns; i++) nr += h->sortkey[i];
  *nhit_rank = nr;
  for (; i < h->n_seeds; i++) nr += h->sortkey[i];
  return nr;
}

so_hitlist *so_hitlist_create(int maxnhits)
{
  so_hitlist *l = (so_hitlist *) calloc(1, sizeof(*l));
  if (!l) return NULL;
  if (maxnhits < 8192) maxnhits = 8192;                  /* HITLST_MINSIZ */
  l->sqdat = (uint64_t *) calloc((size_t) maxnhits, 8);
  l->qmask = (uint8_t *) calloc(1 << 16, 1);
  l->nhits_max = l->nhits_alloc = maxnhits;
  return l;
}

void so_hitlist_delete(so_hitlist *l)
{
  if (l) { free(l->sqdat); free(l->qmask); }
  free(l);
}

/* initHitList (hashhit.c:1262-1296): capacity qlen*ln(qlen)*32 in [8192, INT_MAX];
 * the allocation only grows, in blocks of 16384 (reallocHitList :1232-1248). */
static int hitlist_init(so_hitlist *l, const so_hitinfo *h)
{
  size_t target = (size_t) (h->qlen * log((double) h->qlen) * 32);
  if (target > (size_t) INT_MAX) target = INT_MAX;
  else if (target < 8192) target = 8192;
  if ((int) target > l->nhits_alloc) {
    size_t nsiz = (target + 16384 - 1) / 16384 * 16384;
    uint64_t *p;
    if (nsiz > (size_t) INT_MAX) return SO_OVERFLOW;
    p = (uint64_t *) realloc(l->sqdat, nsiz * 8);
    if (!p) return SO_NOMEM;
    l->sqdat = p;
    l->nhits_alloc = (int) nsiz;
  }
  if (h->qlen >= (1u << 16)) return SO_NOMEM;
  l->qlen = h->qlen;
  l->nhits_max = (int) target;
  l->nhits = 0;
  l->status = 0;
  memset(l->qmask, HQ_NOHIT, l->qlen);                   /* blankHitList :1224 */
  if (h->status & HI_REVERSE) l->status |= 1;
  return SO_SUCCESS;
}

static uint64_t pack_hit(int is_reverse, uint32_t pos, uint32_t q, int nskip)
{
  /* SET_NEXT_SHIFT (hashhit.c:283-288), HASHHIT_HALFBIT = 31 */
  if (is_reverse) return (((uint64_t) pos + q / (uint32_t) nskip) << 31) + q;
  return ((((uint64_t) pos | (1ULL << 32)) - q / (uint32_t) nskip) << 31) + q;
}

/* fillHitListFromHitInfoSegment (hashhit.c:1416-1546), unfiltered branch */
static int fill_segment(so_hitlist *l, so_hitinfo *h, const so_index *ix,
			uint32_t pos_lo, uint32_t pos_hi, uint32_t maxhit, int use_short)
{
  const int is_reverse = h->status & HI_REVERSE;
  const uint32_t n_seeds = (use_short && h->seed_rank > 0) ? h->seed_rank : h->n_seeds;
  uint32_t n;
  int errcode;
  if ((errcode = hitlist_init(l, h))) return errcode;
  for (n = 0; n < n_seeds; n++) {
    const uint32_t sd = use_short ? h->sidx[n] : n;
    const uint32_t *posp;
    uint32_t nhits, nh, i;
    uint64_t *out;
    if (maxhit > 0 && h->sortkey[n] > maxhit) {
      h->qmask[h->qoffs[sd]] = HQ_MULTIHIT;
      continue;
    }
    nhits = fetch_positions(ix, h->posidx[sd], &posp);
    if (h->cix[sd] >= nhits) {
      if (posp[nhits - 1] < pos_lo) continue;
      h->cix[sd] = 0;
    }
    if (posp[h->cix[sd]] > pos_lo) h->cix[sd] = 0;
    posp += h->cix[sd];
    nh = nhits - h->cix[sd];
    for (i = 0; i < nh && posp[i] < pos_lo; i++);
    nh -= i;
    h->cix[sd] += i;
    posp += i;
    if ((uint32_t) l->nhits + nh > (uint32_t) l->nhits_alloc) {
      if (maxhit > 0) return SO_ALLOCBOUNDARY;
      h->qmask[h->qoffs[sd]] = HQ_MULTIHIT;
      continue;
    }
    out = l->sqdat + l->nhits;
    for (i = 0; i < nh && posp[i] < pos_hi; i++)
      out[i] = pack_hit(is_reverse, posp[i], h->qoffs[sd], ix->nskip);
    h->cix[sd] += i;
    l->nhits += (int) i;
  }
  return SO_SUCCESS;
}

/* hashCollectHitsForSegment (hashhit.c:1691-1769) */
int so_collect_hits_segment(so_hitlist *l, so_hitinfo *h, const so_index *ix,
			    uint64_t lo, uint64_t hi, uint32_t nhit_max, int use_short)
{
  int errcode;
  lo /= (uint64_t) ix->nskip;
  if (lo > 0xFFFFFFFFull) return SO_ARGRANGE;
  hi /= (uint64_t) ix->nskip;
  if (hi > 0xFFFFFFFFull) hi = 0xFFFFFFFFull;
  do {
    errcode = fill_segment(l, h, ix, (uint32_t) lo, (uint32_t) hi, nhit_max, use_short);
    nhit_max /= 2;
  } while (errcode == SO_ALLOCBOUNDARY && nhit_max > 16); /* MINHIT_PER_TUPLE */
  if (errcode && errcode != SO_ALLOCBOUNDARY) return errcode;
  so_sort64((uint32_t) l->nhits, l->sqdat);
  l->status |= 2;
  return SO_SUCCESS;
}

/* hashCollectHitsUsingCutoff (hashhit.c:1593-1689) */
int so_collect_hits_cutoff(so_hitlist *l, const so_hitinfo *h, const so_index *ix,
			   uint32_t max_nhit_per_tup)
{
  const uint32_t n_seeds = h->seed_rank ? h->seed_rank : h->n_seeds;
  int errcode, reached;
  if ((errcode = hitlist_init(l, h))) return errcode;
  do {
    uint32_t i;
    reached = 0;
    l->nhits = 0;
    l->status = (h->status & HI_REVERSE) ? 1 : 0;
    memset(l->qmask, HQ_NOHIT, l->qlen);
    for (i = 0; i < n_seeds; i++) {
      const uint32_t nh = h->sortkey[i], sd = h->sidx[i], q = h->qoffs[sd];
      const uint32_t *posp;
      uint32_t j;
      if (nh < 1) continue;
      if (max_nhit_per_tup > 0 && nh > max_nhit_per_tup) { l->qmask[q] = HQ_MULTIHIT; continue; }
      if ((uint64_t) l->nhits + nh > (uint64_t) INT_MAX) return SO_OVERFLOW;
      if ((int) (l->nhits + nh) > l->nhits_max) { reached = 1; break; }
      if (fetch_positions(ix, h->posidx[sd], &posp) != nh) return SO_ASSERT;
      l->qmask[q] = HQ_NORMHIT;
      for (j = 0; j < nh; j++)
	l->sqdat[l->nhits + j] = pack_hit(l->status & 1, posp[j], q, ix->nskip);
      l->nhits += (int) nh;
    }
    max_nhit_per_tup /= 2;
  } while (reached && max_nhit_per_tup > 16);
  so_sort64((uint32_t) l->nhits, l->sqdat);
  l->status |= 2;
  return SO_SUCCESS;
}
