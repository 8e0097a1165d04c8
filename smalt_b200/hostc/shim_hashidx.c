/* shim_hashidx.c - hash index object for the smalt_b200 driver build.
 *
 * Index construction and `.smi` file I/O are NOT on the hot path and stay the reference's
 * CPU code: this translation unit compiles the
 * reference's hashidx.c in place (from the read-only tree on the include path, nothing is
 * copied) and adds ONE accessor that exposes the table arrays so that they can be uploaded
 * to the GPU once (smb_index_upload).  The CPU lookup functions of hashidx.c
 * (hashTableGetKtupleHits, hashTableFetchHitPositions) end up in the object but are never
 * called by the B200 path: all lookups run in seed.cu.
 */
#include <stdlib.h>
#define hashTableSetUp hashTableSetUp_cpu
#include "hashidx.c"
#undef hashTableSetUp

int smbShimIndexBuild(const SeqSet *ssp, const SeqCodec *codecp, int k, int nskip, int typ, int nbits_key,
		      int nbits_lo, uint32_t *npos, uint32_t *nwords, uint32_t *tuplectr_out, float *ms);
int smbShimIndexFetch(uint32_t *idx, uint32_t *pos, uint32_t *wordidx, uint32_t *posidx);

/* hashTableSetUp (hashidx.c:829-998).  The index of a whole sequence set (`smalt index`) is
 * built on the GPU (csrc/index_build.cu) into the same arrays with the same allocation rules, so
 * that hashTableWrite produces the same `.smi` bytes.  The on-the-fly k=5 index of a few intervals
 * (rmap.c:495-517, ~12 KB per pair) is the reference's own builder, as is SMALT_B200_CPU_INDEX=1 (an
 * explicit switch for comparing the two); without a CUDA device the call fails. */
int hashTableSetUp(HashTable *htp, SeqFastq *sqbufp, const SeqSet *ssp, const InterVal *ivp,
		   const SeqCodec *codecp, uint32_t *npos_max, char verbose)
{
  uint32_t npos = 0, nwords = 0, tuplectr = 0;
  float ms = 0.f;
  SETSIZ_t totlen;
  if (ivp != NULL || getenv("SMALT_B200_CPU_INDEX"))
    return hashTableSetUp_cpu(htp, sqbufp, ssp, ivp, codecp, npos_max, verbose);
  seqSetGetSeqNumAndTotLen(&totlen, ssp);
  if (htp->status == HASHSETUP_EMPTY) {
    if (htp->nskip < 1 || (totlen + 1) / htp->nskip > HASHPOS_MAX) return ERRCODE_ASSERT;
  } else {
    hashTableReset(htp, 0);
  }
  if (smbShimIndexBuild(ssp, codecp, htp->wordlen, htp->nskip, htp->typ, htp->nbits_key, htp->nbits_lo,
			&npos, &nwords, &tuplectr, &ms) < 0) {
    /* no silent fallback: the index of a sequence set is built on the GPU or not at all
     * (SMALT_B200_CPU_INDEX=1 asks for the reference's builder explicitly, e.g. to compare the two) */
    fprintf(stderr, "smalt_b200: index construction needs a CUDA device (there is no CPU fallback)\n");
    return ERRCODE_FAILURE;
  }
  if (verbose) fprintf(stderr, "# Hash index built on the GPU (%u positions, %u words, %.1f ms).\n", npos, nwords, ms);
  if (npos_max != NULL) {   /* hashidx.c:881-887 */
    if (*npos_max > 0 && npos > *npos_max) {
      *npos_max = npos;
      return ERRCODE_MAXKPOS;
    }
    *npos_max = npos;
  }
  if (htp->pos == NULL || npos > htp->npos_alloc) {   /* hashidx.c:889-903 */
    size_t n_alloc = ((size_t) npos + 1) / BLKSIZ_IDXPOS + 1;
    n_alloc *= BLKSIZ_IDXPOS;
    if (htp->pos == NULL) {
      ECALLOCP(n_alloc, htp->pos);
      if (!htp->pos) return ERRCODE_NOMEM;
    } else {
      void *hp = EREALLOCP(htp->pos, n_alloc);
      if (!hp) return ERRCODE_NOMEM;
      htp->pos = hp;
    }
    htp->npos_alloc = n_alloc;
  }
  htp->npos = npos;
  if (htp->typ != HASHIDXTYP_PERFECT) {   /* hashidx.c:935-940 */
    htp->nwords = nwords;
    free(htp->wordidx);
    if ((ECALLOCP(2 * ((size_t) nwords + ARRAY_MARGIN), htp->wordidx)) == NULL) return ERRCODE_NOMEM;
    htp->posidx = htp->wordidx + nwords + ARRAY_MARGIN;
  }
  if (smbShimIndexFetch(htp->idx, htp->pos, htp->typ != HASHIDXTYP_PERFECT ? htp->wordidx : NULL,
			htp->typ != HASHIDXTYP_PERFECT ? htp->posidx : NULL))
    return ERRCODE_FAILURE;
  htp->maxpos = (tuplectr > 0) ? tuplectr - 1 : 0;
  htp->status = HASHSETUP_COMPLETE;
  if (verbose) fprintf(stderr, "# Hash table is set up.\n");
  return ERRCODE_SUCCESS;
}

void smbShimHashTableArrays(const HashTable *htp, int *typ, int *wordlen, int *nskip,
			    int *nbits_key, int *nbits_lo, uint32_t *npos, uint32_t *nwords,
			    const uint32_t **idx, const uint32_t **pos,
			    const uint32_t **wordidx, const uint32_t **posidx)
{
  *typ = htp->typ;
  *wordlen = htp->wordlen;
  *nskip = htp->nskip;
  *nbits_key = htp->nbits_key;
  *nbits_lo = htp->nbits_lo;
  *npos = htp->npos;
  *nwords = htp->nwords;
  *idx = htp->idx;
  *pos = htp->pos;
  *wordidx = htp->wordidx;
  *posidx = htp->posidx;
}
