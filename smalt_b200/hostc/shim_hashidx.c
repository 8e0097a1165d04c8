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
#include "hashidx.c"

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
