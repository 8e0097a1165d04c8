/* ref_harness_hashhit.c - TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the reference's hashhit.c *in place* (from /root/reference/src via
 * the include path set by oracle/Makefile; nothing is copied) inside this
 * translation unit so that the private struct _HashHitInfo (hashhit.c:164-213)
 * is visible, and adds one accessor that dumps it into plain arrays.  The
 * resulting object replaces hashhit.o in oracle/_ref/libsmalt_ref.so. */
#include "hashhit.c"

int refh_hitinfo_dump(const HashHitInfo *hip, int maxn, uint32_t *n_seeds,
		      uint32_t *seed_rank, uint32_t *posidx, uint32_t *nhits,
		      uint32_t *qoffs, uint32_t *sortkey, uint32_t *sidx,
		      unsigned char *qmask, unsigned char *status)
{
  uint32_t i;
  *n_seeds = hip->n_seeds;
  *seed_rank = hip->seed_rank;
  *status = hip->status;
  if ((int) hip->n_seeds > maxn || (int) hip->qlen > maxn) return ERRCODE_OVERFLOW;
  for (i = 0; i < hip->n_seeds; i++) {
    posidx[i] = hip->seedp[i].posidx;
    nhits[i] = hip->seedp[i].nhits;
    qoffs[i] = hip->seedp[i].qoffs;
    sortkey[i] = hip->nhitqual_sortkeyp[i];
    sidx[i] = hip->sidxp[i];
  }
  for (i = 0; i < hip->qlen; i++) qmask[i] = hip->qmaskp[i];
  return ERRCODE_SUCCESS;
}
