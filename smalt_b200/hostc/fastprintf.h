/* fastprintf.h - drop-in for fprintf inside the reference's output formatting code.
 *
 * The reference formats every SAM/CIGAR record with a handful of fprintf calls
 * (report.c:192-194 OUFMT_SAM_BEFORE/AFTER, diffstr.c:159-165 one call per CIGAR operation).
 * The block-parallel driver formats records in its worker threads; compiling report.c and
 * diffstr.c with -Dfprintf=smbFastFprintf routes those calls - and only those - to
 * smbFastFprintf, which appends to a per-thread capture buffer when the stream is the one
 * being captured and is a plain vfprintf otherwise.  Same bytes, no format re-parsing by
 * libc, no FILE locking.  Conversions handled natively: %s %c %% and the decimal integer
 * conversions without flags/width (%d %i %u with h, l, ll); anything else is formatted by
 * vsnprintf into the same buffer.
 */
#ifndef SMALT_B200_FASTPRINTF_H
#define SMALT_B200_FASTPRINTF_H
#include <stdio.h>
int smbFastFprintf(FILE *fp, const char *fmt, ...);
/* everything printed to `key` by this thread until smbFastCaptureEnd goes to a memory buffer */
void smbFastCaptureBegin(FILE *key);
/* ends the capture; *buf is malloc'ed (caller frees), *len its length */
int smbFastCaptureEnd(char **buf, size_t *len);
/* direct appends for callers that know what they print (shim_report.c): `n` bytes of room in the
 * capture buffer of this thread when `fp` is the captured stream, NULL otherwise (or when out of
 * memory); smbFastCommit makes the first `used` bytes part of the output */
char *smbFastReserve(FILE *fp, size_t n);
void smbFastCommit(size_t used);
size_t smbFastPutInt(char *p, long long v);
#endif
