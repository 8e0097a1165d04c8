/* Checks smbFastFprintf (smalt_b200/hostc/fastprintf.c) against libc's snprintf for the formats the
 * reference's output code uses (report.c:192-194, diffstr.c:64-65) and a few that take the fallback. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "fastprintf.h"

static int nfail;

#define CHECK(fmt, ...)                                                              \
  do {                                                                               \
    char want[512], *got; size_t len;                                                \
    int nw = snprintf(want, sizeof(want), fmt, __VA_ARGS__), ng;                     \
    smbFastCaptureBegin(stdout);                                                     \
    ng = smbFastFprintf(stdout, fmt, __VA_ARGS__);                                   \
    if (smbFastCaptureEnd(&got, &len) || ng != nw || len != (size_t) nw || memcmp(got, want, len)) { \
      fprintf(stderr, "MISMATCH for \"%s\": want \"%s\" (%d), got \"%.*s\" (%d)\n", fmt, want, nw, (int) len, got ? got : "", ng); \
      nfail++;                                                                       \
    }                                                                                \
    free(got);                                                                       \
  } while (0)

int main(void)
{
  static const int ints[] = {0, 1, -1, 9, 10, 61, 62, 255, 4096, 65535, -32768, INT_MAX, INT_MIN};
  size_t i;
  for (i = 0; i < sizeof(ints) / sizeof(ints[0]); i++) {
    CHECK("%d%c", ints[i], 'M');
    CHECK("%c %d ", 'I', ints[i]);
    CHECK("%s\t%hu\t%s\t%i\t%hi\t", "read/1", (unsigned short) ints[i], "chr1", ints[i], (short) ints[i]);
    CHECK("\t%s\t%i\t%i\t%s\t%s\tNM:i:%i\tAS:i:%i\n", "*", 0, 0, "ACGT", "IIII", ints[i], -ints[i] / 2);
    CHECK("%u %lu %llu %ld %lld", (unsigned) ints[i], (unsigned long) ints[i], (unsigned long long) ints[i],
	  (long) ints[i], (long long) ints[i]);
    CHECK("%5d|%-4d|%03d", ints[i], ints[i] % 100, ints[i] % 10);   /* widths: libc formats the call */
    CHECK("%.3f %s", ints[i] / 7.0, "x");
  }
  CHECK("%s%%%c", "", 'x');
  { /* direct appends */
    char *got, *o; size_t len, n;
    smbFastCaptureBegin(stdout);
    smbFastFprintf(stdout, "%s", "ab");
    o = smbFastReserve(stdout, 64);
    if (!o) { fprintf(stderr, "no room\n"); nfail++; }
    else {
      n = smbFastPutInt(o, -12345);
      o[n++] = '\t';
      n += smbFastPutInt(o + n, 0);
      smbFastCommit(n);
    }
    smbFastFprintf(stdout, "%d%c", 7, 'S');
    if (smbFastReserve(stderr, 8)) { fprintf(stderr, "reserve on a stream that is not captured\n"); nfail++; }
    if (smbFastCaptureEnd(&got, &len) || len != 12 || memcmp(got, "ab-12345\t07S", 12) != 0) {
      fprintf(stderr, "direct append: got \"%.*s\" (%zu)\n", (int) len, got ? got : "", len);
      nfail++;
    }
    free(got);
  }
  if (nfail) return 1;
  puts("ok");
  return 0;
}
