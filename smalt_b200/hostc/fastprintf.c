/* fastprintf.c - see fastprintf.h */
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "fastprintf.h"

#undef fprintf

static __thread struct {
  FILE *fp;
  char *buf;
  size_t len, cap;
  int failed;
} t_cap;

static int cap_reserve(size_t extra)
{
  if (t_cap.len + extra + 1 > t_cap.cap) {
    size_t nc = t_cap.cap ? t_cap.cap : (size_t) 1 << 16;
    char *hp;
    while (nc < t_cap.len + extra + 1) nc *= 2;
    if (!(hp = (char *) realloc(t_cap.buf, nc))) { t_cap.failed = 1; return -1; }
    t_cap.buf = hp;
    t_cap.cap = nc;
  }
  return 0;
}

void smbFastCaptureBegin(FILE *key)
{
  t_cap.fp = key;
  t_cap.buf = NULL;
  t_cap.len = t_cap.cap = 0;
  t_cap.failed = 0;
}

int smbFastCaptureEnd(char **buf, size_t *len)
{
  const int failed = t_cap.failed;
  *buf = t_cap.buf;
  *len = t_cap.len;
  t_cap.fp = NULL;
  t_cap.buf = NULL;
  t_cap.len = t_cap.cap = 0;
  return failed ? -1 : 0;
}

static size_t put_u64(char *p, unsigned long long v)
{
  char tmp[24];
  size_t n = 0, i;
  do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
  for (i = 0; i < n; i++) p[i] = tmp[n - 1 - i];
  return n;
}

static size_t put_i64(char *p, long long v)
{
  if (v < 0) { *p = '-'; return 1 + put_u64(p + 1, 0ULL - (unsigned long long) v); }
  return put_u64(p, (unsigned long long) v);
}

char *smbFastReserve(FILE *fp, size_t n)
{
  if (!fp || fp != t_cap.fp || cap_reserve(n)) return NULL;
  return t_cap.buf + t_cap.len;
}

void smbFastCommit(size_t used) { t_cap.len += used; }

size_t smbFastPutInt(char *p, long long v) { return put_i64(p, v); }

int smbFastFprintf(FILE *fp, const char *fmt, ...)
{
  va_list ap, ap0;
  const char *f;
  size_t start;
  int rv;
  va_start(ap, fmt);
  if (fp != t_cap.fp || !fp) {
    rv = vfprintf(fp, fmt, ap);
    va_end(ap);
    return rv;
  }
  /* one CIGAR operation (diffstr.c:65 CIGAR_EXTF): the call that is made most often */
  if (fmt[0] == '%' && fmt[1] == 'd' && fmt[2] == '%' && fmt[3] == 'c' && !fmt[4]) {
    const int v = va_arg(ap, int), c = va_arg(ap, int);
    size_t n;
    va_end(ap);
    if (cap_reserve(16)) return -1;
    n = put_i64(t_cap.buf + t_cap.len, v);
    t_cap.buf[t_cap.len + n] = (char) c;
    t_cap.len += n + 1;
    return (int) n + 1;
  }
  va_copy(ap0, ap);
  start = t_cap.len;
  for (f = fmt; *f; f++) {
    int lmod = 0; /* -1: h, 1: l, 2: ll */
    if (*f != '%') {
      const char *e = strchr(f, '%');
      const size_t n = e ? (size_t) (e - f) : strlen(f);
      if (cap_reserve(n)) goto fail;
      memcpy(t_cap.buf + t_cap.len, f, n);
      t_cap.len += n;
      f += n - 1;
      continue;
    }
    f++;
    if (*f == 'h') { lmod = -1; f++; }
    else if (*f == 'l') { lmod = 1; f++; if (*f == 'l') { lmod = 2; f++; } }
    switch (*f) {
    case '%':
      if (cap_reserve(1)) goto fail;
      t_cap.buf[t_cap.len++] = '%';
      break;
    case 'c':
      if (cap_reserve(1)) goto fail;
      t_cap.buf[t_cap.len++] = (char) va_arg(ap, int);
      break;
    case 's': {
      const char *s = va_arg(ap, const char *);
      size_t n;
      if (!s) s = "(null)";
      n = strlen(s);
      if (cap_reserve(n)) goto fail;
      memcpy(t_cap.buf + t_cap.len, s, n);
      t_cap.len += n;
      break;
    }
    case 'd': case 'i': {
      long long v;
      if (lmod == 2) v = va_arg(ap, long long);
      else if (lmod == 1) v = va_arg(ap, long);
      else if (lmod == -1) v = (short) va_arg(ap, int);
      else v = va_arg(ap, int);
      if (cap_reserve(24)) goto fail;
      t_cap.len += put_i64(t_cap.buf + t_cap.len, v);
      break;
    }
    case 'u': {
      unsigned long long v;
      if (lmod == 2) v = va_arg(ap, unsigned long long);
      else if (lmod == 1) v = va_arg(ap, unsigned long);
      else if (lmod == -1) v = (unsigned short) va_arg(ap, unsigned int);
      else v = va_arg(ap, unsigned int);
      if (cap_reserve(24)) goto fail;
      t_cap.len += put_u64(t_cap.buf + t_cap.len, v);
      break;
    }
    default: { /* flags, widths, floats, ...: let libc format the whole call */
      int n;
      t_cap.len = start;
      n = vsnprintf(NULL, 0, fmt, ap0);
      if (n < 0 || cap_reserve((size_t) n)) goto fail;
      va_end(ap0);
      va_end(ap);
      va_start(ap, fmt);
      vsnprintf(t_cap.buf + t_cap.len, (size_t) n + 1, fmt, ap);
      va_end(ap);
      t_cap.len += (size_t) n;
      return n;
    }
    }
  }
  va_end(ap0);
  va_end(ap);
  return (int) (t_cap.len - start);
fail:
  va_end(ap0);
  va_end(ap);
  return -1;
}
