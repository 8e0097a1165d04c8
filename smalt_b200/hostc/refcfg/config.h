/* Minimal hand-written configuration for compiling the read-only reference
 * sources under /root/reference/src directly with gcc (smalt_b200/hostc/Makefile (and oracle/Makefile)).
 * Used only to compile reference translation units; defines no behaviour of its own. */
#ifndef SMALT_ORACLE_CONFIG_H
#define SMALT_ORACLE_CONFIG_H
#define HAVE_EMMINTRIN_H 1   /* SSE2 striped Smith-Waterman path (swsimd.c) */
#define HAVE_FLOAT_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_MATH_H 1
#define HAVE_MEMORY_H 1
#define HAVE_PTHREAD_H 1
#define HAVE_SEMAPHORE_H 1
#define HAVE_STDDEF_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRINGS_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_ZLIB 1
#define HAVE_ZLIB_H 1
#define STDC_HEADERS 1
#define PACKAGE "smalt"
#define PACKAGE_NAME "smalt"
#define PACKAGE_VERSION "0.7.6"
#define PACKAGE_STRING "smalt 0.7.6"
#define PACKAGE_BUGREPORT "hp3@sanger.ac.uk"
#define VERSION "0.7.6"
#endif
