/* oracle/ref_shim.c — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * The five host symbols the reference's vendored htscodecs translation units need when they
 * are compiled, unmodified, from where they lie under /root/reference/src/htscodecs
 * (SURVEY.md §8c / Appendix C): codec_alloc_do, codec_free_do (reference src/codec.c:30-63),
 * buf_low_level_malloc, buf_low_level_free (src/buf_struct.h:222-227) and error_assert_failed
 * (src/genozip.h:747).  A plain malloc shim is sufficient: the VBlockP argument is passed through
 * opaque by htscodecs and is NULL here.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdbool.h>
#include <string.h>

void *codec_alloc_do (void *vb, uint64_t size, float grow, unsigned *buf_i, const char *func, uint32_t line)
{
    (void)vb; (void)grow; (void)func; (void)line;
    if (buf_i) *buf_i = 0;
    return malloc (size ? size : 1);
}

void codec_free_do (void *vb, void *addr, const char *func, uint32_t line)
{
    (void)vb; (void)func; (void)line;
    free (addr);
}

void *buf_low_level_malloc (size_t size, bool zero, const char *func, uint32_t line)
{
    (void)func; (void)line;
    return zero ? calloc (1, size ? size : 1) : malloc (size ? size : 1);
}

void buf_low_level_free (void *p, const char *func, uint32_t line)
{
    (void)func; (void)line;
    free (p);
}

void error_assert_failed (const char *func, uint32_t line, const char *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    fprintf (stderr, "oracle/_ref: reference assertion failed in %s:%u: ", func, line);
    vfprintf (stderr, fmt, ap);
    fprintf (stderr, "\n");
    va_end (ap);
    abort ();
}
