/* oracle/hts_port.c — CPU restatement of the htscodecs "4x16" rANS and adaptive arithmetic coders as
 * vendored (and modified) by genozip.  TEST INFRASTRUCTURE ONLY — see oracle/oracle.h.
 *
 * Parity: PINNED.  tests/test_oracle_hts.py compares every entry point byte-for-byte, both directions,
 * against oracle/_ref/libhts_ref.so (the reference's own translation units compiled unmodified).
 *
 * This is a restatement, not a copy: it is organised around "leaf" encoders writing into caller-sized
 * scratch, with the container framing separate.  Each function cites the reference lines it follows
 * (relative to /root/reference/src/htscodecs).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include "oracle.h"

enum { F_ORDER = 1, F_STRIPE = 8, F_NOSZ = 16, F_CAT = 32, F_RLE = 64, F_PACK = 128 };   /* rANS_static4x16.h:38-44 */
#define RANS_L (1u << 15)                                                                /* rANS_word.h:58 */

/* ------------------------------------------------------------------ varint (varint.h:180-300, big-endian base-128) */
static int vput (uint8_t *p, uint32_t v)
{
    int nb = 1;
    for (uint32_t t = v >> 7; t; t >>= 7) nb++;
    for (int k = nb - 1; k >= 0; k--)
        *p++ = (uint8_t)(((v >> (7 * k)) & 0x7f) | (k ? 0x80 : 0));
    return nb;
}

static int vget (const uint8_t *p, const uint8_t *end, uint32_t *v)
{
    const uint8_t *s = p;
    uint32_t acc = 0;
    if (p >= end) { *v = 0; return 0; }
    int limit = 6;                                         /* varint.h:273-279: at most 6 bytes consumed */
    uint8_t c;
    do { c = *p++; acc = (acc << 7) | (c & 0x7f); } while ((c & 0x80) && p < end && --limit > 0);
    *v = acc;
    return (int)(p - s);
}

/* ------------------------------------------------------------------ bounds */
/* The reference evaluates these in double, left to right; gcc -O3 -march=haswell contracts only the
 * leading 1.05*size + C into one fma (seen as vfmadd132sd in the reference object). */
static double bound_base (uint32_t n, int order)
{
    if (order == 0) return fma (1.05, (double)n, 257*3) + 4;
    double t = fma (1.05, (double)n, 257*257*3);
    t += 4; t += 257*3; t += 4;
    return t;
}

uint32_t orc_rans_bound (uint32_t n, int order)            /* rANS_static4x16pr.c:357-369 */
{
    int N = order >> 8; if (!N) N = 4;
    order &= 0xff;
    int sz = (int)(bound_base (n, order) + ((order & F_PACK) ? 1 : 0) + ((order & F_RLE) ? 1 + 257*3 + 4 : 0) + 20 + ((order & F_STRIPE) ? 1 + 5*N : 0));
    return sz + (sz & 1) + 2;
}

uint32_t orc_arith_bound (uint32_t n, int order)           /* arith_dynamic.c:74-80 */
{
    return (uint32_t)(bound_base (n, order) + ((order & F_PACK) ? 1 : 0) + ((order & F_RLE) ? 1 + 257*3 + 4 : 0) + 5);
}

/* ------------------------------------------------------------------ frequency tables */
static uint32_t pow2_ceil (uint32_t v) { v--; v |= v>>1; v |= v>>2; v |= v>>4; v |= v>>8; v |= v>>16; return v + 1; }  /* :102-111 */

/* :113-160.  Scales F[] (sum `size`) so it sums to `tot`.  Note the reference re-uses `size` as the
 * running scaled sum, so the single retry recomputes the multiplier from the *scaled* sum. */
static int scale_freqs (uint32_t *F, int size, uint32_t tot)
{
    if (!size) return 0;
    int retried = 0, top_sym;
    for (;;) {
        uint64_t mul = ((uint64_t)tot << 31) / size + (1 << 30) / size;
        int top_val = 0; top_sym = 0; size = 0;
        for (int j = 0; j < 256; j++) {
            if (!F[j]) continue;
            if ((uint32_t)top_val < F[j]) { top_val = F[j]; top_sym = j; }      /* arg-max on PRE-scaling counts, first max wins */
            F[j] = (uint32_t)((F[j] * mul) >> 31);
            if (!F[j]) F[j] = 1;
            size += F[j];
        }
        int adjust = (int)tot - size;
        if (adjust > 0) { F[top_sym] += adjust; break; }
        if (adjust == 0) break;
        if (F[top_sym] > (uint32_t)-adjust && (retried || F[top_sym] / 2 >= (uint32_t)-adjust)) { F[top_sym] += adjust; break; }
        if (!retried) { retried = 1; continue; }
        adjust += F[top_sym] - 1;
        F[top_sym] = 1;
        for (int j = 0; adjust && j < 256; j++) {
            if (F[j] < 2) continue;
            int d = F[j] > (uint32_t)-adjust;
            int m = d ? adjust : 1 - (int)F[j];
            F[j] += m; adjust -= m;
        }
        break;
    }
    return F[top_sym] > 0 ? 0 : -1;
}

static void shift_freqs (uint32_t *F, uint32_t size, uint32_t max_tot)          /* :165-176 */
{
    if (size == 0 || size == max_tot) return;
    int sh = 0;
    while (size < max_tot) { size *= 2; sh++; }
    for (int i = 0; i < 256; i++) F[i] <<= sh;
}

/* :179-203 — ascending symbol list; after two adjacent present symbols, a count of how many more follow */
static int put_alphabet (uint8_t *p, const uint32_t *F)
{
    uint8_t *s = p;
    int skip = 0;
    for (int j = 0; j < 256; j++) {
        if (!F[j]) continue;
        if (skip) { skip--; continue; }
        *p++ = (uint8_t)j;
        if (j && F[j-1]) {
            int k = j + 1;
            while (k < 256 && F[k]) k++;
            skip = k - (j + 1);
            *p++ = (uint8_t)skip;
        }
    }
    *p++ = 0;
    return (int)(p - s);
}

static int get_alphabet (const uint8_t *p, const uint8_t *end, uint32_t *F)     /* :205-252 */
{
    if (p >= end) return 0;
    const uint8_t *s = p;
    int run = 0, j = *p++;
    do {                                                   /* do-while: a leading symbol 0 is marked too (:215-228) */
        F[j] = 1;
        if (p >= end) return 0;
        if (!run && j + 1 == *p) {
            if (p + 1 >= end) return 0;
            j = *p++; run = *p++;
        }
        else if (run) { run--; if (++j > 255) return 0; }
        else j = *p++;
    } while (j);
    return (int)(p - s);
}

static int put_freqs_o0 (uint8_t *p, const uint32_t *F)                         /* :254-266 */
{
    uint8_t *s = p;
    p += put_alphabet (p, F);
    for (int j = 0; j < 256; j++) if (F[j]) p += vput (p, F[j]);
    return (int)(p - s);
}

/* :292-322 — per-context frequencies for the symbols of F0, zero runs as (0, run-1) */
static int put_freqs_o1 (uint8_t *p, const uint32_t *F0, const uint32_t *F)
{
    uint8_t *s = p;
    int zrun = 0;
    for (int j = 0; j < 256; j++) {
        if (!F0[j]) continue;
        if (F[j]) {
            if (zrun) { p -= zrun - 1; *p++ = (uint8_t)(zrun - 1); zrun = 0; }
            p += vput (p, F[j]);
        }
        else { zrun++; *p++ = 0; }
    }
    if (zrun) { p -= zrun - 1; *p++ = (uint8_t)(zrun - 1); }
    return (int)(p - s);
}

static int get_freqs_o1 (const uint8_t *p, const uint8_t *end, const uint32_t *F0, uint32_t *F, uint32_t *total) /* :324-355 */
{
    if (p >= end) return 0;
    const uint8_t *s = p;
    uint32_t T = 0; int zrun = 0;
    for (int j = 0; j < 256 && p < end; j++) {
        if (!F0[j]) continue;
        uint32_t f;
        if (zrun) { f = 0; zrun--; }
        else {
            p += vget (p, end, &f);
            if (!f) { if (p >= end) return 0; zrun = *p++; }
        }
        F[j] = f; T += f;
    }
    *total = T;
    return (int)(p - s);
}

/* ------------------------------------------------------------------ rANS encoder primitives (rANS_word.h:169-320) */
typedef struct { uint32_t x_max, rcp, bias, cmpl, sh; } EncSym;

static void encsym_init (EncSym *s, uint32_t start, uint32_t freq, uint32_t bits)
{
    s->x_max = ((RANS_L >> bits) << 16) * freq;
    s->cmpl  = (uint16_t)((1u << bits) - freq);
    if (freq < 2) { s->rcp = ~0u; s->sh = 32; s->bias = start + (1u << bits) - 1; }
    else {
        uint32_t k = 0;
        while (freq > (1u << k)) k++;
        s->rcp  = (uint32_t)(((1ull << (k + 31)) + freq - 1) / freq);
        s->sh   = k - 1 + 32;
        s->bias = start;
    }
}

static inline void enc_put (uint32_t *x, uint8_t **pp, const EncSym *s)
{
    uint32_t v = *x;
    if (v >= s->x_max) { *pp -= 2; (*pp)[0] = (uint8_t)v; (*pp)[1] = (uint8_t)(v >> 8); v >>= 16; }
    uint32_t q = (uint32_t)(((uint64_t)v * s->rcp) >> s->sh);
    *x = v + s->bias + q * s->cmpl;
}

static inline void enc_flush (uint32_t x, uint8_t **pp)
{
    *pp -= 4;
    (*pp)[0] = (uint8_t)x; (*pp)[1] = (uint8_t)(x >> 8); (*pp)[2] = (uint8_t)(x >> 16); (*pp)[3] = (uint8_t)(x >> 24);
}

/* Order-0 block (:376-491).  Writes table+payload to out, returns length.  cap must be >= bound. */
static uint32_t rans_o0_block (const uint8_t *in, uint32_t n, uint8_t *out)
{
    if (!n) return 0;                                                            /* :402-403,484-488 */
    uint32_t F[256] = {0};
    for (uint32_t i = 0; i < n; i++) F[in[i]]++;

    uint32_t tot = pow2_ceil (n); if (tot > 4096) tot = 4096;                     /* :409-414 */
    scale_freqs (F, (int)n, tot);
    uint32_t tab = (uint32_t)put_freqs_o0 (out, F);
    scale_freqs (F, (int)tot, 4096);                                             /* :423 */

    EncSym sy[256];
    for (uint32_t j = 0, x = 0; j < 256; j++) if (F[j]) { encsym_init (&sy[j], x, F[j], 12); x += F[j]; }

    size_t scratch_sz = (size_t)(1.05 * n) + 64;
    uint8_t *scratch = malloc (scratch_sz), *end = scratch + scratch_sz, *p = end;
    uint32_t R[4] = { RANS_L, RANS_L, RANS_L, RANS_L };
    for (uint32_t i = n; i-- > 0; )                                              /* symbol i belongs to state i&3; last first (:439-477) */
        enc_put (&R[i & 3], &p, &sy[in[i]]);
    for (int k = 3; k >= 0; k--) enc_flush (R[k], &p);                           /* :479-482 */
    memcpy (out + tab, p, (size_t)(end - p));
    uint32_t len = tab + (uint32_t)(end - p);
    free (scratch);
    return len;
}

/* :626-687.  FP decision between 10- and 12-bit order-1 tables.  The reference object contains these
 * exact fused operations (verified in the -O3 -march=haswell disassembly): t = fma(d,K,-l); e = fma(-F,t,e); e += c. */
static int choose_shift (const uint32_t *F0, uint32_t (*F)[256], const uint32_t *T, int *S)
{
    const double K = 1.539095918623324e-16;
    double e10 = 0, e12 = 0;
    int max_tot = 0;
    for (int i = 0; i < 256; i++) {
        if (!F0[i]) continue;
        int max_val = (int)pow2_ceil (T[i]);
        int ns = 0, sm10 = 0, sm12 = 0;
        for (int j = 0; j < 256; j++) {
            if (F[i][j] && (uint32_t)max_val / F[i][j] > 1024) sm10++;
            if (F[i][j] && (uint32_t)max_val / F[i][j] > 4096) sm12++;
        }
        double l10 = log (1024 + sm10), l12 = log (4096 + sm12);
        double Td = (double)T[i];
        for (int j = 0; j < 256; j++) {
            if (!F[i][j]) continue;
            ns++;
            double Fd = (double)F[i][j];
            int x = (int)((Fd * 1024.0) / Td); if (x < 1) x = 1;
            union { double d; long long ll; } u; u.d = (double)x;
            double t = fma ((double)(u.ll - 4606921278410026770LL), K, -l10);
            e10 = fma (-Fd, t, e10) + 4;
            x = (int)((Fd * 4096.0) / Td); if (x < 1) x = 1;
            u.d = (double)x;
            t = fma ((double)(u.ll - 4606921278410026770LL), K, -l12);
            e12 = fma (-Fd, t, e12) + 6;
        }
        if (ns < 64 && max_val > 128) max_val /= 2;
        if (max_val > 1024) max_val /= 2;
        if (max_val > 4096) max_val = 4096;
        S[i] = max_val;
        if (max_tot < max_val) max_tot = max_val;
    }
    return (e10 / e12 < 1.01 || max_tot <= 1024) ? 10 : 12;
}

/* Order-1 block (:691-860) */
static uint32_t rans_o1_block (const uint8_t *in, uint32_t n, uint8_t *out)
{
    uint32_t (*F)[256] = calloc (256, sizeof *F);
    EncSym (*sy)[256]  = malloc (256 * sizeof *sy);
    uint32_t T[256] = {0}, F0[256] = {0};
    uint32_t q4 = n >> 2;

    /* hist1_4 (utils.h:136-210): F[prev][cur], prev=0 before the first symbol; then the 3 quarter-start fixups (:730-733) */
    { uint8_t prev = 0; for (uint32_t i = 0; i < n; i++) { F[prev][in[i]]++; T[prev]++; prev = in[i]; } }
    F[0][in[1*q4]]++; F[0][in[2*q4]]++; F[0][in[3*q4]]++; T[0] += 3;

    for (uint32_t i = 0; i < n; i++) F0[in[i]] = 1;
    F0[0] = 1;                                                                   /* :741 */

    uint8_t *hdr = out, *p = out + 1;
    p += put_alphabet (p, F0);

    int S[256] = {0};
    int shift = choose_shift (F0, F, T, S);

    for (int i = 0; i < 256; i++) {
        if (!F0[i]) continue;
        int mv = S[i];
        if (shift == 10 && mv > 1024) mv = 1024;
        scale_freqs (F[i], (int)T[i], (uint32_t)mv);
        p += put_freqs_o1 (p, F0, F[i]);
        shift_freqs (F[i], (uint32_t)mv, 1u << shift);
        for (uint32_t j = 0, x = 0; j < 256; j++) { encsym_init (&sy[i][j], x, F[i][j], (uint32_t)shift); x += F[i][j]; }
    }

    *hdr = (uint8_t)(shift << 4);
    if (p - hdr > 1000) {                                                        /* :779-792 try O0-compressing the table */
        uint32_t usz = (uint32_t)(p - (hdr + 1));
        uint8_t *c = malloc (orc_rans_bound (usz, 0));
        uint32_t csz = rans_o0_block (hdr + 1, usz, c);
        if (csz + 6 < (uint32_t)(p - hdr)) {
            uint8_t *w = hdr;
            *w++ |= 1;
            w += vput (w, usz);
            w += vput (w, csz);
            memcpy (w, c, csz);
            p = w + csz;
        }
        free (c);
    }
    uint32_t tab = (uint32_t)(p - out);

    size_t scratch_sz = (size_t)(1.05 * n) + 64;
    uint8_t *scratch = malloc (scratch_sz), *end = scratch + scratch_sz, *w = end;
    uint32_t R[4] = { RANS_L, RANS_L, RANS_L, RANS_L };

    /* chain k covers in[k*q4 .. (k+1)*q4-1]; chain 3 also the remainder (:806-823), processed back to front */
    uint8_t last[4];
    for (int k = 0; k < 3; k++) last[k] = in[(k + 1) * q4 - 1];
    last[3] = in[n - 1];
    for (int64_t i = (int64_t)n - 2; i > (int64_t)4 * q4 - 2; i--) {             /* tail, chain 3 only */
        enc_put (&R[3], &w, &sy[in[i]][last[3]]);
        last[3] = in[i];
    }
    for (int64_t i = (int64_t)q4 - 2; i >= 0; i--)
        for (int k = 3; k >= 0; k--) {
            uint8_t c = in[k * q4 + i];
            enc_put (&R[k], &w, &sy[c][last[k]]);
            last[k] = c;
        }
    for (int k = 3; k >= 0; k--) enc_put (&R[k], &w, &sy[0][last[k]]);           /* :843-846 first symbol in context 0 */
    for (int k = 3; k >= 0; k--) enc_flush (R[k], &w);

    memcpy (out + tab, w, (size_t)(end - w));
    uint32_t len = tab + (uint32_t)(end - w);
    free (scratch); free (F); free (sy);
    return len;
}

/* ------------------------------------------------------------------ rANS decoders */
static inline uint32_t rd32 (const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

static int rans_o0_decode (const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t n)   /* :498-613 */
{
    if (in_len < 16) return -1;
    const uint8_t *p = in, *end = in + in_len;
    uint32_t F[256] = {0}, tot = 0;
    int k = get_alphabet (p, end - 8, F);
    if (!k) return -1;
    p += k;
    for (int j = 0; j < 256; j++) if (F[j]) { p += vget (p, end - 8, &F[j]); tot += F[j]; }    /* decode_freq :268-286 */
    shift_freqs (F, tot, 4096);

    static __thread uint8_t  ssym[4096];
    static __thread uint16_t sfreq[4096], sbase[4096];
    uint32_t x = 0;
    for (int j = 0; j < 256; j++) {
        if (!F[j]) continue;
        if (F[j] > 4096 - x) return -1;
        for (uint32_t y = 0; y < F[j]; y++) { ssym[x + y] = (uint8_t)j; sfreq[x + y] = (uint16_t)F[j]; sbase[x + y] = (uint16_t)y; }
        x += F[j];
    }
    if (x != 4096) return -1;
    if (p + 16 > end) return -1;

    uint32_t R[4];
    for (int s = 0; s < 4; s++) { R[s] = rd32 (p); p += 4; if (R[s] < RANS_L) return -1; }
    for (uint32_t i = 0; i < n; i++) {
        uint32_t *r = &R[i & 3], m = *r & 4095;
        *r = sfreq[m] * (*r >> 12) + sbase[m];
        out[i] = ssym[m];
        if (*r < RANS_L && p + 1 < end) { *r = (*r << 16) | p[0] | (p[1] << 8); p += 2; }
    }
    return 0;
}

static int rans_o1_decode (const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t n)   /* :883-1143 */
{
    if (in_len < 16) return -1;
    const uint8_t *p = in, *end = in + in_len, *tab_end = NULL, *fend = end;
    uint8_t *ctab = NULL;
    uint32_t shift = *p >> 4;
    int rc = -1;
    uint8_t  *sfb = NULL;                                  /* [256][1<<shift] slot -> symbol */
    uint16_t (*fb)[256][2] = NULL;                         /* [ctx][sym] {freq, base} */

    if (*p++ & 1) {                                        /* O0-compressed table (:957-968) */
        uint32_t usz, csz;
        p += vget (p, end, &usz);
        p += vget (p, end, &csz);
        if (csz > end - p - 16) return -1;
        tab_end = p + csz;
        ctab = malloc (usz ? usz : 1);
        if (rans_o0_decode (p, csz, ctab, usz)) goto done;
        p = ctab; fend = ctab + usz;
    }
    if (shift != 10 && shift != 12) goto done;
    sfb = malloc ((size_t)256 << shift);
    fb  = calloc (256, sizeof *fb);

    uint32_t F0[256] = {0};
    int k = get_alphabet (p, fend, F0);
    if (!k) goto done;
    p += k;
    if (p >= fend) goto done;

    for (int i = 0; i < 256; i++) {
        if (!F0[i]) continue;
        uint32_t F[256] = {0}, T = 0;
        k = get_freqs_o1 (p, fend, F0, F, &T);
        if (!k) goto done;
        p += k;
        if (!T) continue;
        shift_freqs (F, T, 1u << shift);
        uint32_t x = 0;
        for (int j = 0; j < 256; j++) {
            if (!F[j]) continue;
            if (F[j] > (1u << shift) - x) goto done;
            memset (sfb + ((size_t)i << shift) + x, j, F[j]);
            fb[i][j][0] = (uint16_t)F[j]; fb[i][j][1] = (uint16_t)x;
            x += F[j];
        }
        if (x != (1u << shift)) goto done;
    }
    if (tab_end) p = tab_end;
    if (p + 16 > end) goto done;

    uint32_t R[4];
    for (int s = 0; s < 4; s++) { R[s] = rd32 (p); p += 4; if (R[s] < RANS_L) goto done; }
    uint32_t q4 = n >> 2, mask = (1u << shift) - 1;
    uint8_t ctx[4] = {0,0,0,0};
    for (uint32_t i = 0; i < q4; i++)
        for (int s = 0; s < 4; s++) {
            /* note the reference decodes all four symbols of a step before renormalising any (:1046-1072);
               renormalisation order is still 0,1,2,3 so the word stream is consumed identically */
            uint32_t m = R[s] & mask;
            uint8_t c = sfb[((size_t)ctx[s] << shift) + m];
            R[s] = fb[ctx[s]][c][0] * (R[s] >> shift) + m - fb[ctx[s]][c][1];
            out[s * q4 + i] = ctx[s] = c;
            if (s == 3)
                for (int r = 0; r < 4; r++)
                    if (R[r] < RANS_L && p + 1 < end) { R[r] = (R[r] << 16) | p[0] | (p[1] << 8); p += 2; }
        }
    for (uint32_t i = 4 * q4; i < n; i++) {                /* remainder on chain 3 (:1076-1083) */
        uint32_t m = R[3] & mask;
        uint8_t c = sfb[((size_t)ctx[3] << shift) + m];
        R[3] = fb[ctx[3]][c][0] * (R[3] >> shift) + m - fb[ctx[3]][c][1];
        out[i] = ctx[3] = c;
        if (R[3] < RANS_L && p + 1 < end) { R[3] = (R[3] << 16) | p[0] | (p[1] << 8); p += 2; }
    }
    rc = 0;
done:
    free (sfb); free (fb); free (ctab);
    return rc;
}

/* ------------------------------------------------------------------ PACK (pack.c:58-154, 168-201, 214-351) */
/* returns packed length; *meta_len bytes of meta written to meta[].  If >16 symbols: data copied, meta_len=1 */
static uint64_t pack_syms (const uint8_t *in, uint64_t n, uint8_t *meta, int *meta_len, uint8_t *out)
{
    int code[256] = {0}, ns = 0;
    for (uint64_t i = 0; i < n; i++) code[in[i]] = 1;
    for (int i = 0; i < 256; i++) if (code[i]) { code[i] = ns++; meta[ns] = (uint8_t)i; }
    meta[0] = (uint8_t)ns;                                 /* 256 wraps to 0 */
    if (ns > 16) { *meta_len = 1; memcpy (out, in, n); return n; }
    *meta_len = ns + 1;
    int bits = ns > 4 ? 4 : ns > 2 ? 2 : ns > 1 ? 1 : 0;
    if (!bits) return 0;
    int per = 8 / bits;
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; i += per) {
        uint8_t b = 0;
        for (int k = 0; k < per && i + k < n; k++) b |= (uint8_t)(code[in[i + k]] << (k * bits));
        out[o++] = b;
    }
    return o;
}

static int unpack_meta (const uint8_t *in, uint32_t in_len, uint8_t *map, int *per_byte)
{
    if (!in_len) return 0;
    unsigned ns = in[0] ? in[0] : 256;
    if      (ns <= 1)  *per_byte = 0;
    else if (ns <= 2)  *per_byte = 8;
    else if (ns <= 4)  *per_byte = 4;
    else if (ns <= 16) *per_byte = 2;
    else { *per_byte = 1; return 1; }
    if (in_len <= 1) return 0;
    unsigned c = 0, j = 1;
    do { map[c++] = in[j++]; } while (c < ns && j < in_len);
    return c < ns ? 0 : (int)j;
}

static int unpack_syms (const uint8_t *in, int64_t in_len, uint8_t *out, uint64_t n, int per, const uint8_t *map)
{
    if (per == 1) { memcpy (out, in, (size_t)in_len); return 0; }
    if (per == 0) { memset (out, map[0], n); return 0; }
    int bits = 8 / per;
    if ((int64_t)((n + per - 1) / per) > in_len) return -1;
    for (uint64_t i = 0; i < n; i++)
        out[i] = map[(in[i / per] >> ((i % per) * bits)) & ((1 << bits) - 1)];
    return 0;
}

/* ------------------------------------------------------------------ rANS container */
int orc_rans_compress (const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_len, int order)   /* :1151-1356 */
{
    if (*out_len < orc_rans_bound (n, order)) return -1;
    if (n <= 20) order &= ~F_STRIPE;

    if (order & F_STRIPE) {                                                     /* :1165-1227 */
        const int N = 4;
        uint8_t *planes = malloc (n);
        uint32_t plen[4], pidx[4];
        for (int i = 0; i < N; i++) { plen[i] = n / N + ((n % N) > (uint32_t)i); pidx[i] = i ? pidx[i-1] + plen[i-1] : 0; }
        for (uint32_t i = 0; i < n; i++) planes[pidx[i % N] + i / N] = in[i];

        uint32_t hdr = 1;
        out[0] = (uint8_t)(order & ~F_NOSZ);
        hdr += vput (out + hdr, n);
        out[hdr++] = N;
        uint8_t *body0 = out + 2 + 5 * (N + 1), *body = body0;
        static const int cand[4] = { 1, 64, 128, 0 };
        for (int i = 0; i < N; i++) {
            int best = 0, last = -1; uint32_t best_sz = n + 10, sz = 0;
            for (int j = 0; j < 4; j++) {
                if ((order & cand[j]) != cand[j]) continue;
                sz = *out_len - (uint32_t)(body - out);
                orc_rans_compress (planes + pidx[i], plen[i], body, &sz, cand[j] | F_NOSZ);
                if (best_sz > sz) { best_sz = sz; best = j; }
                last = j;
            }
            if (best != 3) {                                                    /* reference: best_j != j-1 with j==4 */
                sz = *out_len - (uint32_t)(body - out);
                orc_rans_compress (planes + pidx[i], plen[i], body, &sz, cand[best] | F_NOSZ);
            }
            (void)last;
            body += sz;
            hdr += vput (out + hdr, sz);
        }
        memmove (out + hdr, body0, (size_t)(body - body0));
        *out_len = hdr + (uint32_t)(body - body0);
        free (planes);
        return 0;
    }

    if (order & F_CAT) {                                                        /* :1229-1236 */
        out[0] = F_CAT;
        uint32_t h = 1 + vput (out + 1, n);
        memcpy (out + h, in, n);
        *out_len = h + n;
        return 0;
    }

    int do_pack = order & F_PACK, no_size = order & F_NOSZ;
    uint32_t h = 1, cap = *out_len;
    out[0] = (uint8_t)order;
    if (!no_size) h += vput (out + 1, n);
    order &= 0xf;

    uint8_t *packed = NULL;
    if (do_pack && n) {                                                         /* :1255-1278 */
        int ml;
        packed = malloc ((size_t)n + 1);
        uint8_t meta[260];
        uint64_t plen = pack_syms (in, n, meta, &ml, packed);
        if (ml == 1 && meta[0] > 16) { out[0] &= ~F_PACK; free (packed); packed = NULL; }
        else {
            memcpy (out + h, meta, ml);
            in = packed; n = (uint32_t)plen; h += ml;
            int s = vput (out + h, n);
            h += s; cap -= s;
        }
    }
    else if (do_pack) out[0] &= ~F_PACK;

    /* RLE (:1280-1330) is never requested by genozip's order bytes (codec_htscodecs.c:17-20) nor by the STRIPE
       candidate filter for them; an RLE request is treated as "not worth it" is NOT valid, so refuse. */
    if (order & F_RLE) { free (packed); return -1; }

    cap -= h;
    if (order && n < 8) { out[0] &= ~1; order &= ~1; }                          /* :1333-1336 */

    uint32_t blen = (order == 1) ? rans_o1_block (in, n, out + h) : rans_o0_block (in, n, out + h);
    if (blen >= n) {                                                            /* :1343-1348 CAT fallback */
        out[0] &= ~3;
        out[0] |= F_CAT | no_size;
        memcpy (out + h, in, n);
        blen = n;
    }
    free (packed);
    *out_len = blen + h;
    return 0;
}

int orc_rans_uncompress (const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t *out_len)        /* :1358-1642 */
{
    const uint8_t *end = in + in_len;
    if (!in_len) return -1;

    if (*in & F_STRIPE) {                                                       /* :1366-1439 */
        uint32_t ulen, h = 1;
        h += vget (in + h, end, &ulen);
        if (h >= in_len) return -1;
        uint32_t N = in[h++];
        if (ulen != *out_len || N == 0) return -1;
        uint32_t clen[256], ul[256], idx[256]; uint64_t ctot = 0;
        for (uint32_t i = 0; i < N; i++) {
            ul[i] = ulen / N + ((ulen % N) > i);
            idx[i] = i ? idx[i-1] + ul[i-1] : 0;
            h += vget (in + h, end, &clen[i]);
            ctot += clen[i];
            if (h > in_len || clen[i] > in_len || clen[i] < 1) return -1;
        }
        if (h + ctot > in_len) return -1;
        in_len = h + (uint32_t)ctot;
        uint8_t *planes = malloc (ulen ? ulen : 1);
        for (uint32_t i = 0; i < N; i++) {
            uint32_t ol = ul[i];
            if (in_len < h || orc_rans_uncompress (in + h, in_len - h, planes + idx[i], &ol) || ol != ul[i]) { free (planes); return -1; }
            h += clen[i];
        }
        for (uint32_t j = 0; j < ulen; j++) out[j] = planes[idx[j % N] + j / N];  /* unstripe utils.h:41-73 */
        free (planes);
        return 0;
    }

    int order = *in++; in_len--;
    int do_pack = order & F_PACK, do_rle = order & F_RLE, do_cat = order & F_CAT, no_size = order & F_NOSZ;
    order &= 1;
    if (do_rle) return -1;                                                      /* never produced on this path */

    uint32_t osz;
    if (!no_size) { int s = vget (in, end, &osz); in += s; in_len -= s; }
    else osz = *out_len;
    if (*out_len < osz) return -1;
    *out_len = osz;

    uint8_t map[16] = {0}; int per = 0;
    uint32_t body_ulen = osz;
    uint8_t *tmp = NULL, *dst = out;
    if (do_pack) {
        int ms = unpack_meta (in, in_len, map, &per);
        if (!ms) return -1;
        in += ms; in_len -= ms;
        uint32_t psz; int s = vget (in, end, &psz); in += s; in_len -= s;
        if (psz > osz) return -1;
        body_ulen = psz;
        tmp = malloc (osz ? osz : 1);
        dst = tmp;
    }

    int rc = 0;
    if (in_len) {
        if (do_cat) { if (body_ulen > in_len) rc = -1; else memcpy (dst, in, body_ulen); }
        else rc = order ? rans_o1_decode (in, in_len, dst, body_ulen) : rans_o0_decode (in, in_len, dst, body_ulen);
    }
    else body_ulen = 0;

    if (!rc && do_pack) {
        uint64_t un = (per == 1) ? body_ulen : osz;
        rc = unpack_syms (tmp, body_ulen, out, un, per, map);
        *out_len = (uint32_t)un;
    }
    else if (!rc) *out_len = body_ulen;
    free (tmp);
    return rc;
}

/* ================================================================== adaptive arithmetic coder */

/* ---- range coder (c_range_coder.h:25-127) ---- */
#define RC_TOP   (1u << 24)
#define RC_THRES (255u * RC_TOP)
typedef struct { uint32_t low, code, range, ffnum, cache, carry; uint8_t *out; const uint8_t *in, *in_end; } RC;

static void rc_shift_low (RC *rc)                                               /* :70-88 */
{
    if (rc->low < RC_THRES || rc->carry) {
        *rc->out++ = (uint8_t)(rc->cache + rc->carry);
        while (rc->ffnum) { *rc->out++ = (uint8_t)(rc->carry - 1); rc->ffnum--; }
        rc->cache = rc->low >> 24;
        rc->carry = 0;
    }
    else rc->ffnum++;
    rc->low <<= 8;
}

static void rc_encode (RC *rc, uint32_t cum, uint32_t freq, uint32_t tot)       /* :97-109 */
{
    uint32_t before = rc->low;
    rc->range /= tot;
    rc->low   += cum * rc->range;
    rc->range *= freq;
    rc->carry += rc->low < before;
    while (rc->range < RC_TOP) { rc->range <<= 8; rc_shift_low (rc); }
}

static void rc_start_decode (RC *rc, const uint8_t *in, const uint8_t *in_end)  /* :57-68 */
{
    memset (rc, 0, sizeof *rc);
    rc->range = 0xFFFFFFFFu; rc->in = in; rc->in_end = in_end;
    if (rc->in + 5 > rc->in_end) { rc->in = rc->in_end; return; }
    for (int i = 0; i < 5; i++) rc->code = (rc->code << 8) | *rc->in++;
}

static inline uint32_t rc_get_freq (RC *rc, uint32_t tot)                       /* :111-114 */
{
    return (tot && rc->range >= tot) ? rc->code / (rc->range /= tot) : 0;
}

static void rc_decode (RC *rc, uint32_t cum, uint32_t freq)                     /* :116-126 */
{
    rc->code  -= cum * rc->range;
    rc->range *= freq;
    while (rc->range < RC_TOP) {
        if (rc->in >= rc->in_end) return;
        rc->code = (rc->code << 8) + *rc->in++;
        rc->range <<= 8;
    }
}

/* ---- adaptive model (c_simple_model.h:77-179), up to 258 symbols ---- */
#define M_MAXF  ((1 << 16) - 17)
#define M_STEP  16
typedef struct { uint32_t tot; uint16_t f[260]; uint16_t s[260]; int nsym; } Model;   /* f[0]/s[0] is the sentinel; entries start at 1 */

static void model_init (Model *m, int nsym, int max_sym)                        /* :85-103 */
{
    m->nsym = nsym;
    m->f[0] = M_MAXF; m->s[0] = 0;
    for (int i = 0; i < nsym; i++) { m->s[i + 1] = (uint16_t)i; m->f[i + 1] = i < max_sym ? 1 : 0; }
    m->f[nsym + 1] = 0;                                                          /* F[NSYM].Freq = 0 terminates normalise */
    m->tot = (uint32_t)max_sym;
}

static void model_halve (Model *m)                                              /* :106-116 */
{
    m->tot = 0;
    for (int i = 1; m->f[i]; i++) { m->f[i] -= m->f[i] >> 1; m->tot += m->f[i]; }
}

static inline int model_bump (Model *m, int i)                                  /* shared tail of :131-145 / :160-178; returns symbol */
{
    m->f[i] += M_STEP; m->tot += M_STEP;
    if (m->tot > M_MAXF) model_halve (m);
    if (m->f[i] > m->f[i - 1]) {
        uint16_t tf = m->f[i], ts = m->s[i];
        m->f[i] = m->f[i - 1]; m->s[i] = m->s[i - 1];
        m->f[i - 1] = tf; m->s[i - 1] = ts;
        return ts;
    }
    return m->s[i];
}

static void model_encode (Model *m, RC *rc, uint16_t sym)                       /* :123-146 */
{
    int i = 1; uint32_t acc = 0;
    while (m->s[i] != sym) acc += m->f[i++];
    rc_encode (rc, acc, m->f[i], m->tot);
    model_bump (m, i);
}

static uint16_t model_decode (Model *m, RC *rc)                                 /* :148-179 */
{
    uint32_t freq = rc_get_freq (rc, m->tot);
    if (freq > M_MAXF) return 0;
    int i = 1; uint32_t acc = 0;
    while ((acc += m->f[i]) <= freq) { i++; if (i > m->nsym + 1) return 0; }
    acc -= m->f[i];
    rc_decode (rc, acc, m->f[i]);
    return (uint16_t)model_bump (m, i);
}

/* ---- leaf coders (arith_dynamic.c:92-226 plain, :387-608 RLE) ---- */
#define MAX_RUN 4
static uint32_t arith_block (const uint8_t *in, uint32_t n, uint8_t *out, int order, int rle)
{
    unsigned maxs = 0;
    for (uint32_t i = 0; i < n; i++) if (maxs < in[i]) maxs = in[i];
    maxs++;
    out[0] = (uint8_t)maxs;

    int nctx = order ? 256 : 1;
    Model *lit = malloc (nctx * sizeof *lit), *run = NULL;
    for (int i = 0; i < nctx; i++) model_init (&lit[i], 256, (int)maxs);
    if (rle) { run = malloc (258 * sizeof *run); for (int i = 0; i < 258; i++) model_init (&run[i], 258, MAX_RUN); }

    RC rc; memset (&rc, 0, sizeof rc); rc.range = 0xFFFFFFFFu; rc.out = out + 1;
    uint8_t last = 0;
    for (uint32_t i = 0; i < n; ) {
        model_encode (&lit[order ? last : 0], &rc, in[i]);
        last = in[i++];
        if (!rle) continue;
        int r = 0;
        while (i < n && in[i] == last) { r++; i++; }
        int rctx = last;
        do {                                                                     /* :416-438 */
            int c = r < MAX_RUN ? r : MAX_RUN - 1;
            model_encode (&run[rctx], &rc, (uint16_t)c);
            r -= c;
            if (rctx == last) rctx = 256; else rctx += (rctx < 257);
            if (c == MAX_RUN - 1 && r == 0) model_encode (&run[rctx], &rc, 0);
        } while (r);
    }
    for (int i = 0; i < 5; i++) rc_shift_low (&rc);                              /* RC_FinishEncode */
    free (lit); free (run);
    return (uint32_t)(rc.out - (out + 1)) + 1;
}

static int arith_block_decode (const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t n, int order, int rle)
{
    unsigned maxs = in[0] ? in[0] : 256;
    int nctx = order ? 256 : 1;
    Model *lit = malloc (nctx * sizeof *lit), *run = NULL;
    for (int i = 0; i < nctx; i++) model_init (&lit[i], 256, (int)maxs);
    if (rle) { run = malloc (258 * sizeof *run); for (int i = 0; i < 258; i++) model_init (&run[i], 258, MAX_RUN); }

    RC rc; rc_start_decode (&rc, in + 1, in + in_len);
    uint8_t last = 0;
    for (uint32_t i = 0; i < n; i++) {
        out[i] = (uint8_t)model_decode (&lit[order ? last : 0], &rc);
        last = out[i];
        if (!rle) continue;
        uint32_t r = 0, part; int rctx = last;
        do {                                                                     /* :473-482 / :591-599 */
            part = model_decode (&run[rctx], &rc);
            if (rctx == last) rctx = 256; else rctx += (rctx < 257);
            r += part;
        } while (part == MAX_RUN - 1 && r < n);
        while (r-- && i + 1 < n) out[++i] = last;
    }
    free (lit); free (run);
    return 0;
}

int orc_arith_compress (const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_len, int order)   /* arith_dynamic.c:615-858 */
{
    if (*out_len < orc_arith_bound (n, order)) return -1;
    if (n <= 20) order &= ~F_STRIPE;
    /* the X_CAT early branch (:629-634) is unreachable from genozip's order bytes */

    if (order & F_STRIPE) {                                                     /* :636-768 */
        const int N = 4;
        uint8_t *planes = malloc (n);
        uint32_t plen[4], pidx[4];
        for (int i = 0; i < N; i++) { plen[i] = n / N + ((n % N) > (uint32_t)i); pidx[i] = i ? pidx[i-1] + plen[i-1] : 0; }
        for (uint32_t i = 0; i < n; i++) planes[pidx[i % N] + i / N] = in[i];

        uint32_t hdr = 1;
        out[0] = (uint8_t)(order & ~F_NOSZ);
        hdr += vput (out + hdr, n);
        out[hdr++] = N;
        uint8_t *body0 = out + 2 + 5 * (N + 1), *body = body0;
        static const int cand[4][4] = { {3, 1, 64, 0}, {2, 1, 0, 0}, {2, 1, 128, 0}, {2, 1, 128, 0} };   /* :684-687 */
        for (int i = 0; i < N; i++) {
            const int *m = cand[i < 3 ? i : 3];
            int best = 0, j; uint32_t best_sz = INT_MAX, sz = 0;
            for (j = 1; j <= m[0]; j++) {
                sz = *out_len - (uint32_t)(body - out);
                if ((order & 3) == 0 && (m[j] & 1)) continue;
                orc_arith_compress (planes + pidx[i], plen[i], body, &sz, m[j] | F_NOSZ);
                if (best_sz > sz) { best_sz = sz; best = j; }
            }
            if (best != j - 1) {
                sz = *out_len - (uint32_t)(body - out);
                orc_arith_compress (planes + pidx[i], plen[i], body, &sz, m[best] | F_NOSZ);
            }
            body += sz;
            hdr += vput (out + hdr, sz);
        }
        memmove (out + hdr, body0, (size_t)(body - body0));
        *out_len = hdr + (uint32_t)(body - body0);
        free (planes);
        return 0;
    }

    int do_pack = order & F_PACK, do_rle = order & F_RLE, no_size = order & F_NOSZ;
    uint32_t h = 1;
    out[0] = (uint8_t)order;
    if (!no_size) h += vput (out + 1, n);
    order &= 3;

    uint8_t *packed = NULL;
    if (do_pack && n) {                                                         /* :789-811 */
        int ml; uint8_t meta[260];
        packed = malloc ((size_t)n + 1);
        uint64_t plen = pack_syms (in, n, meta, &ml, packed);
        if (ml == 1 && meta[0] > 16) { out[0] &= ~F_PACK; free (packed); packed = NULL; }
        else {
            memcpy (out + h, meta, ml);
            in = packed; n = (uint32_t)plen; h += ml;
            h += vput (out + h, n);
        }
    }
    else if (do_pack) out[0] &= ~F_PACK;

    if (do_rle && !n) out[0] &= ~F_RLE;                                          /* :813-815 */
    if (order && n < 8) { out[0] &= ~3; order &= ~3; }                           /* :818-821 */

    uint32_t blen = arith_block (in, n, out + h, order == 1, do_rle);
    if (blen >= n) {                                                            /* :847-852 */
        out[0] &= ~(3 | 4);
        out[0] |= F_CAT | no_size;
        memcpy (out + h, in, n);
        blen = n;
    }
    free (packed);
    *out_len = blen + h;
    return 0;
}

int orc_arith_uncompress (const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t *out_len)       /* :860-1104 */
{
    const uint8_t *end = in + in_len;
    if (!in_len) return -1;

    if (*in & F_STRIPE) {
        uint32_t ulen, h = 1;
        h += vget (in + h, end, &ulen);
        if (h >= in_len) return -1;
        uint32_t N = in[h++];
        if (ulen != *out_len || N == 0) return -1;
        uint32_t clen[256], ul[256], idx[256]; uint64_t ctot = 0;
        for (uint32_t i = 0; i < N; i++) {
            ul[i] = ulen / N + ((ulen % N) > i);
            idx[i] = i ? idx[i-1] + ul[i-1] : 0;
            h += vget (in + h, end, &clen[i]);
            ctot += clen[i];
            if (h > in_len || clen[i] > in_len || clen[i] < 1) return -1;
        }
        if (h + ctot > in_len) return -1;
        in_len = h + (uint32_t)ctot;
        uint8_t *planes = malloc (ulen ? ulen : 1);
        for (uint32_t i = 0; i < N; i++) {
            uint32_t ol = ul[i];
            if (in_len < h || orc_arith_uncompress (in + h, in_len - h, planes + idx[i], &ol) || ol != ul[i]) { free (planes); return -1; }
            h += clen[i];
        }
        for (uint32_t j = 0; j < ulen; j++) out[j] = planes[idx[j % N] + j / N];
        free (planes);
        return 0;
    }

    int order = *in++; in_len--;
    int do_pack = order & F_PACK, do_rle = order & F_RLE, do_cat = order & F_CAT, no_size = order & F_NOSZ;
    if (order & 4) return -1;                                                   /* X_EXT (bzip2) never used by genozip */
    order &= 3;

    uint32_t osz;
    if (!no_size) { int s = vget (in, end, &osz); in += s; in_len -= s; }
    else osz = *out_len;
    if (*out_len < osz) return -1;
    *out_len = osz;

    uint8_t map[16] = {0}; int per = 0;
    uint32_t body_ulen = osz;
    uint8_t *tmp = NULL, *dst = out;
    if (do_pack) {
        int ms = unpack_meta (in, in_len, map, &per);
        if (!ms) return -1;
        in += ms; in_len -= ms;
        uint32_t psz; int s = vget (in, end, &psz); in += s; in_len -= s;
        if (psz > osz) return -1;
        body_ulen = psz;
        tmp = malloc (osz ? osz : 1);
        dst = tmp;
    }

    int rc = 0;
    if (in_len) {
        if (do_cat) { if (body_ulen > in_len) rc = -1; else memcpy (dst, in, body_ulen); }
        else rc = arith_block_decode (in, in_len, dst, body_ulen, order == 1, do_rle);
    }
    else body_ulen = 0;

    if (!rc && do_pack) {
        uint64_t un = (per == 1) ? body_ulen : osz;
        rc = unpack_syms (tmp, body_ulen, out, un, per, map);
        *out_len = (uint32_t)un;
    }
    else if (!rc) *out_len = body_ulen;
    free (tmp);
    return rc;
}
