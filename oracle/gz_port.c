/* oracle/gz_port.c — CPU restatement of genozip's own codecs on the hot path: ACGT/XCGT, DOMQ, PBWT, LONGR.
 * TEST INFRASTRUCTURE ONLY — see oracle/oracle.h.
 *
 * Parity status: the ENCODERS are pinned byte-for-byte against the reference's own compiled translation units
 * (src/codec_{acgt,domq,pbwt,longr}.c, unmodified, hosted by oracle/ref_gz_shim.c -> oracle/_ref/libgz_ref.so;
 * tests/test_oracle_gz_ref.py).  The DECODERS restate the reference decoder from its own source (no code shared with
 * the encoders); they are pinned against the reference's own uncompress / reconstruct functions from the same objects
 * (same test file) and by inverting the pinned encoders (tests/test_oracle_gz.py).  Each function follows
 * its reference function line by line (citations below, relative to /root/reference/src).  Interfaces are flat
 * (buffers + line tables) — what the reference reads through VBlock/Context is passed explicitly.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* ================================================================== ACGT / XCGT (codec_acgt.c) */

/* _acgt_encode (reference.c:45-58): A,C,G,T (either case) -> 0..3; IUPAC codes -> their lowest base; the rest 0 */
static uint8_t acgt_code (uint8_t c)
{
    switch (c) {
        case 'C': case 'c': case 'Y': case 'y': case 'S': case 's': case 'B': case 'b': return 1;
        case 'G': case 'g': case 'K': case 'k': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 0;
    }
}

/* _acgt_encode_comp (reference.c:63-75) */
static uint8_t acgt_code_comp (uint8_t c)
{
    switch (c) {
        case 'A': case 'a': return 3;
        case 'C': case 'c': case 'M': case 'm': return 2;
        case 'G': case 'g': case 'R': case 'r': case 'S': case 's': case 'V': case 'v': return 1;
        default: return 0;
    }
}

uint64_t orc_acgt_packed_len (uint64_t n) { return ((2 * n + 63) / 64) * 8; }      /* roundup_bits2bytes64, bits.h:70-71 */

int orc_acgt_pack (const uint8_t *seq, uint64_t n, uint8_t *packed, uint8_t *x)      /* codec_acgt.c:45-55, 64-140 */
{
    memset (packed, 0, orc_acgt_packed_len (n));
    int all_zero = 1;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t c = seq[i];
        packed[i / 4] |= (uint8_t)(acgt_code (c) << (2 * (i % 4)));                  /* LE 64-bit words == LE byte order */
        uint8_t e = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? 0               /* :67-70: XOR with self */
                  : (c == 'a' || c == 'c' || c == 'g' || c == 't') ? 1               /*         XOR with self^1 */
                  : c;                                                               /*         unchanged */
        if (x) x[i] = e;
        if (e) all_zero = 0;
    }
    return all_zero;
}

void orc_acgt_unpack (const uint8_t *packed, const uint8_t *x, uint64_t n, uint8_t *seq)   /* :185-248 */
{
    static const char dec[4] = { 'A', 'C', 'G', 'T' };
    for (uint64_t i = 0; i < n; i++) {
        char b = dec[(packed[i / 4] >> (2 * (i % 4))) & 3];
        if (!x || x[i] == 0) seq[i] = (uint8_t)b;
        else if (x[i] == 1)  seq[i] = (uint8_t)(b + 32);
        else                 seq[i] = x[i];
    }
}

/* ================================================================== DOMQ (codec_domq.c) */
#define FIRST_Q 32
#define NUM_Q   95

typedef struct { uint8_t q; uint32_t count; } QMap;
static int qmap_desc (const void *a, const void *b)                                  /* DESCENDING_SORTER, sorter.h:16-33 */
{
    uint32_t ca = ((const QMap *)a)->count, cb = ((const QMap *)b)->count;
    return -((ca > cb) ? 1 : (ca < cb) ? -1 : 0);
}

void orc_domq_prepare (const uint8_t *txt, const uint64_t *line_off, const uint32_t *line_len, uint32_t n_lines,
                       uint8_t *line_dom, uint8_t *line_diverse, OrcDomqTables *t)
{
    static uint32_t hist[NUM_Q][NUM_Q];
    uint32_t lines_with_dom[NUM_Q] = {0};
    memset (hist, 0, sizeof hist);
    memset (t, 0, sizeof *t);
    t->n_lines = n_lines;

    for (uint32_t li = 0; li < n_lines; li++) {                                       /* codec_domq_calc_histogram :139-178 */
        line_dom[li] = 0; line_diverse[li] = 0;
        uint32_t len = line_len[li];
        if (!len) continue;
        uint32_t lh[NUM_Q] = {0};
        const uint8_t *q = txt + line_off[li];
        for (uint32_t i = 0; i < len; i++) lh[q[i] - FIRST_Q]++;
        uint32_t best = 0; int dom = 0;
        for (int k = 0; k < NUM_Q; k++) if (lh[k] >= best) { best = lh[k]; dom = k; }   /* ties -> higher ASCII */
        if (100 * lh[dom] / len < 85) { line_diverse[li] = 1; t->has_diverse = 1; }
        line_dom[li] = (uint8_t)dom;
        lines_with_dom[dom]++;
        for (int k = 0; k < NUM_Q; k++) hist[dom][k] += lh[k];
    }

    uint8_t dom_to_cdom[NUM_Q] = {0};                                                 /* compact_histogram :180-197 */
    int num_doms = 0;
    for (int k = 0; k < NUM_Q; k++)
        if (lines_with_dom[k]) {
            dom_to_cdom[k] = (uint8_t)num_doms;
            if (num_doms != k) memcpy (hist[num_doms], hist[k], sizeof hist[0]);
            num_doms++;
        }

    uint8_t denormalize[NUM_Q][NUM_Q];                                                /* calc_norm_table :199-247 */
    memset (denormalize, 0, sizeof denormalize);
    int num_norm_qs = 0;
    for (int cd = 0; cd < num_doms; cd++) {
        QMap m[NUM_Q];
        for (int k = 0; k < NUM_Q; k++) { m[k].q = (uint8_t)k; m[k].count = hist[cd][k]; }
        qsort (m, NUM_Q, sizeof (QMap), qmap_desc);                                   /* the libc's tie order is part of the result */
        int r = 0;
        for (; r < NUM_Q && m[r].count; r++) {
            t->normalize[cd * NUM_Q + m[r].q] = (uint8_t)r;
            denormalize[cd][r] = (uint8_t)(m[r].q + FIRST_Q);
        }
        if (r > num_norm_qs) num_norm_qs = r;
    }
    for (int cd = 0; cd < num_doms; cd++)
        for (int r = 0; r < num_norm_qs; r++) t->denorm[cd * num_norm_qs + r] = denormalize[cd][r];
    t->num_norm_qs = (uint8_t)num_norm_qs;
    t->num_doms = (uint8_t)num_doms;
    for (uint32_t li = 0; li < n_lines; li++) if (line_len[li]) line_dom[li] = dom_to_cdom[line_dom[li]];   /* :283-285 */
}

static void add_runs (uint8_t *runs, uint32_t *rl, uint32_t runlen)                   /* codec_domq_add_runs :368-377 */
{
    while (runlen) {
        uint32_t sub = runlen < 254 ? runlen : 254;
        runs[(*rl)++] = (uint8_t)(runlen <= 254 ? sub : 255);
        runlen -= sub;
    }
}

void orc_domq_split (const uint8_t *txt, const uint64_t *line_off, const uint32_t *line_len, uint32_t n_lines,
                     const uint8_t *line_dom, const uint8_t *line_diverse, const OrcDomqTables *t,
                     uint8_t *qual, uint32_t *qual_len, uint8_t *runs, uint32_t *runs_len,
                     uint8_t *mplx, uint32_t *mplx_len, uint8_t *divr, uint32_t *divr_len)
{
    const uint8_t no_doms = t->num_norm_qs;
    uint32_t ql = 0, rl = 0, ml = 0, dl = 0, runlen = 0;
    for (uint32_t li = 0; li < n_lines; li++) {                                       /* :421-468 */
        uint32_t len = line_len[li];
        if (!len) continue;
        const uint8_t *q = txt + line_off[li];
        const uint8_t *norm = t->normalize + line_dom[li] * NUM_Q;                    /* normalise :347-366 */
        if (line_diverse[li]) {
            for (uint32_t i = 0; i < len; i++) divr[dl++] = norm[q[i] - FIRST_Q];
            mplx[ml++] = line_dom[li] | 0x80;
            continue;
        }
        mplx[ml++] = line_dom[li];
        for (uint32_t i = 0; i < len; i++) {
            uint8_t v = norm[q[i] - FIRST_Q];
            if (v == 0) { runlen++; continue; }
            if (runlen) { add_runs (runs, &rl, runlen); runlen = 0; }
            else qual[ql++] = no_doms;
            qual[ql++] = v;
        }
    }
    uint32_t last_len = n_lines ? line_len[n_lines - 1] : 0;
    if (runlen && (rl || runlen < last_len)) {                                        /* :473-482 */
        add_runs (runs, &rl, runlen);
        qual[ql++] = no_doms;
    }
    if (!ql) qual[ql++] = 'X';                                                        /* :497-500 */
    *qual_len = ql; *runs_len = rl; *mplx_len = ml; *divr_len = dl;
}

/* ---- decoder, restated from the reference PIZ code (no code shared with the encoder above) ---- */
typedef struct { const uint8_t *qual; uint32_t qual_len, qnext; uint8_t *runs; uint32_t runs_len, rnext; } DqState;

static uint32_t dq_dom_run (DqState *s, uint8_t dom, uint32_t max_len, uint8_t *out, int *err)   /* :551-588 */
{
    if (s->rnext >= s->runs_len) { *err = 1; return 0; }
    uint8_t *start = s->runs + s->rnext, *r = start;
    while (*r++ == 255) if (r > s->runs + s->runs_len) { *err = 1; return 0; }
    uint32_t nb = (uint32_t)(r - start);
    uint32_t runlen = (nb - 1) * 254 + r[-1];
    if (runlen >= max_len) {                                                          /* codec_domq_shorten_run :529-548 */
        uint32_t next = runlen - max_len;
        uint32_t new_nb = (next + 253) / 254; if (!new_nb) new_nb = 1;
        if (next) { uint8_t m = (uint8_t)(next % 254); start[nb - 1] = m ? m : 254; }
        else start[nb - 1] = 0;
        s->rnext += nb - new_nb;
        runlen = max_len;
    }
    else s->rnext += nb;
    memset (out, dom, runlen);
    return runlen;
}

int orc_domq_reconstruct (const uint8_t *qual, uint32_t qual_len, uint8_t *runs, uint32_t runs_len,
                          const uint8_t *mplx, uint32_t mplx_len, const uint8_t *divr, uint32_t divr_len,
                          const uint8_t *denorm, uint8_t num_norm_qs,
                          const uint32_t *line_len, uint32_t n_lines, uint8_t *out)
{
    DqState s = { qual, qual_len, 0, runs, runs_len, 0 };
    uint32_t mnext = 0, dnext = 0;
    const uint8_t no_dom = num_norm_qs;
    int err = 0;
    for (uint32_t li = 0; li < n_lines; li++) {                                       /* codec_domq_reconstruct :774-809, once per line */
        uint32_t len = line_len[li];
        if (!len) continue;
        if (mplx_len != 1 && mnext >= mplx_len) return -1;
        uint8_t dom_i = (mplx_len == 1) ? mplx[0] : mplx[mnext++];
        const uint8_t *dn = denorm + (uint32_t)(dom_i & 0x7f) * num_norm_qs;
        if (dom_i >> 7) {                                                             /* reconstruct_divr :747-772 */
            if (dnext + len > divr_len) return -2;
            for (uint32_t i = 0; i < len; i++) out[i] = dn[divr[dnext + i]];
            dnext += len; out += len;
            continue;
        }
        uint8_t dom = dn[0];                                                          /* reconstruct_runs :664-745 */
        uint32_t done = 0;
        if (!runs_len) {
            while (done < len && s.qnext + 1 < s.qual_len) {
                if (qual[s.qnext] != no_dom) return -3;
                out[done++] = dn[qual[s.qnext + 1]];
                s.qnext += 2;
            }
            memset (out + done, dom, len - done);
            done = len;
        }
        else while (done < len) {
            if (s.qnext >= s.qual_len) return -4;
            uint8_t v = qual[s.qnext++];
            if (v != no_dom) {
                done += dq_dom_run (&s, dom, len - done, out + done, &err);
                if (err) return -5;
                if (done == len) { s.qnext--; break; }
            }
            else if (s.qual_len == s.qnext) {
                done += dq_dom_run (&s, dom, len - done, out + done, &err);
                if (err) return -6;
                s.qnext--;
                break;
            }
            else v = qual[s.qnext++];
            out[done++] = dn[v];
        }
        if (done != len) return -7;
        out += len;
    }
    return 0;
}

/* ================================================================== PBWT (codec_pbwt.c) */
#define N_SMALL_ALLELES 245                                                           /* vcf.h:750 */

typedef struct { uint8_t allele; uint32_t index; } Perm;

/* codec_pbwt_calculate_permutation :110-154 */
static void pbwt_permute (Perm **perm, Perm **temp, const uint8_t *line, uint32_t n, int first, int zip)
{
    if (first) for (uint32_t i = 0; i < n; i++) (*perm)[i].index = i;
    else {
        int has[256] = {0};
        for (uint32_t i = 0; i < n; i++) has[(*perm)[i].allele] = 1;
        uint8_t order[256]; int no = 0;
        for (uint8_t a = '0'; a != (uint8_t)('0' + N_SMALL_ALLELES); a++) if (has[a]) order[no++] = a;
        static const uint8_t pseudo[5] = { '.', '*', '%', '-', '&' };
        for (int i = 0; i < 5; i++) if (has[pseudo[i]]) order[no++] = pseudo[i];
        uint32_t t = 0;
        for (int k = 0; k < no; k++)
            for (uint32_t i = 0; i < n; i++) if ((*perm)[i].allele == order[k]) (*temp)[t++].index = (*perm)[i].index;
        Perm *sw = *perm; *perm = *temp; *temp = sw;
    }
    if (zip) for (uint32_t i = 0; i < n; i++) (*perm)[i].allele = line[(*perm)[i].index];
}

int orc_pbwt_encode (const uint8_t *ht, uint32_t n_lines, uint32_t w,
                     uint32_t *runs, uint32_t *n_runs, uint32_t *fgrc, uint32_t *n_fgrc)   /* :244-287 */
{
    Perm *perm = calloc (w ? w : 1, sizeof (Perm)), *temp = calloc (w ? w : 1, sizeof (Perm));
    uint32_t nr = 0, nf = 0;
    uint8_t run_allele = 0;
    for (uint32_t li = 0; li < n_lines; li++) {
        pbwt_permute (&perm, &temp, ht + (uint64_t)li * w, w, li == 0, 1);
        int back = li % 2;
        for (uint32_t i = 0; i < w; ) {                                               /* run_len_encode :213-238 */
            uint32_t rl = 0;
            for (; i < w && perm[back ? w - i - 1 : i].allele == run_allele; i++) rl++;
            if (rl) { runs[nr - 1] += rl; continue; }
            uint8_t done = run_allele;
            run_allele = perm[back ? w - i - 1 : i].allele;
            /* update_fgrc :181-210 */
            if (done == '0') {
                if (nf && run_allele == (fgrc[nf - 1] & 0xff)) {
                    uint32_t cnt = (fgrc[nf - 1] >> 8) + 1;
                    if (!(cnt & 0xffffff)) { free (perm); free (temp); return -1; }
                    fgrc[nf - 1] = (fgrc[nf - 1] & 0xff) | (cnt << 8);
                }
                else fgrc[nf++] = run_allele | (1u << 8);
            }
            else if (run_allele != '0') { fgrc[nf++] = run_allele | (1u << 8); runs[nr++] = 0; }
            runs[nr++] = 0;
        }
    }
    uint64_t len = (uint64_t)n_lines * w;                                             /* :274-276 */
    fgrc[nf++] = (uint32_t)(len & 0xffffffffu);
    fgrc[nf++] = (uint32_t)(len >> 32);
    *n_runs = nr; *n_fgrc = nf;
    free (perm); free (temp);
    return 0;
}

int orc_pbwt_decode (uint32_t *runs, uint32_t n_runs, uint32_t *fgrc, uint32_t n_fgrc,
                     uint32_t n_lines, uint8_t *ht, uint64_t *ht_len)                 /* :317-402 */
{
    if (n_fgrc < 2 || !n_lines) return -1;
    n_fgrc -= 2;
    uint64_t len = (uint64_t)fgrc[n_fgrc] | ((uint64_t)fgrc[n_fgrc + 1] << 32);
    uint32_t w = (uint32_t)(len / n_lines);
    *ht_len = len;
    Perm *perm = calloc (w ? w : 1, sizeof (Perm)), *temp = calloc (w ? w : 1, sizeof (Perm));
    uint32_t rnext = 0, fnext = 0;
    uint8_t run_allele = 0;
    uint32_t rows = w ? (uint32_t)(len / w) : 0;
    for (uint32_t li = 0; li < rows; li++) {                                          /* pbwt_decode_one_line :317-369 */
        if (!n_runs) { free (perm); free (temp); return -2; }
        if (!run_allele) run_allele = '0';
        uint8_t *line = ht + (uint64_t)li * w;
        pbwt_permute (&perm, &temp, line, w, li == 0, 0);
        int back = li % 2;
        for (uint32_t i = 0; i < w; ) {
            if (rnext >= n_runs) { free (perm); free (temp); return -3; }
            uint32_t rl = runs[rnext];
            uint32_t part = rl < w - i ? rl : w - i;
            for (uint32_t k = 0; k < part; k++, i++) {
                uint32_t o = back ? w - i - 1 : i;
                line[perm[o].index] = perm[o].allele = run_allele;
            }
            if (part < rl) runs[rnext] -= part;
            if (part == rl) {
                rnext++;
                if (run_allele == '0') {
                    if (fnext < n_fgrc) {
                        run_allele = (uint8_t)(fgrc[fnext] & 0xff);
                        uint32_t cnt = (fgrc[fnext] >> 8) - 1;
                        fgrc[fnext] = (fgrc[fnext] & 0xff) | (cnt << 8);
                        if (!cnt) fnext++;
                    }
                    else run_allele = 0;   /* no more foreground runs: only reachable at the very end */
                }
                else run_allele = '0';
            }
        }
    }
    free (perm); free (temp);
    return 0;
}

/* ================================================================== LONGR (codec_longr.c, codec_longr_alg.c) */
/* channel word layout (codec_longr_alg.c:65-95), LSB first: B:12 | difq:4 | qbin:5 | avg:5 | err_c:2 */
#define LR_NCTX(c)   ((c) & 0x1fffff)          /* B, difq, qbin : 21 bits */
#define LR_Q(c)      (((c) >> 12) & 0x1ff)     /* difq, qbin    :  9 bits */
#define LR_CHAN(c)   (((c) >> 12) & 0xffff)    /* difq,qbin,avg,err_c : 16 bits */
#define LR_DIVR(x)   (((x) + 8) >> 4)          /* DIV_ROUND(x, 4) (:46-47) */

typedef struct {
    uint16_t *avg_sums, *err_sums;             /* [1<<21] (:99-100) */
    uint32_t  err_total[1 << 9];               /* (:101) */
    uint32_t  chan;
    const uint8_t *v2b;
} LrState;

static void lr_init (LrState *s, const uint8_t *v2b)                                  /* codec_longr_alg_init :138-146 */
{
    s->avg_sums = malloc (sizeof (uint16_t) << 21);
    s->err_sums = calloc (1 << 21, sizeof (uint16_t));
    memset (s->err_total, 1 << 4, sizeof s->err_total);                               /* bytes of 0x10 => 0x10101010 per entry */
    for (uint32_t n = 0; n < (1u << 21); n++) s->avg_sums[n] = (uint16_t)(((n >> 16) & 0x1f) << 4);   /* qbin << AVG_SHIFT */
    s->chan = 0; s->v2b = v2b;
}

static void lr_update (LrState *s, uint8_t b, int32_t q1, int32_t q2)                 /* codec_longr_update_state :108-136 */
{
    uint32_t c = s->chan;
    uint32_t nc = LR_NCTX (c), qn = LR_Q (c);
    int32_t err = q1 - LR_DIVR ((int32_t)s->avg_sums[nc]);
    s->avg_sums[nc] = (uint16_t)(s->avg_sums[nc] + err);
    int32_t ae = err < 0 ? -err : err;
    s->err_sums[nc]  = (uint16_t)(s->err_sums[nc] + ae - LR_DIVR ((int32_t)s->err_sums[nc]));
    s->err_total[qn] = s->err_total[qn] + ae - (uint32_t)LR_DIVR (s->err_total[qn]);

    uint32_t B = ((c & 0xfff) << 2 | b) & 0xfff;
    int32_t d = q1 - q2;
    uint32_t il = d < 0 ? (((uint32_t)(-(int64_t)d)) << 1) - 1 : ((uint32_t)d) << 1;  /* INTERLACE, context.h:100 */
    uint32_t difq = il < 15 ? il : 15;
    uint32_t qbin = s->v2b[q1 & 0xff] & 0x1f;
    c = (c & ~0x1fffffu) | B | (difq << 12) | (qbin << 16);
    uint32_t nc2 = LR_NCTX (c), qn2 = LR_Q (c);
    uint32_t avg = s->v2b[LR_DIVR ((int32_t)s->avg_sums[nc2])] & 0x1f;
    uint32_t tot = s->err_total[qn2] >> 0;                                            /* TOTAL_ERR_SHIFT - AVG_SHIFT = 0 */
    uint32_t ae2 = s->err_sums[nc2];
    uint32_t ec = ae2 < (tot >> 1) ? 0 : ae2 < tot ? 1 : ae2 < (tot << 1) ? 2 : 3;
    c = (c & ~(0x7fu << 21)) | (avg << 21) | (ec << 26);
    s->chan = c;
}

static void lr_init_read (LrState *s, const uint8_t *seq, uint32_t len, int rev)      /* codec_longr_alg_init_read :148-159 */
{
    s->chan = 0;
    for (int32_t i = 0; i < 3; i++)
        lr_update (s, rev ? acgt_code_comp ((int32_t)len - 1 - i >= 0 ? seq[len - 1 - i] : 'T')
                          : acgt_code (i < (int32_t)len ? seq[i] : 'A'), 0, 0);
}

void orc_longr_calc_bins (const uint32_t histogram[256], uint64_t num_values, uint8_t v2b[256])   /* codec_longr.c:66-136 */
{
    uint32_t next_val = 0;
    for (unsigned bin = 0; bin < 32; bin++) {
        uint32_t at_least = (uint32_t)(num_values / (32 - bin));
        if (bin < 11) { num_values -= histogram[next_val]; v2b[next_val] = (uint8_t)bin; next_val++; }
        else {
            uint32_t content = 0;
            while (content < at_least && next_val < 256) {
                content += histogram[next_val]; num_values -= histogram[next_val];
                v2b[next_val] = (uint8_t)bin; next_val++;
            }
        }
    }
    memset (v2b + next_val, 31, 256 - next_val);
}

/* len = quality length of each line; seq_len (NULL = len) its sequence length where that differs: a SAM line without quality
   is the single byte ' ' whatever its seq_len (codec_longr.c:188-192) */
int orc_longr_encode2 (const uint8_t *txt, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *len, const uint32_t *seq_len,
                       const uint8_t *is_rev, uint32_t n_lines, const uint8_t v2b[256],
                       uint8_t *values, uint32_t *lens_be)                            /* codec_longr.c:161-247 */
{
    LrState s; lr_init (&s, v2b);
    uint64_t total = 0;
    for (uint32_t li = 0; li < n_lines; li++) total += len[li];
    uint16_t *base_chan = malloc ((total ? total : 1) * sizeof (uint16_t));
    uint32_t *num = calloc (65536, sizeof (uint32_t));
    uint64_t nb = 0;
    for (uint32_t li = 0; li < n_lines; li++) {                                       /* calc_channels :138-159 */
        uint32_t L = len[li];
        if (!L) continue;
        const uint8_t *seq = txt + seq_off[li], *q = txt + qual_off[li];
        int rev = is_rev ? is_rev[li] : 0;
        const uint32_t Ls = seq_len ? seq_len[li] : L;
        lr_init_read (&s, seq, Ls, rev);
        uint8_t prev = 0;
        for (uint32_t k = 0; k < L; k++) {
            uint32_t i = rev ? L - 1 - k : k;
            uint32_t ch = LR_CHAN (s.chan);
            base_chan[nb++] = (uint16_t)ch; num[ch]++;
            uint8_t b = rev ? acgt_code_comp (i >= 3 ? seq[i - 3] : 'T') : acgt_code (i + 3 < Ls ? seq[i + 3] : 'A');
            uint8_t qq = (uint8_t)(q[i] - '!');
            lr_update (&s, b, qq, prev);
            prev = qq;
        }
    }
    uint32_t *next = malloc (65536 * sizeof (uint32_t));                              /* counting sort by channel :193-230 */
    next[0] = 0;
    for (uint32_t c = 1; c < 65536; c++) next[c] = next[c - 1] + num[c - 1];
    nb = 0;
    for (uint32_t li = 0; li < n_lines; li++) {
        uint32_t L = len[li];
        const uint8_t *q = txt + qual_off[li];
        int rev = is_rev ? is_rev[li] : 0;
        for (uint32_t i = 0; i < L; i++, nb++) values[next[base_chan[nb]]++] = (uint8_t)(q[rev ? L - 1 - i : i] - '!');
    }
    for (uint32_t c = 0; c < 65536; c++) lens_be[c] = __builtin_bswap32 (num[c]);      /* BGEN32 :237-240 */
    free (s.avg_sums); free (s.err_sums); free (base_chan); free (num); free (next);
    return 0;
}

int orc_longr_encode (const uint8_t *txt, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *len,
                      const uint8_t *is_rev, uint32_t n_lines, const uint8_t v2b[256],
                      uint8_t *values, uint32_t *lens_be)
{
    return orc_longr_encode2 (txt, seq_off, qual_off, len, NULL, is_rev, n_lines, v2b, values, lens_be);
}

/* missing (NULL or n_lines bytes): set where a line has no quality — its first value is 255 (codec_longr.c:278): the line takes
   that one value, its first output byte is '*' (sam_reconstruct_missing_quality, sam_qual.c:532) and the rest is left alone */
int orc_longr_decode2 (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev,
                       uint32_t n_lines, const uint8_t v2b[256],
                       const uint8_t *values, const uint32_t *lens_be, uint8_t *qual_out, uint8_t *missing);
int orc_longr_decode (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev,
                      uint32_t n_lines, const uint8_t v2b[256],
                      const uint8_t *values, const uint32_t *lens_be, uint8_t *qual_out)
{
    return orc_longr_decode2 (txt, seq_off, len, is_rev, n_lines, v2b, values, lens_be, qual_out, NULL);
}
int orc_longr_decode2 (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev,
                       uint32_t n_lines, const uint8_t v2b[256],
                       const uint8_t *values, const uint32_t *lens_be, uint8_t *qual_out, uint8_t *missing)   /* codec_longr.c:270-373 */
{
    LrState s; lr_init (&s, v2b);
    uint32_t *next = malloc (65536 * sizeof (uint32_t));                              /* reconstruct_init :301-338 */
    uint32_t acc = 0;
    for (uint32_t c = 0; c < 65536; c++) { next[c] = acc; acc += __builtin_bswap32 (lens_be[c]); }
    for (uint32_t li = 0; li < n_lines; li++) {                                       /* recon_one_read :270-296 */
        uint32_t L = len[li];
        const uint8_t *seq = txt + seq_off[li];
        int rev = is_rev ? is_rev[li] : 0;
        lr_init_read (&s, seq, L, rev);
        uint8_t prev = 0;
        for (uint32_t k = 0; k < L; k++) {
            uint32_t i = rev ? L - 1 - k : k;
            uint8_t b = rev ? acgt_code_comp (i >= 3 ? seq[i - 3] : 'T') : acgt_code (i + 3 < L ? seq[i + 3] : 'A');
            uint8_t qq = values[next[LR_CHAN (s.chan)]++];
            lr_update (&s, b, qq, prev);
            if (missing && qq == 255) { missing[li] = 1; qual_out[0] = '*'; break; }   /* RECON_ONE_QUAL: after the state update */
            prev = qq;
            qual_out[i] = (uint8_t)(qq + '!');
        }
        qual_out += L;
    }
    free (s.avg_sums); free (s.err_sums); free (next);
    return 0;
}

/* ================================================================ NORMQ (reference src/codec_normq.c)
 * encode = codec_normq_compress before its sub-codec (:43-62): the lines' quality strings copied into one buffer, reversed where is_rev.
 * decode = codec_normq_reconstruct (:85-106) line by line: a ' ' at the cursor is a line without quality (one byte consumed, '*' written by
 * sam_reconstruct_missing_quality, sam_qual.c:532); out gets len[i] bytes per line (the '*' at the start of a missing line's slot), *used = the
 * stream bytes consumed.  Returns -1 if the stream runs out. */
uint64_t orc_normq_encode (const uint8_t *txt, const uint64_t *line_off, const uint32_t *line_len, const uint8_t *is_rev, uint32_t n_lines, uint8_t *local)
{
    uint64_t next = 0;
    for (uint32_t li = 0; li < n_lines; li++) {
        const uint32_t L = line_len[li];
        if (!L) continue;                                                                 /* :53 */
        const uint8_t *q = txt + line_off[li];
        if (is_rev && is_rev[li]) for (uint32_t i = 0; i < L; i++) local[next + i] = q[L - 1 - i];   /* str_reverse :55 */
        else memcpy (local + next, q, L);
        next += L;
    }
    return next;
}

int orc_normq_decode (const uint8_t *local, uint64_t local_len, const uint32_t *len, const uint8_t *is_rev, uint32_t n_lines, uint8_t *out, uint8_t *missing, uint64_t *used)
{
    uint64_t next = 0;
    for (uint32_t li = 0; li < n_lines; li++) {
        const uint32_t L = len[li];
        if (missing) missing[li] = 0;
        if (!L) continue;
        if (next >= local_len) return -1;
        if (local[next] == ' ') { out[0] = '*'; if (missing) missing[li] = 1; next++; }  /* :91-94 */
        else {
            if (next + L > local_len) return -1;
            if (is_rev && is_rev[li]) for (uint32_t i = 0; i < L; i++) out[i] = local[next + L - 1 - i];   /* :98 */
            else memcpy (out, local + next, L);
            next += L;
        }
        out += L;
    }
    *used = next;
    return 0;
}

/* ================================================================ OQ (reference src/codec_oq.c)
 * mux   = codec_oq_compress before its sub-codec (:54-121): count pass over the QUAL of every line (:61-72), the OQ characters of the lines whose
 *         SEQ.len is not 0 appended to channel QUAL[i] - 33 (:88-99; oq_off = dl->OQ, 0 for a line without OQ:Z), monochar channels (:103-107).
 *         channels = the 94 channels back to back in channel order (count[q] bytes each; what no line wrote is 0).  Returns -1 on a QUAL
 *         character outside '!'..'~'.
 * demux = codec_oq_reconstruct (:126-164) line by line; channels = the channels that exist, back to back (count[q] bytes each).
 *         Returns -1 when a channel is out of data (:152) or a key is out of range. */
#define OQ_CH 94
int orc_oq_mux (const uint8_t *txt, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *oq_off, const uint32_t *seq_len, uint32_t n_lines,
                uint8_t *channels, uint32_t *count, uint8_t *monochars)
{
    uint64_t next[OQ_CH], total = 0;
    memset (count, 0, OQ_CH * sizeof (uint32_t));
    for (uint32_t li = 0; li < n_lines; li++)                                             /* first pass (:61-72) */
        for (uint32_t i = 0; i < qual_len[li]; i++) {
            const unsigned q = (unsigned)txt[qual_off[li] + i] - 33u;
            if (q >= OQ_CH) return -1;
            count[q]++;
        }
    for (int q = 0; q < OQ_CH; q++) { next[q] = total; total += count[q]; }
    memset (channels, 0, total);
    for (uint32_t li = 0; li < n_lines; li++) {                                           /* second pass (:88-99) */
        if (seq_len && !seq_len[li]) continue;                                            /* :94 */
        const uint8_t *qual = txt + qual_off[li], *oq = txt + oq_off[li];
        for (uint32_t i = 0; i < qual_len[li]; i++) channels[next[qual[i] - 33]++] = oq[i];
    }
    uint64_t at = 0;
    for (int q = 0; q < OQ_CH; q++) {                                                     /* :103-107, str_is_monochar (strings.h:176-184) */
        monochars[q] = 0;
        if (count[q]) {
            int same = 1;
            for (uint32_t i = 1; i < count[q] && same; i++) same = channels[at + i] == channels[at];
            if (same) monochars[q] = channels[at];
        }
        at += count[q];
    }
    return 0;
}

int orc_oq_demux (const uint8_t *txt, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *out_off, uint32_t n_lines, uint32_t key_bias,
                  const uint8_t *channels, const uint32_t *count, const uint8_t *monochars, uint8_t *out)
{
    uint64_t next[OQ_CH], after[OQ_CH], total = 0;
    for (int q = 0; q < OQ_CH; q++) { next[q] = total; total += count[q]; after[q] = total; }
    for (uint32_t li = 0; li < n_lines; li++) {
        const uint8_t *qual = txt + qual_off[li]; uint8_t *recon = out + out_off[li];
        for (uint32_t i = 0; i < qual_len[li]; i++) {
            const unsigned ch = (unsigned)qual[i] - key_bias;                             /* :143 */
            if (ch >= OQ_CH) return -1;
            if (monochars[ch]) recon[i] = monochars[ch];                                  /* :145-146 */
            else { if (next[ch] >= after[ch]) return -1; recon[i] = channels[next[ch]++]; }   /* :149-155 */
        }
    }
    return 0;
}

/* ================================================================ matrix transposes of a local buffer
 * zip = dyn_int_transpose (reference src/dyn_int.c:45-105, no copied samples): a local that is not a rectangle is left alone (:75-78, returns 0),
 *       else trans[c * rows + r] = data[r * cols + c] (:88-91), returns 1.
 * piz = BGEN_transpose_u8/16/32_buf (src/buffer.c:364-391): target[r * cols + c] = transposed[c * rows + r], then every element from big endian. */
int orc_local_transpose (void *data, uint64_t n, uint32_t width, uint32_t cols, int piz)
{
    if (!n || !cols) return 0;
    if (n % cols) return piz ? -1 : 0;
    const uint64_t rows = n / cols;
    uint8_t *in = data, *tmp = malloc (n * width);
    for (uint64_t r = 0; r < rows; r++)
        for (uint64_t c = 0; c < cols; c++) {
            const uint8_t *s = piz ? in + (c * rows + r) * width : in + (r * cols + c) * width;
            uint8_t *d       = piz ? tmp + (r * cols + c) * width : tmp + (c * rows + r) * width;
            for (uint32_t k = 0; k < width; k++) d[k] = piz ? s[width - 1 - k] : s[k];          /* BGEN on a little-endian host */
        }
    memcpy (data, tmp, n * width);
    free (tmp);
    return 1;
}

/* ================================================================ b250_zip_generate (reference src/b250.c:202-297)
 * The backward scan over the segmenter's words (type in the last byte, b250_seg_get_wi :64-86), node index -> word index for the words
 * new to this VBlock (context.h:109-112), ONE_UP (:265-266), b250_set_wi in PIZ format (:89-121).  out (len bytes) gets the result
 * right-aligned; returns its length, or -1 if the words do not tile the buffer / a word index cannot be encoded. */
static int32_t b250_get_wi (const uint8_t *b, int64_t p, int *L)
{
    const uint8_t msb = b[p];
    *L = (msb >> 7) == 0 ? 1 : (msb >> 6) == 2 ? 2 : (msb >> 5) == 6 ? 3 : 4;                     /* VARL_BYTES :47 */
    if (p - *L + 1 < 0) return INT32_MIN;
    if (*L == 1) return msb;
    if (*L == 2) { const uint32_t w = b[p - 1] | ((uint32_t)b[p] << 8); return w == 0xBFFE ? -3 : w == 0xBFFF ? -4 : (int32_t)(w & 0x3fff) + 127; }
    if (*L == 3) return (int32_t)((b[p - 2] | ((uint32_t)b[p - 1] << 8) | ((uint32_t)b[p] << 16)) & 0x1fffff) + 16509;
    return (int32_t)((b[p - 3] | ((uint32_t)b[p - 2] << 8) | ((uint32_t)b[p - 1] << 16) | ((uint32_t)b[p] << 24)) & 0x1fffffff);
}
int64_t orc_b250_generate (const uint8_t *b250, uint64_t len, const int32_t *ni2wi, uint32_t n_new, uint32_t ol_len, int one_up_ok, uint8_t *out, uint64_t *n_words)
{
    int64_t src = (int64_t)len - 1, dst = (int64_t)len;                                            /* dst: one past where the next word ends */
    *n_words = 0;
    while (src >= 0) {
        int L, pL = 0;
        int32_t wi = b250_get_wi (b250, src, &L);
        if (wi == INT32_MIN) return -1;
        if (wi >= (int32_t)ol_len) { if ((uint32_t)wi - ol_len >= n_new) return -1; wi = ni2wi[wi - ol_len]; }
        int32_t prev = -1;                                                                          /* WORD_INDEX_NONE */
        if (src - L >= 0) {
            prev = b250_get_wi (b250, src - L, &pL);
            if (prev == INT32_MIN) return -1;
            if (prev >= (int32_t)ol_len) { if ((uint32_t)prev - ol_len >= n_new) return -1; prev = ni2wi[prev - ol_len]; }
        }
        if (one_up_ok && prev >= 0 && wi >= 0 && wi == prev + 1) wi = -2;                          /* WORD_INDEX_ONE_UP */
        uint32_t enc; int n;
        if (wi >= 0 && wi <= 126) { enc = wi; n = 1; }
        else if (wi >= 127 && wi <= 16508) { enc = (2u << 14) | (uint32_t)(wi - 127); n = 2; }
        else if (wi >= 16509 && wi <= 2113660) { enc = (6u << 21) | (uint32_t)(wi - 16509); n = 3; }
        else if (wi > 2113660 && wi <= (1 << 29) - 1) { enc = (7u << 29) | (uint32_t)wi; n = 4; }
        else if (wi == -2) { enc = 127; n = 1; }
        else if (wi == -3) { enc = 0xBFFE; n = 2; }
        else if (wi == -4) { enc = 0xBFFF; n = 2; }
        else return -1;
        dst -= n;
        for (int k = 0; k < n; k++) out[dst + k] = (uint8_t)(enc >> (8 * (n - 1 - k)));
        src -= L; (*n_words)++;
    }
    return (int64_t)len - dst;
}

/* ================================================================ HOMP and T0 (reference src/codec_homp.c, src/codec_t0.c)
 * mode 0 = HOMP (the string is QUAL), 1 = T0 (the string is t0:Z).
 * condense = the first pass of codec_homp_compress (:132-190) / codec_t0_compress (:69-109): the condensed strings back to back, new_len per line.
 * expand   = codec_homp_reconstruct (:213-276) / codec_t0_reconstruct (:137-179) line by line; returns -1 when the stream does not match. */
static unsigned hp_len_at (const uint8_t *seq, uint32_t len, uint32_t i) { uint32_t k = i + 1; while (k < len && seq[k] == seq[i]) k++; return k - i; }   /* strings.h:194-201 */
uint64_t orc_hp_condense (int mode, const uint8_t *txt, const uint64_t *str_off, const uint32_t *str_len, const uint64_t *seq_off, uint32_t n_lines, uint8_t *local, uint32_t *new_len)
{
    uint64_t at = 0;
    for (uint32_t li = 0; li < n_lines; li++) {
        const uint8_t *s = txt + str_off[li], *seq = txt + seq_off[li]; const uint32_t len = str_len[li];
        uint8_t *d = local + at; uint32_t n = 0;
        if (mode == 0 ? len <= 1 : len == 0) { memcpy (d, s, len); n = len; }                      /* homp :137, t0 :75 */
        else for (uint32_t i = 0; i < len; i++) {
            const unsigned h = hp_len_at (seq, len, i);
            if (h > 1) {
                int ok = 1;
                if (mode == 1) { for (unsigned k = 1; k < h && ok; k++) ok = s[i + k] == s[i]; }     /* t0 :84 */
                else { uint8_t prev = 0; for (unsigned k = 0; k < (h + 1) / 2; k++) { uint8_t a = s[i + k], m = s[i + h - 1 - k]; if (a != m || (prev == 'I' && a != 'I')) { ok = 0; break; } prev = a; } }   /* homp :152-165 */
                if (ok) { if (mode == 1) d[n++] = s[i]; else for (unsigned k = 0; k < (h + 1) / 2; k++) { d[n++] = s[i + k]; if (s[i + k] == 'I') break; } }   /* :167-170 */
                else { d[n++] = s[i] | 0x80; for (unsigned k = 1; k < h; k++) d[n++] = s[i + k]; }  /* :171-176 */
                i += h - 1;
            }
            else d[n++] = s[i];
        }
        if (new_len) new_len[li] = n;
        at += n;
    }
    return at;
}
int orc_hp_expand (int mode, const uint8_t *local, uint64_t local_len, const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, uint32_t n_lines, uint8_t *out, uint8_t *missing)
{
    uint64_t next = 0;
    for (uint32_t li = 0; li < n_lines; li++) {
        const uint32_t L = len[li]; const uint8_t *seq = txt + seq_off[li];
        if (missing) missing[li] = 0;
        if (!L) continue;
        if (next >= local_len) return -1;
        if (mode == 0 && local[next] == ' ') { out[0] = '*'; if (missing) missing[li] = 1; next++; out += L; continue; }   /* homp :241-244 */
        uint32_t n = 0;
        for (uint32_t i = 0; i < L; i++) {
            const unsigned h = hp_len_at (seq, L, i);
            if (next >= local_len) return -1;
            if (h > 1) {
                if (local[next] & 0x80) { if (next + h > local_len) return -1; out[n++] = local[next++] & 0x7f; for (unsigned k = 1; k < h; k++) out[n++] = local[next++]; }
                else if (mode == 1) { const uint8_t c = local[next++]; for (unsigned k = 0; k < h; k++) out[n++] = c; }
                else {
                    uint8_t prev = 0;
                    for (unsigned k = 0; k < (h + 1) / 2; k++) { if (prev != 'I') { if (next >= local_len) return -1; prev = local[next++]; } out[n++] = prev; }
                    uint32_t m = n - 1 - (h & 1);
                    for (unsigned k = 0; k < h / 2; k++) out[n++] = out[m--];
                }
                i += h - 1;
            }
            else out[n++] = local[next++];
        }
        out += L;
    }
    return next == local_len ? 0 : -1;
}

/* ================================================================ SMUX (reference src/codec_smux.c)
 * mux   = codec_smux_compress (:180-262); channels = the 5 channels back to back (count[b] bytes each), *n_param = the fifth channel's character when monochar.
 * demux = codec_smux_reconstruct (:273-355) line by line, output in the read's own orientation; a read without quality writes the '*' of
 *         sam_reconstruct_missing_quality at the start of its slot (slots are len[i] bytes).  -1: a channel is out of data. */
static unsigned smux_enc (uint8_t c, int comp) { unsigned k = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; return (comp && k < 4) ? 3 - k : k; }   /* reference.c:78-84 */
int orc_smux_mux (const uint8_t *txt, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *seq_off, const uint32_t *seq_len, const uint8_t *is_rev,
                  uint32_t n_lines, uint8_t *channels, uint32_t *count, uint8_t *n_param)
{
    uint64_t next[5], total = 0;
    memset (count, 0, 5 * sizeof (uint32_t));
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t li = 0; li < n_lines; li++) {
            const uint8_t *qual = txt + qual_off[li], *seq = txt + seq_off[li]; const uint32_t ql = qual_len[li], sl = seq_len[li];
            const int rev = is_rev ? is_rev[li] : 0;
            if (!ql) continue;
            if (!rev)                      for (uint32_t i = 0; i < ql; i++)  { unsigned b = smux_enc (seq[i], 0); if (pass) channels[next[b]++] = qual[i]; else count[b]++; }
            else if (ql == 1 && qual[0] == ' ')                               { unsigned b = smux_enc (seq[sl - 1], 1); if (pass) channels[next[b]++] = ' '; else count[b]++; }
            else                           for (int32_t i = ql - 1; i >= 0; i--) { unsigned b = smux_enc (seq[i], 1); if (pass) channels[next[b]++] = qual[i]; else count[b]++; }
        }
        if (!pass) for (int b = 0; b < 5; b++) { next[b] = total; total += count[b]; }
    }
    *n_param = 0;
    if (count[4]) {                                                                     /* :246-253 */
        const uint8_t *c = channels + total - count[4]; int same = 1;
        for (uint32_t i = 1; i < count[4] && same; i++) same = c[i] == c[0];
        if (same) *n_param = c[0];
    }
    return 0;
}
int orc_smux_demux (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev, const uint64_t *out_off, uint32_t n_lines,
                    const uint8_t *channels, const uint32_t *count, uint8_t n_param, uint8_t *out, uint8_t *missing)
{
    uint64_t next[5], after[5], total = 0;
    for (int b = 0; b < 5; b++) { next[b] = total; total += count[b]; after[b] = total; }
    for (uint32_t li = 0; li < n_lines; li++) {
        const uint8_t *seq = txt + seq_off[li]; uint8_t *recon = out + out_off[li];
        uint32_t L = len[li]; const int rev = is_rev ? is_rev[li] : 0;
        if (missing) missing[li] = 0;
        if (!L) continue;
        if (seq[0] == '*') L = 1;                                                       /* :278-279 */
        for (uint32_t t = 0; t < L; t++) {
            const uint32_t i = rev ? L - 1 - t : t;
            const unsigned b = smux_enc (seq[i], rev);
            if (b == 4 && n_param) recon[i] = n_param;
            else { if (next[b] >= after[b]) return -1; recon[i] = channels[next[b]++]; }
            if (recon[i] == ' ') { recon[i] = 0; recon[0] = '*'; if (missing) missing[li] = 1; break; }   /* :307-310 */
        }
    }
    return 0;
}

/* ================================================================ TMPL (reference src/codec_tmpl.c)
 * mux   = codec_tmpl_compress (:145-210): channels 0..93 by the template, 94 = the excess beyond the template ((ctx+1)->local), back to back.
 * demux = codec_tmpl_reconstruct (:216-259) line by line.  -1: a channel is out of data. */
int orc_tmpl_mux (const uint8_t *txt, const uint64_t *qual_off, const uint32_t *qual_len, uint32_t n_lines, const uint8_t *tmpl, uint32_t tmpl_len, uint8_t *channels, uint32_t *count)
{
    uint64_t next[95], total = 0;
    memset (count, 0, 95 * sizeof (uint32_t));
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t li = 0; li < n_lines; li++) {
            const uint8_t *q = txt + qual_off[li];
            for (uint32_t i = 0; i < qual_len[li]; i++) {
                const unsigned ch = i < tmpl_len ? (unsigned)tmpl[i] - 33u : 94u;
                if (ch > 94) return -1;
                if (pass) channels[next[ch]++] = q[i]; else count[ch]++;
            }
        }
        if (!pass) for (int c = 0; c < 95; c++) { next[c] = total; total += count[c]; }
    }
    return 0;
}
int orc_tmpl_demux (const uint32_t *len, const uint64_t *out_off, uint32_t n_lines, const uint8_t *tmpl, uint32_t tmpl_len, const uint8_t *channels, const uint32_t *count, uint8_t *out)
{
    uint64_t next[95], after[95], total = 0;
    for (int c = 0; c < 95; c++) { next[c] = total; total += count[c]; after[c] = total; }
    for (uint32_t li = 0; li < n_lines; li++)
        for (uint32_t i = 0; i < len[li]; i++) {
            const unsigned ch = i < tmpl_len ? (unsigned)tmpl[i] - 33u : 94u;
            if (ch > 94 || next[ch] >= after[ch]) return -1;
            out[out_off[li] + i] = channels[next[ch]++];
        }
    return 0;
}

/* ================================================================ PACB (reference src/codec_pacb.c)
 * mux   = codec_pacb_compress (:164-236): channel = 7 * np0 + QUAL_get_K_value (:19-27); channels back to back, count[7 * max_np].
 * demux = codec_pacb_reconstruct (:262-326) line by line; a read without quality writes '*' at the start of its slot.  -1: a channel is out of data. */
static unsigned pacb_K (const uint8_t *seq, uint32_t len, uint32_t i)
{
    const uint8_t b = seq[i]; const unsigned at = (b == 'A' || b == 'T');
    if (i > 0 && seq[i - 1] == b) return 6;
    if (i == len - 1 || seq[i + 1] != b) return 4 + at;
    if (i == len - 2 || seq[i + 2] != b) return 2 + at;
    return at;
}
int orc_pacb_mux (const uint8_t *txt, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *seq_off, const uint8_t *np0, uint32_t max_np, uint32_t n_lines,
                  uint8_t *channels, uint32_t *count)
{
    const uint32_t n_ch = 7 * max_np;
    uint64_t next[84], total = 0;
    memset (count, 0, 84 * sizeof (uint32_t));
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t li = 0; li < n_lines; li++) {
            const uint8_t *q = txt + qual_off[li], *seq = txt + seq_off[li]; const uint32_t L = qual_len[li];   /* a read without quality: L = 1 (:135) */
            const unsigned base = (np0 && max_np > 1) ? 7u * np0[li] : 0;
            for (uint32_t i = 0; i < L; i++) {
                const unsigned ch = base + pacb_K (seq, L, i);
                if (ch >= n_ch) return -1;
                if (pass) channels[next[ch]++] = q[i]; else count[ch]++;
            }
        }
        if (!pass) for (uint32_t c = 0; c < n_ch; c++) { next[c] = total; total += count[c]; }
    }
    return 0;
}
int orc_pacb_demux (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *np0, uint32_t max_np, const uint64_t *out_off, uint32_t n_lines,
                    const uint8_t *channels, const uint32_t *count, uint8_t *out)
{
    const uint32_t n_ch = 7 * max_np;
    uint64_t next[84], after[84], total = 0;
    for (uint32_t c = 0; c < n_ch; c++) { next[c] = total; total += count[c]; after[c] = total; }
    for (uint32_t li = 0; li < n_lines; li++) {
        const uint8_t *seq = txt + seq_off[li]; uint8_t *recon = out + out_off[li];
        const unsigned base = (np0 && max_np > 1) ? 7u * np0[li] : 0;
        for (uint32_t i = 0; i < len[li]; i++) {
            const unsigned ch = base + pacb_K (seq, len[li], i);
            if (ch >= n_ch || next[ch] >= after[ch]) return -1;
            const uint8_t score = channels[next[ch]++];
            if (score == ' ') { recon[0] = '*'; break; }                                   /* :313-317 */
            recon[i] = score;
        }
    }
    return 0;
}
