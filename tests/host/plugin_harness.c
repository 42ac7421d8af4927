/* tests/host/plugin_harness.c — a stand-in for the genozip side of the drop-in boundary: TEST INFRASTRUCTURE.
 *
 * It has its own VBlock / Context / Buffer / SectionHeader (the library never looks inside them), registers the two accessor
 * tables (gzb_plugin_host, gzb_plugin_host2) the adapter inside genozip would register, owns a codec table whose simple rows
 * point at the library's plug-in entry points — exactly what CODEC_ARGS (src/codec.h:47-115) would hold — and drives every
 * gzb_codec_* function the way genozip's compute thread does: comp_compress with the soft-fail retry of src/compressor.c:82-110,
 * the line callback of src/codec_htscodecs.c:51-64, piz's uncompress / per-line reconstruct sequence.  The scenarios return what
 * the codecs produced; tests/test_plugin_harness.py compares it with the reference's compiled objects (oracle/_ref).
 *
 * Build: gcc -shared -fPIC plugin_harness.c -I include -L<dir> -l:<libgzb200.so | libgzb200_simt.so>    (tests/test_plugin_harness.py does it) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include "gzb200.h"

struct Buffer  { char *data; uint64_t len, size; uint8_t prm8[8]; };
struct Context { struct Buffer local, packed; bool acgt_no_x; void *state; Codec lcodec, lsubcodec; uint32_t HT_n_lines, ht_per_line; uint8_t table[1 + 95 * 95]; };
union  SectionHeaderUnion { struct { Codec sub_codec; bool acgt_no_x; } h; };
struct VBlock  {
    uint32_t vblock_i, n_lines;
    struct Buffer scratch, txt_out;                       /* txt_out: reconstruction target (vb->txt_data in PIZ) */
    struct Context ctx[8];
    /* line tables of the scenario */
    char **qual, **seq; uint32_t *qual_len, *seq_len; uint8_t *is_rev;
    const uint32_t *recon_lens; const char *seq_txt; uint64_t seq_txt_len; const uint64_t *seq_off;
    int64_t big_allele;
    uint64_t lines_counter[6], time_ns[8];
    int n_missing;
};

static jmp_buf on_abort; static char abort_text[512];
static void h_abort (const char *msg) { snprintf (abort_text, sizeof abort_text, "%s", msg); longjmp (on_abort, 1); }
const char *harness_last_abort (void) { return abort_text; }

static char *buf_alloc_ (struct Buffer *b, uint64_t bytes)
{
    if (bytes > b->size) { b->data = realloc (b->data, bytes + 64); b->size = bytes; }
    return b->data;
}
static void buf_free_ (struct Buffer *b) { free (b->data); memset (b, 0, sizeof *b); }

/* ---------------------------------------------------------------- accessor table 1 */
static uint32_t a_num_lines (VBlockP vb) { return vb->n_lines; }
static uint32_t a_vblock_i (VBlockP vb) { return vb->vblock_i; }
static char *a_buffer_data (BufferP b) { return b->data; }

/* ---------------------------------------------------------------- the codec table (what CODEC_ARGS would hold) */
#define C_NONE 1
#define C_LZMA 4
static GZB_COMPRESS (store_compress)
{
    (void)vb; (void)ctx; (void)header; (void)get_line_cb; (void)name;
    if (*compressed_len < *uncompressed_len) { if (soft_fail) return false; h_abort ("store: buffer too small"); }
    memcpy (compressed, uncompressed, *uncompressed_len); *compressed_len = *uncompressed_len;
    return true;
}
static GZB_UNCOMPRESS (store_uncompress)
{
    (void)vb; (void)ctx; (void)codec; (void)param; (void)sub_codec; (void)name;
    if (compressed_len != uncompressed_len) h_abort ("store: length mismatch");
    memcpy (uncompressed_buf->data, compressed, compressed_len);
}
static uint32_t store_est_size (Codec c, uint64_t n) { (void)c; return (uint32_t)n + 16; }

typedef bool CompressFn (VBlockP, ContextP, SectionHeaderP, const char *, uint32_t *, LocalGetLineCB, char *, uint32_t *, FailType, const char *);
typedef void UncompressFn (VBlockP, ContextP, Codec, uint8_t, const char *, uint32_t, BufferP, uint64_t, Codec, const char *);
typedef uint32_t EstFn (Codec, uint64_t);
static struct { CompressFn *compress; UncompressFn *uncompress; EstFn *est_size; } codec_args[32];
static void codec_table_init (void)
{
    if (codec_args[C_NONE].compress) return;
    codec_args[C_NONE].compress = store_compress; codec_args[C_NONE].uncompress = store_uncompress; codec_args[C_NONE].est_size = store_est_size;
    codec_args[C_LZMA] = codec_args[C_NONE];              /* LZMA stays on the host (out of scope): the harness stores */
#define ROW(id, NAME, unc) codec_args[id].compress = gzb_codec_##NAME##_compress; codec_args[id].uncompress = unc; codec_args[id].est_size = gzb_codec_##NAME##_est_size;
    ROW (GZB_CODEC_RANB, RANB, gzb_codec_rans_uncompress) ROW (GZB_CODEC_RANW, RANW, gzb_codec_rans_uncompress)
    ROW (GZB_CODEC_RANb, RANb, gzb_codec_rans_uncompress) ROW (GZB_CODEC_RANw, RANw, gzb_codec_rans_uncompress)
    ROW (GZB_CODEC_ARTB, ARTB, gzb_codec_arith_uncompress) ROW (GZB_CODEC_ARTW, ARTW, gzb_codec_arith_uncompress)
    ROW (GZB_CODEC_ARTb, ARTb, gzb_codec_arith_uncompress) ROW (GZB_CODEC_ARTw, ARTw, gzb_codec_arith_uncompress)
}

/* ---------------------------------------------------------------- accessor table 2 */
static char *a_local_alloc (VBlockP vb, ContextP c, int s, uint64_t bytes) { (void)vb; return buf_alloc_ (&c[s].local, bytes); }
static char *a_local_data (ContextP c, int s, uint64_t *len) { *len = c[s].local.len; return c[s].local.data; }
static void a_local_set_len (ContextP c, int s, uint64_t len) { c[s].local.len = len; }
static void a_local_free (VBlockP vb, ContextP c, int s) { (void)vb; buf_free_ (&c[s].local); }
static uint8_t *a_local_prm8 (ContextP c, int s) { return &c[s].local.prm8[0]; }
static char *a_scratch_alloc (VBlockP vb, uint64_t bytes) { return bytes ? buf_alloc_ (&vb->scratch, bytes) : vb->scratch.data; }
static void a_scratch_free (VBlockP vb) { buf_free_ (&vb->scratch); }
static bool a_acgt_no_x (ContextP c, int s) { return c[s].acgt_no_x; }
static void a_header_set (SectionHeaderP h, int field, uint32_t v) { if (field == GZB_HDR_SUB_CODEC) h->h.sub_codec = (Codec)v; else if (field == GZB_HDR_ACGT_NO_X) h->h.acgt_no_x = v != 0; }
static Codec g_assign = GZB_CODEC_ARTB;                   /* the scenario's stand-in for codec_assign_best_codec */
static Codec a_assign (VBlockP vb, ContextP c, int s) { (void)vb; c[s].lsubcodec = c[s].local.len >= 50 ? g_assign : C_NONE; return c[s].lsubcodec; }
static bool a_sub_compress (Codec sc, VBlockP vb, ContextP c, SectionHeaderP h, const char *d, uint32_t *len, char *comp, uint32_t *clen, FailType sf, const char *name)
{ return codec_args[sc].compress (vb, c, h, d, len, NULL, comp, clen, sf, name); }
static uint32_t a_sub_est (Codec sc, uint64_t n) { return codec_args[sc].est_size (sc, n); }
static void a_sub_uncompress (Codec sc, VBlockP vb, ContextP c, uint8_t prm, const char *comp, uint32_t clen, BufferP b, uint64_t n, const char *name)
{ buf_alloc_ (b, n); codec_args[sc].uncompress (vb, c, sc, prm, comp, clen, b, n, 0, name); b->len = n; }
static void a_seg_denorm (VBlockP vb, ContextP q, const uint8_t *den, uint32_t len) { (void)vb; q[1].table[0] = (uint8_t)(len / (q[0].local.prm8[0] & 0x7f)); memcpy (q[1].table + 1, den, len); }
static bool a_seq_line (VBlockP vb, ContextP c, uint32_t i, char **s, uint32_t *l, bool *rev) { (void)c; *s = vb->seq[i]; *l = vb->seq_len[i]; if (rev) *rev = vb->is_rev ? vb->is_rev[i] : 0; return true; }
static const uint8_t *a_codec_table (VBlockP vb, ContextP c) { (void)vb; return c[1].table; }       /* LONGR: value_to_bin of ctx+1; DOMQ: table segged into DOMQRUNS */
static void a_pbwt_dims (VBlockP vb, ContextP c, uint32_t *nl, uint32_t *w, int set)
{ struct Context *ht = &vb->ctx[0]; (void)c; if (set) ht->ht_per_line = *w; else { *nl = ht->HT_n_lines; *w = ht->ht_per_line; } }
static void a_add_lines (int which, uint64_t n) { (void)which; (void)n; }
static VBlockP g_vb;
static void a_add_lines_vb (int which, uint64_t n) { if (g_vb) g_vb->lines_counter[which] += n; a_add_lines (which, n); }
static void a_account (VBlockP vb, int which, uint64_t ns) { vb->time_ns[which] += ns; }
static BufferP a_packed_buffer (VBlockP vb, ContextP c, int s, uint64_t bytes) { struct Buffer *b = c[s].acgt_no_x ? &vb->scratch : &c[s].packed; if (bytes) buf_alloc_ (b, bytes); return b; }
static void **a_state (VBlockP vb, ContextP c) { (void)vb; return &c->state; }
static const uint32_t *a_recon_lens (VBlockP vb, ContextP c, uint32_t *n) { (void)c; *n = vb->n_lines; return vb->recon_lens; }
static bool a_recon_seq_table (VBlockP vb, ContextP c, const char **txt, uint64_t *len, const uint64_t **off, const uint8_t **rev)
{ (void)c; *txt = vb->seq_txt; *len = vb->seq_txt_len; *off = vb->seq_off; *rev = vb->is_rev; return vb->seq_txt != NULL; }
static char *a_recon_at (VBlockP vb) { return vb->txt_out.data + vb->txt_out.len; }
static void a_recon_advance (VBlockP vb, int32_t n) { vb->txt_out.len += n; }
static int64_t a_big_allele (VBlockP vb) { return vb->big_allele; }
static bool a_drop (VBlockP vb) { (void)vb; return false; }
static void a_update_line_len (VBlockP vb, ContextP c, uint32_t i, uint32_t n) { (void)c; vb->qual_len[i] = n; }   /* sam_update_qual_len / sam_ultima_update_t0_len */
static void a_missing_quality (VBlockP vb, bool reconstruct) { if (reconstruct) { *a_recon_at (vb) = '*'; vb->txt_out.len++; } vb->n_missing++; }   /* sam_reconstruct_missing_quality */

static void harness_init (void)
{
    static const gzb_plugin_host h1 = { a_num_lines, a_vblock_i, a_buffer_data, h_abort };
    static const gzb_plugin_host2 h2 = {
        a_local_alloc, a_local_data, a_local_set_len, a_local_free, a_local_prm8, a_scratch_alloc, a_scratch_free, a_acgt_no_x, a_header_set,
        a_assign, a_sub_compress, a_sub_est, a_seg_denorm, a_seq_line, a_codec_table, a_pbwt_dims, a_add_lines_vb, a_account,
        a_sub_uncompress, a_packed_buffer, a_state, a_recon_lens, a_recon_seq_table, a_recon_at, a_recon_advance, a_big_allele, a_drop, a_missing_quality, a_update_line_len };
    codec_table_init ();
    gzb_plugin_register (&h1, 0);
    gzb_plugin_register2 (&h2);
}
void harness_set_combining (int on, uint32_t linger_us) { gzb_plugin_set_combining (on, linger_us); }
void harness_shutdown (void) { gzb_plugin_shutdown (); }

static void vb_free (VBlockP vb)
{
    for (int i = 0; i < 8; i++) { buf_free_ (&vb->ctx[i].local); buf_free_ (&vb->ctx[i].packed); }
    buf_free_ (&vb->scratch); buf_free_ (&vb->txt_out);
    free (vb->qual); free (vb->seq); free (vb);
}

/* the line callbacks the segmenter's data types supply (fastq_zip_qual, src/fastq_qual.c:52-77) */
static void cb_qual (VBlockP vb, ContextP ctx, uint32_t i, char **d, uint32_t *l, uint32_t max, bool *rev)
{ (void)ctx; (void)max; *d = vb->qual[i]; *l = vb->qual_len[i]; if (rev) *rev = vb->is_rev ? vb->is_rev[i] : 0; }

/* comp_compress (src/compressor.c:82-110): the compressed buffer starts at 64 bytes when `tight`, and is grown after the soft fail */
static int comp_compress (CompressFn *compress, EstFn *est, Codec codec, VBlockP vb, ContextP ctx, SectionHeaderP header, const char *data, uint32_t *ulen,
                          LocalGetLineCB cb, int tight, char **out, uint32_t *out_len, int *n_soft_fails)
{
    uint32_t cap = est (codec, *ulen);
    if (tight) cap = 64;                                  /* too small on purpose: the first call soft-fails */
    char *z = malloc ((size_t)cap + 16);
    uint32_t clen = cap;
    bool ok = compress (vb, ctx, header, data, ulen, cb, z, &clen, SOFT_FAIL, "harness");
    if (!ok) {                                            /* :90-110: more memory, then once more without soft_fail */
        (*n_soft_fails)++;
        cap = est (codec, *ulen) + 1024; z = realloc (z, (size_t)cap + 16); clen = cap;
        ok = compress (vb, ctx, header, data, ulen, cb, z, &clen, HARD_FAIL, "harness");
        if (!ok) { free (z); return -2; }
    }
    *out = z; *out_len = clen;
    return 0;
}

/* ================================================================ scenarios (called from tests/test_plugin_harness.py) */

/* a simple codec through its plug-in entry points: contiguous or line by line; returns the section and the round trip */
int harness_simple (int codec, const uint8_t *data, uint32_t n, const uint32_t *line_lens, uint32_t n_lines, int tight,
                    uint8_t *comp_out, uint32_t *comp_len, uint8_t *back, int *n_soft_fails)
{
    harness_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof *vb); vb->vblock_i = 1;
    char **lines = NULL;
    if (line_lens) {
        vb->n_lines = n_lines; lines = malloc (n_lines * sizeof *lines); vb->qual = lines; vb->qual_len = (uint32_t *)line_lens;
        uint64_t o = 0; for (uint32_t i = 0; i < n_lines; i++) { lines[i] = (char *)data + o; o += line_lens[i]; }
    }
    union SectionHeaderUnion hdr = { { 0 } };
    uint32_t ulen = n; char *z = NULL; *n_soft_fails = 0;
    int rc = comp_compress (codec_args[codec].compress, codec_args[codec].est_size, (Codec)codec, vb, &vb->ctx[0], &hdr, line_lens ? NULL : (const char *)data, &ulen,
                            line_lens ? cb_qual : NULL, tight, &z, comp_len, n_soft_fails);
    if (rc) { vb->qual = NULL; vb_free (vb); return rc; }
    memcpy (comp_out, z, *comp_len);
    struct Buffer ub = { 0 }; buf_alloc_ (&ub, n);
    codec_args[codec].uncompress (vb, &vb->ctx[0], (Codec)codec, 0, z, *comp_len, &ub, n, 0, "harness");
    memcpy (back, ub.data, n);
    free (z); buf_free_ (&ub); vb->qual_len = NULL; vb_free (vb);
    return 0;
}

/* ACGT / XCGT: codec_acgt_compress on NONREF, then NONREF_X as the next context through XCGT's sub-codec; PIZ: both uncompressors */
int harness_acgt (const uint8_t *seq, uint32_t n, int by_lines, const uint32_t *line_lens, uint32_t n_lines, int sub_codec,
                  uint8_t *packed_out, uint32_t *packed_len, uint8_t *x_out, int *no_x, uint8_t *x_comp, uint32_t *x_comp_len, uint8_t *back, int *n_soft_fails)
{
    harness_init ();
    if (setjmp (on_abort)) return -1;
    g_assign = (Codec)sub_codec;
    VBlockP vb = calloc (1, sizeof *vb); vb->vblock_i = 2;
    struct Context *nonref = &vb->ctx[0];
    char *copy = buf_alloc_ (&nonref->local, n + 16); memcpy (copy, seq, n); nonref->local.len = n;
    char **lines = NULL;
    if (by_lines) {
        vb->n_lines = n_lines; lines = malloc (n_lines * sizeof *lines); vb->qual = lines; vb->qual_len = (uint32_t *)line_lens;
        uint64_t o = 0; for (uint32_t i = 0; i < n_lines; i++) { lines[i] = copy + o; o += line_lens[i]; }
    }
    else { nonref[1].local.data = NULL; }
    /* the adapter overlays NONREF_X.local on NONREF.local for contiguous data (src/codec_acgt.c:97-100): here NONREF_X gets its own buffer */
    union SectionHeaderUnion hdr = { { 0 } };
    uint32_t ulen = n; char *z = NULL; *n_soft_fails = 0;
    int rc = comp_compress (gzb_codec_acgt_compress, gzb_codec_complex_est_size, GZB_CODEC_ACGT, vb, nonref, &hdr, by_lines ? NULL : copy, &ulen,
                            by_lines ? cb_qual : NULL, 0, &z, packed_len, n_soft_fails);
    if (rc) return rc;
    memcpy (packed_out, z, *packed_len);                  /* LZMA is stored by the harness: these are the 2-bit words */
    *no_x = hdr.h.acgt_no_x;
    char *zx = NULL; *x_comp_len = 0;
    if (!*no_x) {
        memcpy (x_out, nonref[1].local.data, n);
        uint32_t xl = n; Codec sc = nonref[1].lsubcodec; int sf = 0;
        rc = comp_compress (codec_args[sc].compress, codec_args[sc].est_size, sc, vb, &nonref[1], &hdr, nonref[1].local.data, &xl, NULL, 1, &zx, x_comp_len, &sf);
        if (rc) return rc;
        *n_soft_fails += sf;
        memcpy (x_comp, zx, *x_comp_len);
    }
    /* PIZ: a fresh VBlock; NONREF.local is allocated by the caller ahead of comp_uncompress */
    VBlockP pv = calloc (1, sizeof *pv); pv->vblock_i = 2;
    struct Context *pn = &pv->ctx[0];
    pn->acgt_no_x = *no_x;
    buf_alloc_ (&pn->local, n + 16); pn->local.len = n;
    gzb_codec_acgt_uncompress (pv, pn, GZB_CODEC_ACGT, 0, z, *packed_len, &pn->local, n, hdr.h.sub_codec, "harness");
    if (!*no_x) {
        buf_alloc_ (&pn[1].local, n + 16);
        gzb_codec_xcgt_uncompress (pv, &pn[1], GZB_CODEC_XCGT, 0, zx, *x_comp_len, &pn[1].local, n, nonref[1].lsubcodec, "harness");
    }
    memcpy (back, pn->local.data, n);
    free (z); free (zx); vb->qual_len = NULL; vb_free (vb); vb_free (pv);
    return 0;
}

/* DOMQ: comp_init at seg time, compress with the soft-fail retry, then the four contexts; PIZ: one reconstruct call per line */
int harness_domq (const uint8_t *txt, const uint64_t *off, const uint32_t *lens, uint32_t n_lines, int sub_codec, int force,
                  uint8_t *qual, uint32_t *qual_len, uint8_t *runs, uint32_t *runs_len, uint8_t *mplx, uint32_t *mplx_len, uint8_t *divr, uint32_t *divr_len,
                  uint8_t *denorm, uint32_t *denorm_len, uint8_t *param, uint8_t *comp, uint32_t *comp_len, uint8_t *back, int *n_soft_fails, uint64_t *counters)
{
    harness_init ();
    if (setjmp (on_abort)) return -1;
    g_assign = (Codec)sub_codec;
    VBlockP vb = calloc (1, sizeof *vb); vb->vblock_i = 3; vb->n_lines = n_lines; g_vb = vb;
    vb->qual = malloc (n_lines * sizeof (char *)); vb->qual_len = (uint32_t *)lens;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) { vb->qual[i] = (char *)txt + off[i]; total += lens[i]; }
    struct Context *q = &vb->ctx[0];
    if (!gzb_codec_domq_comp_init (vb, q, cb_qual, force != 0)) { g_vb = NULL; vb->qual_len = NULL; vb_free (vb); return 1; }   /* not a fit */
    *param = q->local.prm8[0];
    *denorm_len = (uint32_t)q[1].table[0] * (*param & 0x7f); memcpy (denorm, q[1].table + 1, *denorm_len);
    union SectionHeaderUnion hdr = { { 0 } };
    uint32_t ulen = (uint32_t)total; char *z = NULL; *n_soft_fails = 0;
    int rc = comp_compress (gzb_codec_domq_compress, gzb_codec_complex_est_size, GZB_CODEC_DOMQ, vb, q, &hdr, NULL, &ulen, cb_qual, 1, &z, comp_len, n_soft_fails);
    if (rc) return rc;
    memcpy (comp, z, *comp_len);
#define OUT(k, p, l) *l = (uint32_t)q[k].local.len; if (*l) memcpy (p, q[k].local.data, *l);
    OUT (0, qual, qual_len) OUT (1, runs, runs_len) OUT (2, mplx, mplx_len) OUT (3, divr, divr_len)
    memcpy (counters, vb->lines_counter, 4 * sizeof (uint64_t));
    /* PIZ: the sub-codec has already put QUAL.local back (here: it never left); the other three come from their own sections */
    vb->recon_lens = lens;
    buf_alloc_ (&vb->txt_out, total + 64);
    for (uint32_t i = 0; i < n_lines; i++)
        if (lens[i]) gzb_codec_domq_reconstruct (vb, GZB_CODEC_DOMQ, q, lens[i], true);                     /* (empty QUAL lines are not routed to the codec) */
    if (vb->txt_out.len != total) return -5;
    memcpy (back, vb->txt_out.data, total);
    free (z); g_vb = NULL; vb->qual_len = NULL; vb_free (vb);
    return 0;
}

/* NORMQ (SAM / BAM fallback): compress through comp_compress with the soft-fail retry (zip-side lengths: 1 for a line without quality, the
 * byte ' '), then one reconstruct call per line with the read's seq_len; strands through the line callback (ZIP) and recon_seq_table (PIZ) */
int harness_normq (const uint8_t *txt, const uint64_t *off, const uint32_t *zlens, const uint32_t *seq_lens, const uint8_t *is_rev, uint32_t n_lines, int sub_codec,
                   uint8_t *local, uint64_t *local_len, uint8_t *comp, uint32_t *comp_len, uint8_t *back, uint64_t *back_len, int *n_soft_fails, int *n_missing, uint64_t *normq_lines)
{
    harness_init ();
    if (setjmp (on_abort)) return -1;
    g_assign = (Codec)sub_codec;
    VBlockP vb = calloc (1, sizeof *vb); vb->vblock_i = 5; vb->n_lines = n_lines; g_vb = vb;
    vb->qual = malloc (n_lines * sizeof (char *)); vb->qual_len = (uint32_t *)zlens; vb->is_rev = (uint8_t *)is_rev;
    uint64_t total = 0, out_total = 0;
    for (uint32_t i = 0; i < n_lines; i++) { vb->qual[i] = (char *)txt + off[i]; total += zlens[i]; out_total += seq_lens[i]; }
    struct Context *q = &vb->ctx[0];
    q->local.len = total;                                  /* callback-mode locals carry only their total length */
    union SectionHeaderUnion hdr = { { 0 } };
    uint32_t ulen = (uint32_t)total; char *z = NULL; *n_soft_fails = 0;
    int rc = comp_compress (gzb_codec_normq_compress, gzb_codec_complex_est_size, 30 /* CODEC_NORMQ */, vb, q, &hdr, NULL, &ulen, cb_qual, 1, &z, comp_len, n_soft_fails);
    if (rc) return rc;
    memcpy (comp, z, *comp_len);
    *local_len = q->local.len; if (q->local.len) memcpy (local, q->local.data, q->local.len);
    *normq_lines = vb->lines_counter[3];
    /* PIZ: the sub-codec has put QUAL.local back (here: it never left) */
    vb->recon_lens = seq_lens; vb->seq_txt = "";           /* (recon_seq_table: only the strands are looked at) */
    buf_alloc_ (&vb->txt_out, out_total + 64);
    for (uint32_t i = 0; i < n_lines; i++)
        if (seq_lens[i]) gzb_codec_normq_reconstruct (vb, 30, q, seq_lens[i], true);
    *back_len = vb->txt_out.len; memcpy (back, vb->txt_out.data, vb->txt_out.len);
    *n_missing = vb->n_missing;
    free (z); g_vb = NULL; vb->qual_len = NULL; vb->is_rev = NULL; vb_free (vb);
    return 0;
}

/* HOMP (mode 0) / T0 (mode 1): the lines condensed in place with their lengths updated, the sub-codec on the condensed strings, then one reconstruct call per line */
int harness_homp (int mode, const uint8_t *txt, uint64_t txt_len, const uint64_t *str_off, const uint64_t *seq_off, const uint32_t *lens, uint32_t n_lines, int sub_codec,
                  uint8_t *local, uint64_t *local_len, uint32_t *new_lens, uint8_t *comp, uint32_t *comp_len, uint8_t *back, uint64_t *back_len,
                  int *n_soft_fails, int *n_missing, uint64_t *homp_lines)
{
    harness_init ();
    if (setjmp (on_abort)) return -1;
    g_assign = (Codec)sub_codec;
    VBlockP vb = calloc (1, sizeof *vb); vb->vblock_i = 6; vb->n_lines = n_lines; g_vb = vb;
    char *work = malloc (txt_len + 16); memcpy (work, txt, txt_len);             /* the codec condenses the strings in place */
    vb->qual = malloc (n_lines * sizeof (char *)); vb->seq = malloc (n_lines * sizeof (char *));
    vb->qual_len = malloc ((n_lines + 1) * 4); vb->seq_len = malloc ((n_lines + 1) * 4);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) { vb->qual[i] = work + str_off[i]; vb->seq[i] = work + seq_off[i]; vb->qual_len[i] = vb->seq_len[i] = lens[i]; total += lens[i]; }
    struct Context *q = &vb->ctx[0];
    q->local.len = total;                                  /* callback-mode locals carry only their total length */
    union SectionHeaderUnion hdr = { { 0 } };
    uint32_t ulen = (uint32_t)total; char *z = NULL; *n_soft_fails = 0;
    int rc = comp_compress (mode ? gzb_codec_t0_compress : gzb_codec_homp_compress, gzb_codec_complex_est_size, mode ? 28 /* CODEC_T0 */ : 27 /* CODEC_HOMP */, vb, q, &hdr, NULL, &ulen, cb_qual, 1, &z, comp_len, n_soft_fails);
    if (rc) return rc;
    memcpy (comp, z, *comp_len);
    /* what the sub-codec saw = the lines as they are now, one after the other */
    uint64_t at = 0;
    for (uint32_t i = 0; i < n_lines; i++) { memcpy (local + at, vb->qual[i], vb->qual_len[i]); at += vb->qual_len[i]; new_lens[i] = vb->qual_len[i]; }
    *local_len = at;
    if (q->local.len != at) return -5;                     /* ctx->local.len32 kept in step (codec_homp.c:184) */
    *homp_lines = vb->lines_counter[4];
    /* PIZ: the sub-codec put the condensed strings back into the context's local; SEQ of every read is known up front */
    memcpy (buf_alloc_ (&q->local, at + 16), local, at); q->local.len = at;
    vb->recon_lens = lens; vb->seq_txt = (const char *)txt; vb->seq_txt_len = txt_len; vb->seq_off = seq_off;
    buf_alloc_ (&vb->txt_out, total + 64);
    for (uint32_t i = 0; i < n_lines; i++)
        if (lens[i]) (mode ? gzb_codec_t0_reconstruct : gzb_codec_homp_reconstruct) (vb, mode ? 28 : 27, q, lens[i], true);
    *back_len = vb->txt_out.len; memcpy (back, vb->txt_out.data, vb->txt_out.len);
    *n_missing = vb->n_missing;
    free (z); free (work); free (vb->qual_len); free (vb->seq_len); vb->qual_len = NULL; vb->seq_len = NULL; g_vb = NULL; vb_free (vb);
    return 0;
}

/* PBWT: compress, then FGRC as stored (big endian) through uncompress, then one reconstruct call per haplotype */
int harness_pbwt (const uint8_t *ht, uint32_t n_lines, uint32_t w, uint32_t *runs, uint32_t *n_runs, uint32_t *fgrc, uint32_t *n_fgrc,
                  uint8_t *back, char *text, uint32_t *text_len)
{
    harness_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof *vb); vb->vblock_i = 4;
    struct Context *c = &vb->ctx[0];
    const uint64_t len = (uint64_t)n_lines * w;
    memcpy (buf_alloc_ (&c->local, len + 16), ht, len); c->local.len = len; c->HT_n_lines = n_lines; c->ht_per_line = w;
    union SectionHeaderUnion hdr = { { 0 } };
    uint32_t ulen = (uint32_t)len, clen = 0;
    if (!gzb_codec_pbwt_compress (vb, c, &hdr, c->local.data, &ulen, NULL, NULL, &clen, SOFT_FAIL, "harness") || clen) return -2;
    *n_runs = (uint32_t)(c[1].local.len / 4); *n_fgrc = (uint32_t)(c[2].local.len / 4);
    memcpy (runs, c[1].local.data, c[1].local.len); memcpy (fgrc, c[2].local.data, c[2].local.len);
    /* PIZ: contexts 0 = matrix, 1 = RUNS (already uncompressed, host endian), 2 = FGRC: its section body arrives big endian */
    VBlockP pv = calloc (1, sizeof *pv); pv->vblock_i = 4;
    struct Context *p = &pv->ctx[0];
    p->HT_n_lines = n_lines;
    memcpy (buf_alloc_ (&p[1].local, c[1].local.len + 16), c[1].local.data, c[1].local.len); p[1].local.len = c[1].local.len;
    uint32_t *be = malloc (c[2].local.len + 16);
    for (uint32_t i = 0; i < *n_fgrc; i++) be[i] = __builtin_bswap32 (fgrc[i]);
    gzb_codec_pbwt_uncompress (pv, &p[2], GZB_CODEC_PBWT, 0, (char *)be, (uint32_t)c[2].local.len, NULL, 0, 0, "harness");
    if (p->local.len != len || p->ht_per_line != w) return -3;
    memcpy (back, p->local.data, len);
    /* reconstruct: a GT of ploidy 2 per sample with '|' between the haplotypes and a tab between samples, as the container would emit */
    buf_alloc_ (&pv->txt_out, len * 5 + 64);
    pv->big_allele = 12;
    uint64_t cells = 0; for (uint64_t i = 0; i < len; i++) cells += ht[i] != '*';
    for (uint64_t k = 0; k < cells; k++) {
        gzb_codec_pbwt_reconstruct (pv, GZB_CODEC_PBWT, p, 0, true);
        pv->txt_out.data[pv->txt_out.len++] = (k & 1) ? '\t' : '|';
    }
    *text_len = (uint32_t)pv->txt_out.len; memcpy (text, pv->txt_out.data, pv->txt_out.len);
    free (be); vb_free (vb); vb_free (pv);
    return 0;
}

/* LONGR: compress (lengths through the assigned sub-codec, values left in the next context), then one reconstruct call per read */
int harness_longr (const uint8_t *txt, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *seq_len, const uint32_t *qual_len, const uint8_t *is_rev,
                   uint32_t n_lines, const uint8_t *value_to_bin, int sub_codec, uint8_t *values, uint32_t *n_values, uint32_t *lens_be,
                   uint8_t *comp, uint32_t *comp_len, uint8_t *back, int *n_missing)
{
    harness_init ();
    if (setjmp (on_abort)) return -1;
    g_assign = (Codec)sub_codec;
    VBlockP vb = calloc (1, sizeof *vb); vb->vblock_i = 5; vb->n_lines = n_lines; g_vb = vb;
    vb->qual = malloc (n_lines * sizeof (char *)); vb->seq = malloc (n_lines * sizeof (char *));
    vb->qual_len = (uint32_t *)qual_len; vb->seq_len = (uint32_t *)seq_len; vb->is_rev = (uint8_t *)is_rev;
    uint64_t total_q = 0, total_s = 0;
    for (uint32_t i = 0; i < n_lines; i++) { vb->qual[i] = (char *)txt + qual_off[i]; vb->seq[i] = (char *)txt + seq_off[i]; total_q += qual_len[i]; total_s += seq_len[i]; }
    struct Context *c = &vb->ctx[0];
    memcpy (c[1].table, value_to_bin, 256);
    union SectionHeaderUnion hdr = { { 0 } };
    uint32_t ulen = (uint32_t)total_q; char *z = NULL; int sf = 0;
    int rc = comp_compress (gzb_codec_longr_compress, gzb_codec_longr_est_size, GZB_CODEC_LONGR, vb, c, &hdr, NULL, &ulen, cb_qual, 0, &z, comp_len, &sf);
    if (rc) return rc;
    memcpy (comp, z, *comp_len);
    *n_values = (uint32_t)c[1].local.len; memcpy (values, c[1].local.data, c[1].local.len);
    memcpy (lens_be, c[0].local.data, 65536 * 4);
    /* PIZ: lens_ctx.local = the uncompressed section (big endian), values_ctx.local = the values; SEQ of every read known up front */
    vb->recon_lens = seq_len; vb->seq_txt = (const char *)txt; vb->seq_off = seq_off;
    uint64_t txt_len = 0; for (uint32_t i = 0; i < n_lines; i++) { if (seq_off[i] + seq_len[i] > txt_len) txt_len = seq_off[i] + seq_len[i]; if (qual_off[i] + qual_len[i] > txt_len) txt_len = qual_off[i] + qual_len[i]; }
    vb->seq_txt_len = txt_len;
    buf_alloc_ (&vb->txt_out, total_s + 64);
    for (uint32_t i = 0; i < n_lines; i++)
        if (seq_len[i]) gzb_codec_longr_reconstruct (vb, GZB_CODEC_LONGR, c, seq_len[i], true);
    memcpy (back, vb->txt_out.data, vb->txt_out.len);
    *n_missing = vb->n_missing;
    int out_len = (int)vb->txt_out.len;
    free (z); g_vb = NULL; vb->qual_len = NULL; vb->seq_len = NULL; vb->is_rev = NULL; vb_free (vb);
    return out_len >= 0 ? 0 : -6;
}

/* several compute threads submitting one section each through the combiner */
#include <pthread.h>
struct job { gzb_combiner *c; gzb_section s; int rc; };
static void *job_main (void *p) { struct job *j = p; uint64_t t; j->rc = gzb_submit (j->c, &j->s, &t); if (!j->rc) j->rc = gzb_wait (j->c, t, &j->s); return NULL; }
int harness_combine (int codec, const uint8_t *data, const uint32_t *lens, uint32_t n_secs, uint32_t linger_us, uint8_t *out, const uint32_t *out_off, uint32_t *out_lens, uint64_t *n_batches)
{
    gzb_combiner *c = gzb_combiner_create (0, 1, linger_us);
    if (!c) return -1;
    struct job *jobs = calloc (n_secs, sizeof *jobs); pthread_t *th = malloc (n_secs * sizeof *th);
    uint64_t o = 0;
    for (uint32_t i = 0; i < n_secs; i++) {
        jobs[i].c = c; jobs[i].s.codec = codec; jobs[i].s.in = data + o; jobs[i].s.in_len = lens[i]; jobs[i].s.out = out + out_off[i];
        jobs[i].s.out_cap = gzb_est_size (codec, lens[i]); o += lens[i];
        pthread_create (&th[i], NULL, job_main, &jobs[i]);
    }
    int rc = 0;
    for (uint32_t i = 0; i < n_secs; i++) { pthread_join (th[i], NULL); if (jobs[i].rc || jobs[i].s.status) rc = -2; out_lens[i] = jobs[i].s.out_len; }
    *n_batches = gzb_combiner_batches (c);
    gzb_combiner_destroy (c); free (jobs); free (th);
    return rc;
}
