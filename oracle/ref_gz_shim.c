// ref_gz_shim.c — TEST INFRASTRUCTURE.  Hosts the REFERENCE's own codec_domq.c (compiled unmodified from /root/reference/src,
// see oracle/Makefile) outside the genozip program: this file is compiled against the reference's headers (so VBlock, Context,
// Buffer, CodecArgs ... have the reference's exact layouts), hand-constructs the minimal VBlock a compute thread would hand
// to the codec, supplies the ~20 host symbols the object needs (buffer allocation, seg_by_ctx, codec table ...) and exports
// flat entry points:
//
//     ref_domq_encode ()   =  codec_domq_comp_init (force)  +  codec_domq_compress ()          (codec_domq.c:299-323, 379-521)
//     ref_acgt_pack ()     =  codec_acgt_compress () up to (and capturing) its sub-codec call  (codec_acgt.c:64-177)
//     ref_longr_encode ()  =  codec_longr_segconf_calculate_bins () + codec_longr_compress ()  (codec_longr.c:66-264)
//     ref_pbwt_encode ()   =  codec_pbwt_compress ()                                           (codec_pbwt.c:244-287)
//     ref_acgt_unpack ()   =  codec_acgt_uncompress () [+ codec_xcgt_uncompress ()]            (codec_acgt.c:185-248)
//     ref_pbwt_decode ()   =  codec_pbwt_uncompress ()                                         (codec_pbwt.c:372-402)
//     ref_domq_decode ()   =  codec_domq_reconstruct () line by line                           (codec_domq.c:774-809)
//     ref_longr_decode ()  =  codec_longr_reconstruct () read by read                          (codec_longr.c:342-373)
//
// The nucleotide tables _acgt_encode / _acgt_encode_comp are the reference's own (.rodata of its compiled reference.c,
// extracted with objcopy: oracle/Makefile).
// so that the CPU restatement (oracle/gz_port.c) — and through it the CUDA path — is pinned against the reference's compiled
// code rather than against a reading of it.  Only built where /root/reference exists; nothing here is reference source.
#include "genozip.h"
#include "vblock.h"
#include "context.h"
#include "buffer.h"
#include "codec.h"
#include "file.h"
#include "segconf.h"
#include "flags.h"
#include "seg.h"
#include "base64.h"
#include "reconstruct.h"
#include "profiler.h"
#include "sam.h"
#include "sam_private.h"
#include "fastq.h"
#include "vcf.h"
#include "sections.h"
#include "b250.h"
#include <stdarg.h>
#include <setjmp.h>

// ---------------------------------------------------------------- globals the objects reference
Flags flag;
SegConf segconf;
static File the_z_file;
static File the_txt_file;
FileP z_file = &the_z_file, txt_file = NULL;   // (txt_file points at the_txt_file only inside the PACB harness)
FILE *info_stream;
VBlockP evb = NULL;

static jmp_buf on_abort;
static char abort_msg[1024];

noreturn void error_assert_failed (FUNCLINE, rom format, ...)
{
    va_list ap; va_start (ap, format);
    int k = snprintf (abort_msg, sizeof abort_msg, "%s:%u: ", func, code_line);
    vsnprintf (abort_msg + k, sizeof abort_msg - k, format, ap);
    va_end (ap);
    longjmp (on_abort, 1);
}
noreturn void error_asspiz (VBlockP vb, FUNCLINE, rom format, ...)
{
    snprintf (abort_msg, sizeof abort_msg, "%s:%u: ASSPIZ %s", func, code_line, format);
    longjmp (on_abort, 1);
}
const char *ref_gz_last_error (void) { return abort_msg; }

StrText vb_name (VBlockP vb) { StrText t = {}; strcpy (t.s, "VB(shim)"); return t; }
rom codec_name (Codec codec) { return "codec"; }
void show_time_one (VBlockP vb, rom res, uint64_t delta) {}
void codec_show_time (VBlockP vb, rom name, rom subname, Codec codec) {}
bool str_is_zero (STRp(str)) { for (uint32_t i = 0; i < str_len; i++) if (str[i]) return false; return true; }

// ---------------------------------------------------------------- Buffer: malloc-backed, same visible semantics as buffer.c
// (data = memory + 8; buf_free keeps the memory and clears data/len/param; buf_alloc preserves the content)
BufDescType buf_desc (ConstBufferP buf) { BufDescType d = {}; snprintf (d.s, sizeof d.s, "%s size=%"PRIu64" len=%"PRIu64, buf->name ? buf->name : "?", (uint64_t)buf->size, (uint64_t)buf->len); return d; }

void buf_alloc_do (VBlockP vb, BufferP buf, uint64_t requested_size, float grow_at_least_factor, rom name, FUNCLINE)
{
    if (buf->memory && requested_size <= buf->size) { if (!buf->data) buf->data = buf->memory + sizeof (uint64_t); return; }
    uint64_t new_size = (uint64_t)(requested_size * (grow_at_least_factor > 1 ? grow_at_least_factor : 1)) + 64;
    char *mem = calloc (1, new_size + 32);
    if (buf->data && buf->size) memcpy (mem + sizeof (uint64_t), buf->data, buf->size);
    if (buf->memory && !buf->shared) free (buf->memory);
    buf->memory = mem; buf->data = mem + sizeof (uint64_t); buf->size = new_size;
    buf->type = BUF_REGULAR; buf->vb = vb; buf->shared = 0;
    if (name) buf->name = name;
    buf->func = func; buf->code_line = code_line;
}
void buf_free_do (BufferP buf, FUNCLINE)
{
    buf->data = NULL; buf->len = 0; buf->param = 0;
    if (buf->shared) { buf->memory = NULL; buf->size = 0; buf->shared = 0; buf->type = BUF_UNALLOCATED; }   // an overlay lets go of the memory it borrowed
}
void buf_destroy_do (BufferP buf, FUNCLINE)
{
    if (buf->memory && !buf->shared) free (buf->memory);
    memset (buf, 0, sizeof *buf);
}
void buf_copy_do (VBlockP dst_vb, BufferP dst, ConstBufferP src, uint64_t bytes_per_entry, uint64_t src_start_entry, uint64_t max_entries, FUNCLINE, rom dst_name)
{
    uint64_t n = src->len - src_start_entry;
    if (max_entries && max_entries < n) n = max_entries;
    uint64_t w = bytes_per_entry ? bytes_per_entry : 1;
    buf_alloc_do (dst_vb, dst, n * w, 1, dst_name, func, code_line);
    if (n) memcpy (dst->data, src->data + src_start_entry * w, n * w);
    dst->len = n;
}

void buf_set_shared (BufferP buf) { buf->shared = 1; }                     // (the shim never frees shared memory: the harness is short-lived)
void buf_overlay_do (VBlockP vb, BufferP top_buf, BufferP bottom_buf, uint64_t start_in_bottom, bool copy_len, FUNCLINE, rom name)
{
    top_buf->memory = bottom_buf->memory; top_buf->data = bottom_buf->data + start_in_bottom; top_buf->size = bottom_buf->size - start_in_bottom;
    top_buf->type = BUF_REGULAR; top_buf->shared = 1; top_buf->vb = vb; top_buf->name = name;
    if (copy_len) top_buf->len = bottom_buf->len;
}

// codec scratch (codec.c:30-82): plain heap blocks, all released by codec_free_all
static void *codec_blocks[64]; static int n_codec_blocks;
void *codec_alloc_do (VBlockP vb, uint64_t size, float grow_at_least_factor, unsigned *buf_i, FUNCLINE)
{
    ASSERT0 (n_codec_blocks < 64, "shim: too many codec_alloc blocks");
    return codec_blocks[n_codec_blocks++] = malloc (size + 64);
}
void codec_free_do (void *vb, void *addr, FUNCLINE) {}
void codec_free_all (VBlockP vb) { while (n_codec_blocks) free (codec_blocks[--n_codec_blocks]); }
uint32_t codec_complex_est_size (Codec codec, uint64_t uncompressed_len) { return (uint32_t)uncompressed_len + 1024; }
void BGEN_u32_buf (BufferP buf, LocalType *lt) { uint32_t *w = (uint32_t *)buf->data; for (uint64_t i = 0; i < buf->len; i++) w[i] = __builtin_bswap32 (w[i]); }
void ctx_consolidate_stats (VBlockP vb, int parent, ...) {}
StrText line_name (VBlockP vb) { StrText t = {}; strcpy (t.s, "line(shim)"); return t; }
static rom cur_seq; static bool cur_is_rev;                                // the read being reconstructed (LONGR decoder harness)
bool sam_is_last_flags_rev_comp (VBlockP vb) { return cur_is_rev; }
rom sam_piz_get_textual_seq (VBlockP vb) { return cur_seq; }
int64_t reconstruct_from_local_int (VBlockP vb, ContextP ctx, char separator, ReconType reconstruct) { ABORT0 ("shim: reconstruct_from_local_int"); }
uint32_t str_int_ex (int64_t n, char *str, bool add_nul_terminator) { int k = sprintf (str, "%"PRId64, n); return (uint32_t)k; }

// ---------------------------------------------------------------- segmenter / context services used by codec_domq.c
static uint8_t denorm_snip[NUM_CODECS * 0 + 95 * 95 * 2];
static uint32_t denorm_snip_len;
WordIndex seg_by_ctx_ex (VBlockP vb, STRp(snip), ContextP ctx, uint32_t add_bytes, bool *restrict is_new)
{
    denorm_snip_len = snip_len < sizeof denorm_snip ? snip_len : sizeof denorm_snip;
    memcpy (denorm_snip, snip, denorm_snip_len);                           // the de-normalisation table segged into DOMQRUNS (codec_domq.c:241-244)
    return 0;
}
unsigned base64_encode (STR8p(in), char *restrict b64_str) { memcpy (b64_str, in, in_len); return in_len; }   // identity: the table is captured raw
uint32_t base64_decode (STRp(b64_str), STR8c(out)) { memcpy (out, b64_str, b64_str_len); return b64_str_len; }
void ctx_set_ltype (VBlockP vb, int ltype, ...) {}
WordIndex ctx_peek_next_snip (VBlockP vb, ContextP ctx, pSTRp (snip)) { *snip = (rom)denorm_snip; *snip_len = denorm_snip_len; return 0; }   // DOMQRUNS' single snip: the de-normalisation table
// PIZ / consensus-read paths of codec_domq.c that the encoder harness never takes
static bool shim_missing_ok;                                                 // the NORMQ harness reconstructs lines without quality
void sam_reconstruct_missing_quality (VBlockP vb, ReconType reconstruct)
{
    if (!shim_missing_ok) ABORT0 ("shim: sam_reconstruct_missing_quality");
    if (reconstruct) vb->txt_data.data[vb->txt_data.len++] = '*';             // RECONSTRUCT1 ('*') (sam_qual.c:534-535)
}
void sam_xcons_reconstruct_QUAL (VBlockP vb, ContextP ctx, uint32_t qual_len, bool reconstruct) { ABORT0 ("shim: sam_xcons_reconstruct_QUAL"); }
void sam_xcons_split_qual_line (VBlockP vb, BufferP ql_buf) { ABORT0 ("shim: sam_xcons_split_qual_line"); }

// ---------------------------------------------------------------- the codec table: every sub-codec is "store"
static uint8_t *cap_out; static uint32_t cap_len;
static COMPRESS (shim_store)
{
    if (!uncompressed && get_line_cb) {                                    // a sub-codec reads its data line by line (codec_homp.c:202, codec_t0.c:121)
        uint32_t at = 0;
        for (uint32_t li = 0; li < vb->lines.len32; li++) { char *d; uint32_t n; get_line_cb (vb, ctx, li, &d, &n, *uncompressed_len - at, NULL); memcpy (compressed + at, d, n); at += n; }
        *compressed_len = at; cap_out = (uint8_t *)compressed; cap_len = at;
        return true;
    }
    memcpy (compressed, uncompressed, *uncompressed_len);
    *compressed_len = *uncompressed_len;
    cap_out = (uint8_t *)compressed; cap_len = *uncompressed_len;
    return true;
}
static UNCOMPRESS (shim_unstore) { memcpy (uncompressed_buf->data, compressed, compressed_len); }
static uint32_t shim_est_size (Codec codec, uint64_t uncompressed_len) { return (uint32_t)uncompressed_len + 64; }
CodecArgs codec_args[NUM_CODECS];
Codec codec_assign_best_codec (VBlockP vb, ContextP ctx, BufferP non_ctx_data, SectionType st) { return CODEC_NONE; }

static void shim_init (void)
{
    info_stream = stderr;
    memset (&flag, 0, sizeof flag);
    flag.command = ZIP;                                                     // IS_ZIP (genozip.h:405)
    flag.show_time_comp_i = COMP_NONE;                                      // profiler off (profiler.h:119)
    memset (&segconf, 0, sizeof segconf);
    memset (&the_z_file, 0, sizeof the_z_file);
    for (int c = 0; c < NUM_CODECS; c++) { codec_args[c].compress = shim_store; codec_args[c].uncompress = shim_unstore; codec_args[c].est_size = shim_est_size; }
}

// ---------------------------------------------------------------- the lines of the hand-made VBlock
static uint8_t *g_txt; static const uint64_t *g_off; static const uint32_t *g_len;
static COMPRESSOR_CALLBACK (shim_get_line)
{
    *line_data = (char *)g_txt + g_off[vb_line_i];
    *line_data_len = g_len[vb_line_i];
    if (is_rev) *is_rev = 0;
}

// codec_domq_comp_init (force) + codec_domq_compress on n_lines quality strings.  Returns 0, or -1 after an ABORT of the reference code
int ref_domq_encode (const uint8_t *txt, uint64_t txt_len, const uint64_t *line_off, const uint32_t *line_len, uint32_t n_lines,
                     uint8_t *qual, uint32_t *qual_len, uint8_t *runs, uint32_t *runs_len, uint8_t *mplx, uint32_t *mplx_len,
                     uint8_t *divr, uint32_t *divr_len, uint8_t *denorm, uint32_t *denorm_len, uint8_t *param, uint8_t *has_diverse)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1;
    vb->lines.len = n_lines;
    g_txt = malloc (txt_len + 1); memcpy (g_txt, txt, txt_len); g_off = line_off; g_len = line_len;   // the codec normalises the lines in place
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += line_len[i];
    ContextP qual_ctx = CTX (SAM_QUAL);
    for (int k = 0; k < 4; k++) { qual_ctx[k].did_i = SAM_QUAL + k; strcpy (qual_ctx[k].tag_name, k == 0 ? "QUAL" : k == 1 ? "DOMQRUNS" : k == 2 ? "QUALMPLX" : "DIVRQUAL"); }
    qual_ctx->local.len = total;                                            // callback-mode locals carry only their total length

    if (!codec_domq_comp_init (vb, SAM_QUAL, shim_get_line, true)) return -2;

    SectionHeaderCtx header = {};
    uint32_t ulen = (uint32_t)total, clen = 2 * (uint32_t)total + 1024;
    char *comp = malloc (clen);
    cap_out = NULL; cap_len = 0;
    if (!codec_domq_compress (vb, qual_ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_line, comp, &clen, true, "QUAL")) return -3;

    *param = qual_ctx->local.prm8[0]; *has_diverse = qual_ctx->domq_has_diverse;
    memcpy (qual, cap_out, cap_len); *qual_len = cap_len;
    memcpy (runs, qual_ctx[1].local.data, qual_ctx[1].local.len); *runs_len = qual_ctx[1].local.len32;
    memcpy (mplx, qual_ctx[2].local.data, qual_ctx[2].local.len); *mplx_len = qual_ctx[2].local.len32;
    memcpy (divr, qual_ctx[3].local.data, qual_ctx[3].local.len); *divr_len = qual_ctx[3].local.len32;
    memcpy (denorm, denorm_snip, denorm_snip_len); *denorm_len = denorm_snip_len;
    free (comp); free (g_txt);
    for (int k = 0; k < 4; k++) { buf_destroy_do (&qual_ctx[k].local, __FUNCLINE); }
    free (vb);
    return 0;
}

// ================================================================ ACGT
// NONREF.local = the bases; returns the 2-bit words handed to the sub-codec, the exception stream (NONREF_X.local) and acgt_no_x
int ref_acgt_pack (const uint8_t *seq, uint64_t n, uint8_t *packed, uint64_t *packed_len, uint8_t *x, int *no_x)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1;
    ContextP nonref = CTX (FASTQ_NONREF);
    nonref[0].did_i = FASTQ_NONREF; nonref[1].did_i = FASTQ_NONREF + 1;
    strcpy (nonref[0].tag_name, "NONREF"); strcpy (nonref[1].tag_name, "NONREF_X");
    buf_alloc_do (vb, &nonref->local, n + 8, 1, "local", __FUNCLINE);
    memcpy (nonref->local.data, seq, n); nonref->local.len = n;
    SectionHeaderCtx header = {};
    uint32_t ulen = (uint32_t)n, clen = (uint32_t)(n / 4 + 1024);
    char *comp = malloc (clen);
    cap_out = NULL; cap_len = 0;
    if (!codec_acgt_compress (vb, nonref, (SectionHeaderP)&header, nonref->local.data, &ulen, NULL, comp, &clen, true, "NONREF")) return -3;
    memcpy (packed, cap_out, cap_len); *packed_len = cap_len;
    *no_x = header.flags.ctx.acgt_no_x;
    if (!*no_x) memcpy (x, nonref[1].local.data, n); else memset (x, 0, n);
    free (comp); free (vb);
    return 0;
}

// ================================================================ LONGR
static const uint64_t *g_seq_off; static const uint8_t *g_is_rev; static const uint32_t *g_seq_len;
static COMPRESSOR_CALLBACK (shim_get_qual_rev)
{
    *line_data = (char *)g_txt + g_off[vb_line_i];
    *line_data_len = g_len[vb_line_i];
    if (is_rev) *is_rev = g_is_rev ? g_is_rev[vb_line_i] : 0;
}
static void shim_get_seq (VBlockP vb, LineIType vb_line_i, char **line_data, uint32_t *line_data_len, bool *is_rev)
{
    *line_data = (char *)g_txt + g_seq_off[vb_line_i];
    *line_data_len = (g_seq_len ? g_seq_len : g_len)[vb_line_i];
    if (is_rev) *is_rev = g_is_rev ? g_is_rev[vb_line_i] : 0;
}
COMPRESSOR_CALLBACK (fastq_zip_seq) { shim_get_seq (vb, vb_line_i, line_data, line_data_len, is_rev); }
COMPRESSOR_CALLBACK (sam_zip_seq)   { shim_get_seq (vb, vb_line_i, line_data, line_data_len, is_rev); }

// txt holds the SEQ and QUAL strings; is_rev NULL = FASTQ (never reverse-complemented), else SAM-like per-line flags
// seq_len (NULL = len): sequence lengths where they differ from the quality lengths — a SAM line without quality is ' ' (:188-192)
int ref_longr_encode2 (const uint8_t *txt, uint64_t txt_len, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *len, const uint32_t *seq_len,
                       const uint8_t *is_rev, uint32_t n_lines, uint8_t *value_to_bin, uint8_t *values, uint32_t *lens_be);
int ref_longr_encode (const uint8_t *txt, uint64_t txt_len, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *len, const uint8_t *is_rev,
                      uint32_t n_lines, uint8_t *value_to_bin, uint8_t *values, uint32_t *lens_be)
{
    return ref_longr_encode2 (txt, txt_len, seq_off, qual_off, len, NULL, is_rev, n_lines, value_to_bin, values, lens_be);
}
int ref_longr_encode2 (const uint8_t *txt, uint64_t txt_len, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *len, const uint32_t *seq_len,
                       const uint8_t *is_rev, uint32_t n_lines, uint8_t *value_to_bin, uint8_t *values, uint32_t *lens_be)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    evb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines;
    vb->data_type = is_rev ? DT_SAM : DT_FASTQ;
    g_txt = malloc (txt_len + 1); memcpy (g_txt, txt, txt_len); g_off = qual_off; g_seq_off = seq_off; g_len = len; g_is_rev = is_rev; g_seq_len = seq_len;
    vb->txt_data.len = txt_len;                                            // Ltxt: the callbacks' size limit
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += len[i];
    ContextP qual_ctx = CTX (SAM_QUAL);
    for (int k = 0; k < 2; k++) { qual_ctx[k].did_i = SAM_QUAL + k; strcpy (qual_ctx[k].tag_name, k ? "DOMQRUNS" : "QUAL"); }

    segconf.running = true;                                                 // main thread, once per file (codec_longr.c:66-136)
    codec_longr_segconf_calculate_bins (vb, qual_ctx + 1, shim_get_qual_rev);
    segconf.running = false;
    memcpy (value_to_bin, ZCTX (SAM_QUAL + 1)->value_to_bin.data, 256);

    SectionHeaderCtx header = {};
    uint32_t ulen = (uint32_t)total, clen = 65536 * 4 + 4096;
    char *comp = malloc (clen);
    cap_out = NULL; cap_len = 0;
    if (!codec_longr_compress (vb, qual_ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_qual_rev, comp, &clen, true, "QUAL")) return -3;
    if (cap_len != 65536 * 4) return -4;
    memcpy (lens_be, cap_out, cap_len);
    memcpy (values, qual_ctx[1].local.data, total);
    free (comp); free (g_txt); free (vb); free (evb); evb = NULL;
    return 0;
}

// ================================================================ PBWT
int ref_pbwt_encode (const uint8_t *ht, uint32_t n_lines, uint32_t ht_per_line, uint32_t *runs, uint32_t *n_runs, uint32_t *fgrc, uint32_t *n_fgrc)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1;
    ContextP ht_ctx = CTX (FORMAT_GT_HT), runs_ctx = CTX (FORMAT_PBWT_RUNS), fgrc_ctx = CTX (FORMAT_PBWT_FGRC);
    ht_ctx->did_i = FORMAT_GT_HT; runs_ctx->did_i = FORMAT_PBWT_RUNS; fgrc_ctx->did_i = FORMAT_PBWT_FGRC;
    uint64_t n = (uint64_t)n_lines * ht_per_line;
    buf_alloc_do (vb, &ht_ctx->local, n + 8, 1, "local", __FUNCLINE);
    memcpy (ht_ctx->local.data, ht, n); ht_ctx->local.len = n;
    ht_ctx->ht_per_line = ht_per_line; ht_ctx->HT_n_lines = n_lines;
    SectionHeaderCtx header = {};
    uint32_t ulen = (uint32_t)n, clen = 1024;
    char comp[1024];
    if (!codec_pbwt_compress (vb, ht_ctx, (SectionHeaderP)&header, NULL, &ulen, NULL, comp, &clen, true, "HT")) return -3;
    if (clen != 0) return -4;                                               // the matrix itself produces no section (:280-282)
    *n_runs = runs_ctx->local.len32; memcpy (runs, runs_ctx->local.data, 4ull * *n_runs);
    *n_fgrc = fgrc_ctx->local.len32; memcpy (fgrc, fgrc_ctx->local.data, 4ull * *n_fgrc);
    free (vb);
    return 0;
}

// ================================================================ PIZ side
static void shim_init_piz (void)
{
    shim_init ();
    flag.command = PIZ;
    the_z_file.genozip_ver = (Version){ 15, 86 };                           // VER(14), VER2(15,76): the current file format
}

int ref_acgt_unpack (const uint8_t *packed, uint64_t packed_len, const uint8_t *x /* NULL = acgt_no_x */, uint64_t n, uint8_t *seq)
{
    shim_init_piz ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1;
    ContextP nonref = CTX (FASTQ_NONREF);
    nonref[0].did_i = FASTQ_NONREF; nonref[1].did_i = FASTQ_NONREF + 1;
    nonref->flags.acgt_no_x = (x == NULL);
    buf_alloc_do (vb, &nonref->local, n + 8, 1, "local", __FUNCLINE); nonref->local.len = n;        // allocated by the caller (zfile.c:229)
    codec_acgt_uncompress (vb, nonref, CODEC_ACGT, 0, (rom)packed, (uint32_t)packed_len, &nonref->local, n, CODEC_NONE, "NONREF");
    if (x) {
        buf_alloc_do (vb, &nonref[1].local, n + 8, 1, "local", __FUNCLINE); nonref[1].local.len = n;
        codec_xcgt_uncompress (vb, nonref + 1, CODEC_XCGT, 0, (rom)x, (uint32_t)n, &nonref[1].local, n, CODEC_NONE, "NONREF_X");
    }
    memcpy (seq, nonref->local.data, n);
    free (vb);
    return 0;
}

// runs: host-endian uint32 (as after piz_adjust_one_local); fgrc_be: the FGRC section as stored (big-endian words, :380-381)
int ref_pbwt_decode (const uint32_t *runs, uint32_t n_runs, const uint32_t *fgrc_be, uint32_t n_fgrc, uint32_t n_lines, uint8_t *ht, uint64_t *ht_len)
{
    shim_init_piz ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock) + (1 << 20));                   // (codec_pbwt.c casts to the larger VCF VBlock)
    vb->vblock_i = 1; vb->lines.len = n_lines;
    ContextP ht_ctx = CTX (FORMAT_GT_HT), runs_ctx = CTX (FORMAT_PBWT_RUNS), fgrc_ctx = CTX (FORMAT_PBWT_FGRC);
    ht_ctx->did_i = FORMAT_GT_HT; runs_ctx->did_i = FORMAT_PBWT_RUNS; fgrc_ctx->did_i = FORMAT_PBWT_FGRC;
    ht_ctx->HT_n_lines = n_lines;
    buf_alloc_do (vb, &runs_ctx->local, 4ull * n_runs + 8, 1, "local", __FUNCLINE);
    memcpy (runs_ctx->local.data, runs, 4ull * n_runs); runs_ctx->local.len = n_runs;
    buf_alloc_do (vb, &vb->scratch, 4ull * n_fgrc + 8, 1, "scratch", __FUNCLINE);                     // a sub-codec's input lives in vb->scratch
    memcpy (vb->scratch.data, fgrc_be, 4ull * n_fgrc); vb->scratch.len = n_fgrc;
    codec_pbwt_uncompress (vb, fgrc_ctx, CODEC_PBWT, 0, vb->scratch.data, 4 * n_fgrc, &fgrc_ctx->local, 0, CODEC_NONE, "FGRC");
    *ht_len = ht_ctx->local.len;
    memcpy (ht, ht_ctx->local.data, ht_ctx->local.len);
    free (vb);
    return 0;
}

// the four DOMQ streams + table + parameter of one VBlock -> the quality strings of lines of the given lengths
int ref_domq_decode (const uint8_t *qual, uint32_t qual_len, const uint8_t *runs, uint32_t runs_len, const uint8_t *mplx, uint32_t mplx_len,
                     const uint8_t *divr, uint32_t divr_len, const uint8_t *denorm, uint32_t denorm_len, uint8_t param,
                     const uint32_t *line_len, uint32_t n_lines, uint8_t *out)
{
    shim_init_piz ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += line_len[i];
    ContextP c = CTX (SAM_QUAL);
    const uint8_t *src[4] = { qual, runs, mplx, divr }; uint32_t len[4] = { qual_len, runs_len, mplx_len, divr_len };
    for (int k = 0; k < 4; k++) {
        c[k].did_i = SAM_QUAL + k; c[k].is_loaded = true;
        buf_alloc_do (vb, &c[k].local, len[k] + 8, 1, "local", __FUNCLINE);
        memcpy (c[k].local.data, src[k], len[k]); c[k].local.len = len[k];
    }
    c[0].local.prm8[0] = param;
    c[0].dict_id.num = 0;
    memcpy (denorm_snip, denorm, denorm_len); denorm_snip_len = denorm_len;
    buf_alloc_do (vb, &vb->txt_data, total + 64, 1, "txt_data", __FUNCLINE);
    for (uint32_t i = 0; i < n_lines; i++)
        if (line_len[i]) codec_domq_reconstruct (vb, CODEC_DOMQ, c, line_len[i], true);            // (empty QUAL lines are not routed to the codec)
    if (vb->txt_data.len != total) return -5;
    memcpy (out, vb->txt_data.data, total);
    free (vb);
    return 0;
}

// values + big-endian channel lengths + value_to_bin of one VBlock -> the quality strings; is_rev NULL = FASTQ-like (never reversed)
int ref_longr_decode (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev, uint32_t n_lines,
                      const uint8_t *value_to_bin, const uint8_t *values, const uint32_t *lens_be, uint8_t *out)
{
    shim_init_piz ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;      // SEQ and the strand come through the sam_* accessors above
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += len[i];
    ContextP c = CTX (SAM_QUAL);
    c[0].did_i = SAM_QUAL; c[1].did_i = SAM_QUAL + 1;
    buf_alloc_do (vb, &c[0].local, 65536 * 4 + 8, 1, "local", __FUNCLINE);
    memcpy (c[0].local.data, lens_be, 65536 * 4); c[0].local.len = 65536 * 4;
    buf_alloc_do (vb, &c[1].local, total + 8, 1, "local", __FUNCLINE);
    memcpy (c[1].local.data, values, total); c[1].local.len = total;
    ContextP zv = ZCTX (SAM_QUAL + 1);                                      // SEC_COUNTS of the values context: the value-to-bin map as 256 x uint64
    buf_alloc_do (NULL, &zv->counts, 256 * 8, 1, "counts", __FUNCLINE);
    for (int i = 0; i < 256; i++) ((uint64_t *)zv->counts.data)[i] = value_to_bin[i];
    zv->counts.len = 256;
    buf_alloc_do (vb, &vb->txt_data, total + 64, 1, "txt_data", __FUNCLINE);
    for (uint32_t i = 0; i < n_lines; i++) {
        if (!len[i]) continue;
        cur_seq = (rom)txt + seq_off[i]; cur_is_rev = is_rev ? is_rev[i] : false;
        vb->seq_len = len[i];
        codec_longr_reconstruct (vb, CODEC_LONGR, c, len[i], true);
    }
    if (vb->txt_data.len != total) return -5;
    memcpy (out, vb->txt_data.data, total);
    free (vb);
    return 0;
}

// ================================================================ NORMQ (the reference's compiled codec_normq.c)
static const uint8_t *g_rev;
static COMPRESSOR_CALLBACK (shim_get_line_rev)
{
    *line_data = (char *)g_txt + g_off[vb_line_i];
    *line_data_len = g_len[vb_line_i];
    if (is_rev) *is_rev = g_rev ? g_rev[vb_line_i] : 0;
}

// codec_normq_compress on n_lines quality strings: returns what it hands to its sub-codec (QUAL.local after the gather)
int ref_normq_encode (const uint8_t *txt, uint64_t txt_len, const uint64_t *line_off, const uint32_t *line_len, const uint8_t *is_rev, uint32_t n_lines,
                      uint8_t *local, uint64_t *local_len)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines;
    g_txt = (uint8_t *)txt; g_off = line_off; g_len = line_len; g_rev = is_rev;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += line_len[i];
    ContextP ctx = CTX (SAM_QUAL);
    ctx->did_i = SAM_QUAL; strcpy (ctx->tag_name, "QUAL");
    ctx->local.len = total;                                                 // callback-mode locals carry only their total length
    SectionHeaderCtx header = {};
    uint32_t ulen = (uint32_t)total, clen = (uint32_t)total + 1024;
    char *comp = malloc (clen);
    cap_out = NULL; cap_len = 0;
    if (!codec_normq_compress (vb, ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_line_rev, comp, &clen, true, "QUAL")) return -3;
    memcpy (local, cap_out, cap_len); *local_len = cap_len;
    free (comp); buf_destroy_do (&ctx->local, __FUNCLINE); free (vb);
    return 0;
}

// codec_normq_reconstruct line by line: out = what lands in txt_data (a line without quality contributes the one character '*')
int ref_normq_decode (const uint8_t *local, uint64_t local_len, const uint32_t *len, const uint8_t *is_rev, uint32_t n_lines, uint8_t *out, uint64_t *out_len)
{
    shim_init_piz ();
    if (setjmp (on_abort)) { shim_missing_ok = false; return -1; }
    shim_missing_ok = true;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += len[i];
    ContextP c = CTX (SAM_QUAL);
    c->did_i = SAM_QUAL; c->is_loaded = true;
    buf_alloc_do (vb, &c->local, local_len + 8, 1, "local", __FUNCLINE);
    memcpy (c->local.data, local, local_len); c->local.len = local_len;
    buf_alloc_do (vb, &vb->txt_data, total + 64, 1, "txt_data", __FUNCLINE);
    for (uint32_t i = 0; i < n_lines; i++) {
        if (!len[i]) continue;
        CTX (SAM_FLAG)->last_value.i = (is_rev && is_rev[i]) ? 0x10 : 0;      // last_flags.rev_comp (sam_private.h:531)
        codec_normq_reconstruct (vb, CODEC_NORMQ, c, len[i], true);
    }
    shim_missing_ok = false;
    *out_len = vb->txt_data.len;
    memcpy (out, vb->txt_data.data, vb->txt_data.len);
    free (vb);
    return 0;
}

// ================================================================ OQ (the reference's compiled codec_oq.c)
// context services codec_oq.c reaches for: its 94 channel contexts are looked up by dict_id (ctx_get_ctx / ECTX, context.h:174-207);
// the hand-made VBlock has no dict_id map, so both land here: a linear search over the contexts the harness created
#define SHIM_FIRST_DYN_DID 600
ContextP ctx_get_unmapped_ctx (ContextArrayP ca, DataType dt, DictId dict_id, STRp(tag_name))
{
    if (ca->num_contexts < SHIM_FIRST_DYN_DID) ca->num_contexts = SHIM_FIRST_DYN_DID;
    for (Did d = SHIM_FIRST_DYN_DID; d < ca->num_contexts; d++) if (ca->contexts[d].dict_id.num == dict_id.num) return &ca->contexts[d];
    ContextP c = &ca->contexts[ca->num_contexts];
    c->did_i = ca->num_contexts++; c->dict_id = dict_id;
    return c;
}
Did ctx_get_unmapped_existing_did_i (ConstContextArrayP ca, DictId dict_id)
{
    for (Did d = SHIM_FIRST_DYN_DID; d < ca->num_contexts; d++) if (ca->contexts[d].dict_id.num == dict_id.num) return d;
    return DID_NONE;
}
void ctx_update_zctx_txt_len (VBlockP vb, ContextP vctx, int64_t increment) {}
void ctx_consolidate_statsA (VBlockP vb, Did parent, ContextP ctxs[], unsigned num_deps) {}
DisplayPrintId dis_dict_id (DictId dict_id) { DisplayPrintId d = {}; memcpy (d.s, dict_id.id, 8); return d; }
#define SHIM_OQ_DICT_ID_Q(q) DICT_ID_MAKE2_5(((char[]){'O', (q)+33, 'Q',':','Z'}))       /* codec_oq.c:15 */

// codec_oq_compress on the lines of a hand-made SAM VBlock: channels back to back in channel order (count[q] bytes each), the monochars it hands to RANB
int ref_oq_encode (const uint8_t *txt, uint64_t txt_len, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *oq_off, const uint32_t *seq_len,
                   uint32_t n_lines, uint8_t *channels, uint32_t *count, uint8_t *monochars)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->data_type = DT_SAM;
    buf_alloc_do (vb, &vb->txt_data, txt_len + 64, 1, "txt_data", __FUNCLINE);
    memcpy (vb->txt_data.data, txt, txt_len); vb->txt_data.len = txt_len;
    buf_alloc_do (vb, &vb->lines, (uint64_t)(n_lines + 1) * sizeof (ZipDataLineSAM), 1, "lines", __FUNCLINE);
    memset (vb->lines.data, 0, (uint64_t)(n_lines + 1) * sizeof (ZipDataLineSAM)); vb->lines.len = n_lines;
    for (uint32_t i = 0; i < n_lines; i++) {
        ZipDataLineSAM *dl = B(ZipDataLineSAM, vb->lines, i);
        dl->QUAL = (TxtWord){ .index = (uint32_t)qual_off[i], .len = qual_len[i] };
        dl->OQ = (uint32_t)oq_off[i]; dl->SEQ.len = seq_len[i];
    }
    ContextP ctx = CTX (OPTION_OQ_Z);
    ctx->did_i = OPTION_OQ_Z; strcpy (ctx->tag_name, "OQ:Z");
    SectionHeaderCtx header = {};
    uint32_t ulen = 0, clen = 4096;
    char *comp = calloc (1, clen);
    cap_out = NULL; cap_len = 0;
    if (!codec_oq_compress (vb, ctx, (SectionHeaderP)&header, NULL, &ulen, NULL, comp, &clen, true, "OQ:Z")) return -3;
    if (cap_len != 94 || header.sub_codec != CODEC_RANB) return -4;
    memcpy (monochars, cap_out, 94);
    uint64_t at = 0;
    for (int q = 0; q < 94; q++) {
        ContextP c = ctx_get_existing_ctx_do (vb, (DictId)SHIM_OQ_DICT_ID_Q(q), __FUNCLINE);
        if (!c || !c->local.data) { count[q] = 0; continue; }
        count[q] = c->local.len32;
        if (!count[q] && monochars[q])                                      // the reference empties a monochar channel after filling it (:105): its bytes are still there,
            for (uint32_t i = 0; i < n_lines; i++)                          // their number is what the count pass found (:61-72)
                for (uint32_t k = 0; k < qual_len[i]; k++) count[q] += ((uint8_t)txt[qual_off[i] + k] - 33) == q;
        memcpy (channels + at, c->local.data, count[q]); at += count[q];
    }
    free (comp); free (vb);
    return 0;
}

// codec_oq_reconstruct line by line: the QUAL strings are in txt (as SAM_QUAL's last_txt), the OQ of every line lands at the end of txt_data
int ref_oq_decode (const uint8_t *txt, uint64_t txt_len, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *out_off, uint32_t n_lines, uint32_t key_bias,
                   const uint8_t *channels, const uint32_t *count, const uint8_t *monochars, uint8_t *out, uint64_t out_size)
{
    shim_init_piz ();
    if (setjmp (on_abort)) return -1;
    flag.out_dt = key_bias ? DT_SAM : DT_BAM;                               // sam_diff (:131)
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->data_type = DT_SAM; vb->lines.len = n_lines;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += qual_len[i];
    buf_alloc_do (vb, &vb->txt_data, txt_len + total + 64, 1, "txt_data", __FUNCLINE);
    memcpy (vb->txt_data.data, txt, txt_len); vb->txt_data.len = txt_len;
    ContextP ctx = CTX (OPTION_OQ_Z);
    ctx->did_i = OPTION_OQ_Z; ctx->is_loaded = true;
    buf_alloc_do (vb, &ctx->local, 94 + 8, 1, "local", __FUNCLINE);
    memcpy (ctx->local.data, monochars, 94); ctx->local.len = 94;
    uint64_t at = 0;
    for (int q = 0; q < 94; q++) {
        if (!count[q]) continue;                                            // a channel that is not in the file: ECTX returns NULL
        ContextP c = ctx_get_unmapped_ctx (&vb->ca, DT_SAM, (DictId)SHIM_OQ_DICT_ID_Q(q), 0, 0);
        buf_alloc_do (vb, &c->local, count[q] + 8, 1, "local", __FUNCLINE);
        memcpy (c->local.data, channels + at, count[q]); c->local.len = count[q]; at += count[q];
    }
    ContextP qual_ctx = CTX (SAM_QUAL);
    for (uint32_t i = 0; i < n_lines; i++) {
        qual_ctx->last_txt = (TxtWord){ .index = (uint32_t)qual_off[i], .len = qual_len[i] };
        const uint64_t before = vb->txt_data.len;
        codec_oq_reconstruct (vb, CODEC_OQ, ctx, qual_len[i], true);
        if (out_off[i] + qual_len[i] > out_size) return -2;
        memcpy (out + out_off[i], vb->txt_data.data + before, qual_len[i]);
    }
    free (vb);
    return 0;
}

// ================================================================ dyn_int_transpose (the reference's compiled dyn_int.c)
// what dyn_int.o wants from the rest of the program: the local-type table (the reference's own macro), and diagnostics the harness never reaches
const LocalTypeDesc lt_desc[NUM_LOCAL_TYPES] = LOCALTYPE_DESC;
rom lt_name (LocalType lt) { return "lt"; }
rom store_type_name (StoreType store) { return "store"; }
DataTypeProperties dt_props[NUM_DATATYPES], dt_props_def;
FileMode READ = "rb";
StrText char_to_printable (char c) { StrText t = {}; t.s[0] = c; return t; }
StrText1K str_time (void) { StrText1K t = {}; return t; }
StrText1K seg_error (VBlockP vb) { StrText1K t = {}; return t; }
noreturn void error_assertinp_failed (rom format, ...) { snprintf (abort_msg, sizeof abort_msg, "ASSINP %s", format); longjmp (on_abort, 1); }
static uint32_t shim_num_samples;
uint32_t vcf_header_get_num_samples (void) { return shim_num_samples; }
StrText1K str_str_s_ (rom label, STRp(str)) { StrText1K t = {}; return t; }
// lt_desc's file-to-native functions (buffer.c): the table needs their addresses, the harness never calls them
#define SHIM_LT_FN(f) void f (BufferP buf, LocalType *lt) { ABORT0 ("shim: " #f); }
SHIM_LT_FN (BGEN_deinterlace_d8_buf) SHIM_LT_FN (BGEN_deinterlace_d16_buf) SHIM_LT_FN (BGEN_deinterlace_d32_buf) SHIM_LT_FN (BGEN_deinterlace_d64_buf)
SHIM_LT_FN (BGEN_ptranspose_u8_buf) SHIM_LT_FN (BGEN_ptranspose_u16_buf) SHIM_LT_FN (BGEN_ptranspose_u32_buf)
SHIM_LT_FN (BGEN_transpose_u8_buf) SHIM_LT_FN (BGEN_transpose_u16_buf) SHIM_LT_FN (BGEN_transpose_u32_buf)
SHIM_LT_FN (BGEN_u8_buf) SHIM_LT_FN (BGEN_u16_buf) SHIM_LT_FN (BGEN_u64_buf)

// dyn_int_transpose on a local of n elements of `width` bytes: cols > 0 goes to local.n_cols (<= 255), cols_vcf to the number of samples
// of the VCF header (n_cols stays 0).  *transposed = the ltype became LT_UINT*_TR.
int ref_dyn_int_transpose (void *data, uint64_t n, uint32_t width, uint32_t cols, uint32_t cols_vcf, int *transposed)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->data_type = DT_VCF;
    shim_num_samples = cols_vcf;
    ContextP ctx = CTX (VCF_COPY_SAMPLE + 1);
    ctx->did_i = VCF_COPY_SAMPLE + 1; strcpy (ctx->tag_name, "shim");
    ctx->ltype = width == 1 ? LT_UINT8 : width == 2 ? LT_UINT16 : LT_UINT32;
    buf_alloc_do (vb, &ctx->local, n * width + 8, 1, "local", __FUNCLINE);
    memcpy (ctx->local.data, data, n * width); ctx->local.len = n; ctx->local.n_cols = cols;
    dyn_int_transpose (vb, ctx);
    *transposed = ctx->ltype == LT_UINT8_TR || ctx->ltype == LT_UINT16_TR || ctx->ltype == LT_UINT32_TR;
    memcpy (data, ctx->local.data, n * width);
    free (vb);
    return 0;
}

// ================================================================ b250_zip_generate (the reference's compiled b250.c)
bool is_fastq_pair_2 (VBlockP vb) { return false; }
bool fastq_zip_use_pair_identical (DictId dict_id) { return false; }
void ctx_decrement_count (VBlockP vb, ContextP ctx, WordIndex node_index) {}

// b250_zip_generate on a hand-made context: b250 = the segmenter's buffer, nodes = the word indices of the VBlock's new nodes.  out gets the converted
// buffer (it is the tail of the input buffer in the reference), *out_len its length
int ref_b250_generate (const uint8_t *b250, uint64_t len, const int32_t *ni2wi, uint32_t n_new, uint32_t ol_len, uint8_t *out, uint64_t *out_len)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->data_type = DT_SAM;
    ContextP ctx = CTX (SAM_RNAME);
    ctx->did_i = SAM_RNAME; strcpy (ctx->tag_name, "shim"); ctx->dict_id.num = 12345; ctx->nodes_converted = true;
    buf_alloc_do (vb, &ctx->b250, len + 8, 1, "b250", __FUNCLINE);
    memcpy (ctx->b250.data, b250, len); ctx->b250.len = len; ctx->b250.count = 2;                    // (count only matters with all_the_same)
    buf_alloc_do (vb, &ctx->nodes, (uint64_t)n_new * sizeof (WordIndex) + 8, 1, "nodes", __FUNCLINE);
    memcpy (ctx->nodes.data, ni2wi, (uint64_t)n_new * sizeof (WordIndex)); ctx->nodes.len = n_new;
    ctx->ol_nodes.len = ol_len;
    b250_zip_generate (vb, ctx);
    *out_len = ctx->b250.len;
    memcpy (out, ctx->b250.data, ctx->b250.len);
    free (vb);
    return 0;
}

// ================================================================ HOMP and T0 (the reference's compiled codec_homp.c, codec_t0.c)
static uint32_t *g_newlen;                                                  // the lines' lengths as the codec shortens them
static COMPRESSOR_CALLBACK (shim_get_line_newlen)
{
    *line_data = (char *)g_txt + g_off[vb_line_i];
    *line_data_len = g_newlen[vb_line_i];
    if (is_rev) *is_rev = 0;
}
void sam_update_qual_len (VBlockP vb, uint32_t line_i, uint32_t new_len)      { g_newlen[line_i] = new_len; }
void fastq_update_qual_len (VBlockP vb, uint32_t line_i, uint32_t new_len)    { g_newlen[line_i] = new_len; }
void sam_ultima_update_t0_len (VBlockP vb, uint32_t line_i, uint32_t new_len) { g_newlen[line_i] = new_len; }
COMPRESSOR_CALLBACK (sam_zip_t0) { shim_get_line_newlen (vb, ctx, vb_line_i, line_data, line_data_len, maximum_size, is_rev); }
QualHistType did_i_to_qht (Did did_i) { return QHT_QUAL; }

// codec_homp_compress (mode 0) / codec_t0_compress (mode 1) on n_lines strings: what the sub-codec receives (the condensed strings, line by line)
int ref_hp_condense (int mode, const uint8_t *txt, uint64_t txt_len, const uint64_t *str_off, const uint32_t *str_len, const uint64_t *seq_off, uint32_t n_lines,
                     uint8_t *local, uint64_t *local_len, uint32_t *new_len)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;
    g_txt = malloc (txt_len + 1); memcpy (g_txt, txt, txt_len);             // the codec condenses the lines in place
    g_off = str_off; g_len = str_len; g_seq_off = seq_off; g_seq_len = NULL; g_is_rev = NULL;
    g_newlen = new_len; memcpy (g_newlen, str_len, (uint64_t)n_lines * 4);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += str_len[i];
    ContextP ctx = CTX (mode == 0 ? SAM_QUAL : OPTION_t0_Z);
    ctx->did_i = mode == 0 ? SAM_QUAL : OPTION_t0_Z; strcpy (ctx->tag_name, mode == 0 ? "QUAL" : "t0:Z");
    ctx->local.len = total;
    SectionHeaderCtx header = {};
    uint32_t ulen = (uint32_t)total, clen = (uint32_t)total + 4096;
    char *comp = malloc (clen);
    cap_out = NULL; cap_len = 0;
    bool ok = mode == 0 ? codec_homp_compress (vb, ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_line_newlen, comp, &clen, true, "QUAL")
                        : codec_t0_compress   (vb, ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_line_newlen, comp, &clen, true, "t0:Z");
    if (!ok) return -3;
    if (ctx->local.len32 != cap_len) return -4;                             // the codec keeps local.len in step with the lines (:184, t0 :104)
    memcpy (local, cap_out, cap_len); *local_len = cap_len;
    free (comp); free (g_txt); free (vb);
    return 0;
}

// codec_homp_reconstruct / codec_t0_reconstruct line by line
int ref_hp_expand (int mode, const uint8_t *local, uint64_t local_len, const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, uint32_t n_lines, uint8_t *out, uint64_t *out_len)
{
    shim_init_piz ();
    if (setjmp (on_abort)) { shim_missing_ok = false; return -1; }
    shim_missing_ok = true;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;
    flag.out_dt = DT_SAM;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += len[i];
    ContextP c = CTX (mode == 0 ? SAM_QUAL : OPTION_t0_Z);
    c->did_i = mode == 0 ? SAM_QUAL : OPTION_t0_Z; c->is_loaded = true;
    buf_alloc_do (vb, &c->local, local_len + 8, 1, "local", __FUNCLINE);
    memcpy (c->local.data, local, local_len); c->local.len = local_len;
    buf_alloc_do (vb, &vb->txt_data, total + 64, 1, "txt_data", __FUNCLINE);
    uint64_t at = 0;
    for (uint32_t i = 0; i < n_lines; i++) {
        if (!len[i]) continue;
        cur_seq = (rom)txt + seq_off[i]; cur_is_rev = false; vb->seq_len = len[i];
        const uint64_t before = vb->txt_data.len;
        if (mode == 0) codec_homp_reconstruct (vb, CODEC_HOMP, c, len[i], true); else codec_t0_reconstruct (vb, CODEC_T0, c, len[i], true);
        if (c->next_local > local_len) { shim_missing_ok = false; return -2; }
        // a line without quality contributes one '*': lay the lines out in slots of len[i] bytes like the flat entry points do
        memcpy (out + at, vb->txt_data.data + before, vb->txt_data.len - before);
        at += len[i];
    }
    shim_missing_ok = false;
    *out_len = at;
    const int rc = c->next_local == local_len ? 0 : -2;
    free (vb);
    return rc;
}

// ================================================================ SMUX (the reference's compiled codec_smux.c)
COMPRESSOR_CALLBACK (fastq_zip_qual) { shim_get_qual_rev (vb, ctx, vb_line_i, line_data, line_data_len, maximum_size, is_rev); }   // (codec_smux_calc_stats only: never called here)
COMPRESSOR_CALLBACK (sam_zip_qual)   { shim_get_qual_rev (vb, ctx, vb_line_i, line_data, line_data_len, maximum_size, is_rev); }
// codec_smux_compress on n_lines reads: the five channel contexts' locals back to back, the header param
int ref_smux_mux (const uint8_t *txt, uint64_t txt_len, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *seq_off, const uint32_t *seq_len, const uint8_t *is_rev,
                  uint32_t n_lines, uint8_t *channels, uint32_t *count, uint8_t *n_param)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;
    g_txt = (uint8_t *)txt; g_off = qual_off; g_len = qual_len; g_seq_off = seq_off; g_seq_len = seq_len; g_is_rev = is_rev;
    ContextP ctx = CTX (SAM_QUAL);
    ctx->did_i = SAM_QUAL; strcpy (ctx->tag_name, "QUAL"); ctx->dict_id = (DictId)_SAM_QUAL;
    SectionHeaderCtx header = {};
    uint32_t ulen = 0, clen = 64;
    char comp[64] = {};
    if (!codec_smux_compress (vb, ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_qual_rev, comp, &clen, true, "QUAL")) return -3;
    *n_param = header.param;
    uint64_t at = 0;
    for (Did d = SHIM_FIRST_DYN_DID; d < vb->ca.num_contexts; d++) {       // created in the order A, C, G, T, N (decl_smux_ctxs_zip, :16-21)
        ContextP c = &vb->ca.contexts[d]; const int b = d - SHIM_FIRST_DYN_DID;
        if (b >= 5) return -4;
        count[b] = c->local.len32;
        if (b == 4 && header.param && !count[b])                             // dropped (:249-252): its bytes are still there, as many as the lines hold bases that are not A, C, G, T
            for (uint32_t i = 0; i < n_lines; i++) {
                const uint8_t *sq = txt + seq_off[i];
                if (qual_len[i] == 1 && txt[qual_off[i]] == ' ' && is_rev && is_rev[i]) { const uint8_t ch = sq[seq_len[i] - 1]; count[b] += !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T'); }
                else for (uint32_t k = 0; k < qual_len[i]; k++) count[b] += !(sq[k] == 'A' || sq[k] == 'C' || sq[k] == 'G' || sq[k] == 'T');
            }
        if (c->local.data && count[b]) { memcpy (channels + at, c->local.data, count[b]); at += count[b]; }
    }
    free (vb);
    return 0;
}

// codec_smux_reconstruct line by line (SAM in, SAM out): every line's bytes at out_off[i]; a read without quality contributes the one character '*'
int ref_smux_demux (const uint8_t *txt, uint64_t txt_len, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev, const uint64_t *out_off, uint32_t n_lines,
                    const uint8_t *channels, const uint32_t *count, uint8_t n_param, uint8_t *out, uint64_t out_size)
{
    shim_init_piz ();
    if (setjmp (on_abort)) { shim_missing_ok = false; return -1; }
    shim_missing_ok = true;
    flag.out_dt = DT_SAM;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += len[i];
    buf_alloc_do (vb, &vb->txt_data, total + 64, 1, "txt_data", __FUNCLINE);
    ContextP ctx = CTX (SAM_QUAL);
    ctx->did_i = SAM_QUAL; ctx->is_loaded = true; ctx->dict_id = (DictId)_SAM_QUAL; ctx->local.param = n_param;
    static const char base_of[5] = { 'A', 'C', 'G', 'T', 'N' };
    uint64_t at = 0;
    for (int b = 0; b < 5; b++) {
        if (!count[b]) continue;                                            // a channel that is not in the file: ECTX returns NULL
        const char *id = ctx->dict_id.id;
        ContextP c = ctx_get_unmapped_ctx (&vb->ca, DT_SAM, (DictId)DICT_ID_MAKEF_8(((char[]){base_of[b], base_of[b], base_of[b], '-', (id[0] & 0x7f) | 0x40, id[1], id[2], id[3]})), 0, 0);
        buf_alloc_do (vb, &c->local, count[b] + 8, 1, "local", __FUNCLINE);
        memcpy (c->local.data, channels + at, count[b]); c->local.len = count[b]; at += count[b];
    }
    for (uint32_t i = 0; i < n_lines; i++) {
        if (!len[i]) continue;
        cur_seq = (rom)txt + seq_off[i]; cur_is_rev = is_rev ? is_rev[i] : false; vb->seq_len = len[i];
        const uint64_t before = vb->txt_data.len;
        codec_smux_reconstruct (vb, CODEC_SMUX, ctx, len[i], true);
        if (out_off[i] + (vb->txt_data.len - before) > out_size) { shim_missing_ok = false; return -2; }
        memcpy (out + out_off[i], vb->txt_data.data + before, vb->txt_data.len - before);
    }
    shim_missing_ok = false;
    free (vb);
    return 0;
}

// ================================================================ TMPL (the reference's compiled codec_tmpl.c)
ContextP ctx_get_zctx_from_vctx (ConstContextP vctx, bool create_if_missing, bool follow_alias) { return &z_file->ca.contexts[vctx->did_i]; }   // the z-side twin: same slot
void ctx_segconf_set_hard_coded_lcodec (Did did_i, Codec codec) {}
void buf_insert_do (VBlockP vb, BufferP buf, unsigned width, uint64_t insert_at, const void *new_data, uint64_t new_data_len, rom name, FUNCLINE)
{
    if (insert_at != buf->len) ABORT0 ("shim: buf_insert_do only appends");
    buf_alloc_do (vb, buf, (buf->len + new_data_len) * width + 64, 1.5, name, func, code_line);
    memcpy (buf->data + buf->len * width, new_data, new_data_len * width);
    buf->len += new_data_len;
}

// codec_tmpl_segconf_finalize (the template: the most frequent quality per position) + codec_tmpl_compress on n_lines quality strings.
// Returns 1 when the reference finds the template not dominant enough (:67-74), else 0 with the template, the 94 channels and the excess back to back
int ref_tmpl_mux (const uint8_t *txt, uint64_t txt_len, const uint64_t *qual_off, const uint32_t *qual_len, uint32_t n_lines, uint32_t tmpl_len,
                  uint8_t *tmpl_out, uint8_t *channels, uint32_t *count)
{
    shim_init ();
    if (setjmp (on_abort)) return -1;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_FASTQ;
    g_txt = (uint8_t *)txt; g_off = qual_off; g_len = qual_len;
    segconf.std_seq_len = tmpl_len;
    ContextP ctx = CTX (FASTQ_QUAL);
    ctx->did_i = FASTQ_QUAL; strcpy (ctx->tag_name, "QUAL"); (ctx + 1)->did_i = FASTQ_QUAL + 1;
    ContextP zctx = &z_file->ca.contexts[FASTQ_QUAL];
    memset (zctx, 0, 2 * sizeof (Context)); zctx->did_i = FASTQ_QUAL; (zctx + 1)->did_i = FASTQ_QUAL + 1;
    codec_tmpl_segconf_finalize (vb, FASTQ_QUAL, shim_get_line);
    if (!zctx->tmpl_calculated) { free (vb); return 1; }
    memcpy (tmpl_out, (zctx + 1)->template.data, tmpl_len);
    SectionHeaderCtx header = {};
    uint32_t ulen = 0, clen = 64; char comp[64] = {};
    if (!codec_tmpl_compress (vb, ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_line, comp, &clen, true, "QUAL")) return -3;
    uint64_t at = 0;
    DictId tmpl_dict_id = (DictId)DICT_ID_MAKEF_4("tmpl");                   // codec_tmpl.c:77
    for (int q = 0; q < 94; q++) {
        Did d = ctx_get_unmapped_existing_did_i (&vb->ca, sub_dict_id (tmpl_dict_id, q + 33));
        count[q] = d == DID_NONE ? 0 : vb->ca.contexts[d].local.len32;
        if (count[q]) { memcpy (channels + at, vb->ca.contexts[d].local.data, count[q]); at += count[q]; }
    }
    count[94] = (ctx + 1)->local.len32;
    if (count[94]) memcpy (channels + at, (ctx + 1)->local.data, count[94]);
    free (vb);
    return 0;
}

// ================================================================ PACB (the reference's compiled codec_pacb.c)
static const uint8_t *g_np0;
int32_t sam_zip_get_np (VBlockP vb, LineIType line_i) { return g_np0 ? g_np0[line_i] + 1 : 1; }
uint32_t sam_zip_get_seq_len (VBlockP vb, uint32_t line_i)   { return g_len[line_i]; }
uint32_t fastq_zip_get_seq_len (VBlockP vb, uint32_t line_i) { return g_len[line_i]; }
DictId dict_id_make (STRp(str), DictIdType dict_id_type) { ABORT0 ("shim: dict_id_make"); }
bool container_peek_has_item (VBlockP vb, ContextP ctx, DictId item_dict_id, bool consume) { ABORT0 ("shim: container_peek_has_item"); }
int32_t reconstruct_from_ctx_do (VBlockP vb, Did did_i, char sep, ReconType reconstruct, rom func) { ABORT0 ("shim: reconstruct_from_ctx_do"); }
static int32_t g_cur_np;
ValueType reconstruct_peek (VBlockP vb, ContextP ctx, pSTRp(txt)) { ValueType v = { .i = g_cur_np }; return v; }

static void shim_pacb_setup (VBlockP vb, uint32_t max_np)
{
    memset (&the_txt_file, 0, sizeof the_txt_file); the_txt_file.data_type = DT_SAM; txt_file = &the_txt_file;
    if (max_np > 1) bitset_set ((uint64_t *)segconf.has_bits, OPTION_np_i);              // get_max_np (:56-59): MAX_np = 12 with np:i in the file, else 1
    ContextP zctx = &z_file->ca.contexts[SAM_QUAL];
    memset (zctx, 0, sizeof (Context)); zctx->did_i = SAM_QUAL;
    buf_alloc_do (NULL, &zctx->subdicts, (uint64_t)7 * max_np * sizeof (DictId) + 8, 1, "subdicts", __FUNCLINE);
    for (uint32_t c = 0; c < 7 * max_np; c++) B(DictId, zctx->subdicts, 0)[c].num = 0x5041434200ull + c;   // any distinct ids: the channel contexts are found by them
    zctx->subdicts.len = 7 * max_np;
}

// codec_pacb_compress on n_lines reads: the 7 * max_np channel contexts' locals back to back.  max_np must be 1 or 12 (the reference knows no other).
int ref_pacb_mux (const uint8_t *txt, uint64_t txt_len, const uint64_t *qual_off, const uint32_t *qual_len, const uint64_t *seq_off, const uint8_t *np0, uint32_t max_np,
                  uint32_t n_lines, uint8_t *channels, uint32_t *count)
{
    shim_init ();
    if (setjmp (on_abort)) { txt_file = NULL; return -1; }
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;
    g_txt = (uint8_t *)txt; g_off = qual_off; g_len = qual_len; g_seq_off = seq_off; g_seq_len = NULL; g_is_rev = NULL; g_np0 = np0;
    shim_pacb_setup (vb, max_np);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += qual_len[i];
    ContextP ctx = CTX (SAM_QUAL);
    ctx->did_i = SAM_QUAL; strcpy (ctx->tag_name, "QUAL"); ctx->local.len = total;
    SectionHeaderCtx header = {};
    uint32_t ulen = 0, clen = 64; char comp[64] = {};
    if (!codec_pacb_compress (vb, ctx, (SectionHeaderP)&header, NULL, &ulen, shim_get_line, comp, &clen, true, "QUAL")) { txt_file = NULL; return -3; }
    uint64_t at = 0;
    for (uint32_t c = 0; c < 7 * max_np; c++) {
        ContextP sc = &vb->ca.contexts[SHIM_FIRST_DYN_DID + c];                // created in channel order (codec_pacb_init_ctxs :34-35)
        count[c] = sc->local.len32;
        if (count[c]) { memcpy (channels + at, sc->local.data, count[c]); at += count[c]; }
    }
    txt_file = NULL; free (vb);
    return 0;
}

// codec_pacb_reconstruct line by line; every line's bytes at out_off[i] (a read without quality: the one character '*')
int ref_pacb_demux (const uint8_t *txt, uint64_t txt_len, const uint64_t *seq_off, const uint32_t *len, const uint8_t *np0, uint32_t max_np, const uint64_t *out_off, uint32_t n_lines,
                    const uint8_t *channels, const uint32_t *count, uint8_t *out, uint64_t out_size)
{
    shim_init_piz ();
    if (setjmp (on_abort)) { shim_missing_ok = false; txt_file = NULL; return -1; }
    shim_missing_ok = true;
    flag.out_dt = DT_SAM;
    VBlockP vb = calloc (1, sizeof (VBlock));
    vb->vblock_i = 1; vb->lines.len = n_lines; vb->data_type = DT_SAM;
    shim_pacb_setup (vb, max_np);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_lines; i++) total += len[i];
    buf_alloc_do (vb, &vb->txt_data, total + 64, 1, "txt_data", __FUNCLINE);
    ContextP ctx = CTX (SAM_QUAL);
    ctx->did_i = SAM_QUAL; ctx->is_loaded = true;
    CTX (OPTION_np_i)->is_loaded = true;
    ContextP zctx = &z_file->ca.contexts[SAM_QUAL];
    uint64_t at = 0;
    for (uint32_t c = 0; c < 7 * max_np; c++) {                                // every channel context exists (ECTX must not return NULL, :286-289)
        ContextP sc = ctx_get_unmapped_ctx (&vb->ca, DT_SAM, *B(DictId, zctx->subdicts, c), 0, 0);
        buf_alloc_do (vb, &sc->local, count[c] + 8, 1, "local", __FUNCLINE);
        memcpy (sc->local.data, channels + at, count[c]); sc->local.len = count[c]; at += count[c];
    }
    for (uint32_t i = 0; i < n_lines; i++) {
        if (!len[i]) continue;
        cur_seq = (rom)txt + seq_off[i]; cur_is_rev = false; g_cur_np = np0 ? np0[i] + 1 : 1;
        CTX (OPTION_np_i)->last_value.i = g_cur_np; CTX (OPTION_np_i)->last_line_i = vb->line_i;   // np:i of this line was reconstructed before QUAL (codec_pacb_piz_get_np :243-245)
        const uint64_t before = vb->txt_data.len;
        codec_pacb_reconstruct (vb, CODEC_PACB, ctx, len[i], true);
        if (out_off[i] + (vb->txt_data.len - before) > out_size) { shim_missing_ok = false; txt_file = NULL; return -2; }
        memcpy (out + out_off[i], vb->txt_data.data + before, vb->txt_data.len - before);
    }
    shim_missing_ok = false; txt_file = NULL;
    free (vb);
    return 0;
}

// ================================================================ zip_generate_local's transforms: the reference's own macros
// (INTERLACE / DEINTERLACE of context.h:98-101, BGEN16/32/64 of endianness.h) in the loops of buffer.c:337-345, :431-468
int ref_local_transform (int op, void *data, uint64_t n)
{
    switch (op) {
        case 1:  { uint16_t *p = data; for (uint64_t i = 0; i < n; i++) p[i] = BGEN16 (p[i]); return 0; }
        case 2:  { uint32_t *p = data; for (uint64_t i = 0; i < n; i++) p[i] = BGEN32 (p[i]); return 0; }
        case 3:  { uint64_t *p = data; for (uint64_t i = 0; i < n; i++) p[i] = BGEN64 (p[i]); return 0; }
        case 4:  { int8_t  *p = data; for (uint64_t i = 0; i < n; i++) p[i] =         (INTERLACE (int8_t,  p[i])); return 0; }
        case 5:  { int16_t *p = data; for (uint64_t i = 0; i < n; i++) p[i] = BGEN16 (INTERLACE (int16_t, p[i])); return 0; }
        case 6:  { int32_t *p = data; for (uint64_t i = 0; i < n; i++) p[i] = BGEN32 (INTERLACE (int32_t, p[i])); return 0; }
        case 7:  { int64_t *p = data; for (uint64_t i = 0; i < n; i++) p[i] = BGEN64 (INTERLACE (int64_t, p[i])); return 0; }
        case 8:  { uint8_t  *p = data; for (uint64_t i = 0; i < n; i++) { uint8_t  u = p[i];          ((int8_t  *)p)[i] = DEINTERLACE (int8_t,  u); } return 0; }
        case 9:  { uint16_t *p = data; for (uint64_t i = 0; i < n; i++) { uint16_t u = BGEN16 (p[i]); ((int16_t *)p)[i] = DEINTERLACE (int16_t, u); } return 0; }
        case 10: { uint32_t *p = data; for (uint64_t i = 0; i < n; i++) { uint32_t u = BGEN32 (p[i]); ((int32_t *)p)[i] = DEINTERLACE (int32_t, u); } return 0; }
        case 11: { uint64_t *p = data; for (uint64_t i = 0; i < n; i++) { uint64_t u = BGEN64 (p[i]); ((int64_t *)p)[i] = DEINTERLACE (int64_t, u); } return 0; }
        default: return -1;
    }
}
