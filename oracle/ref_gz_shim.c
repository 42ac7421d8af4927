// ref_gz_shim.c — TEST INFRASTRUCTURE.  Hosts the REFERENCE's own codec_domq.c (compiled unmodified from /root/reference/src,
// see oracle/Makefile) outside the genozip program: this file is compiled against the reference's headers (so VBlock, Context,
// Buffer, CodecArgs ... have the reference's exact layouts), hand-constructs the minimal VBlock a compute thread would hand
// to the codec, supplies the ~20 host symbols the object needs (buffer allocation, seg_by_ctx, codec table ...) and exports
// one flat entry point:
//
//     ref_domq_encode ()   =  codec_domq_comp_init (force)  +  codec_domq_compress ()          (codec_domq.c:299-323, 379-521)
//
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
#include "sections.h"
#include <stdarg.h>
#include <setjmp.h>

// ---------------------------------------------------------------- globals the objects reference
Flags flag;
SegConf segconf;
static File the_z_file;
FileP z_file = &the_z_file, txt_file = NULL;
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
WordIndex ctx_peek_next_snip (VBlockP vb, ContextP ctx, pSTRp (snip)) { *snip = NULL; *snip_len = 0; return 0; }
// PIZ / consensus-read paths of codec_domq.c that the encoder harness never takes
void sam_reconstruct_missing_quality (VBlockP vb, ReconType reconstruct) { ABORT0 ("shim: sam_reconstruct_missing_quality"); }
void sam_xcons_reconstruct_QUAL (VBlockP vb, ContextP ctx, uint32_t qual_len, bool reconstruct) { ABORT0 ("shim: sam_xcons_reconstruct_QUAL"); }
void sam_xcons_split_qual_line (VBlockP vb, BufferP ql_buf) { ABORT0 ("shim: sam_xcons_split_qual_line"); }

// ---------------------------------------------------------------- the codec table: every sub-codec is "store"
static uint8_t *cap_out; static uint32_t cap_len;
static COMPRESS (shim_store)
{
    memcpy (compressed, uncompressed, *uncompressed_len);
    *compressed_len = *uncompressed_len;
    cap_out = (uint8_t *)compressed; cap_len = *uncompressed_len;
    return true;
}
static uint32_t shim_est_size (Codec codec, uint64_t uncompressed_len) { return (uint32_t)uncompressed_len + 64; }
CodecArgs codec_args[NUM_CODECS];
Codec codec_assign_best_codec (VBlockP vb, ContextP ctx, BufferP non_ctx_data, SectionType st) { return CODEC_NONE; }

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
    info_stream = stderr;
    flag.show_time_comp_i = COMP_NONE;                                      // profiler off (profiler.h:119)
    for (int c = 0; c < NUM_CODECS; c++) { codec_args[c].compress = shim_store; codec_args[c].est_size = shim_est_size; }
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
