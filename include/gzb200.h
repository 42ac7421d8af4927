/* gzb200.h — C-ABI of libgzb200.so: a B200 (sm_100a) implementation of genozip's per-VBlock codec path.
 *
 * This is the drop-in boundary.  Two layers:
 *
 *  (1) FLAT entry points — plain pointers and sizes, no genozip types, batched ("sections") because one
 *      launch per section is launch-latency-bound.  These are what a cgo/JNI/ctypes binding would bind, and
 *      what the thin adapter inside genozip calls.
 *  (2) PLUG-IN entry points with exactly the reference's COMPRESS / UNCOMPRESS / est_size signatures
 *      (reference src/codec.h:17-40) so that CODEC_ARGS (src/codec.h:47-115) can point at them; genozip's
 *      VBlock/Context/SectionHeader/Buffer are opaque here and are touched only through the accessor table
 *      the adapter registers (gzb_plugin_host).  See INTEGRATION.md for the reference-side stub.
 *
 * There is NO CPU fallback: every entry point fails (non-zero return / GZB_E_NOCUDA) when no CUDA device
 * or kernel image is available.
 *
 * Output contract: for a given (codec, uncompressed bytes) the compressed bytes are identical to those the
 * reference codec function writes (rans_compress_to_4x16 / arith_compress_to and the genozip codecs built
 * on them), and decompression is bit-exact.
 */
#ifndef GZB200_H
#define GZB200_H
#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- codec ids: numeric values of the reference's `Codec` enum (src/genozip.h:322-360) ---- */
enum {
    GZB_CODEC_NONE = 1,
    GZB_CODEC_RANB = 6,  GZB_CODEC_RANW = 7,  GZB_CODEC_RANb = 8,  GZB_CODEC_RANw = 9,
    GZB_CODEC_ACGT = 10, GZB_CODEC_XCGT = 11, GZB_CODEC_DOMQ = 13, GZB_CODEC_PBWT = 15,
    GZB_CODEC_ARTB = 16, GZB_CODEC_ARTW = 17, GZB_CODEC_ARTb = 18, GZB_CODEC_ARTw = 19,
    GZB_CODEC_LONGR = 26,
};

/* ---- status codes ---- */
enum {
    GZB_OK = 0,
    GZB_SOFT_FAIL   = 1,    /* output capacity < est_size: the reference returns false under soft_fail (src/compressor.c:90) */
    GZB_E_NOCUDA    = -1,   /* no device / no sm_100a kernel image: the product never falls back to the CPU */
    GZB_E_CUDA      = -2,   /* a CUDA call failed (see gzb_last_error) */
    GZB_E_BADARG    = -3,
    GZB_E_CORRUPT   = -4,   /* malformed compressed data (the reference ASSERTs, src/codec_htscodecs.c:106-111) */
    GZB_E_UNSUPPORTED = -5, /* valid input that the bulk form of an entry point does not cover (stated at that entry point) */
};

typedef struct gzb_engine gzb_engine;   /* one per (host thread, GPU): a CUDA stream + reusable device arena */

/* flags for the batch calls */
#define GZB_DEVICE_PTRS  1u   /* every pointer is a device pointer on the engine's GPU (inputs resident in HBM) */
#define GZB_OUT_DEVICE   2u   /* bulk OUTPUT streams are device pointers, everything else host (chaining codecs without a PCIe round trip) */
#define GZB_IN_DEVICE    4u   /* bulk INPUT streams are device pointers, everything else host */
/* gzb_section.sflags: per-section override when the batch flags are 0 */
#define GZB_SEC_IN_DEVICE  1u
#define GZB_SEC_OUT_DEVICE 2u

/* ---------------------------------------------------------------- lifecycle (SURVEY §8b "lifecycle") */
int   gzb_device_count (void);                                    /* number of visible CUDA devices, 0 if none */
int   gzb_build_is_emulation (void);                              /* 0 for the product (nvcc, sm_100a); 1 for the test suite's host build of the same
                                                                     sources (tests/host/simt), which no binding may load outside the tests */
int   gzb_engine_create (int device, gzb_engine **out);           /* GZB_E_NOCUDA if no usable device */
void  gzb_engine_destroy (gzb_engine *e);
const char *gzb_last_error (gzb_engine *e);                       /* e may be NULL: creation errors */
void *gzb_engine_stream (gzb_engine *e);                          /* the cudaStream_t all work of this engine is ordered on */
int   gzb_engine_sync (gzb_engine *e);
int   gzb_engine_trim (gzb_engine *e);                            /* frees the engine's grow-only device workspace and pinned staging (they come back on demand) */
int   gzb_vb_device (uint32_t vblock_i, int n_devices);           /* (vblock_i-1) mod n_devices — the dispatcher's round-robin */
/* staging (gzb_vb_stage): asynchronous host<->device transfers on the engine's copy stream, so that the host uploads
 * the next batch's text (vb->txt_data) and fetches the previous batch's sections (vb->z_data) while the current batch's kernels
 * run on buffers it passes with GZB_DEVICE_PTRS.  Host memory should be page-locked.  Uploads and fetches are two queues (the link is
 * full duplex), each in order; gzb_stage_wait returns when every upload / every fetch / every transfer queued so far has completed. */
enum { GZB_STAGE_UPLOADS = 0, GZB_STAGE_FETCHES = 1, GZB_STAGE_ALL = 2 };
int   gzb_stage_upload (gzb_engine *e, void *dst_device, const void *src_host, uint64_t bytes);
int   gzb_stage_fetch  (gzb_engine *e, void *dst_host, const void *src_device, uint64_t bytes);
int   gzb_stage_wait   (gzb_engine *e, int which);
uint64_t gzb_kernel_launches (gzb_engine *e);                     /* kernels launched by this engine so far */
/* Device-time of the dominant chain kernels of the LAST batch call, ms (CUDA events on the engine's stream) */
float gzb_last_chain_ms (gzb_engine *e);
/* the same, split by coder: which = 0 rANS chain kernel, 1 general arithmetic chain kernel (order-1 / RLE leaves), 3 order-0 arithmetic
 * chain kernel, 4 the split arithmetic encoder (bucket + model + code kernels of the long order-1 leaves), 5 the longest of 1, 3, 4
 * (they run side by side), 6 / 7 / 8 the split encoder's bucket / model / code kernels; which = 2: the dominant kernel of the last PBWT / LONGR batch call (the row walk k_pbwt_rows, the channel
 * walk k_longr_channels / k_longr_decode) */
float gzb_last_kernel_ms (gzb_engine *e, int which);

/* ---------------------------------------------------------------- simple codecs: rANS 4x16 and adaptive arithmetic
 * Replaces codec_{RANB,RANW,RANb,RANw,ARTB,ARTW,ARTb,ARTw}_compress (src/codec_htscodecs.c:77-94) →
 * rans_compress_to_4x16 (src/htscodecs/rANS_static4x16pr.c:1151) / arith_compress_to (src/htscodecs/arith_dynamic.c:615),
 * and codec_rans_uncompress / codec_arith_uncompress (src/codec_htscodecs.c:100-129). */
typedef struct {
    int32_t     codec;      /* GZB_CODEC_RANB … GZB_CODEC_ARTw */
    int32_t     status;     /* out: GZB_OK / GZB_SOFT_FAIL / GZB_E_* */
    const void *in;         /* uncompressed bytes (compress) or compressed bytes (uncompress) */
    void       *out;
    uint32_t    in_len;
    uint32_t    out_cap;    /* compress: capacity of out (must be >= gzb_est_size(codec,in_len) or status=GZB_SOFT_FAIL);
                               uncompress: the expected uncompressed length */
    uint32_t    out_len;    /* out: bytes written */
    uint32_t    sflags;     /* GZB_SEC_IN_DEVICE | GZB_SEC_OUT_DEVICE */
} gzb_section;

uint32_t gzb_est_size (int codec, uint64_t uncompressed_len);     /* codec_*_est_size (src/codec_htscodecs.c:26-33): 1 KB + bound */
int gzb_compress_sections   (gzb_engine *e, gzb_section *secs, uint32_t n, uint32_t flags);
int gzb_uncompress_sections (gzb_engine *e, gzb_section *secs, uint32_t n, uint32_t flags);

/* Packed output: the compressed sections of the batch are APPENDED to one buffer in section order, 16-byte aligned, the way
 * zfile_compress_local_data appends a section to vb->z_data (src/zfile.c:229-262) — no per-section capacity of est_size bytes
 * has to exist anywhere.  secs[i].out / out_cap are ignored on entry; on return secs[i].out points into `arena` and out_len is
 * set.  *arena_used = bytes needed; if that exceeds arena_cap nothing is written and the call returns GZB_SOFT_FAIL (grow the
 * buffer and call again, like the soft-fail retry of src/compressor.c:90-110).  `arena` is a device pointer with GZB_DEVICE_PTRS
 * or GZB_OUT_DEVICE, else host memory (one transfer for the whole batch). */
int gzb_compress_sections_packed (gzb_engine *e, gzb_section *secs, uint32_t n, void *arena, uint64_t arena_cap, uint64_t *arena_used, uint32_t flags);

/* ---------------------------------------------------------------- codec assignment by size
 * codec_assign_best_codec (src/codec.c:234-389) without its clock: the first min (len, CODEC_ASSIGN_SAMPLE_SIZE) bytes of every item
 * (src/codec.h:154; one item per context whose codec is still CODEC_UNKNOWN) are compressed with the eight simple codecs of this path
 * in ONE batch; nothing is written, only the compressed lengths come back.  size[k] = body bytes with codec RANB, RANW, RANb, RANw,
 * ARTB, ARTW, ARTb, ARTw (k = 0..7) — the bytes the reference's own test compressions produce.  best = what the sorter's size rules
 * leave (:146-149,167-172): the smallest section (body + the 28-byte SectionHeader the reference's measurement includes, :331-333), equal
 * sizes to the lower Codec value, CODEC_NONE if the bare sample (:325) is not larger than that; GZB_CODEC_UNKNOWN if the sample is
 * below MIN_LEN_FOR_COMPRESSION (:317-318, the section then goes out as CODEC_NONE, compressor.c:56-58).  The reference also weighs
 * clock () time and offers BZ2 / BSC / LZMA: host policy, out of scope — the adapter can still apply its sorter to size[] plus its own
 * timings.  Device pointers with GZB_DEVICE_PTRS / GZB_IN_DEVICE, else host memory. */
#define GZB_ASSIGN_SAMPLE_SIZE        99999u
#define GZB_MIN_LEN_FOR_COMPRESSION   50u
#define GZB_SECTION_HEADER_BYTES      28u
#define GZB_CODEC_UNKNOWN             0
typedef struct {
    const void *data;          /* the context's local / b250 / dict data */
    uint64_t    len;           /* its length in bytes */
    uint32_t    sample_len;    /* out: bytes that were compressed */
    uint32_t    size[8];       /* out: compressed body bytes per codec */
    int32_t     best;          /* out: GZB_CODEC_* */
} gzb_assign_item;
int gzb_assign_codecs (gzb_engine *e, gzb_assign_item *items, uint32_t n, uint32_t flags);

/* n device-to-device copies in one launch (compacting the streams a complex codec produced into right-sized buffers) */
typedef struct { const void *src; void *dst; uint64_t len; } gzb_copy;
int gzb_copy_batch (gzb_engine *e, const gzb_copy *copies, uint32_t n);

/* ---------------------------------------------------------------- Adler-32 of buffers that are in HBM
 * adler32 (1, data, len) of n buffers in one call: the z_digest of every section body (src/compressor.c:151,161 — the reference
 * computes it on the compute thread right after compressing; here the bodies are still in the packed device buffer of
 * gzb_compress_sections_packed), the --verify-codec digest of the uncompressed data (:72-74), or a VBlock's reconstructed text
 * (src/digest.c:62).  `adler` is the value adler32 () returns (the caller applies BGEN32).  Device pointers with
 * GZB_DEVICE_PTRS / GZB_IN_DEVICE, else host memory (uploaded first). */
typedef struct { const void *data; uint64_t len; uint32_t adler; uint32_t reserved; } gzb_digest_item;
int gzb_adler32_batch (gzb_engine *e, gzb_digest_item *items, uint32_t n, uint32_t flags);

/* ---------------------------------------------------------------- zip_generate_local's transforms of a context's local buffer, in place
 * (src/zip.c:167-213 -> src/buffer.c:337-353; PIZ inverses src/buffer.c:431-468, src/local_type.h:70-90) for a batch of buffers:
 *   SWAP16/32/64        BGEN_u16/u32/u64_buf on a little-endian host (LT_UINT*, LT_hex*, LT_FLOAT*); its own inverse
 *   INTERLACE8/16/32/64 interlace_d8_buf, BGEN_interlace_d16/32/64_buf (LT_INT*): INTERLACE (src/context.h:99), then big endian
 *   DEINTERLACE*        BGEN_deinterlace_d*_buf: from big endian, then DEINTERLACE (src/context.h:100)
 * n_elems counts elements of the operation's width; data must be aligned to it.  Device pointers with GZB_DEVICE_PTRS, else host memory. */
enum { GZB_LT_SWAP16 = 1, GZB_LT_SWAP32, GZB_LT_SWAP64, GZB_LT_INTERLACE8, GZB_LT_INTERLACE16, GZB_LT_INTERLACE32, GZB_LT_INTERLACE64,
       GZB_LT_DEINTERLACE8, GZB_LT_DEINTERLACE16, GZB_LT_DEINTERLACE32, GZB_LT_DEINTERLACE64 };
typedef struct { void *data; uint64_t n_elems; int32_t op; int32_t status; } gzb_local_item;
int gzb_local_transform_batch (gzb_engine *e, gzb_local_item *items, uint32_t n, uint32_t flags);

/* Matrix transposes of a local buffer (a rows x cols matrix of 8 / 16 / 32-bit integers, one row per line):
 *   GZB_TR_ZIP  dyn_int_transpose (src/dyn_int.c:45-105, without copied samples): trans[c * rows + r] = data[r * cols + c]; n_elems not a
 *               multiple of cols: left as it is, transposed = 0 (:75-78).  Runs after the endianness transform, as in zip_generate_local (src/zip.c:213-214).
 *   GZB_TR_PIZ  BGEN_transpose_u8/16/32_buf (src/buffer.c:364-391): back to row by row, then from big endian.
 * Device pointers with GZB_DEVICE_PTRS, else host memory. */
enum { GZB_TR_ZIP = 0, GZB_TR_PIZ = 1 };
typedef struct { void *data; uint64_t n_elems; uint32_t cols; uint8_t width /* bytes: 1, 2, 4 */; uint8_t dir; uint8_t transposed /* out */; uint8_t pad; int32_t status; int32_t reserved; } gzb_transpose_item;
int gzb_local_transpose_batch (gzb_engine *e, gzb_transpose_item *items, uint32_t n, uint32_t flags);

/* ---------------------------------------------------------------- b250_zip_generate (src/b250.c:202-297) for a batch of contexts
 * `b250` = the context's b250 buffer as the segmenter left it (little-endian variable-length words, the type in the LAST byte, :137-164);
 * `out` (len bytes, another buffer) receives the PIZ form (big endian, type first), right-aligned as the reference leaves it in its own
 * buffer: the result is the last out_len bytes (the adapter points ctx->b250.data there, :276-281).  On the way node indices >= ol_len
 * become ni2wi[node_index - ol_len] (node_index_to_word_index, src/context.h:109-112) and a word equal to its predecessor + 1 becomes
 * ONE_UP when one_up_ok (nodes.len + ol_nodes.len32 > 1024, :247).  What follows in the reference — the pair-identical test and
 * codec_assign_best_codec (:283-296) — stays with the caller (gzb_assign_codecs for the latter).
 * GZB_E_CORRUPT: the words do not tile the buffer, or a word index that cannot be encoded.  Device pointers with GZB_DEVICE_PTRS. */
typedef struct {
    const void    *b250;     uint64_t len;
    void          *out;
    const int32_t *ni2wi;    /* B(WordIndex, vctx->nodes, i) after conversion */
    uint32_t       n_new;    /* vctx->nodes.len */
    uint32_t       ol_len;   /* vctx->ol_nodes.len32 */
    uint8_t        one_up_ok;
    uint8_t        pad[3];
    int32_t        status;
    uint64_t       out_len;  /* out */
    uint64_t       n_words;  /* out */
} gzb_b250_item;
int gzb_b250_generate_batch (gzb_engine *e, gzb_b250_item *items, uint32_t n, uint32_t flags);

/* ---------------------------------------------------------------- ACGT / XCGT (src/codec_acgt.c)
 * pack:   codec_acgt_compress up to the sub-codec call (:64-163): bases → LE 2-bit words + exception stream.
 *         `packed` receives gzb_acgt_packed_len(n) bytes; `x` (n bytes) may be NULL if the caller declares
 *         acgt_no_x; *x_all_zero is set when the exception stream is all zero (=> header flag acgt_no_x, :136-140).
 * unpack: codec_acgt_uncompress/codec_xcgt_uncompress after their sub-codec call (:185-248). x may be NULL.
 * With host buffers, GZB_OUT_DEVICE (pack) / GZB_IN_DEVICE (unpack) make `x` alone a device pointer: the exception stream
 * stays in HBM between ACGT and its XCGT sub-codec section, like NONREF.local is overlaid on NONREF_X.local in the reference (:97-100). */
uint64_t gzb_acgt_packed_len (uint64_t n_bases);
int gzb_acgt_pack   (gzb_engine *e, const void *seq, uint64_t n_bases, void *packed, void *x, int *x_all_zero, uint32_t flags);
int gzb_acgt_unpack (gzb_engine *e, const void *packed, const void *x, uint64_t n_bases, void *seq, uint32_t flags);

/* Batched forms — all VBlocks of a batch in one launch and one synchronisation (what a dispatcher that hands the
 * GPU many VBlocks at once, SURVEY §8b item 4, calls instead of one gzb_acgt_pack per compute thread).
 * pack:   in  seq, n_bases, packed, x (may be NULL)      out  packed bytes, x bytes, x_all_zero
 * unpack: in  packed, x (NULL = acgt_no_x), n_bases, seq out  seq bytes */
typedef struct gzb_acgt_vb {
    const void *seq;        /* unpack: output buffer (written) */
    uint64_t    n_bases;
    void       *packed;
    void       *x;
    int32_t     x_all_zero; /* pack: out */
    uint32_t    reserved;
} gzb_acgt_vb;
int gzb_acgt_pack_batch   (gzb_engine *e, gzb_acgt_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_acgt_unpack_batch (gzb_engine *e, const gzb_acgt_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- DOMQ (src/codec_domq.c)
 * A VBlock's quality lines are given as a text buffer plus a (offset,len) table (H6 in SURVEY §7: never a per-line callback). */
typedef struct {
    const void     *txt;         /* buffer holding the quality strings (vb->txt_data or ctx->local) */
    uint64_t        txt_len;
    const uint64_t *line_off;    /* n_lines offsets into txt */
    const uint32_t *line_len;    /* n_lines lengths (0 = skip line, :421) */
    uint32_t        n_lines;
    /* outputs of gzb_domq_prepare (codec_domq_prepare_normalize :252-293) */
    uint8_t        *line_dom;    /* n_lines: compacted dom index per line */
    uint8_t        *line_diverse;/* n_lines: 1 if dom < 85% of the line (:141,160) */
    uint8_t         num_norm_qs; /* no_doms marker; section param = num_norm_qs|0x80 (:234) */
    uint8_t         num_doms;
    uint8_t         has_diverse;
    uint8_t         pad;
    uint8_t         denorm[95*95];    /* [num_doms][num_norm_qs] — base64-segged into DOMQRUNS by the host (:241-244) */
    uint8_t         normalize[95*95]; /* [cdom*95 + q-32] */
    /* outputs of gzb_domq_split (codec_domq_compress :379-500, before the sub-codec) */
    void *qual;  uint32_t qual_cap,  qual_len;   /* QUAL.local     — capacity >= 2*total_len+1 */
    void *runs;  uint32_t runs_cap,  runs_len;   /* DOMQRUNS.local — capacity >= total_len+1   */
    void *mplx;  uint32_t mplx_cap,  mplx_len;   /* QUALMPLX.local — capacity >= n_lines       */
    void *divr;  uint32_t divr_cap,  divr_len;   /* DIVRQUAL.local — capacity >= total_len     */
} gzb_domq_vb;

/* prepare: per-line histogram/dom on the GPU, per-dom rank tables on the host with libc qsort (tie order must
 * match the reference's qsort call).  split: normalise + stream split on the GPU.  All VBs of the
 * batch are processed together.  With GZB_DEVICE_PTRS txt/line tables/outputs are device pointers. */
int gzb_domq_prepare (gzb_engine *e, gzb_domq_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_domq_split   (gzb_engine *e, gzb_domq_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* codec_domq_reconstruct (:774-809) for all lines of a VB at once ("pre-decode the whole VB's QUAL
 * into a staging buffer so that the per-line reconstructor is a memcpy").  line_len[] are the seq_len the
 * reconstructor would be called with; out receives the concatenated quality strings. */
typedef struct {
    const void *qual;  uint32_t qual_len;
    const void *runs;  uint32_t runs_len;
    const void *mplx;  uint32_t mplx_len;
    const void *divr;  uint32_t divr_len;
    const uint8_t  *denorm;       /* [num_doms][num_norm_qs] (host memory) */
    uint32_t        denorm_len;
    uint8_t         num_norm_qs;  /* section param & 0x7f */
    const uint32_t *line_len;
    uint32_t        n_lines;
    void           *out;          /* sum(line_len) bytes */
    uint64_t        out_cap;
} gzb_domq_piz_vb;
int gzb_domq_reconstruct (gzb_engine *e, gzb_domq_piz_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- NORMQ (src/codec_normq.c): the fallback quality codec of SAM / BAM
 * gather:      codec_normq_compress before its sub-codec (:43-62): the quality strings of a VBlock copied into one buffer
 *              (QUAL.local), a string reversed where its read is reverse-complemented (is_rev); line_len = what the line callback
 *              returns (1 for a SAM line without quality: the byte ' ').     in  txt, line_off, line_len, is_rev    out  local, local_len
 * reconstruct: codec_normq_reconstruct (:85-106) for every line of a VBlock at once: line_len[i] = the `len` of the i-th call
 *              (the read's seq_len), is_rev[i] = last_flags.rev_comp.  A line whose byte in the stream is ' ' has no quality: it
 *              consumes that one byte, `missing[i]` is set and the first byte of its slot is the '*' of
 *              sam_reconstruct_missing_quality (src/sam_qual.c:532).          in  local, local_len, line_len, is_rev    out  out (line_len[i] bytes per line), missing
 * GZB_E_CORRUPT when the stream does not match the lines.  Device pointers with GZB_DEVICE_PTRS (then local_cap / out_cap bound the work). */
typedef struct gzb_normq_vb {
    const void     *txt;        uint64_t txt_len;
    const uint64_t *line_off;   /* gather */
    const uint32_t *line_len;
    const uint8_t  *is_rev;     /* may be NULL */
    uint32_t        n_lines;
    int32_t         status;
    void           *local;      uint64_t local_cap, local_len;
    void           *out;        uint64_t out_cap;
    uint8_t        *missing;    /* reconstruct, optional */
} gzb_normq_vb;
int gzb_normq_gather      (gzb_engine *e, gzb_normq_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_normq_reconstruct (gzb_engine *e, gzb_normq_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- HOMP and T0 (src/codec_homp.c, src/codec_t0.c): Ultima's homopolymer codecs
 * condense: the first pass of codec_homp_compress (:132-190, mode GZB_HP_HOMP, the string is QUAL) / codec_t0_compress (:69-109, mode
 *           GZB_HP_T0, the string is t0:Z): every homopolymer run of the read's SEQ keeps the part of its string that is not implied — the
 *           first half of a palindrome up to its first 'I' (HOMP), one character of a constant run (T0) — or, when the run does not
 *           comply, its first character | 0x80 and the rest.  The condensed strings come back to back in `local` (what the sub-codec
 *           reads through the line callback after sam_update_qual_len / sam_ultima_update_t0_len), new_len[i] = each line's new length.
 *               in  txt, str_off, str_len, seq_off (SEQ has the string's length), n_lines      out  local, local_len, new_len (optional)
 * expand:   codec_homp_reconstruct (:213-276) / codec_t0_reconstruct (:137-179) for every line of a VBlock in order: str_len[i] = the `len`
 *           of the i-th call.  HOMP: a line whose next byte is ' ' has no quality (one byte consumed, missing[i] set, the '*' of
 *           sam_reconstruct_missing_quality at the start of its slot).  Its deep / translated-to-FASTQ variants (:222-234) are not covered.
 *               in  local, local_len, txt + seq_off (the reconstructed SEQ), str_len      out  out (str_len[i] bytes per line, back to back), missing
 * GZB_E_CORRUPT when the stream does not match the lines.  Device pointers with GZB_DEVICE_PTRS (then local_cap / out_cap bound the work). */
enum { GZB_HP_HOMP = 0, GZB_HP_T0 = 1 };
typedef struct gzb_homp_vb {
    const void     *txt;        uint64_t txt_len;
    const uint64_t *str_off;    /* condense */
    const uint32_t *str_len;
    const uint64_t *seq_off;
    uint32_t        n_lines;
    int32_t         status;
    void           *local;      uint64_t local_cap, local_len;
    uint32_t       *new_len;    /* condense, optional */
    void           *out;        uint64_t out_cap;
    uint8_t        *missing;    /* expand, optional */
} gzb_homp_vb;
int gzb_homp_condense (gzb_engine *e, gzb_homp_vb *vbs, uint32_t n_vbs, int mode, uint32_t flags);
int gzb_homp_expand   (gzb_engine *e, gzb_homp_vb *vbs, uint32_t n_vbs, int mode, uint32_t flags);

/* ---------------------------------------------------------------- OQ (src/codec_oq.c): a read's original quality string OQ:Z multiplexed by its QUAL
 * mux:   codec_oq_compress before its sub-codec (:54-121).  The OQ character at position i of a line goes to channel QUAL[i] - '!' (94
 *        channels); count[] = the QUAL characters of ALL lines per channel (the reference's count pass, :61-72, which sizes the channels),
 *        while only the lines whose seq_len is not 0 are distributed (:94) — bytes nobody writes stay 0; monochars[q] = the character of a
 *        channel that holds one character only, else 0 (:103-107: the reference then drops that channel and stores the 94 monochars with RANB).
 *        oq_off[i] = dl->OQ as it is (0 for a line without OQ:Z: the reference reads the start of txt_data then, :91 — so does this).
 *            in  txt, qual_off, qual_len, oq_off, seq_len (NULL = never 0)      out  channels (channel q at the sum of count[0..q)), count, monochars
 * demux: codec_oq_reconstruct (:126-164) for every line of a VBlock at once: the reconstructed QUAL strings are the keys (key_bias = 33
 *        when they are text, 0 when they are BAM values, :131), a monochar channel yields its character and consumes nothing.
 *            in  txt, qual_off, qual_len, key_bias, channels + count (the channels that exist, back to back; count 0 for the others), monochars, out_off
 *            out out (qual_len[i] bytes at out_off[i] per line)
 * GZB_E_CORRUPT: a QUAL character outside '!'..'~', or a channel out of data (:152).  Device pointers with GZB_DEVICE_PTRS (then
 * channels_cap / out_cap bound the work); with host buffers all of `out` is written (0 where no line lands). */
typedef struct gzb_oq_vb {
    const void     *txt;        uint64_t txt_len;
    const uint64_t *qual_off;
    const uint32_t *qual_len;
    const uint64_t *oq_off;     /* mux */
    const uint32_t *seq_len;    /* mux, may be NULL */
    uint32_t        n_lines;
    int32_t         status;
    uint32_t        key_bias;   /* demux */
    uint32_t        reserved;
    void           *channels;   uint64_t channels_cap;
    uint32_t        count[94];
    uint8_t         monochars[94];
    uint8_t         pad[2];
    void           *out;        uint64_t out_cap;   /* demux */
    const uint64_t *out_off;    /* demux */
} gzb_oq_vb;
int gzb_oq_mux   (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_oq_demux (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- SMUX (src/codec_smux.c): MGI's quality codec — QUAL multiplexed by the base
 * mux:   codec_smux_compress (:180-262): QUAL[i] goes to the channel of SEQ[i] — A, C, G, T, anything else (_nuke_encode, src/reference.c:78-84):
 *        5 channels; a reverse-complemented read (is_rev) is walked from its last position with complemented bases (:236-239); a read
 *        without quality (qual_len 1, ' ') puts its ' ' in the channel of its first base, of its LAST base when reversed (:223-232).
 *        count[] = bytes per channel; n_param = the character of the fifth channel when it holds one character only, else 0 (the
 *        reference then drops that channel and sends the character in the section header's param, :246-253).
 *            in  txt, qual_off, qual_len, seq_off, seq_len, is_rev (NULL = none)      out  channels (channel b at the sum of count[0..b)), count, n_param
 * demux: codec_smux_reconstruct (:273-355) for every line of a VBlock at once (output in the read's own orientation: SAM / BAM / FASTQ out of
 *        FASTQ; not its SAM-to-FASTQ translation, :328-330): qual_len[i] = the `len` of the i-th call (1 when SEQ is "*", :278-279).
 *        A read without quality consumes ONE byte whatever its length — known only from the byte itself — so a VBlock whose channels hold a
 *        ' ' is refused with GZB_E_UNSUPPORTED (reconstruct it line by line); FASTQ never has one.
 *            in  txt + seq_off, qual_len, is_rev, channels + count (those that exist, back to back), n_param, out_off      out  out
 * GZB_E_CORRUPT: a channel out of data (:301).  Device pointers with GZB_DEVICE_PTRS (then channels_cap / out_cap bound the work). */
typedef struct gzb_smux_vb {
    const void     *txt;        uint64_t txt_len;
    const uint64_t *qual_off;   /* mux */
    const uint32_t *qual_len;
    const uint64_t *seq_off;
    const uint32_t *seq_len;    /* mux */
    const uint8_t  *is_rev;     /* may be NULL */
    uint32_t        n_lines;
    int32_t         status;
    void           *channels;   uint64_t channels_cap;
    uint32_t        count[5];
    uint8_t         n_param;
    uint8_t         pad[3];
    void           *out;        uint64_t out_cap;   /* demux */
    const uint64_t *out_off;    /* demux */
} gzb_smux_vb;
int gzb_smux_mux   (gzb_engine *e, gzb_smux_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_smux_demux (gzb_engine *e, gzb_smux_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- TMPL (src/codec_tmpl.c): Element's quality codec — QUAL multiplexed by a template
 * The template (found once per component in segconf, codec_tmpl_segconf_finalize :22-107: the most frequent quality of every read position)
 * is the caller's: tmpl / tmpl_len, host memory.
 * mux:   codec_tmpl_compress (:145-210): QUAL[i] goes to channel tmpl[i] - '!' for i < tmpl_len; what a read has beyond the template goes,
 *        in line order, to the excess stream ((ctx+1)->local, :180-182) — channel 94 here.
 *            in  txt, qual_off, qual_len, tmpl      out  channels (channel q at the sum of count[0..q)), count[95]
 * demux: codec_tmpl_reconstruct (:216-259) for every line of a VBlock at once: qual_len[i] = the `len` of the i-th call.
 *            in  qual_len, tmpl, channels + count, out_off      out  out
 * GZB_E_CORRUPT: a channel out of data.  Device pointers with GZB_DEVICE_PTRS (then channels_cap / out_cap bound the work). */
typedef struct gzb_tmpl_vb {
    const void     *txt;        uint64_t txt_len;   /* mux */
    const uint64_t *qual_off;   /* mux */
    const uint32_t *qual_len;
    uint32_t        n_lines;
    int32_t         status;
    const void     *tmpl;       uint32_t tmpl_len;  uint32_t reserved;
    void           *channels;   uint64_t channels_cap;
    uint32_t        count[95];  uint32_t pad;
    void           *out;        uint64_t out_cap;   /* demux */
    const uint64_t *out_off;    /* demux */
} gzb_tmpl_vb;
int gzb_tmpl_mux   (gzb_engine *e, gzb_tmpl_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_tmpl_demux (gzb_engine *e, gzb_tmpl_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- PACB (src/codec_pacb.c): PacBio's quality codec — QUAL multiplexed by passes and surroundings
 * mux:   codec_pacb_compress (:164-236): QUAL[i] goes to channel 7 * np0 + K, np0[line] = min (np:i, max_np) - 1 (0 for FASTQ and CLR data:
 *        max_np = 1, np0 may be NULL) and K = QUAL_get_K_value (:19-27), one of 7 classes of SEQ around position i; 7 * max_np channels
 *        (max_np <= 12).  A read without quality (qual_len 1, ' ') is one element of a one-base sequence (:135).
 *            in  txt, qual_off, qual_len, seq_off, np0, max_np      out  channels (channel c at the sum of count[0..c)), count[84]
 * demux: codec_pacb_reconstruct (:262-326) for every line of a VBlock at once (the read's own orientation, not the SAM-to-FASTQ translation
 *        :296-303): qual_len[i] = the `len` of the i-th call.  As with SMUX, a VBlock whose channels hold a ' ' (a read without quality
 *        consumes one byte whatever its length, :313-317) is refused with GZB_E_UNSUPPORTED.
 *            in  txt + seq_off, qual_len, np0, max_np, channels + count, out_off      out  out
 * GZB_E_CORRUPT: a channel out of data (:309).  Device pointers with GZB_DEVICE_PTRS (then channels_cap / out_cap bound the work). */
typedef struct gzb_pacb_vb {
    const void     *txt;        uint64_t txt_len;
    const uint64_t *qual_off;   /* mux */
    const uint32_t *qual_len;
    const uint64_t *seq_off;
    const uint8_t  *np0;        /* may be NULL when max_np is 1 */
    uint32_t        n_lines;
    int32_t         status;
    uint32_t        max_np;     uint32_t reserved;
    void           *channels;   uint64_t channels_cap;
    uint32_t        count[84];
    void           *out;        uint64_t out_cap;   /* demux */
    const uint64_t *out_off;    /* demux */
} gzb_pacb_vb;
int gzb_pacb_mux   (gzb_engine *e, gzb_pacb_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_pacb_demux (gzb_engine *e, gzb_pacb_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- PBWT (src/codec_pbwt.c)
 * encode: codec_pbwt_compress (:244-287): haplotype matrix → RUNS (uint32) + FGRC ({allele:8,count:24}; the last
 *         two words are the 64-bit matrix length, :274-276).  Host-endian words.
 * decode: codec_pbwt_uncompress (:372-402): RUNS + FGRC (host-endian) → matrix. */
int gzb_pbwt_encode (gzb_engine *e, const void *ht, uint32_t n_lines, uint32_t ht_per_line,
                     uint32_t *runs, uint32_t runs_cap, uint32_t *n_runs,
                     uint32_t *fgrc, uint32_t fgrc_cap, uint32_t *n_fgrc, uint32_t flags);
int gzb_pbwt_decode (gzb_engine *e, const uint32_t *runs, uint32_t n_runs, const uint32_t *fgrc, uint32_t n_fgrc,
                     uint32_t n_lines, void *ht, uint64_t ht_cap, uint64_t *ht_len, uint32_t flags);

/* Batched forms — the matrices of all VBlocks of a batch in one call: one CTA per VBlock walks its rows, so a batch of >= 148
 * VBlocks (one per SM; several fit) is what fills the GPU.  The per-VBlock outcome is in `status`; the call returns the first
 * non-zero one.
 * encode: in  ht, n_lines, ht_per_line, runs/runs_cap, fgrc/fgrc_cap      out  n_runs, n_fgrc, status
 * decode: in  runs/n_runs, fgrc/n_fgrc, n_lines, ht/ht_cap                 out  ht bytes, ht_len, status */
typedef struct gzb_pbwt_vb {
    void     *ht;           /* the haplotype matrix, n_lines x ht_per_line alleles (encode: read; decode: written) */
    uint64_t  ht_cap;       /* decode: capacity of ht */
    uint64_t  ht_len;       /* decode out: bytes of matrix = the 64-bit length in the last two FGRC words */
    uint32_t  n_lines;      /* ht_ctx->HT_n_lines */
    uint32_t  ht_per_line;  /* encode in */
    uint32_t *runs;  uint32_t runs_cap, n_runs;
    uint32_t *fgrc;  uint32_t fgrc_cap, n_fgrc;
    int32_t   status;
    uint32_t  reserved;
} gzb_pbwt_vb;
int gzb_pbwt_encode_batch (gzb_engine *e, gzb_pbwt_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_pbwt_decode_batch (gzb_engine *e, gzb_pbwt_vb *vbs, uint32_t n_vbs, uint32_t flags);

/* ---------------------------------------------------------------- LONGR (src/codec_longr.c, src/codec_longr_alg.c)
 * encode: codec_longr_compress (:161-247) before its sub-codec: channel per base, stable sort of quals by channel.
 * decode: codec_longr_reconstruct (:342-373) for all reads of a VB. */
typedef struct {
    const void     *txt;        /* buffer holding SEQ and QUAL strings */
    uint64_t        txt_len;
    const uint64_t *seq_off;    /* n_lines */
    const uint64_t *qual_off;   /* n_lines (encode only) */
    const uint32_t *len;        /* n_lines (seq_len == qual_len) */
    const uint8_t  *is_rev;     /* n_lines or NULL */
    uint32_t        n_lines;
    uint8_t         value_to_bin[256];  /* codec_longr_segconf_calculate_bins (:66-136), computed once by the host */
    void           *values;     /* sum(len) bytes: encode out / decode in */
    uint32_t       *lens_be;    /* 65536 big-endian u32: encode out / decode in (:237-240) */
    void           *qual_out;   /* decode: concatenated quality strings, len[i] bytes per line */
    uint8_t        *missing;    /* decode, optional (n_lines): 1 where the line has no quality — its first value is 255 (:278); the line's
                                   first byte is then '*' (sam_reconstruct_missing_quality, src/sam_qual.c:532) and the rest undefined */
    const uint32_t *qual_len;   /* encode, optional (n_lines): quality length where it differs from len — a SAM line without quality is the
                                   one byte ' ' whatever its seq_len (:188,:190-192); NULL = len */
    uint64_t        n_bases;    /* the number of qualities in the VBlock = bytes of `values`.  Host pointers: 0 = sum(len); decode of a VBlock that
                                   has lines without quality passes the length of VALUES.local.  GZB_DEVICE_PTRS: 0 = take txt_len as the bound */
} gzb_longr_vb;
int gzb_longr_encode (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags);
int gzb_longr_decode (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags);
/* codec_longr_segconf_calculate_bins (:66-136): value_to_bin of the 32 equal-population bins from the qualities of one VBlock
 * (histogram on the GPU).  GZB_SOFT_FAIL when there is no quality at all (flag.no_longr, :100-103). */
int gzb_longr_calculate_bins (gzb_engine *e, gzb_longr_vb *vb, uint32_t flags, uint8_t value_to_bin[256]);

/* ================================================================ plug-in layer (reference signatures) */
typedef struct VBlock        *VBlockP;          /* opaque genozip types */
typedef struct Context       *ContextP;
typedef union  SectionHeaderUnion *SectionHeaderP;
typedef struct Buffer        *BufferP;
typedef uint8_t Codec;                           /* src/genozip.h: packed enum */
typedef enum { HARD_FAIL = 0, SOFT_FAIL = 1 } FailType;   /* src/genozip.h */
typedef void LocalGetLineCB (VBlockP vb, ContextP ctx, uint32_t vb_line_i, char **line_data, uint32_t *line_data_len,
                             uint32_t maximum_size, bool *is_rev);                                /* src/genozip.h:673-680 */

/* what the ≤200-line adapter inside genozip registers once: the only code that knows genozip's struct layouts */
typedef struct {
    uint32_t (*vb_num_lines)  (VBlockP vb);                          /* vb->lines.len32 */
    uint32_t (*vb_vblock_i)   (VBlockP vb);                          /* vb->vblock_i */
    char    *(*buffer_data)   (BufferP buf);                         /* buf->data (pre-allocated by the caller, src/zfile.c:229) */
    void     (*abort_msg)     (const char *msg);                     /* ABORT → error_assert_failed (src/error.c:420); NULL = abort() */
} gzb_plugin_host;
void gzb_plugin_register (const gzb_plugin_host *host, int n_devices);

#define GZB_COMPRESS(f) bool f (VBlockP vb, ContextP ctx, SectionHeaderP header, const char *uncompressed, \
    uint32_t *uncompressed_len, LocalGetLineCB get_line_cb, char *compressed, uint32_t *compressed_len, \
    FailType soft_fail, const char *name)                                                          /* src/codec.h:17-27 */
#define GZB_UNCOMPRESS(f) void f (VBlockP vb, ContextP ctx, Codec codec, uint8_t param, const char *compressed, \
    uint32_t compressed_len, BufferP uncompressed_buf, uint64_t uncompressed_len, Codec sub_codec, const char *name) /* src/codec.h:29-38 */

GZB_COMPRESS (gzb_codec_RANB_compress);  GZB_COMPRESS (gzb_codec_RANW_compress);
GZB_COMPRESS (gzb_codec_RANb_compress);  GZB_COMPRESS (gzb_codec_RANw_compress);
GZB_COMPRESS (gzb_codec_ARTB_compress);  GZB_COMPRESS (gzb_codec_ARTW_compress);
GZB_COMPRESS (gzb_codec_ARTb_compress);  GZB_COMPRESS (gzb_codec_ARTw_compress);
GZB_UNCOMPRESS (gzb_codec_rans_uncompress);
GZB_UNCOMPRESS (gzb_codec_arith_uncompress);
uint32_t gzb_codec_RANB_est_size (Codec codec, uint64_t uncompressed_len);   /* src/codec.h:40 */
uint32_t gzb_codec_RANW_est_size (Codec codec, uint64_t uncompressed_len);
uint32_t gzb_codec_RANb_est_size (Codec codec, uint64_t uncompressed_len);
uint32_t gzb_codec_RANw_est_size (Codec codec, uint64_t uncompressed_len);
uint32_t gzb_codec_ARTB_est_size (Codec codec, uint64_t uncompressed_len);
uint32_t gzb_codec_ARTW_est_size (Codec codec, uint64_t uncompressed_len);
uint32_t gzb_codec_ARTb_est_size (Codec codec, uint64_t uncompressed_len);
uint32_t gzb_codec_ARTw_est_size (Codec codec, uint64_t uncompressed_len);


/* ---------------------------------------------------------------- the complex codecs behind the reference's signatures
 * Table rows src/codec.h:99-101,108-109:  DOMQ { codec_domq_compress, USE_SUBCODEC, codec_domq_reconstruct },
 * PBWT { codec_pbwt_compress, codec_pbwt_uncompress, codec_pbwt_reconstruct }, LONGR { codec_longr_compress, USE_SUBCODEC,
 * codec_longr_reconstruct }, ACGT { codec_acgt_compress, codec_acgt_uncompress }, XCGT { USE_SUBCODEC, codec_xcgt_uncompress }.
 * These reach into sibling contexts, the section header, vb->scratch and the codec table, so the adapter inside
 * genozip registers a second accessor table.  `sibling` counts contexts from the one the codec was called with, as the
 * reference's `ctx + 1` does (DOMQ: 0 QUAL, 1 DOMQRUNS, 2 QUALMPLX, 3 DIVRQUAL — declare_domq_contexts, src/codec_domq.c:34-38;
 * ACGT: 0 NONREF, 1 NONREF_X; LONGR: 0 the lengths, 1 the values; PBWT: 0 the matrix, 1 RUNS, 2 FGRC — decl_pbwt_contexts). */
typedef void CodecReconstructFn (VBlockP vb, Codec codec, ContextP ctx, uint32_t len, bool reconstruct);   /* src/codec.h:42-43 */
enum { GZB_HDR_SUB_CODEC = 1, GZB_HDR_ACGT_NO_X = 2 };
typedef struct {
    /* ---- ZIP ---- */
    char    *(*local_alloc)     (VBlockP vb, ContextP ctx, int sibling, uint64_t bytes);    /* buf_alloc (vb, &(ctx+sibling)->local, …) → .data (contents kept) */
    char    *(*local_data)      (ContextP ctx, int sibling, uint64_t *len_bytes);           /* (ctx+sibling)->local.data and its length in bytes */
    void     (*local_set_len)   (ContextP ctx, int sibling, uint64_t len_bytes);
    void     (*local_free)      (VBlockP vb, ContextP ctx, int sibling);                    /* buf_free / buf_destroy */
    uint8_t *(*local_prm8)      (ContextP ctx, int sibling);                                /* &(ctx+sibling)->local.prm8[0] */
    char    *(*scratch_alloc)   (VBlockP vb, uint64_t bytes);                               /* vb->scratch: buf_alloc → .data; bytes = 0: the buffer as it stands */
    void     (*scratch_free)    (VBlockP vb);
    bool     (*ctx_acgt_no_x)   (ContextP ctx, int sibling);                                /* (ctx+sibling)->flags.acgt_no_x */
    void     (*header_set)      (SectionHeaderP header, int field, uint32_t value);         /* GZB_HDR_SUB_CODEC: header->sub_codec; GZB_HDR_ACGT_NO_X: header->flags.ctx.acgt_no_x */
    Codec    (*assign_sub_codec)(VBlockP vb, ContextP ctx, int sibling);                    /* codec_assign_best_codec (vb, ctx+sibling, &local, SEC_LOCAL); never CODEC_UNKNOWN (→ CODEC_NONE);
                                                                                               for ACGT's sibling 1 the adapter also sets lsubcodec_piz / lcodec = XCGT (src/codec_acgt.c:143-152) */
    bool     (*sub_compress)    (Codec c, VBlockP vb, ContextP ctx, SectionHeaderP header, const char *data, uint32_t *len,
                                 char *compressed, uint32_t *compressed_len, FailType soft_fail, const char *name);   /* codec_args[c].compress */
    uint32_t (*sub_est_size)    (Codec c, uint64_t len);                                    /* codec_args[c].est_size */
    void     (*seg_denorm)      (VBlockP vb, ContextP qual_ctx, const uint8_t *denorm, uint32_t len);   /* base64 + seg_by_ctx into DOMQRUNS (src/codec_domq.c:241-244) */
    bool     (*seq_line)        (VBlockP vb, ContextP ctx, uint32_t vb_line_i, char **seq, uint32_t *seq_len, bool *is_rev);   /* fastq_zip_seq / sam_zip_seq (src/codec_longr.c:170) */
    const uint8_t *(*codec_table) (VBlockP vb, ContextP ctx);                              /* the codec's small table, for the ctx the codec was called with.  LONGR: the 256-byte value-to-bin map of ctx+1
                                                                                               (ZCTX(values)->value_to_bin in ZIP, from SEC_COUNTS in PIZ, src/codec_longr.c:316-330).  DOMQ reconstruct: the
                                                                                               de-normalisation table from DOMQRUNS' dictionary: byte 0 = number of doms, then [num_doms][num_norm_qs] */
    void     (*pbwt_dims)       (VBlockP vb, ContextP ht_ctx, uint32_t *n_lines, uint32_t *ht_per_line, int set);   /* ht_ctx->HT_n_lines, ->ht_per_line (get; set != 0: store ht_per_line) */
    void     (*add_lines)       (int which, uint64_t n);                                    /* z_file->domq_lines / longr_lines (src/codec_domq.c:489-492, src/codec_longr.c:181): 0 DOMQ dom, 1 DOMQ diverse, 2 LONGR, 3 NORMQ (z_file->normq_lines, src/codec_normq.c:41), 4 HOMP (z_file->homp_lines, src/codec_homp.c:128-129) */
    void     (*account_time)    (VBlockP vb, int which, uint64_t nanosec);                  /* COPY_TIMER (compressor_domq …), src/profiler.h:18-24: 0 domq 1 acgt 2 xcgt 3 pbwt 4 longr 5 normq 6 homp 7 t0 */
    /* ---- PIZ ---- */
    void     (*sub_uncompress)  (Codec c, VBlockP vb, ContextP ctx, uint8_t param, const char *compressed, uint32_t compressed_len,
                                 BufferP uncompressed_buf, uint64_t uncompressed_len, const char *name);               /* codec_args[c].uncompress */
    BufferP  (*packed_buffer)   (VBlockP vb, ContextP ctx, int sibling, uint64_t bytes);   /* (ctx+sibling)->flags.acgt_no_x ? &vb->scratch : &(ctx+sibling)->packed, allocated to `bytes` when > 0 */
    void    **(*codec_state)    (VBlockP vb, ContextP ctx);                                 /* one pointer slot that lives with the VBlock's context (like lens_ctx->longr_state): the staging of a bulk decode */
    const uint32_t *(*recon_line_lens) (VBlockP vb, ContextP ctx, uint32_t *n_lines);      /* the `len` of every future reconstruct call of this context in this VBlock, in order */
    bool     (*recon_seq_table) (VBlockP vb, ContextP ctx, const char **txt, uint64_t *txt_len, const uint64_t **seq_off, const uint8_t **is_rev);   /* LONGR: every read's SEQ up front (bulk SEQ reconstruct) */
    char    *(*recon_at)        (VBlockP vb);                                               /* BAFTtxt */
    void     (*recon_advance)   (VBlockP vb, int32_t n);                                    /* Ltxt += n (n < 0: remove characters) */
    int64_t  (*pbwt_big_allele) (VBlockP vb);                                               /* reconstruct_from_local_int (vb, CTX(FORMAT_GT_HT_BIG), 0, RECON_OFF) */
    bool     (*drop_curr_line)  (VBlockP vb);
    void     (*missing_quality) (VBlockP vb, bool reconstruct);                             /* sam_reconstruct_missing_quality */
    /* ---- added for HOMP / T0 (may be NULL when those entry points are not used) ---- */
    void     (*update_line_len) (VBlockP vb, ContextP ctx, uint32_t vb_line_i, uint32_t new_len);   /* sam_update_qual_len / fastq_update_qual_len / sam_ultima_update_t0_len (src/codec_homp.c:183, src/codec_t0.c:103) */
} gzb_plugin_host2;
void gzb_plugin_register2 (const gzb_plugin_host2 *host);
void gzb_plugin_shutdown (void);                                   /* destroys the pooled engines (process exit / plug-in unregistered) */

/* ZIP */
bool gzb_codec_domq_comp_init (VBlockP vb, ContextP qual_ctx, LocalGetLineCB get_line_cb, bool force);   /* codec_domq_comp_init (src/codec_domq.c:299-323) */
GZB_COMPRESS (gzb_codec_domq_compress);                            /* src/codec_domq.c:379-521 */
GZB_COMPRESS (gzb_codec_acgt_compress);                            /* src/codec_acgt.c:64-176 */
GZB_COMPRESS (gzb_codec_pbwt_compress);                            /* src/codec_pbwt.c:244-287 */
GZB_COMPRESS (gzb_codec_longr_compress);                           /* src/codec_longr.c:161-264 */
GZB_COMPRESS (gzb_codec_normq_compress);                           /* src/codec_normq.c:31-82 (table row src/codec.h:102) */
GZB_COMPRESS (gzb_codec_homp_compress);                            /* src/codec_homp.c:121-205: the lines condensed in place, their lengths updated, then the sub-codec */
GZB_COMPRESS (gzb_codec_t0_compress);                              /* src/codec_t0.c:58-124 */
uint32_t gzb_codec_complex_est_size (Codec codec, uint64_t uncompressed_len);    /* src/codec.c codec_complex_est_size */
uint32_t gzb_codec_longr_est_size   (Codec codec, uint64_t uncompressed_len);    /* src/codec_longr.c:53-56 */
/* PIZ */
GZB_UNCOMPRESS (gzb_codec_acgt_uncompress);                        /* src/codec_acgt.c:216-248 */
GZB_UNCOMPRESS (gzb_codec_xcgt_uncompress);                        /* src/codec_acgt.c:185-209 */
GZB_UNCOMPRESS (gzb_codec_pbwt_uncompress);                        /* src/codec_pbwt.c:372-402 */
CodecReconstructFn gzb_codec_domq_reconstruct;                     /* src/codec_domq.c:774-809 */
CodecReconstructFn gzb_codec_pbwt_reconstruct;                     /* src/codec_pbwt.c:406-449 */
CodecReconstructFn gzb_codec_longr_reconstruct;                    /* src/codec_longr.c:342-373 */
CodecReconstructFn gzb_codec_homp_reconstruct;                     /* src/codec_homp.c:213-276; recon_seq_table supplies every line's SEQ (deep / SAM-to-FASTQ variants :222-234 not covered) */
CodecReconstructFn gzb_codec_t0_reconstruct;                       /* src/codec_t0.c:137-179 */
CodecReconstructFn gzb_codec_normq_reconstruct;                    /* src/codec_normq.c:85-106; recon_seq_table supplies every line's strand (last_flags.rev_comp) */

/* ---------------------------------------------------------------- combining submission
 * comp_compress calls a codec once per section from every compute thread (src/compressor.c:82-86); one launch per section is
 * launch-latency-bound.  A combiner gathers the sections that several compute threads submit at about the same time into ONE
 * gzb_compress_sections / gzb_uncompress_sections call: every thread submits its section and waits; the first waiter becomes the
 * leader, takes everything that is pending (after giving the others `linger_us` to arrive) and runs the batch for all of them.
 * The plug-in entry points above go through the process-wide combiners when gzb_plugin_set_combining (1) was called. */
typedef struct gzb_combiner gzb_combiner;
gzb_combiner *gzb_combiner_create (int device, int compress /* 1 = compress, 0 = uncompress */, uint32_t linger_us);
void gzb_combiner_destroy (gzb_combiner *c);
int  gzb_submit (gzb_combiner *c, const gzb_section *sec, uint64_t *ticket);         /* host pointers; thread-safe */
int  gzb_wait   (gzb_combiner *c, uint64_t ticket, gzb_section *result);             /* status / out_len of the section; thread-safe */
uint64_t gzb_combiner_batches (gzb_combiner *c);                                      /* batches run so far (sections / batches = the combining factor) */
void gzb_plugin_set_combining (int on, uint32_t linger_us);

#ifdef __cplusplus
}
#endif
#endif
