/* oracle/oracle.h — CPU restatement of genozip's per-VBlock codec path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libgzb200.so) never links, loads or
 * calls anything in oracle/ and has no CPU fallback.
 *
 * Parity status:
 *   - rANS 4x16 / adaptive arithmetic / PACK / STRIPE / CAT (hts_port.c): PINNED — every function is
 *     differential-tested byte-for-byte against oracle/_ref/libhts_ref.so, which is the reference's own
 *     htscodecs translation units compiled unmodified with the reference's flags (oracle/Makefile).
 *   - DOMQ / ACGT / PBWT / LONGR (gz_port.c): PINNED (encoders and decoders) — differential-tested byte-for-byte against
 *     oracle/_ref/libgz_ref.so: the reference's own codec_domq.c, codec_acgt.c, codec_pbwt.c, codec_longr.c compiled
 *     unmodified with the reference's flags and hosted outside the (licence-gated) program by oracle/ref_gz_shim.c,
 *     which is compiled against the reference's headers, hand-makes the VBlock / Contexts a compute thread would
 *     pass and supplies the ~40 host symbols those objects need; the nucleotide tables are the .rodata of the
 *     reference's compiled reference.c (tests/test_oracle_gz_ref.py).  The decoders (restated from the reference's
 *     PIZ code, cited separately) are compared with the reference's own codec_acgt_uncompress / codec_xcgt_uncompress,
 *     codec_pbwt_uncompress, codec_domq_reconstruct and codec_longr_reconstruct hosted the same way, and must invert
 *     the pinned encoders.
 *     Whole-file .genozip identity: parity unpinned (closed licence.o).
 *
 * All citations are relative to /root/reference/src.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* order bytes used by genozip (codec_htscodecs.c:17-20) */
#define ORC_ORDER_B 0x01
#define ORC_ORDER_W 0x19
#define ORC_ORDER_b 0x81
#define ORC_ORDER_w 0x99

/* ---- hts_port.c : htscodecs restatement ---- */
uint32_t orc_rans_bound   (uint32_t n, int order);                         /* rANS_static4x16pr.c:357-369 */
uint32_t orc_arith_bound  (uint32_t n, int order);                         /* arith_dynamic.c:74-80 */
/* return 0 on success, -1 on failure. *out_len: in = capacity, out = bytes written */
int orc_rans_compress     (const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_len, int order);   /* :1151-1356 */
int orc_rans_uncompress   (const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t *out_len);         /* :1358-1642 */
int orc_arith_compress    (const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_len, int order);   /* arith_dynamic.c:615-858 */
int orc_arith_uncompress  (const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t *out_len);         /* arith_dynamic.c:860-1104 */

/* ---- gz_port.c : genozip-specific codecs ---- */

/* ACGT (codec_acgt.c:45-55,64-177).  packed must hold orc_acgt_packed_len(n) bytes; x holds n bytes.
 * returns 1 if the exception stream x is all-zero (=> acgt_no_x), else 0 */
uint64_t orc_acgt_packed_len (uint64_t n_bases);
int  orc_acgt_pack   (const uint8_t *seq, uint64_t n, uint8_t *packed, uint8_t *x);
/* codec_acgt.c:185-248; x may be NULL (acgt_no_x) */
void orc_acgt_unpack (const uint8_t *packed, const uint8_t *x, uint64_t n, uint8_t *seq);

/* DOMQ (codec_domq.c).  Lines are given as (offset,len) into txt. */
typedef struct {
    uint32_t n_lines;
    uint8_t  num_norm_qs;      /* == no_doms marker; header param is num_norm_qs|0x80 (:234) */
    uint8_t  num_doms;
    uint8_t  has_diverse;
    uint8_t  denorm[95*95];    /* [num_doms][num_norm_qs] ASCII (:232-239) */
    uint8_t  normalize[95*95]; /* [cdom*95 + (q-32)] -> rank (:225) */
} OrcDomqTables;

/* codec_domq_prepare_normalize (:252-293): fills line_dom[] (compacted dom), line_diverse[], tables */
void orc_domq_prepare (const uint8_t *txt, const uint64_t *line_off, const uint32_t *line_len, uint32_t n_lines,
                       uint8_t *line_dom, uint8_t *line_diverse, OrcDomqTables *t);
/* codec_domq_compress (:379-521) up to (not including) the sub-codec: the four streams.
 * Buffers must be large enough: qual 2*total+1, runs total+1 (worst cases), mplx n_lines, divr total. */
void orc_domq_split (const uint8_t *txt, const uint64_t *line_off, const uint32_t *line_len, uint32_t n_lines,
                     const uint8_t *line_dom, const uint8_t *line_diverse, const OrcDomqTables *t,
                     uint8_t *qual, uint32_t *qual_len, uint8_t *runs, uint32_t *runs_len,
                     uint8_t *mplx, uint32_t *mplx_len, uint8_t *divr, uint32_t *divr_len);
/* codec_domq_reconstruct (:774-809) line by line, following the reference decoder literally
 * (including the in-place run shortening :529-548).  runs is modified.  returns 0 ok. */
int  orc_domq_reconstruct (const uint8_t *qual, uint32_t qual_len, uint8_t *runs, uint32_t runs_len,
                           const uint8_t *mplx, uint32_t mplx_len, const uint8_t *divr, uint32_t divr_len,
                           const uint8_t *denorm, uint8_t num_norm_qs,
                           const uint32_t *line_len, uint32_t n_lines, uint8_t *out);

/* PBWT (codec_pbwt.c).  runs/fgrc are host-endian uint32 arrays; caps in words. */
int  orc_pbwt_encode (const uint8_t *ht, uint32_t n_lines, uint32_t ht_per_line,
                      uint32_t *runs, uint32_t *n_runs, uint32_t *fgrc, uint32_t *n_fgrc);   /* :244-287 */
int  orc_pbwt_decode (uint32_t *runs, uint32_t n_runs, uint32_t *fgrc, uint32_t n_fgrc,
                      uint32_t n_lines, uint8_t *ht, uint64_t *ht_len);                      /* :317-402 */

/* LONGR (codec_longr.c, codec_longr_alg.c) */
void orc_longr_calc_bins (const uint32_t histogram[256], uint64_t num_values, uint8_t value_to_bin[256]); /* codec_longr.c:66-136 */
/* :161-247: lens_be = 65536 big-endian u32; values = n_quals bytes; is_rev may be NULL */
int  orc_longr_decode2 (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev, uint32_t n_lines,
                        const uint8_t v2b[256], const uint8_t *values, const uint32_t *lens_be, uint8_t *qual_out, uint8_t *missing);
int  orc_longr_encode2 (const uint8_t *txt, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *len, const uint32_t *seq_len,
                        const uint8_t *is_rev, uint32_t n_lines, const uint8_t v2b[256], uint8_t *values, uint32_t *lens_be);
int  orc_longr_encode (const uint8_t *txt, const uint64_t *seq_off, const uint64_t *qual_off, const uint32_t *len,
                       const uint8_t *is_rev, uint32_t n_lines, const uint8_t value_to_bin[256],
                       uint8_t *values, uint32_t *lens_be);
/* :270-373 */
int  orc_longr_decode (const uint8_t *txt, const uint64_t *seq_off, const uint32_t *len, const uint8_t *is_rev,
                       uint32_t n_lines, const uint8_t value_to_bin[256],
                       const uint8_t *values, const uint32_t *lens_be, uint8_t *qual_out /* concatenated */);

#ifdef __cplusplus
}

/* NORMQ (src/codec_normq.c :43-62, :85-106) */
uint64_t orc_normq_encode (const uint8_t *txt, const uint64_t *line_off, const uint32_t *line_len, const uint8_t *is_rev, uint32_t n_lines, uint8_t *local);
int  orc_normq_decode (const uint8_t *local, uint64_t local_len, const uint32_t *len, const uint8_t *is_rev, uint32_t n_lines, uint8_t *out, uint8_t *missing, uint64_t *used);

#endif
