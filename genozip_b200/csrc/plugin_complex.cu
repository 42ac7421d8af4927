// plugin_complex.cu — the complex codecs behind the reference's own signatures (src/codec.h table rows :99-101, :108-109):
//   ZIP  codec_domq_comp_init / codec_domq_compress (src/codec_domq.c:299-323, :379-521), codec_acgt_compress (src/codec_acgt.c:64-176),
//        codec_pbwt_compress (src/codec_pbwt.c:244-287), codec_longr_compress (src/codec_longr.c:161-264)
//   PIZ  codec_acgt_uncompress / codec_xcgt_uncompress (src/codec_acgt.c:185-248), codec_pbwt_uncompress / _reconstruct
//        (src/codec_pbwt.c:372-449), codec_domq_reconstruct (src/codec_domq.c:774-809), codec_longr_reconstruct (src/codec_longr.c:342-373)
// They are adapters over the flat entry points: everything that touches genozip's VBlock / Context / Buffer / SectionHeader goes
// through the accessor table gzb_plugin_host2 the adapter inside genozip registers.  The transforms run on the GPU; what stays
// here is what the reference's functions do around them — buffer allocation in the sibling contexts, the header fields, the
// choice and the call of the sub-codec, the soft-fail re-entry — in the reference's order.
//
// The per-line reconstructors decode the whole VBlock on their FIRST call (the bulk kernels) and hand out one line per call
// afterwards: that needs, up front, what the reference learns line by line — the line lengths (recon_line_lens)
// and, for LONGR, every read's SEQ (recon_seq_table) — which the adapter supplies from its bulk SEQ reconstruction.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <algorithm>
#include "plugin_internal.h"

using namespace gzbp;

namespace {

constexpr Codec CODEC_NONE_ = 1, CODEC_LZMA_ = 4;                  // src/genozip.h:326, :342
constexpr uint32_t MIN_LEN_FOR_COMPRESSION = 50;                   // src/codec.h:15

struct Timer {                                                     // COPY_TIMER (compressor_*): the adapter adds it to vb->profile
    VBlockP vb; int which; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now ();
    ~Timer () {
        if (g_host2.account_time)
            g_host2.account_time (vb, which, (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds> (std::chrono::steady_clock::now () - t0).count ());
    }
};

#define NEED(fn, name) do { if (!g_host2.fn) plugin_abort ("adapter", name, "gzb_plugin_host2." #fn " is not registered"); } while (0)

// the VBlock's lines of a context as a table: the callback returns pointers into one text buffer (vb->txt_data); offsets are
// taken from the lowest pointer seen, so no accessor for the buffer itself is needed
struct Lines {
    std::vector<char *> ptr; std::vector<uint32_t> len; std::vector<uint8_t> rev;
    const char *base = nullptr; uint64_t span = 0, total = 0;
    std::vector<uint64_t> off;
    void finish () {
        const char *lo = nullptr, *hi = nullptr;
        for (size_t i = 0; i < ptr.size (); i++) if (len[i]) {
            if (!lo || ptr[i] < lo) lo = ptr[i];
            if (!hi || ptr[i] + len[i] > hi) hi = ptr[i] + len[i];
        }
        base = lo; span = lo ? (uint64_t)(hi - lo) : 0;
        off.resize (ptr.size ());
        for (size_t i = 0; i < ptr.size (); i++) off[i] = len[i] ? (uint64_t)(ptr[i] - lo) : 0;
    }
};

Lines gather_lines (VBlockP vb, ContextP ctx, LocalGetLineCB cb, const char *name)
{
    if (!g_host.vb_num_lines) plugin_abort ("line table", name, "adapter did not register vb_num_lines");
    const uint32_t n = g_host.vb_num_lines (vb);
    Lines L; L.ptr.resize (n); L.len.resize (n); L.rev.assign (n, 0);
    for (uint32_t i = 0; i < n; i++) {
        char *p = nullptr; uint32_t l = 0; bool r = false;
        cb (vb, ctx, i, &p, &l, 0xffffffffu, &r);
        L.ptr[i] = p; L.len[i] = l; L.rev[i] = r; L.total += l;
    }
    L.finish ();
    return L;
}

// the sub-codec call that ends every complex compressor (src/codec_acgt.c:157-168, src/codec_domq.c:503-520, src/codec_longr.c:251-263)
bool sub_compress (Codec sub, VBlockP vb, ContextP ctx, SectionHeaderP header, const char *data, uint32_t *len,
                   char *compressed, uint32_t *compressed_len, FailType soft_fail, const char *name)
{
    NEED (sub_compress, name);
    return g_host2.sub_compress (sub, vb, ctx, header, data, len, compressed, compressed_len, soft_fail, name);
}

// ---------------------------------------------------------------- DOMQ state between comp_init (seg) and compress
struct DomqState {
    gzb_domq_vb d; Lines lines; std::vector<uint8_t> dom, diverse; bool split_done = false;
};

} // namespace

// ================================================================ ACGT
extern "C" GZB_COMPRESS (gzb_codec_acgt_compress)
{
    NEED (local_alloc, name); NEED (local_data, name); NEED (scratch_alloc, name); NEED (scratch_free, name); NEED (header_set, name); NEED (ctx_acgt_no_x, name);
    Timer tm { vb, 1 };
    bool has_x = !g_host2.ctx_acgt_no_x (ctx, 0);
    uint64_t x_len = 0;
    const uint64_t n = *uncompressed_len, packed_len = gzb_acgt_packed_len (n);
    Codec sub;
    char *packed;
    if (has_x && g_host2.local_data (ctx, 1, &x_len) && x_len) {             // second entry, after soft-failing: continue from the sub-codec (:82-86)
        packed = g_host2.scratch_alloc (vb, 0);
        sub = packed_len >= MIN_LEN_FOR_COMPRESSION ? CODEC_LZMA_ : CODEC_NONE_;
    }
    else {
        packed = g_host2.scratch_alloc (vb, packed_len);
        std::vector<char> gathered;
        char *x = nullptr;
        if (uncompressed) {                                                  // option 1 (:95-111): the exceptions replace the bases IN PLACE; NONREF_X.local is laid over NONREF.local
            if (has_x) x = g_host2.local_alloc (vb, ctx, 1, n);              //   (the adapter overlays: the pointer it returns is `uncompressed` itself)
        }
        else if (get_line_cb) {                                              // option 2 (:114-134)
            if (!has_x) plugin_abort ("acgt", name, "ACGT compression with get_line_cb is only supported with has_x");
            Lines L = gather_lines (vb, ctx, get_line_cb, name);
            if (L.total != n) plugin_abort ("acgt", name, "total length from callbacks != uncompressed_len");
            gathered.resize (n);
            uint64_t o = 0;
            for (size_t i = 0; i < L.ptr.size (); i++) { if (L.len[i]) memcpy (gathered.data () + o, L.ptr[i], L.len[i]); o += L.len[i]; }
            uncompressed = gathered.data ();
            x = g_host2.local_alloc (vb, ctx, 1, n);
        }
        else plugin_abort ("acgt", name, "neither src_data nor callback is provided");
        int all_zero = 0;
        {
            EngineLease E (vb, name);
            if (gzb_acgt_pack (E.e, uncompressed, n, packed, x, &all_zero, 0) != GZB_OK) plugin_abort ("gzb_acgt_pack", name, gzb_last_error (E.e));
        }
        if (has_x) {
            if (all_zero) {                                                  // no exception bases after all (:138-142)
                has_x = false;
                g_host2.header_set (header, GZB_HDR_ACGT_NO_X, 1);
                if (g_host2.local_free) g_host2.local_free (vb, ctx, 1);
            }
            else {
                g_host2.local_set_len (ctx, 1, n);
                NEED (assign_sub_codec, name);
                g_host2.assign_sub_codec (vb, ctx, 1);                       // NONREF_X: lcodec = XCGT, its sub-codec as assigned (:145-155)
            }
        }
        sub = packed_len >= MIN_LEN_FOR_COMPRESSION ? CODEC_LZMA_ : CODEC_NONE_;   // the 2-bit words are little endian as they stand (:159-161)
        g_host2.header_set (header, GZB_HDR_SUB_CODEC, sub);
    }
    uint32_t plen = (uint32_t)packed_len;
    if (!sub_compress (sub, vb, ctx, header, packed, &plen, compressed, compressed_len, soft_fail, name)) return false;
    g_host2.scratch_free (vb);
    return true;
}

extern "C" GZB_UNCOMPRESS (gzb_codec_acgt_uncompress)
{
    (void)codec;
    NEED (packed_buffer, name); NEED (sub_uncompress, name); NEED (ctx_acgt_no_x, name); NEED (local_data, name);
    Timer tm { vb, 1 };
    const uint64_t bytes = gzb_acgt_packed_len (uncompressed_len);            // roundup_bits2bytes64 (2n): whole 64-bit words (:221)
    BufferP packed = g_host2.packed_buffer (vb, ctx, 0, bytes);
    g_host2.sub_uncompress (sub_codec, vb, ctx, param, compressed, compressed_len, packed, bytes, name);     // (:225)
    if (g_host2.ctx_acgt_no_x (ctx, 0)) {                                     // no NONREF_X section follows: decode here (:236-245)
        uint64_t l = 0;
        char *nonref = g_host2.local_data (ctx, 0, &l);
        EngineLease E (vb, name);
        if (gzb_acgt_unpack (E.e, g_host.buffer_data (packed), nullptr, uncompressed_len, nonref, 0) != GZB_OK) plugin_abort ("gzb_acgt_unpack", name, gzb_last_error (E.e));
        if (g_host2.scratch_free) g_host2.scratch_free (vb);
    }
    (void)uncompressed_buf;
}

extern "C" GZB_UNCOMPRESS (gzb_codec_xcgt_uncompress)
{
    (void)codec;
    NEED (packed_buffer, name); NEED (sub_uncompress, name); NEED (local_data, name);
    g_host2.sub_uncompress (sub_codec, vb, ctx, param, compressed, compressed_len, uncompressed_buf, uncompressed_len, name);   // NONREF_X through its own sub-codec (:191)
    Timer tm { vb, 2 };
    uint64_t l = 0;
    char *nonref = g_host2.local_data (ctx, -1, &l);                          // NONREF.local, allocated by the caller (:197)
    BufferP packed = g_host2.packed_buffer (vb, ctx, -1, 0);
    EngineLease E (vb, name);
    if (gzb_acgt_unpack (E.e, g_host.buffer_data (packed), g_host.buffer_data (uncompressed_buf), uncompressed_len, nonref, 0) != GZB_OK)
        plugin_abort ("gzb_acgt_unpack", name, gzb_last_error (E.e));
}

// ================================================================ DOMQ
// codec_domq_comp_init: called from seg_finalize.  The per-line histograms / doms run on the GPU (gzb_domq_prepare); the fit test
// (codec_domq_qual_data_is_a_fit_for_domq :80-131: more than half of the tested lines have a dominant score) is host policy on
// its result.  The de-normalisation table is segged by the adapter.
extern "C" bool gzb_codec_domq_comp_init (VBlockP vb, ContextP qual_ctx, LocalGetLineCB get_line_cb, bool force)
{
    const char *name = "QUAL";
    NEED (codec_state, name); NEED (local_prm8, name); NEED (seg_denorm, name);
    void **slot = g_host2.codec_state (vb, qual_ctx);
    DomqState *S = new DomqState ();
    S->lines = gather_lines (vb, qual_ctx, get_line_cb, name);
    const uint32_t n_lines = (uint32_t)S->lines.len.size ();
    S->dom.assign (n_lines + 1, 0); S->diverse.assign (n_lines + 1, 0);
    gzb_domq_vb &d = S->d; memset (&d, 0, sizeof d);
    d.txt = S->lines.base ? S->lines.base : ""; d.txt_len = S->lines.span; d.line_off = S->lines.off.data (); d.line_len = S->lines.len.data (); d.n_lines = n_lines;
    d.line_dom = S->dom.data (); d.line_diverse = S->diverse.data ();
    {
        EngineLease E (vb, name);
        if (gzb_domq_prepare (E.e, &d, 1, 0) != GZB_OK) plugin_abort ("gzb_domq_prepare", name, gzb_last_error (E.e));
    }
    if (!force) {                                                            // :80-131 (every line is "tested": the GPU has looked at all of them)
        uint32_t tested = 0, with_dom = 0;
        for (uint32_t i = 0; i < n_lines; i++) if (S->lines.len[i]) { tested++; with_dom += !S->diverse[i]; }
        if (!tested || 100.0 * with_dom / tested <= 50.0) { delete S; *slot = nullptr; return false; }
    }
    *g_host2.local_prm8 (qual_ctx, 0) = d.num_norm_qs | 0x80;                 // :234
    g_host2.seg_denorm (vb, qual_ctx, d.denorm, (uint32_t)d.num_doms * d.num_norm_qs);   // :236-244
    *slot = S;
    return true;
}

extern "C" GZB_COMPRESS (gzb_codec_domq_compress)
{
    (void)uncompressed; (void)get_line_cb;
    NEED (codec_state, name); NEED (local_alloc, name); NEED (local_set_len, name); NEED (assign_sub_codec, name); NEED (sub_est_size, name); NEED (header_set, name);
    void **slot = g_host2.codec_state (vb, ctx);
    DomqState *S = (DomqState *)*slot;
    if (!S) plugin_abort ("domq", name, "codec_domq_compress without codec_domq_comp_init");
    static thread_local Codec sub = CODEC_NONE_;
    if (!S->split_done) {                                                    // (a second entry after soft-failing continues at the sub-codec, :392)
        Timer tm { vb, 0 };
        gzb_domq_vb &d = S->d;
        const uint64_t total = S->lines.total;
        d.qual_cap = (uint32_t)(2 * total + 16); d.runs_cap = (uint32_t)(total + 16); d.mplx_cap = d.n_lines + 16; d.divr_cap = (uint32_t)(total + 16);
        std::vector<uint8_t> q (d.qual_cap), r (d.runs_cap), dv (d.divr_cap);  // worst-case staging here; the contexts get the real lengths (:404-419 allocates by estimate and grows)
        d.qual = q.data (); d.runs = r.data (); d.divr = dv.data ();
        d.mplx = g_host2.local_alloc (vb, ctx, 2, d.mplx_cap);
        {
            EngineLease E (vb, name);
            if (gzb_domq_split (E.e, &d, 1, 0) != GZB_OK) plugin_abort ("gzb_domq_split", name, gzb_last_error (E.e));
        }
        memcpy (g_host2.local_alloc (vb, ctx, 0, d.qual_len + 1), d.qual, d.qual_len); g_host2.local_set_len (ctx, 0, d.qual_len);
        memcpy (g_host2.local_alloc (vb, ctx, 1, d.runs_len + 1), d.runs, d.runs_len); g_host2.local_set_len (ctx, 1, d.runs_len);
        g_host2.local_set_len (ctx, 2, d.mplx_len);
        if (d.divr_len) memcpy (g_host2.local_alloc (vb, ctx, 3, d.divr_len + 1), d.divr, d.divr_len);
        g_host2.local_set_len (ctx, 3, d.divr_len);
        if (g_host2.add_lines) {                                             // z_file->domq_lines … (:489-492)
            uint64_t nd = 0; for (uint32_t i = 0; i < d.n_lines; i++) nd += S->diverse[i] && S->lines.len[i];
            uint64_t nl = 0; for (uint32_t i = 0; i < d.n_lines; i++) nl += S->lines.len[i] != 0;
            g_host2.add_lines (0, nl - nd); g_host2.add_lines (1, nd);
        }
        sub = g_host2.assign_sub_codec (vb, ctx, 0);                         // :497-500
        g_host2.header_set (header, GZB_HDR_SUB_CODEC, sub);
        S->split_done = true;
    }
    uint64_t qlen = 0;
    const char *qual = g_host2.local_data (ctx, 0, &qlen);
    *uncompressed_len = (uint32_t)qlen;                                      // :504
    if (*compressed_len < g_host2.sub_est_size (sub, qlen)) {                // :507-511
        if (soft_fail) return false;
        plugin_abort ("domq", name, "compressed buffer too small and soft_fail is off");
    }
    const bool ok = sub_compress (sub, vb, ctx, header, qual, uncompressed_len, compressed, compressed_len, HARD_FAIL, name);
    delete S; *slot = nullptr;
    return ok;
}

namespace { struct ReconStage { std::vector<uint8_t> out, missing; std::vector<uint64_t> off; uint32_t next = 0; uint64_t ht_next = 0; }; }

extern "C" void gzb_codec_domq_reconstruct (VBlockP vb, Codec codec, ContextP ctx, uint32_t len, bool reconstruct)
{
    (void)codec;
    const char *name = "QUAL";
    NEED (codec_state, name); NEED (recon_line_lens, name); NEED (local_data, name); NEED (recon_at, name); NEED (recon_advance, name);
    Timer tm { vb, 0 };
    void **slot = g_host2.codec_state (vb, ctx);
    ReconStage *R = (ReconStage *)*slot;
    if (!R) {                                                                // first line of the VBlock: decode all of them
        R = new ReconStage ();
        uint32_t n_lines = 0;
        const uint32_t *lens = g_host2.recon_line_lens (vb, ctx, &n_lines);
        R->off.resize ((size_t)n_lines + 1); R->off[0] = 0;
        for (uint32_t i = 0; i < n_lines; i++) R->off[i + 1] = R->off[i] + lens[i];
        R->out.resize (R->off[n_lines] + 16);
        gzb_domq_piz_vb p; memset (&p, 0, sizeof p);
        uint64_t l;
        p.qual = g_host2.local_data (ctx, 0, &l); p.qual_len = (uint32_t)l;
        p.runs = g_host2.local_data (ctx, 1, &l); p.runs_len = (uint32_t)l;
        p.mplx = g_host2.local_data (ctx, 2, &l); p.mplx_len = (uint32_t)l;
        p.divr = g_host2.local_data (ctx, 3, &l); p.divr_len = (uint32_t)l;
        NEED (codec_table, name);                                           // (for DOMQ: the de-normalisation table the adapter decoded from the DOMQRUNS dictionary)
        const uint8_t *den = g_host2.codec_table (vb, ctx);
        p.num_norm_qs = *g_host2.local_prm8 (ctx, 0) & 0x7f;
        p.denorm = den + 1; p.denorm_len = (uint32_t)den[0] * p.num_norm_qs;  // den[0] = number of doms, then [num_doms][num_norm_qs]
        p.line_len = lens; p.n_lines = n_lines; p.out = R->out.data (); p.out_cap = R->off[n_lines];
        EngineLease E (vb, name);
        if (gzb_domq_reconstruct (E.e, &p, 1, 0) != GZB_OK) plugin_abort ("gzb_domq_reconstruct", name, gzb_last_error (E.e));
        *slot = R;
    }
    while (R->next + 1 < R->off.size () && R->off[R->next + 1] == R->off[R->next] && len) R->next++;   // (empty lines are not routed to the codec)
    if (R->next + 1 >= R->off.size () || R->off[R->next + 1] - R->off[R->next] != len) plugin_abort ("domq reconstruct", name, "len differs from the line table");
    if (reconstruct) { memcpy (g_host2.recon_at (vb), R->out.data () + R->off[R->next], len); g_host2.recon_advance (vb, (int32_t)len); }
    if (++R->next + 1 == R->off.size ()) { delete R; *slot = nullptr; }
}

// ================================================================ NORMQ (src/codec_normq.c)
namespace { struct NormqState { Codec sub = CODEC_NONE_; }; }

extern "C" GZB_COMPRESS (gzb_codec_normq_compress)
{
    NEED (codec_state, name); NEED (local_alloc, name); NEED (local_set_len, name); NEED (local_data, name); NEED (assign_sub_codec, name); NEED (sub_est_size, name); NEED (header_set, name);
    if (uncompressed || !get_line_cb) plugin_abort ("normq", name, "only callback option is supported");       // :33
    void **slot = g_host2.codec_state (vb, ctx);
    NormqState *S = (NormqState *)*slot;
    if (soft_fail) {                                                         // first entry (:39: a second entry, after soft-failing, continues at the sub-codec)
        Timer tm { vb, 5 };
        Lines Q = gather_lines (vb, ctx, get_line_cb, name);
        const uint32_t n_lines = (uint32_t)Q.len.size ();
        gzb_normq_vb v; memset (&v, 0, sizeof v);
        v.txt = Q.base ? Q.base : ""; v.txt_len = Q.span; v.line_off = Q.off.data (); v.line_len = Q.len.data (); v.is_rev = Q.rev.data (); v.n_lines = n_lines;
        v.local = g_host2.local_alloc (vb, ctx, 0, Q.total + 16); v.local_cap = Q.total + 16;                   // buf_alloc_exact (:45)
        {
            EngineLease E (vb, name);
            if (gzb_normq_gather (E.e, &v, 1, 0) != GZB_OK) plugin_abort ("gzb_normq_gather", name, gzb_last_error (E.e));
        }
        g_host2.local_set_len (ctx, 0, v.local_len);
        if (g_host2.add_lines) g_host2.add_lines (3, n_lines);                                                   // z_file->normq_lines (:41-42)
        delete S; S = new NormqState (); *slot = S;
        S->sub = g_host2.assign_sub_codec (vb, ctx, 0);                                                          // :64-66
        g_host2.header_set (header, GZB_HDR_SUB_CODEC, S->sub);
    }
    else if (!S) plugin_abort ("normq", name, "second entry without a first one");
    uint64_t qlen = 0;
    const char *qual = g_host2.local_data (ctx, 0, &qlen);
    *uncompressed_len = (uint32_t)qlen;                                                                          // :70
    if (*compressed_len < g_host2.sub_est_size (S->sub, qlen)) {                                                 // :73-77
        if (soft_fail) return false;
        plugin_abort ("normq", name, "compressed buffer too small and soft_fail is off");
    }
    const Codec sub = S->sub;
    delete S; *slot = nullptr;
    return sub_compress (sub, vb, ctx, header, qual, uncompressed_len, compressed, compressed_len, HARD_FAIL, name);   // :81
}

extern "C" void gzb_codec_normq_reconstruct (VBlockP vb, Codec codec, ContextP ctx, uint32_t len, bool reconstruct)
{
    (void)codec;
    const char *name = "QUAL";
    NEED (codec_state, name); NEED (recon_line_lens, name); NEED (recon_seq_table, name); NEED (local_data, name); NEED (recon_at, name); NEED (recon_advance, name);
    Timer tm { vb, 5 };
    void **slot = g_host2.codec_state (vb, ctx);
    ReconStage *R = (ReconStage *)*slot;
    if (!R) {                                                                // first line of the VBlock: all of them at once
        R = new ReconStage ();
        uint32_t n_lines = 0;
        const uint32_t *lens = g_host2.recon_line_lens (vb, ctx, &n_lines);
        const char *txt = nullptr; uint64_t txt_len = 0; const uint64_t *seq_off = nullptr; const uint8_t *is_rev = nullptr;
        g_host2.recon_seq_table (vb, ctx, &txt, &txt_len, &seq_off, &is_rev);                                    // (only the strands are used: last_flags.rev_comp of every line, :97)
        R->off.resize ((size_t)n_lines + 1); R->off[0] = 0;
        for (uint32_t i = 0; i < n_lines; i++) R->off[i + 1] = R->off[i] + lens[i];
        R->out.resize (R->off[n_lines] + 16); R->missing.assign ((size_t)n_lines + 1, 0);
        gzb_normq_vb v; memset (&v, 0, sizeof v);
        uint64_t l = 0;
        v.local = g_host2.local_data (ctx, 0, &l); v.local_len = l;
        v.line_len = lens; v.is_rev = is_rev; v.n_lines = n_lines;
        v.out = R->out.data (); v.out_cap = R->off[n_lines] + 16; v.missing = R->missing.data ();
        EngineLease E (vb, name);
        if (gzb_normq_reconstruct (E.e, &v, 1, 0) != GZB_OK) plugin_abort ("gzb_normq_reconstruct", name, gzb_last_error (E.e));
        *slot = R;
    }
    while (R->next + 1 < R->off.size () && R->off[R->next + 1] == R->off[R->next] && len) R->next++;
    if (R->next + 1 >= R->off.size () || R->off[R->next + 1] - R->off[R->next] != len) plugin_abort ("normq reconstruct", name, "len differs from the line table");
    if (R->missing[R->next]) { NEED (missing_quality, name); g_host2.missing_quality (vb, reconstruct); }       // :91-94
    else if (reconstruct) { memcpy (g_host2.recon_at (vb), R->out.data () + R->off[R->next], len); g_host2.recon_advance (vb, (int32_t)len); }   // :96-101
    if (++R->next + 1 == R->off.size ()) { delete R; *slot = nullptr; }
}

// ================================================================ HOMP and T0 (src/codec_homp.c, src/codec_t0.c)
namespace {
struct HompState { Codec sub = CODEC_NONE_; std::vector<char> data; };

// first entry: every line condensed in place and its length updated (:132-190, t0 :69-109), then the sub-codec on the condensed strings — the
// reference hands it the line callback (data = 0): the same bytes, here as one buffer.  A second entry (after a soft fail) goes straight to the sub-codec.
bool homp_like_compress (int mode, VBlockP vb, ContextP ctx, SectionHeaderP header, const char *uncompressed, uint32_t *uncompressed_len, LocalGetLineCB get_line_cb,
                         char *compressed, uint32_t *compressed_len, FailType soft_fail, const char *name)
{
    NEED (codec_state, name); NEED (seq_line, name); NEED (local_set_len, name); NEED (assign_sub_codec, name); NEED (sub_est_size, name); NEED (header_set, name); NEED (update_line_len, name);
    if (uncompressed || !get_line_cb) plugin_abort (mode ? "t0" : "homp", name, "only callback option is supported");
    void **slot = g_host2.codec_state (vb, ctx);
    HompState *S = (HompState *)*slot;
    if (soft_fail) {
        Timer tm { vb, mode ? 7 : 6 };
        Lines Q = gather_lines (vb, ctx, get_line_cb, name);
        const uint32_t n_lines = (uint32_t)Q.len.size ();
        std::vector<uint64_t> seq_off (n_lines), str_off (n_lines); std::vector<uint32_t> new_len (n_lines + 1, 0);
        std::vector<char> txt; txt.reserve (2 * Q.total + 16);
        const uint32_t skip_below = mode ? 1 : 2;                           // homp leaves a quality of length <= 1 alone (:137), t0 an empty string (:75): no SEQ is asked for
        for (uint32_t i = 0; i < n_lines; i++) {
            char *sq = nullptr; uint32_t sl = 0; bool r = false;
            if (Q.len[i] >= skip_below) {
                g_host2.seq_line (vb, ctx, i, &sq, &sl, &r);
                if (sl != Q.len[i]) plugin_abort (mode ? "t0" : "homp", name, "expecting the string's length == seq_len");
            }
            seq_off[i] = txt.size (); txt.insert (txt.end (), sq, sq + sl);
            str_off[i] = txt.size (); txt.insert (txt.end (), Q.ptr[i], Q.ptr[i] + Q.len[i]);
        }
        delete S; S = new HompState (); *slot = S;
        S->data.resize (Q.total + 16);
        gzb_homp_vb v; memset (&v, 0, sizeof v);
        v.txt = txt.data (); v.txt_len = txt.size (); v.str_off = str_off.data (); v.str_len = Q.len.data (); v.seq_off = seq_off.data (); v.n_lines = n_lines;
        v.local = S->data.data (); v.local_cap = S->data.size (); v.new_len = new_len.data ();
        {
            EngineLease E (vb, name);
            if (gzb_homp_condense (E.e, &v, 1, mode, 0) != GZB_OK) plugin_abort ("gzb_homp_condense", name, gzb_last_error (E.e));
        }
        uint64_t at = 0;
        for (uint32_t i = 0; i < n_lines; i++) {                            // in place, like the reference (:143,179-186)
            memcpy (Q.ptr[i], S->data.data () + at, new_len[i]);
            if (new_len[i] != Q.len[i]) g_host2.update_line_len (vb, ctx, i, new_len[i]);
            at += new_len[i];
        }
        S->data.resize (v.local_len);
        g_host2.local_set_len (ctx, 0, v.local_len);                        // ctx->local.len32 -= … (:184)
        if (!mode && g_host2.add_lines) g_host2.add_lines (4, n_lines);     // z_file->homp_lines (:128-129)
        S->sub = g_host2.assign_sub_codec (vb, ctx, 0);                     // :193-195
        g_host2.header_set (header, GZB_HDR_SUB_CODEC, S->sub);
    }
    else if (!S) plugin_abort (mode ? "t0" : "homp", name, "second entry without a first one");
    *uncompressed_len = (uint32_t)S->data.size ();                          // :199
    if (*compressed_len < g_host2.sub_est_size (S->sub, S->data.size ())) { // :202-207
        if (soft_fail) return false;
        plugin_abort (mode ? "t0" : "homp", name, "compressed buffer too small and soft_fail is off");
    }
    const bool ok = sub_compress (S->sub, vb, ctx, header, S->data.data (), uncompressed_len, compressed, compressed_len, HARD_FAIL, name);
    delete S; *slot = nullptr;
    return ok;
}

void homp_like_reconstruct (int mode, VBlockP vb, ContextP ctx, uint32_t len, bool reconstruct)
{
    const char *name = mode ? "t0:Z" : "QUAL";
    NEED (codec_state, name); NEED (recon_line_lens, name); NEED (recon_seq_table, name); NEED (local_data, name); NEED (recon_at, name); NEED (recon_advance, name);
    Timer tm { vb, mode ? 7 : 6 };
    void **slot = g_host2.codec_state (vb, ctx);
    ReconStage *R = (ReconStage *)*slot;
    if (!R) {                                                                // first line of the VBlock: all of them at once
        R = new ReconStage ();
        uint32_t n_lines = 0;
        const uint32_t *lens = g_host2.recon_line_lens (vb, ctx, &n_lines);
        const char *txt = nullptr; uint64_t txt_len = 0; const uint64_t *seq_off = nullptr; const uint8_t *is_rev = nullptr;
        if (!g_host2.recon_seq_table (vb, ctx, &txt, &txt_len, &seq_off, &is_rev) || !seq_off) plugin_abort ("homp reconstruct", name, "the reads' SEQ is needed up front");
        R->off.resize ((size_t)n_lines + 1); R->off[0] = 0;
        for (uint32_t i = 0; i < n_lines; i++) R->off[i + 1] = R->off[i] + lens[i];
        R->out.resize (R->off[n_lines] + 16); R->missing.assign ((size_t)n_lines + 1, 0);
        gzb_homp_vb v; memset (&v, 0, sizeof v);
        uint64_t l = 0;
        v.local = g_host2.local_data (ctx, 0, &l); v.local_len = l;
        v.txt = txt; v.txt_len = txt_len; v.seq_off = seq_off; v.str_len = lens; v.n_lines = n_lines;
        v.out = R->out.data (); v.out_cap = R->off[n_lines] + 16; v.missing = R->missing.data ();
        EngineLease E (vb, name);
        if (gzb_homp_expand (E.e, &v, 1, mode, 0) != GZB_OK) plugin_abort ("gzb_homp_expand", name, gzb_last_error (E.e));
        *slot = R;
    }
    while (R->next + 1 < R->off.size () && R->off[R->next + 1] == R->off[R->next] && len) R->next++;
    if (R->next + 1 >= R->off.size () || R->off[R->next + 1] - R->off[R->next] != len) plugin_abort ("homp reconstruct", name, "len differs from the line table");
    if (R->missing[R->next]) { NEED (missing_quality, name); g_host2.missing_quality (vb, reconstruct); }       // homp :241-244
    else if (reconstruct) { memcpy (g_host2.recon_at (vb), R->out.data () + R->off[R->next], len); g_host2.recon_advance (vb, (int32_t)len); }
    if (++R->next + 1 == R->off.size ()) { delete R; *slot = nullptr; }
}
}

extern "C" GZB_COMPRESS (gzb_codec_homp_compress) { return homp_like_compress (GZB_HP_HOMP, vb, ctx, header, uncompressed, uncompressed_len, get_line_cb, compressed, compressed_len, soft_fail, name); }
extern "C" GZB_COMPRESS (gzb_codec_t0_compress)   { return homp_like_compress (GZB_HP_T0,   vb, ctx, header, uncompressed, uncompressed_len, get_line_cb, compressed, compressed_len, soft_fail, name); }
extern "C" void gzb_codec_homp_reconstruct (VBlockP vb, Codec codec, ContextP ctx, uint32_t len, bool reconstruct) { (void)codec; homp_like_reconstruct (GZB_HP_HOMP, vb, ctx, len, reconstruct); }
extern "C" void gzb_codec_t0_reconstruct   (VBlockP vb, Codec codec, ContextP ctx, uint32_t len, bool reconstruct) { (void)codec; homp_like_reconstruct (GZB_HP_T0,   vb, ctx, len, reconstruct); }

// ================================================================ PBWT
extern "C" GZB_COMPRESS (gzb_codec_pbwt_compress)
{
    (void)uncompressed; (void)uncompressed_len; (void)get_line_cb; (void)compressed; (void)soft_fail; (void)header;
    NEED (pbwt_dims, name); NEED (local_data, name); NEED (local_alloc, name); NEED (local_set_len, name);
    Timer tm { vb, 3 };
    uint32_t n_lines = 0, w = 0;
    g_host2.pbwt_dims (vb, ctx, &n_lines, &w, 0);
    uint64_t ht_len = 0;
    const char *ht = g_host2.local_data (ctx, 0, &ht_len);
    uint64_t rcap = std::max<uint64_t> (w, ht_len / 5) / 4 + 1024, fcap = std::max<uint64_t> (w, ht_len / 30) / 4 + 1024;    // initial allocations of :249-250, in words
    EngineLease E (vb, name);
    for (;;) {                                                               // the reference grows the buffers line by line (:257-259): here, by retrying
        uint32_t *runs = (uint32_t *)g_host2.local_alloc (vb, ctx, 1, rcap * 4), *fgrc = (uint32_t *)g_host2.local_alloc (vb, ctx, 2, fcap * 4);
        uint32_t nr = 0, nf = 0;
        const int rc = gzb_pbwt_encode (E.e, ht, n_lines, w, runs, (uint32_t)rcap, &nr, fgrc, (uint32_t)fcap, &nf, 0);
        if (rc == GZB_OK) { g_host2.local_set_len (ctx, 1, (uint64_t)nr * 4); g_host2.local_set_len (ctx, 2, (uint64_t)nf * 4); break; }
        if (rc != GZB_E_BADARG || rcap > 2 * ht_len + 16) plugin_abort ("gzb_pbwt_encode", name, gzb_last_error (E.e));
        rcap = std::min<uint64_t> (rcap * 4, 2 * ht_len + 17); fcap = std::min<uint64_t> (fcap * 4, ht_len + 17);
    }
    if (g_host2.local_free) g_host2.local_free (vb, ctx, 0);                  // no section for the matrix itself (:281-283)
    *compressed_len = 0;
    return true;
}

extern "C" GZB_UNCOMPRESS (gzb_codec_pbwt_uncompress)
{
    (void)codec; (void)param; (void)uncompressed_buf; (void)uncompressed_len; (void)sub_codec;
    NEED (pbwt_dims, name); NEED (local_data, name); NEED (local_alloc, name); NEED (local_set_len, name);
    Timer tm { vb, 3 };
    const uint32_t n_fgrc = compressed_len / 4;                               // this is the FGRC section, stored big endian (:377-381)
    if (n_fgrc < 2) plugin_abort ("pbwt uncompress", name, "FGRC too short");
    std::vector<uint32_t> fgrc (n_fgrc);
    for (uint32_t i = 0; i < n_fgrc; i++) { uint32_t x; memcpy (&x, compressed + 4 * (size_t)i, 4); fgrc[i] = __builtin_bswap32 (x); }
    const uint64_t len = (uint64_t)fgrc[n_fgrc - 2] | ((uint64_t)fgrc[n_fgrc - 1] << 32);
    uint32_t n_lines = 0, w = 0;
    g_host2.pbwt_dims (vb, ctx, &n_lines, &w, 0);                              // HT_n_lines (ctx = FGRC; the adapter resolves the matrix context)
    if (!n_lines || !len) plugin_abort ("pbwt uncompress", name, "Expecting num_lines and uncompressed_len to be > 0");      // :299-300
    uint64_t runs_bytes = 0;
    const uint32_t *runs = (const uint32_t *)g_host2.local_data (ctx, -1, &runs_bytes);    // RUNS.local, already uncompressed (:376)
    char *ht = g_host2.local_alloc (vb, ctx, -2, len);                         // the matrix context's local (:303)
    uint64_t ht_len = 0;
    EngineLease E (vb, name);
    if (gzb_pbwt_decode (E.e, runs, (uint32_t)(runs_bytes / 4), fgrc.data (), n_fgrc, n_lines, ht, len, &ht_len, 0) != GZB_OK)
        plugin_abort ("gzb_pbwt_decode", name, gzb_last_error (E.e));
    g_host2.local_set_len (ctx, -2, ht_len);
    w = (uint32_t)(ht_len / n_lines);
    g_host2.pbwt_dims (vb, ctx, &n_lines, &w, 1);                              // ht_ctx->ht_per_line (:310)
}

// codec_pbwt_reconstruct: one haplotype per call — text assembly on the matrix codec_pbwt_uncompress left in the context
extern "C" void gzb_codec_pbwt_reconstruct (VBlockP vb, Codec codec, ContextP ctx, uint32_t len, bool reconstruct)
{
    (void)codec; (void)len; (void)reconstruct;
    const char *name = "GT_HT";
    NEED (codec_state, name); NEED (local_data, name); NEED (recon_at, name); NEED (recon_advance, name);
    void **slot = g_host2.codec_state (vb, ctx);
    ReconStage *R = (ReconStage *)*slot;
    if (!R) *slot = R = new ReconStage ();
    uint64_t n = 0;
    const uint8_t *m = (const uint8_t *)g_host2.local_data (ctx, 0, &n);
    uint8_t ht = '*';
    do { ht = m[R->ht_next++]; } while (ht == '*' && R->ht_next < n);         // skip unused spots (:411-414)
    const bool last = R->ht_next >= n;
    if (!(g_host2.drop_curr_line && g_host2.drop_curr_line (vb))) {
        char *at = g_host2.recon_at (vb);
        auto put_int = [&] (int64_t v) { const int k = snprintf (at, 24, "%lld", (long long)v); g_host2.recon_advance (vb, k); };
        if ((ht >= '0' && ht <= '9') || ht == '.') { at[0] = (char)ht; g_host2.recon_advance (vb, 1); }       // :419-421
        else if (ht == '-') g_host2.recon_advance (vb, -1);                                                    // ploidy padding (:424-426)
        else if (ht == '%') {                                                                                  // :431-438
            if (at[-1] == '|' || at[-1] == '/') { at[-1] = '/'; at[0] = '.'; g_host2.recon_advance (vb, 1); }
            else { at[0] = '.'; g_host2.recon_advance (vb, 1); }
        }
        else if (ht == '&') { NEED (pbwt_big_allele, name); put_int (g_host2.pbwt_big_allele (vb) + 245); }   // NUM_SMALL_ALLELES (:440-444)
        else put_int ((uint8_t)(ht - '0'));                                                                    // :446-447
    }
    if (last) { delete R; *slot = nullptr; }
}

// ================================================================ LONGR
extern "C" GZB_COMPRESS (gzb_codec_longr_compress)
{
    (void)uncompressed;
    NEED (seq_line, name); NEED (codec_table, name); NEED (local_alloc, name); NEED (local_set_len, name); NEED (assign_sub_codec, name); NEED (sub_est_size, name); NEED (header_set, name);
    if (!soft_fail) plugin_abort ("longr", name, "second entry not expected");                                // :164
    Codec sub;
    char *lens_local;
    {
        Timer tm { vb, 4 };
        Lines Q = gather_lines (vb, ctx, get_line_cb, name);
        const uint32_t n_lines = (uint32_t)Q.len.size ();
        // the reads' SEQ (fastq_zip_seq / sam_zip_seq, :170): same text buffer as QUAL in FASTQ, but nothing here relies on it — both are copied into one staging text
        std::vector<uint64_t> seq_off (n_lines), qual_off (n_lines); std::vector<uint32_t> slen (n_lines);
        std::vector<uint8_t> rev (n_lines, 0);
        std::vector<char> txt; txt.reserve (2 * Q.total + 16);
        bool any_rev = false;
        for (uint32_t i = 0; i < n_lines; i++) {
            char *s = nullptr; uint32_t sl = 0; bool r = false;
            if (Q.len[i]) g_host2.seq_line (vb, ctx, i, &s, &sl, &r);
            if (Q.len[i] && sl != Q.len[i] && !(Q.len[i] == 1 && Q.ptr[i][0] == ' ')) plugin_abort ("longr", name, "Expecting seq_len == qual_len");   // :190-191
            seq_off[i] = txt.size (); txt.insert (txt.end (), s, s + sl); slen[i] = sl;
            qual_off[i] = txt.size (); txt.insert (txt.end (), Q.ptr[i], Q.ptr[i] + Q.len[i]);
            rev[i] = r; any_rev |= r;
            if (!Q.len[i]) slen[i] = 0;
        }
        if (Q.total != *uncompressed_len) plugin_abort ("longr", name, "Expecting total_len == uncompressed_len");   // :194-195
        gzb_longr_vb v; memset (&v, 0, sizeof v);
        v.txt = txt.data (); v.txt_len = txt.size (); v.seq_off = seq_off.data (); v.qual_off = qual_off.data (); v.len = slen.data (); v.qual_len = Q.len.data ();
        v.is_rev = any_rev ? rev.data () : nullptr; v.n_lines = n_lines;
        memcpy (v.value_to_bin, g_host2.codec_table (vb, ctx), 256);
        v.values = g_host2.local_alloc (vb, ctx, 1, Q.total + 16);                                             // values_ctx->local (:199-200)
        lens_local = g_host2.local_alloc (vb, ctx, 0, 65536 * 4);                                              // lens_ctx->local (:233-235)
        v.lens_be = (uint32_t *)lens_local;
        {
            EngineLease E (vb, name);
            if (gzb_longr_encode (E.e, &v, 1, 0) != GZB_OK) plugin_abort ("gzb_longr_encode", name, gzb_last_error (E.e));
        }
        g_host2.local_set_len (ctx, 1, Q.total);
        g_host2.local_set_len (ctx, 0, 65536 * 4);
        if (g_host2.add_lines) g_host2.add_lines (2, n_lines);                                                 // :181
        sub = g_host2.assign_sub_codec (vb, ctx, 0);                                                           // :244-247
        g_host2.header_set (header, GZB_HDR_SUB_CODEC, sub);
    }
    *uncompressed_len = 65536 * 4;                                                                              // :250
    if (*compressed_len < g_host2.sub_est_size (sub, *uncompressed_len)) plugin_abort ("longr", name, "compressed buffer too small for the lengths");   // :253-255
    return sub_compress (sub, vb, ctx, header, lens_local, uncompressed_len, compressed, compressed_len, HARD_FAIL, name);
}

extern "C" void gzb_codec_longr_reconstruct (VBlockP vb, Codec codec, ContextP ctx, uint32_t len, bool reconstruct)
{
    (void)codec;
    const char *name = "QUAL";
    NEED (codec_state, name); NEED (recon_line_lens, name); NEED (recon_seq_table, name); NEED (local_data, name); NEED (recon_at, name); NEED (recon_advance, name); NEED (codec_table, name);
    Timer tm { vb, 4 };
    void **slot = g_host2.codec_state (vb, ctx);
    ReconStage *R = (ReconStage *)*slot;
    if (!R) {
        R = new ReconStage ();
        uint32_t n_lines = 0;
        const uint32_t *lens = g_host2.recon_line_lens (vb, ctx, &n_lines);
        const char *txt = nullptr; uint64_t txt_len = 0; const uint64_t *seq_off = nullptr; const uint8_t *is_rev = nullptr;
        if (!g_host2.recon_seq_table (vb, ctx, &txt, &txt_len, &seq_off, &is_rev)) plugin_abort ("longr reconstruct", name, "the reads' SEQ are not available up front");
        R->off.resize ((size_t)n_lines + 1); R->off[0] = 0;
        for (uint32_t i = 0; i < n_lines; i++) R->off[i + 1] = R->off[i] + lens[i];
        R->out.resize (R->off[n_lines] + 16); R->missing.assign ((size_t)n_lines + 1, 0);
        gzb_longr_vb v; memset (&v, 0, sizeof v);
        v.txt = txt; v.txt_len = txt_len; v.seq_off = seq_off; v.len = lens; v.is_rev = is_rev; v.n_lines = n_lines;
        memcpy (v.value_to_bin, g_host2.codec_table (vb, ctx), 256);
        uint64_t l = 0;
        v.lens_be = (uint32_t *)g_host2.local_data (ctx, 0, &l);                                               // still big endian, as uncompressed (:303-304 converts in place)
        v.values = g_host2.local_data (ctx, 1, &l); v.n_bases = l;
        v.qual_out = R->out.data (); v.missing = R->missing.data ();
        EngineLease E (vb, name);
        if (gzb_longr_decode (E.e, &v, 1, 0) != GZB_OK) plugin_abort ("gzb_longr_decode", name, gzb_last_error (E.e));
        *slot = R;
    }
    while (R->next + 1 < R->off.size () && R->off[R->next + 1] == R->off[R->next] && len) R->next++;   // (a line without SEQ is not routed to the codec)
    if (R->next + 1 >= R->off.size () || R->off[R->next + 1] - R->off[R->next] != len) plugin_abort ("longr reconstruct", name, "len differs from the line table");
    if (R->missing[R->next]) { NEED (missing_quality, name); g_host2.missing_quality (vb, reconstruct); }       // :368-369
    else {
        memcpy (g_host2.recon_at (vb), R->out.data () + R->off[R->next], len);                                 // (the reference writes even when !reconstruct, :366)
        if (reconstruct) g_host2.recon_advance (vb, (int32_t)len);
    }
    if (++R->next + 1 == R->off.size ()) { delete R; *slot = nullptr; }
}

// codec_complex_est_size (src/codec.c:458-462): room for the pre-processed data (1.1 x) through the arithmetic coder, plus 10 KB
extern "C" uint32_t gzb_codec_complex_est_size (Codec codec, uint64_t uncompressed_len)
{
    (void)codec;
    const uint64_t preprocessed_len = (uint64_t)(uncompressed_len * 1.1);
    return gzb_est_size (GZB_CODEC_ARTB, preprocessed_len) + 10000;
}
extern "C" uint32_t gzb_codec_longr_est_size (Codec codec, uint64_t uncompressed_len)
{
    (void)uncompressed_len;
    return gzb_codec_complex_est_size (codec, 65536 * 4);                    // the length array (:53-56)
}
