// plugin.cu — entry points with exactly the reference's plug-in signatures (src/codec.h:17-40) for the simple codecs, so
// that CODEC_ARGS (src/codec.h:47-115) can point at them:  codec_{RANB,RANW,RANb,RANw,ARTB,ARTW,ARTb,ARTw}_compress
// (src/codec_htscodecs.c:77-94), codec_rans_uncompress / codec_arith_uncompress (:100-129), codec_*_est_size (:26-33).
// genozip's VBlock / Context / Buffer stay opaque: the ≤200-line adapter inside genozip registers four accessors
// (gzb_plugin_register).  Error behaviour mirrors the reference: `false` only for soft_fail with a
// too-small output buffer (src/compressor.c:90); everything else aborts through the host's ABORT.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/gzb200.h"

static gzb_plugin_host g_host = { nullptr, nullptr, nullptr, nullptr };
static int g_n_devices = 1;
static thread_local gzb_engine *tl_engine[64];

extern "C" void gzb_plugin_register (const gzb_plugin_host *host, int n_devices)
{
    if (host) g_host = *host;
    int avail = gzb_device_count ();
    g_n_devices = n_devices > 0 ? (n_devices < avail ? n_devices : avail) : avail;
    if (g_n_devices > 64) g_n_devices = 64;
}

static void plugin_abort (const char *what, const char *name, const char *detail)
{
    char msg[512];
    snprintf (msg, sizeof msg, "gzb200: %s failed for \"%s\": %s", what, name ? name : "?", detail ? detail : "");
    if (g_host.abort_msg) g_host.abort_msg (msg);
    fprintf (stderr, "%s\n", msg);
    abort ();                                                    // there is no CPU fallback (north_star)
}

// device = (vblock_i - 1) mod n_devices — VBlocks are independent, so the dispatcher's compute threads shard them round-robin
static gzb_engine *engine_for (VBlockP vb, const char *name)
{
    if (g_n_devices <= 0) plugin_abort ("device lookup", name, "no CUDA device");
    const uint32_t vblock_i = g_host.vb_vblock_i ? g_host.vb_vblock_i (vb) : 1;
    const int dev = gzb_vb_device (vblock_i, g_n_devices);
    if (!tl_engine[dev] && gzb_engine_create (dev, &tl_engine[dev]) != GZB_OK) plugin_abort ("gzb_engine_create", name, gzb_last_error (nullptr));
    return tl_engine[dev];
}

// codec_hts_compress (src/codec_htscodecs.c:40-74): contiguous data or one line at a time through the callback
static bool hts_compress (int codec, VBlockP vb, ContextP ctx, const char *uncompressed, uint32_t *uncompressed_len,
                          LocalGetLineCB get_line_cb, char *compressed, uint32_t *compressed_len, FailType soft_fail, const char *name)
{
    gzb_engine *e = engine_for (vb, name);
    std::vector<char> gathered;
    if (get_line_cb) {                                           // :51-64
        if (!g_host.vb_num_lines) plugin_abort ("line gather", name, "adapter did not register vb_num_lines");
        gathered.reserve (*uncompressed_len);
        const uint32_t n_lines = g_host.vb_num_lines (vb);
        for (uint32_t line_i = 0; line_i < n_lines; line_i++) {
            char *line = nullptr; uint32_t line_len = 0;
            get_line_cb (vb, ctx, line_i, &line, &line_len, *uncompressed_len - (uint32_t)gathered.size (), nullptr);
            if (line_len) gathered.insert (gathered.end (), line, line + line_len);
        }
        if (gathered.size () != *uncompressed_len) plugin_abort ("line gather", name, "total length from callbacks != uncompressed_len");
        uncompressed = gathered.data ();
    }
    gzb_section s; memset (&s, 0, sizeof s);
    s.codec = codec; s.in = uncompressed; s.in_len = *uncompressed_len; s.out = compressed; s.out_cap = *compressed_len;
    if (gzb_compress_sections (e, &s, 1, 0) != GZB_OK) plugin_abort ("gzb_compress_sections", name, gzb_last_error (e));
    if (s.status == GZB_SOFT_FAIL) {
        if (soft_fail) return false;                             // caller grows z_data and calls again (src/compressor.c:90-110)
        plugin_abort ("compress", name, "output buffer too small and soft_fail is off");
    }
    if (s.status != GZB_OK) plugin_abort ("compress", name, "section status");
    *compressed_len = s.out_len;
    return true;
}

static void hts_uncompress (int codec, VBlockP vb, const char *compressed, uint32_t compressed_len,
                            BufferP uncompressed_buf, uint64_t uncompressed_len, const char *name)
{
    if (!uncompressed_len || !compressed_len) plugin_abort ("uncompress", name, "zero length");          // ASSERTNOTZEROn (:103-104)
    if (!g_host.buffer_data) plugin_abort ("uncompress", name, "adapter did not register buffer_data");
    gzb_engine *e = engine_for (vb, name);
    gzb_section s; memset (&s, 0, sizeof s);
    s.codec = codec; s.in = compressed; s.in_len = compressed_len; s.out = g_host.buffer_data (uncompressed_buf); s.out_cap = (uint32_t)uncompressed_len;
    if (gzb_uncompress_sections (e, &s, 1, 0) != GZB_OK || s.status != GZB_OK || s.out_len != uncompressed_len)
        plugin_abort ("gzb_uncompress_sections", name, gzb_last_error (e));                              // ASSERT (:106-111)
}

#define GZB_COMPRESS_FUNC(NAME, CODEC) \
    extern "C" GZB_COMPRESS (gzb_codec_##NAME##_compress) \
    { (void)header; return hts_compress (CODEC, vb, ctx, uncompressed, uncompressed_len, get_line_cb, compressed, compressed_len, soft_fail, name); } \
    extern "C" uint32_t gzb_codec_##NAME##_est_size (Codec codec, uint64_t uncompressed_len) { (void)codec; return gzb_est_size (CODEC, uncompressed_len); }

GZB_COMPRESS_FUNC (RANB, GZB_CODEC_RANB)
GZB_COMPRESS_FUNC (RANW, GZB_CODEC_RANW)
GZB_COMPRESS_FUNC (RANb, GZB_CODEC_RANb)
GZB_COMPRESS_FUNC (RANw, GZB_CODEC_RANw)
GZB_COMPRESS_FUNC (ARTB, GZB_CODEC_ARTB)
GZB_COMPRESS_FUNC (ARTW, GZB_CODEC_ARTW)
GZB_COMPRESS_FUNC (ARTb, GZB_CODEC_ARTb)
GZB_COMPRESS_FUNC (ARTw, GZB_CODEC_ARTw)

extern "C" GZB_UNCOMPRESS (gzb_codec_rans_uncompress)
{ (void)ctx; (void)codec; (void)param; (void)sub_codec; hts_uncompress (GZB_CODEC_RANB, vb, compressed, compressed_len, uncompressed_buf, uncompressed_len, name); }
extern "C" GZB_UNCOMPRESS (gzb_codec_arith_uncompress)
{ (void)ctx; (void)codec; (void)param; (void)sub_codec; hts_uncompress (GZB_CODEC_ARTB, vb, compressed, compressed_len, uncompressed_buf, uncompressed_len, name); }
