// plugin.cu — entry points with exactly the reference's plug-in signatures (src/codec.h:17-40) for the simple codecs, so
// that CODEC_ARGS (src/codec.h:47-115) can point at them:  codec_{RANB,RANW,RANb,RANw,ARTB,ARTW,ARTb,ARTw}_compress
// (src/codec_htscodecs.c:77-94), codec_rans_uncompress / codec_arith_uncompress (:100-129), codec_*_est_size (:26-33);
// the engine pool the plug-in calls draw from; and the combiners (gzb_submit / gzb_wait) that turn the one-section-per-call
// pattern of comp_compress (src/compressor.c:82-86) into batches.
// genozip's VBlock / Context / Buffer stay opaque: the adapter inside genozip registers accessor tables
// (gzb_plugin_register, gzb_plugin_register2).  Error behaviour mirrors the reference: `false` only for
// soft_fail with a too-small output buffer (src/compressor.c:90); everything else aborts through the host's ABORT.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <deque>
#include <map>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include "plugin_internal.h"

namespace gzbp {

gzb_plugin_host  g_host  = {};
gzb_plugin_host2 g_host2 = {};
int g_n_devices = 1;

void plugin_abort (const char *what, const char *name, const char *detail)
{
    char msg[512];
    snprintf (msg, sizeof msg, "gzb200: %s failed for \"%s\": %s", what, name ? name : "?", detail ? detail : "");
    if (g_host.abort_msg) g_host.abort_msg (msg);
    fprintf (stderr, "%s\n", msg);
    abort ();                                                    // there is no CPU fallback (north_star)
}

// ---------------------------------------------------------------- engine pool
static std::mutex g_pool_mu;
static std::vector<gzb_engine *> g_pool[64];

// device = (vblock_i - 1) mod n_devices — VBlocks are independent, so the dispatcher's compute threads shard them round-robin
EngineLease::EngineLease (VBlockP vb, const char *name)
{
    if (g_n_devices <= 0) plugin_abort ("device lookup", name, "no CUDA device");
    const uint32_t vblock_i = g_host.vb_vblock_i ? g_host.vb_vblock_i (vb) : 1;
    dev = gzb_vb_device (vblock_i, g_n_devices);
    e = nullptr;
    {
        std::lock_guard<std::mutex> lk (g_pool_mu);
        if (!g_pool[dev].empty ()) { e = g_pool[dev].back (); g_pool[dev].pop_back (); }
    }
    if (!e && gzb_engine_create (dev, &e) != GZB_OK) plugin_abort ("gzb_engine_create", name, gzb_last_error (nullptr));
}
EngineLease::~EngineLease ()
{
    std::lock_guard<std::mutex> lk (g_pool_mu);
    g_pool[dev].push_back (e);
}

} // namespace gzbp

using namespace gzbp;

extern "C" void gzb_plugin_register (const gzb_plugin_host *host, int n_devices)
{
    if (host) g_host = *host;
    int avail = gzb_device_count ();
    g_n_devices = n_devices > 0 ? (n_devices < avail ? n_devices : avail) : avail;
    if (g_n_devices > 64) g_n_devices = 64;
}
extern "C" void gzb_plugin_register2 (const gzb_plugin_host2 *host) { if (host) g_host2 = *host; }

// ---------------------------------------------------------------- combiners
struct gzb_combiner {
    int device; bool compress; uint32_t linger_us;
    gzb_engine *eng = nullptr;
    std::mutex mu; std::condition_variable cv;
    struct Item { gzb_section sec; uint64_t ticket; };
    std::deque<Item> pending;                                    // submitted, not yet taken by a leader
    std::map<uint64_t, gzb_section> done;                        // finished, not yet collected
    uint64_t next_ticket = 1, batches = 0;
    bool leader_active = false;
};

extern "C" gzb_combiner *gzb_combiner_create (int device, int compress, uint32_t linger_us)
{
    gzb_combiner *c = new gzb_combiner ();
    c->device = device; c->compress = compress != 0; c->linger_us = linger_us;
    if (gzb_engine_create (device, &c->eng) != GZB_OK) { delete c; return nullptr; }
    return c;
}
extern "C" void gzb_combiner_destroy (gzb_combiner *c) { if (!c) return; gzb_engine_destroy (c->eng); delete c; }
extern "C" uint64_t gzb_combiner_batches (gzb_combiner *c) { std::lock_guard<std::mutex> lk (c->mu); return c->batches; }

extern "C" int gzb_submit (gzb_combiner *c, const gzb_section *sec, uint64_t *ticket)
{
    if (!c || !sec || !ticket) return GZB_E_BADARG;
    std::lock_guard<std::mutex> lk (c->mu);
    *ticket = c->next_ticket++;
    c->pending.push_back ({ *sec, *ticket });
    c->cv.notify_all ();
    return GZB_OK;
}

extern "C" int gzb_wait (gzb_combiner *c, uint64_t ticket, gzb_section *result)
{
    if (!c || !result) return GZB_E_BADARG;
    std::unique_lock<std::mutex> lk (c->mu);
    for (;;) {
        auto it = c->done.find (ticket);
        if (it != c->done.end ()) { *result = it->second; c->done.erase (it); return result->status < 0 ? result->status : GZB_OK; }
        if (c->leader_active || c->pending.empty ()) { c->cv.wait (lk); continue; }
        // become the leader: let the other compute threads arrive, then take everything that is pending
        c->leader_active = true;
        if (c->linger_us) c->cv.wait_for (lk, std::chrono::microseconds (c->linger_us), [] { return false; });
        std::vector<gzb_combiner::Item> batch (c->pending.begin (), c->pending.end ());
        c->pending.clear ();
        lk.unlock ();
        std::vector<gzb_section> secs (batch.size ());
        for (size_t i = 0; i < batch.size (); i++) secs[i] = batch[i].sec;
        const int rc = c->compress ? gzb_compress_sections (c->eng, secs.data (), (uint32_t)secs.size (), 0)
                                   : gzb_uncompress_sections (c->eng, secs.data (), (uint32_t)secs.size (), 0);
        lk.lock ();
        for (size_t i = 0; i < batch.size (); i++) {
            if (rc != GZB_OK && secs[i].status >= 0) secs[i].status = rc;                      // the whole call failed: every section of it did
            c->done[batch[i].ticket] = secs[i];
        }
        c->batches++;
        c->leader_active = false;
        c->cv.notify_all ();
    }
}

// process-wide combiners of the plug-in layer, per device and direction
static bool g_combining = false; static uint32_t g_linger_us = 200;
static std::mutex g_comb_mu;
static gzb_combiner *g_comb[64][2];
extern "C" void gzb_plugin_set_combining (int on, uint32_t linger_us) { g_combining = on != 0; g_linger_us = linger_us; }

int gzbp::run_section (gzb_engine *e, int dev, gzb_section *s, bool compress)
{
    if (!g_combining) return compress ? gzb_compress_sections (e, s, 1, 0) : gzb_uncompress_sections (e, s, 1, 0);
    gzb_combiner *c;
    {
        std::lock_guard<std::mutex> lk (g_comb_mu);
        c = g_comb[dev][compress];
        if (!c) c = g_comb[dev][compress] = gzb_combiner_create (dev, compress, g_linger_us);
    }
    if (!c) return GZB_E_CUDA;
    uint64_t t;
    int rc = gzb_submit (c, s, &t);
    return rc ? rc : gzb_wait (c, t, s);
}

extern "C" void gzb_plugin_shutdown (void)
{
    {
        std::lock_guard<std::mutex> lk (g_pool_mu);
        for (auto &v : g_pool) { for (gzb_engine *e : v) gzb_engine_destroy (e); v.clear (); }
    }
    std::lock_guard<std::mutex> lk (g_comb_mu);
    for (auto &d : g_comb) for (auto &c : d) { gzb_combiner_destroy (c); c = nullptr; }
}

// ---------------------------------------------------------------- simple codecs
// codec_hts_compress (src/codec_htscodecs.c:40-74): contiguous data or one line at a time through the callback
static bool hts_compress (int codec, VBlockP vb, ContextP ctx, const char *uncompressed, uint32_t *uncompressed_len,
                          LocalGetLineCB get_line_cb, char *compressed, uint32_t *compressed_len, FailType soft_fail, const char *name)
{
    EngineLease L (vb, name);
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
    if (run_section (L.e, L.dev, &s, true) != GZB_OK) plugin_abort ("gzb_compress_sections", name, gzb_last_error (L.e));
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
    EngineLease L (vb, name);
    gzb_section s; memset (&s, 0, sizeof s);
    s.codec = codec; s.in = compressed; s.in_len = compressed_len; s.out = g_host.buffer_data (uncompressed_buf); s.out_cap = (uint32_t)uncompressed_len;
    if (run_section (L.e, L.dev, &s, false) != GZB_OK || s.status != GZB_OK || s.out_len != uncompressed_len)
        plugin_abort ("gzb_uncompress_sections", name, gzb_last_error (L.e));                            // ASSERT (:106-111)
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
