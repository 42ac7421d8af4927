// api.cu — C-ABI of libgzb200.so (include/gzb200.h): engine lifecycle, the host-side planner that expands a
// batch of sections into leaves and tiles, device workspace management, staging of host buffers.
//
// Reference interfaces replaced: codec_*_compress / codec_rans_uncompress / codec_arith_uncompress
// (src/codec_htscodecs.c:40-129) and codec_*_est_size (:26-33).  There is no CPU fallback anywhere in this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <numeric>
#include <mutex>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "hts_enc.cuh"
#include "engine.h"

using namespace gzb;

static thread_local std::string g_last_error;       // creation errors, per calling thread (engines are made by the thread that owns them)

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

// ------------------------------------------------------------------------------------------------ bounds (double arithmetic of the reference)
// rans_compress_bound_4x16 (rANS_static4x16pr.c:357-369) and arith_compress_bound (arith_dynamic.c:74-80).  The reference
// object evaluates 1.05*size + C as one fused multiply-add followed by plain additions, left to right.
static double bound_base (uint32_t n, int order)
{
    if (order == 0) return std::fma (1.05, (double)n, 257.0 * 3) + 4;
    double t = std::fma (1.05, (double)n, 257.0 * 257 * 3);
    t += 4; t += 257 * 3; t += 4;
    return t;
}
static uint32_t rans_bound (uint32_t n, int order)
{
    int N = order >> 8; if (!N) N = 4;
    order &= 0xff;
    int sz = (int)(bound_base (n, order) + ((order & F_PACK) ? 1 : 0) + ((order & F_RLE) ? 1 + 257*3 + 4 : 0) + 20 + ((order & F_STRIPE) ? 1 + 5*N : 0));
    return sz + (sz & 1) + 2;
}
static uint32_t arith_bound (uint32_t n, int order)
{
    return (uint32_t)(bound_base (n, order) + ((order & F_PACK) ? 1 : 0) + ((order & F_RLE) ? 1 + 257*3 + 4 : 0) + 5);
}

static bool codec_info (int codec, uint8_t *coder, uint8_t *order)
{
    switch (codec) {
        case GZB_CODEC_RANB: *coder = CODER_RANS;  *order = 0x01; return true;
        case GZB_CODEC_RANW: *coder = CODER_RANS;  *order = 0x19; return true;
        case GZB_CODEC_RANb: *coder = CODER_RANS;  *order = 0x81; return true;
        case GZB_CODEC_RANw: *coder = CODER_RANS;  *order = 0x99; return true;
        case GZB_CODEC_ARTB: *coder = CODER_ARITH; *order = 0x01; return true;
        case GZB_CODEC_ARTW: *coder = CODER_ARITH; *order = 0x19; return true;
        case GZB_CODEC_ARTb: *coder = CODER_ARITH; *order = 0x81; return true;
        case GZB_CODEC_ARTw: *coder = CODER_ARITH; *order = 0x99; return true;
        default: return false;
    }
}

extern "C" uint32_t gzb_est_size (int codec, uint64_t n)
{
    uint8_t coder, order;
    if (!codec_info (codec, &coder, &order) || n > 0xffffffffull) return 0;  // (a section is at most 4 GB - 1: src/sections.h data_uncompressed_len is 32 bits)
    return 1024 + (coder == CODER_RANS ? rans_bound ((uint32_t)n, order) : arith_bound ((uint32_t)n, order));
}

// ------------------------------------------------------------------------------------------------ engine
extern "C" int gzb_build_is_emulation (void)
{
#ifdef GZB_SIMT_EMULATION
    return 1;
#else
    return 0;
#endif
}

extern "C" int gzb_device_count (void)
{
    int n = 0;
    if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError (); return 0; }
    return n;
}

extern "C" int gzb_engine_create (int device, gzb_engine **out)
{
    *out = nullptr;
    int n = gzb_device_count ();
    if (n <= 0 || device < 0 || device >= n) { g_last_error = "no usable CUDA device (the product has no CPU fallback)"; return GZB_E_NOCUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties (&prop, device) != cudaSuccess || prop.major < 10) {
        g_last_error = "device is not sm_100-class: libgzb200 ships sm_100a kernels only"; return GZB_E_NOCUDA;
    }
    cudaGetLastError ();                                                    // (an earlier failure of this thread — an allocation that did not fit, say — is not this engine's)
    gzb_engine *e = new gzb_engine ();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    bool ok = cudaSetDevice (device) == cudaSuccess && cudaStreamCreateWithFlags (&e->stream, cudaStreamNonBlocking) == cudaSuccess;
    cudaEvent_t *evs[8] = { &e->ev0, &e->ev1, &e->ev2, &e->ev3, &e->ev4, &e->ev5, &e->ev6, &e->ev7 };
    for (int i = 0; ok && i < 8; i++) ok = cudaEventCreate (evs[i]) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags (&e->stream2, cudaStreamNonBlocking) == cudaSuccess
            && cudaStreamCreateWithFlags (&e->stream3, cudaStreamNonBlocking) == cudaSuccess
            && cudaStreamCreateWithFlags (&e->stream4, cudaStreamNonBlocking) == cudaSuccess
            && cudaStreamCreateWithFlags (&e->stream_copy, cudaStreamNonBlocking) == cudaSuccess
            && cudaStreamCreateWithFlags (&e->stream_copy2, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {                                                              // a half-made engine would fork its chains onto null streams
        g_last_error = std::string ("engine creation failed: ") + cudaGetErrorString (cudaGetLastError ());
        gzb_engine_destroy (e); return GZB_E_CUDA;
    }
    // log tables for the order-1 table-size decision: must come from the same libm the reference links
    double l10[257], l12[257];
    for (int k = 0; k <= 256; k++) { l10[k] = log ((double)(1024 + k)); l12[k] = log ((double)(4096 + k)); }
    upload_log_tables (l10, l12);
    const cudaError_t err = cudaGetLastError ();
    if (err != cudaSuccess) {
        g_last_error = err == cudaErrorNoKernelImageForDevice || err == cudaErrorInvalidDeviceFunction ? "no sm_100a kernel image for this device"
                                                                                                      : std::string ("engine creation failed: ") + cudaGetErrorString (err);
        gzb_engine_destroy (e); return err == cudaErrorNoKernelImageForDevice || err == cudaErrorInvalidDeviceFunction ? GZB_E_NOCUDA : GZB_E_CUDA;
    }
    *out = e;
    return GZB_OK;
}

extern "C" void gzb_engine_destroy (gzb_engine *e)
{
    if (!e) return;
    cudaSetDevice (e->device);
    if (e->stream) cudaStreamSynchronize (e->stream);
    if (e->ws) cudaFree (e->ws);
    if (e->dq_buf) cudaFree (e->dq_buf);
    if (e->dq_session && e->dq_free) e->dq_free (e->dq_session);
    if (e->pin) cudaFreeHost (e->pin);
    cudaEvent_t evs[8] = { e->ev0, e->ev1, e->ev2, e->ev3, e->ev4, e->ev5, e->ev6, e->ev7 };
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy (ev);
    if (e->stream2) cudaStreamDestroy (e->stream2);
    if (e->stream3) cudaStreamDestroy (e->stream3);
    if (e->stream4) cudaStreamDestroy (e->stream4);
    if (e->stager && e->stager_free) e->stager_free (e->stager);            // (joins the feeder threads before their streams go)
    if (e->stream_copy) { cudaStreamSynchronize (e->stream_copy); cudaStreamDestroy (e->stream_copy); }
    if (e->stream_copy2) { cudaStreamSynchronize (e->stream_copy2); cudaStreamDestroy (e->stream_copy2); }
    if (e->stream) cudaStreamDestroy (e->stream);
    delete e;
}

// give the grow-only workspace and staging back (a batch of another shape follows, or other engines need the memory)
extern "C" int gzb_engine_trim (gzb_engine *e)
{
    if (!e) return GZB_E_BADARG;
    cudaSetDevice (e->device);
    CK (cudaStreamSynchronize (e->stream));
    if (e->ws) { cudaFree (e->ws); e->ws = nullptr; e->ws_cap = 0; }
    if (e->dq_session && e->dq_free) { e->dq_free (e->dq_session); e->dq_session = nullptr; }
    if (e->dq_buf) { cudaFree (e->dq_buf); e->dq_buf = nullptr; e->dq_cap = 0; }
    if (e->pin) { cudaFreeHost (e->pin); e->pin = nullptr; e->pin_cap = 0; }
    return GZB_OK;
}

extern "C" const char *gzb_last_error (gzb_engine *e) { return e ? e->err.c_str () : g_last_error.c_str (); }
extern "C" void *gzb_engine_stream (gzb_engine *e) { return (void *)e->stream; }
extern "C" int gzb_engine_sync (gzb_engine *e) { cudaSetDevice (e->device); CK (cudaStreamSynchronize (e->stream)); return GZB_OK; }
extern "C" int gzb_vb_device (uint32_t vblock_i, int n_devices) { return n_devices > 0 ? (int)((vblock_i ? vblock_i - 1 : 0) % (uint32_t)n_devices) : 0; }
extern "C" uint64_t gzb_kernel_launches (gzb_engine *e) { return e->launches; }
extern "C" float gzb_last_chain_ms (gzb_engine *e) { return e->last_chain_ms; }
extern "C" float gzb_last_kernel_ms (gzb_engine *e, int which)
{
    switch (which) {
        case 0:  return e->last_rans_ms;
        case 1:  return e->last_arith_ms;
        case 2:  return e->last_domain_ms;
        case 3:  return e->last_o0_ms;
        case 4:  return e->last_split_ms;
        case 5:  return e->last_arith_all_ms;
        case 6: case 7: case 8: return e->last_split_part_ms[which - 6];
        default: return 0;
    }
}

int engine_reserve (gzb_engine *e, size_t ws_bytes, size_t pin_bytes)
{
    if (ws_bytes > e->ws_cap) {
        CK (cudaStreamSynchronize (e->stream));
        if (e->ws) cudaFree (e->ws);
        e->ws = nullptr; e->ws_cap = 0;
        size_t want = ws_bytes + ws_bytes / 8 + (1u << 20);
        CK (cudaMalloc (&e->ws, want));
        e->ws_cap = want;
    }
    if (pin_bytes > e->pin_cap) {
        CK (cudaStreamSynchronize (e->stream));
        if (e->pin) cudaFreeHost (e->pin);
        e->pin = nullptr; e->pin_cap = 0;
        size_t want = pin_bytes + pin_bytes / 8 + (1u << 16);
        CK (cudaMallocHost (&e->pin, want));
        e->pin_cap = want;
    }
    return GZB_OK;
}

// ------------------------------------------------------------------------------------------------ compress
namespace {

struct Carver {                        // carves 256-byte aligned regions out of the workspace; first pass measures
    uint8_t *base; size_t off;
    template <typename T> T *take (size_t count) {
        size_t bytes = (count * sizeof (T) + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

constexpr size_t SMALL_COPY = 65536;   // host sections below this are gathered through the pinned staging buffer

uint32_t leaf_out_cap (uint8_t coder, uint32_t order_req, uint32_t n)
{
    double base = 1.06 * n + 2048;
    if (coder == CODER_ARITH) base = (double)n + 1024;                   // the arithmetic leaf stops once its body reaches n bytes
    if (coder == CODER_RANS && (order_req & 1)) {
        double tab = std::min (200.0 * 1024, 3.0 * n + 1024);      // O1 table and its nested-compression scratch
        base += 2 * tab + 1024;
    }
    return ((uint32_t)base + 31) & ~15u;
}

// Host → device input copies.  Large sections go straight from the caller's buffer; runs of consecutive small
// sections (contiguous in the device input region) are gathered into pinned memory and sent as one transfer.
template <typename FD, typename FH, typename FL>
int stage_inputs (gzb_engine *e, uint8_t *stage, uint32_t n, FD dev, FH host, FL len)
{
    cudaStream_t st = e->stream;
    size_t so = 0;
    uint32_t i = 0;
    while (i < n) {
        uint32_t l = len (i);
        if (!l) { i++; continue; }
        if (l >= SMALL_COPY) { CK (cudaMemcpyAsync (dev (i), host (i), l, cudaMemcpyHostToDevice, st)); i++; continue; }
        uint8_t *d0 = dev (i); size_t s0 = so, run = 0;
        while (i < n && len (i) < SMALL_COPY) {
            uint32_t li = len (i);
            if (li) { size_t rel = (size_t)(dev (i) - d0); memcpy (stage + s0 + rel, host (i), li); run = rel + li; }
            i++;
        }
        so = s0 + ((run + 255) & ~(size_t)255);
        CK (cudaMemcpyAsync (d0, stage + s0, run, cudaMemcpyHostToDevice, st));
    }
    return GZB_OK;
}

// Warp jobs over a length-sorted list: an item above BIG bytes is latency-critical and gets a warp of its own; smaller
// items are packed up to 8 per warp (4 lanes each) so that issue slots are shared.  A job never mixes `cls` values.
constexpr uint32_t BIG_LEAF = 32768;
template <typename FL, typename FC> std::vector<uint2> make_jobs (uint32_t n, FL len, FC cls)
{
    std::vector<uint2> jobs;
    uint32_t i = 0;
    while (i < n) {
        if (len (i) > BIG_LEAF) { jobs.push_back (make_uint2 (i, 1)); i++; continue; }
        uint32_t c = 1;
        while (c < 8 && i + c < n && cls (i + c) == cls (i) && len (i + c) <= BIG_LEAF) c++;
        jobs.push_back (make_uint2 (i, c));
        i += c;
    }
    return jobs;
}
int pick_arith_lpw (uint32_t n) { int l = 1; while (l < 32 && (uint64_t)n > 9472ull * l) l *= 2; return l; }

} // namespace

namespace { struct PackOut { void *arena; uint64_t cap; uint64_t *used; bool dev; bool sizes_only = false; }; }   // sizes_only: no output at all, out_len of every section

static int compress_impl (gzb_engine *e, gzb_section *secs, uint32_t n, uint32_t flags, const PackOut *pk)
{
    if (!e) return GZB_E_BADARG;
    if (pk && pk->used) *pk->used = 0;
    if (!n) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devall = flags & GZB_DEVICE_PTRS;
    auto in_dev  = [&] (uint32_t i) { return devall || (secs[i].sflags & GZB_SEC_IN_DEVICE); };
    auto out_dev = [&] (uint32_t i) { return pk ? true : (devall || (secs[i].sflags & GZB_SEC_OUT_DEVICE)); };   // (packed: no per-section staging)

    // ---- plan on the host
    std::vector<EncSection> hs (n);
    std::vector<EncLeaf> hl;
    std::vector<Tile> tiles, stiles;
    std::vector<uint32_t> rlist, alist;
    bool any_pack = false, any_o1 = false;
    size_t in_total = 0, out_total = 0, plane_total = 0, pack_total = 0, outbuf_total = 0;
    size_t small_in = 0;
    for (uint32_t i = 0; i < n; i++) {
        gzb_section &s = secs[i];
        uint8_t coder, order;
        if (!codec_info (s.codec, &coder, &order) || (!s.in && s.in_len) || (!s.out && !pk)) { s.status = GZB_E_BADARG; e->err = "bad section"; return GZB_E_BADARG; }
        EncSection &S = hs[i];
        memset (&S, 0, sizeof S);
        if (pk) s.out_cap = gzb_est_size (s.codec, s.in_len);                 // packed: the buffer as a whole has a capacity, a section has none
        S.n = s.in_len; S.coder = coder; S.order = order; S.out_cap = s.out_cap;
        S.soft_fail = s.out_cap < gzb_est_size (s.codec, s.in_len) - 1024;    // the reference returns NULL → false under soft_fail when *out_size < the bound (rANS_static4x16pr.c:1158); est_size is bound + 1 KB
        S.stripe = (order & F_STRIPE) && S.n > 20;
        S.first_leaf = (uint32_t)hl.size ();
        if (S.soft_fail) { S.n_leaves = 0; continue; }
        if (!in_dev (i))  in_total += (S.n + 15) & ~15ull;
        if (!out_dev (i)) out_total += ((size_t)std::min<uint32_t> (s.out_cap, gzb_est_size (s.codec, s.in_len)) + 15) & ~15ull;
        if (!in_dev (i) && S.n < SMALL_COPY) small_in += (S.n + 15) & ~15ull;
        auto add_leaf = [&] (uint32_t ln, uint32_t order_req, size_t plane_off) {
            EncLeaf L; memset (&L, 0, sizeof L);
            L.n = ln; L.section = i; L.coder = coder; L.order_req = (uint8_t)order_req;
            L.out_cap = leaf_out_cap (coder, order_req, ln);
            L.in = reinterpret_cast<const uint8_t *>(plane_off);              // patched to a pointer after carving
            uint32_t li = (uint32_t)hl.size ();
            hl.push_back (L);
            outbuf_total += L.out_cap;
            if (order_req & F_PACK) { any_pack = true; pack_total += ((size_t)ln + 1 + 15) & ~15ull; }
            if (coder == CODER_RANS && (order_req & 1)) any_o1 = true;
            for (uint32_t off = 0; off < ln; off += TILE) tiles.push_back (Tile { li, off });
            (coder == CODER_RANS ? rlist : alist).push_back (li);
        };
        if (!S.stripe) { add_leaf (S.n, order & ~F_STRIPE, 0); S.n_leaves = 1; }
        else {
            // candidate methods per byte-plane, in the reference's try order (rANS :1203-1214; arith :684-687,726-737)
            static const int rans_m[4] = { 1, 64, 128, 0 };
            static const int arith_m[4][4] = { {3, 1, 64, 0}, {2, 1, 0, 0}, {2, 1, 128, 0}, {2, 1, 128, 0} };
            uint32_t counts = 0, idx = 0;
            for (int p = 0; p < 4; p++) {
                uint32_t plen = S.n / 4 + ((S.n % 4) > (uint32_t)p), nc = 0;
                if (coder == CODER_RANS) {
                    for (int j = 0; j < 4; j++) if ((order & rans_m[j]) == rans_m[j]) { add_leaf (plen, rans_m[j] | F_NOSZ, plane_total + idx); nc++; }
                }
                else for (int j = 1; j <= arith_m[p][0]; j++) {
                    if ((order & 3) == 0 && (arith_m[p][j] & 1)) continue;
                    add_leaf (plen, arith_m[p][j] | F_NOSZ, plane_total + idx); nc++;
                }
                counts |= nc << (4 * p);
                idx += plen;
            }
            S.n_leaves = counts;
            S.planes = reinterpret_cast<uint8_t *>(plane_total);
            for (uint32_t off = 0; off < S.n; off += TILE) stiles.push_back (Tile { i, off });
            plane_total += (S.n + 15) & ~15ull;
        }
    }
    const uint32_t nl = (uint32_t)hl.size ();
    auto by_len = [&] (uint32_t a, uint32_t b) { return hl[a].n > hl[b].n; };
    // rANS: requested order 1 first (a warp job runs one order), longest first within each order
    std::stable_sort (rlist.begin (), rlist.end (), [&] (uint32_t a, uint32_t b) {
        int oa = hl[a].order_req & 1, ob = hl[b].order_req & 1;
        return oa != ob ? oa > ob : hl[a].n > hl[b].n; });
    std::stable_sort (alist.begin (), alist.end (), by_len);
    std::vector<uint2> rjobs = make_jobs ((uint32_t)rlist.size (), [&] (uint32_t i) { return hl[rlist[i]].n; },
                                          [&] (uint32_t i) { return (uint32_t)(hl[rlist[i]].order_req & 1); });

    // the split encoder (arith_split.cu) takes order-1 arithmetic leaves of at least GZB_AR_SPLIT_MIN symbols (default 32768; "off" disables)
    static const uint32_t split_min = [] { const char *v = getenv ("GZB_AR_SPLIT_MIN"); return !v ? 32768u : !strcmp (v, "off") ? 0xffffffffu : (uint32_t)strtoul (v, nullptr, 10); } ();
    uint32_t n_arith_big = 0;
    while (n_arith_big < alist.size () && hl[alist[n_arith_big]].n >= split_min) n_arith_big++;
    // arena estimate for alphabet-dependent tables; grown and replayed on overflow
    size_t arena_est = (size_t)4 << 20;
    for (auto &L : hl) {
        size_t m = std::min<size_t> (256, (size_t)L.n + 1);
        if (L.coder == CODER_RANS) arena_est += (L.order_req & 1) ? std::min<size_t> (m * m * 20 + m * CTXB, 64 * 64 * 20 + 64 * CTXB + (size_t)L.n / 8) + 4096 : 4096 + 64;
        else {
            arena_est += std::min<size_t> ((size_t)256 * 264 * 4, 256 * 72 * 4 + (size_t)L.n / 8) + 258 * 12 * 4 + 64;
            if ((L.order_req & 1) && L.n >= split_min) arena_est += 12 * (size_t)L.n + (257 + 64) * 4 + 64;   // split encoder: positions + records
        }
    }
    if (e->arena_hint > arena_est) arena_est = e->arena_hint;

    for (int attempt = 0; attempt < 4; attempt++) {
        // ---- carve the workspace (two passes: measure, then place)
        Carver c { nullptr, 0 };
        EncPlanDev P; memset (&P, 0, sizeof P);
        uint8_t *d_in = nullptr, *d_out = nullptr, *d_planes = nullptr, *d_pack = nullptr, *d_outbuf = nullptr, *d_arena = nullptr;
        uint32_t *d_hist0 = nullptr; unsigned long long *d_cursor = nullptr; int *d_overflow = nullptr;
        size_t meta_off = 0, meta_bytes = 0;
        for (int pass = 0; pass < 2; pass++) {
            c.off = 0;
            // metadata blob uploaded in one copy: sections | leaves | tiles | stripe tiles | lists
            meta_off = c.off;
            P.sections     = c.take<EncSection> (n);
            P.leaves       = c.take<EncLeaf> (nl ? nl : 1);
            P.tiles        = c.take<Tile> (tiles.size () + 1);
            P.stripe_tiles = c.take<Tile> (stiles.size () + 1);
            P.rans_list    = c.take<uint32_t> (rlist.size () + 1);
            P.rans_jobs    = c.take<uint2> (rjobs.size () + 1);
            P.arith_list   = c.take<uint32_t> (alist.size () + 1);
            meta_bytes = c.off - meta_off;
            P.dyn          = c.take<EncLeafDyn> (nl ? nl : 1);
            P.results      = c.take<SectionResult> (n);
            P.segs         = c.take<CopySeg> ((size_t)n * 16);
            P.stripe_hdr   = c.take<uint8_t> ((size_t)n * 32);
            if (pk) { P.pack_off = c.take<unsigned long long> ((size_t)n + 1); P.pack_cap = pk->cap;
                      P.pack_arena = pk->dev ? (uint8_t *)pk->arena : c.take<uint8_t> (pk->cap + 16); }
            d_cursor       = c.take<unsigned long long> (1);
            d_overflow     = c.take<int> (1);
            P.queue        = c.take<uint32_t> (Q_WORDS);
            d_hist0        = c.take<uint32_t> ((size_t)(nl ? nl : 1) * 256);
            d_in           = c.take<uint8_t> (in_total + 1);
            d_out          = c.take<uint8_t> (out_total + 1);
            d_planes       = c.take<uint8_t> (plane_total + 1);
            d_pack         = c.take<uint8_t> (pack_total + 1);
            d_outbuf       = c.take<uint8_t> (outbuf_total + 1);
            d_arena        = c.take<uint8_t> (arena_est);
            if (pass == 0) {
                int rc = engine_reserve (e, c.off, small_in + 512 * (size_t)n + meta_bytes + 8192);
                if (rc) return rc;
                c.base = e->ws;
            }
        }

        // ---- fill host copies with device pointers
        std::vector<uint8_t> meta (meta_bytes);
        std::vector<EncSection> S2 = hs;
        std::vector<EncLeaf> L2 = hl;
        {
            size_t io = 0, oo = 0, po = 0, bo = 0;
            for (uint32_t i = 0; i < n; i++) {
                EncSection &S = S2[i];
                if (S.soft_fail) continue;
                if (in_dev (i)) S.in = (const uint8_t *)secs[i].in;
                else { S.in = d_in + io; io += (S.n + 15) & ~15ull; }
                if (pk) S.out = nullptr;                                  // destinations relative to the section until k_pack_place
                else if (out_dev (i)) S.out = (uint8_t *)secs[i].out;
                else { S.out = d_out + oo; oo += ((size_t)std::min<uint32_t> (secs[i].out_cap, gzb_est_size (secs[i].codec, secs[i].in_len)) + 15) & ~15ull; }
                if (S.stripe) S.planes = d_planes + (size_t)hs[i].planes;
                uint32_t cnt = S.stripe ? ((S.n_leaves & 15) + ((S.n_leaves >> 4) & 15) + ((S.n_leaves >> 8) & 15) + ((S.n_leaves >> 12) & 15)) : 1;
                for (uint32_t k = 0; k < cnt; k++) {
                    EncLeaf &L = L2[S.first_leaf + k];
                    L.in = S.stripe ? d_planes + (size_t)hl[S.first_leaf + k].in : S.in;
                    L.hist0 = d_hist0 + (size_t)(S.first_leaf + k) * 256;
                    L.outbuf = d_outbuf + bo; bo += L.out_cap;
                    if (L.order_req & F_PACK) { L.packbuf = d_pack + po; po += ((size_t)L.n + 1 + 15) & ~15ull; }
                }
            }
            uint8_t *m = meta.data ();
            auto put = [&] (const void *dev, const void *src, size_t bytes) { if (bytes) memcpy (m + ((const uint8_t *)dev - (e->ws + meta_off)), src, bytes); };
            put (P.sections, S2.data (), n * sizeof (EncSection));
            put (P.leaves, L2.data (), nl * sizeof (EncLeaf));
            put (P.tiles, tiles.data (), tiles.size () * sizeof (Tile));
            put (P.stripe_tiles, stiles.data (), stiles.size () * sizeof (Tile));
            put (P.rans_list, rlist.data (), rlist.size () * 4);
            put (P.rans_jobs, rjobs.data (), rjobs.size () * sizeof (uint2));
            put (P.arith_list, alist.data (), alist.size () * 4);
        }

        P.n_sections = n; P.n_leaves = nl; P.n_tiles = (uint32_t)tiles.size (); P.n_stripe_tiles = (uint32_t)stiles.size ();
        P.n_rans = (uint32_t)rlist.size (); P.n_arith = (uint32_t)alist.size ();
        P.n_arith_big = n_arith_big; P.split_min = split_min;
        P.any_pack = any_pack; P.any_o1 = any_o1;
        P.n_rans_jobs = (uint32_t)rjobs.size (); P.arith_lpw = pick_arith_lpw (P.n_arith);
        P.copy_parts = (n <= 64) ? 32 : (n <= 1024 ? 8 : 2);
        P.arena = Arena { d_arena, (unsigned long long)arena_est, d_cursor, d_overflow };
        P.ev_chain0 = e->ev0; P.ev_chain1 = e->ev1; P.ev_chain2 = e->ev2; P.ev_arith0 = e->ev3; P.st2 = e->stream2; P.ev_o0 = e->ev4; P.st3 = e->stream3;
        P.ev_split = e->ev5; P.st4 = e->stream4; P.sm_count = e->sm_count; P.ev_prof[0] = e->ev6; P.ev_prof[1] = e->ev7;

        // ---- upload: metadata blob, inputs (small ones gathered through pinned staging)
        cudaStream_t st = e->stream;
        memcpy (e->pin, meta.data (), meta_bytes);
        CK (cudaMemcpyAsync (e->ws + meta_off, e->pin, meta_bytes, cudaMemcpyHostToDevice, st));
        {
            int rc = stage_inputs (e, e->pin + ((meta_bytes + 255) & ~(size_t)255), n,
                                   [&] (uint32_t i) { return (uint8_t *)S2[i].in; },
                                   [&] (uint32_t i) { return (const uint8_t *)secs[i].in; },
                                   [&] (uint32_t i) { return (S2[i].soft_fail || in_dev (i)) ? 0u : S2[i].n; });
            if (rc) return rc;
        }
        CK (cudaMemsetAsync (d_cursor, 0, 768, st));                       // cursor | overflow | queue counters (256-byte regions)
        CK (cudaMemsetAsync (d_hist0, 0, (size_t)(nl ? nl : 1) * 1024, st));

        // ---- run
        enc_run (P, st);
        e->launches += P.launches;
        CK (cudaGetLastError ());

        // ---- results
        std::vector<SectionResult> res (n);
        int h_over = 0; unsigned long long h_cursor = 0;
        CK (cudaMemcpyAsync (res.data (), P.results, n * sizeof (SectionResult), cudaMemcpyDeviceToHost, st));
        CK (cudaMemcpyAsync (&h_over, d_overflow, sizeof (int), cudaMemcpyDeviceToHost, st));
        CK (cudaMemcpyAsync (&h_cursor, d_cursor, sizeof h_cursor, cudaMemcpyDeviceToHost, st));
        CK (cudaStreamSynchronize (st));
        if (h_over) {                                                     // tables did not fit: grow the arena and replay the batch
            arena_est = (size_t)h_cursor + (h_cursor >> 2) + ((size_t)1 << 20);
            e->arena_hint = arena_est;
            continue;
        }
        float ms = 0; e->last_rans_ms = e->last_arith_ms = e->last_o0_ms = e->last_split_ms = 0;
        if (P.n_rans_jobs) { cudaEventElapsedTime (&ms, e->ev0, e->ev1); e->last_rans_ms = ms; }
        if (P.n_arith) { cudaEventElapsedTime (&ms, e->ev3, e->ev2); e->last_arith_ms = ms; cudaEventElapsedTime (&ms, e->ev3, e->ev4); e->last_o0_ms = ms; }
        if (P.n_arith && P.n_arith_big) {
            cudaEventElapsedTime (&ms, e->ev3, e->ev5); e->last_split_ms = ms;
            cudaEventElapsedTime (&e->last_split_part_ms[0], e->ev3, e->ev6);      // bucket (from the start of the chain phase), model, code
            cudaEventElapsedTime (&e->last_split_part_ms[1], e->ev6, e->ev7);
            cudaEventElapsedTime (&e->last_split_part_ms[2], e->ev7, e->ev5);
        }
        e->last_arith_all_ms = std::max (e->last_arith_ms, std::max (e->last_o0_ms, e->last_split_ms));
        e->last_chain_ms = std::max (e->last_rans_ms, e->last_arith_all_ms);

        for (uint32_t i = 0; i < n; i++) {
            secs[i].status = res[i].status; secs[i].out_len = res[i].out_len;
            if (res[i].status == 0 && res[i].out_len > secs[i].out_cap) { secs[i].status = GZB_SOFT_FAIL; secs[i].out_len = 0; }
        }
        if (pk) {
            std::vector<unsigned long long> off ((size_t)n + 1);
            CK (cudaMemcpyAsync (off.data (), P.pack_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
            CK (cudaStreamSynchronize (st));
            if (pk->used) *pk->used = off[n];
            if (pk->sizes_only) { for (uint32_t i = 0; i < n; i++) secs[i].out = nullptr; return GZB_OK; }
            if (off[n] > pk->cap) {                                          // nothing was written: the caller grows the buffer and calls again
                for (uint32_t i = 0; i < n; i++) { secs[i].status = GZB_SOFT_FAIL; secs[i].out_len = 0; secs[i].out = nullptr; }
                e->err = "packed output: the buffer is too small";
                return GZB_SOFT_FAIL;
            }
            for (uint32_t i = 0; i < n; i++) secs[i].out = (uint8_t *)pk->arena + off[i];
            if (!pk->dev && off[n]) { CK (cudaMemcpyAsync (pk->arena, P.pack_arena, off[n], cudaMemcpyDeviceToHost, st)); CK (cudaStreamSynchronize (st)); }
            return GZB_OK;
        }
        bool any_d2h = false;
        for (uint32_t i = 0; i < n; i++)
            if (!out_dev (i) && secs[i].status == 0 && secs[i].out_len) {
                CK (cudaMemcpyAsync (secs[i].out, S2[i].out, secs[i].out_len, cudaMemcpyDeviceToHost, st));
                any_d2h = true;
            }
        if (any_d2h) CK (cudaStreamSynchronize (st));
        return GZB_OK;
    }
    e->err = "device arena kept overflowing";
    return GZB_E_CUDA;
}

extern "C" int gzb_compress_sections (gzb_engine *e, gzb_section *secs, uint32_t n, uint32_t flags) { return compress_impl (e, secs, n, flags, nullptr); }

extern "C" int gzb_compress_sections_packed (gzb_engine *e, gzb_section *secs, uint32_t n, void *arena, uint64_t arena_cap, uint64_t *arena_used, uint32_t flags)
{
    if (!arena && arena_cap) return GZB_E_BADARG;
    PackOut pk { arena, arena_cap, arena_used, (flags & (GZB_DEVICE_PTRS | GZB_OUT_DEVICE)) != 0 };
    return compress_impl (e, secs, n, flags, &pk);
}

// ------------------------------------------------------------------------------------------------ codec assignment by size
// codec_assign_best_codec (src/codec.c:234-389) compresses a sample of <= CODEC_ASSIGN_SAMPLE_SIZE bytes of a context's data with every
// generic codec, one after the other, and sorts the results (sorter :128-173).  Here the samples of all the contexts that need a codec
// are compressed with the eight simple codecs of this path as ONE batch — the same bytes the reference's calls produce, so the same
// sizes — and nothing is written anywhere: only the lengths come back.
extern "C" int gzb_assign_codecs (gzb_engine *e, gzb_assign_item *items, uint32_t n, uint32_t flags)
{
    if (!e || (!items && n)) return GZB_E_BADARG;
    static const int codecs[8] = { GZB_CODEC_RANB, GZB_CODEC_RANW, GZB_CODEC_RANb, GZB_CODEC_RANw, GZB_CODEC_ARTB, GZB_CODEC_ARTW, GZB_CODEC_ARTb, GZB_CODEC_ARTw };
    std::vector<gzb_section> secs; secs.reserve ((size_t)n * 8);
    for (uint32_t i = 0; i < n; i++) {
        gzb_assign_item &it = items[i];
        it.sample_len = (uint32_t)std::min<uint64_t> (it.len, GZB_ASSIGN_SAMPLE_SIZE);
        it.best = GZB_CODEC_UNKNOWN;
        for (int k = 0; k < 8; k++) it.size[k] = 0;
        if (it.sample_len < GZB_MIN_LEN_FOR_COMPRESSION) continue;             // "if too small - don't assign" (:317-318): the section goes out as CODEC_NONE (compressor.c:56-58)
        if (!it.data) return GZB_E_BADARG;
        for (int k = 0; k < 8; k++) {
            gzb_section sc; memset (&sc, 0, sizeof sc);
            sc.codec = codecs[k]; sc.in = it.data; sc.in_len = it.sample_len; sc.out_cap = 0xffffffffu;
            secs.push_back (sc);
        }
    }
    if (secs.empty ()) return GZB_OK;
    uint64_t used = 0;
    PackOut pk { nullptr, 0, &used, true, true };
    int rc = compress_impl (e, secs.data (), (uint32_t)secs.size (), flags & (GZB_DEVICE_PTRS | GZB_IN_DEVICE), &pk);
    if (rc) return rc;
    size_t j = 0;
    for (uint32_t i = 0; i < n; i++) {
        gzb_assign_item &it = items[i];
        if (it.sample_len < GZB_MIN_LEN_FOR_COMPRESSION) continue;
        // the sorter's last two rules, which are all that is left of it without the clock: smaller size first, equal sizes -> the lower codec
        // (the non-packing variant, :167-169); CODEC_NONE competes with the sample's own length (:325)
        uint64_t best_size = it.sample_len; int best = GZB_CODEC_NONE;
        for (int k = 0; k < 8; k++, j++) {
            if (secs[j].status != GZB_OK) { e->err = "gzb_assign_codecs: a sample failed to compress"; return secs[j].status; }
            it.size[k] = secs[j].out_len;
            if ((uint64_t)it.size[k] + GZB_SECTION_HEADER_BYTES < best_size) { best_size = (uint64_t)it.size[k] + GZB_SECTION_HEADER_BYTES; best = codecs[k]; }
        }
        it.best = best;
    }
    return GZB_OK;
}

// ------------------------------------------------------------------------------------------------ batched copies on the device
namespace {
__global__ void k_copy_batch (const gzb_copy *cp, uint32_t n)
{
    if (blockIdx.x >= n) return;
    const gzb_copy c = cp[blockIdx.x];
    const uint8_t *s = (const uint8_t *)c.src; uint8_t *d = (uint8_t *)c.dst;
    const uint64_t part = ((c.len + gridDim.y - 1) / gridDim.y + 15) & ~15ull, b = blockIdx.y * part, en = b + part < c.len ? b + part : c.len;
    if (b >= en) return;
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const uint64_t n16 = (en - b) >> 4;
        for (uint64_t i = threadIdx.x; i < n16; i += blockDim.x) reinterpret_cast<uint4 *>(d + b)[i] = reinterpret_cast<const uint4 *>(s + b)[i];
        for (uint64_t i = b + (n16 << 4) + threadIdx.x; i < en; i += blockDim.x) d[i] = s[i];
    }
    else for (uint64_t i = b + threadIdx.x; i < en; i += blockDim.x) d[i] = s[i];
}
}

extern "C" int gzb_copy_batch (gzb_engine *e, const gzb_copy *copies, uint32_t n)
{
    if (!e || (!copies && n)) return GZB_E_BADARG;
    if (!n) return GZB_OK;
    cudaSetDevice (e->device);
    int rc = engine_reserve (e, (size_t)n * sizeof (gzb_copy) + 256, (size_t)n * sizeof (gzb_copy) + 256); if (rc) return rc;
    memcpy (e->pin, copies, (size_t)n * sizeof (gzb_copy));
    CK (cudaMemcpyAsync (e->ws, e->pin, (size_t)n * sizeof (gzb_copy), cudaMemcpyHostToDevice, e->stream));
    uint64_t longest = 0; for (uint32_t i = 0; i < n; i++) longest = std::max<uint64_t> (longest, copies[i].len);
    const uint32_t parts = (uint32_t)std::min<uint64_t> (64, std::max<uint64_t> (1, longest >> 16));
    k_copy_batch<<<dim3 (n, parts), 256, 0, e->stream>>>(reinterpret_cast<const gzb_copy *>(e->ws), n); e->launches++;
    CK (cudaStreamSynchronize (e->stream));                                 // (the pinned staging and the workspace are reused by the next call)
    CK (cudaGetLastError ());
    return GZB_OK;
}

// ------------------------------------------------------------------------------------------------ uncompress
extern "C" int gzb_uncompress_sections (gzb_engine *e, gzb_section *secs, uint32_t n, uint32_t flags)
{
    if (!e) return GZB_E_BADARG;
    if (!n) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devall = flags & GZB_DEVICE_PTRS;
    auto in_dev  = [&] (uint32_t i) { return devall || (secs[i].sflags & GZB_SEC_IN_DEVICE); };
    auto out_dev = [&] (uint32_t i) { return devall || (secs[i].sflags & GZB_SEC_OUT_DEVICE); };

    std::vector<DecSection> hs (n);
    std::vector<uint32_t> order_idx (n);
    size_t in_total = 0, out_total = 0, aux_total = 0, small_in = 0;
    uint32_t n_rans_sec = 0, n_arith_sec = 0;
    for (uint32_t i = 0; i < n; i++) {
        gzb_section &s = secs[i];
        uint8_t coder, order;
        // the reference asserts on zero lengths (codec_htscodecs.c:103-104, 120-121)
        if (!codec_info (s.codec, &coder, &order) || !s.in || !s.out || !s.in_len || !s.out_cap || s.out_cap >= 0x7fffffffu) {   // (the reference refuses out_sz >= INT_MAX)
            s.status = GZB_E_BADARG; e->err = "bad section"; return GZB_E_BADARG;
        }
        DecSection &S = hs[i]; memset (&S, 0, sizeof S);
        S.in_len = s.in_len; S.n = s.out_cap; S.coder = coder;
        if (!in_dev (i)) in_total += (S.in_len + 15) & ~15ull;
        if (!out_dev (i)) out_total += (S.n + 15) & ~15ull;
        aux_total += (S.n + 15) & ~15ull;
        if (!in_dev (i) && S.in_len < SMALL_COPY) small_in += (S.in_len + 15) & ~15ull;
        (coder == CODER_RANS ? n_rans_sec : n_arith_sec)++;
        order_idx[i] = i;
    }
    std::stable_sort (order_idx.begin (), order_idx.end (), [&] (uint32_t a, uint32_t b) { return hs[a].n > hs[b].n; });
    std::vector<uint32_t> rlist, alist;
    for (uint32_t k = 0; k < n; k++) { uint32_t i = order_idx[k]; for (int j = 0; j < 4; j++) (hs[i].coder == CODER_RANS ? rlist : alist).push_back (4 * i + j); }
    std::vector<uint2> rjobs = make_jobs ((uint32_t)rlist.size (), [&] (uint32_t i) { return hs[rlist[i] >> 2].n; }, [&] (uint32_t) { return 0u; });

    size_t arena_est = (size_t)4 << 20;
    for (auto &S : hs) arena_est += (S.coder == CODER_RANS) ? std::min<size_t> ((size_t)1400 << 10, 96 * 1024 + (size_t)S.in_len * 8) + 4 * 16384
                                                             : std::min<size_t> ((size_t)4 * 272 * 1024, 4 * 80 * 1024 + (size_t)S.n / 4);
    if (e->arena_hint_dec > arena_est) arena_est = e->arena_hint_dec;

    for (int attempt = 0; attempt < 4; attempt++) {
        Carver c { nullptr, 0 };
        DecPlanDev P; memset (&P, 0, sizeof P);
        uint8_t *d_in = nullptr, *d_out = nullptr, *d_planes = nullptr, *d_tmp = nullptr, *d_arena = nullptr;
        unsigned long long *d_cursor = nullptr; int *d_overflow = nullptr;
        size_t meta_off = 0, meta_bytes = 0;
        for (int pass = 0; pass < 2; pass++) {
            c.off = 0;
            meta_off = c.off;
            P.sections   = c.take<DecSection> (n);
            P.rans_list  = c.take<uint32_t> (rlist.size () + 1);
            P.rans_jobs  = c.take<uint2> (rjobs.size () + 1);
            P.arith_list = c.take<uint32_t> (alist.size () + 1);
            meta_bytes = c.off - meta_off;
            P.leaves     = c.take<DecLeaf> ((size_t)4 * n);
            P.results    = c.take<SectionResult> (n);
            d_cursor     = c.take<unsigned long long> (1);
            d_overflow   = c.take<int> (1);
            P.queue      = c.take<uint32_t> (Q_WORDS);
            d_in         = c.take<uint8_t> (in_total + 1);
            d_out        = c.take<uint8_t> (out_total + 1);
            d_planes     = c.take<uint8_t> (aux_total);
            d_tmp        = c.take<uint8_t> (aux_total);
            d_arena      = c.take<uint8_t> (arena_est);
            if (pass == 0) {
                int rc = engine_reserve (e, c.off, small_in + 512 * (size_t)n + meta_bytes + 8192);
                if (rc) return rc;
                c.base = e->ws;
            }
        }
        std::vector<DecSection> S2 = hs;
        {
            size_t io = 0, oo = 0, ao = 0;
            for (uint32_t i = 0; i < n; i++) {
                DecSection &S = S2[i];
                if (in_dev (i)) S.in = (const uint8_t *)secs[i].in; else { S.in = d_in + io; io += (S.in_len + 15) & ~15ull; }
                if (out_dev (i)) S.out = (uint8_t *)secs[i].out; else { S.out = d_out + oo; oo += (S.n + 15) & ~15ull; }
                S.planes = d_planes + ao; S.tmp = d_tmp + ao; ao += (S.n + 15) & ~15ull;
            }
        }
        std::vector<uint8_t> meta (meta_bytes);
        auto put = [&] (const void *dev, const void *src, size_t bytes) { if (bytes) memcpy (meta.data () + ((const uint8_t *)dev - (e->ws + meta_off)), src, bytes); };
        put (P.sections, S2.data (), n * sizeof (DecSection));
        put (P.rans_list, rlist.data (), rlist.size () * 4);
        put (P.rans_jobs, rjobs.data (), rjobs.size () * sizeof (uint2));
        put (P.arith_list, alist.data (), alist.size () * 4);

        P.n_sections = n; P.n_rans = (uint32_t)rlist.size (); P.n_arith = (uint32_t)alist.size ();
        P.n_rans_jobs = (uint32_t)rjobs.size (); P.arith_lpw = pick_arith_lpw (P.n_arith);
        P.parts = (n <= 64) ? 32 : (n <= 1024 ? 8 : 2);
        P.arena = Arena { d_arena, (unsigned long long)arena_est, d_cursor, d_overflow };
        P.ev_chain0 = e->ev0; P.ev_chain1 = e->ev1; P.ev_chain2 = e->ev2; P.ev_arith0 = e->ev3; P.st2 = e->stream2; P.ev_o0 = e->ev4; P.st3 = e->stream3; P.sm_count = e->sm_count;
        P.st4 = e->stream4; P.ev_long = e->ev5; P.long_min = chain_tune ().long_min; P.n_long_cand = 0;
        while (P.n_long_cand < alist.size () && hs[alist[P.n_long_cand] >> 2].n >= P.long_min) P.n_long_cand++;

        cudaStream_t st = e->stream;
        memcpy (e->pin, meta.data (), meta_bytes);
        CK (cudaMemcpyAsync (e->ws + meta_off, e->pin, meta_bytes, cudaMemcpyHostToDevice, st));
        {
            int rc = stage_inputs (e, e->pin + ((meta_bytes + 255) & ~(size_t)255), n,
                                   [&] (uint32_t i) { return (uint8_t *)S2[i].in; },
                                   [&] (uint32_t i) { return (const uint8_t *)secs[i].in; },
                                   [&] (uint32_t i) { return in_dev (i) ? 0u : S2[i].in_len; });
            if (rc) return rc;
        }
        CK (cudaMemsetAsync (d_cursor, 0, 768, st));                       // cursor | overflow | queue counters (256-byte regions)

        dec_run (P, st);
        e->launches += P.launches;
        CK (cudaGetLastError ());

        std::vector<SectionResult> res (n);
        int h_over = 0; unsigned long long h_cursor = 0;
        CK (cudaMemcpyAsync (res.data (), P.results, n * sizeof (SectionResult), cudaMemcpyDeviceToHost, st));
        CK (cudaMemcpyAsync (&h_over, d_overflow, sizeof (int), cudaMemcpyDeviceToHost, st));
        CK (cudaMemcpyAsync (&h_cursor, d_cursor, sizeof h_cursor, cudaMemcpyDeviceToHost, st));
        for (uint32_t i = 0; i < n; i++) if (!out_dev (i)) CK (cudaMemcpyAsync (secs[i].out, S2[i].out, S2[i].n, cudaMemcpyDeviceToHost, st));
        CK (cudaStreamSynchronize (st));
        if (h_over) { arena_est = (size_t)h_cursor + (h_cursor >> 2) + ((size_t)1 << 20); e->arena_hint_dec = arena_est; continue; }
        float ms = 0; e->last_rans_ms = e->last_arith_ms = e->last_o0_ms = e->last_split_ms = 0;
        if (P.n_rans_jobs) { cudaEventElapsedTime (&ms, e->ev0, e->ev1); e->last_rans_ms = ms; }
        if (P.n_arith) { cudaEventElapsedTime (&ms, e->ev3, e->ev2); e->last_arith_ms = ms; cudaEventElapsedTime (&ms, e->ev3, e->ev4); e->last_o0_ms = ms; }
        if (P.n_arith && P.n_long_cand) { cudaEventElapsedTime (&ms, e->ev3, e->ev5); e->last_split_ms = ms; }   // (which = 4 on the decode side: k_arith_decode_long)
        e->last_arith_all_ms = std::max (e->last_arith_ms, std::max (e->last_o0_ms, e->last_split_ms));
        e->last_chain_ms = std::max (e->last_rans_ms, e->last_arith_all_ms);
        int rc = GZB_OK;
        for (uint32_t i = 0; i < n; i++) {
            secs[i].status = res[i].status; secs[i].out_len = res[i].status ? 0 : res[i].out_len;
            if (res[i].status) { rc = GZB_E_CORRUPT; e->err = "malformed compressed section"; }
        }
        return rc;
    }
    e->err = "device arena kept overflowing";
    return GZB_E_CUDA;
}
