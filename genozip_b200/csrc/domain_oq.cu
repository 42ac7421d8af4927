// domain_oq.cu — the OQ codec of SAM / BAM (src/codec_oq.c): the original quality string OQ:Z of a read is multiplexed by the read's
// QUAL — the OQ character at position i goes to channel QUAL[i] - '!' (94 channels, each then an ordinary local section) — because
// recalibration maps almost every (OQ, context) to one QUAL, so a channel is nearly constant.  A channel that holds one character
// only is dropped and the character kept in the 94-byte OQ:Z.local (:103-107).
//
//   mux    codec_oq_compress before its sub-codec (:54-121): count pass over the QUAL of ALL lines (:61-72), mux pass over the lines
//          whose SEQ.len is not 0 (:88-99), monochar test (:103-107)
//   demux  codec_oq_reconstruct (:126-164) for every line of a VBlock at once
//
// Both are a STABLE distribution by key, the shape of arith_split.cu's bucket kernel: one CTA per VBlock, 32 warps owning 32
// consecutive ranges of lines; a count pass, a scan over (channel, warp), then every warp walks its lines again 32 characters at a
// time and ranks equal keys by lane (__match_any_sync), which keeps the order inside the chunk; cursors per (warp, channel) in shared memory.
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "engine.h"

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr int OQ_CH = 94, OQ_WARPS = 32, OQ_PAD = 96;

struct OqVb {
    const uint8_t  *txt;
    const uint64_t *qual_off;
    const uint32_t *qual_len;
    const uint64_t *oq_off;      // mux
    const uint32_t *seq_len;     // mux, may be NULL
    uint8_t        *out;         // demux
    const uint64_t *out_off;     // demux
    uint8_t        *chan;
    uint32_t       *count;       // [94]  mux: out; demux: in
    uint8_t        *mono;        // [94]  mux: out; demux: in
    uint32_t       *info;        // [0] error
    uint32_t        n_lines, key_bias;
    unsigned long long chan_cap;
};

// one chunk of <= 32 characters of a line: ranks among equal keys, in lane order.  Returns the key (OQ_PAD + lane for an inactive lane).
__device__ __forceinline__ uint32_t oq_rank (uint32_t key, bool act, int lane, uint32_t &peers)
{
    const uint32_t k = act ? key : (uint32_t)OQ_PAD + 32u + lane;
    peers = __match_any_sync (0xffffffffu, k);
    return k;
}

template <int DEMUX>
__global__ void __launch_bounds__(OQ_WARPS * 32) k_oq (const OqVb *vbs)
{
    const OqVb &V = vbs[blockIdx.x];
    __shared__ uint32_t cur[OQ_WARPS][OQ_PAD];        // per warp and channel: characters distributed (count, then cursor)
    __shared__ uint32_t all[OQ_WARPS][OQ_PAD];        // mux: characters of every line, distributed or not (:61-72 counts them all)
    __shared__ uint32_t tot[OQ_PAD], base[OQ_PAD];
    __shared__ uint8_t  s_mono[OQ_PAD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < OQ_WARPS * OQ_PAD; i += OQ_WARPS * 32) { (&cur[0][0])[i] = 0; (&all[0][0])[i] = 0; }
    if (tid < OQ_PAD) s_mono[tid] = (DEMUX && tid < OQ_CH) ? V.mono[tid] : 0;
    __syncthreads ();
    const uint32_t per = (V.n_lines + OQ_WARPS - 1) / OQ_WARPS;
    const uint32_t l0 = min (V.n_lines, (uint32_t)warp * per), l1 = min (V.n_lines, l0 + per);
    bool bad = false;
    // ---- count
    for (uint32_t l = l0; l < l1; l++) {
        const uint32_t len = V.qual_len[l];
        const uint8_t *q = V.txt + V.qual_off[l];
        const bool dist = DEMUX || !V.seq_len || V.seq_len[l] != 0;
        for (uint32_t i0 = 0; i0 < len; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool act = i < len;
            uint32_t key = act ? (uint32_t)q[i] - V.key_bias : 0;
            if (act && key >= (uint32_t)OQ_CH) { bad = true; key = 0; }
            uint32_t peers;
            const uint32_t k = oq_rank (key, act, lane, peers);
            if (act && (peers & ((1u << lane) - 1)) == 0) {                 // the lowest lane of a group of equal keys counts the group
                if (!DEMUX) all[warp][k] += __popc (peers);
                if (dist && !(DEMUX && s_mono[k])) cur[warp][k] += __popc (peers);
            }
            __syncwarp ();
        }
    }
    if (bad) V.info[0] = 1;
    __syncthreads ();
    // ---- channel offsets and the cursors of every warp
    if (tid < OQ_PAD) {
        uint32_t t = 0;
        if (tid < OQ_CH) { if (DEMUX) t = V.count[tid]; else for (int w = 0; w < OQ_WARPS; w++) t += all[w][tid]; }
        tot[tid] = t;
    }
    __syncthreads ();
    if (warp == 0) {                                                        // exclusive scan of 96 totals: 3 per lane
        uint32_t v[3], s = 0;
        for (int k = 0; k < 3; k++) { v[k] = tot[3 * lane + k]; s += v[k]; }
        uint32_t inc = s;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        uint32_t x = inc - s;
        for (int k = 0; k < 3; k++) { base[3 * lane + k] = x; x += v[k]; }
        if (lane == 31 && (unsigned long long)x > V.chan_cap) V.info[0] = 2;
    }
    __syncthreads ();
    if (tid < OQ_CH) {
        uint32_t x = base[tid];
        for (int w = 0; w < OQ_WARPS; w++) { const uint32_t c = cur[w][tid]; cur[w][tid] = x; x += c; }
        if (DEMUX && x > base[tid] + tot[tid]) V.info[0] = 3;               // "channel is out of data" (:152-153)
        if (!DEMUX) V.count[tid] = tot[tid];
    }
    __syncthreads ();
    if (V.info[0]) return;
    // ---- distribute
    for (uint32_t l = l0; l < l1; l++) {
        const uint32_t len = V.qual_len[l];
        if (!DEMUX && V.seq_len && V.seq_len[l] == 0) continue;             // :94
        const uint8_t *q = V.txt + V.qual_off[l];
        const uint8_t *oq = DEMUX ? nullptr : V.txt + V.oq_off[l];          // (dl->OQ = 0 reads the start of the text, as the reference does, :91)
        uint8_t *o = DEMUX ? V.out + V.out_off[l] : nullptr;
        for (uint32_t i0 = 0; i0 < len; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool act = i < len;
            const uint32_t key = act ? (uint32_t)q[i] - V.key_bias : 0;
            uint32_t peers;
            const uint32_t k = oq_rank (key, act, lane, peers);
            const bool mono = DEMUX && act && s_mono[k];
            const uint32_t b = (act && !mono) ? cur[warp][k] : 0;
            __syncwarp ();
            if (act) {
                const uint32_t pos = b + __popc (peers & ((1u << lane) - 1));
                if (DEMUX) o[i] = mono ? s_mono[k] : V.chan[pos];
                else V.chan[pos] = oq[i];
                if (!mono && (peers >> lane) == 1) cur[warp][k] = b + __popc (peers);   // the highest lane of the group moves the cursor
            }
            __syncwarp ();
        }
    }
    if (DEMUX) return;
    // ---- monochar channels (:103-107): str_is_monochar over the channel's count_q bytes (what no line wrote stays 0)
    __syncthreads ();
    for (int k = warp; k < OQ_CH; k += OQ_WARPS) {
        const uint32_t n = tot[k];
        uint8_t m = 0;
        if (n) {
            const uint8_t *c = V.chan + base[k];
            const uint8_t first = c[0];
            bool same = true;
            for (uint32_t i0 = 0; i0 < n && same; i0 += 32) {
                const uint32_t i = i0 + lane;
                same = __all_sync (0xffffffffu, i >= n || c[i] == first);
            }
            m = same ? first : 0;
        }
        if (lane == 0) V.mono[k] = m;
    }
}

struct Carver {
    uint8_t *base; size_t off;
    template <typename T> T *take (size_t count) {
        size_t bytes = (count * sizeof (T) + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

int oq_run (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags, int demux)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    std::vector<OqVb> h (n_vbs);
    std::vector<uint64_t> total (n_vbs, 0), chan_bytes (n_vbs, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_oq_vb &S = vbs[v]; S.status = GZB_OK;
        if ((S.n_lines && (!S.qual_off || !S.qual_len || (demux ? (!S.out_off || !S.out) : !S.oq_off))) || (!S.txt && S.txt_len) || (!S.channels && S.channels_cap)) return GZB_E_BADARG;
        if (!devptr) for (uint32_t i = 0; i < S.n_lines; i++) total[v] += S.qual_len[i];
        else total[v] = demux ? S.out_cap : S.channels_cap;
        if (demux) for (int k = 0; k < OQ_CH; k++) chan_bytes[v] += S.count[k];
        else chan_bytes[v] = devptr ? S.channels_cap : total[v];
        if (chan_bytes[v] > S.channels_cap || (demux && !devptr && total[v] > S.out_cap)) { e->err = "OQ: a buffer is too small"; return GZB_E_BADARG; }
        if (chan_bytes[v] > 0xffffffffull) { e->err = "OQ: more than 4 GB of channels in a VBlock"; return GZB_E_BADARG; }
    }
    Carver c { nullptr, 0 };
    OqVb *d_vbs = nullptr; uint32_t *d_info = nullptr, *d_count = nullptr; uint8_t *d_mono = nullptr;
    const size_t desc_bytes = ((size_t)n_vbs * sizeof (OqVb) + 255) & ~(size_t)255, meta_bytes = (size_t)n_vbs * (OQ_PAD * 4 + OQ_PAD + 16);
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<OqVb> (n_vbs); d_info = c.take<uint32_t> ((size_t)n_vbs * 4);
        d_count = c.take<uint32_t> ((size_t)n_vbs * OQ_PAD); d_mono = c.take<uint8_t> ((size_t)n_vbs * OQ_PAD);
        for (uint32_t v = 0; v < n_vbs; v++) {
            const gzb_oq_vb &S = vbs[v]; OqVb &D = h[v];
            D.n_lines = S.n_lines; D.key_bias = demux ? S.key_bias : 33u; D.chan_cap = S.channels_cap;
            D.info = d_info ? d_info + 4 * (size_t)v : nullptr;
            D.count = d_count ? d_count + (size_t)OQ_PAD * v : nullptr; D.mono = d_mono ? d_mono + (size_t)OQ_PAD * v : nullptr;
            D.txt      = devptr ? (const uint8_t *)S.txt : c.take<uint8_t> (S.txt_len + 16);
            D.qual_off = devptr ? S.qual_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.qual_len = devptr ? S.qual_len : c.take<uint32_t> ((size_t)S.n_lines + 1);
            D.oq_off   = demux ? nullptr : devptr ? S.oq_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.seq_len  = (demux || !S.seq_len) ? nullptr : devptr ? S.seq_len : c.take<uint32_t> ((size_t)S.n_lines + 1);
            D.out_off  = !demux ? nullptr : devptr ? S.out_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.out      = !demux ? nullptr : devptr ? (uint8_t *)S.out : c.take<uint8_t> (S.out_cap + 16);
            D.chan     = devptr ? (uint8_t *)S.channels : c.take<uint8_t> (chan_bytes[v] + 16);
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, desc_bytes + meta_bytes + 512); if (rc) return rc; c.base = e->ws; }
    }
    uint32_t *p_count = reinterpret_cast<uint32_t *>(e->pin + desc_bytes);
    uint8_t  *p_mono  = e->pin + desc_bytes + (size_t)n_vbs * OQ_PAD * 4;
    uint32_t *p_info  = reinterpret_cast<uint32_t *>(e->pin + desc_bytes + (size_t)n_vbs * (OQ_PAD * 4 + OQ_PAD));
    for (uint32_t v = 0; v < n_vbs; v++) {
        const gzb_oq_vb &S = vbs[v]; OqVb &D = h[v];
        memset (p_count + (size_t)OQ_PAD * v, 0, OQ_PAD * 4); memset (p_mono + (size_t)OQ_PAD * v, 0, OQ_PAD);
        if (demux) { memcpy (p_count + (size_t)OQ_PAD * v, S.count, OQ_CH * 4); memcpy (p_mono + (size_t)OQ_PAD * v, S.monochars, OQ_CH); }
        if (devptr) { if (!demux && chan_bytes[v]) CK (cudaMemsetAsync (D.chan, 0, chan_bytes[v], st)); continue; }
        if (S.n_lines) {
            CK (cudaMemcpyAsync ((void *)D.qual_off, S.qual_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
            CK (cudaMemcpyAsync ((void *)D.qual_len, S.qual_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
            if (D.oq_off)  CK (cudaMemcpyAsync ((void *)D.oq_off, S.oq_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
            if (D.seq_len) CK (cudaMemcpyAsync ((void *)D.seq_len, S.seq_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
            if (D.out_off) CK (cudaMemcpyAsync ((void *)D.out_off, S.out_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
        }
        if (S.txt_len) CK (cudaMemcpyAsync ((void *)D.txt, S.txt, S.txt_len, cudaMemcpyHostToDevice, st));
        if (demux) {
            if (chan_bytes[v]) CK (cudaMemcpyAsync (D.chan, S.channels, chan_bytes[v], cudaMemcpyHostToDevice, st));
            if (S.out_cap) CK (cudaMemsetAsync (D.out, 0, S.out_cap, st));     // (host buffers: the whole of `out` comes back, zero where no line lands)
        }
        else if (chan_bytes[v]) CK (cudaMemsetAsync (D.chan, 0, chan_bytes[v], st));
    }
    memcpy (e->pin, h.data (), (size_t)n_vbs * sizeof (OqVb));              // descriptors through the pinned staging (stage.cu says why)
    CK (cudaMemcpyAsync (d_vbs, e->pin, (size_t)n_vbs * sizeof (OqVb), cudaMemcpyHostToDevice, st));
    CK (cudaMemcpyAsync (d_count, p_count, (size_t)n_vbs * OQ_PAD * 4, cudaMemcpyHostToDevice, st));
    CK (cudaMemcpyAsync (d_mono, p_mono, (size_t)n_vbs * OQ_PAD, cudaMemcpyHostToDevice, st));
    CK (cudaMemsetAsync (d_info, 0, (size_t)n_vbs * 16, st));
    if (demux) k_oq<1><<<n_vbs, OQ_WARPS * 32, 0, st>>>(d_vbs); else k_oq<0><<<n_vbs, OQ_WARPS * 32, 0, st>>>(d_vbs);
    e->launches++;
    CK (cudaMemcpyAsync (p_info, d_info, (size_t)n_vbs * 16, cudaMemcpyDeviceToHost, st));
    if (!demux) {
        CK (cudaMemcpyAsync (p_count, d_count, (size_t)n_vbs * OQ_PAD * 4, cudaMemcpyDeviceToHost, st));
        CK (cudaMemcpyAsync (p_mono, d_mono, (size_t)n_vbs * OQ_PAD, cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    int rc = GZB_OK;
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_oq_vb &S = vbs[v];
        const uint32_t err = p_info[4 * v];
        if (err) {
            S.status = err == 2 ? GZB_E_BADARG : GZB_E_CORRUPT; rc = S.status;
            e->err = err == 1 ? "OQ: a QUAL character outside '!'..'~'" : err == 2 ? "OQ: the channel buffer is too small" : "OQ: a channel is out of data";
            continue;
        }
        if (!demux) {
            memcpy (S.count, p_count + (size_t)OQ_PAD * v, OQ_CH * 4); memcpy (S.monochars, p_mono + (size_t)OQ_PAD * v, OQ_CH);
            uint64_t nb = 0; for (int k = 0; k < OQ_CH; k++) nb += S.count[k];
            if (!devptr && nb) CK (cudaMemcpyAsync (S.channels, h[v].chan, nb, cudaMemcpyDeviceToHost, st));
        }
        else if (!devptr && S.out_cap) CK (cudaMemcpyAsync (S.out, h[v].out, S.out_cap, cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    return rc;
}

} // namespace

extern "C" int gzb_oq_mux   (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags) { return oq_run (e, vbs, n_vbs, flags, 0); }
extern "C" int gzb_oq_demux (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags) { return oq_run (e, vbs, n_vbs, flags, 1); }
