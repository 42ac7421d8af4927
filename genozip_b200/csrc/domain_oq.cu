// domain_oq.cu — the OQ codec of SAM / BAM (src/codec_oq.c): the original quality string OQ:Z of a read is multiplexed by the read's
// QUAL — the OQ character at position i goes to channel QUAL[i] - '!' (94 channels, each then an ordinary local section) — because
// recalibration maps almost every (OQ, context) to one QUAL, so a channel is nearly constant.  A channel that holds one character
// only is dropped and the character kept in the 94-byte OQ:Z.local (:103-107).
//
//   mux    codec_oq_compress before its sub-codec (:54-121): count pass over the QUAL of ALL lines (:61-72), mux pass over the lines
//          whose SEQ.len is not 0 (:88-99), monochar test (:103-107)
//   demux  codec_oq_reconstruct (:126-164) for every line of a VBlock at once
//
// SMUX (src/codec_smux.c), the quality codec of MGI reads, is the same operation with other keys: QUAL[i] goes to the channel of the BASE at
// position i (A, C, G, T, anything else: 5 channels), a reverse-complemented read walked backwards with complemented bases (:217-239); only the
// fifth channel is tested for monochar (its character then travels in the section header's param, :246-253).
// TMPL (src/codec_tmpl.c), the quality codec of Element reads, again: QUAL[i] goes to the channel of TEMPLATE[i] - '!', the template being
// the most frequent quality of every read position (found once, in segconf); what a read has beyond the template goes to one more stream.
//
// PACB (src/codec_pacb.c), the quality codec of PacBio reads: the channel of QUAL[i] is 7 * (min (np, max_np) - 1) + K, np the read's number of
// passes and K one of 7 classes of the base's surroundings (inside / at the start of a homopolymer of 1, 2, 3+ bases, A/T or C/G, :19-27).
//
// All four are a STABLE distribution by key, the shape of arith_split.cu's bucket kernel: one CTA per VBlock, 32 warps owning 32
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
    const uint64_t *a_off;       // OQ: the QUAL string (the keys).  SMUX: the QUAL string (the values; mux only)
    const uint32_t *a_len;       //     its length.                  SMUX mux: QUAL's length; demux: the `len` of the reconstruct call
    const uint64_t *b_off;       // OQ mux: the OQ string (dl->OQ).   SMUX: the SEQ string (the keys)
    const uint32_t *b_len;       // OQ mux: SEQ.len or NULL (0 = the line is not distributed, :94).  SMUX mux: SEQ's length
    const uint8_t  *is_rev;      // SMUX, may be NULL
    uint8_t        *out;         // demux
    const uint64_t *out_off;     // demux
    uint8_t        *chan;
    uint32_t       *count;       // [94]  mux: out; demux: in
    uint8_t        *mono;        // [94]  mux: out; demux: in
    uint32_t       *info;        // [0] error
    uint32_t        n_lines, key_bias, kind, n_ch;   // kind 0 OQ (94 channels), 1 SMUX (5), 2 TMPL (94 + the excess stream)
    unsigned long long chan_cap;
    const uint8_t  *tmpl; uint32_t tmpl_len;         // TMPL
    const uint8_t  *np0;                             // PACB: per line min (np, max_np) - 1, or NULL (FASTQ, CLR: always 0)
};

__device__ __forceinline__ uint32_t pacb_K (const uint8_t *seq, uint32_t len, uint32_t i)      // QUAL_get_K_value (src/codec_pacb.c:19-27)
{
    const uint8_t b = seq[i];
    const uint32_t at = (b == 'A' || b == 'T') ? 1u : 0u;
    if (i > 0 && seq[i - 1] == b) return 6;
    if (i == len - 1 || seq[i + 1] != b) return 4 + at;
    if (i == len - 2 || seq[i + 2] != b) return 2 + at;
    return at;
}
__device__ __forceinline__ uint32_t smux_enc (uint8_t c, bool comp)        // _nuke_encode / _nuke_encode_comp (src/reference.c:78-84)
{
    const uint32_t k = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
    return (comp && k < 4) ? 3u - k : k;
}

// One line as the kernel walks it: n elements in the order the reference visits them; element t has the key key (t), its value is read from
// (mux) or written to (demux) position at (t) of the line's string.
struct OqLine {
    const uint8_t *keys; const uint8_t *vals; uint8_t *dst;
    uint32_t n, rev, last_key_only, bias, kind, tmpl_len;
    bool dist;
    __device__ __forceinline__ uint32_t at (uint32_t t) const { return rev ? n - 1 - t : t; }
    __device__ __forceinline__ uint32_t key (uint32_t t) const
    {
        if (kind == 0) return (uint32_t)keys[t] - bias;
        if (kind == 2) return t < tmpl_len ? (uint32_t)keys[t] - 33u : 94u;      // keys = the template (codec_tmpl.c:171-178); 94 = the excess stream
        if (kind == 3) return bias + pacb_K (keys, n, t);                       // bias = 7 * np0 of the line (codec_pacb.c:138-140)
        return last_key_only ? smux_enc (keys[last_key_only - 1], true) : smux_enc (keys[at (t)], rev);
    }
};
template <int DEMUX> __device__ __forceinline__ OqLine oq_line (const OqVb &V, uint32_t l)
{
    OqLine L; L.kind = V.kind; L.bias = V.key_bias; L.rev = 0; L.last_key_only = 0; L.dist = true; L.dst = nullptr; L.vals = nullptr; L.tmpl_len = V.tmpl_len;
    if (V.kind == 3) {
        L.n = V.a_len[l]; L.keys = V.txt + V.b_off[l]; L.bias = V.np0 ? 7u * V.np0[l] : 0;
        if (DEMUX) L.dst = V.out + V.out_off[l]; else L.vals = V.txt + V.a_off[l];
        return L;
    }
    if (V.kind == 2) {
        L.n = V.a_len[l]; L.keys = V.tmpl;
        if (DEMUX) L.dst = V.out + V.out_off[l]; else L.vals = V.txt + V.a_off[l];
        return L;
    }
    if (V.kind == 0) {
        L.n = V.a_len[l]; L.keys = V.txt + V.a_off[l];
        if (DEMUX) L.dst = V.out + V.out_off[l];
        else { L.vals = V.txt + V.b_off[l]; L.dist = !V.b_len || V.b_len[l] != 0; }   // (dl->OQ = 0 reads the start of the text, as the reference does, codec_oq.c:91)
        return L;
    }
    L.keys = V.txt + V.b_off[l]; L.rev = V.is_rev ? V.is_rev[l] : 0;
    if (DEMUX) { L.n = V.a_len[l]; if (L.n && L.keys[0] == '*') L.n = 1; L.dst = V.out + V.out_off[l]; return L; }      // codec_smux.c:278-279
    L.n = V.a_len[l]; L.vals = V.txt + V.a_off[l];
    if (L.n == 1 && L.rev && L.vals[0] == ' ') { L.last_key_only = V.b_len[l]; L.rev = 0; }   // a reversed read without quality: the channel of its LAST base (:207-208,229-232)
    return L;
}

template <int DEMUX>
__global__ void __launch_bounds__(OQ_WARPS * 32) k_oq (const OqVb *vbs)
{
    const OqVb &V = vbs[blockIdx.x];
    __shared__ uint32_t cur[OQ_WARPS][OQ_PAD];        // per warp and channel: characters distributed (count, then cursor)
    __shared__ uint32_t all[OQ_WARPS][OQ_PAD];        // mux: characters of every line, distributed or not (codec_oq.c:61-72 counts them all)
    __shared__ uint32_t tot[OQ_PAD], base[OQ_PAD];
    __shared__ uint8_t  s_mono[OQ_PAD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_ch = V.n_ch;
    for (int i = tid; i < OQ_WARPS * OQ_PAD; i += OQ_WARPS * 32) { (&cur[0][0])[i] = 0; (&all[0][0])[i] = 0; }
    if (tid < OQ_PAD) s_mono[tid] = (DEMUX && (uint32_t)tid < n_ch) ? V.mono[tid] : 0;
    __syncthreads ();
    const uint32_t per = (V.n_lines + OQ_WARPS - 1) / OQ_WARPS;
    const uint32_t l0 = min (V.n_lines, (uint32_t)warp * per), l1 = min (V.n_lines, l0 + per);
    bool bad = false;
    // ---- count
    for (uint32_t l = l0; l < l1; l++) {
        const OqLine L = oq_line<DEMUX> (V, l);
        for (uint32_t t0 = 0; t0 < L.n; t0 += 32) {
            const uint32_t t = t0 + lane;
            const bool act = t < L.n;
            uint32_t key = act ? L.key (t) : 0;
            if (act && key >= n_ch) { bad = true; key = 0; }
            const uint32_t k = act ? key : (uint32_t)OQ_PAD + 32u + lane;
            const uint32_t peers = __match_any_sync (0xffffffffu, k);
            if (act && (peers & ((1u << lane) - 1)) == 0) {                 // the lowest lane of a group of equal keys counts the group
                if (!DEMUX) all[warp][k] += __popc (peers);
                if (L.dist && !(DEMUX && s_mono[k])) cur[warp][k] += __popc (peers);
            }
            __syncwarp ();
        }
    }
    if (bad) V.info[0] = 1;
    __syncthreads ();
    // ---- channel offsets and the cursors of every warp
    if (tid < OQ_PAD) {
        uint32_t t = 0;
        if ((uint32_t)tid < n_ch) { if (DEMUX) t = V.count[tid]; else for (int w = 0; w < OQ_WARPS; w++) t += all[w][tid]; }
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
    if ((uint32_t)tid < n_ch) {
        uint32_t x = base[tid];
        for (int w = 0; w < OQ_WARPS; w++) { const uint32_t c = cur[w][tid]; cur[w][tid] = x; x += c; }
        if (DEMUX && x > base[tid] + tot[tid]) V.info[0] = 3;               // "channel is out of data" (codec_oq.c:152-153, codec_smux.c:301)
        if (!DEMUX) V.count[tid] = tot[tid];
    }
    __syncthreads ();
    if (V.info[0]) return;
    // ---- distribute
    bool blank = false;
    for (uint32_t l = l0; l < l1; l++) {
        const OqLine L = oq_line<DEMUX> (V, l);
        if (!L.dist) continue;                                              // codec_oq.c:94
        for (uint32_t t0 = 0; t0 < L.n; t0 += 32) {
            const uint32_t t = t0 + lane;
            const bool act = t < L.n;
            const uint32_t key = act ? L.key (t) : 0;
            const uint32_t k = act ? key : (uint32_t)OQ_PAD + 32u + lane;
            const uint32_t peers = __match_any_sync (0xffffffffu, k);
            const bool mono = DEMUX && act && s_mono[k];
            const uint32_t b = (act && !mono) ? cur[warp][k] : 0;
            __syncwarp ();
            if (act) {
                const uint32_t pos = b + __popc (peers & ((1u << lane) - 1));
                if (DEMUX) { const uint8_t c = mono ? s_mono[k] : V.chan[pos]; L.dst[L.at (t)] = c; if ((V.kind == 1 || V.kind == 3) && c == ' ') blank = true; }
                else V.chan[pos] = L.last_key_only ? (uint8_t)' ' : L.vals[L.at (t)];
                if (!mono && (peers >> lane) == 1) cur[warp][k] = b + __popc (peers);   // the highest lane of the group moves the cursor
            }
            __syncwarp ();
        }
    }
    if (blank) V.info[0] = 4;                                               // a read without quality (codec_smux.c:307-310): how much a line consumes then depends on the data
    if (DEMUX) return;
    // ---- monochar channels (codec_oq.c:103-107; SMUX: the fifth channel only, codec_smux.c:246): str_is_monochar over the channel's bytes (what no line wrote stays 0)
    __syncthreads ();
    for (uint32_t k = warp; k < n_ch; k += OQ_WARPS) {
        const uint32_t n = tot[k];
        uint8_t m = 0;
        if (n && (V.kind == 0 || (V.kind == 1 && k == 4))) {
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

// what both public structs boil down to
struct OqHost {
    const void *txt; uint64_t txt_len;
    const uint64_t *a_off; const uint32_t *a_len; const uint64_t *b_off; const uint32_t *b_len; const uint8_t *is_rev;
    uint32_t n_lines, key_bias, kind, n_ch;
    void *channels; uint64_t channels_cap; uint32_t *count; uint8_t *mono;
    void *out; uint64_t out_cap; const uint64_t *out_off;
    int32_t *status;
    const uint8_t *tmpl = nullptr; uint32_t tmpl_len = 0;                    // TMPL: host memory always
    const uint8_t *np0 = nullptr;                                            // PACB
};

int oq_run (gzb_engine *e, std::vector<OqHost> &vbs, uint32_t flags, int demux)
{
    const uint32_t n_vbs = (uint32_t)vbs.size ();
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    std::vector<OqVb> h (n_vbs);
    std::vector<uint64_t> total (n_vbs, 0), chan_bytes (n_vbs, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        OqHost &S = vbs[v]; *S.status = GZB_OK;
        const bool need_a_off = S.kind == 0 || !demux, need_b_off = S.kind == 1 || S.kind == 3 || (S.kind == 0 && !demux);
        if (S.kind == 2 && S.tmpl_len && !S.tmpl) return GZB_E_BADARG;
        if ((S.n_lines && (!S.a_len || (need_a_off && !S.a_off) || (need_b_off && !S.b_off) || (S.kind == 1 && !demux && !S.b_len) || (demux && (!S.out_off || !S.out)))) ||
            (!S.txt && S.txt_len) || (!S.channels && S.channels_cap)) return GZB_E_BADARG;
        if (!devptr) for (uint32_t i = 0; i < S.n_lines; i++) total[v] += S.a_len[i];
        else total[v] = demux ? S.out_cap : S.channels_cap;
        if (demux) for (uint32_t k = 0; k < S.n_ch; k++) chan_bytes[v] += S.count[k];
        else chan_bytes[v] = devptr ? S.channels_cap : total[v];
        if (chan_bytes[v] > S.channels_cap || (demux && !devptr && total[v] > S.out_cap)) { e->err = "OQ / SMUX: a buffer is too small"; return GZB_E_BADARG; }
        if (chan_bytes[v] > 0xffffffffull) { e->err = "OQ / SMUX: more than 4 GB of channels in a VBlock"; return GZB_E_BADARG; }
    }
    Carver c { nullptr, 0 };
    OqVb *d_vbs = nullptr; uint32_t *d_info = nullptr, *d_count = nullptr; uint8_t *d_mono = nullptr;
    const size_t desc_bytes = ((size_t)n_vbs * sizeof (OqVb) + 255) & ~(size_t)255, meta_bytes = (size_t)n_vbs * (OQ_PAD * 4 + OQ_PAD + 16);
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<OqVb> (n_vbs); d_info = c.take<uint32_t> ((size_t)n_vbs * 4);
        d_count = c.take<uint32_t> ((size_t)n_vbs * OQ_PAD); d_mono = c.take<uint8_t> ((size_t)n_vbs * OQ_PAD);
        for (uint32_t v = 0; v < n_vbs; v++) {
            const OqHost &S = vbs[v]; OqVb &D = h[v];
            D.n_lines = S.n_lines; D.key_bias = S.key_bias; D.chan_cap = S.channels_cap; D.kind = S.kind; D.n_ch = S.n_ch;
            D.info = d_info ? d_info + 4 * (size_t)v : nullptr;
            D.count = d_count ? d_count + (size_t)OQ_PAD * v : nullptr; D.mono = d_mono ? d_mono + (size_t)OQ_PAD * v : nullptr;
            D.txt     = devptr ? (const uint8_t *)S.txt : c.take<uint8_t> (S.txt_len + 16);
            D.a_len   = devptr ? S.a_len : c.take<uint32_t> ((size_t)S.n_lines + 1);
            D.a_off   = !S.a_off ? nullptr : devptr ? S.a_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.b_off   = !S.b_off ? nullptr : devptr ? S.b_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.b_len   = !S.b_len ? nullptr : devptr ? S.b_len : c.take<uint32_t> ((size_t)S.n_lines + 1);
            D.is_rev  = !S.is_rev ? nullptr : devptr ? S.is_rev : c.take<uint8_t> ((size_t)S.n_lines + 1);
            D.out_off = !demux ? nullptr : devptr ? S.out_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.out     = !demux ? nullptr : devptr ? (uint8_t *)S.out : c.take<uint8_t> (S.out_cap + 16);
            D.chan    = devptr ? (uint8_t *)S.channels : c.take<uint8_t> (chan_bytes[v] + 16);
            D.tmpl    = S.kind == 2 ? c.take<uint8_t> ((size_t)S.tmpl_len + 16) : nullptr; D.tmpl_len = S.tmpl_len;
            D.np0     = !S.np0 ? nullptr : devptr ? S.np0 : c.take<uint8_t> ((size_t)S.n_lines + 1);
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, desc_bytes + meta_bytes + 512); if (rc) return rc; c.base = e->ws; }
    }
    uint32_t *p_count = reinterpret_cast<uint32_t *>(e->pin + desc_bytes);
    uint8_t  *p_mono  = e->pin + desc_bytes + (size_t)n_vbs * OQ_PAD * 4;
    uint32_t *p_info  = reinterpret_cast<uint32_t *>(e->pin + desc_bytes + (size_t)n_vbs * (OQ_PAD * 4 + OQ_PAD));
    for (uint32_t v = 0; v < n_vbs; v++) {
        const OqHost &S = vbs[v]; OqVb &D = h[v];
        memset (p_count + (size_t)OQ_PAD * v, 0, OQ_PAD * 4); memset (p_mono + (size_t)OQ_PAD * v, 0, OQ_PAD);
        if (demux) { memcpy (p_count + (size_t)OQ_PAD * v, S.count, S.n_ch * 4); memcpy (p_mono + (size_t)OQ_PAD * v, S.mono, S.n_ch); }
        if (S.kind == 2 && S.tmpl_len) CK (cudaMemcpyAsync ((void *)D.tmpl, S.tmpl, S.tmpl_len, cudaMemcpyHostToDevice, st));   // (pageable, tiny)
        if (devptr) { if (!demux && chan_bytes[v]) CK (cudaMemsetAsync (D.chan, 0, chan_bytes[v], st)); continue; }
        if (S.n_lines) {
            CK (cudaMemcpyAsync ((void *)D.a_len, S.a_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
            if (D.a_off)   CK (cudaMemcpyAsync ((void *)D.a_off, S.a_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
            if (D.b_off)   CK (cudaMemcpyAsync ((void *)D.b_off, S.b_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
            if (D.b_len)   CK (cudaMemcpyAsync ((void *)D.b_len, S.b_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
            if (D.is_rev)  CK (cudaMemcpyAsync ((void *)D.is_rev, S.is_rev, S.n_lines, cudaMemcpyHostToDevice, st));
            if (D.np0)     CK (cudaMemcpyAsync ((void *)D.np0, S.np0, S.n_lines, cudaMemcpyHostToDevice, st));
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
        OqHost &S = vbs[v];
        const uint32_t err = p_info[4 * v];
        if (err) {
            *S.status = err == 2 ? GZB_E_BADARG : err == 4 ? GZB_E_UNSUPPORTED : GZB_E_CORRUPT; rc = *S.status;
            e->err = err == 1 ? "OQ / SMUX: a key outside the channels" : err == 2 ? "OQ / SMUX: the channel buffer is too small" : err == 3 ? "OQ / SMUX: a channel is out of data"
                   : "SMUX / PACB: a read without quality — how much such a line consumes depends on the data: reconstruct these VBlocks line by line";
            continue;
        }
        if (!demux) {
            memcpy (S.count, p_count + (size_t)OQ_PAD * v, S.n_ch * 4); memcpy (S.mono, p_mono + (size_t)OQ_PAD * v, S.n_ch);
            uint64_t nb = 0; for (uint32_t k = 0; k < S.n_ch; k++) nb += S.count[k];
            if (!devptr && nb) CK (cudaMemcpyAsync (S.channels, h[v].chan, nb, cudaMemcpyDeviceToHost, st));
        }
        else if (!devptr && S.out_cap) CK (cudaMemcpyAsync (S.out, h[v].out, S.out_cap, cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    return rc;
}

int oq_public (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags, int demux)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    std::vector<OqHost> h (n_vbs);
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_oq_vb &S = vbs[v];
        h[v] = OqHost { S.txt, S.txt_len, S.qual_off, S.qual_len, demux ? nullptr : S.oq_off, demux ? nullptr : S.seq_len, nullptr, S.n_lines, demux ? S.key_bias : 33u, 0, (uint32_t)OQ_CH,
                        S.channels, S.channels_cap, S.count, S.monochars, S.out, S.out_cap, S.out_off, &S.status };
    }
    return oq_run (e, h, flags, demux);
}
int smux_public (gzb_engine *e, gzb_smux_vb *vbs, uint32_t n_vbs, uint32_t flags, int demux)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    std::vector<OqHost> h (n_vbs);
    std::vector<uint8_t> mono ((size_t)n_vbs * 8, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_smux_vb &S = vbs[v];
        if (demux) mono[8 * (size_t)v + 4] = S.n_param;
        h[v] = OqHost { S.txt, S.txt_len, demux ? nullptr : S.qual_off, S.qual_len, S.seq_off, demux ? nullptr : S.seq_len, S.is_rev, S.n_lines, 0, 1, 5,
                        S.channels, S.channels_cap, S.count, &mono[8 * (size_t)v], S.out, S.out_cap, S.out_off, &S.status };
    }
    const int rc = oq_run (e, h, flags, demux);
    if (!demux) for (uint32_t v = 0; v < n_vbs; v++) vbs[v].n_param = mono[8 * (size_t)v + 4];
    return rc;
}

int tmpl_public (gzb_engine *e, gzb_tmpl_vb *vbs, uint32_t n_vbs, uint32_t flags, int demux)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    std::vector<OqHost> h (n_vbs);
    std::vector<uint8_t> mono ((size_t)n_vbs * OQ_PAD, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_tmpl_vb &S = vbs[v];
        h[v] = OqHost { S.txt, S.txt_len, demux ? nullptr : S.qual_off, S.qual_len, nullptr, nullptr, nullptr, S.n_lines, 33u, 2, 95,
                        S.channels, S.channels_cap, S.count, &mono[(size_t)OQ_PAD * v], S.out, S.out_cap, S.out_off, &S.status };
        h[v].tmpl = (const uint8_t *)S.tmpl; h[v].tmpl_len = S.tmpl_len;
        for (uint32_t i = 0; i < S.tmpl_len; i++) if (h[v].tmpl[i] < 33 || h[v].tmpl[i] > 126) { e->err = "TMPL: a template character outside '!'..'~'"; return GZB_E_BADARG; }
    }
    return oq_run (e, h, flags, demux);
}

int pacb_public (gzb_engine *e, gzb_pacb_vb *vbs, uint32_t n_vbs, uint32_t flags, int demux)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    std::vector<OqHost> h (n_vbs);
    std::vector<uint8_t> mono ((size_t)n_vbs * OQ_PAD, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_pacb_vb &S = vbs[v];
        if (S.max_np < 1 || S.max_np > 12 || (S.max_np > 1 && S.n_lines && !S.np0)) { e->err = "PACB: max_np must be 1..12 (MAX_np), np0 given when it is above 1"; return GZB_E_BADARG; }
        h[v] = OqHost { S.txt, S.txt_len, demux ? nullptr : S.qual_off, S.qual_len, S.seq_off, nullptr, nullptr, S.n_lines, 0, 3, 7u * S.max_np,
                        S.channels, S.channels_cap, S.count, &mono[(size_t)OQ_PAD * v], S.out, S.out_cap, S.out_off, &S.status };
        h[v].np0 = S.max_np > 1 ? S.np0 : nullptr;
    }
    return oq_run (e, h, flags, demux);
}

} // namespace

extern "C" int gzb_pacb_mux   (gzb_engine *e, gzb_pacb_vb *vbs, uint32_t n_vbs, uint32_t flags) { return pacb_public (e, vbs, n_vbs, flags, 0); }
extern "C" int gzb_pacb_demux (gzb_engine *e, gzb_pacb_vb *vbs, uint32_t n_vbs, uint32_t flags) { return pacb_public (e, vbs, n_vbs, flags, 1); }
extern "C" int gzb_tmpl_mux   (gzb_engine *e, gzb_tmpl_vb *vbs, uint32_t n_vbs, uint32_t flags) { return tmpl_public (e, vbs, n_vbs, flags, 0); }
extern "C" int gzb_tmpl_demux (gzb_engine *e, gzb_tmpl_vb *vbs, uint32_t n_vbs, uint32_t flags) { return tmpl_public (e, vbs, n_vbs, flags, 1); }
extern "C" int gzb_oq_mux     (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags)   { return oq_public (e, vbs, n_vbs, flags, 0); }
extern "C" int gzb_oq_demux   (gzb_engine *e, gzb_oq_vb *vbs, uint32_t n_vbs, uint32_t flags)   { return oq_public (e, vbs, n_vbs, flags, 1); }
extern "C" int gzb_smux_mux   (gzb_engine *e, gzb_smux_vb *vbs, uint32_t n_vbs, uint32_t flags) { return smux_public (e, vbs, n_vbs, flags, 0); }
extern "C" int gzb_smux_demux (gzb_engine *e, gzb_smux_vb *vbs, uint32_t n_vbs, uint32_t flags) { return smux_public (e, vbs, n_vbs, flags, 1); }
