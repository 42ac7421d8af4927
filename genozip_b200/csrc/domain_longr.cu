// domain_longr.cu — LONGR long-read quality model on sm_100a.
//
// Reference functions replaced (relative to /root/reference/src): codec_longr_compress before its sub-codec
// (codec_longr.c:161-247: codec_longr_calc_channels :138-159, counting sort :193-230, lens :232-240),
// codec_longr_reconstruct for all reads of a VBlock (:270-373), the state machine codec_longr_update_state /
// _alg_init / _alg_init_read (codec_longr_alg.c:108-159), codec_longr_segconf_calculate_bins (codec_longr.c:66-136).
//
// What the algorithm is, seen from a GPU.  Every base is one EVENT: a read-modify-write of two tables at a context that the
// INPUT alone determines — u = (last six bases : 12 | capped interlaced difference of the two previous qualities : 4 |
// bin of the previous quality : 5), and Q = the upper 9 bits of u:
//      (avg, err) = st[u];  t = tot[Q]                      the base's channel = (u >> 12) | bin(avg) << 9 | class(err, t) << 14
//      st[u] <- update (avg, err, q);  tot[Q] <- update (t, |q - avg|)
// (the three updates codec_longr_alg_init_read makes at the start of a read are events without a channel).  The tables are
// carried from read to read through the VBlock, so a VBlock is one serial stream of events; VBlocks are independent.
//
//   encode  k_longr_channels: ONE WARP PER VBLOCK takes 32 consecutive events at a time.  The contexts of all 32 are known up
//           front, so their table entries are fetched together; events of a tile that share a context (or a Q) are chained in
//           order inside the warp (__match_any_sync groups, one round per member), everything else is parallel.  The loads of the
//           next tile are issued as soon as this tile's stores are, and overlap the rest of its work.  st is one 32-bit word per
//           context (avg | err << 16), tot lives in shared memory.  The stable counting sort by channel (:193-230) is
//           k_longr_prefix + k_longr_place (one warp per VBlock, ranks inside a tile from __match_any_sync, one atomic per
//           channel and tile, consumed one tile later).
//   decode  k_longr_decode: the next quality is only known once the previous base's channel is, so a VBlock decodes one base
//           at a time — lane 0 of one warp per VBlock; the other 31 lanes ask L2 for the table words of the contexts the
//           step after next may land in (one candidate quality each).  The chain per base is two dependent memory round trips (st[u], then
//           the channel's cursor) instead of the reference's four; the cursor of a channel is one 64-bit word that carries the
//           index AND the next four qualities of the channel (refilled off the critical path).  Throughput comes from the
//           number of VBlocks in flight.
#include <cstring>
#include <vector>
#include <string>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "gzb_tma.cuh"
#include "engine.h"

using namespace gzb;

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr uint32_t NCTX = 1u << 21, NCHAN = 1u << 16, NQ9 = 1u << 9;
constexpr uint32_t FULL = 0xffffffffu;
// k_longr_decode: the idle lanes ask L2 for the table words of the contexts the step after next may land in.  Measured on B200
// (1 184 VBlocks of 2 M bases): piz 2.60 GB/s without, 2.34 GB/s with the hints, and 2.8 KB of DRAM reads per base (ncu,
// profiles/r02_pbwt_longr.md: 412 GB for 148 M bases) — the tables of a thousand VBlocks are 9.5 GB, a hinted word is evicted
// before the walk gets to it.  Off.
constexpr bool LR_L2_HINTS = false;

struct LrVb {
    const uint8_t *txt; const uint64_t *seq_off, *qual_off; const uint32_t *len; const uint8_t *is_rev;
    const uint32_t *qlen;                // encode: quality length where it differs from len (a line without quality: ' '), or nullptr
    uint32_t n_lines;
    uint32_t *st;                        // [1<<21] avg_sums | err_sums << 16 (codec_longr_alg.c:99-100)
    uint32_t *chan_num;                  // [65536] bases per channel, then the exclusive prefix next_of_chan
    unsigned long long *cur;             // decode: [65536] cursor of each channel: next index << 32 | up to four next values
    uint16_t *base_chan;                 // encode: channel of each base, in processing order
    uint8_t  *values; uint32_t *lens_be; uint8_t *qual_out; uint8_t *missing;
    uint32_t *err;                       // != 0: damaged LENS / VALUES
    uint64_t  total;
    uint8_t   v2b[256];
};

__device__ __forceinline__ uint32_t acgt_code (uint32_t c)                  // _acgt_encode (reference.c:45-58)
{
    switch (c) {
        case 'C': case 'c': case 'Y': case 'y': case 'S': case 's': case 'B': case 'b': return 1;
        case 'G': case 'g': case 'K': case 'k': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 0;
    }
}
__device__ __forceinline__ uint32_t acgt_code_comp (uint32_t c)             // _acgt_encode_comp (reference.c:63-75)
{
    switch (c) {
        case 'A': case 'a': return 3;
        case 'C': case 'c': case 'M': case 'm': return 2;
        case 'G': case 'g': case 'R': case 'r': case 'S': case 's': case 'V': case 'v': return 1;
        default: return 0;
    }
}

// the three table updates of one event (codec_longr_update_state :112-118); returns the new st word
__device__ __forceinline__ uint32_t lr_st_update (uint32_t st, int32_t q, uint32_t &abs_err)
{
    const int32_t avg = (int32_t)(st & 0xffffu), er = (int32_t)(st >> 16);
    const int32_t d = q - ((avg + 8) >> 4);
    abs_err = (uint32_t)(d < 0 ? -d : d);
    return ((uint32_t)(avg + d) & 0xffffu) | (((uint32_t)(er + (int32_t)abs_err - ((er + 8) >> 4)) & 0xffffu) << 16);
}
__device__ __forceinline__ uint32_t lr_tot_update (uint32_t t, uint32_t abs_err) { return t + abs_err - ((t + 8u) >> 4); }
// avg : 5 | err_c : 2 of the channel (:128-135), from the table words as they are BEFORE this event's update
__device__ __forceinline__ uint32_t lr_chan_hi (uint32_t st, uint32_t t, const uint8_t *v2b)
{
    const uint32_t avg = v2b[(((st & 0xffffu) + 8u) >> 4) & 0xffu] & 0x1fu, er = st >> 16;
    const uint32_t ec = er < (t >> 1) ? 0u : er < t ? 1u : er < (t << 1) ? 2u : 3u;
    return avg | (ec << 5);
}
__device__ __forceinline__ uint32_t lr_difq (int32_t q1, int32_t q2)         // INTERLACE (context.h:100), capped (:122)
{
    const int32_t d = q1 - q2;
    const uint32_t il = d < 0 ? (((uint32_t)(-d)) << 1) - 1 : ((uint32_t)d) << 1;
    return il < 15 ? il : 15;
}

// state tables: chan_avgs_sums[n] = qbin(n) << AVG_SHIFT (:138-146); err sums 0; channel counters 0
__global__ void k_longr_init (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.y];
    uint4 *st4 = reinterpret_cast<uint4 *>(V.st);
    for (uint32_t n4 = blockIdx.x * blockDim.x + threadIdx.x; n4 < NCTX / 4; n4 += gridDim.x * blockDim.x) {
        const uint32_t v = (((n4 * 4) >> 16) & 0x1f) << 4;                  // (the bin is the same for four neighbours)
        st4[n4] = make_uint4 (v, v, v, v);
        if (n4 < NCHAN / 4) reinterpret_cast<uint4 *>(V.chan_num)[n4] = make_uint4 (0, 0, 0, 0);
    }
}

// value of lane (lane - d) of `cur`, continued into the previous tile's `prev`
__device__ __forceinline__ uint32_t lr_up (uint32_t cur, uint32_t prev, int d, int lane)
{
    const uint32_t a = __shfl_up_sync (FULL, cur, d), b = __shfl_sync (FULL, prev, (lane - d) & 31);
    return lane >= d ? a : b;
}

struct LrTileIn { uint32_t P, Qv; };     // per lane: base code of processing position e-1, quality of base e-3 (0 where there is none)

__device__ __forceinline__ LrTileIn lr_tile_load (const uint8_t *seq, const uint8_t *q, uint32_t L, uint32_t Lq, bool rev, uint32_t e, const uint8_t *lut)
{
    LrTileIn t; t.P = 0; t.Qv = 0;
    if (e >= 1 && e - 1 < L) t.P = lut[(rev ? 256 : 0) + seq[rev ? L - e : e - 1]];
    if (e >= 3 && e - 3 < Lq) t.Qv = (uint8_t)(q[rev ? Lq + 2 - e : e - 3] - '!');
    return t;
}

// encode, first pass: the channel of every base and the number of bases per channel (codec_longr_calc_channels :138-159 for
// every line, :185-203).  One warp per VBlock.
__global__ void __launch_bounds__(32) k_longr_channels (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.x];
    __shared__ uint32_t tot[NQ9];
    __shared__ uint8_t v2b[256], lut[512];
    const int lane = threadIdx.x;
    const uint32_t lt = (1u << lane) - 1;
    for (int i = lane; i < (int)NQ9; i += 32) tot[i] = 0x10101010u;          // memset (.., 1<<4, ..) on uint32 (:142)
    for (int i = lane; i < 256; i += 32) { v2b[i] = V.v2b[i]; lut[i] = (uint8_t)acgt_code (i); lut[256 + i] = (uint8_t)acgt_code_comp (i); }
    __syncwarp ();
    uint32_t *st = V.st;
    uint64_t nb = 0;
    for (uint32_t li = 0; li < V.n_lines; li++) {
        const uint32_t L = V.len[li], Lq = V.qlen ? V.qlen[li] : L;
        if (!Lq) continue;                                                  // (:188)
        const uint8_t *seq = V.txt + V.seq_off[li], *q = V.txt + V.qual_off[li];
        const bool rev = V.is_rev ? V.is_rev[li] : false;
        const uint32_t nE = Lq + 3;                                         // three events without a base, then one per quality
        uint32_t Pprev = 0, W1prev = 0, Qprev = 0;
        LrTileIn in = lr_tile_load (seq, q, L, Lq, rev, lane, lut);
        // context and table word of the first tile
        uint32_t u, val; int32_t q1;
        {
            const uint32_t W1 = in.P | (lr_up (in.P, Pprev, 1, lane) << 2), W2 = W1 | (lr_up (W1, W1prev, 2, lane) << 4), B = (W2 | (lr_up (W1, W1prev, 4, lane) << 8)) & 0xfffu;
            const uint32_t qm1 = lr_up (in.Qv, Qprev, 1, lane), qm2 = lr_up (in.Qv, Qprev, 2, lane);
            u = B | (lr_difq ((int32_t)qm1, (int32_t)qm2) << 12) | (lane == 0 ? 0u : (uint32_t)(v2b[qm1] & 0x1f) << 16);
            q1 = (int32_t)in.Qv; Pprev = in.P; W1prev = W1; Qprev = in.Qv;
            val = lane < nE ? st[u] : 0;
        }
        for (uint32_t e0 = 0; e0 < nE; e0 += 32) {
            const uint32_t e = e0 + lane;
            const bool active = e < nE;
            const bool more = e0 + 32 < nE;
            LrTileIn nx; nx.P = 0; nx.Qv = 0;
            if (more) nx = lr_tile_load (seq, q, L, Lq, rev, e + 32, lut);      // (in flight while this tile's entries arrive)

            // ---- st: events of the tile that share a context are chained in order
            const uint32_t g = __match_any_sync (FULL, active ? u : (0x80000000u | lane));
            const uint32_t rank = __popc (g & lt);
            const int prevlane = rank ? 31 - __clz ((int)(g & lt)) : lane;
            uint32_t ae, nv = lr_st_update (val, q1, ae);
            const uint32_t maxrank = __reduce_max_sync (FULL, rank);
            for (uint32_t r = 1; r <= maxrank; r++) {
                const uint32_t pv = __shfl_sync (FULL, nv, prevlane);
                if (rank == r) { val = pv; nv = lr_st_update (val, q1, ae); }
            }
            if (active && !(g >> lane >> 1)) st[u] = nv;                    // the last event of a context stores
            __syncwarp ();                                                  // (the next tile's loads, by other lanes, see these stores)

            // ---- the next tile's contexts and loads
            uint32_t u2 = 0, val2 = 0; int32_t q1n = 0;
            if (more) {
                const uint32_t W1 = nx.P | (lr_up (nx.P, Pprev, 1, lane) << 2), W2 = W1 | (lr_up (W1, W1prev, 2, lane) << 4), B = (W2 | (lr_up (W1, W1prev, 4, lane) << 8)) & 0xfffu;
                const uint32_t qm1 = lr_up (nx.Qv, Qprev, 1, lane), qm2 = lr_up (nx.Qv, Qprev, 2, lane);
                u2 = B | (lr_difq ((int32_t)qm1, (int32_t)qm2) << 12) | ((uint32_t)(v2b[qm1] & 0x1f) << 16);
                q1n = (int32_t)nx.Qv; Pprev = nx.P; W1prev = W1; Qprev = nx.Qv;
                if (e + 32 < nE) val2 = st[u2];
            }

            // ---- tot: events that share a Q are chained in order
            const uint32_t Qn = (u >> 12) & 0x1ffu;
            const uint32_t gq = __match_any_sync (FULL, active ? Qn : (0x8000u | lane));
            const uint32_t qrank = __popc (gq & lt), maxq = __reduce_max_sync (FULL, qrank);
            uint32_t t_pre = 0;
            for (uint32_t r = 0; r <= maxq; r++) {
                if (active && qrank == r) { t_pre = tot[Qn]; tot[Qn] = lr_tot_update (t_pre, ae); }
                __syncwarp ();
            }

            // ---- channels of the tile's bases, bases per channel
            const uint32_t ch = ((u >> 12) & 0x1ffu) | (lr_chan_hi (val, t_pre, v2b) << 9);
            const bool isbase = active && e >= 3;
            const uint32_t gc = __match_any_sync (FULL, isbase ? ch : (0x10000u | lane));
            if (isbase) {
                V.base_chan[nb + (e - 3)] = (uint16_t)ch;
                if (!(gc & lt)) atomicAdd (&V.chan_num[ch], (uint32_t)__popc (gc));
            }
            u = u2; val = val2; q1 = q1n;
        }
        nb += Lq;
    }
}

// lens (BGEN32, :237-240) and the exclusive prefix next_of_chan (:193-196); one CTA per VBlock.  Decode: the lengths are
// untrusted — their sum must be the number of bases.
__global__ void __launch_bounds__(1024) k_longr_prefix (const LrVb *vbs, int write_lens)
{
    const LrVb &V = vbs[blockIdx.x];
    __shared__ uint32_t sm[33];
    uint64_t acc = 0;
    for (uint32_t base = 0; base < NCHAN; base += 1024) {
        const uint32_t c = base + threadIdx.x;
        uint32_t v = write_lens ? V.chan_num[c] : __byte_perm (V.lens_be[c], 0, 0x0123);
        if (write_lens) V.lens_be[c] = __byte_perm (v, 0, 0x0123);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sm[warp] = inc;
        __syncthreads ();
        if (warp == 0) { uint32_t w = sm[lane], wi = w; for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (FULL, wi, o); if (lane >= o) wi += t; } sm[lane] = wi - w; if (lane == 31) sm[32] = wi; }
        __syncthreads ();
        const uint64_t start = acc + sm[warp] + inc - v;
        V.chan_num[c] = (uint32_t)(start < V.total ? start : V.total);
        if (!write_lens) {
            // cursor of the channel: next index, and the values up to the next 4-byte boundary
            const uint32_t idx = V.chan_num[c];
            uint32_t vals = 0;
            const uint32_t va = (uint32_t)((uintptr_t)V.values & 3);
            for (uint32_t k = idx; k < ((idx + va + 4) & ~3u) - va && k < V.total; k++) vals |= (uint32_t)V.values[k] << (8 * (k - idx));
            V.cur[c] = ((unsigned long long)idx << 32) | vals;
        }
        acc += sm[32];
        __syncthreads ();
    }
    if (!write_lens && threadIdx.x == 0 && acc != V.total) *V.err = 1;
}

// encode, second pass: stable scatter of the qualities into their channel segments (:205-230).  One warp per VBlock; the rank
// of a base among the tile's bases of its channel comes from __match_any_sync, the segment position from one atomic per channel
// and tile whose result is consumed one tile later.
__global__ void __launch_bounds__(32) k_longr_place (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.x];
    const int lane = threadIdx.x;
    const uint32_t lt = (1u << lane) - 1;
    uint64_t nb = 0;
    // pending tile
    bool p_act = false; uint32_t p_old = 0, p_rank = 0, p_q = 0; int p_leader = 0;
    for (uint32_t li = 0; li < V.n_lines; li++) {
        const uint32_t L = V.qlen ? V.qlen[li] : V.len[li];
        if (!L) continue;
        const uint8_t *q = V.txt + V.qual_off[li];
        const bool rev = V.is_rev ? V.is_rev[li] : false;
        for (uint32_t k0 = 0; k0 < L; k0 += 32) {
            const uint32_t k = k0 + lane;
            const bool act = k < L;
            const uint32_t ch = act ? V.base_chan[nb + k] : 0, qv = act ? (uint8_t)(q[rev ? L - 1 - k : k] - '!') : 0;
            const uint32_t g = __match_any_sync (FULL, act ? ch : (0x10000u | lane));
            const int leader = __ffs ((int)g) - 1;
            uint32_t old = 0;
            if (act && lane == leader) old = atomicAdd (&V.chan_num[ch], (uint32_t)__popc (g));
            // the tile before: its atomics have had a tile's time to come back
            const uint32_t po = __shfl_sync (FULL, p_old, p_leader);
            if (p_act) V.values[po + p_rank] = (uint8_t)p_q;
            __syncwarp ();                                                  // atomics of different lanes on one counter stay in tile order
            p_act = act; p_old = old; p_rank = __popc (g & lt); p_q = qv; p_leader = leader;
        }
        nb += L;
    }
    const uint32_t po = __shfl_sync (FULL, p_old, p_leader);
    if (p_act) V.values[po + p_rank] = (uint8_t)p_q;
}

// codec_longr_recon_one_read (:270-296) for every read of a VBlock; one warp per VBlock, lane 0 walks
__global__ void __launch_bounds__(32) k_longr_decode (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.x];
    __shared__ uint32_t tot[NQ9];
    __shared__ uint8_t v2b[256], lut[512];
    const int lane = threadIdx.x;
    for (int i = lane; i < (int)NQ9; i += 32) tot[i] = 0x10101010u;
    for (int i = lane; i < 256; i += 32) { v2b[i] = V.v2b[i]; lut[i] = (uint8_t)acgt_code (i); lut[256 + i] = (uint8_t)acgt_code_comp (i); }
    __syncwarp ();
    uint32_t *st = V.st; unsigned long long *cur = V.cur;
    const uint8_t *values = V.values; const uint64_t total = V.total;
    const uint32_t va = (uint32_t)((uintptr_t)values & 3);                  // refills are aligned 4-byte loads
    uint8_t *out = V.qual_out;
    uint32_t bad = 0;
    // a cursor whose refill is still in flight: written back when the next base has issued its own loads
    uint32_t p_ch = 0xffffffffu, p_idx = 0, p_vals = 0;
    for (uint32_t li = 0; li < V.n_lines; li++) {
        const uint32_t L = V.len[li];
        if (!L) continue;
        const uint8_t *seq = V.txt + V.seq_off[li];
        const bool rev = V.is_rev ? V.is_rev[li] : false;
        const uint8_t *lu = lut + (rev ? 256 : 0);
        uint32_t B = 0, u = 0, u_prev = 0xffffffffu, nv_prev = 0;
        int32_t prev = 0, qq = 0;
        bool missing = false;
        for (uint32_t e = 0; e < L + 3; e++) {
            if (lane == 0) {                                                // ---- the walk: lane 0
                uint32_t val = u == u_prev ? nv_prev : st[u];
                const uint32_t Qn = (u >> 12) & 0x1ffu, t_pre = tot[Qn];
                qq = 0;
                if (e >= 3) {
                    const uint32_t ch = Qn | (lr_chan_hi (val, t_pre, v2b) << 9);
                    uint32_t idx, vals;
                    if (ch == p_ch) { idx = p_idx; vals = p_vals; p_ch = 0xffffffffu; }
                    else {
                        const unsigned long long c = cur[ch];
                        if (p_ch != 0xffffffffu) { cur[p_ch] = ((unsigned long long)p_idx << 32) | p_vals; p_ch = 0xffffffffu; }
                        idx = (uint32_t)(c >> 32); vals = (uint32_t)c;
                    }
                    qq = (int32_t)(vals & 0xffu);
                    if (idx >= total) { bad = 1; qq = 0; }
                    idx++; vals >>= 8;
                    if (!((idx + va) & 3u)) { p_vals = idx < total ? *reinterpret_cast<const uint32_t *>(values + idx) : 0; p_ch = ch; p_idx = idx; }   // refill, off the chain
                    else cur[ch] = ((unsigned long long)idx << 32) | vals;
                    const uint32_t k = e - 3;
                    out[rev ? L - 1 - k : k] = (uint8_t)(qq + '!');
                }
                uint32_t ae;
                const uint32_t nv = lr_st_update (val, qq, ae);
                st[u] = nv; tot[Qn] = lr_tot_update (t_pre, ae);
                u_prev = u; nv_prev = nv;
                // the context of the next event: base code of processing position e
                const uint32_t b = e < L ? lu[seq[rev ? L - 1 - e : e]] : 0;
                B = ((B << 2) | b) & 0xfffu;
                u = B | (lr_difq (qq, prev) << 12) | ((uint32_t)(v2b[qq & 0xff] & 0x1f) << 16);
                prev = qq;
            }
            // ---- the other lanes: the context AFTER the next one depends on the next quality, which is not known yet — but it is
            //      most likely within 16 of this one.  Each lane asks L2 for the table word of one candidate, so that the walk finds
            //      it there (an L2 hit instead of a DRAM round trip on the chain).  A wrong guess costs nothing but the traffic.
            const int32_t qb = __shfl_sync (FULL, qq, 0);
            const uint32_t Bb = __shfl_sync (FULL, B, 0);
            missing = e >= 3 && qb == 255;                                  // 255 + '!' == ' ': the line has no quality (:278)
            if (missing) break;                                             // after the state update, like RECON_ONE_QUAL
            if (LR_L2_HINTS && e + 1 < L + 3) {
                const uint32_t b1 = e + 1 < L ? lu[seq[rev ? L - 2 - e : e + 1]] : 0;
                const int32_t cq = min (93, max (0, qb - 15 + lane));
                prefetch_l2 (st + ((((Bb << 2) | b1) & 0xfffu) | (lr_difq (cq, qb) << 12) | ((uint32_t)(v2b[cq] & 0x1f) << 16)));
            }
        }
        if (lane == 0) {
            if (missing) out[0] = '*';                                      // sam_reconstruct_missing_quality (sam_qual.c:532-541); the rest of the line is undefined
            if (V.missing) V.missing[li] = missing;
        }
        out += L;
    }
    if (lane == 0) {
        if (p_ch != 0xffffffffu) cur[p_ch] = ((unsigned long long)p_idx << 32) | p_vals;
        if (bad) *V.err = 2;
    }
}

// histogram of the quality values of a VBlock's lines (add_to_histogram, codec_longr.c:60-64)
__global__ void k_longr_hist (const LrVb *vbs, uint32_t *hist)
{
    const LrVb &V = vbs[0];
    __shared__ uint32_t h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads ();
    for (uint32_t li = blockIdx.x; li < V.n_lines; li += gridDim.x) {
        const uint32_t L = V.qlen ? V.qlen[li] : V.len[li];
        const uint8_t *q = V.txt + V.qual_off[li];
        if (L == 1 && q[0] == ' ') continue;                                // IS_SPACE: a missing quality (:88)
        for (uint32_t k = threadIdx.x; k < L; k += blockDim.x) atomicAdd (&h[(uint8_t)(q[k] - '!')], 1u);
    }
    __syncthreads ();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) if (h[i]) atomicAdd (&hist[i], h[i]);
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

int longr_run (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags, int mode /* 0 encode, 1 decode, 2 histogram */, uint32_t *hist_out)
{
    if (!e || !vbs) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    const bool encode = mode == 0, decode = mode == 1;
    cudaStream_t st = e->stream;
    std::vector<LrVb> h (n_vbs);
    std::vector<uint64_t> total (n_vbs, 0), out_len (n_vbs, 0);     // qualities in the VBlock (= bytes of values); decode: bytes of qual_out
    for (uint32_t v = 0; v < n_vbs; v++) {
        if (devptr) total[v] = out_len[v] = vbs[v].n_bases ? vbs[v].n_bases : vbs[v].txt_len;   // lengths live on the device: the caller's count, else bound by the text size
        else {
            for (uint32_t i = 0; i < vbs[v].n_lines; i++) { total[v] += (mode != 1 && vbs[v].qual_len) ? vbs[v].qual_len[i] : vbs[v].len[i]; out_len[v] += vbs[v].len[i]; }
            if (mode == 1 && vbs[v].n_bases) { if (vbs[v].n_bases > total[v]) return GZB_E_BADARG; total[v] = vbs[v].n_bases; }   // fewer values than bases: lines without quality
        }
        if (total[v] >= (1ull << 32)) { e->err = "LONGR: more than 4 G qualities in a VBlock"; return GZB_E_BADARG; }
    }
    Carver c { nullptr, 0 };
    LrVb *d_vbs = nullptr; uint32_t *d_err = nullptr, *d_hist = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<LrVb> (n_vbs); d_err = c.take<uint32_t> (n_vbs); d_hist = c.take<uint32_t> (256);
        for (uint32_t v = 0; v < n_vbs; v++) {
            LrVb &D = h[v]; const gzb_longr_vb &S = vbs[v];
            memset (&D, 0, sizeof D);
            D.n_lines = S.n_lines; D.total = total[v];
            D.err = d_err ? d_err + v : nullptr;
            memcpy (D.v2b, S.value_to_bin, 256);
            D.txt      = devptr ? (const uint8_t *)S.txt : c.take<uint8_t> (S.txt_len + 16);
            D.seq_off  = devptr ? S.seq_off : c.take<uint64_t> (S.n_lines + 1);
            D.qual_off = devptr ? S.qual_off : (!decode ? c.take<uint64_t> (S.n_lines + 1) : nullptr);
            D.len      = devptr ? S.len : c.take<uint32_t> (S.n_lines + 1);
            D.is_rev   = S.is_rev ? (devptr ? S.is_rev : c.take<uint8_t> (S.n_lines + 1)) : nullptr;
            D.qlen     = (!decode && S.qual_len) ? (devptr ? S.qual_len : c.take<uint32_t> (S.n_lines + 1)) : nullptr;
            if (mode == 2) continue;
            D.st = c.take<uint32_t> (NCTX); D.chan_num = c.take<uint32_t> (NCHAN);
            D.values   = devptr ? (uint8_t *)S.values : c.take<uint8_t> (total[v] + 16);
            D.lens_be  = devptr ? S.lens_be : c.take<uint32_t> (NCHAN);
            if (decode) {
                D.cur = c.take<unsigned long long> (NCHAN);
                D.qual_out = devptr ? (uint8_t *)S.qual_out : c.take<uint8_t> (out_len[v] + 16);
                D.missing  = S.missing ? (devptr ? S.missing : c.take<uint8_t> (S.n_lines + 1)) : nullptr;
            }
            else D.base_chan = c.take<uint16_t> (total[v] + 1);
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
    }
    for (uint32_t v = 0; v < n_vbs; v++) {
        LrVb &D = h[v]; const gzb_longr_vb &S = vbs[v];
        if (!devptr) {
            if (S.txt_len) CK (cudaMemcpyAsync ((void *)D.txt, S.txt, S.txt_len, cudaMemcpyHostToDevice, st));
            if (S.n_lines) {
                CK (cudaMemcpyAsync ((void *)D.seq_off, S.seq_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
                if (!decode) CK (cudaMemcpyAsync ((void *)D.qual_off, S.qual_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
                CK (cudaMemcpyAsync ((void *)D.len, S.len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
                if (S.is_rev) CK (cudaMemcpyAsync ((void *)D.is_rev, S.is_rev, S.n_lines, cudaMemcpyHostToDevice, st));
                if (D.qlen) CK (cudaMemcpyAsync ((void *)D.qlen, S.qual_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
            }
            if (decode) {
                if (total[v]) CK (cudaMemcpyAsync (D.values, S.values, total[v], cudaMemcpyHostToDevice, st));
                CK (cudaMemcpyAsync (D.lens_be, S.lens_be, NCHAN * 4, cudaMemcpyHostToDevice, st));
            }
        }
    }
    CK (cudaMemcpyAsync (d_vbs, h.data (), n_vbs * sizeof (LrVb), cudaMemcpyHostToDevice, st));
    CK (cudaMemsetAsync (d_err, 0, 4 * (size_t)n_vbs, st));
    if (mode == 2) {
        CK (cudaMemsetAsync (d_hist, 0, 1024, st));
        k_longr_hist<<<592, 256, 0, st>>>(d_vbs, d_hist); e->launches++;
        CK (cudaMemcpyAsync (hist_out, d_hist, 1024, cudaMemcpyDeviceToHost, st));
        CK (cudaStreamSynchronize (st));
        CK (cudaGetLastError ());
        return GZB_OK;
    }
    k_longr_init<<<dim3 (64, n_vbs), 256, 0, st>>>(d_vbs);
    if (encode) {
        cudaEventRecord (e->ev0, st);
        k_longr_channels<<<n_vbs, 32, 0, st>>>(d_vbs);
        cudaEventRecord (e->ev1, st);
        k_longr_prefix<<<n_vbs, 1024, 0, st>>>(d_vbs, 1);
        k_longr_place<<<n_vbs, 32, 0, st>>>(d_vbs);
        e->launches += 4;
    }
    else {
        k_longr_prefix<<<n_vbs, 1024, 0, st>>>(d_vbs, 0);
        cudaEventRecord (e->ev0, st);
        k_longr_decode<<<n_vbs, 32, 0, st>>>(d_vbs);
        cudaEventRecord (e->ev1, st);
        e->launches += 3;
    }
    std::vector<uint32_t> errs (n_vbs, 0);
    CK (cudaMemcpyAsync (errs.data (), d_err, 4 * (size_t)n_vbs, cudaMemcpyDeviceToHost, st));
    if (!devptr)
        for (uint32_t v = 0; v < n_vbs; v++) {
            if (encode) {
                if (total[v]) CK (cudaMemcpyAsync (vbs[v].values, h[v].values, total[v], cudaMemcpyDeviceToHost, st));
                CK (cudaMemcpyAsync (vbs[v].lens_be, h[v].lens_be, NCHAN * 4, cudaMemcpyDeviceToHost, st));
            }
            else {
                if (out_len[v]) CK (cudaMemcpyAsync (vbs[v].qual_out, h[v].qual_out, out_len[v], cudaMemcpyDeviceToHost, st));
                if (vbs[v].missing && vbs[v].n_lines) CK (cudaMemcpyAsync (vbs[v].missing, h[v].missing, vbs[v].n_lines, cudaMemcpyDeviceToHost, st));
            }
        }
    CK (cudaStreamSynchronize (st));
    CK (cudaGetLastError ());
    cudaEventElapsedTime (&e->last_domain_ms, e->ev0, e->ev1);
    for (uint32_t v = 0; v < n_vbs; v++)
        if (errs[v]) { e->err = errs[v] == 1 ? "LONGR: channel lengths do not add up to the number of qualities" : "LONGR: a channel runs past the end of the values"; return GZB_E_CORRUPT; }
    return GZB_OK;
}

} // namespace

extern "C" int gzb_longr_encode (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags) { return longr_run (e, vbs, n_vbs, flags, 0, nullptr); }
extern "C" int gzb_longr_decode (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags) { return longr_run (e, vbs, n_vbs, flags, 1, nullptr); }

// codec_longr_segconf_calculate_bins (codec_longr.c:66-136): histogram of the VBlock's qualities on the GPU, the 32 bins on the host
extern "C" int gzb_longr_calculate_bins (gzb_engine *e, gzb_longr_vb *vb, uint32_t flags, uint8_t value_to_bin[256])
{
    if (!e || !vb || !value_to_bin) return GZB_E_BADARG;
    uint32_t histogram[256];
    int rc = longr_run (e, vb, 1, flags, 2, histogram);
    if (rc) return rc;
    uint64_t num_values = 0;
    for (int i = 0; i < 256; i++) num_values += histogram[i];
    if (!num_values) return GZB_SOFT_FAIL;                                  // flag.no_longr (:100-103)
    const unsigned NUM_BINS = 32, fixed = 11;
    uint32_t next_val = 0;
    for (unsigned bin_i = 0; bin_i < NUM_BINS; bin_i++) {
        const uint64_t at_least = num_values / (NUM_BINS - bin_i);
        if (bin_i < fixed) { num_values -= histogram[next_val]; value_to_bin[next_val] = (uint8_t)bin_i; next_val++; }
        else {
            uint64_t bin_content = 0;
            while (bin_content < at_least && next_val < 256) {
                bin_content += histogram[next_val]; num_values -= histogram[next_val];
                value_to_bin[next_val] = (uint8_t)bin_i; next_val++;
            }
        }
    }
    memset (value_to_bin + next_val, NUM_BINS - 1, 256 - next_val);
    return GZB_OK;
}
