// domain_fastq.cu — ACGT/XCGT 2-bit sequence packing and DOMQ dominant-quality modelling on sm_100a.
//
// Reference functions replaced (all relative to /root/reference/src):
//   ACGT: codec_acgt_compress up to its sub-codec call (codec_acgt.c:45-55, 64-163), codec_acgt_uncompress /
//         codec_xcgt_uncompress after theirs (:185-248).
//   DOMQ: codec_domq_prepare_normalize (codec_domq.c:252-293: per-line histogram + dom :139-178, compaction
//         :180-197, per-dom rank tables :199-247 — the qsort stays on the host),
//         codec_domq_compress up to its sub-codec call (:379-500), codec_domq_reconstruct (:774-809) for
//         all lines of a VBlock at once.
//
// These are the bandwidth-shaped kernels of the FASTQ path: bytes in, bytes out, one or two passes.
#include <cstring>
#include <cstdlib>
#include <vector>
#include <string>
#include <algorithm>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "engine.h"

using namespace gzb;

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

struct Carver {
    uint8_t *base; size_t off;
    template <typename T> T *take (size_t count) {
        size_t bytes = (count * sizeof (T) + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

// ================================================================================================ ACGT
// _acgt_encode (reference.c:45-58): ACGT either case -> 0..3, IUPAC -> lowest participating base, everything else 0
__device__ __forceinline__ uint32_t acgt_code (uint32_t c)
{
    switch (c) {
        case 'C': case 'c': case 'Y': case 'y': case 'S': case 's': case 'B': case 'b': return 1;
        case 'G': case 'g': case 'K': case 'k': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 0;
    }
}

// one thread = 32 bases = one little-endian 64-bit word (codec_acgt.c:45-55) + 32 exception bytes (:67-70,104-107)
__device__ __forceinline__ void acgt_pack_body (const uint8_t * __restrict__ seq, uint64_t n, uint64_t * __restrict__ packed, uint8_t * __restrict__ x, int *x_nonzero)
{
    __shared__ uint8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (uint8_t)acgt_code (i);
    __syncthreads ();
    const uint64_t nwords = (2 * n + 63) / 64;
    uint32_t any = 0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t base = w * 32;
        uint64_t word = 0;
        const bool full = base + 32 <= n;
        if (full && ((reinterpret_cast<uintptr_t>(seq) + base) & 15) == 0) {
            const uint4 *p = reinterpret_cast<const uint4 *>(seq + base);
            uint4 v[2] = { p[0], p[1] };
            const uint32_t *u = reinterpret_cast<const uint32_t *>(v);
            uint32_t xe[8];
            #pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t cc = u[k], xo = 0;
                // four bases at once when they are all upper-case A C G T (the rule): code = ((c >> 1) ^ (c >> 2)) & 3
                // ('A' 0x41 -> 0, 'C' 0x43 -> 1, 'G' 0x47 -> 2, 'T' 0x54 -> 3), checked by mapping the codes back to characters
                const uint32_t code4 = ((cc >> 1) ^ (cc >> 2)) & 0x03030303u;               // one 2-bit code per byte
                const uint32_t sel = (code4 & 0x3u) | ((code4 >> 4) & 0x30u) | ((code4 >> 8) & 0x300u) | ((code4 >> 12) & 0x3000u);   // the four codes as byte selectors
                if (__byte_perm (0x54474341u, 0, sel) == cc) {
                    const uint32_t two = code4 | (code4 >> 6);                               // bytes 0,1 -> bits 0..3 ; bytes 2,3 -> bits 16..19
                    word |= (uint64_t)((two | (two >> 12)) & 0xffu) << (8 * k);
                    xe[k] = 0;
                    continue;
                }
                #pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t c = (cc >> (8 * b)) & 0xff;
                    word |= (uint64_t)lut[c] << (2 * (4 * k + b));
                    uint32_t up = c & 0xdf;
                    uint32_t e = (up == 'A' || up == 'C' || up == 'G' || up == 'T') ? (c >> 5) & 1 : c;   // upper -> 0, lower -> 1, else verbatim
                    xo |= e << (8 * b);
                }
                xe[k] = xo; any |= xo;
            }
            if (x) {
                if (((reinterpret_cast<uintptr_t>(x) + base) & 15) == 0) {
                    uint4 *q = reinterpret_cast<uint4 *>(x + base);
                    q[0] = make_uint4 (xe[0], xe[1], xe[2], xe[3]); q[1] = make_uint4 (xe[4], xe[5], xe[6], xe[7]);
                }
                else for (int k = 0; k < 32; k++) x[base + k] = (uint8_t)(xe[k >> 2] >> (8 * (k & 3)));
            }
        }
        else {
            for (int k = 0; k < 32 && base + k < n; k++) {
                uint32_t c = seq[base + k];
                word |= (uint64_t)lut[c] << (2 * k);
                uint32_t up = c & 0xdf;
                uint32_t e = (up == 'A' || up == 'C' || up == 'G' || up == 'T') ? (c >> 5) & 1 : c;
                if (x) x[base + k] = (uint8_t)e;
                any |= e;
            }
        }
        packed[w] = word;
    }
    any = __reduce_or_sync (0xffffffffu, any);
    if (any && (threadIdx.x & 31) == 0) atomicOr (x_nonzero, 1);
}

__global__ void k_acgt_pack (const uint8_t *seq, uint64_t n, uint64_t *packed, uint8_t *x, int *x_nonzero) { acgt_pack_body (seq, n, packed, x, x_nonzero); }

// a batch of VBlocks in one launch: blockIdx.y = VBlock
struct AcgtD { const uint8_t *seq; uint64_t n; uint64_t *packed; uint8_t *x; int *flag; };
__global__ void k_acgt_pack_batch (const AcgtD *d) { const AcgtD a = d[blockIdx.y]; if (a.n) acgt_pack_body (a.seq, a.n, a.packed, a.x, a.flag); }

// codec_acgt_uncompress / codec_xcgt_uncompress (:185-248): one thread = 32 bases
__device__ __forceinline__ void acgt_unpack_body (const uint64_t * __restrict__ packed, const uint8_t * __restrict__ x, uint64_t n, uint8_t * __restrict__ seq)
{
    const uint64_t nwords = (2 * n + 63) / 64;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t base = w * 32, word = packed[w];
        const bool fast = base + 32 <= n && ((reinterpret_cast<uintptr_t>(seq) + base) & 15) == 0 &&
                          (!x || ((reinterpret_cast<uintptr_t>(x) + base) & 15) == 0);
        if (fast) {
            uint32_t xe[8] = {0,0,0,0,0,0,0,0};
            if (x) { const uint4 *p = reinterpret_cast<const uint4 *>(x + base); uint4 a = p[0], b = p[1];
                     xe[0]=a.x; xe[1]=a.y; xe[2]=a.z; xe[3]=a.w; xe[4]=b.x; xe[5]=b.y; xe[6]=b.z; xe[7]=b.w; }
            uint32_t o[8];
            #pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t ow = 0;
                #pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t code = (uint32_t)(word >> (2 * (4 * k + b))) & 3;
                    uint32_t base_c = (0x54474341u >> (8 * code)) & 0xff;         // "ACGT"
                    uint32_t e = (xe[k] >> (8 * b)) & 0xff;
                    uint32_t c = e == 0 ? base_c : e == 1 ? base_c + 32 : e;
                    ow |= c << (8 * b);
                }
                o[k] = ow;
            }
            uint4 *q = reinterpret_cast<uint4 *>(seq + base);
            q[0] = make_uint4 (o[0], o[1], o[2], o[3]); q[1] = make_uint4 (o[4], o[5], o[6], o[7]);
        }
        else for (int k = 0; k < 32 && base + k < n; k++) {
            uint32_t code = (uint32_t)(word >> (2 * k)) & 3;
            uint32_t base_c = (0x54474341u >> (8 * code)) & 0xff;
            uint32_t e = x ? x[base + k] : 0;
            seq[base + k] = (uint8_t)(e == 0 ? base_c : e == 1 ? base_c + 32 : e);
        }
    }
}

__global__ void k_acgt_unpack (const uint64_t *packed, const uint8_t *x, uint64_t n, uint8_t *seq) { acgt_unpack_body (packed, x, n, seq); }
__global__ void k_acgt_unpack_batch (const AcgtD *d) { const AcgtD a = d[blockIdx.y]; if (a.n) acgt_unpack_body (a.packed, a.x, a.n, const_cast<uint8_t *>(a.seq)); }

// ================================================================================================ DOMQ
constexpr int NQ = 95, FIRST_Q = 32;              // printable qualities ' '..'~' (codec_domq.c:31-33)

struct DqVb {                                      // device view of one VBlock's quality lines
    const uint8_t  *txt;
    const uint64_t *line_off;
    const uint32_t *line_len;
    uint8_t  *line_dom;        // raw dom (q-32) after the histogram pass, compacted dom after prepare
    uint8_t  *line_diverse;
    uint32_t *hist;            // [95][95] per-dom histograms, then [95] lines_with_dom
    uint32_t *nd_off;          // per line: offset in the concatenation of non-diverse lines
    uint32_t *dv_off;          // per line: offset in DIVRQUAL
    uint32_t *mx_idx;          // per line: index in QUALMPLX
    uint8_t  *E;               // normalised non-diverse concatenation
    uint8_t  *qual, *runs, *mplx, *divr;
    uint32_t *lens;            // [8]: qual_len, runs_len, mplx_len, divr_len, M (non-diverse total), last_line_len
    void     *tiles, *toff;    // per tile of E: DqTile (k_dqs_count), DqTileOff (k_dqs_scan)
    uint32_t  n_lines;
    uint8_t   no_doms;
    uint8_t   pad[3];
    uint8_t   normalize[NQ * NQ];   // [cdom*95 + q-32] -> rank
    uint8_t   dom_to_cdom[NQ];
};

// ---- pass 1: per-line histogram, dom, diversity; per-dom histograms (codec_domq_calc_histogram :139-178)
// One THREAD per line (a line is ~150 qualities, far too short for a warp: the earlier warp-per-line pass spent 2.2 warp
// instructions per byte on match/ballot work).  The thread reads the aligned 16-byte blocks of its line, the next block in flight while
// it counts the current one, and counts without a branch (32 lines in lock step diverge on anything that depends on the data — a
// version that kept the current run in registers executed every lane's run ends — and sm_100 has no byte-wise SIMD min/max): every byte
// is one shared-memory atomic add whose result nobody waits for.  A thread owns 64 words: character c adds 1 << 16 * (c >> 6 & 1) to
// word c & 63, i.e. two 16-bit counters per word, the 32 lanes always in 32 different banks; bytes of the first and last block that are
// not the line's are turned into 0x7f first, a counter nobody reads (like those below ' ').  One pass over the 95 counters then finds
// every line's arg-max (ties -> the higher quality, :153-158) and, by ballot, the qualities any of the warp's 32 lines uses; only
// those are added into the doms' histograms — one REDUX per quality when the 32 lines share their dom, which is the common case,
// accumulated per warp in shared memory and flushed to the VBlock's histogram when the dom changes.  Lines above DQ_LONG_LINE
// qualities are counted by the warp as a whole.
// Measured (profiles/r02_domq_kernels.md, r02_domq_full.md; 128 VBlocks = 1.77 GB per launch): 2.65 ms with half the instructions of the
// warp-per-line pass (5.3 ms in round 1's terms per 128 VBlocks: 2.64 ms per 64).  8-bit load / add / store counters (lines <= 255 only)
// took 2.45 ms at twice the resident warps: the pass is bound by L1/shared-memory wavefronts (a 16-byte load per lane touches 32 lines,
// plus one or two shared-memory accesses per byte), not by occupancy or issue slots — ncu: wavefront pipe 46 %, short scoreboard 14 and
// mio throttle 8 of 34 stall cycles per instruction.
constexpr int LINES_PER_BLOCK = 512;
constexpr uint32_t DQ_LONG_LINE = 4096;
__device__ __forceinline__ void dq_flush_acc (const DqVb &V, uint32_t *acc, uint32_t dom, int lane)
{
    for (int k = lane; k < 96; k += 32) {
        const uint32_t c = acc[k];
        if (c) { atomicAdd (&V.hist[k < NQ ? dom * NQ + k : NQ * NQ + dom], c); acc[k] = 0; }
    }
    __syncwarp ();
}
__global__ void __launch_bounds__(128) k_domq_linehist (const DqVb *vbs, const uint32_t *blk_vb, const uint32_t *blk_first)
{
    const DqVb &V = vbs[blk_vb[blockIdx.x]];
    const uint32_t first = blk_first[blockIdx.x], last = min (first + LINES_PER_BLOCK, V.n_lines);
    __shared__ uint32_t cnt[64 * 128];                 // [character & 63][thread]: low half = that character, high half = character + 64
    __shared__ uint32_t wacc[4][96];                   // per warp: [0..94] histogram of the lines of dom `cur`, [95] their number
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 64 * 128; i += 128) cnt[i] = 0;
    for (int i = lane; i < 96; i += 32) wacc[warp][i] = 0;
    __syncthreads ();
    uint32_t *const mine = cnt + tid;                  // mine[(c & 63) * 128]
    uint32_t *const acc = wacc[warp];
    uint32_t cur = 0;
    #define DQ_COUNT(c_) { const uint32_t b_ = (c_); atomicAdd (&mine[(b_ & 63) * 128], 1u << ((b_ & 64) >> 2)); }
    #define DQ_COUNTER(k_) ((mine[(((k_) + FIRST_Q) & 63) * 128] >> ((((k_) + FIRST_Q) & 64) >> 2)) & 0xffff)
    for (uint32_t g0 = first + warp * 32; g0 < last; g0 += 128) {
        const uint32_t gl = g0 + lane;
        const uint32_t len = gl < last ? V.line_len[gl] : 0;
        const uint64_t off = gl < last ? V.line_off[gl] : 0;
        const bool own = len && len <= DQ_LONG_LINE;
        if (own) {
            const uint32_t skip = (uint32_t)((uintptr_t)(V.txt + off) & 15), end = skip + len, nblk = (end + 15) >> 4;
            const uint4 *blk = reinterpret_cast<const uint4 *>(V.txt + off - skip);   // (an aligned block that holds a byte of the line lies inside the line's allocation)
            uint4 nxt = blk[0];
            for (uint32_t b = 0; b < nblk; b++) {
                const uint4 q = nxt;
                if (b + 1 < nblk) nxt = blk[b + 1];
                uint32_t w[4] = { q.x, q.y, q.z, q.w };
                if (b == 0 || b + 1 == nblk) {
                    #pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int p0 = (int)(b * 16) + 4 * j;
                        const int lead = max (0, min (4, (int)skip - p0)), upto = max (0, min (4, (int)end - p0));      // bytes [lead, upto) of the word are the line's
                        const uint32_t keep = (uint32_t)((0xffffffffull << (8 * lead)) & ~(0xffffffffull << (8 * upto)));
                        w[j] = (w[j] & keep) | (0x7f7f7f7fu & ~keep);
                    }
                }
                #pragma unroll
                for (int j = 0; j < 4; j++) { DQ_COUNT (w[j] & 0x7f); DQ_COUNT ((w[j] >> 8) & 0x7f); DQ_COUNT ((w[j] >> 16) & 0x7f); DQ_COUNT ((w[j] >> 24) & 0x7f); }
            }
        }
        __syncwarp ();
        // arg-max of every line, and the qualities in use
        uint32_t bc = 0, bq = 0, used[3] = { 0, 0, 0 };
        #pragma unroll
        for (int k = 0; k < NQ; k++) {
            const uint32_t c = own ? DQ_COUNTER (k) : 0;
            if (c >= bc) { bc = c; bq = k; }
            if (__any_sync (0xffffffffu, c)) used[k >> 5] |= 1u << (k & 31);
        }
        uint32_t my_dom = 0, my_div = 0;
        if (own) { my_dom = bq; my_div = (100u * bc / len < 85u) ? 1 : 0; }                       // DOMQ_THRESHOLD (:141,160)
        const uint32_t act = __ballot_sync (0xffffffffu, own);
        if (act) {
            const uint32_t d0 = __shfl_sync (0xffffffffu, my_dom, __ffs (act) - 1);
            const bool uniform = __ballot_sync (0xffffffffu, own && my_dom == d0) == act;
            if (uniform && d0 != cur) { dq_flush_acc (V, acc, cur, lane); cur = d0; }
            #pragma unroll
            for (int wd = 0; wd < 3; wd++)
                for (uint32_t m = used[wd]; m; m &= m - 1) {
                    const uint32_t k = wd * 32 + __ffs (m) - 1;
                    const uint32_t c = own ? DQ_COUNTER (k) : 0;
                    if (uniform) { const uint32_t sum = __reduce_add_sync (0xffffffffu, c); if (lane == 0) acc[k] += sum; }
                    else if (c) atomicAdd (&V.hist[my_dom * NQ + k], c);
                }
            if (uniform) { if (lane == 0) acc[95] += __popc (act); __syncwarp (); }
            else if (own) atomicAdd (&V.hist[NQ * NQ + my_dom], 1u);
            // clear the words of the qualities in use (a thread counts 4 lines of <= 4096 characters in its CTA's life: what characters
            // outside ' '..'~' leave behind in the other counters cannot carry from a low half into a high one)
            if (own) {
                #pragma unroll
                for (int wd = 0; wd < 3; wd++)
                    for (uint32_t m = used[wd]; m; m &= m - 1) mine[((wd * 32 + __ffs (m) - 1 + FIRST_Q) & 63) * 128] = 0;
            }
        }
        // the long lines of these 32, one at a time, by the whole warp: equal qualities of a 32-byte chunk are counted once, by their lowest lane
        uint32_t longs = __ballot_sync (0xffffffffu, len > DQ_LONG_LINE);
        if (longs) { dq_flush_acc (V, acc, cur, lane); }
        while (longs) {
            const int t = __ffs (longs) - 1; longs &= longs - 1;
            const uint32_t ll = __shfl_sync (0xffffffffu, len, t);
            const uint8_t *q = V.txt + __shfl_sync (0xffffffffu, off, t);
            for (uint32_t i0 = 0; i0 < ll; i0 += 32) {
                const uint32_t i = i0 + lane;
                const uint32_t v = i < ll ? (uint32_t)q[i] - FIRST_Q : 0x100u + lane;
                const uint32_t peers = __match_any_sync (0xffffffffu, v);
                if (v < (uint32_t)NQ && (peers & ((1u << lane) - 1)) == 0) acc[v] += __popc (peers);
                __syncwarp ();
            }
            uint32_t lc = 0, lq = 0;
            for (int k = lane; k < NQ; k += 32) { uint32_t c = acc[k]; if (c >= lc) { lc = c; lq = k; } }
            for (int o = 16; o; o >>= 1) {
                uint32_t oc = __shfl_xor_sync (0xffffffffu, lc, o), oq = __shfl_xor_sync (0xffffffffu, lq, o);
                if (oc > lc || (oc == lc && oq > lq)) { lc = oc; lq = oq; }
            }
            if (lane == t) { my_dom = lq; my_div = (100u * lc / ll < 85u) ? 1 : 0; }
            if (lane == 0) acc[95] = 1;
            __syncwarp ();
            dq_flush_acc (V, acc, lq, lane);
        }
        if (gl < last) { V.line_dom[gl] = (uint8_t)my_dom; V.line_diverse[gl] = (uint8_t)my_div; }
    }
    #undef DQ_COUNT
    #undef DQ_COUNTER
    dq_flush_acc (V, acc, cur, lane);
}

// ---- block scan helpers (512 threads)
__device__ __forceinline__ uint32_t warp_incl_sum (uint32_t v, int lane)
{
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}
__device__ __forceinline__ uint32_t warp_incl_max (uint32_t v, int lane)
{
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, v, o); if (lane >= o) v = max (v, t); }
    return v;
}
// exclusive block sum; *total = block total.  sm needs 17 words.  All threads call.
__device__ uint32_t block_excl_sum (uint32_t v, uint32_t *sm, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = warp_incl_sum (v, lane);
    if (lane == 31) sm[warp] = inc;
    __syncthreads ();
    if (warp == 0) { uint32_t w = lane < nw ? sm[lane] : 0; uint32_t wi = warp_incl_sum (w, lane); if (lane < nw) sm[lane] = wi - w; if (lane == nw - 1) sm[16] = wi; }
    __syncthreads ();
    uint32_t r = sm[warp] + inc - v;
    *total = sm[16];
    __syncthreads ();
    return r;
}
__device__ uint32_t block_excl_max (uint32_t v, uint32_t *sm, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = warp_incl_max (v, lane);
    uint32_t excl_in_warp = __shfl_up_sync (0xffffffffu, inc, 1); if (lane == 0) excl_in_warp = 0;
    if (lane == 31) sm[warp] = inc;
    __syncthreads ();
    if (warp == 0) { uint32_t w = lane < nw ? sm[lane] : 0; uint32_t wi = warp_incl_max (w, lane); uint32_t we = __shfl_up_sync (0xffffffffu, wi, 1); if (lane == 0) we = 0;
                     if (lane < nw) sm[lane] = we; if (lane == nw - 1) sm[16] = wi; }
    __syncthreads ();
    uint32_t r = max (sm[warp], excl_in_warp);
    *total = sm[16];
    __syncthreads ();
    return r;
}

// ---- pass 2a: per-line output offsets (one CTA per VB)
__global__ void __launch_bounds__(512) k_domq_lineoffsets (const DqVb *vbs)
{
    const DqVb &V = vbs[blockIdx.x];
    __shared__ uint32_t sm[17];
    uint32_t nd = 0, dv = 0, mx = 0, t;
    for (uint32_t base = 0; base < V.n_lines; base += 512) {
        uint32_t li = base + threadIdx.x;
        uint32_t len = li < V.n_lines ? V.line_len[li] : 0;
        bool div = len && V.line_diverse[li];
        uint32_t a = block_excl_sum (div ? 0 : len, sm, &t); uint32_t ta = t;
        uint32_t b = block_excl_sum (div ? len : 0, sm, &t); uint32_t tb = t;
        uint32_t c = block_excl_sum (len ? 1 : 0, sm, &t);   uint32_t tc = t;
        if (li < V.n_lines) { V.nd_off[li] = nd + a; V.dv_off[li] = dv + b; V.mx_idx[li] = mx + c; }
        nd += ta; dv += tb; mx += tc;
    }
    if (threadIdx.x == 0) { V.lens[4] = nd; V.lens[3] = dv; V.lens[2] = mx; V.lens[5] = V.n_lines ? V.line_len[V.n_lines - 1] : 0; }
}

// ---- pass 2b: normalise every line with its dom's rank table (:347-366) into E / DIVRQUAL, write QUALMPLX
__global__ void __launch_bounds__(256) k_domq_normalize (const DqVb *vbs, const uint32_t *blk_vb, const uint32_t *blk_first)
{
    const DqVb &V = vbs[blk_vb[blockIdx.x]];
    const uint32_t first = blk_first[blockIdx.x], last = min (first + LINES_PER_BLOCK, V.n_lines);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 64 consecutive lines per warp, 32 at a time: the lanes fetch the 32 lines' descriptors together, then the warp copies line by line
    for (uint32_t g0 = first + warp * (LINES_PER_BLOCK / 8); g0 < min (last, first + (warp + 1) * (LINES_PER_BLOCK / 8)); g0 += 32) {
        const uint32_t gl = g0 + lane, gend = min (last, first + (warp + 1) * (LINES_PER_BLOCK / 8));
        uint32_t my_len = 0, my_cdom = 0, my_div = 0, my_dst = 0; uint64_t my_off = 0;
        if (gl < gend) {
            my_len = V.line_len[gl]; my_off = V.line_off[gl];
            if (my_len) { my_cdom = V.dom_to_cdom[V.line_dom[gl]]; my_div = V.line_diverse[gl]; my_dst = my_div ? V.dv_off[gl] : V.nd_off[gl]; }
        }
        const uint32_t cnt = min (32u, gend - g0);
        for (uint32_t t = 0; t < cnt; t++) {
            const uint32_t len = __shfl_sync (0xffffffffu, my_len, t);
            if (!len) continue;
            const uint32_t cdom = __shfl_sync (0xffffffffu, my_cdom, t), div = __shfl_sync (0xffffffffu, my_div, t);
            const uint8_t *norm = V.normalize + cdom * NQ;
            const uint8_t *q = V.txt + __shfl_sync (0xffffffffu, my_off, t);
            uint8_t *dst = (div ? V.divr : V.E) + __shfl_sync (0xffffffffu, my_dst, t);
            for (uint32_t i = lane; i < len; i += 32) dst[i] = norm[q[i] - FIRST_Q];
        }
        if (gl < gend && my_len) { V.mplx[V.mx_idx[gl]] = (uint8_t)(my_cdom | (my_div ? 0x80 : 0)); V.line_dom[gl] = (uint8_t)my_cdom; }   // :436,446; ql->dom becomes cdom (:283-285)
    }
}

// ---- pass 3: stream split over the non-diverse concatenation E (:421-500), tile-parallel.  What an element of E emits depends on
// what precedes it only through (a) whether the previous element is a dom, which is E[i-1] itself, and (b) for a non-dom that ends a dom
// run, where the previous non-dom is — inside a tile for every non-dom but the tile's first.  So: k_dqs_count sums each tile's
// QUAL / DOMQRUNS bytes leaving the tile's first run open, k_dqs_scan (one CTA per VBlock) closes the open runs with a running maximum over
// the tiles, turns the sums into offsets and emits the trailing run, and k_dqs_write writes every tile at its offsets.
constexpr uint32_t DQ_TILE = 16384;                   // 256 threads x 64 elements
struct DqTile { uint32_t cq, cr, first, last1; };     // QUAL bytes; DOMQRUNS bytes without the first run's; index of the first non-dom with a run open before it (or ~0); index + 1 of the last non-dom (0 = none)
struct DqTileOff { uint32_t q, r, ln, first; };       // offsets into QUAL / DOMQRUNS, index + 1 of the last non-dom before the tile, DqTile.first

__device__ __forceinline__ uint32_t put_run_bytes (uint8_t *runs, uint32_t pos, uint32_t r)      // codec_domq_add_runs :368-377
{
    while (r) { uint32_t sub = r < 254 ? r : 254; runs[pos++] = (uint8_t)(r <= 254 ? sub : 255); r -= sub; }
    return pos;
}

// The 64 elements of one thread as bit masks (bit j = element i0 + j): `nd` the non-doms, `ends` the non-doms that end a dom run (their
// predecessor is a dom), and the index + 1 of the last non-dom of the tile before them (0 = none in this tile).  256 threads.  sm: 17 words.
struct DqChunk { uint64_t nd, ends; uint32_t i0, tile_last, tile_last_total; };
// bit k = byte k of w is not 0, for bytes <= 0x80 (E holds ranks < 95): no carry leaves a byte; the multiplication gathers bits 0, 8, 16, 24 at 24..27
__device__ __forceinline__ uint32_t dq_nonzero4 (uint32_t w) { return (((((w + 0x7f7f7f7fu) & 0x80808080u) >> 7) * 0x01020408u) >> 24) & 0xf; }
__device__ __forceinline__ uint64_t dq_nonzero16 (const uint4 q) { return dq_nonzero4 (q.x) | (dq_nonzero4 (q.y) << 4) | (dq_nonzero4 (q.z) << 8) | (dq_nonzero4 (q.w) << 12); }
__device__ DqChunk dq_load_chunk (const DqVb &V, uint32_t M, uint32_t base, uint32_t *sm, uint8_t *s_last)
{
    DqChunk c;
    c.i0 = base + threadIdx.x * 64;
    const uint32_t n = c.i0 < M ? min (64u, M - c.i0) : 0;
    c.nd = 0;
    if (n) {                                                             // E is 256-byte aligned and padded by 64
        const uint4 *p = reinterpret_cast<const uint4 *>(V.E + c.i0);
        const uint4 q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3];
        c.nd = dq_nonzero16 (q0) | (dq_nonzero16 (q1) << 16) | (dq_nonzero16 (q2) << 32) | (dq_nonzero16 (q3) << 48);
        if (n < 64) c.nd &= (1ull << n) - 1;
    }
    s_last[threadIdx.x] = (uint8_t)(n < 64 || (c.nd >> 63));              // is my last element a non-dom (past the end counts as one)
    const uint32_t my_last = c.nd ? c.i0 + 64 - __clzll ((long long)c.nd) : 0;
    c.tile_last = block_excl_max (my_last, sm, &c.tile_last_total);      // (its barriers also publish s_last)
    const uint32_t prev_dom = threadIdx.x ? !s_last[threadIdx.x - 1] : (base ? V.E[base - 1] == 0 : 0);
    c.ends = c.nd & ~((c.nd << 1) | (prev_dom ? 0 : 1));
    return c;
}
// QUAL bytes << 16 | DOMQRUNS bytes of the chunk.  A run that ends at a non-dom of this chunk began inside the chunk (one byte) unless it is
// the chunk's first non-dom: then it reaches back to `ln`, or — ln = 0 — to an earlier tile: *open = its index, not counted.
__device__ __forceinline__ uint32_t dq_chunk_bytes (const DqChunk &c, uint32_t ln, uint32_t *open)
{
    const uint32_t n_ends = __popcll (c.ends);
    uint32_t cq = 2 * __popcll (c.nd) - n_ends, cr = n_ends;
    if (c.ends && (c.ends & (0ull - c.ends)) == (c.nd & (0ull - c.nd))) {     // the first run end is the first non-dom
        const uint32_t idx = c.i0 + __ffsll ((long long)c.ends) - 1;
        if (ln) cr += (idx - ln + 253) / 254 - 1;
        else { cr -= 1; *open = idx; }
    }
    return (cq << 16) | cr;
}

__global__ void __launch_bounds__(256) k_dqs_count (const DqVb *vbs)
{
    const DqVb &V = vbs[blockIdx.y];
    const uint32_t M = V.lens[4], base = blockIdx.x * DQ_TILE;
    if (base >= M) return;
    __shared__ uint32_t sm[17];
    __shared__ uint8_t s_last[256];
    __shared__ uint32_t s_first;
    if (threadIdx.x == 0) s_first = 0xffffffffu;
    const DqChunk c = dq_load_chunk (V, M, base, sm, s_last);
    uint32_t open = 0xffffffffu, tot;
    const uint32_t mine = dq_chunk_bytes (c, c.tile_last, &open);
    if (open != 0xffffffffu) s_first = open;                              // at most one thread: the one with the tile's first non-dom
    block_excl_sum (mine, sm, &tot);
    if (threadIdx.x == 0) reinterpret_cast<DqTile *>(V.tiles)[blockIdx.x] = DqTile { tot >> 16, tot & 0xffff, s_first, c.tile_last_total };
}

__global__ void __launch_bounds__(512) k_dqs_scan (const DqVb *vbs)
{
    const DqVb &V = vbs[blockIdx.x];
    const DqTile *T = reinterpret_cast<const DqTile *>(V.tiles); DqTileOff *O = reinterpret_cast<DqTileOff *>(V.toff);
    __shared__ uint32_t sm[17];
    const uint32_t M = V.lens[4], n_tiles = (M + DQ_TILE - 1) / DQ_TILE;
    uint32_t qpos = 0, rpos = 0, last_nd1 = 0, t;
    for (uint32_t b = 0; b < n_tiles; b += 512) {
        const uint32_t k = b + threadIdx.x;
        DqTile d = k < n_tiles ? T[k] : DqTile { 0, 0, 0xffffffffu, 0 };
        const uint32_t ln = max (block_excl_max (d.last1, sm, &t), last_nd1); const uint32_t tl = t;
        if (d.first != 0xffffffffu) d.cr += (d.first - ln + 253) / 254;
        const uint32_t oq = block_excl_sum (d.cq, sm, &t) + qpos; const uint32_t tq = t;
        const uint32_t orr = block_excl_sum (d.cr, sm, &t) + rpos; const uint32_t tr = t;
        if (k < n_tiles) O[k] = DqTileOff { oq, orr, ln, d.first };
        qpos += tq; rpos += tr; last_nd1 = max (last_nd1, tl);
    }
    if (threadIdx.x == 0) {
        uint32_t runlen = M - last_nd1;                                            // trailing dom run (:473-482)
        if (runlen && (rpos || runlen < V.lens[5])) { rpos = put_run_bytes (V.runs, rpos, runlen); V.qual[qpos++] = V.no_doms; }
        if (!qpos) V.qual[qpos++] = 'X';                                           // :497-500
        V.lens[0] = qpos; V.lens[1] = rpos;
    }
}

__global__ void __launch_bounds__(256) k_dqs_write (const DqVb *vbs)
{
    const DqVb &V = vbs[blockIdx.y];
    const uint32_t M = V.lens[4], base = blockIdx.x * DQ_TILE;
    if (base >= M) return;
    __shared__ uint32_t sm[17];
    __shared__ uint8_t s_last[256];
    const DqTileOff o = reinterpret_cast<const DqTileOff *>(V.toff)[blockIdx.x];
    const uint8_t no_doms = V.no_doms;
    const DqChunk c = dq_load_chunk (V, M, base, sm, s_last);
    // as counted by k_dqs_count (the tile's open run left out), so that the offsets agree: that run's bytes go to every chunk after it
    uint32_t open = 0xffffffffu, t;
    const uint32_t mine = dq_chunk_bytes (c, c.tile_last, &open);
    const uint32_t ex = block_excl_sum (mine, sm, &t);
    uint32_t oq = o.q + (ex >> 16), orr = o.r + (ex & 0xffff);
    if (o.first != 0xffffffffu && o.first < c.i0) orr += (o.first - o.ln + 253) / 254;
    uint32_t ln = max (c.tile_last, o.ln);
    for (uint64_t m = c.nd; m; m &= m - 1) {
        const int j = __ffsll ((long long)m) - 1;
        if ((c.ends >> j) & 1) orr = put_run_bytes (V.runs, orr, c.i0 + j - ln);
        else V.qual[oq++] = no_doms;
        V.qual[oq++] = V.E[c.i0 + j];                                     // (in L1 since dq_load_chunk)
        ln = c.i0 + j + 1;
    }
}

// ================================================================================================ DOMQ PIZ
struct DqPiz {
    const uint8_t *qual, *runs, *mplx, *divr;
    const uint32_t *line_len;
    uint8_t  *out;
    uint8_t  *E;               // M bytes, zero-filled (= dom rank) before literals are scattered
    uint32_t *cum;             // inclusive prefix of decoded run lengths, one per run
    uint32_t *nd_off, *dv_off, *out_off;
    uint8_t  *dom_i;
    uint32_t *info;            // [0] M, [1] D, [2] err, [3] n_runs
    uint32_t  qual_len, runs_len, mplx_len, divr_len, n_lines;
    uint8_t   no_dom;
    uint8_t   pad[3];
    uint8_t   denorm[NQ * NQ];
    uint32_t  denorm_len;
};

// per-line: mux byte (:790-791), diverse flag (:795-796), offsets.  One CTA per VB.
__global__ void __launch_bounds__(512) k_dqp_lines (const DqPiz *vbs)
{
    const DqPiz &V = vbs[blockIdx.x];
    __shared__ uint32_t sm[17];
    uint32_t nd = 0, dv = 0, mx = 0, oo = 0, t;
    for (uint32_t base = 0; base < V.n_lines; base += 512) {
        uint32_t li = base + threadIdx.x;
        uint32_t len = li < V.n_lines ? V.line_len[li] : 0;
        uint32_t c = block_excl_sum (len ? 1 : 0, sm, &t); uint32_t tc = t;
        uint8_t d = 0;
        if (len) { uint32_t k = (V.mplx_len == 1) ? 0 : mx + c; d = k < V.mplx_len ? V.mplx[k] : 0; if (k >= V.mplx_len) V.info[2] = 1; }
        bool div = len && (d >> 7);
        uint32_t a = block_excl_sum (div ? 0 : len, sm, &t); uint32_t ta = t;
        uint32_t b = block_excl_sum (div ? len : 0, sm, &t); uint32_t tb = t;
        uint32_t o = block_excl_sum (len, sm, &t);           uint32_t to = t;
        if (li < V.n_lines) { V.nd_off[li] = nd + a; V.dv_off[li] = dv + b; V.out_off[li] = oo + o; V.dom_i[li] = d; }
        nd += ta; dv += tb; mx += tc; oo += to;
    }
    if (threadIdx.x == 0) { V.info[0] = nd; V.info[1] = dv; if (dv > V.divr_len) V.info[2] = 1; }
}

// DOMQRUNS bytes -> inclusive prefix of run lengths (a run = 255* then one byte != 255; value 254*(n-1)+last, :560-565)
__global__ void __launch_bounds__(512) k_dqp_runs (const DqPiz *vbs)
{
    const DqPiz &V = vbs[blockIdx.x];
    __shared__ uint32_t sm[17];
    uint32_t nruns = 0, start1 = 0, cum = 0, t;          // start1 = (index after the previous run end)
    for (uint32_t base = 0; base < V.runs_len; base += 512) {
        uint32_t i = base + threadIdx.x;
        uint32_t b = i < V.runs_len ? V.runs[i] : 255;
        bool end = i < V.runs_len && b != 255;
        uint32_t k = block_excl_sum (end ? 1 : 0, sm, &t); uint32_t tk = t;
        uint32_t tm;
        uint32_t st = max (block_excl_max (end ? i + 1 : 0, sm, &tm), start1);
        uint32_t val = end ? 254 * (i - st) + b : 0;
        uint32_t tv;
        uint32_t cv = block_excl_sum (val, sm, &tv);
        if (end) V.cum[nruns + k] = cum + cv + val;
        nruns += tk; cum += tv; start1 = max (start1, tm);
    }
    if (threadIdx.x == 0) V.info[3] = nruns;
}

// QUAL bytes -> scatter literals into E (:693-735).  A byte equal to no_dom is a marker; a literal not preceded by a
// marker consumes the next run; a marker in the very last position is the final-run indicator.
__global__ void __launch_bounds__(512) k_dqp_literals (const DqPiz *vbs)
{
    const DqPiz &V = vbs[blockIdx.x];
    __shared__ uint32_t sm[17];
    const uint32_t M = V.info[0], nruns = V.info[3], no_dom = V.no_dom;
    // no DOMQRUNS at all: only (marker, literal) pairs are consumed (:676-690) — a lone 'X' is not a literal
    const uint32_t qn = V.qual_len;
    uint32_t lits = 0, cons = 0, t;
    for (uint32_t base = 0; base < qn; base += 512) {
        uint32_t j = base + threadIdx.x;
        uint32_t b = j < qn ? V.qual[j] : no_dom;
        bool prev_marker = j > 0 && j < qn && V.qual[j - 1] == no_dom;
        bool is_lit = j < qn && b != no_dom && (V.runs_len || prev_marker);
        bool consumes = V.runs_len && is_lit && !prev_marker;
        uint32_t li = block_excl_sum (is_lit ? 1 : 0, sm, &t); uint32_t tl = t;
        uint32_t ci = block_excl_sum (consumes ? 1 : 0, sm, &t); uint32_t tc = t;
        if (is_lit) {
            uint32_t k = cons + ci;                                               // runs consumed before this literal
            uint32_t before = consumes ? (k < nruns ? V.cum[k] : 0xffffffffu) : (k ? (k - 1 < nruns ? V.cum[k - 1] : 0xffffffffu) : 0);
            uint64_t pos = (uint64_t)lits + li + before;
            if (before == 0xffffffffu || pos >= M) V.info[2] = 1;
            else V.E[pos] = (uint8_t)b;
        }
        lits += tl; cons += tc;
    }
}

// de-normalise every line (:698,731 and reconstruct_divr :752-766): one warp per line
__global__ void __launch_bounds__(256) k_dqp_denorm (const DqPiz *vbs, const uint32_t *blk_vb, const uint32_t *blk_first)
{
    const DqPiz &V = vbs[blk_vb[blockIdx.x]];
    const uint32_t first = blk_first[blockIdx.x], last = min (first + LINES_PER_BLOCK, V.n_lines);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (V.info[2]) return;
    for (uint32_t li = first + warp; li < last; li += 8) {
        const uint32_t len = V.line_len[li];
        if (!len) continue;
        const uint32_t d = V.dom_i[li];
        const uint32_t row = (d & 0x7f) * V.no_dom;
        if (row + V.no_dom > V.denorm_len) { V.info[2] = 1; continue; }
        const uint8_t *dn = V.denorm + row;
        const uint8_t *src = (d >> 7) ? V.divr + V.dv_off[li] : V.E + V.nd_off[li];
        uint8_t *dst = V.out + V.out_off[li];
        for (uint32_t i = lane; i < len; i += 32) { uint32_t r = src[i]; dst[i] = r < V.no_dom ? dn[r] : 0; }
    }
}

void line_blocks (const std::vector<uint32_t> &n_lines, std::vector<uint32_t> &bvb, std::vector<uint32_t> &bfirst)
{
    for (uint32_t v = 0; v < n_lines.size (); v++)
        for (uint32_t f = 0; f < n_lines[v]; f += LINES_PER_BLOCK) { bvb.push_back (v); bfirst.push_back (f); }
}

typedef struct { uint8_t q; uint32_t count; } QMap;
int qmap_desc (const void *a, const void *b)                                       // DESCENDING_SORTER (sorter.h:16-33)
{
    uint32_t ca = ((const QMap *)a)->count, cb = ((const QMap *)b)->count;
    return -((ca > cb) ? 1 : (ca < cb) ? -1 : 0);
}

} // namespace

// ================================================================================================ C-ABI: ACGT
extern "C" uint64_t gzb_acgt_packed_len (uint64_t n) { return ((2 * n + 63) / 64) * 8; }

extern "C" int gzb_acgt_pack (gzb_engine *e, const void *seq, uint64_t n, void *packed, void *x, int *x_all_zero, uint32_t flags)
{
    if (!e || (!seq && n) || !packed) return GZB_E_BADARG;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS, xdev = !devptr && (flags & GZB_OUT_DEVICE) && x;   // exception stream stays in HBM for its sub-codec
    const uint64_t plen = gzb_acgt_packed_len (n);
    cudaStream_t st = e->stream;
    Carver c { nullptr, 0 };
    uint8_t *d_seq = nullptr, *d_packed = nullptr, *d_x = nullptr; int *d_flag = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_flag = c.take<int> (1);
        if (!devptr) { d_seq = c.take<uint8_t> (n + 32); d_packed = c.take<uint8_t> (plen + 32); d_x = xdev ? (uint8_t *)x : c.take<uint8_t> (n + 32); }
        if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
    }
    if (devptr) { d_seq = (uint8_t *)seq; d_packed = (uint8_t *)packed; d_x = (uint8_t *)x; }
    else if (n) CK (cudaMemcpyAsync (d_seq, seq, n, cudaMemcpyHostToDevice, st));
    CK (cudaMemsetAsync (d_flag, 0, sizeof (int), st));
    if (n) {
        uint64_t nwords = plen / 8;
        uint32_t grid = (uint32_t)std::min<uint64_t> ((nwords + 255) / 256, (uint64_t)e->sm_count * 16);
        k_acgt_pack<<<grid, 256, 0, st>>>(d_seq, n, reinterpret_cast<uint64_t *>(d_packed), (devptr && !x) ? nullptr : d_x, d_flag);
        e->launches++;
    }
    int h_flag = 0;
    CK (cudaMemcpyAsync (&h_flag, d_flag, sizeof (int), cudaMemcpyDeviceToHost, st));
    if (!devptr && plen) CK (cudaMemcpyAsync (packed, d_packed, plen, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    if (x_all_zero) *x_all_zero = !h_flag;
    if (!devptr && !xdev && x && n && h_flag) { CK (cudaMemcpyAsync (x, d_x, n, cudaMemcpyDeviceToHost, st)); CK (cudaStreamSynchronize (st)); }
    else if (!devptr && !xdev && x && n) memset (x, 0, n);                                   // all-zero exception stream: no transfer needed
    return GZB_OK;
}

extern "C" int gzb_acgt_unpack (gzb_engine *e, const void *packed, const void *x, uint64_t n, void *seq, uint32_t flags)
{
    if (!e || (!packed && n) || !seq) return GZB_E_BADARG;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS, xdev = !devptr && (flags & GZB_IN_DEVICE) && x;
    const uint64_t plen = gzb_acgt_packed_len (n);
    cudaStream_t st = e->stream;
    Carver c { nullptr, 0 };
    uint8_t *d_seq = nullptr, *d_packed = nullptr, *d_x = nullptr;
    if (!devptr) {
        for (int pass = 0; pass < 2; pass++) {
            c.off = 0;
            d_seq = c.take<uint8_t> (n + 32); d_packed = c.take<uint8_t> (plen + 32); d_x = c.take<uint8_t> (n + 32);
            if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
        }
        if (plen) CK (cudaMemcpyAsync (d_packed, packed, plen, cudaMemcpyHostToDevice, st));
        if (xdev) d_x = (uint8_t *)x;
        else if (x && n) CK (cudaMemcpyAsync (d_x, x, n, cudaMemcpyHostToDevice, st));
    }
    else { d_seq = (uint8_t *)seq; d_packed = (uint8_t *)packed; d_x = (uint8_t *)x; }
    if (n) {
        uint64_t nwords = plen / 8;
        uint32_t grid = (uint32_t)std::min<uint64_t> ((nwords + 255) / 256, (uint64_t)e->sm_count * 16);
        k_acgt_unpack<<<grid, 256, 0, st>>>(reinterpret_cast<const uint64_t *>(d_packed), x ? d_x : nullptr, n, d_seq);
        e->launches++;
    }
    if (!devptr && n) CK (cudaMemcpyAsync (seq, d_seq, n, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    return GZB_OK;
}

// Batched forms: every VBlock of the batch in ONE launch (and one stream synchronisation) instead of one per VBlock.
// Host-pointer mode stages the whole batch in the engine workspace; with GZB_OUT_DEVICE (pack) / GZB_IN_DEVICE (unpack)
// the exception streams stay in the caller's device buffers, as in the single-VBlock calls.
extern "C" int gzb_acgt_pack_batch (gzb_engine *e, gzb_acgt_vb *vbs, uint32_t n_vbs, uint32_t flags)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS, xdev_f = !devptr && (flags & GZB_OUT_DEVICE);
    cudaStream_t st = e->stream;
    Carver c { nullptr, 0 };
    AcgtD *d_desc = nullptr; int *d_flags = nullptr;
    std::vector<AcgtD> h (n_vbs);
    uint64_t max_words = 0;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_desc = c.take<AcgtD> (n_vbs); d_flags = c.take<int> (n_vbs);
        for (uint32_t v = 0; v < n_vbs; v++) {
            const gzb_acgt_vb &a = vbs[v];
            if ((!a.seq && a.n_bases) || !a.packed) return GZB_E_BADARG;
            const uint64_t plen = gzb_acgt_packed_len (a.n_bases);
            max_words = std::max<uint64_t> (max_words, plen / 8);
            AcgtD &D = h[v];
            D.n = a.n_bases;
            if (devptr) { D.seq = (const uint8_t *)a.seq; D.packed = (uint64_t *)a.packed; D.x = (uint8_t *)a.x; }
            else {
                D.seq = c.take<uint8_t> (a.n_bases + 32); D.packed = reinterpret_cast<uint64_t *>(c.take<uint8_t> (plen + 32));
                D.x = (xdev_f && a.x) ? (uint8_t *)a.x : c.take<uint8_t> (a.n_bases + 32);
            }
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, (size_t)n_vbs * (sizeof (AcgtD) + sizeof (int)) + 4096); if (rc) return rc; c.base = e->ws; }
    }
    for (uint32_t v = 0; v < n_vbs; v++) h[v].flag = d_flags + v;
    AcgtD *p_desc = reinterpret_cast<AcgtD *>(e->pin); int *p_flags = reinterpret_cast<int *>(e->pin + (size_t)n_vbs * sizeof (AcgtD));
    memcpy (p_desc, h.data (), (size_t)n_vbs * sizeof (AcgtD));
    CK (cudaMemcpyAsync (d_desc, p_desc, (size_t)n_vbs * sizeof (AcgtD), cudaMemcpyHostToDevice, st));
    CK (cudaMemsetAsync (d_flags, 0, (size_t)n_vbs * sizeof (int), st));
    if (!devptr) for (uint32_t v = 0; v < n_vbs; v++) if (vbs[v].n_bases) CK (cudaMemcpyAsync (const_cast<uint8_t *>(h[v].seq), vbs[v].seq, vbs[v].n_bases, cudaMemcpyHostToDevice, st));
    if (max_words) {
        dim3 grid ((uint32_t)std::min<uint64_t> ((max_words + 255) / 256, 4096), n_vbs);
        k_acgt_pack_batch<<<grid, 256, 0, st>>>(d_desc);
        e->launches++;
    }
    CK (cudaMemcpyAsync (p_flags, d_flags, (size_t)n_vbs * sizeof (int), cudaMemcpyDeviceToHost, st));
    if (!devptr) for (uint32_t v = 0; v < n_vbs; v++) {
        const uint64_t plen = gzb_acgt_packed_len (vbs[v].n_bases);
        if (plen) CK (cudaMemcpyAsync (vbs[v].packed, h[v].packed, plen, cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    bool more = false;
    for (uint32_t v = 0; v < n_vbs; v++) {
        vbs[v].x_all_zero = !p_flags[v];
        const bool xdev = xdev_f && vbs[v].x;
        if (!devptr && !xdev && vbs[v].x && vbs[v].n_bases) {
            if (p_flags[v]) { CK (cudaMemcpyAsync (vbs[v].x, h[v].x, vbs[v].n_bases, cudaMemcpyDeviceToHost, st)); more = true; }
            else memset (vbs[v].x, 0, vbs[v].n_bases);                       // all-zero exception stream: no transfer needed
        }
    }
    if (more) CK (cudaStreamSynchronize (st));
    return GZB_OK;
}

extern "C" int gzb_acgt_unpack_batch (gzb_engine *e, const gzb_acgt_vb *vbs, uint32_t n_vbs, uint32_t flags)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS, xdev_f = !devptr && (flags & GZB_IN_DEVICE);
    cudaStream_t st = e->stream;
    Carver c { nullptr, 0 };
    AcgtD *d_desc = nullptr;
    std::vector<AcgtD> h (n_vbs);
    uint64_t max_words = 0;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_desc = c.take<AcgtD> (n_vbs);
        for (uint32_t v = 0; v < n_vbs; v++) {
            const gzb_acgt_vb &a = vbs[v];
            if ((!a.packed && a.n_bases) || !a.seq) return GZB_E_BADARG;
            const uint64_t plen = gzb_acgt_packed_len (a.n_bases);
            max_words = std::max<uint64_t> (max_words, plen / 8);
            AcgtD &D = h[v];
            D.n = a.n_bases; D.flag = nullptr;
            if (devptr) { D.seq = (const uint8_t *)a.seq; D.packed = (uint64_t *)a.packed; D.x = (uint8_t *)a.x; }
            else {
                D.seq = c.take<uint8_t> (a.n_bases + 32); D.packed = reinterpret_cast<uint64_t *>(c.take<uint8_t> (plen + 32));
                D.x = !a.x ? nullptr : xdev_f ? (uint8_t *)a.x : c.take<uint8_t> (a.n_bases + 32);
            }
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, (size_t)n_vbs * sizeof (AcgtD) + 4096); if (rc) return rc; c.base = e->ws; }
    }
    AcgtD *p_desc = reinterpret_cast<AcgtD *>(e->pin);
    memcpy (p_desc, h.data (), (size_t)n_vbs * sizeof (AcgtD));
    CK (cudaMemcpyAsync (d_desc, p_desc, (size_t)n_vbs * sizeof (AcgtD), cudaMemcpyHostToDevice, st));
    if (!devptr) for (uint32_t v = 0; v < n_vbs; v++) {
        const uint64_t plen = gzb_acgt_packed_len (vbs[v].n_bases);
        if (plen) CK (cudaMemcpyAsync (h[v].packed, vbs[v].packed, plen, cudaMemcpyHostToDevice, st));
        if (vbs[v].x && !xdev_f && vbs[v].n_bases) CK (cudaMemcpyAsync (h[v].x, vbs[v].x, vbs[v].n_bases, cudaMemcpyHostToDevice, st));
    }
    if (max_words) {
        dim3 grid ((uint32_t)std::min<uint64_t> ((max_words + 255) / 256, 4096), n_vbs);
        k_acgt_unpack_batch<<<grid, 256, 0, st>>>(d_desc);
        e->launches++;
    }
    if (!devptr) for (uint32_t v = 0; v < n_vbs; v++) if (vbs[v].n_bases) CK (cudaMemcpyAsync (const_cast<void *>(vbs[v].seq), h[v].seq, vbs[v].n_bases, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    return GZB_OK;
}

// ================================================================================================ C-ABI: DOMQ (ZIP)
// The device state of a prepare call (staged text, line tables, per-line dom/diverse) lives in the engine's DOMQ
// session buffer until the matching split call, so a host-buffer caller uploads its quality text once.
namespace {

struct DqLayout {
    std::vector<DqVb> h;           // host copies of the device descriptors
    DqVb *d_vbs = nullptr;
    uint32_t *d_bvb = nullptr, *d_bfirst = nullptr; uint32_t n_blocks = 0, max_tiles = 0;
    uint32_t *d_hist = nullptr, *d_lens = nullptr;   // [n_vbs][NQ*NQ+NQ] histograms, [n_vbs][8] stream lengths
    size_t pin_vbs = 0, pin_vbs_bytes = 0, pin_blocks = 0, pin_landing = 0;   // offsets into the engine's pinned staging: descriptors | line blocks | histograms, lengths
    std::vector<uint8_t *> d_linedom, d_linediv;
};

int dq_stage (gzb_engine *e, gzb_domq_vb *vbs, uint32_t n_vbs, bool devptr, bool outdev, DqLayout &L)
{
    std::vector<uint32_t> nl (n_vbs), bvb, bfirst;
    for (uint32_t v = 0; v < n_vbs; v++) nl[v] = vbs[v].n_lines;
    line_blocks (nl, bvb, bfirst);
    L.n_blocks = (uint32_t)bvb.size ();
    Carver c { nullptr, 0 };
    L.h.assign (n_vbs, DqVb ());
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        L.d_vbs = c.take<DqVb> (n_vbs);
        L.d_bvb = c.take<uint32_t> (bvb.size () + 1); L.d_bfirst = c.take<uint32_t> (bfirst.size () + 1);
        L.d_hist = c.take<uint32_t> ((size_t)n_vbs * (NQ * NQ + NQ));       // the histograms and the stream lengths of all VBlocks, contiguous: one memset, one transfer
        L.d_lens = c.take<uint32_t> ((size_t)n_vbs * 8);
        for (uint32_t v = 0; v < n_vbs; v++) {
            DqVb &D = L.h[v]; const gzb_domq_vb &S = vbs[v];
            uint64_t tot = 0;   // total quality bytes is unknown on the host in device mode: bound by txt_len
            tot = S.txt_len;
            D.n_lines = S.n_lines;
            D.txt      = devptr ? (const uint8_t *)S.txt : c.take<uint8_t> (S.txt_len + 16);
            D.line_off = devptr ? S.line_off : c.take<uint64_t> (S.n_lines + 1);
            D.line_len = devptr ? S.line_len : c.take<uint32_t> (S.n_lines + 1);
            D.line_dom = devptr ? S.line_dom : c.take<uint8_t> (S.n_lines + 1);
            D.line_diverse = devptr ? S.line_diverse : c.take<uint8_t> (S.n_lines + 1);
            D.hist   = L.d_hist ? L.d_hist + (size_t)v * (NQ * NQ + NQ) : nullptr;
            D.nd_off = c.take<uint32_t> (S.n_lines + 1); D.dv_off = c.take<uint32_t> (S.n_lines + 1); D.mx_idx = c.take<uint32_t> (S.n_lines + 1);
            D.E      = c.take<uint8_t> (tot + 64);
            D.tiles  = c.take<DqTile> (tot / DQ_TILE + 2); D.toff = c.take<DqTileOff> (tot / DQ_TILE + 2);
            L.max_tiles = std::max (L.max_tiles, (uint32_t)(tot / DQ_TILE + 1));
            D.lens   = L.d_lens ? L.d_lens + (size_t)v * 8 : nullptr;
            D.qual   = (devptr || outdev) ? (uint8_t *)S.qual : c.take<uint8_t> (2 * tot + 16);
            D.runs   = (devptr || outdev) ? (uint8_t *)S.runs : c.take<uint8_t> (tot + 16);
            D.mplx   = (devptr || outdev) ? (uint8_t *)S.mplx : c.take<uint8_t> (S.n_lines + 16);
            D.divr   = (devptr || outdev) ? (uint8_t *)S.divr : c.take<uint8_t> (tot + 16);
        }
        if (pass == 0) {
            if (c.off > e->dq_cap) {
                CK (cudaStreamSynchronize (e->stream));
                if (e->dq_buf) cudaFree (e->dq_buf);
                e->dq_buf = nullptr; e->dq_cap = 0;
                CK (cudaMalloc (&e->dq_buf, c.off + (c.off >> 3)));
                e->dq_cap = c.off + (c.off >> 3);
            }
            c.base = e->dq_buf;
        }
    }
    // Descriptors go up through the engine's page-locked staging: a transfer from pageable memory is staged by the driver and waits
    // behind whatever bulk upload is in progress on the link (measured beside a staged upload: this call 76 ms instead of 11 ms).
    cudaStream_t st = e->stream;
    const size_t nb4 = bvb.size () * 4, nb4a = (nb4 + 255) & ~(size_t)255;
    L.pin_vbs = 0; L.pin_vbs_bytes = ((size_t)n_vbs * sizeof (DqVb) + 255) & ~(size_t)255;
    L.pin_blocks = L.pin_vbs + L.pin_vbs_bytes; L.pin_landing = L.pin_blocks + 2 * nb4a;
    int rc = engine_reserve (e, 0, L.pin_landing + (size_t)n_vbs * ((NQ * NQ + NQ) * 4 + 32) + 512); if (rc) return rc;
    if (nb4) {
        memcpy (e->pin + L.pin_blocks, bvb.data (), nb4); memcpy (e->pin + L.pin_blocks + nb4a, bfirst.data (), nb4);
        CK (cudaMemcpyAsync (L.d_bvb, e->pin + L.pin_blocks, nb4, cudaMemcpyHostToDevice, st));
        CK (cudaMemcpyAsync (L.d_bfirst, e->pin + L.pin_blocks + nb4a, nb4, cudaMemcpyHostToDevice, st));
    }
    return GZB_OK;
}

} // namespace

extern "C" int gzb_domq_prepare (gzb_engine *e, gzb_domq_vb *vbs, uint32_t n_vbs, uint32_t flags)
{
    if (!e || !vbs) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    if (n_vbs > 65535) { e->err = "gzb_domq_prepare: at most 65535 VBlocks per call"; return GZB_E_BADARG; }
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS, outdev = flags & GZB_OUT_DEVICE;
    cudaStream_t st = e->stream;
    DqLayout *L = new DqLayout ();
    delete reinterpret_cast<DqLayout *>(e->dq_session); e->dq_session = L;
    e->dq_free = [] (void *p) { delete reinterpret_cast<DqLayout *>(p); };
    int rc = dq_stage (e, vbs, n_vbs, devptr, outdev, *L);
    if (rc) return rc;
    for (uint32_t v = 0; v < n_vbs; v++) {
        DqVb &D = L->h[v];
        if (!devptr) {
            if (vbs[v].txt_len) CK (cudaMemcpyAsync ((void *)D.txt, vbs[v].txt, vbs[v].txt_len, cudaMemcpyHostToDevice, st));
            if (D.n_lines) {
                CK (cudaMemcpyAsync ((void *)D.line_off, vbs[v].line_off, (size_t)D.n_lines * 8, cudaMemcpyHostToDevice, st));
                CK (cudaMemcpyAsync ((void *)D.line_len, vbs[v].line_len, (size_t)D.n_lines * 4, cudaMemcpyHostToDevice, st));
            }
        }
    }
    const size_t hist_bytes = (size_t)n_vbs * (NQ * NQ + NQ) * 4;
    uint32_t *const hist = reinterpret_cast<uint32_t *>(e->pin + L->pin_landing), *const lens = reinterpret_cast<uint32_t *>(e->pin + L->pin_landing + hist_bytes);
    CK (cudaMemsetAsync (L->d_hist, 0, hist_bytes, st));
    memcpy (e->pin + L->pin_vbs, L->h.data (), n_vbs * sizeof (DqVb));
    CK (cudaMemcpyAsync (L->d_vbs, e->pin + L->pin_vbs, n_vbs * sizeof (DqVb), cudaMemcpyHostToDevice, st));
    if (L->n_blocks) { k_domq_linehist<<<L->n_blocks, 128, 0, st>>>(L->d_vbs, L->d_bvb, L->d_bfirst); e->launches++; }
    CK (cudaMemcpyAsync (hist, L->d_hist, hist_bytes, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));

    // host: compaction (:180-197) and per-dom rank tables (:199-247).  qsort's order among equal counts is whatever
    // this libc does — the same call with the same comparator the reference makes.
    for (uint32_t v = 0; v < n_vbs; v++) {
        uint32_t (*H)[NQ] = reinterpret_cast<uint32_t (*)[NQ]>(hist + (size_t)v * (NQ * NQ + NQ));
        const uint32_t *lwd = hist + (size_t)v * (NQ * NQ + NQ) + NQ * NQ;
        gzb_domq_vb &S = vbs[v]; DqVb &D = L->h[v];
        memset (S.denorm, 0, sizeof S.denorm); memset (S.normalize, 0, sizeof S.normalize); memset (D.dom_to_cdom, 0, NQ);
        int num_doms = 0;
        for (int k = 0; k < NQ; k++) if (lwd[k]) { D.dom_to_cdom[k] = (uint8_t)num_doms; if (num_doms != k) memcpy (H[num_doms], H[k], sizeof H[0]); num_doms++; }
        uint8_t denormalize[NQ][NQ]; memset (denormalize, 0, sizeof denormalize);
        int num_norm_qs = 0;
        for (int cd = 0; cd < num_doms; cd++) {
            QMap m[NQ];
            for (int k = 0; k < NQ; k++) { m[k].q = (uint8_t)k; m[k].count = H[cd][k]; }
            qsort (m, NQ, sizeof (QMap), qmap_desc);
            int r = 0;
            for (; r < NQ && m[r].count; r++) { S.normalize[cd * NQ + m[r].q] = (uint8_t)r; denormalize[cd][r] = (uint8_t)(m[r].q + FIRST_Q); }
            if (r > num_norm_qs) num_norm_qs = r;
        }
        for (int cd = 0; cd < num_doms; cd++) for (int r = 0; r < num_norm_qs; r++) S.denorm[cd * num_norm_qs + r] = denormalize[cd][r];
        S.num_norm_qs = (uint8_t)num_norm_qs; S.num_doms = (uint8_t)num_doms;
        D.no_doms = (uint8_t)num_norm_qs;
        memcpy (D.normalize, S.normalize, sizeof D.normalize);
    }
    // per-line outputs: compacted dom + diverse flag
    memcpy (e->pin + L->pin_vbs, L->h.data (), n_vbs * sizeof (DqVb));       // (the first upload of the descriptors completed before the histograms came back)
    CK (cudaMemcpyAsync (L->d_vbs, e->pin + L->pin_vbs, n_vbs * sizeof (DqVb), cudaMemcpyHostToDevice, st));
    k_domq_lineoffsets<<<n_vbs, 512, 0, st>>>(L->d_vbs); e->launches++;
    if (L->n_blocks) { k_domq_normalize<<<L->n_blocks, 256, 0, st>>>(L->d_vbs, L->d_bvb, L->d_bfirst); e->launches++; }
    CK (cudaMemcpyAsync (lens, L->d_lens, (size_t)n_vbs * 32, cudaMemcpyDeviceToHost, st));
    for (uint32_t v = 0; v < n_vbs; v++) {
        if (!devptr && vbs[v].n_lines) {
            if (vbs[v].line_dom) CK (cudaMemcpyAsync (vbs[v].line_dom, L->h[v].line_dom, vbs[v].n_lines, cudaMemcpyDeviceToHost, st));
            if (vbs[v].line_diverse) CK (cudaMemcpyAsync (vbs[v].line_diverse, L->h[v].line_diverse, vbs[v].n_lines, cudaMemcpyDeviceToHost, st));
        }
    }
    CK (cudaStreamSynchronize (st));
    for (uint32_t v = 0; v < n_vbs; v++) {
        vbs[v].has_diverse = lens[(size_t)v * 8 + 3] != 0;
        vbs[v].mplx_len = lens[(size_t)v * 8 + 2]; vbs[v].divr_len = lens[(size_t)v * 8 + 3];
    }
    L->h.shrink_to_fit ();
    e->dq_n_vbs = n_vbs; e->dq_devptr = devptr;
    return GZB_OK;
}

extern "C" int gzb_domq_split (gzb_engine *e, gzb_domq_vb *vbs, uint32_t n_vbs, uint32_t flags)
{
    if (!e || !vbs) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS, outdev = flags & GZB_OUT_DEVICE;
    DqLayout *L = reinterpret_cast<DqLayout *>(e->dq_session);
    if (!L || e->dq_n_vbs != n_vbs || e->dq_devptr != devptr) { e->err = "gzb_domq_split must follow gzb_domq_prepare on the same batch"; return GZB_E_BADARG; }
    cudaStream_t st = e->stream;
    k_dqs_count<<<dim3 (L->max_tiles, n_vbs), 256, 0, st>>>(L->d_vbs);
    k_dqs_scan<<<n_vbs, 512, 0, st>>>(L->d_vbs);
    k_dqs_write<<<dim3 (L->max_tiles, n_vbs), 256, 0, st>>>(L->d_vbs); e->launches += 3;
    int rc = engine_reserve (e, 0, (size_t)n_vbs * 32 + 256); if (rc) return rc;
    uint32_t *const lens = reinterpret_cast<uint32_t *>(e->pin);
    CK (cudaMemcpyAsync (lens, L->d_lens, (size_t)n_vbs * 32, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_domq_vb &S = vbs[v]; const uint32_t *l = lens + (size_t)v * 8;
        S.qual_len = l[0]; S.runs_len = l[1]; S.mplx_len = l[2]; S.divr_len = l[3];
        if (!devptr && !outdev) {
            if (S.qual_len > S.qual_cap || S.runs_len > S.runs_cap || S.mplx_len > S.mplx_cap || S.divr_len > S.divr_cap) { e->err = "DOMQ output capacity too small"; return GZB_E_BADARG; }
            if (S.qual_len) CK (cudaMemcpyAsync (S.qual, L->h[v].qual, S.qual_len, cudaMemcpyDeviceToHost, st));
            if (S.runs_len) CK (cudaMemcpyAsync (S.runs, L->h[v].runs, S.runs_len, cudaMemcpyDeviceToHost, st));
            if (S.mplx_len) CK (cudaMemcpyAsync (S.mplx, L->h[v].mplx, S.mplx_len, cudaMemcpyDeviceToHost, st));
            if (S.divr_len) CK (cudaMemcpyAsync (S.divr, L->h[v].divr, S.divr_len, cudaMemcpyDeviceToHost, st));
        }
    }
    CK (cudaStreamSynchronize (st));
    return GZB_OK;
}

// ================================================================================================ C-ABI: DOMQ (PIZ)
extern "C" int gzb_domq_reconstruct (gzb_engine *e, gzb_domq_piz_vb *vbs, uint32_t n_vbs, uint32_t flags)
{
    if (!e || !vbs) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS, indev = devptr || (flags & GZB_IN_DEVICE);
    cudaStream_t st = e->stream;
    std::vector<uint32_t> nl (n_vbs), bvb, bfirst;
    std::vector<uint64_t> total (n_vbs, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        nl[v] = vbs[v].n_lines;
        if (vbs[v].num_norm_qs == 0 || vbs[v].denorm_len > NQ * NQ) { e->err = "bad DOMQ denorm table"; return GZB_E_BADARG; }
        total[v] = vbs[v].out_cap;
    }
    line_blocks (nl, bvb, bfirst);
    std::vector<DqPiz> h (n_vbs);
    Carver c { nullptr, 0 };
    DqPiz *d_vbs = nullptr; uint32_t *d_bvb = nullptr, *d_bfirst = nullptr, *d_info = nullptr;
    // pinned staging (see dq_stage): descriptors | line blocks | per-VBlock results
    const size_t nb4 = bvb.size () * 4, nb4a = (nb4 + 255) & ~(size_t)255;
    const size_t pin_vbs_bytes = ((size_t)n_vbs * sizeof (DqPiz) + 255) & ~(size_t)255, pin_blocks = pin_vbs_bytes, pin_info = pin_blocks + 2 * nb4a;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<DqPiz> (n_vbs); d_bvb = c.take<uint32_t> (bvb.size () + 1); d_bfirst = c.take<uint32_t> (bfirst.size () + 1);
        d_info = c.take<uint32_t> ((size_t)n_vbs * 8);
        for (uint32_t v = 0; v < n_vbs; v++) {
            DqPiz &D = h[v]; const gzb_domq_piz_vb &S = vbs[v];
            D.qual_len = S.qual_len; D.runs_len = S.runs_len; D.mplx_len = S.mplx_len; D.divr_len = S.divr_len; D.n_lines = S.n_lines;
            D.no_dom = S.num_norm_qs; D.denorm_len = S.denorm_len;
            memcpy (D.denorm, S.denorm, S.denorm_len);
            D.qual = indev ? (const uint8_t *)S.qual : c.take<uint8_t> (S.qual_len + 16);
            D.runs = indev ? (const uint8_t *)S.runs : c.take<uint8_t> (S.runs_len + 16);
            D.mplx = indev ? (const uint8_t *)S.mplx : c.take<uint8_t> (S.mplx_len + 16);
            D.divr = indev ? (const uint8_t *)S.divr : c.take<uint8_t> (S.divr_len + 16);
            D.line_len = devptr ? S.line_len : c.take<uint32_t> (S.n_lines + 1);
            D.out  = devptr ? (uint8_t *)S.out : c.take<uint8_t> (total[v] + 16);
            D.E    = c.take<uint8_t> (total[v] + 16);
            D.cum  = c.take<uint32_t> (S.runs_len + 1);
            D.nd_off = c.take<uint32_t> (S.n_lines + 1); D.dv_off = c.take<uint32_t> (S.n_lines + 1); D.out_off = c.take<uint32_t> (S.n_lines + 1);
            D.dom_i = c.take<uint8_t> (S.n_lines + 1);
            D.info = d_info ? d_info + (size_t)v * 8 : nullptr;
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, pin_info + (size_t)n_vbs * 32 + 512); if (rc) return rc; c.base = e->ws; }
    }
    for (uint32_t v = 0; v < n_vbs; v++) {
        DqPiz &D = h[v]; const gzb_domq_piz_vb &S = vbs[v];
        if (!indev) {
            if (S.qual_len) CK (cudaMemcpyAsync ((void *)D.qual, S.qual, S.qual_len, cudaMemcpyHostToDevice, st));
            if (S.runs_len) CK (cudaMemcpyAsync ((void *)D.runs, S.runs, S.runs_len, cudaMemcpyHostToDevice, st));
            if (S.mplx_len) CK (cudaMemcpyAsync ((void *)D.mplx, S.mplx, S.mplx_len, cudaMemcpyHostToDevice, st));
            if (S.divr_len) CK (cudaMemcpyAsync ((void *)D.divr, S.divr, S.divr_len, cudaMemcpyHostToDevice, st));
        }
        if (!devptr && S.n_lines) CK (cudaMemcpyAsync ((void *)D.line_len, S.line_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
        CK (cudaMemsetAsync (D.E, 0, total[v] + 16, st));
    }
    CK (cudaMemsetAsync (d_info, 0, (size_t)n_vbs * 32, st));
    memcpy (e->pin, h.data (), n_vbs * sizeof (DqPiz));
    CK (cudaMemcpyAsync (d_vbs, e->pin, n_vbs * sizeof (DqPiz), cudaMemcpyHostToDevice, st));
    if (nb4) {
        memcpy (e->pin + pin_blocks, bvb.data (), nb4); memcpy (e->pin + pin_blocks + nb4a, bfirst.data (), nb4);
        CK (cudaMemcpyAsync (d_bvb, e->pin + pin_blocks, nb4, cudaMemcpyHostToDevice, st));
        CK (cudaMemcpyAsync (d_bfirst, e->pin + pin_blocks + nb4a, nb4, cudaMemcpyHostToDevice, st));
    }
    k_dqp_lines<<<n_vbs, 512, 0, st>>>(d_vbs);
    k_dqp_runs<<<n_vbs, 512, 0, st>>>(d_vbs);
    k_dqp_literals<<<n_vbs, 512, 0, st>>>(d_vbs);
    if (!bvb.empty ()) k_dqp_denorm<<<(uint32_t)bvb.size (), 256, 0, st>>>(d_vbs, d_bvb, d_bfirst);
    e->launches += 4;
    uint32_t *const info = reinterpret_cast<uint32_t *>(e->pin + pin_info);
    CK (cudaMemcpyAsync (info, d_info, (size_t)n_vbs * 32, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    for (uint32_t v = 0; v < n_vbs; v++) {
        const uint32_t *I = info + (size_t)v * 8;
        if (I[2] || (uint64_t)I[0] + I[1] > vbs[v].out_cap) { e->err = "malformed DOMQ streams"; return GZB_E_CORRUPT; }
        if (!devptr && (I[0] + I[1])) CK (cudaMemcpyAsync (vbs[v].out, h[v].out, (size_t)I[0] + I[1], cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    return GZB_OK;
}
