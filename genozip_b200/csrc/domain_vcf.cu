// domain_vcf.cu — PBWT genotype-matrix transform on sm_100a.
//
// Reference functions replaced (relative to /root/reference/src/codec_pbwt.c):
//   codec_pbwt_compress (:244-287) = per row: permutation update (codec_pbwt_calculate_permutation :110-154), boustrophedon
//   traversal (:265) and run-length encoding into RUNS + FGRC (codec_pbwt_run_len_encode :213-238, _udpate_fgrc :181-210);
//   codec_pbwt_uncompress (:372-402) + pbwt_decode_one_line (:317-369).
//
// Rows are serial (row r's permutation is the stable partition of row r-1's by row r-1's alleles); columns are parallel and
// VBlocks are independent.  A batch call runs ONE CTA PER VBLOCK's matrix (k_pbwt_rows): the permutation (16-bit indices) and
// the permuted row live in shared memory, the rows stream global → shared through a ring of 1-D bulk copies (TMA) that runs
// PB_RING rows ahead of the CTA, and a row costs three block barriers: the gather, ONE packed 64-bit block scan that carries the
// counts of up to three alleles and the run-boundary count together, and the scatter of the next permutation.
//
//   encode: k_pbwt_rows<0> writes one record (position in traversal order, allele) per run boundary; k_pbwt_emit turns the records
//           into RUNS and FGRC with three prefix sums — the reference's state machine (:181-238) is a run-length encoding of the
//           foreground runs plus a zero-length background run between two adjacent foreground runs, so nothing in it is serial.
//   decode: k_pbwt_dec_prefix / k_pbwt_dec_alleles / k_pbwt_expand turn RUNS + FGRC into the alleles in traversal order (fully
//           parallel), k_pbwt_rows<1> walks the rows and un-permutes them.
//
// Matrices whose row does not fit in shared memory (ht_per_line > ~14 000) or that are longer than 4 GB take the round-1 kernels
// (k_pbwt_encode_wide / k_pbwt_decode_wide: one CTA, permutation in global memory), one VBlock at a time.
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "gzb_tma.cuh"
#include "engine.h"

using namespace gzb;

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr int      PBT        = 256;               // threads of the row kernel
constexpr int      PB_RING    = 8;                 // rows in flight global → shared
constexpr size_t   PB_SMEM_MAX = 200 * 1024;       // dynamic shared memory the row kernel may ask for
constexpr int      PB_THREADS = 1024;              // the wide kernels and the scans
constexpr uint32_t PB_TILE    = 4096;              // positions per CTA of k_pbwt_expand

enum : uint32_t { PBE_CAP = 1, PBE_FGCOUNT = 2, PBE_COVER = 3 };

struct PbVb {
    const uint8_t *src;          // rows kernel input: the matrix (encode) or the alleles in traversal order (decode)
    uint8_t  *dst;               // decode: the matrix
    uint64_t  len;               // n_lines * w
    uint32_t  n_lines, w;
    uint32_t *runs, *fgrc;       // encode: outputs; decode: inputs
    uint32_t  runs_cap, fgrc_cap;
    uint32_t  n_runs, n_fgrc;    // decode (n_fgrc without the two length words)
    uint32_t *bpos; uint8_t *bal; uint32_t bcap;      // encode: run-boundary records
    uint32_t *fgs;               // encode: foreground ordinal of each FGRC entry's first run
    uint64_t *cumpos; uint32_t *fgcum; uint8_t *ral;  // decode: inclusive prefix of RUNS / of FGRC counts, allele of run k
    uint32_t *gperm;             // wide path: 4*w words of global scratch
    uint32_t *result;            // [0] n_runs [1] n_fgrc [2] error [3] boundary records
    uint32_t  wide;              // 1 = not for the row kernel
    uint32_t  pad;
};

// order in which alleles are grouped (:122-130): '0' .. 255, 0 .. 36 (uint8 wrap-around) are 0..244, then . * % - & ; any other
// byte (never written by the segmenter; the reference drops such columns from the permutation) is grouped last
__device__ __forceinline__ uint32_t pb_ord (uint32_t a)
{
    const uint32_t t = (a - 48u) & 0xffu;
    if (t < 245u) return t;
    return 245u + (uint32_t)((0x50355155542ull >> (4 * (t - 245u))) & 15u);       // bytes 37..47: % & ' ( ) * + , - . /
}

__device__ __forceinline__ uint64_t pb_scan64 (uint64_t v, uint64_t (*ws)[PBT / 32], uint32_t &flip, uint64_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
    uint64_t *s = ws[flip]; flip ^= 1;           // (double-buffered: one barrier per scan)
    if (lane == 31) s[warp] = inc;
    __syncthreads ();
    uint64_t before = 0, tot = 0;
    for (int k = 0; k < PBT / 32; k++) { const uint64_t x = s[k]; tot += x; if (k < warp) before += x; }
    total = tot;
    return before + inc - v;
}

constexpr uint32_t NOKEY = 0xffffu;

// SIMD helpers on the eight alleles a thread holds in one 64-bit word
__device__ __forceinline__ uint32_t pb_cnt_eq (uint32_t lo, uint32_t hi, uint32_t key4, uint32_t vlo, uint32_t vhi)   // bytes of (lo, hi) equal to key among the valid ones
{
    return __popc (__vcmpeq4 (lo, key4) & vlo) + __popc (__vcmpeq4 (hi, key4) & vhi);
}

// One CTA per VBlock.  A thread owns a contiguous chunk of C8 (a multiple of 8) columns; it handles them eight at a time as
// one 64-bit word of alleles and one 128-bit word of permutation indices.  A row whose alleles are all among '0' '1' '2' (the
// common case) takes the SIMD path: three barriers — the gather (which also tells, through its OR, whether any other allele
// turned up), one packed scan, the scatter.  Any other row takes the general passes over the set of alleles present.
template <int DEC>
__global__ void __launch_bounds__(PBT) k_pbwt_rows (const PbVb *vbs)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint64_t ws[2][PBT / 32];
    __shared__ uint32_t bm[8];
    __shared__ uint64_t bar[PB_RING];
    const PbVb &V = vbs[blockIdx.x];
    const uint32_t w = V.w, n_lines = V.n_lines;
    if (V.wide || !w || !n_lines) return;
    const uint8_t *src = V.src; const uint64_t len = V.len;
    const uint32_t slot = tma_slot_bytes (w), wp = (w + 7) & ~7u;
    uint16_t *P   = reinterpret_cast<uint16_t *>(smem);                      // two permutations of w 16-bit indices (stride wp)
    uint8_t  *al  = smem + 4 * (size_t)wp;                                   // the permuted row
    uint8_t  *ring = al + ((wp + 16 + 15) & ~15u);
    uint8_t  *orow = ring + (size_t)PB_RING * slot;                          // decode: the un-permuted row, staged for aligned stores
    const int tid = threadIdx.x;
    const uint32_t C8 = ((w + PBT * 8 - 1) / (PBT * 8)) * 8, lo = min (w, tid * C8), hi = min (w, lo + C8);

    if (tid < 8) bm[tid] = 0;
    if (tid == 0) { for (int s = 0; s < PB_RING; s++) mbar_init (&bar[s], 1); mbar_fence_init (); }
    for (uint32_t i = tid; i < wp; i += PBT) P[i] = (uint16_t)i;             // first line: identity (:146-148)
    __syncthreads ();
    if (tid == 0)
        for (uint32_t r = 0; r < PB_RING && r < n_lines; r++) {
            const uint8_t *g = src + (uint64_t)r * w;
            if (tma_superset_ok (g, w, src, len)) tma_row_issue (ring + (size_t)r * slot, g, w, &bar[r]); else mbar_arrive (&bar[r]);
        }
    GZB_TMA_EMU_ISSUED ();

    uint32_t flip = 0, nbnd = 0, err = 0;
    uint8_t carry = 0;                                                       // allele of the run that is open when a row starts
    for (uint32_t r = 0; r < n_lines; r++) {
        const uint32_t s = r % PB_RING, cur = r & 1;
        const uint16_t *Pc = P + (size_t)cur * wp; uint16_t *Pn = P + (size_t)(cur ^ 1) * wp;
        const uint8_t *g = src + (uint64_t)r * w;
        const bool staged = tma_superset_ok (g, w, src, len);
        mbar_wait (&bar[s], (r / PB_RING) & 1);
        const uint8_t *row = staged ? ring + (size_t)s * slot + ((uintptr_t)g & 15) : g;
        const bool back = r & 1, part = r + 1 < n_lines;
        uint8_t *so = nullptr;
        if (DEC) so = orow + ((uintptr_t)(V.dst + (uint64_t)r * w) & 15);

        // ---- the row in permuted order (:151-153 / :340-343), eight columns at a time
        uint32_t other = 0;
        for (uint32_t i0 = lo; i0 < hi; i0 += 8) {
            const uint32_t nv = min (8u, hi - i0);
            const uint4 pv = *reinterpret_cast<const uint4 *>(Pc + i0);
            const uint32_t pw[4] = { pv.x, pv.y, pv.z, pv.w };
            uint32_t alo = 0x30303030u, ahi = 0x30303030u;                   // (columns past the end read as '0' and are masked where it matters)
            #pragma unroll
            for (uint32_t j = 0; j < 8; j++)
                if (j < nv) {
                    const uint32_t idx = (pw[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                    uint32_t a;
                    if (DEC) { a = row[back ? w - 1 - (i0 + j) : i0 + j]; so[idx] = (uint8_t)a; }
                    else a = row[idx];
                    if (j < 4) alo = (alo & ~(0xffu << (8 * j))) | (a << (8 * j)); else ahi = (ahi & ~(0xffu << (8 * (j - 4)))) | (a << (8 * (j - 4)));
                }
            *reinterpret_cast<uint2 *>(al + i0) = make_uint2 (alo, ahi);
            other |= __vcmpgtu4 (__vsub4 (alo, 0x30303030u), 0x02020202u) | __vcmpgtu4 (__vsub4 (ahi, 0x30303030u), 0x02020202u);
        }
        const int general = __syncthreads_or (other != 0);                   // B1: al (and orow) complete; the ring slot is free
        if (tid == 0 && r + PB_RING < n_lines) {
            const uint8_t *g2 = src + (uint64_t)(r + PB_RING) * w;
            if (tma_superset_ok (g2, w, src, len)) tma_row_issue (ring + (size_t)s * slot, g2, w, &bar[s]); else mbar_arrive (&bar[s]);
        }

        if (DEC) {                                                           // the finished row: head bytes, 16-byte body, tail bytes
            uint8_t *gd = V.dst + (uint64_t)r * w;
            const uint32_t h = min (w, (16u - (uint32_t)((uintptr_t)gd & 15)) & 15u), n16 = (w - h) >> 4, t0 = h + (n16 << 4);
            if ((uint32_t)tid < h) gd[tid] = so[tid];
            for (uint32_t c = tid; c < n16; c += PBT) reinterpret_cast<uint4 *>(gd + h)[c] = reinterpret_cast<const uint4 *>(so + h)[c];
            if ((uint32_t)tid < w - t0) gd[t0 + tid] = so[t0 + tid];
        }

        if (!general) {
            // ---- SIMD path: alleles '0' '1' '2' only.  Counts of the three and of the run boundaries of the chunk, one scan,
            //      then the stable partition into the next permutation (:132-141) and the boundary records (:213-238)
            uint32_t c0 = 0, c1 = 0, c2 = 0, cb = 0;
            if (part || !DEC)
                for (uint32_t i0 = lo; i0 < hi; i0 += 8) {
                    const uint32_t nv = min (8u, hi - i0);
                    const uint2 a = *reinterpret_cast<const uint2 *>(al + i0);
                    const uint32_t vlo = nv >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - nv))), vhi = nv > 4 ? (0x01010101u >> (8 * (8 - nv))) : 0u;
                    const uint32_t tl = __vsub4 (a.x, 0x30303030u), th = __vsub4 (a.y, 0x30303030u);
                    c0 += pb_cnt_eq (tl, th, 0u, vlo, vhi); c1 += pb_cnt_eq (tl, th, 0x01010101u, vlo, vhi); c2 += pb_cnt_eq (tl, th, 0x02020202u, vlo, vhi);
                    if (!DEC) {
                        // neighbour in traversal order: the column before (forward rows) or after (backward rows); across the chunk
                        // edge it comes from shared memory, at the row's first column in traversal order from the open run
                        uint32_t nlo, nhi;
                        if (!back) {
                            const uint32_t pb = i0 ? al[i0 - 1] : carry;
                            nlo = (a.x << 8) | pb; nhi = (a.y << 8) | (a.x >> 24);
                        }
                        else {
                            const uint32_t sb = i0 + nv < w ? al[i0 + nv] : carry;        // successor of the last valid column
                            uint32_t xl = a.x, xh = a.y;
                            if (nv < 8) { if (nv < 4) xl = (xl & ~(0xffu << (8 * nv))) | (sb << (8 * nv)); else if (nv > 4) xh = (xh & ~(0xffu << (8 * (nv - 4)))) | (sb << (8 * (nv - 4))); else xh = (xh & ~0xffu) | sb; }
                            nlo = (xl >> 8) | (xh << 24); nhi = (xh >> 8) | (nv == 8 ? sb << 24 : 0u);
                        }
                        cb += __popc (__vcmpne4 (a.x, nlo) & vlo) + __popc (__vcmpne4 (a.y, nhi) & vhi);
                    }
                }
            if (part || !DEC) {
                uint64_t tot;
                const uint64_t ex = pb_scan64 ((uint64_t)c0 | ((uint64_t)c1 << 16) | ((uint64_t)c2 << 32) | ((uint64_t)cb << 48), ws, flip, tot);
                const uint32_t t0 = (uint32_t)tot & 0xffff, t1 = (uint32_t)(tot >> 16) & 0xffff;
                uint32_t p0 = (uint32_t)ex & 0xffff, p1 = t0 + ((uint32_t)(ex >> 16) & 0xffff), p2 = t0 + t1 + ((uint32_t)(ex >> 32) & 0xffff);
                if (part)
                    for (uint32_t i0 = lo; i0 < hi; i0 += 8) {
                        const uint32_t nv = min (8u, hi - i0);
                        const uint2 a = *reinterpret_cast<const uint2 *>(al + i0);
                        const uint4 pv = *reinterpret_cast<const uint4 *>(Pc + i0);
                        const uint32_t pw[4] = { pv.x, pv.y, pv.z, pv.w };
                        #pragma unroll
                        for (uint32_t j = 0; j < 8; j++)
                            if (j < nv) {
                                const uint32_t t = (((j < 4 ? a.x : a.y) >> (8 * (j & 3))) - 0x30u) & 0xffu;
                                const uint16_t idx = (uint16_t)(pw[j >> 1] >> (16 * (j & 1)));
                                const uint32_t d = t == 0 ? p0++ : t == 1 ? p1++ : p2++;
                                Pn[d] = idx;
                            }
                    }
                if (!DEC) {
                    const uint32_t tb = (uint32_t)(tot >> 48), eb = (uint32_t)(ex >> 48);
                    if (cb) {                                                // (few: a thread's boundaries are written one by one, in traversal order)
                        uint32_t bi = nbnd + (back ? tb - eb - cb : eb);
                        for (uint32_t k = 0; k < hi - lo; k++) {
                            const uint32_t i = back ? hi - 1 - k : lo + k;
                            const uint8_t c = al[i], nb = back ? (i + 1 < w ? al[i + 1] : carry) : (i ? al[i - 1] : carry);
                            if (c != nb) {
                                if (bi < V.bcap) { V.bpos[bi] = (uint32_t)((uint64_t)r * w + (back ? w - 1 - i : i)); V.bal[bi] = c; } else err = PBE_CAP;
                                bi++;
                            }
                        }
                    }
                    nbnd += tb;
                }
            }
        }
        else {
            // ---- general path: the set of alleles of the row, then passes of up to three alleles each; the first pass also counts
            //      and records the run boundaries
            if (tid < 8) bm[tid] = 0;
            __syncthreads ();
            uint32_t m0 = 0;
            for (uint32_t i = lo; i < hi; i++) { const uint32_t o = pb_ord (al[i]); if (o < 32) m0 |= 1u << o; else atomicOr (&bm[o >> 5], 1u << (o & 31)); }
            m0 = __reduce_or_sync (0xffffffffu, m0);
            if ((tid & 31) == 0 && m0) atomicOr (&bm[0], m0);
            __syncthreads ();
            uint32_t base = 0, k0 = NOKEY, k1 = NOKEY, k2 = NOKEY, nk = 0;
            bool first = true;
            for (uint32_t wd = 0; wd <= 8; wd++) {
                uint32_t bits = (wd < 8 && part) ? bm[wd] : 0;
                while (bits || (wd == 8 && (nk || (first && !DEC)))) {
                    if (bits) {
                        const uint32_t key = wd * 32 + (__ffs ((int)bits) - 1);
                        bits &= bits - 1;
                        if (nk == 0) k0 = key; else if (nk == 1) k1 = key; else k2 = key;
                        if (++nk < 3) continue;
                    }
                    uint32_t c0 = 0, c1 = 0, c2 = 0, cb = 0;
                    for (uint32_t i = lo; i < hi; i++) { const uint32_t o = pb_ord (al[i]); c0 += o == k0; c1 += o == k1; c2 += o == k2; }
                    if (first && !DEC)
                        for (uint32_t j = lo; j < hi; j++) {
                            const uint8_t c = al[back ? w - 1 - j : j], p = j ? al[back ? w - j : j - 1] : carry;
                            cb += c != p;
                        }
                    uint64_t tot;
                    const uint64_t ex = pb_scan64 ((uint64_t)c0 | ((uint64_t)c1 << 16) | ((uint64_t)c2 << 32) | ((uint64_t)cb << 48), ws, flip, tot);
                    const uint32_t t0 = (uint32_t)tot & 0xffff, t1 = (uint32_t)(tot >> 16) & 0xffff, t2 = (uint32_t)(tot >> 32) & 0xffff;
                    uint32_t p0 = base + ((uint32_t)ex & 0xffff), p1 = base + t0 + ((uint32_t)(ex >> 16) & 0xffff), p2 = base + t0 + t1 + ((uint32_t)(ex >> 32) & 0xffff);
                    if (part)
                        for (uint32_t i = lo; i < hi; i++) {
                            const uint32_t o = pb_ord (al[i]);
                            if (o == k0) Pn[p0++] = Pc[i]; else if (o == k1) Pn[p1++] = Pc[i]; else if (o == k2) Pn[p2++] = Pc[i];
                        }
                    if (first && !DEC) {
                        uint32_t bi = nbnd + (uint32_t)(ex >> 48);
                        for (uint32_t j = lo; j < hi; j++) {
                            const uint8_t c = al[back ? w - 1 - j : j], p = j ? al[back ? w - j : j - 1] : carry;
                            if (c != p) {
                                if (bi < V.bcap) { V.bpos[bi] = (uint32_t)((uint64_t)r * w + j); V.bal[bi] = c; } else err = PBE_CAP;
                                bi++;
                            }
                        }
                        nbnd += (uint32_t)(tot >> 48);
                    }
                    base += t0 + t1 + t2;
                    first = false; nk = 0; k0 = k1 = k2 = NOKEY;
                }
            }
        }
        carry = al[back ? 0 : w - 1];
        __syncthreads ();                                                    // B4: the next permutation is complete, al may be overwritten
    }
    if (!DEC) {
        if (err) atomicMax (&V.result[2], err);
        if (tid == 0) V.result[3] = nbnd;
    }
}

// two-level exclusive scan of 64-bit values over a 1024-thread block
__device__ uint64_t pb_block_scan64 (uint64_t v, uint64_t *sm /*[33]*/, uint64_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sm[warp] = inc;
    __syncthreads ();
    if (warp == 0) {
        const uint64_t x = sm[lane]; uint64_t xi = x;
        for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync (0xffffffffu, xi, o); if (lane >= o) xi += t; }
        sm[lane] = xi - x;
        if (lane == 31) sm[32] = xi;
    }
    __syncthreads ();
    const uint64_t r = sm[warp] + inc - v;
    *total = sm[32];
    __syncthreads ();
    return r;
}

// boundary records → RUNS + FGRC.  Record k opens a run of allele y = bal[k] after a run of x = bal[k-1] (x = 0 before the first):
//   x == '0'            one RUNS entry; FGRC: one more run of y, or a new entry            (codec_pbwt_udpate_fgrc :184-198)
//   x != '0', y != '0'  a zero-length background entry, then the run; FGRC: a new entry   (:201-204)
//   y == '0'            one RUNS entry                                                     (:207-209)
// Two foreground runs in a row always differ, so FGRC is the run-length encoding of the foreground runs' alleles.
__global__ void __launch_bounds__(PB_THREADS) k_pbwt_emit (const PbVb *vbs)
{
    const PbVb &V = vbs[blockIdx.x];
    if (V.wide || !V.w || !V.n_lines) return;
    __shared__ uint64_t sm[33];
    __shared__ uint32_t s_err;
    const int tid = threadIdx.x;
    if (tid == 0) s_err = V.result[2];
    __syncthreads ();
    if (s_err) return;
    const uint32_t NB = V.result[3];
    uint32_t nr = 0, nfg = 0, ne = 0, err = 0;
    for (uint32_t b0 = 0; b0 < NB; b0 += PB_THREADS) {
        const uint32_t k = b0 + tid;
        uint32_t en = 0, fg = 0, hd = 0; uint8_t y = 0;
        if (k < NB) {
            y = V.bal[k];
            const uint8_t x = k ? V.bal[k - 1] : 0;
            en = (x != '0' && y != '0') ? 2 : 1;
            fg = y != '0';
            hd = fg && (x != '0' || k < 2 || V.bal[k - 2] != y);
        }
        uint64_t tot;
        const uint64_t ex = pb_block_scan64 ((uint64_t)en | ((uint64_t)fg << 21) | ((uint64_t)hd << 42), sm, &tot);
        if (k < NB) {
            const uint32_t ri = nr + ((uint32_t)ex & 0x1fffff), m = nfg + ((uint32_t)(ex >> 21) & 0x1fffff), ei = ne + (uint32_t)(ex >> 42);
            const uint32_t run = (k + 1 < NB ? V.bpos[k + 1] : (uint32_t)V.len) - V.bpos[k];
            if (ri + en > V.runs_cap) err = PBE_CAP;
            else if (en == 2) { V.runs[ri] = 0; V.runs[ri + 1] = run; }
            else V.runs[ri] = run;
            if (hd) { if (ei + 3 > V.fgrc_cap) err = PBE_CAP; else { V.fgs[ei] = m; V.fgrc[ei] = y; } }
        }
        nr += (uint32_t)tot & 0x1fffff; nfg += (uint32_t)(tot >> 21) & 0x1fffff; ne += (uint32_t)(tot >> 42);
    }
    if (ne + 2 > V.fgrc_cap) err = PBE_CAP;
    if (err) atomicMax (&s_err, err);
    __syncthreads ();
    if (s_err) { if (tid == 0) V.result[2] = s_err; return; }
    for (uint32_t i = tid; i < ne; i += PB_THREADS) {
        const uint32_t cnt = (i + 1 < ne ? V.fgs[i + 1] : nfg) - V.fgs[i];
        if (cnt > 0xffffffu) atomicMax (&V.result[2], (uint32_t)PBE_FGCOUNT);  // the reference asserts when the 24-bit count wraps (:193)
        V.fgrc[i] = (V.fgrc[i] & 0xff) | (cnt << 8);
    }
    if (tid == 0) {
        V.fgrc[ne] = (uint32_t)(V.len & 0xffffffffu); V.fgrc[ne + 1] = (uint32_t)(V.len >> 32);          // :274-276
        V.result[0] = nr; V.result[1] = ne + 2;
    }
}

// ---------------------------------------------------------------------------------------------- decode
__global__ void __launch_bounds__(PB_THREADS) k_pbwt_dec_prefix (const PbVb *vbs)   // one CTA per VBlock: prefix sums of run lengths and FGRC counts
{
    const PbVb &V = vbs[blockIdx.x];
    __shared__ uint64_t sm[33];
    uint64_t acc = 0;
    for (uint32_t base = 0; base < V.n_runs; base += PB_THREADS) {
        const uint32_t i = base + threadIdx.x; const uint64_t v = i < V.n_runs ? V.runs[i] : 0; uint64_t tot;
        const uint64_t ex = pb_block_scan64 (v, sm, &tot);
        if (i < V.n_runs) V.cumpos[i] = acc + ex + v;
        acc += tot;
    }
    if (threadIdx.x == 0 && acc < V.len) V.result[2] = PBE_COVER;           // the runs must cover the matrix (:363-365)
    uint64_t a2 = 0;
    for (uint32_t base = 0; base < V.n_fgrc; base += PB_THREADS) {
        const uint32_t i = base + threadIdx.x; const uint64_t v = i < V.n_fgrc ? (V.fgrc[i] >> 8) : 0; uint64_t tot;
        const uint64_t ex = pb_block_scan64 (v, sm, &tot);
        if (i < V.n_fgrc) V.fgcum[i] = (uint32_t)min ((uint64_t)0xffffffffu, a2 + ex + v);
        a2 += tot;
    }
}

__global__ void k_pbwt_dec_alleles (const PbVb *vbs)  // RUNS alternate background('0') / foreground, starting with background (:326)
{
    const PbVb &V = vbs[blockIdx.y];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < V.n_runs; k += gridDim.x * blockDim.x) {
        if (!(k & 1)) { V.ral[k] = '0'; continue; }
        const uint32_t j = k >> 1;                                          // ordinal of this foreground run
        uint32_t lo = 0, hi = V.n_fgrc;                                     // first group with fgcum > j (:353-358)
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (V.fgcum[mid] > j) hi = mid; else lo = mid + 1; }
        V.ral[k] = lo < V.n_fgrc ? (uint8_t)(V.fgrc[lo] & 0xff) : 0;
    }
}

__device__ __forceinline__ uint32_t pb_first_after (const uint64_t *cum, uint32_t lo, uint32_t hi, uint64_t p)   // first k in [lo,hi) with cum[k] > p, else hi
{
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cum[mid] > p) hi = mid; else lo = mid + 1; }
    return lo;
}

// alleles in traversal order: position p belongs to the first run whose inclusive prefix exceeds p (zero-length runs never match)
__global__ void __launch_bounds__(256) k_pbwt_expand (const PbVb *vbs, uint8_t *const *pal)
{
    const PbVb &V = vbs[blockIdx.y];
    if (V.wide) return;
    __shared__ uint32_t s_lo, s_hi;
    uint8_t *out = pal[blockIdx.y];
    for (uint64_t t0 = (uint64_t)blockIdx.x * PB_TILE; t0 < V.len; t0 += (uint64_t)gridDim.x * PB_TILE) {
        const uint64_t t1 = min (V.len, t0 + PB_TILE);
        if (threadIdx.x == 0) s_lo = pb_first_after (V.cumpos, 0, V.n_runs, t0);
        if (threadIdx.x == 32) s_hi = pb_first_after (V.cumpos, 0, V.n_runs, t1 - 1);
        __syncthreads ();
        const uint64_t p0 = t0 + (uint64_t)threadIdx.x * 16;
        if (p0 < t1) {
            uint32_t k = pb_first_after (V.cumpos, s_lo, min (V.n_runs, s_hi + 1), p0);
            uint32_t o[4] = { 0, 0, 0, 0 };
            uint64_t end = k < V.n_runs ? V.cumpos[k] : ~0ull; uint32_t a = k < V.n_runs ? V.ral[k] : 0;
            for (uint32_t b = 0; b < 16 && p0 + b < t1; b++) {
                while (p0 + b >= end) { k++; end = k < V.n_runs ? V.cumpos[k] : ~0ull; a = k < V.n_runs ? V.ral[k] : 0; }
                o[b >> 2] |= a << (8 * (b & 3));
            }
            *reinterpret_cast<uint4 *>(out + p0) = make_uint4 (o[0], o[1], o[2], o[3]);     // (the buffer has 16 bytes of slack)
        }
        __syncthreads ();
    }
}

// ---------------------------------------------------------------------------------------------- wide matrices (round-1 kernels)
__device__ uint32_t pb_block_excl_sum (uint32_t v, uint32_t *sm, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sm[warp] = inc;
    __syncthreads ();
    if (warp == 0) {
        uint32_t w = sm[lane], wi = w;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, wi, o); if (lane >= o) wi += t; }
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads ();
    uint32_t r = sm[warp] + inc - v;
    *total = sm[32];
    __syncthreads ();
    return r;
}

__device__ void pb_key_order (const uint8_t *has, uint8_t *order, uint32_t *n_order)
{
    uint32_t no = 0;
    for (uint32_t a = '0'; a != (uint32_t)(('0' + 245) & 0xff); a = (a + 1) & 0xff) if (has[a]) order[no++] = (uint8_t)a;
    const uint8_t pseudo[5] = { '.', '*', '%', '-', '&' };
    for (int i = 0; i < 5; i++) if (has[pseudo[i]]) order[no++] = pseudo[i];
    *n_order = no;
}

__device__ void pb_partition (const uint32_t *perm, uint32_t *tmp, const uint8_t *al, uint32_t w,
                              uint8_t *has, uint8_t *order, uint32_t *n_order, uint32_t *sm)
{
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += PB_THREADS) has[i] = 0;
    __syncthreads ();
    for (uint32_t i = tid; i < w; i += PB_THREADS) has[al[i]] = 1;
    __syncthreads ();
    if (tid == 0) pb_key_order (has, order, n_order);
    __syncthreads ();
    const uint32_t C = (w + PB_THREADS - 1) / PB_THREADS, lo = min (w, tid * C), hi = min (w, lo + C);
    uint32_t base = 0;
    for (uint32_t kx = 0; kx < *n_order; kx++) {
        const uint8_t key = order[kx];
        uint32_t cnt = 0;
        for (uint32_t i = lo; i < hi; i++) cnt += al[i] == key;
        uint32_t tot;
        uint32_t pos = base + pb_block_excl_sum (cnt, sm, &tot);
        for (uint32_t i = lo; i < hi; i++) if (al[i] == key) tmp[pos++] = perm[i];
        base += tot;
    }
    __syncthreads ();
}

__global__ void __launch_bounds__(PB_THREADS) k_pbwt_encode_wide (const PbVb *vbs, uint32_t vb_i)
{
    const PbVb &P = vbs[vb_i];
    const uint32_t w = P.w;
    uint32_t *perm = P.gperm, *tmp = perm + w;
    uint8_t  *al   = reinterpret_cast<uint8_t *>(P.gperm + 2 * (size_t)w);
    uint32_t *bpos = P.gperm + 2 * (size_t)w + ((w + 3) / 4);
    __shared__ uint32_t sm[33];
    __shared__ uint8_t has[256], order[256];
    __shared__ uint32_t n_order, s_nb;
    __shared__ uint32_t st_nr, st_nf, st_err; __shared__ uint8_t st_allele;
    const int tid = threadIdx.x;
    if (tid == 0) { st_nr = 0; st_nf = 0; st_err = 0; st_allele = 0; }
    for (uint32_t i = tid; i < w; i += PB_THREADS) perm[i] = i;
    __syncthreads ();
    for (uint32_t r = 0; r < P.n_lines; r++) {
        const uint8_t *line = P.src + (size_t)r * w;
        for (uint32_t i = tid; i < w; i += PB_THREADS) al[i] = line[perm[i]];
        __syncthreads ();
        const bool back = r & 1;
        const uint8_t carry = st_allele;
        const uint32_t C = (w + PB_THREADS - 1) / PB_THREADS, lo = min (w, tid * C), hi = min (w, lo + C);
        uint32_t cnt = 0;
        for (uint32_t i = lo; i < hi; i++) {
            const uint8_t cur = al[back ? w - 1 - i : i];
            const uint8_t prev = i ? al[back ? w - i : i - 1] : carry;
            cnt += cur != prev;
        }
        uint32_t tot;
        uint32_t pos = pb_block_excl_sum (cnt, sm, &tot);
        for (uint32_t i = lo; i < hi; i++) {
            const uint8_t cur = al[back ? w - 1 - i : i];
            const uint8_t prev = i ? al[back ? w - i : i - 1] : carry;
            if (cur != prev) bpos[pos++] = i;
        }
        if (tid == 0) s_nb = tot;
        __syncthreads ();
        if (tid == 0 && !st_err) {                                          // codec_pbwt_run_len_encode (:213-238), one boundary at a time
            uint32_t nr = st_nr, nf = st_nf; uint8_t run_allele = st_allele;
            uint32_t prev_pos = 0;
            for (uint32_t b = 0; b <= s_nb; b++) {
                const uint32_t p = b < s_nb ? bpos[b] : w;
                if (p > prev_pos && nr) P.runs[nr - 1] += p - prev_pos;
                if (b == s_nb) break;
                if (nr + 2 > P.runs_cap || nf + 3 > P.fgrc_cap) { st_err = PBE_CAP; break; }
                const uint8_t done = run_allele;
                run_allele = al[back ? w - 1 - p : p];
                if (done == '0') {
                    if (nf && run_allele == (P.fgrc[nf - 1] & 0xff)) {
                        const uint32_t c = (P.fgrc[nf - 1] >> 8) + 1;
                        if (!(c & 0xffffffu)) { st_err = PBE_FGCOUNT; break; }
                        P.fgrc[nf - 1] = (P.fgrc[nf - 1] & 0xff) | (c << 8);
                    }
                    else P.fgrc[nf++] = run_allele | (1u << 8);
                }
                else if (run_allele != '0') { P.fgrc[nf++] = run_allele | (1u << 8); P.runs[nr++] = 0; }
                P.runs[nr++] = 0;
                prev_pos = p;
            }
            st_nr = nr; st_nf = nf; st_allele = run_allele;
        }
        __syncthreads ();
        if (r + 1 < P.n_lines) {
            pb_partition (perm, tmp, al, w, has, order, &n_order, sm);
            uint32_t *t = perm; perm = tmp; tmp = t;
        }
    }
    if (tid == 0) {
        uint32_t nf = st_nf;
        if (nf + 2 <= P.fgrc_cap) { P.fgrc[nf++] = (uint32_t)(P.len & 0xffffffffu); P.fgrc[nf++] = (uint32_t)(P.len >> 32); } else st_err = PBE_CAP;
        P.result[0] = st_nr; P.result[1] = nf; P.result[2] = st_err;
    }
}

__global__ void __launch_bounds__(PB_THREADS) k_pbwt_decode_wide (const PbVb *vbs, uint32_t vb_i)
{
    const PbVb &P = vbs[vb_i];
    const uint32_t w = P.w;
    uint32_t *perm = P.gperm, *tmp = perm + w;
    uint8_t  *al   = reinterpret_cast<uint8_t *>(P.gperm + 2 * (size_t)w);
    __shared__ uint32_t sm[33];
    __shared__ uint8_t has[256], order[256];
    __shared__ uint32_t n_order, s_klo;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < w; i += PB_THREADS) perm[i] = i;
    if (tid == 0) s_klo = 0;
    __syncthreads ();
    for (uint32_t r = 0; r < P.n_lines; r++) {
        uint8_t *line = P.dst + (size_t)r * w;
        const bool back = r & 1;
        const uint64_t p0 = (uint64_t)r * w;
        const uint32_t klo = s_klo;
        __syncthreads ();
        for (uint32_t i = tid; i < w; i += PB_THREADS) {
            const uint64_t p = p0 + i;
            const uint32_t k = pb_first_after (P.cumpos, klo, P.n_runs, p);
            const uint8_t a = k < P.n_runs ? P.ral[k] : 0;
            const uint32_t o = back ? w - 1 - i : i;
            al[o] = a; line[perm[o]] = a;
            if (i == w - 1) s_klo = k;
        }
        __syncthreads ();
        if (r + 1 < P.n_lines) {
            pb_partition (perm, tmp, al, w, has, order, &n_order, sm);
            uint32_t *t = perm; perm = tmp; tmp = t;
        }
    }
}

size_t pb_rows_smem (uint32_t w, bool dec)
{
    const size_t slot = tma_slot_bytes (w), wp = (w + 7) & ~7u;
    return 4 * wp + ((wp + 16 + 15) & ~(size_t)15) + (size_t)PB_RING * slot + (dec ? slot : 0) + 16;
}
bool pb_is_wide (uint32_t w, uint64_t len) { return w > 65535 || len >= (1ull << 32) || pb_rows_smem (w, true) > PB_SMEM_MAX; }

struct Carver {
    uint8_t *base; size_t off;
    template <typename T> T *take (size_t count) {
        size_t bytes = (count * sizeof (T) + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

int pb_status (gzb_engine *e, uint32_t err)
{
    if (!err) return GZB_OK;
    if (err == PBE_COVER) { e->err = "PBWT: runs do not cover the matrix"; return GZB_E_CORRUPT; }
    e->err = err == PBE_FGCOUNT ? "PBWT: more than 0xffffff consecutive foreground runs of one allele" : "PBWT: RUNS/FGRC capacity too small";
    return GZB_E_BADARG;
}

} // namespace

extern "C" int gzb_pbwt_encode_batch (gzb_engine *e, gzb_pbwt_vb *vbs, uint32_t n_vbs, uint32_t flags)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    for (uint32_t v = 0; v < n_vbs; v++)
        if (!vbs[v].ht || !vbs[v].runs || !vbs[v].fgrc || !vbs[v].ht_per_line) return GZB_E_BADARG;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    std::vector<PbVb> h (n_vbs);
    Carver c { nullptr, 0 };
    PbVb *d_vbs = nullptr; uint32_t *d_res = nullptr;
    size_t smem = 0; bool any_narrow = false;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<PbVb> (n_vbs); d_res = c.take<uint32_t> (4 * (size_t)n_vbs);
        for (uint32_t v = 0; v < n_vbs; v++) {
            PbVb &D = h[v]; const gzb_pbwt_vb &S = vbs[v];
            memset (&D, 0, sizeof D);
            D.n_lines = S.n_lines; D.w = S.ht_per_line; D.len = (uint64_t)S.n_lines * S.ht_per_line;
            D.runs_cap = S.runs_cap; D.fgrc_cap = S.fgrc_cap;
            D.wide = pb_is_wide (D.w, D.len);
            D.result = d_res ? d_res + 4 * (size_t)v : nullptr;
            if (devptr) { D.src = (const uint8_t *)S.ht; D.runs = S.runs; D.fgrc = S.fgrc; }
            else { D.src = c.take<uint8_t> (D.len + 16); D.runs = c.take<uint32_t> ((size_t)S.runs_cap + 1); D.fgrc = c.take<uint32_t> ((size_t)S.fgrc_cap + 1); }
            if (D.wide) D.gperm = c.take<uint32_t> (4 * (size_t)D.w + 64);
            else {
                D.bcap = S.runs_cap;
                D.bpos = c.take<uint32_t> ((size_t)D.bcap + 1); D.bal = c.take<uint8_t> ((size_t)D.bcap + 16); D.fgs = c.take<uint32_t> ((size_t)S.fgrc_cap + 1);
                smem = std::max (smem, pb_rows_smem (D.w, false)); any_narrow = true;
            }
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
    }
    if (!devptr)
        for (uint32_t v = 0; v < n_vbs; v++)
            if (h[v].len) CK (cudaMemcpyAsync ((void *)h[v].src, vbs[v].ht, h[v].len, cudaMemcpyHostToDevice, st));
    CK (cudaMemcpyAsync (d_vbs, h.data (), n_vbs * sizeof (PbVb), cudaMemcpyHostToDevice, st));
    CK (cudaMemsetAsync (d_res, 0, 16 * (size_t)n_vbs, st));
    if (any_narrow) {
        CK (cudaFuncSetAttribute (k_pbwt_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_SMEM_MAX));
        cudaEventRecord (e->ev0, st);
        k_pbwt_rows<0><<<n_vbs, PBT, smem, st>>>(d_vbs);
        cudaEventRecord (e->ev1, st);
        k_pbwt_emit<<<n_vbs, PB_THREADS, 0, st>>>(d_vbs);
        e->launches += 2;
    }
    for (uint32_t v = 0; v < n_vbs; v++)
        if (h[v].wide && h[v].n_lines) { k_pbwt_encode_wide<<<1, PB_THREADS, 0, st>>>(d_vbs, v); e->launches++; }
    std::vector<uint32_t> res (4 * (size_t)n_vbs);
    CK (cudaMemcpyAsync (res.data (), d_res, 16 * (size_t)n_vbs, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    CK (cudaGetLastError ());
    e->last_domain_ms = 0;
    if (any_narrow) cudaEventElapsedTime (&e->last_domain_ms, e->ev0, e->ev1);
    int rc = GZB_OK;
    for (uint32_t v = 0; v < n_vbs; v++) {
        uint32_t *r = &res[4 * (size_t)v];
        if (!h[v].n_lines) {                                                // an empty matrix: only the length words (:274-276)
            if (vbs[v].fgrc_cap < 2) r[2] = PBE_CAP;
            else { r[0] = 0; r[1] = 2; const uint32_t z[2] = { 0, 0 }; CK (cudaMemcpyAsync (h[v].fgrc, z, 8, cudaMemcpyHostToDevice, st)); }
        }
        vbs[v].status = pb_status (e, r[2]);
        if (vbs[v].status) { if (!rc) rc = vbs[v].status; vbs[v].n_runs = vbs[v].n_fgrc = 0; continue; }
        vbs[v].n_runs = r[0]; vbs[v].n_fgrc = r[1];
        if (!devptr) {
            if (r[0]) CK (cudaMemcpyAsync (vbs[v].runs, h[v].runs, (size_t)r[0] * 4, cudaMemcpyDeviceToHost, st));
            if (r[1]) CK (cudaMemcpyAsync (vbs[v].fgrc, h[v].fgrc, (size_t)r[1] * 4, cudaMemcpyDeviceToHost, st));
        }
    }
    CK (cudaStreamSynchronize (st));
    if (rc) pb_status (e, rc == GZB_E_CORRUPT ? PBE_COVER : PBE_CAP);
    return rc;
}

extern "C" int gzb_pbwt_decode_batch (gzb_engine *e, gzb_pbwt_vb *vbs, uint32_t n_vbs, uint32_t flags)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    for (uint32_t v = 0; v < n_vbs; v++)
        if (!vbs[v].ht || !vbs[v].runs || !vbs[v].fgrc || vbs[v].n_fgrc < 2 || !vbs[v].n_lines || !vbs[v].n_runs) return GZB_E_BADARG;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    // the matrix length travels in the last two FGRC words (:293-301)
    std::vector<uint64_t> lens (n_vbs);
    if (devptr) {
        int rc = engine_reserve (e, 0, 8 * (size_t)n_vbs + 64); if (rc) return rc;
        for (uint32_t v = 0; v < n_vbs; v++) CK (cudaMemcpyAsync (e->pin + 8 * (size_t)v, vbs[v].fgrc + (vbs[v].n_fgrc - 2), 8, cudaMemcpyDeviceToHost, st));
        CK (cudaStreamSynchronize (st));
        for (uint32_t v = 0; v < n_vbs; v++) { uint32_t t[2]; memcpy (t, e->pin + 8 * (size_t)v, 8); lens[v] = (uint64_t)t[0] | ((uint64_t)t[1] << 32); }
    }
    else for (uint32_t v = 0; v < n_vbs; v++) lens[v] = (uint64_t)vbs[v].fgrc[vbs[v].n_fgrc - 2] | ((uint64_t)vbs[v].fgrc[vbs[v].n_fgrc - 1] << 32);
    for (uint32_t v = 0; v < n_vbs; v++)
        if (!lens[v] || lens[v] > vbs[v].ht_cap || lens[v] % vbs[v].n_lines) { e->err = "PBWT: bad matrix length"; vbs[v].status = GZB_E_CORRUPT; return GZB_E_CORRUPT; }

    std::vector<PbVb> h (n_vbs);
    std::vector<uint8_t *> hpal (n_vbs, nullptr);
    Carver c { nullptr, 0 };
    PbVb *d_vbs = nullptr; uint32_t *d_res = nullptr; uint8_t **d_pal = nullptr;
    size_t smem = 0; bool any_narrow = false; uint32_t max_runs = 0; uint64_t max_len = 0;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<PbVb> (n_vbs); d_res = c.take<uint32_t> (4 * (size_t)n_vbs); d_pal = c.take<uint8_t *> (n_vbs);
        for (uint32_t v = 0; v < n_vbs; v++) {
            PbVb &D = h[v]; const gzb_pbwt_vb &S = vbs[v];
            memset (&D, 0, sizeof D);
            D.n_lines = S.n_lines; D.len = lens[v]; D.w = (uint32_t)(lens[v] / S.n_lines);
            D.n_runs = S.n_runs; D.n_fgrc = S.n_fgrc - 2;
            D.wide = pb_is_wide (D.w, D.len);
            D.result = d_res ? d_res + 4 * (size_t)v : nullptr;
            D.cumpos = c.take<uint64_t> ((size_t)S.n_runs + 1); D.fgcum = c.take<uint32_t> ((size_t)S.n_fgrc + 1); D.ral = c.take<uint8_t> ((size_t)S.n_runs + 16);
            if (devptr) { D.runs = S.runs; D.fgrc = S.fgrc; D.dst = (uint8_t *)S.ht; }
            else { D.runs = c.take<uint32_t> ((size_t)S.n_runs + 1); D.fgrc = c.take<uint32_t> ((size_t)S.n_fgrc + 1); D.dst = c.take<uint8_t> (D.len + 16); }
            if (D.wide) D.gperm = c.take<uint32_t> (4 * (size_t)D.w + 64);
            else {
                hpal[v] = c.take<uint8_t> (D.len + 32); D.src = hpal[v];
                smem = std::max (smem, pb_rows_smem (D.w, true)); any_narrow = true; max_len = std::max (max_len, D.len);
            }
            max_runs = std::max (max_runs, S.n_runs);
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
    }
    if (!devptr)
        for (uint32_t v = 0; v < n_vbs; v++) {
            CK (cudaMemcpyAsync (h[v].runs, vbs[v].runs, (size_t)vbs[v].n_runs * 4, cudaMemcpyHostToDevice, st));
            CK (cudaMemcpyAsync (h[v].fgrc, vbs[v].fgrc, (size_t)vbs[v].n_fgrc * 4, cudaMemcpyHostToDevice, st));
        }
    CK (cudaMemcpyAsync (d_vbs, h.data (), n_vbs * sizeof (PbVb), cudaMemcpyHostToDevice, st));
    CK (cudaMemcpyAsync (d_pal, hpal.data (), n_vbs * sizeof (uint8_t *), cudaMemcpyHostToDevice, st));
    CK (cudaMemsetAsync (d_res, 0, 16 * (size_t)n_vbs, st));
    k_pbwt_dec_prefix<<<n_vbs, PB_THREADS, 0, st>>>(d_vbs);
    k_pbwt_dec_alleles<<<dim3 (std::min<uint32_t> ((max_runs + 255) / 256, 1024), n_vbs), 256, 0, st>>>(d_vbs);
    e->launches += 2;
    if (any_narrow) {
        const uint32_t tiles = (uint32_t)std::min<uint64_t> ((max_len + PB_TILE - 1) / PB_TILE, 8192);
        k_pbwt_expand<<<dim3 (tiles, n_vbs), 256, 0, st>>>(d_vbs, d_pal);
        CK (cudaFuncSetAttribute (k_pbwt_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_SMEM_MAX));
        cudaEventRecord (e->ev0, st);
        k_pbwt_rows<1><<<n_vbs, PBT, smem, st>>>(d_vbs);
        cudaEventRecord (e->ev1, st);
        e->launches += 2;
    }
    for (uint32_t v = 0; v < n_vbs; v++)
        if (h[v].wide) { k_pbwt_decode_wide<<<1, PB_THREADS, 0, st>>>(d_vbs, v); e->launches++; }
    std::vector<uint32_t> res (4 * (size_t)n_vbs);
    CK (cudaMemcpyAsync (res.data (), d_res, 16 * (size_t)n_vbs, cudaMemcpyDeviceToHost, st));
    if (!devptr)
        for (uint32_t v = 0; v < n_vbs; v++) CK (cudaMemcpyAsync (vbs[v].ht, h[v].dst, h[v].len, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    CK (cudaGetLastError ());
    e->last_domain_ms = 0;
    if (any_narrow) cudaEventElapsedTime (&e->last_domain_ms, e->ev0, e->ev1);
    int rc = GZB_OK;
    for (uint32_t v = 0; v < n_vbs; v++) {
        vbs[v].status = pb_status (e, res[4 * (size_t)v + 2]);
        vbs[v].ht_len = vbs[v].status ? 0 : lens[v];
        if (vbs[v].status && !rc) rc = vbs[v].status;
    }
    if (rc) pb_status (e, PBE_COVER);
    return rc;
}

extern "C" int gzb_pbwt_encode (gzb_engine *e, const void *ht, uint32_t n_lines, uint32_t ht_per_line,
                                uint32_t *runs, uint32_t runs_cap, uint32_t *n_runs,
                                uint32_t *fgrc, uint32_t fgrc_cap, uint32_t *n_fgrc, uint32_t flags)
{
    if (!e || !ht || !runs || !fgrc || !n_runs || !n_fgrc || !ht_per_line) return GZB_E_BADARG;
    gzb_pbwt_vb vb; memset (&vb, 0, sizeof vb);
    vb.ht = (void *)ht; vb.n_lines = n_lines; vb.ht_per_line = ht_per_line;
    vb.runs = runs; vb.runs_cap = runs_cap; vb.fgrc = fgrc; vb.fgrc_cap = fgrc_cap;
    const int rc = gzb_pbwt_encode_batch (e, &vb, 1, flags);
    if (rc) return rc;
    *n_runs = vb.n_runs; *n_fgrc = vb.n_fgrc;
    return GZB_OK;
}

extern "C" int gzb_pbwt_decode (gzb_engine *e, const uint32_t *runs, uint32_t n_runs, const uint32_t *fgrc, uint32_t n_fgrc,
                                uint32_t n_lines, void *ht, uint64_t ht_cap, uint64_t *ht_len, uint32_t flags)
{
    if (!e || !runs || !fgrc || !ht || !ht_len || n_fgrc < 2 || !n_lines || !n_runs) return GZB_E_BADARG;
    gzb_pbwt_vb vb; memset (&vb, 0, sizeof vb);
    vb.ht = ht; vb.ht_cap = ht_cap; vb.n_lines = n_lines;
    vb.runs = (uint32_t *)runs; vb.n_runs = n_runs; vb.fgrc = (uint32_t *)fgrc; vb.n_fgrc = n_fgrc;
    const int rc = gzb_pbwt_decode_batch (e, &vb, 1, flags);
    if (rc) return rc;
    *ht_len = vb.ht_len;
    return GZB_OK;
}
