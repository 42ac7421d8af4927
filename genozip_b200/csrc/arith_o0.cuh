// arith_o0.cuh — order-0 leaves of the adaptive arithmetic coder (arith_compress_O0 / arith_uncompress_O0,
// reference arith_dynamic.c:92-152): ONE model for the whole leaf, so it can live where a warp searches it cheaply.
//
//   shared memory (per warp)   E[256]    the reference's frequency-sorted list  Freq | Symbol << 16   (c_simple_model.h:77-82)
//                              POS[256]  symbol -> position in E (encoder only)
//   registers (per lane)       s, base   sum of the lane's 8 entries E[8l .. 8l+7] and the sum of everything before them,
//                                        maintained incrementally (+STEP per symbol, one fix-up when the bubble step
//                                        crosses a lane boundary, a rebuild when the model is halved)
//
// The reference's linear search "AccFreq += F[i].Freq until it exceeds code/range" (c_simple_model.h:148-158) becomes: one
// multiply + ballot finds the lane whose 8 entries contain the symbol, then at most 8 uniform steps over that lane's
// entries (two broadcast 16-byte shared loads) — no division of code by range, no global memory, no warp scan.
// The encoder finds the position through POS.  Device only; parity is checked by the -m gpu tests.
#pragma once
#include "arith_model.cuh"

namespace gzb {

constexpr uint32_t AR0_E_WORDS = 264;                      // 256 entries + 8 padding entries (Freq 0)
constexpr uint32_t AR0_SMEM_BYTES = AR0_E_WORDS * 4 + 256; // E + POS

struct Ar0 {
    uint32_t s, base;            // per lane
    uint32_t tot; float rtot;    // uniform: TotFreq and its reciprocal rounded down
    uint32_t e0;                 // uniform copy of E[0], the top entry
};

__device__ __forceinline__ void ar0_init (uint32_t *E, uint8_t *POS, uint32_t maxs, int lane, Ar0 &a)   // c_simple_model.h:85-103
{
    for (uint32_t i = lane; i < AR0_E_WORDS; i += 32) E[i] = i < maxs ? (1u | (i << 16)) : 0xffff0000u;
    if (POS) for (uint32_t i = lane; i < 256; i += 32) POS[i] = (uint8_t)i;
    __syncwarp ();
    const uint32_t lo = 8u * lane;
    a.s = lo >= maxs ? 0u : min (8u, maxs - lo);
    a.base = min (lo, maxs);
    a.tot = maxs; a.rtot = ar_rcp_below (maxs);
    a.e0 = 1u;
}

// normalize (c_simple_model.h:106-116) after entry p was bumped, then the bubble step (:140-145); rebuilds the lane sums
static __device__ __noinline__ Ar0 ar0_halve (uint32_t *E, uint8_t *POS, uint32_t maxs, Ar0 a, uint32_t p, int lane)   // state by value: a reference would pin the caller's registers in local memory
{
    __syncwarp ();
    const uint32_t lo = 8u * lane;
    uint32_t sum = 0;
    #pragma unroll
    for (uint32_t j = 0; j < 8; j++) {
        const uint32_t idx = lo + j;
        if (idx < maxs) { const uint32_t v = E[idx]; uint32_t g = v & 0xffffu; g -= g >> 1; E[idx] = (v & 0xffff0000u) | g; sum += g; }
    }
    __syncwarp ();
    uint32_t inc = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
    a.s = sum; a.base = inc - sum;
    a.tot = __shfl_sync (0xffffffffu, inc, 31); a.rtot = ar_rcp_below (a.tot);
    if (p) {
        const uint32_t en = E[p], prev = E[p - 1];
        __syncwarp ();
        if ((en & 0xffffu) > (prev & 0xffffu)) {
            if (lane == 0) {
                E[p - 1] = en; E[p] = prev;
                if (POS) { POS[en >> 16] = (uint8_t)(p - 1); POS[prev >> 16] = (uint8_t)p; }
            }
            if ((p & 7) == 0) {
                const uint32_t d = (en & 0xffffu) - (prev & 0xffffu), owner = p >> 3;
                if ((uint32_t)lane == owner - 1) a.s += d;
                if ((uint32_t)lane == owner) { a.s -= d; a.base += d; }
            }
        }
    }
    __syncwarp ();
    a.e0 = E[0];
    AR_READS_DONE ();
    return a;
}

// entry p (current value e, predecessor prev when p > 0) was coded: Freq += STEP, TotFreq += STEP, halve, bubble (:131-145)
__device__ __forceinline__ void ar0_update (uint32_t *E, uint8_t *POS, uint32_t maxs, Ar0 &a, uint32_t p, uint32_t e, uint32_t prev, int lane)
{
    // Every lane computes the same values; lane 0 alone stores them, and a __syncwarp () stands between its stores and the other
    // lanes' next reads (the reads before this call ended with AR_READS_DONE): no two lanes ever touch a word of the model without a
    // barrier in between (compute-sanitizer racecheck, profiles/r02_sanitizer.md).
    const uint32_t en = e + AR_STEP, owner = p >> 3;
    if (a.tot + AR_STEP > AR_MAXF) { if (lane == 0) E[p] = en; const Ar0 t = ar0_halve (E, POS, maxs, a, p, lane); a = t; return; }
    a.tot += AR_STEP; a.rtot = ar_rcp_below (a.tot);
    if ((uint32_t)lane == owner) a.s += AR_STEP; else if ((uint32_t)lane > owner) a.base += AR_STEP;
    if (p && (en & 0xffffu) > (prev & 0xffffu)) {
        if (lane == 0) {
            E[p - 1] = en; E[p] = prev;
            if (POS) { POS[en >> 16] = (uint8_t)(p - 1); POS[prev >> 16] = (uint8_t)p; }
        }
        if ((p & 7) == 0) {                                                  // the bubble step crossed a lane boundary
            const uint32_t d = (en & 0xffffu) - (prev & 0xffffu);
            if ((uint32_t)lane == owner - 1) a.s += d;
            if ((uint32_t)lane == owner) { a.s -= d; a.base += d; }
        }
        if (p == 1) a.e0 = en;
    }
    else { if (lane == 0) E[p] = en; if (p == 0) a.e0 = en; }
    __syncwarp ();
}

// copies the model into the global layout of arith_model.cuh (for the reference-exact tail after a corrupt / truncated stream)
__device__ __forceinline__ void ar0_export (uint32_t *E, const Ar0 &a, uint32_t *m, uint32_t maxs, int lane)
{
    if (lane == 0) E[0] = a.e0;                                              // (the top entry lives in a.e0 while it keeps being coded)
    __syncwarp ();
    for (uint32_t i = lane; i < maxs; i += 32) m[4 + i] = E[i];
    if (lane == 0) ar_store_head (m, a.tot, a.rtot);
    __syncwarp ();
}

// ---- decoder ---------------------------------------------------------------------------------------------------------------
// returns the number of symbols decoded before an anomaly (reference error return or dry input); n = all done
__device__ __forceinline__ uint32_t ar0_decode_run (uint32_t *E, uint32_t maxs, Ar0 &a, ArDec &rc, ArOut &o, uint32_t n, int lane)
{
    uint32_t i = 0;
    while (i < n) {
        const uint32_t r = ar_div (rc.range, a.tot, a.rtot);
        const uint32_t t1 = (a.e0 & 0xffffu) * r;
        uint32_t sym;
        if (rc.code < t1 && a.tot + AR_STEP <= AR_MAXF) {                    // the top entry again: registers + one shared store
            rc.range = t1;
            a.e0 += AR_STEP;                                                // (registers only: E[0] is written back when another entry is looked at)
            a.tot += AR_STEP; a.rtot = ar_rcp_below (a.tot);
            if (lane == 0) a.s += AR_STEP; else a.base += AR_STEP;
            sym = a.e0 >> 16;
        }
        else {
            if (lane == 0) E[0] = a.e0;
            __syncwarp ();
            const uint32_t T = (a.base + a.s) * r;                          // <= TotFreq * r <= range: no overflow
            const uint32_t ball = __ballot_sync (0xffffffffu, T > rc.code);
            if (!ball) { rc.range = r; ar_out_put (o, 0); return i + 1; }   // code/r >= TotFreq: the reference returns symbol 0 (:153-161)
            const int owner = __ffs (ball) - 1;
            uint32_t acc = __shfl_sync (0xffffffffu, a.base, owner);
            const uint4 v0 = *reinterpret_cast<const uint4 *>(E + 8 * owner), v1 = *reinterpret_cast<const uint4 *>(E + 8 * owner + 4);
            uint32_t e = 0, prev = 0, j = 8;
            do {                                                            // at most 8 uniform steps over the owner's entries
                #define AR0_STEP(J, EJ, EP) { const uint32_t f_ = (EJ) & 0xffffu; if (rc.code < (acc + f_) * r) { e = (EJ); prev = (EP); j = J; break; } acc += f_; }
                AR0_STEP (0, v0.x, 0u) AR0_STEP (1, v0.y, v0.x) AR0_STEP (2, v0.z, v0.y) AR0_STEP (3, v0.w, v0.z)
                AR0_STEP (4, v1.x, v0.w) AR0_STEP (5, v1.y, v1.x) AR0_STEP (6, v1.z, v1.y) AR0_STEP (7, v1.w, v1.z)
                #undef AR0_STEP
            } while (0);
            if (j == 8) { rc.range = r; ar_out_put (o, 0); return i + 1; }  // cannot happen: the owner's inclusive threshold exceeds code
            const uint32_t p = 8u * owner + j;
            if (j == 0 && p) prev = E[p - 1];
            rc.code -= acc * r; rc.range = (e & 0xffffu) * r;
            sym = e >> 16;
            AR_READS_DONE ();
            ar0_update (E, nullptr, maxs, a, p, e, prev, lane);
        }
        ar_out_put (o, sym);
        i++;
        if (rc.range < AR_TOP && !ar_dec_renorm (rc)) return i;             // the input ran dry
    }
    return n;
}

// ---- encoder ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ar0_encode_sym (uint32_t *E, uint8_t *POS, uint32_t maxs, Ar0 &a, ArEnc &rc, uint32_t sym, int lane)
{
    const uint32_t r = ar_div (rc.range, a.tot, a.rtot);
    if ((a.e0 >> 16) == sym && a.tot + AR_STEP <= AR_MAXF) {
        rc.range = (a.e0 & 0xffffu) * r;
        a.e0 += AR_STEP;                                                    // (registers only: E[0] is written back when another entry is looked at)
        a.tot += AR_STEP; a.rtot = ar_rcp_below (a.tot);
        if (lane == 0) a.s += AR_STEP; else a.base += AR_STEP;
        return;
    }
    if (lane == 0) E[0] = a.e0;
    __syncwarp ();
    const uint32_t p = POS[sym], owner = p >> 3, j = p & 7;
    const uint32_t e = E[p], prev = p ? E[p - 1] : 0u;
    uint32_t acc = __shfl_sync (0xffffffffu, a.base, owner);
    const uint4 v0 = *reinterpret_cast<const uint4 *>(E + 8 * owner), v1 = *reinterpret_cast<const uint4 *>(E + 8 * owner + 4);
    switch (j) {                                                            // frequencies before entry j inside its lane
        case 7: acc += v1.z & 0xffffu;
        case 6: acc += v1.y & 0xffffu;
        case 5: acc += v1.x & 0xffffu;
        case 4: acc += v0.w & 0xffffu;
        case 3: acc += v0.z & 0xffffu;
        case 2: acc += v0.y & 0xffffu;
        case 1: acc += v0.x & 0xffffu;
        default: break;
    }
    const uint32_t before = rc.low;
    rc.low += acc * r; rc.range = (e & 0xffffu) * r;
    rc.carry += rc.low < before;
    AR_READS_DONE ();
    ar0_update (E, POS, maxs, a, p, e, prev, lane);
}

__device__ __forceinline__ uint32_t ar0_encode_leaf (uint32_t *E, uint8_t *POS, uint32_t maxs, const uint8_t *in, uint32_t n, uint8_t *out, int lane)
{
    Ar0 a; ar0_init (E, POS, maxs, lane, a);
    out[0] = (uint8_t)maxs;                                                  // arith_dynamic.c:105-110 (256 wraps to 0)
    ArEnc rc; rc.low = 0; rc.range = 0xffffffffu; rc.ffnum = 0; rc.cache = 0; rc.carry = 0; rc.out = out + 1;
    const uint8_t *limit = out + n + 8;
    uint32_t s_next = n ? __ldg (in) : 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t s = s_next;
        if (i + 1 < n) s_next = __ldg (in + i + 1);
        ar0_encode_sym (E, POS, maxs, a, rc, s, lane);
        if (rc.range < AR_TOP) {
            do { rc.range <<= 8; ar_shift_low (rc); } while (rc.range < AR_TOP);
            if (rc.out + rc.ffnum > limit) return n + 1;
        }
    }
    for (int i = 0; i < 5; i++) ar_shift_low (rc);                          // RC_FinishEncode
    return (uint32_t)(rc.out - out);
}

} // namespace gzb
