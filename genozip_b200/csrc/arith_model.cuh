// arith_model.cuh — the htscodecs adaptive arithmetic coder (reference src/htscodecs/c_simple_model.h:77-179 model,
// c_range_coder.h:46-126 range coder, arith_dynamic.c:92-226 / 387-608 loops) as ONE dependency chain per leaf, written
// for the latency of a single warp:
//
//   * model layout made for vector loads:   word 0 TotFreq | word 1 float bits of a reciprocal of TotFreq rounded DOWN |
//     words 4.. entries  Freq | Symbol << 16  (approximately sorted by frequency, like the reference's list), padded with
//     Freq 0 / Symbol 0xffff entries to a multiple of 8;
//   * the model of the CURRENT context (TotFreq, reciprocal, first four entries) lives in registers; every update is
//     written through to memory, so a context switch is two loads and a repeated context costs none;
//   * range / TotFreq is one float multiply with the stored reciprocal plus an exact integer correction;
//   * the decoder never divides code by range: "AccFreq <= code / range" (c_simple_model.h:156) is evaluated as
//     "AccFreq * range <= code" on the first four entries;
//   * symbols beyond the first four entries are located by the whole warp, 8 entries per lane;
//   * everything else (halving, bubble step beyond the cached entries) works on memory and reloads the registers.
//
// All lanes of the warp carry the coder state redundantly (uniform execution); stores of identical values to identical
// addresses merge into one transaction.  The same source builds for the host with one "lane" (tests/host_arith.cpp):
// the CPU suite checks this logic against the oracle without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) || defined(GZB_SIMT_EMULATION)       // (the second: tests/host/simt, g++ build of this source for the SIMT emulator)
  #define AR_FN __device__ __forceinline__
  #define AR_SLOW static __device__ __noinline__
#else
  #include <string.h>
  #define AR_FN static inline
  #define AR_SLOW static
  struct uint2 { uint32_t x, y; };
  struct uint4 { uint32_t x, y, z, w; };
  static inline uint32_t __ldg (const uint8_t *p) { return *p; }
  static inline void __syncwarp () {}
  static inline uint32_t __byte_perm (uint32_t a, uint32_t b, uint32_t s)
  {
      const uint64_t v = ((uint64_t)b << 32) | a; uint32_t r = 0;
      for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
      return r;
  }
#endif

namespace gzb {

// All 32 lanes of the warp run the chain redundantly and store identical values to the models.  A lane must therefore not
// store to a model before every lane has read it: AR_READS_DONE () — a __syncwarp () — stands after every group of model
// reads that a store may follow.  (Round 1 relied on the lanes of a converged warp issuing together and compiled the marker
// to nothing; measured on B200 the real barrier costs 0.3 % of the step — gpurun_out/ab_main*.json of round 2 — so the code
// no longer depends on more than CUDA guarantees.  -DGZB_READS_DONE_LOCKSTEP restores the old build for an A/B run.)
#if defined(GZB_READS_DONE_LOCKSTEP) && !defined(GZB_SIMT_EMULATION)
  #define AR_READS_DONE()
#else
  #define AR_READS_DONE() __syncwarp ()
#endif

#define AR_MAXF  65519u          // MAX_FREQ = (1<<16)-17 (c_simple_model.h:70)
#define AR_STEP  16u             // STEP (:73)
#define AR_TOP   (1u << 24)      // TOP (c_range_coder.h:22)

AR_FN uint32_t ar_stride (uint32_t maxs) { return 4 + ((maxs + 7) & ~7u); }
constexpr uint32_t AR_RUN_STRIDE = 12;           // run-length models: 4 live symbols (MAX_RUN, arith_dynamic.c:383)

// ---- reciprocal and division -------------------------------------------------------------------------------------
// 1/x rounded safely DOWN: MUFU.RCP (<= 1 ulp) scaled by (1 - 5e-7); relative deficit < 7e-7
AR_FN float ar_rcp_below (uint32_t x)
{
#ifdef __CUDA_ARCH__
    float r;
    asm ("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__uint2float_ru (x)));
    return __fmul_rz (r, 0.9999995f);
#else
    return (float)(0.999999 / (double)x);
#endif
}

// exact a / d for d < 2^17 given rd <= 1/d: the float estimate never exceeds the quotient; its deficit is corrected
// in integers (one step almost always: the deficit is <= q * 1e-6 + 1)
AR_FN uint32_t ar_div (uint32_t a, uint32_t d, float rd)
{
#ifdef __CUDA_ARCH__
    uint32_t q = __float2uint_rz (__fmul_rz (__uint2float_rz (a), rd));
    uint32_t r = a - q * d;
    if (r >= d) {
        q++; r -= d;
        if (r >= d) {                                                       // small divisors (young models): a second float round
            const uint32_t q2 = __float2uint_rz (__fmul_rz (__uint2float_rz (r), rd));
            q += q2; r -= q2 * d;
            while (r >= d) { q++; r -= d; }
        }
    }
    return q;
#else
    (void)rd;
    return a / d;
#endif
}

// ---- warp primitives (one lane on the host) ------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
AR_FN uint32_t ar_wsum (uint32_t v) { return __reduce_add_sync (0xffffffffu, v); }
#else
AR_FN uint32_t ar_wsum (uint32_t v) { return v; }
#endif

AR_FN uint32_t ar_f2u (float f) { uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint (f);
#else
    memcpy (&u, &f, 4);
#endif
    return u; }
AR_FN float ar_u2f (uint32_t u) { float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float (u);
#else
    memcpy (&f, &u, 4);
#endif
    return f; }

// ---- model memory ------------------------------------------------------------------------------------------------------
AR_FN void ar_model_init (uint32_t *m, uint32_t maxs)                       // c_simple_model.h:85-103
{
    m[0] = maxs; m[1] = ar_f2u (ar_rcp_below (maxs)); m[2] = 0; m[3] = 0;
    for (uint32_t i = 0; i < maxs; i++) m[4 + i] = 1u | (i << 16);
    const uint32_t st = ar_stride (maxs);
    for (uint32_t i = 4 + maxs; i < st; i++) m[i] = 0xffff0000u;            // Freq 0: never selected; Symbol 0xffff: never matches
}

struct ArCache { uint32_t tot; float rtot; uint32_t e0, e1, e2, e3; };      // the current context's model head, in registers

AR_FN void ar_load (const uint32_t *m, ArCache &c)
{
    const uint2 h = *reinterpret_cast<const uint2 *>(m);
    const uint4 v = *reinterpret_cast<const uint4 *>(m + 4);
    c.tot = h.x; c.rtot = ar_u2f (h.y); c.e0 = v.x; c.e1 = v.y; c.e2 = v.z; c.e3 = v.w;
    AR_READS_DONE ();
}

AR_FN void ar_store_head (uint32_t *m, uint32_t tot, float rtot)
{
    uint2 h; h.x = tot; h.y = ar_f2u (rtot);
    *reinterpret_cast<uint2 *>(m) = h;
}

AR_FN void ar_flush (uint32_t *m, const ArCache &c)                        // registers -> memory (write-back of the run loop's updates)
{
    ar_store_head (m, c.tot, c.rtot);
    uint4 v; v.x = c.e0; v.y = c.e1; v.z = c.e2; v.w = c.e3;
    *reinterpret_cast<uint4 *>(m + 4) = v;
}

// Update of entry p directly on memory: Freq += STEP, TotFreq += STEP, halve everything past MAX_FREQ (normalize,
// :106-116), one bubble step towards the front (:140-145).  e = current value of entry p, tot = current TotFreq.
// Used for entries beyond the cached four and for every update that triggers the halving.  Leaves memory
// authoritative; the caller reloads its registers.
AR_SLOW void ar_update_mem (uint32_t *m, uint32_t maxs, uint32_t p, uint32_t e, uint32_t tot, int lane)
{
    uint32_t en = e + AR_STEP;
    tot += AR_STEP;
    if (tot > AR_MAXF) {
        m[4 + p] = en;
        __syncwarp ();
        uint32_t sum = 0;
#ifdef __CUDA_ARCH__
        for (uint32_t j = lane; j < maxs; j += 32)
#else
        for (uint32_t j = 0; j < maxs; j++)
#endif
        { const uint32_t v = m[4 + j]; uint32_t g = v & 0xffffu; g -= g >> 1; m[4 + j] = (v & 0xffff0000u) | g; sum += g; }
        tot = ar_wsum (sum);
        __syncwarp ();
        en = m[4 + p];
    }
    if (p) {
        const uint32_t prev = m[4 + p - 1];
        AR_READS_DONE ();
        if ((en & 0xffffu) > (prev & 0xffffu)) { m[4 + p - 1] = en; m[4 + p] = prev; }
        else m[4 + p] = en;
    }
    else m[4] = en;
    ar_store_head (m, tot, ar_rcp_below (tot));
    (void)lane;
}

// ---- warp-wide searches beyond the cached entries: lane l owns entries 8l .. 8l+7 ----------------------------------------
// decoder: the first entry p whose inclusive cumulative frequency times r exceeds code — the reference's
// "AccFreq += Freq until it exceeds code / r" (c_simple_model.h:156) without dividing.  Returns p (>= maxs: none — corrupt
// input), acc = cumulative frequency before p, e = entry p, prev = entry p-1 (p > 0).
// One scan of the 32 lane sums locates the owning lane; its 8 entries are then walked by all lanes together (uniform,
// early exit), so no lane-indexed selection and no dependent load follows.
// (results by value: a reference into the caller's frame would put those values in local memory)
struct ArHit { uint32_t p, acc, e, prev; };
AR_SLOW ArHit ar_find_code (const uint32_t *m, uint32_t maxs, uint32_t code, uint32_t r, int lane)
{
    ArHit h; h.p = maxs; h.acc = 0; h.e = 0; h.prev = 0;
    uint32_t acc = 0, e = 0, prev = 0;
#ifdef __CUDA_ARCH__
    const uint32_t j0 = 8u * lane;
    uint4 a = make_uint4 (0, 0, 0, 0), b = a;
    if (j0 < maxs) { a = *reinterpret_cast<const uint4 *>(m + 4 + j0); b = *reinterpret_cast<const uint4 *>(m + 8 + j0); }
    const uint32_t s = ((a.x & 0xffffu) + (a.y & 0xffffu)) + ((a.z & 0xffffu) + (a.w & 0xffffu)) +
                       ((b.x & 0xffffu) + (b.y & 0xffffu)) + ((b.z & 0xffffu) + (b.w & 0xffffu));
    uint32_t inc = s;                                                       // inclusive scan of the lane sums
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
    const uint32_t ball = __ballot_sync (0xffffffffu, j0 < maxs && inc * r > code);      // inc <= TotFreq: inc * r <= range, no overflow
    if (!ball) return h;
    const uint32_t owner = __ffs (ball) - 1;
    acc = __shfl_sync (0xffffffffu, inc - s, owner);
    const uint32_t *q = m + 4 + 8 * owner;
    const uint4 v0 = *reinterpret_cast<const uint4 *>(q), v1 = *reinterpret_cast<const uint4 *>(q + 4);
    const uint32_t pl = q[-1];                                              // entry before the lane's first (owner 0: header padding, unused)
    AR_READS_DONE ();
    uint32_t j = 8;
    do {
        #define AR_WALK(J, EJ, EP) { const uint32_t f_ = (EJ) & 0xffffu; if (code < (acc + f_) * r) { e = (EJ); prev = (EP); j = J; break; } acc += f_; }
        AR_WALK (0, v0.x, pl)   AR_WALK (1, v0.y, v0.x) AR_WALK (2, v0.z, v0.y) AR_WALK (3, v0.w, v0.z)
        AR_WALK (4, v1.x, v0.w) AR_WALK (5, v1.y, v1.x) AR_WALK (6, v1.z, v1.y) AR_WALK (7, v1.w, v1.z)
        #undef AR_WALK
    } while (0);
    if (j == 8) return h;                                                  // cannot happen: the owner's inclusive threshold exceeds code
    h.p = 8 * owner + j; h.acc = acc; h.e = e; h.prev = prev;
    return h;
#else
    (void)lane;
    acc = 0; prev = 0;
    for (uint32_t p = 0; p < maxs; p++) {
        e = m[4 + p];
        if (code < (acc + (e & 0xffffu)) * r) { h.p = p; h.acc = acc; h.e = e; h.prev = prev; return h; }
        acc += e & 0xffffu; prev = e;
    }
    return h;
#endif
}

// encoder: the entry holding `sym` (always present for sym < maxs); prev = the entry before it (p > 0)
AR_SLOW ArHit ar_find_sym (const uint32_t *m, uint32_t maxs, uint32_t sym, int lane)
{
    ArHit h; h.p = maxs; h.acc = 0; h.e = 0; h.prev = 0;
    uint32_t acc = 0, e = 0, prev = 0;
#ifdef __CUDA_ARCH__
    const uint32_t j0 = 8u * lane;
    uint4 a = make_uint4 (0xffff0000u, 0xffff0000u, 0xffff0000u, 0xffff0000u), b = a;
    if (j0 < maxs) { a = *reinterpret_cast<const uint4 *>(m + 4 + j0); b = *reinterpret_cast<const uint4 *>(m + 8 + j0); }
    const uint32_t lastprev = __shfl_up_sync (0xffffffffu, b.w, 1);         // the previous lane's last entry
    uint32_t mj = 8, esel = 0, psel = 0, part = 0;                          // first match inside the lane's 8 entries; frequencies before it
    if ((b.w >> 16) == sym) { mj = 7; esel = b.w; psel = b.z; }
    if ((b.z >> 16) == sym) { mj = 6; esel = b.z; psel = b.y; }
    if ((b.y >> 16) == sym) { mj = 5; esel = b.y; psel = b.x; }
    if ((b.x >> 16) == sym) { mj = 4; esel = b.x; psel = a.w; }
    if ((a.w >> 16) == sym) { mj = 3; esel = a.w; psel = a.z; }
    if ((a.z >> 16) == sym) { mj = 2; esel = a.z; psel = a.y; }
    if ((a.y >> 16) == sym) { mj = 1; esel = a.y; psel = a.x; }
    if ((a.x >> 16) == sym) { mj = 0; esel = a.x; psel = lastprev; }
    if (mj > 0) part += a.x & 0xffffu;
    if (mj > 1) part += a.y & 0xffffu;
    if (mj > 2) part += a.z & 0xffffu;
    if (mj > 3) part += a.w & 0xffffu;
    if (mj > 4) part += b.x & 0xffffu;
    if (mj > 5) part += b.y & 0xffffu;
    if (mj > 6) part += b.z & 0xffffu;
    if (mj > 7) part += b.w & 0xffffu;
    const uint32_t hit = __ballot_sync (0xffffffffu, mj < 8);
    if (!hit) return h;
    const int w = __ffs (hit) - 1;
    h.p = 8u * w + __shfl_sync (0xffffffffu, mj, w);
    h.acc = ar_wsum (lane <= w ? part : 0u);
    h.e = __shfl_sync (0xffffffffu, esel, w);
    h.prev = __shfl_sync (0xffffffffu, psel, w);
    return h;
#else
    (void)lane;
    acc = 0; prev = 0;
    for (uint32_t p = 0; p < maxs; p++) {
        e = m[4 + p];
        if ((e >> 16) == sym) { h.p = p; h.acc = acc; h.e = e; h.prev = prev; return h; }
        acc += e & 0xffffu; prev = e;
    }
    return h;
#endif
}

// entry p >= 4 (beyond the cached four) was coded: update in memory and in the cached head; the halving goes the long way
#define AR_BUMP_DEEP(P, E, PREV)                                                                                \
    {                                                                                                           \
        if (c.tot + AR_STEP > AR_MAXF) { ar_update_mem (m, maxs, P, E, c.tot, lane); stale = true; }            \
        else {                                                                                                  \
            const uint32_t en_ = (E) + AR_STEP;                                                                 \
            c.tot += AR_STEP; c.rtot = ar_rcp_below (c.tot);                                                    \
            ar_store_head (m, c.tot, c.rtot);                                                                   \
            if ((en_ & 0xffffu) > ((PREV) & 0xffffu)) { m[4 + (P) - 1] = en_; m[4 + (P)] = (PREV); if ((P) == 4) c.e3 = en_; } \
            else m[4 + (P)] = en_;                                                                              \
        }                                                                                                       \
    }

// ---- cached fast-path update ---------------------------------------------------------------------------------------------
// entry K (0..3) of the cached model was coded: registers and memory are updated identically unless the halving
// triggers, in which case memory is updated and `stale` asks the caller to reload.
#define AR_BUMP_CASE(K, EK, EPREV)                                                                              \
    {                                                                                                           \
        if (c.tot + AR_STEP > AR_MAXF) { ar_update_mem (m, maxs, K, EK, c.tot, lane); stale = true; }           \
        else {                                                                                                  \
            c.tot += AR_STEP; EK += AR_STEP;                                                                    \
            c.rtot = ar_rcp_below (c.tot);                                                                      \
            if (K > 0 && (EK & 0xffffu) > (EPREV & 0xffffu)) { const uint32_t t_ = EK; EK = EPREV; EPREV = t_; } \
            ar_store_head (m, c.tot, c.rtot);                                                                   \
            uint4 v_; v_.x = c.e0; v_.y = c.e1; v_.z = c.e2; v_.w = c.e3;                                       \
            *reinterpret_cast<uint4 *>(m + 4) = v_;                                                             \
        }                                                                                                       \
    }

// ---- decoder ---------------------------------------------------------------------------------------------------------------
struct ArDec { uint32_t code, range, ipos, ilen; const uint8_t *in; };

// One symbol (SIMPLE_MODEL_decodeSymbol :148-179 + RC_GetFreq/RC_Decode, c_range_coder.h:111-126).  Returns the symbol;
// `anomaly` is set when the reference's error return was taken (symbol 0, range divided, nothing else changes).
// ANY: also handles range < TotFreq (possible only after the input ran dry): the reference then codes entry 0 without
// dividing the range.
template <bool ANY>
AR_FN uint32_t ar_decode_sym (uint32_t *m, uint32_t maxs, ArCache &c, ArDec &rc, int lane, bool &stale, bool &anomaly, uint32_t r_in = 0)
{
    uint32_t sym;
    const bool forced = ANY && rc.range < c.tot;
    const uint32_t r = ANY ? (forced ? rc.range : ar_div (rc.range, c.tot, c.rtot)) : r_in;   // !ANY: the caller has divided already
    const uint32_t f0 = c.e0 & 0xffffu;
    const uint32_t t1 = f0 * r;
    if (rc.code < t1 || forced) {
        rc.range = forced ? rc.range * f0 : t1;
        sym = c.e0 >> 16;
        AR_BUMP_CASE (0, c.e0, c.e0)
    }
    else {
        const uint32_t f1 = c.e1 & 0xffffu, f2 = c.e2 & 0xffffu, f3 = c.e3 & 0xffffu;
        const uint32_t t2 = t1 + f1 * r, t3 = t2 + f2 * r, t4 = t3 + f3 * r;
        if (rc.code < t2)      { rc.code -= t1; rc.range = f1 * r; sym = c.e1 >> 16; AR_BUMP_CASE (1, c.e1, c.e0) }
        else if (rc.code < t3) { rc.code -= t2; rc.range = f2 * r; sym = c.e2 >> 16; AR_BUMP_CASE (2, c.e2, c.e1) }
        else if (rc.code < t4) { rc.code -= t3; rc.range = f3 * r; sym = c.e3 >> 16; AR_BUMP_CASE (3, c.e3, c.e2) }
        else {
            const ArHit h = ar_find_code (m, maxs, rc.code, r, lane);
            const uint32_t p = h.p, acc = h.acc, e = h.e, prev = h.prev;
            if (p >= maxs) { rc.range = r; anomaly = true; return 0; }    // code / r >= TotFreq (or > MAX_FREQ): :153-154, :160-161
            rc.code -= acc * r; rc.range = (e & 0xffffu) * r;
            sym = e >> 16;
            AR_BUMP_DEEP (p, e, prev)
        }
    }
    return sym;
}

// RC_Decode's renormalisation (:119-125); false = the input ran dry with range still below TOP
AR_FN bool ar_dec_renorm (ArDec &rc)
{
    while (rc.range < AR_TOP) {
        if (rc.ipos >= rc.ilen) return false;
        rc.code = (rc.code << 8) + __ldg (rc.in + rc.ipos++);
        rc.range <<= 8;
    }
    return true;
}

// Output bytes are gathered in a 32-bit window and written one aligned word at a time (a byte store per symbol from
// hundreds of concurrent leaves is what the L2 write path chokes on); head and tail bytes go out singly.
struct ArOut {
    uint8_t *wptr;              // address of the aligned word being filled
    uint32_t pos, head, win;    // pos = (out & 3) + symbols written; head = out & 3
};
AR_FN void ar_out_init (ArOut &o, uint8_t *out) { o.head = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 3); o.wptr = out - o.head; o.pos = o.head; o.win = 0; }
AR_FN void ar_out_put (ArOut &o, uint32_t b)
{
    o.win = __byte_perm (o.win, b, 0x4321);
    o.pos++;
    if ((o.pos & 3) == 0) {
        if (o.pos == 4 && o.head) { for (uint32_t t = o.head; t < 4; t++) o.wptr[t] = (uint8_t)(o.win >> (8 * t)); }
        else *reinterpret_cast<uint32_t *>(o.wptr) = o.win;
        o.wptr += 4;
    }
}
AR_FN void ar_out_flush (ArOut &o)
{
    const uint32_t tail = o.pos & 3;                                        // bytes after the last aligned word boundary
    const uint32_t first = (o.pos < 4) ? o.head : 0;                        // never touch bytes before the leaf's output
    for (uint32_t t = first; t < tail; t++) o.wptr[t] = (uint8_t)(o.win >> (8 * (4 - tail + t)));
}

AR_FN void ar_dec_start (ArDec &rc, const uint8_t *body, uint32_t body_len)   // RC_StartDecode (c_range_coder.h:57-68); body[0] = max_sym
{
    rc.range = 0xffffffffu; rc.code = 0; rc.in = body; rc.ipos = 1; rc.ilen = body_len;
    if (rc.ipos + 5 > rc.ilen) rc.ipos = rc.ilen;
    else for (int i = 0; i < 5; i++) rc.code = (rc.code << 8) | body[rc.ipos++];
}

// symbols i .. n-1 after an anomaly (reference error return, or input exhausted): the reference's exact (odd) behaviour,
// every symbol through memory
template <bool O1>
AR_FN void ar_decode_tail (uint32_t *lit, uint32_t maxs, ArDec &rc, ArOut &o, uint32_t i, uint32_t n, uint32_t ctx, int lane)
{
    const uint32_t stride = ar_stride (maxs);
    uint32_t *m = lit + (O1 ? ctx : 0) * stride;
    ArCache c;
    for (; i < n; i++) {
        bool stale = false, anomaly = false;
        ar_load (m, c);
        const uint32_t s = ar_decode_sym<true> (m, maxs, c, rc, lane, stale, anomaly);
        if (!anomaly) ar_dec_renorm (rc);
        ar_out_put (o, s);
        if (O1) m = lit + s * stride;
    }
}

// arith_uncompress_O0 / O1 (arith_dynamic.c:129-152, 200-226) and the RLE variants (:451-493, :564-608)
template <bool O1>
AR_FN void ar_decode_leaf (uint32_t *lit, uint32_t maxs, bool rle, const uint8_t *body, uint32_t body_len, uint8_t *out, uint32_t n, int lane, uint32_t run4 = 2)
{
    const uint32_t stride = ar_stride (maxs);
    uint32_t *run = lit + (O1 ? 256 : 1) * stride;
    ArDec rc; ar_dec_start (rc, body, body_len);
    ArOut o; ar_out_init (o, out);
    ArCache c;
    uint32_t ctx = 0, i = 0;
    uint32_t *m = lit;
    if (!rle) {
        // fast loop: invariant range >= TOP (so range >= TotFreq); left for good at the first anomaly.
        // RUN STEP: the context's top entry is decoded again and (order 1) it is the context itself — by far the most
        // frequent case of a low-entropy stream.  It touches registers only: the model head is written back (ar_flush)
        // when another kind of symbol turns up.
        ar_load (m, c);
        bool ok = true, dirty = false;
        bool selfloop = !O1 || (c.e0 >> 16) == ctx;
        uint32_t skip4 = 0, fail4 = 0;                                       // back-off of the 4-step speculation below
        while (i < n && ok) {
            // FOUR RUN STEPS AT ONCE, speculatively.  A run step leaves the code alone (the top entry's cumulative frequency is 0)
            // and only shrinks the range, so four of them are a straight line of four divisions; the halving is excluded up front.
            // The ranges only shrink (freq <= TotFreq), so the steps 1 .. k are what the single steps would have done iff code < g_k
            // and no renormalisation was due before step k (g_(k-1) >= TOP): the longest such prefix is taken (mode 2; mode 1 takes all
            // four or nothing), the renormalisation after step k is the ordinary one.  Nothing has been changed if k = 0.
            if (run4 && selfloop && skip4 == 0 && i + 4 <= n && c.tot + 4 * AR_STEP <= AR_MAXF && (run4 == 2 || ((o.pos & 3) == 0 && o.pos >= 4))) {
                const uint32_t f0 = c.e0 & 0xffffu;
                const float rt1 = ar_rcp_below (c.tot + AR_STEP), rt2 = ar_rcp_below (c.tot + 2 * AR_STEP), rt3 = ar_rcp_below (c.tot + 3 * AR_STEP);
                const uint32_t g1 = f0 * ar_div (rc.range, c.tot, c.rtot);
                const uint32_t g2 = (f0 + AR_STEP) * ar_div (g1, c.tot + AR_STEP, rt1);
                const uint32_t g3 = (f0 + 2 * AR_STEP) * ar_div (g2, c.tot + 2 * AR_STEP, rt2);
                const uint32_t g4 = (f0 + 3 * AR_STEP) * ar_div (g3, c.tot + 3 * AR_STEP, rt3);
                uint32_t k, gk;
                if (run4 == 2) {
                    if (rc.code < g4 && g3 >= AR_TOP)      { k = 4; gk = g4; }
                    else if (rc.code < g3 && g2 >= AR_TOP) { k = 3; gk = g3; }
                    else if (rc.code < g2 && g1 >= AR_TOP) { k = 2; gk = g2; }
                    else if (rc.code < g1)                 { k = 1; gk = g1; }
                    else                                   { k = 0; gk = 0; }
                }
                else { k = (rc.code < g4 && g4 >= AR_TOP) ? 4 : 0; gk = g4; }
                if (k) {
                    rc.range = gk;
                    c.e0 += k * AR_STEP; c.tot += k * AR_STEP; c.rtot = ar_rcp_below (c.tot);
                    dirty = true;
                    if (k == 4 && (o.pos & 3) == 0 && o.pos >= 4) { *reinterpret_cast<uint32_t *>(o.wptr) = (c.e0 >> 16) * 0x01010101u; o.wptr += 4; o.pos += 4; }
                    else for (uint32_t j = 0; j < k; j++) ar_out_put (o, c.e0 >> 16);
                    i += k; fail4 = 0;
                    if (rc.range < AR_TOP) ok = ar_dec_renorm (rc);
                    else if (k < 4) skip4 = 1;                              // the run ended on another symbol: that one goes the single-step way
                    continue;
                }
                fail4 = fail4 < 4 ? fail4 + 1 : 4; skip4 = run4 == 2 ? 2u : 1u << fail4;   // single steps before the next attempt
            }
            else if (skip4) skip4--;
            const uint32_t r = ar_div (rc.range, c.tot, c.rtot);
            const uint32_t t1 = (c.e0 & 0xffffu) * r;
            if (rc.code < t1 && selfloop && c.tot + AR_STEP <= AR_MAXF) {
                rc.range = t1;
                c.e0 += AR_STEP; c.tot += AR_STEP; c.rtot = ar_rcp_below (c.tot);
                dirty = true;
                ar_out_put (o, c.e0 >> 16);
                i++;
                if (rc.range < AR_TOP) ok = ar_dec_renorm (rc);
                continue;
            }
            if (dirty) { ar_flush (m, c); dirty = false; }
            bool stale = false, anomaly = false;
            const uint32_t s = ar_decode_sym<false> (m, maxs, c, rc, lane, stale, anomaly, r);
            if (anomaly) ok = false;
            else if (rc.range < AR_TOP) ok = ar_dec_renorm (rc);
            ar_out_put (o, s);
            i++;
            if (O1 && s != ctx) { ctx = s; m = lit + s * stride; ar_load (m, c); }
            else if (stale) ar_load (m, c);
            selfloop = !O1 || (c.e0 >> 16) == ctx;
        }
        if (dirty) ar_flush (m, c);
        ar_decode_tail<O1> (lit, maxs, rc, o, i, n, ctx, lane);
    }
    else {
        uint32_t last = 0;
        for (; i < n; i++) {
            bool stale = false, anomaly = false;
            m = lit + (O1 ? last : 0) * stride;
            ar_load (m, c);
            const uint32_t s = ar_decode_sym<true> (m, maxs, c, rc, lane, stale, anomaly);
            if (!anomaly) ar_dec_renorm (rc);
            ar_out_put (o, s);
            last = s;
            uint32_t r = 0, part, rctx = last;                              // arith_dynamic.c:473-482 / :591-599
            do {
                uint32_t *rm = run + rctx * AR_RUN_STRIDE;
                anomaly = false;
                ar_load (rm, c);
                part = ar_decode_sym<true> (rm, 4, c, rc, lane, stale, anomaly);
                if (!anomaly) ar_dec_renorm (rc);
                if (rctx == last) rctx = 256; else rctx += (rctx < 257);
                r += part;
            } while (part == 3 && r < n);
            while (r-- && i + 1 < n) { ++i; ar_out_put (o, last); }
        }
    }
    ar_out_flush (o);
}

// ---- encoder ---------------------------------------------------------------------------------------------------------------
struct ArEnc { uint32_t low, range, ffnum, cache, carry; uint8_t *out; };

AR_FN void ar_shift_low (ArEnc &rc)                                          // c_range_coder.h:70-88
{
    if (rc.low < (255u << 24) || rc.carry) {
        *rc.out = (uint8_t)(rc.cache + rc.carry);
        for (uint32_t i = 0; i < rc.ffnum; i++) rc.out[1 + i] = (uint8_t)(rc.carry - 1);
        rc.out += 1 + rc.ffnum; rc.ffnum = 0;
        rc.cache = rc.low >> 24;
        rc.carry = 0;
    }
    else rc.ffnum++;
    rc.low <<= 8;
}

// SIMPLE_MODEL_encodeSymbol (:123-146) + RC_Encode (c_range_coder.h:97-109) without its renormalisation loop
AR_FN void ar_encode_sym (uint32_t *m, uint32_t maxs, ArCache &c, ArEnc &rc, uint32_t sym, int lane, bool &stale)
{
    const uint32_t r = ar_div (rc.range, c.tot, c.rtot);
    const uint32_t before = rc.low;
    if ((c.e0 >> 16) == sym) {
        rc.range = (c.e0 & 0xffffu) * r;
        AR_BUMP_CASE (0, c.e0, c.e0)
        return;
    }
    const uint32_t f0 = c.e0 & 0xffffu;
    if ((c.e1 >> 16) == sym) {
        rc.low += f0 * r; rc.range = (c.e1 & 0xffffu) * r;
        AR_BUMP_CASE (1, c.e1, c.e0)
    }
    else if ((c.e2 >> 16) == sym) {
        rc.low += (f0 + (c.e1 & 0xffffu)) * r; rc.range = (c.e2 & 0xffffu) * r;
        AR_BUMP_CASE (2, c.e2, c.e1)
    }
    else if ((c.e3 >> 16) == sym) {
        rc.low += (f0 + (c.e1 & 0xffffu) + (c.e2 & 0xffffu)) * r; rc.range = (c.e3 & 0xffffu) * r;
        AR_BUMP_CASE (3, c.e3, c.e2)
    }
    else {
        const ArHit h = ar_find_sym (m, maxs, sym, lane);
        const uint32_t p = h.p, acc = h.acc, e = h.e, prev = h.prev;
        if (p >= maxs) return;                                              // cannot happen for a symbol < maxs
        rc.low += acc * r; rc.range = (e & 0xffffu) * r;
        AR_BUMP_DEEP (p, e, prev)
    }
    rc.carry += rc.low < before;
}

// arith_compress_O0 / O1 (arith_dynamic.c:92-126, 157-197) and the RLE variants (:387-448, :496-561).  Returns the body
// length, or n + 1 once the body is certain to reach the input length (it is then discarded for a raw copy, :847-852;
// stopping early also bounds the scratch a hostile, expanding input can touch).
template <bool O1>
AR_FN uint32_t ar_encode_leaf (uint32_t *lit, uint32_t maxs, bool rle, const uint8_t *in, uint32_t n, uint8_t *out, int lane)
{
    const uint32_t stride = ar_stride (maxs);
    uint32_t *run = lit + (O1 ? 256 : 1) * stride;
    out[0] = (uint8_t)maxs;                                                  // arith_dynamic.c:105-110 (256 wraps to 0)
    ArEnc rc; rc.low = 0; rc.range = 0xffffffffu; rc.ffnum = 0; rc.cache = 0; rc.carry = 0; rc.out = out + 1;
    const uint8_t *limit = out + n + 8;
    ArCache c;
    uint32_t *m = lit;
    if (!rle) {
        uint32_t ctx = 0;
        ar_load (m, c);
        bool dirty = false;
        uint32_t s_next = n ? __ldg (in) : 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t s = s_next;
            if (i + 1 < n) s_next = __ldg (in + i + 1);
            // RUN STEP (see ar_decode_leaf): the symbol is the context's top entry and the context does not change
            if ((c.e0 >> 16) == s && (!O1 || s == ctx) && c.tot + AR_STEP <= AR_MAXF) {
                const uint32_t r = ar_div (rc.range, c.tot, c.rtot);
                rc.range = (c.e0 & 0xffffu) * r;
                c.e0 += AR_STEP; c.tot += AR_STEP; c.rtot = ar_rcp_below (c.tot);
                dirty = true;
            }
            else {
                if (dirty) { ar_flush (m, c); dirty = false; }
                // the next symbol's context is this symbol: its model is fetched while this symbol is coded (this symbol
                // only touches the current context's model)
                ArCache nx; nx = c;
                const bool sw = O1 && s != ctx;
                if (sw) ar_load (lit + s * stride, nx);
                bool stale = false;
                ar_encode_sym (m, maxs, c, rc, s, lane, stale);
                if (sw) { c = nx; ctx = s; m = lit + s * stride; }
                else if (stale) ar_load (m, c);
            }
            if (rc.range < AR_TOP) {
                do { rc.range <<= 8; ar_shift_low (rc); } while (rc.range < AR_TOP);
                if (rc.out + rc.ffnum > limit) return n + 1;
            }
        }
    }
    else {
        uint32_t last = 0;
        for (uint32_t i = 0; i < n; ) {
            if (rc.out + rc.ffnum > limit) return n + 1;
            bool stale = false;
            const uint32_t s = __ldg (in + i);
            m = lit + (O1 ? last : 0) * stride;
            ar_load (m, c);
            ar_encode_sym (m, maxs, c, rc, s, lane, stale);
            while (rc.range < AR_TOP) { rc.range <<= 8; ar_shift_low (rc); }
            last = s; i++;
            uint32_t r = 0;                                                 // :413-438 run length in base-4 digits
            while (i < n && __ldg (in + i) == last) { r++; i++; }
            uint32_t rctx = last;
            do {
                const uint32_t d = r < 4 ? r : 3;
                uint32_t *rm = run + rctx * AR_RUN_STRIDE;
                ar_load (rm, c);
                ar_encode_sym (rm, 4, c, rc, d, lane, stale);
                while (rc.range < AR_TOP) { rc.range <<= 8; ar_shift_low (rc); }
                r -= d;
                if (rctx == last) rctx = 256; else rctx += (rctx < 257);
                if (d == 3 && r == 0) {
                    rm = run + rctx * AR_RUN_STRIDE;
                    ar_load (rm, c);
                    ar_encode_sym (rm, 4, c, rc, 0, lane, stale);
                    while (rc.range < AR_TOP) { rc.range <<= 8; ar_shift_low (rc); }
                }
            } while (r);
        }
    }
    for (int i = 0; i < 5; i++) ar_shift_low (rc);                          // RC_FinishEncode
    return (uint32_t)(rc.out - out);
}

} // namespace gzb
