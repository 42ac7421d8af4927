// arith_chain.cu — the adaptive arithmetic coder's serial loop, one leaf per WARP.
//
// The bitstream fixes one dependency chain per leaf (reference arith_dynamic.c:92-226, 387-608; range coder
// c_range_coder.h:46-126; model c_simple_model.h:123-179), so the chain itself cannot be split.  What the other 31 lanes
// of the warp can do is the model's linear search — the reference walks the frequency-sorted list entry by entry
// (:127-130, :156-158) — here a warp looks at 32 entries per load: ballot for the match, a warp reduction for the
// cumulative frequency.  All lanes carry the range-coder state redundantly (uniform SIMT execution costs nothing);
// lane 0 performs the byte I/O and the model write-back.
#include "gzb_internal.cuh"
#include "hts_enc.cuh"
#include "arith_model.cuh"

namespace gzb {

// 1/x rounded safely DOWN: MUFU.RCP (<= 1 ulp, one instruction) scaled by (1 - 5e-7); relative deficit < 7e-7
__device__ __forceinline__ float rcp_below (float x)
{
    float r;
    asm ("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return __fmul_rz (r, 0.9999995f);
}

// exact 32-bit division by a divisor < 2^17: float estimates that never exceed the truth + small corrections
__device__ __forceinline__ uint32_t div_small (uint32_t a, uint32_t d)
{
    const float rd = rcp_below (__uint2float_ru (d));
    uint32_t q = __float2uint_rz (__fmul_rz (__uint2float_rz (a), rd));
    uint32_t r = a - q * d;
    const uint32_t q2 = __float2uint_rz (__fmul_rz (__uint2float_rz (r), rd));
    q += q2; r -= q2 * d;
    while (r >= d) { q++; r -= d; }
    return q;
}

// exact 32-bit division whose QUOTIENT is small (< 2^17): RC_GetFreq's code / range (c_range_coder.h:111-114)
__device__ __forceinline__ uint32_t div_smallq (uint32_t a, uint32_t d)
{
    const float rd = rcp_below (__uint2float_ru (d));
    uint32_t q = __float2uint_rz (__fmul_rz (__uint2float_rz (a), rd));
    uint32_t r = a - q * d;
    while (r >= d) { q++; r -= d; }
    return q;
}

// ---- warp-wide model access -------------------------------------------------------------------------------------
// find `sym`: returns the model word index of its entry, its value in e and the cumulative frequency before it in acc
__device__ __forceinline__ uint32_t warp_find_sym (const uint32_t *m, uint32_t maxs, uint32_t sym, int lane, uint32_t &e, uint32_t &acc)
{
    acc = 0;
    for (uint32_t base = 0; ; base += 32) {
        const uint32_t j = base + lane;
        const uint32_t v = j < maxs ? m[4 + j] : 0;
        const uint32_t hit = __ballot_sync (0xffffffffu, j < maxs && (v >> 16) == sym);
        if (hit) {
            const int w = __ffs (hit) - 1;
            acc += __reduce_add_sync (0xffffffffu, lane < w ? (v & 0xffffu) : 0u);
            e = __shfl_sync (0xffffffffu, v, w);
            return 4 + base + w;
        }
        acc += __reduce_add_sync (0xffffffffu, v & 0xffffu);
        if (base + 32 >= maxs) { e = 0; return 0; }                       // cannot happen for a symbol < maxs
    }
}

// find the entry whose cumulative range contains freq; 0 = exhausted (corrupt stream)
__device__ __forceinline__ uint32_t warp_find_freq (const uint32_t *m, uint32_t maxs, uint32_t freq, int lane, uint32_t &e, uint32_t &acc)
{
    acc = 0;
    for (uint32_t base = 0; ; base += 32) {
        const uint32_t j = base + lane;
        const uint32_t v = j < maxs ? m[4 + j] : 0;
        uint32_t c = v & 0xffffu;                                           // inclusive prefix over the 32 entries
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, c, o); if (lane >= o) c += t; }
        const uint32_t hit = __ballot_sync (0xffffffffu, j < maxs && acc + c > freq);
        if (hit) {
            const int w = __ffs (hit) - 1;
            e = __shfl_sync (0xffffffffu, v, w);
            acc += __shfl_sync (0xffffffffu, c, w) - (e & 0xffffu);
            return 4 + base + w;
        }
        acc += __shfl_sync (0xffffffffu, c, 31);
        if (base + 32 >= maxs) { e = 0; return 0; }
    }
}

// Freq += STEP, halve past MAX_FREQ, one bubble step (c_simple_model.h:131-145).  Lane 0 writes; halving is warp-wide.
__device__ __forceinline__ void warp_model_bump (uint32_t *m, uint32_t maxs, uint32_t i, uint32_t e, uint32_t tot, int lane)
{
    uint32_t f = (e & 0xffffu) + AR_STEP;
    tot += AR_STEP;
    if (tot > AR_MAXF) {                                                    // normalize (:106-116), warp-uniform branch
        m[i] = (e & 0xffff0000u) | f;
        __syncwarp ();
        uint32_t sum = 0;
        for (uint32_t base = 0; base < maxs; base += 32) {
            const uint32_t j = base + lane;
            if (j < maxs) { uint32_t v = m[4 + j]; uint32_t g = v & 0xffffu; g -= g >> 1; m[4 + j] = (v & 0xffff0000u) | g; sum += g; }
        }
        tot = __reduce_add_sync (0xffffffffu, sum);
        __syncwarp ();
        f = m[i] & 0xffffu;
    }
    // every lane writes the same words (one merged transaction); each lane then reads back its own stores, so the common
    // path needs no warp synchronisation
    const uint32_t prev = m[i - 1];
    m[0] = tot;
    if (f > (prev & 0xffffu)) { m[i - 1] = (e & 0xffff0000u) | f; m[i] = prev; }
    else m[i] = (e & 0xffff0000u) | f;
}

// ---- encoder ---------------------------------------------------------------------------------------------------
struct RCEnc { uint32_t low, range, ffnum, cache, carry; uint8_t *out; };

__device__ __forceinline__ void rc_shift_low (RCEnc &rc, int lane)          // c_range_coder.h:70-88
{
    if (rc.low < (255u << 24) || rc.carry) {
        *rc.out = (uint8_t)(rc.cache + rc.carry);                           // all lanes store the same byte: one merged transaction
        for (uint32_t i = 0; i < rc.ffnum; i++) rc.out[1 + i] = (uint8_t)(rc.carry - 1);
        rc.out += 1 + rc.ffnum; rc.ffnum = 0;
        rc.cache = rc.low >> 24;
        rc.carry = 0;
    }
    else rc.ffnum++;
    rc.low <<= 8;
}

__device__ __forceinline__ void warp_encode (uint32_t *m, uint32_t maxs, RCEnc &rc, uint32_t sym, int lane)   // :123-146 + RC_Encode :97-109
{
    uint32_t e, acc, i;
    const uint32_t tot = m[0];
    const uint4 v = *reinterpret_cast<const uint4 *>(m + 4);              // the model is approximately sorted by frequency: the
    if      ((v.x >> 16) == sym) { e = v.x; acc = 0; i = 4; }               // symbol is almost always among the first four entries
    else if ((v.y >> 16) == sym) { e = v.y; acc = v.x & 0xffffu; i = 5; }   // (padding entries carry symbol 0xffff: they never match)
    else if ((v.z >> 16) == sym) { e = v.z; acc = (v.x & 0xffffu) + (v.y & 0xffffu); i = 6; }
    else if ((v.w >> 16) == sym) { e = v.w; acc = (v.x & 0xffffu) + (v.y & 0xffffu) + (v.z & 0xffffu); i = 7; }
    else i = warp_find_sym (m, maxs, sym, lane, e, acc);
    const uint32_t before = rc.low;
    rc.range = div_small (rc.range, tot);
    rc.low   += acc * rc.range;
    rc.range *= e & 0xffffu;
    rc.carry += rc.low < before;
    warp_model_bump (m, maxs, i, e, tot, lane);
    while (rc.range < (1u << 24)) { rc.range <<= 8; rc_shift_low (rc, lane); }
}

__global__ void __launch_bounds__(128) k_arith_encode (const EncLeaf *leaves, EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list)
{
    const int lane = threadIdx.x & 31;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= n_list) return;
    const uint32_t li = list[slot];
    const EncLeaf &L = leaves[li];
    EncLeafDyn &D = dyn[li];
    const uint32_t n = D.eff_n, maxs = D.nsym, stride = ar_stride (maxs);
    const uint8_t * __restrict__ in = D.eff_in;
    const bool o1 = D.eff_order, rle = (D.hdr[0] & F_RLE) != 0;
    uint32_t *lit = D.models, *run = lit + (o1 ? 256 : 1) * stride;
    uint8_t *out = L.outbuf;
    if (!lit) return;
    __builtin_assume (__isGlobal (lit)); __builtin_assume (__isGlobal (out)); __builtin_assume (__isGlobal (in));
    if (lane == 0) out[0] = (uint8_t)maxs;                                 // arith_dynamic.c:105-110 (256 wraps to 0)
    RCEnc rc; rc.low = 0; rc.range = 0xffffffffu; rc.ffnum = 0; rc.cache = 0; rc.carry = 0; rc.out = out + 1;
    // A body that reaches the input length is discarded for a raw copy (arith_dynamic.c:847-852), so encoding stops
    // as soon as that is certain; this also bounds the scratch a hostile (expanding) input can touch.
    const uint8_t *limit = out + n + 8;
    bool expanded = false;
    uint32_t last = 0;
    for (uint32_t i = 0; i < n; ) {
        if (rc.out + rc.ffnum > limit) { expanded = true; break; }
        const uint32_t s = __ldg (in + i);
        warp_encode (lit + (o1 ? last : 0) * stride, maxs, rc, s, lane);
        last = s; i++;
        if (!rle) continue;
        uint32_t r = 0;                                                   // :413-438 run length in base-4 digits
        while (i < n && __ldg (in + i) == last) { r++; i++; }
        uint32_t rctx = last;
        do {
            const uint32_t c = r < 4 ? r : 3;
            warp_encode (run + rctx * AR_RUN_STRIDE, 4, rc, c, lane);
            r -= c;
            if (rctx == last) rctx = 256; else rctx += (rctx < 257);
            if (c == 3 && r == 0) warp_encode (run + rctx * AR_RUN_STRIDE, 4, rc, 0, lane);
        } while (r);
    }
    if (!expanded) for (int i = 0; i < 5; i++) rc_shift_low (rc, lane);   // RC_FinishEncode
    if (lane == 0) {
        D.tab_len = expanded ? n + 1 : (uint32_t)(rc.out - out);          // whole body at the front of outbuf
        D.payload_len = 0;
    }
}

void launch_arith_encode (EncPlanDev &P, cudaStream_t st)
{
    k_arith_encode<<<(P.n_arith + 3) / 4, 128, 0, st>>>(P.leaves, P.dyn, P.arith_list, P.n_arith);
}

// ---- decoder ---------------------------------------------------------------------------------------------------
struct RCDec { uint32_t code, range; const uint8_t *in, *end; };

__device__ __forceinline__ uint32_t warp_decode (uint32_t *m, uint32_t maxs, RCDec &rc, int lane)   // c_simple_model.h:148-179
{
    const uint32_t tot = m[0];
    uint32_t freq = 0;
    if (tot && rc.range >= tot) {                                           // RC_GetFreq (c_range_coder.h:111-114)
        rc.range = div_small (rc.range, tot);
        freq = (rc.code >> 17) >= rc.range ? rc.code / rc.range : div_smallq (rc.code, rc.range);   // quotient < 2^17 for any valid stream
    }
    if (freq > AR_MAXF) return 0;
    uint32_t e, acc, i;
    const uint4 v = *reinterpret_cast<const uint4 *>(m + 4);
    const uint32_t c0 = v.x & 0xffffu, c1 = c0 + (v.y & 0xffffu), c2 = c1 + (v.z & 0xffffu), c3 = c2 + (v.w & 0xffffu);
    if      (c0 > freq) { e = v.x; acc = 0;  i = 4; }                        // padding entries have Freq 0: they never extend the range
    else if (c1 > freq) { e = v.y; acc = c0; i = 5; }
    else if (c2 > freq) { e = v.z; acc = c1; i = 6; }
    else if (c3 > freq) { e = v.w; acc = c2; i = 7; }
    else i = warp_find_freq (m, maxs, freq, lane, e, acc);
    if (!i) return 0;
    rc.code  -= acc * rc.range;
    rc.range *= e & 0xffffu;
    warp_model_bump (m, maxs, i, e, tot, lane);
    while (rc.range < (1u << 24)) {                                         // RC_Decode (:116-126)
        if (rc.in >= rc.end) break;
        rc.code = (rc.code << 8) + __ldg (rc.in++);
        rc.range <<= 8;
    }
    return e >> 16;
}

__global__ void __launch_bounds__(128) k_arith_decode (DecLeaf *leaves, const uint32_t *list, uint32_t n_list)
{
    const int lane = threadIdx.x & 31;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= n_list) return;
    DecLeaf &L = leaves[list[slot]];
    if (!L.valid || L.err || L.cat || !L.body_ulen || !L.models) return;
    const uint32_t n = L.body_ulen, maxs = L.nsym, stride = ar_stride (maxs);
    const bool o1 = L.order == 1, rle = L.rle;
    uint32_t *lit = L.models, *run = lit + (o1 ? 256 : 1) * stride;
    uint8_t *out = L.dst;
    __builtin_assume (__isGlobal (lit)); __builtin_assume (__isGlobal (out)); __builtin_assume (__isGlobal (L.body));
    RCDec rc; rc.range = 0xffffffffu; rc.code = 0; rc.in = L.body + 1; rc.end = L.body + L.body_len;
    if (rc.in + 5 > rc.end) rc.in = rc.end;                               // RC_StartDecode (c_range_coder.h:57-68)
    else for (int i = 0; i < 5; i++) rc.code = (rc.code << 8) | *rc.in++;
    // Output bytes are gathered in a 32-bit window and written one aligned word at a time (a byte store per symbol from
    // hundreds of concurrent leaves is what the L2 write path chokes on); head and tail bytes go out singly.
    uint32_t last = 0, win = 0;
    const uint32_t head_end = (uint32_t)((4 - ((uintptr_t)out & 3)) & 3);   // bytes before the first aligned word are stored singly
    #define PUT_BYTE(idx, b) do { win = (win >> 8) | ((uint32_t)(b) << 24); \
        if (lane == 0) { const uintptr_t A = reinterpret_cast<uintptr_t>(out + (idx)); \
                         if ((A & 3) == 3 && (idx) >= 3) *reinterpret_cast<uint32_t *>(A - 3) = win; \
                         else if ((idx) < head_end) out[idx] = (uint8_t)(b); } } while (0)
    uint32_t i = 0;
    for (; i < n; i++) {
        const uint32_t s = warp_decode (lit + (o1 ? last : 0) * stride, maxs, rc, lane);
        PUT_BYTE (i, s);
        last = s;
        if (!rle) continue;
        uint32_t r = 0, part, rctx = last;                                // arith_dynamic.c:473-482 / :591-599
        do {
            part = warp_decode (run + rctx * AR_RUN_STRIDE, 4, rc, lane);
            if (rctx == last) rctx = 256; else rctx += (rctx < 257);
            r += part;
        } while (part == 3 && r < n);
        while (r-- && i + 1 < n) { ++i; PUT_BYTE (i, last); }
    }
    #undef PUT_BYTE
    if (lane == 0) {                                                      // tail: bytes after the last aligned word boundary
        const uint32_t tail = (uint32_t)(((uintptr_t)(out + n)) & 3);
        for (uint32_t t = 0; t < tail && t < n; t++) out[n - 1 - t] = (uint8_t)(win >> (24 - 8 * t));
    }
}

void launch_arith_decode (DecPlanDev &P, cudaStream_t st)
{
    k_arith_decode<<<(P.n_arith + 3) / 4, 128, 0, st>>>(P.leaves, P.arith_list, P.n_arith);
}

} // namespace gzb
