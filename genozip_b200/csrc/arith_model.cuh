// arith_model.cuh — the adaptive frequency model of the htscodecs arithmetic coder (reference SIMPLE_MODEL,
// src/htscodecs/c_simple_model.h:77-179) in a layout made for 16-byte loads:
//   word 0      TotFreq
//   word 3      sentinel (Freq = MAX_FREQ: never swapped past, :98-99)
//   word 4..    live entries  freq | symbol << 16  (approximately sorted by frequency)
//   then one zero word (terminates normalise, :101) and padding to a multiple of 4 words.
// The symbol being coded is almost always among the first four entries, which arrive in one load.
#pragma once
#include <stdint.h>

namespace gzb {

#define AR_MAXF  65519u          // MAX_FREQ = (1<<16)-17 (c_simple_model.h:70)
#define AR_STEP  16u             // STEP (:73)

__host__ __device__ __forceinline__ uint32_t ar_stride (uint32_t maxs) { return (maxs + 5 + 3) & ~3u; }
constexpr uint32_t AR_RUN_STRIDE = 12;           // run-length models: 4 live symbols (MAX_RUN, arith_dynamic.c:383)

__device__ __forceinline__ void ar_model_init (uint32_t *m, uint32_t maxs)                 // :85-103
{
    m[0] = maxs; m[1] = 0; m[2] = 0; m[3] = AR_MAXF | 0xffff0000u;
    for (uint32_t i = 0; i < maxs; i++) m[4 + i] = 1u | (i << 16);
    const uint32_t st = ar_stride (maxs);
    for (uint32_t i = 4 + maxs; i < st; i++) m[i] = 0xffff0000u;          // Freq 0 (terminates normalise) and a symbol that never matches
}

// bump the coded entry i (holding e): Freq += STEP, halve everything past MAX_FREQ, one bubble step towards the front
__device__ __forceinline__ void ar_model_bump (uint32_t *m, uint32_t i, uint32_t e, uint32_t tot)   // :131-145 / :164-178
{
    uint32_t f = (e & 0xffffu) + AR_STEP;
    tot += AR_STEP;
    if (tot > AR_MAXF) {                                                                     // normalize (:106-116)
        m[i] = (e & 0xffff0000u) | f;
        tot = 0;
        for (uint32_t j = 4; (m[j] & 0xffffu); j++) { uint32_t g = m[j] & 0xffffu; g -= g >> 1; m[j] = (m[j] & 0xffff0000u) | g; tot += g; }
        f = m[i] & 0xffffu;
    }
    m[0] = tot;
    const uint32_t prev = m[i - 1];
    if (f > (prev & 0xffffu)) { m[i - 1] = (e & 0xffff0000u) | f; m[i] = prev; }
    else m[i] = (e & 0xffff0000u) | f;
}

// locate `sym`: returns its index, the entry in e and the cumulative frequency before it in acc
__device__ __forceinline__ uint32_t ar_find_sym (const uint32_t *m, uint32_t sym, uint32_t &e, uint32_t &acc)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(m + 4);
    if ((v.x >> 16) == sym) { e = v.x; acc = 0; return 4; }
    if ((v.y >> 16) == sym) { e = v.y; acc = v.x & 0xffffu; return 5; }
    if ((v.z >> 16) == sym) { e = v.z; acc = (v.x & 0xffffu) + (v.y & 0xffffu); return 6; }
    acc = (v.x & 0xffffu) + (v.y & 0xffffu) + (v.z & 0xffffu);
    if ((v.w >> 16) == sym) { e = v.w; return 7; }
    acc += v.w & 0xffffu;
    uint32_t i = 8; e = m[8];
    while ((e >> 16) != sym) { acc += e & 0xffffu; e = m[++i]; }
    return i;
}

// locate the entry whose cumulative range contains `freq`; returns 0 when the model is exhausted (corrupt input)
__device__ __forceinline__ uint32_t ar_find_freq (const uint32_t *m, uint32_t freq, uint32_t &e, uint32_t &acc)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(m + 4);
    uint32_t a = v.x & 0xffffu;
    if (a > freq) { e = v.x; acc = 0; return 4; }
    uint32_t b = a + (v.y & 0xffffu);
    if (b > freq) { e = v.y; acc = a; return 5; }
    uint32_t c = b + (v.z & 0xffffu);
    if (c > freq) { e = v.z; acc = b; return 6; }
    uint32_t d = c + (v.w & 0xffffu);
    if (d > freq) { e = v.w; acc = c; return 7; }
    uint32_t i = 8; acc = d;
    for (;;) {
        e = m[i];
        if (!e) return 0;
        if (acc + (e & 0xffffu) > freq) return i;
        acc += e & 0xffffu; i++;
    }
}

} // namespace gzb
