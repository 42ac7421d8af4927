// arith_chain.cu — the adaptive arithmetic coder's serial loop, one leaf per WARP.
//
// The bitstream fixes one dependency chain per leaf (reference arith_dynamic.c:92-226, 387-608; range coder
// c_range_coder.h:46-126; model c_simple_model.h:123-179), so the chain itself cannot be split and a leaf advances at
// the latency of one warp's dependent instruction stream.  arith_model.cuh keeps that stream short: the current
// context's model head in registers, one float multiply + integer correction per division, no code/range division on
// the common path, warp-wide search (8 entries per lane) only beyond the first four entries.
#include "gzb_internal.cuh"
#include "hts_enc.cuh"
#include "arith_model.cuh"
#include "arith_o0.cuh"
#include <stdlib.h>
#include <string.h>

namespace gzb {

// Order-0 leaves without the run-length models (the byte planes of the STRIPE codecs, mostly) keep their single model in
// SHARED memory and search it with per-lane running sums: arith_o0.cuh.

// Two kernels per direction, launched over the same leaf list on two streams: each warp looks at its leaf's class and
// leaves at once if it belongs to the other kernel (separate kernels = separate register allocation for the two loops).
__device__ __forceinline__ bool ar_is_o0_class (bool o1, bool rle) { return !o1 && !rle; }
__device__ __forceinline__ bool ar_is_long_leaf (bool o1, bool rle, uint32_t n, uint32_t long_min);

template <bool O0CLASS>
__global__ void __launch_bounds__(128) k_arith_encode_t (const EncLeaf *leaves, EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list, uint32_t *queue)
{
    const int lane = threadIdx.x & 31;
    __shared__ __align__(16) uint8_t s_model[O0CLASS ? 4 : 1][O0CLASS ? AR0_SMEM_BYTES : 16];
    for (;;) {
        const uint32_t slot = queue_take (queue, lane);                     // persistent warps: the next leaf of the longest-first list
        if (slot >= n_list) return;
        const uint32_t li = list[slot];
        const EncLeaf &L = leaves[li];
        EncLeafDyn &D = dyn[li];
        const uint32_t n = D.eff_n, maxs = D.nsym;
        const uint8_t * __restrict__ in = D.eff_in;
        const bool o1 = D.eff_order, rle = (D.hdr[0] & F_RLE) != 0;
        uint32_t *lit = D.models;
        uint8_t *out = L.outbuf;
        if (!lit || ar_is_o0_class (o1, rle) != O0CLASS || D.split_pos) continue;   // (split_pos: the leaf goes through arith_split.cu)
        __builtin_assume (__isGlobal (lit)); __builtin_assume (__isGlobal (out)); __builtin_assume (__isGlobal (in));
        uint32_t len;
        if (O0CLASS) {
            uint8_t *sm = s_model[threadIdx.x >> 5];
            __syncwarp ();
            len = ar0_encode_leaf (reinterpret_cast<uint32_t *>(sm), sm + AR0_E_WORDS * 4, maxs, in, n, out, lane);
        }
        else len = o1 ? ar_encode_leaf<true> (lit, maxs, rle, in, n, out, lane) : ar_encode_leaf<false> (lit, maxs, rle, in, n, out, lane);
        if (lane == 0) {
            D.tab_len = len;                                               // whole body at the front of outbuf (n + 1 = expanded)
            D.payload_len = 0;
        }
    }
}

// CTAs per SM of the persistent chain kernels and the other switches of the chain phase, from the environment (A/B runs):
//   GZB_AR_CTAS     general arithmetic kernel (order 1 / RLE leaves), 4 warps per CTA       default 2
//                   (2 / 4 / 8 / 16 per SM measure the same within run-to-run noise; 2 leaves registers for the kernels beside them)
//   GZB_AR0_CTAS    order-0 arithmetic kernel                                               default 2
//   GZB_AR_RUN4     the decoder tries four run steps at once: 0 never, 1 all four or nothing, 2 the longest valid prefix   default 2
//   GZB_AR_SPLIT_STREAM  the split encoder runs on its own stream beside the general kernel default 1
//   GZB_AR_LONG_MIN order-1 leaves of at least this many symbols are decoded by k_arith_decode_long (off = none)   default off
//                   (measured, 768 VBlocks: alone the mirror takes the longest leaf from 332 to 305 / 286 ms (16 / 32 entries); beside the
//                   rANS and order-0 kernels its shared memory costs them more residency than it gains: piz 445 -> 539 ms)
//   GZB_AR_LONG_ENT entries per context it mirrors in shared memory: 16 or 32                  default 16
const ChainTune &chain_tune ()                                               // (read at every call: a sweep inside one process changes the variables between batches)
{
    static thread_local ChainTune c;
    auto geti = [] (const char *name, int dflt, int lo, int hi) {
        const char *v = getenv (name);
        if (!v || !*v) return dflt;
        const int x = atoi (v);
        return x < lo ? lo : x > hi ? hi : x;
    };
    c.arith_ctas = geti ("GZB_AR_CTAS", 2, 1, 16);
    c.arith_o0_ctas = geti ("GZB_AR0_CTAS", 2, 1, 16);
    c.run4 = geti ("GZB_AR_RUN4", 2, 0, 2);
    c.split_stream = geti ("GZB_AR_SPLIT_STREAM", 1, 0, 1);
    c.long_ent = geti ("GZB_AR_LONG_ENT", 16, 16, 32) >= 32 ? 32 : 16;
    { const char *v = getenv ("GZB_AR_LONG_MIN"); c.long_min = !v || !*v || !strcmp (v, "off") ? 0xffffffffu : (uint32_t)strtoul (v, nullptr, 10); }
    return c;
}

static uint32_t persistent_grid (uint32_t n_list, int sm_count, int ctas_per_sm)
{
    const uint32_t want = (n_list + 3) / 4, cap = (uint32_t)(sm_count > 0 ? sm_count : 148) * (uint32_t)ctas_per_sm;
    return want < cap ? want : cap;
}

void launch_arith_encode (EncPlanDev &P, cudaStream_t st)
{
    k_arith_encode_t<false><<<persistent_grid (P.n_arith, P.sm_count, chain_tune ().arith_ctas), 128, 0, st>>>(P.leaves, P.dyn, P.arith_list, P.n_arith, P.queue + Q_ARITH);
}
void launch_arith_encode_o0 (EncPlanDev &P, cudaStream_t st)
{
    k_arith_encode_t<true><<<persistent_grid (P.n_arith, P.sm_count, chain_tune ().arith_o0_ctas), 128, 0, st>>>(P.leaves, P.dyn, P.arith_list, P.n_arith, P.queue + Q_ARITH_O0);
}

template <bool O0CLASS>
__global__ void __launch_bounds__(128) k_arith_decode_t (DecLeaf *leaves, const uint32_t *list, uint32_t n_list, uint32_t *queue, uint32_t run4,
                                                         uint32_t n_long_cand, uint32_t long_min)
{
    const int lane = threadIdx.x & 31;
    __shared__ __align__(16) uint8_t s_model[O0CLASS ? 4 : 1][O0CLASS ? AR0_SMEM_BYTES : 16];
    for (;;) {
        const uint32_t slot = queue_take (queue, lane);
        if (slot >= n_list) return;
        DecLeaf &L = leaves[list[slot]];
        if (!L.valid || L.err || L.cat || !L.body_ulen || !L.models) continue;
        const uint32_t n = L.body_ulen, maxs = L.nsym;
        const bool o1 = L.order == 1, rle = L.rle;
        if (ar_is_o0_class (o1, rle) != O0CLASS) continue;
        if (!O0CLASS && slot < n_long_cand && ar_is_long_leaf (o1, rle, n, long_min)) continue;   // k_arith_decode_long's
        uint32_t *lit = L.models;
        uint8_t *out = L.dst;
        const uint8_t * __restrict__ body = L.body;
        __builtin_assume (__isGlobal (lit)); __builtin_assume (__isGlobal (out)); __builtin_assume (__isGlobal (body));
        if (O0CLASS) {
            uint32_t *E = reinterpret_cast<uint32_t *>(s_model[threadIdx.x >> 5]);
            __syncwarp ();
            Ar0 a; ar0_init (E, nullptr, maxs, lane, a);
            ArDec rc; ar_dec_start (rc, body, L.body_len);
            ArOut o; ar_out_init (o, out);
            const uint32_t done = ar0_decode_run (E, maxs, a, rc, o, n, lane);
            if (done < n) {                                                     // corrupt / truncated stream: finish exactly like the reference, through memory
                ar0_export (E, a, lit, maxs, lane);
                ar_decode_tail<false> (lit, maxs, rc, o, done, n, 0, lane);
            }
            ar_out_flush (o);
        }
        else if (o1) ar_decode_leaf<true> (lit, maxs, rle, body, L.body_len, out, n, lane, run4);
        else         ar_decode_leaf<false> (lit, maxs, rle, body, L.body_len, out, n, lane, run4);
    }
}

// ------------------------------------------------------------------------------------------------ long order-1 leaves (decode)
// In a batch the models of all leaves together (264 KB per order-1 leaf) are several times the L2, so every context switch of a
// long leaf — a fifth of its symbols — waits for DRAM, and the launch lasts as long as its longest leaf.  This kernel keeps a
// MIRROR of every context's model head and first ENT entries in shared memory (one warp = one leaf = one CTA, 256 x (4 + ENT)
// words): reads of that part come from the mirror, every store goes to both, so global memory stays complete and authoritative
// (the halving, the search beyond the mirror and the reference's behaviour on damaged streams run on it unchanged, the mirror of
// the context is then refreshed).  A context switch is a shared-memory load; a symbol among the first ENT entries is found by
// one scan of the lanes' frequencies.
__device__ __forceinline__ bool ar_is_long_leaf (bool o1, bool rle, uint32_t n, uint32_t long_min) { return o1 && !rle && n >= long_min; }

template <int ENT>
__device__ __forceinline__ void arl_mirror_ctx (const uint32_t *m, uint32_t *s, uint32_t maxs, int lane)
{
    __syncwarp ();
    for (uint32_t w = lane; w < 4 + ENT; w += 32) s[w] = (w < 4 + maxs) ? m[w] : 0xffff0000u;
    __syncwarp ();
}

#define ARL_ST(IDX, V) { m[4 + (IDX)] = (V); if ((IDX) < (uint32_t)ENT) s[4 + (IDX)] = (V); }
#define ARL_BUMP_CASE(K, EK, EPREV)                                                                             \
    {                                                                                                           \
        if (c.tot + AR_STEP > AR_MAXF) { ar_update_mem (m, maxs, K, EK, c.tot, lane); stale = true; }           \
        else {                                                                                                  \
            c.tot += AR_STEP; EK += AR_STEP;                                                                    \
            c.rtot = ar_rcp_below (c.tot);                                                                      \
            if (K > 0 && (EK & 0xffffu) > (EPREV & 0xffffu)) { const uint32_t t_ = EK; EK = EPREV; EPREV = t_; } \
            ar_flush (m, c); ar_flush (s, c);                                                                   \
        }                                                                                                       \
    }
#define ARL_BUMP_DEEP(P, E, PREV)                                                                               \
    {                                                                                                           \
        if (c.tot + AR_STEP > AR_MAXF) { ar_update_mem (m, maxs, P, E, c.tot, lane); stale = true; }            \
        else {                                                                                                  \
            const uint32_t en_ = (E) + AR_STEP;                                                                 \
            c.tot += AR_STEP; c.rtot = ar_rcp_below (c.tot);                                                    \
            ar_store_head (m, c.tot, c.rtot); ar_store_head (s, c.tot, c.rtot);                                 \
            if ((en_ & 0xffffu) > ((PREV) & 0xffffu)) { ARL_ST ((P) - 1, en_) ARL_ST ((P), (PREV)) if ((P) == 4) c.e3 = en_; } \
            else ARL_ST ((P), en_)                                                                              \
        }                                                                                                       \
    }

// one symbol that is not a run step (ar_decode_sym<false> on the mirrored model); r = range / TotFreq
template <int ENT>
__device__ __forceinline__ uint32_t arl_decode_sym (uint32_t *m, uint32_t *s, uint32_t maxs, ArCache &c, ArDec &rc, int lane, bool &stale, bool &anomaly, uint32_t r)
{
    uint32_t sym;
    const uint32_t f0 = c.e0 & 0xffffu;
    const uint32_t t1 = f0 * r;
    if (rc.code < t1) { rc.range = t1; sym = c.e0 >> 16; ARL_BUMP_CASE (0, c.e0, c.e0) return sym; }
    const uint32_t f1 = c.e1 & 0xffffu, f2 = c.e2 & 0xffffu, f3 = c.e3 & 0xffffu;
    const uint32_t t2 = t1 + f1 * r, t3 = t2 + f2 * r, t4 = t3 + f3 * r;
    if (rc.code < t2)      { rc.code -= t1; rc.range = f1 * r; sym = c.e1 >> 16; ARL_BUMP_CASE (1, c.e1, c.e0) }
    else if (rc.code < t3) { rc.code -= t2; rc.range = f2 * r; sym = c.e2 >> 16; ARL_BUMP_CASE (2, c.e2, c.e1) }
    else if (rc.code < t4) { rc.code -= t3; rc.range = f3 * r; sym = c.e3 >> 16; ARL_BUMP_CASE (3, c.e3, c.e2) }
    else {
        // among the mirrored entries: lane l looks at entry l; the first lane whose inclusive cumulative frequency times r exceeds
        // code owns the symbol (c_simple_model.h:156 without the division, as in ar_find_code)
        const uint32_t en = lane < ENT ? s[4 + lane] : 0xffff0000u;
        const uint32_t f = en & 0xffffu;
        uint32_t inc = f;
        #pragma unroll
        for (int o = 1; o < ENT; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const uint32_t ball = __ballot_sync (0xffffffffu, lane < ENT && inc * r > rc.code);
        AR_READS_DONE ();
        if (ball) {
            const uint32_t p = __ffs (ball) - 1;                            // >= 4: the first four thresholds are t1 .. t4
            const uint32_t acc = __shfl_sync (0xffffffffu, inc - f, p), e = __shfl_sync (0xffffffffu, en, p);
            const uint32_t prev = __shfl_sync (0xffffffffu, en, (p + 31) & 31);
            rc.code -= acc * r; rc.range = (e & 0xffffu) * r;
            sym = e >> 16;
            ARL_BUMP_DEEP (p, e, prev)
        }
        else {                                                              // beyond the mirror: the complete model in global memory
            const ArHit h = ar_find_code (m, maxs, rc.code, r, lane);
            const uint32_t p = h.p, acc = h.acc, e = h.e, prev = h.prev;
            if (p >= maxs) { rc.range = r; anomaly = true; return 0; }
            rc.code -= acc * r; rc.range = (e & 0xffffu) * r;
            sym = e >> 16;
            ARL_BUMP_DEEP (p, e, prev)
        }
    }
    return sym;
}

template <int ENT>
__global__ void __launch_bounds__(32) k_arith_decode_long (DecLeaf *leaves, const uint32_t *list, uint32_t n_cand, uint32_t long_min, uint32_t *queue, uint32_t run4)
{
    extern __shared__ __align__(16) uint32_t s_mirror[];                    // [256][4 + ENT]
    constexpr uint32_t W = 4 + ENT;
    const int lane = threadIdx.x & 31;
    for (;;) {
        const uint32_t slot = queue_take (queue, lane);
        if (slot >= n_cand) return;
        DecLeaf &L = leaves[list[slot]];
        if (!L.valid || L.err || L.cat || !L.body_ulen || !L.models) continue;
        const uint32_t n = L.body_ulen, maxs = L.nsym;
        if (!ar_is_long_leaf (L.order == 1, L.rle, n, long_min)) continue;
        uint32_t *lit = L.models;
        uint8_t *out = L.dst;
        const uint8_t * __restrict__ body = L.body;
        __builtin_assume (__isGlobal (lit)); __builtin_assume (__isGlobal (out)); __builtin_assume (__isGlobal (body));
        const uint32_t stride = ar_stride (maxs);
        for (uint32_t cx = 0; cx < maxs; cx++) arl_mirror_ctx<ENT> (lit + cx * stride, s_mirror + cx * W, maxs, lane);
        ArDec rc; ar_dec_start (rc, body, L.body_len);
        ArOut o; ar_out_init (o, out);
        ArCache c;
        uint32_t ctx = 0, i = 0;
        uint32_t *m = lit, *s = s_mirror;
        ar_load (s, c);
        bool ok = true, dirty = false;
        bool selfloop = (c.e0 >> 16) == ctx;
        uint32_t skip4 = 0, fail4 = 0;
        while (i < n && ok) {                                               // the loop of ar_decode_leaf<true>, model reads from the mirror
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
            if (dirty) { ar_flush (m, c); ar_flush (s, c); dirty = false; }
            bool stale = false, anomaly = false;
            const uint32_t sy = arl_decode_sym<ENT> (m, s, maxs, c, rc, lane, stale, anomaly, r);
            if (anomaly) ok = false;
            else if (rc.range < AR_TOP) ok = ar_dec_renorm (rc);
            ar_out_put (o, sy);
            i++;
            if (stale) arl_mirror_ctx<ENT> (m, s, maxs, lane);              // the halving went through global memory
            if (sy != ctx) { ctx = sy; m = lit + sy * stride; s = s_mirror + sy * W; ar_load (s, c); }
            else if (stale) ar_load (s, c);
            selfloop = (c.e0 >> 16) == ctx;
        }
        if (dirty) ar_flush (m, c);
        ar_decode_tail<true> (lit, maxs, rc, o, i, n, ctx, lane);
        ar_out_flush (o);
        __syncwarp ();
    }
}

void launch_arith_decode_long (DecPlanDev &P, cudaStream_t st)
{
    if (!P.n_long_cand) return;
    const int ent = chain_tune ().long_ent;
    const uint32_t smem = 256u * (4u + (uint32_t)ent) * 4u;
    const uint32_t per_sm = (uint32_t)((227u * 1024u) / (smem + 1024u));
    const uint32_t cap = (uint32_t)(P.sm_count > 0 ? P.sm_count : 148) * per_sm, grid = P.n_long_cand < cap ? P.n_long_cand : cap;
    if (ent == 32) k_arith_decode_long<32><<<grid, 32, smem, st>>>(P.leaves, P.arith_list, P.n_long_cand, P.long_min, P.queue + Q_ARITH_LONG, (uint32_t)chain_tune ().run4);
    else           k_arith_decode_long<16><<<grid, 32, smem, st>>>(P.leaves, P.arith_list, P.n_long_cand, P.long_min, P.queue + Q_ARITH_LONG, (uint32_t)chain_tune ().run4);
}

void launch_arith_decode (DecPlanDev &P, cudaStream_t st)
{
    k_arith_decode_t<false><<<persistent_grid (P.n_arith, P.sm_count, chain_tune ().arith_ctas), 128, 0, st>>>(P.leaves, P.arith_list, P.n_arith, P.queue + Q_ARITH, (uint32_t)chain_tune ().run4, P.n_long_cand, P.long_min);
}
void launch_arith_decode_o0 (DecPlanDev &P, cudaStream_t st)
{
    k_arith_decode_t<true><<<persistent_grid (P.n_arith, P.sm_count, chain_tune ().arith_o0_ctas), 128, 0, st>>>(P.leaves, P.arith_list, P.n_arith, P.queue + Q_ARITH_O0, 0u, 0u, 0u);
}

} // namespace gzb
