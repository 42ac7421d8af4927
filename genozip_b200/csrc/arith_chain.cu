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

namespace gzb {

// Order-0 leaves without the run-length models (the byte planes of the STRIPE codecs, mostly) keep their single model in
// SHARED memory and search it with per-lane running sums: arith_o0.cuh.

// Two kernels per direction, launched over the same leaf list on two streams: each warp looks at its leaf's class and
// leaves at once if it belongs to the other kernel (separate kernels = separate register allocation for the two loops).
__device__ __forceinline__ bool ar_is_o0_class (bool o1, bool rle) { return !o1 && !rle; }

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
//   GZB_AR_CTAS     general arithmetic kernel (order 1 / RLE leaves), 4 warps per CTA       default 4
//   GZB_AR0_CTAS    order-0 arithmetic kernel                                               default 4
//   GZB_AR_RUN4     the decoder tries four run steps at once                                default 1
//   GZB_AR_SPLIT_STREAM  the split encoder runs on its own stream beside the general kernel default 1
const ChainTune &chain_tune ()                                               // (read at every call: a sweep inside one process changes the variables between batches)
{
    static thread_local ChainTune c;
    auto geti = [] (const char *name, int dflt, int lo, int hi) {
        const char *v = getenv (name);
        if (!v || !*v) return dflt;
        const int x = atoi (v);
        return x < lo ? lo : x > hi ? hi : x;
    };
    c.arith_ctas = geti ("GZB_AR_CTAS", 4, 1, 16);
    c.arith_o0_ctas = geti ("GZB_AR0_CTAS", 4, 1, 16);
    c.run4 = geti ("GZB_AR_RUN4", 1, 0, 1);
    c.split_stream = geti ("GZB_AR_SPLIT_STREAM", 1, 0, 1);
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
__global__ void __launch_bounds__(128) k_arith_decode_t (DecLeaf *leaves, const uint32_t *list, uint32_t n_list, uint32_t *queue, uint32_t run4)
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
        else if (o1) ar_decode_leaf<true> (lit, maxs, rle, body, L.body_len, out, n, lane, run4 != 0);
        else         ar_decode_leaf<false> (lit, maxs, rle, body, L.body_len, out, n, lane, run4 != 0);
    }
}

void launch_arith_decode (DecPlanDev &P, cudaStream_t st)
{
    k_arith_decode_t<false><<<persistent_grid (P.n_arith, P.sm_count, chain_tune ().arith_ctas), 128, 0, st>>>(P.leaves, P.arith_list, P.n_arith, P.queue + Q_ARITH, (uint32_t)chain_tune ().run4);
}
void launch_arith_decode_o0 (DecPlanDev &P, cudaStream_t st)
{
    k_arith_decode_t<true><<<persistent_grid (P.n_arith, P.sm_count, chain_tune ().arith_o0_ctas), 128, 0, st>>>(P.leaves, P.arith_list, P.n_arith, P.queue + Q_ARITH_O0, 0u);
}

} // namespace gzb
