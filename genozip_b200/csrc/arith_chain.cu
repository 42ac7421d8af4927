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

namespace gzb {

// Order-0 leaves without the run-length models (the byte planes of the STRIPE codecs, mostly) keep their single model in
// SHARED memory: ~1 KB per warp, ~25-cycle accesses instead of L1/L2 round trips through global memory, and no
// dependence on what the write-through L1 does with a line that has just been stored to.
constexpr uint32_t AR_SMEM_WORDS = 4 + 256 + 8;

__device__ __forceinline__ void ar_model_init_warp (uint32_t *m, uint32_t maxs, int lane)
{
    const uint32_t st = ar_stride (maxs);
    for (uint32_t i = lane; i < st; i += 32) {
        uint32_t v;
        if (i == 0) v = maxs; else if (i == 1) v = ar_f2u (ar_rcp_below (maxs)); else if (i < 4) v = 0;
        else if (i - 4 < maxs) v = 1u | ((i - 4) << 16); else v = 0xffff0000u;
        m[i] = v;
    }
    __syncwarp ();
}

__global__ void __launch_bounds__(128) k_arith_encode (const EncLeaf *leaves, EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list)
{
    const int lane = threadIdx.x & 31;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= n_list) return;
    const uint32_t li = list[slot];
    const EncLeaf &L = leaves[li];
    EncLeafDyn &D = dyn[li];
    const uint32_t n = D.eff_n, maxs = D.nsym;
    const uint8_t * __restrict__ in = D.eff_in;
    const bool o1 = D.eff_order, rle = (D.hdr[0] & F_RLE) != 0;
    uint32_t *lit = D.models;
    uint8_t *out = L.outbuf;
    if (!lit) return;
    __shared__ __align__(16) uint32_t s_model[4][AR_SMEM_WORDS];
    __builtin_assume (__isGlobal (lit)); __builtin_assume (__isGlobal (out)); __builtin_assume (__isGlobal (in));
    uint32_t len;
    if (!o1 && !rle) {
        uint32_t *sm = s_model[threadIdx.x >> 5];
        ar_model_init_warp (sm, maxs, lane);
        len = ar_encode_leaf<false> (sm, maxs, false, in, n, out, lane);
    }
    else len = o1 ? ar_encode_leaf<true> (lit, maxs, rle, in, n, out, lane) : ar_encode_leaf<false> (lit, maxs, rle, in, n, out, lane);
    if (lane == 0) {
        D.tab_len = len;                                                   // whole body at the front of outbuf (n + 1 = expanded)
        D.payload_len = 0;
    }
}

void launch_arith_encode (EncPlanDev &P, cudaStream_t st)
{
    k_arith_encode<<<(P.n_arith + 3) / 4, 128, 0, st>>>(P.leaves, P.dyn, P.arith_list, P.n_arith);
}

__global__ void __launch_bounds__(128) k_arith_decode (DecLeaf *leaves, const uint32_t *list, uint32_t n_list)
{
    const int lane = threadIdx.x & 31;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= n_list) return;
    DecLeaf &L = leaves[list[slot]];
    if (!L.valid || L.err || L.cat || !L.body_ulen || !L.models) return;
    const uint32_t n = L.body_ulen, maxs = L.nsym;
    const bool o1 = L.order == 1, rle = L.rle;
    uint32_t *lit = L.models;
    uint8_t *out = L.dst;
    const uint8_t * __restrict__ body = L.body;
    __shared__ __align__(16) uint32_t s_model[4][AR_SMEM_WORDS];
    __builtin_assume (__isGlobal (lit)); __builtin_assume (__isGlobal (out)); __builtin_assume (__isGlobal (body));
    if (!o1 && !rle) {
        uint32_t *sm = s_model[threadIdx.x >> 5];
        ar_model_init_warp (sm, maxs, lane);
        ar_decode_leaf<false> (sm, maxs, false, body, L.body_len, out, n, lane);
    }
    else if (o1) ar_decode_leaf<true> (lit, maxs, rle, body, L.body_len, out, n, lane);
    else         ar_decode_leaf<false> (lit, maxs, rle, body, L.body_len, out, n, lane);
}

void launch_arith_decode (DecPlanDev &P, cudaStream_t st)
{
    k_arith_decode<<<(P.n_arith + 3) / 4, 128, 0, st>>>(P.leaves, P.arith_list, P.n_arith);
}

} // namespace gzb
