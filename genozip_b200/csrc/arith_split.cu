// arith_split.cu — the adaptive arithmetic ENCODER of a long order-1 leaf, taken apart.
//
// In the encoder the adaptive models evolve with the symbols alone: what the range coder does never feeds back into them
// (c_simple_model.h:123-146 reads and bumps the model; c_range_coder.h:97-109 only consumes cumFreq / freq / totFreq).  And
// under order 1 the 256 context models do not see each other: context c sees, in order, exactly the symbols that follow a c.
// So the one chain "model step + coder step" per symbol of arith_chain.cu splits into
//   A  per context, in parallel: its sub-sequence through its own model -> one record (cumFreq, freq, totFreq) per symbol;
//   B  per leaf: the range coder's recurrence over the records — a division, two multiplies, the carry and the renormalisation.
// The bytes are the reference's: the same models see the same symbols in the same order, the same coder sees the same triples.
// The decoder cannot be split this way (the symbol depends on the code).
//
//   k_ar_split_bucket   one CTA per leaf: positions grouped by context, in order (a stable counting sort by the previous symbol)
//   k_ar_split_model    one warp per (leaf, context): stage A, on arith_model.cuh's model code
//   k_ar_split_code     one warp per leaf: stage B
#include "gzb_internal.cuh"
#include "hts_enc.cuh"
#include "arith_model.cuh"

namespace gzb {

constexpr int SPLIT_WARPS = 32;                  // warps of the bucket CTA; each owns a contiguous segment of the leaf

__device__ __forceinline__ bool ar_split_leaf (const EncLeafDyn &D) { return D.split_pos != nullptr; }

// ------------------------------------------------------------------------------------------------ positions by context
__global__ void __launch_bounds__(SPLIT_WARPS * 32) k_ar_split_bucket (EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list)
{
    if (blockIdx.x >= n_list) return;
    EncLeafDyn &D = dyn[list[blockIdx.x]];
    if (!ar_split_leaf (D)) return;
    __shared__ uint32_t cur[SPLIT_WARPS][256];                               // counts, then write cursors, per warp segment and context
    __shared__ uint32_t tot[256], cnt[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = D.eff_n;
    const uint8_t * __restrict__ in = D.eff_in;
    uint32_t *pos = D.split_pos, *start = D.split_start;
    const uint32_t seg = (((n + SPLIT_WARPS - 1) / SPLIT_WARPS) + 31) & ~31u;
    const uint32_t b = min (n, (uint32_t)warp * seg), e = min (n, b + seg);
    for (int i = tid; i < SPLIT_WARPS * 256; i += SPLIT_WARPS * 32) (&cur[0][0])[i] = 0;
    __syncthreads ();
    for (uint32_t i = b + lane; i < e; i += 32) atomicAdd (&cur[warp][i ? in[i - 1] : 0], 1u);   // the context of symbol i (arith_dynamic.c:176: last = 0 at the start)
    __syncthreads ();
    if (tid < 256) { uint32_t t = 0; for (int w = 0; w < SPLIT_WARPS; w++) t += cur[w][tid]; tot[tid] = t; cnt[tid] = t; }
    __syncthreads ();
    if (tid < 256) {                                                         // contexts by decreasing length of their sub-sequence: stage A takes the long ones first
        const uint32_t mine = cnt[tid];
        uint32_t rank = 0;
        for (int c = 0; c < 256; c++) { const uint32_t o = cnt[c]; rank += (o > mine) || (o == mine && c < tid); }
        reinterpret_cast<uint8_t *>(start + 257)[rank] = (uint8_t)tid;
    }
    if (warp == 0) {                                                         // exclusive scan of the 256 totals: 8 per lane
        uint32_t s = 0, v[8];
        for (int k = 0; k < 8; k++) { v[k] = tot[8 * lane + k]; s += v[k]; }
        uint32_t inc = s;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        uint32_t x = inc - s;
        for (int k = 0; k < 8; k++) { tot[8 * lane + k] = x; x += v[k]; }
    }
    __syncthreads ();
    if (tid < 256) {
        uint32_t x = tot[tid];
        start[tid] = x;
        for (int w = 0; w < SPLIT_WARPS; w++) { const uint32_t c = cur[w][tid]; cur[w][tid] = x; x += c; }
        if (tid == 255) start[256] = x;
    }
    __syncthreads ();
    for (uint32_t i0 = b; i0 < e; i0 += 32) {                                // in order within the segment: ranks among equal contexts by lane
        const uint32_t i = i0 + lane;
        const bool act = i < e;
        const uint32_t key = act ? (i ? in[i - 1] : 0u) : 256u + lane;
        const uint32_t peers = __match_any_sync (0xffffffffu, key);
        const uint32_t base = act ? cur[warp][key] : 0;
        __syncwarp ();
        if (act) {
            pos[base + __popc (peers & ((1u << lane) - 1))] = i;
            if ((peers >> lane) == 1) cur[warp][key] = base + __popc (peers);   // the highest lane of the group moves the cursor
        }
        __syncwarp ();
    }
}

// ------------------------------------------------------------------------------------------------ stage A
// One symbol of one context through its model: SIMPLE_MODEL_encodeSymbol (c_simple_model.h:123-146) up to the call of RC_Encode,
// whose arguments are returned: x = cumFreq | freq << 16, y = totFreq (all below 2^16: MAX_FREQ, :70).
__device__ __forceinline__ uint2 ar_model_rec (uint32_t *m, uint32_t maxs, ArCache &c, uint32_t sym, int lane, bool &stale)
{
    uint2 rec; rec.y = c.tot;
    const uint32_t f0 = c.e0 & 0xffffu;
    if ((c.e0 >> 16) == sym)      { rec.x = f0 << 16; AR_BUMP_CASE (0, c.e0, c.e0) }
    else if ((c.e1 >> 16) == sym) { rec.x = f0 | (c.e1 << 16); AR_BUMP_CASE (1, c.e1, c.e0) }
    else if ((c.e2 >> 16) == sym) { rec.x = (f0 + (c.e1 & 0xffffu)) | (c.e2 << 16); AR_BUMP_CASE (2, c.e2, c.e1) }
    else if ((c.e3 >> 16) == sym) { rec.x = (f0 + (c.e1 & 0xffffu) + (c.e2 & 0xffffu)) | (c.e3 << 16); AR_BUMP_CASE (3, c.e3, c.e2) }
    else {
        const ArHit h = ar_find_sym (m, maxs, sym, lane);
        const uint32_t p = h.p, e = h.e, prev = h.prev;
        if (p >= maxs) { rec.x = 1u << 16; return rec; }                    // cannot happen for a symbol < maxs
        rec.x = h.acc | (e << 16);
        AR_BUMP_DEEP (p, e, prev)
    }
    return rec;
}

// Persistent warps take (rank, leaf) items from a counter, rank-major: every leaf's longest sub-sequence first, then every leaf's
// second longest, ... — a (leaf, context) grid in index order would reach the last leaf's hot context only after all the others'
// cold ones.  Within a sub-sequence a run of the context's top symbol is emitted by the lanes together: symbol k of the run
// has cumFreq 0, freq f0 + 16k, totFreq tot + 16k (c_simple_model.h:123-146 applied k times; the halving bounds the run).
__global__ void __launch_bounds__(128) k_ar_split_model (EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list, uint32_t *queue)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n_items = n_list * 256u;
    for (;;) {
        const uint32_t q = queue_take (queue, lane);
        if (q >= n_items) return;
        const uint32_t slot = q % n_list, rank = q / n_list;
        EncLeafDyn &D = dyn[list[slot]];
        if (!ar_split_leaf (D)) continue;
        const uint32_t maxs = D.nsym;
        const uint32_t ctx = reinterpret_cast<const uint8_t *>(D.split_start + 257)[rank];
        if (ctx >= maxs) continue;
        const uint32_t j0 = D.split_start[ctx], j1 = D.split_start[ctx + 1];
        if (j0 == j1) continue;
        const uint8_t * __restrict__ in = D.eff_in;
        const uint32_t * __restrict__ pos = D.split_pos;
        uint2 *recs = D.split_rec;
        uint32_t *m = D.models + ctx * ar_stride (maxs);
        ArCache c; ar_load (m, c);
        bool dirty = false;
        uint32_t nx_p = 0, nx_s = 0;
        if (j0 + lane < j1) { nx_p = pos[j0 + lane]; nx_s = in[nx_p]; }
        for (uint32_t j = j0; j < j1; j += 32) {
            const uint32_t cnt = min (32u, j1 - j);
            const uint32_t my_p = nx_p, my_s = nx_s;                           // 32 symbols of the sub-sequence at once; the next 32 are fetched under this group's work
            if (j + 32 + lane < j1) { nx_p = pos[j + 32 + lane]; nx_s = in[nx_p]; }
            uint32_t my_x = 0, my_y = 0;
            uint32_t t = 0;
            while (t < cnt) {
                const uint32_t top = c.e0 >> 16;
                const uint32_t hit = __ballot_sync (0xffffffffu, (uint32_t)lane < cnt && my_s == top) >> t;
                uint32_t run = hit == 0xffffffffu ? 32u : (uint32_t)__ffs (~hit) - 1;   // symbols t .. t+run-1 are the top entry again
                const uint32_t room = c.tot + AR_STEP <= AR_MAXF ? (AR_MAXF - c.tot) / AR_STEP : 0;
                run = min (run, room);
                if (run) {
                    const uint32_t k = (uint32_t)lane - t;
                    if (k < run) { my_x = (c.e0 + AR_STEP * k) << 16; my_y = c.tot + AR_STEP * k; }
                    c.e0 += AR_STEP * run; c.tot += AR_STEP * run;
                    dirty = true;
                    t += run;
                    continue;
                }
                if (dirty) { c.rtot = ar_rcp_below (c.tot); ar_flush (m, c); dirty = false; }
                const uint32_t sy = __shfl_sync (0xffffffffu, my_s, t);
                bool stale = false;
                const uint2 rec = ar_model_rec (m, maxs, c, sy, lane, stale);
                if (stale) ar_load (m, c);
                if ((uint32_t)lane == t) { my_x = rec.x; my_y = rec.y; }
                t++;
            }
            if ((uint32_t)lane < cnt) recs[my_p] = make_uint2 (my_x, my_y);
        }
        if (dirty) { c.rtot = ar_rcp_below (c.tot); ar_flush (m, c); }
    }
}

// ------------------------------------------------------------------------------------------------ stage B
// The loop-carried chain is range -> I2F -> FMUL -> F2I -> IMAD -> ISETP -> IMAD.  Everything else is kept off it: the 32 records of a
// group and the reciprocals of their totals are staged in shared memory (one 16-byte load per symbol, issued one symbol ahead; the
// next group's records are fetched from global memory one group ahead), the quotient's correction is a predicated increment (a
// deficit above one — young models only — goes the long way through ar_div).
__global__ void __launch_bounds__(128) k_ar_split_code (const EncLeaf *leaves, EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list)
{
    __shared__ uint4 s_rec[4][2][32];                                        // per warp, two groups: cumFreq | freq << 16, totFreq, reciprocal, -
    const int lane = threadIdx.x & 31;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= n_list) return;
    const uint32_t li = list[slot];
    EncLeafDyn &D = dyn[li];
    if (!ar_split_leaf (D)) return;
    const uint32_t n = D.eff_n;
    const uint2 * __restrict__ recs = D.split_rec;
    uint8_t *out = leaves[li].outbuf;
    out[0] = (uint8_t)D.nsym;                                                // arith_dynamic.c:162-167 (256 wraps to 0)
    ArEnc rc; rc.low = 0; rc.range = 0xffffffffu; rc.ffnum = 0; rc.cache = 0; rc.carry = 0; rc.out = out + 1;
    const uint8_t *limit = out + n + 8;
    uint32_t len = 0;
    bool full = false;
    uint2 nx = make_uint2 (1u << 16, 1u);
    if ((uint32_t)lane < n) nx = recs[lane];
    for (uint32_t i0 = 0; i0 < n && !full; i0 += 32) {
        const uint32_t cnt = min (32u, n - i0);
        uint4 *buf = s_rec[threadIdx.x >> 5][(i0 >> 5) & 1];
        const uint2 my = nx;
        if (i0 + 32 + lane < n) nx = recs[i0 + 32 + lane];
        buf[lane] = make_uint4 (my.x, my.y, ar_f2u (ar_rcp_below (my.y)), 0u);
        __syncwarp ();
        uint4 r_next = buf[0];
        for (uint32_t t = 0; t < cnt; t++) {
            const uint4 r_ = r_next;
            r_next = buf[(t + 1) & 31];
            const uint32_t tot = r_.y;
#ifdef __CUDA_ARCH__
            uint32_t q = __float2uint_rz (__fmul_rz (__uint2float_rz (rc.range), ar_u2f (r_.z)));
            const uint32_t rem = rc.range - q * tot;                         // RC_Encode (c_range_coder.h:97-109): range / totFreq, exactly
            if (rem >= tot) { if (rem - tot >= tot) q = ar_div (rc.range, tot, ar_u2f (r_.z)); else q++; }
#else
            const uint32_t q = rc.range / tot;
#endif
            const uint32_t before = rc.low;
            rc.low += (r_.x & 0xffffu) * q; rc.range = (r_.x >> 16) * q;
            rc.carry += rc.low < before;
            if (rc.range < AR_TOP) {
                do { rc.range <<= 8; ar_shift_low (rc); } while (rc.range < AR_TOP);
                if (rc.out + rc.ffnum > limit) { full = true; break; }       // certain to reach the input length: stored raw (arith_dynamic.c:847-852)
            }
        }
    }
    if (full) len = n + 1;
    else { for (int i = 0; i < 5; i++) ar_shift_low (rc); len = (uint32_t)(rc.out - out); }   // RC_FinishEncode
    if (lane == 0) { D.tab_len = len; D.payload_len = 0; }
}

void launch_arith_encode_split (EncPlanDev &P, cudaStream_t st)
{
    if (!P.n_arith_big) return;
    k_ar_split_bucket<<<P.n_arith_big, SPLIT_WARPS * 32, 0, st>>>(P.dyn, P.arith_list, P.n_arith_big);
    if (P.ev_prof[0]) cudaEventRecord (P.ev_prof[0], st);
    {
        const uint32_t want = P.n_arith_big * 64u, cap = (uint32_t)(P.sm_count > 0 ? P.sm_count : 148) * 8u;   // 8 CTAs = 32 warps per SM (64 registers per thread)
        k_ar_split_model<<<want < cap ? want : cap, 128, 0, st>>>(P.dyn, P.arith_list, P.n_arith_big, P.queue + Q_SPLIT_MODEL);
    }
    if (P.ev_prof[1]) cudaEventRecord (P.ev_prof[1], st);
    k_ar_split_code<<<(P.n_arith_big + 3) / 4, 128, 0, st>>>(P.leaves, P.dyn, P.arith_list, P.n_arith_big);
}

} // namespace gzb
