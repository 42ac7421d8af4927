// rans_chain.cu — the latency-bound inner loops of the rANS 4x16 coder: one leaf = 4 lanes = the 4 interleaved states.
//
// The bitstream fixes 4 dependency chains per leaf (reference rANS_static4x16pr.c:434-482, 798-851 encode;
// :555-604, 1021-1083 decode), so a big leaf is walked by ONE warp at the speed of its dependent instruction
// stream.  What matters is therefore (1) instructions per step — steps run in fully unrolled, guard-free blocks of 4,
// table entries for the next block are loaded while the current one executes, read-only data goes through
// __ldg — and (2) running every leaf of every VBlock of the batch at once: big leaves get a warp each, small leaves
// are packed 8 per warp so they share issue slots (warp "jobs", planned by the host).
//
// Shared output / input pointer of the four states: reproduced exactly with a 4-wide ballot per step.
//   encode: states are served 3,2,1,0 within a step and the stream grows backwards (:453-456, :832-835):
//           an emitting lane k writes its 16-bit word at  wp - 2*popc(emitters with index >= k)
//   decode: states renormalise 0,1,2,3 and the stream is read forwards (:578-594, :1062-1066):
//           a renormalising lane k reads the word at     poff + 2*popc(renormalisers with index < k)
#include "gzb_internal.cuh"
#include "hts_enc.cuh"

namespace gzb {

// ================================================================================================ encode
struct EncLane {
    const uint8_t * __restrict__ in;
    const EncSym  * __restrict__ tab;    // O0: shared-memory table by symbol; O1: global table by rank pair
    const uint8_t *rank;                 // shared memory
    uint8_t *end;
    uint32_t n, nsym, shift;
    bool valid;
};

__device__ __forceinline__ EncSym __ldg4 (const EncSym *p)
{
    const uint4 v = __ldg (reinterpret_cast<const uint4 *>(p));
    EncSym e; e.x_max = v.x; e.rcp = v.y; e.bias = v.z; e.cmpl_sh = v.w;
    return e;
}

__device__ __forceinline__ void enc_step (uint32_t &x, uint8_t *&wp, const EncSym &e, bool act, int k, int gshift)
{
    const bool emit = act && x >= e.x_max;
    const uint32_t g = (__ballot_sync (0xffffffffu, emit) >> gshift) & 0xfu;
    if (emit) { *reinterpret_cast<uint16_t *>(wp - 2 * __popc (g >> k)) = (uint16_t)x; x >>= 16; }
    wp -= 2 * __popc (g);
    const uint32_t q = __umulhi (x, e.rcp) >> (e.cmpl_sh >> 16);
    const uint32_t nx = x + e.bias + q * (e.cmpl_sh & 0xffffu);
    x = act ? nx : x;
}

// Order 0.  Step s covers symbols 4*(S-1-s) .. +3 (symbol i belongs to state i&3; the last symbols first, :439-477).
__device__ __forceinline__ uint32_t encode_o0 (const EncLane &f, const EncSym *stab, int lane)
{
    const int k = lane & 3, gshift = lane & ~3;
    uint32_t x = RANS_L;
    uint8_t *wp = f.end;
    const uint32_t steps = f.valid ? (f.n + 3) >> 2 : 0;
    uint32_t maxsteps = steps;
    for (int o = 16; o; o >>= 1) maxsteps = max (maxsteps, __shfl_xor_sync (0xffffffffu, maxsteps, o));
    uint32_t s = 0;
    while (s < maxsteps) {
        const bool active = f.valid && s < steps;
        // fast region: every active lane has 4 full steps ahead (s >= 1 when the first step is the partial remainder)
        uint32_t lim = active ? (((f.n & 3) && s == 0) ? 0u : steps) : 0xffffffffu;
        for (int o = 16; o; o >>= 1) lim = min (lim, __shfl_xor_sync (0xffffffffu, lim, o));
        if (lim != 0xffffffffu && s + 4 <= lim) {
            const uint8_t *ip = active ? f.in + 4 * (steps - 1 - s) + k : f.in;
            EncSym e0, e1, e2, e3;
            if (active) { e0 = stab[__ldg (ip)]; e1 = stab[__ldg (ip - 4)]; e2 = stab[__ldg (ip - 8)]; e3 = stab[__ldg (ip - 12)]; }
            for (; s + 4 <= lim; s += 4) {
                const EncSym c0 = e0, c1 = e1, c2 = e2, c3 = e3;
                ip -= 16;
                if (active && s + 8 <= lim) { e0 = stab[__ldg (ip)]; e1 = stab[__ldg (ip - 4)]; e2 = stab[__ldg (ip - 8)]; e3 = stab[__ldg (ip - 12)]; }
                enc_step (x, wp, c0, active, k, gshift);
                enc_step (x, wp, c1, active, k, gshift);
                enc_step (x, wp, c2, active, k, gshift);
                enc_step (x, wp, c3, active, k, gshift);
            }
        }
        else {                                                               // one guarded step (leaf heads and tails)
            bool act = false; EncSym e; e.x_max = 0xffffffffu; e.rcp = 0; e.bias = 0; e.cmpl_sh = 0;
            if (active) {
                const uint32_t idx = 4 * (steps - 1 - s) + k;
                if (idx < f.n) { act = true; e = stab[f.in[idx]]; }
            }
            enc_step (x, wp, e, act, k, gshift);
            s++;
        }
    }
    if (f.valid) {                                                           // RansEncFlush 3,2,1,0 (:479-482); 2-byte aligned only
        uint16_t *w = reinterpret_cast<uint16_t *>(wp - 4 * (4 - k));
        w[0] = (uint16_t)x; w[1] = (uint16_t)(x >> 16);
    }
    return f.valid ? (uint32_t)(f.end - wp) + 16 : 0;
}

// Order 1.  Lane k walks its quarter backwards, pos = pstart .. k*q4, then one step in context 0 (:806-846); chain 3 also
// owns the remainder, so lanes 0-2 join `delay` steps later.
__device__ __forceinline__ uint32_t encode_o1 (const EncLane &f, const uint8_t *srank, int lane)
{
    const int k = lane & 3, gshift = lane & ~3;
    uint32_t x = RANS_L;
    uint8_t *wp = f.end;
    const uint32_t q4 = f.n >> 2, r = f.n & 3;
    const uint32_t steps = f.valid ? q4 + r : 0;
    const uint32_t len = (k == 3) ? q4 + r : q4, delay = (k == 3) ? 0 : r;
    const uint32_t pstart = (k == 3) ? f.n - 2 : (k + 1) * q4 - 2;
    const uint32_t ns = f.nsym;
    uint32_t l = f.valid ? srank[f.in[pstart + 1]] : 0;
    uint32_t maxsteps = steps;
    for (int o = 16; o; o >>= 1) maxsteps = max (maxsteps, __shfl_xor_sync (0xffffffffu, maxsteps, o));
    // ---- the hot transition: a symbol that follows itself with probability >= 63/64 (e.g. the ACGT exception stream, 99.9 %
    // zeros).  Its encoder symbol is kept in registers and a block of 4 steps in which every lane of the warp codes it is
    // run without any table access: one aligned word load brings the lane's 4 input bytes, one compare recognises them.
    uint32_t hrank = 0xffffffffu, hot4b = 0;
    EncSym H; H.x_max = 0; H.rcp = 0; H.bias = 0; H.cmpl_sh = 0;
    {
        const uint32_t size = 1u << f.shift;
        uint32_t best = 0;
        if (f.valid) for (uint32_t rr = k; rr < ns; rr += 4) {
            const EncSym e = __ldg4 (f.tab + rr * ns + rr);
            const uint32_t fr = size - (e.cmpl_sh & 0xffffu);
            if (fr <= size && fr >= size - (size >> 6) && e.x_max == ((RANS_L >> f.shift) << 16) * fr) best = max (best, (fr << 8) | rr);
        }
        best = max (best, __shfl_xor_sync (0xffffffffu, best, 1));
        best = max (best, __shfl_xor_sync (0xffffffffu, best, 2));
        uint32_t b = 0;                                                      // the byte value of that rank (rank 0 is symbol 0: forced present, :741)
        if (best) {
            hrank = best & 0xffu;
            H = __ldg4 (f.tab + hrank * ns + hrank);
            if (hrank) for (uint32_t c = k ? k : 4; c < 256; c += 4) if (srank[c] == hrank) b = c;
        }
        b = max (b, __shfl_xor_sync (0xffffffffu, b, 1));                   // (warp-wide shuffles stay outside divergent code)
        b = max (b, __shfl_xor_sync (0xffffffffu, b, 2));
        hot4b = b * 0x01010101u;
    }
    const bool hot_on = __any_sync (0xffffffffu, hrank != 0xffffffffu);
    const uint32_t hsh = H.cmpl_sh >> 16, hcmpl = H.cmpl_sh & 0xffffu;
    uint32_t s = 0;
    while (s < maxsteps) {
        const bool active = f.valid && s < steps;
        // fast region [s, lim): all 4 lanes of every active group regular and not at their final context-0 step:
        // from s >= r (lanes 0-2 have joined) to steps-1 (exclusive)
        uint32_t lim = active ? (s >= r ? steps - 1 : 0u) : 0xffffffffu;
        for (int o = 16; o; o >>= 1) lim = min (lim, __shfl_xor_sync (0xffffffffu, lim, o));
        if (hot_on && lim != 0xffffffffu && s + 16 <= lim) {
            // super-blocks of 16 steps.  The lane's 16 input bytes in[ip-15 .. ip] come as 4 funnel-shifted aligned words
            // (sliding window, loaded one super-block ahead); a super-block in which every lane of the warp sees only the
            // hot symbol runs from registers: 4 instructions per step, one vote per 16 steps.  The state grows
            // monotonically under the hot symbol, so testing the value before the last step covers the emit test of all 16.
            const uint8_t *ip = active ? f.in + (pstart - (s - delay)) : f.in + 15;
            const uintptr_t A = reinterpret_cast<uintptr_t>(ip) - 3;
            const uint32_t sh8 = 8u * (uint32_t)(A & 3);
            const uint32_t *wp32 = reinterpret_cast<const uint32_t *>(A & ~(uintptr_t)3);     // aligned word holding in[ip-3]
            uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0, wlast = 0;         // n_j: bytes of sub-block j (byte 3 = first coded); wlast: lowest word loaded
            #define LOAD_SUPER() { const uint32_t a0 = __ldg (wp32), a1 = __ldg (wp32 - 1), a2 = __ldg (wp32 - 2), a3 = __ldg (wp32 - 3); \
                                   n0 = __funnelshift_r (a0, wlast, sh8); n1 = __funnelshift_r (a1, a0, sh8); n2 = __funnelshift_r (a2, a1, sh8); \
                                   n3 = __funnelshift_r (a3, a2, sh8); wlast = a3; wp32 -= 4; }
            if (active) { if (sh8) wlast = __ldg (wp32 + 1); LOAD_SUPER (); }
            uint32_t backoff = 0, hskip = 0, hfail = 0;
            for (; s + 16 <= lim; s += 16) {
                const uint32_t b0 = n0, b1 = n1, b2 = n2, b3 = n3;
                if (active && s + 32 <= lim) LOAD_SUPER ();
                bool done = false;
                if (hskip == 0) {
                    const bool ishot = !active || (b0 == hot4b && b1 == hot4b && b2 == hot4b && b3 == hot4b && l == hrank);
                    if (__all_sync (0xffffffffu, ishot)) {
                        uint32_t y = x, ylast = x;
                        #pragma unroll
                        for (int t = 0; t < 16; t++) { ylast = y; y = y + H.bias + (__umulhi (y, H.rcp) >> hsh) * hcmpl; }
                        if (!__any_sync (0xffffffffu, active && ylast >= H.x_max)) { if (active) x = y; done = true; }
                    }
                    if (done) hfail = 0; else { hfail = min (2 * hfail + 1, 7u); hskip = hfail; }
                }
                else hskip--;
                if (done) continue;
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t bytes = j == 0 ? b0 : j == 1 ? b1 : j == 2 ? b2 : b3;
                    EncSym c0, c1, c2, c3;
                    c0.x_max = c1.x_max = c2.x_max = c3.x_max = 0xffffffffu; c0.rcp = c1.rcp = c2.rcp = c3.rcp = 0;
                    c0.bias = c1.bias = c2.bias = c3.bias = 0; c0.cmpl_sh = c1.cmpl_sh = c2.cmpl_sh = c3.cmpl_sh = 0;
                    if (active) {
                        const uint32_t r0 = srank[bytes >> 24], r1 = srank[(bytes >> 16) & 0xffu], r2 = srank[(bytes >> 8) & 0xffu], r3 = srank[bytes & 0xffu];
                        c0 = __ldg4 (f.tab + r0 * ns + l); c1 = __ldg4 (f.tab + r1 * ns + r0); c2 = __ldg4 (f.tab + r2 * ns + r1); c3 = __ldg4 (f.tab + r3 * ns + r2);
                        l = r3;
                    }
                    bool exact = backoff != 0;
                    if (!exact) {
                        uint32_t y = x; bool emit = false;
                        #define SPEC(c) { emit |= y >= c.x_max; y = y + c.bias + (__umulhi (y, c.rcp) >> (c.cmpl_sh >> 16)) * (c.cmpl_sh & 0xffffu); }
                        SPEC (c0) SPEC (c1) SPEC (c2) SPEC (c3)
                        #undef SPEC
                        if (__any_sync (0xffffffffu, emit && active)) { exact = true; backoff = 5; }
                        else if (active) x = y;
                    }
                    if (exact) {
                        backoff--;
                        enc_step (x, wp, c0, active, k, gshift);
                        enc_step (x, wp, c1, active, k, gshift);
                        enc_step (x, wp, c2, active, k, gshift);
                        enc_step (x, wp, c3, active, k, gshift);
                    }
                }
            }
            #undef LOAD_SUPER
        }
        else if (lim != 0xffffffffu && s + 4 <= lim) {
            const uint8_t *ip = active ? f.in + (pstart - (s - delay)) : f.in + 3;
            EncSym e0, e1, e2, e3;
            #define LOAD_O1() { const uint32_t r0 = srank[__ldg (ip)], r1 = srank[__ldg (ip - 1)], r2 = srank[__ldg (ip - 2)], r3 = srank[__ldg (ip - 3)]; \
                                e0 = __ldg4 (f.tab + r0 * ns + l); e1 = __ldg4 (f.tab + r1 * ns + r0); e2 = __ldg4 (f.tab + r2 * ns + r1); e3 = __ldg4 (f.tab + r3 * ns + r2); l = r3; }
            if (active) LOAD_O1 ();
            // Low-entropy streams (e.g. the ACGT exception stream) renormalise once in hundreds of steps.  A block is first run
            // speculatively WITHOUT the shared-pointer logic (no ballot, no popc, no store: 6 instead of ~20 instructions per
            // step); only if some lane of the warp would have emitted is it replayed exactly.  After a failed speculation the
            // next 16 blocks go straight to the exact path, so dense streams pay ~6%.
            uint32_t backoff = 0;
            for (; s + 4 <= lim; s += 4) {
                const EncSym c0 = e0, c1 = e1, c2 = e2, c3 = e3;
                ip -= 4;
                if (active && s + 8 <= lim) LOAD_O1 ();
                bool exact = backoff != 0;
                if (!exact) {
                    uint32_t y = x; bool emit = false;
                    #define SPEC(c) { emit |= y >= c.x_max; y = y + c.bias + (__umulhi (y, c.rcp) >> (c.cmpl_sh >> 16)) * (c.cmpl_sh & 0xffffu); }
                    SPEC (c0) SPEC (c1) SPEC (c2) SPEC (c3)
                    #undef SPEC
                    if (__any_sync (0xffffffffu, emit && active)) { exact = true; backoff = 17; }
                    else if (active) x = y;
                }
                if (exact) {
                    backoff--;
                    enc_step (x, wp, c0, active, k, gshift);
                    enc_step (x, wp, c1, active, k, gshift);
                    enc_step (x, wp, c2, active, k, gshift);
                    enc_step (x, wp, c3, active, k, gshift);
                }
            }
            #undef LOAD_O1
        }
        else {
            bool act = false; EncSym e; e.x_max = 0xffffffffu; e.rcp = 0; e.bias = 0; e.cmpl_sh = 0;
            const int t0 = (int)s - (int)delay;
            if (f.valid && t0 >= 0 && t0 < (int)len) {
                const uint32_t cr = ((uint32_t)t0 == len - 1) ? srank[0] : srank[f.in[pstart - t0]];
                act = true; e = __ldg4 (f.tab + cr * ns + l); l = cr;
            }
            enc_step (x, wp, e, act, k, gshift);
            s++;
        }
    }
    if (f.valid) {
        uint16_t *w = reinterpret_cast<uint16_t *>(wp - 4 * (4 - k));
        w[0] = (uint16_t)x; w[1] = (uint16_t)(x >> 16);
    }
    return f.valid ? (uint32_t)(f.end - wp) + 16 : 0;
}

__global__ void __launch_bounds__(32) k_rans_encode (const EncLeaf *leaves, EncLeafDyn *dyn, const uint32_t *order_list, const uint2 *jobs, uint32_t n_jobs)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];                      // 8 x (EncSym[256] + rank[256])
    EncSym  (*s_tab)[256]  = reinterpret_cast<EncSym (*)[256]>(s_dyn);
    uint8_t (*s_rank)[256] = reinterpret_cast<uint8_t (*)[256]>(s_dyn + (size_t)8 * 256 * sizeof (EncSym));
    if (blockIdx.x >= n_jobs) return;
    const uint2 job = jobs[blockIdx.x];
    const int lane = threadIdx.x, grp = lane >> 2, k = lane & 3;
    EncLane f; f.valid = false; f.in = nullptr; f.tab = nullptr; f.rank = nullptr; f.end = nullptr; f.n = f.nsym = 0; f.shift = 12;
    uint32_t li = 0; bool o1 = false;
    if ((uint32_t)grp < job.y) {
        li = order_list[job.x + grp];
        const EncLeaf &L = leaves[li];
        const EncLeafDyn &D = dyn[li];
        if (D.eff_n && D.symtab) {
            f.valid = true; f.in = D.eff_in; f.n = D.eff_n; f.nsym = D.nsym; f.tab = D.symtab; f.rank = D.rank; f.shift = D.shift;
            o1 = D.eff_order;
            f.end = L.outbuf + (L.out_cap & ~1u);
        }
    }
    for (int g = 0; g < (int)job.y; g++) {                                  // stage rank maps (and O0 tables) in shared memory
        const bool v = __shfl_sync (0xffffffffu, (int)f.valid, g * 4);
        if (!v) continue;
        const bool go1 = __shfl_sync (0xffffffffu, (int)o1, g * 4);
        const unsigned long long rp = __shfl_sync (0xffffffffu, (unsigned long long)f.rank, g * 4);
        const unsigned long long tp = __shfl_sync (0xffffffffu, (unsigned long long)f.tab, g * 4);
        for (int i = lane; i < 256; i += 32) s_rank[g][i] = reinterpret_cast<const uint8_t *>(rp)[i];
        if (!go1) for (int i = lane; i < 256; i += 32) s_tab[g][i] = reinterpret_cast<const EncSym *>(tp)[i];
    }
    __syncwarp ();
    // a job is planned for one requested order; leaves demoted to order 0 by the "<8 symbols" rule (:1333-1336) run first
    const bool any_o1 = __any_sync (0xffffffffu, f.valid && o1), any_o0 = __any_sync (0xffffffffu, f.valid && !o1);
    uint32_t plen = 0;
    if (any_o0) { EncLane f0 = f; f0.valid = f.valid && !o1; const uint32_t p = encode_o0 (f0, s_tab[grp], lane); if (f0.valid) plen = p; }
    if (any_o1) { EncLane f1 = f; f1.valid = f.valid && o1;  const uint32_t p = encode_o1 (f1, s_rank[grp], lane); if (f1.valid) plen = p; }
    if ((uint32_t)grp < job.y && k == 0) dyn[li].payload_len = plen;
}

void launch_rans_encode (EncPlanDev &P, cudaStream_t st)
{
    k_rans_encode<<<P.n_rans_jobs, 32, 8 * (256 * sizeof (EncSym) + 256), st>>>(P.leaves, P.dyn, P.rans_list, P.rans_jobs, P.n_rans_jobs);
}

// ================================================================================================ decode
struct DecLane {
    const uint8_t * __restrict__ body;
    uint8_t *out;
    const uint2    *lut;        // O0 (shared memory for single-leaf jobs)
    const uint32_t * __restrict__ lut1;   // O1 merged LUT
    const uint8_t  *symof;      // O1 row -> symbol (shared memory)
    uint32_t n, body_len, shift, row0, nctx;
    bool valid;
};

__device__ __forceinline__ void dec_renorm (uint32_t &x, uint32_t &poff, bool act, const uint8_t *body, uint32_t body_len, int k, int gshift)
{
    const bool need = act && x < RANS_L;
    const uint32_t g = (__ballot_sync (0xffffffffu, need) >> gshift) & 0xfu;
    if (need) {
        const uint32_t a = poff + 2 * __popc (g & ((1u << k) - 1));
        if (a + 1 < body_len) x = (x << 16) | __ldg (body + a) | (__ldg (body + a + 1) << 8);   // RansDecRenormSafe (rANS_word.h:397-405)
    }
    poff += 2 * __popc (g);
}

template <bool SLUT> __device__ __forceinline__ void decode_o0 (const DecLane &d, const uint2 *slut, uint32_t poff, int lane)
{
    const int k = lane & 3, gshift = lane & ~3;
    uint32_t x = 0;
    if (d.valid) { const uint8_t *q = d.body + poff + 4 * k; x = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24); poff += 16; }
    const uint32_t steps = d.valid ? (d.n + 3) >> 2 : 0, full = d.valid ? d.n >> 2 : 0;
    uint32_t maxsteps = steps;
    for (int o = 16; o; o >>= 1) maxsteps = max (maxsteps, __shfl_xor_sync (0xffffffffu, maxsteps, o));
    uint8_t *op = d.out + k;
    uint32_t s = 0;
    while (s < maxsteps) {
        const bool active = d.valid && s < steps;
        uint32_t lim = active ? full : 0xffffffffu;                          // steps in which all 4 lanes of an active group decode
        for (int o = 16; o; o >>= 1) lim = min (lim, __shfl_xor_sync (0xffffffffu, lim, o));
        if (lim != 0xffffffffu && s + 4 <= lim) {
            for (; s + 4 <= lim; s += 4) {
                #pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (active) {
                        const uint2 e = SLUT ? slut[x & 4095] : __ldg (d.lut + (x & 4095));
                        x = (e.x >> 16) * (x >> 12) + e.y;
                        op[4 * t] = (uint8_t)e.x;
                    }
                    dec_renorm (x, poff, active, d.body, d.body_len, k, gshift);
                }
                op += 16;
            }
        }
        else {
            const bool act = active && 4 * s + k < d.n;
            if (act) { const uint2 e = SLUT ? slut[x & 4095] : __ldg (d.lut + (x & 4095)); x = (e.x >> 16) * (x >> 12) + e.y; *op = (uint8_t)e.x; }
            op += 4;
            dec_renorm (x, poff, act, d.body, d.body_len, k, gshift);
            s++;
        }
    }
}

__device__ __forceinline__ void decode_o1 (const DecLane &d, const uint8_t *ssym, uint32_t poff, int lane)
{
    const int k = lane & 3, gshift = lane & ~3;
    uint32_t x = 0;
    if (d.valid) { const uint8_t *q = d.body + poff + 4 * k; x = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24); poff += 16; }
    const uint32_t q4 = d.n >> 2, r = d.n - 4 * q4;
    const uint32_t steps = d.valid ? q4 + r : 0;                             // chains 0-2 decode q4 symbols; chain 3 also the remainder (:1076-1083)
    uint32_t maxsteps = steps;
    for (int o = 16; o; o >>= 1) maxsteps = max (maxsteps, __shfl_xor_sync (0xffffffffu, maxsteps, o));
    const uint32_t shift = d.shift, mask = (1u << shift) - 1;
    uint32_t coff = d.row0 << shift;                                         // LUT row of the previous symbol; context 0 first (:1029)
    // Each chain writes its own quarter of the output.  Bytes are gathered in a 32-bit window and stored one aligned word
    // at a time: per-symbol byte stores from hundreds of concurrent leaves saturate the L2 write path long before anything else.
    uint8_t *op = d.out + (size_t)k * q4;
    const uint8_t *op0 = op;
    const uint32_t head_end = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(op) & 3)) & 3);
    uint32_t win = 0;
    #define PUT_SLOW(b) do { win = __byte_perm (win, (b), 0x4321); const uintptr_t A = reinterpret_cast<uintptr_t>(op); const uint32_t cnt = (uint32_t)(op - op0); \
                             if ((A & 3) == 3 && cnt >= 3) *reinterpret_cast<uint32_t *>(A - 3) = win; else if (cnt < head_end) *op = (uint8_t)(b); op++; } while (0)
    // ---- the hot transition (see encode_o1): a symbol that follows itself with probability >= 63/64 is decoded from
    // registers — slot range [hotB, hotB + hotF) of LUT row hot_coff — without touching the LUT.  Found by probing the
    // middle slot of every row (an interval longer than half the table contains it) and verified on both of its ends.
    uint32_t hotF = 0, hotB = 0, hot_coff = 0xffffffffu, hot4 = 0;
    {
        const uint32_t size = 1u << shift, mid = size >> 1;
        uint32_t best = 0;
        if (d.valid) for (uint32_t rr = k; rr < d.nctx; rr += 4) {
            const uint32_t e = __ldg (d.lut1 + (rr << shift) + mid);
            const uint32_t F = ((e >> 8) & 0xfffu) + 1;
            if ((e & 0xffu) == rr && F >= size - (size >> 6)) best = max (best, (F << 8) | rr);
        }
        best = max (best, __shfl_xor_sync (0xffffffffu, best, 1));
        best = max (best, __shfl_xor_sync (0xffffffffu, best, 2));
        if (best) {
            const uint32_t rr = best & 0xffu, F = best >> 8;
            const uint32_t e = __ldg (d.lut1 + (rr << shift) + mid);
            const uint32_t B = mid - (e >> 20);
            if (B <= mid && B + F <= size) {
                const uint32_t e_lo = __ldg (d.lut1 + (rr << shift) + B), e_hi = __ldg (d.lut1 + (rr << shift) + B + F - 1);
                const uint32_t want = rr | ((F - 1) << 8);
                if (e_lo == want && e_hi == (want | ((F - 1) << 20))) { hotF = F; hotB = B; hot_coff = rr << shift; hot4 = (uint32_t)ssym[rr] * 0x01010101u; }
            }
        }
    }
    const bool hot_on = __any_sync (0xffffffffu, hotF != 0);
    uint32_t s = 0;
    while (s < maxsteps) {
        const bool active = d.valid && s < steps;
        uint32_t lim = active ? q4 : 0xffffffffu;
        for (int o = 16; o; o >>= 1) lim = min (lim, __shfl_xor_sync (0xffffffffu, lim, o));
        if (lim != 0xffffffffu && s >= 4 && s + 4 <= lim) {
            const uint32_t ph = (uint32_t)(reinterpret_cast<uintptr_t>(op) & 3);   // constant over the region: op advances by 4 per block
            const bool st0 = ph == 3, st1 = ph == 2, st2 = ph == 1, st3 = ph == 0;
            // Speculation as in the encoder: a block is first decoded without the shared read pointer logic; if any lane of the
            // warp dropped below the renormalisation bound the block is replayed exactly (the speculative word stores are
            // simply overwritten).  16 exact blocks follow a failed speculation.
            uint32_t backoff = 0, hskip = 0, hfail = 0;
            const uint32_t hsh = 8u * (4u - ph);                             // bytes of a hot block that complete the lane's current word
            const uint32_t hcm = (1u << shift) - hotF;
            for (; s + 4 <= lim; s += 4) {
                if (hot_on && s + 16 <= lim) {
                    // 16 hot steps from registers: x' = F*(x >> shift) + (x & mask) - B  =  x - (x >> shift)*(size - F) - B, a
                    // two-instruction dependency per step.  The state only shrinks, so the renormalisation test of the last
                    // step covers all 16; the slot test (x & mask) - B < F is tracked as a running maximum.
                    if (hskip == 0) {
                        uint32_t y = x, dmax = 0;
                        #pragma unroll
                        for (int t = 0; t < 16; t++) {
                            dmax = max (dmax, (y & mask) - hotB);
                            y = (y - hotB) - (y >> shift) * hcm;
                        }
                        const bool bad = active && (dmax >= hotF || y < RANS_L || coff != hot_coff);
                        if (!__any_sync (0xffffffffu, bad)) {
                            if (active) {
                                x = y;
                                uint32_t *w = reinterpret_cast<uint32_t *>(op - ph);
                                w[0] = __funnelshift_rc (win, hot4, hsh); w[1] = hot4; w[2] = hot4; w[3] = hot4;
                                win = hot4; op += 16;
                            }
                            hfail = 0; s += 12;
                            continue;
                        }
                        hfail = min (2 * hfail + 1, 15u); hskip = hfail;
                    }
                    else hskip--;
                }
                bool exact = backoff != 0;
                if (!exact) {
                    uint32_t y = x, yc = coff, yw = win; bool low = false;
                    if (active) {
                        #pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const uint32_t m = y & mask;
                            const uint32_t e = __ldg (d.lut1 + yc + m);
                            const uint32_t ys = y >> shift;
                            y = ((e >> 8) & 0xfffu) * ys + (ys + (e >> 20));
                            low |= y < RANS_L;
                            yc = (e & 0xffu) << shift;
                            yw = __byte_perm (yw, (uint32_t)ssym[e & 0xffu], 0x4321);
                            const bool stt = t == 0 ? st0 : t == 1 ? st1 : t == 2 ? st2 : st3;
                            if (stt) *reinterpret_cast<uint32_t *>(op + t - 3) = yw;
                        }
                    }
                    if (__any_sync (0xffffffffu, low)) { exact = true; backoff = 17; }
                    else { x = y; coff = yc; win = yw; }
                }
                if (exact) {
                    backoff--;
                    #pragma unroll
                    for (int t = 0; t < 4; t++) {
                        if (active) {
                            const uint32_t m = x & mask;
                            const uint32_t e = __ldg (d.lut1 + coff + m);
                            const uint32_t xs = x >> shift;
                            x = ((e >> 8) & 0xfffu) * xs + (xs + (e >> 20));
                            coff = (e & 0xffu) << shift;
                            win = __byte_perm (win, (uint32_t)ssym[e & 0xffu], 0x4321);
                            const bool stt = t == 0 ? st0 : t == 1 ? st1 : t == 2 ? st2 : st3;
                            if (stt) *reinterpret_cast<uint32_t *>(op + t - 3) = win;
                        }
                        dec_renorm (x, poff, active, d.body, d.body_len, k, gshift);
                    }
                }
                if (active) op += 4;                                         // a finished group's pointer must stay put for its tail flush
            }
        }
        else {
            const bool act = active && (s < q4 || k == 3);
            if (act) {
                const uint32_t m = x & mask;
                const uint32_t e = __ldg (d.lut1 + coff + m);
                const uint32_t xs = x >> shift;
                x = ((e >> 8) & 0xfffu) * xs + (xs + (e >> 20));
                coff = (e & 0xffu) << shift;
                PUT_SLOW ((uint32_t)ssym[e & 0xffu]);
            }
            dec_renorm (x, poff, act, d.body, d.body_len, k, gshift);
            s++;
        }
    }
    #undef PUT_SLOW
    if (d.valid) {                                                           // bytes after the last aligned word boundary
        const uint32_t cnt = (uint32_t)(op - op0), tail = (uint32_t)(reinterpret_cast<uintptr_t>(op) & 3);
        for (uint32_t t = 0; t < tail && t < cnt; t++) op[-1 - (int)t] = (uint8_t)(win >> (24 - 8 * t));
    }
}

__global__ void __launch_bounds__(32) k_rans_decode (DecLeaf *leaves, const uint32_t *list, const uint2 *jobs, uint32_t n_jobs)
{
    __shared__ uint2 s_lut[4096];                                            // order-0 LUT of a single-leaf job
    __shared__ uint8_t s_symof[8][256];
    if (blockIdx.x >= n_jobs) return;
    const uint2 job = jobs[blockIdx.x];
    const int lane = threadIdx.x, grp = lane >> 2;
    DecLane d; d.valid = false; d.body = nullptr; d.out = nullptr; d.lut = nullptr; d.lut1 = nullptr; d.symof = nullptr;
    d.n = d.body_len = d.row0 = d.nctx = 0; d.shift = 12;
    bool o1 = false; uint32_t poff = 0;
    const DecLeaf *Lp = nullptr;
    if ((uint32_t)grp < job.y) {
        const DecLeaf &L = leaves[list[job.x + grp]];
        if (L.valid && !L.err && !L.cat && L.body_ulen && (L.lut || L.lut1)) {
            d.valid = true; o1 = L.order; d.body = L.body; d.body_len = L.body_len; d.out = L.dst; d.n = L.body_ulen;
            poff = L.payload_off; d.lut = L.lut; d.lut1 = L.lut1; d.shift = L.shift; d.row0 = L.ctxrank[0]; d.nctx = L.nctx;
            Lp = &L;
        }
    }
    for (int g = 0; g < (int)job.y; g++) {                                  // row -> symbol maps of order-1 leaves
        const bool v = __shfl_sync (0xffffffffu, (int)(d.valid && o1), g * 4);
        if (!v) continue;
        const unsigned long long lp = __shfl_sync (0xffffffffu, (unsigned long long)Lp, g * 4);
        for (int i = lane; i < 256; i += 32) s_symof[g][i] = reinterpret_cast<const DecLeaf *>(lp)->symof[i];
    }
    bool slut = false;
    if (job.y == 1) {
        slut = __shfl_sync (0xffffffffu, (int)(d.valid && !o1), 0);
        if (slut) {
            const unsigned long long lp = __shfl_sync (0xffffffffu, (unsigned long long)d.lut, 0);
            for (int i = lane; i < 4096; i += 32) s_lut[i] = reinterpret_cast<const uint2 *>(lp)[i];
        }
    }
    __syncwarp ();
    const bool any_o1 = __any_sync (0xffffffffu, d.valid && o1), any_o0 = __any_sync (0xffffffffu, d.valid && !o1);
    if (any_o0) {
        DecLane d0 = d; d0.valid = d.valid && !o1;
        if (slut) decode_o0<true> (d0, s_lut, poff, lane); else decode_o0<false> (d0, s_lut, poff, lane);
    }
    if (any_o1) { DecLane d1 = d; d1.valid = d.valid && o1; decode_o1 (d1, s_symof[grp & 7], poff, lane); }
}

void launch_rans_decode (DecPlanDev &P, cudaStream_t st)
{
    k_rans_decode<<<P.n_rans_jobs, 32, 0, st>>>(P.leaves, P.rans_list, P.rans_jobs, P.n_rans_jobs);
}

} // namespace gzb
