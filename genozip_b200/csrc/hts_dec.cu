// hts_dec.cu — decode side of the htscodecs "4x16" rANS / adaptive-arithmetic containers on sm_100a.
//
// Bit-exact inverse of reference rans_uncompress_to_4x16 (src/htscodecs/rANS_static4x16pr.c:1358-1642) and
// arith_uncompress_to (src/htscodecs/arith_dynamic.c:860-1104) for every container genozip's order bytes can
// produce (PACK, STRIPE N=4, CAT, NOSZ, arithmetic RLE).  Phases over all sections of a batch:
//
//   parse containers → build decode tables / init models → rANS chains / arithmetic chains → raw copies
//   → unpack → unstripe
#include "gzb_internal.cuh"
#include "hts_enc.cuh"
#include "arith_model.cuh"

namespace gzb {

__device__ __forceinline__ uint32_t get_varint (const uint8_t *p, const uint8_t *end, uint32_t *v)   // varint.h:267-300
{
    const uint8_t *s = p;
    uint32_t acc = 0; int limit = 6; uint8_t c;
    if (p >= end) { *v = 0; return 0; }
    do { c = *p++; acc = (acc << 7) | (c & 0x7f); } while ((c & 0x80) && p < end && --limit > 0);
    *v = acc;
    return (uint32_t)(p - s);
}

// ------------------------------------------------------------------------------------------------ container parse
// one non-STRIPE container (:1441-1551 / arith_dynamic.c:945-1025)
__device__ int parse_leaf (const uint8_t *in, uint32_t in_len, uint32_t expect, uint8_t coder,
                           uint8_t *fin, uint8_t *tmp, DecLeaf &L)
{
    const uint8_t *end = in + in_len;
    L.valid = 1; L.coder = coder; L.err = 0;
    L.lut = nullptr; L.lut1 = nullptr; L.models = nullptr;
    if (!in_len) return -1;
    uint32_t flags = *in++;
    if (flags & F_STRIPE) return -1;                                     // nested STRIPE is never produced
    if (coder == CODER_RANS && (flags & F_RLE)) return -1;               // rANS RLE is never requested by genozip
    L.order = (uint8_t)(flags & (coder == CODER_ARITH ? 3u : 1u));
    L.cat = (flags & F_CAT) != 0; L.rle = (flags & F_RLE) != 0; L.pack = (flags & F_PACK) != 0;
    uint32_t osz = expect;
    if (!(flags & F_NOSZ)) { in += get_varint (in, end, &osz); if (osz != expect) return -1; }
    L.ulen = osz; L.body_ulen = osz; L.per_byte = 1;
    L.fin = fin; L.dst = fin;
    if (L.pack) {                                                        // hts_unpack_meta (pack.c:168-201)
        if (in >= end) return -1;
        uint32_t ns = in[0] ? in[0] : 256;
        if (ns <= 16) {
            L.per_byte = ns <= 1 ? 0 : ns <= 2 ? 8 : ns <= 4 ? 4 : 2;
            if (in + 1 + ns > end) return -1;
            for (uint32_t i = 0; i < 16; i++) L.map[i] = i < ns ? in[1 + i] : 0;
            in += 1 + ns;
        }
        else { L.per_byte = 1; in += 1; }
        uint32_t psz;
        in += get_varint (in, end, &psz);
        if (psz > osz) return -1;
        L.body_ulen = psz;
        L.dst = tmp;
    }
    L.body = in;
    L.body_len = (uint32_t)(end - in);
    if (coder == CODER_ARITH && (flags & F_EXT) && L.body_len && !L.cat) return -1;   // bzip2 payload: never written by genozip, refused by the reference's
                                                                         // build too (arith_dynamic.c:1050-1058); without a payload (:1037) or under CAT (:1043) the flag is never looked at
    if (!L.body_len) {                                                   // :1598-1601: nothing is decoded, tmp1_size = 0 ...
        if (!(L.pack && L.per_byte == 0)) return -1;                     // ... so the result is empty (or hts_unpack fails, pack.c:242) and the plug-in's
        L.body_ulen = 0;                                                 //     out_len == uncompressed_len check aborts (codec_htscodecs.c:111,126); only a
    }                                                                    //     one-symbol PACK map rebuilds the output from no payload at all
    if (L.cat && L.body_ulen > L.body_len) return -1;
    return 0;
}

__global__ void k_dec_parse (const DecSection *secs, DecLeaf *leaves, SectionResult *res, uint32_t n_secs)
{
    uint32_t si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= n_secs) return;
    const DecSection &S = secs[si];
    DecLeaf *L = leaves + 4 * (size_t)si;
    for (int i = 0; i < 4; i++) L[i].valid = 0;
    res[si].out_len = S.n; res[si].status = 0;
    if (!S.in_len) { res[si].status = -4; return; }
    if (!(S.in[0] & F_STRIPE)) {
        if (parse_leaf (S.in, S.in_len, S.n, S.coder, S.out, S.tmp, L[0])) { L[0].valid = 0; res[si].status = -4; }
        return;
    }
    // STRIPE (:1366-1439): flags, varint ulen, N, N x varint clen, N sub-containers
    const uint8_t *end = S.in + S.in_len;
    uint32_t h = 1, ulen;
    h += get_varint (S.in + h, end, &ulen);
    if (h >= S.in_len || ulen != S.n) { res[si].status = -4; return; }
    uint32_t N = S.in[h++];
    if (N != 4) { res[si].status = -4; return; }                         // genozip always writes N = 4 (:1166-1167)
    uint32_t clen[4], tot = 0;
    for (int i = 0; i < 4; i++) {
        h += get_varint (S.in + h, end, &clen[i]);
        if (h > S.in_len || clen[i] > S.in_len || clen[i] < 1) { res[si].status = -4; return; }
        tot += clen[i];
    }
    if ((uint64_t)h + tot > S.in_len) { res[si].status = -4; return; }
    uint32_t idx = 0;
    const uint32_t end_off = h + tot;
    for (int i = 0; i < 4; i++) {
        uint32_t ul = ulen / 4 + ((ulen % 4) > (uint32_t)i);
        // the reference hands each sub-decoder everything up to the end of the STRIPE container (:1425)
        if (parse_leaf (S.in + h, end_off - h, ul, S.coder, S.planes + idx, S.tmp + idx, L[i])) { res[si].status = -4; L[i].valid = 0; }
        h += clen[i];
        idx += ul;
    }
}

// ------------------------------------------------------------------------------------------------ rANS decode tables
struct DTabSmem {
    uint32_t F[256];
    uint32_t F0[256];
    uint32_t start[256];
    uint32_t scan[8];
    uint32_t T, ok, consumed, nctx;
    const uint8_t *p;
};

// decode_alphabet (:205-252).  returns bytes consumed, 0 on error
__device__ uint32_t get_alphabet (const uint8_t *p, const uint8_t *end, uint32_t *F)
{
    if (p >= end) return 0;
    const uint8_t *s = p;
    int run = 0, j = *p++;
    do {
        F[j] = 1;
        if (p >= end) return 0;
        if (!run && j + 1 == *p) { if (p + 1 >= end) return 0; j = *p++; run = *p++; }
        else if (run) { run--; if (++j > 255) return 0; }
        else j = *p++;
    } while (j);
    return (uint32_t)(p - s);
}

// exclusive prefix over sm.F by symbol → sm.start; returns total (all threads)
__device__ uint32_t scan_F (DTabSmem &sm)
{
    const int tid = threadIdx.x;
    uint32_t f = sm.F[tid], v = f;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, v, o); if ((tid & 31) >= o) v += t; }
    if ((tid & 31) == 31) sm.scan[tid >> 5] = v;
    __syncthreads ();
    uint32_t base = 0, total = 0;
    for (int w = 0; w < 8; w++) { if (w < (tid >> 5)) base += sm.scan[w]; total += sm.scan[w]; }
    sm.start[tid] = base + v - f;
    __syncthreads ();
    return total;
}

// order-0 table at p (:526-549): fills lut[4096]; returns bytes consumed (0 = error).  Whole CTA.
__device__ uint32_t build_dec_o0 (DTabSmem &sm, const uint8_t *p, const uint8_t *end, uint2 *lut)
{
    const int tid = threadIdx.x;
    sm.F[tid] = 0;
    __syncthreads ();
    if (tid == 0) {
        const uint8_t *s = p;
        uint32_t k = get_alphabet (p, end, sm.F);
        uint64_t tot = 0;                                                  // a varint is any 32-bit value: sums must not wrap (the reference checks
        bool ok = k != 0;                                                  // every F[j] against what is left of TOTFREQ, :538)
        p += k;
        if (k) for (int j = 0; j < 256; j++) if (sm.F[j]) { p += get_varint (p, end, &sm.F[j]); if (sm.F[j] > 4096) ok = false; tot += sm.F[j]; }
        sm.ok = ok && tot <= 4096;
        sm.T = (uint32_t)(tot <= 4096 ? tot : 0); sm.consumed = (uint32_t)(p - s);
    }
    __syncthreads ();
    if (!sm.ok || !sm.T || sm.T > 4096 || (sm.T & (sm.T - 1))) return 0;
    int sh = 0; for (uint32_t t = sm.T; t < 4096; t *= 2) sh++;         // normalise_freq_shift (:165-176)
    sm.F[tid] <<= sh;
    __syncthreads ();
    uint32_t total = scan_F (sm);
    if (total != 4096) return 0;
    uint32_t f = sm.F[tid], st = sm.start[tid];
    if (st + f > 4096) f = 0;                                              // (cannot happen once the sums are exact; the arena must never be overrun)
    for (uint32_t y = 0; y < f; y++) lut[st + y] = make_uint2 ((uint32_t)tid | (f << 16), y);
    __syncthreads ();
    return sm.consumed;
}

// 4 lanes = 4 states decode n symbols from `in` (payload at off) with an order-0 LUT; used for the compressed O1 table
__device__ void rans_o0_decode_lanes (const uint8_t *body, uint32_t body_len, uint32_t off, const uint2 *lut,
                                      uint8_t *out, uint32_t n, int lane, bool valid)
{
    const int k = lane & 3, gshift = lane & ~3;
    uint32_t x = 0, poff = off + 16;
    if (valid) { const uint8_t *q = body + off + 4 * k; x = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24); }
    uint32_t steps = valid ? (n + 3) >> 2 : 0, maxsteps = steps;
    for (int o = 16; o; o >>= 1) maxsteps = max (maxsteps, __shfl_xor_sync (0xffffffffu, maxsteps, o));
    for (uint32_t s = 0; s < maxsteps; s++) {
        uint32_t idx = 4 * s + k;
        bool act = valid && s < steps && idx < n;
        if (act) {
            uint2 e = lut[x & 4095];
            x = (e.x >> 16) * (x >> 12) + e.y;
            out[idx] = (uint8_t)e.x;
        }
        bool need = act && x < RANS_L;
        uint32_t g = (__ballot_sync (0xffffffffu, need) >> gshift) & 0xfu;
        if (need) {
            uint32_t a = poff + 2 * __popc (g & ((1u << k) - 1));
            if (a + 1 < body_len) x = (x << 16) | body[a] | (body[a + 1] << 8);          // RansDecRenormSafe
        }
        poff += 2 * __popc (g);
    }
}

// RansDecInit + "if (R < RANS_BYTE_L) goto err" for the four states (:555-558, :1023-1026)
__device__ bool rans_states_low (const uint8_t *q)
{
    for (int k = 0; k < 4; k++) {
        const uint32_t x = q[4 * k] | (q[4 * k + 1] << 8) | (q[4 * k + 2] << 16) | ((uint32_t)q[4 * k + 3] << 24);
        if (x < RANS_L) return true;
    }
    return false;
}

// one CTA per leaf slot
__global__ void __launch_bounds__(256) k_dec_tables (DecLeaf *leaves, SectionResult *res, uint32_t n_slots, Arena arena)
{
    if (blockIdx.x >= n_slots) return;
    DecLeaf &L = leaves[blockIdx.x];
    if (!L.valid || L.coder != CODER_RANS || L.cat || !L.body_ulen) return;
    __shared__ DTabSmem sm;
    __shared__ uint8_t *s_ptr[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *body = L.body, *end = body + L.body_len;
    const uint32_t si = blockIdx.x >> 2;
    bool fail = false;

    if (L.body_len < 16) fail = true;
    else if (!L.order) {                                                  // ---- order 0 (:498-558)
        if (tid == 0) s_ptr[0] = arena.alloc (4096 * 8);
        __syncthreads ();
        if (!s_ptr[0]) return;                                            // arena overflow: host replays the batch
        uint32_t used = build_dec_o0 (sm, body, end - 8, reinterpret_cast<uint2 *>(s_ptr[0]));
        if (!used || used + 16 > L.body_len || rans_states_low (body + used)) fail = true;
        else if (tid == 0) { L.lut = reinterpret_cast<uint2 *>(s_ptr[0]); L.payload_off = used; }
    }
    else {                                                                // ---- order 1 (:883-1019)
        const uint32_t shift = body[0] >> 4;
        const bool comp = body[0] & 1;
        const uint8_t *p = body + 1, *fend = end, *tab_end = nullptr;
        if (shift != 10 && shift != 12) fail = true;
        if (!fail && comp) {
            uint32_t usz, csz;
            p += get_varint (p, end, &usz);
            p += get_varint (p, end, &csz);
            if (p + csz + 16 > end || usz > (1u << 20) || csz < 16) fail = true;
            if (!fail) {
                if (tid == 0) { s_ptr[0] = arena.alloc (4096 * 8); s_ptr[1] = arena.alloc (usz + 16); }
                __syncthreads ();
                if (!s_ptr[0] || !s_ptr[1]) return;
                uint2 *nlut = reinterpret_cast<uint2 *>(s_ptr[0]);
                uint32_t used = build_dec_o0 (sm, p, p + csz - 8, nlut);
                if (!used || used + 16 > csz || rans_states_low (p + used)) fail = true;
                else {
                    if (warp == 0) rans_o0_decode_lanes (p, csz, used, nlut, s_ptr[1], usz, lane, lane < 4);
                    __syncthreads ();
                    tab_end = p + csz;
                    p = s_ptr[1]; fend = s_ptr[1] + usz;
                }
            }
        }
        if (!fail) {
            // alphabet of all symbols, then per present context the frequencies (:970-1011)
            sm.F0[tid] = 0;
            __syncthreads ();
            if (tid == 0) {
                uint32_t k = get_alphabet (p, fend, sm.F0);
                sm.ok = k != 0 && p + k < fend;
                sm.p = p + k;
                uint32_t nc = 0;
                for (int i = 0; i < 256; i++) if (sm.F0[i]) nc++;
                sm.nctx = nc;
                s_ptr[2] = nullptr;
                if (sm.ok) s_ptr[2] = arena.alloc (((unsigned long long)nc << shift) * 4);
                uint32_t rr = 0;
                for (int i = 0; i < 256; i++) if (sm.F0[i]) { L.ctxrank[i] = (uint8_t)rr; L.symof[rr] = (uint8_t)i; rr++; } else L.ctxrank[i] = 0;
            }
            __syncthreads ();
            if (!sm.ok) fail = true;
            else if (!s_ptr[2]) return;
        }
        if (!fail) {
            uint32_t *lut1 = reinterpret_cast<uint32_t *>(s_ptr[2]);
            uint32_t row = 0;
            for (int i = 0; i < 256 && !fail; i++) {
                if (!sm.F0[i]) continue;                                  // uniform: F0 is shared
                sm.F[tid] = 0;
                __syncthreads ();
                if (tid == 0) {                                           // decode_freq_d (:324-355)
                    const uint8_t *q = sm.p;
                    uint64_t T = 0; int zrun = 0; bool ok = q < fend;           // (64-bit: a varint is any 32-bit value, :1003)
                    for (int j = 0; ok && j < 256 && q < fend; j++) {
                        if (!sm.F0[j]) continue;
                        uint32_t f;
                        if (zrun) { f = 0; zrun--; }
                        else {
                            q += get_varint (q, fend, &f);
                            if (!f) { if (q >= fend) { ok = false; break; } zrun = *q++; }
                        }
                        if (f > (1u << shift)) ok = false;
                        sm.F[j] = f; T += f;
                    }
                    sm.ok = ok && T <= (1u << shift); sm.T = (uint32_t)(T <= (1u << shift) ? T : 0); sm.p = q;
                }
                __syncthreads ();
                if (!sm.ok) { fail = true; break; }
                const uint32_t T = sm.T;
                if (T) {
                    if (T > (1u << shift) || (T & (T - 1))) { fail = true; break; }
                    int sh = 0; for (uint32_t t = T; t < (1u << shift); t *= 2) sh++;
                    sm.F[tid] <<= sh;
                    __syncthreads ();
                    uint32_t total = scan_F (sm);
                    if (total != (1u << shift)) { fail = true; break; }
                    uint32_t f = sm.F[tid], st = sm.start[tid];
                    if (st + f > (1u << shift)) f = 0;
                    uint32_t *srow = lut1 + ((size_t)row << shift);
                    const uint32_t base_e = (uint32_t)L.ctxrank[tid] | ((f - 1) << 8);
                    for (uint32_t y = 0; y < f; y++) srow[st + y] = base_e | (y << 20);
                }
                else {
                    // a context without frequencies (:994-997): a valid stream never enters it.  The reference's decoder would read
                    // whatever an earlier call left in its table; here a damaged stream finds a defined row (row 0's symbol,
                    // next context row 0) instead of arena leftovers that could point the next look-up outside the table
                    uint32_t *srow = lut1 + ((size_t)row << shift);
                    for (uint32_t y = tid; y < (1u << shift); y += 256) srow[y] = 0;
                }
                row++;
                __syncthreads ();
            }
            if (!fail) {
                const uint8_t *pay = tab_end ? tab_end : sm.p;
                if (pay + 16 > end || rans_states_low (pay)) fail = true;
                else if (tid == 0) { L.lut1 = lut1; L.nctx = (uint16_t)sm.nctx; L.shift = (uint8_t)shift; L.payload_off = (uint32_t)(pay - body); }
            }
        }
    }
    if (fail && tid == 0) { L.err = -4; res[si].status = -4; }
}

// ------------------------------------------------------------------------------------------------ arithmetic decoder
__global__ void k_arith_dec_init (DecLeaf *leaves, uint32_t n_slots, Arena arena)
{
    if (blockIdx.x >= n_slots) return;
    DecLeaf &L = leaves[blockIdx.x];
    if (!L.valid || L.coder != CODER_ARITH || L.cat || !L.body_ulen || L.err) return;
    __shared__ uint32_t *s_m;
    const uint32_t maxs = L.body[0] ? L.body[0] : 256, stride = ar_stride (maxs), nctx = L.order ? 256 : 1;
    if (threadIdx.x == 0) { s_m = reinterpret_cast<uint32_t *>(arena.alloc (((unsigned long long)nctx * stride + 258 * AR_RUN_STRIDE) * 4)); L.models = s_m; L.nsym = (uint16_t)maxs; }
    __syncthreads ();
    if (!s_m) return;
    for (uint32_t c = threadIdx.x; c < nctx; c += blockDim.x) ar_model_init (s_m + c * stride, maxs);
    if (L.rle) for (uint32_t c = threadIdx.x; c < 258; c += blockDim.x) ar_model_init (s_m + nctx * stride + c * AR_RUN_STRIDE, 4);
}

// ------------------------------------------------------------------------------------------------ post passes
// raw (CAT) bodies; grid (slots, parts)
__global__ void k_dec_cat (const DecLeaf *leaves, uint32_t n_slots)
{
    if (blockIdx.x >= n_slots) return;
    const DecLeaf &L = leaves[blockIdx.x];
    if (!L.valid || L.err || !L.cat) return;
    uint32_t n = L.body_ulen, part = ((n + gridDim.y - 1) / gridDim.y + 15) & ~15u;
    uint32_t b = blockIdx.y * part, e = min (b + part, n);
    for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) L.dst[i] = L.body[i];
}

// hts_unpack (pack.c:214-351); grid (slots, parts)
__global__ void k_dec_unpack (const DecLeaf *leaves, SectionResult *res, uint32_t n_slots)
{
    if (blockIdx.x >= n_slots) return;
    const DecLeaf &L = leaves[blockIdx.x];
    if (!L.valid || L.err || !L.pack) return;
    const uint32_t per = L.per_byte;
    uint32_t n = (per == 1) ? L.body_ulen : L.ulen;
    if (per == 1 && n != L.ulen) { if (!threadIdx.x && !blockIdx.y) res[blockIdx.x >> 2].status = -4; return; }
    if (per > 1 && (uint64_t)(n + per - 1) / per > L.body_ulen) { if (!threadIdx.x && !blockIdx.y) res[blockIdx.x >> 2].status = -4; return; }
    uint32_t part = ((n + gridDim.y - 1) / gridDim.y + 15) & ~15u;
    uint32_t b = blockIdx.y * part, e = min (b + part, n);
    const uint8_t *src = L.dst; uint8_t *dst = L.fin;
    if (per == 1) { for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) dst[i] = src[i]; return; }
    if (per == 0) { for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) dst[i] = L.map[0]; return; }
    const uint32_t bits = 8 / per, m = (1u << bits) - 1;
    for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x)
        dst[i] = L.map[(src[i / per] >> ((i % per) * bits)) & m];
}

// unstripe (utils.h:41-73); grid (sections, parts)
__global__ void k_dec_unstripe (const DecSection *secs, const SectionResult *res, uint32_t n_secs)
{
    if (blockIdx.x >= n_secs) return;
    const DecSection &S = secs[blockIdx.x];
    if (res[blockIdx.x].status || !S.in_len || !(S.in[0] & F_STRIPE)) return;
    const uint32_t n = S.n;
    uint32_t idx[4], acc = 0;
    for (int j = 0; j < 4; j++) { idx[j] = acc; acc += n / 4 + ((n % 4) > (uint32_t)j); }
    uint32_t part = ((n + gridDim.y - 1) / gridDim.y + 15) & ~15u;
    uint32_t b = blockIdx.y * part, e = min (b + part, n);
    for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) S.out[i] = S.planes[idx[i & 3] + (i >> 2)];
}

// ------------------------------------------------------------------------------------------------ launcher
#define LAUNCH(kern, grid, block, ...) do { kern<<<(grid), (block), 0, st>>>(__VA_ARGS__); P.launches++; } while (0)

void dec_run (DecPlanDev &P, cudaStream_t st)
{
    const uint32_t ns = P.n_sections, nslots = 4 * ns;
    if (!ns) return;
    LAUNCH (k_dec_parse, (ns + 127) / 128, 128, P.sections, P.leaves, P.results, ns);
    if (P.n_rans)  LAUNCH (k_dec_tables, nslots, 256, P.leaves, P.results, nslots, P.arena);
    if (P.n_arith) LAUNCH (k_arith_dec_init, nslots, 256, P.leaves, nslots, P.arena);
    // The rANS and the arithmetic leaves are independent and both kernels are latency-bound (a handful of warps per SM),
    // so they run concurrently: the arithmetic kernel is forked onto the engine's second stream and joined afterwards.
    cudaEventRecord (P.ev_chain0, st);
    if (P.n_arith) {
        cudaStreamWaitEvent (P.st2, P.ev_chain0, 0);
        cudaEventRecord (P.ev_arith0, P.st2);
        if (P.n_long_cand) {                                               // the long order-1 leaves first, on a stream of their own
            cudaStreamWaitEvent (P.st4, P.ev_chain0, 0);
            launch_arith_decode_long (P, P.st4); P.launches++;
            cudaEventRecord (P.ev_long, P.st4);
        }
        launch_arith_decode (P, P.st2); P.launches++;
        cudaEventRecord (P.ev_chain2, P.st2);
        cudaStreamWaitEvent (P.st3, P.ev_chain0, 0);
        launch_arith_decode_o0 (P, P.st3); P.launches++;
        cudaEventRecord (P.ev_o0, P.st3);
    }
    if (P.n_rans_jobs) { launch_rans_decode (P, st); P.launches++; }
    cudaEventRecord (P.ev_chain1, st);
    if (P.n_arith) { cudaStreamWaitEvent (st, P.ev_chain2, 0); cudaStreamWaitEvent (st, P.ev_o0, 0); if (P.n_long_cand) cudaStreamWaitEvent (st, P.ev_long, 0); }
    dim3 g (nslots, P.parts), gs (ns, P.parts);
    LAUNCH (k_dec_cat, g, 256, P.leaves, nslots);
    LAUNCH (k_dec_unpack, g, 256, P.leaves, P.results, nslots);
    LAUNCH (k_dec_unstripe, gs, 256, P.sections, P.results, ns);
}

} // namespace gzb
