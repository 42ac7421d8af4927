// hts_enc.cu — encode side of the htscodecs "4x16" rANS / adaptive-arithmetic containers on sm_100a.
//
// Byte-identical to reference rans_compress_to_4x16 (src/htscodecs/rANS_static4x16pr.c:1151-1356) and
// arith_compress_to (src/htscodecs/arith_dynamic.c:615-858) for genozip's order bytes 0x01/0x19/0x81/0x99
// (src/codec_htscodecs.c:17-20).  A batch of sections is expanded by the host planner (api.cu) into leaves;
// the kernels here run over *all leaves of the batch* per phase, with no host round trip between phases:
//
//   transpose (STRIPE planes) → hist0 → pack decide → pack → hist0' → leaf prep → hist1 → tables
//   → rANS chains / arithmetic chains → leaf final (CAT rule) → section final (STRIPE selection) → segment copy
//
// The chain kernels are dependency-bound (4 chains per rANS leaf, 1 per arithmetic leaf — fixed by the
// bitstream); all other passes are bandwidth-shaped.
#include "gzb_internal.cuh"
#include "hts_enc.cuh"
#include "arith_model.cuh"

namespace gzb {

__constant__ double c_log_o1_10[257];   // log(1024 + k), k = 0..256, computed by the host's libm (compute_shift :647)
__constant__ double c_log_o1_12[257];   // log(4096 + k)                                               (:648)

void upload_log_tables (const double *l10, const double *l12)
{
    cudaMemcpyToSymbol (c_log_o1_10, l10, 257 * sizeof (double));
    cudaMemcpyToSymbol (c_log_o1_12, l12, 257 * sizeof (double));
}

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ uint32_t pow2_ceil (uint32_t v) { v--; v |= v>>1; v |= v>>2; v |= v>>4; v |= v>>8; v |= v>>16; return v + 1; }

__device__ __forceinline__ int put_varint (uint8_t *p, uint32_t v)       // varint.h:203-237 (big-endian base 128)
{
    int nb = v < (1u<<7) ? 1 : v < (1u<<14) ? 2 : v < (1u<<21) ? 3 : v < (1u<<28) ? 4 : 5;
    for (int k = nb - 1; k >= 0; k--) *p++ = (uint8_t)(((v >> (7*k)) & 0x7f) | (k ? 0x80 : 0));
    return nb;
}

__device__ __forceinline__ EncSym make_encsym (uint32_t start, uint32_t freq, uint32_t bits)   // RansEncSymbolInit, rANS_word.h:190-255
{
    EncSym s;
    s.x_max = ((RANS_L >> bits) << 16) * freq;
    uint32_t cmpl = ((1u << bits) - freq) & 0xffff;
    if (freq < 2) { s.rcp = 0xffffffffu; s.bias = start + (1u << bits) - 1; s.cmpl_sh = cmpl; }
    else {
        uint32_t k = 32 - __clz (freq - 1);                               // ceil(log2(freq))
        s.rcp  = (uint32_t)(((1ull << (k + 31)) + freq - 1) / freq);
        s.bias = start;
        s.cmpl_sh = cmpl | ((k - 1) << 16);
    }
    return s;
}

// normalise_freq (rANS_static4x16pr.c:113-160) on a compact row of `len` counters; serial, one thread.
// Literal semantics incl. the re-use of `size` as the scaled running sum on the single retry.
__device__ int scale_freqs (uint32_t *F, int len, int size, uint32_t tot)
{
    if (!size) return 0;
    int retried = 0, top = 0;
    for (;;) {
        uint64_t mul = ((uint64_t)tot << 31) / (uint32_t)size + (uint32_t)((1 << 30) / size);
        uint32_t top_val = 0; top = 0; size = 0;
        for (int j = 0; j < len; j++) {
            uint32_t f = F[j];
            if (!f) continue;
            if (top_val < f) { top_val = f; top = j; }
            f = (uint32_t)((f * mul) >> 31);
            if (!f) f = 1;
            F[j] = f;
            size += (int)f;
        }
        int adjust = (int)tot - size;
        if (adjust > 0) { F[top] += adjust; break; }
        if (adjust == 0) break;
        uint32_t need = (uint32_t)-adjust;
        if (F[top] > need && (retried || F[top] / 2 >= need)) { F[top] += adjust; break; }
        if (!retried) { retried = 1; continue; }
        adjust += (int)F[top] - 1;
        F[top] = 1;
        for (int j = 0; adjust && j < len; j++) {
            if (F[j] < 2) continue;
            int d = F[j] > (uint32_t)-adjust;
            int m = d ? adjust : 1 - (int)F[j];
            F[j] += m; adjust -= m;
        }
        break;
    }
    return F[top] > 0 ? 0 : -1;
}

// encode_alphabet (:179-203) over a 256-entry presence array
template <typename T> __device__ int put_alphabet (uint8_t *p, const T *present)
{
    uint8_t *s = p;
    int skip = 0;
    for (int j = 0; j < 256; j++) {
        if (!present[j]) continue;
        if (skip) { skip--; continue; }
        *p++ = (uint8_t)j;
        if (j && present[j-1]) {
            int k = j + 1;
            while (k < 256 && present[k]) k++;
            skip = k - (j + 1);
            *p++ = (uint8_t)skip;
        }
    }
    *p++ = 0;
    return (int)(p - s);
}

// ------------------------------------------------------------------------------------------------ STRIPE transpose
// plane j of a section = bytes j, j+4, j+8 … (rANS_static4x16pr.c:1178-1190)
__global__ void k_stripe_transpose (const EncSection *secs, const Tile *tiles, uint32_t n_tiles)
{
    if (blockIdx.x >= n_tiles) return;
    const Tile t = tiles[blockIdx.x];
    const EncSection &S = secs[t.leaf];
    const uint32_t n = S.n;
    uint32_t idx[4], acc = 0;
    for (int j = 0; j < 4; j++) { idx[j] = acc; acc += n / 4 + ((n % 4) > (uint32_t)j); }
    uint32_t end = min (t.off + TILE, n);
    for (uint32_t i = t.off + threadIdx.x; i < end; i += blockDim.x)
        S.planes[idx[i & 3] + (i >> 2)] = S.in[i];
}

// ------------------------------------------------------------------------------------------------ order-0 histogram
// pass 0: over the leaf input; pass 1: over packbuf, only for leaves whose PACK succeeded
__global__ void __launch_bounds__(256) k_hist0 (const EncLeaf *leaves, const EncLeafDyn *dyn, const Tile *tiles, uint32_t n_tiles, int pass)
{
    if (blockIdx.x >= n_tiles) return;
    const Tile t = tiles[blockIdx.x];
    const EncLeaf &L = leaves[t.leaf];
    const uint8_t * __restrict__ src; uint32_t n;
    if (pass == 0) { src = L.in; n = L.n; }
    else { if (!dyn[t.leaf].packed) return; src = dyn[t.leaf].eff_in; n = dyn[t.leaf].eff_n; }
    if (t.off >= n) return;
    // One 256-bin histogram per warp in shared memory.  The streams of this path are low-entropy (one symbol is most of the
    // tile), and 32 lanes adding to the same shared counter serialise — so the tile's first byte is counted in registers
    // (4 bytes per SIMD compare) and only the other symbols go through shared-memory atomics.
    __shared__ uint32_t h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads ();
    const uint32_t end = min (t.off + TILE, n), len = end - t.off;
    const uint8_t *p = src + t.off;
    uint32_t *my = h[threadIdx.x >> 5];
    const uint32_t hot = p[0], hot4 = hot * 0x01010101u;
    uint32_t cnt_hot = 0;
    const uint32_t head = min (len, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15));
    const uint32_t nvec = (len - head) >> 4, tail0 = head + (nvec << 4);
    #define HIST_BYTE(B) { const uint32_t b_ = (B); if (b_ == hot) cnt_hot++; else atomicAdd (&my[b_], 1u); }
    if (threadIdx.x < head) HIST_BYTE (p[threadIdx.x])
    if (tail0 + threadIdx.x < len) HIST_BYTE (p[tail0 + threadIdx.x])          // fewer than 16 bytes
    const uint4 *pv = reinterpret_cast<const uint4 *>(p + head);
    for (uint32_t i = threadIdx.x; i < nvec; i += 256) {
        const uint4 v = __ldg (pv + i);
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (__vcmpeq4 (w[k], hot4) == 0xffffffffu) cnt_hot += 4;
            else { HIST_BYTE (w[k] & 0xffu) HIST_BYTE ((w[k] >> 8) & 0xffu) HIST_BYTE ((w[k] >> 16) & 0xffu) HIST_BYTE (w[k] >> 24) }
        }
    }
    #undef HIST_BYTE
    cnt_hot = __reduce_add_sync (0xffffffffu, cnt_hot);
    if ((threadIdx.x & 31) == 0 && cnt_hot) atomicAdd (&my[hot], cnt_hot);
    __syncthreads ();
    for (int i = threadIdx.x; i < 256; i += 256) {
        uint32_t v = 0;
        #pragma unroll
        for (int k = 0; k < 8; k++) v += h[k][i];
        if (v) atomicAdd (&L.hist0[i], v);
    }
}

// ------------------------------------------------------------------------------------------------ PACK decision + container header
// one thread per leaf.  hts_pack (pack.c:58-154) symbol census and the header bytes of rans_compress_to_4x16
// :1238-1278 / arith_compress_to :770-815.
__global__ void k_pack_decide (const EncLeaf *leaves, EncLeafDyn *dyn, uint32_t n_leaves)
{
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_leaves) return;
    const EncLeaf &L = leaves[li];
    EncLeafDyn &D = dyn[li];
    uint32_t flags = L.order_req, h = 1;
    D.eff_in = L.in; D.eff_n = L.n; D.packed = 0; D.cat = 0; D.per_byte = 1;
    D.tab_len = D.payload_len = 0;
    if (!(flags & F_NOSZ)) h += put_varint (D.hdr + 1, L.n);

    if ((flags & F_PACK) && L.n) {
        int ns = 0;
        for (int i = 0; i < 256; i++) if (L.hist0[i]) { D.code[i] = (uint8_t)ns; D.hdr[h + 1 + ns] = (uint8_t)i; ns++; }
        if (ns > 16 && ns != 256) flags &= ~F_PACK;                        // :1260-1264
        else if (ns == 256) {                                             // count byte wraps to 0, data stored verbatim
            D.hdr[h] = 0; h += 1;
            h += put_varint (D.hdr + h, L.n);
        }
        else {
            D.hdr[h] = (uint8_t)ns; h += ns + 1;
            int per = ns > 4 ? 2 : ns > 2 ? 4 : ns > 1 ? 8 : 0;
            uint32_t plen = per ? (L.n + per - 1) / per : 0;
            h += put_varint (D.hdr + h, plen);
            D.packed = 1; D.per_byte = (uint8_t)per;
            D.eff_in = L.packbuf; D.eff_n = plen;
            for (int i = 0; i < 256; i++) L.hist0[i] = 0;                  // re-counted over the packed bytes (pass 1)
        }
    }
    else if (flags & F_PACK) flags &= ~F_PACK;                             // :1276-1278
    if (L.coder == CODER_ARITH && (flags & F_RLE) && !D.eff_n) flags &= ~F_RLE;   // arith_dynamic.c:813-815
    D.hdr[0] = (uint8_t)flags;
    D.hdr_len = h;
}

__global__ void k_pack (const EncLeaf *leaves, const EncLeafDyn *dyn, const Tile *tiles, uint32_t n_tiles)
{
    if (blockIdx.x >= n_tiles) return;
    const Tile t = tiles[blockIdx.x];
    const EncLeafDyn &D = dyn[t.leaf];
    if (!D.packed || !D.per_byte) return;
    const EncLeaf &L = leaves[t.leaf];
    __shared__ uint8_t code[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) code[i] = D.code[i];
    __syncthreads ();
    const int per = D.per_byte, bits = 8 / per;
    uint32_t end = min (t.off + TILE, L.n);                                 // TILE is a multiple of 8
    for (uint32_t o = t.off / per + threadIdx.x; o * per < end; o += blockDim.x) {
        uint32_t b = 0, base = o * per;
        for (int k = 0; k < per && base + k < L.n; k++) b |= (uint32_t)code[L.in[base + k]] << (k * bits);
        L.packbuf[o] = (uint8_t)b;
    }
}

// ------------------------------------------------------------------------------------------------ leaf prep
// one CTA per leaf, after the final hist0: effective order (:1333-1336), symbol ranks, zero the O1 counters.
__global__ void k_leaf_prep (const EncLeaf *leaves, EncLeafDyn *dyn, uint32_t n_leaves, Arena arena)
{
    uint32_t li = blockIdx.x;
    if (li >= n_leaves) return;
    const EncLeaf &L = leaves[li];
    EncLeafDyn &D = dyn[li];
    __shared__ uint32_t s_nsym;
    __shared__ uint32_t *s_h1;
    uint32_t flags = D.hdr[0];
    uint32_t order = flags & (L.coder == CODER_ARITH ? 3u : 1u);
    if (order && D.eff_n < 8) { flags &= (L.coder == CODER_ARITH ? ~3u : ~1u); order = 0; }
    __syncthreads ();
    if (threadIdx.x == 0) {
        D.hdr[0] = (uint8_t)flags; D.eff_order = (uint8_t)order;
        int ns = 0;
        for (int i = 0; i < 256; i++)
            if (L.hist0[i] || (i == 0 && order && L.coder == CODER_RANS)) D.rank[i] = (uint8_t)ns++;   // O1: symbol 0 forced present (:741)
            else D.rank[i] = 0;
        D.nsym = (uint16_t)ns; s_nsym = ns;
        D.hist1 = nullptr; D.symtab = nullptr; D.ctxbytes = nullptr; D.models = nullptr; D.split_pos = nullptr; D.split_start = nullptr; D.split_rec = nullptr;
        if (L.coder == CODER_RANS && D.eff_n) {
            if (order) {
                D.hist1    = reinterpret_cast<uint32_t *>(arena.alloc ((unsigned long long)ns * ns * 4));
                D.symtab   = reinterpret_cast<EncSym *>(arena.alloc ((unsigned long long)ns * ns * sizeof (EncSym)));
                D.ctxbytes = arena.alloc ((unsigned long long)ns * CTXB);
                if (!D.hist1 || !D.symtab || !D.ctxbytes) { D.hist1 = nullptr; D.symtab = nullptr; D.ctxbytes = nullptr; }
            }
            else D.symtab = reinterpret_cast<EncSym *>(arena.alloc (256 * sizeof (EncSym)));
        }
        s_h1 = D.hist1;
    }
    __syncthreads ();
    if (s_h1) {
        uint32_t cells = s_nsym * s_nsym;
        for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) s_h1[i] = 0;
    }
}

// ------------------------------------------------------------------------------------------------ order-1 histogram
// hist1_4 (utils.h:136-210): F[prev][cur], prev = 0 before the first byte.  Compact [rank][rank] counters,
// accumulated in shared memory when nsym^2 fits, flushed with global atomics.
constexpr uint32_t H1_SMEM_CELLS = 10240;    // 40 KB of u32: nsym <= 101
__global__ void __launch_bounds__(256) k_hist1 (const EncLeaf *leaves, const EncLeafDyn *dyn, const Tile *tiles, uint32_t n_tiles)
{
    if (blockIdx.x >= n_tiles) return;
    const Tile t = tiles[blockIdx.x];
    const EncLeaf &L = leaves[t.leaf];
    const EncLeafDyn &D = dyn[t.leaf];
    if (L.coder != CODER_RANS || !D.eff_order || t.off >= D.eff_n || !D.hist1) return;
    __shared__ uint32_t h[H1_SMEM_CELLS];
    __shared__ uint8_t rank[256];
    const uint32_t ns = D.nsym, cells = ns * ns;
    const bool in_smem = cells <= H1_SMEM_CELLS;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) rank[i] = D.rank[i];
    if (in_smem) for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) h[i] = 0;
    __syncthreads ();
    const uint8_t * __restrict__ src = D.eff_in;
    const uint32_t end = min (t.off + TILE, D.eff_n), len = end - t.off;
    const uint8_t *p = src + t.off;
    // As in k_hist0: the pair (first byte of the tile, itself) — nearly every pair of a low-entropy stream such as the
    // ACGT exception stream — is counted in registers, 16 bytes per compare; everything else goes through the atomics.
    const uint32_t hot = p[0], hot4 = hot * 0x01010101u, hot_cell = rank[hot] * ns + rank[hot];
    uint32_t cnt_hot = 0;
    #define H1_PAIR(PB, B, IDX) { const uint32_t pb_ = (PB), b_ = (B); \
        if (pb_ == hot && b_ == hot && (IDX)) cnt_hot++; \
        else { const uint32_t cell_ = ((IDX) ? rank[pb_] : rank[0]) * ns + rank[b_]; if (in_smem) atomicAdd (&h[cell_], 1u); else atomicAdd (&D.hist1[cell_], 1u); } }
    const uint32_t head = min (len, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15));
    const uint32_t nvec = (len - head) >> 4, tail0 = head + (nvec << 4);
    if (threadIdx.x < head) { const uint32_t idx = t.off + threadIdx.x; H1_PAIR (idx ? src[idx - 1] : 0u, src[idx], idx) }
    if (tail0 + threadIdx.x < len) { const uint32_t idx = t.off + tail0 + threadIdx.x; H1_PAIR (idx ? src[idx - 1] : 0u, src[idx], idx) }
    const uint4 *pv = reinterpret_cast<const uint4 *>(p + head);
    for (uint32_t i = threadIdx.x; i < nvec; i += 256) {
        const uint4 v = __ldg (pv + i);
        const uint32_t idx0 = t.off + head + (i << 4);
        uint32_t pb = idx0 ? src[idx0 - 1] : 0u;
        if (idx0 && pb == hot && v.x == hot4 && v.y == hot4 && v.z == hot4 && v.w == hot4) { cnt_hot += 16; continue; }
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
        #pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t b = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
            H1_PAIR (pb, b, idx0 + k)
            pb = b;
        }
    }
    #undef H1_PAIR
    cnt_hot = __reduce_add_sync (0xffffffffu, cnt_hot);
    if ((threadIdx.x & 31) == 0 && cnt_hot) { if (in_smem) atomicAdd (&h[hot_cell], cnt_hot); else atomicAdd (&D.hist1[hot_cell], cnt_hot); }
    if (in_smem) {
        __syncthreads ();
        for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) if (h[i]) atomicAdd (&D.hist1[i], h[i]);
    }
}

// ------------------------------------------------------------------------------------------------ rANS chain encoder
// One leaf = 4 lanes = the 4 interleaved states.  Each lane runs its own state; the shared, backwards-growing
// output pointer of the reference (:453-456, :832-835: states are served 3,2,1,0 within a step) is reproduced with a
// 4-wide ballot: a lane that renormalises writes its 16-bit word at  end - used - 2*popc(emitters with index >= mine).
struct ChainIn {
    const uint8_t *in; const EncSym *tab; const uint8_t *rank; uint8_t *end;
    uint32_t n, nsym; bool valid, o1;
};

__device__ __forceinline__ void rans_step (uint32_t &x, uint32_t &used, uint8_t *end, bool act, const EncSym &e, int k, int gshift)
{
    bool emit = act && x >= e.x_max;
    uint32_t g = (__ballot_sync (0xffffffffu, emit) >> gshift) & 0xfu;
    if (emit) {
        uint32_t r = __popc (g >> k);
        *reinterpret_cast<uint16_t *>(end - used - 2 * r) = (uint16_t)x;
        x >>= 16;
    }
    used += 2 * __popc (g);
    if (act) {
        uint32_t q = __umulhi (x, e.rcp) >> (e.cmpl_sh >> 16);
        x = x + e.bias + q * (e.cmpl_sh & 0xffffu);
    }
}

// All 32 lanes of the warp must call this (ballots use the full mask).  Returns payload bytes (same in all 4 lanes of a group).
__device__ uint32_t rans_encode_warp (const ChainIn &c, int lane)
{
    const int k = lane & 3, gshift = lane & ~3;
    uint32_t x = RANS_L, used = 0;
    uint32_t steps = 0, q4 = c.n >> 2, r = c.n & 3;
    if (c.valid) steps = c.o1 ? q4 + r : (c.n + 3) >> 2;
    uint32_t maxsteps = steps;
    for (int o = 16; o; o >>= 1) maxsteps = max (maxsteps, __shfl_xor_sync (0xffffffffu, maxsteps, o));

    if (!c.o1) {
        // step s covers symbols base..base+3, base = 4*(steps-1-s); lane k takes base+k (symbol i belongs to state i&3, :439-477)
        for (uint32_t s0 = 0; s0 < maxsteps; s0 += 4) {
            EncSym e[4]; bool act[4];
            #pragma unroll
            for (int t = 0; t < 4; t++) {
                uint32_t s = s0 + t;
                act[t] = false;
                if (s < steps) {
                    uint32_t idx = 4 * (steps - 1 - s) + k;
                    if (idx < c.n) { act[t] = true; e[t] = c.tab[c.in[idx]]; }
                }
            }
            #pragma unroll
            for (int t = 0; t < 4; t++) rans_step (x, used, c.end, act[t], e[t], k, gshift);
        }
    }
    else {
        // lane k walks its quarter backwards: pos = pstart … k*q4, then one step in context 0 (:806-846).
        // chain 3 also owns the remainder, so lanes 0-2 join r steps later.
        const uint32_t pstart = (k == 3) ? c.n - 2 : (k + 1) * q4 - 2;
        const uint32_t len = c.valid ? ((k == 3) ? q4 + r : q4) : 0;
        const uint32_t delay = (k == 3) ? 0 : r;
        uint32_t l = 0;
        if (len) l = c.rank[c.in[pstart + 1]];
        for (uint32_t s0 = 0; s0 < maxsteps; s0 += 4) {
            EncSym e[4]; bool act[4];
            #pragma unroll
            for (int t = 0; t < 4; t++) {
                uint32_t s = s0 + t;
                act[t] = false;
                if (s >= delay && s - delay < len) {
                    uint32_t u = s - delay;
                    uint32_t cr = (u == len - 1) ? c.rank[0] : c.rank[c.in[pstart - u]];
                    act[t] = true;
                    e[t] = c.tab[cr * c.nsym + l];
                    l = cr;
                }
            }
            #pragma unroll
            for (int t = 0; t < 4; t++) rans_step (x, used, c.end, act[t], e[t], k, gshift);
        }
    }
    // RansEncFlush in order 3,2,1,0 (:479-482): state k lands at end-used-4*(4-k)
    if (c.valid) {
        uint16_t *w = reinterpret_cast<uint16_t *>(c.end - used - 4 * (4 - k));   // 2-byte aligned only
        w[0] = (uint16_t)x; w[1] = (uint16_t)(x >> 16);
    }
    return c.valid ? used + 16 : 0;
}

// ------------------------------------------------------------------------------------------------ frequency tables + encoder symbols
// one CTA (256 threads) per rANS leaf.
struct TabSmem {
    uint32_t F[256];
    uint32_t T[256];
    int      S[256];
    uint16_t sm10[256], sm12[256], ns[256];
    uint32_t len[256], off[256];
    uint8_t  present[256];
    uint8_t  symof[256];           // rank -> symbol
    EncSym   ntab[256];            // nested order-0 table (compressed O1 header)
    uint8_t  nbytes[800];          // nested frequency-table bytes
    uint32_t scan[8];
    int      shift;
    uint32_t tab_len, n_tab_len, n_payload;
};

// order-0 block front end (:405-432): counts (by symbol) -> table bytes at `out`, symbols into tab[256] (indexed by symbol)
__device__ uint32_t build_o0 (TabSmem &sm, uint32_t n, uint8_t *out, EncSym *tab)
{
    const int tid = threadIdx.x;
    __syncthreads ();
    if (tid == 0) {
        uint32_t tot = pow2_ceil (n); if (tot > 4096) tot = 4096;
        scale_freqs (sm.F, 256, (int)n, tot);
        uint8_t *p = out;
        p += put_alphabet (p, sm.F);
        for (int j = 0; j < 256; j++) if (sm.F[j]) p += put_varint (p, sm.F[j]);
        sm.tab_len = (uint32_t)(p - out);
        scale_freqs (sm.F, 256, (int)tot, 4096);
    }
    __syncthreads ();
    // exclusive prefix of F over the 256 symbols
    uint32_t f = sm.F[tid], v = f;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, v, o); if ((tid & 31) >= o) v += t; }
    if ((tid & 31) == 31) sm.scan[tid >> 5] = v;
    __syncthreads ();
    uint32_t base = 0;
    for (int w = 0; w < (tid >> 5); w++) base += sm.scan[w];
    if (f) tab[tid] = make_encsym (base + v - f, f, 12);
    __syncthreads ();
    return sm.tab_len;
}

__global__ void __launch_bounds__(256) k_tables (const EncLeaf *leaves, EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list)
{
    if (blockIdx.x >= n_list) return;
    const uint32_t li = list[blockIdx.x];
    const EncLeaf &L = leaves[li];
    EncLeafDyn &D = dyn[li];
    if (!D.eff_n || !D.symtab) return;
    __shared__ TabSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (!D.eff_order) {                                                   // ---- order 0 (:376-432)
        sm.F[tid] = L.hist0[tid];
        uint32_t tl = build_o0 (sm, D.eff_n, L.outbuf, D.symtab);
        if (tid == 0) D.tab_len = tl;
        return;
    }

    // ---- order 1 (:691-796)
    const uint32_t ns = D.nsym, n = D.eff_n, q4 = n >> 2;
    uint32_t *H = D.hist1;
    sm.present[tid] = (L.hist0[tid] || tid == 0) ? 1 : 0;
    if (sm.present[tid]) sm.symof[D.rank[tid]] = (uint8_t)tid;
    if (tid == 0) for (int q = 1; q < 4; q++) H[D.rank[0] * ns + D.rank[D.eff_in[q * q4]]]++;      // :730-733
    __syncthreads ();

    // per context: total, #symbols, and how many get bumped to 1 at 10 / 12 bits (:635-645)
    for (uint32_t r = warp; r < ns; r += 8) {
        uint32_t tot = 0;
        for (uint32_t j = lane; j < ns; j += 32) tot += H[r * ns + j];
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync (0xffffffffu, tot, o);
        uint32_t mv = pow2_ceil (tot), a = 0, b = 0, c = 0;
        for (uint32_t j = lane; j < ns; j += 32) {
            uint32_t h = H[r * ns + j];
            if (h) { c++; if (mv / h > 1024) a++; if (mv / h > 4096) b++; }
        }
        for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync (0xffffffffu, a, o); b += __shfl_xor_sync (0xffffffffu, b, o); c += __shfl_xor_sync (0xffffffffu, c, o); }
        if (lane == 0) { sm.T[r] = tot; sm.sm10[r] = (uint16_t)a; sm.sm12[r] = (uint16_t)b; sm.ns[r] = (uint16_t)c; }
    }
    __syncthreads ();

    // compute_shift (:626-687): the entropy sums are order-dependent double arithmetic; the reference object
    // evaluates  t = fma(d, K, -l);  e = fma(-F, t, e) + c  — reproduced exactly, serially, in (context, symbol) order.
    if (tid == 0) {
        const double K = 1.539095918623324e-16;
        double e10 = 0, e12 = 0;
        int max_tot = 0;
        for (uint32_t r = 0; r < ns; r++) {
            int mv = (int)pow2_ceil (sm.T[r]);
            double l10 = c_log_o1_10[sm.sm10[r]], l12 = c_log_o1_12[sm.sm12[r]];
            double Td = (double)sm.T[r];
            for (uint32_t j = 0; j < ns; j++) {
                uint32_t h = H[r * ns + j];
                if (!h) continue;
                double Fd = (double)h;
                int x = __double2int_rz (__ddiv_rn (__dmul_rn (Fd, 1024.0), Td)); if (x < 1) x = 1;
                double t = __fma_rn (__ll2double_rn (__double_as_longlong ((double)x) - 4606921278410026770LL), K, -l10);
                e10 = __dadd_rn (__fma_rn (-Fd, t, e10), 4.0);
                x = __double2int_rz (__ddiv_rn (__dmul_rn (Fd, 4096.0), Td)); if (x < 1) x = 1;
                t = __fma_rn (__ll2double_rn (__double_as_longlong ((double)x) - 4606921278410026770LL), K, -l12);
                e12 = __dadd_rn (__fma_rn (-Fd, t, e12), 6.0);
            }
            if (sm.ns[r] < 64 && mv > 128) mv /= 2;
            if (mv > 1024) mv /= 2;
            if (mv > 4096) mv = 4096;
            sm.S[r] = mv;
            if (max_tot < mv) max_tot = mv;
        }
        sm.shift = (__ddiv_rn (e10, e12) < 1.01 || max_tot <= 1024) ? 10 : 12;
    }
    __syncthreads ();
    const int shift = sm.shift;

    // per context (thread r): normalise, serialise, shift to 1<<shift, build encoder symbols (:750-776)
    if ((uint32_t)tid < ns) {
        const uint32_t r = tid;
        uint32_t *row = H + r * ns;
        int mv = sm.S[r];
        if (shift == 10 && mv > 1024) mv = 1024;
        scale_freqs (row, (int)ns, (int)sm.T[r], (uint32_t)mv);
        // encode_freq_d (:292-322): every rank is a symbol of F0; zero runs as (0, run-1)
        uint8_t *p0 = D.ctxbytes + r * CTXB, *p = p0;
        int zrun = 0;
        for (uint32_t j = 0; j < ns; j++) {
            uint32_t f = row[j];
            if (f) { if (zrun) { p -= zrun - 1; *p++ = (uint8_t)(zrun - 1); zrun = 0; } p += put_varint (p, f); }
            else { zrun++; *p++ = 0; }
        }
        if (zrun) { p -= zrun - 1; *p++ = (uint8_t)(zrun - 1); }
        sm.len[r] = (uint32_t)(p - p0);
        // normalise_freq_shift (:165-176) then symbols
        int sh = 0;
        if (mv != 0 && (uint32_t)mv != (1u << shift)) { uint32_t s = mv; while (s < (1u << shift)) { s *= 2; sh++; } }
        uint32_t x = 0;
        EncSym *trow = D.symtab + r * ns;
        for (uint32_t j = 0; j < ns; j++) {
            uint32_t f = row[j] << sh;
            trow[j] = make_encsym (x, f, (uint32_t)shift);
            x += f;
        }
    }
    __syncthreads ();
    if (tid == 0) {
        uint8_t *p = L.outbuf;
        *p++ = (uint8_t)(shift << 4);
        p += put_alphabet (p, sm.present);
        uint32_t o = (uint32_t)(p - L.outbuf);
        for (uint32_t r = 0; r < ns; r++) { sm.off[r] = o; o += sm.len[r]; }
        sm.tab_len = o;
    }
    __syncthreads ();
    if ((uint32_t)tid < ns) {
        const uint8_t *src = D.ctxbytes + tid * CTXB;
        uint8_t *dst = L.outbuf + sm.off[tid];
        for (uint32_t i = 0; i < sm.len[tid]; i++) dst[i] = src[i];
    }
    __syncthreads ();

    // optional order-0 compression of the table itself (:779-792)
    uint32_t tab_len = sm.tab_len;
    if (tab_len > 1000) {
        const uint32_t usz = tab_len - 1;
        const uint8_t *u = L.outbuf + 1;
        sm.F[tid] = 0;
        __syncthreads ();
        for (uint32_t i = tid; i < usz; i += 256) atomicAdd (&sm.F[u[i]], 1u);
        uint32_t ntab = build_o0 (sm, usz, sm.nbytes, sm.ntab);
        uint8_t *nend = L.outbuf + (L.out_cap & ~1u);
        if (warp == 0) {
            ChainIn c; c.valid = lane < 4; c.o1 = false; c.in = u; c.n = usz; c.nsym = 0; c.tab = sm.ntab; c.rank = nullptr; c.end = nend;
            uint32_t pl = rans_encode_warp (c, lane);
            if (lane == 0) sm.n_payload = pl;
        }
        __syncthreads ();
        const uint32_t csz = ntab + sm.n_payload;
        if (csz + 6 < tab_len) {
            __shared__ uint32_t s_hdr;
            if (tid == 0) {
                uint8_t *w = L.outbuf;
                *w++ |= 1;
                w += put_varint (w, usz);
                w += put_varint (w, csz);
                s_hdr = (uint32_t)(w - L.outbuf);
            }
            __syncthreads ();
            uint8_t *w = L.outbuf + s_hdr;
            for (uint32_t i = tid; i < ntab; i += 256) w[i] = sm.nbytes[i];
            const uint8_t *pay = nend - sm.n_payload;
            for (uint32_t i = tid; i < sm.n_payload; i += 256) w[ntab + i] = pay[i];
            tab_len = s_hdr + csz;
        }
        __syncthreads ();
    }
    if (tid == 0) { D.tab_len = tab_len; D.shift = (uint8_t)shift; }
}

// ------------------------------------------------------------------------------------------------ adaptive arithmetic coder
// Model layout per context (words): [0] TotFreq, [1] sentinel, [2 .. 2+maxs) entries (freq | symbol<<16), then a zero
// terminator — the reference's SIMPLE_MODEL (c_simple_model.h:77-103) restricted to the live entries.
// model initialisation for all arithmetic leaves: one CTA per leaf
__global__ void k_arith_init (const EncLeaf *leaves, EncLeafDyn *dyn, const uint32_t *list, uint32_t n_list, Arena arena, uint32_t split_min)
{
    if (blockIdx.x >= n_list) return;
    const uint32_t li = list[blockIdx.x];
    const EncLeaf &L = leaves[li];
    EncLeafDyn &D = dyn[li];
    __shared__ uint32_t s_max;
    __shared__ uint32_t *s_m;
    const uint32_t nctx = D.eff_order ? 256 : 1;
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (int i = 255; i >= 0; i--) if (L.hist0[i]) { m = i; break; }
        s_max = m + 1; D.nsym = (uint16_t)(m + 1);
        D.models = s_m = reinterpret_cast<uint32_t *>(arena.alloc (((unsigned long long)nctx * ar_stride (m + 1) + 258 * AR_RUN_STRIDE) * 4));
        if (s_m && D.eff_order && !(D.hdr[0] & F_RLE) && D.eff_n >= split_min) {       // long order-1 leaf: the split encoder (arith_split.cu)
            uint32_t *sp = reinterpret_cast<uint32_t *>(arena.alloc (4ull * D.eff_n));
            uint32_t *ss = reinterpret_cast<uint32_t *>(arena.alloc (4ull * (257 + 64)));     // start[257] + the contexts ranked by length (256 bytes)
            uint2    *sr = reinterpret_cast<uint2 *>(arena.alloc (8ull * D.eff_n));
            if (sp && ss && sr) { D.split_pos = sp; D.split_start = ss; D.split_rec = sr; }   // (else: the arena overflowed and the batch is replayed)
        }
    }
    __syncthreads ();
    if (!s_m) return;
    const uint32_t maxs = s_max, stride = ar_stride (maxs);
    uint32_t *lit = s_m;
    for (uint32_t c = threadIdx.x; c < nctx; c += blockDim.x) ar_model_init (lit + c * stride, maxs);
    if (D.hdr[0] & F_RLE) {
        uint32_t *run = lit + nctx * stride;
        for (uint32_t c = threadIdx.x; c < 258; c += blockDim.x) ar_model_init (run + c * AR_RUN_STRIDE, 4);
    }
}

// ------------------------------------------------------------------------------------------------ leaf / section finalisation
__global__ void k_leaf_final (const EncLeaf *leaves, EncLeafDyn *dyn, uint32_t n_leaves)
{
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_leaves) return;
    const EncLeaf &L = leaves[li];
    EncLeafDyn &D = dyn[li];
    uint32_t body = D.tab_len + D.payload_len;
    if (body >= D.eff_n) {                                                // :1343-1348 / arith_dynamic.c:847-852
        uint32_t f = D.hdr[0], nosz = L.order_req & F_NOSZ;
        f &= (L.coder == CODER_ARITH) ? ~(3u | F_EXT) : ~3u;
        f |= F_CAT | nosz;
        D.hdr[0] = (uint8_t)f;
        D.cat = 1;
        body = D.eff_n;
    }
    D.total_len = D.hdr_len + body;
}

constexpr int SEGS_PER_SECTION = 16;

__device__ int leaf_segments (const EncLeaf &L, const EncLeafDyn &D, uint8_t *dst, CopySeg *segs)
{
    int k = 0;
    segs[k++] = CopySeg { D.hdr, dst, D.hdr_len, 0 };
    dst += D.hdr_len;
    if (D.cat) segs[k++] = CopySeg { D.eff_in, dst, D.eff_n, 0 };
    else {
        segs[k++] = CopySeg { L.outbuf, dst, D.tab_len, 0 };
        segs[k++] = CopySeg { L.outbuf + (L.out_cap & ~1u) - D.payload_len, dst + D.tab_len, D.payload_len, 0 };
    }
    return k;
}

// one thread per section: STRIPE method selection (:1201-1222 / arith_dynamic.c:676-763) and the copy plan
__global__ void k_section_final (const EncSection *secs, const EncLeaf *leaves, const EncLeafDyn *dyn,
                                 SectionResult *res, CopySeg *segs, uint8_t *stripe_hdr, uint32_t n_secs)
{
    uint32_t si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= n_secs) return;
    const EncSection &S = secs[si];
    CopySeg *sg = segs + (size_t)si * SEGS_PER_SECTION;
    for (int i = 0; i < SEGS_PER_SECTION; i++) sg[i].len = 0;
    if (S.soft_fail) { res[si].out_len = 0; res[si].status = 1; return; }
    if (!S.stripe) {
        const uint32_t li = S.first_leaf;
        leaf_segments (leaves[li], dyn[li], S.out, sg);
        res[si].out_len = dyn[li].total_len; res[si].status = 0;
        return;
    }
    // STRIPE: header = flags (NOSZ cleared), varint n, N, then varint clen per plane; candidates of a plane are consecutive
    // leaves in the reference's try-order; the smallest wins, the first on ties (strict '>').
    uint8_t *h = stripe_hdr + (size_t)si * 32;
    uint32_t hl = 0;
    h[hl++] = (uint8_t)(S.order & ~F_NOSZ);
    hl += put_varint (h + hl, S.n);
    h[hl++] = 4;
    uint32_t chosen[4], li = S.first_leaf, per_plane = S.n_leaves;         // n_leaves packs the 4 candidate counts, 4 bits each
    for (int p = 0; p < 4; p++) {
        uint32_t nc = (per_plane >> (4 * p)) & 0xf, best = li;
        uint32_t best_sz = (S.coder == CODER_RANS) ? S.n + 10 : 0x7fffffffu;
        for (uint32_t c = 0; c < nc; c++)
            if (best_sz > dyn[li + c].total_len) { best_sz = dyn[li + c].total_len; best = li + c; }
        chosen[p] = best;
        hl += put_varint (h + hl, dyn[best].total_len);
        li += nc;
    }
    int k = 0;
    uint8_t *dst = S.out;
    sg[k++] = CopySeg { h, dst, hl, 0 };
    dst += hl;
    for (int p = 0; p < 4; p++) {
        k += leaf_segments (leaves[chosen[p]], dyn[chosen[p]], dst, sg + k);
        dst += dyn[chosen[p]].total_len;
    }
    res[si].out_len = (uint32_t)(dst - S.out); res[si].status = 0;
}

// packed output: the sections of a batch are appended to one buffer in section order, as zfile_compress_local_data appends
// a section to vb->z_data (src/zfile.c:229-262).  One CTA: offsets = exclusive prefix of the 16-byte aligned section lengths;
// the copy plan's destinations (relative to their section while S.out is NULL) are then moved into the buffer.  If the batch
// does not fit, nothing is copied and the host reports the size that is needed.
__global__ void __launch_bounds__(1024) k_pack_place (const SectionResult *res, CopySeg *segs, unsigned long long *off, uint32_t n_secs,
                                                      uint8_t *arena, unsigned long long cap)
{
    __shared__ unsigned long long sm[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long acc = 0;
    for (uint32_t base = 0; base < n_secs; base += 1024) {
        const uint32_t si = base + threadIdx.x;
        const unsigned long long v = si < n_secs && res[si].status == 0 ? ((unsigned long long)res[si].out_len + 15ull) & ~15ull : 0;
        unsigned long long inc = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sm[warp] = inc;
        __syncthreads ();
        if (warp == 0) {
            const unsigned long long x = sm[lane]; unsigned long long xi = x;
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync (0xffffffffu, xi, o); if (lane >= o) xi += t; }
            sm[lane] = xi - x;
            if (lane == 31) sm[32] = xi;
        }
        __syncthreads ();
        if (si < n_secs) off[si] = acc + sm[warp] + inc - v;
        acc += sm[32];
        __syncthreads ();
    }
    if (threadIdx.x == 0) off[n_secs] = acc;
    const bool fits = acc <= cap;
    for (uint32_t k = threadIdx.x; k < n_secs * SEGS_PER_SECTION; k += 1024) {
        if (!fits) segs[k].len = 0;
        else if (segs[k].len) segs[k].dst = arena + off[k / SEGS_PER_SECTION] + (size_t)segs[k].dst;
    }
}

// grid (n_segs, parts): each segment is split into gridDim.y parts
__global__ void k_copy_segs (const CopySeg *segs, uint32_t n_segs)
{
    if (blockIdx.x >= n_segs) return;
    const CopySeg s = segs[blockIdx.x];
    if (!s.len) return;
    uint32_t part = (s.len + gridDim.y - 1) / gridDim.y;
    part = (part + 15) & ~15u;
    uint32_t b = blockIdx.y * part, e = min (b + part, s.len);
    for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) s.dst[i] = s.src[i];
}

// ------------------------------------------------------------------------------------------------ launcher
#define LAUNCH(kern, grid, block, ...) do { if ((grid) > 0) { kern<<<(grid), (block), 0, st>>>(__VA_ARGS__); P.launches++; } } while (0)

void enc_run (EncPlanDev &P, cudaStream_t st)
{
    const uint32_t nl = P.n_leaves, ns = P.n_sections;
    if (!ns) return;
    if (P.n_stripe_tiles) LAUNCH (k_stripe_transpose, P.n_stripe_tiles, 256, P.sections, P.stripe_tiles, P.n_stripe_tiles);
    LAUNCH (k_hist0, P.n_tiles, 256, P.leaves, P.dyn, P.tiles, P.n_tiles, 0);
    LAUNCH (k_pack_decide, (nl + 127) / 128, 128, P.leaves, P.dyn, nl);
    if (P.any_pack) {
        LAUNCH (k_pack, P.n_tiles, 256, P.leaves, P.dyn, P.tiles, P.n_tiles);
        LAUNCH (k_hist0, P.n_tiles, 256, P.leaves, P.dyn, P.tiles, P.n_tiles, 1);
    }
    LAUNCH (k_leaf_prep, nl, 256, P.leaves, P.dyn, nl, P.arena);
    if (P.n_rans) {
        if (P.any_o1) LAUNCH (k_hist1, P.n_tiles, 256, P.leaves, P.dyn, P.tiles, P.n_tiles);
        LAUNCH (k_tables, P.n_rans, 256, P.leaves, P.dyn, P.rans_list, P.n_rans);
    }
    if (P.n_arith) LAUNCH (k_arith_init, P.n_arith, 256, P.leaves, P.dyn, P.arith_list, P.n_arith, P.arena, P.split_min);
    // The rANS and the arithmetic leaves are independent and both kernels are latency-bound (a handful of warps per SM),
    // so they run concurrently: the arithmetic kernel is forked onto the engine's second stream and joined afterwards.
    cudaEventRecord (P.ev_chain0, st);
    if (P.n_arith) {
        cudaStreamWaitEvent (P.st2, P.ev_chain0, 0);
        cudaEventRecord (P.ev_arith0, P.st2);
        // the split encoder's three kernels (the long order-1 leaves) go first and, unless switched off, on a stream of their own:
        // behind the general kernel on one stream they would wait for its last leaf before they start
        const bool own = chain_tune ().split_stream && P.st4;
        if (P.n_arith_big && own) {
            cudaStreamWaitEvent (P.st4, P.ev_chain0, 0);
            launch_arith_encode_split (P, P.st4); P.launches += 3;
            cudaEventRecord (P.ev_split, P.st4);
        }
        launch_arith_encode (P, P.st2); P.launches++;
        if (P.n_arith_big && !own) { launch_arith_encode_split (P, P.st2); P.launches += 3; cudaEventRecord (P.ev_split, P.st2); }
        cudaEventRecord (P.ev_chain2, P.st2);
        cudaStreamWaitEvent (P.st3, P.ev_chain0, 0);
        launch_arith_encode_o0 (P, P.st3); P.launches++;
        cudaEventRecord (P.ev_o0, P.st3);
    }
    if (P.n_rans_jobs) { launch_rans_encode (P, st); P.launches++; }
    cudaEventRecord (P.ev_chain1, st);
    if (P.n_arith) { cudaStreamWaitEvent (st, P.ev_chain2, 0); cudaStreamWaitEvent (st, P.ev_o0, 0); if (P.n_arith_big) cudaStreamWaitEvent (st, P.ev_split, 0); }
    LAUNCH (k_leaf_final, (nl + 127) / 128, 128, P.leaves, P.dyn, nl);
    LAUNCH (k_section_final, (ns + 127) / 128, 128, P.sections, P.leaves, P.dyn, P.results, P.segs, P.stripe_hdr, ns);
    if (P.pack_off) LAUNCH (k_pack_place, 1, 1024, P.results, P.segs, P.pack_off, ns, P.pack_arena, P.pack_cap);
    dim3 g (ns * SEGS_PER_SECTION, P.copy_parts);
    k_copy_segs<<<g, 256, 0, st>>>(P.segs, ns * SEGS_PER_SECTION); P.launches++;
}

} // namespace gzb
