// b250.cu — b250_zip_generate (reference src/b250.c:202-297) for a batch of contexts: the conversion of a context's b250 buffer from
// the form the segmenter appends (little-endian variable-length words whose TYPE sits in the LAST byte, so that the segmenter can
// take the last word back, :137-182) to the form PIZ reads forwards (big endian, type first, :299-327), with, on the way,
//   - the node indices of words new to this VBlock turned into word indices (node_index_to_word_index, src/context.h:109-112),
//   - a word that is its predecessor + 1 replaced by the one-byte ONE_UP (only in dictionaries above 1024 words, :247,265-266).
// The reference walks the buffer backwards, one word at a time, in place.  Here:
//   k_b250_parse    a chunk of 64 bytes can be entered by the backward walk at 4 positions only (a word is 1..4 bytes): for each, where the
//                   walk leaves the chunk and how many words it saw — every chunk of every context at once, nothing depends on anything
//   k_b250_compose  one thread per context chains the chunks' entry -> exit maps from the end of the buffer (a few thousand steps)
//   k_b250_sizes    every chunk again, from its real entry: the converted value of each word (its predecessor is one more step of the
//                   walk), the bytes it will take
//   k_b250_offsets  one thread per context: where each chunk's output ends
//   k_b250_write    every chunk a third time: the words in PIZ format, right-aligned in `out` like the reference's result in its buffer
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "engine.h"

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr uint32_t B2_CH = 64;                        // bytes per chunk; chunks are counted from the END of the buffer
constexpr int32_t WI_ONE_UP = -2, WI_EMPTY = -3, WI_MISSING = -4;                 // src/context.h:25-28
constexpr int32_t VARL_MIN_2B = 127, VARL_MAX_2B = 127 + (1 << 14) - 1 - 2, VARL_MIN_3B = VARL_MAX_2B + 1, VARL_MAX_3B = VARL_MIN_3B + (1 << 21) - 1,
                  VARL_MAX_4B = (1 << 29) - 1;                                      // src/b250.c:31-39

struct B2Item {
    const uint8_t *in; uint8_t *out; const int32_t *ni2wi;
    unsigned long long len, first_chunk, n_chunks;
    uint32_t n_new, ol_len, one_up_ok;
    uint32_t *info;          // [0] error, [1] n_words lo, [2] n_words hi, [3..4] out_len
};
struct B2Chunk { uint8_t exit[4]; uint8_t cnt[4]; };  // per entry position j (the walk enters at the chunk's end - 1 - j): exit position in the chunk below, words seen
struct B2Real  { uint32_t entry; uint32_t out_bytes; unsigned long long out_end; };   // the real entry; bytes written by the chunk; bytes written by all the chunks after it (towards the end)

__device__ __forceinline__ uint32_t varl_bytes (uint32_t msb) { return (msb >> 7) == 0 ? 1 : (msb >> 6) == 2 ? 2 : (msb >> 5) == 6 ? 3 : 4; }   // VARL_BYTES :47
// b250_seg_get_wi (:64-86): the word whose last byte is at p
__device__ __forceinline__ int32_t seg_get_wi (const uint8_t *b, long long p, uint32_t L)
{
    if (L == 1) return b[p];
    if (L == 2) { const uint32_t w = b[p - 1] | ((uint32_t)b[p] << 8); return w == 0xBFFE ? WI_EMPTY : w == 0xBFFF ? WI_MISSING : (int32_t)(w & 0x3fff) + VARL_MIN_2B; }
    if (L == 3) return (int32_t)((b[p - 2] | ((uint32_t)b[p - 1] << 8) | ((uint32_t)b[p] << 16)) & 0x1fffff) + VARL_MIN_3B;
    return (int32_t)((b[p - 3] | ((uint32_t)b[p - 2] << 8) | ((uint32_t)b[p - 1] << 16) | ((uint32_t)b[p] << 24)) & 0x1fffffff);
}
__device__ __forceinline__ int32_t converted_wi (const B2Item &I, const uint8_t *b, long long p, uint32_t L)     // get_converted_wi :184-196
{
    const int32_t wi = seg_get_wi (b, p, L);
    if (wi >= (int32_t)I.ol_len) { const uint32_t k = (uint32_t)wi - I.ol_len; return k < I.n_new ? I.ni2wi[k] : -100; }   // (-100: a node index outside the VBlock's nodes)
    return wi;
}
// b250_set_wi (:89-121) in PIZ format: the encoding and its length; 0 = not encodable (the reference aborts)
__device__ __forceinline__ uint32_t piz_enc (int32_t wi, uint32_t &enc)
{
    if (wi >= 0 && wi <= 126)                  { enc = (uint32_t)wi; return 1; }
    if (wi >= VARL_MIN_2B && wi <= VARL_MAX_2B) { enc = (2u << 14) | (uint32_t)(wi - VARL_MIN_2B); return 2; }
    if (wi >= VARL_MIN_3B && wi <= VARL_MAX_3B) { enc = (6u << 21) | (uint32_t)(wi - VARL_MIN_3B); return 3; }
    if (wi > VARL_MAX_3B && wi <= VARL_MAX_4B)  { enc = (7u << 29) | (uint32_t)wi; return 4; }
    if (wi == WI_ONE_UP)  { enc = 127; return 1; }
    if (wi == WI_EMPTY)   { enc = 0xBFFE; return 2; }
    if (wi == WI_MISSING) { enc = 0xBFFF; return 2; }
    enc = 0; return 0;
}
// chunk c of an item covers [lo, hi): hi = len - c * 64
__device__ __forceinline__ void chunk_range (const B2Item &I, unsigned long long c, long long &lo, long long &hi)
{
    hi = (long long)I.len - (long long)(c * B2_CH); lo = hi > (long long)B2_CH ? hi - B2_CH : 0;
}

__global__ void __launch_bounds__(128) k_b250_parse (const B2Item *items, const uint32_t *chunk_item, B2Chunk *chunks, unsigned long long n_chunks)
{
    const unsigned long long g = (unsigned long long)blockIdx.x * 128 + threadIdx.x;
    if (g >= n_chunks) return;
    const B2Item &I = items[chunk_item[g >> 6]];                            // chunks are listed in groups of 64
    const unsigned long long c = g - I.first_chunk;
    if (c >= I.n_chunks) return;
    long long lo, hi; chunk_range (I, c, lo, hi);
    B2Chunk r;
    for (int j = 0; j < 4; j++) {
        long long p = hi - 1 - j; uint32_t n = 0;
        while (p >= lo) { p -= varl_bytes (I.in[p]); n++; }
        r.exit[j] = (uint8_t)(p < lo - 4 ? 4 : lo - 1 - p);                 // 0..3; the walk can leave the buffer's start at -1 only (exit 0 with lo = 0), checked by the composer
        r.cnt[j] = (uint8_t)n;
    }
    chunks[g] = r;
}

__global__ void k_b250_compose (const B2Item *items, const B2Chunk *chunks, B2Real *real, uint32_t n_items)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const B2Item &I = items[i];
    uint32_t entry = 0; unsigned long long words = 0;
    for (unsigned long long c = 0; c < I.n_chunks; c++) {
        const B2Chunk k = chunks[I.first_chunk + c];
        real[I.first_chunk + c].entry = entry;
        long long lo, hi; chunk_range (I, c, lo, hi);
        if (hi - 1 - (long long)entry < lo) { entry -= (uint32_t)(hi - lo); continue; }   // a chunk shorter than the entry offset (the first bytes of the buffer): the walk passes over it
        words += k.cnt[entry];
        entry = k.exit[entry];
    }
    if (entry != 0) I.info[0] = 1;                                          // "src in backward scan exceeded start of b250 array" (:273)
    I.info[1] = (uint32_t)words; I.info[2] = (uint32_t)(words >> 32);
}

// the walk over one chunk from its real entry.  f (p, L, value, prev_value): value = the converted word index of the word ending at p,
// prev_value = that of the word before it in the buffer (WORD_INDEX_NONE = -1 for the first word)
template <typename F> __device__ __forceinline__ void b250_walk (const B2Item &I, unsigned long long c, uint32_t entry, F f)
{
    long long lo, hi; chunk_range (I, c, lo, hi);
    long long p = hi - 1 - (long long)entry;
    if (p < lo) return;
    uint32_t L = varl_bytes (I.in[p]);
    if (p - (long long)L + 1 < 0) { I.info[0] = 1; return; }
    int32_t v = converted_wi (I, I.in, p, L);
    while (p >= lo) {
        const long long q = p - L;                                           // the end of the word before
        int32_t pv = -1; uint32_t pL = 0;
        if (q >= 0) {
            pL = varl_bytes (I.in[q]);
            if (q - (long long)pL + 1 < 0) { I.info[0] = 1; return; }
            pv = converted_wi (I, I.in, q, pL);
        }
        int32_t out = v;
        if (I.one_up_ok && pv >= 0 && v >= 0 && v == pv + 1) out = WI_ONE_UP;       // :265-266
        f (p, L, out);
        p = q; L = pL; v = pv;
    }
}

__global__ void __launch_bounds__(128) k_b250_sizes (const B2Item *items, const uint32_t *chunk_item, B2Real *real, unsigned long long n_chunks)
{
    const unsigned long long g = (unsigned long long)blockIdx.x * 128 + threadIdx.x;
    if (g >= n_chunks) return;
    const B2Item &I = items[chunk_item[g >> 6]];
    const unsigned long long c = g - I.first_chunk;
    if (c >= I.n_chunks || I.info[0]) return;
    uint32_t bytes = 0; bool bad = false;
    b250_walk (I, c, real[g].entry, [&] (long long, uint32_t, int32_t out) { uint32_t enc; const uint32_t n = piz_enc (out, enc); if (!n) bad = true; bytes += n; });
    if (bad) I.info[0] = 2;                                                  // "wi ∉ [-4..-2,0..2^29)" (:108)
    real[g].out_bytes = bytes;
}

__global__ void k_b250_offsets (const B2Item *items, B2Real *real, uint32_t n_items)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const B2Item &I = items[i];
    if (I.info[0]) return;
    unsigned long long at = 0;
    for (unsigned long long c = 0; c < I.n_chunks; c++) { real[I.first_chunk + c].out_end = at; at += real[I.first_chunk + c].out_bytes; }
    I.info[3] = (uint32_t)at; I.info[4] = (uint32_t)(at >> 32);
}

__global__ void __launch_bounds__(128) k_b250_write (const B2Item *items, const uint32_t *chunk_item, const B2Real *real, unsigned long long n_chunks)
{
    const unsigned long long g = (unsigned long long)blockIdx.x * 128 + threadIdx.x;
    if (g >= n_chunks) return;
    const B2Item &I = items[chunk_item[g >> 6]];
    const unsigned long long c = g - I.first_chunk;
    if (c >= I.n_chunks || I.info[0]) return;
    uint8_t *dst = I.out + I.len - real[g].out_end;                          // one past the last byte this chunk writes
    b250_walk (I, c, real[g].entry, [&] (long long, uint32_t, int32_t out) {
        uint32_t enc; const uint32_t n = piz_enc (out, enc);
        dst -= n;
        for (uint32_t k = 0; k < n; k++) dst[k] = (uint8_t)(enc >> (8 * (n - 1 - k)));   // big endian: the type first (:113-116)
    });
}

} // namespace

extern "C" int gzb_b250_generate_batch (gzb_engine *e, gzb_b250_item *items, uint32_t n, uint32_t flags)
{
    if (!e || (!items && n)) return GZB_E_BADARG;
    if (!n) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    std::vector<B2Item> h (n);
    std::vector<uint32_t> chunk_item;                                       // one entry per group of 64 chunks
    unsigned long long n_chunks = 0; size_t data_bytes = 0;
    auto al = [] (size_t x) { return (x + 255) & ~(size_t)255; };
    for (uint32_t i = 0; i < n; i++) {
        gzb_b250_item &it = items[i];
        it.status = GZB_OK; it.out_len = 0; it.n_words = 0;
        if ((it.len && (!it.b250 || !it.out)) || (it.n_new && !it.ni2wi)) { it.status = GZB_E_BADARG; e->err = "bad b250 item"; return GZB_E_BADARG; }
        h[i] = B2Item ();
        h[i].len = it.len; h[i].n_new = it.n_new; h[i].ol_len = it.ol_len; h[i].one_up_ok = it.one_up_ok;
        h[i].n_chunks = (it.len + B2_CH - 1) / B2_CH;
        n_chunks = (n_chunks + 63) & ~63ull; h[i].first_chunk = n_chunks;
        chunk_item.insert (chunk_item.end (), (size_t)((h[i].n_chunks + 63) / 64), i);
        n_chunks += h[i].n_chunks;
        if (!devptr) data_bytes += 2 * al (it.len + 16) + al ((size_t)it.n_new * 4 + 16);
    }
    const unsigned long long padded = (n_chunks + 63) & ~63ull;
    const size_t o_ci = al (n * sizeof (B2Item)), o_info = o_ci + al ((chunk_item.size () + 1) * 4), o_chunks = o_info + al ((size_t)n * 32),
                 o_real = o_chunks + al ((size_t)padded * sizeof (B2Chunk)), o_data = o_real + al ((size_t)padded * sizeof (B2Real));
    int rc = engine_reserve (e, o_data + data_bytes, o_info + (size_t)n * 32 + 256); if (rc) return rc;
    size_t cur = o_data;
    for (uint32_t i = 0; i < n; i++) {
        gzb_b250_item &it = items[i];
        h[i].info = reinterpret_cast<uint32_t *>(e->ws + o_info) + 8 * (size_t)i;
        if (devptr) { h[i].in = (const uint8_t *)it.b250; h[i].out = (uint8_t *)it.out; h[i].ni2wi = it.ni2wi; continue; }
        h[i].in = e->ws + cur; if (it.len) CK (cudaMemcpyAsync (e->ws + cur, it.b250, it.len, cudaMemcpyHostToDevice, st)); cur += al (it.len + 16);
        h[i].out = e->ws + cur; cur += al (it.len + 16);
        h[i].ni2wi = reinterpret_cast<const int32_t *>(e->ws + cur); if (it.n_new) CK (cudaMemcpyAsync (e->ws + cur, it.ni2wi, (size_t)it.n_new * 4, cudaMemcpyHostToDevice, st)); cur += al ((size_t)it.n_new * 4 + 16);
    }
    memcpy (e->pin, h.data (), n * sizeof (B2Item));
    memcpy (e->pin + o_ci, chunk_item.data (), chunk_item.size () * 4);
    CK (cudaMemcpyAsync (e->ws, e->pin, n * sizeof (B2Item), cudaMemcpyHostToDevice, st));
    if (!chunk_item.empty ()) CK (cudaMemcpyAsync (e->ws + o_ci, e->pin + o_ci, chunk_item.size () * 4, cudaMemcpyHostToDevice, st));
    CK (cudaMemsetAsync (e->ws + o_info, 0, (size_t)n * 32, st));
    const B2Item *d_items = reinterpret_cast<const B2Item *>(e->ws); const uint32_t *d_ci = reinterpret_cast<const uint32_t *>(e->ws + o_ci);
    B2Chunk *d_chunks = reinterpret_cast<B2Chunk *>(e->ws + o_chunks); B2Real *d_real = reinterpret_cast<B2Real *>(e->ws + o_real);
    if (padded) {
        const uint32_t gc = (uint32_t)((padded + 127) / 128), gi = (n + 127) / 128;
        k_b250_parse<<<gc, 128, 0, st>>>(d_items, d_ci, d_chunks, padded);
        k_b250_compose<<<gi, 128, 0, st>>>(d_items, d_chunks, d_real, n);
        k_b250_sizes<<<gc, 128, 0, st>>>(d_items, d_ci, d_real, padded);
        k_b250_offsets<<<gi, 128, 0, st>>>(d_items, d_real, n);
        k_b250_write<<<gc, 128, 0, st>>>(d_items, d_ci, d_real, padded);
        e->launches += 5;
    }
    uint32_t *info = reinterpret_cast<uint32_t *>(e->pin + o_info);
    CK (cudaMemcpyAsync (info, e->ws + o_info, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    int ret = GZB_OK;
    for (uint32_t i = 0; i < n; i++) {
        gzb_b250_item &it = items[i]; const uint32_t *f = info + 8 * (size_t)i;
        if (f[0]) { it.status = GZB_E_CORRUPT; ret = GZB_E_CORRUPT; e->err = f[0] == 1 ? "b250: the backward scan does not end at the start of the buffer" : "b250: a word index that cannot be encoded"; continue; }
        it.n_words = f[1] | ((uint64_t)f[2] << 32); it.out_len = f[3] | ((uint64_t)f[4] << 32);
        if (!devptr && it.out_len) CK (cudaMemcpyAsync ((uint8_t *)it.out + it.len - it.out_len, h[i].out + it.len - it.out_len, it.out_len, cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    return ret;
}
