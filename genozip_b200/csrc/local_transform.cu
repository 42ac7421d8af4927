// local_transform.cu — zip_generate_local's in-place transforms of a context's `local` buffer, and their PIZ inverses, for a batch of
// buffers that are already in HBM (reference src/zip.c:167-213; the loops are src/buffer.c:337-353 and :431-468, the arithmetic
// the macros INTERLACE / DEINTERLACE of src/context.h:98-101):
//   LT_UINT16/32/64 (and hex / float types)   BGEN_u*_buf: byte swap on a little-endian host (its own inverse)
//   LT_INT8/16/32/64                          interlace_d8_buf / BGEN_interlace_d*_buf: n >= 0 -> 2n, n < 0 -> -2n - 1, then big endian
//                                             BGEN_deinterlace_d*_buf (PIZ, local_type.h:76-82): the inverse
// Pure bandwidth work on data the codec path reads next (N read + N written, in place): one pass instead of a host pass per section.
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gzb200.h"
#include "engine.h"
#include "gzb_internal.cuh"

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr uint32_t LT_CHUNK = 1u << 16;         // elements per CTA

struct LtItem { void *data; unsigned long long n; int op; unsigned long long first_chunk; };

__device__ __forceinline__ uint16_t bswap (uint16_t v) { return (uint16_t)((v << 8) | (v >> 8)); }
__device__ __forceinline__ uint32_t bswap (uint32_t v) { return __byte_perm (v, 0, 0x0123); }
__device__ __forceinline__ unsigned long long bswap (unsigned long long v) { return ((unsigned long long)bswap ((uint32_t)v) << 32) | bswap ((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint8_t bswap (uint8_t v) { return v; }

// INTERLACE (src/context.h:99): (n < 0) ? ((unsigned)(-n) << 1) - 1 : (unsigned)n << 1, in the width of the type (-(-128) wraps to 128 in 8 bits)
template <typename U, typename S> __device__ __forceinline__ U interlace (U raw) { const S n = (S)raw; return n < 0 ? (U)(((U)(0 - (U)n) << 1) - 1) : (U)((U)n << 1); }
// DEINTERLACE (:100): (u & 1) ? -((u >> 1) + 1) : u >> 1
template <typename U, typename S> __device__ __forceinline__ U deinterlace (U u) { return (u & 1) ? (U)(0 - (U)((u >> 1) + 1)) : (U)(u >> 1); }

template <typename U, typename S> __device__ __forceinline__ void lt_apply (U *p, unsigned long long b, unsigned long long en, int kind)
{
    for (unsigned long long i = b + threadIdx.x; i < en; i += blockDim.x) {
        const U v = p[i];
        p[i] = kind == 0 ? bswap (v) : kind == 1 ? bswap (interlace<U, S> (v)) : deinterlace<U, S> (bswap (v));
    }
}

__global__ void __launch_bounds__(256) k_local_transform (const LtItem *items, const uint32_t *chunk_item)
{
    const LtItem it = items[chunk_item[blockIdx.x]];
    const unsigned long long b = ((unsigned long long)blockIdx.x - it.first_chunk) * LT_CHUNK, en = min (it.n, b + LT_CHUNK);
    switch (it.op) {
        case GZB_LT_SWAP16:        lt_apply<uint16_t, int16_t> ((uint16_t *)it.data, b, en, 0); break;
        case GZB_LT_SWAP32:        lt_apply<uint32_t, int32_t> ((uint32_t *)it.data, b, en, 0); break;
        case GZB_LT_SWAP64:        lt_apply<unsigned long long, long long> ((unsigned long long *)it.data, b, en, 0); break;
        case GZB_LT_INTERLACE8:    lt_apply<uint8_t, int8_t> ((uint8_t *)it.data, b, en, 1); break;
        case GZB_LT_INTERLACE16:   lt_apply<uint16_t, int16_t> ((uint16_t *)it.data, b, en, 1); break;
        case GZB_LT_INTERLACE32:   lt_apply<uint32_t, int32_t> ((uint32_t *)it.data, b, en, 1); break;
        case GZB_LT_INTERLACE64:   lt_apply<unsigned long long, long long> ((unsigned long long *)it.data, b, en, 1); break;
        case GZB_LT_DEINTERLACE8:  lt_apply<uint8_t, int8_t> ((uint8_t *)it.data, b, en, 2); break;
        case GZB_LT_DEINTERLACE16: lt_apply<uint16_t, int16_t> ((uint16_t *)it.data, b, en, 2); break;
        case GZB_LT_DEINTERLACE32: lt_apply<uint32_t, int32_t> ((uint32_t *)it.data, b, en, 2); break;
        case GZB_LT_DEINTERLACE64: lt_apply<unsigned long long, long long> ((unsigned long long *)it.data, b, en, 2); break;
        default: break;
    }
}

uint32_t lt_width (int op)
{
    switch (op) {
        case GZB_LT_INTERLACE8: case GZB_LT_DEINTERLACE8: return 1;
        case GZB_LT_SWAP16: case GZB_LT_INTERLACE16: case GZB_LT_DEINTERLACE16: return 2;
        case GZB_LT_SWAP32: case GZB_LT_INTERLACE32: case GZB_LT_DEINTERLACE32: return 4;
        case GZB_LT_SWAP64: case GZB_LT_INTERLACE64: case GZB_LT_DEINTERLACE64: return 8;
        default: return 0;
    }
}

} // namespace

extern "C" int gzb_local_transform_batch (gzb_engine *e, gzb_local_item *items, uint32_t n, uint32_t flags)
{
    if (!e || (!items && n)) return GZB_E_BADARG;
    if (!n) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    std::vector<LtItem> h (n);
    std::vector<uint32_t> chunk_item;
    size_t host_bytes = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t w = lt_width (items[i].op);
        if (!w || (!items[i].data && items[i].n_elems) || ((uintptr_t)items[i].data & (w - 1))) { items[i].status = GZB_E_BADARG; e->err = "bad local transform item"; return GZB_E_BADARG; }
        items[i].status = GZB_OK;
        h[i].n = items[i].n_elems; h[i].op = items[i].op; h[i].first_chunk = chunk_item.size ();
        chunk_item.insert (chunk_item.end (), (size_t)((items[i].n_elems + LT_CHUNK - 1) / LT_CHUNK), i);
        if (!devptr) host_bytes += (items[i].n_elems * w + 255) & ~(size_t)255;
    }
    const size_t nch = chunk_item.size ();
    auto al = [] (size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_ci = al (n * sizeof (LtItem)), o_data = o_ci + al ((nch + 1) * 4);
    int rc = engine_reserve (e, o_data + host_bytes, o_data + 256); if (rc) return rc;
    cudaStream_t st = e->stream;
    size_t cur = o_data;
    for (uint32_t i = 0; i < n; i++) {
        const size_t bytes = items[i].n_elems * lt_width (items[i].op);
        if (devptr) h[i].data = items[i].data;
        else { h[i].data = e->ws + cur; if (bytes) CK (cudaMemcpyAsync (e->ws + cur, items[i].data, bytes, cudaMemcpyHostToDevice, st)); cur += (bytes + 255) & ~(size_t)255; }
    }
    memcpy (e->pin, h.data (), n * sizeof (LtItem));
    memcpy (e->pin + o_ci, chunk_item.data (), nch * 4);
    CK (cudaMemcpyAsync (e->ws, e->pin, n * sizeof (LtItem), cudaMemcpyHostToDevice, st));
    if (nch) {
        CK (cudaMemcpyAsync (e->ws + o_ci, e->pin + o_ci, nch * 4, cudaMemcpyHostToDevice, st));
        k_local_transform<<<(uint32_t)nch, 256, 0, st>>>(reinterpret_cast<const LtItem *>(e->ws), reinterpret_cast<const uint32_t *>(e->ws + o_ci)); e->launches++;
    }
    if (!devptr) for (uint32_t i = 0; i < n; i++) {
        const size_t bytes = items[i].n_elems * lt_width (items[i].op);
        if (bytes) CK (cudaMemcpyAsync (items[i].data, h[i].data, bytes, cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    CK (cudaGetLastError ());
    return GZB_OK;
}


// ------------------------------------------------------------------------------------------------ matrix transposes of a local buffer
// dyn_int_transpose (src/dyn_int.c:45-105, the case without copied samples): a local of rows x cols integers (one row per line, one
// column per sample or per item of an array field) is stored column by column, so that a column's similar values are neighbours for the
// codec: trans[c * rows + r] = data[r * cols + c]; a local that is not a rectangle is left alone (:75-78).  PIZ: BGEN_transpose_u8/16/32_buf
// (src/buffer.c:364-391): back to row by row, then from big endian.  Out of place through the engine's workspace (the reference goes
// through vb->scratch), 32 x 32 tiles through shared memory so that both the reads and the writes are coalesced.
namespace {

struct TrItem { const void *in; void *out; unsigned long long R, C; uint32_t width, swap; unsigned long long first_tile, tiles_c; };

template <typename U> __device__ __forceinline__ void tr_tile (const TrItem &it, unsigned long long tile, U (*sm)[33])
{
    const unsigned long long tr = tile / it.tiles_c, tc = tile % it.tiles_c;
    const U *in = reinterpret_cast<const U *>(it.in); U *out = reinterpret_cast<U *>(it.out);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                        // 256 threads: 8 rows of the tile at a time
    for (int k = ty; k < 32; k += 8) {
        const unsigned long long r = tr * 32 + k, c = tc * 32 + tx;
        if (r < it.R && c < it.C) sm[k][tx] = in[r * it.C + c];
    }
    __syncthreads ();
    for (int k = ty; k < 32; k += 8) {
        const unsigned long long c = tc * 32 + k, r = tr * 32 + tx;                // out is C x R
        if (r < it.R && c < it.C) { const U v = sm[tx][k]; out[c * it.R + r] = it.swap ? bswap (v) : v; }
    }
    __syncthreads ();
}

__global__ void __launch_bounds__(256) k_local_transpose (const TrItem *items, const uint32_t *tile_item, unsigned long long n_tiles)
{
    __shared__ uint32_t sm32[32][33];
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const TrItem it = items[tile_item[t >> 10]];                                // tiles are listed in groups of 1024
        const unsigned long long tile = t - it.first_tile;
        if (it.width == 1)      tr_tile<uint8_t>  (it, tile, reinterpret_cast<uint8_t (*)[33]>(sm32));
        else if (it.width == 2) tr_tile<uint16_t> (it, tile, reinterpret_cast<uint16_t (*)[33]>(sm32));
        else                    tr_tile<uint32_t> (it, tile, sm32);
    }
}

} // namespace

extern "C" int gzb_local_transpose_batch (gzb_engine *e, gzb_transpose_item *items, uint32_t n, uint32_t flags)
{
    if (!e || (!items && n)) return GZB_E_BADARG;
    if (!n) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    std::vector<TrItem> h (n);
    std::vector<uint32_t> tile_item;                                                // one entry per group of 1024 tiles
    unsigned long long n_tiles = 0;
    size_t bytes_total = 0;
    for (uint32_t i = 0; i < n; i++) {
        gzb_transpose_item &it = items[i];
        const uint32_t w = it.width;
        if ((w != 1 && w != 2 && w != 4) || !it.cols || it.dir > 1 || (!it.data && it.n_elems) || ((uintptr_t)it.data & (w - 1))) { it.status = GZB_E_BADARG; e->err = "bad transpose item"; return GZB_E_BADARG; }
        it.status = GZB_OK; it.transposed = 0;
        h[i] = TrItem ();
        if (!it.n_elems) continue;
        if (it.n_elems % it.cols) {                                                  // not a rectangle
            if (it.dir == GZB_TR_PIZ) { it.status = GZB_E_CORRUPT; e->err = "transposed local is not a rectangle"; return GZB_E_CORRUPT; }
            continue;                                                               // ZIP: left as it is (dyn_int.c:75-78)
        }
        const unsigned long long rows = it.n_elems / it.cols;
        // ZIP reads rows x cols and writes cols x rows; PIZ reads cols x rows and writes rows x cols
        h[i].R = it.dir == GZB_TR_ZIP ? rows : it.cols; h[i].C = it.dir == GZB_TR_ZIP ? it.cols : rows;
        h[i].width = w; h[i].swap = it.dir == GZB_TR_PIZ && w > 1;
        h[i].tiles_c = (h[i].C + 31) / 32;
        n_tiles = (n_tiles + 1023) & ~1023ull;                                      // every item starts a new group
        h[i].first_tile = n_tiles;
        const unsigned long long t = ((h[i].R + 31) / 32) * h[i].tiles_c;
        tile_item.insert (tile_item.end (), (size_t)((t + 1023) / 1024), i);
        n_tiles += t;
        it.transposed = 1;
        bytes_total += ((size_t)it.n_elems * w + 255) & ~(size_t)255;
    }
    if (!n_tiles) return GZB_OK;
    auto al = [] (size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_ti = al (n * sizeof (TrItem)), o_data = o_ti + al ((tile_item.size () + 1) * 4);
    int rc = engine_reserve (e, o_data + bytes_total * (devptr ? 1 : 2), o_data + 256); if (rc) return rc;
    cudaStream_t st = e->stream;
    size_t cur = o_data;
    for (uint32_t i = 0; i < n; i++) {
        if (!items[i].transposed) continue;
        const size_t bytes = (size_t)items[i].n_elems * items[i].width, ab = al (bytes);
        if (devptr) h[i].in = items[i].data;
        else { h[i].in = e->ws + cur; CK (cudaMemcpyAsync (e->ws + cur, items[i].data, bytes, cudaMemcpyHostToDevice, st)); cur += ab; }
        h[i].out = e->ws + cur; cur += ab;
    }
    // an item whose tile range ends inside a group of 1024 shares the group's entry with nobody (the next item starts a new group), but the
    // kernel's loop runs over the padding too: give the padding tiles of a group to their item as tiles past its end
    memcpy (e->pin, h.data (), n * sizeof (TrItem));
    memcpy (e->pin + o_ti, tile_item.data (), tile_item.size () * 4);
    CK (cudaMemcpyAsync (e->ws, e->pin, n * sizeof (TrItem), cudaMemcpyHostToDevice, st));
    CK (cudaMemcpyAsync (e->ws + o_ti, e->pin + o_ti, tile_item.size () * 4, cudaMemcpyHostToDevice, st));
    const unsigned long long padded = (n_tiles + 1023) & ~1023ull;
    const uint32_t grid = (uint32_t)std::min<unsigned long long> (padded, 148ull * 16);
    k_local_transpose<<<grid, 256, 0, st>>>(reinterpret_cast<const TrItem *>(e->ws), reinterpret_cast<const uint32_t *>(e->ws + o_ti), padded); e->launches++;
    for (uint32_t i = 0; i < n; i++) {
        if (!items[i].transposed) continue;
        const size_t bytes = (size_t)items[i].n_elems * items[i].width;
        CK (cudaMemcpyAsync (items[i].data, h[i].out, bytes, devptr ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    }
    CK (cudaStreamSynchronize (st));
    CK (cudaGetLastError ());
    return GZB_OK;
}
