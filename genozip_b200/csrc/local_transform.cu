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
